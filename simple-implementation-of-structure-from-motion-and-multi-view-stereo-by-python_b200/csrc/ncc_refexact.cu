// K0 prep_gray_stack and K1 ncc_score_refexact ("Mode A") for sm_100a.
//
// Mode A is the literal behaviour of the reference scorer
// (MVS2.py:62-77 + MVS2.py:39-43 + HarrisFeatures.py:116-133 + utils.py:241-244):
// project the centre with the REFERENCE view's camera, truncate, cut the same
// (2*wid+1)^2 gray window out of every view, NCC each against the reference view's
// window, keep views with ncc > thr, average them.
//
// Hardware mapping (DESIGN.md section "K1"): one warp per hypothesis.  A window row
// is 11 bytes at an arbitrary column, i.e. inside one 4-byte-aligned 16-byte chunk.
// Because every view is sampled at the same (row, col) and rows are 128-byte
// pitched, the chunk has the SAME byte alignment in every view, so the reference
// window and the view windows are compared word against word with one byte mask and
// no realignment.  lane = (sub = lane>>2, word = lane&3):
//   pass A(view)      : row = sub (0..7), 32 lanes x 4 B = 8 rows of one view
//   pass B(view pair) : rows 8..10 of view 2q (lanes 0..11) and 2q+1 (lanes 16..27)
// Per word three dp4a give sum(w), sum(w*w), sum(w*ref) on packed u8, exact in int32.
// Per-view totals are produced by a transposing butterfly (one shuffle per view and
// quantity instead of five), the ratio is taken in fp64 so that the strict
// threshold test and the oracle agree bit for bit.  Tensor cores are not used: this
// is a gather-bound integer reduction with no dense contraction.
#include "mvs_common.cuh"

#define FULL 0xffffffffu

// ---------------------------------------------------------------------------------
// K0: RGB u8 [V,H,W,3] -> gray u8 [V,H,pitch].  cv2.cvtColor(BGR2GRAY) applied to an
// RGB-ordered array (HarrisFeatures.py:124-125 fed by main.py:18):
//   g = (R*3735 + G*19235 + B*9798 + 16384) >> 15
// HBM-bound streaming kernel: 3 B read + 1 B written per pixel.
// ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) prep_gray_stack(const uint8_t* __restrict__ rgb, uint8_t* __restrict__ gray, int H,
                                                       int W, int64_t pitch, int64_t rows_total) {
    // one thread = 4 consecutive pixels of one row (12 B in, 4 B out)
    const int quads = (W + 3) >> 2;
    const int64_t total = rows_total * quads;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t row = i / quads;
        const int q = (int)(i - row * quads);
        const uint8_t* src = rgb + (row * W + q * 4) * 3;
        uint32_t out = 0;
        const int npx = min(4, W - q * 4);
        if (npx == 4 && ((reinterpret_cast<uintptr_t>(src) & 3) == 0)) {
            const uint32_t a = __ldg(reinterpret_cast<const uint32_t*>(src));
            const uint32_t b = __ldg(reinterpret_cast<const uint32_t*>(src) + 1);
            const uint32_t c = __ldg(reinterpret_cast<const uint32_t*>(src) + 2);
            // bytes: a = R0 G0 B0 R1 | b = G1 B1 R2 G2 | c = B2 R3 G3 B3
            const uint32_t r0 = a & 255, g0 = (a >> 8) & 255, b0 = (a >> 16) & 255, r1 = a >> 24;
            const uint32_t g1 = b & 255, b1 = (b >> 8) & 255, r2 = (b >> 16) & 255, g2 = b >> 24;
            const uint32_t b2 = c & 255, r3 = (c >> 8) & 255, g3 = (c >> 16) & 255, b3 = c >> 24;
            out = ((r0 * 3735u + g0 * 19235u + b0 * 9798u + 16384u) >> 15) |
                  (((r1 * 3735u + g1 * 19235u + b1 * 9798u + 16384u) >> 15) << 8) |
                  (((r2 * 3735u + g2 * 19235u + b2 * 9798u + 16384u) >> 15) << 16) |
                  (((r3 * 3735u + g3 * 19235u + b3 * 9798u + 16384u) >> 15) << 24);
        } else {
            for (int k = 0; k < npx; ++k) {
                const uint32_t r = src[3 * k], g = src[3 * k + 1], b = src[3 * k + 2];
                out |= ((r * 3735u + g * 19235u + b * 9798u + 16384u) >> 15) << (8 * k);
            }
        }
        *reinterpret_cast<uint32_t*>(gray + row * pitch + q * 4) = out;
    }
}

int mvs_launch_gray(mvs_ctx* ctx, const uint8_t* d_rgb, cudaStream_t s) {
    const int64_t rows = (int64_t)ctx->V * ctx->H;
    const int64_t total = rows * ((ctx->W + 3) >> 2);
    int blocks = (int)((total + 255) / 256);
    const int cap = ctx->sm_count * 16;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    prep_gray_stack<<<blocks, 256, 0, s>>>(d_rgb, ctx->d_gray, ctx->H, ctx->W, ctx->pitch, rows);
    ctx->launches++;
    MVS_CUDA_CHECK(cudaGetLastError());
    return MVS_OK;
}

// ---------------------------------------------------------------------------------
// Projection, bit-compatible with cv2.projectPoints as called by utils.py:241-244:
//   X = (r0*c0 + r1*c1 + r2*c2) + t  (left to right, no FMA contraction),
//   z = z ? 1/z : 1;  x = (X*z)*fx + cx.
// ---------------------------------------------------------------------------------
__device__ __forceinline__ void project_ref(const CamProj& cam, double c0, double c1, double c2, double& x, double& y) {
    const double X =
        __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(cam.r[0], c0), __dmul_rn(cam.r[1], c1)), __dmul_rn(cam.r[2], c2)), cam.t[0]);
    const double Y =
        __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(cam.r[3], c0), __dmul_rn(cam.r[4], c1)), __dmul_rn(cam.r[5], c2)), cam.t[1]);
    const double Z =
        __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(cam.r[6], c0), __dmul_rn(cam.r[7], c1)), __dmul_rn(cam.r[8], c2)), cam.t[2]);
    const double iz = (Z != 0.0) ? __ddiv_rn(1.0, Z) : 1.0;
    x = __dadd_rn(__dmul_rn(__dmul_rn(X, iz), cam.fx), cam.cx);
    y = __dadd_rn(__dmul_rn(__dmul_rn(Y, iz), cam.fy), cam.cy);
}

// int() truncation toward zero + the asymmetric bounds rule of HarrisFeatures.py:128.
// Non-finite projections are rejected (the reference would raise inside int()).
__device__ __forceinline__ bool window_anchor(double x, double y, int H, int W, int wid, int& row, int& col) {
    if (!(isfinite(x) && isfinite(y))) {
        row = col = 0;
        return false;
    }
    const double lim = 1073741824.0;
    row = (int)fmin(fmax(y, -lim), lim);
    col = (int)fmin(fmax(x, -lim), lim);
    return (row - wid >= 0) && (row + wid + 1 < H) && (col - wid > 0) && (col + wid + 1 < W);
}

__device__ __forceinline__ int dp4a_u(uint32_t a, uint32_t b, int c) { return (int)__dp4a(a, b, (unsigned)c); }

// Final ratio in fp64, same operation order as oracle/mode_a.py::score.
__device__ __forceinline__ double ncc_from_sums(int n, int S, int SS, int SAB, int Sr, int var_r, bool& defined) {
    const int var_i = n * SS - S * S;                 // exact: <= 121*121*255^2 < 2^31
    const int num = n * SAB - S * Sr;
    defined = (var_i != 0) && (var_r != 0);
    const double den = (double)var_i * (double)var_r;
    return ((double)num / sqrt(den)) * ((double)n / (double)(n - 1));
}

// view-in-group index held by a lane after the transposing butterfly below
__device__ __forceinline__ int lane_view16(int lane) {
    return ((lane >> 4) & 1) + 8 * ((lane >> 3) & 1) + 4 * ((lane >> 2) & 1) + 2 * ((lane >> 1) & 1);
}

// Reduce 16 per-view partials (a[0..15], one per view of the group) plus the 8 tail
// partials (tail[q] belongs to view 2q + (lane>>4)) across the warp.  On return every
// lane holds the complete sum of view lane_view16(lane).  16 shuffles.
__device__ __forceinline__ int butterfly16(const int (&a)[16], const int (&tail)[8], int lane) {
    int b[8];
    const bool h16 = lane & 16;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        const int keep = h16 ? a[2 * q + 1] : a[2 * q];
        const int send = h16 ? a[2 * q] : a[2 * q + 1];
        b[q] = keep + tail[q] + __shfl_xor_sync(FULL, send, 16);
    }
    int c4[4];
    const bool h8 = lane & 8;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int keep = h8 ? b[q + 4] : b[q];
        const int send = h8 ? b[q] : b[q + 4];
        c4[q] = keep + __shfl_xor_sync(FULL, send, 8);
    }
    int d2[2];
    const bool h4 = lane & 4;
#pragma unroll
    for (int q = 0; q < 2; ++q) {
        const int keep = h4 ? c4[q + 2] : c4[q];
        const int send = h4 ? c4[q] : c4[q + 2];
        d2[q] = keep + __shfl_xor_sync(FULL, send, 4);
    }
    const bool h2 = lane & 2;
    const int keep = h2 ? d2[1] : d2[0];
    const int send = h2 ? d2[0] : d2[1];
    int e = keep + __shfl_xor_sync(FULL, send, 2);
    e += __shfl_xor_sync(FULL, e, 1);
    return e;
}

struct GroupSums {
    int S, SS, SAB;
};

// Gather + reduce one group of 16 views (g16*16 .. g16*16+15) for the warp's
// hypothesis.  pA/pB: this lane's byte address inside view 0 for pass A / pass B.
__device__ __forceinline__ GroupSums gather_group16(const uint8_t* __restrict__ pA, const uint8_t* __restrict__ pB,
                                                    int64_t vstride, int V, int g16, uint32_t m, uint32_t refA,
                                                    uint32_t refB, bool activeB, int lane) {
    int aS[16], aSS[16], aSAB[16];
    int tS[8], tSS[8], tSAB[8];
    const int v0 = g16 * 16;
    uint32_t wA[16], wB[8];
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        const int v = v0 + j;
        wA[j] = (v < V) ? (__ldg(reinterpret_cast<const uint32_t*>(pA + (int64_t)v * vstride)) & m) : 0u;
    }
    const int hi = (lane >> 4) & 1;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        const int v = v0 + 2 * q + hi;
        wB[q] = (activeB && v < V) ? (__ldg(reinterpret_cast<const uint32_t*>(pB + (int64_t)v * vstride)) & m) : 0u;
    }
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        aS[j] = dp4a_u(wA[j], 0x01010101u, 0);
        aSS[j] = dp4a_u(wA[j], wA[j], 0);
        aSAB[j] = dp4a_u(wA[j], refA, 0);
    }
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        tS[q] = dp4a_u(wB[q], 0x01010101u, 0);
        tSS[q] = dp4a_u(wB[q], wB[q], 0);
        tSAB[q] = dp4a_u(wB[q], refB, 0);
    }
    GroupSums r;
    r.S = butterfly16(aS, tS, lane);
    r.SS = butterfly16(aSS, tSS, lane);
    r.SAB = butterfly16(aSAB, tSAB, lane);
    return r;
}

// ---------------------------------------------------------------------------------
// K1 (wid = 5): one warp per hypothesis, grid-stride over hypotheses.
// ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(256, 2)
    ncc_score_refexact_w5(const uint8_t* __restrict__ gray, const CamProj* __restrict__ cams, int V, int H, int W,
                          int64_t pitch, int64_t vstride, int64_t N, const double* __restrict__ c,
                          const int32_t* __restrict__ ref, double thr, uint64_t* __restrict__ vis_out,
                          double* __restrict__ avg_out, int32_t* __restrict__ count_out, double* __restrict__ xy_out,
                          float* __restrict__ ncc_out) {
    constexpr int WID = 5;
    constexpr int NPIX = 121;
    const int lane = threadIdx.x & 31;
    const int word = lane & 3;
    const int sub = lane >> 2;
    const int64_t warp0 = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int mask_words32 = 2 * ((V + 63) >> 6);          // 32-bit chunks per hypothesis in vis_out
    const int npairs = (V + 31) >> 5;                      // passes of 32 views

    for (int64_t h = warp0; h < N; h += nwarps) {
        const int r = __ldg(ref + h);
        double x = nan(""), y = nan("");
        int row = 0, col = 0;
        bool valid = false;
        if (r >= 0 && r < V) {
            const double c0 = __ldg(c + 3 * h), c1 = __ldg(c + 3 * h + 1), c2 = __ldg(c + 3 * h + 2);
            project_ref(cams[r], c0, c1, c2, x, y);
            valid = window_anchor(x, y, H, W, WID, row, col);
        }
        if (lane == 0 && xy_out) {
            xy_out[2 * h] = x;
            xy_out[2 * h + 1] = y;
        }
        if (!valid) {                                      // getDescFeatures -> [None]: V = [], avg = 0
            if (lane < mask_words32) reinterpret_cast<uint32_t*>(vis_out)[h * mask_words32 + lane] = 0u;
            if (lane == 0) {
                count_out[h] = 0;
                if (avg_out) avg_out[h] = 0.0;
            }
            if (ncc_out)
                for (int v = lane; v < V; v += 32) ncc_out[h * V + v] = nanf("");
            continue;
        }
        const int o = (col - WID) & 3;                     // byte offset of the window inside its 16-B chunk
        const int a0 = (col - WID) - o;
        // byte mask of this lane's word: byte b is in the window iff o <= 4*word+b <= o+10
        uint32_t m = 0;
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const int pos = 4 * word + b;
            if (pos >= o && pos <= o + 2 * WID) m |= 0xffu << (8 * b);
        }
        const int64_t base = (int64_t)(row - WID) * pitch + a0 + 4 * word;
        const uint8_t* pA = gray + base + (int64_t)sub * pitch;
        const bool activeB = (sub & 3) < 3;
        const uint8_t* pB = gray + base + (int64_t)(8 + (sub & 3)) * pitch;
        const uint32_t refA = __ldg(reinterpret_cast<const uint32_t*>(pA + (int64_t)r * vstride)) & m;
        const uint32_t refB = activeB ? (__ldg(reinterpret_cast<const uint32_t*>(pB + (int64_t)r * vstride)) & m) : 0u;
        // reference-window sums (tail rows counted once: lanes 0..11)
        const uint32_t refB1 = (lane < 16) ? refB : 0u;
        int Sr = dp4a_u(refA, 0x01010101u, dp4a_u(refB1, 0x01010101u, 0));
        int SSr = dp4a_u(refA, refA, dp4a_u(refB1, refB1, 0));
        Sr = __reduce_add_sync(FULL, Sr);
        SSr = __reduce_add_sync(FULL, SSr);
        const int var_r = NPIX * SSr - Sr * Sr;

        double acc = 0.0;
        int count = 0;
        uint32_t mychunk = 0;
        for (int p = 0; p < npairs; ++p) {
            GroupSums ga = gather_group16(pA, pB, vstride, V, 2 * p, m, refA, refB, activeB, lane);
            GroupSums gb = ga;
            if ((2 * p + 1) * 16 < V) gb = gather_group16(pA, pB, vstride, V, 2 * p + 1, m, refA, refB, activeB, lane);
            const bool odd = lane & 1;
            const int S = odd ? gb.S : ga.S, SS = odd ? gb.SS : ga.SS, SAB = odd ? gb.SAB : ga.SAB;
            const int pos = (odd ? 16 : 0) + lane_view16(lane);        // bit inside this 32-view chunk
            const int v = p * 32 + pos;
            bool defined;
            const double val = ncc_from_sums(NPIX, S, SS, SAB, Sr, var_r, defined);
            const bool scored = (v < V) && (v != r) && defined;
            const bool vis = scored && (val > thr);
            if (vis) acc += val;
            const uint32_t chunk = __reduce_or_sync(FULL, vis ? (1u << pos) : 0u);
            count += __popc(chunk);
            if (lane == p) mychunk = chunk;                 // p < 32 chunks <=> V <= 1024
            if (ncc_out && v < V) ncc_out[h * V + v] = scored ? (float)val : nanf("");
        }
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) acc += __shfl_xor_sync(FULL, acc, s);
        if (lane < mask_words32) reinterpret_cast<uint32_t*>(vis_out)[h * mask_words32 + lane] = mychunk;
        if (lane == 0) {
            count_out[h] = count;
            if (avg_out) avg_out[h] = count > 0 ? acc / (double)count : 0.0;
        }
    }
}

// ---------------------------------------------------------------------------------
// Generic-wid variant (wid 1..7): one warp per hypothesis, lanes over views, byte
// loads.  Not tuned: the reference hard-codes wid = 5 (MVS2.py:64,69).
// ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
    ncc_score_refexact_any(const uint8_t* __restrict__ gray, const CamProj* __restrict__ cams, int V, int H, int W,
                           int64_t pitch, int64_t vstride, int64_t N, const double* __restrict__ c,
                           const int32_t* __restrict__ ref, double thr, int wid, uint64_t* __restrict__ vis_out,
                           double* __restrict__ avg_out, int32_t* __restrict__ count_out, double* __restrict__ xy_out,
                           float* __restrict__ ncc_out) {
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int mask_words32 = 2 * ((V + 63) >> 6);
    const int k = 2 * wid + 1;
    const int n = k * k;
    for (int64_t h = warp0; h < N; h += nwarps) {
        const int r = __ldg(ref + h);
        double x = nan(""), y = nan("");
        int row = 0, col = 0;
        bool valid = false;
        if (r >= 0 && r < V) {
            project_ref(cams[r], __ldg(c + 3 * h), __ldg(c + 3 * h + 1), __ldg(c + 3 * h + 2), x, y);
            valid = window_anchor(x, y, H, W, wid, row, col);
        }
        if (lane == 0 && xy_out) {
            xy_out[2 * h] = x;
            xy_out[2 * h + 1] = y;
        }
        double acc = 0.0;
        int count = 0;
        uint32_t mychunk = 0;
        const uint8_t* wr = gray + (int64_t)r * vstride + (int64_t)(row - wid) * pitch + (col - wid);
        long long Sr = 0, SSr = 0;
        if (valid) {
            for (int i = 0; i < k; ++i)
                for (int j = 0; j < k; ++j) {
                    const long long a = wr[(int64_t)i * pitch + j];
                    Sr += a;
                    SSr += a * a;
                }
        }
        const long long var_r = n * SSr - Sr * Sr;
        for (int p = 0; p * 32 < V; ++p) {
            const int v = p * 32 + lane;
            bool vis = false, scored = false;
            double val = 0.0;
            if (valid && v < V && v != r) {
                const uint8_t* wv = gray + (int64_t)v * vstride + (int64_t)(row - wid) * pitch + (col - wid);
                long long S = 0, SS = 0, SAB = 0;
                for (int i = 0; i < k; ++i)
                    for (int j = 0; j < k; ++j) {
                        const long long a = wv[(int64_t)i * pitch + j];
                        const long long b = wr[(int64_t)i * pitch + j];
                        S += a;
                        SS += a * a;
                        SAB += a * b;
                    }
                const long long var_i = n * SS - S * S;
                const long long num = n * SAB - S * Sr;
                scored = (var_i != 0) && (var_r != 0);
                val = ((double)num / sqrt((double)var_i * (double)var_r)) * ((double)n / (double)(n - 1));
                vis = scored && (val > thr);
            }
            if (vis) acc += val;
            const uint32_t chunk = __ballot_sync(FULL, vis);
            count += __popc(chunk);
            if (lane == p) mychunk = chunk;
            if (ncc_out && v < V) ncc_out[h * V + v] = scored ? (float)val : nanf("");
        }
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) acc += __shfl_xor_sync(FULL, acc, s);
        if (lane < mask_words32) reinterpret_cast<uint32_t*>(vis_out)[h * mask_words32 + lane] = mychunk;
        if (lane == 0) {
            count_out[h] = count;
            if (avg_out) avg_out[h] = count > 0 ? acc / (double)count : 0.0;
        }
    }
}

int mvs_launch_score_refexact(mvs_ctx* ctx, int64_t N, const double* c, const int32_t* ref, double thr, int wid,
                              uint64_t* vis, double* avg, int32_t* count, double* xy, float* ncc, cudaStream_t s) {
    if (N == 0) return MVS_OK;
    const int warps_per_block = 8;
    int64_t want = (N + warps_per_block - 1) / warps_per_block;
    const int64_t cap = (int64_t)ctx->sm_count * 2 * 4;       // 2 resident CTAs/SM, x4 for tail balance
    const int blocks = (int)(want < cap ? want : cap);
    if (wid == 5) {
        ncc_score_refexact_w5<<<blocks, 256, 0, s>>>(ctx->d_gray, ctx->d_cam, ctx->V, ctx->H, ctx->W, ctx->pitch,
                                                     ctx->vstride, N, c, ref, thr, vis, avg, count, xy, ncc);
    } else {
        ncc_score_refexact_any<<<blocks, 256, 0, s>>>(ctx->d_gray, ctx->d_cam, ctx->V, ctx->H, ctx->W, ctx->pitch,
                                                      ctx->vstride, N, c, ref, thr, wid, vis, avg, count, xy, ncc);
    }
    ctx->launches++;
    MVS_CUDA_CHECK(cudaGetLastError());
    return MVS_OK;
}
