// K0 prep_gray4 and K1 ncc_score_gather ("Mode A") for sm_100a.
//
// Mode A is the literal behaviour of the reference scorer
// (MVS2.py:62-77 + MVS2.py:39-43 + HarrisFeatures.py:116-133 + utils.py:241-244):
// project the centre with the REFERENCE view's camera, truncate, cut the same
// (2*wid+1)^2 gray window out of every view, NCC each against the reference view's
// window, keep views with ncc > thr, average them.
//
// Hardware mapping (DESIGN.md section "K1").  Every view is sampled at the SAME
// (row, col) (MVS2.py:68), so the resident stack is stored view-interleaved,
// u8 [H][G][Vp][4]: one window row of ALL views is one contiguous run of at most
// NG * 4*Vp bytes, 16-byte aligned.  A hypothesis is owned by LPH lanes (LPH = 4, 8, 16
// or 32, the power of two >= Vp/4), 32/LPH hypotheses per warp.  Lane q owns the four
// views 4q..4q+3: per window row and pixel group it issues ONE 16-byte load (4 views x 4
// pixels), and per 32-bit word three dp4a give sum(w), sum(w*w), sum(w*ref) on packed
// u8 -- exact in int32.  There is NO cross-lane reduction of the per-view sums: each
// lane finishes its own four views, the ratio is taken in fp64 in the oracle's
// operation order so that the strict threshold test agrees bit for bit.  The reference
// window (K x NG masked words) is staged in shared memory once per hypothesis and read
// back as one broadcast LDS.128 per window row.  Tensor cores are not used: this is a
// gather-bound integer reduction with no dense contraction.
#include "mvs_common.cuh"

#define FULL 0xffffffffu

// ---------------------------------------------------------------------------------
// K0: RGB u8 [V,H,W,3] -> gray u8 [H][G][Vp][4].  cv2.cvtColor(BGR2GRAY) applied to an
// RGB-ordered array (HarrisFeatures.py:124-125 fed by main.py:18):
//   g = (R*3735 + G*19235 + B*9798 + 16384) >> 15
// One thread = 4 consecutive pixels of one view (12 B in, 4 B out), views fastest so
// that the interleaved stores coalesce.  Runs once per context.
// ---------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t gray_px(uint32_t r, uint32_t g, uint32_t b) {
    return (r * 3735u + g * 19235u + b * 9798u + 16384u) >> 15;
}

__global__ void __launch_bounds__(256) prep_gray4(const uint8_t* __restrict__ rgb, uint8_t* __restrict__ gray4, int V, int H,
                                                  int W, int Gw, int64_t gstride, int64_t rowpitch) {
    const int64_t total = (int64_t)H * Gw * V;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int v = (int)(i % V);
        const int64_t rg = i / V;
        const int g = (int)(rg % Gw);
        const int row = (int)(rg / Gw);
        const uint8_t* src = rgb + (((int64_t)v * H + row) * W + g * 4) * 3;
        uint32_t out = 0;
        const int npx = min(4, W - g * 4);
        if (npx == 4 && ((reinterpret_cast<uintptr_t>(src) & 3) == 0)) {
            const uint32_t a = __ldg(reinterpret_cast<const uint32_t*>(src));
            const uint32_t b = __ldg(reinterpret_cast<const uint32_t*>(src) + 1);
            const uint32_t c = __ldg(reinterpret_cast<const uint32_t*>(src) + 2);
            // bytes: a = R0 G0 B0 R1 | b = G1 B1 R2 G2 | c = B2 R3 G3 B3
            out = gray_px(a & 255, (a >> 8) & 255, (a >> 16) & 255) |
                  (gray_px(a >> 24, b & 255, (b >> 8) & 255) << 8) |
                  (gray_px((b >> 16) & 255, b >> 24, c & 255) << 16) |
                  (gray_px((c >> 8) & 255, (c >> 16) & 255, c >> 24) << 24);
        } else {
            for (int k = 0; k < npx; ++k) out |= gray_px(src[3 * k], src[3 * k + 1], src[3 * k + 2]) << (8 * k);
        }
        *reinterpret_cast<uint32_t*>(gray4 + row * rowpitch + g * gstride + 4 * v) = out;
    }
}

int mvs_launch_gray(mvs_ctx* ctx, const uint8_t* d_rgb, cudaStream_t s) {
    const int Gw = (ctx->W + 3) >> 2;
    const int64_t total = (int64_t)ctx->H * Gw * ctx->V;
    int64_t blocks = (total + 255) / 256;
    const int64_t cap = (int64_t)ctx->sm_count * 16;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    prep_gray4<<<(int)blocks, 256, 0, s>>>(d_rgb, ctx->d_gray, ctx->V, ctx->H, ctx->W, Gw, ctx->gstride, ctx->rowpitch);
    ctx->launches++;
    MVS_CUDA_CHECK(cudaGetLastError());
    return MVS_OK;
}

// Resident interleaved stack -> dense planar [V,H,W] (parity/debug path of mvs_download_gray).
__global__ void __launch_bounds__(256) unpack_gray4(const uint8_t* __restrict__ gray4, uint8_t* __restrict__ planar, int V,
                                                    int H, int W, int64_t gstride, int64_t rowpitch) {
    const int64_t total = (int64_t)V * H * W;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int col = (int)(i % W);
        const int64_t vr = i / W;
        const int row = (int)(vr % H);
        const int v = (int)(vr / H);
        planar[i] = gray4[row * rowpitch + (col >> 2) * gstride + 4 * v + (col & 3)];
    }
}

int mvs_launch_unpack_gray(mvs_ctx* ctx, uint8_t* d_planar, cudaStream_t s) {
    const int64_t total = (int64_t)ctx->V * ctx->H * ctx->W;
    int64_t blocks = (total + 255) / 256;
    const int64_t cap = (int64_t)ctx->sm_count * 16;
    if (blocks > cap) blocks = cap;
    unpack_gray4<<<(int)blocks, 256, 0, s>>>(ctx->d_gray, d_planar, ctx->V, ctx->H, ctx->W, ctx->gstride, ctx->rowpitch);
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

// Bytes of pixel group g (window-relative positions 4g..4g+3) that lie inside the
// K-pixel run starting at offset o (0..3) of group 0.
__device__ __forceinline__ uint32_t group_mask(int o, int K, int g) {
    int lo = o - 4 * g;
    lo = lo < 0 ? 0 : lo;
    int hi = o + K - 1 - 4 * g;
    hi = hi > 3 ? 3 : hi;
    if (lo > hi) return 0u;
    return (0xffffffffu << (8 * lo)) & (0xffffffffu >> (8 * (3 - hi)));
}

// Final ratio in fp64, same operation order as oracle/mode_a.py::score.  All integer
// terms are exact: n*SS, S*S <= 225^2*255^2 need 64 bits at wid = 7.
__device__ __forceinline__ double ncc_from_sums(int n, int S, int SS, int SAB, int Sr, long long var_r, bool& defined) {
    const long long var_i = (long long)n * SS - (long long)S * S;
    const long long num = (long long)n * SAB - (long long)S * Sr;
    defined = (var_i != 0) && (var_r != 0);
    const double den = (double)var_i * (double)var_r;
    return ((double)num / sqrt(den)) * ((double)n / (double)(n - 1));
}

template <int LPH>
__device__ __forceinline__ uint32_t hyp_mask(int sub) {
    if constexpr (LPH == 32) {
        return FULL;
    } else {
        return ((1u << LPH) - 1u) << (sub * LPH);
    }
}

// ---------------------------------------------------------------------------------
// K1: LPH lanes per hypothesis, 32/LPH hypotheses per warp, grid-stride.
// ---------------------------------------------------------------------------------
template <int WID, int LPH>
__global__ void __launch_bounds__(256, 2)
    ncc_score_gather(const uint8_t* __restrict__ gray4, const CamProj* __restrict__ cams, int V, int Q, int H, int W,
                     int64_t gstride, int64_t rowpitch, int64_t N, const double* __restrict__ c,
                     const int32_t* __restrict__ ref, double thr, uint64_t* __restrict__ vis_out,
                     double* __restrict__ avg_out, int32_t* __restrict__ count_out, double* __restrict__ xy_out,
                     float* __restrict__ ncc_out) {
    constexpr int K = 2 * WID + 1;
    constexpr int NG = (K + 6) / 4;            // pixel groups a K-pixel run at offset 0..3 can touch
    constexpr int NPIX = K * K;
    constexpr int HPW = 32 / LPH;
    __shared__ __align__(16) uint32_t s_ref[8][HPW][K][NG];

    const int lane = threadIdx.x & 31;
    const int wib = threadIdx.x >> 5;
    const int sub = lane / LPH;
    const int lih = lane % LPH;
    const uint32_t hmask = hyp_mask<LPH>(sub);
    const int64_t warp0 = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int mask_words32 = 2 * ((V + 63) >> 6);          // 32-bit chunks per hypothesis in vis_out
    const int passes = (LPH == 32) ? (Q + 31) >> 5 : 1;
    uint32_t(*sref)[NG] = s_ref[wib][sub];

    for (int64_t h0 = warp0 * HPW; h0 < N; h0 += nwarps * HPW) {
        const int64_t h = h0 + sub;
        if (h >= N) continue;                              // the whole lane group leaves together
        const int r = __ldg(ref + h);
        double x = nan(""), y = nan("");
        int row = 0, col = 0;
        bool valid = false;
        if (r >= 0 && r < V) {
            const double c0 = __ldg(c + 3 * h), c1 = __ldg(c + 3 * h + 1), c2 = __ldg(c + 3 * h + 2);
            project_ref(cams[r], c0, c1, c2, x, y);
            valid = window_anchor(x, y, H, W, WID, row, col);
        }
        if (lih == 0 && xy_out) {
            xy_out[2 * h] = x;
            xy_out[2 * h + 1] = y;
        }
        if (!valid) {                                      // getDescFeatures -> [None]: V = [], avg = 0
            for (int w = lih; w < mask_words32; w += LPH) reinterpret_cast<uint32_t*>(vis_out)[h * mask_words32 + w] = 0u;
            if (lih == 0) {
                count_out[h] = 0;
                if (avg_out) avg_out[h] = 0.0;
            }
            if (ncc_out)
                for (int v = lih; v < V; v += LPH) ncc_out[h * V + v] = nanf("");
            continue;
        }
        const int o = (col - WID) & 3;                     // offset of the window inside its first pixel group
        const uint8_t* base = gray4 + (int64_t)(row - WID) * rowpitch + (int64_t)((col - WID) >> 2) * gstride;
        uint32_t m[NG];
#pragma unroll
        for (int g = 0; g < NG; ++g) m[g] = group_mask(o, K, g);

        // ---- reference window: masked words to shared memory, its sums reduced over the lane group
        int Sr = 0, SSr = 0;
        __syncwarp(hmask);                                 // previous hypothesis' readers are done
        for (int idx = lih; idx < K * NG; idx += LPH) {
            const int rr = idx / NG, g = idx - rr * NG;
            const uint32_t mg = group_mask(o, K, g);
            uint32_t w = 0u;
            if (mg) w = __ldg(reinterpret_cast<const uint32_t*>(base + rr * rowpitch + g * gstride + 4 * r)) & mg;
            sref[rr][g] = w;
            Sr = dp4a_u(w, 0x01010101u, Sr);
            SSr = dp4a_u(w, w, SSr);
        }
        Sr = __reduce_add_sync(hmask, Sr);
        SSr = __reduce_add_sync(hmask, SSr);
        __syncwarp(hmask);
        const long long var_r = (long long)NPIX * SSr - (long long)Sr * Sr;

        double acc = 0.0;
        int count = 0;
        uint32_t myword = 0;
        for (int p = 0; p < passes; ++p) {
            const int qq = p * LPH + lih;                  // this lane's quad: views 4qq..4qq+3
            const bool act = qq < Q;
            int S[4] = {0, 0, 0, 0}, SS[4] = {0, 0, 0, 0}, SAB[4] = {0, 0, 0, 0};
            const uint8_t* pq = base + (int64_t)qq * 16;
#pragma unroll
            for (int rr = 0; rr < K; ++rr) {
                uint32_t rw[NG];
                if constexpr (NG == 4) {
                    const uint4 t4 = *reinterpret_cast<const uint4*>(&sref[rr][0]);
                    rw[0] = t4.x; rw[1] = t4.y; rw[2] = t4.z; rw[3] = t4.w;
                } else {
#pragma unroll
                    for (int g = 0; g < NG; ++g) rw[g] = sref[rr][g];
                }
#pragma unroll
                for (int g = 0; g < NG; ++g) {
                    if (m[g] != 0u && act) {
                        const uint4 w4 = __ldg(reinterpret_cast<const uint4*>(pq + rr * rowpitch + g * gstride));
                        const uint32_t w[4] = {w4.x & m[g], w4.y & m[g], w4.z & m[g], w4.w & m[g]};
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            S[k] = dp4a_u(w[k], 0x01010101u, S[k]);
                            SS[k] = dp4a_u(w[k], w[k], SS[k]);
                            SAB[k] = dp4a_u(w[k], rw[g], SAB[k]);
                        }
                    }
                }
            }
            // ---- this lane finishes its own four views
            uint32_t nib = 0;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int v = 4 * qq + k;
                bool defined;
                const double val = ncc_from_sums(NPIX, S[k], SS[k], SAB[k], Sr, var_r, defined);
                const bool scored = act && (v < V) && (v != r) && defined;
                const bool vis = scored && (val > thr);
                if (vis) {
                    acc += val;
                    nib |= 1u << k;
                }
                if (ncc_out && act && v < V) ncc_out[h * V + v] = scored ? (float)val : nanf("");
            }
            // lane group covers LPH*4 mask bits per pass = max(LPH/8, 1) 32-bit words
            constexpr int WPP = LPH >= 8 ? LPH / 8 : 1;
#pragma unroll
            for (int w = 0; w < WPP; ++w) {
                const uint32_t contrib = ((lih >> 3) == w) ? (nib << (4 * (lih & 7))) : 0u;
                const uint32_t word = __reduce_or_sync(hmask, contrib);
                count += __popc(word);
                if (lih == p * WPP + w) myword = word;     // word index < 32 <=> V <= 1024
            }
        }
#pragma unroll
        for (int s = LPH / 2; s > 0; s >>= 1) acc += __shfl_xor_sync(hmask, acc, s);
        for (int w = lih; w < mask_words32; w += LPH)
            reinterpret_cast<uint32_t*>(vis_out)[h * mask_words32 + w] = (w == lih) ? myword : 0u;
        if (lih == 0) {
            count_out[h] = count;
            if (avg_out) avg_out[h] = count > 0 ? acc / (double)count : 0.0;
        }
    }
}

template <int WID>
static void launch_gather(mvs_ctx* ctx, int blocks_cap, int64_t N, const double* c, const int32_t* ref, double thr,
                          uint64_t* vis, double* avg, int32_t* count, double* xy, float* ncc, cudaStream_t s) {
    const int Q = ctx->Q;
#define MVS_LAUNCH(LPH)                                                                                              \
    do {                                                                                                             \
        const int64_t per_block = 8 * (32 / LPH);                                                                    \
        int64_t want = (N + per_block - 1) / per_block;                                                              \
        const int blocks = (int)(want < blocks_cap ? want : blocks_cap);                                             \
        ncc_score_gather<WID, LPH><<<blocks, 256, 0, s>>>(ctx->d_gray, ctx->d_cam, ctx->V, Q, ctx->H, ctx->W,         \
                                                          ctx->gstride, ctx->rowpitch, N, c, ref, thr, vis, avg,     \
                                                          count, xy, ncc);                                           \
    } while (0)
    if (Q <= 4)
        MVS_LAUNCH(4);
    else if (Q <= 8)
        MVS_LAUNCH(8);
    else if (Q <= 16)
        MVS_LAUNCH(16);
    else
        MVS_LAUNCH(32);
#undef MVS_LAUNCH
}

int mvs_launch_score_refexact(mvs_ctx* ctx, int64_t N, const double* c, const int32_t* ref, double thr, int wid,
                              uint64_t* vis, double* avg, int32_t* count, double* xy, float* ncc, cudaStream_t s) {
    if (N == 0) return MVS_OK;
    const int cap = ctx->sm_count * 2 * 4;                 // 2 resident CTAs/SM, x4 for tail balance
    switch (wid) {
        case 1: launch_gather<1>(ctx, cap, N, c, ref, thr, vis, avg, count, xy, ncc, s); break;
        case 2: launch_gather<2>(ctx, cap, N, c, ref, thr, vis, avg, count, xy, ncc, s); break;
        case 3: launch_gather<3>(ctx, cap, N, c, ref, thr, vis, avg, count, xy, ncc, s); break;
        case 4: launch_gather<4>(ctx, cap, N, c, ref, thr, vis, avg, count, xy, ncc, s); break;
        case 5: launch_gather<5>(ctx, cap, N, c, ref, thr, vis, avg, count, xy, ncc, s); break;
        case 6: launch_gather<6>(ctx, cap, N, c, ref, thr, vis, avg, count, xy, ncc, s); break;
        case 7: launch_gather<7>(ctx, cap, N, c, ref, thr, vis, avg, count, xy, ncc, s); break;
        default: mvs_set_error("wid %d not supported (1..7)", wid); return MVS_ERR_ARG;
    }
    ctx->launches++;
    MVS_CUDA_CHECK(cudaGetLastError());
    return MVS_OK;
}
