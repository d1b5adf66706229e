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
// pixels) and four dp4a against the masked reference word give sum(w*ref) on packed u8,
// exact in int32.  sum(w) and n*sum(w^2)-sum(w)^2 of every view do not depend on the
// reference view: they are box sums, computed once per (wid) into two resident maps
// (K0b) and read back with one 8-byte and one 16-byte load per lane.  There is NO
// cross-lane reduction of per-view sums: each lane finishes its own four views in fp64.
// The reference window (K x NG masked words) is staged in shared memory once per
// hypothesis and read back as one broadcast LDS.128 per window row.  Hypotheses arrive
// ordered by anchor tile (bin.cu), so neighbouring lane groups hit the same L1 lines.
// Tensor cores are not used: this is a gather-bound integer reduction with no dense
// contraction.
#include "project.cuh"
#include "scan.cuh"
#include <stdlib.h>

#define FULL 0xffffffffu
#ifndef MVS_K1_ITERS
#define MVS_K1_ITERS 4          // block pairs per lane group and CTA pass: chunk = 8 warps x groups x 2 x ITERS positions
#endif

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

__global__ void __launch_bounds__(256) prep_gray4(const uint8_t* __restrict__ rgb, uint8_t* __restrict__ gray4, int v0, int nv,
                                                  int H, int W, int Gw, int64_t gstride, int64_t rowpitch) {
    // rgb holds the views [v0, v0 + nv) only (one upload chunk, or the whole stack)
    const int64_t total = (int64_t)H * Gw * nv;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int vl = (int)(i % nv);
        const int64_t rg = i / nv;
        const int g = (int)(rg % Gw);
        const int row = (int)(rg / Gw);
        const uint8_t* src = rgb + (((int64_t)vl * H + row) * W + g * 4) * 3;
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
        *reinterpret_cast<uint32_t*>(gray4 + row * rowpitch + g * gstride + 4 * (v0 + vl)) = out;
    }
}

// convert the views [v0, v0 + nv) whose RGB pixels start at d_rgb
int mvs_launch_gray(mvs_ctx* ctx, const uint8_t* d_rgb, int v0, int nv, cudaStream_t s) {
    const int Gw = (ctx->W + 3) >> 2;
    const int64_t total = (int64_t)ctx->H * Gw * nv;
    int64_t blocks = (total + 255) / 256;
    const int64_t cap = (int64_t)ctx->sm_count * 16;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    prep_gray4<<<(int)blocks, 256, 0, s>>>(d_rgb, ctx->d_gray, v0, nv, ctx->H, ctx->W, Gw, ctx->gstride, ctx->rowpitch);
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
// K0b: per-anchor window sums of every view (box sums, independent of the reference
// view): smap = sum(w), vmap = n*sum(w^2) - sum(w)^2, for every anchor that passes the
// bounds rule of HarrisFeatures.py:128.  One thread = one anchor x one view quad.
// ---------------------------------------------------------------------------------
template <int WID>
__global__ void __launch_bounds__(256)
    build_window_maps(const uint8_t* __restrict__ gray4, uint16_t* __restrict__ smap, uint32_t* __restrict__ vmap, int Vp,
                      int Q, int H, int W, int64_t gstride, int64_t rowpitch) {
    constexpr int K = 2 * WID + 1;
    constexpr int NG = (K + 6) / 4;
    constexpr uint32_t NPIX = K * K;
    const int64_t total = (int64_t)H * W * Q;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int q = (int)(i % Q);
        const int64_t px = i / Q;
        const int col = (int)(px % W);
        const int row = (int)(px / W);
        if (!((row - WID >= 0) && (row + WID + 1 < H) && (col - WID > 0) && (col + WID + 1 < W))) continue;
        const int o = (col - WID) & 3;
        const uint8_t* pq = gray4 + (int64_t)(row - WID) * rowpitch + (int64_t)((col - WID) >> 2) * gstride + q * 16;
        uint32_t S[4] = {0, 0, 0, 0}, SS[4] = {0, 0, 0, 0};
#pragma unroll
        for (int g = 0; g < NG; ++g) {
            const uint32_t mg = group_mask(o, K, g);
            if (mg == 0u) continue;
#pragma unroll
            for (int rr = 0; rr < K; ++rr) {
                const uint4 w4 = __ldg(reinterpret_cast<const uint4*>(pq + rr * rowpitch + g * gstride));
                const uint32_t w[4] = {w4.x & mg, w4.y & mg, w4.z & mg, w4.w & mg};
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    S[k] = __dp4a(w[k], 0x01010101u, S[k]);
                    SS[k] = __dp4a(w[k], w[k], SS[k]);
                }
            }
        }
        // n*SS and S*S are <= 225^2 * 255^2 < 2^32 and n*SS >= S*S: exact in u32
        uint4 var;
        var.x = NPIX * SS[0] - S[0] * S[0];
        var.y = NPIX * SS[1] - S[1] * S[1];
        var.z = NPIX * SS[2] - S[2] * S[2];
        var.w = NPIX * SS[3] - S[3] * S[3];
        uint2 s2;
        s2.x = S[0] | (S[1] << 16);
        s2.y = S[2] | (S[3] << 16);
        *reinterpret_cast<uint4*>(vmap + px * Vp + 4 * q) = var;
        *reinterpret_cast<uint2*>(smap + px * Vp + 4 * q) = s2;
    }
}

int mvs_build_window_maps(mvs_ctx* ctx, int wid, cudaStream_t s) {
    if (ctx->maps_wid == wid && ctx->d_smap && ctx->d_vmap) return MVS_OK;
    if (ctx->maps_wid >= 0) MVS_CUDA_CHECK(cudaDeviceSynchronize());   // rebuilding for another wid: no reader may be in flight
    const size_t n = (size_t)ctx->H * ctx->W * ctx->Vp;
    int rc;
    if ((rc = mvs_ensure((void**)&ctx->d_smap, &ctx->smap_bytes, n * sizeof(uint16_t) + 256, "window-sum map")) != MVS_OK ||
        (rc = mvs_ensure((void**)&ctx->d_vmap, &ctx->vmap_bytes, n * sizeof(uint32_t) + 256, "window-variance map")) != MVS_OK)
        return rc;
    ctx->maps_wid = -1;
    MVS_CUDA_CHECK(cudaMemsetAsync(ctx->d_smap, 0, n * sizeof(uint16_t), s));
    MVS_CUDA_CHECK(cudaMemsetAsync(ctx->d_vmap, 0, n * sizeof(uint32_t), s));
    const int64_t total = (int64_t)ctx->H * ctx->W * ctx->Q;
    int64_t blocks = (total + 255) / 256;
    const int64_t cap = (int64_t)ctx->sm_count * 32;
    if (blocks > cap) blocks = cap;
#define MVS_MAPS(WID_)                                                                                              \
    case WID_:                                                                                                      \
        build_window_maps<WID_><<<(int)blocks, 256, 0, s>>>(ctx->d_gray, ctx->d_smap, ctx->d_vmap, ctx->Vp, ctx->Q, \
                                                            ctx->H, ctx->W, ctx->gstride, ctx->rowpitch);           \
        break;
    switch (wid) {
        MVS_MAPS(1) MVS_MAPS(2) MVS_MAPS(3) MVS_MAPS(4) MVS_MAPS(5) MVS_MAPS(6) MVS_MAPS(7)
        default: mvs_set_error("wid %d not supported (1..7)", wid); return MVS_ERR_ARG;
    }
#undef MVS_MAPS
    ctx->launches++;
    MVS_CUDA_CHECK(cudaGetLastError());
    // The maps are shared by every later call on ANY stream: the build (once per context and wid, ~0.1 ms) is
    // waited for here, so that no other stream can ever read half-built maps.  (Not possible while `s` is being
    // captured into a CUDA graph: build the maps with one eager call before capturing.)
    cudaStreamCaptureStatus capst = cudaStreamCaptureStatusNone;
    MVS_CUDA_CHECK(cudaStreamIsCapturing(s, &capst));
    if (capst == cudaStreamCaptureStatusNone) MVS_CUDA_CHECK(cudaStreamSynchronize(s));
    ctx->maps_wid = wid;
    return MVS_OK;
}

template <int LPH>
__device__ __forceinline__ uint32_t hyp_mask(int sub) {
    if constexpr (LPH == 32) {
        return FULL;
    } else {
        return ((1u << LPH) - 1u) << (sub * LPH);
    }
}

// ncc = (num / sqrt(var_i * var_r)) * n/(n-1), oracle/mode_a.py::score operation order,
// with the correctly rounded fp64 sqrt and division.  Out of line: it only runs for values
// closer than 1e-9 to the threshold.
__device__ __noinline__ double ncc_exact(double num, double var_i, double var_r, double cn) {
    return (num / sqrt(var_i * var_r)) * cn;
}

// u32 / s32 -> fp64 on the fp64 pipe (magic-number add) instead of the conversion unit
__device__ __forceinline__ double u32_to_double(uint32_t u) {
    return __hiloint2double(0x43300000, (int)u) - 4503599627370496.0;                 // 2^52 + u - 2^52
}
__device__ __forceinline__ double s32_to_double(int i) {
    return __hiloint2double(0x43300000, i ^ 0x80000000) - 4503601774854144.0;         // 2^52 + 2^31
}

// Fast path of the ratio: y0 = rsqrt.approx(den) (MUFU.RSQ64H, ~2^-22) plus one Newton
// step gives rsqrt(den) to ~1e-13 relative.  var_r_scaled = var_r * ((n-1)/n)^2 folds the
// n/(n-1) factor in, so val = num * rsqrt(var_i * var_r_scaled).  The caller re-does any
// value within 1e-9 of the threshold with ncc_exact, so the strict '>' decision is always
// the oracle's; the value itself differs from the oracle's by < 1e-12.
__device__ __forceinline__ double ncc_fast(double num, double var_i, double var_r_scaled) {
    const double den = var_i * var_r_scaled;
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(den));
    const double t = y * y;
    const double e = fma(-0.5 * den, t, 1.5);
    return num * (y * e);
}

struct ScoreArgs {
    const uint8_t* gray4;
    const uint16_t* smap;
    const uint32_t* vmap;
    int V, Vp, Q, W;
    int64_t gstride, rowpitch;
    const int32_t* ref;
    double thr;
    uint64_t* vis_out;
    double* avg_out;
    int32_t* count_out;
    float* ncc_out;
};

// Score NB hypotheses (NB = 1 or 2) that share the window row AND the first pixel group,
// i.e. the same (row, (col - WID) >> 2): they read exactly the same 16-byte quads of every
// view, so each quad is loaded ONCE and multiplied against NB reference windows.  GS is the
// compile-time byte stride between pixel groups (4*Vp) or 0 = read it from the arguments.
template <int WID, int LPH, int NB, int GS, bool WANT_NCC, int UNR = 0>
__device__ __forceinline__ void score_block(const ScoreArgs& A, const uint32_t (&anchor)[NB], const int64_t (&h)[NB],
                                            uint32_t (*sref)[2 * WID + 1][(2 * WID + 7) / 4], int lih, uint32_t hmask) {
    constexpr int K = 2 * WID + 1;
    constexpr int NG = (K + 6) / 4;            // pixel groups a K-pixel run at offset 0..3 can touch
    constexpr int NPIX = K * K;
    const int64_t gstride = GS ? (int64_t)GS : A.gstride;
    const int mask_words32 = 2 * ((A.V + 63) >> 6);        // 32-bit chunks per hypothesis in vis_out
    const int passes = (LPH == 32) ? (A.Q + 31) >> 5 : 1;
    const double cn = (double)NPIX / (double)(NPIX - 1);
    const double thr = A.thr;

    const int row = (int)(anchor[0] >> 16);
    const int cg = ((int)(anchor[0] & 0xffffu) - WID) >> 2;
    const uint8_t* base = A.gray4 + (int64_t)(row - WID) * A.rowpitch + (int64_t)cg * gstride;
    int r[NB], o[NB], Sr[NB];
    int64_t mi[NB];
    double var_r[NB], var_rs[NB];
    bool need_last = false;                                // only the last pixel group can be empty
#pragma unroll
    for (int b = 0; b < NB; ++b) {
        const int col = (int)(anchor[b] & 0xffffu);
        r[b] = __ldg(A.ref + h[b]);
        o[b] = (col - WID) & 3;
        mi[b] = ((int64_t)row * A.W + col) * A.Vp;
        Sr[b] = (int)__ldg(A.smap + mi[b] + r[b]);
        var_r[b] = (double)__ldg(A.vmap + mi[b] + r[b]);
        var_rs[b] = var_r[b] * ((double)((NPIX - 1) * (NPIX - 1)) / (double)(NPIX * NPIX));
        need_last |= group_mask(o[b], K, NG - 1) != 0u;
    }
    // an unused last group re-reads group 0 (an L1 hit) against zero reference words
    const uint32_t lastoff = need_last ? (uint32_t)((NG - 1) * gstride) : 0u;

    // ---- reference windows: masked words to shared memory
    __syncwarp(hmask);                                     // the previous block's readers are done
#pragma unroll
    for (int b = 0; b < NB; ++b) {
        if constexpr (LPH % NG == 0) {                     // a lane always stages the same pixel group
            const int g = lih % NG;
            const uint32_t mg = group_mask(o[b], K, g);
            const uint8_t* pr = base + g * gstride + 4 * r[b] + (lih / NG) * A.rowpitch;
            for (int rr = lih / NG; rr < K; rr += LPH / NG) {
                sref[b][rr][g] = mg ? (__ldg(reinterpret_cast<const uint32_t*>(pr)) & mg) : 0u;
                pr += (LPH / NG) * A.rowpitch;
            }
        } else {
            for (int idx = lih; idx < K * NG; idx += LPH) {
                const int rr = idx / NG, g = idx - rr * NG;
                const uint32_t mg = group_mask(o[b], K, g);
                uint32_t w = 0u;
                if (mg) w = __ldg(reinterpret_cast<const uint32_t*>(base + rr * A.rowpitch + g * gstride + 4 * r[b])) & mg;
                sref[b][rr][g] = w;
            }
        }
    }
    __syncwarp(hmask);

    double acc[NB];
    int count[NB];
    uint32_t myword[NB];
#pragma unroll
    for (int b = 0; b < NB; ++b) {
        acc[b] = 0.0;
        count[b] = 0;
        myword[b] = 0u;
    }
    for (int p = 0; p < passes; ++p) {
        const int qq = p * LPH + lih;                      // this lane's quad: views 4qq..4qq+3
        const bool act = qq < A.Q;
        const int qc = act ? qq : A.Q - 1;                 // idle lanes shadow the last quad (same L1 lines)
        uint2 s2[NB];
        uint4 v4[NB];
#pragma unroll
        for (int b = 0; b < NB; ++b) {
            s2[b] = __ldg(reinterpret_cast<const uint2*>(A.smap + mi[b] + 4 * qc));
            v4[b] = __ldg(reinterpret_cast<const uint4*>(A.vmap + mi[b] + 4 * qc));
        }
        int SAB[NB][4];
#pragma unroll
        for (int b = 0; b < NB; ++b)
#pragma unroll
            for (int k = 0; k < 4; ++k) SAB[b][k] = 0;
        const uint8_t* prow = base + (int64_t)qc * 16;
        // UNR window rows per loop body (0 = all K): the fully unrolled pair + single instances of the 32-lane variants
        // are 730 + 501 instructions and miss the instruction cache (ncu "no instruction" stall 2.4 per issue at 128 views)
#pragma unroll(UNR ? UNR : K)
        for (int rr = 0; rr < K; ++rr) {
            uint32_t rw[NB][NG];
#pragma unroll
            for (int b = 0; b < NB; ++b) {
                if constexpr (NG == 4) {
                    const uint4 t4 = *reinterpret_cast<const uint4*>(&sref[b][rr][0]);
                    rw[b][0] = t4.x; rw[b][1] = t4.y; rw[b][2] = t4.z; rw[b][3] = t4.w;
                } else {
#pragma unroll
                    for (int g = 0; g < NG; ++g) rw[b][g] = sref[b][rr][g];
                }
            }
#pragma unroll
            for (int g = 0; g < NG; ++g) {
                const uint8_t* pw = (g == NG - 1) ? prow + lastoff : prow + g * gstride;
                const uint4 w4 = __ldg(reinterpret_cast<const uint4*>(pw));
#pragma unroll
                for (int b = 0; b < NB; ++b) {
                    SAB[b][0] = dp4a_u(w4.x, rw[b][g], SAB[b][0]);
                    SAB[b][1] = dp4a_u(w4.y, rw[b][g], SAB[b][1]);
                    SAB[b][2] = dp4a_u(w4.z, rw[b][g], SAB[b][2]);
                    SAB[b][3] = dp4a_u(w4.w, rw[b][g], SAB[b][3]);
                }
            }
            prow += A.rowpitch;
        }
        // ---- this lane finishes its own four views of every hypothesis of the block
#pragma unroll
        for (int b = 0; b < NB; ++b) {
            const int Sv[4] = {(int)(s2[b].x & 0xffffu), (int)(s2[b].x >> 16), (int)(s2[b].y & 0xffffu), (int)(s2[b].y >> 16)};
            const uint32_t var[4] = {v4[b].x, v4[b].y, v4[b].z, v4[b].w};
            // views this lane may score: inside V (padding views have zero variance anyway), not the
            // reference view itself, and only when the reference window has variance
            uint32_t okmask = (act && var_r[b] != 0.0) ? 0xfu : 0u;
            if ((r[b] >> 2) == qq) okmask &= ~(1u << (r[b] & 3));
            if (4 * qq + 3 >= A.V) okmask &= (1u << max(A.V - 4 * qq, 0)) - 1u;
            auto numerator = [&](int k) {
                // exact integers: n*SAB, S*Sr <= 225^2 * 255^2 need 64 bits at wid > 5
                return (WID <= 5) ? s32_to_double(NPIX * SAB[b][k] - Sv[k] * Sr[b])
                                  : (double)((long long)NPIX * SAB[b][k] - (long long)Sv[k] * Sr[b]);
            };
            uint32_t near = 0u, nib = 0u;
            float dump[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const double val = ncc_fast(numerator(k), u32_to_double(var[k]), var_rs[b]);
                if (var[k] == 0u) okmask &= ~(1u << k);
                const bool vis = ((okmask >> k) & 1u) && (val > thr);
                acc[b] += vis ? val : 0.0;
                nib |= vis ? (1u << k) : 0u;
                if (fabs(val - thr) < 1e-9) near |= 1u << k;
                if constexpr (WANT_NCC) dump[k] = (float)val;
            }
            near &= okmask;
            if (__builtin_expect(near != 0u, 0)) {
                // rare: a value within 1e-9 of the threshold is redone with the correctly rounded sqrt and
                // division and the provisional decision / sum contribution replaced
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    if (!((near >> k) & 1u)) continue;
                    const double num = numerator(k), vd = u32_to_double(var[k]);
                    const double fast = ncc_fast(num, vd, var_rs[b]);
                    const double exact = ncc_exact(num, vd, var_r[b], cn);
                    if ((nib >> k) & 1u) acc[b] -= fast;
                    nib &= ~(1u << k);
                    if (exact > thr) {
                        acc[b] += exact;
                        nib |= 1u << k;
                    }
                    if constexpr (WANT_NCC) dump[k] = (float)exact;
                }
            }
            if constexpr (WANT_NCC) {
                if (act) {
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        if (4 * qq + k < A.V) A.ncc_out[h[b] * A.V + 4 * qq + k] = ((okmask >> k) & 1u) ? dump[k] : nanf("");
                }
            }
            // the lane group covers LPH*4 mask bits per pass = max(LPH/8, 1) 32-bit words
            constexpr int WPP = LPH >= 8 ? LPH / 8 : 1;
#pragma unroll
            for (int w = 0; w < WPP; ++w) {
                const uint32_t contrib = ((lih >> 3) == w) ? (nib << (4 * (lih & 7))) : 0u;
                const uint32_t word = __reduce_or_sync(hmask, contrib);
                count[b] += __popc(word);
                if (lih == p * WPP + w) myword[b] = word;  // word index < 32 <=> V <= 1024
            }
        }
    }
#pragma unroll
    for (int b = 0; b < NB; ++b) {
        double a = acc[b];
#pragma unroll
        for (int s = LPH / 2; s > 0; s >>= 1) a += __shfl_xor_sync(hmask, a, s);
        for (int w = lih; w < mask_words32; w += LPH)
            reinterpret_cast<uint32_t*>(A.vis_out)[h[b] * mask_words32 + w] = (w == lih) ? myword[b] : 0u;
        if (lih == 0) {
            A.count_out[h[b]] = count[b];
            if (A.avg_out) A.avg_out[h[b]] = count[b] > 0 ? a / (double)count[b] : 0.0;
        }
    }
}

// ---------------------------------------------------------------------------------
// K1: LPH lanes per hypothesis PAIR, 32/LPH pairs per warp.  A CTA walks chunks of
// consecutive positions of the (row, pixel-group)-ordered batch so that its lane groups
// share L1 lines; two neighbouring positions with the same (row, pixel group) are scored
// as one block with shared loads.  anchors[i] = row<<16|col of position i
// (MVS_ANCHOR_INVALID: rejected, result already written by bin_project); order[i] =
// hypothesis index (NULL: identity) -- packed as entries[i] = (index, anchor) for ordered batches.
// ---------------------------------------------------------------------------------
template <int WID, int LPH, int GS, int MINB, bool WANT_NCC, int UNR = 0>
__global__ void __launch_bounds__(256, MINB)
    ncc_score_gather(const ScoreArgs A, int64_t N, const uint32_t* __restrict__ anchors, const uint2* __restrict__ entries) {
    constexpr int K = 2 * WID + 1;
    constexpr int NG = (K + 6) / 4;
    constexpr int HPW = 32 / LPH;
    constexpr int ITERS = MVS_K1_ITERS;
    constexpr int CHUNK = 8 * HPW * 2 * ITERS;
    __shared__ __align__(16) uint32_t s_ref[8][HPW][2][K][NG];

    const int lane = threadIdx.x & 31;
    const int wib = threadIdx.x >> 5;
    const int sub = lane / LPH;
    const int lih = lane % LPH;
    const uint32_t hmask = hyp_mask<LPH>(sub);
    uint32_t(*sref)[K][NG] = s_ref[wib][sub];

    for (int64_t i0 = (int64_t)blockIdx.x * CHUNK; i0 < N; i0 += (int64_t)gridDim.x * CHUNK) {
        for (int it = 0; it < ITERS; ++it) {
            const int64_t i = i0 + 2 * ((it * 8 + wib) * HPW + sub);
            if (i >= N) continue;                          // the whole lane group leaves together
            uint32_t a0, a1 = MVS_ANCHOR_INVALID;
            int64_t h0 = i, h1 = i + 1;
            if (entries) {                                 // ordered batch: (hypothesis index, anchor) pairs
                const uint2 e0 = __ldg(entries + i);
                h0 = e0.x;
                a0 = e0.y;
                if (i + 1 < N) {
                    const uint2 e1 = __ldg(entries + i + 1);
                    h1 = e1.x;
                    a1 = e1.y;
                }
            } else {
                a0 = __ldg(anchors + i);
                if (i + 1 < N) a1 = __ldg(anchors + i + 1);
            }
            const bool ok0 = a0 != MVS_ANCHOR_INVALID, ok1 = a1 != MVS_ANCHOR_INVALID;
            const bool same = ok0 && ok1 && ((a0 >> 16) == (a1 >> 16)) &&
                              ((((int)(a0 & 0xffffu) - WID) >> 2) == (((int)(a1 & 0xffffu) - WID) >> 2));
            if (same) {
                const uint32_t aa[2] = {a0, a1};
                const int64_t hh[2] = {h0, h1};
                score_block<WID, LPH, 2, GS, WANT_NCC, UNR>(A, aa, hh, sref, lih, hmask);
            } else {
                const uint32_t a2[2] = {a0, a1};
                const int64_t h2[2] = {h0, h1};
#pragma unroll 1
                for (int t = 0; t < 2; ++t) {              // one code instance for both singles
                    if (a2[t] == MVS_ANCHOR_INVALID) continue;
                    const uint32_t aa[1] = {a2[t]};
                    const int64_t hh[1] = {h2[t]};
                    score_block<WID, LPH, 1, GS, WANT_NCC, UNR>(A, aa, hh, sref, lih, hmask);
                }
            }
        }
    }
}



// ---------------------------------------------------------------------------------
// K1 for 33..48 views (dinoRing's 48, the temple-shaped 47): "quad + pair" lanes.
//
// EIGHT lanes own a block of two hypotheses, four blocks per warp.  Lane l owns the view quad
// 4l..4l+3 (one LDG.128 per window row and pixel group) AND the view pair 32+2l, 33+2l (one
// LDG.64), i.e. six views: every lane is busy at 48 views (the 16-lane mapping idles a quarter of
// them) and a warp instruction returns only bytes that are used -- 6 data-pipe wavefronts per four
// blocks instead of 4 per two.  Control flow is WARP-UNIFORM: each lane group walks its own run of
// PER consecutive positions of the ordered batch, pairing a position with its successor when both
// read the same quads (same window row and first pixel group); a lone position is scored as a
// block whose second member is a masked duplicate, a finished group idles on duplicates of its last
// block.  All shuffles therefore use the full mask (no WARPSYNC/ENDCOLLECTIVE sequences around
// partial-mask collectives) and stay inside their 8-lane segment.
// ---------------------------------------------------------------------------------
#ifndef MVS_K6_PER
#define MVS_K6_PER 8            // positions per lane-group run: measured 0.331 ms (8) / 0.343 (4) / 0.358 (16) / 0.363 (32) / 0.437 (64) per 2^20 hypotheses
#endif

__device__ __forceinline__ uint32_t seg8_or(uint32_t v) {
    v |= __shfl_xor_sync(FULL, v, 1);
    v |= __shfl_xor_sync(FULL, v, 2);
    v |= __shfl_xor_sync(FULL, v, 4);
    return v;
}

template <int WID, int GS, int MINB, bool WANT_NCC>
__global__ void __launch_bounds__(256, MINB)
    ncc_score_gather6(const ScoreArgs A, int64_t N, const uint32_t* __restrict__ anchors, const uint2* __restrict__ entries,
                      int PER) {
    constexpr int K = 2 * WID + 1;
    constexpr int NG = (K + 6) / 4;
    constexpr int NPIX = K * K;
    const int CHUNK = 32 * PER;                            // 8 warps x 4 lane groups x PER positions
    __shared__ __align__(16) uint32_t s_ref[8][4][2][K][NG];

    const int lane = threadIdx.x & 31;
    const int wib = threadIdx.x >> 5;
    const int grp = lane >> 3;
    const int lih = lane & 7;
    uint32_t(*sref)[K][NG] = s_ref[wib][grp];
    const int64_t gstride = GS ? (int64_t)GS : A.gstride;
    const int Q2 = (A.Vp - 32) >> 1;                       // lanes that also own a view pair (1..8)
    const int lp = lih < Q2 ? lih : Q2 - 1;                // lanes past the last pair shadow it (same L1 lines)
    const double cn = (double)NPIX / (double)(NPIX - 1);
    const double thr = A.thr;
    // views this lane may score at all: inside V, and the pair only on lanes that own one
    uint32_t lane_views = 0u;
#pragma unroll
    for (int k = 0; k < 4; ++k)
        if (4 * lih + k < A.V) lane_views |= 1u << k;
#pragma unroll
    for (int k = 0; k < 2; ++k)
        if (lih < Q2 && 32 + 2 * lih + k < A.V) lane_views |= 16u << k;

    for (int64_t i0 = (int64_t)blockIdx.x * CHUNK; i0 < N; i0 += (int64_t)gridDim.x * CHUNK) {
        int64_t pos = i0 + (int64_t)((wib * 4 + grp) * PER);
        const int64_t end = (pos + PER < N) ? pos + PER : N;
        while (__any_sync(FULL, pos < end)) {
            // ---- this group's next block: position pos, paired with pos + 1 when both read the same quads
            uint32_t a[2] = {MVS_ANCHOR_INVALID, MVS_ANCHOR_INVALID};
            int64_t h[2] = {0, 0};
            if (pos < end) {
                if (entries) {
                    const uint2 e0 = __ldg(entries + pos);
                    h[0] = e0.x;
                    a[0] = e0.y;
                    if (pos + 1 < end) {
                        const uint2 e1 = __ldg(entries + pos + 1);
                        h[1] = e1.x;
                        a[1] = e1.y;
                    }
                } else {
                    h[0] = pos;
                    a[0] = __ldg(anchors + pos);
                    if (pos + 1 < end) {
                        h[1] = pos + 1;
                        a[1] = __ldg(anchors + pos + 1);
                    }
                }
            }
            bool live[2];
            live[0] = a[0] != MVS_ANCHOR_INVALID;
            live[1] = live[0] && a[1] != MVS_ANCHOR_INVALID && ((a[0] >> 16) == (a[1] >> 16)) &&
                      ((((int)(a[0] & 0xffffu) - WID) >> 2) == (((int)(a[1] & 0xffffu) - WID) >> 2));
            pos += live[1] ? 2 : 1;
            if (!__any_sync(FULL, live[0])) continue;      // nothing to score in the whole warp (rejected hypotheses)
            if (!live[0]) a[0] = ((uint32_t)WID << 16) | (uint32_t)(WID + 1);   // idle group: a window that is always inside
            if (!live[1]) {
                a[1] = a[0];
                h[1] = h[0];
            }

            const int row = (int)(a[0] >> 16);
            const int cg = ((int)(a[0] & 0xffffu) - WID) >> 2;
            const uint8_t* base = A.gray4 + (int64_t)(row - WID) * A.rowpitch + (int64_t)cg * gstride;
            int r[2], o[2], Sr[2];
            int64_t mi[2];
            double var_r[2], var_rs[2];
            bool need_last = false;
#pragma unroll
            for (int b = 0; b < 2; ++b) {
                const int col = (int)(a[b] & 0xffffu);
                r[b] = live[0] ? __ldg(A.ref + h[b]) : 0;
                o[b] = (col - WID) & 3;
                mi[b] = ((int64_t)row * A.W + col) * A.Vp;
                Sr[b] = (int)__ldg(A.smap + mi[b] + r[b]);
                var_r[b] = (double)__ldg(A.vmap + mi[b] + r[b]);
                var_rs[b] = var_r[b] * ((double)((NPIX - 1) * (NPIX - 1)) / (double)(NPIX * NPIX));
                need_last |= group_mask(o[b], K, NG - 1) != 0u;
            }
            const uint32_t lastoff = need_last ? (uint32_t)((NG - 1) * gstride) : 0u;

            // ---- reference windows: masked words to shared memory
            __syncwarp();
#pragma unroll
            for (int b = 0; b < 2; ++b) {
                if constexpr (NG == 4) {                   // a lane always stages the same pixel group
                    const int g = lih & 3;
                    const uint32_t mg = group_mask(o[b], K, g);
                    const uint8_t* pr = base + g * gstride + 4 * r[b] + (lih >> 2) * A.rowpitch;
#pragma unroll
                    for (int rr0 = 0; rr0 < K; rr0 += 2) {
                        const int rr = rr0 + (lih >> 2);
                        if (rr < K) sref[b][rr][g] = __ldg(reinterpret_cast<const uint32_t*>(pr)) & mg;
                        pr += 2 * A.rowpitch;
                    }
                } else {
                    for (int idx = lih; idx < K * NG; idx += 8) {
                        const int rr = idx / NG, g = idx - rr * NG;
                        const uint32_t mg = group_mask(o[b], K, g);
                        sref[b][rr][g] = __ldg(reinterpret_cast<const uint32_t*>(base + rr * A.rowpitch + g * gstride + 4 * r[b])) & mg;
                    }
                }
            }
            __syncwarp();

            // ---- window sums of this lane's six views (maps) and the dot products
            uint2 s4[2];
            uint32_t s2[2];
            uint4 v4[2];
            uint2 v2[2];
#pragma unroll
            for (int b = 0; b < 2; ++b) {
                s4[b] = __ldg(reinterpret_cast<const uint2*>(A.smap + mi[b] + 4 * lih));
                v4[b] = __ldg(reinterpret_cast<const uint4*>(A.vmap + mi[b] + 4 * lih));
                s2[b] = __ldg(reinterpret_cast<const uint32_t*>(A.smap + mi[b] + 32 + 2 * lp));
                v2[b] = __ldg(reinterpret_cast<const uint2*>(A.vmap + mi[b] + 32 + 2 * lp));
            }
            int SAB[2][6];
#pragma unroll
            for (int b = 0; b < 2; ++b)
#pragma unroll
                for (int k = 0; k < 6; ++k) SAB[b][k] = 0;
            const uint8_t* p4 = base + 16 * lih;
            const uint8_t* p2 = base + 128 + 8 * lp;
#pragma unroll
            for (int rr = 0; rr < K; ++rr) {
                uint32_t rw[2][NG];
#pragma unroll
                for (int b = 0; b < 2; ++b) {
                    if constexpr (NG == 4) {
                        const uint4 t4 = *reinterpret_cast<const uint4*>(&sref[b][rr][0]);
                        rw[b][0] = t4.x; rw[b][1] = t4.y; rw[b][2] = t4.z; rw[b][3] = t4.w;
                    } else {
#pragma unroll
                        for (int g = 0; g < NG; ++g) rw[b][g] = sref[b][rr][g];
                    }
                }
#pragma unroll
                for (int g = 0; g < NG; ++g) {
                    const uint32_t off = (g == NG - 1) ? lastoff : (uint32_t)(g * gstride);
                    const uint4 w4 = __ldg(reinterpret_cast<const uint4*>(p4 + off));
                    const uint2 w2 = __ldg(reinterpret_cast<const uint2*>(p2 + off));
#pragma unroll
                    for (int b = 0; b < 2; ++b) {
                        SAB[b][0] = dp4a_u(w4.x, rw[b][g], SAB[b][0]);
                        SAB[b][1] = dp4a_u(w4.y, rw[b][g], SAB[b][1]);
                        SAB[b][2] = dp4a_u(w4.z, rw[b][g], SAB[b][2]);
                        SAB[b][3] = dp4a_u(w4.w, rw[b][g], SAB[b][3]);
                        SAB[b][4] = dp4a_u(w2.x, rw[b][g], SAB[b][4]);
                        SAB[b][5] = dp4a_u(w2.y, rw[b][g], SAB[b][5]);
                    }
                }
                p4 += A.rowpitch;
                p2 += A.rowpitch;
            }

            // ---- this lane finishes its own six views of both hypotheses
#pragma unroll
            for (int b = 0; b < 2; ++b) {
                const int Sv[6] = {(int)(s4[b].x & 0xffffu), (int)(s4[b].x >> 16), (int)(s4[b].y & 0xffffu), (int)(s4[b].y >> 16),
                                   (int)(s2[b] & 0xffffu),   (int)(s2[b] >> 16)};
                const uint32_t var[6] = {v4[b].x, v4[b].y, v4[b].z, v4[b].w, v2[b].x, v2[b].y};
                uint32_t okmask = (live[b] && var_r[b] != 0.0) ? lane_views : 0u;
                if ((r[b] >> 2) == lih) okmask &= ~(1u << (r[b] & 3));                       // the reference view itself
                if (r[b] >= 32 && ((r[b] - 32) >> 1) == lih) okmask &= ~(16u << (r[b] & 1));
                auto numerator = [&](int k) {
                    return (WID <= 5) ? s32_to_double(NPIX * SAB[b][k] - Sv[k] * Sr[b])
                                      : (double)((long long)NPIX * SAB[b][k] - (long long)Sv[k] * Sr[b]);
                };
                uint32_t near = 0u, bits = 0u;
                double acc = 0.0;
                float dump[6];
#pragma unroll
                for (int k = 0; k < 6; ++k) {
                    const double val = ncc_fast(numerator(k), u32_to_double(var[k]), var_rs[b]);
                    if (var[k] == 0u) okmask &= ~(1u << k);
                    const bool vis = ((okmask >> k) & 1u) && (val > thr);
                    acc += vis ? val : 0.0;
                    bits |= vis ? (1u << k) : 0u;
                    if (fabs(val - thr) < 1e-9) near |= 1u << k;
                    if constexpr (WANT_NCC) dump[k] = (float)val;
                }
                near &= okmask;
                if (__builtin_expect(near != 0u, 0)) {
                    // rare: a value within 1e-9 of the threshold is redone with the correctly rounded sqrt and
                    // division and the provisional decision / sum contribution replaced
#pragma unroll
                    for (int k = 0; k < 6; ++k) {
                        if (!((near >> k) & 1u)) continue;
                        const double num = numerator(k), vd = u32_to_double(var[k]);
                        const double fast = ncc_fast(num, vd, var_rs[b]);
                        const double exact = ncc_exact(num, vd, var_r[b], cn);
                        if ((bits >> k) & 1u) acc -= fast;
                        bits &= ~(1u << k);
                        if (exact > thr) {
                            acc += exact;
                            bits |= 1u << k;
                        }
                        if constexpr (WANT_NCC) dump[k] = (float)exact;
                    }
                }
                if constexpr (WANT_NCC) {
                    if (live[b]) {
#pragma unroll
                        for (int k = 0; k < 6; ++k) {
                            const int v = k < 4 ? 4 * lih + k : 32 + 2 * lih + (k - 4);
                            if ((lane_views >> k) & 1u) A.ncc_out[h[b] * A.V + v] = ((okmask >> k) & 1u) ? dump[k] : nanf("");
                        }
                    }
                }
                // visible mask: bits 4l..4l+3 of the low word, bits 2l, 2l+1 of the high word
                const uint32_t lo = seg8_or((bits & 15u) << (4 * lih));
                const uint32_t hi = seg8_or((bits >> 4) << (2 * lih));
#pragma unroll
                for (int sft = 4; sft > 0; sft >>= 1) acc += __shfl_xor_sync(FULL, acc, sft);
                const int count = __popc(lo) + __popc(hi);
                if (live[b] && lih == 0) {
                    reinterpret_cast<uint2*>(A.vis_out)[h[b]] = make_uint2(lo, hi);
                    A.count_out[h[b]] = count;
                    if (A.avg_out) A.avg_out[h[b]] = count > 0 ? acc / (double)count : 0.0;
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------
// gather_probe: the measured CEILING of K1's memory side ("the builder must measure the
// L2 gather ceiling with a micro-benchmark").  Same traversal of the ordered batch, same lane
// mapping, same pairing of positions that share (row, pixel group), and exactly the loads K1
// issues for a block -- the 16-byte view quads of every window row and pixel group, the
// reference-window words, the two map entries per lane, ref / smap / vmap of the reference view --
// but NO arithmetic: every loaded word is XOR-folded into one register so that the loads stay
// live, and nothing is stored.  bench.py times it on the same inputs right after K1;
// roofline.frac = probe time / K1 time.
// ---------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t fold4(uint32_t acc, const uint4& w) {
    return (acc ^ w.x ^ w.y) ^ (w.z ^ w.w);
}

template <int WID, int LPH, int NB, int GS>
__device__ __forceinline__ uint32_t probe_block(const ScoreArgs& A, const uint32_t (&anchor)[NB], const int64_t (&h)[NB],
                                                int lih) {
    constexpr int K = 2 * WID + 1;
    constexpr int NG = (K + 6) / 4;
    const int64_t gstride = GS ? (int64_t)GS : A.gstride;
    const int passes = (LPH == 32) ? (A.Q + 31) >> 5 : 1;
    const int row = (int)(anchor[0] >> 16);
    const int cg = ((int)(anchor[0] & 0xffffu) - WID) >> 2;
    const uint8_t* base = A.gray4 + (int64_t)(row - WID) * A.rowpitch + (int64_t)cg * gstride;
    uint32_t acc = 0u;
    int64_t mi[NB];
    bool need_last = false;
#pragma unroll
    for (int b = 0; b < NB; ++b) {
        const int col = (int)(anchor[b] & 0xffffu);
        const int r = __ldg(A.ref + h[b]);
        mi[b] = ((int64_t)row * A.W + col) * A.Vp;
        acc ^= (uint32_t)__ldg(A.smap + mi[b] + r) ^ __ldg(A.vmap + mi[b] + r);
        need_last |= group_mask((col - WID) & 3, K, NG - 1) != 0u;
        // the reference-window words K1 stages in shared memory
        for (int idx = lih; idx < K * NG; idx += LPH) {
            const int rr = idx / NG, g = idx - rr * NG;
            acc ^= __ldg(reinterpret_cast<const uint32_t*>(base + rr * A.rowpitch + g * gstride + 4 * r));
        }
    }
    const uint32_t lastoff = need_last ? (uint32_t)((NG - 1) * gstride) : 0u;
    for (int p = 0; p < passes; ++p) {
        const int qq = p * LPH + lih;
        const int qc = qq < A.Q ? qq : A.Q - 1;
#pragma unroll
        for (int b = 0; b < NB; ++b) {
            const uint2 s2 = __ldg(reinterpret_cast<const uint2*>(A.smap + mi[b] + 4 * qc));
            const uint4 v4 = __ldg(reinterpret_cast<const uint4*>(A.vmap + mi[b] + 4 * qc));
            acc = fold4(acc ^ s2.x ^ s2.y, v4);
        }
        const uint8_t* prow = base + (int64_t)qc * 16;
#pragma unroll
        for (int rr = 0; rr < K; ++rr) {
#pragma unroll
            for (int g = 0; g < NG; ++g) {
                const uint8_t* pw = (g == NG - 1) ? prow + lastoff : prow + g * gstride;
                acc = fold4(acc, __ldg(reinterpret_cast<const uint4*>(pw)));
            }
            prow += A.rowpitch;
        }
    }
    return acc;
}

template <int WID, int LPH, int GS, int MINB>
__global__ void __launch_bounds__(256, MINB)
    gather_probe(const ScoreArgs A, int64_t N, const uint32_t* __restrict__ anchors, const uint2* __restrict__ entries,
                 uint32_t* __restrict__ sink) {
    constexpr int HPW = 32 / LPH;
    constexpr int ITERS = MVS_K1_ITERS;
    constexpr int CHUNK = 8 * HPW * 2 * ITERS;
    const int lane = threadIdx.x & 31;
    const int wib = threadIdx.x >> 5;
    const int sub = lane / LPH;
    const int lih = lane % LPH;
    uint32_t acc = 0u;
    for (int64_t i0 = (int64_t)blockIdx.x * CHUNK; i0 < N; i0 += (int64_t)gridDim.x * CHUNK) {
        for (int it = 0; it < ITERS; ++it) {
            const int64_t i = i0 + 2 * ((it * 8 + wib) * HPW + sub);
            if (i >= N) continue;
            uint32_t a0, a1 = MVS_ANCHOR_INVALID;
            int64_t h0 = i, h1 = i + 1;
            if (entries) {
                const uint2 e0 = __ldg(entries + i);
                h0 = e0.x;
                a0 = e0.y;
                if (i + 1 < N) {
                    const uint2 e1 = __ldg(entries + i + 1);
                    h1 = e1.x;
                    a1 = e1.y;
                }
            } else {
                a0 = __ldg(anchors + i);
                if (i + 1 < N) a1 = __ldg(anchors + i + 1);
            }
            const bool ok0 = a0 != MVS_ANCHOR_INVALID, ok1 = a1 != MVS_ANCHOR_INVALID;
            const bool same = ok0 && ok1 && ((a0 >> 16) == (a1 >> 16)) &&
                              ((((int)(a0 & 0xffffu) - WID) >> 2) == (((int)(a1 & 0xffffu) - WID) >> 2));
            if (same) {
                const uint32_t aa[2] = {a0, a1};
                const int64_t hh[2] = {h0, h1};
                acc ^= probe_block<WID, LPH, 2, GS>(A, aa, hh, lih);
            } else {
                if (ok0) {
                    const uint32_t aa[1] = {a0};
                    const int64_t hh[1] = {h0};
                    acc ^= probe_block<WID, LPH, 1, GS>(A, aa, hh, lih);
                }
                if (ok1) {
                    const uint32_t aa[1] = {a1};
                    const int64_t hh[1] = {h1};
                    acc ^= probe_block<WID, LPH, 1, GS>(A, aa, hh, lih);
                }
            }
        }
    }
    if (acc == 0x9e3779b9u && sink) sink[blockIdx.x] = acc;       // keeps the loads live; practically never taken
}


// gather_probe6: the loads of ncc_score_gather6 (same walk, same pairing, same addresses), no arithmetic.
template <int WID, int GS, int MINB>
__global__ void __launch_bounds__(256, MINB)
    gather_probe6(const ScoreArgs A, int64_t N, const uint32_t* __restrict__ anchors, const uint2* __restrict__ entries,
                  uint32_t* __restrict__ sink, int PER) {
    constexpr int K = 2 * WID + 1;
    constexpr int NG = (K + 6) / 4;
    const int CHUNK = 32 * PER;
    const int lane = threadIdx.x & 31;
    const int wib = threadIdx.x >> 5;
    const int grp = lane >> 3;
    const int lih = lane & 7;
    const int64_t gstride = GS ? (int64_t)GS : A.gstride;
    const int Q2 = (A.Vp - 32) >> 1;
    const int lp = lih < Q2 ? lih : Q2 - 1;
    uint32_t acc = 0u;
    for (int64_t i0 = (int64_t)blockIdx.x * CHUNK; i0 < N; i0 += (int64_t)gridDim.x * CHUNK) {
        int64_t pos = i0 + (int64_t)((wib * 4 + grp) * PER);
        const int64_t end = (pos + PER < N) ? pos + PER : N;
        while (__any_sync(FULL, pos < end)) {
            uint32_t a[2] = {MVS_ANCHOR_INVALID, MVS_ANCHOR_INVALID};
            int64_t h[2] = {0, 0};
            if (pos < end) {
                if (entries) {
                    const uint2 e0 = __ldg(entries + pos);
                    h[0] = e0.x;
                    a[0] = e0.y;
                    if (pos + 1 < end) {
                        const uint2 e1 = __ldg(entries + pos + 1);
                        h[1] = e1.x;
                        a[1] = e1.y;
                    }
                } else {
                    h[0] = pos;
                    a[0] = __ldg(anchors + pos);
                    if (pos + 1 < end) {
                        h[1] = pos + 1;
                        a[1] = __ldg(anchors + pos + 1);
                    }
                }
            }
            bool live[2];
            live[0] = a[0] != MVS_ANCHOR_INVALID;
            live[1] = live[0] && a[1] != MVS_ANCHOR_INVALID && ((a[0] >> 16) == (a[1] >> 16)) &&
                      ((((int)(a[0] & 0xffffu) - WID) >> 2) == (((int)(a[1] & 0xffffu) - WID) >> 2));
            pos += live[1] ? 2 : 1;
            if (!__any_sync(FULL, live[0])) continue;
            if (!live[0]) a[0] = ((uint32_t)WID << 16) | (uint32_t)(WID + 1);
            if (!live[1]) {
                a[1] = a[0];
                h[1] = h[0];
            }
            const int row = (int)(a[0] >> 16);
            const int cg = ((int)(a[0] & 0xffffu) - WID) >> 2;
            const uint8_t* base = A.gray4 + (int64_t)(row - WID) * A.rowpitch + (int64_t)cg * gstride;
            bool need_last = false;
#pragma unroll
            for (int b = 0; b < 2; ++b) {
                const int col = (int)(a[b] & 0xffffu);
                const int r = live[0] ? __ldg(A.ref + h[b]) : 0;
                const int64_t mi = ((int64_t)row * A.W + col) * A.Vp;
                acc ^= (uint32_t)__ldg(A.smap + mi + r) ^ __ldg(A.vmap + mi + r);
                need_last |= group_mask((col - WID) & 3, K, NG - 1) != 0u;
                for (int idx = lih; idx < K * NG; idx += 8) {
                    const int rr = idx / NG, g = idx - rr * NG;
                    acc ^= __ldg(reinterpret_cast<const uint32_t*>(base + rr * A.rowpitch + g * gstride + 4 * r));
                }
                const uint2 s4 = __ldg(reinterpret_cast<const uint2*>(A.smap + mi + 4 * lih));
                const uint4 v4 = __ldg(reinterpret_cast<const uint4*>(A.vmap + mi + 4 * lih));
                const uint32_t s2 = __ldg(reinterpret_cast<const uint32_t*>(A.smap + mi + 32 + 2 * lp));
                const uint2 v2 = __ldg(reinterpret_cast<const uint2*>(A.vmap + mi + 32 + 2 * lp));
                acc = fold4(acc ^ s4.x ^ s4.y ^ s2 ^ v2.x ^ v2.y, v4);
            }
            const uint32_t lastoff = need_last ? (uint32_t)((NG - 1) * gstride) : 0u;
            const uint8_t* p4 = base + 16 * lih;
            const uint8_t* p2 = base + 128 + 8 * lp;
#pragma unroll
            for (int rr = 0; rr < K; ++rr) {
#pragma unroll
                for (int g = 0; g < NG; ++g) {
                    const uint32_t off = (g == NG - 1) ? lastoff : (uint32_t)(g * gstride);
                    const uint2 w2 = __ldg(reinterpret_cast<const uint2*>(p2 + off));
                    acc = fold4(acc ^ w2.x ^ w2.y, __ldg(reinterpret_cast<const uint4*>(p4 + off)));
                }
                p4 += A.rowpitch;
                p2 += A.rowpitch;
            }
        }
    }
    if (acc == 0x9e3779b9u && sink) sink[blockIdx.x] = acc;
}

#ifndef MVS_K6_MINB
#define MVS_K6_MINB 2          // resident CTAs per SM: measured 0.364 ms (2: 128 registers, no spills) / 0.382 (3) / 0.470 (4) per 2^20 hypotheses
#endif

// CTAs of a K1 launch: at most 8 per resident slot (the CTAs walk the chunks with a grid stride); a launch that is one of
// ctx->k1_share concurrent range launches of the same batch (overlapped exchange) gets its share of that cap, so that
// the batch as a whole is scored by the same number of CTAs as a single launch.  MVS_K1_CAPMUL: tuning knob.
static int64_t k1_grid(const mvs_ctx* ctx, int64_t want, int minb) {
    static int mul = 0;
    if (mul == 0) {
        const char* e = getenv("MVS_K1_CAPMUL");
        mul = e ? atoi(e) : 8;
        if (mul < 1) mul = 8;
    }
    int64_t cap = (int64_t)ctx->sm_count * minb * mul;
    if (ctx->k1_share > 1) cap = (cap + ctx->k1_share - 1) / ctx->k1_share;
    if (cap < 1) cap = 1;
    return want < cap ? want : cap;
}

template <int WID, int GS, int MINB = MVS_K6_MINB>
static int launch_gather6(mvs_ctx* ctx, const ScoreArgs& A, int64_t N, const uint32_t* anchors, const uint2* entries,
                          cudaStream_t s) {
    static int per = 0;                                    // MVS_K6_PER: positions per lane-group run (tuning knob)
    if (per == 0) {
        const char* e = getenv("MVS_K6_PER");
        per = e ? atoi(e) : MVS_K6_PER;
        if (per < 2 || per > 1024) per = MVS_K6_PER;
    }
    const int64_t chunk = 32 * (int64_t)per;
    const int64_t want = (N + chunk - 1) / chunk;
    const int blocks = (int)k1_grid(ctx, want, MINB);
    if (ctx->probe_gather)
        gather_probe6<WID, GS, MINB><<<blocks, 256, 0, s>>>(A, N, anchors, entries, (uint32_t*)ctx->d_bin_hist, per);
    else if (A.ncc_out)
        ncc_score_gather6<WID, GS, MINB, true><<<blocks, 256, 0, s>>>(A, N, anchors, entries, per);
    else
        ncc_score_gather6<WID, GS, MINB, false><<<blocks, 256, 0, s>>>(A, N, anchors, entries, per);
    return MVS_OK;
}

// MVS_K1_LEGACY=1: the 16-lane mapping for 33..48 views (kept for comparison)
static int k1_legacy() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("MVS_K1_LEGACY");
        v = e ? atoi(e) : 0;
    }
    return v;
}

#ifndef MVS_K1_MINB
#define MVS_K1_MINB 4
#endif

template <int WID, int LPH, int GS, int MINB = MVS_K1_MINB, int UNR = 0>
static int launch_gather_gs(mvs_ctx* ctx, const ScoreArgs& A, int64_t N, const uint32_t* anchors, const uint2* entries,
                            cudaStream_t s) {
    // the per-view NCC dump is a parity/debug output: its stores are compiled out of the hot variant
    auto kern = A.ncc_out ? ncc_score_gather<WID, LPH, GS, MINB, true, UNR> : ncc_score_gather<WID, LPH, GS, MINB, false, UNR>;
    const int64_t chunk = 8 * (32 / LPH) * 2 * MVS_K1_ITERS;
    const int64_t want = (N + chunk - 1) / chunk;
    const int blocks = (int)k1_grid(ctx, want, MINB);
    if (ctx->probe_gather)                                 // measurement hook: loads only, no results (mvs_probe_gather)
        gather_probe<WID, LPH, GS, MINB><<<blocks, 256, 0, s>>>(A, N, anchors, entries, (uint32_t*)ctx->d_bin_hist);
    else
        kern<<<blocks, 256, 0, s>>>(A, N, anchors, entries);
    return MVS_OK;
}

// tuning knob for experiments on the headline configuration (resident CTAs per SM)
static int k1_minb_override() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("MVS_K1_MINB");
        v = e ? atoi(e) : 0;
    }
    return v;
}

template <int WID>
static int launch_gather(mvs_ctx* ctx, const ScoreArgs& A, int64_t N, const uint32_t* anchors, const uint2* entries,
                         cudaStream_t s) {
    const int Q = ctx->Q;
    if (Q <= 4) return launch_gather_gs<WID, 4, 0>(ctx, A, N, anchors, entries, s);
    if (Q <= 8) return launch_gather_gs<WID, 8, 0>(ctx, A, N, anchors, entries, s);
    // 33..48 views: eight lanes per block, six views per lane (needs a window that fits the image for its idle groups)
    if (Q > 8 && Q <= 12 && !k1_legacy() && ctx->H >= 2 * WID + 2 && ctx->W >= 2 * WID + 3) {
        if (WID == 5 && Q == 12) {
            static int mb = -1;                            // MVS_K6_MINB: resident CTAs per SM (tuning knob)
            if (mb < 0) {
                const char* e = getenv("MVS_K6_MINB");
                mb = e ? atoi(e) : 0;
            }
            if (mb == 3) return launch_gather6<5, 192, 3>(ctx, A, N, anchors, entries, s);
            if (mb == 4) return launch_gather6<5, 192, 4>(ctx, A, N, anchors, entries, s);
            return launch_gather6<5, 192>(ctx, A, N, anchors, entries, s);
        }
        return launch_gather6<WID, 0>(ctx, A, N, anchors, entries, s);
    }
    if (Q <= 16) {
        // the reference's own configuration (wid 5, dinoRing's 48 views): group stride as an immediate
        if (WID == 5 && Q == 12) {
            const int mb = k1_minb_override();
            if (mb == 2) return launch_gather_gs<5, 16, 192, 2>(ctx, A, N, anchors, entries, s);
            if (mb == 4) return launch_gather_gs<5, 16, 192, 4>(ctx, A, N, anchors, entries, s);
            if (mb == 3) return launch_gather_gs<5, 16, 192, 3>(ctx, A, N, anchors, entries, s);
            return launch_gather_gs<5, 16, 192, 4>(ctx, A, N, anchors, entries, s);
        }
        return launch_gather_gs<WID, 16, 0>(ctx, A, N, anchors, entries, s);
    }
    if (WID == 5 && (Q == 32 || Q == 64)) {
        const int mb = k1_minb_override();                 // MVS_K1_MINB: resident CTAs per SM (tuning knob)
        // MVS_K1_UNR: window rows per loop body of the 128-view variant (tuning knob).  Measured per 2^20 hypotheses on
        // ring128_1080p: all 11 rows unrolled 1.075 ms, 1 row 1.034, 2 rows 1.012-1.016, 3 rows 1.015, 4 rows 1.048, 6 rows
        // 1.100 -- the unrolled pair + single instances miss the instruction cache; ring256_4k (HBM-resident stack, two
        // passes of 32 quads) wants the loads of all rows in flight instead: 2.00 ms unrolled, 2.15-2.19 with 2, 4 or 6 rows.
        static int unr = -2;
        if (unr == -2) {
            const char* e = getenv("MVS_K1_UNR");
            unr = e ? atoi(e) : -1;
        }
        if (Q == 32) {
            if (mb == 2) return launch_gather_gs<5, 32, 512, 2>(ctx, A, N, anchors, entries, s);
            if (mb == 3) return launch_gather_gs<5, 32, 512, 3>(ctx, A, N, anchors, entries, s);
            if (unr == 0) return launch_gather_gs<5, 32, 512>(ctx, A, N, anchors, entries, s);
            if (unr == 1) return launch_gather_gs<5, 32, 512, MVS_K1_MINB, 1>(ctx, A, N, anchors, entries, s);
            if (unr == 3) return launch_gather_gs<5, 32, 512, MVS_K1_MINB, 3>(ctx, A, N, anchors, entries, s);
            return launch_gather_gs<5, 32, 512, MVS_K1_MINB, 2>(ctx, A, N, anchors, entries, s);
        }
        if (mb == 2) return launch_gather_gs<5, 32, 1024, 2>(ctx, A, N, anchors, entries, s);
        if (mb == 3) return launch_gather_gs<5, 32, 1024, 3>(ctx, A, N, anchors, entries, s);
        if (unr == 2) return launch_gather_gs<5, 32, 1024, MVS_K1_MINB, 2>(ctx, A, N, anchors, entries, s);
        return launch_gather_gs<5, 32, 1024>(ctx, A, N, anchors, entries, s);
    }
    return launch_gather_gs<WID, 32, 0>(ctx, A, N, anchors, entries, s);
}

// K1 alone over positions [p0, p1) of the batch mvs_bin_hypotheses left in ctx->d_bin_* (`sort` as passed there).  No
// profiling bracket: mvs_launch_score_refexact and the overlapped score + publish (exchange.cu) wrap it.  A sub-range
// of an ordered batch is the same kernels on an offset view of the (hypothesis index, anchor) entries.
int mvs_launch_k1(mvs_ctx* ctx, int64_t N, const int32_t* ref, double thr, int wid, uint64_t* vis, double* avg,
                  int32_t* count, float* ncc, bool sort, int64_t p0, int64_t p1, cudaStream_t s) {
    if (p1 > N) p1 = N;
    if (p0 >= p1) return MVS_OK;
    if (!sort && p0 != 0) {                                // input-order batches are indexed by position: no sub-ranges
        mvs_set_error("mvs_launch_k1: a position range needs an ordered batch");
        return MVS_ERR_ARG;
    }
    const uint32_t* anchors = ctx->d_bin_anchor;
    const uint2* entries = sort ? (const uint2*)ctx->d_bin_entry + p0 : nullptr;
    ScoreArgs A;
    A.gray4 = ctx->d_gray; A.smap = ctx->d_smap; A.vmap = ctx->d_vmap;
    A.V = ctx->V; A.Vp = ctx->Vp; A.Q = ctx->Q; A.W = ctx->W;
    A.gstride = ctx->gstride; A.rowpitch = ctx->rowpitch;
    A.ref = ref; A.thr = thr; A.vis_out = vis; A.avg_out = avg; A.count_out = count; A.ncc_out = ncc;
    const int64_t n = p1 - p0;
    int rc;
    switch (wid) {
        case 1: rc = launch_gather<1>(ctx, A, n, anchors, entries, s); break;
        case 2: rc = launch_gather<2>(ctx, A, n, anchors, entries, s); break;
        case 3: rc = launch_gather<3>(ctx, A, n, anchors, entries, s); break;
        case 4: rc = launch_gather<4>(ctx, A, n, anchors, entries, s); break;
        case 5: rc = launch_gather<5>(ctx, A, n, anchors, entries, s); break;
        case 6: rc = launch_gather<6>(ctx, A, n, anchors, entries, s); break;
        default: rc = launch_gather<7>(ctx, A, n, anchors, entries, s); break;
    }
    ctx->launches++;
    MVS_CUDA_CHECK(cudaGetLastError());
    return rc;
}

int mvs_launch_score_refexact(mvs_ctx* ctx, int64_t N, const double* c, const int32_t* ref, double thr, int wid,
                              uint64_t* vis, double* avg, int32_t* count, double* xy, float* ncc, cudaStream_t s) {
    if (N == 0) return MVS_OK;
    if (wid < 1 || wid > 7) {
        mvs_set_error("wid %d not supported (1..7)", wid);
        return MVS_ERR_ARG;
    }
    int rc;
    if ((rc = mvs_build_window_maps(ctx, wid, s)) != MVS_OK) return rc;
    const bool sort = N >= MVS_SORT_MIN;
    if ((rc = mvs_bin_hypotheses(ctx, N, c, ref, wid, sort, vis, avg, count, xy, ncc, s)) != MVS_OK) return rc;
    const int pslot = (int)(ctx->prof_n % MVS_PROF_RING);
    if (ctx->profile) MVS_CUDA_CHECK(cudaEventRecord(ctx->prof_ev[2 * pslot], s));
    rc = mvs_launch_k1(ctx, N, ref, thr, wid, vis, avg, count, ncc, sort, 0, N, s);
    if (ctx->profile) {
        MVS_CUDA_CHECK(cudaEventRecord(ctx->prof_ev[2 * pslot + 1], s));
        ctx->prof_n++;
    }
    return rc;
}
