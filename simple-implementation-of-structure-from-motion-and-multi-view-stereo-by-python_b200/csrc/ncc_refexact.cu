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

// ncc = (num / sqrt(var_i * var_r)) * n/(n-1), oracle/mode_a.py::score operation order.
// Fast path: num * rsqrt(den) * n/(n-1) is within a few ulp of that; only a value closer
// than 1e-9 to the threshold is recomputed with the correctly rounded sqrt and division,
// so the strict '>' decision is always the oracle's.
__device__ __forceinline__ double ncc_ratio(double num, double var_i, double var_r, double cn, double thr) {
    const double den = var_i * var_r;
    double val = (num * rsqrt(den)) * cn;
    if (__builtin_expect(fabs(val - thr) < 1e-9, 0)) val = (num / sqrt(den)) * cn;
    return val;
}

// ---------------------------------------------------------------------------------
// K1: LPH lanes per hypothesis, 32/LPH hypotheses per warp.  A CTA walks chunks of
// CHUNK consecutive positions of the (tile-ordered) batch so that its lane groups share
// L1 lines.  anchors[i] = row<<16|col of position i (MVS_ANCHOR_INVALID: rejected,
// result already written by bin_project); order[i] = hypothesis index (NULL: identity).
// ---------------------------------------------------------------------------------
template <int WID, int LPH>
__global__ void __launch_bounds__(256, 2)
    ncc_score_gather(const uint8_t* __restrict__ gray4, const uint16_t* __restrict__ smap, const uint32_t* __restrict__ vmap,
                     int V, int Vp, int Q, int W, int64_t gstride, int64_t rowpitch, int64_t N,
                     const uint32_t* __restrict__ anchors, const int32_t* __restrict__ order,
                     const int32_t* __restrict__ ref, double thr, uint64_t* __restrict__ vis_out,
                     double* __restrict__ avg_out, int32_t* __restrict__ count_out, float* __restrict__ ncc_out) {
    constexpr int K = 2 * WID + 1;
    constexpr int NG = (K + 6) / 4;            // pixel groups a K-pixel run at offset 0..3 can touch
    constexpr int NPIX = K * K;
    constexpr int HPW = 32 / LPH;
    constexpr int ITERS = 8;
    constexpr int CHUNK = 8 * HPW * ITERS;
    __shared__ __align__(16) uint32_t s_ref[8][HPW][K][NG];

    const int lane = threadIdx.x & 31;
    const int wib = threadIdx.x >> 5;
    const int sub = lane / LPH;
    const int lih = lane % LPH;
    const uint32_t hmask = hyp_mask<LPH>(sub);
    const int mask_words32 = 2 * ((V + 63) >> 6);          // 32-bit chunks per hypothesis in vis_out
    const int passes = (LPH == 32) ? (Q + 31) >> 5 : 1;
    const double cn = (double)NPIX / (double)(NPIX - 1);
    uint32_t(*sref)[NG] = s_ref[wib][sub];

    for (int64_t i0 = (int64_t)blockIdx.x * CHUNK; i0 < N; i0 += (int64_t)gridDim.x * CHUNK) {
        for (int it = 0; it < ITERS; ++it) {
            const int64_t i = i0 + (it * 8 + wib) * HPW + sub;
            if (i >= N) continue;                          // the whole lane group leaves together
            const uint32_t a = __ldg(anchors + i);
            if (a == MVS_ANCHOR_INVALID) continue;
            const int64_t h = order ? (int64_t)__ldg(order + i) : i;
            const int r = __ldg(ref + h);
            const int row = (int)(a >> 16), col = (int)(a & 0xffffu);
            const int o = (col - WID) & 3;                 // offset of the window inside its first pixel group
            const uint8_t* base = gray4 + (int64_t)(row - WID) * rowpitch + (int64_t)((col - WID) >> 2) * gstride;
            const int64_t mi = ((int64_t)row * W + col) * Vp;
            const int Sr = (int)__ldg(smap + mi + r);
            const double var_r = (double)__ldg(vmap + mi + r);
            // byte masks of the window inside each pixel group; empty groups re-read group 0
            // (an L1 hit) against a zero reference word instead of branching
            uint32_t m[NG];
            int64_t goff[NG];
#pragma unroll
            for (int g = 0; g < NG; ++g) {
                m[g] = group_mask(o, K, g);
                goff[g] = m[g] ? g * gstride : 0;
            }
            // ---- reference window: masked words to shared memory
            __syncwarp(hmask);                             // previous hypothesis' readers are done
            for (int idx = lih; idx < K * NG; idx += LPH) {
                const int rr = idx / NG, g = idx - rr * NG;
                const uint32_t mg = group_mask(o, K, g);
                uint32_t w = 0u;
                if (mg) w = __ldg(reinterpret_cast<const uint32_t*>(base + rr * rowpitch + g * gstride + 4 * r)) & mg;
                sref[rr][g] = w;
            }
            __syncwarp(hmask);

            double acc = 0.0;
            int count = 0;
            uint32_t myword = 0;
            for (int p = 0; p < passes; ++p) {
                const int qq = p * LPH + lih;              // this lane's quad: views 4qq..4qq+3
                const bool act = qq < Q;
                const int qc = act ? qq : Q - 1;           // idle lanes shadow the last quad (same L1 lines)
                const uint2 s2 = __ldg(reinterpret_cast<const uint2*>(smap + mi + 4 * qc));
                const uint4 v4 = __ldg(reinterpret_cast<const uint4*>(vmap + mi + 4 * qc));
                int SAB[4] = {0, 0, 0, 0};
                const uint8_t* pq = base + (int64_t)qc * 16;
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
                        const uint4 w4 = __ldg(reinterpret_cast<const uint4*>(pq + rr * rowpitch + goff[g]));
                        SAB[0] = dp4a_u(w4.x, rw[g], SAB[0]);
                        SAB[1] = dp4a_u(w4.y, rw[g], SAB[1]);
                        SAB[2] = dp4a_u(w4.z, rw[g], SAB[2]);
                        SAB[3] = dp4a_u(w4.w, rw[g], SAB[3]);
                    }
                }
                // ---- this lane finishes its own four views
                const int Sv[4] = {(int)(s2.x & 0xffffu), (int)(s2.x >> 16), (int)(s2.y & 0xffffu), (int)(s2.y >> 16)};
                const uint32_t var[4] = {v4.x, v4.y, v4.z, v4.w};
                uint32_t nib = 0;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int v = 4 * qq + k;
                    // exact integers: n*SAB, S*Sr <= 225^2 * 255^2 need 64 bits at wid > 5
                    const double num = (WID <= 5) ? (double)(NPIX * SAB[k] - Sv[k] * Sr)
                                                  : (double)((long long)NPIX * SAB[k] - (long long)Sv[k] * Sr);
                    const bool defined = (var[k] != 0u) && (var_r != 0.0);
                    const double val = ncc_ratio(num, (double)var[k], var_r, cn, thr);
                    const bool scored = act && (v < V) && (v != r) && defined;
                    const bool vis = scored && (val > thr);
                    if (vis) {
                        acc += val;
                        nib |= 1u << k;
                    }
                    if (ncc_out && act && v < V) ncc_out[h * V + v] = scored ? (float)val : nanf("");
                }
                // the lane group covers LPH*4 mask bits per pass = max(LPH/8, 1) 32-bit words
                constexpr int WPP = LPH >= 8 ? LPH / 8 : 1;
#pragma unroll
                for (int w = 0; w < WPP; ++w) {
                    const uint32_t contrib = ((lih >> 3) == w) ? (nib << (4 * (lih & 7))) : 0u;
                    const uint32_t word = __reduce_or_sync(hmask, contrib);
                    count += __popc(word);
                    if (lih == p * WPP + w) myword = word; // word index < 32 <=> V <= 1024
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
}

template <int WID, int LPH>
static int launch_gather_lph(mvs_ctx* ctx, int64_t N, const uint32_t* anchors, const int32_t* order, const int32_t* ref,
                             double thr, uint64_t* vis, double* avg, int32_t* count, float* ncc, cudaStream_t s) {
    auto kern = ncc_score_gather<WID, LPH>;
    static bool configured = false;                        // per instantiation: prefer L1 over shared memory
    if (!configured) {
        cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxL1);
        configured = true;
    }
    const int64_t chunk = 8 * (32 / LPH) * 8;
    const int64_t want = (N + chunk - 1) / chunk;
    const int64_t cap = (int64_t)ctx->sm_count * 2 * 8;
    const int blocks = (int)(want < cap ? want : cap);
    kern<<<blocks, 256, 0, s>>>(ctx->d_gray, ctx->d_smap, ctx->d_vmap, ctx->V, ctx->Vp, ctx->Q, ctx->W, ctx->gstride,
                                ctx->rowpitch, N, anchors, order, ref, thr, vis, avg, count, ncc);
    return MVS_OK;
}

template <int WID>
static int launch_gather(mvs_ctx* ctx, int64_t N, const uint32_t* anchors, const int32_t* order, const int32_t* ref,
                         double thr, uint64_t* vis, double* avg, int32_t* count, float* ncc, cudaStream_t s) {
    const int Q = ctx->Q;
    if (Q <= 4) return launch_gather_lph<WID, 4>(ctx, N, anchors, order, ref, thr, vis, avg, count, ncc, s);
    if (Q <= 8) return launch_gather_lph<WID, 8>(ctx, N, anchors, order, ref, thr, vis, avg, count, ncc, s);
    if (Q <= 16) return launch_gather_lph<WID, 16>(ctx, N, anchors, order, ref, thr, vis, avg, count, ncc, s);
    return launch_gather_lph<WID, 32>(ctx, N, anchors, order, ref, thr, vis, avg, count, ncc, s);
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
    const uint32_t* anchors = sort ? ctx->d_bin_sanchor : ctx->d_bin_anchor;
    const int32_t* order = sort ? ctx->d_bin_order : nullptr;
    const int pslot = (int)(ctx->prof_n % MVS_PROF_RING);
    if (ctx->profile) MVS_CUDA_CHECK(cudaEventRecord(ctx->prof_ev[2 * pslot], s));
    switch (wid) {
        case 1: rc = launch_gather<1>(ctx, N, anchors, order, ref, thr, vis, avg, count, ncc, s); break;
        case 2: rc = launch_gather<2>(ctx, N, anchors, order, ref, thr, vis, avg, count, ncc, s); break;
        case 3: rc = launch_gather<3>(ctx, N, anchors, order, ref, thr, vis, avg, count, ncc, s); break;
        case 4: rc = launch_gather<4>(ctx, N, anchors, order, ref, thr, vis, avg, count, ncc, s); break;
        case 5: rc = launch_gather<5>(ctx, N, anchors, order, ref, thr, vis, avg, count, ncc, s); break;
        case 6: rc = launch_gather<6>(ctx, N, anchors, order, ref, thr, vis, avg, count, ncc, s); break;
        default: rc = launch_gather<7>(ctx, N, anchors, order, ref, thr, vis, avg, count, ncc, s); break;
    }
    if (ctx->profile) {
        MVS_CUDA_CHECK(cudaEventRecord(ctx->prof_ev[2 * pslot + 1], s));
        ctx->prof_n++;
    }
    ctx->launches++;
    MVS_CUDA_CHECK(cudaGetLastError());
    return rc;
}
