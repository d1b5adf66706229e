// Seed-patch stage on the device: the loop of DensePointsWithMVS2 over the SfM tracks
// (MVS2.py:208-260) as four batched steps --
//   seed_triangulate : every (reference observation, other observation) pair of every track through
//                      cv2.triangulatePoints (utils.py:238-239): 4x4 DLT system + OpenCV's one-sided
//                      Jacobi SVD, restated (OpenCV is a third-party dependency outside the reference
//                      tree; the restatement agrees with cv2 4.13 to 4e-16 on dinoRing's seeds,
//                      oracle/triangulate.py); then c = X/w, dist = |c - O|, n = (O - c)/dist
//                      (MVS2.py:241-247)
//   (K1 at MIN_NCC 0.4 on the whole candidate list, MVS2.py:255)
//   seed_select      : per track the candidate with the smallest heap key (dist, c0, c1, c2, R)
//                      (MVS2.py:14,253-260: nearest first, first with enough visible views wins)
//   (order-preserving compaction into patch records + cell fill, MVS2.py:257-259)
// fp64 with explicit round-to-nearest operations (no FMA contraction), like the scalar code it restates.
#include "scan.cuh"

__device__ __forceinline__ double m_(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double a_(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double s_(double a, double b) { return __dsub_rn(a, b); }

// right singular vector of the smallest singular value of the 4x4 matrix whose COLUMNS are the rows of At
__device__ void jacobi_null4(double (&At)[4][4], double (&out)[4]) {
    double Vt[4][4] = {{1, 0, 0, 0}, {0, 1, 0, 0}, {0, 0, 1, 0}, {0, 0, 0, 1}};
    double W[4];
    const double eps = 2.220446049250313e-16 * 10;
    for (int i = 0; i < 4; ++i) {
        double sd = 0.0;
        for (int k = 0; k < 4; ++k) sd = a_(sd, m_(At[i][k], At[i][k]));
        W[i] = sd;
    }
    for (int iter = 0; iter < 30; ++iter) {
        bool changed = false;
        for (int i = 0; i < 3; ++i)
            for (int j = i + 1; j < 4; ++j) {
                double a = W[i], b = W[j], p = 0.0;
                for (int k = 0; k < 4; ++k) p = a_(p, m_(At[i][k], At[j][k]));
                if (fabs(p) <= m_(eps, sqrt(m_(a, b)))) continue;
                p = m_(p, 2.0);
                const double beta = s_(a, b), gamma = hypot(p, beta);
                double c, s;
                if (beta < 0) {
                    const double delta = m_(s_(gamma, beta), 0.5);
                    s = sqrt(__ddiv_rn(delta, gamma));
                    c = __ddiv_rn(p, m_(m_(gamma, s), 2.0));
                } else {
                    c = sqrt(__ddiv_rn(a_(gamma, beta), m_(gamma, 2.0)));
                    s = __ddiv_rn(p, m_(m_(gamma, c), 2.0));
                }
                a = b = 0.0;
                for (int k = 0; k < 4; ++k) {
                    const double t0 = a_(m_(c, At[i][k]), m_(s, At[j][k]));
                    const double t1 = a_(m_(-s, At[i][k]), m_(c, At[j][k]));
                    At[i][k] = t0;
                    At[j][k] = t1;
                    a = a_(a, m_(t0, t0));
                    b = a_(b, m_(t1, t1));
                }
                W[i] = a;
                W[j] = b;
                changed = true;
                for (int k = 0; k < 4; ++k) {
                    const double t0 = a_(m_(c, Vt[i][k]), m_(s, Vt[j][k]));
                    const double t1 = a_(m_(-s, Vt[i][k]), m_(c, Vt[j][k]));
                    Vt[i][k] = t0;
                    Vt[j][k] = t1;
                }
            }
        if (!changed) break;
    }
    for (int i = 0; i < 4; ++i) {
        double sd = 0.0;
        for (int k = 0; k < 4; ++k) sd = a_(sd, m_(At[i][k], At[i][k]));
        W[i] = sqrt(sd);
    }
    int order[4] = {0, 1, 2, 3};
    for (int i = 0; i < 3; ++i) {                          // OpenCV's selection sort, decreasing
        int j = i;
        for (int k = i + 1; k < 4; ++k)
            if (W[order[j]] < W[order[k]]) j = k;
        const int t = order[i];
        order[i] = order[j];
        order[j] = t;
    }
    for (int k = 0; k < 4; ++k) out[k] = Vt[order[3]][k];
}

// X4 [n,4] homogeneous; optionally the seed candidate derived from it: c [n,3], nrm [n,3], dist [n] w.r.t.
// the centre of view va (MVS2.py:241-247)
__global__ void __launch_bounds__(64)
    seed_triangulate(int64_t n, const int32_t* __restrict__ va, const int32_t* __restrict__ vb, const double* __restrict__ xa,
                     const double* __restrict__ xb, const double* __restrict__ P, int V, const CamGeom* __restrict__ geom,
                     double* __restrict__ X4, double* __restrict__ c_out, double* __restrict__ n_out, double* __restrict__ dist_out) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int a = va[i], b = vb[i];
    double un[4] = {nan(""), nan(""), nan(""), nan("")};
    if (a >= 0 && a < V && b >= 0 && b < V) {
        const double* P1 = P + 12 * a;
        const double* P2 = P + 12 * b;
        const double x1 = xa[2 * i], y1 = xa[2 * i + 1], x2 = xb[2 * i], y2 = xb[2 * i + 1];
        double At[4][4];                                    // At[k][row] = A[row][k]
        for (int k = 0; k < 4; ++k) {
            At[k][0] = s_(m_(x1, P1[8 + k]), P1[k]);
            At[k][1] = s_(m_(y1, P1[8 + k]), P1[4 + k]);
            At[k][2] = s_(m_(x2, P2[8 + k]), P2[k]);
            At[k][3] = s_(m_(y2, P2[8 + k]), P2[4 + k]);
        }
        jacobi_null4(At, un);
    }
    if (X4)
        for (int k = 0; k < 4; ++k) X4[4 * i + k] = un[k];
    if (c_out) {
        double c[3];
        for (int k = 0; k < 3; ++k) c[k] = (un[3] == 0.0) ? m_(0.0, un[k]) : __ddiv_rn(un[k], un[3]);   // MVS2.py:241-244
        double d[3] = {0, 0, 0}, dist = nan("");
        if (a >= 0 && a < V) {
            const CamGeom& g = geom[a];
            // distance(c, O) (utils.py:246-247), n = (O - c)/dist (MVS2.py:245-246)
            const double e0 = s_(c[0], g.C[0]), e1 = s_(c[1], g.C[1]), e2 = s_(c[2], g.C[2]);
            dist = sqrt(a_(a_(m_(e0, e0), m_(e1, e1)), m_(e2, e2)));
            d[0] = __ddiv_rn(s_(g.C[0], c[0]), dist);
            d[1] = __ddiv_rn(s_(g.C[1], c[1]), dist);
            d[2] = __ddiv_rn(s_(g.C[2], c[2]), dist);
        }
        for (int k = 0; k < 3; ++k) {
            c_out[3 * i + k] = c[k];
            n_out[3 * i + k] = d[k];
        }
        dist_out[i] = dist;
    }
}

// one thread per track: smallest (dist, c0, c1, c2, R) among the candidates with count >= bound
__global__ void __launch_bounds__(128)
    seed_select(int64_t n_tracks, const int64_t* __restrict__ cand_off, const double* __restrict__ dist,
                const double* __restrict__ c, const int32_t* __restrict__ ref, const int32_t* __restrict__ count, int bound,
                uint8_t* __restrict__ gate) {
    const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= n_tracks) return;
    int64_t best = -1;
    for (int64_t i = cand_off[t]; i < cand_off[t + 1]; ++i) {
        if (count[i] < bound) continue;
        if (best < 0) {
            best = i;
            continue;
        }
        // Python tuple comparison: first differing component decides (NaN compares false either way)
        const double ka[4] = {dist[i], c[3 * i], c[3 * i + 1], c[3 * i + 2]};
        const double kb[4] = {dist[best], c[3 * best], c[3 * best + 1], c[3 * best + 2]};
        bool less = false, decided = false;
        for (int q = 0; q < 4 && !decided; ++q) {
            if (ka[q] != kb[q]) {
                less = ka[q] < kb[q];
                decided = true;
            }
        }
        if (!decided) less = ref[i] < ref[best];
        if (less) best = i;
    }
    if (best >= 0) gate[best] = 1;
}

static int upload(void** d, const void* h, size_t bytes, cudaStream_t s) {
    if (cudaMalloc(d, bytes ? bytes : 16) != cudaSuccess) {
        cudaGetLastError();
        mvs_set_error("seed stage: device allocation of %zu bytes failed", bytes);
        return MVS_ERR_NOMEM;
    }
    if (bytes) MVS_CUDA_CHECK(cudaMemcpyAsync(*d, h, bytes, cudaMemcpyHostToDevice, s));
    return MVS_OK;
}

extern "C" int mvs_triangulate(mvs_ctx* ctx, int64_t n, const int32_t* view_a, const int32_t* view_b, const double* xa,
                               const double* xb, const double* P, double* X4) {
    if (!ctx) { mvs_set_error("mvs_triangulate: null context"); return MVS_ERR_ARG; }
    if (n < 0 || !P || (n > 0 && (!view_a || !view_b || !xa || !xb || !X4))) { mvs_set_error("mvs_triangulate: null argument"); return MVS_ERR_ARG; }
    if (n == 0) return MVS_OK;
    MVS_CUDA_CHECK(cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->own_stream;
    void *d_va = nullptr, *d_vb = nullptr, *d_xa = nullptr, *d_xb = nullptr, *d_P = nullptr, *d_X = nullptr;
    int rc = MVS_OK;
    if ((rc = upload(&d_va, view_a, sizeof(int32_t) * n, s)) == MVS_OK && (rc = upload(&d_vb, view_b, sizeof(int32_t) * n, s)) == MVS_OK &&
        (rc = upload(&d_xa, xa, sizeof(double) * 2 * n, s)) == MVS_OK && (rc = upload(&d_xb, xb, sizeof(double) * 2 * n, s)) == MVS_OK &&
        (rc = upload(&d_P, P, sizeof(double) * 12 * ctx->V, s)) == MVS_OK) {
        if (cudaMalloc(&d_X, sizeof(double) * 4 * n) != cudaSuccess) {
            cudaGetLastError();
            rc = MVS_ERR_NOMEM;
        } else {
            seed_triangulate<<<(unsigned)((n + 63) / 64), 64, 0, s>>>(n, (const int32_t*)d_va, (const int32_t*)d_vb, (const double*)d_xa,
                                                                     (const double*)d_xb, (const double*)d_P, ctx->V, ctx->d_geom,
                                                                     (double*)d_X, nullptr, nullptr, nullptr);
            ctx->launches++;
            if (cudaMemcpyAsync(X4, d_X, sizeof(double) * 4 * n, cudaMemcpyDeviceToHost, s) != cudaSuccess ||
                cudaStreamSynchronize(s) != cudaSuccess) {
                mvs_set_error("mvs_triangulate failed: %s", cudaGetErrorString(cudaGetLastError()));
                rc = MVS_ERR_CUDA;
            }
        }
    }
    void* bufs[] = {d_va, d_vb, d_xa, d_xb, d_P, d_X};
    for (void* b : bufs)
        if (b) cudaFree(b);
    return rc;
}

extern "C" int mvs_seed_stage(mvs_ctx* ctx, int64_t n_tracks, const int64_t* offsets, const double* obs, const double* P,
                              double min_ncc, int wid, int bound, void* seeds, int64_t* n_seeds, void* stream) {
    if (!ctx) { mvs_set_error("mvs_seed_stage: null context"); return MVS_ERR_ARG; }
    if (n_tracks < 0 || !n_seeds || !P || (n_tracks > 0 && (!offsets || !obs || !seeds))) {
        mvs_set_error("mvs_seed_stage: null argument");
        return MVS_ERR_ARG;
    }
    *n_seeds = 0;
    if (n_tracks == 0) return MVS_OK;
    MVS_CUDA_CHECK(cudaSetDevice(ctx->device));
    cudaStream_t s = (cudaStream_t)stream;
    // ---- host: candidate list = (first observation, k-th observation) of every track, k >= 1
    int64_t n = 0;
    int64_t* cand_off = (int64_t*)malloc(sizeof(int64_t) * (n_tracks + 1));
    if (!cand_off) return MVS_ERR_NOMEM;
    for (int64_t t = 0; t < n_tracks; ++t) {
        cand_off[t] = n;
        const int64_t len = offsets[t + 1] - offsets[t];
        if (len < 0) { free(cand_off); mvs_set_error("mvs_seed_stage: offsets must be non-decreasing"); return MVS_ERR_ARG; }
        if (len > 1) n += len - 1;
    }
    cand_off[n_tracks] = n;
    if (n == 0) { free(cand_off); return MVS_OK; }
    int32_t* h_v = (int32_t*)malloc(sizeof(int32_t) * 2 * n);
    double* h_x = (double*)malloc(sizeof(double) * 4 * n);
    if (!h_v || !h_x) { free(cand_off); free(h_v); free(h_x); return MVS_ERR_NOMEM; }
    for (int64_t t = 0, i = 0; t < n_tracks; ++t) {
        const int64_t lo = offsets[t], hi = offsets[t + 1];
        for (int64_t k = lo + 1; k < hi; ++k, ++i) {
            h_v[i] = (int32_t)obs[3 * lo];
            h_v[n + i] = (int32_t)obs[3 * k];
            h_x[2 * i] = obs[3 * lo + 1];
            h_x[2 * i + 1] = obs[3 * lo + 2];
            h_x[2 * n + 2 * i] = obs[3 * k + 1];
            h_x[2 * n + 2 * i + 1] = obs[3 * k + 2];
        }
    }
    const size_t mw = (size_t)((ctx->V + 63) / 64);
    void *d_v = nullptr, *d_x = nullptr, *d_P = nullptr, *d_off = nullptr;
    uint8_t* d_work = nullptr;
    int rc;
    if ((rc = upload(&d_v, h_v, sizeof(int32_t) * 2 * n, s)) != MVS_OK || (rc = upload(&d_x, h_x, sizeof(double) * 4 * n, s)) != MVS_OK ||
        (rc = upload(&d_P, P, sizeof(double) * 12 * ctx->V, s)) != MVS_OK ||
        (rc = upload(&d_off, cand_off, sizeof(int64_t) * (n_tracks + 1), s)) != MVS_OK)
        goto done;
    {
        // c, nrm [n,3] | dist, avg [n] | xy [n,2] | vis [n,mw] | count [n] i32 | gate [n] u8 | n_out i64
        const size_t bytes = sizeof(double) * (3 + 3 + 1 + 1 + 2 + mw) * n + sizeof(int32_t) * n + n + 64;
        if (cudaMalloc(&d_work, bytes) != cudaSuccess) {
            cudaGetLastError();
            mvs_set_error("mvs_seed_stage: device allocation of %zu bytes failed", bytes);
            rc = MVS_ERR_NOMEM;
            goto done;
        }
        double* d_c = (double*)d_work;
        double* d_n = d_c + 3 * n;
        double* d_dist = d_n + 3 * n;
        double* d_avg = d_dist + n;
        double* d_xy = d_avg + n;
        uint64_t* d_vis = (uint64_t*)(d_xy + 2 * n);
        int64_t* d_nout = (int64_t*)(d_vis + mw * n);
        int32_t* d_cnt = (int32_t*)(d_nout + 1);
        uint8_t* d_gate = (uint8_t*)(d_cnt + n);
        const int32_t* d_ref = (const int32_t*)d_v;
        seed_triangulate<<<(unsigned)((n + 63) / 64), 64, 0, s>>>(n, d_ref, d_ref + n, (const double*)d_x, (const double*)d_x + 2 * n,
                                                                 (const double*)d_P, ctx->V, ctx->d_geom, nullptr, d_c, d_n, d_dist);
        ctx->launches++;
        if ((rc = mvs_launch_score_refexact(ctx, n, d_c, d_ref, min_ncc, wid, d_vis, d_avg, d_cnt, d_xy, nullptr, s)) != MVS_OK) goto done;
        if (cudaMemsetAsync(d_gate, 0, n, s) != cudaSuccess) { rc = MVS_ERR_CUDA; goto done; }
        seed_select<<<(unsigned)((n_tracks + 127) / 128), 128, 0, s>>>(n_tracks, (const int64_t*)d_off, d_dist, d_c, d_ref, d_cnt, bound,
                                                                      d_gate);
        ctx->launches++;
        if ((rc = mvs_launch_compact(ctx, n, 0, d_c, d_n, d_ref, d_vis, d_avg, d_cnt, d_xy, d_gate, bound, seeds, n_tracks, d_nout,
                                     nullptr, nullptr, s)) != MVS_OK)
            goto done;
        if (cudaMemcpyAsync(n_seeds, d_nout, sizeof(int64_t), cudaMemcpyDeviceToHost, s) != cudaSuccess ||
            cudaStreamSynchronize(s) != cudaSuccess) {
            mvs_set_error("mvs_seed_stage failed: %s", cudaGetErrorString(cudaGetLastError()));
            rc = MVS_ERR_CUDA;
            goto done;
        }
        if (ctx->d_cells && *n_seeds > 0) rc = mvs_cells_fill(ctx, seeds, *n_seeds, stream);   // MVS2.py:258-259
        if (rc == MVS_OK && cudaStreamSynchronize(s) != cudaSuccess) rc = MVS_ERR_CUDA;
    }
done:
    free(cand_off);
    free(h_v);
    free(h_x);
    void* bufs[] = {d_v, d_x, d_P, d_off, d_work};
    for (void* b : bufs)
        if (b) cudaFree(b);
    return rc;
}
