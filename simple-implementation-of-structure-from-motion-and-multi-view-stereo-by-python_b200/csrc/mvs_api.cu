// C-ABI entry points (include/mvs_ncc.h): context lifetime, camera preparation,
// host/device buffer handling.  Kernels live in the other .cu files.
#include <math.h>
#include <stdarg.h>
#include <stdlib.h>

#include "mvs_common.cuh"

static thread_local char g_err[512] = "";

void mvs_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

extern "C" const char* mvs_last_error(void) { return g_err; }
extern "C" int mvs_abi_version(void) { return MVS_ABI_VERSION; }

// ---------------------------------------------------------------------------------
// Rodrigues round trip R' = Rodrigues(Rodrigues(R)) (utils.py:242-243), restated from
// OpenCV's published algorithm: orthonormalise (OpenCV: U*Vt of the SVD; here the
// same polar factor by Newton iteration X <- (X + X^-T)/2), take the axis-angle
// vector, rebuild the matrix.  Host code, fp64, runs once per view at load time.
// ---------------------------------------------------------------------------------
static void inv_transpose3(const double* a, double* o) {
    const double c00 = a[4] * a[8] - a[5] * a[7], c01 = a[5] * a[6] - a[3] * a[8], c02 = a[3] * a[7] - a[4] * a[6];
    const double c10 = a[2] * a[7] - a[1] * a[8], c11 = a[0] * a[8] - a[2] * a[6], c12 = a[1] * a[6] - a[0] * a[7];
    const double c20 = a[1] * a[5] - a[2] * a[4], c21 = a[2] * a[3] - a[0] * a[5], c22 = a[0] * a[4] - a[1] * a[3];
    const double det = a[0] * c00 + a[1] * c01 + a[2] * c02;
    const double id = 1.0 / det;
    // inverse = adj/det with adj = cof^T, so inverse-transpose = cof/det
    o[0] = c00 * id; o[1] = c01 * id; o[2] = c02 * id;
    o[3] = c10 * id; o[4] = c11 * id; o[5] = c12 * id;
    o[6] = c20 * id; o[7] = c21 * id; o[8] = c22 * id;
}

static void rodrigues_roundtrip(const double* Rin, double* out) {
    double Q[9], T[9];
    memcpy(Q, Rin, sizeof(Q));
    for (int it = 0; it < 12; ++it) {
        inv_transpose3(Q, T);
        double delta = 0.0;
        for (int i = 0; i < 9; ++i) {
            const double nv = 0.5 * (Q[i] + T[i]);
            delta = fmax(delta, fabs(nv - Q[i]));
            Q[i] = nv;
        }
        if (delta < 1e-17) break;
    }
    double rx = Q[7] - Q[5], ry = Q[2] - Q[6], rz = Q[3] - Q[1];
    const double s = sqrt((rx * rx + ry * ry + rz * rz) * 0.25);
    double c = (Q[0] + Q[4] + Q[8] - 1.0) * 0.5;
    c = c > 1.0 ? 1.0 : (c < -1.0 ? -1.0 : c);
    const double theta = acos(c);
    double v[3];
    if (s < 1e-5) {
        if (c > 0) {
            v[0] = v[1] = v[2] = 0.0;
        } else {
            double tx = sqrt(fmax((Q[0] + 1) * 0.5, 0.0));
            double ty = sqrt(fmax((Q[4] + 1) * 0.5, 0.0)) * (Q[1] < 0 ? -1.0 : 1.0);
            double tz = sqrt(fmax((Q[8] + 1) * 0.5, 0.0)) * (Q[2] < 0 ? -1.0 : 1.0);
            if (fabs(tx) < fabs(ty) && fabs(tx) < fabs(tz) && ((Q[5] > 0) != (ty * tz > 0))) tz = -tz;
            const double k = theta / sqrt(tx * tx + ty * ty + tz * tz);
            v[0] = tx * k; v[1] = ty * k; v[2] = tz * k;
        }
    } else {
        const double k = theta / (2.0 * s);
        v[0] = rx * k; v[1] = ry * k; v[2] = rz * k;
    }
    const double th = sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
    if (th < 2.220446049250313e-16) {
        const double I[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
        memcpy(out, I, sizeof(I));
        return;
    }
    const double kx = v[0] / th, ky = v[1] / th, kz = v[2] / th;
    const double ct = cos(th), st = sin(th), c1 = 1.0 - ct;
    out[0] = ct + c1 * kx * kx;      out[1] = c1 * kx * ky - st * kz; out[2] = c1 * kx * kz + st * ky;
    out[3] = c1 * kx * ky + st * kz; out[4] = ct + c1 * ky * ky;      out[5] = c1 * ky * kz - st * kx;
    out[6] = c1 * kx * kz - st * ky; out[7] = c1 * ky * kz + st * kx; out[8] = ct + c1 * kz * kz;
}

static int ensure_stage(mvs_ctx* ctx, size_t bytes) {
    if (bytes <= ctx->stage_bytes) return MVS_OK;
    if (ctx->d_stage) cudaFree(ctx->d_stage);
    ctx->d_stage = nullptr;
    ctx->stage_bytes = 0;
    size_t want = bytes + bytes / 4 + 4096;
    if (cudaMalloc(&ctx->d_stage, want) != cudaSuccess) {
        cudaGetLastError();
        mvs_set_error("cudaMalloc of %zu staging bytes failed", want);
        return MVS_ERR_NOMEM;
    }
    ctx->stage_bytes = want;
    return MVS_OK;
}

extern "C" int mvs_create(mvs_ctx** out, int device, int V, int H, int W, const uint8_t* rgb, int rgb_on_device,
                          const double* K, const double* R, const double* Rrt, const double* t) {
    if (!out || !rgb || !K || !R || !t || V < 1 || V > 1024 || H < 1 || W < 1 || H > 65535 || W > 65535) {
        mvs_set_error("mvs_create: bad argument (need 1 <= V <= 1024, 1 <= H, W <= 65535, non-null rgb/K/R/t)");
        return MVS_ERR_ARG;
    }
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        mvs_set_error("mvs_create: no CUDA device is usable; this library has no CPU fallback");
        return MVS_ERR_CUDA;
    }
    if (device < 0 || device >= ndev) {
        mvs_set_error("mvs_create: device %d out of range (0..%d)", device, ndev - 1);
        return MVS_ERR_ARG;
    }
    MVS_CUDA_CHECK(cudaSetDevice(device));
    cudaDeviceProp prop;
    MVS_CUDA_CHECK(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) {
        mvs_set_error("mvs_create: device %d is sm_%d%d; this library is built for sm_100a (B200) only", device, prop.major,
                      prop.minor);
        return MVS_ERR_CUDA;
    }
    mvs_ctx* ctx = (mvs_ctx*)calloc(1, sizeof(mvs_ctx));
    if (!ctx) return MVS_ERR_NOMEM;
    ctx->device = device;
    ctx->V = V;
    ctx->H = H;
    ctx->W = W;
    ctx->Vp = (V + 3) / 4 * 4;
    ctx->Q = ctx->Vp / 4;
    ctx->G = (W + 3) / 4 + MVS_GROUP_PAD;
    ctx->gstride = 4 * (int64_t)ctx->Vp;
    ctx->rowpitch = ctx->G * ctx->gstride;
    ctx->sm_count = prop.multiProcessorCount;
    ctx->maps_wid = -1;
    int rc = MVS_OK;
    const size_t gray_bytes = (size_t)ctx->rowpitch * H + 256;
    uint8_t* d_stage2[2] = {nullptr, nullptr};
    CamProj* hp = (CamProj*)malloc(sizeof(CamProj) * V);
    CamGeom* hg = (CamGeom*)malloc(sizeof(CamGeom) * V);
    ctx->h_rrt = (double*)malloc(sizeof(double) * 9 * V);
    ctx->h_centres = (double*)malloc(sizeof(double) * 3 * V);
    if (!hp || !hg || !ctx->h_rrt || !ctx->h_centres) {
        rc = MVS_ERR_NOMEM;
        goto fail;
    }
    for (int v = 0; v < V; ++v) {
        const double* Rf = R + 9 * v;
        double* rr = ctx->h_rrt + 9 * v;
        if (Rrt)
            memcpy(rr, Rrt + 9 * v, sizeof(double) * 9);
        else
            rodrigues_roundtrip(Rf, rr);
        memcpy(hp[v].r, rr, sizeof(double) * 9);
        memcpy(hg[v].rf, Rf, sizeof(double) * 9);
        for (int i = 0; i < 3; ++i) hp[v].t[i] = t[3 * v + i];
        hp[v].fx = hg[v].fx = K[9 * v + 0];
        hp[v].fy = hg[v].fy = K[9 * v + 4];
        hp[v].cx = hg[v].cx = K[9 * v + 2];
        hp[v].cy = hg[v].cy = K[9 * v + 5];
        // C = -(R^T t) with the FILE rotation (MVS2.py:188-189), left-to-right sums
        for (int i = 0; i < 3; ++i) {
            const double s = Rf[0 * 3 + i] * t[3 * v + 0] + Rf[1 * 3 + i] * t[3 * v + 1] + Rf[2 * 3 + i] * t[3 * v + 2];
            hg[v].C[i] = ctx->h_centres[3 * v + i] = -s;
        }
    }
    if (cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaMalloc(&ctx->d_gray, gray_bytes) != cudaSuccess || cudaMalloc(&ctx->d_cam, sizeof(CamProj) * V) != cudaSuccess ||
        cudaMalloc(&ctx->d_geom, sizeof(CamGeom) * V) != cudaSuccess) {
        mvs_set_error("mvs_create: device allocation failed: %s", cudaGetErrorString(cudaGetLastError()));
        rc = MVS_ERR_NOMEM;
        goto fail;
    }
    if (cudaMemsetAsync(ctx->d_gray, 0, gray_bytes, ctx->own_stream) != cudaSuccess ||
        cudaMemcpyAsync(ctx->d_cam, hp, sizeof(CamProj) * V, cudaMemcpyHostToDevice, ctx->own_stream) != cudaSuccess ||
        cudaMemcpyAsync(ctx->d_geom, hg, sizeof(CamGeom) * V, cudaMemcpyHostToDevice, ctx->own_stream) != cudaSuccess) {
        mvs_set_error("mvs_create: camera upload failed: %s", cudaGetErrorString(cudaGetLastError()));
        rc = MVS_ERR_CUDA;
        goto fail;
    }
    if (rgb_on_device) {
        // the caller's stream ordering is unknown: make its writes visible first
        if (cudaDeviceSynchronize() != cudaSuccess) { rc = MVS_ERR_CUDA; goto fail; }
        rc = mvs_launch_gray(ctx, rgb, 0, V, ctx->own_stream);
    } else {
        // Load path (main.py:7-20 hands over pageable NumPy images): chunks of whole views go through two
        // PINNED host buffers and two device staging buffers -- the CPU fills pinned buffer k+1 while the DMA
        // engine uploads chunk k on a copy stream and prep_gray4 converts chunk k-1 on the compute stream.
        // No full-size RGB copy ever exists on the device (6.4 GB for 256 x 4K).
        const size_t view_bytes = (size_t)H * W * 3;
        size_t target = 32u << 20;                            // ~32 MB per chunk
        if (const char* e = getenv("MVS_UPLOAD_CHUNK_MB")) {
            const long mb = atol(e);
            if (mb > 0) target = (size_t)mb << 20;
        }
        int per = (int)(target / view_bytes);
        per = per < 1 ? 1 : (per > V ? V : per);
        const size_t chunk_bytes = view_bytes * per;
        uint8_t* h_pin[2] = {nullptr, nullptr};
        cudaStream_t copy_s = nullptr;
        cudaEvent_t up[2] = {nullptr, nullptr}, done[2] = {nullptr, nullptr};
        bool ok = cudaStreamCreateWithFlags(&copy_s, cudaStreamNonBlocking) == cudaSuccess;
        for (int i = 0; i < 2 && ok; ++i)
            ok = cudaMallocHost(&h_pin[i], chunk_bytes) == cudaSuccess && cudaMalloc(&d_stage2[i], chunk_bytes) == cudaSuccess &&
                 cudaEventCreateWithFlags(&up[i], cudaEventDisableTiming) == cudaSuccess &&
                 cudaEventCreateWithFlags(&done[i], cudaEventDisableTiming) == cudaSuccess;
        if (!ok) {
            mvs_set_error("mvs_create: staging allocation (2 x %zu bytes pinned + device) failed: %s", chunk_bytes,
                          cudaGetErrorString(cudaGetLastError()));
            rc = MVS_ERR_NOMEM;
        }
        for (int v0 = 0, k = 0; rc == MVS_OK && v0 < V; v0 += per, ++k) {
            const int nv = V - v0 < per ? V - v0 : per;
            const int sl = k & 1;
            if (k >= 2 && cudaEventSynchronize(done[sl]) != cudaSuccess) { rc = MVS_ERR_CUDA; break; }   // slot free again
            memcpy(h_pin[sl], rgb + (size_t)v0 * view_bytes, view_bytes * nv);
            if (cudaMemcpyAsync(d_stage2[sl], h_pin[sl], view_bytes * nv, cudaMemcpyHostToDevice, copy_s) != cudaSuccess ||
                cudaEventRecord(up[sl], copy_s) != cudaSuccess || cudaStreamWaitEvent(ctx->own_stream, up[sl], 0) != cudaSuccess) {
                mvs_set_error("mvs_create: RGB upload failed: %s", cudaGetErrorString(cudaGetLastError()));
                rc = MVS_ERR_CUDA;
                break;
            }
            if ((rc = mvs_launch_gray(ctx, d_stage2[sl], v0, nv, ctx->own_stream)) != MVS_OK) break;
            if (cudaEventRecord(done[sl], ctx->own_stream) != cudaSuccess) { rc = MVS_ERR_CUDA; break; }
        }
        if (cudaStreamSynchronize(ctx->own_stream) != cudaSuccess && rc == MVS_OK) rc = MVS_ERR_CUDA;
        for (int i = 0; i < 2; ++i) {
            if (h_pin[i]) cudaFreeHost(h_pin[i]);
            if (up[i]) cudaEventDestroy(up[i]);
            if (done[i]) cudaEventDestroy(done[i]);
        }
        if (copy_s) cudaStreamDestroy(copy_s);
    }
    if (rc != MVS_OK) goto fail;
    if (cudaStreamSynchronize(ctx->own_stream) != cudaSuccess) {
        mvs_set_error("mvs_create: gray conversion failed: %s", cudaGetErrorString(cudaGetLastError()));
        rc = MVS_ERR_CUDA;
        goto fail;
    }
    for (int i = 0; i < 2; ++i)
        if (d_stage2[i]) cudaFree(d_stage2[i]);
    free(hp);
    free(hg);
    *out = ctx;
    return MVS_OK;
fail:
    for (int i = 0; i < 2; ++i)
        if (d_stage2[i]) cudaFree(d_stage2[i]);
    free(hp);
    free(hg);
    mvs_destroy(ctx);
    return rc;
}

extern "C" int mvs_destroy(mvs_ctx* ctx) {
    if (!ctx) return MVS_OK;
    cudaSetDevice(ctx->device);
    if (ctx->d_gray) cudaFree(ctx->d_gray);
    if (ctx->d_cam) cudaFree(ctx->d_cam);
    if (ctx->d_geom) cudaFree(ctx->d_geom);
    if (ctx->d_stage) cudaFree(ctx->d_stage);
    if (ctx->d_tiles) cudaFree(ctx->d_tiles);
    mvs_pmvs_release(ctx);
    void* more[] = {ctx->d_smap, ctx->d_vmap, ctx->d_bin_hist, ctx->d_bin_key, ctx->d_bin_rank, ctx->d_bin_entry,
                    ctx->d_bin_anchor, ctx->d_bin_scan};
    for (void* b : more)
        if (b) cudaFree(b);
    void* bufs[] = {ctx->d_cells, ctx->d_claim, ctx->d_counts, ctx->d_scan, ctx->cand_slot, ctx->cand_parent, ctx->cand_c,
                    ctx->cand_n, ctx->cand_ref, ctx->cand_px, ctx->cand_vis, ctx->cand_avg, ctx->cand_count, ctx->cand_xy,
                    ctx->cand_gate};
    for (void* b : bufs)
        if (b) cudaFree(b);
    void* more2[] = {ctx->d_out, ctx->d_inbox, ctx->d_round_n, ctx->d_barrier_state, ctx->d_ticket, ctx->d_live, ctx->d_bin_part};
    for (void* b : more2)
        if (b) cudaFree(b);
    for (int i = 0; i < 8; ++i) {
        if (ctx->x_side[i]) cudaStreamDestroy(ctx->x_side[i]);
        if (ctx->x_ev[i]) cudaEventDestroy(ctx->x_ev[i]);
    }
    if (ctx->x_fork) cudaEventDestroy(ctx->x_fork);
    if (ctx->h_pinned) cudaFreeHost(ctx->h_pinned);
    for (int i = 0; i < 2; ++i)
        if (ctx->ev_round[i]) cudaEventDestroy(ctx->ev_round[i]);
    for (int i = 0; i < 2 * MVS_PROF_RING; ++i)
        if (ctx->prof_ev[i]) cudaEventDestroy(ctx->prof_ev[i]);
    if (ctx->in_stream) cudaStreamDestroy(ctx->in_stream);
    if (ctx->out_stream) cudaStreamDestroy(ctx->out_stream);
    for (int i = 0; i < 3; ++i) {
        if (ctx->ev_in[i]) cudaEventDestroy(ctx->ev_in[i]);
        if (ctx->ev_done[i]) cudaEventDestroy(ctx->ev_done[i]);
        if (ctx->ev_out[i]) cudaEventDestroy(ctx->ev_out[i]);
    }
    if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
    free(ctx->h_rrt);
    free(ctx->h_centres);
    free(ctx);
    return MVS_OK;
}

extern "C" int mvs_get_info(const mvs_ctx* ctx, int* V, int* H, int* W, int64_t* pitch) {
    if (!ctx) { mvs_set_error("null context"); return MVS_ERR_ARG; }
    if (V) *V = ctx->V;
    if (H) *H = ctx->H;
    if (W) *W = ctx->W;
    if (pitch) *pitch = ctx->rowpitch;
    return MVS_OK;
}

extern "C" int mvs_download_gray(mvs_ctx* ctx, uint8_t* out_host) {
    if (!ctx || !out_host) { mvs_set_error("mvs_download_gray: null argument"); return MVS_ERR_ARG; }
    MVS_CUDA_CHECK(cudaSetDevice(ctx->device));
    const size_t bytes = (size_t)ctx->V * ctx->H * ctx->W;
    int rc = ensure_stage(ctx, bytes);
    if (rc != MVS_OK) return rc;
    rc = mvs_launch_unpack_gray(ctx, (uint8_t*)ctx->d_stage, ctx->own_stream);
    if (rc != MVS_OK) return rc;
    MVS_CUDA_CHECK(cudaMemcpyAsync(out_host, ctx->d_stage, bytes, cudaMemcpyDeviceToHost, ctx->own_stream));
    MVS_CUDA_CHECK(cudaStreamSynchronize(ctx->own_stream));
    return MVS_OK;
}

extern "C" int mvs_get_cameras(const mvs_ctx* ctx, double* Rrt_host, double* centres_host) {
    if (!ctx) { mvs_set_error("null context"); return MVS_ERR_ARG; }
    if (Rrt_host) memcpy(Rrt_host, ctx->h_rrt, sizeof(double) * 9 * ctx->V);
    if (centres_host) memcpy(centres_host, ctx->h_centres, sizeof(double) * 3 * ctx->V);
    return MVS_OK;
}

extern "C" int mvs_profile_enable(mvs_ctx* ctx, int on) {
    if (!ctx) { mvs_set_error("null context"); return MVS_ERR_ARG; }
    MVS_CUDA_CHECK(cudaSetDevice(ctx->device));
    if (on && !ctx->prof_ev[0]) {
        for (int i = 0; i < 2 * MVS_PROF_RING; ++i) MVS_CUDA_CHECK(cudaEventCreate(&ctx->prof_ev[i]));
    }
    ctx->profile = on ? 1 : 0;
    if (on) ctx->prof_n = 0;
    return MVS_OK;
}

extern "C" int mvs_profile_score_ms(mvs_ctx* ctx, float* mean_ms, int* n_kernels) {
    if (!ctx || !mean_ms) { mvs_set_error("mvs_profile_score_ms: null argument"); return MVS_ERR_ARG; }
    if (!ctx->prof_ev[0] || ctx->prof_n == 0) {
        mvs_set_error("mvs_profile_score_ms: no scoring kernel has been timed (call mvs_profile_enable first)");
        return MVS_ERR_STATE;
    }
    MVS_CUDA_CHECK(cudaSetDevice(ctx->device));
    const int n = (int)(ctx->prof_n < MVS_PROF_RING ? ctx->prof_n : MVS_PROF_RING);
    double sum = 0.0;
    for (int k = 0; k < n; ++k) {
        const int slot = (int)((ctx->prof_n - 1 - k) % MVS_PROF_RING);
        float ms = 0.f;
        MVS_CUDA_CHECK(cudaEventSynchronize(ctx->prof_ev[2 * slot + 1]));
        MVS_CUDA_CHECK(cudaEventElapsedTime(&ms, ctx->prof_ev[2 * slot], ctx->prof_ev[2 * slot + 1]));
        sum += ms;
    }
    *mean_ms = (float)(sum / n);
    if (n_kernels) *n_kernels = n;
    return MVS_OK;
}

extern "C" int mvs_profile_probe(mvs_ctx* ctx, int on) {
    if (!ctx) { mvs_set_error("null context"); return MVS_ERR_ARG; }
    ctx->probe_gather = on ? 1 : 0;
    return MVS_OK;
}

extern "C" int64_t mvs_launch_count(const mvs_ctx* ctx) { return ctx ? ctx->launches : 0; }

static inline size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

extern "C" int mvs_score_batch(mvs_ctx* ctx, int mode, int64_t N, const double* c, const double* nrm, const int32_t* ref,
                               double min_ncc, int wid, uint64_t* vis_mask, double* avg, int32_t* count, double* xy,
                               float* ncc, int on_device, void* stream) {
    if (!ctx) { mvs_set_error("mvs_score_batch: null context"); return MVS_ERR_ARG; }
    if (N < 0 || (N > 0 && (!c || !ref || !vis_mask || !count))) {
        mvs_set_error("mvs_score_batch: c, ref, vis_mask and count are required");
        return MVS_ERR_ARG;
    }
    if (mode == MVS_MODE_PMVS)
        return mvs_score_pmvs(ctx, N, c, nrm, ref, nullptr, min_ncc, 2 * wid + 1, 0, 0, 0, vis_mask, avg, count, xy, ncc,
                              nullptr, nullptr, on_device, stream);
    if (mode != MVS_MODE_REFEXACT) {
        mvs_set_error("mvs_score_batch: unknown mode %d", mode);
        return MVS_ERR_ARG;
    }
    if (wid < 1 || wid > 7) { mvs_set_error("mvs_score_batch: wid must be in 1..7 (got %d)", wid); return MVS_ERR_ARG; }
    (void)nrm;
    if (N == 0) return MVS_OK;
    MVS_CUDA_CHECK(cudaSetDevice(ctx->device));
    const int V = ctx->V;
    const size_t mw = (size_t)((V + 63) / 64);
    if (on_device) {
        return mvs_launch_score_refexact(ctx, N, c, ref, min_ncc, wid, vis_mask, avg, count, xy, ncc, (cudaStream_t)stream);
    }
    // host mode: a three-stage pipeline over chunks of the batch -- H2D of chunk k+1, the kernels of
    // chunk k and D2H of chunk k-1 run concurrently on three streams (PCIe is full duplex), three
    // staging slots.  Synchronises before returning.
    static int64_t chunk_pref = 0;                          // MVS_HOST_CHUNK: tuning knob (hypotheses per chunk)
    if (chunk_pref == 0) {
        const char* e = getenv("MVS_HOST_CHUNK");
        chunk_pref = e ? atoll(e) : (1 << 18);
        if (chunk_pref < 1024) chunk_pref = 1 << 18;
    }
    const int64_t CH = N > 2 * chunk_pref ? chunk_pref : N;  // hypotheses per chunk
    const size_t b_c = align256(sizeof(double) * 3 * CH), b_ref = align256(sizeof(int32_t) * CH);
    const size_t b_vis = align256(sizeof(uint64_t) * mw * CH), b_avg = align256(sizeof(double) * CH);
    const size_t b_cnt = align256(sizeof(int32_t) * CH), b_xy = align256(sizeof(double) * 2 * CH);
    const size_t b_ncc = ncc ? align256(sizeof(float) * (size_t)V * CH) : 0;
    const size_t slot = b_c + b_ref + b_vis + b_avg + b_cnt + b_xy + b_ncc;
    const int nslots = N > CH ? 3 : 1;
    int rc = ensure_stage(ctx, slot * nslots);
    if (rc != MVS_OK) return rc;
    if (!ctx->in_stream) {
        MVS_CUDA_CHECK(cudaStreamCreateWithFlags(&ctx->in_stream, cudaStreamNonBlocking));
        MVS_CUDA_CHECK(cudaStreamCreateWithFlags(&ctx->out_stream, cudaStreamNonBlocking));
        for (int i = 0; i < 3; ++i) {
            MVS_CUDA_CHECK(cudaEventCreateWithFlags(&ctx->ev_in[i], cudaEventDisableTiming));
            MVS_CUDA_CHECK(cudaEventCreateWithFlags(&ctx->ev_done[i], cudaEventDisableTiming));
            MVS_CUDA_CHECK(cudaEventCreateWithFlags(&ctx->ev_out[i], cudaEventDisableTiming));
        }
    }
    // window maps first, so that the first chunk's kernels do not wait behind their build
    if ((rc = mvs_build_window_maps(ctx, wid, ctx->own_stream)) != MVS_OK) return rc;
    int64_t k = 0;
    for (int64_t lo = 0; lo < N; lo += CH, ++k) {
        const int64_t n = (N - lo < CH) ? (N - lo) : CH;
        const int sl = (int)(k % nslots);
        uint8_t* p = (uint8_t*)ctx->d_stage + slot * sl;
        double* d_c = (double*)p; p += b_c;
        int32_t* d_ref = (int32_t*)p; p += b_ref;
        uint64_t* d_vis = (uint64_t*)p; p += b_vis;
        double* d_avg = (double*)p; p += b_avg;
        int32_t* d_cnt = (int32_t*)p; p += b_cnt;
        double* d_xy = (double*)p; p += b_xy;
        float* d_ncc = ncc ? (float*)p : nullptr;
        // slot reuse: chunk k-3's results must have left the slot
        if (k >= nslots) MVS_CUDA_CHECK(cudaStreamWaitEvent(ctx->in_stream, ctx->ev_out[sl], 0));
        MVS_CUDA_CHECK(cudaMemcpyAsync(d_c, c + 3 * lo, sizeof(double) * 3 * n, cudaMemcpyHostToDevice, ctx->in_stream));
        MVS_CUDA_CHECK(cudaMemcpyAsync(d_ref, ref + lo, sizeof(int32_t) * n, cudaMemcpyHostToDevice, ctx->in_stream));
        MVS_CUDA_CHECK(cudaEventRecord(ctx->ev_in[sl], ctx->in_stream));
        MVS_CUDA_CHECK(cudaStreamWaitEvent(ctx->own_stream, ctx->ev_in[sl], 0));
        rc = mvs_launch_score_refexact(ctx, n, d_c, d_ref, min_ncc, wid, d_vis, d_avg, d_cnt, d_xy, d_ncc, ctx->own_stream);
        if (rc != MVS_OK) return rc;
        MVS_CUDA_CHECK(cudaEventRecord(ctx->ev_done[sl], ctx->own_stream));
        MVS_CUDA_CHECK(cudaStreamWaitEvent(ctx->out_stream, ctx->ev_done[sl], 0));
        cudaStream_t so = ctx->out_stream;
        MVS_CUDA_CHECK(cudaMemcpyAsync(vis_mask + mw * lo, d_vis, sizeof(uint64_t) * mw * n, cudaMemcpyDeviceToHost, so));
        MVS_CUDA_CHECK(cudaMemcpyAsync(count + lo, d_cnt, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, so));
        if (avg) MVS_CUDA_CHECK(cudaMemcpyAsync(avg + lo, d_avg, sizeof(double) * n, cudaMemcpyDeviceToHost, so));
        if (xy) MVS_CUDA_CHECK(cudaMemcpyAsync(xy + 2 * lo, d_xy, sizeof(double) * 2 * n, cudaMemcpyDeviceToHost, so));
        if (ncc) MVS_CUDA_CHECK(cudaMemcpyAsync(ncc + (size_t)V * lo, d_ncc, sizeof(float) * (size_t)V * n, cudaMemcpyDeviceToHost, so));
        MVS_CUDA_CHECK(cudaEventRecord(ctx->ev_out[sl], so));
    }
    MVS_CUDA_CHECK(cudaStreamSynchronize(ctx->out_stream));
    MVS_CUDA_CHECK(cudaStreamSynchronize(ctx->own_stream));
    return MVS_OK;
}

extern "C" int mvs_compact_accepted_p2p(mvs_ctx* ctx, int64_t N, int64_t index_base, const double* c, const double* nrm,
                                        const int32_t* ref, const uint64_t* vis_mask, const double* avg,
                                        const int32_t* count, const double* xy, const uint8_t* gate, int bound,
                                        void* const* peer_records, int64_t* const* peer_counts, int rank, int world,
                                        int wire, int64_t capacity, void* stream) {
    if (!ctx) { mvs_set_error("mvs_compact_accepted_p2p: null context"); return MVS_ERR_ARG; }
    if (wire != MVS_WIRE_FULL && wire != MVS_WIRE_COMPACT) { mvs_set_error("mvs_compact_accepted_p2p: unknown wire format %d", wire); return MVS_ERR_ARG; }
    if (N < 0 || capacity < 0 || !peer_records || !peer_counts || world < 1 || world > MVS_MAX_PEERS || rank < 0 ||
        rank >= world || (N > 0 && (!c || !ref || !vis_mask || !avg || !count || !xy))) {
        mvs_set_error("mvs_compact_accepted_p2p: need 1 <= world <= %d, 0 <= rank < world, the peer tables and c, ref, "
                      "vis_mask, avg, count, xy", MVS_MAX_PEERS);
        return MVS_ERR_ARG;
    }
    for (int d = 0; d < world; ++d)
        if (!peer_records[d] || !peer_counts[d]) { mvs_set_error("mvs_compact_accepted_p2p: null peer pointer %d", d); return MVS_ERR_ARG; }
    MVS_CUDA_CHECK(cudaSetDevice(ctx->device));
    return mvs_launch_compact_p2p(ctx, N, index_base, c, nrm, ref, vis_mask, avg, count, xy, gate, bound, peer_records,
                                  peer_counts, rank, world, wire, capacity, nullptr, nullptr, (cudaStream_t)stream);
}

extern "C" int mvs_wire_bytes(const mvs_ctx* ctx, int wire) {
    if (!ctx) return 0;
    const int mwb = 8 * ((ctx->V + 63) / 64);
    if (wire == MVS_WIRE_COMPACT) return 56 + mwb;
    if (wire == MVS_WIRE_FULL) return (int)sizeof(mvs_patch_record) + mwb;
    return 0;
}

extern "C" int mvs_records_expand(mvs_ctx* ctx, int wire, const void* wire_records, int64_t n, void* records, void* stream) {
    if (!ctx) { mvs_set_error("mvs_records_expand: null context"); return MVS_ERR_ARG; }
    if (n < 0 || (n > 0 && (!wire_records || !records))) { mvs_set_error("mvs_records_expand: null buffer"); return MVS_ERR_ARG; }
    MVS_CUDA_CHECK(cudaSetDevice(ctx->device));
    if (wire == MVS_WIRE_FULL) {
        if (n > 0 && wire_records != records)
            MVS_CUDA_CHECK(cudaMemcpyAsync(records, wire_records, (size_t)n * mvs_wire_bytes(ctx, wire), cudaMemcpyDeviceToDevice,
                                           (cudaStream_t)stream));
        return MVS_OK;
    }
    if (wire != MVS_WIRE_COMPACT) { mvs_set_error("mvs_records_expand: unknown wire format %d", wire); return MVS_ERR_ARG; }
    return mvs_launch_records_expand(ctx, wire_records, n, records, (cudaStream_t)stream);
}

extern "C" int mvs_score_pmvs(mvs_ctx* ctx, int64_t N, const double* c, const double* nrm, const int32_t* ref,
                              const uint64_t* cand, double min_ncc, int mu, int flags, int group, int bound,
                              uint64_t* vis_mask, double* avg, int32_t* count, double* xy, float* ncc, int32_t* best_idx,
                              double* best_avg, int on_device, void* stream) {
    if (!ctx) { mvs_set_error("mvs_score_pmvs: null context"); return MVS_ERR_ARG; }
    const bool reduce = (flags & MVS_PMVS_REDUCE_TO_REFEXACT) != 0;
    if (N < 0 || (N > 0 && (!c || !ref || (!nrm && !reduce)))) {
        mvs_set_error("mvs_score_pmvs: c, ref and (unless reducing to Mode A) nrm are required");
        return MVS_ERR_ARG;
    }
    if (mu != 3 && mu != 5 && mu != 7 && mu != 9 && mu != 11) {
        mvs_set_error("mvs_score_pmvs: mu must be 3, 5, 7, 9 or 11 (got %d)", mu);
        return MVS_ERR_ARG;
    }
    if (group > 1 && !best_idx) { mvs_set_error("mvs_score_pmvs: best_idx is required when group > 1"); return MVS_ERR_ARG; }
    if (group <= 1 && (!vis_mask || !count)) {
        mvs_set_error("mvs_score_pmvs: vis_mask and count are required without selection");
        return MVS_ERR_ARG;
    }
    if (N == 0) return MVS_OK;
    MVS_CUDA_CHECK(cudaSetDevice(ctx->device));
    if (on_device)
        return mvs_launch_score_pmvs(ctx, N, c, nrm, ref, cand, min_ncc, mu, flags, group, bound, vis_mask, avg, count, xy,
                                     ncc, best_idx, best_avg, (cudaStream_t)stream);
    // host mode: stage in, run, stage out, synchronise
    const int V = ctx->V;
    const size_t mw = (size_t)((V + 63) / 64);
    const int g = group > 1 ? group : 1;
    const int64_t n_sets = (N + g - 1) / g;
    const size_t b_c = align256(sizeof(double) * 3 * N), b_n = nrm ? b_c : 0, b_ref = align256(sizeof(int32_t) * N);
    const size_t b_cand = cand ? align256(sizeof(uint64_t) * mw * N) : 0;
    const size_t b_vis = vis_mask ? align256(sizeof(uint64_t) * mw * N) : 0, b_avg = avg ? align256(sizeof(double) * N) : 0;
    const size_t b_cnt = count ? align256(sizeof(int32_t) * N) : 0, b_xy = xy ? align256(sizeof(double) * 2 * N) : 0;
    const size_t b_ncc = ncc ? align256(sizeof(float) * (size_t)V * N) : 0;
    const size_t b_bi = best_idx ? align256(sizeof(int32_t) * n_sets) : 0, b_ba = best_avg ? align256(sizeof(double) * n_sets) : 0;
    int rc = ensure_stage(ctx, b_c + b_n + b_ref + b_cand + b_vis + b_avg + b_cnt + b_xy + b_ncc + b_bi + b_ba);
    if (rc != MVS_OK) return rc;
    cudaStream_t s = ctx->own_stream;
    uint8_t* p = (uint8_t*)ctx->d_stage;
    auto take = [&](size_t bytes) -> void* { void* q = bytes ? (void*)p : nullptr; p += bytes; return q; };
    double* d_c = (double*)take(b_c);
    double* d_n = (double*)take(b_n);
    int32_t* d_ref = (int32_t*)take(b_ref);
    uint64_t* d_cand = (uint64_t*)take(b_cand);
    uint64_t* d_vis = (uint64_t*)take(b_vis);
    double* d_avg = (double*)take(b_avg);
    int32_t* d_cnt = (int32_t*)take(b_cnt);
    double* d_xy = (double*)take(b_xy);
    float* d_ncc = (float*)take(b_ncc);
    int32_t* d_bi = (int32_t*)take(b_bi);
    double* d_ba = (double*)take(b_ba);
    MVS_CUDA_CHECK(cudaMemcpyAsync(d_c, c, sizeof(double) * 3 * N, cudaMemcpyHostToDevice, s));
    if (nrm) MVS_CUDA_CHECK(cudaMemcpyAsync(d_n, nrm, sizeof(double) * 3 * N, cudaMemcpyHostToDevice, s));
    MVS_CUDA_CHECK(cudaMemcpyAsync(d_ref, ref, sizeof(int32_t) * N, cudaMemcpyHostToDevice, s));
    if (cand) MVS_CUDA_CHECK(cudaMemcpyAsync(d_cand, cand, sizeof(uint64_t) * mw * N, cudaMemcpyHostToDevice, s));
    rc = mvs_launch_score_pmvs(ctx, N, d_c, d_n, d_ref, d_cand, min_ncc, mu, flags, group, bound, d_vis, d_avg, d_cnt, d_xy,
                               d_ncc, d_bi, d_ba, s);
    if (rc != MVS_OK) return rc;
    if (vis_mask) MVS_CUDA_CHECK(cudaMemcpyAsync(vis_mask, d_vis, sizeof(uint64_t) * mw * N, cudaMemcpyDeviceToHost, s));
    if (count) MVS_CUDA_CHECK(cudaMemcpyAsync(count, d_cnt, sizeof(int32_t) * N, cudaMemcpyDeviceToHost, s));
    if (avg) MVS_CUDA_CHECK(cudaMemcpyAsync(avg, d_avg, sizeof(double) * N, cudaMemcpyDeviceToHost, s));
    if (xy) MVS_CUDA_CHECK(cudaMemcpyAsync(xy, d_xy, sizeof(double) * 2 * N, cudaMemcpyDeviceToHost, s));
    if (ncc) MVS_CUDA_CHECK(cudaMemcpyAsync(ncc, d_ncc, sizeof(float) * (size_t)V * N, cudaMemcpyDeviceToHost, s));
    if (best_idx) MVS_CUDA_CHECK(cudaMemcpyAsync(best_idx, d_bi, sizeof(int32_t) * n_sets, cudaMemcpyDeviceToHost, s));
    if (best_avg) MVS_CUDA_CHECK(cudaMemcpyAsync(best_avg, d_ba, sizeof(double) * n_sets, cudaMemcpyDeviceToHost, s));
    MVS_CUDA_CHECK(cudaStreamSynchronize(s));
    return MVS_OK;
}

extern "C" int mvs_select_best(mvs_ctx* ctx, int64_t N, int group, const double* avg, const int32_t* count, int bound,
                               int32_t* best_idx, double* best_avg, void* stream) {
    if (!ctx) { mvs_set_error("mvs_select_best: null context"); return MVS_ERR_ARG; }
    if (N < 0 || group < 1 || (N > 0 && (!avg || !count || !best_idx))) {
        mvs_set_error("mvs_select_best: need group >= 1 and avg, count, best_idx");
        return MVS_ERR_ARG;
    }
    MVS_CUDA_CHECK(cudaSetDevice(ctx->device));
    return mvs_launch_select_best(ctx, N, group, avg, count, bound, best_idx, best_avg, (cudaStream_t)stream);
}

extern "C" int mvs_record_bytes(const mvs_ctx* ctx) {
    return ctx ? (int)(sizeof(mvs_patch_record) + 8 * ((ctx->V + 63) / 64)) : 0;
}

extern "C" int mvs_compact_accepted(mvs_ctx* ctx, int64_t N, int64_t index_base, const double* c, const double* nrm,
                                    const int32_t* ref, const uint64_t* vis_mask, const double* avg, const int32_t* count,
                                    const double* xy, const uint8_t* gate, int bound, void* records, int64_t capacity,
                                    int64_t* n_out, void* stream) {
    if (!ctx) { mvs_set_error("mvs_compact_accepted: null context"); return MVS_ERR_ARG; }
    if (N < 0 || capacity < 0 || !n_out || (N > 0 && (!c || !ref || !vis_mask || !avg || !count || !xy || !records))) {
        mvs_set_error("mvs_compact_accepted: c, ref, vis_mask, avg, count, xy, records and n_out are required");
        return MVS_ERR_ARG;
    }
    MVS_CUDA_CHECK(cudaSetDevice(ctx->device));
    return mvs_launch_compact(ctx, N, index_base, c, nrm, ref, vis_mask, avg, count, xy, gate, bound, records, capacity,
                              n_out, nullptr, nullptr, (cudaStream_t)stream);
}
