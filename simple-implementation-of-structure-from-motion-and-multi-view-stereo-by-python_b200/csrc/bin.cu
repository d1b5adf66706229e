// Projection + spatial binning of a hypothesis batch (feeds K1, ncc_refexact.cu).
//
// Mode A samples every view at the reference view's projection (MVS2.py:63,68), so two
// hypotheses whose anchors are a few pixels apart read almost the same bytes of the
// view-interleaved stack, whatever their reference views are.  Ordering a batch by
// (8x8-pixel anchor tile, row, first pixel group of the window) turns the per-hypothesis
// gather (V*121 bytes each) into L1 hits -- a tile's bytes are fetched from L2 about once
// per CTA instead of once per hypothesis -- and puts hypotheses that read exactly the
// same 16-byte quads next to each other, so K1 loads them once for two hypotheses.
//   bin_project : one thread per hypothesis -- fp64 projection in cv2's operation order
//                 (utils.py:241-244), int() truncation and the bounds rule of
//                 HarrisFeatures.py:128; writes xy, the packed anchor, the empty result of
//                 rejected hypotheses; counts the hypothesis into its tile's histogram
//   (exclusive scan of the histogram, scan.cu)
//   bin_scatter : entry[pos] = (hypothesis index, anchor)
// The order inside a tile depends on atomic arrival, the RESULTS do not: each hypothesis
// is scored independently and written to its own output slot.
#include "project.cuh"
#include "scan.cuh"

// anchor tile = 2^TR rows x 2^TC pixel groups (8 rows x 2 groups = 8 x 8 pixels measured best)
#ifndef MVS_TILE_ROWS_LOG2
#define MVS_TILE_ROWS_LOG2 3
#endif
#ifndef MVS_TILE_GROUPS_LOG2
#define MVS_TILE_GROUPS_LOG2 1
#endif

__global__ void __launch_bounds__(256)
    bin_project(const CamProj* __restrict__ cams, int V, int H, int W, int wid, int64_t N, const double* __restrict__ c,
                const int32_t* __restrict__ ref, int tiles_x, int n_tiles, int32_t* __restrict__ hist,
                int32_t* __restrict__ key, int32_t* __restrict__ rank, uint32_t* __restrict__ anchor,
                uint64_t* __restrict__ vis_out, double* __restrict__ avg_out, int32_t* __restrict__ count_out,
                double* __restrict__ xy_out, float* __restrict__ ncc_out, int cams_in_smem) {
    // cameras go to shared memory once per CTA: every thread reads the 128-byte record of ITS reference view.  The rows
    // are padded to 17 doubles: with the natural 16-double pitch field k of EVERY view falls on the same two banks, and a
    // warp whose lanes have ~24 different reference views paid a ~24-way bank conflict on each of its 16 field reads.
    constexpr int PITCH = 17;
    extern __shared__ __align__(16) unsigned char s_cam_raw[];
    const double* s_cam = nullptr;
    if (cams_in_smem) {
        double* dst = reinterpret_cast<double*>(s_cam_raw);
        const double* src = reinterpret_cast<const double*>(cams);
        constexpr int F = (int)(sizeof(CamProj) / sizeof(double));      // 16
        for (int i = threadIdx.x; i < V * F; i += blockDim.x) dst[(i / F) * PITCH + (i % F)] = src[i];
        __syncthreads();
        s_cam = dst;
    }
    const int mw = (V + 63) >> 6;
    for (int64_t h = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; h < N; h += (int64_t)gridDim.x * blockDim.x) {
        const int r = __ldg(ref + h);
        double x = nan(""), y = nan("");
        int row = 0, col = 0;
        bool valid = false;
        if (r >= 0 && r < V) {
            if (s_cam) {
                CamProj cam;
                const double* rowp = s_cam + r * PITCH;
#pragma unroll
                for (int i = 0; i < 9; ++i) cam.r[i] = rowp[i];
#pragma unroll
                for (int i = 0; i < 3; ++i) cam.t[i] = rowp[9 + i];
                cam.fx = rowp[12]; cam.fy = rowp[13]; cam.cx = rowp[14]; cam.cy = rowp[15];
                project_ref(cam, __ldg(c + 3 * h), __ldg(c + 3 * h + 1), __ldg(c + 3 * h + 2), x, y);
            } else {
                project_ref(cams[r], __ldg(c + 3 * h), __ldg(c + 3 * h + 1), __ldg(c + 3 * h + 2), x, y);
            }
            valid = window_anchor(x, y, H, W, wid, row, col);
        }
        if (xy_out) {
            xy_out[2 * h] = x;
            xy_out[2 * h + 1] = y;
        }
        anchor[h] = valid ? (((uint32_t)row << 16) | (uint32_t)col) : MVS_ANCHOR_INVALID;
        if (!valid) {                                      // getDescFeatures -> [None]: V = [], avg = 0
            for (int w = 0; w < mw; ++w) vis_out[h * mw + w] = 0ull;
            count_out[h] = 0;
            if (avg_out) avg_out[h] = 0.0;
            if (ncc_out)
                for (int v = 0; v < V; ++v) ncc_out[h * V + v] = nanf("");
        }
        if (hist) {
            // bins = (row, first pixel group of the window), ordered tile by tile (8 rows x 2 groups)
            const int cg = (col - wid) >> 2;
            constexpr int TR = MVS_TILE_ROWS_LOG2, TC = MVS_TILE_GROUPS_LOG2;
            const int k = valid ? (((row >> TR) * tiles_x + (cg >> TC)) << (TR + TC)) + ((row & ((1 << TR) - 1)) << TC) +
                                      (cg & ((1 << TC) - 1))
                                : n_tiles;
            key[h] = k;
            rank[h] = atomicAdd(hist + k, 1);
        }
    }
}

__global__ void __launch_bounds__(256)
    bin_scatter(int64_t N, const int32_t* __restrict__ hist_excl, const int32_t* __restrict__ key,
                const int32_t* __restrict__ rank, const uint32_t* __restrict__ anchor, uint2* __restrict__ entry,
                uint8_t* __restrict__ part, int part_size) {
    for (int64_t h = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; h < N; h += (int64_t)gridDim.x * blockDim.x) {
        const int pos = hist_excl[key[h]] + rank[h];
        entry[pos] = make_uint2((uint32_t)h, anchor[h]);   // one 8-byte scattered store per hypothesis
        // partitioned publish (exchange.cu): which position range -- i.e. which of K1's launches -- scores this hypothesis
        if (part) part[h] = (uint8_t)(pos / part_size);
    }
}

int mvs_bin_hypotheses(mvs_ctx* ctx, int64_t N, const double* c, const int32_t* ref, int wid, bool sort, uint64_t* vis,
                       double* avg, int32_t* count, double* xy, float* ncc, cudaStream_t s, uint8_t* part, int64_t part_size) {
    int rc;
    if (N >= (1ll << 31)) {
        mvs_set_error("batches of 2^31 or more hypotheses are not supported (got %lld)", (long long)N);
        return MVS_ERR_ARG;
    }
    if ((rc = mvs_ensure((void**)&ctx->d_bin_anchor, &ctx->bin_anchor_bytes, sizeof(uint32_t) * N, "anchors")) != MVS_OK)
        return rc;
    const int tiles_x = (((ctx->W + 3) >> 2) >> MVS_TILE_GROUPS_LOG2) + 1, tiles_y = (ctx->H >> MVS_TILE_ROWS_LOG2) + 1;
    const int n_tiles = (tiles_x * tiles_y) << (MVS_TILE_ROWS_LOG2 + MVS_TILE_GROUPS_LOG2);   // bins: one per (row, pixel group)
    if (sort) {
        if ((rc = mvs_ensure((void**)&ctx->d_bin_hist, &ctx->bin_hist_bytes, sizeof(int32_t) * (n_tiles + 2), "tile histogram")) != MVS_OK ||
            (rc = mvs_ensure((void**)&ctx->d_bin_key, &ctx->bin_key_bytes, sizeof(int32_t) * N, "tile keys")) != MVS_OK ||
            (rc = mvs_ensure((void**)&ctx->d_bin_rank, &ctx->bin_rank_bytes, sizeof(int32_t) * N, "tile ranks")) != MVS_OK ||
            (rc = mvs_ensure((void**)&ctx->d_bin_entry, &ctx->bin_entry_bytes, sizeof(uint2) * N, "ordered entries")) != MVS_OK ||
            (rc = mvs_ensure((void**)&ctx->d_bin_scan, &ctx->bin_scan_bytes, sizeof(int64_t) * ((n_tiles + 1 + 1023) / 1024 + 2), "scan")) != MVS_OK)
            return rc;
        MVS_CUDA_CHECK(cudaMemsetAsync(ctx->d_bin_hist, 0, sizeof(int32_t) * (n_tiles + 2), s));
    }
    int64_t blocks = (N + 255) / 256;
    const int64_t cap = (int64_t)ctx->sm_count * 16;
    if (blocks > cap) blocks = cap;
    const size_t cam_bytes = 17 * sizeof(double) * (size_t)ctx->V;      // padded rows, see bin_project
    const int cams_in_smem = cam_bytes <= 48 * 1024 && N >= 4096;      // up to 384 views; tiny batches skip the staging
    bin_project<<<(int)blocks, 256, cams_in_smem ? cam_bytes : 0, s>>>(ctx->d_cam, ctx->V, ctx->H, ctx->W, wid, N, c, ref, tiles_x,
                                                                      n_tiles, sort ? ctx->d_bin_hist : nullptr, ctx->d_bin_key,
                                                                      ctx->d_bin_rank, ctx->d_bin_anchor, vis, avg, count, xy, ncc,
                                                                      cams_in_smem);
    ctx->launches++;
    if (sort) {
        int64_t* d_total = ctx->d_bin_scan + (n_tiles + 1 + 1023) / 1024;
        if ((rc = mvs_exclusive_scan_i32(ctx->d_bin_hist, n_tiles + 1, ctx->d_bin_scan, d_total, s)) != MVS_OK) return rc;
        ctx->launches += mvs_scan_launches(n_tiles + 1);
        bin_scatter<<<(int)blocks, 256, 0, s>>>(N, ctx->d_bin_hist, ctx->d_bin_key, ctx->d_bin_rank, ctx->d_bin_anchor,
                                                (uint2*)ctx->d_bin_entry, part, (int)part_size);
        ctx->launches++;
    }
    MVS_CUDA_CHECK(cudaGetLastError());
    return MVS_OK;
}
