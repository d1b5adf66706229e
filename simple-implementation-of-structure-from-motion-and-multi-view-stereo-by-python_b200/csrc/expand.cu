// K3 expand_candidates + K4 accept/commit: the reference's patch_expansion
// (MVS2.py:308-404) restructured into synchronous rounds, and its CellTable
// (MVS2.py:80-120) as a byte grid in HBM.
//
// Round semantics (DESIGN.md "Rounds"; oracle/expansion.py::expand_round is the CPU
// restatement): every frontier patch is expanded against the ROUND-START table; slots
// s = (f*V + v)*4 + k are de-duplicated per tested cell (lowest s wins, by an epoch-
// tagged atomicMax so the claim grid never needs clearing); survivors are scored and
// gated; accepted records are committed in slot order.  All fp64 geometry uses
// explicit round-to-nearest intrinsics (no FMA contraction) so that candidates are
// bit-identical to the NumPy oracle and hence truncate to the same pixels.
#include "scan.cuh"

#define FULL 0xffffffffu

__device__ __forceinline__ double xmul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double xadd(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double xsub(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ double xdiv(double a, double b) { return __ddiv_rn(a, b); }
__device__ __forceinline__ double dot3(double a0, double a1, double a2, double b0, double b1, double b2) {
    return xadd(xadd(xmul(a0, b0), xmul(a1, b1)), xmul(a2, b2));
}

__device__ __forceinline__ const mvs_patch_record* rec_at(const uint8_t* base, int64_t i, int rec_bytes) {
    return reinterpret_cast<const mvs_patch_record*>(base + i * rec_bytes);
}
__device__ __forceinline__ const uint64_t* rec_vis(const mvs_patch_record* r) {
    return reinterpret_cast<const uint64_t*>(r + 1);
}

// CellTable.which_cell (MVS2.py:113-114): floor(x / cell_size) in fp64
__device__ __forceinline__ bool which_cell(double x, double y, int cs, int& ci, int& cj) {
    if (!(isfinite(x) && isfinite(y))) return false;
    const double fi = floor(xdiv(x, (double)cs)), fj = floor(xdiv(y, (double)cs));
    if (fabs(fi) > 1e9 || fabs(fj) > 1e9) return false;
    ci = (int)fi;
    cj = (int)fj;
    return true;
}

__constant__ int c_di[4] = {-1, -1, 1, 1};      // loop order of MVS2.py:331-332
__constant__ int c_dj[4] = {-1, 1, -1, 1};

// claim value: newer epochs always win over stale ones; inside an epoch the LOWEST slot wins
__device__ __forceinline__ unsigned long long claim_value(unsigned epoch, long long slot) {
    return ((unsigned long long)epoch << 40) | (unsigned long long)((1ll << 40) - 1 - slot);
}

// pass 0: claim cells; pass 1: count surviving slots; pass 2: emit candidates.
// ONE WARP per frontier patch, lanes span its views (lane v handles view v, v + 32, ...): the reference's
// loop order (view ascending, then the four diagonals, MVS2.py:328-332) is the lane order, so the number of
// surviving slots before a lane's is a warp prefix sum and the candidates of a patch land in ascending slot
// order without any per-thread serial walk over the visible set.
template <int PASS>
__global__ void __launch_bounds__(256)
    expand_slots(const uint8_t* __restrict__ frontier, int64_t F, int rec_bytes, int V, int cs, int wc, int hc,
                 const uint8_t* __restrict__ cells, unsigned long long* __restrict__ claim, unsigned epoch,
                 int32_t* __restrict__ counts /*[F]*/, const CamGeom* __restrict__ geom, int64_t* __restrict__ cand_slot,
                 int64_t* __restrict__ cand_parent, double* __restrict__ cand_c, double* __restrict__ cand_n,
                 int32_t* __restrict__ cand_ref, int32_t* __restrict__ cand_px, const int64_t* __restrict__ d_F = nullptr,
                 int64_t F_clamp = 0, uint8_t* __restrict__ live_buf = nullptr, int live_pitch = 0) {
    const int lane = threadIdx.x & 31;
    const int64_t f = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    if (d_F) {
        // frontier size still on the device (the commit that produced it has not been read back): the grid
        // covers an upper bound `F`, warps past the real size only zero their count
        const int64_t Fd = *d_F < F_clamp ? *d_F : F_clamp;
        if (f >= Fd) {
            if (PASS == 1 && f < F && lane == 0) counts[f] = 0;
            return;
        }
    }
    if (f >= F) return;
    const mvs_patch_record* p = rec_at(frontier, f, rec_bytes);
    const uint64_t* vis = rec_vis(p);
    int ci = 0, cj = 0;
    const bool ok = which_cell(p->xy[0], p->xy[1], cs, ci, cj);
    int64_t base = (PASS == 2) ? (int64_t)counts[f] : 0;          // candidates of this patch emitted so far
    int total = 0;
    for (int v0 = 0; v0 < V; v0 += 32) {
        const int v = v0 + lane;
        const bool seen = ok && v < V && ((vis[v >> 6] >> (v & 63)) & 1ull);
        unsigned live = 0u;                                       // bit k: diagonal k of view v survives
        if (PASS == 2 && live_buf) {
            // the count pass left the surviving diagonals of every (patch, view): no second probe of cells and claims
            live = (v < V) ? live_buf[f * live_pitch + v] : 0u;
        } else if (seen) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int ti = ci + c_di[k], tj = cj + c_dj[k];
                if (ti < 0 || ti >= wc || tj < 0 || tj >= hc) continue;                  // is_vacant: out of range
                const int64_t cell = ((int64_t)v * wc + ti) * hc + tj;
                if (!cells[cell]) continue;                                              // MVS2.py:333
                const long long slot = ((long long)f * V + v) * 4 + k;
                const unsigned long long mine = claim_value(epoch, slot);
                if (PASS == 0) {
                    atomicMax(claim + cell, mine);
                    continue;
                }
                if (claim[cell] == mine) live |= 1u << k;                                // else a lower slot tests this cell
            }
        }
        if (PASS == 0) continue;
        if (PASS == 1 && live_buf && v < V) live_buf[f * live_pitch + v] = (uint8_t)live;
        const int n_live = __popc(live);
        int incl = n_live;
#pragma unroll
        for (int sft = 1; sft < 32; sft <<= 1) {
            const int y = __shfl_up_sync(FULL, incl, sft);
            if (lane >= sft) incl += y;
        }
        const int chunk_total = __shfl_sync(FULL, incl, 31);
        if (PASS == 2 && live) {
            // only WHICH slots survive is decided here; their geometry is one thread per candidate (expand_geometry)
            int64_t out = base + (incl - n_live);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                if (!((live >> k) & 1u)) continue;
                cand_slot[out] = ((long long)f * V + v) * 4 + k;
                cand_parent[out] = f;
                ++out;
            }
        }
        base += chunk_total;
        total += chunk_total;
    }
    if (PASS == 1 && lane == 0) counts[f] = total;
}

// candidate geometry, MVS2.py:334-358, one thread per surviving slot (slot = (f*V + v)*4 + k)
__global__ void __launch_bounds__(128)
    expand_geometry(const uint8_t* __restrict__ frontier, int rec_bytes, int V, int cs, int64_t M,
                    const int64_t* __restrict__ cand_slot, const CamGeom* __restrict__ geom, double* __restrict__ cand_c,
                    double* __restrict__ cand_n, int32_t* __restrict__ cand_ref, int32_t* __restrict__ cand_px,
                    uint8_t* __restrict__ gate = nullptr, double dist_limit = 0.0) {
    const int64_t out = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (out >= M) return;
    const long long slot = cand_slot[out];
    const int k = (int)(slot & 3);
    const int v = (int)((slot >> 2) % V);
    const int64_t f = (slot >> 2) / V;
    const mvs_patch_record* p = rec_at(frontier, f, rec_bytes);
    int ci = 0, cj = 0;
    which_cell(p->xy[0], p->xy[1], cs, ci, cj);                                        // a surviving slot has a valid cell
    const CamGeom& g = geom[v];
    const int di = c_di[k];
    const double u = xmul((double)cs, xadd((double)(ci + di), 0.5));
    const double vv = xmul((double)cs, xadd((double)(cj + di), 0.5));                  // sic: di on both axes (MVS2.py:334)
    const double a0 = xsub(u, g.cx), a1 = xsub(vv, g.cy), a2 = xdiv(xadd(g.fx, g.fy), 2.0);
    // R^T a + C  (sic: "+ C", MVS2.py:353)
    const double P0 = xadd(dot3(g.rf[0], g.rf[3], g.rf[6], a0, a1, a2), g.C[0]);
    const double P1 = xadd(dot3(g.rf[1], g.rf[4], g.rf[7], a0, a1, a2), g.C[1]);
    const double P2 = xadd(dot3(g.rf[2], g.rf[5], g.rf[8], a0, a1, a2), g.C[2]);
    const double nrm = sqrt(dot3(P0, P1, P2, P0, P1, P2));
    const double d0 = xdiv(P0, nrm), d1 = xdiv(P1, nrm), d2 = xdiv(P2, nrm);
    // ray_plane_intersection (MVS2.py:302-306) with origin O = C
    const double dot_out = dot3(d0, d1, d2, p->n[0], p->n[1], p->n[2]);
    const double w0 = xsub(p->c[0], g.C[0]), w1 = xsub(p->c[1], g.C[1]), w2 = xsub(p->c[2], g.C[2]);
    const double tpar = xdiv(dot3(w0, w1, w2, p->n[0], p->n[1], p->n[2]), dot_out);
    const double X0 = xadd(g.C[0], xmul(tpar, d0)), X1 = xadd(g.C[1], xmul(tpar, d1)), X2 = xadd(g.C[2], xmul(tpar, d2));
    const double q0 = xsub(g.C[0], X0), q1 = xsub(g.C[1], X1), q2 = xsub(g.C[2], X2);
    const double dist = sqrt(dot3(q0, q1, q2, q0, q1, q2));
    cand_c[3 * out] = X0; cand_c[3 * out + 1] = X1; cand_c[3 * out + 2] = X2;
    cand_n[3 * out] = xdiv(q0, dist); cand_n[3 * out + 1] = xdiv(q1, dist); cand_n[3 * out + 2] = xdiv(q2, dist);
    cand_ref[out] = v;
    cand_px[2 * out] = (int)u;
    cand_px[2 * out + 1] = (int)vv;
    if (gate) {
        // the accept gate of MVS2.py:369 depends on geometry only (same operations as expand_gate)
        const double n0 = xdiv(q0, dist), n1 = xdiv(q1, dist), n2 = xdiv(q2, dist);
        const double e0 = xsub(p->c[0], X0), e1 = xsub(p->c[1], X1), e2 = xsub(p->c[2], X2);
        const double a = dot3(e0, e1, e2, p->n[0], p->n[1], p->n[2]);
        const double b = dot3(e0, e1, e2, n0, n1, n2);
        const bool neigh = fabs(xadd(a, b)) < 0.1;
        const double dd = sqrt(dot3(e0, e1, e2, e0, e1, e2));
        gate[out] = (neigh && dd < dist_limit) ? 1 : 0;
    }
}

// accept gate of MVS2.py:369 without the visible_ct clause (applied by the compaction):
// is_patch_neighbor(parent, cand, 0.1) and distance(parent.c, cand.c) < 0.05/scale
__global__ void __launch_bounds__(256)
    expand_gate(const uint8_t* __restrict__ frontier, int rec_bytes, int64_t begin, int64_t end,
                const int64_t* __restrict__ cand_parent, const double* __restrict__ cand_c,
                const double* __restrict__ cand_n, double dist_limit, uint8_t* __restrict__ gate) {
    const int64_t m = begin + blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (m >= end) return;
    const mvs_patch_record* p = rec_at(frontier, cand_parent[m], rec_bytes);
    const double w0 = xsub(p->c[0], cand_c[3 * m]), w1 = xsub(p->c[1], cand_c[3 * m + 1]), w2 = xsub(p->c[2], cand_c[3 * m + 2]);
    const double a = dot3(w0, w1, w2, p->n[0], p->n[1], p->n[2]);
    const double b = dot3(w0, w1, w2, cand_n[3 * m], cand_n[3 * m + 1], cand_n[3 * m + 2]);
    const bool neigh = fabs(xadd(a, b)) < 0.1;
    const double dist = sqrt(dot3(w0, w1, w2, w0, w1, w2));
    gate[m] = (neigh && dist < dist_limit) ? 1 : 0;          // NaN compares false, as in Python
}

// commit, step 1: keep[i] = 0 for a dj=+1 record whose dj=-1 sibling (slot-1) is the previous record
__global__ void __launch_bounds__(256)
    commit_flags(const uint8_t* __restrict__ recs, int64_t n, int rec_bytes, int32_t* __restrict__ keep) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const long long slot = rec_at(recs, i, rec_bytes)->index;
    bool k = true;
    if ((slot & 1) && i > 0 && rec_at(recs, i - 1, rec_bytes)->index == slot - 1) k = false;
    keep[i] = k ? 1 : 0;
}

// commit, step 2: copy kept records to the next frontier (order preserved) and clear their cells
__global__ void __launch_bounds__(256)
    commit_apply(const uint8_t* __restrict__ recs, int64_t n, int rec_bytes, const int32_t* __restrict__ offsets,
                 const int64_t* __restrict__ total, uint8_t* __restrict__ next, int V, int cs, int wc, int hc,
                 uint8_t* __restrict__ cells) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int64_t off = offsets[i];
    const int64_t nxt = (i + 1 < n) ? (int64_t)offsets[i + 1] : *total;
    if (nxt == off) return;                                   // dropped sibling
    const mvs_patch_record* r = rec_at(recs, i, rec_bytes);
    if (next) {
        const uint32_t* src = reinterpret_cast<const uint32_t*>(r);
        uint32_t* dst = reinterpret_cast<uint32_t*>(next + off * rec_bytes);
        for (int q = 0; q < rec_bytes / 4; ++q) dst[q] = src[q];
    }
    int ci, cj;
    if (!which_cell(r->xy[0], r->xy[1], cs, ci, cj)) return;
    if (ci < 0 || ci >= wc || cj < 0 || cj >= hc) return;     // the reference would stop in pdb here
    const uint64_t* vis = rec_vis(r);
    for (int w = 0; w < (V + 63) / 64; ++w) {
        uint64_t bits = vis[w];
        while (bits) {
            const int v = w * 64 + __ffsll((long long)bits) - 1;
            bits &= bits - 1;
            if (v < V) cells[((int64_t)v * wc + ci) * hc + cj] = 0;     // MVS2.py:105
        }
    }
}

__global__ void __launch_bounds__(256)
    cells_fill_kernel(const uint8_t* __restrict__ recs, int64_t n, int rec_bytes, int V, int cs, int wc, int hc,
                      uint8_t* __restrict__ cells) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const mvs_patch_record* r = rec_at(recs, i, rec_bytes);
    int ci, cj;
    if (!which_cell(r->xy[0], r->xy[1], cs, ci, cj)) return;
    if (ci < 0 || ci >= wc || cj < 0 || cj >= hc) return;
    const uint64_t* vis = rec_vis(r);
    for (int w = 0; w < (V + 63) / 64; ++w) {
        uint64_t bits = vis[w];
        while (bits) {
            const int v = w * 64 + __ffsll((long long)bits) - 1;
            bits &= bits - 1;
            if (v < V) cells[((int64_t)v * wc + ci) * hc + cj] = 0;
        }
    }
}

// ---------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------
static int rec_bytes_of(const mvs_ctx* ctx) { return (int)(sizeof(mvs_patch_record) + 8 * ((ctx->V + 63) / 64)); }

extern "C" int mvs_cells_init(mvs_ctx* ctx, int cell_size, const uint8_t* table_host) {
    if (!ctx || cell_size < 1) { mvs_set_error("mvs_cells_init: bad argument"); return MVS_ERR_ARG; }
    MVS_CUDA_CHECK(cudaSetDevice(ctx->device));
    ctx->cell_size = cell_size;
    ctx->wc = (ctx->W - 1 + cell_size - 1) / cell_size;       // ceil((W-1)/cs), MVS2.py:88
    ctx->hc = (ctx->H - 1 + cell_size - 1) / cell_size;
    const size_t ncell = (size_t)ctx->V * ctx->wc * ctx->hc;
    int rc = mvs_ensure((void**)&ctx->d_cells, &ctx->cells_bytes, ncell + 16, "cell table");
    if (rc != MVS_OK) return rc;
    rc = mvs_ensure((void**)&ctx->d_claim, &ctx->claim_bytes, ncell * sizeof(unsigned long long), "claim grid");
    if (rc != MVS_OK) return rc;
    if (table_host)
        MVS_CUDA_CHECK(cudaMemcpy(ctx->d_cells, table_host, ncell, cudaMemcpyHostToDevice));
    else
        MVS_CUDA_CHECK(cudaMemset(ctx->d_cells, 1, ncell));
    MVS_CUDA_CHECK(cudaMemset(ctx->d_claim, 0, ncell * sizeof(unsigned long long)));
    ctx->epoch = 0;
    ctx->n_cand = 0;
    return MVS_OK;
}

extern "C" int mvs_cells_shape(const mvs_ctx* ctx, int* cell_size, int* wc, int* hc) {
    if (!ctx || !ctx->d_cells) { mvs_set_error("cell table not initialised"); return MVS_ERR_STATE; }
    if (cell_size) *cell_size = ctx->cell_size;
    if (wc) *wc = ctx->wc;
    if (hc) *hc = ctx->hc;
    return MVS_OK;
}

extern "C" int mvs_cells_download(mvs_ctx* ctx, uint8_t* table_host) {
    if (!ctx || !ctx->d_cells || !table_host) { mvs_set_error("cell table not initialised"); return MVS_ERR_STATE; }
    MVS_CUDA_CHECK(cudaSetDevice(ctx->device));
    MVS_CUDA_CHECK(cudaDeviceSynchronize());
    MVS_CUDA_CHECK(cudaMemcpy(table_host, ctx->d_cells, (size_t)ctx->V * ctx->wc * ctx->hc, cudaMemcpyDeviceToHost));
    return MVS_OK;
}

extern "C" int mvs_cells_fill(mvs_ctx* ctx, const void* records, int64_t n, void* stream) {
    if (!ctx || !ctx->d_cells) { mvs_set_error("cell table not initialised"); return MVS_ERR_STATE; }
    if (n < 0 || (n > 0 && !records)) { mvs_set_error("mvs_cells_fill: bad argument"); return MVS_ERR_ARG; }
    if (n == 0) return MVS_OK;
    MVS_CUDA_CHECK(cudaSetDevice(ctx->device));
    cells_fill_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        (const uint8_t*)records, n, rec_bytes_of(ctx), ctx->V, ctx->cell_size, ctx->wc, ctx->hc, ctx->d_cells);
    ctx->launches++;
    MVS_CUDA_CHECK(cudaGetLastError());
    return MVS_OK;
}

static int ensure_candidates(mvs_ctx* ctx, int64_t M) {
    const size_t mw = (size_t)((ctx->V + 63) / 64);
    int rc;
    if ((rc = mvs_ensure((void**)&ctx->cand_slot, &ctx->cand_cap[0], sizeof(int64_t) * M, "candidates")) != MVS_OK) return rc;
    if ((rc = mvs_ensure((void**)&ctx->cand_parent, &ctx->cand_cap[1], sizeof(int64_t) * M, "candidates")) != MVS_OK) return rc;
    if ((rc = mvs_ensure((void**)&ctx->cand_c, &ctx->cand_cap[2], sizeof(double) * 3 * M, "candidates")) != MVS_OK) return rc;
    if ((rc = mvs_ensure((void**)&ctx->cand_n, &ctx->cand_cap[3], sizeof(double) * 3 * M, "candidates")) != MVS_OK) return rc;
    if ((rc = mvs_ensure((void**)&ctx->cand_ref, &ctx->cand_cap[4], sizeof(int32_t) * M, "candidates")) != MVS_OK) return rc;
    if ((rc = mvs_ensure((void**)&ctx->cand_px, &ctx->cand_cap[5], sizeof(int32_t) * 2 * M, "candidates")) != MVS_OK) return rc;
    if ((rc = mvs_ensure((void**)&ctx->cand_vis, &ctx->cand_cap[6], sizeof(uint64_t) * mw * M, "scores")) != MVS_OK) return rc;
    if ((rc = mvs_ensure((void**)&ctx->cand_avg, &ctx->cand_cap[7], sizeof(double) * M, "scores")) != MVS_OK) return rc;
    if ((rc = mvs_ensure((void**)&ctx->cand_count, &ctx->cand_cap[8], sizeof(int32_t) * M, "scores")) != MVS_OK) return rc;
    if ((rc = mvs_ensure((void**)&ctx->cand_xy, &ctx->cand_cap[9], sizeof(double) * 2 * M, "scores")) != MVS_OK) return rc;
    if ((rc = mvs_ensure((void**)&ctx->cand_gate, &ctx->cand_cap[10], M, "gate")) != MVS_OK) return rc;
    return MVS_OK;
}

extern "C" int mvs_round_generate(mvs_ctx* ctx, const void* frontier, int64_t F, int64_t* n_candidates, void* stream) {
    if (!ctx || !ctx->d_cells) { mvs_set_error("mvs_round_generate: cell table not initialised"); return MVS_ERR_STATE; }
    if (F < 0 || !n_candidates || (F > 0 && !frontier)) { mvs_set_error("mvs_round_generate: bad argument"); return MVS_ERR_ARG; }
    MVS_CUDA_CHECK(cudaSetDevice(ctx->device));
    cudaStream_t s = (cudaStream_t)stream;
    ctx->n_cand = 0;
    *n_candidates = 0;
    if (F == 0) return MVS_OK;
    if (F * (int64_t)ctx->V * 4 >= (1ll << 40)) { mvs_set_error("frontier too large for the slot encoding"); return MVS_ERR_ARG; }
    int rc;
    if ((rc = mvs_ensure((void**)&ctx->d_counts, &ctx->counts_bytes, sizeof(int32_t) * F, "slot counts")) != MVS_OK) return rc;
    if ((rc = mvs_ensure((void**)&ctx->d_scan, &ctx->scan_bytes, sizeof(int64_t) * ((F + 1023) / 1024 + 2), "scan")) != MVS_OK) return rc;
    ctx->epoch++;
    const int rb = rec_bytes_of(ctx);
    const unsigned blocks = (unsigned)((F + 7) / 8);                 // one warp per frontier patch
    int64_t* d_total = ctx->d_scan + (F + 1023) / 1024;
    expand_slots<0><<<blocks, 256, 0, s>>>((const uint8_t*)frontier, F, rb, ctx->V, ctx->cell_size, ctx->wc, ctx->hc, ctx->d_cells,
                                          ctx->d_claim, ctx->epoch, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr,
                                          nullptr, nullptr);
    expand_slots<1><<<blocks, 256, 0, s>>>((const uint8_t*)frontier, F, rb, ctx->V, ctx->cell_size, ctx->wc, ctx->hc, ctx->d_cells,
                                          ctx->d_claim, ctx->epoch, ctx->d_counts, nullptr, nullptr, nullptr, nullptr, nullptr,
                                          nullptr, nullptr);
    ctx->launches += 2;
    MVS_CUDA_CHECK(cudaGetLastError());
    if ((rc = mvs_exclusive_scan_i32(ctx->d_counts, F, ctx->d_scan, d_total, s)) != MVS_OK) return rc;
    ctx->launches += mvs_scan_launches(F);
    int64_t M = 0;
    MVS_CUDA_CHECK(cudaMemcpyAsync(&M, d_total, sizeof(int64_t), cudaMemcpyDeviceToHost, s));
    MVS_CUDA_CHECK(cudaStreamSynchronize(s));
    if (M >= (1ll << 31)) { mvs_set_error("more than 2^31 candidates in one round"); return MVS_ERR_ARG; }
    if (M > 0) {
        if ((rc = ensure_candidates(ctx, M)) != MVS_OK) return rc;
        expand_slots<2><<<blocks, 256, 0, s>>>((const uint8_t*)frontier, F, rb, ctx->V, ctx->cell_size, ctx->wc, ctx->hc,
                                              ctx->d_cells, ctx->d_claim, ctx->epoch, ctx->d_counts, ctx->d_geom, ctx->cand_slot,
                                              ctx->cand_parent, ctx->cand_c, ctx->cand_n, ctx->cand_ref, ctx->cand_px);
        expand_geometry<<<(unsigned)((M + 127) / 128), 128, 0, s>>>((const uint8_t*)frontier, rb, ctx->V, ctx->cell_size, M,
                                                                   ctx->cand_slot, ctx->d_geom, ctx->cand_c, ctx->cand_n,
                                                                   ctx->cand_ref, ctx->cand_px);
        ctx->launches += 2;
        MVS_CUDA_CHECK(cudaGetLastError());
    }
    ctx->n_cand = M;
    *n_candidates = M;
    return MVS_OK;
}

// shared body of mvs_round_score / mvs_round_score_p2p: score the shard and evaluate the gate
static int round_score_shard(mvs_ctx* ctx, const void* frontier, int64_t begin, int64_t end, double min_ncc, int wid,
                             double scale, cudaStream_t s, bool gate_done = false) {
    const int64_t n = end - begin;
    if (n == 0) return MVS_OK;
    const size_t mw = (size_t)((ctx->V + 63) / 64);
    int rc = mvs_launch_score_refexact(ctx, n, ctx->cand_c + 3 * begin, ctx->cand_ref + begin, min_ncc, wid,
                                       ctx->cand_vis + mw * begin, ctx->cand_avg + begin, ctx->cand_count + begin,
                                       ctx->cand_xy + 2 * begin, nullptr, s);
    if (rc != MVS_OK) return rc;
    if (gate_done) return MVS_OK;                              // mvs_expand_run: expand_geometry already wrote the gate
    expand_gate<<<(unsigned)((n + 255) / 256), 256, 0, s>>>((const uint8_t*)frontier, rec_bytes_of(ctx), begin, end,
                                                           ctx->cand_parent, ctx->cand_c, ctx->cand_n, 0.05 / scale,
                                                           ctx->cand_gate);
    ctx->launches++;
    MVS_CUDA_CHECK(cudaGetLastError());
    return MVS_OK;
}

extern "C" int mvs_round_score(mvs_ctx* ctx, const void* frontier, int64_t begin, int64_t end, double min_ncc, int wid,
                               int bound, double scale, void* records, int64_t capacity, int64_t* n_out, void* stream) {
    if (!ctx || !ctx->d_cells) { mvs_set_error("mvs_round_score: cell table not initialised"); return MVS_ERR_STATE; }
    if (begin < 0 || end < begin || end > ctx->n_cand || !n_out || capacity < 0 || (end > begin && (!records || !frontier))) {
        mvs_set_error("mvs_round_score: bad shard [%lld, %lld) of %lld candidates", (long long)begin, (long long)end,
                      (long long)ctx->n_cand);
        return MVS_ERR_ARG;
    }
    MVS_CUDA_CHECK(cudaSetDevice(ctx->device));
    cudaStream_t s = (cudaStream_t)stream;
    const int64_t n = end - begin;
    if (n == 0) {
        MVS_CUDA_CHECK(cudaMemsetAsync(n_out, 0, sizeof(int64_t), s));
        return MVS_OK;
    }
    const size_t mw = (size_t)((ctx->V + 63) / 64);
    int rc = round_score_shard(ctx, frontier, begin, end, min_ncc, wid, scale, s);
    if (rc != MVS_OK) return rc;
    return mvs_launch_compact(ctx, n, 0, ctx->cand_c + 3 * begin, ctx->cand_n + 3 * begin, ctx->cand_ref + begin,
                              ctx->cand_vis + mw * begin, ctx->cand_avg + begin, ctx->cand_count + begin,
                              ctx->cand_xy + 2 * begin, ctx->cand_gate + begin, bound, records, capacity, n_out,
                              ctx->cand_slot + begin, ctx->cand_px + 2 * begin, s);
}

// Phase 2 with the exchange fused in: the passing records of this shard are stored straight into every
// GPU's inbox (mvs_compact_accepted_p2p semantics, index = slot id), an empty shard publishes a zero count.
extern "C" int mvs_round_score_p2p(mvs_ctx* ctx, const void* frontier, int64_t begin, int64_t end, double min_ncc, int wid,
                                   int bound, double scale, void* const* peer_records, int64_t* const* peer_counts,
                                   int rank, int world, int wire, int64_t capacity, void* stream) {
    if (!ctx || !ctx->d_cells) { mvs_set_error("mvs_round_score_p2p: cell table not initialised"); return MVS_ERR_STATE; }
    if (begin < 0 || end < begin || end > ctx->n_cand || capacity < end - begin || !peer_records || !peer_counts ||
        world < 1 || world > MVS_MAX_PEERS || rank < 0 || rank >= world || (end > begin && !frontier) ||
        (wire != MVS_WIRE_FULL && wire != MVS_WIRE_COMPACT)) {
        mvs_set_error("mvs_round_score_p2p: bad shard [%lld, %lld) of %lld candidates, capacity %lld, rank %d of %d",
                      (long long)begin, (long long)end, (long long)ctx->n_cand, (long long)capacity, rank, world);
        return MVS_ERR_ARG;
    }
    MVS_CUDA_CHECK(cudaSetDevice(ctx->device));
    cudaStream_t s = (cudaStream_t)stream;
    const int64_t n = end - begin;
    const size_t mw = (size_t)((ctx->V + 63) / 64);
    int rc = round_score_shard(ctx, frontier, begin, end, min_ncc, wid, scale, s);
    if (rc != MVS_OK) return rc;
    return mvs_launch_compact_p2p(ctx, n, 0, ctx->cand_c + 3 * begin, ctx->cand_n + 3 * begin, ctx->cand_ref + begin,
                                  ctx->cand_vis + mw * begin, ctx->cand_avg + begin, ctx->cand_count + begin,
                                  ctx->cand_xy + 2 * begin, ctx->cand_gate + begin, bound, peer_records, peer_counts, rank,
                                  world, wire, capacity, ctx->cand_slot + begin, ctx->cand_px + 2 * begin, s);
}

extern "C" int mvs_round_commit(mvs_ctx* ctx, const void* records, int64_t n, void* next_frontier, int64_t* n_next,
                                void* stream) {
    if (!ctx || !ctx->d_cells) { mvs_set_error("mvs_round_commit: cell table not initialised"); return MVS_ERR_STATE; }
    if (n < 0 || !n_next || (n > 0 && !records)) { mvs_set_error("mvs_round_commit: bad argument"); return MVS_ERR_ARG; }
    MVS_CUDA_CHECK(cudaSetDevice(ctx->device));
    cudaStream_t s = (cudaStream_t)stream;
    if (n == 0) {
        MVS_CUDA_CHECK(cudaMemsetAsync(n_next, 0, sizeof(int64_t), s));
        return MVS_OK;
    }
    int rc;
    if ((rc = mvs_ensure((void**)&ctx->d_counts, &ctx->counts_bytes, sizeof(int32_t) * n, "commit flags")) != MVS_OK) return rc;
    if ((rc = mvs_ensure((void**)&ctx->d_scan, &ctx->scan_bytes, sizeof(int64_t) * ((n + 1023) / 1024 + 2), "scan")) != MVS_OK) return rc;
    const int rb = rec_bytes_of(ctx);
    const unsigned blocks = (unsigned)((n + 255) / 256);
    commit_flags<<<blocks, 256, 0, s>>>((const uint8_t*)records, n, rb, ctx->d_counts);
    if ((rc = mvs_exclusive_scan_i32(ctx->d_counts, n, ctx->d_scan, n_next, s)) != MVS_OK) return rc;
    commit_apply<<<blocks, 256, 0, s>>>((const uint8_t*)records, n, rb, ctx->d_counts, n_next, (uint8_t*)next_frontier, ctx->V,
                                       ctx->cell_size, ctx->wc, ctx->hc, ctx->d_cells);
    ctx->launches += 2 + mvs_scan_launches(n);
    MVS_CUDA_CHECK(cudaGetLastError());
    return MVS_OK;
}

extern "C" int mvs_round_candidates(mvs_ctx* ctx, int64_t* slot, int64_t* parent, double* c, double* nrm, int32_t* ref) {
    if (!ctx) { mvs_set_error("null context"); return MVS_ERR_ARG; }
    MVS_CUDA_CHECK(cudaSetDevice(ctx->device));
    MVS_CUDA_CHECK(cudaDeviceSynchronize());
    const int64_t M = ctx->n_cand;
    if (M == 0) return MVS_OK;
    if (slot) MVS_CUDA_CHECK(cudaMemcpy(slot, ctx->cand_slot, sizeof(int64_t) * M, cudaMemcpyDeviceToHost));
    if (parent) MVS_CUDA_CHECK(cudaMemcpy(parent, ctx->cand_parent, sizeof(int64_t) * M, cudaMemcpyDeviceToHost));
    if (c) MVS_CUDA_CHECK(cudaMemcpy(c, ctx->cand_c, sizeof(double) * 3 * M, cudaMemcpyDeviceToHost));
    if (nrm) MVS_CUDA_CHECK(cudaMemcpy(nrm, ctx->cand_n, sizeof(double) * 3 * M, cudaMemcpyDeviceToHost));
    if (ref) MVS_CUDA_CHECK(cudaMemcpy(ref, ctx->cand_ref, sizeof(int32_t) * M, cudaMemcpyDeviceToHost));
    return MVS_OK;
}


// ---------------------------------------------------------------------------------------
// The whole expansion loop on the host side of the C ABI (replaces the while-loop of MVS2.py:321-404 for
// ALL rounds in one call): no Python between rounds and ONE host synchronisation per round.
//   round k:  expand_slots<2> (candidates)  ->  score this GPU's shard + gate  ->  publish the minimal wire
//             into every GPU's inbox  ->  [device barrier]  ->  commit from the wire (next frontier appended
//             to the context's output buffer, cells filled)  ->  expand_slots<0,1> + scan of round k+1 run
//             SPECULATIVELY on the device-side frontier size  ->  one read-back of {accepted, next M}.
// Every GPU executes the same sequence on the same candidate list, so all ranks stay in lockstep without
// any host communication.
// ---------------------------------------------------------------------------------------
int mvs_launch_publish(mvs_ctx* ctx, int64_t N, const uint64_t* vis, const double* avg, const int32_t* count,
                       const uint8_t* gate, int bound, void* const* peer_inbox, int rank, int world, int64_t capacity,
                       int parity, cudaStream_t s);
int mvs_launch_commit_wire(mvs_ctx* ctx, const void* inbox_local, int world, int64_t capacity, int parity, int64_t M,
                           void* next_frontier, int64_t* d_n_next, cudaStream_t s);
int64_t mvs_exchange_sub_bytes(const mvs_ctx* ctx, int64_t capacity);

// Everything the host needs to know about a round, written by ONE tiny kernel straight into pinned (device-mapped)
// host memory: {accepted, candidates of the next round, kept per rank, barrier error flag}.  Replaces three small
// device->host copies (and a fourth for the barrier flag), each of which cost a few microseconds of a ~160 us round.
__global__ void __launch_bounds__(32) round_report(const int64_t* __restrict__ d_n_next, const int64_t* __restrict__ d_M_next,
                                                   const uint8_t* __restrict__ inbox_half, int64_t region_bytes, int64_t sub_bytes,
                                                   int world, const int* __restrict__ barrier_err,
                                                   volatile int64_t* __restrict__ host_out) {
    const int t = threadIdx.x;
    if (t == 0) host_out[0] = *d_n_next;
    if (t == 1) host_out[1] = d_M_next ? *d_M_next : 0;
    if (t >= 2 && t < 2 + world) {
        // kept candidates of rank t - 2: one header, or one per position range of a partitioned publish
        const uint8_t* reg = inbox_half + (int64_t)(t - 2) * region_bytes;
        const int used = (int)(reinterpret_cast<const int64_t*>(reg)[1] >> 48);
        int64_t kept = 0;
        for (int k = 0; k < (used > 1 ? used : 1); ++k) kept += *reinterpret_cast<const int64_t*>(reg + (int64_t)k * sub_bytes);
        host_out[t] = kept;
    }
    if (t == 2 + world) host_out[t] = barrier_err ? (int64_t)*barrier_err : 0;
    __threadfence_system();
}

// grow a device buffer, KEEPING its first `keep` bytes
static int ensure_keep(void** p, size_t* cap, size_t bytes, size_t keep, cudaStream_t s, const char* what) {
    if (bytes <= *cap && *p) return MVS_OK;
    void* q = nullptr;
    const size_t want = bytes * 2 + 4096;
    if (cudaMalloc(&q, want) != cudaSuccess) {
        cudaGetLastError();
        mvs_set_error("device allocation of %zu bytes for %s failed", want, what);
        return MVS_ERR_NOMEM;
    }
    if (*p && keep) {
        MVS_CUDA_CHECK(cudaMemcpyAsync(q, *p, keep, cudaMemcpyDeviceToDevice, s));
        MVS_CUDA_CHECK(cudaStreamSynchronize(s));
    }
    if (*p) cudaFree(*p);
    *p = q;
    *cap = want;
    return MVS_OK;
}

// claim + count passes of candidate generation for a frontier whose size is host-known (d_F == nullptr) or
// still on the device (grid over the upper bound F_upper); leaves the candidate count in *d_total
static int generate_count(mvs_ctx* ctx, const void* frontier, int64_t F_upper, const int64_t* d_F, int64_t F_clamp,
                          int64_t** d_total_out, cudaStream_t s) {
    int rc;
    if ((rc = mvs_ensure((void**)&ctx->d_counts, &ctx->counts_bytes, sizeof(int32_t) * F_upper, "slot counts")) != MVS_OK) return rc;
    if ((rc = mvs_ensure((void**)&ctx->d_scan, &ctx->scan_bytes, sizeof(int64_t) * ((F_upper + 1023) / 1024 + 2), "scan")) != MVS_OK) return rc;
    if ((rc = mvs_ensure((void**)&ctx->d_live, &ctx->live_bytes, (size_t)F_upper * ctx->V, "surviving slots")) != MVS_OK) return rc;
    ctx->epoch++;
    const int rb = rec_bytes_of(ctx);
    const unsigned blocks = (unsigned)((F_upper + 7) / 8);           // one warp per frontier patch
    int64_t* d_total = ctx->d_scan + (F_upper + 1023) / 1024;
    expand_slots<0><<<blocks, 256, 0, s>>>((const uint8_t*)frontier, F_upper, rb, ctx->V, ctx->cell_size, ctx->wc, ctx->hc,
                                          ctx->d_cells, ctx->d_claim, ctx->epoch, nullptr, nullptr, nullptr, nullptr, nullptr,
                                          nullptr, nullptr, nullptr, d_F, F_clamp);
    expand_slots<1><<<blocks, 256, 0, s>>>((const uint8_t*)frontier, F_upper, rb, ctx->V, ctx->cell_size, ctx->wc, ctx->hc,
                                          ctx->d_cells, ctx->d_claim, ctx->epoch, ctx->d_counts, nullptr, nullptr, nullptr,
                                          nullptr, nullptr, nullptr, nullptr, d_F, F_clamp, ctx->d_live, ctx->V);
    ctx->launches += 2;
    MVS_CUDA_CHECK(cudaGetLastError());
    if ((rc = mvs_exclusive_scan_i32(ctx->d_counts, F_upper, ctx->d_scan, d_total, s)) != MVS_OK) return rc;
    ctx->launches += mvs_scan_launches(F_upper);
    *d_total_out = d_total;
    return MVS_OK;
}

extern "C" int mvs_expand_run(mvs_ctx* ctx, const void* seeds, int64_t n_seeds, const mvs_expand_params* prm,
                              mvs_round_stat* stats, int max_stats, int* n_rounds, int64_t* n_patches, void* stream) {
    if (!ctx || !ctx->d_cells) { mvs_set_error("mvs_expand_run: cell table not initialised (mvs_cells_init)"); return MVS_ERR_STATE; }
    if (!prm || n_seeds < 0 || (n_seeds > 0 && !seeds) || !n_rounds || !n_patches || prm->wid < 1 || prm->wid > 7 ||
        prm->world < 1 || prm->world > MVS_MAX_PEERS || prm->rank < 0 || prm->rank >= prm->world ||
        (prm->world > 1 && (!prm->peer_inbox || !prm->peer_flags || prm->capacity < 1))) {
        mvs_set_error("mvs_expand_run: bad argument (need params, 1 <= wid <= 7, 0 <= rank < world <= %d and, for world > 1, "
                      "the inbox / flag tables and a capacity)", MVS_MAX_PEERS);
        return MVS_ERR_ARG;
    }
    MVS_CUDA_CHECK(cudaSetDevice(ctx->device));
    cudaStream_t s = (cudaStream_t)stream;
    const int world = prm->world, rank = prm->rank;
    const int rb = rec_bytes_of(ctx);
    const size_t mw = (size_t)((ctx->V + 63) / 64);
    const int64_t max_rounds = prm->max_rounds < 0 ? (1ll << 40) : prm->max_rounds;
    const int64_t max_iter = prm->max_iterations < 0 ? (1ll << 60) : prm->max_iterations;
    const int64_t max_patches = prm->max_patches < 0 ? (1ll << 60) : prm->max_patches;
    *n_rounds = 0;
    *n_patches = 0;
    ctx->n_out = 0;
    int rc;
    if (!ctx->h_pinned) MVS_CUDA_CHECK(cudaMallocHost(&ctx->h_pinned, 256));
    int64_t* h_back = (int64_t*)ctx->h_pinned;                 // {n_next, next M} + per-rank kept counts
    if (prm->timing && !ctx->ev_round[0]) {
        MVS_CUDA_CHECK(cudaEventCreate(&ctx->ev_round[0]));
        MVS_CUDA_CHECK(cudaEventCreate(&ctx->ev_round[1]));
    }
    int64_t* d_n_next = nullptr;
    if ((rc = mvs_ensure((void**)&ctx->d_round_n, &ctx->round_n_bytes, 64, "round counters")) != MVS_OK) return rc;
    d_n_next = ctx->d_round_n;

    int64_t F = n_seeds, expanded = 0, T = 0;                  // T = accepted records so far (offset of the next frontier)
    const uint8_t* frontier = (const uint8_t*)seeds;
    bool frontier_in_out = false;
    int64_t frontier_off = 0;
    if (max_iter < F) F = max_iter;
    if (F == 0 || max_rounds == 0) return MVS_OK;
    if (F * (int64_t)ctx->V * 4 >= (1ll << 40)) { mvs_set_error("frontier too large for the slot encoding"); return MVS_ERR_ARG; }
    int64_t* d_total = nullptr;
    if ((rc = generate_count(ctx, frontier, F, nullptr, 0, &d_total, s)) != MVS_OK) return rc;
    int64_t M = 0;
    MVS_CUDA_CHECK(cudaMemcpyAsync(&h_back[1], d_total, sizeof(int64_t), cudaMemcpyDeviceToHost, s));
    MVS_CUDA_CHECK(cudaStreamSynchronize(s));
    M = h_back[1];

    for (int64_t k = 0; k < max_rounds; ++k) {
        if (F == 0) break;
        if (M >= (1ll << 31)) { mvs_set_error("more than 2^31 candidates in one round"); return MVS_ERR_ARG; }
        expanded += F;
        if (prm->timing) MVS_CUDA_CHECK(cudaEventRecord(ctx->ev_round[0], s));
        int64_t n_next = 0, M_next = 0, passed = 0;
        int64_t F_next_upper = 0;
        if (M > 0) {
            const int64_t shard_cap = world > 1 ? prm->capacity : M;
            if (world > 1 && (M + world - 1) / world > shard_cap) {
                mvs_set_error("mvs_expand_run: round %lld has %lld candidates, more than the inbox capacity of %lld per GPU x %d GPUs",
                              (long long)k, (long long)M, (long long)shard_cap, world);
                return MVS_ERR_STATE;
            }
            if ((rc = ensure_candidates(ctx, M)) != MVS_OK) return rc;
            if ((rc = ensure_keep((void**)&ctx->d_out, &ctx->out_bytes, (size_t)(T + M) * rb, (size_t)T * rb, s, "accepted patches")) != MVS_OK)
                return rc;
            if (frontier_in_out) frontier = ctx->d_out + frontier_off * rb;   // the buffer may have moved
            // ---- candidates of this round
            expand_slots<2><<<(unsigned)((F + 7) / 8), 256, 0, s>>>(frontier, F, rb, ctx->V, ctx->cell_size, ctx->wc, ctx->hc,
                                                                        ctx->d_cells, ctx->d_claim, ctx->epoch, ctx->d_counts,
                                                                        ctx->d_geom, ctx->cand_slot, ctx->cand_parent, ctx->cand_c,
                                                                        ctx->cand_n, ctx->cand_ref, ctx->cand_px, nullptr, 0, ctx->d_live,
                                                                        ctx->V);
            expand_geometry<<<(unsigned)((M + 127) / 128), 128, 0, s>>>(frontier, rb, ctx->V, ctx->cell_size, M, ctx->cand_slot,
                                                                       ctx->d_geom, ctx->cand_c, ctx->cand_n, ctx->cand_ref,
                                                                       ctx->cand_px, ctx->cand_gate, 0.05 / prm->scale);
            ctx->launches += 2;
            MVS_CUDA_CHECK(cudaGetLastError());
            ctx->n_cand = M;
            // ---- this GPU's shard: score, gate, publish
            const int64_t begin = (M * rank) / world, end = (M * (rank + 1)) / world;
            void* local_tab[1];
            void* const* inbox_tab = prm->peer_inbox;
            const void* inbox_local;
            if (world == 1) {
                const size_t need = (size_t)mvs_exchange_bytes(ctx, 1, shard_cap);
                if ((rc = mvs_ensure((void**)&ctx->d_inbox, &ctx->inbox_bytes, need, "local inbox")) != MVS_OK) return rc;
                local_tab[0] = ctx->d_inbox;
                inbox_tab = local_tab;
                inbox_local = ctx->d_inbox;
            } else {
                inbox_local = prm->peer_inbox[rank];
            }
            const int parity = (int)(k & 1);
            // score + publish (expand_geometry already wrote the gate); large shards with mvs_exchange_set_parts(P > 1):
            // K1 in P launches, the publish of one position range under the scoring of the next
            if ((rc = mvs_launch_score_publish(ctx, end - begin, ctx->cand_c + 3 * begin, ctx->cand_ref + begin, prm->min_ncc, prm->wid,
                                               ctx->cand_vis + mw * begin, ctx->cand_avg + begin, ctx->cand_count + begin,
                                               ctx->cand_xy + 2 * begin, ctx->cand_gate + begin, prm->bound, inbox_tab, rank, world,
                                               shard_cap, parity, s)) != MVS_OK)
                return rc;
            if (world > 1 && (rc = mvs_p2p_barrier(ctx, prm->peer_flags, rank, world, s)) != MVS_OK) return rc;
            // ---- commit (identical on every GPU): next frontier appended to the output buffer
            uint8_t* next = ctx->d_out + T * rb;
            if ((rc = mvs_launch_commit_wire(ctx, inbox_local, world, shard_cap, parity, M, next, d_n_next, s)) != MVS_OK) return rc;
            // ---- speculative claim + count passes of the NEXT round on the device-side frontier size
            const int64_t remaining = max_iter - expanded;
            const int64_t* d_M_next = nullptr;
            if (k + 1 < max_rounds && remaining > 0) {
                F_next_upper = M;                              // accepted <= candidates
                if ((rc = generate_count(ctx, next, F_next_upper, d_n_next, remaining, &d_total, s)) != MVS_OK) return rc;
                d_M_next = d_total;
            }
            // ---- one read-back per round
            const int64_t region_bytes = mvs_exchange_bytes(ctx, world, shard_cap) / (2 * world);
            round_report<<<1, 32, 0, s>>>(d_n_next, d_M_next, (const uint8_t*)inbox_local + (size_t)parity * world * region_bytes,
                                          region_bytes, mvs_exchange_sub_bytes(ctx, shard_cap), world,
                                          (world > 1 && ctx->d_barrier_state) ? (const int*)((const uint8_t*)ctx->d_barrier_state + 8) : nullptr,
                                          h_back);
            ctx->launches++;
            MVS_CUDA_CHECK(cudaGetLastError());
            if (prm->timing) MVS_CUDA_CHECK(cudaEventRecord(ctx->ev_round[1], s));
            MVS_CUDA_CHECK(cudaStreamSynchronize(s));
            n_next = h_back[0];
            M_next = h_back[1];
            for (int r = 0; r < world; ++r) passed += h_back[2 + r];
            if (world > 1 && h_back[2 + world] != 0) {
                mvs_set_error("mvs_expand_run: a peer GPU did not reach the round barrier (round %lld)", (long long)k);
                return MVS_ERR_STATE;
            }
        } else if (prm->timing) {
            MVS_CUDA_CHECK(cudaEventRecord(ctx->ev_round[1], s));
            MVS_CUDA_CHECK(cudaStreamSynchronize(s));
        }
        if (stats && *n_rounds < max_stats) {
            mvs_round_stat& st = stats[*n_rounds];
            st.frontier = F;
            st.candidates = M;
            st.passed = passed;
            st.accepted = n_next;
            st.ms = 0.f;
            if (prm->timing) cudaEventElapsedTime(&st.ms, ctx->ev_round[0], ctx->ev_round[1]);
        }
        ++*n_rounds;
        frontier_in_out = true;
        frontier_off = T;
        frontier = ctx->d_out + T * rb;
        T += n_next;
        ctx->n_out = T;
        const int64_t remaining = max_iter - expanded;
        F = n_next < remaining ? n_next : remaining;
        M = M_next;
        if (T >= max_patches) break;
    }
    *n_patches = T;
    return MVS_OK;
}

extern "C" int mvs_expand_result(mvs_ctx* ctx, void* records, int64_t offset, int64_t n, int on_device, void* stream) {
    if (!ctx) { mvs_set_error("mvs_expand_result: null context"); return MVS_ERR_ARG; }
    if (offset < 0 || n < 0 || offset + n > ctx->n_out || (n > 0 && !records)) {
        mvs_set_error("mvs_expand_result: records [%lld, %lld) requested, %lld available", (long long)offset, (long long)(offset + n),
                      (long long)ctx->n_out);
        return MVS_ERR_ARG;
    }
    if (n == 0) return MVS_OK;
    MVS_CUDA_CHECK(cudaSetDevice(ctx->device));
    const size_t rb = (size_t)rec_bytes_of(ctx);
    MVS_CUDA_CHECK(cudaMemcpyAsync(records, ctx->d_out + offset * rb, n * rb, on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost,
                                   (cudaStream_t)stream));
    if (!on_device) MVS_CUDA_CHECK(cudaStreamSynchronize((cudaStream_t)stream));
    return MVS_OK;
}
