// CellTable.filter_out_outlier (MVS2.py:132-158) -- the PMVS-style filtering pass the reference ships but
// leaves disabled as "very very slow" (MVS2.py:280-281): O(#Q^2) per cell in Python.
//
// Reference semantics, restated for patches whose visible-set entries share one (x, y) (every patch the
// reference's scorer creates, MVS2.py:74): a patch p lives in the lists Q[(v, ci, cj)] of every view v of its
// visible set, (ci, cj) = cell of its (x, y), with multiplicity len(p.V) (MVS2.py:106-107).  The scan visits
// cells in (view, ci, cj) order; a non-vacant cell computes
//     threshold = mean over the list entries of (1 - avg_ncc)              (sequential fp64 sum, MVS2.py:140-143)
//     outliers  = { p2 : exists p1 != p2 in the list with not is_patch_neighbor(p1, p2, 0.2)
//                        and len(p2.V) * p2.avg_ncc < threshold }          (MVS2.py:145-149)
// and removes every outlier from ALL its lists (MVS2.py:150-157) before the scan goes on.  A removal at
// (v, ci, cj) only touches lists (v', ci, cj) of the SAME cell position, so cell positions are independent
// and the sequential dependency is the ascending view order inside one position: one thread per (ci, cj)
// column walks its views.  Lists keep insertion order (ascending patch index).
// Declared divergence: a non-vacant cell whose list has become empty makes the reference raise
// ZeroDivisionError (MVS2.py:143); here it is skipped and counted.
#include "scan.cuh"

__device__ __forceinline__ double fm(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double fa(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double fs(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ double fdot(const double* a, double b0, double b1, double b2) {
    return fa(fa(fm(a[0], b0), fm(a[1], b1)), fm(a[2], b2));
}

__device__ __forceinline__ const mvs_patch_record* frec(const uint8_t* base, int64_t i, int rb) {
    return reinterpret_cast<const mvs_patch_record*>(base + i * rb);
}

__device__ __forceinline__ bool sees(const mvs_patch_record* r, int v) {
    return (reinterpret_cast<const uint64_t*>(r + 1)[v >> 6] >> (v & 63)) & 1ull;
}

// column id (ci * hc + cj) of every patch, -1 when its (x, y) is outside the table; histogram
__global__ void __launch_bounds__(256) filter_keys(const uint8_t* __restrict__ recs, int64_t n, int rb, int cs, int wc, int hc,
                                                   int32_t* __restrict__ key, int32_t* __restrict__ hist) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const mvs_patch_record* r = frec(recs, i, rb);
    int k = -1;
    if (isfinite(r->xy[0]) && isfinite(r->xy[1])) {
        const double fi = floor(__ddiv_rn(r->xy[0], (double)cs)), fj = floor(__ddiv_rn(r->xy[1], (double)cs));
        if (fi >= 0 && fi < wc && fj >= 0 && fj < hc) k = (int)fi * hc + (int)fj;
    }
    key[i] = k;
    if (k >= 0) atomicAdd(hist + k, 1);
}

__global__ void __launch_bounds__(256) filter_fill(int64_t n, const int32_t* __restrict__ key, const int32_t* __restrict__ start,
                                                   int32_t* __restrict__ cursor, int32_t* __restrict__ list) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n || key[i] < 0) return;
    list[start[key[i]] + atomicAdd(cursor + key[i], 1)] = (int32_t)i;
}

__global__ void __launch_bounds__(64)
    filter_columns(const uint8_t* __restrict__ recs, int rb, int V, int wc, int hc, const uint8_t* __restrict__ cells,
                   const int32_t* __restrict__ start, const int32_t* __restrict__ cursor, int32_t* __restrict__ list,
                   uint8_t* __restrict__ removed, uint8_t* __restrict__ mark, unsigned long long* __restrict__ counters) {
    const int col = blockIdx.x * blockDim.x + threadIdx.x;
    if (col >= wc * hc) return;
    const int L = cursor[col];
    if (L == 0) {
        // cells that are non-vacant without any patch (the reference would divide by zero, MVS2.py:143)
        unsigned long long e = 0;
        for (int v = 0; v < V; ++v) e += cells[(int64_t)v * wc * hc + col] == 0;
        if (e) atomicAdd(counters + 1, e);
        return;
    }
    int32_t* mine = list + start[col];
    for (int a = 1; a < L; ++a) {                          // insertion order = ascending patch index
        const int32_t x = mine[a];
        int b = a - 1;
        while (b >= 0 && mine[b] > x) {
            mine[b + 1] = mine[b];
            --b;
        }
        mine[b + 1] = x;
    }
    unsigned long long n_removed = 0, n_empty = 0;
    for (int v = 0; v < V; ++v) {
        if (cells[(int64_t)v * wc * hc + col] != 0) continue;            // vacant cell: not visited (MVS2.py:136)
        double thr = 0.0;
        long long entries = 0;
        for (int a = 0; a < L; ++a) {
            const int32_t p = mine[a];
            if (removed[p]) continue;
            const mvs_patch_record* r = frec(recs, p, rb);
            if (!sees(r, v)) continue;
            const double term = fs(1.0, r->avg);
            for (int m = 0; m < r->count; ++m) thr = fa(thr, term);       // one list entry per element of p.V
            entries += r->count;
        }
        if (entries == 0) {
            ++n_empty;
            continue;
        }
        thr = __ddiv_rn(thr, (double)entries);
        bool any = false;
        for (int a = 0; a < L; ++a) {
            const int32_t p2 = mine[a];
            if (removed[p2]) continue;
            const mvs_patch_record* r2 = frec(recs, p2, rb);
            if (!sees(r2, v)) continue;
            if (!(fm((double)r2->count, r2->avg) < thr)) continue;
            for (int b = 0; b < L; ++b) {
                const int32_t p1 = mine[b];
                if (p1 == p2 || removed[p1]) continue;
                const mvs_patch_record* r1 = frec(recs, p1, rb);
                if (!sees(r1, v)) continue;
                // is_patch_neighbor(p1, p2) (MVS2.py:298-299, threshold 0.2)
                const double d[3] = {fs(r1->c[0], r2->c[0]), fs(r1->c[1], r2->c[1]), fs(r1->c[2], r2->c[2])};
                const double q = fabs(fa(fdot(d, r1->n[0], r1->n[1], r1->n[2]), fdot(d, r2->n[0], r2->n[1], r2->n[2])));
                if (!(q < 0.2)) {
                    mark[p2] = 1;
                    any = true;
                    break;
                }
            }
        }
        if (any) {
            for (int a = 0; a < L; ++a) {
                const int32_t p = mine[a];
                if (mark[p] && !removed[p]) {
                    removed[p] = 1;
                    ++n_removed;
                }
            }
        }
    }
    if (n_removed) atomicAdd(counters, n_removed);
    if (n_empty) atomicAdd(counters + 1, n_empty);
}

extern "C" int mvs_cells_filter(mvs_ctx* ctx, const void* records, int64_t n, uint8_t* removed, int64_t* counts_host,
                                void* stream) {
    if (!ctx || !ctx->d_cells) { mvs_set_error("mvs_cells_filter: cell table not initialised"); return MVS_ERR_STATE; }
    if (n < 0 || n >= (1ll << 31) || !counts_host || (n > 0 && (!records || !removed))) {
        mvs_set_error("mvs_cells_filter: bad argument");
        return MVS_ERR_ARG;
    }
    counts_host[0] = counts_host[1] = 0;
    MVS_CUDA_CHECK(cudaSetDevice(ctx->device));
    cudaStream_t s = (cudaStream_t)stream;
    const int ncol = ctx->wc * ctx->hc;
    const int rb = (int)(sizeof(mvs_patch_record) + 8 * ((ctx->V + 63) / 64));
    // scan scratch | total | counters [2] | key [n] | list [n] | start [ncol] | cursor [ncol] | mark [n]
    const size_t n_scan = (size_t)(ncol + 1023) / 1024 + 2;
    const size_t bytes = sizeof(int64_t) * (n_scan + 1) + 16 + sizeof(int32_t) * (2 * (size_t)n + 2 * (size_t)ncol) + (size_t)n + 64;
    uint8_t* w = nullptr;
    if (cudaMalloc(&w, bytes) != cudaSuccess) {
        cudaGetLastError();
        mvs_set_error("mvs_cells_filter: device allocation of %zu bytes failed", bytes);
        return MVS_ERR_NOMEM;
    }
    int64_t* d_scan = (int64_t*)w;                          // 8-byte aligned parts first
    int64_t* d_total = d_scan + n_scan;
    unsigned long long* d_counters = (unsigned long long*)(d_total + 1);
    int32_t* d_key = (int32_t*)(d_counters + 2);
    int32_t* d_list = d_key + n;
    int32_t* d_start = d_list + n;
    int32_t* d_cursor = d_start + ncol;
    uint8_t* d_mark = (uint8_t*)(d_cursor + ncol);
    int rc = MVS_OK;
    const unsigned nb = (unsigned)((n + 255) / 256);
    if (cudaMemsetAsync(w, 0, bytes, s) != cudaSuccess || (n > 0 && cudaMemsetAsync(removed, 0, (size_t)n, s) != cudaSuccess)) {
        mvs_set_error("mvs_cells_filter: memset failed: %s", cudaGetErrorString(cudaGetLastError()));
        cudaFree(w);
        return MVS_ERR_CUDA;
    }
    if (n > 0) {
        filter_keys<<<nb, 256, 0, s>>>((const uint8_t*)records, n, rb, ctx->cell_size, ctx->wc, ctx->hc, d_key, d_start);
        rc = mvs_exclusive_scan_i32(d_start, ncol, d_scan, d_total, s);
        if (rc == MVS_OK) {
            filter_fill<<<nb, 256, 0, s>>>(n, d_key, d_start, d_cursor, d_list);
            ctx->launches += 2 + mvs_scan_launches(ncol);
        }
    }
    if (rc == MVS_OK) {
        filter_columns<<<(unsigned)((ncol + 63) / 64), 64, 0, s>>>((const uint8_t*)records, rb, ctx->V, ctx->wc, ctx->hc, ctx->d_cells,
                                                                  d_start, d_cursor, d_list, removed, d_mark, d_counters);
        ctx->launches++;
        if (cudaMemcpyAsync(counts_host, d_counters, 2 * sizeof(int64_t), cudaMemcpyDeviceToHost, s) != cudaSuccess ||
            cudaStreamSynchronize(s) != cudaSuccess) {
            mvs_set_error("mvs_cells_filter failed: %s", cudaGetErrorString(cudaGetLastError()));
            rc = MVS_ERR_CUDA;
        }
    }
    cudaFree(w);
    return rc;
}
