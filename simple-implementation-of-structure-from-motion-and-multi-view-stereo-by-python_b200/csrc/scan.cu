#include "scan.cuh"

#define FULL 0xffffffffu

int mvs_ensure(void** p, size_t* cap, size_t bytes, const char* what) {
    if (bytes <= *cap && *p) return MVS_OK;
    if (*p) cudaFree(*p);
    *p = nullptr;
    *cap = 0;
    const size_t want = bytes + bytes / 2 + 4096;
    if (cudaMalloc(p, want) != cudaSuccess) {
        cudaGetLastError();
        mvs_set_error("device allocation of %zu bytes for %s failed", want, what);
        return MVS_ERR_NOMEM;
    }
    *cap = want;
    return MVS_OK;
}

// block-wide inclusive scan of one int per thread (1024 threads); returns inclusive value,
// *block_total valid in all threads after the call
__device__ __forceinline__ int block_scan_incl(int v, int* wsum /*[32] shared*/, int* block_total) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    int x = v;
#pragma unroll
    for (int s = 1; s < 32; s <<= 1) {
        const int y = __shfl_up_sync(FULL, x, s);
        if (lane >= s) x += y;
    }
    if (lane == 31) wsum[w] = x;
    __syncthreads();
    if (w == 0) {
        int ws = wsum[lane];
#pragma unroll
        for (int s = 1; s < 32; s <<= 1) {
            const int y = __shfl_up_sync(FULL, ws, s);
            if (lane >= s) ws += y;
        }
        wsum[lane] = ws;
    }
    __syncthreads();
    const int res = x + (w > 0 ? wsum[w - 1] : 0);
    *block_total = wsum[31];
    __syncthreads();
    return res;
}

__global__ void __launch_bounds__(1024) scan_tiles(int32_t* __restrict__ a, int64_t n, int64_t* __restrict__ tile_tot) {
    __shared__ int wsum[32];
    const int64_t i = (int64_t)blockIdx.x * 1024 + threadIdx.x;
    const int v = (i < n) ? a[i] : 0;
    int tot;
    const int incl = block_scan_incl(v, wsum, &tot);
    if (i < n) a[i] = incl - v;
    if (threadIdx.x == 0) tile_tot[blockIdx.x] = tot;
}

__global__ void __launch_bounds__(1024) scan_tile_totals(int64_t* __restrict__ tile_tot, int T, int64_t* __restrict__ total) {
    // single CTA, sequential over chunks of 1024 tiles; int64 carry
    __shared__ int wsum[32];
    __shared__ int64_t carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < T; base += 1024) {
        const int i = base + threadIdx.x;
        const int v = (i < T) ? (int)tile_tot[i] : 0;
        int tot;
        const int incl = block_scan_incl(v, wsum, &tot);
        const int64_t c0 = carry;
        if (i < T) tile_tot[i] = c0 + incl - v;
        __syncthreads();
        if (threadIdx.x == 0) carry = c0 + tot;
        __syncthreads();
    }
    if (threadIdx.x == 0) *total = carry;
}

// offsets become int64-safe only through the tile base; per-element result stays int32 when
// the grand total fits (callers check totals < 2^31)
__global__ void __launch_bounds__(1024) scan_add_base(int32_t* __restrict__ a, int64_t n, const int64_t* __restrict__ tile_tot) {
    const int64_t i = (int64_t)blockIdx.x * 1024 + threadIdx.x;
    if (i < n) a[i] += (int32_t)tile_tot[blockIdx.x];
}

// Up to 4096 elements: one CTA, one launch (thread t owns 4 consecutive elements).
#define MVS_SCAN_SMALL_MAX 4096
__global__ void __launch_bounds__(1024) scan_small(int32_t* __restrict__ a, int n, int64_t* __restrict__ total) {
    __shared__ int wsum[32];
    const int lo = threadIdx.x * 4;
    int v[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) v[k] = (lo + k < n) ? a[lo + k] : 0;
    const int sum = v[0] + v[1] + v[2] + v[3];
    int tot;
    int run = block_scan_incl(sum, wsum, &tot) - sum;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        if (lo + k < n) a[lo + k] = run;
        run += v[k];
    }
    if (threadIdx.x == 0) *total = tot;
}

// Up to MVS_SCAN_MID_TILES tiles of 1024: two launches -- scan_tiles, then every CTA reduces the totals of the
// tiles before it itself (a few hundred values) and adds the base; the last CTA writes the grand total.
#define MVS_SCAN_MID_TILES 2048
__global__ void __launch_bounds__(1024) scan_add_base_reduce(int32_t* __restrict__ a, int64_t n, const int64_t* __restrict__ tile_tot,
                                                             int T, int64_t* __restrict__ total) {
    __shared__ long long wsum[32];
    __shared__ long long s_base;
    long long part = 0;
    for (int i = threadIdx.x; i < (int)blockIdx.x; i += 1024) part += tile_tot[i];
#pragma unroll
    for (int sft = 16; sft > 0; sft >>= 1) part += __shfl_xor_sync(FULL, part, sft);
    if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = part;
    __syncthreads();
    if (threadIdx.x == 0) {
        long long b = 0;
        for (int w = 0; w < 32; ++w) b += wsum[w];
        s_base = b;
        if ((int)blockIdx.x == T - 1) *total = b + tile_tot[T - 1];
    }
    __syncthreads();
    const int64_t i = (int64_t)blockIdx.x * 1024 + threadIdx.x;
    if (i < n) a[i] += (int32_t)s_base;
}

int mvs_exclusive_scan_i32(int32_t* a, int64_t n, int64_t* tile_scratch, int64_t* total, cudaStream_t s) {
    if (n == 0) {
        MVS_CUDA_CHECK(cudaMemsetAsync(total, 0, sizeof(int64_t), s));
        return MVS_OK;
    }
    if (n <= MVS_SCAN_SMALL_MAX) {
        scan_small<<<1, 1024, 0, s>>>(a, (int)n, total);
        MVS_CUDA_CHECK(cudaGetLastError());
        return MVS_OK;
    }
    if ((n + 1023) / 1024 <= MVS_SCAN_MID_TILES) {
        const int Tm = (int)((n + 1023) / 1024);
        scan_tiles<<<Tm, 1024, 0, s>>>(a, n, tile_scratch);
        scan_add_base_reduce<<<Tm, 1024, 0, s>>>(a, n, tile_scratch, Tm, total);
        MVS_CUDA_CHECK(cudaGetLastError());
        return MVS_OK;
    }
    const int T = (int)((n + 1023) / 1024);
    scan_tiles<<<T, 1024, 0, s>>>(a, n, tile_scratch);
    scan_tile_totals<<<1, 1024, 0, s>>>(tile_scratch, T, total);
    scan_add_base<<<T, 1024, 0, s>>>(a, n, tile_scratch);
    MVS_CUDA_CHECK(cudaGetLastError());
    return MVS_OK;
}
