// ctNcc (MVS2.py:39-43) for a batch of descriptor pairs: out[m] = n/(n-1) * Pearson(a[m], b[m])
// on uint8 descriptors of length n, exact integer sums, fp64 ratio; NaN when either
// descriptor has zero variance (the reference divides by a zero std).  One warp per pair.
#include "mvs_common.cuh"

#define FULL 0xffffffffu

__global__ void __launch_bounds__(256)
    ncc_pairs_kernel(int64_t M, int n, const uint8_t* __restrict__ a, const uint8_t* __restrict__ b, double* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t m = warp0; m < M; m += nwarps) {
        long long sa = 0, sb = 0, saa = 0, sbb = 0, sab = 0;
        for (int i = lane; i < n; i += 32) {
            const long long x = a[m * n + i], y = b[m * n + i];
            sa += x; sb += y; saa += x * x; sbb += y * y; sab += x * y;
        }
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) {
            sa += __shfl_xor_sync(FULL, sa, s);
            sb += __shfl_xor_sync(FULL, sb, s);
            saa += __shfl_xor_sync(FULL, saa, s);
            sbb += __shfl_xor_sync(FULL, sbb, s);
            sab += __shfl_xor_sync(FULL, sab, s);
        }
        if (lane == 0) {
            const long long va = n * saa - sa * sa, vb = n * sbb - sb * sb;
            const long long num = n * sab - sa * sb;
            out[m] = (va == 0 || vb == 0) ? nan("") : ((double)num / sqrt((double)va * (double)vb)) * ((double)n / (double)(n - 1));
        }
    }
}

extern "C" int mvs_ncc_pairs(int device, int64_t M, int n, const uint8_t* a, const uint8_t* b, double* out, int on_device,
                             void* stream) {
    if (M < 0 || n < 2 || n > 65536 || (M > 0 && (!a || !b || !out))) {
        mvs_set_error("mvs_ncc_pairs: need M >= 0, 2 <= n <= 65536 and non-null buffers");
        return MVS_ERR_ARG;
    }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        mvs_set_error("mvs_ncc_pairs: no CUDA device is usable; this library has no CPU fallback");
        return MVS_ERR_CUDA;
    }
    if (M == 0) return MVS_OK;
    MVS_CUDA_CHECK(cudaSetDevice(device));
    const int blocks = (int)((M + 7) / 8 < 4096 ? (M + 7) / 8 : 4096);
    if (on_device) {
        ncc_pairs_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(M, n, a, b, out);
        MVS_CUDA_CHECK(cudaGetLastError());
        return MVS_OK;
    }
    uint8_t *da = nullptr, *db = nullptr;
    double* dout = nullptr;
    const size_t bytes = (size_t)M * n;
    int rc = MVS_OK;
    if (cudaMalloc(&da, bytes) != cudaSuccess || cudaMalloc(&db, bytes) != cudaSuccess ||
        cudaMalloc(&dout, sizeof(double) * M) != cudaSuccess) {
        mvs_set_error("mvs_ncc_pairs: device allocation failed");
        cudaGetLastError();
        rc = MVS_ERR_NOMEM;
    } else if (cudaMemcpy(da, a, bytes, cudaMemcpyHostToDevice) != cudaSuccess ||
               cudaMemcpy(db, b, bytes, cudaMemcpyHostToDevice) != cudaSuccess) {
        mvs_set_error("mvs_ncc_pairs: upload failed: %s", cudaGetErrorString(cudaGetLastError()));
        rc = MVS_ERR_CUDA;
    } else {
        ncc_pairs_kernel<<<blocks, 256>>>(M, n, da, db, dout);
        if (cudaMemcpy(out, dout, sizeof(double) * M, cudaMemcpyDeviceToHost) != cudaSuccess) {
            mvs_set_error("mvs_ncc_pairs: kernel or download failed: %s", cudaGetErrorString(cudaGetLastError()));
            rc = MVS_ERR_CUDA;
        }
    }
    cudaFree(da);
    cudaFree(db);
    cudaFree(dout);
    return rc;
}
