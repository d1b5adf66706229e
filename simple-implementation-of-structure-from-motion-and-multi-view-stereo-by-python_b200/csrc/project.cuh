// Reference-view projection and window anchoring shared by the binning and scoring kernels.
#pragma once
#include "mvs_common.cuh"

// Projection, bit-compatible with cv2.projectPoints as called by utils.py:241-244:
//   X = (r0*c0 + r1*c1 + r2*c2) + t  (left to right, no FMA contraction),
//   z = z ? 1/z : 1;  x = (X*z)*fx + cx.
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
