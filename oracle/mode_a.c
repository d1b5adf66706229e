/*
 * Mode A ("reference-exact") photo-consistency scorer, restated in plain C (fp64 / exact integers).
 *
 * TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): a second, independent checker beside the NumPy
 * restatement oracle/mode_a.py -- only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may
 * load it; the product path never does (tests/test_abi_cpu.py).  Built by oracle/build_c.py (gcc, no FMA
 * contraction) into oracle/_build/libmodea.so; pinned to the golden vectors the reference itself produced
 * (tests/test_oracle_cpu.py::test_c_restatement_*).
 *
 * Follows, step by step (file:line under the reference root):
 *   MVS2.py:62-77             MyPatch.photo_consistenecy_test -- every view sampled at the REFERENCE view's
 *                             projection of c (MVS2.py:68), strict '>' threshold, avg over the visible views
 *   MVS2.py:39-43             ctNcc -- z-scores with the population std, sum / (n - 1): n/(n-1) x Pearson;
 *                             zero variance -> NaN -> not visible
 *   HarrisFeatures.py:116-133 getDescFeatures -- gray = cv2 BGR2GRAY applied to an RGB array,
 *                             g = (R*3735 + G*19235 + B*9798 + 16384) >> 15; int() truncation; bounds
 *                             row-wid >= 0, row+wid+1 < H, col-wid > 0 (strict), col+wid+1 < W
 *   utils.py:241-244          projectPoint -- cv2.projectPoints' operation order on the Rodrigues round trip
 *                             R' of the file rotation (passed in by the caller): z = z ? 1/z : 1,
 *                             x = (X*z)*fx + cx
 * Declared divergence (as in the NumPy oracle): a non-finite projection is rejected (the reference raises).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>

#ifdef _OPENMP
#include <omp.h>
#endif

/* HarrisFeatures.py:124-125 on RGB-ordered pixels: rgb [npix,3] u8 -> gray [npix] u8 */
void modea_gray(const uint8_t* rgb, int64_t npix, uint8_t* gray) {
    for (int64_t i = 0; i < npix; ++i) {
        const int r = rgb[3 * i], g = rgb[3 * i + 1], b = rgb[3 * i + 2];
        gray[i] = (uint8_t)((r * 3735 + g * 19235 + b * 9798 + 16384) >> 15);
    }
}

/* utils.py:241-244 with cv2.projectPoints' operation order; Rrt [9] row-major, t [3], k4 = fx, fy, cx, cy */
static void project_ref(const double* Rrt, const double* t, const double* k4, const double* c, double* x, double* y) {
    const double X = Rrt[0] * c[0] + Rrt[1] * c[1] + Rrt[2] * c[2] + t[0];
    const double Y = Rrt[3] * c[0] + Rrt[4] * c[1] + Rrt[5] * c[2] + t[1];
    const double Z = Rrt[6] * c[0] + Rrt[7] * c[1] + Rrt[8] * c[2] + t[2];
    const double iz = Z != 0.0 ? 1.0 / Z : 1.0;
    *x = (X * iz) * k4[0] + k4[2];
    *y = (Y * iz) * k4[1] + k4[3];
}

/*
 * N hypotheses.  gray [V,H,W] u8 (planar), Rrt [V,9], t [V,3], k4 [V,4], c [N,3], ref [N].
 * Outputs: vis [N,V] u8 (0/1), avg [N], count [N], xy [N,2] (NaN for an invalid reference view),
 * ncc [N,V] f64 or NULL (NaN where the reference yields no score, and at the reference view).
 * threads <= 0: all cores (OpenMP), 1: the scalar port.
 */
int modea_score(const uint8_t* gray, int V, int H, int W, const double* Rrt, const double* t, const double* k4, int64_t N,
                const double* c, const int32_t* ref, double thr, int wid, uint8_t* vis, double* avg, int32_t* count,
                double* xy, double* ncc, int threads) {
    if (V < 1 || H < 1 || W < 1 || wid < 0 || N < 0) return -1;
    const int K = 2 * wid + 1;
    const int64_t n = (int64_t)K * K;
    const double scale = (double)n / ((double)n - 1.0);
#ifdef _OPENMP
    if (threads > 0) omp_set_num_threads(threads);
#pragma omp parallel for schedule(dynamic, 64)
#endif
    for (int64_t h = 0; h < N; ++h) {
        uint8_t* vh = vis + h * V;
        for (int v = 0; v < V; ++v) {
            vh[v] = 0;
            if (ncc) ncc[h * V + v] = NAN;
        }
        avg[h] = 0.0;
        count[h] = 0;
        const int r = ref[h];
        double x = NAN, y = NAN;
        if (r >= 0 && r < V) project_ref(Rrt + 9 * r, t + 3 * r, k4 + 4 * r, c + 3 * h, &x, &y);
        xy[2 * h] = x;
        xy[2 * h + 1] = y;
        if (!(isfinite(x) && isfinite(y))) continue;                         /* declared divergence */
        const double lim = 1099511627776.0;                                  /* 2^40: clamp before the cast */
        const int64_t col = (int64_t)trunc(x > lim ? lim : (x < -lim ? -lim : x));
        const int64_t row = (int64_t)trunc(y > lim ? lim : (y < -lim ? -lim : y));
        if (!(row - wid >= 0 && row + wid + 1 < H && col - wid > 0 && col + wid + 1 < W)) continue;   /* HarrisFeatures.py:128 */
        /* the reference window and its exact sums */
        const uint8_t* gr = gray + ((int64_t)r * H + (row - wid)) * W + (col - wid);
        int64_t Sr = 0, SSr = 0;
        for (int a = 0; a < K; ++a)
            for (int b = 0; b < K; ++b) {
                const int64_t w = gr[(int64_t)a * W + b];
                Sr += w;
                SSr += w * w;
            }
        const int64_t var_r = n * SSr - Sr * Sr;
        double acc = 0.0;
        int cnt = 0;
        for (int v = 0; v < V; ++v) {
            if (v == r) continue;
            const uint8_t* gv = gray + ((int64_t)v * H + (row - wid)) * W + (col - wid);   /* MVS2.py:68: same (row, col) */
            int64_t S = 0, SS = 0, SAB = 0;
            for (int a = 0; a < K; ++a)
                for (int b = 0; b < K; ++b) {
                    const int64_t w = gv[(int64_t)a * W + b], wr = gr[(int64_t)a * W + b];
                    S += w;
                    SS += w * w;
                    SAB += w * wr;
                }
            const int64_t var = n * SS - S * S;
            if (var == 0 || var_r == 0) continue;                            /* NaN in the reference: never visible */
            const double val = (double)(n * SAB - S * Sr) / sqrt((double)var * (double)var_r) * scale;
            if (ncc) ncc[h * V + v] = val;
            if (val > thr) {                                                 /* strict */
                vh[v] = 1;
                acc += val;
                ++cnt;
            }
        }
        count[h] = cnt;
        avg[h] = cnt > 0 ? acc / (double)cnt : 0.0;
    }
    return 0;
}
