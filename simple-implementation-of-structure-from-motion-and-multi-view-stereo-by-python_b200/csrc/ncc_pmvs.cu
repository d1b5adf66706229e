// K2 ncc_score_pmvs ("Mode B") for sm_100a: per-view projection, oriented mu x mu grid,
// bilinear taps, mean multi-view NCC, on-chip argmax over hypothesis sets.
//
// This scorer does NOT exist in the reference (its MVS2.py:62-77 samples every view at the
// reference camera's projection with a truncated window); it is the one BASELINE.json's
// north_star describes and its normative spec is oracle/mode_b.py.  With
// MVS_PMVS_REDUCE_TO_REFEXACT it collapses onto the reference's behaviour (same camera for
// every view, image-aligned integer lattice, no interpolation, the bounds rule of
// HarrisFeatures.py:128) and must then agree with K1.
//
// Hardware mapping (DESIGN.md section "K2"): ONE WARP PER CTA, one hypothesis (or one hypothesis SET when
// selecting) per warp iteration.
//   * staging: per 32 views, lane v prepares view v -- the projection of the patch centre through
//     K_v (R'_v c + t_v) in fp64 and the per-step increments RELATIVE to it in fp32 -- into shared memory; a tap
//     position is then 6 FMA + 1 reciprocal, accurate to ~1e-6 px;
//   * phase 1, lanes span the mu*mu samples (ceil(mu^2/32) per lane): the four bilinear taps of a sample come
//     from ONE texture-gather instruction (tld4) on a gather-enabled 2-D ATLAS that holds every view as a tile
//     (the texture path, not the LSU), as normalised floats; interpolation weights are exact fp32, not the
//     sampler's 8-bit ones; the interpolated values of 16 views go to shared memory [view][sample]; a half whose
//     16 views are all "safe" (footprint provably inside the image) runs without per-tap bounds tests;
//   * phase 2, lanes span the VIEWS: lane L sums half of the samples of view L & 15 on pivot-shifted values
//     (x - x[0], so low-variance windows keep full fp32 precision); one shuffle per quantity joins the halves;
//   * the best hypothesis of a set (highest mean NCC among those with >= bound visible views, lowest index on
//     ties) is tracked in registers and written once per set.
// Tensor cores are not used: a gather-bound reduction has no dense contraction.
#include "project.cuh"
#include "scan.cuh"
#include <stdlib.h>

#define FULL 0xffffffffu
#define PMVS_VAR_MIN (1e-3f / 65025.0f)        // oracle/mode_b.py VAR_MIN in (grey/255)^2

// Cameras as the staging lanes read them: lane v stages view v, so the fields are stored FIELD-MAJOR
// (structure of arrays): a warp-wide read of one field is one coalesced run.  With the 128-byte CamProj records
// every such read touched 32 different lines -- ~14 data-pipe wavefronts per load, ~670 per hypothesis, more than
// the shared-memory traffic of the whole kernel.
//   cam64 [16][V] double: r0..r8, t0..t2, fx, fy, cx, cy        cam32 [13][V] float: r0..r8, fx, fy, cx, cy
#define PMVS_CAM64_FIELDS 16
#define PMVS_CAM32_FIELDS 13

// Texture handles of the views, read with a warp-uniform index: a handle that is provably
// uniform lets TLD4 take it from a uniform register (no per-lane "waterfall" loop).  One table
// per device; re-uploaded when another context scored last (mvs_launch_score_pmvs).
#define PMVS_MAX_VIEWS 1024
__constant__ cudaTextureObject_t c_tex[PMVS_MAX_VIEWS];
__constant__ float2 c_off[PMVS_MAX_VIEWS];         // tile origin of each view inside its atlas

struct PmvsArgs {
    const CamProj* cams;
    const double* cam64;       // [PMVS_CAM64_FIELDS][V]
    const float* cam32;        // [PMVS_CAM32_FIELDS][V]
    const cudaTextureObject_t* tex;
    const float2* off;         // [V] tile origin of each view inside its atlas
    int V, H, W;
    int flags;
    int cam_stride;            // = V, the field pitch of cam64 / cam32.  A kernel argument of ITS OWN: when the staging
                               // lanes' address arithmetic used A.V itself, ptxas kept A.V in a vector register and computed the
                               // texture-handle index min(view, A.V - 1) there too -- the handle then lost its uniform register
                               // and every TLD4 of the multi-atlas variants got a waterfall loop (tests/test_abi_cpu.py)
    int group;                 // hypotheses per selection set (<= 1: no selection)
    int bound;
    float thr;
    const double* c;
    const double* nrm;
    const int32_t* ref;
    const uint64_t* cand;      // optional [N, mw] candidate-view mask
    uint64_t* vis_out;
    double* avg_out;
    int32_t* count_out;
    double* xy_out;
    float* ncc_out;
    int32_t* best_idx;
    double* best_avg;
};

// 1/x as ONE MUFU.RCP (__fdividef(1, x) compiles to five instructions: a denormal-range test, a scaling and a
// rescaling around the same MUFU.RCP; the denominators here are depths of ~0.5, never denormal)
__device__ __forceinline__ float rcp_approx(float x) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) v += __shfl_xor_sync(FULL, v, s);
    return v;
}

// One view's staged parameters.  A sample at grid offset (a, b) has homogeneous image
// coordinates h0 + a*hx + b*hy; its pixel position is written RELATIVE to the projection
// (uc, vc) of the patch centre, which the staging lane knows in fp64:
//   u = uc + (a*gxu + b*gyu) / z,  gxu = hx.X - uc*hx.Z, ...,  z = Z0 + a*hxZ + b*hyZ
// so fp32 only ever carries offsets of a few pixels (error ~1e-6 px instead of ~5e-5 px).
// uc, vc are split into integer part and fraction.  Everything is stored divided by Z0 (z/Z0 = 1 + a*ex + b*ey),
// which leaves TEN floats: every lane fetches them per view as two LDS.128 and one LDS.64.
struct ViewAffine {
    float iu, fu, iv, fv;
    float ex, ey, gxu, gyu;      // ex = hxZ/Z0, ey = hyZ/Z0 (NaN when the centre is not in front of the view)
    float gxv, gyv;
};

// Bilinear sample through the gather path.  (u0, v0) = integer tap origin (pixel centres at
// integers), (fu, fv) = fractions.  Returns the value in [0,1] and whether all four taps are
// inside the image.
__device__ __forceinline__ float tap4(const bool checked, cudaTextureObject_t tex, float2 off, float uc1, float vc1, float fu,
                                      float fv, bool front, float wm2, float hm2, bool& ok) {
    // (uc1, vc1) = tap origin + tile origin + 1: the centre of the 2x2 gather footprint inside the atlas.
    // Bounds are tested against the VIEW's tile, not the atlas; an unchecked ("safe") view is known to be
    // inside, and the texture's clamp addressing makes the fetch of a rejected tap harmless.
    ok = !checked || (front && (uc1 >= off.x + 1.0f) && (uc1 <= off.x + 1.0f + wm2) && (vc1 >= off.y + 1.0f) &&
                      (vc1 <= off.y + 1.0f + hm2));
    // components: x=(0,1) y=(1,1) z=(1,0) w=(0,0) as (column offset, row offset)
    const float4 g = tex2Dgather<float4>(tex, uc1, vc1, 0);
    const float top = fmaf(fu, g.z - g.w, g.w);
    const float bot = fmaf(fu, g.y - g.x, g.x);
    return fmaf(fv, bot - top, top);
}

template <int MU, bool REDUCE_A, int MINB, bool ONE_ATLAS, int VB>
__global__ void __launch_bounds__(32, MINB) ncc_score_pmvs(const PmvsArgs A, int64_t N) {
    static_assert(VB == 16 || VB == 32, "views per reduction batch");
    constexpr int NS = MU * MU;
    constexpr int SPL = (NS + 31) / 32;                    // samples per lane
    constexpr float HALF = 0.5f * (MU - 1);
    // ONE WARP PER CTA: every loop bound and view index below then derives from blockIdx and the
    // kernel arguments only, i.e. is provably warp-uniform, which lets the texture handle travel
    // in a uniform register (no per-lane waterfall loop around TLD4) and keeps the gathers of a
    // batch of views in flight together.
    // 32 views of the current block + the reference view; three arrays (not 48-byte records with 8 bytes of padding):
    // the 264 bytes decide whether 31 or 32 warps fit an SM at mu = 7
    __shared__ float4 s_va[33], s_vb[33];                  // iu fu iv fv | ex ey gxu gyu
    __shared__ float2 s_vc[33];                            // gxv gyv
    // samples of 16 views.  Row stride = 2 (mod 32) and the second half-warp starts at an odd sample
    // offset, so the 32 lanes of phase 2 (16 views x 2 sample halves) hit 32 different banks.
    // Phase 2 reads a view's samples four at a time (VB = 32: every lane ONE view and all its samples; VB = 16: lane L
    // and lane L + 16 share view L, the first and the second part of its sample quads); the row stride is a multiple of
    // 4 floats with an odd quotient, so the eight lanes of a quarter-warp read eight different 16-byte bank groups.
    constexpr int NS4 = (NS + 3) / 4;                      // sample quads per view
    constexpr int VSTRIDE = 4 * (NS4 | 1);
    __shared__ __align__(16) float s_val[VB][VSTRIDE];
    __shared__ __align__(16) float s_dref[4 * NS4];        // pivot-shifted samples of the reference view
    // Projection of the CURRENT centre through the first 64 views (floor and fraction of uc, vc; depth), kept across consecutive
    // hypotheses with the same centre and reference view -- the normals of one depth in a depth x normal set -- so
    // that the fp64 part of the staging (12 loads, 3 dot products, a division) runs once per centre, not per normal.
    // Lane l only ever reads back what it wrote itself (views l and 32 + l): no synchronisation.
    __shared__ float4 s_cuv[64];                           // floor(uc), uc - floor(uc), floor(vc), vc - floor(vc)
    __shared__ float s_czv[64];

    // ONE_ATLAS (every BASELINE shape up to 128 x 1080p): all views are tiles of one texture, so its handle is read
    // from the constant bank ONCE instead of once per view (LDCU + index clamp + R2UR per gather batch entry)
    const cudaTextureObject_t tex0 = c_tex[0];
    const int lane = threadIdx.x;
    constexpr int wib = 0;
    const int64_t warp0 = blockIdx.x;
    const int64_t nwarps = gridDim.x;
    const int mw = (A.V + 63) >> 6;
    const int group = A.group > 1 ? A.group : 1;
    const int64_t n_sets = (N + group - 1) / group;
    const float cn = (float)NS / (float)(NS - 1);
    constexpr bool reduce_a = REDUCE_A;
    const float wm2 = (float)(A.W - 2), hm2 = (float)(A.H - 2);

    // this lane's sample offsets on the grid (m = k*MU + j)
    float aj[SPL], ak[SPL];
    bool live[SPL];
#pragma unroll
    for (int q = 0; q < SPL; ++q) {
        const int m = lane + 32 * q;
        live[q] = m < NS;
        const int mm = live[q] ? m : 0;
        aj[q] = (float)(mm % MU) - HALF;
        ak[q] = (float)(mm / MU) - HALF;
    }

    for (int64_t set = warp0; set < n_sets; set += nwarps) {
        float best_key = -INFINITY;
        int best_i = -1;
        double pc0 = 0.0, pc1 = 0.0, pc2 = 0.0;            // centre / reference view whose projections are cached
        int pr = -1;
        unsigned cached = 0u;                              // bit b: the cache of view block b belongs to (pc, pr)
        double px = 0.0, py = 0.0, pZc = 0.0;              // ... and its projection through the reference view
        for (int gi = 0; gi < group; ++gi) {
            const int64_t h = set * group + gi;
            if (h >= N) break;
            // ---- reference view: projection of the centre (fp64, cv2 order, as Mode A), patch axes
            const int r = __ldg(A.ref + h);
            const double c0 = __ldg(A.c + 3 * h), c1 = __ldg(A.c + 3 * h + 1), c2 = __ldg(A.c + 3 * h + 2);
            double x = nan(""), y = nan("");
            bool hyp_ok = (r >= 0) && (r < A.V);
            float ex[3] = {0, 0, 0}, ey[3] = {0, 0, 0}, step = 0.0f;
            int row = 0, col = 0;
            // consecutive hypotheses with the same centre and reference view (the normals of one depth): the fp64
            // projections of the centre are taken over from the previous hypothesis
            const bool same_centre = !reduce_a && r == pr && c0 == pc0 && c1 == pc1 && c2 == pc2;
            if (!same_centre) {
                cached = 0u;
                pc0 = c0; pc1 = c1; pc2 = c2; pr = r;
            }
            if (hyp_ok) {
                const CamProj& cr = A.cams[r];
                if (!same_centre) {
                    project_ref(cr, c0, c1, c2, px, py);
                    pZc = cr.r[6] * c0 + cr.r[7] * c1 + cr.r[8] * c2 + cr.t[2];
                }
                x = px; y = py;
                const double Zc = pZc;
                if (reduce_a) {
                    hyp_ok = window_anchor(x, y, A.H, A.W, (MU - 1) / 2, row, col);
                } else {
                    const float n0 = (float)__ldg(A.nrm + 3 * h), n1 = (float)__ldg(A.nrm + 3 * h + 1),
                                n2 = (float)__ldg(A.nrm + 3 * h + 2);
                    const float nn = sqrtf(n0 * n0 + n1 * n1 + n2 * n2);
                    const float inn = 1.0f / nn;
                    const float nh[3] = {n0 * inn, n1 * inn, n2 * inn};
                    const float a0 = (float)cr.r[0], a1 = (float)cr.r[1], a2 = (float)cr.r[2];
                    const float dn = a0 * nh[0] + a1 * nh[1] + a2 * nh[2];
                    const float p0 = a0 - dn * nh[0], p1 = a1 - dn * nh[1], p2 = a2 - dn * nh[2];
                    const float al = sqrtf(p0 * p0 + p1 * p1 + p2 * p2);
                    const float ial = 1.0f / al;
                    ex[0] = p0 * ial; ex[1] = p1 * ial; ex[2] = p2 * ial;
                    ey[0] = ex[1] * nh[2] - ex[2] * nh[1];
                    ey[1] = ex[2] * nh[0] - ex[0] * nh[2];
                    ey[2] = ex[0] * nh[1] - ex[1] * nh[0];
                    step = (float)(Zc / (0.5 * (cr.fx + cr.fy)));
                    hyp_ok = isfinite(c0) && isfinite(c1) && isfinite(c2) && isfinite(nn) && (nn > 0.0f) && isfinite(Zc) &&
                             (Zc > 0.0) && (al >= 1e-9f);
                }
            }
            if (lane == 0 && A.xy_out) {
                A.xy_out[2 * h] = x;
                A.xy_out[2 * h + 1] = y;
            }

            // prepare view v for this hypothesis into slot `slot` (executed by one lane per view)
            // cidx >= 0: cache slot of this (lane, view block); use_cache: that slot holds this centre's projection
            auto stage_view = [&](int v, int slot, int cidx, bool use_cache) -> bool {
                const int cam = reduce_a ? r : v;          // MVS2.py:68: the reference camera for every view
                const double* c64 = A.cam64 + cam;
                const float* c32 = A.cam32 + cam;
                const int Vs = A.cam_stride;
                float4 cuv;                                 // the centre's projection, split: integer part and fraction
                float Zvf;
                if (use_cache) {
                    cuv = s_cuv[cidx];
                    Zvf = s_czv[cidx];
                } else {
                    const double cvcx = c64[14 * Vs], cvcy = c64[15 * Vs];
                    const double cvfx = c64[12 * Vs], cvfy = c64[13 * Vs];
                    const double Xc = c64[0] * c0 + c64[1 * Vs] * c1 + c64[2 * Vs] * c2 + c64[9 * Vs];
                    const double Yc = c64[3 * Vs] * c0 + c64[4 * Vs] * c1 + c64[5 * Vs] * c2 + c64[10 * Vs];
                    const double Zv = c64[6 * Vs] * c0 + c64[7 * Vs] * c1 + c64[8 * Vs] * c2 + c64[11 * Vs];
                    const double izd = 1.0 / Zv;
                    const double uc = (cvfx * Xc + cvcx * Zv) * izd;
                    const double vc = (cvfy * Yc + cvcy * Zv) * izd;
                    const double ucf = floor(uc), vcf = floor(vc);
                    cuv = make_float4((float)ucf, (float)(uc - ucf), (float)vcf, (float)(vc - vcf));
                    Zvf = (float)Zv;
                    if (cidx >= 0) {
                        s_cuv[cidx] = cuv;
                        s_czv[cidx] = Zvf;
                    }
                }
                ViewAffine va;
                // integer parts carry the view's tile origin inside the atlas and the +1 that addresses the
                // centre of the 2x2 gather footprint, so a tap coordinate is iu + floor(.) with no further adds
                const float2 org = A.off[v];
                va.iu = cuv.x + org.x + 1.0f; va.fu = cuv.y;
                va.iv = cuv.z + org.y + 1.0f; va.fv = cuv.w;
                const float Z0 = Zvf;
                const float iZ0 = Z0 > 0.0f ? rcp_approx(Z0) : nanf("");   // behind the camera: every tap test fails on NaN
                const float f0 = c32[0], f1 = c32[1 * Vs], f2 = c32[2 * Vs], f3 = c32[3 * Vs], f4 = c32[4 * Vs], f5 = c32[5 * Vs],
                            f6 = c32[6 * Vs], f7 = c32[7 * Vs], f8 = c32[8 * Vs], cffx = c32[9 * Vs], cffy = c32[10 * Vs];
                const float rx0 = f0 * ex[0] + f1 * ex[1] + f2 * ex[2];
                const float rx1 = f3 * ex[0] + f4 * ex[1] + f5 * ex[2];
                const float rx2 = f6 * ex[0] + f7 * ex[1] + f8 * ex[2];
                const float ry0 = f0 * ey[0] + f1 * ey[1] + f2 * ey[2];
                const float ry1 = f3 * ey[0] + f4 * ey[1] + f5 * ey[2];
                const float ry2 = f6 * ey[0] + f7 * ey[1] + f8 * ey[2];
                // hx.X - uc*hx.Z = step*(fx*rx0 + (cx - uc)*rx2): the principal point cancels against uc
                // (fp32 is enough here: du, dv only multiply the perspective terms rx2, ry2 ~ 1e-3 of the others)
                const float du = (c32[11 * Vs] - cuv.x) - cuv.y, dv = (c32[12 * Vs] - cuv.z) - cuv.w;
                const float hxZ = step * rx2, hyZ = step * ry2;
                const float gxu = step * fmaf(du, rx2, cffx * rx0), gyu = step * fmaf(du, ry2, cffx * ry0);
                const float gxv = step * fmaf(dv, rx2, cffy * rx1), gyv = step * fmaf(dv, ry2, cffy * ry1);
                va.ex = hxZ * iZ0; va.ey = hyZ * iZ0;
                va.gxu = gxu * iZ0; va.gyu = gyu * iZ0;
                va.gxv = gxv * iZ0; va.gyv = gyv * iZ0;
                s_va[slot] = make_float4(va.iu, va.fu, va.iv, va.fv);
                s_vb[slot] = make_float4(va.ex, va.ey, va.gxu, va.gyu);
                s_vc[slot] = make_float2(va.gxv, va.gyv);
                // "safe" view: the whole mu x mu footprint provably stays inside the image and in front of
                // the camera (bound on the tap offsets from the staged increments, 0.01 px of slack), so its
                // taps need no per-tap bounds test
                const float zmin = Z0 - HALF * (fabsf(hxZ) + fabsf(hyZ));
                const float izm = rcp_approx(zmin);            // (an IEEE division's slow path would break the warp-uniformity proof below)
                const float ru = HALF * (fabsf(gxu) + fabsf(gyu)) * izm + 0.01f;
                const float rv = HALF * (fabsf(gxv) + fabsf(gyv)) * izm + 0.01f;
                const float ucl = cuv.x + cuv.y, vcl = cuv.z + cuv.w;
                return !reduce_a && (zmin > 0.0f) && (ucl - ru >= 0.0f) && (ucl + ru <= wm2 + 0.98f) && (vcl - rv >= 0.0f) &&
                       (vcl + rv <= hm2 + 0.98f);
            };

            // Issue the taps of the view staged in `slot` (no warp-synchronous operation in here,
            // so the gathers of a whole batch of views are in flight together).
            auto issue_view = [&](const bool checked, int slot, cudaTextureObject_t tex, float2 off, float (&val)[SPL], bool& ok_all) {
                const float4 q0 = s_va[slot], q1 = s_vb[slot];            // iu fu iv fv | ex ey gxu gyu
                const float2 q2 = s_vc[slot];                             // gxv gyv
                ok_all = true;
#pragma unroll
                for (int q = 0; q < SPL; ++q) {
                    float u0, v0, fu, fv;
                    bool front = true;
                    if (reduce_a) {
                        u0 = (float)col + aj[q] + off.x + 1.0f;
                        v0 = (float)row + ak[q] + off.y + 1.0f;
                        fu = fv = 0.0f;
                    } else {
                        const float z = fmaf(ak[q], q1.y, fmaf(aj[q], q1.x, 1.0f));   // depth relative to the centre's
                        const float iz = rcp_approx(z);
                        const float tu = fmaf(fmaf(ak[q], q1.w, aj[q] * q1.z), iz, q0.y);
                        const float tv = fmaf(fmaf(ak[q], q2.y, aj[q] * q2.x), iz, q0.w);
                        // (floor on the FMA pipe through the 1.5 * 2^23 trick instead of FRND was measured: 4.71 vs 4.73 ms at
                        // mu = 5, 6.51 vs 6.22 ms at mu = 7 -- not kept)
                        const float flu = floorf(tu), flv = floorf(tv);
                        u0 = q0.x + flu;
                        v0 = q0.z + flv;
                        fu = tu - flu;
                        fv = tv - flv;
                        front = z > 0.0f;
                    }
                    bool ok;
                    val[q] = tap4(checked, tex, off, u0, v0, fu, fv, front, wm2, hm2, ok);
                    if (checked) ok_all &= ok || !live[q];
                }
            };

            double acc = 0.0;
            int count = 0;
            float Sr = 0.0f, ssr = 0.0f;
            for (int w32 = 0; w32 < 2 * mw; ++w32) {       // 32 views per mask word
                uint32_t word = 0u;
                bool safe_lo = false, safe_hi = false;
                if (hyp_ok) {                              // uniform across the warp
                    // ---- stage this block's 32 views (one lane each); block 0 also stages the reference view
                    __syncwarp();
                    const bool same_c = !reduce_a && w32 < 2 && ((cached >> w32) & 1u);
                    const bool safe = stage_view(min(w32 * 32 + lane, A.V - 1), lane, w32 < 2 ? w32 * 32 + lane : -1, same_c);   // lanes past the last view shadow it
                    if (w32 < 2) cached |= 1u << w32;      // (uniform) this block's cache now belongs to the current centre
                    // one vote per half: every view of the half safe -> its taps skip the bounds tests
                    safe_lo = __all_sync(FULL, safe || lane >= 16);
                    safe_hi = __all_sync(FULL, safe || lane < 16);
                    if (w32 == 0 && r >= 32 && lane == 0) stage_view(r, 32, -1, false);
                    __syncwarp();
                    if (w32 == 0) {
                        // ---- the reference view's own samples
                        bool ok_all;
                        float val[SPL];
                        issue_view(true, r < 32 ? r : 32, A.tex[r], A.off[r], val, ok_all);
                        hyp_ok = __all_sync(FULL, ok_all);
                        const float pivot = __shfl_sync(FULL, val[0], 0);
                        float s = 0.0f, ss = 0.0f;
#pragma unroll
                        for (int q = 0; q < SPL; ++q) {
                            const float dq = live[q] ? val[q] - pivot : 0.0f;
                            if (live[q]) s_dref[lane + 32 * q] = dq;
                            s += dq;
                            ss = fmaf(dq, dq, ss);
                        }
                        Sr = warp_sum(s);
                        ssr = warp_sum(ss) - Sr * Sr * (1.0f / NS);  // sum of squared deviations of the reference samples
                    }
                }
                if (hyp_ok) {
                    const uint32_t cand32 =
                        A.cand ? (uint32_t)(__ldg(A.cand + h * mw + (w32 >> 1)) >> (32 * (w32 & 1))) : 0xffffffffu;
#pragma unroll 1
                    for (int bt = 0; bt < 32 / VB; ++bt) {                   // reduction batches of VB views
                        if (w32 * 32 + bt * VB >= A.V) break;
                        constexpr int TB = SPL == 1 ? 16 : (SPL == 2 ? 8 : 4);
                        uint32_t bad = 0u;                                   // bit j: some tap of view j of the batch is outside
                        __syncwarp();                                        // the previous batch's readers are done
#pragma unroll 1
                        for (int hh = 0; hh < VB / 16; ++hh) {
                        const int half = bt * (VB / 16) + hh;                // which 16 views of the block of 32
                        const int vbase = w32 * 32 + half * 16;
                        if (vbase >= A.V) break;
                        // ---- phase 1: lanes span the samples.  Taps of TB views are issued together, the
                        // interpolated values go to shared memory [view][sample].
                        if (half ? safe_hi : safe_lo) {                      // safe half: taps without bounds tests
#pragma unroll 1                                             // (unrolled, the mu = 7 loop body outgrows the instruction cache:
                                                             //  ncu "no instruction" stall 2.0 per issue)
                        for (int j0 = 0; j0 < 16; j0 += TB) {
                            float val[TB][SPL];
                            bool okl[TB];
#pragma unroll
                            for (int jj = 0; jj < TB; ++jj)
                                // views past the last one re-sample it (branch-free); their sums are never scored
                                issue_view(false, half * 16 + j0 + jj, ONE_ATLAS ? tex0 : c_tex[min(vbase + j0 + jj, A.V - 1)],
                                           make_float2(0.0f, 0.0f), val[jj], okl[jj]);      // the tile origin only enters the bounds tests
#pragma unroll
                            for (int jj = 0; jj < TB; ++jj) {
                                if (!okl[jj]) bad |= 1u << (hh * 16 + j0 + jj);
#pragma unroll
                                for (int q = 0; q < SPL; ++q)
                                    if (live[q]) s_val[hh * 16 + j0 + jj][lane + 32 * q] = val[jj][q];
                            }
                        }
                        } else {
#pragma unroll 1                                             // (unrolled, the mu = 7 loop body outgrows the instruction cache:
                                                             //  ncu "no instruction" stall 2.0 per issue)
                        for (int j0 = 0; j0 < 16; j0 += TB) {
                            float val[TB][SPL];
                            bool okl[TB];
#pragma unroll
                            for (int jj = 0; jj < TB; ++jj)
                                // views past the last one re-sample it (branch-free); their sums are never scored
                                issue_view(true, half * 16 + j0 + jj, ONE_ATLAS ? tex0 : c_tex[min(vbase + j0 + jj, A.V - 1)],
                                           c_off[min(vbase + j0 + jj, A.V - 1)], val[jj], okl[jj]);
#pragma unroll
                            for (int jj = 0; jj < TB; ++jj) {
                                if (!okl[jj]) bad |= 1u << (hh * 16 + j0 + jj);
#pragma unroll
                                for (int q = 0; q < SPL; ++q)
                                    if (live[q]) s_val[hh * 16 + j0 + jj][lane + 32 * q] = val[jj][q];
                            }
                        }
                        }
                        }
                        const uint32_t usable = ~__reduce_or_sync(FULL, bad);
                        __syncwarp();
                        if (VB == 32) {
                            // ---- phase 2: lane L owns view L of the block and sums all its samples, four per load
                            // (pivot = the view's first sample, so low-variance windows keep full fp32 precision)
                            const float4* rowq = reinterpret_cast<const float4*>(s_val[lane]);
                            const float4* refq = reinterpret_cast<const float4*>(s_dref);
                            float pivot = 0.0f, Sd = 0.0f, SSd = 0.0f, SAB = 0.0f;
#pragma unroll
                            for (int g = 0; g < NS4; ++g) {
                                const float4 xq = rowq[g], dq = refq[g];
                                const float xs[4] = {xq.x, xq.y, xq.z, xq.w}, ds[4] = {dq.x, dq.y, dq.z, dq.w};
                                if (g == 0) pivot = xq.x;
#pragma unroll
                                for (int e = 0; e < 4; ++e) {
                                    if (4 * g + e < NS && 4 * g + e > 0) {   // (the pivot's own deviation is 0)
                                        const float dd = xs[e] - pivot;
                                        Sd += dd;
                                        SSd = fmaf(dd, dd, SSd);
                                        SAB = fmaf(dd, ds[e], SAB);
                                    }
                                }
                            }
                            const int v = w32 * 32 + lane;
                            const float ss = SSd - Sd * Sd * (1.0f / NS);
                            const float cov = SAB - Sd * Sr * (1.0f / NS);
                            const bool scored = (v < A.V) && (v != r) && ((usable >> lane) & 1u) && ((cand32 >> lane) & 1u) &&
                                                (ss * (1.0f / NS) >= PMVS_VAR_MIN) && (ssr * (1.0f / NS) >= PMVS_VAR_MIN);
                            const float val_ncc = cov * rsqrtf(ss * ssr) * cn;
                            const bool vis = scored && (val_ncc > A.thr);
                            if (vis) acc += (double)val_ncc;
                            word = __ballot_sync(FULL, vis);
                            if (A.ncc_out && v < A.V) A.ncc_out[h * A.V + v] = scored ? val_ncc : nanf("");
                        } else {
                        // ---- phase 2: lanes span the VIEWS.  Lanes L and L + 16 sum the first Q0 and the remaining
                        // sample quads of view L (pivot = the view's first sample, so low-variance windows keep full
                        // fp32 precision); one shuffle per quantity joins the two parts -- no butterfly.
                        const int half = bt, vbase = w32 * 32 + half * 16;
                        const int vj = lane & 15;
                        constexpr int Q0 = (NS4 + 1) / 2;                    // quads of the first part (all complete)
                        constexpr int R1 = NS - 4 * Q0;                      // samples of the second part
                        const bool lo = lane < 16;
                        const float pivot = s_val[vj][0];
                        const float4* rowq = reinterpret_cast<const float4*>(s_val[vj]) + (lo ? 0 : Q0);
                        const float4* refq = reinterpret_cast<const float4*>(s_dref) + (lo ? 0 : Q0);
                        float Sd = 0.0f, SSd = 0.0f, SAB = 0.0f;
#pragma unroll
                        for (int g = 0; g < Q0; ++g) {
                            float4 xq = make_float4(0.f, 0.f, 0.f, 0.f), dq = xq;
                            if (4 * g < R1 || lo) {                          // (quads past the second part's end are not read)
                                xq = rowq[g];
                                dq = refq[g];
                            }
                            const float xs[4] = {xq.x, xq.y, xq.z, xq.w}, ds[4] = {dq.x, dq.y, dq.z, dq.w};
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                if (4 * g + e < R1 || lo) {                  // compile-time true for the samples both parts have
                                    const float dd = xs[e] - pivot;
                                    Sd += dd;
                                    SSd = fmaf(dd, dd, SSd);
                                    SAB = fmaf(dd, ds[e], SAB);
                                }
                            }
                        }
                        Sd += __shfl_xor_sync(FULL, Sd, 16);
                        SSd += __shfl_xor_sync(FULL, SSd, 16);
                        SAB += __shfl_xor_sync(FULL, SAB, 16);
                        const int v = vbase + vj;
                        const float ss = SSd - Sd * Sd * (1.0f / NS);
                        const float cov = SAB - Sd * Sr * (1.0f / NS);
                        const bool scored = (v < A.V) && (v != r) && ((usable >> vj) & 1u) &&
                                            ((cand32 >> (half * 16 + vj)) & 1u) && (ss * (1.0f / NS) >= PMVS_VAR_MIN) &&
                                            (ssr * (1.0f / NS) >= PMVS_VAR_MIN);
                        const float val_ncc = cov * rsqrtf(ss * ssr) * cn;
                        const bool vis = scored && (val_ncc > A.thr) && (lane < 16);   // lanes 16..31 hold duplicates
                        if (vis) acc += (double)val_ncc;
                        word |= __ballot_sync(FULL, vis) << (half * 16);
                        if (A.ncc_out && v < A.V && lane < 16) A.ncc_out[h * A.V + v] = scored ? val_ncc : nanf("");
                        }
                    }
                } else if (A.ncc_out) {
                    for (int v = w32 * 32 + lane; v < min(A.V, w32 * 32 + 32); v += 32) A.ncc_out[h * A.V + v] = nanf("");
                }
                count += __popc(word);
                if (lane == 0 && A.vis_out) reinterpret_cast<uint32_t*>(A.vis_out)[h * 2 * mw + w32] = word;
            }
#pragma unroll
            for (int s = 16; s > 0; s >>= 1) acc += __shfl_xor_sync(FULL, acc, s);
            const double avg = count > 0 ? acc / (double)count : 0.0;
            if (lane == 0) {
                if (A.count_out) A.count_out[h] = count;
                if (A.avg_out) A.avg_out[h] = avg;
            }
            if (count >= A.bound && (float)avg > best_key) {   // strict '>': lowest index wins ties
                best_key = (float)avg;
                best_i = gi;
            }
        }
        if (lane == 0 && A.best_idx) {
            A.best_idx[set] = best_i;
            if (A.best_avg) A.best_avg[set] = best_i >= 0 ? (double)best_key : 0.0;
        }
    }
}

// Argmax over consecutive hypothesis sets for already scored batches (either mode):
// key = avg if count >= bound else -inf, lowest index on ties, -1 when none qualifies.
__global__ void __launch_bounds__(256) select_best_sets(const double* __restrict__ avg, const int32_t* __restrict__ count,
                                                        int64_t N, int group, int bound, int32_t* __restrict__ best_idx,
                                                        double* __restrict__ best_avg) {
    const int64_t n_sets = (N + group - 1) / group;
    for (int64_t set = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; set < n_sets; set += (int64_t)gridDim.x * blockDim.x) {
        double bk = -INFINITY;
        int bi = -1;
        for (int gi = 0; gi < group; ++gi) {
            const int64_t h = set * group + gi;
            if (h >= N) break;
            if (count[h] >= bound && avg[h] > bk) {
                bk = avg[h];
                bi = gi;
            }
        }
        best_idx[set] = bi;
        if (best_avg) best_avg[set] = bi >= 0 ? bk : 0.0;
    }
}

int mvs_launch_select_best(mvs_ctx* ctx, int64_t N, int group, const double* avg, const int32_t* count, int bound,
                           int32_t* best_idx, double* best_avg, cudaStream_t s) {
    if (N == 0) return MVS_OK;
    const int64_t n_sets = (N + group - 1) / group;
    int64_t blocks = (n_sets + 255) / 256;
    const int64_t cap = (int64_t)ctx->sm_count * 8;
    if (blocks > cap) blocks = cap;
    select_best_sets<<<(int)blocks, 256, 0, s>>>(avg, count, N, group, bound, best_idx, best_avg);
    ctx->launches++;
    MVS_CUDA_CHECK(cudaGetLastError());
    return MVS_OK;
}

// ---------------------------------------------------------------------------------
// Texture path set-up (once per context, on the first Mode B call): the gray images of all views
// are tiled into as few gather-enabled 2-D CUDA arrays ("atlases") as the device's texture size
// limit allows -- ONE for every BASELINE shape up to 128 x 1080p, three for 256 x 4K.  One
// texture object per view was measured 1.6x slower: every TLD4 of a warp then names a different
// texture header (47 views x 16 warps per SM) and the header cache thrashes.  A view is addressed
// by its tile offset; taps are validated against the view's own bounds, so tiles never bleed.
// ---------------------------------------------------------------------------------
int mvs_pmvs_prepare(mvs_ctx* ctx, cudaStream_t s) {
    if (ctx->pmvs_ready) return MVS_OK;
    const int V = ctx->V, H = ctx->H, W = ctx->W;
    int rc = MVS_OK;
    uint8_t* d_planar = nullptr;
    CamProj* hp = nullptr;
    double* h64 = nullptr;
    float* h32 = nullptr;
    cudaTextureObject_t* htex = nullptr;
    int max_w = 0, max_h = 0;
    if (cudaDeviceGetAttribute(&max_w, cudaDevAttrMaxTexture2DGatherWidth, ctx->device) != cudaSuccess ||
        cudaDeviceGetAttribute(&max_h, cudaDevAttrMaxTexture2DGatherHeight, ctx->device) != cudaSuccess || max_w < W ||
        max_h < H) {
        cudaGetLastError();
        mvs_set_error("Mode B set-up: a %d x %d view does not fit a gather texture (limit %d x %d)", W, H, max_w, max_h);
        return MVS_ERR_ARG;
    }
    if (const char* e = getenv("MVS_PMVS_MAX_TEX")) {       // test knob: a smaller limit forces several atlases
        const int lim = atoi(e);
        if (lim >= W && lim >= H) {
            max_w = max_w < lim ? max_w : lim;
            max_h = max_h < lim ? max_h : lim;
        }
    }
    const int fit_x = max_w / W, fit_y = max_h / H;
    const int tiles_x = V < fit_x ? V : fit_x;                         // tiles per atlas row
    const int rows_all = (V + tiles_x - 1) / tiles_x;                  // tile rows needed in total
    const int rows_per = rows_all < fit_y ? rows_all : fit_y;          // tile rows per atlas
    const int per_atlas = tiles_x * rows_per;
    const int n_atlas = (V + per_atlas - 1) / per_atlas;
    ctx->pmvs_n_atlas = n_atlas;
    ctx->pmvs_arrays = (cudaArray_t*)calloc(n_atlas, sizeof(cudaArray_t));
    ctx->pmvs_atlas_tex = (cudaTextureObject_t*)calloc(n_atlas, sizeof(cudaTextureObject_t));
    ctx->pmvs_tex_host = (cudaTextureObject_t*)calloc(V, sizeof(cudaTextureObject_t));
    ctx->pmvs_off_host = (float*)calloc(2 * (size_t)V, sizeof(float));
    hp = (CamProj*)malloc(sizeof(CamProj) * V);
    h64 = (double*)calloc((size_t)PMVS_CAM64_FIELDS * V, sizeof(double));
    h32 = (float*)calloc((size_t)PMVS_CAM32_FIELDS * V, sizeof(float));
    if (!ctx->pmvs_arrays || !ctx->pmvs_atlas_tex || !ctx->pmvs_tex_host || !ctx->pmvs_off_host || !hp || !h64 || !h32) {
        rc = MVS_ERR_NOMEM;
        goto done;
    }
    htex = ctx->pmvs_tex_host;
    if (cudaMalloc(&d_planar, (size_t)V * H * W) != cudaSuccess) {
        cudaGetLastError();
        mvs_set_error("Mode B set-up: cudaMalloc of %zu planar bytes failed", (size_t)V * H * W);
        rc = MVS_ERR_NOMEM;
        goto done;
    }
    if ((rc = mvs_launch_unpack_gray(ctx, d_planar, s)) != MVS_OK) goto done;
    {
        cudaChannelFormatDesc fmt = cudaCreateChannelDesc<unsigned char>();
        for (int a = 0; a < n_atlas; ++a) {
            const int v_lo = a * per_atlas, v_hi = (v_lo + per_atlas < V) ? v_lo + per_atlas : V;
            const int rows = (v_hi - v_lo + tiles_x - 1) / tiles_x;
            if (cudaMallocArray(&ctx->pmvs_arrays[a], &fmt, (size_t)tiles_x * W, (size_t)rows * H, cudaArrayTextureGather) !=
                cudaSuccess) {
                mvs_set_error("Mode B set-up: cudaMallocArray (%d x %d atlas) failed: %s", tiles_x * W, rows * H,
                              cudaGetErrorString(cudaGetLastError()));
                rc = MVS_ERR_NOMEM;
                goto done;
            }
            cudaResourceDesc rd;
            memset(&rd, 0, sizeof(rd));
            rd.resType = cudaResourceTypeArray;
            rd.res.array.array = ctx->pmvs_arrays[a];
            cudaTextureDesc td;
            memset(&td, 0, sizeof(td));
            td.addressMode[0] = td.addressMode[1] = cudaAddressModeClamp;
            td.filterMode = cudaFilterModePoint;
            td.readMode = cudaReadModeNormalizedFloat;
            td.normalizedCoords = 0;
            if (cudaCreateTextureObject(&ctx->pmvs_atlas_tex[a], &rd, &td, nullptr) != cudaSuccess) {
                mvs_set_error("Mode B set-up: cudaCreateTextureObject failed: %s", cudaGetErrorString(cudaGetLastError()));
                rc = MVS_ERR_CUDA;
                goto done;
            }
            for (int v = v_lo; v < v_hi; ++v) {
                const int tx = (v - v_lo) % tiles_x, ty = (v - v_lo) / tiles_x;
                if (cudaMemcpy2DToArrayAsync(ctx->pmvs_arrays[a], (size_t)tx * W, (size_t)ty * H, d_planar + (size_t)v * H * W, W, W,
                                             H, cudaMemcpyDeviceToDevice, s) != cudaSuccess) {
                    mvs_set_error("Mode B set-up: copy to atlas failed: %s", cudaGetErrorString(cudaGetLastError()));
                    rc = MVS_ERR_CUDA;
                    goto done;
                }
                htex[v] = ctx->pmvs_atlas_tex[a];
                ctx->pmvs_off_host[2 * v] = (float)(tx * W);
                ctx->pmvs_off_host[2 * v + 1] = (float)(ty * H);
            }
        }
    }
    if (cudaMemcpy(hp, ctx->d_cam, sizeof(CamProj) * V, cudaMemcpyDeviceToHost) != cudaSuccess) { rc = MVS_ERR_CUDA; goto done; }
    for (int v = 0; v < V; ++v) {
        for (int i = 0; i < 9; ++i) {
            h64[(size_t)i * V + v] = hp[v].r[i];
            h32[(size_t)i * V + v] = (float)hp[v].r[i];
        }
        for (int i = 0; i < 3; ++i) h64[(size_t)(9 + i) * V + v] = hp[v].t[i];
        const double k4[4] = {hp[v].fx, hp[v].fy, hp[v].cx, hp[v].cy};
        for (int i = 0; i < 4; ++i) {
            h64[(size_t)(12 + i) * V + v] = k4[i];
            h32[(size_t)(9 + i) * V + v] = (float)k4[i];
        }
    }
    if (cudaMalloc(&ctx->d_pmvs_off, sizeof(float) * 2 * V) != cudaSuccess ||
        cudaMalloc(&ctx->d_pmvs_tex, sizeof(cudaTextureObject_t) * V) != cudaSuccess ||
        cudaMalloc(&ctx->d_pmvs_camf, (sizeof(double) * PMVS_CAM64_FIELDS + sizeof(float) * PMVS_CAM32_FIELDS) * (size_t)V) != cudaSuccess) {
        cudaGetLastError();
        rc = MVS_ERR_NOMEM;
        goto done;
    }
    if (cudaMemcpyAsync(ctx->d_pmvs_off, ctx->pmvs_off_host, sizeof(float) * 2 * V, cudaMemcpyHostToDevice, s) != cudaSuccess ||
        cudaMemcpyAsync(ctx->d_pmvs_tex, htex, sizeof(cudaTextureObject_t) * V, cudaMemcpyHostToDevice, s) != cudaSuccess ||
        cudaMemcpyAsync(ctx->d_pmvs_camf, h64, sizeof(double) * PMVS_CAM64_FIELDS * (size_t)V, cudaMemcpyHostToDevice, s) != cudaSuccess ||
        cudaMemcpyAsync((double*)ctx->d_pmvs_camf + (size_t)PMVS_CAM64_FIELDS * V, h32, sizeof(float) * PMVS_CAM32_FIELDS * (size_t)V,
                        cudaMemcpyHostToDevice, s) != cudaSuccess ||
        cudaStreamSynchronize(s) != cudaSuccess) {
        mvs_set_error("Mode B set-up: upload failed: %s", cudaGetErrorString(cudaGetLastError()));
        rc = MVS_ERR_CUDA;
        goto done;
    }
    ctx->pmvs_ready = 1;
    {
        static int64_t serial = 0;                 // distinguishes a new context allocated at a recycled address
        ctx->pmvs_serial = ++serial;
    }
done:
    if (d_planar) cudaFree(d_planar);
    free(hp);
    free(h64);
    free(h32);
    return rc;
}

void mvs_pmvs_release(mvs_ctx* ctx) {
    if (ctx->pmvs_atlas_tex) {
        for (int a = 0; a < ctx->pmvs_n_atlas; ++a)
            if (ctx->pmvs_atlas_tex[a]) cudaDestroyTextureObject(ctx->pmvs_atlas_tex[a]);
        free(ctx->pmvs_atlas_tex);
        ctx->pmvs_atlas_tex = nullptr;
    }
    if (ctx->pmvs_arrays) {
        for (int a = 0; a < ctx->pmvs_n_atlas; ++a)
            if (ctx->pmvs_arrays[a]) cudaFreeArray(ctx->pmvs_arrays[a]);
        free(ctx->pmvs_arrays);
        ctx->pmvs_arrays = nullptr;
    }
    free(ctx->pmvs_tex_host);
    ctx->pmvs_tex_host = nullptr;
    free(ctx->pmvs_off_host);
    ctx->pmvs_off_host = nullptr;
    ctx->pmvs_n_atlas = 0;
    if (ctx->d_pmvs_off) cudaFree(ctx->d_pmvs_off);
    ctx->d_pmvs_off = nullptr;
    if (ctx->d_pmvs_tex) cudaFree(ctx->d_pmvs_tex);
    if (ctx->d_pmvs_camf) cudaFree(ctx->d_pmvs_camf);
    ctx->d_pmvs_tex = nullptr;
    ctx->d_pmvs_camf = nullptr;
    ctx->pmvs_ready = 0;
}

int mvs_launch_score_pmvs(mvs_ctx* ctx, int64_t N, const double* c, const double* nrm, const int32_t* ref,
                          const uint64_t* cand, double thr, int mu, int flags, int group, int bound, uint64_t* vis,
                          double* avg, int32_t* count, double* xy, float* ncc, int32_t* best_idx, double* best_avg,
                          cudaStream_t s) {
    if (N == 0) return MVS_OK;
    int rc;
    if ((rc = mvs_pmvs_prepare(ctx, s)) != MVS_OK) return rc;
    // the constant-memory handle table belongs to the context that scored last on this device
    static const mvs_ctx* table_owner[64] = {nullptr};
    static int64_t table_serial[64] = {0};
    if (ctx->device < 64 && (table_owner[ctx->device] != ctx || table_serial[ctx->device] != ctx->pmvs_serial)) {
        // another context's Mode B kernel may still be reading the per-device table on another stream: drain the
        // device before handing the table over (only when contexts alternate on one GPU; a CUDA graph captured for
        // one context must not be replayed after another context scored in Mode B -- re-capture it)
        if (table_owner[ctx->device] != nullptr) MVS_CUDA_CHECK(cudaDeviceSynchronize());
        MVS_CUDA_CHECK(cudaMemcpyToSymbolAsync(c_tex, ctx->pmvs_tex_host, sizeof(cudaTextureObject_t) * ctx->V, 0,
                                               cudaMemcpyHostToDevice, s));
        MVS_CUDA_CHECK(cudaMemcpyToSymbolAsync(c_off, ctx->pmvs_off_host, sizeof(float) * 2 * ctx->V, 0, cudaMemcpyHostToDevice, s));
        table_owner[ctx->device] = ctx;
        table_serial[ctx->device] = ctx->pmvs_serial;
    }
    PmvsArgs A;
    A.cams = ctx->d_cam;
    A.cam64 = (const double*)ctx->d_pmvs_camf;
    A.cam32 = (const float*)((const double*)ctx->d_pmvs_camf + (size_t)PMVS_CAM64_FIELDS * ctx->V);
    A.tex = (const cudaTextureObject_t*)ctx->d_pmvs_tex;
    A.off = (const float2*)ctx->d_pmvs_off;
    A.V = ctx->V; A.H = ctx->H; A.W = ctx->W; A.cam_stride = ctx->V;
    A.flags = flags; A.group = group; A.bound = bound; A.thr = (float)thr;
    A.c = c; A.nrm = nrm; A.ref = ref; A.cand = cand;
    A.vis_out = vis; A.avg_out = avg; A.count_out = count; A.xy_out = xy; A.ncc_out = ncc;
    A.best_idx = best_idx; A.best_avg = best_avg;
    const int g = group > 1 ? group : 1;
    const int64_t n_sets = (N + g - 1) / g;
    int64_t blocks = n_sets;                               // one warp per CTA
    const int64_t cap = (int64_t)ctx->sm_count * 32 * 4;
    if (blocks > cap) blocks = cap;
    const int pslot = (int)(ctx->prof_n % MVS_PROF_RING);
    if (ctx->profile) MVS_CUDA_CHECK(cudaEventRecord(ctx->prof_ev[2 * pslot], s));
    static int minb = -1;                                  // MVS_K2_MINB: resident warps per SM (tuning knob)
    if (minb < 0) {
        const char* e = getenv("MVS_K2_MINB");
        minb = e ? atoi(e) : 32;                           // 32 one-warp CTAs per SM (64 registers) measured best
    }
    const bool one = ctx->pmvs_n_atlas == 1;
    // views per reduction batch: 32 up to mu = 5 (3.44 vs 3.66 ms per 2^20 hypotheses at mu = 5), 16 from mu = 7 on, where the
    // 32-row sample buffer costs a quarter of the resident warps (6.25 vs 5.99 ms).  MVS_K2_VB overrides (tuning knob).
    static int vb_env = -1;
    if (vb_env < 0) {
        const char* e = getenv("MVS_K2_VB");
        vb_env = e ? atoi(e) : 0;
    }
    const int vb = vb_env == 16 || vb_env == 32 ? vb_env : (mu <= 5 ? 32 : 16);
#define PMVS_GO(MU_, RA_, MB_, OA_, VB_)                                                                       \
    do {                                                                                                   \
        static bool carve = false;                         /* shared memory of 32 one-warp CTAs needs the largest carve-out */ \
        if (!carve) {                                                                                      \
            cudaFuncSetAttribute(ncc_score_pmvs<MU_, RA_, MB_, OA_, VB_>, cudaFuncAttributePreferredSharedMemoryCarveout, \
                                 cudaSharedmemCarveoutMaxShared);                                          \
            carve = true;                                                                                  \
        }                                                                                                  \
        ncc_score_pmvs<MU_, RA_, MB_, OA_, VB_><<<(int)blocks, 32, 0, s>>>(A, N);                          \
    } while (0)
#define PMVS_LAUNCH(MU_)                                                                                   \
    case MU_:                                                                                              \
        if (flags & MVS_PMVS_REDUCE_TO_REFEXACT)                                                           \
            PMVS_GO(MU_, true, 16, false, 16);                                                             \
        else if (minb == 16)                                                                               \
            PMVS_GO(MU_, false, 16, false, 16);                                                            \
        else if (minb == 24)                                                                               \
            PMVS_GO(MU_, false, 24, false, 16);                                                            \
        else if (one && vb == 32)                                                                          \
            PMVS_GO(MU_, false, 32, true, 32);                                                             \
        else if (one)                                                                                      \
            PMVS_GO(MU_, false, 32, true, 16);                                                             \
        else if (vb == 32)                                                                                 \
            PMVS_GO(MU_, false, 32, false, 32);                                                            \
        else                                                                                               \
            PMVS_GO(MU_, false, 32, false, 16);                                                            \
        break;
    switch (mu) {
        PMVS_LAUNCH(3) PMVS_LAUNCH(5) PMVS_LAUNCH(7) PMVS_LAUNCH(9) PMVS_LAUNCH(11)
        default: mvs_set_error("Mode B grid size mu = %d not supported (3, 5, 7, 9, 11)", mu); return MVS_ERR_ARG;
    }
#undef PMVS_GO
#undef PMVS_LAUNCH
    if (ctx->profile) {
        MVS_CUDA_CHECK(cudaEventRecord(ctx->prof_ev[2 * pslot + 1], s));
        ctx->prof_n++;
    }
    ctx->launches++;
    MVS_CUDA_CHECK(cudaGetLastError());
    return MVS_OK;
}
