// Shared definitions of the B200 NCC scorer library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "mvs_ncc.h"

#define MVS_ABI_VERSION 1
#define MVS_GROUP_PAD 8
#define MVS_MAX_PEERS 16
#define MVS_PROF_RING 64
#define MVS_ANCHOR_INVALID 0xffffffffu
#define MVS_SORT_MIN 8192        // batches smaller than this are scored in input order

// One view's projection parameters as the scorer reads them (128 B, fp64).
// r = Rodrigues round trip of the file rotation (utils.py:242-243).
struct __align__(16) CamProj {
    double r[9];
    double t[3];
    double fx, fy, cx, cy;
};

// Expansion-side geometry of one view (MVS2.py:188-189, 334-358): FILE rotation and
// centre C = -R^T t.
struct __align__(16) CamGeom {
    double rf[9];
    double C[3];
    double fx, fy, cx, cy;
};

struct mvs_ctx {
    int device;
    int V, H, W;
    // Resident gray stack, VIEW-INTERLEAVED: u8 [H][G][Vp][4] -- four consecutive pixels of
    // one view form a 32-bit word, the words of all views for the same four pixels are
    // adjacent.  Byte (v, row, col) lives at row*rowpitch + (col>>2)*gstride + 4*v + (col&3).
    // Mode A samples every view at the SAME (row, col) (MVS2.py:68), so one window row of all
    // views is a single contiguous run of <= NG*gstride bytes.
    int Vp;               // V rounded up to a multiple of 4 (padding views are zero)
    int Q;                // Vp / 4: 16-byte quads (4 views x 4 pixels) per pixel group
    int G;                // pixel groups per row, ceil(W/4) + MVS_GROUP_PAD zero groups
    int64_t gstride;      // bytes per pixel group = 4 * Vp
    int64_t rowpitch;     // bytes per image row = G * gstride
    uint8_t* d_gray;      // [H, G, Vp, 4] + 256 B tail pad
    // Per-anchor window sums of every view for half window maps_wid (built on first use):
    //   smap u16 [H][W][Vp] = sum(w),  vmap u32 [H][W][Vp] = n*sum(w*w) - sum(w)^2  (exact)
    uint16_t* d_smap;
    uint32_t* d_vmap;
    size_t smap_bytes, vmap_bytes;
    int maps_wid;
    // spatial binning scratch (bin.cu): hypotheses ordered by 8x8-pixel anchor tile
    int32_t* d_bin_hist;   // [tiles + 2]
    int32_t* d_bin_key;    // [N]
    int32_t* d_bin_rank;   // [N]
    void* d_bin_entry;     // [N] uint2 (hypothesis index, anchor) per ordered position
    uint32_t* d_bin_anchor;   // [N] row<<16 | col per hypothesis (MVS_ANCHOR_INVALID = rejected)
    int64_t* d_bin_scan;
    size_t bin_hist_bytes, bin_key_bytes, bin_rank_bytes, bin_entry_bytes, bin_anchor_bytes,
        bin_scan_bytes;
    // optional CUDA-event bracket around the scoring kernel (mvs_profile_enable)
    int profile;
    int probe_gather;     // mvs_probe_gather: the scoring launch runs the loads-only ceiling probe instead of K1
    cudaEvent_t prof_ev[2 * MVS_PROF_RING];
    int64_t prof_n;       // scoring kernels bracketed since mvs_profile_enable(1)
    // Mode B texture path (ncc_pmvs.cu), created on the first Mode B call
    int pmvs_ready;
    int64_t pmvs_serial;
    int pmvs_n_atlas;
    cudaArray_t* pmvs_arrays;              // [n_atlas] gather-enabled 2-D arrays, views tiled inside
    cudaTextureObject_t* pmvs_atlas_tex;   // [n_atlas]
    cudaTextureObject_t* pmvs_tex_host;    // [V] the atlas handle of each view
    float* pmvs_off_host;                  // [V,2] tile origin of each view
    void* d_pmvs_tex;                      // [V] cudaTextureObject_t
    void* d_pmvs_off;                      // [V] float2
    void* d_pmvs_camf;                     // cameras field-major: double [16][V] then float [13][V]
    CamProj* d_cam;       // [V]
    CamGeom* d_geom;      // [V]
    double* h_rrt;        // [V,9] host copy
    double* h_centres;    // [V,3]
    int sm_count;
    int64_t launches;
    // host-mode staging (grown on demand)
    void* d_stage;
    size_t stage_bytes;
    cudaStream_t own_stream;
    // host-mode pipeline (mvs_score_batch with host buffers): copy-in / compute / copy-out streams,
    // three staging slots
    cudaStream_t in_stream, out_stream;
    cudaEvent_t ev_in[3], ev_done[3], ev_out[3];
    // compaction scratch
    int32_t* d_tiles;
    size_t tile_bytes;
    // cell table (CellTable, MVS2.py:80-120) and round state
    int cell_size, wc, hc;
    uint8_t* d_cells;                 // [V, wc, hc], 1 = vacant
    size_t cells_bytes;
    unsigned long long* d_claim;      // [V, wc, hc] epoch-tagged slot claims
    size_t claim_bytes;
    unsigned epoch;
    int32_t* d_counts;
    size_t counts_bytes;
    int64_t* d_scan;
    size_t scan_bytes;
    int64_t n_cand;
    int64_t* cand_slot;
    int64_t* cand_parent;
    double* cand_c;
    double* cand_n;
    int32_t* cand_ref;
    int32_t* cand_px;
    uint64_t* cand_vis;
    double* cand_avg;
    int32_t* cand_count;
    double* cand_xy;
    uint8_t* cand_gate;
    size_t cand_cap[11];
    // fused round pipeline (mvs_expand_run) and exchange (exchange.cu)
    uint8_t* d_out;                   // accepted patch records of all rounds, in commit order
    size_t out_bytes;
    int64_t n_out;
    uint8_t* d_inbox;                 // world == 1: the local "inbox" of the minimal wire
    size_t inbox_bytes;
    int64_t* d_round_n;
    size_t round_n_bytes;
    void* d_barrier_state;            // {u64 reserved, int error}
    void* d_ticket;                   // last-CTA tickets of publish_count_scan (one per part)
    // overlapped exchange (mvs_exchange_set_parts): one K1 launch + publish per position range, range k on side stream k
    // (descending priority), the last range on the caller's stream
    int xparts;                       // 1 = off
    int k1_share;                     // > 1 while the range launches of one batch are being enqueued (K1's grid cap is shared)
    int64_t xparts_min;               // smallest shard that is partitioned
    uint8_t* d_bin_part;              // [N] position range of every hypothesis of the current ordered batch
    size_t bin_part_bytes;
    cudaStream_t x_side[8];
    cudaEvent_t x_ev[8], x_fork;
    uint8_t* d_live;                  // [F, V] surviving diagonals of every (frontier patch, view), count pass -> emit pass
    size_t live_bytes;
    void* h_pinned;                   // small pinned read-back area
    cudaEvent_t ev_round[2];
};

int mvs_p2p_barrier_failed(mvs_ctx* ctx, void* stream);

void mvs_set_error(const char* fmt, ...);

#define MVS_CUDA_CHECK(expr)                                                               \
    do {                                                                                   \
        cudaError_t _e = (expr);                                                           \
        if (_e != cudaSuccess) {                                                           \
            mvs_set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, \
                          __LINE__);                                                       \
            return MVS_ERR_CUDA;                                                           \
        }                                                                                  \
    } while (0)

// kernels / launchers implemented in the .cu files
int mvs_launch_gray(mvs_ctx* ctx, const uint8_t* d_rgb, int v0, int nv, cudaStream_t s);
int mvs_launch_unpack_gray(mvs_ctx* ctx, uint8_t* d_planar, cudaStream_t s);
int mvs_launch_score_refexact(mvs_ctx* ctx, int64_t N, const double* c, const int32_t* ref, double thr, int wid,
                              uint64_t* vis, double* avg, int32_t* count, double* xy, float* ncc, cudaStream_t s);
int mvs_build_window_maps(mvs_ctx* ctx, int wid, cudaStream_t s);
// project + validate every hypothesis (writes xy and, for rejected ones, the empty result),
// optionally order them by anchor tile; leaves anchors / ordered entries in ctx->d_bin_*
// part != nullptr (ordered batches only): part[h] = ordered position of hypothesis h / part_size
int mvs_bin_hypotheses(mvs_ctx* ctx, int64_t N, const double* c, const int32_t* ref, int wid, bool sort, uint64_t* vis,
                       double* avg, int32_t* count, double* xy, float* ncc, cudaStream_t s, uint8_t* part = nullptr,
                       int64_t part_size = 1);
// K1 alone over positions [p0, p1) of the batch mvs_bin_hypotheses left in the context (sort as passed there)
int mvs_launch_k1(mvs_ctx* ctx, int64_t N, const int32_t* ref, double thr, int wid, uint64_t* vis, double* avg,
                  int32_t* count, float* ncc, bool sort, int64_t p0, int64_t p1, cudaStream_t s);
// score + publish of one shard; with mvs_exchange_set_parts(P > 1) every position range of the ordered batch is its own
// K1 launch on a stream of its own priority, and the accept decisions of range k travel over NVLink while the ranges
// behind it are still being scored (exchange.cu)
int mvs_launch_score_publish(mvs_ctx* ctx, int64_t N, const double* c, const int32_t* ref, double thr, int wid, uint64_t* vis,
                             double* avg, int32_t* count, double* xy, const uint8_t* gate, int bound, void* const* peer_inbox,
                             int rank, int world, int64_t capacity, int parity, cudaStream_t s);
#define MVS_MAX_PARTS 8
int mvs_launch_score_pmvs(mvs_ctx* ctx, int64_t N, const double* c, const double* nrm, const int32_t* ref,
                          const uint64_t* cand, double thr, int mu, int flags, int group, int bound, uint64_t* vis,
                          double* avg, int32_t* count, double* xy, float* ncc, int32_t* best_idx, double* best_avg,
                          cudaStream_t s);
int mvs_launch_select_best(mvs_ctx* ctx, int64_t N, int group, const double* avg, const int32_t* count, int bound,
                           int32_t* best_idx, double* best_avg, cudaStream_t s);
void mvs_pmvs_release(mvs_ctx* ctx);
int mvs_launch_compact_p2p(mvs_ctx* ctx, int64_t N, int64_t index_base, const double* c, const double* nrm,
                           const int32_t* ref, const uint64_t* vis, const double* avg, const int32_t* count,
                           const double* xy, const uint8_t* gate, int bound, void* const* peer_records,
                           int64_t* const* peer_counts, int rank, int world, int wire, int64_t capacity,
                           const int64_t* index_arr, const int32_t* px, cudaStream_t s);
int mvs_launch_records_expand(mvs_ctx* ctx, const void* wire, int64_t n, void* records, cudaStream_t s);
int mvs_launch_compact(mvs_ctx* ctx, int64_t N, int64_t index_base, const double* c, const double* nrm, const int32_t* ref,
                       const uint64_t* vis, const double* avg, const int32_t* count, const double* xy, const uint8_t* gate,
                       int bound, void* records, int64_t capacity, int64_t* d_n_out, const int64_t* index_arr,
                       const int32_t* px, cudaStream_t s);
