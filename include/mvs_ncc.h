/*
 * mvs_ncc.h -- C ABI of the B200 photo-consistency (NCC) scorer.
 *
 * This is the drop-in boundary for the one hot path this repo accelerates: the
 * multi-view-stereo scoring of the reference's MVS2.py.  The reference is pure
 * Python and has no FFI of its own; each entry point below names the reference
 * interface (file:line under the reference root) whose work it replaces.  The
 * Python binding a maintainer adds is a ctypes stub -- see INTEGRATION.md.
 *
 * Conventions
 *   - return value: 0 = MVS_OK, negative = error; mvs_last_error() gives the text
 *     for the calling thread.  No exceptions cross the ABI.
 *   - a context owns, on ONE GPU: the gray image stack, the cameras, the cell
 *     table and scratch buffers.  Callers own every I/O buffer they pass.
 *   - every I/O pointer of a call is either a host pointer or a device pointer,
 *     chosen per call by `on_device`.  Host mode copies in, runs, copies out and
 *     synchronises before returning; device mode only enqueues work on `stream`
 *     (a cudaStream_t passed as void*; NULL = the legacy default stream).
 *   - one context per GPU; calls on one context must be serialised by the caller
 *     (thread-compatible, not thread-safe).  Device-mode calls of one context share
 *     its scratch buffers (ordering scratch, window maps, candidate arrays): enqueue
 *     them on ONE stream, or order the streams with events -- two streams running
 *     calls of the same context concurrently is a data race.  A CUDA graph captured
 *     from such calls replays with the buffer addresses of capture time: run the
 *     sequence once eagerly first (scratch buffers grow on demand) and re-capture
 *     after any call that handles a larger batch.
 *   - there is NO CPU fallback: every entry point fails with MVS_ERR_CUDA when no
 *     sm_100 device is usable.
 */
#ifndef MVS_NCC_H
#define MVS_NCC_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MVS_OK 0
#define MVS_ERR_ARG (-1)
#define MVS_ERR_CUDA (-2)
#define MVS_ERR_NOMEM (-3)
#define MVS_ERR_STATE (-4)

/* scoring modes for mvs_score_batch */
#define MVS_MODE_REFEXACT 0 /* "Mode A": literal behaviour of MVS2.py:62-77 */
#define MVS_MODE_PMVS 1     /* "Mode B": per-view projection, oriented mu x mu bilinear grid */

/* flags for mvs_score_pmvs */
#define MVS_PMVS_REDUCE_TO_REFEXACT 1 /* collapse Mode B onto the reference's behaviour: the reference
                                         camera for every view (MVS2.py:68), image-aligned integer lattice
                                         around (int(x), int(y)), no interpolation, the bounds rule of
                                         HarrisFeatures.py:128 -- must agree with MVS_MODE_REFEXACT */

/* record formats on the wire between GPUs (mvs_compact_accepted_p2p, mvs_records_expand) */
#define MVS_WIRE_FULL 0    /* mvs_patch_record + visible mask: mvs_record_bytes() */
#define MVS_WIRE_COMPACT 1 /* only what a peer cannot recompute -- c[3] f64, avg f64, index i64, ref i32,
                              px[2] i32, pad i32, visible mask: 56 + 8*ceil(V/64) bytes.  Valid for patches
                              whose normal is the unit vector from c to their reference camera and whose
                              x, y is the reference projection of c, i.e. every patch the reference's MVS
                              creates (MVS2.py:247, 357-358, 74); mvs_records_expand rebuilds n, xy and
                              count bit-identically on the receiving GPU */

typedef struct mvs_ctx mvs_ctx;

/* ABI version of this header (checked by the Python loader). */
int mvs_abi_version(void);

/* Text of the last error raised on the calling thread ("" if none). */
const char* mvs_last_error(void);

/*
 * Load the image stack and the cameras into HBM once.
 * Replaces: main.py:7-20 (read_imgs: list of H x W x 3 uint8 RGB arrays),
 *           utils.py:56-81 (read_pars: K 3x3, R 3x3, t 3x1 per view),
 *           HarrisFeatures.py:124-125 (the per-call full-image gray conversion,
 *           done here once on the device, bit-exact to cv2's BGR2GRAY applied to
 *           an RGB array), utils.py:242 (cv2.Rodrigues round trip of R).
 *   rgb  [V,H,W,3] uint8, host or device (rgb_on_device)
 *   K    [V,9] row-major, only K[0],K[4],K[2],K[5] (fx,fy,cx,cy) enter the path
 *   R    [V,9] row-major rotation as written in the *_par.txt file
 *   Rrt  [V,9] the Rodrigues round trip of R computed by the caller (e.g. with
 *        cv2, which makes projections bit-identical to utils.py:241-244), or NULL
 *        to let the library compute it (agrees with cv2 to ~1e-14 per entry)
 *   t    [V,3]
 */
int mvs_create(mvs_ctx** out, int device, int V, int H, int W, const uint8_t* rgb, int rgb_on_device,
               const double* K, const double* R, const double* Rrt, const double* t);

int mvs_destroy(mvs_ctx* ctx);

/* Geometry of the resident stack: views, rows, cols, and the byte pitch of one image row of
 * the view-interleaved layout u8 [H][G][Vp][4] (all views of a row are adjacent). */
int mvs_get_info(const mvs_ctx* ctx, int* V, int* H, int* W, int64_t* pitch);

/* Copy the resident gray stack back to the host as a dense [V,H,W] uint8 array. */
int mvs_download_gray(mvs_ctx* ctx, uint8_t* out_host);

/* Cameras as the device uses them: Rrt [V,9] and centres C = -R^T t [V,3] (file R,
 * MVS2.py:188-189).  Either pointer may be NULL. */
int mvs_get_cameras(const mvs_ctx* ctx, double* Rrt_host, double* centres_host);

/*
 * Score N patch hypotheses.
 * Replaces: MyPatch.photo_consistenecy_test (MVS2.py:62-77) with ctNcc
 *           (MVS2.py:39-43), getDescFeatures (HarrisFeatures.py:116-133) and
 *           projectPoint (utils.py:241-244), for a whole batch.
 *   mode      MVS_MODE_REFEXACT | MVS_MODE_PMVS
 *   c   [N,3] patch centres (patch.c)
 *   nrm [N,3] patch normals (patch.n); unused (may be NULL) in MVS_MODE_REFEXACT,
 *             exactly as the reference never reads it
 *   ref [N]   reference view index (patch.R)
 *   min_ncc   MIN_NCC (0.4 at MVS2.py:255, 0.7 at MVS2.py:362); strict '>'
 *   wid       half window (reference: hard-coded 5); window = (2*wid+1)^2 pixels.
 *             In MVS_MODE_PMVS the grid is mu x mu with mu = 2*wid+1.
 * outputs (any may be NULL except vis_mask and count)
 *   vis_mask [N, ceil(V/64)] uint64, bit v set <=> view v would be appended to
 *             patch.V (MVS2.py:72-74)
 *   avg  [N]  patch.avg_ncc_score (MVS2.py:73,75-76): mean NCC over visible views, 0 if none
 *   count [N] patch.visible_ct()
 *   xy  [N,2] the unrounded reference-view projection (x = column, y = row) stored
 *             in every patch.V entry (MVS2.py:74)
 *   ncc [N,V] float32, every per-view NCC (NaN where the reference computes none
 *             or NaN: reference view itself, zero-variance window, out of bounds);
 *             for parity tests, NULL on the hot path
 */
int mvs_score_batch(mvs_ctx* ctx, int mode, int64_t N, const double* c, const double* nrm, const int32_t* ref,
                    double min_ncc, int wid, uint64_t* vis_mask, double* avg, int32_t* count, double* xy, float* ncc,
                    int on_device, void* stream);

/*
 * Mode B ("PMVS-style") scoring with optional on-chip selection over hypothesis sets.
 * Replaces: nothing in the reference -- MVS2.py:62-77 never projects into the other views,
 *           never interpolates and never reads patch.n.  This is the scorer
 *           BASELINE.json's north_star describes; its normative spec is oracle/mode_b.py
 *           and it is reported as an extension.  Same outputs as mvs_score_batch.
 *   c, nrm, ref   as mvs_score_batch (nrm required unless MVS_PMVS_REDUCE_TO_REFEXACT)
 *   cand [N, ceil(V/64)] optional candidate-view mask (NULL = every view)
 *   mu        grid size (3, 5, 7, 9 or 11): mu x mu samples, one grid step ~ one pixel in
 *             the reference view, every view sampled bilinearly through ITS OWN camera
 *   group     > 1: hypotheses [s*group, (s+1)*group) form selection set s (e.g. the depth x
 *             normal variants of one cell); best_idx[s] = index inside the set of the member
 *             with the highest avg among those with count >= bound (lowest index on ties,
 *             -1 when none), best_avg[s] = its avg.  The per-hypothesis outputs (vis_mask,
 *             avg, count, xy, ncc) may then all be NULL: only one record per set leaves the SM.
 *   best_idx [ceil(N/group)] int32, best_avg [ceil(N/group)] float64 (either may be NULL)
 */
int mvs_score_pmvs(mvs_ctx* ctx, int64_t N, const double* c, const double* nrm, const int32_t* ref,
                   const uint64_t* cand, double min_ncc, int mu, int flags, int group, int bound, uint64_t* vis_mask,
                   double* avg, int32_t* count, double* xy, float* ncc, int32_t* best_idx, double* best_avg,
                   int on_device, void* stream);

/*
 * Selection over hypothesis sets for an already scored batch of either mode (SURVEY appendix
 * A.7; the reference has no refinement/argmax step -- north_star extension): same rule as the
 * `group` argument of mvs_score_pmvs.  DEVICE pointers; enqueued on `stream`.
 */
int mvs_select_best(mvs_ctx* ctx, int64_t N, int group, const double* avg, const int32_t* count, int bound,
                    int32_t* best_idx, double* best_avg, void* stream);

/*
 * One accepted patch as exchanged between GPUs after a round: the fields of the
 * reference's MyPatch (MVS2.py:45-60) that later rounds read.  A record is
 * sizeof(mvs_patch_record) + 8*ceil(V/64) bytes: the struct followed by the visible
 * set as a bit mask (patch.V; every entry shares xy, MVS2.py:74).
 */
typedef struct mvs_patch_record {
    double c[3];   /* patch.c */
    double n[3];   /* patch.n */
    double xy[2];  /* x, y stored in every patch.V entry */
    double avg;    /* patch.avg_ncc_score */
    int32_t ref;   /* patch.R */
    int32_t count; /* patch.visible_ct() */
    int64_t index; /* global candidate index / expansion slot id (deterministic order across GPUs) */
    int32_t px[2]; /* int(u), int(v) of the cell centre the patch was cast through: the pixel
                      CellTable.get_color reads (MVS2.py:116-117,357); -1 when not applicable */
} mvs_patch_record;

/* Bytes per record for this context: sizeof(mvs_patch_record) + 8*ceil(V/64). */
int mvs_record_bytes(const mvs_ctx* ctx);

/*
 * Keep the hypotheses that pass the caller-side accept test and pack them, in input
 * order, into patch records.
 * Replaces: the accept branch of patch_expansion (MVS2.py:369: visible_ct >= bound and
 *           neighbour / distance tests; MVS2.py:401-403: enqueue) and of the seed loop
 *           (MVS2.py:256-257), for a batch.
 *   gate [N] uint8 or NULL: extra per-hypothesis condition computed by the caller side
 *   keeps hypothesis i  <=>  count[i] >= bound && (gate == NULL || gate[i])
 *   records   capacity * mvs_record_bytes() bytes; n_out receives the number kept
 *             (records beyond capacity are counted but not written)
 * All pointers are DEVICE pointers (n_out included); work is enqueued on `stream`.
 */
int mvs_compact_accepted(mvs_ctx* ctx, int64_t N, int64_t index_base, const double* c, const double* nrm,
                         const int32_t* ref, const uint64_t* vis_mask, const double* avg, const int32_t* count,
                         const double* xy, const uint8_t* gate, int bound, void* records, int64_t capacity,
                         int64_t* n_out, void* stream);

/*
 * The same compaction FUSED with the per-round all-gather over peer memory (NVLink P2P stores):
 * every GPU of the box owns an inbox of world * capacity records and world int64 counts; this
 * rank's kept records are written, in input order, to region `rank` of EVERY inbox and its count to
 * slot `rank` of every count array, from inside the compaction kernel -- no collective call, no host
 * round trip for the payload size.  Replaces: mvs_compact_accepted followed by an all-gather of
 * (counts, records); the reference has no counterpart (single process).
 *   peer_records [world]  HOST array of DEVICE pointers, entry g = base of GPU g's inbox as mapped
 *                         into THIS process (CUDA IPC / symmetric memory; entry `rank` = local)
 *   peer_counts  [world]  likewise for the int64 count arrays
 *   wire                  MVS_WIRE_FULL | MVS_WIRE_COMPACT: record format written (region stride =
 *                         capacity * mvs_wire_bytes(ctx, wire)); nrm, xy are not read for COMPACT
 * The caller must order this call against the peers' use of their inboxes (a barrier across the
 * GPUs before the call: inboxes free; after it: records visible).  Other pointers: DEVICE.
 */
int mvs_compact_accepted_p2p(mvs_ctx* ctx, int64_t N, int64_t index_base, const double* c, const double* nrm,
                             const int32_t* ref, const uint64_t* vis_mask, const double* avg, const int32_t* count,
                             const double* xy, const uint8_t* gate, int bound, void* const* peer_records,
                             int64_t* const* peer_counts, int rank, int world, int wire, int64_t capacity, void* stream);

/* Bytes per record of a wire format for this context (0: unknown format). */
int mvs_wire_bytes(const mvs_ctx* ctx, int wire);

/*
 * Receiver side: rebuild n full patch records (mvs_record_bytes() each) from wire records.
 * DEVICE pointers; enqueued on `stream`.  MVS_WIRE_FULL is a plain copy.
 */
int mvs_records_expand(mvs_ctx* ctx, int wire, const void* wire_records, int64_t n, void* records, void* stream);

/*
 * Cell table (CellTable, MVS2.py:80-120): one vacancy byte per (view, x-cell, y-cell),
 * shape [V, ceil((W-1)/cs), ceil((H-1)/cs)] exactly as the reference's list of bool
 * arrays (MVS2.py:88), 1 = vacant.  table_host NULL = all vacant.
 */
int mvs_cells_init(mvs_ctx* ctx, int cell_size, const uint8_t* table_host);
int mvs_cells_shape(const mvs_ctx* ctx, int* cell_size, int* wc, int* hc);
int mvs_cells_download(mvs_ctx* ctx, uint8_t* table_host);

/*
 * CellTable.fill_with_point (MVS2.py:98-107) for a batch of patch records: clears cell
 * (v, floor(x/cs), floor(y/cs)) for every view v in the record's visible set.
 * records: DEVICE pointer to n records.
 */
int mvs_cells_fill(mvs_ctx* ctx, const void* records, int64_t n, void* stream);

/*
 * CellTable.filter_out_outlier (MVS2.py:132-158; shipped by the reference but disabled as "very very slow",
 * MVS2.py:280-281) for n patch records in INSERTION order (seeds first, then the expansion's patches in
 * commit order): removed[i] = 1 for every patch the reference's scan would delete from its Q lists.  One
 * thread per cell position walks the views in ascending order, exactly the reference's dependency (see
 * csrc/filter.cu).  A non-vacant cell with an empty list (the reference raises ZeroDivisionError there) is
 * skipped and counted.  The cell table itself is not modified (the reference does not modify it either).
 *   records DEVICE, removed DEVICE [n] uint8, counts_host HOST [2] = {patches removed, empty non-vacant cells}
 * Synchronises `stream`.
 */
int mvs_cells_filter(mvs_ctx* ctx, const void* records, int64_t n, uint8_t* removed, int64_t* counts_host, void* stream);

/*
 * cv2.triangulatePoints (utils.py:238-239) for n correspondences: x in view_a / view_b [n,2], projection
 * matrices P [V,12] = K [R|t] (utils.py:234-236) -> homogeneous X4 [n,4].  HOST pointers; synchronous.
 * OpenCV's algorithm (4x4 DLT system, one-sided Jacobi SVD) restated in fp64 on the device.
 */
int mvs_triangulate(mvs_ctx* ctx, int64_t n, const int32_t* view_a, const int32_t* view_b, const double* xa, const double* xb,
                    const double* P, double* X4);

/*
 * Seed-patch stage (MVS2.py:208-260) for SfM tracks given as a flat observation list: track t owns
 * obs[offsets[t] .. offsets[t+1]) rows of (view, x, y); its first observation is the reference, every
 * further one is triangulated against it, all candidates are scored in one batch at `min_ncc` (0.4,
 * MVS2.py:255) and the nearest candidate with >= bound visible views wins (heap key MVS2.py:14).  The
 * winners become patch records in track order (index = candidate number, i.e. observation index minus
 * (track + 1)); when a cell table is initialised their cells are filled (MVS2.py:258-259).
 *   offsets HOST [n_tracks+1], obs HOST [n_obs,3], P HOST [V,12]; seeds DEVICE (capacity n_tracks records);
 *   n_seeds HOST.  Synchronises `stream`.
 */
int mvs_seed_stage(mvs_ctx* ctx, int64_t n_tracks, const int64_t* offsets, const double* obs, const double* P, double min_ncc,
                   int wid, int bound, void* seeds, int64_t* n_seeds, void* stream);

/*
 * One synchronous expansion round, phase 1: candidate generation.
 * Replaces: patch_expansion's candidate loop (MVS2.py:328-361) for a whole frontier.
 * For every frontier patch f, every view v in its visible set and every diagonal
 * k = (di,dj) in ((-1,-1),(-1,1),(1,-1),(1,1)): slot s = (f*V + v)*4 + k is live when
 * cell (v, ci+di, cj+dj) is vacant in the round-start table (MVS2.py:333); among live
 * slots testing the same cell only the lowest s survives.  Survivors become candidates
 * in ascending slot order with c = ray/plane hit (MVS2.py:334-356, including the
 * reference's (ci+di, cj+di) pixel and "+C" ray quirks), n = (O-c)/|O-c|, ref = v.
 *   frontier  DEVICE pointer to F patch records
 *   n_candidates  HOST pointer; the call synchronises `stream` to produce it
 * The candidate arrays stay inside the context until the next generate call.
 */
int mvs_round_generate(mvs_ctx* ctx, const void* frontier, int64_t F, int64_t* n_candidates, void* stream);

/*
 * Phase 2: score candidates [begin, end) (this GPU's shard) with Mode A, apply the
 * accept test of MVS2.py:369 (visible_ct >= bound, is_patch_neighbor(thr 0.1),
 * distance(parent.c, c) < 0.05/scale) and pack the passing ones, in slot order, into
 * records (index = slot id).  records/n_out: DEVICE pointers.
 */
int mvs_round_score(mvs_ctx* ctx, const void* frontier, int64_t begin, int64_t end, double min_ncc, int wid, int bound,
                    double scale, void* records, int64_t capacity, int64_t* n_out, void* stream);

/*
 * Phase 2 with the exchange fused in (see mvs_compact_accepted_p2p): the passing records of this
 * GPU's shard are stored, in slot order, into region `rank` of EVERY GPU's inbox and their number
 * into slot `rank` of every count array; an empty shard publishes a zero count.  `capacity` (records
 * per region) must be >= end - begin.  Replaces: mvs_round_score + the all-gather of the records.
 */
int mvs_round_score_p2p(mvs_ctx* ctx, const void* frontier, int64_t begin, int64_t end, double min_ncc, int wid, int bound,
                        double scale, void* const* peer_records, int64_t* const* peer_counts, int rank, int world,
                        int wire, int64_t capacity, void* stream);

/*
 * Phase 3 (identical on every GPU, on the gathered records of ALL shards in ascending
 * slot order): drop a dj=+1 record whose dj=-1 sibling (slot-1) also passed -- the
 * `break` of MVS2.py:404 --, fill the cells of the kept ones (MVS2.py:401-402) and
 * write them as the next frontier.  All pointers DEVICE pointers.
 */
int mvs_round_commit(mvs_ctx* ctx, const void* records, int64_t n, void* next_frontier, int64_t* n_next, void* stream);

/*
 * Minimal wire between the GPUs of one box (exchange.cu).  In a round every GPU holds the same candidate
 * list and scores its shard; a peer only lacks, per candidate, whether it passed, its visible set and its
 * mean NCC.  A shard's wire is: header {int64 kept, int64 n}, one {u32 bits, u32 prefix} word per 32
 * candidates, and {f64 avg, u64 vis[ceil(V/64)]} per PASSED candidate in candidate order.  An inbox holds
 * two parity halves (double buffering: one barrier per round) of `world` regions of `capacity` candidates.
 * Replaces: MVS_WIRE_COMPACT records for rounds (they also carried c, ref, px, slot -- resident on every GPU).
 * The reference has no counterpart (single process).
 */
int64_t mvs_exchange_bytes(const mvs_ctx* ctx, int world, int64_t capacity);

/*
 * Publish the accept decisions of N scored hypotheses (count >= bound && gate) into region `rank` of the
 * `parity` half of EVERY GPU's inbox (peer-mapped DEVICE pointers, HOST array of `world` entries; entry
 * `rank` is this GPU's own inbox).  Plain stores over NVLink from inside the compaction kernels; follow with
 * mvs_p2p_barrier before any GPU reads its inbox.  All data pointers DEVICE; enqueued on `stream`.
 * Replaces: the accept branch of MVS2.py:369,401-403 + the per-round all-gather of SURVEY 8(e).
 */
int mvs_publish_accepted(mvs_ctx* ctx, int64_t N, const uint64_t* vis_mask, const double* avg, const int32_t* count,
                         const uint8_t* gate, int bound, void* const* peer_inbox, int rank, int world, int64_t capacity,
                         int parity, void* stream);

/*
 * Overlap of the exchange with the scoring.  mvs_exchange_set_parts(ctx, P, min_batch) (1 <= P <= 8; the default 1 is
 * off; min_batch <= 0: 2^17) makes mvs_score_publish -- and mvs_expand_run -- cut a shard of at least min_batch
 * candidates into P consecutive POSITION ranges of the tile-ordered batch (anchor-tile ranges: the L1 reuse inside a
 * range survives).  Range k is its own K1 launch followed by the publish of its accept decisions (compaction + NVLink
 * stores) on a side stream of priority k, the last range on the caller's stream; the stream priorities make the GPU
 * work through the ranges in order without draining between them, so the publish of range k runs while the ranges
 * behind it are still being scored and only the last range's stores are exposed.  Such a round's region holds P
 * sub-regions (one per range: words over all candidates, entries of the range's own) and says so in its header; the
 * commit ORs them.  Results are identical for every P.  Call it with the same P on every GPU, BEFORE
 * mvs_exchange_bytes sizes the inboxes.
 * mvs_score_publish = mvs_score_batch(MVS_MODE_REFEXACT, on_device) + mvs_publish_accepted in one stream-ordered
 * (CUDA-graph capturable) sequence; all data pointers DEVICE.
 * Replaces: nothing in the reference (single process); the accept branch it feeds is MVS2.py:369,401-403.
 */
int mvs_exchange_set_parts(mvs_ctx* ctx, int parts, int64_t min_batch);
int mvs_score_publish(mvs_ctx* ctx, int64_t N, const double* c, const int32_t* ref, double min_ncc, int wid,
                      uint64_t* vis_mask, double* avg, int32_t* count, double* xy, const uint8_t* gate, int bound,
                      void* const* peer_inbox, int rank, int world, int64_t capacity, int parity, void* stream);

/*
 * Device-side barrier across the GPUs of the box: one tiny kernel, no host round trip, capturable in a
 * CUDA graph.  peer_flags: HOST array of `world` DEVICE pointers, entry g = GPU g's array of `world + 1`
 * uint64 words (slots 0..world-1: one flag per peer, slot world: that GPU's own epoch counter; ALL zero-initialised
 * by the caller before the first barrier) as mapped into this process.  Every rank must call it the same number of
 * times on the same arrays.  A rank that waits ~15 s for a peer gives up and raises a sticky error that
 * mvs_p2p_barrier_failed reports (it never hangs the GPU).
 */
int mvs_p2p_barrier(mvs_ctx* ctx, void* const* peer_flags, int rank, int world, void* stream);
int mvs_p2p_barrier_failed(mvs_ctx* ctx, void* stream);

/*
 * The whole expansion in ONE call (all rounds; one host synchronisation per round, no Python between
 * rounds).  Replaces: the while-loop of patch_expansion (MVS2.py:321-404) restructured into synchronous
 * rounds (DESIGN.md "Rounds"): per round mvs_round_generate -> score this GPU's shard + accept gate ->
 * mvs_publish_accepted -> [mvs_p2p_barrier] -> commit from the wire.  With world > 1 every GPU runs the
 * same call on the same seeds and cell table; results do not depend on `world`.
 */
typedef struct mvs_expand_params {
    double min_ncc;          /* MIN_NCC of MVS2.py:362 (0.7) */
    double scale;            /* args.scale of MVS2.py:369 */
    int32_t wid;             /* half window (5) */
    int32_t bound;           /* visible_lower_bound (MVS2.py:200-203) */
    int64_t max_rounds;      /* < 0: unlimited */
    int64_t max_iterations;  /* patches EXPANDED, the reference's `iteration < 100000` (MVS2.py:321); < 0: unlimited */
    int64_t max_patches;     /* stop after the round that reaches this many accepted patches; < 0: unlimited */
    int32_t rank, world;     /* this GPU's shard; world == 1: no exchange */
    void* const* peer_inbox; /* world > 1: see mvs_publish_accepted */
    void* const* peer_flags; /* world > 1: see mvs_p2p_barrier */
    int64_t capacity;        /* world > 1: candidates per inbox region; a round needs ceil(M / world) <= capacity */
    int32_t timing;          /* != 0: CUDA events around every round -> mvs_round_stat.ms */
    int32_t reserved;
} mvs_expand_params;

typedef struct mvs_round_stat {
    int64_t frontier, candidates, passed, accepted;
    float ms;
    int32_t reserved;
} mvs_round_stat;

/*   seeds      DEVICE pointer to n_seeds patch records (already filled into the cell table)
 *   stats      HOST array of max_stats entries (may be NULL); n_rounds, n_patches: HOST
 * The accepted patches of all rounds stay in the context, in commit order (= round order, slot order
 * inside a round); fetch them with mvs_expand_result. */
int mvs_expand_run(mvs_ctx* ctx, const void* seeds, int64_t n_seeds, const mvs_expand_params* params, mvs_round_stat* stats,
                   int max_stats, int* n_rounds, int64_t* n_patches, void* stream);
/* Copy accepted records [offset, offset + n) of the last mvs_expand_run to `records` (HOST, or DEVICE with on_device). */
int mvs_expand_result(mvs_ctx* ctx, void* records, int64_t offset, int64_t n, int on_device, void* stream);

/* Debug/parity access to the candidates of the last mvs_round_generate: any pointer may be
 * NULL; HOST pointers.  slot [M] int64, parent [M] int64, c [M,3], nrm [M,3], ref [M] int32. */
int mvs_round_candidates(mvs_ctx* ctx, int64_t* slot, int64_t* parent, double* c, double* nrm, int32_t* ref);

/*
 * NCC of M descriptor pairs.
 * Replaces: ctNcc(desc1, desc2) (MVS2.py:39-43) for uint8 descriptors of length n:
 * out[m] = sum(z(a_m) * z(b_m)) / (n - 1) with population-std z-scores, i.e. n/(n-1) times the
 * Pearson correlation; NaN when a descriptor has zero variance.  No context needed.
 *   a, b [M,n] uint8, out [M] float64; host or device pointers (on_device)
 */
int mvs_ncc_pairs(int device, int64_t M, int n, const uint8_t* a, const uint8_t* b, double* out, int on_device,
                  void* stream);

/*
 * Measurement hooks (bench.py): while enabled, every mvs_score_batch / mvs_round_score
 * brackets its scoring kernel (K1 alone, without the projection/binning launches before
 * it) with two CUDA events on the stream the kernel is launched on (a ring of the last 64
 * launches).  mvs_profile_score_ms waits for those kernels and returns their MEAN duration
 * in milliseconds and how many were averaged; mvs_profile_enable(ctx, 1) restarts the count.
 */
int mvs_profile_enable(mvs_ctx* ctx, int on);
int mvs_profile_score_ms(mvs_ctx* ctx, float* mean_ms, int* n_kernels);

/*
 * Gather-ceiling probe (the measured L1/L2 gather ceiling that the roofline of the
 * L2-resident configurations is reported against).  While on, the scoring launch of
 * mvs_score_batch(MVS_MODE_REFEXACT) is replaced by a loads-only kernel that walks the ordered batch
 * exactly as K1 does and issues exactly K1's loads (16-byte view quads of every window row and pixel
 * group, reference-window words, map entries) with no arithmetic and no stores; the outputs of such a
 * call are NOT written.  Timed through mvs_profile_enable / mvs_profile_score_ms like K1 itself.
 * Replaces: nothing in the reference (measurement only).
 */
int mvs_profile_probe(mvs_ctx* ctx, int on);

/* Number of kernels this library has launched on ctx since creation (for bench.py's
 * gpu_launches claim). */
int64_t mvs_launch_count(const mvs_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif /* MVS_NCC_H */
