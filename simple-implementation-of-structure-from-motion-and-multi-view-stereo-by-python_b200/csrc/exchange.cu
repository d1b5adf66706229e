// Per-round exchange of accept decisions between the GPUs of one box, and the commit that consumes it.
//
// In a round every GPU holds the SAME candidate list (mvs_round_generate runs replicated on the full
// frontier) and scores only its shard.  What a peer cannot recompute about a candidate is therefore tiny:
// whether it passed, its visible set and its mean NCC.  The "minimal wire" of a shard is
//   header   int64 kept, int64 n                      (16 B)
//   words    {u32 bits, u32 prefix} per 32 candidates  bit j = candidate begin+32w+j passed;
//                                                      prefix = number of passed candidates before word w
//   entries  {f64 avg, u64 vis[mw]} per PASSED candidate, in candidate order
// i.e. 8 + 8*ceil(V/64) bytes per accepted candidate + 2 bits per candidate (the 64-byte
// MVS_WIRE_COMPACT records also carried c, ref, px and the slot id, which every peer already has).
// publish_* write a shard's wire into region `rank` of EVERY GPU's inbox with plain stores over NVLink
// (peer-mapped pointers; entries staged in shared memory and written as coalesced 16-byte runs), one
// device-side flag barrier (p2p_barrier, capturable in a CUDA graph) orders them, and commit_wire_*
// rebuild full patch records for the kept candidates from the resident candidate arrays.  Inboxes are
// double-buffered by round parity, so ONE barrier per round suffices.  With world == 1 the same kernels
// run on a context-owned local inbox: one code path for any GPU count.
// The reference has no counterpart (single process); what an accepted patch must carry is
// MVS2.py:401-403 (fill cells for every hit, enqueue).
#include "project.cuh"
#include "scan.cuh"
#include <stdlib.h>

#define FULL 0xffffffffu
#define XTILE 1024         // candidates per CTA of the publish kernels (256 threads x 4)

struct WireLayout {
    int64_t region_bytes;  // one source rank's region (large enough for either layout below)
    int64_t ent_off;       // byte offset of the entries inside a region / a part's sub-region
    int64_t capacity;      // candidates per region
    int64_t sub_bytes;     // partitioned layout: bytes of one part's sub-region {header, words over ALL candidates, entries}
    int64_t part_cap;      // partitioned layout: positions (hence entries at most) per part
    int wb;                // bytes per entry
    int mw;
    int parts;             // the context's mvs_exchange_set_parts value (1: single layout only)
};

// A region is written in ONE of two layouts, chosen by the sender round by round and named in its header
// (hdr[1] = n | parts_used << 48):
//   single       header | words | entries                                  (parts_used = 0 or 1)
//   partitioned  P sub-regions of sub_bytes, one per K1 launch (position range of the ordered batch); the words of
//                sub-region k cover ALL n candidates but only carry the bits of the candidates scored by launch k,
//                its entries are those candidates' in candidate order.  passed(i) = OR over the parts.
#define WIRE_PARTS_SHIFT 48
#define WIRE_N_MASK ((1ll << WIRE_PARTS_SHIFT) - 1)

__host__ __device__ static inline int64_t align_up(int64_t x, int64_t a) { return (x + a - 1) / a * a; }

static WireLayout wire_layout(const mvs_ctx* ctx, int64_t capacity) {
    WireLayout L;
    L.mw = (ctx->V + 63) / 64;
    L.wb = 8 + 8 * L.mw;
    L.capacity = capacity;
    const int64_t nw = (capacity + 31) / 32;
    L.ent_off = align_up(16 + 8 * nw, 256);
    L.region_bytes = align_up(L.ent_off + capacity * L.wb, 256);
    L.parts = ctx->xparts > 1 ? ctx->xparts : 1;
    L.part_cap = align_up((capacity + L.parts - 1) / L.parts, 1024);
    L.sub_bytes = align_up(L.ent_off + L.part_cap * L.wb, 256);
    if (L.parts > 1 && L.parts * L.sub_bytes > L.region_bytes) L.region_bytes = L.parts * L.sub_bytes;
    return L;
}

int64_t mvs_exchange_sub_bytes(const mvs_ctx* ctx, int64_t capacity) { return wire_layout(ctx, capacity).sub_bytes; }

// positions per K1 launch for a batch of N hypotheses in P parts (a multiple of every K1 variant's chunk)
static inline int64_t part_size_of(int64_t N, int P) { return align_up((N + P - 1) / P, 1024); }

extern "C" int mvs_exchange_set_parts(mvs_ctx* ctx, int parts, int64_t min_batch) {
    if (!ctx) { mvs_set_error("mvs_exchange_set_parts: null context"); return MVS_ERR_ARG; }
    if (parts < 1 || parts > MVS_MAX_PARTS) {
        mvs_set_error("mvs_exchange_set_parts: 1 <= parts <= %d", MVS_MAX_PARTS);
        return MVS_ERR_ARG;
    }
    ctx->xparts = parts;
    ctx->xparts_min = min_batch > 0 ? min_batch : (1ll << 17);
    return MVS_OK;
}

extern "C" int64_t mvs_exchange_bytes(const mvs_ctx* ctx, int world, int64_t capacity) {
    if (!ctx || world < 1 || capacity < 1) return 0;
    return 2 * (int64_t)world * wire_layout(ctx, capacity).region_bytes;
}

struct PeerInbox {
    uint8_t* base[MVS_MAX_PEERS];     // region `rank` of the current parity half in every GPU's inbox
    int world;
};

// which candidates a publish launch covers: all (used <= 1) or those of position range `id`
struct PartSel {
    const uint8_t* part;   // [N] position range of every candidate (ordered batches), nullptr: range = index / size
    int64_t size;
    int id, used;
};

__device__ __forceinline__ bool pass_flag(const int32_t* count, const uint8_t* gate, int bound, int64_t i, int64_t N,
                                          const PartSel& S) {
    if (i >= N) return false;
    // (membership first: the results of the other ranges may still be in flight)
    if (S.used > 1 && (S.part ? (int)S.part[i] : (int)(i / S.size)) != S.id) return false;
    return count[i] >= bound && (gate == nullptr || gate[i] != 0);
}

// Per-tile counts of passed candidates AND, in the CTA that finishes last (ticket counter), the exclusive scan of the
// tile counts + the header {kept, n} to every inbox: one launch instead of two.
__global__ void __launch_bounds__(256)
    publish_count_scan(const int32_t* __restrict__ count, const uint8_t* __restrict__ gate, int bound, int64_t N,
                       int32_t* __restrict__ tile_counts, int T, const PeerInbox P, unsigned* __restrict__ ticket,
                       const PartSel S) {
    __shared__ int warp_sum[8];
    __shared__ bool s_last;
    __shared__ long long s_carry;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int tile = blockIdx.x; tile < T; tile += gridDim.x) {     // (a few CTAs walk all tiles when the publish shares the GPU with K1)
        const int64_t base = (int64_t)tile * XTILE;
        int k = 0;
#pragma unroll
        for (int j = 0; j < 4; ++j) k += pass_flag(count, gate, bound, base + j * 256 + threadIdx.x, N, S) ? 1 : 0;
        k = __reduce_add_sync(FULL, k);
        if (lane == 0) warp_sum[w] = k;
        __syncthreads();
        if (threadIdx.x == 0) {
            int sum = 0;
            for (int q = 0; q < 8; ++q) sum += warp_sum[q];
            tile_counts[tile] = sum;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        __threadfence();
        s_last = atomicAdd(ticket, 1u) == gridDim.x - 1;
        s_carry = 0;
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();                                       // the other CTAs' counts are visible
    for (int b0 = 0; b0 < T; b0 += 256) {
        const int i = b0 + threadIdx.x;
        const int v = (i < T) ? tile_counts[i] : 0;
        int x = v;
#pragma unroll
        for (int sft = 1; sft < 32; sft <<= 1) {
            const int y = __shfl_up_sync(FULL, x, sft);
            if (lane >= sft) x += y;
        }
        if (lane == 31) warp_sum[w] = x;
        __syncthreads();
        int before = 0, total = 0;
        for (int q = 0; q < 8; ++q) {
            if (q < w) before += warp_sum[q];
            total += warp_sum[q];
        }
        if (i < T) tile_counts[i] = (int32_t)(s_carry + before + (x - v));
        __syncthreads();
        if (threadIdx.x == 0) s_carry += total;
        __syncthreads();
    }
    if (threadIdx.x < P.world) {
        int64_t* hdr = reinterpret_cast<int64_t*>(P.base[threadIdx.x]);
        hdr[0] = s_carry;
        hdr[1] = N | ((int64_t)(S.used > 1 ? S.used : 0) << WIRE_PARTS_SHIFT);
    }
    if (threadIdx.x == 0) *ticket = 0u;                    // ready for the next launch
}

// words + entries of one tile of XTILE candidates to every inbox
__global__ void __launch_bounds__(256)
    publish_scatter(const int32_t* __restrict__ count, const uint8_t* __restrict__ gate, int bound, int64_t N,
                    const int32_t* __restrict__ tile_offsets, const uint64_t* __restrict__ vis, const double* __restrict__ avg,
                    const PeerInbox P, const WireLayout L, const PartSel S, int T) {
    extern __shared__ __align__(16) uint8_t s_ent[];       // 256 entries
    __shared__ int warp_base[8];
    __shared__ __align__(16) uint2 s_words[32];            // the tile's 32 words
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int ent_words = L.wb >> 3;
    for (int tile = blockIdx.x; tile < T; tile += gridDim.x) {
    const int64_t base = (int64_t)tile * XTILE;
    int64_t out = tile_offsets[tile];
    for (int j = 0; j < 4; ++j) {
        const int64_t i = base + j * 256 + threadIdx.x;
        const bool k = pass_flag(count, gate, bound, i, N, S);
        const unsigned b = __ballot_sync(FULL, k);
        if (lane == 0) warp_base[w] = __popc(b);
        __syncthreads();
        int before = 0, total = 0;
        for (int q = 0; q < 8; ++q) {
            if (q < w) before += warp_base[q];
            total += warp_base[q];
        }
        if (lane == 0) s_words[j * 8 + w] = make_uint2(b, (uint32_t)(out + before));
        if (k) {
            const int slot = before + __popc(b & ((1u << lane) - 1u));
            uint64_t* e = reinterpret_cast<uint64_t*>(s_ent + (size_t)slot * L.wb);
            e[0] = (uint64_t)__double_as_longlong(avg[i]);
            for (int q = 0; q < L.mw; ++q) e[1 + q] = vis[i * L.mw + q];
        }
        __syncthreads();
        const int nwords = total * ent_words;
        const int64_t off = L.ent_off + out * L.wb;
        if ((L.wb & 15) == 0) {                            // 16-byte stores (V <= 64: 16-byte entries)
            const uint4* src = reinterpret_cast<const uint4*>(s_ent);
            for (int t = threadIdx.x; t < (nwords >> 1); t += 256) {
                const uint4 wv = src[t];
                for (int d = 0; d < P.world; ++d) reinterpret_cast<uint4*>(P.base[d] + off)[t] = wv;
            }
        } else {
            const uint64_t* src = reinterpret_cast<const uint64_t*>(s_ent);
            for (int t = threadIdx.x; t < nwords; t += 256) {
                const uint64_t wv = src[t];
                for (int d = 0; d < P.world; ++d) reinterpret_cast<uint64_t*>(P.base[d] + off)[t] = wv;
            }
        }
        out += total;
        __syncthreads();
    }
    // the tile's 32 words (256 bytes) as 16 coalesced 16-byte stores per inbox
    if (threadIdx.x < 16) {
        const int64_t w0 = base >> 5;                      // first word of this tile
        const int64_t nw = (N + 31) >> 5;
        const uint4 wv = reinterpret_cast<const uint4*>(s_words)[threadIdx.x];
        if (w0 + 2 * threadIdx.x < nw) {                   // pairs of words; the tail pair may hold one unused word (inside the region)
            for (int d = 0; d < P.world; ++d) reinterpret_cast<uint4*>(P.base[d] + 16 + 8 * w0)[threadIdx.x] = wv;
        }
    }
    __syncthreads();                                       // s_words / s_ent are reused by the next tile
    }
}

// part_id / parts_used / part / part_size: the position range this launch covers (parts_used <= 1: everything, single layout)
static int launch_publish_part(mvs_ctx* ctx, int64_t N, const uint64_t* vis, const double* avg, const int32_t* count,
                               const uint8_t* gate, int bound, void* const* peer_inbox, int rank, int world, int64_t capacity,
                               int parity, int part_id, int parts_used, const uint8_t* part, int64_t part_size, cudaStream_t s,
                               int max_ctas = 0) {
    const WireLayout L = wire_layout(ctx, capacity);
    PeerInbox P;
    memset(&P, 0, sizeof(P));
    P.world = world;
    for (int d = 0; d < world; ++d)
        P.base[d] = (uint8_t*)peer_inbox[d] + ((int64_t)(parity & 1) * world + rank) * L.region_bytes +
                    (parts_used > 1 ? (int64_t)part_id * L.sub_bytes : 0);
    PartSel S;
    S.part = part; S.size = part_size > 0 ? part_size : 1; S.id = part_id; S.used = parts_used > 1 ? parts_used : 1;
    const int T = (int)((N + XTILE - 1) / XTILE);
    int32_t* tiles = ctx->d_tiles + (size_t)part_id * (T + 1);          // (sized by the caller for all parts)
    unsigned* ticket = (unsigned*)ctx->d_ticket + part_id;
    // one CTA per tile.  (max_ctas > 0 caps the grid -- the kernels walk the tiles with a grid stride -- for a publish that
    // shares the GPU with K1; measured with 16 .. 128 CTAs the publish itself became the long pole (0.93 .. 0.48 ms per
    // step against 0.40), so no caller uses it)
    int grid = T > 0 ? T : 1;                              // T == 0: an empty shard
    if (max_ctas > 0 && grid > max_ctas) grid = max_ctas;
    publish_count_scan<<<grid, 256, 0, s>>>(count, gate, bound, N, tiles, T, P, ticket, S);
    ctx->launches++;
    if (T > 0) {
        const size_t smem = (size_t)256 * L.wb;
        if (smem > 40 * 1024) MVS_CUDA_CHECK(cudaFuncSetAttribute(publish_scatter, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        publish_scatter<<<grid, 256, smem, s>>>(count, gate, bound, N, tiles, vis, avg, P, L, S, T);
        ctx->launches++;
    }
    MVS_CUDA_CHECK(cudaGetLastError());
    return MVS_OK;
}

static int ensure_publish_scratch(mvs_ctx* ctx, int64_t N, int parts, cudaStream_t s) {
    const int T = (int)((N + XTILE - 1) / XTILE);
    int rc;
    if ((rc = mvs_ensure((void**)&ctx->d_tiles, &ctx->tile_bytes, sizeof(int32_t) * (size_t)(T + 1) * parts, "publish scratch")) != MVS_OK)
        return rc;
    if (!ctx->d_ticket) {
        MVS_CUDA_CHECK(cudaMalloc(&ctx->d_ticket, 64));
        MVS_CUDA_CHECK(cudaMemsetAsync(ctx->d_ticket, 0, 64, s));
    }
    return MVS_OK;
}

int mvs_launch_publish(mvs_ctx* ctx, int64_t N, const uint64_t* vis, const double* avg, const int32_t* count,
                       const uint8_t* gate, int bound, void* const* peer_inbox, int rank, int world, int64_t capacity,
                       int parity, cudaStream_t s) {
    int rc;
    if ((rc = ensure_publish_scratch(ctx, N, 1, s)) != MVS_OK) return rc;
    return launch_publish_part(ctx, N, vis, avg, count, gate, bound, peer_inbox, rank, world, capacity, parity, 0, 1, nullptr, 1, s);
}

// ---------------------------------------------------------------------------------------
// Score + publish of one shard with the exchange OVERLAPPED with the scoring.  The tile-ordered batch is cut into P
// consecutive position ranges (anchor-tile ranges, so the L1 reuse inside a range survives); range k is its own K1
// launch followed by the publish of its accept decisions -- compaction + NVLink stores into every GPU's inbox -- on
// side stream k, the last range on the caller's stream.  The launches do not depend on each other, so they are all
// enqueued at once and the hardware works through them back to back WITHOUT draining the GPU in between (the CTAs of
// the next launch fill the slots the previous one frees); the P launches share the grid cap of one launch, so the
// batch is scored by as many CTAs as before.  The publish of a finished range then runs while the other ranges are
// still being scored, and only the publish of the range that finishes last is exposed.  OFF by default
// (mvs_exchange_set_parts): measured per step of 2^20 hypotheses per GPU, plain / two ranges / four ranges -- 0.399 /
// 0.408 / 0.423 ms on one GPU (nothing to hide there), 0.512 / 0.517 / 0.532 ms on eight; with MVS_XMODE=1 (descending
// stream priorities) 0.502 ms on eight.  K1's two CTAs per SM fill the register file, so a publish CTA only runs where
// a K1 CTA has retired: at equal priority the publish queues behind the next range's CTAs, with priority it fragments
// the SMs.  Measured and rejected before that: the ranges as consecutive launches on ONE stream
// (every extra launch costs ~40 us of drain and ramp-up: K1 0.326 -> 0.367 ms for two ranges); ONE K1 launch whose warps report finished chunks into per-range counters for
// a gate kernel on the side stream (the __threadfence before every report costs as much: 0.326 -> 0.363 ms); side
// streams of DESCENDING priority (the high-priority publish CTAs fragment the register file K1's two CTAs per SM
// fill completely: 0.420 vs 0.408 ms); range launches of one chunk per CTA (0.438 ms: 1.7 chunks per CTA through the
// shared grid cap amortise the CTA start-up).
// The fork / join is stream-ordered (events), so the sequence is capturable in a CUDA graph like the plain one.
// P <= 1, small shards: the plain sequence.
// ---------------------------------------------------------------------------------------
int mvs_launch_score_publish(mvs_ctx* ctx, int64_t N, const double* c, const int32_t* ref, double thr, int wid, uint64_t* vis,
                             double* avg, int32_t* count, double* xy, const uint8_t* gate, int bound, void* const* peer_inbox,
                             int rank, int world, int64_t capacity, int parity, cudaStream_t s) {
    int rc;
    const int P = ctx->xparts;
    if (P <= 1 || N < ctx->xparts_min || N < 2048ll * P || N < MVS_SORT_MIN || ctx->probe_gather) {
        if (N > 0 && (rc = mvs_launch_score_refexact(ctx, N, c, ref, thr, wid, vis, avg, count, xy, nullptr, s)) != MVS_OK) return rc;
        return mvs_launch_publish(ctx, N, vis, avg, count, gate, bound, peer_inbox, rank, world, capacity, parity, s);
    }
    if (wid < 1 || wid > 7) {
        mvs_set_error("wid %d not supported (1..7)", wid);
        return MVS_ERR_ARG;
    }
    if ((rc = mvs_build_window_maps(ctx, wid, s)) != MVS_OK) return rc;
    if ((rc = ensure_publish_scratch(ctx, N, P, s)) != MVS_OK) return rc;
    if ((rc = mvs_ensure((void**)&ctx->d_bin_part, &ctx->bin_part_bytes, (size_t)N, "position ranges")) != MVS_OK) return rc;
    static int xmode = -1;                                 // MVS_XMODE (measurement knob): 0 = every side stream at the default (lowest)
    if (xmode < 0) {                                       // priority, 1 = side streams of descending priority (range 0 highest)
        const char* e = getenv("MVS_XMODE");
        xmode = e ? atoi(e) : 0;
    }
    if (!ctx->x_fork) {
        int lo = 0, hi = 0;                                // (numerically: hi <= lo, hi = the highest priority)
        MVS_CUDA_CHECK(cudaDeviceGetStreamPriorityRange(&lo, &hi));
        for (int i = 0; i < MVS_MAX_PARTS; ++i) {
            int prio = hi + i < lo ? hi + i : lo;
            if (xmode != 1) prio = lo;
            MVS_CUDA_CHECK(cudaStreamCreateWithPriority(&ctx->x_side[i], cudaStreamNonBlocking, prio));
            MVS_CUDA_CHECK(cudaEventCreateWithFlags(&ctx->x_ev[i], cudaEventDisableTiming));
        }
        MVS_CUDA_CHECK(cudaEventCreateWithFlags(&ctx->x_fork, cudaEventDisableTiming));
    }
    const int64_t S = part_size_of(N, P);                  // positions per range: a multiple of 1024, hence of K1's chunk
    if ((rc = mvs_bin_hypotheses(ctx, N, c, ref, wid, true, vis, avg, count, xy, nullptr, s, ctx->d_bin_part, S)) != MVS_OK) return rc;
    const int pslot = (int)(ctx->prof_n % MVS_PROF_RING);
    if (ctx->profile) MVS_CUDA_CHECK(cudaEventRecord(ctx->prof_ev[2 * pslot], s));
    MVS_CUDA_CHECK(cudaEventRecord(ctx->x_fork, s));
    ctx->k1_share = P;                                     // the P range launches share the grid cap of one launch
    // ---- ranges 0 .. P-2 on the side streams (enqueued in order: a tool that serialises kernels still runs them in order)
    for (int k = 0; k + 1 < P; ++k) {
        cudaStream_t q = ctx->x_side[k];
        MVS_CUDA_CHECK(cudaStreamWaitEvent(q, ctx->x_fork, 0));
        if ((rc = mvs_launch_k1(ctx, N, ref, thr, wid, vis, avg, count, nullptr, true, (int64_t)k * S, (int64_t)(k + 1) * S, q)) != MVS_OK) {
            ctx->k1_share = 0;
            return rc;
        }
        if ((rc = launch_publish_part(ctx, N, vis, avg, count, gate, bound, peer_inbox, rank, world, capacity, parity, k, P,
                                      ctx->d_bin_part, S, q)) != MVS_OK) {
            ctx->k1_share = 0;
            return rc;
        }
        MVS_CUDA_CHECK(cudaEventRecord(ctx->x_ev[k], q));
    }
    // ---- the last range on the caller's stream, then join
    rc = mvs_launch_k1(ctx, N, ref, thr, wid, vis, avg, count, nullptr, true, (int64_t)(P - 1) * S, N, s);
    ctx->k1_share = 0;
    if (rc != MVS_OK) return rc;
    if (ctx->profile) {                                    // K1 alone: the lowest-priority range ends last
        MVS_CUDA_CHECK(cudaEventRecord(ctx->prof_ev[2 * pslot + 1], s));
        ctx->prof_n++;
    }
    if ((rc = launch_publish_part(ctx, N, vis, avg, count, gate, bound, peer_inbox, rank, world, capacity, parity, P - 1, P,
                                  ctx->d_bin_part, S, s)) != MVS_OK)
        return rc;
    for (int k = 0; k + 1 < P; ++k) MVS_CUDA_CHECK(cudaStreamWaitEvent(s, ctx->x_ev[k], 0));
    return MVS_OK;
}

extern "C" int mvs_score_publish(mvs_ctx* ctx, int64_t N, const double* c, const int32_t* ref, double min_ncc, int wid,
                                 uint64_t* vis_mask, double* avg, int32_t* count, double* xy, const uint8_t* gate, int bound,
                                 void* const* peer_inbox, int rank, int world, int64_t capacity, int parity, void* stream) {
    if (!ctx) { mvs_set_error("mvs_score_publish: null context"); return MVS_ERR_ARG; }
    if (N < 0 || capacity < N || !peer_inbox || world < 1 || world > MVS_MAX_PEERS || rank < 0 || rank >= world ||
        (N > 0 && (!c || !ref || !vis_mask || !avg || !count))) {
        mvs_set_error("mvs_score_publish: need 0 <= N <= capacity, 1 <= world <= %d, 0 <= rank < world, the inbox table, c, ref "
                      "and vis_mask, avg, count", MVS_MAX_PEERS);
        return MVS_ERR_ARG;
    }
    for (int d = 0; d < world; ++d)
        if (!peer_inbox[d]) { mvs_set_error("mvs_score_publish: null inbox pointer %d", d); return MVS_ERR_ARG; }
    MVS_CUDA_CHECK(cudaSetDevice(ctx->device));
    return mvs_launch_score_publish(ctx, N, c, ref, min_ncc, wid, vis_mask, avg, count, xy, gate, bound, peer_inbox, rank, world,
                                    capacity, parity, (cudaStream_t)stream);
}

extern "C" int mvs_publish_accepted(mvs_ctx* ctx, int64_t N, const uint64_t* vis_mask, const double* avg, const int32_t* count,
                                    const uint8_t* gate, int bound, void* const* peer_inbox, int rank, int world,
                                    int64_t capacity, int parity, void* stream) {
    if (!ctx) { mvs_set_error("mvs_publish_accepted: null context"); return MVS_ERR_ARG; }
    if (N < 0 || capacity < N || !peer_inbox || world < 1 || world > MVS_MAX_PEERS || rank < 0 || rank >= world ||
        (N > 0 && (!vis_mask || !avg || !count))) {
        mvs_set_error("mvs_publish_accepted: need 0 <= N <= capacity, 1 <= world <= %d, 0 <= rank < world, the inbox table and "
                      "vis_mask, avg, count", MVS_MAX_PEERS);
        return MVS_ERR_ARG;
    }
    for (int d = 0; d < world; ++d)
        if (!peer_inbox[d]) { mvs_set_error("mvs_publish_accepted: null inbox pointer %d", d); return MVS_ERR_ARG; }
    MVS_CUDA_CHECK(cudaSetDevice(ctx->device));
    return mvs_launch_publish(ctx, N, vis_mask, avg, count, gate, bound, peer_inbox, rank, world, capacity, parity,
                              (cudaStream_t)stream);
}

// ---------------------------------------------------------------------------------------
// Device-side barrier across the GPUs of the box: every rank stores its epoch into slot `rank` of every
// GPU's flag array (release, system scope, after a system-wide fence that orders the publish kernels'
// stores before it) and waits until all slots of its own array reached the epoch.  One tiny kernel, no
// host round trip, capturable in a CUDA graph (the epoch lives in device memory).  A rank that waits
// longer than ~15 s gives up and raises the context's error flag instead of hanging the GPU.
// ---------------------------------------------------------------------------------------
struct PeerFlags {
    unsigned long long* flags[MVS_MAX_PEERS];
    int world, rank;
};

__global__ void __launch_bounds__(32) p2p_barrier(const PeerFlags P, unsigned long long* __restrict__ epoch_ctr,
                                                  int* __restrict__ err) {
    __shared__ unsigned long long s_epoch;
    if (threadIdx.x == 0) s_epoch = ++(*epoch_ctr);
    __syncwarp();
    const unsigned long long epoch = s_epoch;
    const int d = threadIdx.x;
    if (d < P.world) {
        __threadfence_system();
        asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(P.flags[d] + P.rank), "l"(epoch) : "memory");
        const unsigned long long* mine = P.flags[P.rank] + d;
        const long long t0 = clock64();
        unsigned long long seen = 0;
        for (;;) {
            asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(seen) : "l"(mine) : "memory");
            if (seen >= epoch) break;
            if (clock64() - t0 > 30000000000ll) {          // ~15 s at 2 GHz: a peer is gone
                *err = 1;
                break;
            }
            __nanosleep(64);
        }
    }
    __syncwarp();
    __threadfence_system();
}

extern "C" int mvs_p2p_barrier(mvs_ctx* ctx, void* const* peer_flags, int rank, int world, void* stream) {
    if (!ctx) { mvs_set_error("mvs_p2p_barrier: null context"); return MVS_ERR_ARG; }
    if (!peer_flags || world < 1 || world > MVS_MAX_PEERS || rank < 0 || rank >= world) {
        mvs_set_error("mvs_p2p_barrier: need 1 <= world <= %d and 0 <= rank < world", MVS_MAX_PEERS);
        return MVS_ERR_ARG;
    }
    MVS_CUDA_CHECK(cudaSetDevice(ctx->device));
    if (!ctx->d_barrier_state) {
        MVS_CUDA_CHECK(cudaMalloc(&ctx->d_barrier_state, 64));
        MVS_CUDA_CHECK(cudaMemset(ctx->d_barrier_state, 0, 64));
    }
    PeerFlags P;
    memset(&P, 0, sizeof(P));
    P.world = world;
    P.rank = rank;
    for (int d = 0; d < world; ++d) {
        if (!peer_flags[d]) { mvs_set_error("mvs_p2p_barrier: null flag pointer %d", d); return MVS_ERR_ARG; }
        P.flags[d] = (unsigned long long*)peer_flags[d];
    }
    // the epoch counter lives in slot `world` of this GPU's OWN flag array: it is zeroed together with the flags, so
    // the ranks' epochs agree by construction whatever else the contexts did before
    p2p_barrier<<<1, 32, 0, (cudaStream_t)stream>>>(P, P.flags[rank] + world, (int*)((uint8_t*)ctx->d_barrier_state + 8));
    ctx->launches++;
    MVS_CUDA_CHECK(cudaGetLastError());
    return MVS_OK;
}

// 0 = fine, 1 = a barrier gave up waiting for a peer since the context was created (synchronises `stream`)
extern "C" int mvs_p2p_barrier_failed(mvs_ctx* ctx, void* stream) {
    if (!ctx) return 0;
    if (!ctx->d_barrier_state) return 0;
    int e = 0;
    if (cudaMemcpyAsync(&e, (uint8_t*)ctx->d_barrier_state + 8, sizeof(int), cudaMemcpyDeviceToHost, (cudaStream_t)stream) != cudaSuccess ||
        cudaStreamSynchronize((cudaStream_t)stream) != cudaSuccess)
        return 1;
    return e;
}

// ---------------------------------------------------------------------------------------
// Commit from the wire (identical on every GPU): candidates [0, M) of the round, shard r =
// [M*r/world, M*(r+1)/world) scored by rank r, its wire in region r of the local inbox.
//   flags : keep[i] = passed(i) and not (dj = +1 candidate whose dj = -1 sibling slot-1 also passed) -- the
//           `break` of MVS2.py:404
//   (exclusive scan)
//   apply : kept candidates become full patch records of the next frontier, in slot order, and clear cell
//           (v, floor(x/cs), floor(y/cs)) of every visible view (MVS2.py:401-402)
// n = (O_ref - c)/|O_ref - c| is resident (candidate generation), x, y is re-projected with the
// scorer's own operations (bit-identical), count = popcount(vis).
// ---------------------------------------------------------------------------------------
struct WireView {
    const uint8_t* half;               // current parity half of the local inbox
    int64_t begin[MVS_MAX_PEERS + 1];  // shard bounds
    int world;
};

__device__ __forceinline__ bool wire_passed(const WireView& Wv, const WireLayout& L, int64_t i, const uint8_t*& region,
                                            uint32_t& bits, uint32_t& prefix, int& bitpos) {
    int r = 0;
    while (r + 1 < Wv.world && i >= Wv.begin[r + 1]) ++r;
    const int64_t j = i - Wv.begin[r];
    region = Wv.half + (int64_t)r * L.region_bytes;
    bitpos = (int)(j & 31);
    // the sender's layout of this round: single, or one sub-region per K1 launch (a candidate passed in at most one)
    const int used = (int)(*reinterpret_cast<const int64_t*>(region + 8) >> WIRE_PARTS_SHIFT);
    for (int k = 0;;) {
        const uint2 wd = *reinterpret_cast<const uint2*>(region + 16 + 8 * (j >> 5));
        bits = wd.x;
        prefix = wd.y;
        if ((bits >> bitpos) & 1u) return true;
        if (++k >= used) return false;
        region += L.sub_bytes;
    }
}

__global__ void __launch_bounds__(256)
    commit_wire_flags(const WireView Wv, const WireLayout L, int64_t M, const int64_t* __restrict__ slot,
                      int32_t* __restrict__ keep) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= M) return;
    const uint8_t* region;
    uint32_t bits, prefix;
    int bp;
    bool k = wire_passed(Wv, L, i, region, bits, prefix, bp);
    if (k && i > 0) {
        const long long s = slot[i];
        if ((s & 1) && slot[i - 1] == s - 1) {
            const uint8_t* r2;
            uint32_t b2, p2;
            int bp2;
            if (wire_passed(Wv, L, i - 1, r2, b2, p2, bp2)) k = false;
        }
    }
    keep[i] = k ? 1 : 0;
}

__device__ __forceinline__ bool cell_of(double x, double y, int cs, int& ci, int& cj) {
    if (!(isfinite(x) && isfinite(y))) return false;
    const double fi = floor(__ddiv_rn(x, (double)cs)), fj = floor(__ddiv_rn(y, (double)cs));
    if (fabs(fi) > 1e9 || fabs(fj) > 1e9) return false;
    ci = (int)fi;
    cj = (int)fj;
    return true;
}

__global__ void __launch_bounds__(256)
    commit_wire_apply(const WireView Wv, const WireLayout L, int64_t M, const int32_t* __restrict__ offsets,
                      const int64_t* __restrict__ total, const int64_t* __restrict__ slot, const double* __restrict__ cand_c,
                      const double* __restrict__ cand_n, const int32_t* __restrict__ cand_ref,
                      const int32_t* __restrict__ cand_px, const CamProj* __restrict__ cams, int V, int cs, int wc, int hc,
                      uint8_t* __restrict__ cells, uint8_t* __restrict__ next, int rec_bytes) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= M) return;
    const int64_t off = offsets[i];
    const int64_t nxt = (i + 1 < M) ? (int64_t)offsets[i + 1] : *total;
    if (nxt == off) return;                                // not passed, or the dropped sibling
    const uint8_t* region;
    uint32_t bits, prefix;
    int bp;
    wire_passed(Wv, L, i, region, bits, prefix, bp);
    const int64_t e = (int64_t)prefix + __popc(bits & ((1u << bp) - 1u));
    const uint64_t* ent = reinterpret_cast<const uint64_t*>(region + L.ent_off + e * L.wb);
    mvs_patch_record* r = reinterpret_cast<mvs_patch_record*>(next + off * rec_bytes);
    const double c0 = cand_c[3 * i], c1 = cand_c[3 * i + 1], c2 = cand_c[3 * i + 2];
    const int v = cand_ref[i];
    double x, y;
    project_ref(cams[v], c0, c1, c2, x, y);
    r->c[0] = c0; r->c[1] = c1; r->c[2] = c2;
    r->n[0] = cand_n[3 * i]; r->n[1] = cand_n[3 * i + 1]; r->n[2] = cand_n[3 * i + 2];
    r->xy[0] = x; r->xy[1] = y;
    r->avg = __longlong_as_double((long long)ent[0]);
    r->ref = v;
    r->index = slot[i];
    r->px[0] = cand_px[2 * i];
    r->px[1] = cand_px[2 * i + 1];
    uint64_t* rv = reinterpret_cast<uint64_t*>(r + 1);
    int cnt = 0;
    int ci = 0, cj = 0;
    const bool in = cell_of(x, y, cs, ci, cj) && ci >= 0 && ci < wc && cj >= 0 && cj < hc;
    for (int q = 0; q < L.mw; ++q) {
        uint64_t b = ent[1 + q];
        rv[q] = b;
        cnt += __popcll(b);
        while (in && b) {
            const int vv = q * 64 + __ffsll((long long)b) - 1;
            b &= b - 1;
            if (vv < V) cells[((int64_t)vv * wc + ci) * hc + cj] = 0;     // MVS2.py:105
        }
    }
    r->count = cnt;
}

int mvs_launch_commit_wire(mvs_ctx* ctx, const void* inbox_local, int world, int64_t capacity, int parity, int64_t M,
                           void* next_frontier, int64_t* d_n_next, cudaStream_t s) {
    if (M == 0) {
        MVS_CUDA_CHECK(cudaMemsetAsync(d_n_next, 0, sizeof(int64_t), s));
        return MVS_OK;
    }
    const WireLayout L = wire_layout(ctx, capacity);
    WireView Wv;
    memset(&Wv, 0, sizeof(Wv));
    Wv.world = world;
    Wv.half = (const uint8_t*)inbox_local + (int64_t)(parity & 1) * world * L.region_bytes;
    for (int r = 0; r <= world; ++r) Wv.begin[r] = (M * r) / world;
    int rc;
    if ((rc = mvs_ensure((void**)&ctx->d_counts, &ctx->counts_bytes, sizeof(int32_t) * M, "commit flags")) != MVS_OK) return rc;
    if ((rc = mvs_ensure((void**)&ctx->d_scan, &ctx->scan_bytes, sizeof(int64_t) * ((M + 1023) / 1024 + 2), "scan")) != MVS_OK) return rc;
    const unsigned blocks = (unsigned)((M + 255) / 256);
    commit_wire_flags<<<blocks, 256, 0, s>>>(Wv, L, M, ctx->cand_slot, ctx->d_counts);
    if ((rc = mvs_exclusive_scan_i32(ctx->d_counts, M, ctx->d_scan, d_n_next, s)) != MVS_OK) return rc;
    commit_wire_apply<<<blocks, 256, 0, s>>>(Wv, L, M, ctx->d_counts, d_n_next, ctx->cand_slot, ctx->cand_c, ctx->cand_n,
                                            ctx->cand_ref, ctx->cand_px, ctx->d_cam, ctx->V, ctx->cell_size, ctx->wc, ctx->hc,
                                            ctx->d_cells, (uint8_t*)next_frontier,
                                            (int)(sizeof(mvs_patch_record) + 8 * ((ctx->V + 63) / 64)));
    ctx->launches += 2 + mvs_scan_launches(M);
    MVS_CUDA_CHECK(cudaGetLastError());
    return MVS_OK;
}
