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

#define FULL 0xffffffffu
#define XTILE 1024         // candidates per CTA of the publish kernels (256 threads x 4)

struct WireLayout {
    int64_t region_bytes;  // one source rank's region
    int64_t ent_off;       // byte offset of the entries inside a region
    int64_t capacity;      // candidates per region
    int wb;                // bytes per entry
    int mw;
};

__host__ __device__ static inline int64_t align_up(int64_t x, int64_t a) { return (x + a - 1) / a * a; }

static WireLayout wire_layout(const mvs_ctx* ctx, int64_t capacity) {
    WireLayout L;
    L.mw = (ctx->V + 63) / 64;
    L.wb = 8 + 8 * L.mw;
    L.capacity = capacity;
    const int64_t nw = (capacity + 31) / 32;
    L.ent_off = align_up(16 + 8 * nw, 256);
    L.region_bytes = align_up(L.ent_off + capacity * L.wb, 256);
    return L;
}

extern "C" int64_t mvs_exchange_bytes(const mvs_ctx* ctx, int world, int64_t capacity) {
    if (!ctx || world < 1 || capacity < 1) return 0;
    return 2 * (int64_t)world * wire_layout(ctx, capacity).region_bytes;
}

struct PeerInbox {
    uint8_t* base[MVS_MAX_PEERS];     // region `rank` of the current parity half in every GPU's inbox
    int world;
};

__device__ __forceinline__ bool pass_flag(const int32_t* count, const uint8_t* gate, int bound, int64_t i, int64_t N) {
    return i < N && count[i] >= bound && (gate == nullptr || gate[i] != 0);
}

// Per-tile counts of passed candidates AND, in the CTA that finishes last (ticket counter), the exclusive scan of the
// tile counts + the header {kept, n} to every inbox: one launch instead of two.
__global__ void __launch_bounds__(256)
    publish_count_scan(const int32_t* __restrict__ count, const uint8_t* __restrict__ gate, int bound, int64_t N,
                       int32_t* __restrict__ tile_counts, int T, const PeerInbox P, unsigned* __restrict__ ticket) {
    __shared__ int warp_sum[8];
    __shared__ bool s_last;
    __shared__ long long s_carry;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int64_t base = (int64_t)blockIdx.x * XTILE;
    int k = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) k += pass_flag(count, gate, bound, base + j * 256 + threadIdx.x, N) ? 1 : 0;
    k = __reduce_add_sync(FULL, k);
    if (lane == 0) warp_sum[w] = k;
    __syncthreads();
    if (threadIdx.x == 0) {
        int sum = 0;
        for (int q = 0; q < 8; ++q) sum += warp_sum[q];
        if ((int)blockIdx.x < T) tile_counts[blockIdx.x] = sum;
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
        hdr[1] = N;
    }
    if (threadIdx.x == 0) *ticket = 0u;                    // ready for the next launch
}

// words + entries of one tile of XTILE candidates to every inbox
__global__ void __launch_bounds__(256)
    publish_scatter(const int32_t* __restrict__ count, const uint8_t* __restrict__ gate, int bound, int64_t N,
                    const int32_t* __restrict__ tile_offsets, const uint64_t* __restrict__ vis, const double* __restrict__ avg,
                    const PeerInbox P, const WireLayout L) {
    extern __shared__ __align__(16) uint8_t s_ent[];       // 256 entries
    __shared__ int warp_base[8];
    __shared__ __align__(16) uint2 s_words[32];            // the tile's 32 words
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int64_t base = (int64_t)blockIdx.x * XTILE;
    int64_t out = tile_offsets[blockIdx.x];
    const int ent_words = L.wb >> 3;
    for (int j = 0; j < 4; ++j) {
        const int64_t i = base + j * 256 + threadIdx.x;
        const bool k = pass_flag(count, gate, bound, i, N);
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
}

int mvs_launch_publish(mvs_ctx* ctx, int64_t N, const uint64_t* vis, const double* avg, const int32_t* count,
                       const uint8_t* gate, int bound, void* const* peer_inbox, int rank, int world, int64_t capacity,
                       int parity, cudaStream_t s) {
    const WireLayout L = wire_layout(ctx, capacity);
    PeerInbox P;
    memset(&P, 0, sizeof(P));
    P.world = world;
    for (int d = 0; d < world; ++d)
        P.base[d] = (uint8_t*)peer_inbox[d] + ((int64_t)(parity & 1) * world + rank) * L.region_bytes;
    const int T = (int)((N + XTILE - 1) / XTILE);
    int rc;
    if ((rc = mvs_ensure((void**)&ctx->d_tiles, &ctx->tile_bytes, sizeof(int32_t) * (size_t)(T + 1), "publish scratch")) != MVS_OK)
        return rc;
    if (!ctx->d_ticket) {
        MVS_CUDA_CHECK(cudaMalloc(&ctx->d_ticket, 64));
        MVS_CUDA_CHECK(cudaMemsetAsync(ctx->d_ticket, 0, 64, s));
    }
    publish_count_scan<<<T > 0 ? T : 1, 256, 0, s>>>(count, gate, bound, N, ctx->d_tiles, T, P, (unsigned*)ctx->d_ticket);   // T == 0: an empty shard
    ctx->launches++;
    if (T > 0) {
        const size_t smem = (size_t)256 * L.wb;
        if (smem > 40 * 1024) MVS_CUDA_CHECK(cudaFuncSetAttribute(publish_scatter, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        publish_scatter<<<T, 256, smem, s>>>(count, gate, bound, N, ctx->d_tiles, vis, avg, P, L);
        ctx->launches++;
    }
    MVS_CUDA_CHECK(cudaGetLastError());
    return MVS_OK;
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
    if (!ctx || !ctx->d_barrier_state) return 0;
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
    const uint2 wd = *reinterpret_cast<const uint2*>(region + 16 + 8 * (j >> 5));
    bits = wd.x;
    prefix = wd.y;
    bitpos = (int)(j & 31);
    return (bits >> bitpos) & 1u;
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
