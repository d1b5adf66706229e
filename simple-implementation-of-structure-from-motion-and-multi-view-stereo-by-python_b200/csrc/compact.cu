// K4a accept_compact: order-preserving stream compaction of accepted hypotheses into
// fixed-size patch records (the payload of the per-round all-gather).
//
// Replaces the accept branch of the reference's expansion loop (MVS2.py:369,401-403:
// "if visible_ct >= bound and ...: fill cells, enqueue") for a whole batch: a
// hypothesis is kept iff count >= bound and its optional caller-side gate byte is
// set.  Output order = input order (ascending global index), so the result does not
// depend on how the batch was sharded across GPUs.
//
// Three small launches: per-tile counts, one-block exclusive scan, scatter (records are
// staged in shared memory and written out coalesced).
// HBM-bound; algorithmic bytes = 4 B/hypothesis read + record_bytes per accepted.
#include "mvs_common.cuh"

#define TILE 1024          // hypotheses per CTA (256 threads x 4)

// Destinations of the compacted records.  world == 1: the caller's own buffer.  world > 1: the
// inbox of EVERY GPU of the box (peer-mapped pointers): this rank's records go to region `rank` of
// each inbox, `capacity` records per region, and its count to slot `rank` of each count array -- the
// compaction IS the all-gather (stores over NVLink), no collective call and no host round trip.
struct PeerList {
    uint8_t* rec[MVS_MAX_PEERS];
    int64_t* cnt[MVS_MAX_PEERS];
    int world, rank;
    int wire;              // MVS_WIRE_FULL | MVS_WIRE_COMPACT: format of the records written
};

// MVS_WIRE_COMPACT: only what a peer cannot recompute.  Every patch the reference's MVS creates has
// n = (O_ref - c)/|O_ref - c| (MVS2.py:247, 357-358) and carries the reference projection of c as its
// x, y (MVS2.py:74); count = popcount(visible set).  56 + 8*ceil(V/64) bytes instead of 96 + 8*ceil(V/64).
struct WireCompact {
    double c[3];
    double avg;
    int64_t index;
    int32_t ref;
    int32_t px[2];
    int32_t pad;
};
static_assert(sizeof(WireCompact) == 56, "wire layout");
#define FULL 0xffffffffu

__device__ __forceinline__ bool keep_flag(const int32_t* count, const uint8_t* gate, int bound, int64_t i, int64_t N) {
    return i < N && count[i] >= bound && (gate == nullptr || gate[i] != 0);
}

__global__ void __launch_bounds__(256) compact_count(const int32_t* __restrict__ count, const uint8_t* __restrict__ gate,
                                                     int bound, int64_t N, int32_t* __restrict__ tile_counts) {
    __shared__ int warp_sum[8];
    const int64_t base = (int64_t)blockIdx.x * TILE;
    int k = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) k += keep_flag(count, gate, bound, base + j * 256 + threadIdx.x, N) ? 1 : 0;
    k = __reduce_add_sync(FULL, k);
    if ((threadIdx.x & 31) == 0) warp_sum[threadIdx.x >> 5] = k;
    __syncthreads();
    if (threadIdx.x == 0) {
        int s = 0;
        for (int w = 0; w < 8; ++w) s += warp_sum[w];
        tile_counts[blockIdx.x] = s;
    }
}

// single CTA: exclusive scan of tile_counts[0..T) in place; total -> *n_out
__global__ void __launch_bounds__(1024) compact_scan(int32_t* __restrict__ tile_counts, int T, const PeerList P) {
    __shared__ int64_t carry;
    __shared__ int wsum[32];
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int base = 0; base < T; base += 1024) {
        const int i = base + threadIdx.x;
        const int v = (i < T) ? tile_counts[i] : 0;
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
        const int64_t excl = carry + (w > 0 ? wsum[w - 1] : 0) + (x - v);
        if (i < T) tile_counts[i] = (int32_t)excl;
        __syncthreads();
        if (threadIdx.x == 1023) carry = excl + v;
        __syncthreads();
    }
    if (threadIdx.x < P.world) P.cnt[threadIdx.x][P.rank] = carry;   // every GPU learns this rank's count
}

// Records are assembled in shared memory (a kept hypothesis writes its own record there) and
// then streamed out as one contiguous, coalesced run of 8-byte words per 256 hypotheses.
__global__ void __launch_bounds__(256)
    compact_scatter(const int32_t* __restrict__ count, const uint8_t* __restrict__ gate, int bound, int64_t N,
                    const int32_t* __restrict__ tile_offsets, int64_t index_base, const double* __restrict__ c,
                    const double* __restrict__ nrm, const int32_t* __restrict__ ref, const uint64_t* __restrict__ vis,
                    const double* __restrict__ avg, const double* __restrict__ xy, int mw, const PeerList P,
                    int rec_bytes, int64_t capacity, const int64_t* __restrict__ index_arr, const int32_t* __restrict__ px) {
    extern __shared__ __align__(16) uint8_t s_rec[];       // 256 records
    __shared__ int warp_base[8];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int64_t base = (int64_t)blockIdx.x * TILE;
    int64_t out = tile_offsets[blockIdx.x];
    const int rec_words = rec_bytes >> 3;
    for (int j = 0; j < 4; ++j) {
        const int64_t i = base + j * 256 + threadIdx.x;
        const bool k = keep_flag(count, gate, bound, i, N);
        const unsigned b = __ballot_sync(FULL, k);
        if (lane == 0) warp_base[w] = __popc(b);
        __syncthreads();
        int before = 0, total = 0;
        for (int q = 0; q < 8; ++q) {
            if (q < w) before += warp_base[q];
            total += warp_base[q];
        }
        if (k && P.wire == MVS_WIRE_COMPACT) {
            const int slot = before + __popc(b & ((1u << lane) - 1u));
            WireCompact* r = reinterpret_cast<WireCompact*>(s_rec + (size_t)slot * rec_bytes);
            r->c[0] = c[3 * i]; r->c[1] = c[3 * i + 1]; r->c[2] = c[3 * i + 2];
            r->avg = avg[i];
            r->index = index_arr ? index_arr[i] : index_base + i;
            r->ref = ref[i];
            r->px[0] = px ? px[2 * i] : -1;
            r->px[1] = px ? px[2 * i + 1] : -1;
            r->pad = 0;
            uint64_t* rv = reinterpret_cast<uint64_t*>(r + 1);
            for (int q = 0; q < mw; ++q) rv[q] = vis[i * mw + q];
        } else if (k) {
            const int slot = before + __popc(b & ((1u << lane) - 1u));
            mvs_patch_record* r = reinterpret_cast<mvs_patch_record*>(s_rec + (size_t)slot * rec_bytes);
            r->c[0] = c[3 * i]; r->c[1] = c[3 * i + 1]; r->c[2] = c[3 * i + 2];
            if (nrm) { r->n[0] = nrm[3 * i]; r->n[1] = nrm[3 * i + 1]; r->n[2] = nrm[3 * i + 2]; }
            else { r->n[0] = r->n[1] = r->n[2] = 0.0; }
            r->xy[0] = xy[2 * i]; r->xy[1] = xy[2 * i + 1];
            r->avg = avg[i];
            r->ref = ref[i];
            r->count = count[i];
            r->index = index_arr ? index_arr[i] : index_base + i;
            r->px[0] = px ? px[2 * i] : -1;
            r->px[1] = px ? px[2 * i + 1] : -1;
            uint64_t* rv = reinterpret_cast<uint64_t*>(r + 1);
            for (int q = 0; q < mw; ++q) rv[q] = vis[i * mw + q];
        }
        __syncthreads();
        // records beyond `capacity` are counted but not written
        int64_t room = capacity - out;
        room = room < 0 ? 0 : room;
        const int nrec = (int)(room < total ? room : total);
        const int nwords = nrec * rec_words;
        const int64_t off = ((int64_t)P.rank * capacity + out) * rec_bytes;
        if ((rec_bytes & 15) == 0) {                       // 16-byte stores (64-byte compact wire records at V <= 64)
            const uint4* src = reinterpret_cast<const uint4*>(s_rec);
            for (int t = threadIdx.x; t < (nwords >> 1); t += 256) {
                const uint4 wv = src[t];
                for (int d = 0; d < P.world; ++d) reinterpret_cast<uint4*>(P.rec[d] + off)[t] = wv;
            }
        } else {
            const uint64_t* src = reinterpret_cast<const uint64_t*>(s_rec);
            for (int t = threadIdx.x; t < nwords; t += 256) {
                const uint64_t wv = src[t];
                for (int d = 0; d < P.world; ++d) reinterpret_cast<uint64_t*>(P.rec[d] + off)[t] = wv;
            }
        }
        out += total;
        __syncthreads();
    }
}

static int launch_compact_peers(mvs_ctx* ctx, int64_t N, int64_t index_base, const double* c, const double* nrm,
                                const int32_t* ref, const uint64_t* vis, const double* avg, const int32_t* count,
                                const double* xy, const uint8_t* gate, int bound, const PeerList& P, int64_t capacity,
                                const int64_t* index_arr, const int32_t* px, cudaStream_t s) {
    const int T = (int)((N + TILE - 1) / TILE);
    if ((size_t)T * sizeof(int32_t) > ctx->tile_bytes) {
        if (ctx->d_tiles) cudaFree(ctx->d_tiles);
        ctx->d_tiles = nullptr;
        ctx->tile_bytes = 0;
        const size_t want = (size_t)T * sizeof(int32_t) * 2 + 4096;
        if (cudaMalloc(&ctx->d_tiles, want) != cudaSuccess) {
            cudaGetLastError();
            mvs_set_error("compaction scratch allocation of %zu bytes failed", want);
            return MVS_ERR_NOMEM;
        }
        ctx->tile_bytes = want;
    }
    const int mw = (ctx->V + 63) / 64;
    if (T > 0) {
        compact_count<<<T, 256, 0, s>>>(count, gate, bound, N, ctx->d_tiles);
        ctx->launches++;
    }
    compact_scan<<<1, 1024, 0, s>>>(ctx->d_tiles, T, P);          // T == 0: just publishes a zero count
    ctx->launches++;
    if (T > 0) {
        const int rec_bytes = (int)((P.wire == MVS_WIRE_COMPACT ? sizeof(WireCompact) : sizeof(mvs_patch_record)) + 8 * mw);
        const size_t smem = (size_t)256 * rec_bytes;
        if (smem > 48 * 1024)
            MVS_CUDA_CHECK(cudaFuncSetAttribute(compact_scatter, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        compact_scatter<<<T, 256, smem, s>>>(count, gate, bound, N, ctx->d_tiles, index_base, c, nrm, ref, vis, avg, xy, mw, P,
                                             rec_bytes, capacity, index_arr, px);
        ctx->launches++;
    }
    MVS_CUDA_CHECK(cudaGetLastError());
    return MVS_OK;
}

int mvs_launch_compact(mvs_ctx* ctx, int64_t N, int64_t index_base, const double* c, const double* nrm, const int32_t* ref,
                       const uint64_t* vis, const double* avg, const int32_t* count, const double* xy, const uint8_t* gate,
                       int bound, void* records, int64_t capacity, int64_t* d_n_out, const int64_t* index_arr,
                       const int32_t* px, cudaStream_t s) {
    PeerList P;
    memset(&P, 0, sizeof(P));
    P.world = 1;
    P.rank = 0;
    P.wire = MVS_WIRE_FULL;
    P.rec[0] = (uint8_t*)records;
    P.cnt[0] = d_n_out;
    return launch_compact_peers(ctx, N, index_base, c, nrm, ref, vis, avg, count, xy, gate, bound, P, capacity, index_arr, px, s);
}

int mvs_launch_compact_p2p(mvs_ctx* ctx, int64_t N, int64_t index_base, const double* c, const double* nrm,
                           const int32_t* ref, const uint64_t* vis, const double* avg, const int32_t* count,
                           const double* xy, const uint8_t* gate, int bound, void* const* peer_records,
                           int64_t* const* peer_counts, int rank, int world, int wire, int64_t capacity,
                           const int64_t* index_arr, const int32_t* px, cudaStream_t s) {
    PeerList P;
    memset(&P, 0, sizeof(P));
    P.world = world;
    P.rank = rank;
    P.wire = wire;
    for (int d = 0; d < world; ++d) {
        P.rec[d] = (uint8_t*)peer_records[d];
        P.cnt[d] = peer_counts[d];
    }
    return launch_compact_peers(ctx, N, index_base, c, nrm, ref, vis, avg, count, xy, gate, bound, P, capacity, index_arr,
                                px, s);
}

// ---------------------------------------------------------------------------------
// Receiver side of MVS_WIRE_COMPACT: rebuild full patch records.  n = (O_ref - c)/|O_ref - c| with
// the same correctly rounded operations as the candidate generator (expand.cu, MVS2.py:357-358),
// x, y = the reference projection in cv2's operation order (project.cuh) -- bit-identical to what
// the sender held.
// ---------------------------------------------------------------------------------
#include "project.cuh"

__global__ void __launch_bounds__(256)
    records_expand(const uint8_t* __restrict__ wire, int64_t n, int mw, const CamProj* __restrict__ cams,
                   const CamGeom* __restrict__ geom, int V, uint8_t* __restrict__ records) {
    const int wb = (int)sizeof(WireCompact) + 8 * mw, rb = (int)sizeof(mvs_patch_record) + 8 * mw;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const WireCompact* w = reinterpret_cast<const WireCompact*>(wire + i * wb);
        mvs_patch_record* r = reinterpret_cast<mvs_patch_record*>(records + i * rb);
        const int v = w->ref;
        const double c0 = w->c[0], c1 = w->c[1], c2 = w->c[2];
        r->c[0] = c0; r->c[1] = c1; r->c[2] = c2;
        double x = nan(""), y = nan(""), n0 = 0.0, n1 = 0.0, n2 = 0.0;
        if (v >= 0 && v < V) {
            const CamGeom& g = geom[v];
            const double q0 = __dsub_rn(g.C[0], c0), q1 = __dsub_rn(g.C[1], c1), q2 = __dsub_rn(g.C[2], c2);
            const double dist = sqrt(__dadd_rn(__dadd_rn(__dmul_rn(q0, q0), __dmul_rn(q1, q1)), __dmul_rn(q2, q2)));
            n0 = __ddiv_rn(q0, dist); n1 = __ddiv_rn(q1, dist); n2 = __ddiv_rn(q2, dist);
            project_ref(cams[v], c0, c1, c2, x, y);
        }
        r->n[0] = n0; r->n[1] = n1; r->n[2] = n2;
        r->xy[0] = x; r->xy[1] = y;
        r->avg = w->avg;
        r->ref = v;
        r->index = w->index;
        r->px[0] = w->px[0];
        r->px[1] = w->px[1];
        const uint64_t* wv = reinterpret_cast<const uint64_t*>(w + 1);
        uint64_t* rv = reinterpret_cast<uint64_t*>(r + 1);
        int cnt = 0;
        for (int q = 0; q < mw; ++q) {
            rv[q] = wv[q];
            cnt += __popcll(wv[q]);
        }
        r->count = cnt;
    }
}

int mvs_launch_records_expand(mvs_ctx* ctx, const void* wire, int64_t n, void* records, cudaStream_t s) {
    if (n == 0) return MVS_OK;
    int64_t blocks = (n + 255) / 256;
    const int64_t cap = (int64_t)ctx->sm_count * 16;
    if (blocks > cap) blocks = cap;
    records_expand<<<(int)blocks, 256, 0, s>>>((const uint8_t*)wire, n, (ctx->V + 63) / 64, ctx->d_cam, ctx->d_geom, ctx->V,
                                               (uint8_t*)records);
    ctx->launches++;
    MVS_CUDA_CHECK(cudaGetLastError());
    return MVS_OK;
}
