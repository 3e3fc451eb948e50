// sgr_binning.cu — tile binning: per-tile offsets (scan), work lists, and the per-tile depth sort that emits the
// depth-ordered id lists (upstream's point_list) and, per 8x4 pixel block of every tile, the depth-ordered list of
// (Gaussian id, tile-list position, quarter mask) entries consumed by the blend kernels.
//
// Replaces upstream's InclusiveSum + global 64-bit DeviceRadixSort + identifyTileRanges.  Upstream
// sorts ONE global array keyed (tile << 32 | depth bits); the stable sort resolves (tile, depth) ties by ascending
// Gaussian index.  Here instances are first counting-sorted by tile (atomics in sgr_preprocess.cu), then every tile is
// sorted independently on the 64-bit key (depth bits << 32 | gaussian index) — the same total order, so the per-tile
// lists are identical to upstream's ranges (bit-exact against the oracle's point_list).
//
// Per-tile sort = one MSD bucket pass over the tile's key range in shared memory followed by an exact rank inside
// each (small) bucket; linear work in the list length, no power-of-two padding.
#include <cstdlib>

#include "sgr_common.cuh"

namespace sgr {
namespace {

// ------------------------------------------------------------------------------------------------ plan
// One single-CTA kernel per chunk (replaces upstream's InclusiveSum + the num_rendered device->host copy, and plans
// everything that depends only on the per-tile instance counts):
//   1. exclusive scan of tile_cnt -> tile_off, instance range reserved on the device (overflow -> chunk dropped);
//   2. + 3. ONE work list of the non-empty tiles by descending size class: the forward blend kernel walks it longest-
//      processing-time first, the sort kernels split it at n_big (tiles with >= kSmallSortCap instances come first);
//      plus the list of empty tiles (background only);
//   4. the chunk's slice of the backward item lists (kept in `state`; the items themselves are pushed by the forward
//      blend, which knows how many records of every block segment really blended).
struct PlanArgs {
    int n;                        // tiles in the chunk (renders * tiles per render)
    int first_chunk;              // != 0: initialise the status header
    unsigned long long capacity;
    unsigned int item_region;     // first backward-item slot this chunk may use, before adding block records / kSegB
    unsigned long long blk_capacity;
    unsigned int* tile_cnt;       // chunk slice
    unsigned int* tile_off;       // chunk slice
    unsigned int* cursor;         // chunk scatter cursors (zeroed here)
    StateHeader* header;
    WorkCounts* wc;
    ChunkPlan* plan;
    unsigned int *work_blend, *work_empty;
};

constexpr int kScanThreads = 1024;
constexpr int kSizeClasses = 66;      // 2 * (bit length of the count) + next-lower bit; class 0 = empty

__device__ __forceinline__ int size_class(unsigned int c) {
    if (c == 0) return 0;
    const int msb = 31 - __clz(c);
    const int half = msb ? int((c >> (msb - 1)) & 1u) : 0;
    return 2 * (msb + 1) + half;
}

constexpr int kPlanIpt = 8;          // tiles per thread and strip of the scan sweep

__global__ void __launch_bounds__(kScanThreads) plan_kernel(PlanArgs a) {
    // The kernel is a latency chain on one SM, so it is written as TWO sweeps over tile_cnt whose loads are all
    // independent (one exposed L2 latency each), with the header fields fetched ahead:
    //   sweep 1 (coalesced): totals (instances, empty tiles, longest list, non-empty tiles) and the size-class histogram;
    //   thread 0: device-side reservation of the instance range, class starts, counters;
    //   sweep 2 (strips of 1024 * kPlanIpt tiles, carried scan): tile_off, zeroed cursors, and the two
    //   work lists — empty tiles at their scanned position, non-empty tiles at a shared-memory cursor of their class.
    // Both scanned quantities travel in one 64-bit word: instances (high 40 bits) and empty tiles (low 24 bits).
    __shared__ unsigned long long s_warp[32];
    __shared__ unsigned long long s_total, s_base;
    __shared__ unsigned long long s_red[32];
    __shared__ unsigned int s_max, s_nonempty, s_dropped;
    __shared__ unsigned int s_hist[kSizeClasses], s_start[kSizeClasses];
    const int t = threadIdx.x;
    // header fields of earlier chunks (their kernels precede this one in the stream), fetched ahead of the sweeps
    unsigned long long h_cursor = 0, h_capacity = 0, h_required = 0, h_blk_required = 0, h_blk_capacity = 0;
    unsigned int h_overflow = 0, h_max = 0, h_nonempty = 0;
    if (t == 0) {
        if (a.first_chunk) {
            h_capacity = a.capacity; h_blk_capacity = a.blk_capacity;
        } else {
            h_cursor = a.header->inst_cursor; h_capacity = a.header->capacity; h_required = a.header->inst_required;
            h_blk_required = a.header->blk_required; h_blk_capacity = a.header->blk_capacity;
            h_overflow = a.header->overflow; h_max = a.header->max_tile_instances; h_nonempty = a.header->nonempty_tiles;
        }
        s_max = 0; s_nonempty = 0;
    }
    if (t < kSizeClasses) s_hist[t] = 0;
    __syncthreads();
    // ---- sweep 1
    {
        unsigned int mx = 0, ne = 0;
        unsigned long long sum = 0;
        for (int k0 = 0; k0 < a.n; k0 += kScanThreads * kPlanIpt) {
            unsigned int c[kPlanIpt];
#pragma unroll
            for (int j = 0; j < kPlanIpt; ++j) {
                const int k = k0 + j * kScanThreads + t;
                c[j] = k < a.n ? a.tile_cnt[k] : 0xffffffffu;
            }
#pragma unroll
            for (int j = 0; j < kPlanIpt; ++j) {
                if (c[j] == 0xffffffffu) continue;          // beyond the chunk (a real count never reaches 2^32 - 1)
                sum += (static_cast<unsigned long long>(c[j]) << 24) + (c[j] == 0 ? 1ull : 0ull);
                mx = max(mx, c[j]);
                if (c[j] != 0) { ++ne; atomicAdd(&s_hist[size_class(c[j])], 1u); }
            }
        }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
            sum += __shfl_xor_sync(0xffffffffu, sum, d);
            mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, d));
            ne += __shfl_xor_sync(0xffffffffu, ne, d);
        }
        if ((t & 31) == 0) { s_red[t >> 5] = sum; atomicMax(&s_max, mx); atomicAdd(&s_nonempty, ne); }
    }
    __syncthreads();
    if (t == 0) {
        unsigned long long tot = 0;
#pragma unroll
        for (int w = 0; w < 32; ++w) tot += s_red[w];
        const unsigned long long total = tot >> 24;
        const bool fits = h_cursor + total <= h_capacity;
        s_dropped = fits ? 0u : 1u;
        s_base = h_cursor;
        a.header->inst_required = h_required + total;
        a.header->capacity = h_capacity;
        a.header->overflow = h_overflow | (fits ? 0u : 1u);
        a.header->max_tile_instances = fits ? max(h_max, s_max) : h_max;
        a.header->nonempty_tiles = h_nonempty + (fits ? s_nonempty : 0u);
        a.header->pad = 0;
        a.header->inst_cursor = fits ? h_cursor + total : h_cursor;
        if (a.first_chunk) { a.header->blk_required = 0; a.header->blk_capacity = h_blk_capacity; }
        a.wc->chunk_instances = fits ? static_cast<unsigned int>(total) : 0u;
        a.wc->chunk_dropped = fits ? 0u : 1u;
        unsigned int r1 = 0;
        for (int c = kSizeClasses - 1; c >= 1; --c) { s_start[c] = r1; r1 += s_hist[c]; }
        s_start[0] = 0;
        a.wc->n_big = fits ? s_start[size_class(kSmallSortCap) - 1] : 0u;   // tiles of the classes >= class(kSmallSortCap)
        a.wc->sort_cursor = 0;
        a.wc->n_blend = fits ? r1 : 0u;
        a.wc->n_empty = fits ? static_cast<unsigned int>(tot & 0xffffffull) : unsigned(a.n);
        a.wc->blend_cursor = 0;
        a.wc->empty_cursor = 0;
        a.wc->n_dense = 0;
        a.wc->dense_cursor = 0;
        // the block-record cursor is final for all earlier chunks here (their sorts precede this kernel in the stream)
#pragma unroll
        for (int c = 0; c < kBwdClasses; ++c) a.plan->n_items[c] = 0;
        const unsigned long long used = min(h_blk_required, h_blk_capacity);   // records really stored
        a.plan->item_base = a.item_region + static_cast<unsigned int>(used / kSegB);
        a.plan->cursor = 0;
    }
    __syncthreads();
    const bool dropped = s_dropped != 0;
    const unsigned int base = static_cast<unsigned int>(s_base);
    // ---- sweep 2: warp w of a strip owns 32 * kPlanIpt consecutive tiles, lane l the tiles base + 32 j + l — every
    // load and store of the sweep is coalesced (a single SM executes this kernel: its load/store unit is the limit)
    const int warp = t >> 5, lane = t & 31;
    unsigned long long carry = 0;                 // (instances << 24 | empty tiles) of the strips before this one
    for (int k0 = 0; k0 < a.n; k0 += kScanThreads * kPlanIpt) {
        const int wbase = k0 + warp * 32 * kPlanIpt;
        unsigned int c[kPlanIpt];
#pragma unroll
        for (int j = 0; j < kPlanIpt; ++j) {
            const int k = wbase + 32 * j + lane;
            c[j] = k < a.n ? a.tile_cnt[k] : 0xffffffffu;
        }
        unsigned long long ex[kPlanIpt];           // exclusive prefix inside the warp's range
        unsigned long long wsum = 0;
#pragma unroll
        for (int j = 0; j < kPlanIpt; ++j) {
            const unsigned long long v =
                c[j] == 0xffffffffu ? 0ull : (static_cast<unsigned long long>(c[j]) << 24) + (c[j] == 0 ? 1ull : 0ull);
            unsigned long long inc = v;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const unsigned long long u = __shfl_up_sync(0xffffffffu, inc, d);
                if (lane >= d) inc += u;
            }
            ex[j] = wsum + inc - v;
            wsum += __shfl_sync(0xffffffffu, inc, 31);
        }
        __syncthreads();                          // s_warp / s_total of the previous strip have been read
        if (lane == 0) s_warp[warp] = wsum;
        __syncthreads();
        if (t < 32) {
            unsigned long long w = s_warp[t], wi = w;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const unsigned long long u = __shfl_up_sync(0xffffffffu, wi, d);
                if (t >= d) wi += u;
            }
            s_warp[t] = wi - w;                   // exclusive warp prefix
            if (t == 31) s_total = wi;
        }
        __syncthreads();
        const unsigned long long wex = carry + s_warp[warp];
#pragma unroll
        for (int j = 0; j < kPlanIpt; ++j) {
            if (c[j] == 0xffffffffu) continue;
            const int k = wbase + 32 * j + lane;
            const unsigned long long e64 = wex + ex[j];
            a.cursor[k] = 0;
            if (dropped) {                        // render(s) emitted as background; reported via the status block
                a.tile_cnt[k] = 0;
                a.tile_off[k] = base;
                a.work_empty[k] = k;
            } else {
                a.tile_off[k] = base + static_cast<unsigned int>(e64 >> 24);
                if (c[j] == 0) a.work_empty[static_cast<unsigned int>(e64 & 0xffffffull)] = k;
                else a.work_blend[atomicAdd(&s_start[size_class(c[j])], 1u)] = k;
            }
        }
        carry += s_total;
    }
}

// ------------------------------------------------------------------------------------------------ sort
struct SortArgs {
    int N, num_tiles, tiles_x, render_base;
    const unsigned int* tile_off;    // global arrays
    const unsigned int* tile_cnt;
    const unsigned long long* keys;
    unsigned long long* keys_w;      // the first key buffer, writable (scratch of spilled lists once their keys are dead)
    unsigned long long* keys_tmp;    // second key buffer [cap]: bucket-ordered keys of lists beyond the shared memory
    unsigned int* sorted_ids;
    float4 *rec0, *rec1, *rec2;      // tile-level records in depth order
    const float4 *g0, *g1, *g2;
    const unsigned int* work;        // chunk-local indices of the non-empty tiles, longest list first
    WorkCounts* wc;                  // n_big: the first n_big tiles go to the big kernel; sort_cursor: small queue
    unsigned int big_smem_keys;      // key capacity of the big kernel's shared-memory buffer
    int simple;                      // SGR_FLAG_SIMPLE_BLEND: no block lists
    // block lists (simple == 0)
    StateHeader* header;
    unsigned int *blk_off, *blk_cnt;
    uint2* bidx;                     // block-list entries (Gaussian id, position in the tile list << 4 | quarter mask)
    unsigned long long blk_capacity;
    unsigned int* dense_items;       // block items of the dense lists (sgr_common.cuh::kDenseEntries)
};

// Experiment build (-DSGR_SORT_TIMING, tools/sort_timing.py): per-tile phase timestamps of the tile sorts.
#ifdef SGR_SORT_TIMING
constexpr int kSortMarks = 10;
__device__ unsigned long long g_sort_marks[4096][kSortMarks + 2];
__device__ unsigned int g_sort_mark_cursor;
#define SORT_MARK(i) do { if (threadIdx.x == 0 && mark_slot < 4096u) { unsigned long long tt; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tt)); g_sort_marks[mark_slot][i] = tt; } } while (0)
#else
#define SORT_MARK(i) do {} while (0)
#endif

template <int THREADS>
__device__ __forceinline__ void block_minmax(unsigned long long& mn, unsigned long long& mx, unsigned long long* s_red) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        const unsigned long long a = __shfl_xor_sync(0xffffffffu, mn, d);
        const unsigned long long b = __shfl_xor_sync(0xffffffffu, mx, d);
        mn = a < mn ? a : mn;
        mx = b > mx ? b : mx;
    }
    const int w = threadIdx.x >> 5;
    __syncthreads();                               // s_red may still be read from a previous use
    if ((threadIdx.x & 31) == 0) { s_red[2 * w] = mn; s_red[2 * w + 1] = mx; }
    __syncthreads();
    if (w == 0) {                                  // second stage: one warp reduces the per-warp results
        const int l = threadIdx.x;
        unsigned long long rmn = l < THREADS / 32 ? s_red[2 * l] : ~0ull;
        unsigned long long rmx = l < THREADS / 32 ? s_red[2 * l + 1] : 0ull;
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
            const unsigned long long a = __shfl_xor_sync(0xffffffffu, rmn, d);
            const unsigned long long b = __shfl_xor_sync(0xffffffffu, rmx, d);
            rmn = a < rmn ? a : rmn;
            rmx = b > rmx ? b : rmx;
        }
        if (l == 0) { s_red[2 * (THREADS / 32)] = rmn; s_red[2 * (THREADS / 32) + 1] = rmx; }
    }
    __syncthreads();
    mn = s_red[2 * (THREADS / 32)];
    mx = s_red[2 * (THREADS / 32) + 1];
}

// In-place exclusive scan of hist[NB] (NB = THREADS * PER).
template <int THREADS, int NB>
__device__ __forceinline__ void block_exclusive_scan(unsigned int* hist, unsigned int* s_warp) {
    constexpr int PER = NB / THREADS;
    static_assert(NB % THREADS == 0, "bucket count must be a multiple of the block size");
    const int t = threadIdx.x;
    unsigned int v[PER], sum = 0;
#pragma unroll
    for (int j = 0; j < PER; ++j) { v[j] = hist[t * PER + j]; sum += v[j]; }
    unsigned int inc = sum;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const unsigned int u = __shfl_up_sync(0xffffffffu, inc, d);
        if ((t & 31) >= d) inc += u;
    }
    if ((t & 31) == 31) s_warp[t >> 5] = inc;
    __syncthreads();
    if (t < 32) {
        unsigned int w = (t < THREADS / 32) ? s_warp[t] : 0u, wi = w;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const unsigned int u = __shfl_up_sync(0xffffffffu, wi, d);
            if (t >= d) wi += u;
        }
        if (t < THREADS / 32) s_warp[t] = wi - w;
    }
    __syncthreads();
    unsigned int run = inc - sum + s_warp[t >> 5];
#pragma unroll
    for (int j = 0; j < PER; ++j) { hist[t * PER + j] = run; run += v[j]; }
    __syncthreads();
}

// Sorts the n keys of one tile: writes the depth-ordered ids (upstream's point_list) and then either the tile's
// gathered 48-byte records (SGR_FLAG_SIMPLE_BLEND) or the eight per-block entry lists.  kb: n-element key buffer (shared or global).
// KPT > 0: the tile has at most THREADS * KPT keys and every thread keeps its keys in registers, so the three passes
// over the unsorted keys (range, histogram, scatter) cost ONE exposed global-memory latency instead of three (the
// per-tile sort is a latency chain, not a bandwidth problem); KPT == 0 re-reads the keys from global memory.
template <int THREADS, int NB, int KPT>
__device__ __forceinline__ void sort_one_tile(const SortArgs& a, int tile_local, unsigned long long* kb,
                                              unsigned int* cpre_buf, unsigned int* hist, unsigned int* s_warp, unsigned long long* s_red) {
    const int t = threadIdx.x;
    const int rl = tile_local / a.num_tiles;
    const size_t tg = size_t(a.render_base) * a.num_tiles + tile_local;      // global tile index
    const unsigned int n = a.tile_cnt[tg];
    const size_t off = a.tile_off[tg];
    const unsigned long long* keys = a.keys + off;
    constexpr int KR = KPT > 0 ? KPT : 1;
    unsigned long long kr[KR];
#ifdef SGR_SORT_TIMING
    __shared__ unsigned int s_mark_slot;
    if (t == 0) {
        s_mark_slot = atomicAdd(&g_sort_mark_cursor, 1u);
        if (s_mark_slot < 4096u) { g_sort_marks[s_mark_slot][kSortMarks] = n; g_sort_marks[s_mark_slot][kSortMarks + 1] = THREADS; }
    }
    __syncthreads();
    const unsigned int mark_slot = s_mark_slot;
#endif
    SORT_MARK(0);

    // 0. key range
    unsigned long long mn = ~0ull, mx = 0ull;
    if (KPT > 0) {
#pragma unroll
        for (int j = 0; j < KR; ++j) {
            const unsigned int k = t + j * THREADS;
            kr[j] = k < n ? keys[k] : 0ull;
        }
#pragma unroll
        for (int j = 0; j < KR; ++j)
            if (t + j * THREADS < n) { mn = kr[j] < mn ? kr[j] : mn; mx = kr[j] > mx ? kr[j] : mx; }
    } else {
        for (unsigned int k = t; k < n; k += THREADS) {
            const unsigned long long key = keys[k];
            mn = key < mn ? key : mn;
            mx = key > mx ? key : mx;
        }
    }
    for (int k = t; k < NB; k += THREADS) hist[k] = 0;     // visible after the barriers inside block_minmax
    block_minmax<THREADS>(mn, mx, s_red);
    const unsigned long long range = mx - mn;
    const int bits = range ? 64 - __clzll(static_cast<long long>(range)) : 0;
    constexpr int LOG_NB = (NB == 1024) ? 10 : (NB == 2048 ? 11 : 12);
    static_assert((1 << LOG_NB) == NB, "NB must be 1024, 2048 or 4096");
    const int shift = max(0, bits - LOG_NB);

    SORT_MARK(1);
    // 1. bucket histogram
    if (KPT > 0) {
#pragma unroll
        for (int j = 0; j < KR; ++j)
            if (t + j * THREADS < n) atomicAdd(&hist[static_cast<unsigned int>((kr[j] - mn) >> shift)], 1u);
    } else {
        for (unsigned int k = t; k < n; k += THREADS)
            atomicAdd(&hist[static_cast<unsigned int>((keys[k] - mn) >> shift)], 1u);
    }
    __syncthreads();
    SORT_MARK(2);
    // 2. bucket starts
    block_exclusive_scan<THREADS, NB>(hist, s_warp);
    SORT_MARK(3);
    // 3. scatter into bucket order (arbitrary order inside a bucket); afterwards hist[b] = end of bucket b
    if (KPT > 0) {
#pragma unroll
        for (int j = 0; j < KR; ++j)
            if (t + j * THREADS < n)
                kb[atomicAdd(&hist[static_cast<unsigned int>((kr[j] - mn) >> shift)], 1u)] = kr[j];
    } else {
        for (unsigned int k = t; k < n; k += THREADS) {
            const unsigned long long key = keys[k];
            const unsigned int pos = atomicAdd(&hist[static_cast<unsigned int>((key - mn) >> shift)], 1u);
            kb[pos] = key;
        }
    }
    __syncthreads();
    SORT_MARK(4);
    // 4a. exact rank inside the bucket -> final position of the id (upstream's point_list).
    const size_t gb = size_t(rl) * a.N;
    const int tile = tile_local - rl * a.num_tiles;
    const float X0 = float((tile % a.tiles_x) * kTile), Y0 = float((tile / a.tiles_x) * kTile);
#pragma unroll 4
    for (unsigned int p = t; p < n; p += THREADS) {
        const unsigned long long key = kb[p];
        const unsigned int b = static_cast<unsigned int>((key - mn) >> shift);
        const unsigned int s = b ? hist[b - 1] : 0u;
        const unsigned int e = hist[b];
        unsigned int cnt = 0;
        for (unsigned int j = s; j < e; ++j) cnt += (kb[j] < key) ? 1u : 0u;
        a.sorted_ids[off + s + cnt] = static_cast<unsigned int>(key & 0xffffffffull);
    }
    __syncthreads();                             // the CTA's own global stores are visible to it behind the barrier
    SORT_MARK(5);
    // 4b. thread = sorted position: the instance's cull mask for this tile from word 0 of its Gaussian's record
    //     (one random 16-byte load; the ids are read back coalesced).  The key buffer is dead after the ranking: it
    //     takes the ids and the quarter masks in sorted order for the block-list emission.  SGR_FLAG_SIMPLE_BLEND
    //     instead copies the 48-byte records into depth order (the tile-level stream of the upstream-shaped kernels;
    //     word 2 of rec0 becomes the position in the tile list, upstream's contributor index).
    unsigned int* ids_s = reinterpret_cast<unsigned int*>(kb);
    unsigned int* masks = ids_s + n;
#pragma unroll 4
    for (unsigned int p = t; p < n; p += THREADS) {
        const unsigned int id = a.sorted_ids[off + p];
        float4 v0 = __ldg(a.g0 + gb + id);
        if (a.simple) {
            const float4 v1 = __ldg(a.g1 + gb + id);
            const float4 v2 = __ldg(a.g2 + gb + id);
            v0.z = __uint_as_float(p);
            a.rec0[off + p] = v0;
            a.rec1[off + p] = v1;
            a.rec2[off + p] = v2;
        } else {
            ids_s[p] = id;
            masks[p] = quarter_mask(v0.x, v0.y, v0.z, X0, Y0);
        }
    }
    __syncthreads();
    SORT_MARK(6);
    if (a.simple) return;

    // 5. block lists.  Every instance is appended to the list of each 8x4 pixel block of the tile that its
    //    conservative alpha >= 1/255 extent touches (quarter_mask), keeping the depth order: per chunk of 32 sorted
    //    instances (lane = record) one ballot per block, a scan of the per-chunk counts, then the copies.  A block-list
    //    entry is 8 bytes: (Gaussian id, position in the tile list << 4 | the block's 4-bit quarter mask); the blend
    //    kernels gather the Gaussians' records themselves (cp.async) from the per-(render, Gaussian) arrays, so the
    //    48-byte records are never copied per instance.  Every list starts at a multiple of 4 entries (the index
    //    batches travel by 1-D TMA: 16-byte granularity).
    const size_t tgb = tg * kBlocksPerTile;
    const unsigned int nchunks = (n + 31u) >> 5;
    unsigned int* cpre = cpre_buf;                // [nchunks][8] counts -> exclusive prefixes
    const int warp = t >> 5, lane = t & 31;
    constexpr int kWarps = THREADS / 32;
    for (unsigned int c0 = warp; c0 < nchunks; c0 += 4 * kWarps) {      // four chunks in flight per warp
        unsigned int mk[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const unsigned int p = 32u * (c0 + k * kWarps) + lane;
            mk[k] = p < n ? masks[p] : 0u;
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const unsigned int c = c0 + k * kWarps;
            if (c >= nchunks) break;
            unsigned int mine = 0;
#pragma unroll
            for (int b = 0; b < kBlocksPerTile; ++b) {
                const unsigned int bal = __ballot_sync(0xffffffffu, (mk[k] >> (4 * b)) & 0xfu);
                if (lane == b) mine = __popc(bal);
            }
            if (lane < kBlocksPerTile) cpre[c * kBlocksPerTile + lane] = mine;
        }
    }
    __syncthreads();
    SORT_MARK(7);
    if (warp < kBlocksPerTile) {                 // warp w: exclusive scan of block w's chunk counts
        unsigned int running = 0;
        for (unsigned int c0 = 0; c0 < nchunks; c0 += 32) {
            const unsigned int c = c0 + lane;
            const unsigned int v = c < nchunks ? cpre[c * kBlocksPerTile + warp] : 0u;
            unsigned int inc = v;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const unsigned int u = __shfl_up_sync(0xffffffffu, inc, d);
                if (lane >= d) inc += u;
            }
            if (c < nchunks) cpre[c * kBlocksPerTile + warp] = running + inc - v;
            running += __shfl_sync(0xffffffffu, inc, 31);
        }
        if (lane == 0) s_warp[warp] = running;
    }
    __syncthreads();
    if (t == 0) {                                // reserve the tile's block-list entries (bump allocation)
        unsigned int total = 0;
#pragma unroll
        for (int b = 0; b < kBlocksPerTile; ++b) total += (s_warp[b] + 3u) & ~3u;
        const unsigned long long base = atomicAdd(&a.header->blk_required, static_cast<unsigned long long>(total));
        const bool fits = base + total <= a.blk_capacity;
        if (!fits) atomicOr(&a.header->overflow, 2u);     // the tile's blocks stay empty; reported via the status
        unsigned int run = static_cast<unsigned int>(base);
#pragma unroll
        for (int b = 0; b < kBlocksPerTile; ++b) {
            a.blk_off[tgb + b] = fits ? run : 0u;
            a.blk_cnt[tgb + b] = fits ? s_warp[b] : 0u;
            s_warp[8 + b] = run;
            run += (s_warp[b] + 3u) & ~3u;
        }
        s_warp[16] = fits ? 1u : 0u;
    }
    __syncthreads();
    if (t < kBlocksPerTile && s_warp[16] && s_warp[t] >= kDenseEntries)      // a dense list: queued separately
        a.dense_items[atomicAdd(&a.wc->n_dense, 1u)] = unsigned(tile_local) * kBlocksPerTile + t;
    SORT_MARK(8);
    if (s_warp[16]) {
        const unsigned int lt = (1u << lane) - 1u;
        for (unsigned int c0 = warp; c0 < nchunks; c0 += 4 * kWarps) {      // four chunks in flight per warp
            unsigned int mk[4], idk[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const unsigned int p = 32u * (c0 + k * kWarps) + lane;
                mk[k] = p < n ? masks[p] : 0u;
                idk[k] = p < n ? ids_s[p] : 0u;
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const unsigned int c = c0 + k * kWarps;
                if (c >= nchunks) break;
                const unsigned int p = 32u * c + lane;
                // lane b < 8 holds where block b's entries of this chunk start
                const unsigned int mybase = lane < kBlocksPerTile ? s_warp[8 + lane] + cpre[c * kBlocksPerTile + lane] : 0u;
#pragma unroll
                for (int b = 0; b < kBlocksPerTile; ++b) {
                    const unsigned int nib = (mk[k] >> (4 * b)) & 0xfu;
                    const unsigned int bal = __ballot_sync(0xffffffffu, nib);
                    const unsigned int base = __shfl_sync(0xffffffffu, mybase, b);
                    if (nib) a.bidx[base + __popc(bal & lt)] = make_uint2(idk[k], (p << 4) | nib);
                }
            }
        }
        // the pad entries (up to 3 per list) are read by the 16-byte granular index copies: quarter mask 0
        if (t < kBlocksPerTile) {
            const unsigned int cnt_b = s_warp[t];
            for (unsigned int k = cnt_b; k < ((cnt_b + 3u) & ~3u); ++k) a.bidx[size_t(s_warp[8 + t]) + k] = make_uint2(0u, 0u);
        }
    }
    __syncthreads();                             // the buffers are reused by the CTA's next tile
    SORT_MARK(9);
}

#ifndef SGR_SORT_SMALL_MIN_CTAS
#define SGR_SORT_SMALL_MIN_CTAS 4
#endif
__global__ void __launch_bounds__(kSmallSortThreads, SGR_SORT_SMALL_MIN_CTAS) sort_small_kernel(SortArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];                // kSmallSortSmem bytes
    unsigned long long* kb = reinterpret_cast<unsigned long long*>(smem_raw);                   // kSmallSortCap keys
    unsigned int* hist = reinterpret_cast<unsigned int*>(kb + kSmallSortCap);                   // kSmallSortBuckets
    __shared__ unsigned int s_warp[32];
    __shared__ unsigned long long s_red[2 * (kSmallSortThreads / 32) + 2];
    __shared__ unsigned int s_item;
    const unsigned int first = a.wc->n_big, nw = a.wc->n_blend;
    for (;;) {                                   // dynamic queue, longest lists first
        __syncthreads();
        if (threadIdx.x == 0) s_item = first + atomicAdd(&a.wc->sort_cursor, 1u);
        __syncthreads();
        const unsigned int w = s_item;
        if (w >= nw) break;
        const int tile_local = a.work[w];
        const size_t tg = size_t(a.render_base) * a.num_tiles + tile_local;
        const unsigned int n = a.tile_cnt[tg];
        if (n <= 4u * kSmallSortThreads)
            sort_one_tile<kSmallSortThreads, kSmallSortBuckets, 4>(a, tile_local, kb, hist, hist, s_warp, s_red);
        else
            sort_one_tile<kSmallSortThreads, kSmallSortBuckets, kSmallSortCap / kSmallSortThreads>(
                a, tile_local, kb, hist, hist, s_warp, s_red);
    }
    // Programmatic dependent launch (launch_sort_tiles): this grid started before sort_big_kernel finished.  The
    // blend behind it in the stream is ordered after THIS grid only, so every CTA waits here for the long-list
    // grid to complete and flush before it exits: completion of sort_small then implies completion of sort_big
    // (also under graph capture, where the pair becomes a programmatic edge followed by a plain edge).  Without the
    // attribute (SGR_SORT_MODE experiments) the wait returns immediately.
    asm volatile("griddepcontrol.wait;" ::: "memory");
}

__global__ void __launch_bounds__(kBigSortThreads) sort_big_kernel(SortArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    unsigned long long* kb_s = reinterpret_cast<unsigned long long*>(smem_raw);                 // big_smem_keys
    unsigned int* hist = reinterpret_cast<unsigned int*>(kb_s + a.big_smem_keys);               // kBigSortBuckets
    // Programmatic dependent launch: the short-list kernel behind this one in the stream may start as soon as every
    // CTA of this grid is resident (it does not consume this kernel's output) — see launch_sort_tiles.
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    __shared__ unsigned int s_warp[32];
    __shared__ unsigned long long s_red[2 * (kBigSortThreads / 32) + 2];
    const unsigned int nw = a.wc->n_big;
    for (unsigned int w = blockIdx.x; w < nw; w += gridDim.x) {
        const int tile_local = a.work[w];
        const size_t tg = size_t(a.render_base) * a.num_tiles + tile_local;
        const unsigned int n = a.tile_cnt[tg];
        // lists that do not fit in shared memory use their segment of the second global key buffer
        const bool spill = n > a.big_smem_keys;
        unsigned long long* kb = spill ? a.keys_tmp + a.tile_off[tg] : kb_s;
        // per-chunk block counts: the bucket histogram (dead after the ranking, 4 bytes per bucket >= 1 byte per key);
        // a spilled list uses its own segment of the first key buffer, dead once the keys sit in kb
        unsigned int* cpre = spill ? reinterpret_cast<unsigned int*>(a.keys_w + a.tile_off[tg]) : hist;
        if (n <= 8u * kBigSortThreads && !spill)
            sort_one_tile<kBigSortThreads, kBigSortBuckets, 8>(a, tile_local, kb, cpre, hist, s_warp, s_red);
        else
            sort_one_tile<kBigSortThreads, kBigSortBuckets, 0>(a, tile_local, kb, cpre, hist, s_warp, s_red);
    }
}

}  // namespace

#ifdef SGR_SORT_TIMING
extern "C" int sgr_debug_sort_marks(unsigned long long* out, unsigned int* count) {   // experiment build only
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(count, g_sort_mark_cursor, sizeof(unsigned int));
    cudaMemcpyFromSymbol(out, g_sort_marks, sizeof(unsigned long long) * 4096 * (kSortMarks + 2));
    unsigned int zero = 0;
    cudaMemcpyToSymbol(g_sort_mark_cursor, &zero, sizeof(zero));
    return 0;
}
#endif

cudaError_t launch_plan(const ChunkCtx& c) {
    PlanArgs a;
    const size_t base = size_t(c.render_base) * c.g.num_tiles;
    a.n = c.num_renders * c.g.num_tiles;
    a.first_chunk = c.render_base == 0 ? 1 : 0;
    a.capacity = c.p->max_instances;
    // every chunk owns the item slots [8 * render_base * T + chunk_index + block records before it / kSegB, ...) of
    // every class: a chunk with n tiles and m block records emits at most 8 n + m / kSegB items
    a.item_region = static_cast<unsigned int>(base * kBlocksPerTile + size_t(c.chunk_index));
    a.blk_capacity = c.blk_capacity;
    a.tile_cnt = c.tile_cnt + base;
    a.tile_off = c.tile_off + base;
    a.cursor = c.cursor;
    a.header = c.header;
    a.wc = c.work_counts;
    a.plan = c.plan;
    a.work_blend = c.work_blend;
    a.work_empty = c.work_empty;
    plan_kernel<<<1, kScanThreads, 0, c.stream>>>(a);
    return cudaGetLastError();
}

cudaError_t launch_sort_tiles(const ChunkCtx& c) {
    SortArgs a;
    a.N = c.g.N; a.num_tiles = c.g.num_tiles; a.tiles_x = c.g.tiles_x; a.render_base = c.render_base;
    a.tile_off = c.tile_off; a.tile_cnt = c.tile_cnt; a.keys = c.keys; a.sorted_ids = c.sorted_ids;
    a.rec0 = c.rec0; a.rec1 = c.rec1; a.rec2 = c.rec2; a.g0 = c.g0; a.g1 = c.g1; a.g2 = c.g2;
    a.work = c.work_blend; a.wc = c.work_counts;
    a.keys_tmp = c.keys_tmp; a.keys_w = c.keys;
    a.simple = (c.p->flags & SGR_FLAG_SIMPLE_BLEND) ? 1 : 0;
    a.header = c.header; a.blk_off = c.blk_off; a.blk_cnt = c.blk_cnt;
    a.bidx = c.bidx; a.blk_capacity = c.blk_capacity; a.dense_items = c.dense_items;
    const int dslot = current_device_slot();
    static int num_sms_dev[kMaxDevices] = {};
    static bool attr_set_dev[kMaxDevices] = {};
    static int small_per_sm_dev[kMaxDevices] = {};
    int& num_sms = num_sms_dev[dslot];
    bool& attr_set = attr_set_dev[dslot];
    int& small_per_sm = small_per_sm_dev[dslot];
    if (num_sms == 0) {
        cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dslot);
        if (num_sms <= 0) num_sms = 148;
    }
    // Shared-memory key buffer of the big kernel: sized from the caller's hint of the longest list (a CTA that takes
    // the whole SM's shared memory keeps the small-tile CTAs off that SM); longer lists spill to global memory.
    unsigned int big_keys = kBigSortSmemCap;
    if (c.p->max_tile_instances_hint > 0) {
        const unsigned long long want = (unsigned long long)c.p->max_tile_instances_hint * 5 / 4;
        big_keys = unsigned(want < 8192 ? 8192 : (want > kBigSortSmemCap ? kBigSortSmemCap : want));
    }
    a.big_smem_keys = big_keys;
    const size_t big_smem = size_t(big_keys) * 8 + size_t(kBigSortBuckets) * 4;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(sort_big_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             int(size_t(kBigSortSmemCap) * 8 + size_t(kBigSortBuckets) * 4));
        if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(sort_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(kSmallSortSmem));
        if (e != cudaSuccess) return e;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&small_per_sm, sort_small_kernel, kSmallSortThreads, kSmallSortSmem);
        if (e != cudaSuccess) return e;
        if (small_per_sm <= 0) small_per_sm = 1;
        attr_set = true;
    }
    const int total_tiles = c.num_renders * c.g.num_tiles;
    cudaError_t e;
    static int mode = -1;      // experiment switch: 0 PDL overlap (default), 1 serial, 2 small only, 3 big only, 4 two streams
    if (mode < 0) { const char* m = getenv("SGR_SORT_MODE"); mode = m ? atoi(m) : 0; }
    if (mode >= 1 && mode <= 3) {
        if (mode != 2) sort_big_kernel<<<min(num_sms, total_tiles), kBigSortThreads, big_smem, c.stream>>>(a);
        if (mode != 3) sort_small_kernel<<<min(num_sms * small_per_sm, total_tiles), kSmallSortThreads, kSmallSortSmem, c.stream>>>(a);
        return cudaGetLastError();
    }
    // The long-list kernel goes first (it is the longer pole and a 1024-thread CTA cannot share an SM with the
    // short-list CTAs, so it must not queue behind them); the short-list kernel follows in the SAME stream as a
    // programmatic dependent launch: it starts once all long-list CTAs are resident and fills the remaining SMs.  Its
    // CTAs execute griddepcontrol.wait before they exit, so the next launch in the stream (the blend), which is
    // ordered behind the short-list grid, is transitively ordered behind the long-list grid as well.
    if (mode == 0) {
        sort_big_kernel<<<min(num_sms, total_tiles), kBigSortThreads, big_smem, c.stream>>>(a);
        if ((e = cudaGetLastError()) != cudaSuccess) return e;
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(min(num_sms * small_per_sm, total_tiles));
        cfg.blockDim = dim3(kSmallSortThreads);
        cfg.dynamicSmemBytes = kSmallSortSmem;
        cfg.stream = c.stream;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        return cudaLaunchKernelEx(&cfg, sort_small_kernel, a);
    }
    // mode 4 (experiment): the same overlap with a side stream and events instead of the dependent launch
    static cudaStream_t side_dev[kMaxDevices] = {};
    static cudaEvent_t ev_fork_dev[kMaxDevices] = {}, ev_join_dev[kMaxDevices] = {};
    cudaStream_t& side = side_dev[dslot];
    cudaEvent_t &ev_fork = ev_fork_dev[dslot], &ev_join = ev_join_dev[dslot];
    if (!side) {
        if ((e = cudaStreamCreateWithFlags(&side, cudaStreamNonBlocking)) != cudaSuccess) return e;
        if ((e = cudaEventCreateWithFlags(&ev_fork, cudaEventDisableTiming)) != cudaSuccess) return e;
        if ((e = cudaEventCreateWithFlags(&ev_join, cudaEventDisableTiming)) != cudaSuccess) return e;
    }
    if ((e = cudaEventRecord(ev_fork, c.stream)) != cudaSuccess) return e;
    if ((e = cudaStreamWaitEvent(side, ev_fork, 0)) != cudaSuccess) return e;
    sort_big_kernel<<<min(num_sms, total_tiles), kBigSortThreads, big_smem, c.stream>>>(a);
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    sort_small_kernel<<<min(num_sms * small_per_sm, total_tiles), kSmallSortThreads, kSmallSortSmem, side>>>(a);
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    if ((e = cudaEventRecord(ev_join, side)) != cudaSuccess) return e;
    return cudaStreamWaitEvent(c.stream, ev_join, 0);
}

}  // namespace sgr
