// sgr_common.cuh — shared definitions of the sm_100a rasteriser kernels (internal; the public ABI is include/sgr.h).
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/sgr.h"

namespace sgr {

constexpr int kTile = 16;                 // BLOCK_X = BLOCK_Y of the published algorithm
constexpr int kTilePixels = kTile * kTile;
constexpr int kDefaultRendersPerChunk = 16;
constexpr int kBlocksPerTile = 8;            // a 16x16 tile = eight 8x4 pixel blocks (blk = (y block) * 2 + (x block))
// Block lists: the per-tile sort emits, for every 8x4 pixel block of a tile, the depth-ordered sub-list of the tile's
// instances whose conservative extent touches the block, as 8-byte entries (Gaussian id, position in the tile list
// << 4 | the block's 4-bit quarter mask).  The blend kernels stream the entries and gather exactly those Gaussians'
// 48-byte records from the per-(render, Gaussian) arrays the preprocess kernel wrote (cp.async) — a record is written
// once per (render, Gaussian), never copied per tile or per block, and a warp never fetches or culls a record that
// cannot touch its pixels.
// Backward work granularity: a block list is replayed in independent segments of kSegB block records (multiple of the
// blend kernels' batch).  The forward blend checkpoints every pixel's running state at the segment boundaries of
// lists longer than one segment; slot (block list, s) = blk_off / (kSegB / 2) + s, s < #segments, the last slot
// holding the final state (a list of m > kSegB records spans at least ceil(m / kSegB) slots of that numbering).
#ifndef SGR_SEGB
#define SGR_SEGB 256
#endif
constexpr int kSegB = SGR_SEGB;
constexpr int kCkptPerSlot = 32;             // one entry per pixel of the block
constexpr int kBwdClasses = 4;               // backward items by number of records that really blended (most first)
// Dense blocks.  A warp walks a block list at its own issue rate — about 0.18 instructions per cycle next to three
// other warps on its scheduler, 0.27 alone — so the ~100 longest lists of a launch (tools/fwd_trace.py: 140-160 us
// each, started at t = 0) outlast the throughput-bound part of the forward blend (everything else is done after
// ~105 us).  The tile sort therefore queues the block lists of at least kDenseEntries entries separately, and the first
// ceil(#dense / kDenseWarps) CTAs of the forward blend run only kDenseWarps warps (one or two per scheduler) on them;
// their other warps sleep until the CTA's dense walks are over.  A scheduling decision only: no arithmetic changes.
#ifndef SGR_FWD_DENSE_ENTRIES
#define SGR_FWD_DENSE_ENTRIES 1536
#endif
#ifndef SGR_FWD_DENSE_WARPS
#define SGR_FWD_DENSE_WARPS 4
#endif
#ifndef SGR_FWD_DENSE_MATES
#define SGR_FWD_DENSE_MATES 4
#endif
#ifndef SGR_FWD_DENSE_EMPTY
#define SGR_FWD_DENSE_EMPTY 0
#endif
constexpr unsigned int kDenseEntries = SGR_FWD_DENSE_ENTRIES;
constexpr int kDenseWarps = SGR_FWD_DENSE_WARPS;
constexpr int kDenseMates = SGR_FWD_DENSE_MATES;            // further warps of a dense CTA that work on ordinary items meanwhile
constexpr bool kDenseWaitersDoEmpty = SGR_FWD_DENSE_EMPTY;  // the waiting warps process background tiles before they sleep

// ------------------------------------------------------------------------------------------------
// Device-resident status / counters at the start of `state`.
// The first sizeof(SgrStatus) bytes mirror SgrStatus (include/sgr.h) so a caller may also fetch the status with its
// own asynchronous copy of the head of `state`.
struct alignas(256) StateHeader {
    unsigned long long inst_required;      // total instances all renders need (even when overflowing)
    unsigned long long capacity;
    unsigned int overflow;                 // bit 0: instances, bit 1: block records
    unsigned int max_tile_instances;
    unsigned int nonempty_tiles;
    unsigned int pad;
    unsigned long long blk_required;       // block records all renders need (bump cursor of the tile sorts)
    unsigned long long blk_capacity;
    unsigned long long inst_cursor;        // instances consumed so far (global offset of the next chunk)
};
static_assert(sizeof(SgrStatus) == 48, "SgrStatus layout is part of the ABI");

// Backward work lists of one chunk (kept in `state`; filled by the forward blend, consumed by the backward blend):
// class c holds n_items[c] items at bwd_items[c * items_stride + item_base ...].
struct ChunkPlan {
    unsigned int n_items[kBwdClasses];
    unsigned int item_base, cursor, pad0, pad1;
};

// Layout of `state` (kept forward -> backward) for a problem shape.
struct StateLayout {
    uint64_t header, tile_off, tile_cnt, tile_time, n_contrib, clamp_mask, sorted_ids, g0, g1, g2, rec0, rec1, rec2, blk_off,
        blk_cnt, blk_eff, bidx, ck0, ck1, plan, bwd_items, bwd_items_stride, total;
};
// Layout of `scratch` (valid only inside one call).
struct ScratchLayout {
    uint64_t keys, keys_tmp, rect, cursor, work_blend, work_empty, work_counts, dense_items, loss_part, accum, total;
};

__host__ __device__ inline uint64_t align_up(uint64_t x, uint64_t a = 256) { return (x + a - 1) / a * a; }

inline int tiles_x(int W) { return (W + kTile - 1) / kTile; }
inline int tiles_y(int H) { return (H + kTile - 1) / kTile; }
inline uint64_t ckpt_slots(uint64_t capB) { return capB / (kSegB / 2) + 2; }

// `simple`: SGR_FLAG_SIMPLE_BLEND keeps tile-level copies of the records (the upstream-shaped kernels walk whole tile
// lists) and no block lists; the default path keeps block lists and no tile-level records.
inline StateLayout make_state_layout(int B, int V, int N, int H, int W, uint64_t cap, uint64_t capB, bool simple) {
    const uint64_t R = uint64_t(B) * V, T = uint64_t(tiles_x(W)) * tiles_y(H), P = uint64_t(H) * W;
    const uint64_t rec_cap = simple ? cap : 0, blk_cap = simple ? 0 : capB;
    StateLayout L;
    uint64_t o = 0;
    L.header = o;      o = align_up(o + sizeof(StateHeader));
    L.tile_off = o;    o = align_up(o + R * T * 4);
    L.tile_cnt = o;    o = align_up(o + R * T * 4);
    L.tile_time = o;   o = align_up(o + R * T * 8);     // (start ns, duration ns) of the forward blend per tile
    L.n_contrib = o;   o = align_up(o + R * P * 4);
    L.clamp_mask = o;  o = align_up(o + R * P);         // bit c: channel c of the pixel was clamped (SGR_FLAG_CLAMP_COLOR)
    L.sorted_ids = o;  o = align_up(o + cap * 4);
    L.g0 = o;          o = align_up(o + R * N * 16);    // per-(render, Gaussian) records written by the preprocess kernel:
    L.g1 = o;          o = align_up(o + R * N * 16);    // (x, y, extent, power threshold), (-A/2, -B, -C/2, opacity),
    L.g2 = o;          o = align_up(o + R * N * 16);    // (r, g, b, depth)
    L.rec0 = o;        o = align_up(o + rec_cap * 16);
    L.rec1 = o;        o = align_up(o + rec_cap * 16);
    L.rec2 = o;        o = align_up(o + rec_cap * 16);
    L.blk_off = o;     o = align_up(o + R * T * kBlocksPerTile * 4);
    L.blk_cnt = o;     o = align_up(o + R * T * kBlocksPerTile * 4);
    L.blk_eff = o;     o = align_up(o + R * T * kBlocksPerTile * 4);
    L.bidx = o;        o = align_up(o + blk_cap * 8);
    L.ck0 = o;         o = align_up(o + ckpt_slots(blk_cap) * kCkptPerSlot * 16);   // (T, C0, C1, C2) per pixel and slot
    L.ck1 = o;         o = align_up(o + ckpt_slots(blk_cap) * kCkptPerSlot * 4);    // D
    L.plan = o;        o = align_up(o + R * sizeof(ChunkPlan));                     // one per chunk (at most R chunks)
    // backward items: a chunk of n tiles and m block records emits at most 8 n + m / kSegB items per class
    L.bwd_items_stride = R * T * kBlocksPerTile + blk_cap / kSegB + R + 1;
    L.bwd_items = o;   o = align_up(o + L.bwd_items_stride * kBwdClasses * 8);
    L.total = o;
    return L;
}

constexpr int kAccumPlanes = 10;          // mean2D.xy, conic A/B/C, opacity, rgb, depth
// The gradient accumulators are one 48-byte row per (render, Gaussian): (mean2D.x, mean2D.y, conic A, conic B |
// conic C, opacity, r, g | b, depth, -, -), so that the blend backward adds a Gaussian's terms with three 16-byte vector
// atomics (red.global.add.v4.f32) instead of ten scalar ones into ten planes.
constexpr int kAccumStride = 12;

inline ScratchLayout make_scratch_layout(int B, int V, int N, int H, int W, uint64_t cap, int rpc) {
    const uint64_t R = uint64_t(B) * V, T = uint64_t(tiles_x(W)) * tiles_y(H);
    const uint64_t Rc = (uint64_t(rpc) < R) ? uint64_t(rpc) : R;
    ScratchLayout L;
    uint64_t o = 0;
    L.keys = o;        o = align_up(o + cap * 8);
    L.keys_tmp = o;    o = align_up(o + cap * 8);
    L.rect = o;        o = align_up(o + Rc * N * 8);
    L.cursor = o;      o = align_up(o + Rc * T * 4);
    L.work_blend = o;  o = align_up(o + Rc * T * 4);
    L.work_empty = o;  o = align_up(o + Rc * T * 4);
    L.work_counts = o; o = align_up(o + 256);
    L.dense_items = o; o = align_up(o + Rc * T * kBlocksPerTile * 4);              // block items of the dense lists
    L.loss_part = o;   o = align_up(o + Rc * T * 8 * 4);                           // fused loss: one partial per work item
    L.accum = o;       o = align_up(o + Rc * N * 4 * kAccumStride);
    L.total = o;
    return L;
}

// Per-chunk work-list counters (device).
struct WorkCounts {
    unsigned int n_big;        // tiles with >= kSmallSortCap instances: the first n_big entries of the blend list
    unsigned int sort_cursor;  // dynamic queue head of the small-tile sort
    unsigned int chunk_instances;
    unsigned int chunk_dropped;   // != 0: chunk did not fit into max_instances
    unsigned int n_blend;      // non-empty tiles, longest lists first (blend work list)
    unsigned int n_empty;      // tiles without instances (background only)
    unsigned int blend_cursor; // dynamic work queue heads of the persistent blend kernels
    unsigned int empty_cursor;
    unsigned int n_dense;      // dense block lists (queued by the tile sorts)
    unsigned int dense_cursor;
};

constexpr int kSmallSortCap = 4096;       // lists shorter than this are sorted in a 36 KB shared-memory CTA (power of two)
constexpr int kSmallSortThreads = 256;
constexpr int kSmallSortBuckets = 1024;
constexpr int kBigSortThreads = 1024;
constexpr int kBigSortBuckets = 4096;
constexpr int kBigSortSmemCap = 16 * 1024;  // instances whose keys fit in shared memory next to the histogram (<= 4 * kBigSortBuckets:
                                            // the dead histogram later holds one byte per key of block counts)
constexpr size_t kSmallSortSmem = size_t(kSmallSortCap) * 8 + size_t(kSmallSortBuckets) * 4;    // keys, histogram

// ------------------------------------------------------------------------------------------------
// Per-render constants handed to kernels by value.
struct RenderGeom {
    int N, H, W, tiles_x, tiles_y, num_tiles, V;
    float tanfovx, tanfovy;
};

// ------------------------------------------------------------------------------------------------
// exp_spec: bit-identical to oracle/sgr_oracle.cpp::exp_spec (fixed sequence of IEEE fp32 ops).
// exp_core is the same sequence without the range / NaN checks, valid for x in [-87, 88] (and NaN -> NaN):
// rint() and the float->int conversion are done with the 1.5 * 2^23 magic constant, which yields the identical
// round-half-even integer n for |x * log2(e)| < 2^22 and leaves n in the low mantissa bits of `tn`.
__device__ __forceinline__ float exp_core(float x) {
    const float t = __fmul_rn(x, 1.44269502162933349609375f);
    const float tn = __fadd_rn(t, 12582912.0f);
    const float n = __fsub_rn(tn, 12582912.0f);
    float r = __fmaf_rn(n, -0.693145751953125f, x);
    r = __fmaf_rn(n, -1.428606765330187045e-06f, r);
    float p = 1.98412701e-4f;
    p = __fmaf_rn(p, r, 1.38888892e-3f);
    p = __fmaf_rn(p, r, 8.33333377e-3f);
    p = __fmaf_rn(p, r, 4.16666679e-2f);
    p = __fmaf_rn(p, r, 1.66666672e-1f);
    p = __fmaf_rn(p, r, 0.5f);
    p = __fmaf_rn(p, r, 1.0f);
    p = __fmaf_rn(p, r, 1.0f);
    return __int_as_float(__float_as_int(p) + (__float_as_int(tn) << 23));
}

// exp_fast: exp(x) = 2^(x * log2(e)) on the SFU (MUFU.EX2, ~2 ulp) — what upstream's own `exp(power)` compiles to.
// One FMUL + one MUFU instead of exp_core's 12 fma-pipe instructions; not bit-reproducible on a CPU, so the blend
// kernels built on it are compared with the oracle to a tolerance (kExactExp = false, the default) while the
// exp_core instantiation (SGR_FLAG_EXACT_EXP) stays bit-exact.
__device__ __forceinline__ float exp_fast(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(__fmul_rn(x, 1.44269502162933349609375f)));
    return y;
}

__device__ __forceinline__ float exp_spec(float x) {
    if (x != x) return x;
    if (x < -87.0f) return 0.0f;
    if (x > 88.0f) return __int_as_float(0x7f800000);
    return exp_core(x);
}

// power = -1/2 (A dx^2 + C dy^2) - B dx dy in the spec's fixed FMA pattern (oracle: gauss_power), on the pre-scaled
// conic (hA, nB, hC) = (-A/2, -B, -C/2) that the preprocess kernel stores (exact scalings).
__device__ __forceinline__ float gauss_power(float hA, float nB, float hC, float dx, float dy) {
    return __fmaf_rn(dx, __fmul_rn(hA, dx), __fmul_rn(dy, __fmaf_rn(hC, dy, __fmul_rn(nB, dx))));
}

// Conservative alpha >= 1/255 extent (|dx| <= ex, |dy| <= ey) packed as two halves by the preprocess kernel.
__device__ __forceinline__ float2 unpack_extent(float packed) {
    const unsigned int u = __float_as_uint(packed);
    return make_float2(__half2float(__ushort_as_half(static_cast<unsigned short>(u & 0xffffu))),
                       __half2float(__ushort_as_half(static_cast<unsigned short>(u >> 16))));
}

// Quarter mask of one (Gaussian, tile) instance: bit (blk * 4 + q) is set when the Gaussian's conservative extent
// overlaps the pixel centres of 4x2 quarter q of 8x4 pixel block blk of the 16x16 tile at (X0, Y0).
// blk = (y block 0..3) * 2 + (x block 0..1), q = (x half) | (y half) << 1 — the blend kernels' work decomposition.
// NaN coordinates give an empty mask (all comparisons false): such a record can never pass the alpha test.
__device__ __forceinline__ unsigned int quarter_mask(float x, float y, float packed_ext, float X0, float Y0) {
    // Pixel centres are integers, so "xh >= first centre of the column" is floor(xh) >= it, and "xl <= last centre" is
    // ceil(xl) <= it: the overlapped columns / rows are integer ranges (exact; the clamps only keep the conversions
    // in range and map NaN to an empty range), turned into the mask with shifts and one carry-free multiply.
    const float2 ext = unpack_extent(packed_ext);
    const int X0i = __float2int_rz(X0), Y0i = __float2int_rz(Y0);
    const int cxl = __float2int_ru(fminf(fmaxf(x - ext.x, -1.0e6f), 1.0e6f)) - X0i;
    const int fxh = __float2int_rd(fminf(fmaxf(x + ext.x, -1.0e6f), 1.0e6f)) - X0i;
    const int cyl = __float2int_ru(fminf(fmaxf(y - ext.y, -1.0e6f), 1.0e6f)) - Y0i;
    const int fyh = __float2int_rd(fminf(fmaxf(y + ext.y, -1.0e6f), 1.0e6f)) - Y0i;
    const int clo = max(0, cxl >> 2), chi = min(3, fxh >> 2);      // 4-pixel columns qc: clo <= qc <= chi
    const int rlo = max(0, cyl >> 1), rhi = min(7, fyh >> 1);      // 2-pixel rows qr
    if (clo > chi || rlo > rhi) return 0u;
    const unsigned int col4 = (2u << chi) - (1u << clo);           // bits clo..chi
    const unsigned int row = (col4 & 3u) | ((col4 & 0xcu) << 2);   // column qc -> bit (qc & 1) | (qc >> 1) << 2
    const unsigned int r8 = (2u << rhi) - (1u << rlo);             // bits rlo..rhi
    const unsigned int t2 = (r8 & 0x03u) | ((r8 & 0x0cu) << 6) | ((r8 & 0x30u) << 12) | ((r8 & 0xc0u) << 18);
    const unsigned int m = (t2 & 0x01010101u) | ((t2 & 0x02020202u) << 1);   // row qr -> bit (qr >> 1) * 8 + (qr & 1) * 2
    return row * m;                                                // disjoint bit groups: the product is the OR of the shifts
}

constexpr float kAlphaMin = 1.0f / 255.0f;
constexpr float kAlphaMax = 0.99f;
constexpr float kTMin = 0.0001f;

// streaming 128-bit loads/stores
__device__ __forceinline__ float4 ldg_f4(const float4* p) { return __ldg(p); }

// Per-device launch caches (function attributes and occupancy are per device; one process may drive several GPUs).
constexpr int kMaxDevices = 64;
inline int current_device_slot() {
    int dev = 0;
    cudaGetDevice(&dev);
    return (dev >= 0 && dev < kMaxDevices) ? dev : 0;
}

// host-side launch wrappers implemented in the .cu files -------------------------------------------
struct ChunkCtx {
    RenderGeom g;
    int render_base;          // first render of the chunk
    int num_renders;          // renders in the chunk
    const SgrProblem* p;
    // state
    StateHeader* header;
    unsigned int* tile_off;   // [R*T]
    unsigned int* tile_cnt;   // [R*T]
    uint2* tile_time;         // [R*T]
    unsigned int* n_contrib;  // [R*P]
    unsigned int* sorted_ids; // [cap]
    float4 *g0, *g1, *g2;     // [Rc*N] this chunk's slice of the per-(render, Gaussian) records
    float4 *rec0, *rec1, *rec2;   // [cap] tile-level copies in depth order (SGR_FLAG_SIMPLE_BLEND only)
    unsigned char* clamp_mask; // [R*P]
    unsigned int *blk_off, *blk_cnt, *blk_eff;   // [R*T*8] block lists: start, records, records the backward replays
    uint2* bidx;                                 // [capB] block-list entries (Gaussian id, tile-list position << 4 | quarter mask)
    unsigned long long blk_capacity;
    float4* ck0;              // [ckpt_slots][32] forward checkpoints (T, C0, C1, C2)
    float* ck1;               // [ckpt_slots][32] forward checkpoints D
    ChunkPlan* plan;          // this chunk's backward plan
    uint2* bwd_items;         // [kBwdClasses][bwd_items_stride] backward (block, segment) items of all chunks
    unsigned long long bwd_items_stride;
    int chunk_index;
    // scratch
    unsigned long long* keys; // [cap]
    unsigned long long* keys_tmp; // [cap] bucket-ordered keys of tile lists that exceed the sort's shared memory
    uint2* rect;              // [Rc*N] packed tile rectangle
    unsigned int* cursor;     // [Rc*T]
    unsigned int *work_blend, *work_empty;
    WorkCounts* work_counts;
    unsigned int* dense_items; // [Rc*T*8] chunk-local tile * 8 + block of the dense block lists
    float* loss_part;         // [Rc*T*8] fused-loss partial sums, one per (render, tile, pixel block)
    float* accum;             // [Rc*N][kAccumStride]
    // optional fused loss (forward)
    const float *loss_target, *loss_mask;
    float* loss_dL_dcolor;
    float loss_scale;
    cudaStream_t stream;
};

cudaError_t launch_preprocess(const ChunkCtx& c, int32_t* radii);
cudaError_t launch_plan(const ChunkCtx& c);   // tile offsets, sort / blend / backward-segment work lists, status header
cudaError_t launch_scatter(const ChunkCtx& c);
cudaError_t launch_sort_tiles(const ChunkCtx& c);
cudaError_t launch_blend_forward(const ChunkCtx& c, float* out_color, float* out_depth, float* out_alpha,
                                 float* out_feed);
cudaError_t launch_loss_reduce(const ChunkCtx& c, float* loss_out);   // fused loss: sum of the chunk's partials
cudaError_t launch_blend_backward(const ChunkCtx& c, const SgrBackwardArgs& b);
cudaError_t launch_preprocess_backward(const ChunkCtx& c, const SgrBackwardArgs& a);
cudaError_t launch_mark_visible(const float* means3D, int N, const float* view, uint8_t* visible, cudaStream_t s);
cudaError_t launch_cov3d_from_scale_rot(const float* scales, const float* rots, float mod, int N, float* cov6,
                                        cudaStream_t s);
cudaError_t launch_cov3d_from_scale_rot_backward(const float* scales, const float* rots, float mod, int N,
                                                 const float* dcov, float* dscales, float* drots, cudaStream_t s);

cudaError_t launch_prep_cov3d(const float* s_raw, const float* rot, const float* dist2, long long n, bool bf16,
                              float* cov6, cudaStream_t s);
cudaError_t launch_prep_cov3d_backward(const float* s_raw, const float* rot, const float* dist2, long long n, bool bf16,
                                       const float* dcov, float* ds_raw, float* drot, cudaStream_t s);
cudaError_t launch_sh_colors(const float* means, const float* shs, const float* campos, int N, int deg, int max_coeffs,
                             float* colors, uint8_t* clamped, cudaStream_t s);
cudaError_t launch_sh_colors_backward(const float* means, const float* shs, const float* campos, int N, int deg,
                                      int max_coeffs, const uint8_t* clamped, const float* dcolors, float* dshs,
                                      float* dmeans, cudaStream_t s);

}  // namespace sgr
