// sgr_common.cuh — shared definitions of the sm_100a rasteriser kernels (internal; the public ABI is include/sgr.h).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/sgr.h"

namespace sgr {

constexpr int kTile = 16;                 // BLOCK_X = BLOCK_Y of the published algorithm
constexpr int kTilePixels = kTile * kTile;
constexpr int kDefaultRendersPerChunk = 16;

// ------------------------------------------------------------------------------------------------
// Device-resident status / counters at the start of `state`.
// The first sizeof(SgrStatus) bytes mirror SgrStatus (include/sgr.h) so a caller may also fetch the status with its
// own asynchronous copy of the head of `state`.
struct alignas(256) StateHeader {
    unsigned long long inst_required;      // total instances all renders need (even when overflowing)
    unsigned long long capacity;
    unsigned int overflow;
    unsigned int max_tile_instances;
    unsigned int nonempty_tiles;
    unsigned int pad;
    unsigned long long inst_cursor;        // instances consumed so far (global offset of the next chunk)
};
static_assert(sizeof(SgrStatus) == 32, "SgrStatus layout is part of the ABI");

// Layout of `state` (kept forward -> backward) for a problem shape.
struct StateLayout {
    uint64_t header, tile_off, tile_cnt, tile_time, n_contrib, sorted_ids, rec0, rec1, rec2, total;
};
// Layout of `scratch` (valid only inside one call).
struct ScratchLayout {
    uint64_t keys, g0, g1, g2, rect, cursor, work_small, work_big, work_blend, work_empty, work_counts, accum, total;
};

__host__ __device__ inline uint64_t align_up(uint64_t x, uint64_t a = 256) { return (x + a - 1) / a * a; }

inline int tiles_x(int W) { return (W + kTile - 1) / kTile; }
inline int tiles_y(int H) { return (H + kTile - 1) / kTile; }

inline StateLayout make_state_layout(int B, int V, int N, int H, int W, uint64_t cap) {
    (void)N;
    const uint64_t R = uint64_t(B) * V, T = uint64_t(tiles_x(W)) * tiles_y(H), P = uint64_t(H) * W;
    StateLayout L;
    uint64_t o = 0;
    L.header = o;      o = align_up(o + sizeof(StateHeader));
    L.tile_off = o;    o = align_up(o + R * T * 4);
    L.tile_cnt = o;    o = align_up(o + R * T * 4);
    L.tile_time = o;   o = align_up(o + R * T * 8);     // (start ns, duration ns) of the forward blend per tile
    L.n_contrib = o;   o = align_up(o + R * P * 4);
    L.sorted_ids = o;  o = align_up(o + cap * 4);
    L.rec0 = o;        o = align_up(o + cap * 16);
    L.rec1 = o;        o = align_up(o + cap * 16);
    L.rec2 = o;        o = align_up(o + cap * 16);
    L.total = o;
    return L;
}

constexpr int kAccumPlanes = 10;          // mean2D.xy, conic A/B/C, opacity, rgb, depth

inline ScratchLayout make_scratch_layout(int B, int V, int N, int H, int W, uint64_t cap, int rpc) {
    const uint64_t R = uint64_t(B) * V, T = uint64_t(tiles_x(W)) * tiles_y(H);
    const uint64_t Rc = (uint64_t(rpc) < R) ? uint64_t(rpc) : R;
    ScratchLayout L;
    uint64_t o = 0;
    L.keys = o;        o = align_up(o + cap * 8);
    L.g0 = o;          o = align_up(o + Rc * N * 16);
    L.g1 = o;          o = align_up(o + Rc * N * 16);
    L.g2 = o;          o = align_up(o + Rc * N * 16);
    L.rect = o;        o = align_up(o + Rc * N * 8);
    L.cursor = o;      o = align_up(o + Rc * T * 4);
    L.work_small = o;  o = align_up(o + Rc * T * 4);
    L.work_big = o;    o = align_up(o + Rc * T * 4);
    L.work_blend = o;  o = align_up(o + Rc * T * 4);
    L.work_empty = o;  o = align_up(o + Rc * T * 4);
    L.work_counts = o; o = align_up(o + 256);
    L.accum = o;       o = align_up(o + Rc * N * 4 * kAccumPlanes);
    L.total = o;
    return L;
}

// Per-chunk work-list counters (device).
struct WorkCounts {
    unsigned int n_small;      // tiles with 1..kSmallSortCap instances
    unsigned int n_big;        // tiles with more
    unsigned int chunk_instances;
    unsigned int chunk_dropped;   // != 0: chunk did not fit into max_instances
    unsigned int n_blend;      // non-empty tiles, longest lists first (blend work list)
    unsigned int n_empty;      // tiles without instances (background only)
    unsigned int blend_cursor; // dynamic work queue heads of the persistent blend kernels
    unsigned int empty_cursor;
};

constexpr int kSmallSortCap = 4096;       // instances sorted in a 40 KB shared-memory CTA
constexpr int kSmallSortThreads = 256;
constexpr int kSmallSortBuckets = 1024;
constexpr int kBigSortThreads = 1024;
constexpr int kBigSortBuckets = 4096;
constexpr int kBigSortSmemCap = 26 * 1024;  // instances whose keys fit in shared memory next to the histogram

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

constexpr float kAlphaMin = 1.0f / 255.0f;
constexpr float kAlphaMax = 0.99f;
constexpr float kTMin = 0.0001f;

// streaming 128-bit loads/stores
__device__ __forceinline__ float4 ldg_f4(const float4* p) { return __ldg(p); }

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
    float4 *rec0, *rec1, *rec2;   // [cap]
    // scratch
    unsigned long long* keys; // [cap]
    float4 *g0, *g1, *g2;     // [Rc*N]
    uint2* rect;              // [Rc*N] packed tile rectangle
    unsigned int* cursor;     // [Rc*T]
    unsigned int *work_small, *work_big, *work_blend, *work_empty;
    WorkCounts* work_counts;
    float* accum;             // [kAccumPlanes][Rc*N]
    cudaStream_t stream;
};

cudaError_t launch_preprocess(const ChunkCtx& c, int32_t* radii);
cudaError_t launch_scan_tiles(const ChunkCtx& c);
cudaError_t launch_worklist(const ChunkCtx& c);     // blend/empty work lists from tile_cnt (forward and backward)
cudaError_t launch_scatter(const ChunkCtx& c);
cudaError_t launch_sort_tiles(const ChunkCtx& c);
cudaError_t launch_blend_forward(const ChunkCtx& c, float* out_color, float* out_depth, float* out_alpha);
cudaError_t launch_blend_backward(const ChunkCtx& c, const float* out_alpha, const float* dL_dcolor,
                                  const float* dL_ddepth, const float* dL_dalpha);
cudaError_t launch_preprocess_backward(const ChunkCtx& c, const SgrBackwardArgs& a);
cudaError_t launch_mark_visible(const float* means3D, int N, const float* view, uint8_t* visible, cudaStream_t s);
cudaError_t launch_cov3d_from_scale_rot(const float* scales, const float* rots, float mod, int N, float* cov6,
                                        cudaStream_t s);
cudaError_t launch_cov3d_from_scale_rot_backward(const float* scales, const float* rots, float mod, int N,
                                                 const float* dcov, float* dscales, float* drots, cudaStream_t s);

}  // namespace sgr
