// sgr_blend.cu — the alpha-compositing kernels (forward and backward) of the B200 rasteriser.
//
// Replaces upstream renderCUDA forward/backward (third-party diff_gaussian_rasterization; call site
// /root/reference/core/gaussians/gs.py:99-106).  Design (DESIGN.md "blend"):
//   * the unit of work is one 4x2 pixel QUARTER of one 8x4 pixel block of one tile of one render.  Every WARP is
//     autonomous: it pops work items from a device-side queue ordered longest-list-first
//     (sgr_binning.cu::plan_kernel), streams the BLOCK's depth-ordered 48-byte records (three float4 streams; the
//     per-tile sort emits one contiguous list per block that holds only the instances whose extent touches the
//     block) through its own shared-memory ring with 1-D TMA bulk copies (cp.async.bulk) completing on its own
//     mbarriers, and never waits for another warp — no block-wide barrier, no producer/consumer hand-off, early exit
//     as soon as its 8 pixels are finished;
//   * lane = (pixel, slot): 8 pixels x 4 slots.  A batch of 128 records is culled against the quarter (lane =
//     record: the per-record quarter bit built by the tile sort, one ballot per round, compaction of the survivors'
//     batch-local indices into a sentinel-padded list); a "super-trip" then evaluates FOUR consecutive survivors for
//     the 8 pixels at once.  The sequential part of compositing — the transmittance T — crosses the slots with a
//     two-step shuffle prefix over (1 - alpha) that does not depend on T, so the T-dependent chain is one multiply
//     per four Gaussians and dense lists are no longer a single warp's dependent-issue chain; colour / depth /
//     weight are accumulated per (pixel, slot) lane and combined once per item.  Culled records would have been
//     skipped by the alpha test, so every survivor is evaluated with the straightforward kernel's arithmetic;
//   * exp(power) is a template switch: kExact = the oracle's exp_spec sequence and the oracle's sequential
//     test_T = fma(-alpha, T, T) chain evaluated redundantly by the four slot lanes of a pixel (SGR_FLAG_EXACT_EXP:
//     T, the alpha / T thresholds and therefore n_contrib are bit-identical to oracle/sgr_oracle.cpp; colour, depth
//     and alpha differ from it only by the order of the four partial sums, ~1e-7); default = the SFU's ex2 like
//     upstream's own exp() and the prefix product (~1e-6 from the oracle);
//   * per batch the forward clears the quarter bits of the records that did not blend in the quarter (the backward
//     culls on the exact set), checkpoints the running state every kSegB records and pushes one backward work item
//     per (quarter, segment) that blended anything, classed by the number of records that blended;
//   * backward: work items are (block, quarter, kSegB-record segment), resumed from the forward's checkpoints; the
//     same lanes walk the survivors back to front, four per super-trip.  With u_j = w_j (c_j . dL/dC + z_j dL/dD) the
//     per-pixel recurrences are scalar: T_k = T_(k+1) / (1 - alpha_k) (prefix product across the slots) and
//     U_k = sum_(j behind k) u_j (prefix sum), dL/dalpha_k = T_k (c_k . g) - (U_k + T_final (bg . g - dL/dA)) / (1 -
//     alpha_k); the nine gradient terms of a (Gaussian, quarter) pair are reduced over its 8 pixel lanes with a
//     reduce-scatter butterfly and leave as ONE atomic per component (no shared-memory stash, no role flip).
//     Gradient arithmetic is free to use FMA (tolerance, not bit-exact);
//   * optional epilogues: clamp (+ mask for the backward), masked L1 loss + dL/dcolour (SgrForwardArgs::loss_*), the
//     factor-2 bilinear LPIPS feed.
//
// Compiled with --fmad=false: the per-pixel expressions are the oracle's (oracle/sgr_oracle.cpp::blend_forward /
// blend_backward) evaluated in the same order.
#include <cstdio>
#include <cstdlib>
#include <type_traits>

#include "sgr_common.cuh"

namespace sgr {
namespace {

#ifndef SGR_FWD_BATCH
#define SGR_FWD_BATCH 128
#endif
#ifndef SGR_FWD_STAGES
#define SGR_FWD_STAGES 2
#endif
#ifndef SGR_BWD_BATCH
#define SGR_BWD_BATCH 128
#endif
#ifndef SGR_BWD_STAGES
#define SGR_BWD_STAGES 2
#endif
constexpr int kFwdBatch = SGR_FWD_BATCH;       // records per ring stage (culled 32 at a time: lane = record)
constexpr int kBwdBatch = SGR_BWD_BATCH;
constexpr int kFwdStages = SGR_FWD_STAGES;     // per-warp TMA ring depth
constexpr int kBwdStages = SGR_BWD_STAGES;
static_assert(kSegB % kFwdBatch == 0 && kSegB % kBwdBatch == 0, "segments are whole batches");
constexpr int kWarpsPerCta = 8;
constexpr int kBlendThreads = kWarpsPerCta * 32;
constexpr int kBlockW = 8, kBlockH = 4;
constexpr int kSlots = 4;                      // consecutive survivors evaluated per super-trip (lane = pixel + 8 * slot)
#ifndef SGR_FWD_GROUPS
#define SGR_FWD_GROUPS 2
#endif
#ifndef SGR_BWD_GROUPS
#define SGR_BWD_GROUPS 1
#endif
constexpr int kFwdGroups = SGR_FWD_GROUPS;     // super-trips written load-first per loop iteration (ILP)
constexpr int kBwdGroups = SGR_BWD_GROUPS;
constexpr unsigned int kFull = 0xffffffffu;
template <bool kExact>
__device__ __forceinline__ float blend_exp(float x) { return kExact ? exp_core(x) : exp_fast(x); }

// ------------------------------------------------------------------------------------------------ PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {}
}
// 1-D TMA: global -> shared bulk copy, completion signalled on `bar` (bytes multiple of 16, 16-byte aligned).
__device__ __forceinline__ void tma_load_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ float rcp_approx(float x) {       // MUFU.RCP, 1 ulp; gradients are compared with a tolerance
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ unsigned long long global_timer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

constexpr int kListPad = 16;                   // the group loops read up to 4 * groups - 1 entries past the count

template <int kNumStages, int kBatch>
struct WarpSmem {
    static constexpr int stages = kNumStages;
    float4 r0[kNumStages][kBatch + 1];          // slot kBatch of every stage = the sentinel record (never blends)
    float4 r1[kNumStages][kBatch + 1];
    float4 r2[kNumStages][kBatch + 1];
    // batch-local indices of the quarter's survivors, sentinel-filled
    unsigned char list[kBatch + kListPad];
    unsigned char hit[kBatch + kListPad];       // forward: hit[j] != 0 <=> some pixel of the quarter blended record j
    uint64_t full[kNumStages];
};

// Cull a batch of m <= kBatch block records against the warp's 4x2 quarter `q` of the block: 32 records per round
// (lane = record) read bit q of their quarter mask (low bits of rec0.z, built by the tile sort from
// sgr_common.cuh::quarter_mask and refined by the forward), one ballot per round, warp-parallel compaction of the
// survivors' batch-local indices into list[...] (ascending; descending with kReverse — the backward walks back to
// front); unused list entries point at the sentinel record.  bits[r] != 0: record 32 r + lane is a candidate.
template <bool kReverse, int kBatch>
__device__ __forceinline__ unsigned int cull_batch(const float4* r0, unsigned int m, int q, unsigned char* list, int lane,
                                                   unsigned int (&bits)[kBatch / 32]) {
    constexpr int R = kBatch / 32;
    const unsigned int lt = kReverse ? ~((2u << lane) - 1u) : (1u << lane) - 1u;   // lanes before / after this one
    const unsigned int* words = reinterpret_cast<const unsigned int*>(r0);
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const unsigned int e = 32u * r + lane;
        bits[r] = (e < m) ? ((words[4 * e + 2] >> q) & 1u) : 0u;
    }
    {                             // every list entry the compaction does not overwrite points at the sentinel record
        constexpr unsigned int fill = kBatch * 0x01010101u;
        uint4* lw = reinterpret_cast<uint4*>(list);
        constexpr int kVecs = (kBatch + kListPad) / 16;
        if (lane < kVecs) lw[lane] = make_uint4(fill, fill, fill, fill);
        static_assert(kVecs <= 32, "one store per lane fills the list");
    }
    unsigned int mq[R];
#pragma unroll
    for (int r = 0; r < R; ++r) mq[r] = __ballot_sync(kFull, bits[r]);
    __syncwarp();                 // the sentinel fill is ordered before the compaction stores
    unsigned int n = 0;
#pragma unroll
    for (int k = 0; k < R; ++k) {
        const int r = kReverse ? R - 1 - k : k;                 // descending lists walk the rounds from the back
        if (bits[r]) list[n + __popc(mq[r] & lt)] = static_cast<unsigned char>(32u * r + lane);
        n += __popc(mq[r]);
    }
    __syncwarp();
    return n;
}

// Per-warp TMA ring: `issued` / `consumed` count batches over the whole kernel (stage = k % stages,
// parity = (k / stages) & 1), so the mbarriers never need re-initialisation between work items.
template <typename Smem>
__device__ __forceinline__ void ring_issue(Smem& sm, unsigned int issued, const float4* g0, const float4* g1,
                                           const float4* g2, unsigned int m, int lane) {
    if (lane == 0) {
        const int s = issued % Smem::stages;
        const uint32_t bytes = m * 16u;
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // earlier generic reads of the slot are done
        mbar_arrive_expect_tx(&sm.full[s], 3u * bytes);
        tma_load_1d(sm.r0[s], g0, bytes, &sm.full[s]);
        tma_load_1d(sm.r1[s], g1, bytes, &sm.full[s]);
        tma_load_1d(sm.r2[s], g2, bytes, &sm.full[s]);
    }
}

// Pops one work item for the calling warp, or 0xffffffff when the queue is empty.
__device__ __forceinline__ unsigned int pop_item(unsigned int* cursor, unsigned int n_items, int lane) {
    unsigned int w = 0;
    if (lane == 0) w = atomicAdd(cursor, 1u);
    w = __shfl_sync(kFull, w, 0);
    return w < n_items ? w : 0xffffffffu;
}

// Sum / product / max of a per-(pixel, slot) value over the four slot lanes of a pixel (lane = pixel + 8 * slot).
__device__ __forceinline__ float slots_sum(float v) {
    v += __shfl_xor_sync(kFull, v, 8);
    return v + __shfl_xor_sync(kFull, v, 16);
}
__device__ __forceinline__ float slots_prod(float v) {
    v *= __shfl_xor_sync(kFull, v, 8);
    return v * __shfl_xor_sync(kFull, v, 16);
}
__device__ __forceinline__ unsigned int slots_max(unsigned int v) {
    v = max(v, __shfl_xor_sync(kFull, v, 8));
    return max(v, __shfl_xor_sync(kFull, v, 16));
}

// ------------------------------------------------------------------------------------------------ forward
struct FwdArgs {
    RenderGeom g;
    int render_base;
    const unsigned int* blk_off;  // [R*T*8] block lists (sgr_binning.cu step 5)
    const unsigned int* blk_cnt;
    unsigned int* q_eff;          // [R*T*32] out: records of the block list the backward replays for the quarter
    const float4 *rec0, *rec1, *rec2;   // block records
    unsigned int* rec0_words;     // rec0 as words: the forward refines the quarter masks (word 2 of every record)
    int refine_masks;
    const float* bg;
    unsigned int* n_contrib;
    unsigned char* clamp_mask;    // [R*P] written when clamp_color
    uint2* tile_time;
    float *out_color, *out_depth, *out_alpha;
    float* out_feed;              // optional LPIPS feed [R,3,H/2,W/2]
    float4* ck0;                  // per-pixel checkpoints (T, C0, C1, C2) at segment boundaries of long lists
    float* ck1;                   // ... and D
    ChunkPlan* plan;              // backward items are pushed here (refine_masks only)
    uint2* bwd_items;
    unsigned long long bwd_items_stride;
    const unsigned int *work_blend, *work_empty;
    WorkCounts* wc;
    int clamp_color;
    // optional fused loss (include/sgr.h SgrForwardArgs::loss_*)
    const float *loss_target, *loss_mask;
    float* loss_dL_dcolor;
    float* loss_part;
    float loss_scale;
};

// Backward item classes by the number of records of the (quarter, segment) that blended.
__device__ __forceinline__ int bwd_class(unsigned int hits) {
    return hits >= 3u * kSegB / 4 ? 0 : hits >= 3u * kSegB / 8 ? 1 : hits >= kSegB / 8 ? 2 : 3;
}

// Fused loss epilogue of one pixel: |clamp(c) * m - t * m| summed over the three channels; writes
// d loss / d (unclamped colour) = sign(diff) * m * scale where the clamp did not saturate (torch's clamp / abs rules).
__device__ __forceinline__ float loss_pixel(const FwdArgs& a, size_t rP, size_t P, size_t pix, float c0, float c1,
                                            float c2) {
    const float m = a.loss_mask ? a.loss_mask[rP + pix] : 1.0f;
    const float* tg = a.loss_target + 3 * rP;
    float* dl = a.loss_dL_dcolor + 3 * rP;
    const float gs = m * a.loss_scale;
    float part = 0.0f;
    const float cs[3] = {c0, c1, c2};
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
        const float c = cs[ch];
        const float cl = fminf(fmaxf(c, 0.0f), 1.0f);
        const float diff = cl * m - tg[ch * P + pix] * m;
        part += fabsf(diff);
        const bool pass = c >= 0.0f && c <= 1.0f;
        dl[ch * P + pix] = pass ? (diff > 0.0f ? gs : diff < 0.0f ? -gs : 0.0f) : 0.0f;
    }
    return part;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
    return v;
}

// Output epilogue of one pixel; all 32 lanes call it.  `writer`: this lane stores the pixel (one lane per pixel).
// kRowXor = lane distance of the pixel one row below (8: lane = 8x4 block pixel; 4: lane & 7 = 4x2 quarter pixel).
// clamp_color: gs.py:107 `clamp(0, 1)` fused, with the clamp mask kept for the backward (bit c = channel c saturated,
// torch's rule: the gradient passes where 0 <= c <= 1).  out_feed: the LPIPS input of whole_loss.py:132-136 — a
// factor-2 bilinear resize (align_corners=False) is the mean over 2x2 pixels, here two shuffles.
template <int kRowXor>
__device__ __forceinline__ void write_pixel(const FwdArgs& a, int r, size_t P, int px, int py, bool writer, float c0,
                                            float c1, float c2, float D, float Wt, unsigned int last) {
    const size_t pix = size_t(py) * a.g.W + px;
    if (a.clamp_color) {
        const unsigned int m = (c0 >= 0.0f && c0 <= 1.0f ? 0u : 1u) | (c1 >= 0.0f && c1 <= 1.0f ? 0u : 2u) |
                               (c2 >= 0.0f && c2 <= 1.0f ? 0u : 4u);
        c0 = fminf(fmaxf(c0, 0.0f), 1.0f); c1 = fminf(fmaxf(c1, 0.0f), 1.0f); c2 = fminf(fmaxf(c2, 0.0f), 1.0f);
        if (writer) a.clamp_mask[size_t(r) * P + pix] = static_cast<unsigned char>(m);
    }
    if (a.out_feed) {                           // even H, W: a 2x2 cell is entirely inside or outside the image
        float s0 = c0 + __shfl_xor_sync(kFull, c0, 1), s1 = c1 + __shfl_xor_sync(kFull, c1, 1),
              s2 = c2 + __shfl_xor_sync(kFull, c2, 1);
        s0 += __shfl_xor_sync(kFull, s0, kRowXor); s1 += __shfl_xor_sync(kFull, s1, kRowXor);
        s2 += __shfl_xor_sync(kFull, s2, kRowXor);
        if (writer && (px & 1) == 0 && (py & 1) == 0) {
            const size_t Pq = P >> 2;
            const size_t q = size_t(py >> 1) * (a.g.W >> 1) + (px >> 1);
            float* of = a.out_feed + size_t(r) * 3 * Pq;
            of[q] = 0.5f * s0 - 1.0f; of[Pq + q] = 0.5f * s1 - 1.0f; of[2 * Pq + q] = 0.5f * s2 - 1.0f;
        }
    }
    if (writer) {
        float* oc = a.out_color + size_t(r) * 3 * P;
        oc[pix] = c0; oc[P + pix] = c1; oc[2 * P + pix] = c2;
        a.out_depth[size_t(r) * P + pix] = D;
        a.out_alpha[size_t(r) * P + pix] = Wt;
        a.n_contrib[size_t(r) * P + pix] = last;
    }
}

using FwdSmem = WarpSmem<kFwdStages, kFwdBatch>;
using BwdSmem = WarpSmem<kBwdStages, kBwdBatch>;

#ifndef SGR_FWD_MIN_CTAS
#define SGR_FWD_MIN_CTAS 2
#endif
#ifndef SGR_BWD_MIN_CTAS
#define SGR_BWD_MIN_CTAS 2
#endif

template <bool kExact>
__global__ void __launch_bounds__(kBlendThreads, SGR_FWD_MIN_CTAS) blend_forward_kernel(FwdArgs a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    FwdSmem& sm = reinterpret_cast<FwdSmem*>(smem_raw)[warp];
    const size_t P = size_t(a.g.H) * a.g.W;
    const float bg0 = a.bg[0], bg1 = a.bg[1], bg2 = a.bg[2];
    const unsigned int n_items = a.wc->n_blend * kItemsPerTile, n_empty_items = a.wc->n_empty * kBlocksPerTile;
    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < kFwdStages; ++s) mbar_init(&sm.full[s], 1);
        fence_barrier_init();
    }
    __syncwarp();
    const int pl = lane & 7, slot = lane >> 3;   // pixel of the quarter (x = pl & 3, y = pl >> 2), survivor slot
    unsigned int issued = 0, consumed = 0;       // ring counters (warp-uniform)
    const bool refine = a.refine_masks != 0;
    for (int w = lane; w < kFwdBatch + kListPad; w += 32) sm.hit[w] = 0;
    if (lane < kFwdStages) {      // the sentinel record: opacity 0 at power 0 -> alpha 0; never valid in exact mode
        sm.r0[lane][kFwdBatch] = make_float4(0.0f, 0.0f, 0.0f, __int_as_float(0x7f800000));
        sm.r1[lane][kFwdBatch] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        sm.r2[lane][kFwdBatch] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    }
    __syncwarp();

    // ---------------- quarters of tiles with instances
    for (;;) {
        const unsigned int item = pop_item(&a.wc->blend_cursor, n_items, lane);
        if (item == 0xffffffffu) break;
        const unsigned int tile_local = a.work_blend[item / kItemsPerTile];
        const int blk = (item % kItemsPerTile) / kQuartersPerBlock;
        const int q = item % kQuartersPerBlock;
        const int rl = tile_local / a.g.num_tiles;
        const int tile = tile_local - rl * a.g.num_tiles;
        const int r = a.render_base + rl;
        const size_t tg = size_t(r) * a.g.num_tiles + tile;
        const size_t bi = tg * kBlocksPerTile + blk;
        const size_t item_slot = (size_t(tile_local) * kBlocksPerTile + blk) * kQuartersPerBlock + q;
        const unsigned int n = a.blk_cnt[bi];                       // records of this block's list
        const size_t off = a.blk_off[bi];
        const unsigned int nb = (n + kFwdBatch - 1) / kFwdBatch;
        const int tx = tile % a.g.tiles_x, ty = tile / a.g.tiles_x;
        const int qx0 = tx * kTile + (blk & 1) * kBlockW + (q & 1) * 4, qy0 = ty * kTile + (blk >> 1) * kBlockH + (q >> 1) * 2;
        if (qx0 >= a.g.W || qy0 >= a.g.H) {                        // quarter entirely outside the image
            if (lane == 0) {
                if (a.loss_target) a.loss_part[item_slot] = 0.0f;
                if (refine) a.q_eff[bi * kQuartersPerBlock + q] = 0u;
            }
            continue;
        }
        const unsigned long long t_begin = global_timer_ns();
        const int px = qx0 + (pl & 3), py = qy0 + (pl >> 2);
        const bool inside = px < a.g.W && py < a.g.H;
        const float pxf = float(px), pyf = float(py);
        const float4 *g0 = a.rec0 + off, *g1 = a.rec1 + off, *g2 = a.rec2 + off;

        // T and done are per pixel (identical in the pixel's four slot lanes); the sums are per (pixel, slot)
        float T = 1.0f, C0 = 0.0f, C1 = 0.0f, C2 = 0.0f, D = 0.0f, Wt = 0.0f;
        float fix = 1.0f;                        // default mode: (1 - alpha) of the slots that blended in the pixel's last super-trip
        unsigned int last = 0;                   // position in the tile list + 1 of the last record this lane blended
        bool done = !inside;
        unsigned int b_issued = 0;
        unsigned int seg_hits = 0, eff = 0;      // records of the current segment that blended; last such record + 1
        while (b_issued < nb && b_issued < unsigned(kFwdStages - 1)) {
            ring_issue(sm, issued, g0 + b_issued * kFwdBatch, g1 + b_issued * kFwdBatch, g2 + b_issued * kFwdBatch,
                       min(unsigned(kFwdBatch), n - b_issued * kFwdBatch), lane);
            ++issued; ++b_issued;
        }
        for (unsigned int b = 0; b < nb; ++b) {
            if (b_issued < nb) {                 // refill the slot consumed in the previous iteration
                ring_issue(sm, issued, g0 + b_issued * kFwdBatch, g1 + b_issued * kFwdBatch, g2 + b_issued * kFwdBatch,
                           min(unsigned(kFwdBatch), n - b_issued * kFwdBatch), lane);
                ++issued; ++b_issued;
            }
            const int s = consumed % kFwdStages;
            mbar_wait(&sm.full[s], (consumed / kFwdStages) & 1);
            ++consumed;
            const unsigned int m = min(unsigned(kFwdBatch), n - b * kFwdBatch);
            const unsigned int cbase = b * kFwdBatch;
            const float4* r0 = sm.r0[s];
            const float4* r1 = sm.r1[s];
            const float4* r2 = sm.r2[s];
            unsigned int qbits[kFwdBatch / 32];
            const unsigned int cnt = cull_batch<false, kFwdBatch>(r0, m, q, sm.list, lane, qbits);
            // One super-trip = the quarter's next four survivors (slot = which of the four) for its 8 pixels.  List
            // entries past the count point at the sentinel record (alpha = 0): no bounds checks in the loop.
            for (unsigned int t0 = 0; t0 < cnt; t0 += kSlots * kFwdGroups) {
                float4 q0[kFwdGroups], q1[kFwdGroups];
                unsigned int jj[kFwdGroups];
                float al[kFwdGroups];
#pragma unroll
                for (int g = 0; g < kFwdGroups; ++g) {
                    jj[g] = sm.list[t0 + kSlots * g + slot];
                    q0[g] = r0[jj[g]];
                    q1[g] = r1[jj[g]];
                }
#pragma unroll
                for (int g = 0; g < kFwdGroups; ++g) {
                    const float dx = q0[g].x - pxf, dy = q0[g].y - pyf;
                    const float power = gauss_power(q1[g].x, q1[g].y, q1[g].z, dx, dy);
                    // exact mode: exp_core needs power >= the record's threshold (>= -87); below it alpha < 1/255 anyway
                    const bool valid = kExact ? (!(power > 0.0f) && !(power < q0[g].w)) : !(power > 0.0f);
                    const float alpha = fminf(kAlphaMax, q1[g].w * blend_exp<kExact>(valid ? power : 0.0f));
                    al[g] = (valid && !(alpha < kAlphaMin)) ? alpha : 0.0f;
                }
#pragma unroll
                for (int g = 0; g < kFwdGroups; ++g) {
                    float Tb;                    // transmittance in front of this lane's record
                    bool blend;
                    if (kExact) {
                        // the oracle's sequential chain over the four slots, evaluated by every slot lane of the pixel
                        float Tk = T;
                        bool dn = done;
                        Tb = T; blend = false;
#pragma unroll
                        for (int k = 0; k < kSlots; ++k) {
                            const float ak = __shfl_sync(kFull, al[g], pl + 8 * k);
                            const bool ok = !dn && ak != 0.0f;
                            const float test_T = __fmaf_rn(-ak, Tk, Tk);
                            const bool stop = ok && (test_T < kTMin);
                            const bool bl = ok && !stop;
                            if (k == slot) { Tb = Tk; blend = bl; }
                            Tk = bl ? test_T : Tk;
                            dn = dn || stop;
                        }
                        T = Tk; done = dn;
                    } else {
                        // prefix product of (1 - alpha) across the slots (independent of T), then one multiply by T
                        const float om = 1.0f - al[g];
                        float inc = om;
                        float v = __shfl_up_sync(kFull, inc, 8);
                        inc = slot >= 1 ? inc * v : inc;
                        v = __shfl_up_sync(kFull, inc, 16);
                        inc = slot >= 2 ? inc * v : inc;
                        float exc = __shfl_up_sync(kFull, inc, 8);
                        exc = slot >= 1 ? exc : 1.0f;
                        const float tot = __shfl_sync(kFull, inc, pl + 24);
                        Tb = T * exc;
                        const bool pass = !(T * inc < kTMin);       // monotone over the slots: a stop holds for all behind
                        blend = !done && al[g] != 0.0f && pass;
                        const float T3 = T * tot;
                        const bool stop_now = !done && (T3 < kTMin);
                        // a pixel that stops inside this super-trip keeps T and remembers which slots still blended
                        fix = stop_now ? (blend ? om : 1.0f) : fix;
                        T = (done || stop_now) ? T : T3;
                        done = done || stop_now;
                    }
                    const float4 q2 = r2[jj[g]];
                    const float w = blend ? al[g] * Tb : 0.0f;      // adding c * 0 leaves the sums bit-unchanged
                    C0 = __fmaf_rn(q2.x, w, C0);
                    C1 = __fmaf_rn(q2.y, w, C1);
                    C2 = __fmaf_rn(q2.z, w, C2);
                    Wt += w;
                    D = __fmaf_rn(q2.w, w, D);
                    // n_contrib counts positions in the TILE's list (upstream's contributor index): bits 4.. of word 2
                    last = blend ? (__float_as_uint(q0[g].z) >> 4) + 1u : last;
                    // same-value stores of the record's lanes to one byte: benign
                    if (refine && blend) sm.hit[jj[g]] = 1;
                }
            }
            __syncwarp();
            if (refine) {
                // Mask refinement for the backward pass: clear the quarter's bit of the candidates that no pixel of
                // the quarter blended (alpha test, finished pixels) — the backward walk culls on the same word.  The
                // four quarters of a block share its records: atomic AND.
                if (cnt != 0u) {
#pragma unroll
                    for (int rr = 0; rr < kFwdBatch / 32; ++rr) {
                        const unsigned int e = 32u * rr + lane;
                        const unsigned int h = sm.hit[e];
                        if (qbits[rr] && !h) atomicAnd(a.rec0_words + 4 * (off + cbase + e) + 2, ~(1u << q));
                        if (h) sm.hit[e] = 0;
                        const unsigned int hb = __ballot_sync(kFull, h != 0u);
                        if (hb) { seg_hits += __popc(hb); eff = cbase + 32u * rr + (32u - __clz(hb)); }
                    }
                    __syncwarp();
                }
                // end of a backward segment (or of the list): one work item if anything blended in it
                const unsigned int end = cbase + m;
                if (end % kSegB == 0u || end == n) {
                    if (seg_hits != 0u && lane == 0) {
                        const int cls = bwd_class(seg_hits);
                        const unsigned int sl = atomicAdd(&a.plan->n_items[cls], 1u);
                        a.bwd_items[size_t(cls) * a.bwd_items_stride + a.plan->item_base + sl] =
                            make_uint2(unsigned(item_slot), (end - 1u) / kSegB);
                    }
                    seg_hits = 0u;
                }
            }
            // lists longer than one backward segment: checkpoint the running state at every segment boundary
            if (n > unsigned(kSegB) && ((b + 1) * kFwdBatch) % kSegB == 0 && (b + 1) * kFwdBatch < n) {
                const float c0 = slots_sum(C0), c1 = slots_sum(C1), c2 = slots_sum(C2), dd = slots_sum(D);
                const float Tc = kExact ? T : T * slots_prod(fix);
                if (slot == 0) {
                    const size_t ci = (off / (kSegB / 2) + (b + 1) * kFwdBatch / kSegB - 1) * 32 + q * 8 + pl;
                    a.ck0[ci] = make_float4(Tc, c0, c1, c2);
                    a.ck1[ci] = dd;
                }
            }
            if (__all_sync(kFull, done)) break;
        }
        // combine the four slot lanes of every pixel
        C0 = slots_sum(C0); C1 = slots_sum(C1); C2 = slots_sum(C2); D = slots_sum(D); Wt = slots_sum(Wt);
        last = slots_max(last);
        if (!kExact) T *= slots_prod(fix);
        if (n > unsigned(kSegB) && slot == 0) {       // final state, read by the backward's non-final segments
            const size_t ci = (off / (kSegB / 2) + (n + kSegB - 1) / kSegB - 1) * 32 + q * 8 + pl;
            a.ck0[ci] = make_float4(T, C0, C1, C2);
            a.ck1[ci] = D;
        }
        if (refine) {
            // an early exit (all pixels finished) inside a segment leaves its item unpushed: push it now
            if (seg_hits != 0u && lane == 0) {
                const int cls = bwd_class(seg_hits);
                const unsigned int sl = atomicAdd(&a.plan->n_items[cls], 1u);
                a.bwd_items[size_t(cls) * a.bwd_items_stride + a.plan->item_base + sl] =
                    make_uint2(unsigned(item_slot), (eff - 1u) / kSegB);
            }
            if (lane == 0) a.q_eff[bi * kQuartersPerBlock + q] = eff;
        }
        // drain: copies already in flight must land before their slots are reused by the next item
        while (consumed < issued) {
            mbar_wait(&sm.full[consumed % kFwdStages], (consumed / kFwdStages) & 1);
            ++consumed;
        }
        __syncwarp();
        const bool writer = inside && slot == 0;
        if (a.loss_target) {
            float part = 0.0f;
            if (writer) part = loss_pixel(a, size_t(r) * P, P, size_t(py) * a.g.W + px, C0 + T * bg0, C1 + T * bg1, C2 + T * bg2);
            part = warp_sum(part);
            if (lane == 0) a.loss_part[item_slot] = part;
        }
        write_pixel<4>(a, r, P, px, py, writer, C0 + T * bg0, C1 + T * bg1, C2 + T * bg2, D, Wt, last);
        if (a.tile_time && lane == 0) {          // diagnostics: earliest start, duration of the slowest quarter
            const unsigned long long now = global_timer_ns();
            if (blk == 0 && q == 0) a.tile_time[tg].x = (unsigned int)t_begin;
            atomicMax(&a.tile_time[tg].y, (unsigned int)(now - t_begin));
        }
    }

    // ---------------- blocks of tiles without instances: background only (lane = pixel of the 8x4 block)
    // whole tiles are popped (8 uniform items per global atomic round trip)
    for (unsigned int item = 0xffffffffu;;) {
        if (item == 0xffffffffu || (item % kBlocksPerTile) == kBlocksPerTile - 1) {
            unsigned int w = 0;
            if (lane == 0) w = atomicAdd(&a.wc->empty_cursor, unsigned(kBlocksPerTile));
            item = __shfl_sync(kFull, w, 0);
        } else {
            ++item;
        }
        if (item >= n_empty_items) break;
        const unsigned int tile_local = a.work_empty[item / kBlocksPerTile];
        const int blk = item % kBlocksPerTile;
        const int rl = tile_local / a.g.num_tiles;
        const int tile = tile_local - rl * a.g.num_tiles;
        const int r = a.render_base + rl;
        const int tx = tile % a.g.tiles_x, ty = tile / a.g.tiles_x;
        const int px = tx * kTile + (blk & 1) * kBlockW + (lane & 7), py = ty * kTile + (blk >> 1) * kBlockH + (lane >> 3);
        const bool inside = px < a.g.W && py < a.g.H;
        if (a.loss_target) {
            float part = 0.0f;
            if (inside) part = loss_pixel(a, size_t(r) * P, P, size_t(py) * a.g.W + px, bg0, bg1, bg2);
            part = warp_sum(part);
            // one partial per (block, quarter) slot: the block's sum goes to quarter 0
            if (lane < kQuartersPerBlock)
                a.loss_part[(size_t(tile_local) * kBlocksPerTile + blk) * kQuartersPerBlock + lane] = lane == 0 ? part : 0.0f;
        }
        write_pixel<8>(a, r, P, px, py, inside, bg0, bg1, bg2, 0.0f, 0.0f, 0u);
    }
}

// ------------------------------------------------------------------------------------------------ backward
struct BwdArgs {
    RenderGeom g;
    int render_base;
    const unsigned int* blk_off;
    const unsigned int* blk_cnt;
    const unsigned int* q_eff;
    const unsigned int* bids;
    const float4 *rec0, *rec1, *rec2;     // block records (masks refined by the forward)
    const float* bg;
    const unsigned int* n_contrib;
    const unsigned char* clamp_mask;      // NULL: the forward did not clamp
    const float *out_alpha, *dL_dcolor, *dL_ddepth, *dL_dalpha;
    const float* loss_dL_dcolor;          // fused-loss gradient (already zero where the clamp saturated), or NULL
    const float* dL_dfeed;                // gradient w.r.t. the LPIPS feed [R,3,H/2,W/2], or NULL
    float* accum;
    size_t plane;
    const float4* ck0;            // forward checkpoints (sgr_common.cuh::kSegB)
    const float* ck1;
    const uint2* bwd_items;       // [kBwdClasses][stride]; this chunk's items start at plan->item_base in every class
    unsigned long long bwd_items_stride;
    ChunkPlan* plan;
    const float* dL_scale;        // device scalar multiplying loss_dL_dcolor, or NULL
};

// Reduce-scatter of 8 per-lane values over the 8 pixel lanes of a slot group (lanes pl = 0..7): after the three
// exchange levels lane pl holds the full sum of value `pl`.  7 shuffles instead of 24.
__device__ __forceinline__ float reduce_scatter8(const float (&v)[8], int pl) {
    const bool h4 = (pl & 4) != 0, h2 = (pl & 2) != 0, h1 = (pl & 1) != 0;
    float w4[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {                // lanes with pl & 4 keep values 4..7
        const float send = h4 ? v[i] : v[i + 4];
        const float keep = h4 ? v[i + 4] : v[i];
        w4[i] = keep + __shfl_xor_sync(kFull, send, 4);
    }
    float w2[2];
#pragma unroll
    for (int i = 0; i < 2; ++i) {                // ... pl & 2 keep the upper pair
        const float send = h2 ? w4[i] : w4[i + 2];
        const float keep = h2 ? w4[i + 2] : w4[i];
        w2[i] = keep + __shfl_xor_sync(kFull, send, 2);
    }
    const float send = h1 ? w2[0] : w2[1];
    const float keep = h1 ? w2[1] : w2[0];
    return keep + __shfl_xor_sync(kFull, send, 1);
}

// Work item = (block list, quarter, segment): the quarter replays the list entries [lo, hi) of its segment back to
// front, lo = segment * kSegB, hi = min(lo + kSegB, q_eff) (q_eff = the last record any of its pixels blended, + 1).
// A pixel whose contributors end inside the segment starts from its final state (T = 1 - alpha_out, nothing behind);
// a pixel that continues behind the segment starts from the forward's checkpoints: T = T_ck and
// U = (C_fin - C_ck) . dL/dC + (D_fin - D_ck) dL/dD — so the long face / hand lists do not serialise on one warp.
template <bool kDepthAlphaGrads, bool kExact>
__global__ void __launch_bounds__(kBlendThreads, SGR_BWD_MIN_CTAS) blend_backward_kernel(BwdArgs a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    BwdSmem& sm = reinterpret_cast<BwdSmem*>(smem_raw)[warp];
    const size_t P = size_t(a.g.H) * a.g.W;
    const float bg0 = a.bg[0], bg1 = a.bg[1], bg2 = a.bg[2];
    unsigned int cls_end[kBwdClasses];            // items are popped class by class, most blended records first
    {
        unsigned int run = 0;
#pragma unroll
        for (int c = 0; c < kBwdClasses; ++c) { run += a.plan->n_items[c]; cls_end[c] = run; }
    }
    const unsigned int n_items = cls_end[kBwdClasses - 1];
    const unsigned int item_base = a.plan->item_base;
    const float ddelx_dx = 0.5f * float(a.g.W), ddely_dy = 0.5f * float(a.g.H);
    const float gscale = a.dL_scale ? *a.dL_scale : 1.0f;
    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < kBwdStages; ++s) mbar_init(&sm.full[s], 1);
        fence_barrier_init();
    }
    __syncwarp();
    const int pl = lane & 7, slot = lane >> 3;
    if (lane < kBwdStages) {      // the sentinel record: opacity 0 -> alpha 0; threshold +inf -> never valid (exact mode)
        sm.r0[lane][kBwdBatch] = make_float4(0.0f, 0.0f, 0.0f, __int_as_float(0x7f800000));
        sm.r1[lane][kBwdBatch] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        sm.r2[lane][kBwdBatch] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    }
    __syncwarp();
    unsigned int issued = 0, consumed = 0;
    // reduce_scatter8 leaves component pl in lane pl: 0 mean2D.x, 1 mean2D.y, 2..4 conic, 5 opacity, 6..7 colour r, g
    // (colour b and the depth term take a plain butterfly); this is the lane's accumulator plane and its scale
    const float out_scale = pl == 0 ? ddelx_dx : pl == 1 ? ddely_dy : (pl >= 2 && pl <= 4) ? -0.5f : 1.0f;

    for (;;) {
        const unsigned int item = pop_item(&a.plan->cursor, n_items, lane);
        if (item == 0xffffffffu) break;
        int cls = 0;
#pragma unroll
        for (int c = 0; c < kBwdClasses - 1; ++c) cls += item >= cls_end[c] ? 1 : 0;
        const unsigned int in_cls = item - (cls ? cls_end[cls - 1] : 0u);
        const uint2 ws = a.bwd_items[size_t(cls) * a.bwd_items_stride + item_base + in_cls];
        const unsigned int tile_local = ws.x / kItemsPerTile;
        const int blk = (ws.x % kItemsPerTile) / kQuartersPerBlock;
        const int q = ws.x % kQuartersPerBlock;
        const unsigned int lo = ws.y * unsigned(kSegB);
        const int rl = tile_local / a.g.num_tiles;
        const int tile = tile_local - rl * a.g.num_tiles;
        const int r = a.render_base + rl;
        const size_t tg = size_t(r) * a.g.num_tiles + tile;
        const size_t bi = tg * kBlocksPerTile + blk;
        const int tx = tile % a.g.tiles_x, ty = tile / a.g.tiles_x;
        const int qx0 = tx * kTile + (blk & 1) * kBlockW + (q & 1) * 4, qy0 = ty * kTile + (blk >> 1) * kBlockH + (q >> 1) * 2;
        const int px = qx0 + (pl & 3), py = qy0 + (pl >> 2);
        const bool inside = px < a.g.W && py < a.g.H;
        const size_t pix = size_t(py) * a.g.W + px;
        const unsigned int last = inside ? a.n_contrib[size_t(r) * P + pix] : 0u;     // position in the TILE list + 1
        const unsigned int eff = a.q_eff[bi * kQuartersPerBlock + q];
        if (eff <= lo) continue;
        const unsigned int hi = min(lo + unsigned(kSegB), eff);
        const size_t off = a.blk_off[bi];
        const unsigned int n = a.blk_cnt[bi];
        float T_final = 1.0f, dp0 = 0, dp1 = 0, dp2 = 0, ddep = 0, dalp = 0;
        if (inside) {
            T_final = 1.0f - a.out_alpha[size_t(r) * P + pix];
            // d loss / d (unclamped colour): the caller's gradients w.r.t. the returned image / LPIPS feed, masked
            // where the forward's clamp saturated, plus the fused loss's own gradient
            if (a.dL_dcolor) {
                const float* dc = a.dL_dcolor + size_t(r) * 3 * P;
                dp0 = dc[pix]; dp1 = dc[P + pix]; dp2 = dc[2 * P + pix];
            }
            if (a.dL_dfeed) {
                const size_t Pq = P >> 2;
                const float* df = a.dL_dfeed + size_t(r) * 3 * Pq + size_t(py >> 1) * (a.g.W >> 1) + (px >> 1);
                dp0 = fmaf(0.5f, df[0], dp0); dp1 = fmaf(0.5f, df[Pq], dp1); dp2 = fmaf(0.5f, df[2 * Pq], dp2);
            }
            if (a.clamp_mask) {
                const unsigned int cm = a.clamp_mask[size_t(r) * P + pix];
                if (cm & 1u) dp0 = 0.0f;
                if (cm & 2u) dp1 = 0.0f;
                if (cm & 4u) dp2 = 0.0f;
            }
            if (a.loss_dL_dcolor) {
                const float* dc = a.loss_dL_dcolor + size_t(r) * 3 * P;
                dp0 = fmaf(dc[pix], gscale, dp0); dp1 = fmaf(dc[P + pix], gscale, dp1); dp2 = fmaf(dc[2 * P + pix], gscale, dp2);
            }
            if (kDepthAlphaGrads) {
                if (a.dL_ddepth) ddep = a.dL_ddepth[size_t(r) * P + pix];
                if (a.dL_dalpha) dalp = a.dL_dalpha[size_t(r) * P + pix];
            }
        }
        const unsigned int nb = (hi - lo + kBwdBatch - 1) / kBwdBatch;
        const float pxf = float(px), pyf = float(py);
        const float4 *g0 = a.rec0 + off + lo, *g1 = a.rec1 + off + lo, *g2 = a.rec2 + off + lo;
        // per-pixel state (identical in the pixel's four slot lanes): transmittance in front of the records handled
        // so far (walking back to front) and U = sum over the records behind of w_j (c_j . dL/dC + z_j dL/dD)
        float T = T_final, U = 0.0f;
        // A pixel has contributors behind this segment iff its last contributor sits at or behind the first record
        // after the segment (records are in tile-list order; word 2 >> 4 = position in the tile list).
        bool resume = false;
        if (lo + unsigned(kSegB) < n) {
            const unsigned int next_pos = __float_as_uint(__ldg(&a.rec0[off + lo + kSegB].z)) >> 4;
            resume = last > next_pos;
        }
        if (resume) {                            // resume from the checkpoints
            const size_t slot0 = off / (kSegB / 2);
            const size_t ci = (slot0 + ws.y) * 32 + q * 8 + pl;
            const size_t fi = (slot0 + (n + kSegB - 1) / kSegB - 1) * 32 + q * 8 + pl;
            const float4 ck = a.ck0[ci], fin = a.ck0[fi];
            T = ck.x;
            U = (fin.y - ck.y) * dp0;
            U = fmaf(fin.z - ck.z, dp1, U);
            U = fmaf(fin.w - ck.w, dp2, U);
            if (kDepthAlphaGrads) U = fmaf(a.ck1[fi] - a.ck1[ci], ddep, U);
        }
        const float bg_dot = (bg0 * dp0 + bg1 * dp1) + bg2 * dp2;
        const float K = T_final * (bg_dot - dalp);     // dL/dalpha_k = T_k (c_k . g) - (U_k + K) / (1 - alpha_k)
        float* acc = a.accum + size_t(rl) * a.g.N;
        const unsigned int* ids = a.bids + off + lo;

        // walk step k handles list batch (nb - 1 - k)
        unsigned int b_issued = 0;
        while (b_issued < nb && b_issued < unsigned(kBwdStages - 1)) {     // prefetch depth of the ring
            const unsigned int lb = nb - 1 - b_issued;
            ring_issue(sm, issued, g0 + lb * kBwdBatch, g1 + lb * kBwdBatch, g2 + lb * kBwdBatch,
                       min(unsigned(kBwdBatch), hi - lo - lb * kBwdBatch), lane);
            ++issued; ++b_issued;
        }
        for (unsigned int b = 0; b < nb; ++b) {
            if (b_issued < nb) {
                const unsigned int lb = nb - 1 - b_issued;
                ring_issue(sm, issued, g0 + lb * kBwdBatch, g1 + lb * kBwdBatch, g2 + lb * kBwdBatch,
                           min(unsigned(kBwdBatch), hi - lo - lb * kBwdBatch), lane);
                ++issued; ++b_issued;
            }
            const int s = consumed % kBwdStages;
            mbar_wait(&sm.full[s], (consumed / kBwdStages) & 1);
            ++consumed;
            const unsigned int bbase = (nb - 1 - b) * kBwdBatch;          // segment-local index of the batch's first record
            const unsigned int m = min(unsigned(kBwdBatch), hi - lo - bbase);
            const float4* r0 = sm.r0[s];
            const float4* r1 = sm.r1[s];
            const float4* r2 = sm.r2[s];
            unsigned int qbits[kBwdBatch / 32];
            const unsigned int cnt = cull_batch<true, kBwdBatch>(r0, m, q, sm.list, lane, qbits);
            // super-trip: slot s handles the quarter's (t0 + s)-th survivor from the back
            for (unsigned int t0 = 0; t0 < cnt; t0 += kSlots * kBwdGroups) {
#pragma unroll
                for (int g = 0; g < kBwdGroups; ++g) {
                    const unsigned int j = sm.list[t0 + kSlots * g + slot];
                    const float4 q0 = r0[j], q1 = r1[j], q2 = r2[j];
                    const float dx = q0.x - pxf, dy = q0.y - pyf;
                    const float power = gauss_power(q1.x, q1.y, q1.z, dx, dy);
                    const bool has = (__float_as_uint(q0.z) >> 4) < last;
                    const bool valid = has && (kExact ? (!(power > 0.0f) && !(power < q0.w)) : !(power > 0.0f));
                    const float G = blend_exp<kExact>(valid ? power : 0.0f);
                    const float alpha = fminf(kAlphaMax, q1.w * G);
                    const float al = (valid && !(alpha < kAlphaMin)) ? alpha : 0.0f;   // 0: the pixel did not blend it
                    // c . g of this lane's record
                    float cg = q2.x * dp0;
                    cg = fmaf(q2.y, dp1, cg);
                    cg = fmaf(q2.z, dp2, cg);
                    if (kDepthAlphaGrads) cg = fmaf(q2.w, ddep, cg);
                    // T_k = T / prod over the slots up to mine of (1 - alpha): inclusive prefix product of 1 / (1 - alpha)
                    const float rinv = rcp_approx(1.0f - al);
                    float pinc = rinv;
                    float v = __shfl_up_sync(kFull, pinc, 8);
                    pinc = slot >= 1 ? pinc * v : pinc;
                    v = __shfl_up_sync(kFull, pinc, 16);
                    pinc = slot >= 2 ? pinc * v : pinc;
                    const float Tk = T * pinc;
                    const float w = al * Tk;                 // blend weight of this record in this pixel
                    const float u = w * cg;
                    // U_k = U + sum of u over the slots before mine (they are behind my record)
                    float uinc = u;
                    v = __shfl_up_sync(kFull, uinc, 8);
                    uinc = slot >= 1 ? uinc + v : uinc;
                    v = __shfl_up_sync(kFull, uinc, 16);
                    uinc = slot >= 2 ? uinc + v : uinc;
                    const float Uk = U + (uinc - u);
                    const float dal = al != 0.0f ? fmaf(Tk, cg, -(Uk + K) * rinv) : 0.0f;    // dL/dalpha
                    T *= __shfl_sync(kFull, pinc, pl + 24);
                    U += __shfl_sync(kFull, uinc, pl + 24);
                    // gradient terms of (this pixel, this record)
                    const float dL_dG = q1.w * dal;
                    const float gdx = G * dx, gdy = G * dy;
                    float t[8];
                    // -gdx*A - gdy*B = 2*gdx*hA + gdy*nB   (rec1 = (-A/2, -B, -C/2, o))
                    t[0] = dL_dG * fmaf(2.0f * q1.x, gdx, q1.y * gdy);
                    t[1] = dL_dG * fmaf(2.0f * q1.z, gdy, q1.y * gdx);
                    t[2] = gdx * dx * dL_dG;
                    t[3] = gdx * dy * dL_dG;
                    t[4] = gdy * dy * dL_dG;
                    t[5] = G * dal;
                    t[6] = w * dp0;
                    t[7] = w * dp1;
                    float t8 = w * dp2, t9 = kDepthAlphaGrads ? w * ddep : 0.0f;
                    // reduce over the 8 pixels of the quarter: lane pl ends up with component pl of its slot's record;
                    // components 8 (and 9) take a plain butterfly
                    const float red = reduce_scatter8(t, pl);
                    t8 += __shfl_xor_sync(kFull, t8, 4); t8 += __shfl_xor_sync(kFull, t8, 2); t8 += __shfl_xor_sync(kFull, t8, 1);
                    if (kDepthAlphaGrads) {
                        t9 += __shfl_xor_sync(kFull, t9, 4); t9 += __shfl_xor_sync(kFull, t9, 2); t9 += __shfl_xor_sync(kFull, t9, 1);
                    }
                    if (j < unsigned(kBwdBatch)) {           // not the sentinel
                        const unsigned int id = __ldg(ids + bbase + j);
                        float* gacc = acc + id;
                        if (red != 0.0f) atomicAdd(gacc + size_t(pl) * a.plane, red * out_scale);
                        if (pl == 0 && t8 != 0.0f) atomicAdd(gacc + 8 * a.plane, t8);
                        if (kDepthAlphaGrads && pl == 1 && t9 != 0.0f) atomicAdd(gacc + 9 * a.plane, t9);
                    }
                }
            }
            __syncwarp();
        }
        __syncwarp();
    }
}

// Fused loss: fixed-order sum of the chunk's per-item partials (double), added to the device scalar.
__global__ void __launch_bounds__(1024) loss_reduce_kernel(const float* part, unsigned int n, float* loss_out, int first,
                                                           float scale) {
    __shared__ double sh[1024];
    double acc = 0.0;
    for (unsigned int i = threadIdx.x; i < n; i += 1024) acc += double(part[i]);
    sh[threadIdx.x] = acc;
    __syncthreads();
    for (int o = 512; o > 0; o >>= 1) {
        if (int(threadIdx.x) < o) sh[threadIdx.x] += sh[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) *loss_out = (first ? 0.0f : *loss_out) + float(sh[0] * double(scale));
}

int g_num_sms[kMaxDevices] = {};
int num_sms() {
    const int d = current_device_slot();
    if (g_num_sms[d] == 0) {
        cudaDeviceGetAttribute(&g_num_sms[d], cudaDevAttrMultiProcessorCount, d);
        if (g_num_sms[d] <= 0) g_num_sms[d] = 148;
    }
    return g_num_sms[d];
}

// Opts the kernel into its dynamic shared memory size and returns the resident CTAs per SM (cached per kernel).
template <typename K>
cudaError_t prepare_kernel(K kernel, size_t smem, int* per_sm) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
    if (e != cudaSuccess) return e;
    int n = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kernel, kBlendThreads, smem);
    if (e != cudaSuccess) return e;
    *per_sm = n > 0 ? n : 1;
    return cudaSuccess;
}

template <bool kDA, bool kExact>
cudaError_t launch_bwd_variant(const BwdArgs& a, long long want, cudaStream_t stream) {
    constexpr size_t smem = sizeof(BwdSmem) * kWarpsPerCta;
    static int per_sm_dev[kMaxDevices] = {};
    int& per_sm = per_sm_dev[current_device_slot()];
    if (per_sm == 0) {
        cudaError_t e = prepare_kernel(blend_backward_kernel<kDA, kExact>, smem, &per_sm);
        if (e != cudaSuccess) return e;
    }
    blend_backward_kernel<kDA, kExact><<<int(min((long long)num_sms() * per_sm, want)), kBlendThreads, smem, stream>>>(a);
    return cudaGetLastError();
}

}  // namespace

cudaError_t launch_blend_forward(const ChunkCtx& c, float* out_color, float* out_depth, float* out_alpha,
                                 float* out_feed) {
    FwdArgs a;
    a.g = c.g; a.render_base = c.render_base; a.blk_off = c.blk_off; a.blk_cnt = c.blk_cnt; a.q_eff = c.q_eff;
    a.rec0 = c.brec0; a.rec1 = c.brec1; a.rec2 = c.brec2; a.bg = c.p->bg; a.n_contrib = c.n_contrib;
    a.rec0_words = reinterpret_cast<unsigned int*>(c.brec0);
    a.refine_masks = (c.p->flags & SGR_FLAG_FORWARD_ONLY) ? 0 : 1;
    a.tile_time = (c.p->flags & SGR_FLAG_TILE_TIMING) ? c.tile_time : nullptr;
    a.out_color = out_color; a.out_depth = out_depth; a.out_alpha = out_alpha; a.out_feed = out_feed;
    a.clamp_mask = c.clamp_mask;
    a.ck0 = c.ck0; a.ck1 = c.ck1; a.plan = c.plan; a.bwd_items = c.bwd_items; a.bwd_items_stride = c.bwd_items_stride;
    a.work_blend = c.work_blend; a.work_empty = c.work_empty; a.wc = c.work_counts;
    a.clamp_color = ((c.p->flags & SGR_FLAG_CLAMP_COLOR) || c.loss_target) ? 1 : 0;
    a.loss_target = c.loss_target; a.loss_mask = c.loss_mask; a.loss_dL_dcolor = c.loss_dL_dcolor;
    a.loss_part = c.loss_part; a.loss_scale = c.loss_scale;
    constexpr size_t smem = sizeof(FwdSmem) * kWarpsPerCta;
    const bool exact = (c.p->flags & SGR_FLAG_EXACT_EXP) != 0;
    static int per_sm_dev[kMaxDevices][2] = {};
    int& per_sm = per_sm_dev[current_device_slot()][exact ? 1 : 0];
    if (per_sm == 0) {
        cudaError_t e = exact ? prepare_kernel(blend_forward_kernel<true>, smem, &per_sm)
                              : prepare_kernel(blend_forward_kernel<false>, smem, &per_sm);
        if (e != cudaSuccess) return e;
    }
    const long long items = (long long)c.num_renders * c.g.num_tiles * kItemsPerTile;
    static int ctas_override = -1;               // experiment hook: SGR_FWD_CTAS_PER_SM
    if (ctas_override < 0) { const char* v = getenv("SGR_FWD_CTAS_PER_SM"); ctas_override = v ? atoi(v) : 0; }
    const int use_per_sm = ctas_override > 0 ? min(ctas_override, per_sm) : per_sm;
    const int grid = int(min((long long)num_sms() * use_per_sm, (items + kWarpsPerCta - 1) / kWarpsPerCta));
    if (exact) blend_forward_kernel<true><<<grid, kBlendThreads, smem, c.stream>>>(a);
    else blend_forward_kernel<false><<<grid, kBlendThreads, smem, c.stream>>>(a);
    return cudaGetLastError();
}

cudaError_t launch_loss_reduce(const ChunkCtx& c, float* loss_out) {
    const unsigned int n = unsigned(c.num_renders) * c.g.num_tiles * kItemsPerTile;
    loss_reduce_kernel<<<1, 1024, 0, c.stream>>>(c.loss_part, n, loss_out, c.render_base == 0 ? 1 : 0, c.loss_scale);
    return cudaGetLastError();
}

cudaError_t launch_blend_backward(const ChunkCtx& c, const SgrBackwardArgs& b) {
    BwdArgs a;
    a.g = c.g; a.render_base = c.render_base; a.blk_off = c.blk_off; a.blk_cnt = c.blk_cnt; a.q_eff = c.q_eff;
    a.bids = c.bids; a.rec0 = c.brec0; a.rec1 = c.brec1; a.rec2 = c.brec2; a.bg = c.p->bg;
    a.n_contrib = c.n_contrib; a.out_alpha = b.out_alpha; a.dL_dcolor = b.dL_dcolor; a.dL_ddepth = b.dL_ddepth;
    a.dL_dalpha = b.dL_dalpha; a.loss_dL_dcolor = b.loss_dL_dcolor; a.dL_dfeed = b.dL_dlpips_feed;
    a.clamp_mask = ((c.p->flags & SGR_FLAG_CLAMP_COLOR) || b.fused_clamp) ? c.clamp_mask : nullptr;
    a.accum = c.accum; a.plane = size_t(c.num_renders) * c.g.N;
    a.ck0 = c.ck0; a.ck1 = c.ck1; a.bwd_items = c.bwd_items; a.bwd_items_stride = c.bwd_items_stride;
    a.plan = c.plan; a.dL_scale = b.dL_dcolor_scale;
    const long long items = (long long)c.num_renders * c.g.num_tiles * kItemsPerTile;
    const long long want = (items + kWarpsPerCta - 1) / kWarpsPerCta;
    const bool exact = (c.p->flags & SGR_FLAG_EXACT_EXP) != 0;
    if (b.dL_ddepth || b.dL_dalpha)
        return exact ? launch_bwd_variant<true, true>(a, want, c.stream) : launch_bwd_variant<true, false>(a, want, c.stream);
    return exact ? launch_bwd_variant<false, true>(a, want, c.stream) : launch_bwd_variant<false, false>(a, want, c.stream);
}

}  // namespace sgr
