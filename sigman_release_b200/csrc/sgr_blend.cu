// sgr_blend.cu — the alpha-compositing kernels (forward and backward) of the B200 rasteriser.
//
// Replaces upstream renderCUDA forward/backward (third-party diff_gaussian_rasterization; call site
// /root/reference/core/gaussians/gs.py:99-106).  Design (DESIGN.md "blend"):
//   * the unit of work is one 8x4 pixel block of one tile of one render.  Every WARP is autonomous: it pops work items
//     from a device-side queue ordered longest-list-first (sgr_binning.cu::plan_kernel), streams the BLOCK's
//     depth-ordered list of 8-byte entries (Gaussian id, tile-list position << 4 | quarter mask; the per-tile sort
//     emits one contiguous list per block that holds only the instances whose extent touches the block) and gathers
//     those Gaussians' 48-byte records (three float4 arrays written once per (render, Gaussian) by the preprocess
//     kernel) into its own shared-memory ring with asynchronous copies (cp.async; the backward's index ring is fed
//     by 1-D TMA bulk copies completing on its own mbarriers), and never waits for another warp — no block-wide
//     barrier, no producer/consumer hand-off, early exit as soon as its 32 pixels are finished;
//   * a block is four 4x2 quarters.  A batch of 128 records is culled in straight-line code (lane = record): the
//     per-record 4-bit quarter masks built by the tile sort are read, one ballot per (round, quarter), and the
//     survivors' batch-local indices are compacted into one sentinel-padded list per quarter; every quarter then
//     walks only its own survivors (lane = pixel), so up to four different Gaussians are evaluated per trip.  Culled
//     records would have been skipped by the alpha test, so the per-pixel arithmetic, the contributor index and every
//     output bit equal the straightforward kernel's;
//   * forward trips run in groups of eight (four at the end of a list): eight independent alpha chains written
//     load-first, then the sequential compositing recurrence; per batch the forward clears the mask bits of the
//     (quarter, record) pairs that did not blend (a plain store: the block owns its records) so that the backward
//     culls on the exact set, checkpoints the running state every kSegB records and pushes one backward work item
//     per segment that blended anything, classed by the number of records that blended;
//   * backward: work items are (block, kSegB-record segment), resumed from the forward's checkpoints; the same
//     walk in reverse over descending lists with phases A+B (alpha, G, then the per-pixel transmittance / colour
//     recurrences producing dL/dalpha and the blend weight, stashed) and C: the roles flip to lane = (Gaussian,
//     quarter) pair, each lane summing the gradient terms of its quarter's 8 pixels in registers (XOR-swizzled stash,
//     no shuffles) before three 16-byte vector atomics into the Gaussian's accumulator row (red.global.add.v4.f32).
//     Gradient arithmetic is free to use FMA (tolerance, not bit-exact);
//   * the block lists of at least kDenseEntries entries are queued separately by the tile sort and walked on "dense"
//     CTAs that keep one dense walk + one ordinary warp per scheduler awake (sgr_common.cuh): a warp is latency-bound,
//     and the longest walks otherwise outlast the rest of the launch;
//   * optional epilogue: clamp + masked L1 loss + dL/dcolour (SgrForwardArgs::loss_*).
//
// Compiled with --fmad=false: the per-pixel expressions are the oracle's (oracle/sgr_oracle.cpp::blend_forward /
// blend_backward) evaluated in the same order.  exp(power) is a template switch: kExact = the oracle's exp_spec
// sequence (SGR_FLAG_EXACT_EXP: colour, depth, alpha and n_contrib bit-exact), default = the SFU's ex2 like upstream's
// own exp() (sgr_common.cuh::exp_fast; ~1e-6 from the oracle).
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
constexpr int kFwdBatch = SGR_FWD_BATCH;       // records per ring stage (culled 32 at a time: lane = record), forward
#ifndef SGR_BWD_BATCH
#define SGR_BWD_BATCH 128
#endif
#ifndef SGR_BWD_STAGES
#define SGR_BWD_STAGES 1
#endif
constexpr int kBwdBatch = SGR_BWD_BATCH;       // ... backward
constexpr int kFwdStages = SGR_FWD_STAGES;     // per-warp TMA ring depth, forward
constexpr int kBwdStages = SGR_BWD_STAGES;     // backward: one 128-record stage (the stash takes the rest of the budget)
constexpr int kWarpsPerCta = 8;                // backward: two CTAs of 8 warps per SM
constexpr int kBlendThreads = kWarpsPerCta * 32;
// forward: ONE CTA of 16 warps per SM: the CTAs that take the dense block lists (sgr_common.cuh::kDenseEntries) keep
// their other warps asleep meanwhile, so that a dense walk has a scheduler (almost) to itself
constexpr int kFwdWarps = 16;
constexpr int kFwdThreads = kFwdWarps * 32;
constexpr int kBlockW = 8, kBlockH = 4;        // pixel block of one warp (four 4x2 quarters walk separate lists)
#ifndef SGR_BWD_SLOTS
#define SGR_BWD_SLOTS 16
#endif
constexpr int kBwdSlots = SGR_BWD_SLOTS;       // backward: trips per phase pass (8 or 16: the stash swizzle)
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
// 16-byte vector reduction (sm_90+): four float adds to one aligned 16-byte chunk in one L2 transaction.
__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ unsigned long long global_timer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

constexpr int kListPad = 16;                   // keeps the rows 16-byte aligned

constexpr int kIxStages = 2;                   // backward index batches: one being walked, one in flight

// Forward: a 2-stage ring of gathered records; the index entries travel through registers.
template <int kNumStages, int kBatch>
struct FwdWarpSmem {
    float4 r0[kNumStages][kBatch + 1];          // gathered records; slot kBatch of every stage = the sentinel record
    float4 r1[kNumStages][kBatch + 1];
    float4 r2[kNumStages][kBatch + 1];
    // per quarter: batch-local indices of the survivors, sentinel-filled; the trip loops read whole groups of 4 / 8
    // entries starting at multiples of their size, which stays inside kBatch — the pad is a safety margin
    unsigned char list[4][kBatch + kListPad];
    unsigned int hit[kBatch + 1];               // byte q of word j != 0 <=> quarter q blended record j
};
// Backward: one stage of gathered records, the index entries in a TMA-fed ring, the phase B -> C stash.
template <int kDepth, int kBatch>
struct BwdWarpSmem {
    float4 r0[1][kBatch + 1];
    float4 r1[1][kBatch + 1];
    float4 r2[1][kBatch + 1];
    alignas(16) uint2 ix[kIxStages][kBatch];    // block-list entries (Gaussian id, tile-list position << 4 | quarter mask)
    float stash[2][kDepth][32];
    float dpix[4][32];                          // dL/dcolor (3) and dL/ddepth of the block's pixels
    unsigned char list[4][kBatch + kListPad];
    uint64_t ixbar[kIxStages];
};

// Cull a batch of m <= kBatch block-list entries against the four 4x2 quarters of the warp's 8x4 pixel block:
// 32 entries per round (lane = entry) read their 4-bit quarter mask (low bits of the entry, built by the tile sort
// from sgr_common.cuh::quarter_mask and refined by the forward), one ballot per quarter, warp-parallel compaction of the survivors'
// batch-local indices into list[quarter][...] (ascending; descending with kReverse — the backward walks back to front);
// unused list entries point at the sentinel record.  Returns the four survivor counts.
template <bool kReverse, int kBatch>
__device__ __forceinline__ uint4 cull_batch(const uint2 (&ent)[kBatch / 32], unsigned int m,
                                            unsigned char (*list)[kBatch + kListPad],
                                            int lane, unsigned int (&bits)[kBatch / 32]) {
    // Straight-line code (the per-batch overhead is latency, not work): all mask words are loaded first, then all
    // ballots, then the compaction stores.  ent[r] = this lane's entry 32 * r + lane of the batch (word y = position << 4 | mask), bits[r] = its mask.
    constexpr int R = kBatch / 32;
    const unsigned int lt = kReverse ? ~((2u << lane) - 1u) : (1u << lane) - 1u;   // lanes before / after this one
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const unsigned int e = 32u * r + lane;
        bits[r] = (e < m) ? (ent[r].y & 0xfu) : 0u;
    }
    {                             // every list entry the compaction does not overwrite points at the sentinel record
        constexpr unsigned int fill = kBatch * 0x01010101u;
        uint4* lw = reinterpret_cast<uint4*>(&list[0][0]);
        constexpr int kVecs = 4 * (kBatch + kListPad) / 16;       // the four rows including their pads
#pragma unroll
        for (int v = 0; v < kVecs; v += 32)
            if (v + lane < kVecs) lw[v + lane] = make_uint4(fill, fill, fill, fill);
    }
    unsigned int mq[4][R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
#pragma unroll
        for (int q = 0; q < 4; ++q) mq[q][r] = __ballot_sync(kFull, bits[r] & (1u << q));
    }
    __syncwarp();                 // the sentinel fill is ordered before the compaction stores
    unsigned int n[4] = {0, 0, 0, 0};
#pragma unroll
    for (int k = 0; k < R; ++k) {
        const int r = kReverse ? R - 1 - k : k;                 // descending lists walk the rounds from the back
        const unsigned int e = 32u * r + lane;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            if (bits[r] & (1u << q)) list[q][n[q] + __popc(mq[q][r] & lt)] = static_cast<unsigned char>(e);
            n[q] += __popc(mq[q][r]);
        }
    }
    __syncwarp();
    return make_uint4(n[0], n[1], n[2], n[3]);
}

// Per-warp index ring: the k-th index batch of the kernel's lifetime goes to stage k % kIxStages and completes phase
// (k / kIxStages) & 1 of that stage's mbarrier, so the barriers never need re-initialisation between work items.
// One 1-D TMA bulk copy of the batch's entries, rounded up to a multiple of 4 (lists start 16-byte aligned and are
// padded to a multiple of 4 entries by the tile sort).
template <typename Smem>
__device__ __forceinline__ void ix_issue(Smem& sm, unsigned int k, const uint2* src, unsigned int m, int lane) {
    if (lane == 0) {
        const int st = k % kIxStages;
        const uint32_t bytes = ((m + 3u) & ~3u) * 8u;
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // earlier generic reads of the stage are done
        mbar_arrive_expect_tx(&sm.ixbar[st], bytes);
        tma_load_1d(sm.ix[st], src, bytes, &sm.ixbar[st]);
    }
}
template <typename Smem>
__device__ __forceinline__ void ix_wait(Smem& sm, unsigned int k) {
    mbar_wait(&sm.ixbar[k % kIxStages], (k / kIxStages) & 1);
}

__device__ __forceinline__ void cp_async16(void* dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// Gathers the records of a batch of block-list entries from the render's per-Gaussian arrays into ring stage rs
// (lane = entry; asynchronous 16-byte copies, one commit group per batch).  Entries whose quarter mask is 0 (pads,
// and — in the backward — records the forward found not to blend) are not fetched: the cull never selects them.
template <int kRounds, typename Smem>
__device__ __forceinline__ void gather_batch(Smem& sm, int rs, const uint2 (&ent)[kRounds], unsigned int m,
                                             const float4* t0, const float4* t1, const float4* t2, int lane) {
#pragma unroll
    for (int rr = 0; rr < kRounds; ++rr) {
        const unsigned int e = 32u * rr + lane;
        if (e < m && (ent[rr].y & 0xfu)) {
            const unsigned int p = ent[rr].x;
            cp_async16(&sm.r0[rs][e], t0 + p);
            cp_async16(&sm.r1[rs][e], t1 + p);
            cp_async16(&sm.r2[rs][e], t2 + p);
        }
    }
    cp_async_commit();
}

// Once a batch's gather has landed: word 2 of rec0 (the packed extent, used by the tile sort only) is replaced by the
// record's position in the tile list (upstream's contributor index), which the walk compares with / reports as n_contrib.
template <int kRounds, typename Smem>
__device__ __forceinline__ void patch_positions(Smem& sm, int rs, const uint2 (&ent)[kRounds], unsigned int m, int lane) {
#pragma unroll
    for (int rr = 0; rr < kRounds; ++rr) {
        const unsigned int e = 32u * rr + lane;
        if (e < m && (ent[rr].y & 0xfu)) sm.r0[rs][e].z = __uint_as_float(ent[rr].y >> 4);
    }
}

// This lane's entries 32 * r + lane of a batch of m <= kBatch block-list entries (coalesced loads; 0 beyond the batch).
template <int kRounds>
__device__ __forceinline__ void load_entries(uint2 (&ent)[kRounds], const uint2* src, unsigned int m, int lane) {
#pragma unroll
    for (int rr = 0; rr < kRounds; ++rr) {
        const unsigned int e = 32u * rr + lane;
        ent[rr] = e < m ? __ldg(src + e) : make_uint2(0u, 0u);
    }
}

// Pops one work item for the calling warp: (chunk-local tile index, pixel block) or 0xffffffff when the queue is empty.
__device__ __forceinline__ unsigned int pop_item(unsigned int* cursor, unsigned int n_items, int lane) {
    unsigned int w = 0;
    if (lane == 0) w = atomicAdd(cursor, 1u);
    w = __shfl_sync(kFull, w, 0);
    return w < n_items ? w : 0xffffffffu;
}


// ------------------------------------------------------------------------------------------------ forward
struct FwdArgs {
    RenderGeom g;
    int render_base;
    const unsigned int* blk_off;  // [R*T*8] block lists (sgr_binning.cu step 5)
    const unsigned int* blk_cnt;
    unsigned int* blk_eff;        // out: records of the block list the backward has to replay
    const float4 *g0, *g1, *g2;   // [Rc*N] the chunk's per-(render, Gaussian) records
    uint2* bidx;                  // block-list entries; the forward refines their quarter masks in place
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
    const unsigned int* dense_items;   // block items of the dense lists (queued by the tile sort)
    WorkCounts* wc;
    int clamp_color;
    // optional fused loss (include/sgr.h SgrForwardArgs::loss_*)
    const float *loss_target, *loss_mask;
    float* loss_dL_dcolor;
    float* loss_part;
    float loss_scale;
};

// Backward item classes by the number of records of the segment that blended in at least one quarter.
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

// Output epilogue of one pixel (all 32 lanes of the block call it; lane = pixel (x = lane & 7, y = lane >> 3)).
// clamp_color: gs.py:107 `clamp(0, 1)` fused, with the clamp mask kept for the backward (bit c = channel c saturated,
// torch's rule: the gradient passes where 0 <= c <= 1).  out_feed: the LPIPS input of whole_loss.py:132-136 — a
// factor-2 bilinear resize (align_corners=False) is the mean over 2x2 pixels, here two shuffles inside the block.
__device__ __forceinline__ void write_pixel(const FwdArgs& a, int r, size_t P, int px, int py, bool inside, int lane,
                                            float c0, float c1, float c2, float D, float Wt, unsigned int last) {
    (void)lane;
    const size_t pix = size_t(py) * a.g.W + px;
    if (a.clamp_color) {
        const unsigned int m = (c0 >= 0.0f && c0 <= 1.0f ? 0u : 1u) | (c1 >= 0.0f && c1 <= 1.0f ? 0u : 2u) |
                               (c2 >= 0.0f && c2 <= 1.0f ? 0u : 4u);
        c0 = fminf(fmaxf(c0, 0.0f), 1.0f); c1 = fminf(fmaxf(c1, 0.0f), 1.0f); c2 = fminf(fmaxf(c2, 0.0f), 1.0f);
        if (inside) a.clamp_mask[size_t(r) * P + pix] = static_cast<unsigned char>(m);
    }
    if (a.out_feed) {                           // even H, W: a 2x2 cell is entirely inside or outside the image
        float s0 = c0 + __shfl_xor_sync(kFull, c0, 1), s1 = c1 + __shfl_xor_sync(kFull, c1, 1),
              s2 = c2 + __shfl_xor_sync(kFull, c2, 1);
        s0 += __shfl_xor_sync(kFull, s0, 8); s1 += __shfl_xor_sync(kFull, s1, 8); s2 += __shfl_xor_sync(kFull, s2, 8);
        if (inside && (px & 1) == 0 && (py & 1) == 0) {
            const size_t Pq = P >> 2;
            const size_t q = size_t(py >> 1) * (a.g.W >> 1) + (px >> 1);
            float* of = a.out_feed + size_t(r) * 3 * Pq;
            of[q] = 0.5f * s0 - 1.0f; of[Pq + q] = 0.5f * s1 - 1.0f; of[2 * Pq + q] = 0.5f * s2 - 1.0f;
        }
    }
    if (inside) {
        float* oc = a.out_color + size_t(r) * 3 * P;
        oc[pix] = c0; oc[P + pix] = c1; oc[2 * P + pix] = c2;
        a.out_depth[size_t(r) * P + pix] = D;
        a.out_alpha[size_t(r) * P + pix] = Wt;
        a.n_contrib[size_t(r) * P + pix] = last;
    }
}

using FwdSmem = FwdWarpSmem<kFwdStages, kFwdBatch>;

// Experiment build (-DSGR_FWD_TRACE, tools/fwd_trace.py): per work item of the forward blend (begin, end, entries, SM).
#ifdef SGR_FWD_TRACE
constexpr unsigned int kFwdTraceCap = 1u << 17;
__device__ unsigned long long g_fwd_trace[kFwdTraceCap][4];
#endif
using BwdSmem = BwdWarpSmem<kBwdSlots, kBwdBatch>;

#ifndef SGR_FWD_GROUP8
#define SGR_FWD_GROUP8 1
#endif
#ifndef SGR_FWD_MIN_CTAS
#define SGR_FWD_MIN_CTAS 1
#endif
#ifndef SGR_BWD_MIN_CTAS
#define SGR_BWD_MIN_CTAS 2
#endif
template <bool kExact>
__global__ void __launch_bounds__(kFwdThreads, SGR_FWD_MIN_CTAS) blend_forward_kernel(FwdArgs a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ int s_dense_busy;                  // warps of this CTA that still walk dense lists
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const unsigned int n_dense = a.wc->n_dense;
    // at most a third of the CTAs turn dense (a launch whose lists are mostly dense is throughput-bound, not tail-
    // bound); dense lists beyond their capacity are taken by whoever runs out of ordinary items
    const bool dense_cta = blockIdx.x * unsigned(kDenseWarps) < n_dense && blockIdx.x * 3u < gridDim.x;
    bool dense_mode = dense_cta && warp < kDenseWarps;
    if (threadIdx.x == 0) s_dense_busy = dense_cta ? kDenseWarps : 0;
    __syncthreads();
    FwdSmem& sm = reinterpret_cast<FwdSmem*>(smem_raw)[warp];
    const size_t P = size_t(a.g.H) * a.g.W;
    const float bg0 = a.bg[0], bg1 = a.bg[1], bg2 = a.bg[2];
    const unsigned int n_items = a.wc->n_blend * kBlocksPerTile, n_empty_items = a.wc->n_empty * kBlocksPerTile;
    const int qsel = ((lane >> 2) & 1) | ((lane >> 3) & 2);        // 4x2 pixel quarter of this lane: x half | 2 * y half
    const unsigned int qmask = 0x00000f0fu << ((lane & 4) | (lane & 16));   // the quarter's 8 lanes
    const unsigned char* mylist = sm.list[qsel];
    const bool refine = a.refine_masks != 0;
    unsigned char* hit_bytes = reinterpret_cast<unsigned char*>(sm.hit) + qsel;       // indexed with 4 * j
#pragma unroll
    for (int w = 0; w < kFwdBatch; w += 32) sm.hit[w + lane] = 0u;
    if (lane < kFwdStages) {      // the sentinel record: threshold +inf -> never valid, alpha 0, colour 0
        sm.r0[lane][kFwdBatch] = make_float4(0.0f, 0.0f, 0.0f, __int_as_float(0x7f800000));
        sm.r1[lane][kFwdBatch] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        sm.r2[lane][kFwdBatch] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    }
    __syncwarp();

    // ---------------- tiles without instances: background only.  Whole tiles are popped; the fused loss's loads of
    // four blocks are issued together (the phase is a chain of memory latencies, not work).  Called by every warp
    // after the main queue, and before it by the warps that would otherwise sleep beside a dense walk.
    const bool vec_ok = (a.g.W & 3) == 0 &&
        ((reinterpret_cast<uintptr_t>(a.out_color) | reinterpret_cast<uintptr_t>(a.out_depth) |
          reinterpret_cast<uintptr_t>(a.out_alpha) | reinterpret_cast<uintptr_t>(a.n_contrib) |
          reinterpret_cast<uintptr_t>(a.clamp_mask) | reinterpret_cast<uintptr_t>(a.loss_target) |
          reinterpret_cast<uintptr_t>(a.loss_mask) | reinterpret_cast<uintptr_t>(a.loss_dL_dcolor) |
          reinterpret_cast<uintptr_t>(a.out_feed)) & 15) == 0;       // (null pointers count as aligned)
    auto empty_tiles = [&]() {
        for (;;) {
            unsigned int w = 0;
            if (lane == 0) w = atomicAdd(&a.wc->empty_cursor, 1u);
            w = __shfl_sync(kFull, w, 0);
            if (w * kBlocksPerTile >= n_empty_items) break;
            const unsigned int tile_local = a.work_empty[w];
            const int rl = tile_local / a.g.num_tiles;
            const int tile = tile_local - rl * a.g.num_tiles;
            const int r = a.render_base + rl;
            const int tx = tile % a.g.tiles_x, ty = tile / a.g.tiles_x;
            const size_t rP = size_t(r) * P;
            if (vec_ok && (tx + 1) * kTile <= a.g.W && (ty + 1) * kTile <= a.g.H) {
                // A tile entirely inside the image, rows 16-byte aligned: lane = (row lane >> 2 of a half tile, four
                // pixels (lane & 3) * 4 ..), everything as 16-byte accesses, the loads of both half tiles first.
                const int x4 = tx * kTile + (lane & 3) * 4;
                float4 tg[2][3], mk[2];
                size_t pix[2];
#pragma unroll
                for (int hf = 0; hf < 2; ++hf) {
                    pix[hf] = size_t(ty * kTile + hf * 8 + (lane >> 2)) * a.g.W + x4;
                    if (a.loss_target) {
                        mk[hf] = a.loss_mask ? *reinterpret_cast<const float4*>(a.loss_mask + rP + pix[hf])
                                             : make_float4(1.0f, 1.0f, 1.0f, 1.0f);
#pragma unroll
                        for (int ch = 0; ch < 3; ++ch)
                            tg[hf][ch] = *reinterpret_cast<const float4*>(a.loss_target + 3 * rP + ch * P + pix[hf]);
                    }
                }
                const float cs[3] = {bg0, bg1, bg2};
                float cl[3];
                unsigned int cmask = 0;
#pragma unroll
                for (int ch = 0; ch < 3; ++ch) {
                    const bool pass = cs[ch] >= 0.0f && cs[ch] <= 1.0f;
                    cmask |= pass ? 0u : (1u << ch);
                    cl[ch] = a.clamp_color ? fminf(fmaxf(cs[ch], 0.0f), 1.0f) : cs[ch];
                }
                float part = 0.0f;
#pragma unroll
                for (int hf = 0; hf < 2; ++hf) {
                    float* oc = a.out_color + 3 * rP + pix[hf];
#pragma unroll
                    for (int ch = 0; ch < 3; ++ch)
                        *reinterpret_cast<float4*>(oc + ch * P) = make_float4(cl[ch], cl[ch], cl[ch], cl[ch]);
                    *reinterpret_cast<float4*>(a.out_depth + rP + pix[hf]) = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
                    *reinterpret_cast<float4*>(a.out_alpha + rP + pix[hf]) = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
                    *reinterpret_cast<uint4*>(a.n_contrib + rP + pix[hf]) = make_uint4(0u, 0u, 0u, 0u);
                    if (a.clamp_color) *reinterpret_cast<unsigned int*>(a.clamp_mask + rP + pix[hf]) = cmask * 0x01010101u;
                    if (a.out_feed && ((lane >> 2) & 1) == 0) {      // 2x2 means of a constant image: even rows write
                        const size_t Pq = P >> 2;
                        const size_t q = size_t((ty * kTile + hf * 8 + (lane >> 2)) >> 1) * (a.g.W >> 1) + (x4 >> 1);
                        float* of = a.out_feed + size_t(r) * 3 * Pq + q;
#pragma unroll
                        for (int ch = 0; ch < 3; ++ch) {
                            const float v = 0.5f * (4.0f * cl[ch]) - 1.0f;
                            *reinterpret_cast<float2*>(of + ch * Pq) = make_float2(v, v);
                        }
                    }
                    if (a.loss_target) {
                        const float mks[4] = {mk[hf].x, mk[hf].y, mk[hf].z, mk[hf].w};
#pragma unroll
                        for (int ch = 0; ch < 3; ++ch) {          // loss_pixel() on the background colour
                            const float tgs[4] = {tg[hf][ch].x, tg[hf][ch].y, tg[hf][ch].z, tg[hf][ch].w};
                            const bool pass = cs[ch] >= 0.0f && cs[ch] <= 1.0f;
                            float dl[4];
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                const float c1 = fminf(fmaxf(cs[ch], 0.0f), 1.0f);
                                const float diff = c1 * mks[k] - tgs[k] * mks[k];
                                const float gs = mks[k] * a.loss_scale;
                                part += fabsf(diff);
                                dl[k] = pass ? (diff > 0.0f ? gs : diff < 0.0f ? -gs : 0.0f) : 0.0f;
                            }
                            *reinterpret_cast<float4*>(a.loss_dL_dcolor + 3 * rP + ch * P + pix[hf]) =
                                make_float4(dl[0], dl[1], dl[2], dl[3]);
                        }
                    }
                }
                if (a.loss_target) {                  // one partial for the tile (slot of block 0), zeros for the others
                    part = warp_sum(part);
                    if (lane < kBlocksPerTile) a.loss_part[size_t(tile_local) * kBlocksPerTile + lane] = lane == 0 ? part : 0.0f;
                }
                continue;
            }
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                float tg[4][3], mk[4];
                int pxs[4], pys[4];
                bool in[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int blk = half * 4 + k;
                    pxs[k] = tx * kTile + (blk & 1) * kBlockW + (lane & 7);
                    pys[k] = ty * kTile + (blk >> 1) * kBlockH + (lane >> 3);
                    in[k] = pxs[k] < a.g.W && pys[k] < a.g.H;
                    if (a.loss_target && in[k]) {
                        const size_t pix = size_t(pys[k]) * a.g.W + pxs[k];
                        mk[k] = a.loss_mask ? a.loss_mask[rP + pix] : 1.0f;
#pragma unroll
                        for (int ch = 0; ch < 3; ++ch) tg[k][ch] = a.loss_target[3 * rP + ch * P + pix];
                    }
                }
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int blk = half * 4 + k;
                    if (a.loss_target) {
                        float part = 0.0f;
                        if (in[k]) {
                            const size_t pix = size_t(pys[k]) * a.g.W + pxs[k];
                            const float gs = mk[k] * a.loss_scale;
                            const float cs[3] = {bg0, bg1, bg2};
#pragma unroll
                            for (int ch = 0; ch < 3; ++ch) {    // loss_pixel() on the background colour
                                const float c = cs[ch];
                                const float cl = fminf(fmaxf(c, 0.0f), 1.0f);
                                const float diff = cl * mk[k] - tg[k][ch] * mk[k];
                                part += fabsf(diff);
                                const bool pass = c >= 0.0f && c <= 1.0f;
                                a.loss_dL_dcolor[3 * rP + ch * P + pix] = pass ? (diff > 0.0f ? gs : diff < 0.0f ? -gs : 0.0f) : 0.0f;
                            }
                        }
                        part = warp_sum(part);
                        if (lane == 0) a.loss_part[size_t(tile_local) * kBlocksPerTile + blk] = part;
                    }
                    write_pixel(a, r, P, pxs[k], pys[k], in[k], lane, bg0, bg1, bg2, 0.0f, 0.0f, 0u);
                }
            }
        }
    };

    // ---------------- blocks of tiles with instances
    bool main_done = false;
    for (;;) {
        unsigned int bitem;                       // chunk-local tile * 8 + block
        bool take_dense = dense_mode;
        if (dense_mode) {
            const unsigned int d = pop_item(&a.wc->dense_cursor, n_dense, lane);
            if (d == 0xffffffffu) {               // no dense list left: wake the CTA's other warps, join the main queue
                dense_mode = false;
                if (lane == 0) atomicSub(&s_dense_busy, 1);
                continue;
            }
            bitem = a.dense_items[d];
        } else {
            int wait = 0;                         // (warp-uniform: read by one lane)
            if (dense_cta && warp >= kDenseWarps + kDenseMates)
                wait = __shfl_sync(kFull, *reinterpret_cast<volatile int*>(&s_dense_busy), 0);
            if (wait > 0) {
                // this CTA's schedulers belong to its dense walks for now
                if (kDenseWaitersDoEmpty) empty_tiles();
                if (lane == 0)
                    while (*reinterpret_cast<volatile int*>(&s_dense_busy) > 0) __nanosleep(1000);
                __syncwarp();
            }
            const unsigned int item = main_done ? 0xffffffffu : pop_item(&a.wc->blend_cursor, n_items, lane);
            if (item == 0xffffffffu) {            // ordinary items are gone: help with the dense lists, if any are left
                main_done = true;
                const unsigned int d = pop_item(&a.wc->dense_cursor, n_dense, lane);
                if (d == 0xffffffffu) break;
                bitem = a.dense_items[d];
                take_dense = true;
            } else {
                bitem = a.work_blend[item / kBlocksPerTile] * kBlocksPerTile + item % kBlocksPerTile;
            }
        }
        const unsigned int tile_local = bitem / kBlocksPerTile;
        const int blk = bitem % kBlocksPerTile;
        const int rl = tile_local / a.g.num_tiles;
        const int tile = tile_local - rl * a.g.num_tiles;
        const int r = a.render_base + rl;
        const size_t tg = size_t(r) * a.g.num_tiles + tile;
        const size_t bi = tg * kBlocksPerTile + blk;
        const unsigned int n = a.blk_cnt[bi];                       // records of this block's list
        if (!take_dense && n >= kDenseEntries) continue;            // queued as a dense list (same test as the tile sort)
        const size_t off = a.blk_off[bi];
        const unsigned int nb = (n + kFwdBatch - 1) / kFwdBatch;
        const int tx = tile % a.g.tiles_x, ty = tile / a.g.tiles_x;
        const int bx0 = tx * kTile + (blk & 1) * kBlockW, by0 = ty * kTile + (blk >> 1) * kBlockH;
        if (bx0 >= a.g.W || by0 >= a.g.H) {                        // block entirely outside the image
            if (lane == 0) {
                if (a.loss_target) a.loss_part[size_t(tile_local) * kBlocksPerTile + blk] = 0.0f;
                if (refine) a.blk_eff[bi] = 0u;
            }
            continue;
        }
        const unsigned long long t_begin = global_timer_ns();
        const int px = bx0 + (lane & 7), py = by0 + (lane >> 3);
        const bool inside = px < a.g.W && py < a.g.H;
        const float pxf = float(px), pyf = float(py);
        const size_t goff = size_t(rl) * a.g.N;
        const float4 *g0 = a.g0 + goff, *g1 = a.g1 + goff, *g2 = a.g2 + goff;
        const uint2* lst = a.bidx + off;
        if (a.loss_target && inside) {          // the fused loss reads its target at the END of the item: fetch it to L2 now
            const size_t pix = size_t(py) * a.g.W + px;
            const float* tgp = a.loss_target + 3 * size_t(r) * P + pix;
            asm volatile("prefetch.global.L2 [%0];" ::"l"(tgp));
            asm volatile("prefetch.global.L2 [%0];" ::"l"(tgp + P));
            asm volatile("prefetch.global.L2 [%0];" ::"l"(tgp + 2 * P));
            if (a.loss_mask) asm volatile("prefetch.global.L2 [%0];" ::"l"(a.loss_mask + size_t(r) * P + pix));
        }

        float T = 1.0f, C0 = 0.0f, C1 = 0.0f, C2 = 0.0f, D = 0.0f, Wt = 0.0f;
        unsigned int last = 0;
        bool done = !inside;
        unsigned int seg_hits = 0, eff = 0;      // records of the current segment that blended; last such record + 1
        // Pipeline per batch b: this lane's block-list entries are loaded two batches ahead (registers), the records
        // they point at are gathered one batch ahead (while batch b - 1 is walked), and the batch is walked once its
        // gather group has completed.
        constexpr int kR = kFwdBatch / 32;
        uint2 ent0[kR], ent1[kR], ent2[kR];          // entries of batch b, b + 1, b + 2
        load_entries(ent0, lst, min(unsigned(kFwdBatch), n), lane);
        load_entries(ent1, lst + kFwdBatch, nb > 1 ? min(unsigned(kFwdBatch), n - kFwdBatch) : 0u, lane);
        if (nb) gather_batch(sm, 0, ent0, min(unsigned(kFwdBatch), n), g0, g1, g2, lane);
        for (unsigned int b = 0; b < nb; ++b) {
            load_entries(ent2, lst + (b + 2) * kFwdBatch, b + 2 < nb ? min(unsigned(kFwdBatch), n - (b + 2) * kFwdBatch) : 0u, lane);
            if (b + 1 < nb) {
                gather_batch(sm, (b + 1) % kFwdStages, ent1, min(unsigned(kFwdBatch), n - (b + 1) * kFwdBatch),
                             g0, g1, g2, lane);
                cp_async_wait<1>();              // everything but the group just committed: batch b has landed
            } else {
                cp_async_wait<0>();
            }
            const int s = b % kFwdStages;
            const unsigned int m = min(unsigned(kFwdBatch), n - b * kFwdBatch);
            const unsigned int cbase = b * kFwdBatch;
            patch_positions(sm, s, ent0, m, lane);       // (each lane patches the records it gathered itself)
            __syncwarp();
            const float4* r0 = sm.r0[s];
            const float4* r1 = sm.r1[s];
            const float4* r2 = sm.r2[s];
            uint4 cnt;
            unsigned int qbits[kFwdBatch / 32];
            cnt = cull_batch<false, kFwdBatch>(ent0, m, sm.list, lane, qbits);
            // quarters whose 8 pixels are all finished need no further evaluation
            const unsigned int dmask = __ballot_sync(kFull, done);
            const unsigned int my_n = ((dmask & qmask) == qmask) ? 0u : (qsel == 0 ? cnt.x : qsel == 1 ? cnt.y : qsel == 2 ? cnt.z : cnt.w);
            const int total = int(__reduce_max_sync(kFull, my_n));
            // One trip = every quarter evaluates its next survivor (lane = pixel).  Four trips per iteration, written
            // load-first so that the four independent alpha chains interleave ahead of the sequential compositing
            // recurrence.  List entries past a quarter's count point at the sentinel record (alpha = 0): no bounds
            // checks in the loop.  Quarters that are finished still walk (ok = false for all their lanes).
            unsigned int lastj = 0xffffffffu;
            // A group of G trips: G independent alpha chains (loads first), then the compositing recurrence with the
            // colour records fetched as they are needed.  Groups of 8 while at least 5 trips remain (a single warp's
            // issue rate is bounded by the dependent-issue latency, so the long lists that run alone at the end of the
            // launch need the extra instruction-level parallelism), a group of 4 for the rest.
            auto trip_group = [&](int t0, auto G_tag) {
                constexpr int G = decltype(G_tag)::value;
                float4 q0[G], q1[G];
                unsigned int jj[G];
#pragma unroll
                for (int w4 = 0; w4 < G / 4; ++w4) {
                    const unsigned int packed = *reinterpret_cast<const unsigned int*>(mylist + t0 + 4 * w4);
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        jj[4 * w4 + u] = (packed >> (8 * u)) & 0xffu;
                        q0[4 * w4 + u] = r0[jj[4 * w4 + u]];
                        q1[4 * w4 + u] = r1[jj[4 * w4 + u]];
                    }
                }
                float al[G];
#pragma unroll
                for (int u = 0; u < G; ++u) {
                    const float dx = q0[u].x - pxf, dy = q0[u].y - pyf;
                    const float power = gauss_power(q1[u].x, q1[u].y, q1[u].z, dx, dy);
                    // exact mode: exp_core needs power >= the record's threshold (>= -87; below it alpha < 1/255
                    // anyway); the SFU exponential takes any argument
                    const bool valid = kExact ? (!(power > 0.0f) && !(power < q0[u].w)) : !(power > 0.0f);
                    const float alpha = fminf(kAlphaMax, q1[u].w * blend_exp<kExact>(kExact ? (valid ? power : 0.0f) : power));
                    al[u] = valid ? alpha : 0.0f;
                }
#pragma unroll
                for (int u = 0; u < G; ++u) {
                    const float4 q2 = r2[jj[u]];
                    const bool ok = !done && !(al[u] < kAlphaMin);
                    const float test_T = __fmaf_rn(-al[u], T, T);
                    const bool stop = ok && (test_T < kTMin);
                    const bool blend = ok && !stop;
                    const float w = blend ? al[u] * T : 0.0f;      // adding c * 0 leaves the sums bit-unchanged
                    C0 = __fmaf_rn(q2.x, w, C0);
                    C1 = __fmaf_rn(q2.y, w, C1);
                    C2 = __fmaf_rn(q2.z, w, C2);
                    if (kExact) Wt += w;                           // default mode: alpha = 1 - T at the end
                    D = __fmaf_rn(q2.w, w, D);
                    T = blend ? test_T : T;
                    lastj = blend ? jj[u] : lastj;
                    done = done || stop;
                    // same-value stores of the quarter's lanes to one byte: benign
                    if (refine && blend) hit_bytes[4u * jj[u]] = 1;
                }
            };
            {
                int t0 = 0;
#if SGR_FWD_GROUP8
                for (; total - t0 > 4; t0 += 8) trip_group(t0, std::integral_constant<int, 8>{});
#endif
                for (; t0 < total; t0 += 4) trip_group(t0, std::integral_constant<int, 4>{});
            }
            // n_contrib counts positions in the TILE's list (upstream's contributor index): word 2 of rec0 (patched)
            if (lastj != 0xffffffffu) last = __float_as_uint(r0[lastj].z) + 1u;
            __syncwarp();
            // Mask refinement for the backward pass: clear the (block, quarter) bits of the records that no pixel of
            // the quarter blended (alpha test, finished pixels) — the backward walk culls on the same word.
            if (refine) {
                if ((cnt.x | cnt.y | cnt.z | cnt.w) != 0u) {
                    unsigned int hw[kFwdBatch / 32];
#pragma unroll
                    for (int rr = 0; rr < kFwdBatch / 32; ++rr) hw[rr] = sm.hit[32 * rr + lane];
#pragma unroll
                    for (int rr = 0; rr < kFwdBatch / 32; ++rr) {
                        const unsigned int e = 32u * rr + lane;
                        const unsigned int exact = ((hw[rr] & 0x01010101u) * 0x10204080u) >> 28;
                        const unsigned int clear = qbits[rr] & ~exact;      // qbits = 0 beyond the batch
                        // the block owns its records: a plain store of the refined word
                        if (clear) a.bidx[off + cbase + e].y = ent0[rr].y & ~clear;
                        if (hw[rr]) sm.hit[e] = 0u;
                        const unsigned int hb = __ballot_sync(kFull, exact != 0u);
                        if (hb) { seg_hits += __popc(hb); eff = cbase + 32u * rr + (32u - __clz(hb)); }
                    }
                    __syncwarp();
                }
                // end of a backward segment (or of the list): one work item if anything blended in it
                const unsigned int end = cbase + m;
                if (end % kSegB == 0u || end == n) {
                    if (seg_hits != 0u && lane == 0) {
                        const int cls = bwd_class(seg_hits);
                        const unsigned int slot = atomicAdd(&a.plan->n_items[cls], 1u);
                        a.bwd_items[size_t(cls) * a.bwd_items_stride + a.plan->item_base + slot] =
                            make_uint2(unsigned(tile_local) * kBlocksPerTile + blk, (end - 1u) / kSegB);
                    }
                    seg_hits = 0u;
                }
            }
            // lists longer than one backward segment: checkpoint the running state at every segment boundary
            if (n > unsigned(kSegB) && ((b + 1) * kFwdBatch) % kSegB == 0 && (b + 1) * kFwdBatch < n) {
                const size_t ci = (off / (kSegB / 2) + (b + 1) * kFwdBatch / kSegB - 1) * 32 + lane;
                a.ck0[ci] = make_float4(T, C0, C1, C2);
                a.ck1[ci] = D;
            }
            if (__all_sync(kFull, done)) break;
#pragma unroll
            for (int rr = 0; rr < kR; ++rr) { ent0[rr] = ent1[rr]; ent1[rr] = ent2[rr]; }
        }
        if (n > unsigned(kSegB)) {                    // final state, read by the backward's non-final segments
            const size_t ci = (off / (kSegB / 2) + (n + kSegB - 1) / kSegB - 1) * 32 + lane;
            a.ck0[ci] = make_float4(T, C0, C1, C2);
            a.ck1[ci] = D;
        }
        if (refine) {
            // an early exit (all pixels finished) inside a segment leaves its item unpushed: push it now
            if (seg_hits != 0u && lane == 0) {
                const int cls = bwd_class(seg_hits);
                const unsigned int slot = atomicAdd(&a.plan->n_items[cls], 1u);
                a.bwd_items[size_t(cls) * a.bwd_items_stride + a.plan->item_base + slot] =
                    make_uint2(unsigned(tile_local) * kBlocksPerTile + blk, (eff - 1u) / kSegB);
            }
            if (lane == 0) a.blk_eff[bi] = eff;
        }
        cp_async_wait<0>();                      // early exit: a gather in flight must land before its stage is reused
        __syncwarp();
        if (a.loss_target) {
            float part = 0.0f;
            if (inside) part = loss_pixel(a, size_t(r) * P, P, size_t(py) * a.g.W + px, C0 + T * bg0, C1 + T * bg1, C2 + T * bg2);
            part = warp_sum(part);
            if (lane == 0) a.loss_part[size_t(tile_local) * kBlocksPerTile + blk] = part;
        }
        if (!kExact) Wt = 1.0f - T;                  // sum of the blend weights up to rounding (the oracle sums them)
        write_pixel(a, r, P, px, py, inside, lane, C0 + T * bg0, C1 + T * bg1, C2 + T * bg2, D, Wt, last);
#ifdef SGR_FWD_TRACE
        if (lane == 0 && bitem < kFwdTraceCap) {
            unsigned int smid;
            asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
            g_fwd_trace[bitem][0] = t_begin; g_fwd_trace[bitem][1] = global_timer_ns();
            g_fwd_trace[bitem][2] = n; g_fwd_trace[bitem][3] = (smid << 8) | unsigned(warp);
        }
#endif
        if (a.tile_time && lane == 0) {          // diagnostics: start of block 0, duration of the slowest block
            const unsigned long long now = global_timer_ns();
            if (blk == 0) a.tile_time[tg].x = (unsigned int)t_begin;
            atomicMax(&a.tile_time[tg].y, (unsigned int)(now - t_begin));
        }
    }

    empty_tiles();
}

// ------------------------------------------------------------------------------------------------ backward
struct BwdArgs {
    RenderGeom g;
    int render_base;
    const unsigned int* blk_off;
    const unsigned int* blk_cnt;
    const unsigned int* blk_eff;
    const uint2* bidx;                    // block-list entries (quarter masks refined by the forward)
    const float4 *g0, *g1, *g2;           // [Rc*N] the chunk's per-(render, Gaussian) records
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

// Work item = (block list, segment): the block replays the list entries [lo, hi) of its segment back to front,
// lo = segment * kSegB, hi = min(lo + kSegB, blk_eff) (blk_eff = the last record any pixel blended, + 1).
// A pixel whose contributors end inside the segment starts from its final state (T = 1 - alpha_out as in A.5, nothing
// behind); a pixel that continues behind the segment starts from the forward's checkpoints: T = T_ck and
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
        for (int s = 0; s < kIxStages; ++s) mbar_init(&sm.ixbar[s], 1);
        fence_barrier_init();
    }
    __syncwarp();
    const int qsel = ((lane >> 2) & 1) | ((lane >> 3) & 2);        // 4x2 pixel quarter of this lane: x half | 2 * y half
    if (lane < kBwdStages) {      // the sentinel record: threshold +inf -> never valid
        sm.r0[lane][kBwdBatch] = make_float4(0.0f, 0.0f, 0.0f, __int_as_float(0x7f800000));
        sm.r1[lane][kBwdBatch] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        sm.r2[lane][kBwdBatch] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    }
    __syncwarp();
    float (*stA)[32] = sm.stash[0];             // dL/dalpha * G (every geometric term carries this product)
    float (*stW)[32] = sm.stash[1];             // blend weight alpha * T (0 = the pixel did not blend this Gaussian)
    const unsigned char* mylist = sm.list[qsel];
    unsigned int ix_issued = 0, ix_waited = 0;   // index ring counters over the kernel's lifetime (warp-uniform)
    // Stash rows are XOR-swizzled by trip, column = lane ^ swz(t): conflict-free both for lane = pixel (phases A, B)
    // and for lane = (trip, quarter) pairs reading one pixel of their quarter (phase C).

    for (;;) {
        const unsigned int item = pop_item(&a.plan->cursor, n_items, lane);
        if (item == 0xffffffffu) break;
        int cls = 0;
#pragma unroll
        for (int c = 0; c < kBwdClasses - 1; ++c) cls += item >= cls_end[c] ? 1 : 0;
        const unsigned int in_cls = item - (cls ? cls_end[cls - 1] : 0u);
        const uint2 ws = a.bwd_items[size_t(cls) * a.bwd_items_stride + item_base + in_cls];
        const unsigned int tile_local = ws.x / kBlocksPerTile;
        const int blk = ws.x % kBlocksPerTile;
        const unsigned int lo = ws.y * unsigned(kSegB);
        const int rl = tile_local / a.g.num_tiles;
        const int tile = tile_local - rl * a.g.num_tiles;
        const int r = a.render_base + rl;
        const size_t tg = size_t(r) * a.g.num_tiles + tile;
        const size_t bi = tg * kBlocksPerTile + blk;
        const int tx = tile % a.g.tiles_x, ty = tile / a.g.tiles_x;
        const int bx0 = tx * kTile + (blk & 1) * kBlockW, by0 = ty * kTile + (blk >> 1) * kBlockH;
        const int px = bx0 + (lane & 7), py = by0 + (lane >> 3);
        const bool inside = px < a.g.W && py < a.g.H;
        const size_t pix = size_t(py) * a.g.W + px;
        const unsigned int last = inside ? a.n_contrib[size_t(r) * P + pix] : 0u;     // position in the TILE list + 1
        const unsigned int eff = a.blk_eff[bi];
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
        sm.dpix[0][lane] = dp0; sm.dpix[1][lane] = dp1; sm.dpix[2][lane] = dp2; sm.dpix[3][lane] = ddep;
        __syncwarp();
        const unsigned int nb = (hi - lo + kBwdBatch - 1) / kBwdBatch;
        const float pxf = float(px), pyf = float(py);
        const float wx0 = float(bx0), wy0 = float(by0);
        const size_t goff = size_t(rl) * a.g.N;
        const float4 *g0 = a.g0 + goff, *g1 = a.g1 + goff, *g2 = a.g2 + goff;
        const uint2* lst = a.bidx + off + lo;
        float T = T_final;
        float U = 0.0f;
        // A pixel has contributors behind this segment iff its last contributor sits at or behind the first record
        // after the segment (records are in tile-list order; word 2 >> 4 = position in the tile list).
        bool resume = false;
        if (lo + unsigned(kSegB) < n) {
            const unsigned int next_pos = __ldg(a.bidx + off + lo + kSegB).y >> 4;
            resume = last > next_pos;
        }
        if (resume) {                            // resume from the checkpoints
            const size_t slot0 = off / (kSegB / 2);
            const size_t ci = (slot0 + ws.y) * 32 + lane;
            const size_t fi = (slot0 + (n + kSegB - 1) / kSegB - 1) * 32 + lane;
            const float4 ck = a.ck0[ci], fin = a.ck0[fi];
            // T = the forward's running transmittance at the segment end, U = what the records behind contribute
            T = ck.x;
            U = (fin.y - ck.y) * dp0;
            U = fmaf(fin.z - ck.z, dp1, U);
            U = fmaf(fin.w - ck.w, dp2, U);
            if (kDepthAlphaGrads) U = fmaf(a.ck1[fi] - a.ck1[ci], ddep, U);
        }
        const float bg_dot = (bg0 * dp0 + bg1 * dp1) + bg2 * dp2;
        const float K = T_final * (bg_dot - dalp);
        float* acc = a.accum + size_t(rl) * a.g.N * kAccumStride;

        // walk step k handles list batch (nb - 1 - k): its index entries arrive by TMA one step ahead; its records
        // (and Gaussian ids) are gathered at the start of the step — only those the forward found to blend
        const unsigned int ixk0 = ix_issued;
        if (nb) {
            const unsigned int lb = nb - 1;
            ix_issue(sm, ix_issued, lst + lb * kBwdBatch, min(unsigned(kBwdBatch), hi - lo - lb * kBwdBatch), lane);
            ++ix_issued;
        }
        for (unsigned int b = 0; b < nb; ++b) {
            if (b + 1 < nb) {
                const unsigned int lb = nb - 2 - b;
                ix_issue(sm, ix_issued, lst + lb * kBwdBatch, unsigned(kBwdBatch), lane);
                ++ix_issued;
            }
            const unsigned int cbase = lo + (nb - 1 - b) * kBwdBatch;     // list index of the batch's first record
            const unsigned int m = min(unsigned(kBwdBatch), hi - cbase);
            ix_wait(sm, ixk0 + b); ++ix_waited;
            const uint2* ixs = sm.ix[(ixk0 + b) % kIxStages];     // stays resident while the batch is walked
            uint2 ent[kBwdBatch / 32];
#pragma unroll
            for (int rr = 0; rr < kBwdBatch / 32; ++rr) {
                const unsigned int e = 32u * rr + lane;
                ent[rr] = e < m ? ixs[e] : make_uint2(0u, 0u);
            }
            gather_batch(sm, 0, ent, m, g0, g1, g2, lane);
            cp_async_wait<0>();
            patch_positions(sm, 0, ent, m, lane);
            __syncwarp();
            const float4* r0 = sm.r0[0];
            const float4* r1 = sm.r1[0];
            const float4* r2 = sm.r2[0];
            unsigned int qbits[kBwdBatch / 32];
            const uint4 cnt = cull_batch<true, kBwdBatch>(ent, m, sm.list, lane, qbits);
            const int total = int(max(max(cnt.x, cnt.y), max(cnt.z, cnt.w)));
            // trip t of the batch handles the quarter's survivor number (my_n - 1 - t): back to front
            for (int base = 0; base < total; base += kBwdSlots) {
                const int trips = min(kBwdSlots, total - base);
                // ---- phases A + B: alpha / G of four trips (independent, load-first), then the sequential per-pixel
                // recurrences -> dL/dalpha and blend weight per trip, stashed for phase C.  The lists are descending
                // (trip t = the quarter's t-th survivor from the back) and sentinel-padded.
                for (int t0 = 0; t0 < trips; t0 += 4) {
                    const unsigned int packed = *reinterpret_cast<const unsigned int*>(mylist + base + t0);
                    float4 q0[4], q1[4], q2[4];
                    bool has[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const unsigned int j = (packed >> (8 * u)) & 0xffu;
                        q0[u] = r0[j];
                        has[u] = __float_as_uint(q0[u].z) < last;
                        q1[u] = r1[j];
                        q2[u] = r2[j];
                    }
                    float al[4], gg[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const float dx = q0[u].x - pxf, dy = q0[u].y - pyf;
                        const float power = gauss_power(q1[u].x, q1[u].y, q1[u].z, dx, dy);
                        const bool valid = has[u] && !(power > 0.0f) && !(power < q0[u].w);
                        gg[u] = blend_exp<kExact>(valid ? power : 0.0f);
                        const float alpha = fminf(kAlphaMax, q1[u].w * gg[u]);
                        al[u] = (valid && !(alpha < kAlphaMin)) ? alpha : 0.0f;
                    }
                    // The per-pixel recurrences are two scalars: with cg_k = c_k . dL/dC (+ z_k dL/dD),
                    //   T_k = T_(k+1) / (1 - alpha_k),   U_k = sum over the records behind k of w_j cg_j,
                    //   dL/dalpha_k = T_k cg_k - (U_k + K) / (1 - alpha_k),   K = T_final (bg . dL/dC - dL/dA).
                    float dl[4], wg[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const float alpha = al[u];
                        const bool on = alpha != 0.0f;
                        const float inv_om = rcp_approx(1.0f - alpha);
                        const float Tn = T * inv_om;
                        float cg = q2[u].x * dp0;
                        cg = fmaf(q2[u].y, dp1, cg);
                        cg = fmaf(q2[u].z, dp2, cg);
                        if (kDepthAlphaGrads) cg = fmaf(q2[u].w, ddep, cg);
                        const float w = alpha * Tn;
                        dl[u] = on ? fmaf(Tn, cg, -(U + K) * inv_om) : 0.0f;
                        wg[u] = on ? w : 0.0f;
                        U = on ? fmaf(w, cg, U) : U;
                        T = on ? Tn : T;
                    }
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int t = t0 + u;
                        const int col = lane ^ ((t & 3) | ((t & 4) << 1));
                        stA[t][col] = dl[u] * gg[u];
                        stW[t][col] = wg[u];
                    }
                }
                __syncwarp();
                // ---- phase C: the roles flip to lane = (trip, quarter) pair, i.e. one Gaussian of one quarter; the
                // lane sums the gradient terms of the quarter's 8 pixels in registers (no shuffles) and issues the
                // atomics.  The pass's pairs are enumerated quarter by quarter and taken 32 at a time.
                {
                    const int c0n = min(max(int(cnt.x) - base, 0), kBwdSlots), c1n = min(max(int(cnt.y) - base, 0), kBwdSlots);
                    const int c2n = min(max(int(cnt.z) - base, 0), kBwdSlots), c3n = min(max(int(cnt.w) - base, 0), kBwdSlots);
                    const int e1 = c0n + c1n, e2 = e1 + c2n, npairs = e2 + c3n;
                    for (int pbase = 0; pbase < npairs; pbase += 32) {
                        const int pi = pbase + lane;
                        const bool cvalid = pi < npairs;
                        const int cq = pi < c0n ? 0 : pi < e1 ? 1 : pi < e2 ? 2 : 3;
                        const int ct = cvalid ? pi - (cq == 0 ? 0 : cq == 1 ? c0n : cq == 2 ? e1 : e2) : 0;
                        const unsigned int j = cvalid ? sm.list[cq][base + ct] : 0u;
                        const float4 q0 = r0[j];
                        const float4 q1 = r1[j];
                        const int cswz = (ct & 3) | ((ct & 4) << 1);
                        const int cpl0 = ((cq & 1) << 2) | ((cq & 2) << 3);           // first pixel lane of quarter cq
                        const float bxq = wx0 + float((cq & 1) * 4), byq = wy0 + float((cq >> 1) * 2);
                        const float hA2 = 2.0f * q1.x, hC2 = 2.0f * q1.z;
                        float s0 = 0, s1 = 0, s2 = 0, s3 = 0, s4 = 0, s5 = 0, s6 = 0, s7 = 0, s8 = 0, s9 = 0;
#pragma unroll
                        for (int p = 0; p < 8; ++p) {
                            const int pl = cpl0 + (p & 3) + ((p >> 2) << 3);       // pixel lane inside the quarter
                            const int col = pl ^ cswz;
                            const float w = cvalid ? stW[ct][col] : 0.0f;
                            const float qv = cvalid ? stA[ct][col] : 0.0f;          // dL/dalpha * G
                            const float dx = q0.x - (bxq + float(p & 3)), dy = q0.y - (byq + float(p >> 2));
                            const float Q = q1.w * qv;                               // dL/dG * G
                            const float qx = Q * dx, qy = Q * dy;
                            // -dx*A - dy*B = 2*dx*hA + dy*nB   (rec1 = (-A/2, -B, -C/2, o))
                            s0 = fmaf(hA2, qx, fmaf(q1.y, qy, s0));
                            s1 = fmaf(hC2, qy, fmaf(q1.y, qx, s1));
                            s2 = fmaf(qx, dx, s2);
                            s3 = fmaf(qx, dy, s3);
                            s4 = fmaf(qy, dy, s4);
                            s5 += qv;
                            s6 = fmaf(w, sm.dpix[0][pl], s6);
                            s7 = fmaf(w, sm.dpix[1][pl], s7);
                            s8 = fmaf(w, sm.dpix[2][pl], s8);
                            if (kDepthAlphaGrads) s9 = fmaf(w, sm.dpix[3][pl], s9);
                        }
                        if (cvalid) {
                            float* g = acc + size_t(ixs[j].x) * kAccumStride;   // the record's Gaussian id -> its row
                            red_add_v4(g, s0 * ddelx_dx, s1 * ddely_dy, -0.5f * s2, -0.5f * s3);
                            red_add_v4(g + 4, -0.5f * s4, s5, s6, s7);
                            if (kDepthAlphaGrads) red_add_v4(g + 8, s8, s9, 0.0f, 0.0f);
                            else atomicAdd(g + 8, s8);
                        }
                    }
                }
                __syncwarp();
            }
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
cudaError_t prepare_kernel(K kernel, size_t smem, int* per_sm, int threads = kBlendThreads) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
    if (e != cudaSuccess) return e;
    int n = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kernel, threads, smem);
    if (e != cudaSuccess) return e;
    *per_sm = n > 0 ? n : 1;
    return cudaSuccess;
}

}  // namespace

cudaError_t launch_blend_forward(const ChunkCtx& c, float* out_color, float* out_depth, float* out_alpha,
                                 float* out_feed) {
    FwdArgs a;
    a.g = c.g; a.render_base = c.render_base; a.blk_off = c.blk_off; a.blk_cnt = c.blk_cnt; a.blk_eff = c.blk_eff;
    a.g0 = c.g0; a.g1 = c.g1; a.g2 = c.g2; a.bidx = c.bidx;
    a.bg = c.p->bg; a.n_contrib = c.n_contrib;
    a.refine_masks = (c.p->flags & SGR_FLAG_FORWARD_ONLY) ? 0 : 1;
    a.tile_time = (c.p->flags & SGR_FLAG_TILE_TIMING) ? c.tile_time : nullptr;
    a.out_color = out_color; a.out_depth = out_depth; a.out_alpha = out_alpha; a.out_feed = out_feed;
    a.clamp_mask = c.clamp_mask;
    a.ck0 = c.ck0; a.ck1 = c.ck1; a.plan = c.plan; a.bwd_items = c.bwd_items; a.bwd_items_stride = c.bwd_items_stride;
    a.work_blend = c.work_blend; a.work_empty = c.work_empty; a.wc = c.work_counts; a.dense_items = c.dense_items;
    a.clamp_color = ((c.p->flags & SGR_FLAG_CLAMP_COLOR) || c.loss_target) ? 1 : 0;
    a.loss_target = c.loss_target; a.loss_mask = c.loss_mask; a.loss_dL_dcolor = c.loss_dL_dcolor;
    a.loss_part = c.loss_part; a.loss_scale = c.loss_scale;
    constexpr size_t smem = sizeof(FwdSmem) * kFwdWarps;
    const bool exact = (c.p->flags & SGR_FLAG_EXACT_EXP) != 0;
    static int per_sm_dev[kMaxDevices][2] = {};
    int& per_sm = per_sm_dev[current_device_slot()][exact ? 1 : 0];
    if (per_sm == 0) {
        cudaError_t e = exact ? prepare_kernel(blend_forward_kernel<true>, smem, &per_sm, kFwdThreads)
                              : prepare_kernel(blend_forward_kernel<false>, smem, &per_sm, kFwdThreads);
        if (e != cudaSuccess) return e;
    }
    const long long items = (long long)c.num_renders * c.g.num_tiles * kBlocksPerTile;
    static int ctas_override = -1;               // experiment hook: SGR_FWD_CTAS_PER_SM
    if (ctas_override < 0) { const char* v = getenv("SGR_FWD_CTAS_PER_SM"); ctas_override = v ? atoi(v) : 0; }
    const int use_per_sm = ctas_override > 0 ? min(ctas_override, per_sm) : per_sm;
    const int grid = int(min((long long)num_sms() * use_per_sm, (items + kFwdWarps - 1) / kFwdWarps));
    if (exact) blend_forward_kernel<true><<<grid, kFwdThreads, smem, c.stream>>>(a);
    else blend_forward_kernel<false><<<grid, kFwdThreads, smem, c.stream>>>(a);
    return cudaGetLastError();
}

#ifdef SGR_FWD_TRACE
extern "C" int sgr_debug_fwd_trace(unsigned long long* out, unsigned int rows) {   // experiment build only
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(out, g_fwd_trace, sizeof(unsigned long long) * 4 * (rows < kFwdTraceCap ? rows : kFwdTraceCap));
    return 0;
}
#endif

cudaError_t launch_loss_reduce(const ChunkCtx& c, float* loss_out) {
    const unsigned int n = unsigned(c.num_renders) * c.g.num_tiles * kBlocksPerTile;
    loss_reduce_kernel<<<1, 1024, 0, c.stream>>>(c.loss_part, n, loss_out, c.render_base == 0 ? 1 : 0, c.loss_scale);
    return cudaGetLastError();
}

namespace {
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

cudaError_t launch_blend_backward(const ChunkCtx& c, const SgrBackwardArgs& b) {
    BwdArgs a;
    a.g = c.g; a.render_base = c.render_base; a.blk_off = c.blk_off; a.blk_cnt = c.blk_cnt; a.blk_eff = c.blk_eff;
    a.bidx = c.bidx;
    a.g0 = c.g0; a.g1 = c.g1; a.g2 = c.g2; a.bg = c.p->bg;
    a.n_contrib = c.n_contrib; a.out_alpha = b.out_alpha; a.dL_dcolor = b.dL_dcolor; a.dL_ddepth = b.dL_ddepth;
    a.dL_dalpha = b.dL_dalpha; a.loss_dL_dcolor = b.loss_dL_dcolor; a.dL_dfeed = b.dL_dlpips_feed;
    a.clamp_mask = ((c.p->flags & SGR_FLAG_CLAMP_COLOR) || b.fused_clamp) ? c.clamp_mask : nullptr;
    a.accum = c.accum; a.plane = size_t(c.num_renders) * c.g.N;
    a.ck0 = c.ck0; a.ck1 = c.ck1; a.bwd_items = c.bwd_items; a.bwd_items_stride = c.bwd_items_stride;
    a.plan = c.plan; a.dL_scale = b.dL_dcolor_scale;
    const long long items = (long long)c.num_renders * c.g.num_tiles * kBlocksPerTile;
    const long long want = (items + kWarpsPerCta - 1) / kWarpsPerCta;
    const bool exact = (c.p->flags & SGR_FLAG_EXACT_EXP) != 0;
    if (b.dL_ddepth || b.dL_dalpha)
        return exact ? launch_bwd_variant<true, true>(a, want, c.stream) : launch_bwd_variant<true, false>(a, want, c.stream);
    return exact ? launch_bwd_variant<false, true>(a, want, c.stream) : launch_bwd_variant<false, false>(a, want, c.stream);
}

}  // namespace sgr

