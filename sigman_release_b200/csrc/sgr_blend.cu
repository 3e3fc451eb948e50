// sgr_blend.cu — the alpha-compositing kernels (forward and backward) of the B200 rasteriser.
//
// Replaces upstream renderCUDA forward/backward (third-party diff_gaussian_rasterization; call site
// /root/reference/core/gaussians/gs.py:99-106).  Design (DESIGN.md "blend"):
//   * persistent CTAs (a multiple of the SM count) pull (render, tile) work items from a device-side queue that is
//     ordered longest-list-first (sgr_binning.cu::worklist_kernel);
//   * the tile's depth-ordered 48-byte records (three float4 streams, contiguous per tile because the per-tile sort
//     gathers them) are staged into a shared-memory ring with 1-D TMA bulk copies (cp.async.bulk) completing on
//     mbarriers, issued by a dedicated producer warp; the eight consumer warps run decoupled from each other;
//   * every consumer warp owns an 8x4 pixel block made of four 4x2 quarters.  It first tests 32 records at a time
//     (lane = record) against each quarter with the conservative alpha >= 1/255 extent computed in the preprocess
//     kernel (four ballots); then every quarter walks only its own survivors (lane = pixel), so up to four different
//     Gaussians are evaluated per trip.  Culled records would have been skipped by the alpha test, so the per-pixel
//     arithmetic, the contributor index and every output bit equal the straightforward kernel's;
//   * the walk is split into phases through a per-warp shared-memory stash so that only the inherently sequential
//     part sits on the dependent chain: phase A evaluates alpha for up to 16 trips (independent iterations, unrolled
//     and interleaved by the compiler), phase B composites front to back reading the stashed alphas;
//   * backward: the same walk in reverse with three phases — A: alpha, B: the per-pixel transmittance / colour
//     recurrences producing dL/dalpha and the blend weight, C: per-Gaussian partial sums reduced over the quarter's 8
//     pixels with a transposing butterfly (lane k ends up with component k) and one atomic per component.
//
// Compiled with --fmad=false: the per-pixel expressions are the oracle's (oracle/sgr_oracle.cpp::blend_forward /
// blend_backward) evaluated in the same order, which makes colour, depth, alpha and n_contrib bit-exact.
#include <cuda_fp16.h>

#include <cstdio>

#include "sgr_common.cuh"

namespace sgr {
namespace {

constexpr int kChunk = 128;                    // records per ring stage
constexpr int kStages = 4;
constexpr int kConsumerWarps = 8;              // 8 warps x (8x4 pixels) = one 16x16 tile
constexpr int kBlendThreads = (kConsumerWarps + 1) * 32;
constexpr int kBlockW = 8, kBlockH = 4;        // pixel block of one warp (four 4x2 quarters walk separate lists)
constexpr int kSlots = 16;                     // trips per phase pass (depth of the per-warp stash)
constexpr unsigned int kFull = 0xffffffffu;

// ------------------------------------------------------------------------------------------------ PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_inval(uint64_t* bar) {
    asm volatile("mbarrier.inval.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
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
__device__ __forceinline__ unsigned long long global_timer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

template <int kStashes>
struct BlendSmem {
    float4 r0[kStages][kChunk];
    float4 r1[kStages][kChunk];
    float4 r2[kStages][kChunk];
    float stash[kConsumerWarps][kStashes][kSlots][32];
    unsigned char list[kConsumerWarps][4][kChunk];   // per warp and quarter: chunk-local indices of the survivors
    uint64_t full[kStages];
    uint64_t empty[kStages];
    unsigned int work;                          // current work item (chunk-local tile index) or 0xffffffff
    unsigned int warps_done;
    unsigned int max_last;                      // backward: largest n_contrib of the tile
};

__device__ __forceinline__ float2 unpack_extent(float packed) {
    const unsigned int u = __float_as_uint(packed);
    return make_float2(__half2float(__ushort_as_half(static_cast<unsigned short>(u & 0xffffu))),
                       __half2float(__ushort_as_half(static_cast<unsigned short>(u >> 16))));
}

// Cull the chunk's m records against the four 4x2 quarters of the warp's 8x4 block at (wx0, wy0): lane = record, one
// ballot per quarter, warp-parallel compaction of the survivors' chunk-local indices into list[quarter][...] (ascending).
// `limit`: records at or beyond it are ignored (backward: beyond the warp's largest n_contrib).  Returns the four counts.
__device__ __forceinline__ uint4 cull_chunk(const float4* r0, unsigned int m, unsigned int limit, float wx0, float wy0,
                                            unsigned char (*list)[kChunk], int lane) {
    unsigned int n0 = 0, n1 = 0, n2 = 0, n3 = 0;
    const unsigned int lt = (1u << lane) - 1u;
    for (unsigned int sub = 0; sub < m; sub += 32) {
        const unsigned int e = sub + lane;
        bool px0 = false, px1 = false, py0 = false, py1 = false;
        if (e < m && e < limit) {
            const float4 q = r0[e];
            const float2 ext = unpack_extent(q.z);
            const float xl = q.x - ext.x, xh = q.x + ext.x, yl = q.y - ext.y, yh = q.y + ext.y;
            px0 = (xh >= wx0) && (xl <= wx0 + 3.0f);
            px1 = (xh >= wx0 + 4.0f) && (xl <= wx0 + 7.0f);
            py0 = (yh >= wy0) && (yl <= wy0 + 1.0f);
            py1 = (yh >= wy0 + 2.0f) && (yl <= wy0 + 3.0f);
        }
        const unsigned int m0 = __ballot_sync(kFull, px0 && py0);
        const unsigned int m1 = __ballot_sync(kFull, px1 && py0);
        const unsigned int m2 = __ballot_sync(kFull, px0 && py1);
        const unsigned int m3 = __ballot_sync(kFull, px1 && py1);
        if (px0 && py0) list[0][n0 + __popc(m0 & lt)] = static_cast<unsigned char>(e);
        if (px1 && py0) list[1][n1 + __popc(m1 & lt)] = static_cast<unsigned char>(e);
        if (px0 && py1) list[2][n2 + __popc(m2 & lt)] = static_cast<unsigned char>(e);
        if (px1 && py1) list[3][n3 + __popc(m3 & lt)] = static_cast<unsigned char>(e);
        n0 += __popc(m0); n1 += __popc(m1); n2 += __popc(m2); n3 += __popc(m3);
    }
    __syncwarp();
    return make_uint4(n0, n1, n2, n3);
}

// ------------------------------------------------------------------------------------------------ forward
struct FwdArgs {
    RenderGeom g;
    int render_base;
    const unsigned int* tile_off;
    const unsigned int* tile_cnt;
    const float4 *rec0, *rec1, *rec2;
    const float* bg;
    unsigned int* n_contrib;
    uint2* tile_time;
    float *out_color, *out_depth, *out_alpha;
    const unsigned int *work_blend, *work_empty;
    WorkCounts* wc;
    int clamp_color;
};

using FwdSmem = BlendSmem<1>;
using BwdSmem = BlendSmem<2>;

__global__ void __launch_bounds__(kBlendThreads) blend_forward_kernel(FwdArgs a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    FwdSmem& sm = *reinterpret_cast<FwdSmem*>(smem_raw);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const bool is_producer = warp == kConsumerWarps;
    const size_t P = size_t(a.g.H) * a.g.W;
    const float bg0 = a.bg[0], bg1 = a.bg[1], bg2 = a.bg[2];
    const unsigned int n_blend = a.wc->n_blend, n_empty = a.wc->n_empty;

    // ---------------- tiles with instances
    unsigned long long t_begin = 0;
    size_t timed_tile = 0;
    for (bool first_tile = true;; first_tile = false) {
        __syncthreads();                         // previous tile fully retired; no bulk copy in flight
        if (tid == 0) {
            const unsigned long long now = global_timer_ns();
            if (!first_tile) a.tile_time[timed_tile] = make_uint2((unsigned int)t_begin, (unsigned int)(now - t_begin));
            t_begin = now;
            const unsigned int w = atomicAdd(&a.wc->blend_cursor, 1u);
            sm.work = (w < n_blend) ? a.work_blend[w] : 0xffffffffu;
            sm.warps_done = 0;
#pragma unroll
            for (int s = 0; s < kStages; ++s) {
                if (!first_tile) { mbar_inval(&sm.full[s]); mbar_inval(&sm.empty[s]); }
                mbar_init(&sm.full[s], 1); mbar_init(&sm.empty[s], kConsumerWarps);
            }
            fence_barrier_init();
        }
        __syncthreads();
        const unsigned int tile_local = sm.work;
        if (tile_local == 0xffffffffu) break;
        const int rl = tile_local / a.g.num_tiles;
        const int tile = tile_local - rl * a.g.num_tiles;
        const int r = a.render_base + rl;
        const size_t tg = size_t(r) * a.g.num_tiles + tile;
        timed_tile = tg;
        const unsigned int n = a.tile_cnt[tg];
        const size_t off = a.tile_off[tg];
        const unsigned int nchunks = (n + kChunk - 1) / kChunk;

        if (is_producer) {
            if (lane == 0) {
                unsigned int issued = 0;
                for (unsigned int c = 0; c < nchunks; ++c) {
                    const int s = c % kStages;
                    if (c >= kStages) {
                        const uint32_t par = ((c / kStages) - 1) & 1;
                        bool stop = false;
                        while (!mbar_try_wait(&sm.empty[s], par)) {
                            if (*reinterpret_cast<volatile unsigned int*>(&sm.warps_done) == kConsumerWarps) { stop = true; break; }
                        }
                        if (stop) break;
                    }
                    if (*reinterpret_cast<volatile unsigned int*>(&sm.warps_done) == kConsumerWarps) break;
                    const unsigned int m = min(unsigned(kChunk), n - c * kChunk);
                    const uint32_t bytes = m * 16u;
                    mbar_arrive_expect_tx(&sm.full[s], 3u * bytes);
                    tma_load_1d(sm.r0[s], a.rec0 + off + size_t(c) * kChunk, bytes, &sm.full[s]);
                    tma_load_1d(sm.r1[s], a.rec1 + off + size_t(c) * kChunk, bytes, &sm.full[s]);
                    tma_load_1d(sm.r2[s], a.rec2 + off + size_t(c) * kChunk, bytes, &sm.full[s]);
                    issued = c + 1;
                }
                // drain: every issued copy must have landed before the barriers are re-initialised
                const unsigned int first = issued > unsigned(kStages) ? issued - kStages : 0u;
                for (unsigned int c = first; c < issued; ++c) mbar_wait(&sm.full[c % kStages], (c / kStages) & 1);
            }
            continue;
        }

        // ---- consumer warp: 8x4 pixel block
        const int tx = tile % a.g.tiles_x, ty = tile / a.g.tiles_x;
        const int bx0 = tx * kTile + (warp & 1) * kBlockW, by0 = ty * kTile + (warp >> 1) * kBlockH;
        const int px = bx0 + (lane & 7), py = by0 + (lane >> 3);
        const bool inside = px < a.g.W && py < a.g.H;
        const float pxf = float(px), pyf = float(py);
        const float wx0 = float(bx0), wy0 = float(by0);
        const int qsel = ((lane >> 2) & 1) | ((lane >> 3) & 2);        // 4x2 pixel quarter of this lane: x half | 2 * y half
        const unsigned int qmask = 0x00000f0fu << ((lane & 4) | (lane & 16));   // the quarter's 8 lanes
        float (*stash)[32] = sm.stash[warp][0];
        unsigned char (*mylists)[kChunk] = sm.list[warp];
        const unsigned char* mylist = sm.list[warp][qsel];
        float T = 1.0f, C0 = 0.0f, C1 = 0.0f, C2 = 0.0f, D = 0.0f, Wt = 0.0f;
        unsigned int last = 0;
        bool done = !inside;
        bool warp_done = __all_sync(kFull, done);
        if (warp_done && lane == 0) atomicAdd(&sm.warps_done, 1u);
#ifdef SGR_PROFILE_WARPS
        long long pf_wait = 0, pf_cull = 0, pf_a = 0, pf_b = 0, pf_t0 = clock64(), pf_t;
        int pf_trips = 0, pf_passes = 0;
#define PF_MARK(acc) do { long long now__ = clock64(); acc += now__ - pf_t; pf_t = now__; } while (0)
#else
#define PF_MARK(acc) do {} while (0)
#endif
        for (unsigned int c = 0; c < nchunks; ++c) {
            const int s = c % kStages;
            const uint32_t par = (c / kStages) & 1;
            if (warp_done) {
                // keep the ring turning for the other warps; leave as soon as every warp is done
                bool all_done = false;
                while (!mbar_try_wait(&sm.full[s], par)) {
                    if (*reinterpret_cast<volatile unsigned int*>(&sm.warps_done) == kConsumerWarps) { all_done = true; break; }
                }
                if (all_done) break;
                __syncwarp();
                if (lane == 0) mbar_arrive(&sm.empty[s]);
                continue;
            }
#ifdef SGR_PROFILE_WARPS
            pf_t = clock64();
#endif
            mbar_wait(&sm.full[s], par);
            PF_MARK(pf_wait);
            const unsigned int m = min(unsigned(kChunk), n - c * kChunk);
            const unsigned int cbase = c * kChunk;
            const float4* r0 = sm.r0[s];
            const float4* r1 = sm.r1[s];
            const float4* r2 = sm.r2[s];
            {
                const uint4 cnt = cull_chunk(r0, m, m, wx0, wy0, mylists, lane);
                // quarters whose 8 pixels are all finished need no further evaluation
                const unsigned int dmask = __ballot_sync(kFull, done);
                const unsigned int my_n = ((dmask & qmask) == qmask) ? 0u : (qsel == 0 ? cnt.x : qsel == 1 ? cnt.y : qsel == 2 ? cnt.z : cnt.w);
                const int total = int(__reduce_max_sync(kFull, my_n));
                PF_MARK(pf_cull);
                for (int base = 0; base < total; base += kSlots) {
                    const int trips = min(kSlots, total - base);
                    // ---- phase A: alpha of the next `trips` survivors of this lane's quarter.  Trips are independent;
                    // blocks of four are written load-first so the compiler interleaves the four dependent chains.
                    for (int t0 = 0; t0 < trips; t0 += 4) {
                        const unsigned int packed = *reinterpret_cast<const unsigned int*>(mylist + base + t0);
                        float4 q0[4], q1[4];
                        bool has[4];
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            has[u] = unsigned(base + t0 + u) < my_n;
                            const unsigned int j = has[u] ? ((packed >> (8 * u)) & 0xffu) : 0u;
                            q0[u] = r0[j];
                            q1[u] = r1[j];
                        }
                        float al[4];
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            const float dx = q0[u].x - pxf, dy = q0[u].y - pyf;
                            const float power = gauss_power(q1[u].x, q1[u].y, q1[u].z, dx, dy);
                            const bool valid = has[u] && !(power > 0.0f) && !(power < q0[u].w);
                            const float alpha = fminf(kAlphaMax, q1[u].w * exp_core(valid ? power : 0.0f));
                            al[u] = valid ? alpha : 0.0f;
                        }
#pragma unroll
                        for (int u = 0; u < 4; ++u) stash[t0 + u][lane] = al[u];
                    }
                    __syncwarp();
#ifdef SGR_PROFILE_WARPS
                    pf_trips += trips; ++pf_passes;
#endif
                    PF_MARK(pf_a);
                    // ---- phase B: front-to-back compositing (the sequential part), predicated, loads hoisted
                    for (int t0 = 0; t0 < trips; t0 += 4) {
                        const unsigned int packed = *reinterpret_cast<const unsigned int*>(mylist + base + t0);
                        float al[4];
                        float4 q2[4];
                        unsigned int idx[4];
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            const unsigned int j = (unsigned(base + t0 + u) < my_n) ? ((packed >> (8 * u)) & 0xffu) : 0u;
                            al[u] = stash[t0 + u][lane];
                            q2[u] = r2[j];
                            idx[u] = cbase + j + 1;
                        }
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            const bool ok = !done && !(al[u] < kAlphaMin);
                            const float test_T = T * (1.0f - al[u]);
                            const bool stop = ok && (test_T < kTMin);
                            const bool blend = ok && !stop;
                            const float w0 = q2[u].x * al[u], w1 = q2[u].y * al[u], w2 = q2[u].z * al[u], w3 = q2[u].w * al[u];
                            C0 = blend ? C0 + w0 * T : C0;
                            C1 = blend ? C1 + w1 * T : C1;
                            C2 = blend ? C2 + w2 * T : C2;
                            Wt = blend ? Wt + al[u] * T : Wt;
                            D = blend ? D + w3 * T : D;
                            T = blend ? test_T : T;
                            last = blend ? idx[u] : last;
                            done = done || stop;
                        }
                    }
                    __syncwarp();
                    PF_MARK(pf_b);
                    if (__all_sync(kFull, done)) { warp_done = true; break; }
                }
            }
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(&sm.empty[s]);
                if (warp_done) atomicAdd(&sm.warps_done, 1u);
            }
        }
#ifdef SGR_PROFILE_WARPS
        if (lane == 0 && n > 8000)
            printf("tile r%d t%d n=%u warp %d: total %lld wait %lld cull %lld A %lld B %lld trips %d passes %d done=%d\n", r, tile,
                   n, warp, clock64() - pf_t0, pf_wait, pf_cull, pf_a, pf_b, pf_trips, pf_passes, int(warp_done));
#endif
        if (inside) {
            const size_t pix = size_t(py) * a.g.W + px;
            float c0 = C0 + T * bg0, c1 = C1 + T * bg1, c2 = C2 + T * bg2;
            if (a.clamp_color) {
                c0 = fminf(fmaxf(c0, 0.0f), 1.0f); c1 = fminf(fmaxf(c1, 0.0f), 1.0f); c2 = fminf(fmaxf(c2, 0.0f), 1.0f);
            }
            float* oc = a.out_color + size_t(r) * 3 * P;
            oc[pix] = c0; oc[P + pix] = c1; oc[2 * P + pix] = c2;
            a.out_depth[size_t(r) * P + pix] = D;
            a.out_alpha[size_t(r) * P + pix] = Wt;
            a.n_contrib[size_t(r) * P + pix] = last;
        }
    }

    // ---------------- tiles without instances: background only
    float e0 = bg0, e1 = bg1, e2 = bg2;
    if (a.clamp_color) { e0 = fminf(fmaxf(e0, 0.0f), 1.0f); e1 = fminf(fmaxf(e1, 0.0f), 1.0f); e2 = fminf(fmaxf(e2, 0.0f), 1.0f); }
    for (;;) {
        __syncthreads();
        if (tid == 0) {
            const unsigned int w = atomicAdd(&a.wc->empty_cursor, 1u);
            sm.work = (w < n_empty) ? a.work_empty[w] : 0xffffffffu;
        }
        __syncthreads();
        const unsigned int tile_local = sm.work;
        if (tile_local == 0xffffffffu) break;
        if (tid >= kTilePixels) continue;
        const int rl = tile_local / a.g.num_tiles;
        const int tile = tile_local - rl * a.g.num_tiles;
        const int r = a.render_base + rl;
        const int tx = tile % a.g.tiles_x, ty = tile / a.g.tiles_x;
        const int px = tx * kTile + (tid & 15), py = ty * kTile + (tid >> 4);
        if (px < a.g.W && py < a.g.H) {
            const size_t pix = size_t(py) * a.g.W + px;
            float* oc = a.out_color + size_t(r) * 3 * P;
            oc[pix] = e0; oc[P + pix] = e1; oc[2 * P + pix] = e2;
            a.out_depth[size_t(r) * P + pix] = 0.0f;
            a.out_alpha[size_t(r) * P + pix] = 0.0f;
            a.n_contrib[size_t(r) * P + pix] = 0u;
        }
    }
}

// ------------------------------------------------------------------------------------------------ backward
struct BwdArgs {
    RenderGeom g;
    int render_base;
    const unsigned int* tile_off;
    const unsigned int* tile_cnt;
    const unsigned int* sorted_ids;
    const float4 *rec0, *rec1, *rec2;
    const float* bg;
    const unsigned int* n_contrib;
    const float *out_alpha, *dL_dcolor, *dL_ddepth, *dL_dalpha;
    float* accum;
    size_t plane;
    const unsigned int* work_blend;
    WorkCounts* wc;
};

// Chunks are walked from the back of the list; chunk c of the walk is list chunk (nchunks - 1 - c).
template <bool kDepthAlphaGrads>
__global__ void __launch_bounds__(kBlendThreads) blend_backward_kernel(BwdArgs a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    BwdSmem& sm = *reinterpret_cast<BwdSmem*>(smem_raw);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const bool is_producer = warp == kConsumerWarps;
    const size_t P = size_t(a.g.H) * a.g.W;
    const float bg0 = a.bg[0], bg1 = a.bg[1], bg2 = a.bg[2];
    const unsigned int n_blend = a.wc->n_blend;
    const float ddelx_dx = 0.5f * float(a.g.W), ddely_dy = 0.5f * float(a.g.H);

    for (bool first_tile = true;; first_tile = false) {
        __syncthreads();
        if (tid == 0) {
            const unsigned int w = atomicAdd(&a.wc->blend_cursor, 1u);
            sm.work = (w < n_blend) ? a.work_blend[w] : 0xffffffffu;
            sm.max_last = 0;
#pragma unroll
            for (int s = 0; s < kStages; ++s) {
                if (!first_tile) { mbar_inval(&sm.full[s]); mbar_inval(&sm.empty[s]); }
                mbar_init(&sm.full[s], 1); mbar_init(&sm.empty[s], kConsumerWarps);
            }
            fence_barrier_init();
        }
        __syncthreads();
        const unsigned int tile_local = sm.work;
        if (tile_local == 0xffffffffu) break;
        const int rl = tile_local / a.g.num_tiles;
        const int tile = tile_local - rl * a.g.num_tiles;
        const int r = a.render_base + rl;
        const size_t tg = size_t(r) * a.g.num_tiles + tile;
        const unsigned int n = a.tile_cnt[tg];
        const size_t off = a.tile_off[tg];

        // per-pixel inputs (consumer threads) and the tile's replay length
        const int tx = tile % a.g.tiles_x, ty = tile / a.g.tiles_x;
        const int bx0 = tx * kTile + (warp & 1) * kBlockW, by0 = ty * kTile + (warp >> 1) * kBlockH;
        const int px = bx0 + (lane & 7), py = by0 + (lane >> 3);
        const bool inside = !is_producer && px < a.g.W && py < a.g.H;
        unsigned int last = 0;
        float T_final = 1.0f, dp0 = 0, dp1 = 0, dp2 = 0, ddep = 0, dalp = 0;
        if (inside) {
            const size_t pix = size_t(py) * a.g.W + px;
            last = a.n_contrib[size_t(r) * P + pix];
            T_final = 1.0f - a.out_alpha[size_t(r) * P + pix];
            const float* dc = a.dL_dcolor + size_t(r) * 3 * P;
            dp0 = dc[pix]; dp1 = dc[P + pix]; dp2 = dc[2 * P + pix];
            if (kDepthAlphaGrads) {
                if (a.dL_ddepth) ddep = a.dL_ddepth[size_t(r) * P + pix];
                if (a.dL_dalpha) dalp = a.dL_dalpha[size_t(r) * P + pix];
            }
        }
        const unsigned int wmax = __reduce_max_sync(kFull, last);
        if (lane == 0 && wmax) atomicMax(&sm.max_last, wmax);
        __syncthreads();
        const unsigned int n_eff = min(n, sm.max_last);
        const unsigned int nchunks = (n_eff + kChunk - 1) / kChunk;

        if (is_producer) {
            if (lane == 0) {
                for (unsigned int c = 0; c < nchunks; ++c) {
                    const int s = c % kStages;
                    if (c >= kStages) mbar_wait(&sm.empty[s], ((c / kStages) - 1) & 1);
                    const unsigned int lc = nchunks - 1 - c;
                    const unsigned int m = min(unsigned(kChunk), n_eff - lc * kChunk);
                    const uint32_t bytes = m * 16u;
                    mbar_arrive_expect_tx(&sm.full[s], 3u * bytes);
                    tma_load_1d(sm.r0[s], a.rec0 + off + size_t(lc) * kChunk, bytes, &sm.full[s]);
                    tma_load_1d(sm.r1[s], a.rec1 + off + size_t(lc) * kChunk, bytes, &sm.full[s]);
                    tma_load_1d(sm.r2[s], a.rec2 + off + size_t(lc) * kChunk, bytes, &sm.full[s]);
                    // (Gaussian ids are 4-byte entries at a 4-byte aligned segment start: not TMA-able; the consumers
                    //  read the id of an entry that contributed straight from global memory.)
                }
            }
            continue;
        }

        const float pxf = float(px), pyf = float(py);
        const float wx0 = float(bx0), wy0 = float(by0);
        const int qsel = ((lane >> 2) & 1) | ((lane >> 3) & 2);        // 4x2 pixel quarter of this lane: x half | 2 * y half
        const unsigned int qbase = (lane & 4) | (lane & 16);          // first lane of the quarter
        const unsigned int qmask = 0x00000f0fu << qbase;              // the quarter's 8 lanes (lane bits 0, 1, 3)
        unsigned char (*mylists)[kChunk] = sm.list[warp];
        const unsigned char* mylist = sm.list[warp][qsel];
        const bool p0 = (lane & 1) != 0, p1 = (lane & 2) != 0, p3 = (lane & 8) != 0;
        const int comp = (p3 ? 4 : 0) | (p1 ? 2 : 0) | (p0 ? 1 : 0);  // component this lane owns after the butterfly
        float (*stA)[32] = sm.stash[warp][0];
        float (*stW)[32] = sm.stash[warp][1];
        float T = T_final;
        float ar0 = 0, ar1 = 0, ar2 = 0, adr = 0, aar = 0, last_alpha = 0, lc0 = 0, lc1 = 0, lc2 = 0, last_depth = 0;
        const float bg_dot = (bg0 * dp0 + bg1 * dp1) + bg2 * dp2;
        float* acc = a.accum + size_t(rl) * a.g.N;
        const unsigned int* ids = a.sorted_ids + off;

        for (unsigned int c = 0; c < nchunks; ++c) {
            const int s = c % kStages;
            mbar_wait(&sm.full[s], (c / kStages) & 1);
            const unsigned int lcn = nchunks - 1 - c;
            const unsigned int cbase = lcn * kChunk;
            const unsigned int m = min(unsigned(kChunk), n_eff - cbase);
            const float4* r0 = sm.r0[s];
            const float4* r1 = sm.r1[s];
            const float4* r2 = sm.r2[s];
            if (cbase < wmax) {                               // some pixel of this warp replays entries of the chunk
                const uint4 cnt = cull_chunk(r0, m, wmax - cbase, wx0, wy0, mylists, lane);
                const unsigned int my_n = qsel == 0 ? cnt.x : qsel == 1 ? cnt.y : qsel == 2 ? cnt.z : cnt.w;
                const int total = int(max(max(cnt.x, cnt.y), max(cnt.z, cnt.w)));
                // trip t of the chunk handles the quarter's survivor number (my_n - 1 - t): back to front
                for (int base = 0; base < total; base += kSlots) {
                    const int trips = min(kSlots, total - base);
                    // ---- phase A: alpha (0 = no blend); load-first blocks of four so the independent chains interleave
                    for (int t0 = 0; t0 < trips; t0 += 4) {
                        float4 q0[4], q1[4];
                        bool has[4];
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            const int k = int(my_n) - 1 - (base + t0 + u);
                            const unsigned int j = k >= 0 ? mylist[k] : 0u;
                            has[u] = k >= 0 && (cbase + j < last);
                            q0[u] = r0[j];
                            q1[u] = r1[j];
                        }
                        float al[4];
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            const float dx = q0[u].x - pxf, dy = q0[u].y - pyf;
                            const float power = gauss_power(q1[u].x, q1[u].y, q1[u].z, dx, dy);
                            const bool valid = has[u] && !(power > 0.0f) && !(power < q0[u].w);
                            const float alpha = fminf(kAlphaMax, q1[u].w * exp_core(valid ? power : 0.0f));
                            al[u] = (valid && !(alpha < kAlphaMin)) ? alpha : 0.0f;
                        }
#pragma unroll
                        for (int u = 0; u < 4; ++u) stA[t0 + u][lane] = al[u];
                    }
                    __syncwarp();
                    // ---- phase B: the sequential per-pixel recurrences -> dL/dalpha and blend weight per trip
                    for (int t0 = 0; t0 < trips; t0 += 4) {
                        float al[4], dl[4], wg[4];
                        float4 q2[4];
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            const int k = int(my_n) - 1 - (base + t0 + u);
                            al[u] = stA[t0 + u][lane];
                            q2[u] = r2[k >= 0 ? mylist[k] : 0u];
                        }
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            const float alpha = al[u];
                            float wgt = 0.0f, dL_dal = 0.0f;
                            if (alpha != 0.0f) {
                                const float om = 1.0f - alpha;
                                T = T / om;
                                wgt = alpha * T;
                                const float ola = 1.0f - last_alpha;
                                ar0 = last_alpha * lc0 + ola * ar0; lc0 = q2[u].x;
                                dL_dal += (q2[u].x - ar0) * dp0;
                                ar1 = last_alpha * lc1 + ola * ar1; lc1 = q2[u].y;
                                dL_dal += (q2[u].y - ar1) * dp1;
                                ar2 = last_alpha * lc2 + ola * ar2; lc2 = q2[u].z;
                                dL_dal += (q2[u].z - ar2) * dp2;
                                if (kDepthAlphaGrads) {
                                    adr = last_alpha * last_depth + ola * adr; last_depth = q2[u].w;
                                    dL_dal += (q2[u].w - adr) * ddep;
                                    aar = last_alpha + ola * aar;
                                    dL_dal += (1.0f - aar) * dalp;
                                }
                                dL_dal *= T;
                                last_alpha = alpha;
                                dL_dal += (-T_final / om) * bg_dot;
                            }
                            dl[u] = dL_dal;
                            wg[u] = wgt;
                        }
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            stA[t0 + u][lane] = dl[u];
                            stW[t0 + u][lane] = wg[u];
                        }
                    }
                    __syncwarp();
                    // ---- phase C: per-Gaussian partial sums over the quarter's 8 pixels (independent trips)
                    for (int t = 0; t < trips; ++t) {
                        const int k = int(my_n) - 1 - (base + t);
                        const unsigned int j = k >= 0 ? mylist[k] : 0u;
                        const float wgt = stW[t][lane];
                        const float dL_dal = stA[t][lane];
                        const float4 q0 = r0[j];
                        const float4 q1 = r1[j];
                        const bool active = wgt != 0.0f;
                        const unsigned int act = __ballot_sync(kFull, active);
                        if (act == 0) continue;
                        const float dx = q0.x - pxf, dy = q0.y - pyf;
                        const float power = gauss_power(q1.x, q1.y, q1.z, dx, dy);
                        const float G = exp_core(active ? power : 0.0f);
                        const float dL_dG = q1.w * dL_dal;
                        const float gdx = G * dx, gdy = G * dy;
                        const float dG_ddelx = -gdx * q1.x - gdy * q1.y;
                        const float dG_ddely = -gdy * q1.z - gdx * q1.y;
                        float v0 = active ? dL_dG * dG_ddelx * ddelx_dx : 0.0f;
                        float v1 = active ? dL_dG * dG_ddely * ddely_dy : 0.0f;
                        float v2 = active ? -0.5f * gdx * dx * dL_dG : 0.0f;
                        float v3 = active ? -0.5f * gdx * dy * dL_dG : 0.0f;
                        float v4 = active ? -0.5f * gdy * dy * dL_dG : 0.0f;
                        float v5 = active ? G * dL_dal : 0.0f;
                        float v6 = wgt * dp0, v7 = wgt * dp1, v8 = wgt * dp2, v9 = wgt * ddep;
                        // transposing butterfly over lane bits 0, 1, 3 for v0..v7
                        const float b0 = (p0 ? v1 : v0) + __shfl_xor_sync(kFull, p0 ? v0 : v1, 1);
                        const float b1 = (p0 ? v3 : v2) + __shfl_xor_sync(kFull, p0 ? v2 : v3, 1);
                        const float b2 = (p0 ? v5 : v4) + __shfl_xor_sync(kFull, p0 ? v4 : v5, 1);
                        const float b3 = (p0 ? v7 : v6) + __shfl_xor_sync(kFull, p0 ? v6 : v7, 1);
                        const float c0 = (p1 ? b1 : b0) + __shfl_xor_sync(kFull, p1 ? b0 : b1, 2);
                        const float c1 = (p1 ? b3 : b2) + __shfl_xor_sync(kFull, p1 ? b2 : b3, 2);
                        const float d0 = (p3 ? c1 : c0) + __shfl_xor_sync(kFull, p3 ? c0 : c1, 8);
                        v8 += __shfl_xor_sync(kFull, v8, 1);
                        v8 += __shfl_xor_sync(kFull, v8, 2);
                        v8 += __shfl_xor_sync(kFull, v8, 8);
                        if (kDepthAlphaGrads) {
                            v9 += __shfl_xor_sync(kFull, v9, 1);
                            v9 += __shfl_xor_sync(kFull, v9, 2);
                            v9 += __shfl_xor_sync(kFull, v9, 8);
                        }
                        if (act & qmask) {
                            const unsigned int id = __ldg(ids + cbase + j);
                            if (d0 != 0.0f) atomicAdd(acc + size_t(comp) * a.plane + id, d0);
                            if (comp == 0 && v8 != 0.0f) atomicAdd(acc + size_t(8) * a.plane + id, v8);
                            if (kDepthAlphaGrads && comp == 1 && v9 != 0.0f) atomicAdd(acc + size_t(9) * a.plane + id, v9);
                        }
                    }
                    __syncwarp();
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&sm.empty[s]);
        }
    }
}

int g_num_sms = 0;
int num_sms() {
    if (g_num_sms == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
        if (g_num_sms <= 0) g_num_sms = 148;
    }
    return g_num_sms;
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

}  // namespace

cudaError_t launch_blend_forward(const ChunkCtx& c, float* out_color, float* out_depth, float* out_alpha) {
    FwdArgs a;
    a.g = c.g; a.render_base = c.render_base; a.tile_off = c.tile_off; a.tile_cnt = c.tile_cnt;
    a.rec0 = c.rec0; a.rec1 = c.rec1; a.rec2 = c.rec2; a.bg = c.p->bg; a.n_contrib = c.n_contrib;
    a.tile_time = c.tile_time;
    a.out_color = out_color; a.out_depth = out_depth; a.out_alpha = out_alpha;
    a.work_blend = c.work_blend; a.work_empty = c.work_empty; a.wc = c.work_counts;
    a.clamp_color = (c.p->flags & SGR_FLAG_CLAMP_COLOR) ? 1 : 0;
    static int per_sm = 0;
    if (per_sm == 0) {
        cudaError_t e = prepare_kernel(blend_forward_kernel, sizeof(FwdSmem), &per_sm);
        if (e != cudaSuccess) return e;
    }
    const int total_tiles = c.num_renders * c.g.num_tiles;
    const int grid = min(num_sms() * per_sm, total_tiles);
    blend_forward_kernel<<<grid, kBlendThreads, sizeof(FwdSmem), c.stream>>>(a);
    return cudaGetLastError();
}

cudaError_t launch_blend_backward(const ChunkCtx& c, const float* out_alpha, const float* dL_dcolor,
                                  const float* dL_ddepth, const float* dL_dalpha) {
    BwdArgs a;
    a.g = c.g; a.render_base = c.render_base; a.tile_off = c.tile_off; a.tile_cnt = c.tile_cnt;
    a.sorted_ids = c.sorted_ids; a.rec0 = c.rec0; a.rec1 = c.rec1; a.rec2 = c.rec2; a.bg = c.p->bg;
    a.n_contrib = c.n_contrib; a.out_alpha = out_alpha; a.dL_dcolor = dL_dcolor; a.dL_ddepth = dL_ddepth;
    a.dL_dalpha = dL_dalpha; a.accum = c.accum; a.plane = size_t(c.num_renders) * c.g.N;
    a.work_blend = c.work_blend; a.wc = c.work_counts;
    const int total_tiles = c.num_renders * c.g.num_tiles;
    if (dL_ddepth || dL_dalpha) {
        static int per_sm = 0;
        if (per_sm == 0) {
            cudaError_t e = prepare_kernel(blend_backward_kernel<true>, sizeof(BwdSmem), &per_sm);
            if (e != cudaSuccess) return e;
        }
        blend_backward_kernel<true><<<min(num_sms() * per_sm, total_tiles), kBlendThreads, sizeof(BwdSmem), c.stream>>>(a);
    } else {
        static int per_sm = 0;
        if (per_sm == 0) {
            cudaError_t e = prepare_kernel(blend_backward_kernel<false>, sizeof(BwdSmem), &per_sm);
            if (e != cudaSuccess) return e;
        }
        blend_backward_kernel<false><<<min(num_sms() * per_sm, total_tiles), kBlendThreads, sizeof(BwdSmem), c.stream>>>(a);
    }
    return cudaGetLastError();
}

}  // namespace sgr
