// sgr_blend.cu — the alpha-compositing kernels (forward and backward) of the B200 rasteriser.
//
// Replaces upstream renderCUDA forward/backward (third-party diff_gaussian_rasterization; call site
// /root/reference/core/gaussians/gs.py:99-106).  Design (DESIGN.md "blend"):
//   * persistent CTAs (a multiple of the SM count) pull (render, tile) work items from a device-side queue that is
//     ordered longest-list-first (sgr_binning.cu::worklist_kernel);
//   * the tile's depth-ordered 48-byte records (three float4 streams, contiguous per tile because the per-tile sort
//     gathers them) are staged into a shared-memory ring with 1-D TMA bulk copies (cp.async.bulk) completing on
//     mbarriers, issued by a dedicated producer warp; the eight consumer warps run decoupled from each other;
//   * every consumer warp owns an 8x4 pixel block.  It first tests 32 records at a time (lane = record) against the
//     block with the conservative alpha >= 1/255 extent computed in the preprocess kernel, ballots, and evaluates
//     only the surviving records (lane = pixel).  Culled records would have been skipped by the alpha test, so the
//     per-pixel arithmetic, the contributor index and every output bit equal the straightforward kernel's;
//   * backward: the same walk in reverse; per-Gaussian partial sums are reduced across the 32 pixels of the warp with
//     shuffles before a single atomic per component.
//
// Compiled with --fmad=false: the per-pixel expressions are the oracle's (oracle/sgr_oracle.cpp::blend_forward /
// blend_backward) evaluated in the same order, which makes colour, depth, alpha and n_contrib bit-exact.
#include <cuda_fp16.h>

#include "sgr_common.cuh"

namespace sgr {
namespace {

constexpr int kChunk = 128;                    // records per ring stage
constexpr int kStages = 4;
constexpr int kConsumerWarps = 8;              // 8 warps x (8x4 pixels) = one 16x16 tile
constexpr int kBlendThreads = (kConsumerWarps + 1) * 32;
constexpr int kBlockW = 8, kBlockH = 4;        // pixel block of one warp

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

struct BlendSmem {
    float4 r0[kStages][kChunk];
    float4 r1[kStages][kChunk];
    float4 r2[kStages][kChunk];
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

// ------------------------------------------------------------------------------------------------ forward
struct FwdArgs {
    RenderGeom g;
    int render_base;
    const unsigned int* tile_off;
    const unsigned int* tile_cnt;
    const float4 *rec0, *rec1, *rec2;
    const float* bg;
    unsigned int* n_contrib;
    float *out_color, *out_depth, *out_alpha;
    const unsigned int *work_blend, *work_empty;
    WorkCounts* wc;
    int clamp_color;
};

__global__ void __launch_bounds__(kBlendThreads) blend_forward_kernel(FwdArgs a) {
    __shared__ __align__(128) BlendSmem sm;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const bool is_producer = warp == kConsumerWarps;
    const size_t P = size_t(a.g.H) * a.g.W;
    const float bg0 = a.bg[0], bg1 = a.bg[1], bg2 = a.bg[2];
    const unsigned int n_blend = a.wc->n_blend, n_empty = a.wc->n_empty;

    // ---------------- tiles with instances
    for (bool first_tile = true;; first_tile = false) {
        __syncthreads();                         // previous tile fully retired; no bulk copy in flight
        if (tid == 0) {
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
        const float wx0 = float(bx0), wx1 = float(bx0 + kBlockW - 1), wy0 = float(by0), wy1 = float(by0 + kBlockH - 1);
        float T = 1.0f, C0 = 0.0f, C1 = 0.0f, C2 = 0.0f, D = 0.0f, Wt = 0.0f;
        unsigned int last = 0;
        bool done = !inside;
        bool warp_done = __all_sync(0xffffffffu, done);
        if (warp_done && lane == 0) atomicAdd(&sm.warps_done, 1u);
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
            mbar_wait(&sm.full[s], par);
            const unsigned int m = min(unsigned(kChunk), n - c * kChunk);
            const unsigned int cbase = c * kChunk;
            for (unsigned int sub = 0; sub < m; sub += 32) {
                const unsigned int e = sub + lane;
                bool pass = false;
                if (e < m) {
                    const float4 q = sm.r0[s][e];
                    const float2 ext = unpack_extent(q.z);
                    pass = (q.x + ext.x >= wx0) && (q.x - ext.x <= wx1) && (q.y + ext.y >= wy0) && (q.y - ext.y <= wy1);
                }
                unsigned int mask = __ballot_sync(0xffffffffu, pass);
                while (mask) {
                    const unsigned int j = sub + (__ffs(mask) - 1);
                    mask &= mask - 1;
                    if (done) continue;
                    const float4 q0 = sm.r0[s][j];
                    const float4 q1 = sm.r1[s][j];
                    const float dx = q0.x - pxf, dy = q0.y - pyf;
                    const float power = gauss_power(q1.x, q1.y, q1.z, dx, dy);
                    if (power > 0.0f || power < q0.w) continue;
                    const float alpha = fminf(kAlphaMax, q1.w * exp_spec(power));
                    if (alpha < kAlphaMin) continue;
                    const float test_T = T * (1.0f - alpha);
                    if (test_T < kTMin) { done = true; continue; }
                    const float4 q2 = sm.r2[s][j];
                    C0 += q2.x * alpha * T;
                    C1 += q2.y * alpha * T;
                    C2 += q2.z * alpha * T;
                    Wt += alpha * T;
                    D += q2.w * alpha * T;
                    T = test_T;
                    last = cbase + j + 1;
                }
                if (__all_sync(0xffffffffu, done)) { warp_done = true; break; }
            }
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(&sm.empty[s]);
                if (warp_done) atomicAdd(&sm.warps_done, 1u);
            }
        }
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

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
    return v;
}

// Chunks are walked from the back of the list; chunk c of the walk is list chunk (nchunks - 1 - c).
template <bool kDepthAlphaGrads>
__global__ void __launch_bounds__(kBlendThreads) blend_backward_kernel(BwdArgs a) {
    __shared__ __align__(128) BlendSmem sm;
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
        unsigned int wmax = last;
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) wmax = max(wmax, __shfl_xor_sync(0xffffffffu, wmax, d));
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
        const float wx0 = float(bx0), wx1 = float(bx0 + kBlockW - 1), wy0 = float(by0), wy1 = float(by0 + kBlockH - 1);
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
            if (cbase < wmax) {                               // some pixel of this warp replays entries of the chunk
                const unsigned int nsub = (m + 31) / 32;
                for (unsigned int sb = nsub; sb-- > 0;) {
                    const unsigned int sub = sb * 32;
                    const unsigned int e = sub + lane;
                    bool pass = false;
                    if (e < m && cbase + e < wmax) {
                        const float4 q = sm.r0[s][e];
                        const float2 ext = unpack_extent(q.z);
                        pass = (q.x + ext.x >= wx0) && (q.x - ext.x <= wx1) && (q.y + ext.y >= wy0) && (q.y - ext.y <= wy1);
                    }
                    unsigned int mask = __ballot_sync(0xffffffffu, pass);
                    while (mask) {
                        const unsigned int hb = 31 - __clz(mask);     // back to front
                        mask &= ~(1u << hb);
                        const unsigned int j = sub + hb;
                        const unsigned int contributor = cbase + j;
                        float v0 = 0, v1 = 0, v2 = 0, v3 = 0, v4 = 0, v5 = 0, v6 = 0, v7 = 0, v8 = 0, v9 = 0;
                        bool active = false;
                        if (contributor < last) {
                            const float4 q0 = sm.r0[s][j];
                            const float4 q1 = sm.r1[s][j];
                            const float dx = q0.x - pxf, dy = q0.y - pyf;
                            const float power = gauss_power(q1.x, q1.y, q1.z, dx, dy);
                            if (!(power > 0.0f) && !(power < q0.w)) {
                                const float G = exp_spec(power);
                                const float alpha = fminf(kAlphaMax, q1.w * G);
                                if (!(alpha < kAlphaMin)) {
                                    active = true;
                                    const float4 q2 = sm.r2[s][j];
                                    T = T / (1.0f - alpha);
                                    const float w = alpha * T;
                                    float dL_dal = 0.0f;
                                    ar0 = last_alpha * lc0 + (1.0f - last_alpha) * ar0; lc0 = q2.x;
                                    dL_dal += (q2.x - ar0) * dp0; v6 = w * dp0;
                                    ar1 = last_alpha * lc1 + (1.0f - last_alpha) * ar1; lc1 = q2.y;
                                    dL_dal += (q2.y - ar1) * dp1; v7 = w * dp1;
                                    ar2 = last_alpha * lc2 + (1.0f - last_alpha) * ar2; lc2 = q2.z;
                                    dL_dal += (q2.z - ar2) * dp2; v8 = w * dp2;
                                    if (kDepthAlphaGrads) {
                                        adr = last_alpha * last_depth + (1.0f - last_alpha) * adr; last_depth = q2.w;
                                        dL_dal += (q2.w - adr) * ddep; v9 = w * ddep;
                                        aar = last_alpha + (1.0f - last_alpha) * aar;
                                        dL_dal += (1.0f - aar) * dalp;
                                    }
                                    dL_dal *= T;
                                    last_alpha = alpha;
                                    dL_dal += (-T_final / (1.0f - alpha)) * bg_dot;
                                    const float dL_dG = q1.w * dL_dal;
                                    const float gdx = G * dx, gdy = G * dy;
                                    const float dG_ddelx = -gdx * q1.x - gdy * q1.y;
                                    const float dG_ddely = -gdy * q1.z - gdx * q1.y;
                                    v0 = dL_dG * dG_ddelx * ddelx_dx;
                                    v1 = dL_dG * dG_ddely * ddely_dy;
                                    v2 = -0.5f * gdx * dx * dL_dG;
                                    v3 = -0.5f * gdx * dy * dL_dG;
                                    v4 = -0.5f * gdy * dy * dL_dG;
                                    v5 = G * dL_dal;
                                }
                            }
                        }
                        if (__any_sync(0xffffffffu, active)) {
                            const unsigned int id = __ldg(ids + contributor);
                            v0 = warp_sum(v0); v1 = warp_sum(v1); v2 = warp_sum(v2); v3 = warp_sum(v3); v4 = warp_sum(v4);
                            v5 = warp_sum(v5); v6 = warp_sum(v6); v7 = warp_sum(v7); v8 = warp_sum(v8);
                            if (kDepthAlphaGrads) v9 = warp_sum(v9);
                            // lanes 0..9 each add one component
                            float mine = v0;
                            mine = lane == 1 ? v1 : mine; mine = lane == 2 ? v2 : mine; mine = lane == 3 ? v3 : mine;
                            mine = lane == 4 ? v4 : mine; mine = lane == 5 ? v5 : mine; mine = lane == 6 ? v6 : mine;
                            mine = lane == 7 ? v7 : mine; mine = lane == 8 ? v8 : mine; mine = lane == 9 ? v9 : mine;
                            const int nplanes = kDepthAlphaGrads ? kAccumPlanes : kAccumPlanes - 1;
                            if (lane < nplanes && mine != 0.0f) atomicAdd(acc + size_t(lane) * a.plane + id, mine);
                        }
                    }
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

template <typename K>
int resident_ctas(K kernel) {
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kBlendThreads, 0) != cudaSuccess || per_sm <= 0) per_sm = 1;
    return per_sm;
}

}  // namespace

cudaError_t launch_blend_forward(const ChunkCtx& c, float* out_color, float* out_depth, float* out_alpha) {
    FwdArgs a;
    a.g = c.g; a.render_base = c.render_base; a.tile_off = c.tile_off; a.tile_cnt = c.tile_cnt;
    a.rec0 = c.rec0; a.rec1 = c.rec1; a.rec2 = c.rec2; a.bg = c.p->bg; a.n_contrib = c.n_contrib;
    a.out_color = out_color; a.out_depth = out_depth; a.out_alpha = out_alpha;
    a.work_blend = c.work_blend; a.work_empty = c.work_empty; a.wc = c.work_counts;
    a.clamp_color = (c.p->flags & SGR_FLAG_CLAMP_COLOR) ? 1 : 0;
    static int per_sm = 0;
    if (per_sm == 0) per_sm = resident_ctas(blend_forward_kernel);
    const int total_tiles = c.num_renders * c.g.num_tiles;
    const int grid = min(num_sms() * per_sm, total_tiles);
    blend_forward_kernel<<<grid, kBlendThreads, 0, c.stream>>>(a);
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
        if (per_sm == 0) per_sm = resident_ctas(blend_backward_kernel<true>);
        blend_backward_kernel<true><<<min(num_sms() * per_sm, total_tiles), kBlendThreads, 0, c.stream>>>(a);
    } else {
        static int per_sm = 0;
        if (per_sm == 0) per_sm = resident_ctas(blend_backward_kernel<false>);
        blend_backward_kernel<false><<<min(num_sms() * per_sm, total_tiles), kBlendThreads, 0, c.stream>>>(a);
    }
    return cudaGetLastError();
}

}  // namespace sgr
