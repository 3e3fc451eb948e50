// sgr_preprocess.cu — per-Gaussian kernels: 3D->2D projection + covariance conic + tile rectangle (forward),
// instance scatter, and the per-Gaussian backward.
//
// THIS FILE IS COMPILED WITH --fmad=false: every index-determining expression (radius, pixel centre, tile
// rectangle, conic) is evaluated with the same IEEE fp32 operations in the same order as the CPU oracle
// (oracle/sgr_oracle.cpp::preprocess / compute_cov2d), which makes radii and tile rectangles bit-exact.
//
// Replaces upstream preprocessCUDA / duplicateWithKeys / computeCov2DCUDA+preprocessCUDA(backward)
// (third-party diff_gaussian_rasterization; call site /root/reference/core/gaussians/gs.py:99-106;
// SURVEY.md A.2, A.3, A.6).
#include <cuda_fp16.h>

#include "sgr_common.cuh"
#include "sgr_cov2d.cuh"

namespace sgr {
namespace {

// Cooperative load of `count` floats: 128-bit loads when the block's slice is 16-byte aligned.
__device__ __forceinline__ void stage_floats(const float* __restrict__ src, float* dst, int count) {
    if ((reinterpret_cast<uintptr_t>(src) & 15) == 0 && (count & 3) == 0) {
        const float4* s4 = reinterpret_cast<const float4*>(src);
        float4* d4 = reinterpret_cast<float4*>(dst);
        for (int k = threadIdx.x; k < (count >> 2); k += blockDim.x) d4[k] = __ldg(s4 + k);
    } else {
        for (int k = threadIdx.x; k < count; k += blockDim.x) dst[k] = __ldg(src + k);
    }
}

struct PreArgs {
    RenderGeom g;
    int render_base;
    const float *means, *cov, *colors, *opac, *view, *proj;
    float4 *g0, *g1, *g2;
    uint2* rect;
    int32_t* radii;
    unsigned int* tile_cnt;
};

constexpr int kPreThreads = 256;
#ifndef SGR_PRE_ITERS
#define SGR_PRE_ITERS 2
#endif
constexpr int kPreIters = SGR_PRE_ITERS;     // sub-batches of 256 Gaussians per CTA (amortises the histogram flush)
constexpr int kHistTiles = 4096;             // tiles per render whose counters fit the shared-memory histogram

// Tile counters: the (Gaussian, tile) instances of the CTA's 2048 Gaussians are first counted in shared memory and
// flushed with one global atomic per touched tile — the dense head / hand tiles otherwise serialise tens of thousands
// of same-address atomics in L2.
__global__ void __launch_bounds__(kPreThreads) preprocess_kernel(PreArgs a) {
    __shared__ float s_view[16], s_proj[16];
    __shared__ __align__(16) float s_mean[kPreThreads * 3];
    __shared__ __align__(16) float s_cov[kPreThreads * 6];
    __shared__ __align__(16) float s_col[kPreThreads * 3];
    __shared__ __align__(16) float s_op[kPreThreads];
    extern __shared__ unsigned int s_hist[];     // [num_tiles] when num_tiles <= kHistTiles

    const int rl = blockIdx.y;
    const int r = a.render_base + rl;
    const int b = r / a.g.V;
    const int N = a.g.N;
    const bool use_hist = a.g.num_tiles <= kHistTiles;
    if (threadIdx.x < 16) s_view[threadIdx.x] = __ldg(a.view + size_t(r) * 16 + threadIdx.x);
    else if (threadIdx.x < 32) s_proj[threadIdx.x - 16] = __ldg(a.proj + size_t(r) * 16 + threadIdx.x - 16);
    if (use_hist)
        for (int k = threadIdx.x; k < a.g.num_tiles; k += kPreThreads) s_hist[k] = 0;
    unsigned int* gcnt = a.tile_cnt + size_t(r) * a.g.num_tiles;
    unsigned int* cnt = use_hist ? s_hist : gcnt;

  for (int it = 0; it < kPreIters; ++it) {
    const int i0 = (blockIdx.x * kPreIters + it) * kPreThreads;
    if (i0 >= N) break;
    const int nblk = min(kPreThreads, N - i0);
    const size_t gbase = size_t(b) * N + i0;
    __syncthreads();                             // previous sub-batch fully consumed (and s_view / s_hist ready)
    stage_floats(a.means + gbase * 3, s_mean, nblk * 3);
    stage_floats(a.cov + gbase * 6, s_cov, nblk * 6);
    stage_floats(a.colors + gbase * 3, s_col, nblk * 3);
    stage_floats(a.opac + gbase, s_op, nblk);
    __syncthreads();

    const int t = threadIdx.x;
    if (t >= nblk) continue;
    const int i = i0 + t;
    const size_t oi = size_t(rl) * N + i;          // chunk-local record index
    const int W = a.g.W, H = a.g.H;
    const int gx = a.g.tiles_x, gy = a.g.tiles_y;

    int radius_i = 0;
    uint2 rect = make_uint2(0u, 0u);
    const float m[3] = {s_mean[3 * t], s_mean[3 * t + 1], s_mean[3 * t + 2]};
    const float* view = s_view;
    const float* proj = s_proj;
    const float pvz = view[2] * m[0] + view[6] * m[1] + view[10] * m[2] + view[14];
    if (pvz > 0.2f) {
        const float hx = proj[0] * m[0] + proj[4] * m[1] + proj[8] * m[2] + proj[12];
        const float hy = proj[1] * m[0] + proj[5] * m[1] + proj[9] * m[2] + proj[13];
        const float hw = proj[3] * m[0] + proj[7] * m[1] + proj[11] * m[2] + proj[15];
        const float pw = 1.0f / (hw + 0.0000001f);
        const float projx = hx * pw, projy = hy * pw;
        Cov2D c2;
        compute_cov2d(m, &s_cov[6 * t], view, a.g.tanfovx, a.g.tanfovy, W, H, c2);
        const float det = c2.a * c2.c - c2.b * c2.b;
        if (det != 0.0f) {
            const float det_inv = 1.0f / det;
            const float mid = 0.5f * (c2.a + c2.c);
            const float dd = mid * mid - det;
            const float disc = sqrtf(fmaxf(0.1f, dd));
            const float l1 = mid + disc, l2 = mid - disc;
            const float radius = ceilf(3.0f * sqrtf(fmaxf(l1, l2)));
            const float px = float(((double(projx) + 1.0) * double(W) - 1.0) * 0.5);
            const float py = float(((double(projy) + 1.0) * double(H) - 1.0) * 0.5);
            const int rminx = min(gx, max(0, __float2int_rz((px - radius) / float(kTile))));
            const int rminy = min(gy, max(0, __float2int_rz((py - radius) / float(kTile))));
            const int rmaxx = min(gx, max(0, __float2int_rz((px + radius + float(kTile - 1)) / float(kTile))));
            const int rmaxy = min(gy, max(0, __float2int_rz((py + radius + float(kTile - 1)) / float(kTile))));
            if ((rmaxx - rminx) * (rmaxy - rminy) != 0) {
                radius_i = __float2int_rz(radius);
                rect = make_uint2(uint32_t(rminx) | (uint32_t(rminy) << 16), uint32_t(rmaxx) | (uint32_t(rmaxy) << 16));
                const float opac = s_op[t];
                // Conservative extent of the region where this Gaussian can pass the alpha >= 1/255 test
                // (DESIGN.md "warp-level culling"): |dx| <= ex, |dy| <= ey.  Not index-determining.
                float ex = -1.0e30f, ey = -1.0e30f;          // opacity < 1/255 (or NaN): can never blend
                if (opac >= kAlphaMin) {
                    ex = 1.0e30f; ey = 1.0e30f;              // default: no culling
                    const float lmin = mid - sqrtf(fmaxf(0.0f, dd));
                    if (lmin > 0.0f && c2.a > 0.0f && c2.c > 0.0f) {
                        const float gsl = 4.0e-6f * ((c2.a + c2.c) / lmin);   // fp32 evaluation error of the quadratic form
                        if (gsl < 0.5f) {
                            const float tau = (logf(255.0f * opac) + 2.0e-4f) / (1.0f - gsl);
                            ex = sqrtf(2.0f * tau * c2.a) * 1.0001f + 0.01f;
                            ey = sqrtf(2.0f * tau * c2.c) * 1.0001f + 0.01f;
                        }
                    }
                }
                // rec0 = (x, y, half2(ex, ey) rounded up, pthr): power < pthr can never reach alpha >= 1/255
                // (clamped to -87 so that exp_core's argument range holds: below it exp_spec returns 0 anyway)
                const float pthr = (opac >= kAlphaMin) ? fmaxf(-(logf(255.0f * opac) + 1.0e-3f), -87.0f) : 1.0e30f;
                const unsigned int exy = uint32_t(__half_as_ushort(__float2half_ru(ex))) |
                                         (uint32_t(__half_as_ushort(__float2half_ru(ey))) << 16);
                a.g0[oi] = make_float4(px, py, __uint_as_float(exy), pthr);
                // rec1 = (-A/2, -B, -C/2, opacity) with conic (A, B, C) = (c, -b, a) / det as in the oracle
                const float cA = c2.c * det_inv, cB = -c2.b * det_inv, cC = c2.a * det_inv;
                a.g1[oi] = make_float4(-0.5f * cA, -cB, -0.5f * cC, opac);
                a.g2[oi] = make_float4(s_col[3 * t], s_col[3 * t + 1], s_col[3 * t + 2], pvz);
                for (int y = rminy; y < rmaxy; ++y)
                    for (int x = rminx; x < rmaxx; ++x) atomicAdd(cnt + y * gx + x, 1u);
            }
        }
    }
    a.radii[size_t(r) * N + i] = radius_i;
    a.rect[oi] = rect;
  }
    if (use_hist) {
        __syncthreads();
        for (int k = threadIdx.x; k < a.g.num_tiles; k += kPreThreads) {
            const unsigned int c = s_hist[k];
            if (c) atomicAdd(gcnt + k, c);
        }
    }
}

#ifndef SGR_SCATTER_ITERS
#define SGR_SCATTER_ITERS 3
#endif
constexpr int kScatterIters = SGR_SCATTER_ITERS;   // sub-batches of 256 Gaussians per scatter CTA (1048 CTAs for 8 x 100 K:
                                                   // one wave; with 2 the 1568 CTAs ran as 1.3 waves: 23 -> 21 us)

struct ScatterArgs {
    RenderGeom g;
    int render_base;
    const float4* g2;
    const uint2* rect;
    const unsigned int* tile_off;
    unsigned int* cursor;
    unsigned long long* keys;
    const WorkCounts* wc;
};

// duplicateWithKeys: one (depth bits << 32 | gaussian id) key per covered tile, appended to the tile's segment.
// A CTA handles kScatterIters * 256 Gaussians of one render: it counts its instances per tile in shared memory, reserves a
// contiguous range per touched tile with ONE global atomic, and then places its keys with shared-memory atomics.
__global__ void __launch_bounds__(256) scatter_kernel(ScatterArgs a) {
    extern __shared__ unsigned int s_sc[];       // cnt[num_tiles], base[num_tiles] when num_tiles <= kHistTiles
    if (a.wc->chunk_dropped) return;
    const int rl = blockIdx.y;
    const int T = a.g.num_tiles;
    const bool use_hist = T <= kHistTiles;
    unsigned int* s_cnt = s_sc;
    unsigned int* s_base = s_sc + T;
    const size_t tb = size_t(a.render_base + rl) * T;
    unsigned int* cur = a.cursor + size_t(rl) * T;
    const int i_lo = blockIdx.x * kScatterIters * 256;
    if (use_hist) {
        for (int k = threadIdx.x; k < T; k += 256) s_cnt[k] = 0;
        __syncthreads();
        for (int it = 0; it < kScatterIters; ++it) {
            const int i = i_lo + it * 256 + threadIdx.x;
            if (i >= a.g.N) break;
            const uint2 rc = a.rect[size_t(rl) * a.g.N + i];
            const int rminx = rc.x & 0xffff, rminy = rc.x >> 16, rmaxx = rc.y & 0xffff, rmaxy = rc.y >> 16;
            for (int y = rminy; y < rmaxy; ++y)
                for (int x = rminx; x < rmaxx; ++x) atomicAdd(&s_cnt[y * a.g.tiles_x + x], 1u);
        }
        __syncthreads();
        for (int k = threadIdx.x; k < T; k += 256) {
            const unsigned int c = s_cnt[k];
            if (c) s_base[k] = a.tile_off[tb + k] + atomicAdd(cur + k, c);
            s_cnt[k] = 0;
        }
        __syncthreads();
    }
    for (int it = 0; it < kScatterIters; ++it) {
        const int i = i_lo + it * 256 + threadIdx.x;
        if (i >= a.g.N) break;
        const size_t oi = size_t(rl) * a.g.N + i;
        const uint2 rc = a.rect[oi];
        const int rminx = rc.x & 0xffff, rminy = rc.x >> 16, rmaxx = rc.y & 0xffff, rmaxy = rc.y >> 16;
        if (rmaxx <= rminx || rmaxy <= rminy) continue;
        const unsigned long long key =
            (static_cast<unsigned long long>(__float_as_uint(a.g2[oi].w)) << 32) | static_cast<unsigned int>(i);
        for (int y = rminy; y < rmaxy; ++y)
            for (int x = rminx; x < rmaxx; ++x) {
                const int t = y * a.g.tiles_x + x;
                const size_t slot = use_hist ? size_t(s_base[t]) + atomicAdd(&s_cnt[t], 1u)
                                             : size_t(a.tile_off[tb + t]) + atomicAdd(cur + t, 1u);
                a.keys[slot] = key;
            }
    }
}

}  // namespace

cudaError_t launch_preprocess(const ChunkCtx& c, int32_t* radii) {
    PreArgs a;
    a.g = c.g; a.render_base = c.render_base;
    a.means = c.p->means3D; a.cov = c.p->cov3D; a.colors = c.p->colors; a.opac = c.p->opacities;
    a.view = c.p->viewmatrix; a.proj = c.p->projmatrix;
    a.g0 = c.g0; a.g1 = c.g1; a.g2 = c.g2; a.rect = c.rect; a.radii = radii; a.tile_cnt = c.tile_cnt;
    const int per_cta = kPreThreads * kPreIters;
    dim3 grid((c.g.N + per_cta - 1) / per_cta, c.num_renders);
    const size_t smem = c.g.num_tiles <= kHistTiles ? size_t(c.g.num_tiles) * 4 : 0;
    preprocess_kernel<<<grid, kPreThreads, smem, c.stream>>>(a);
    return cudaGetLastError();
}

cudaError_t launch_scatter(const ChunkCtx& c) {
    ScatterArgs a;
    a.g = c.g; a.render_base = c.render_base; a.g2 = c.g2; a.rect = c.rect; a.tile_off = c.tile_off;
    a.cursor = c.cursor; a.keys = c.keys; a.wc = c.work_counts;
    const int per_cta = 256 * kScatterIters;
    dim3 grid((c.g.N + per_cta - 1) / per_cta, c.num_renders);
    const size_t smem = c.g.num_tiles <= kHistTiles ? size_t(c.g.num_tiles) * 8 : 0;
    scatter_kernel<<<grid, 256, smem, c.stream>>>(a);
    return cudaGetLastError();
}

}  // namespace sgr
