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

namespace sgr {
namespace {

struct Cov2D {
    float a, b, c;
    float M[2][3];
    float tx, ty, tz, xmul, ymul, fx, fy;
};

// oracle: compute_cov2d
__device__ __forceinline__ void compute_cov2d(const float* m, const float* S6, const float* view, float tanfovx,
                                              float tanfovy, int W, int H, Cov2D& o) {
    const float fx = float(W) / (2.0f * tanfovx);
    const float fy = float(H) / (2.0f * tanfovy);
    float tx = view[0] * m[0] + view[4] * m[1] + view[8] * m[2] + view[12];
    float ty = view[1] * m[0] + view[5] * m[1] + view[9] * m[2] + view[13];
    const float tz = view[2] * m[0] + view[6] * m[1] + view[10] * m[2] + view[14];
    const float limx = 1.3f * tanfovx;
    const float limy = 1.3f * tanfovy;
    const float txtz = tx / tz;
    const float tytz = ty / tz;
    o.xmul = (txtz < -limx || txtz > limx) ? 0.0f : 1.0f;
    o.ymul = (tytz < -limy || tytz > limy) ? 0.0f : 1.0f;
    // std::min(limx, std::max(-limx, v)) semantics (NaN -> -lim), written with explicit selects
    float cx = (-limx < txtz) ? txtz : -limx;
    cx = (cx < limx) ? cx : limx;
    float cy = (-limy < tytz) ? tytz : -limy;
    cy = (cy < limy) ? cy : limy;
    tx = cx * tz;
    ty = cy * tz;
    const float j00 = fx / tz;
    const float j02 = -(fx * tx) / (tz * tz);
    const float j11 = fy / tz;
    const float j12 = -(fy * ty) / (tz * tz);
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        o.M[0][k] = view[4 * k + 0] * j00 + view[4 * k + 2] * j02;
        o.M[1][k] = view[4 * k + 1] * j11 + view[4 * k + 2] * j12;
    }
    const float S[3][3] = {{S6[0], S6[1], S6[2]}, {S6[1], S6[3], S6[4]}, {S6[2], S6[4], S6[5]}};
    float A[2][3];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) A[i][j] = o.M[i][0] * S[0][j] + o.M[i][1] * S[1][j] + o.M[i][2] * S[2][j];
    const float c00 = A[0][0] * o.M[0][0] + A[0][1] * o.M[0][1] + A[0][2] * o.M[0][2];
    const float c01 = A[1][0] * o.M[0][0] + A[1][1] * o.M[0][1] + A[1][2] * o.M[0][2];
    const float c11 = A[1][0] * o.M[1][0] + A[1][1] * o.M[1][1] + A[1][2] * o.M[1][2];
    o.a = c00 + 0.3f;
    o.b = c01;
    o.c = c11 + 0.3f;
    o.tx = tx; o.ty = ty; o.tz = tz; o.fx = fx; o.fy = fy;
}

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
constexpr int kPreIters = 8;                 // sub-batches of 256 Gaussians per CTA (amortises the histogram flush)
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
// A CTA handles kPreIters * 256 Gaussians of one render: it counts its instances per tile in shared memory, reserves a
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
    const int i_lo = blockIdx.x * kPreIters * 256;
    if (use_hist) {
        for (int k = threadIdx.x; k < T; k += 256) s_cnt[k] = 0;
        __syncthreads();
        for (int it = 0; it < kPreIters; ++it) {
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
    for (int it = 0; it < kPreIters; ++it) {
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

// ------------------------------------------------------------------------------------------------ backward
struct PreBwdArgs {
    RenderGeom g;
    int render_base, num_renders;
    const float *means, *cov, *view, *proj;
    const int32_t* radii;
    const float* accum;          // [kAccumPlanes][Rc*N]
    size_t plane;                // Rc*N
    float *d_means3D, *d_cov3D, *d_colors, *d_opac, *d_means2D;
};

// oracle: preprocess_backward.  One thread per (subject, Gaussian); loops over the subject's views inside the
// chunk in view order and adds the sum to the per-subject gradients (deterministic).
__global__ void __launch_bounds__(256) preprocess_backward_kernel(PreBwdArgs a) {
    const int N = a.g.N, V = a.g.V;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const int b0 = a.render_base / V;
    const int b = b0 + blockIdx.y;
    const int r_lo = max(a.render_base, b * V);
    const int r_hi = min(a.render_base + a.num_renders, (b + 1) * V);
    if (r_lo >= r_hi) return;
    const size_t gi = size_t(b) * N + i;
    const float m[3] = {a.means[3 * gi], a.means[3 * gi + 1], a.means[3 * gi + 2]};
    float S6[6];
#pragma unroll
    for (int k = 0; k < 6; ++k) S6[k] = a.cov[6 * gi + k];
    float gm[3] = {0, 0, 0}, gcov[6] = {0, 0, 0, 0, 0, 0}, gcol[3] = {0, 0, 0}, gop = 0;
    for (int r = r_lo; r < r_hi; ++r) {
        const size_t oi = size_t(r - a.render_base) * N + i;
        const bool vis = a.radii[size_t(r) * N + i] > 0;
        if (a.d_means2D) {
            float* o = a.d_means2D + (size_t(r) * N + i) * 3;
            o[0] = vis ? a.accum[0 * a.plane + oi] : 0.0f;
            o[1] = vis ? a.accum[1 * a.plane + oi] : 0.0f;
            o[2] = 0.0f;
        }
        if (!vis) continue;
        const float g2x = a.accum[0 * a.plane + oi], g2y = a.accum[1 * a.plane + oi];
        const float dA = a.accum[2 * a.plane + oi], dB = a.accum[3 * a.plane + oi], dC = a.accum[4 * a.plane + oi];
        gop += a.accum[5 * a.plane + oi];
        gcol[0] += a.accum[6 * a.plane + oi]; gcol[1] += a.accum[7 * a.plane + oi]; gcol[2] += a.accum[8 * a.plane + oi];
        const float dz = a.accum[9 * a.plane + oi];
        const float* view = a.view + size_t(r) * 16;
        const float* proj = a.proj + size_t(r) * 16;
        Cov2D c2;
        compute_cov2d(m, S6, view, a.g.tanfovx, a.g.tanfovy, a.g.W, a.g.H, c2);
        const float ca = c2.a, cb = c2.b, cc = c2.c;
        const float denom = ca * cc - cb * cb;
        const float denom2inv = 1.0f / ((denom * denom) + 0.0000001f);
        float dL_da = 0, dL_db = 0, dL_dc = 0;
        const float(*M)[3] = c2.M;
        if (denom2inv != 0.0f) {
            dL_da = denom2inv * (-cc * cc * dA + 2.0f * cb * cc * dB + (denom - ca * cc) * dC);
            dL_dc = denom2inv * (-ca * ca * dC + 2.0f * ca * cb * dB + (denom - ca * cc) * dA);
            dL_db = denom2inv * 2.0f * (cb * cc * dA - (denom + 2.0f * cb * cb) * dB + ca * cb * dC);
            gcov[0] += M[0][0] * M[0][0] * dL_da + M[0][0] * M[1][0] * dL_db + M[1][0] * M[1][0] * dL_dc;
            gcov[3] += M[0][1] * M[0][1] * dL_da + M[0][1] * M[1][1] * dL_db + M[1][1] * M[1][1] * dL_dc;
            gcov[5] += M[0][2] * M[0][2] * dL_da + M[0][2] * M[1][2] * dL_db + M[1][2] * M[1][2] * dL_dc;
            gcov[1] += 2.0f * M[0][0] * M[0][1] * dL_da + (M[0][0] * M[1][1] + M[0][1] * M[1][0]) * dL_db + 2.0f * M[1][0] * M[1][1] * dL_dc;
            gcov[2] += 2.0f * M[0][0] * M[0][2] * dL_da + (M[0][0] * M[1][2] + M[0][2] * M[1][0]) * dL_db + 2.0f * M[1][0] * M[1][2] * dL_dc;
            gcov[4] += 2.0f * M[0][2] * M[0][1] * dL_da + (M[0][1] * M[1][2] + M[0][2] * M[1][1]) * dL_db + 2.0f * M[1][1] * M[1][2] * dL_dc;
        }
        const float S[3][3] = {{S6[0], S6[1], S6[2]}, {S6[1], S6[3], S6[4]}, {S6[2], S6[4], S6[5]}};
        float dM[2][3];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const float m0s = M[0][0] * S[k][0] + M[0][1] * S[k][1] + M[0][2] * S[k][2];
            const float m1s = M[1][0] * S[k][0] + M[1][1] * S[k][1] + M[1][2] * S[k][2];
            dM[0][k] = 2.0f * m0s * dL_da + m1s * dL_db;
            dM[1][k] = 2.0f * m1s * dL_dc + m0s * dL_db;
        }
        const float dJ00 = view[0] * dM[0][0] + view[4] * dM[0][1] + view[8] * dM[0][2];
        const float dJ02 = view[2] * dM[0][0] + view[6] * dM[0][1] + view[10] * dM[0][2];
        const float dJ11 = view[1] * dM[1][0] + view[5] * dM[1][1] + view[9] * dM[1][2];
        const float dJ12 = view[2] * dM[1][0] + view[6] * dM[1][1] + view[10] * dM[1][2];
        const float tz = 1.0f / c2.tz, tz2 = tz * tz, tz3 = tz2 * tz;
        const float dtx = c2.xmul * -c2.fx * tz2 * dJ02;
        const float dty = c2.ymul * -c2.fy * tz2 * dJ12;
        const float dtz = -c2.fx * tz2 * dJ00 - c2.fy * tz2 * dJ11 + (2.0f * c2.fx * c2.tx) * tz3 * dJ02 +
                          (2.0f * c2.fy * c2.ty) * tz3 * dJ12;
        float g0 = view[0] * dtx + view[1] * dty + view[2] * dtz;
        float g1 = view[4] * dtx + view[5] * dty + view[6] * dtz;
        float g2 = view[8] * dtx + view[9] * dty + view[10] * dtz;
        const float hx = proj[0] * m[0] + proj[4] * m[1] + proj[8] * m[2] + proj[12];
        const float hy = proj[1] * m[0] + proj[5] * m[1] + proj[9] * m[2] + proj[13];
        const float hw = proj[3] * m[0] + proj[7] * m[1] + proj[11] * m[2] + proj[15];
        const float mw = 1.0f / (hw + 0.0000001f);
        const float mul1 = hx * mw * mw, mul2 = hy * mw * mw;
        g0 += (proj[0] * mw - proj[3] * mul1) * g2x + (proj[1] * mw - proj[3] * mul2) * g2y;
        g1 += (proj[4] * mw - proj[7] * mul1) * g2x + (proj[5] * mw - proj[7] * mul2) * g2y;
        g2 += (proj[8] * mw - proj[11] * mul1) * g2x + (proj[9] * mw - proj[11] * mul2) * g2y;
        const float mul3 = view[2] * m[0] + view[6] * m[1] + view[10] * m[2] + view[14];
        g0 += (view[2] - view[3] * mul3) * dz;
        g1 += (view[6] - view[7] * mul3) * dz;
        g2 += (view[10] - view[11] * mul3) * dz;
        gm[0] += g0; gm[1] += g1; gm[2] += g2;
    }
    // chunks of one call run back to back on one stream: plain read-modify-write is race free and ordered
#pragma unroll
    for (int k = 0; k < 3; ++k) a.d_means3D[3 * gi + k] += gm[k];
#pragma unroll
    for (int k = 0; k < 6; ++k) a.d_cov3D[6 * gi + k] += gcov[k];
#pragma unroll
    for (int k = 0; k < 3; ++k) a.d_colors[3 * gi + k] += gcol[k];
    a.d_opac[gi] += gop;
}

__global__ void mark_visible_kernel(const float* __restrict__ means, int N, const float* __restrict__ view,
                                    uint8_t* __restrict__ visible) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const float z = view[2] * means[3 * i] + view[6] * means[3 * i + 1] + view[10] * means[3 * i + 2] + view[14];
    visible[i] = (z > 0.2f) ? 1 : 0;
}

// computeCov3D: Sigma = Rm diag((mod*s)^2) Rm^T, Rm from the unnormalised quaternion (r,x,y,z).
__device__ __forceinline__ void quat_rot(const float* q, float Rm[3][3]) {
    const float r = q[0], x = q[1], y = q[2], z = q[3];
    Rm[0][0] = 1.0f - 2.0f * (y * y + z * z); Rm[0][1] = 2.0f * (x * y - r * z); Rm[0][2] = 2.0f * (x * z + r * y);
    Rm[1][0] = 2.0f * (x * y + r * z); Rm[1][1] = 1.0f - 2.0f * (x * x + z * z); Rm[1][2] = 2.0f * (y * z - r * x);
    Rm[2][0] = 2.0f * (x * z - r * y); Rm[2][1] = 2.0f * (y * z + r * x); Rm[2][2] = 1.0f - 2.0f * (x * x + y * y);
}

__global__ void cov3d_kernel(const float* __restrict__ scales, const float* __restrict__ rots, float mod, int N,
                             float* __restrict__ cov6) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    float Rm[3][3];
    quat_rot(rots + 4 * i, Rm);
    const float s[3] = {mod * scales[3 * i], mod * scales[3 * i + 1], mod * scales[3 * i + 2]};
    float L[3][3];
#pragma unroll
    for (int j = 0; j < 3; ++j)
#pragma unroll
        for (int k = 0; k < 3; ++k) L[j][k] = Rm[j][k] * s[k];
    float* o = cov6 + 6 * i;
    o[0] = L[0][0] * L[0][0] + L[0][1] * L[0][1] + L[0][2] * L[0][2];
    o[1] = L[0][0] * L[1][0] + L[0][1] * L[1][1] + L[0][2] * L[1][2];
    o[2] = L[0][0] * L[2][0] + L[0][1] * L[2][1] + L[0][2] * L[2][2];
    o[3] = L[1][0] * L[1][0] + L[1][1] * L[1][1] + L[1][2] * L[1][2];
    o[4] = L[1][0] * L[2][0] + L[1][1] * L[2][1] + L[1][2] * L[2][2];
    o[5] = L[2][0] * L[2][0] + L[2][1] * L[2][1] + L[2][2] * L[2][2];
}

__global__ void cov3d_backward_kernel(const float* __restrict__ scales, const float* __restrict__ rots, float mod,
                                      int N, const float* __restrict__ dcov, float* __restrict__ dscales,
                                      float* __restrict__ drots) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const float* q = rots + 4 * i;
    float Rm[3][3];
    quat_rot(q, Rm);
    const float s[3] = {mod * scales[3 * i], mod * scales[3 * i + 1], mod * scales[3 * i + 2]};
    const float* d = dcov + 6 * i;
    // symmetric gradient matrix: the stored off-diagonal value appears twice in Sigma
    const float G[3][3] = {{d[0], 0.5f * d[1], 0.5f * d[2]}, {0.5f * d[1], d[3], 0.5f * d[4]}, {0.5f * d[2], 0.5f * d[4], d[5]}};
    float L[3][3], dLm[3][3];
#pragma unroll
    for (int j = 0; j < 3; ++j)
#pragma unroll
        for (int k = 0; k < 3; ++k) L[j][k] = Rm[j][k] * s[k];
    // Sigma = L L^T  ->  dL/dL = 2 G L
#pragma unroll
    for (int j = 0; j < 3; ++j)
#pragma unroll
        for (int k = 0; k < 3; ++k) dLm[j][k] = 2.0f * (G[j][0] * L[0][k] + G[j][1] * L[1][k] + G[j][2] * L[2][k]);
    float dR[3][3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        dscales[3 * i + k] = mod * (dLm[0][k] * Rm[0][k] + dLm[1][k] * Rm[1][k] + dLm[2][k] * Rm[2][k]);
#pragma unroll
        for (int j = 0; j < 3; ++j) dR[j][k] = dLm[j][k] * s[k];
    }
    const float r = q[0], x = q[1], y = q[2], z = q[3];
    drots[4 * i + 0] = 2.0f * (-z * dR[0][1] + y * dR[0][2] + z * dR[1][0] - x * dR[1][2] - y * dR[2][0] + x * dR[2][1]);
    drots[4 * i + 1] = 2.0f * (y * dR[0][1] + z * dR[0][2] + y * dR[1][0] - 2.0f * x * dR[1][1] - r * dR[1][2] +
                               z * dR[2][0] + r * dR[2][1] - 2.0f * x * dR[2][2]);
    drots[4 * i + 2] = 2.0f * (-2.0f * y * dR[0][0] + x * dR[0][1] + r * dR[0][2] + x * dR[1][0] + z * dR[1][2] -
                               r * dR[2][0] + z * dR[2][1] - 2.0f * y * dR[2][2]);
    drots[4 * i + 3] = 2.0f * (-2.0f * z * dR[0][0] - r * dR[0][1] + x * dR[0][2] + r * dR[1][0] - 2.0f * z * dR[1][1] +
                               y * dR[1][2] + x * dR[2][0] + y * dR[2][1]);
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
    const int per_cta = 256 * kPreIters;
    dim3 grid((c.g.N + per_cta - 1) / per_cta, c.num_renders);
    const size_t smem = c.g.num_tiles <= kHistTiles ? size_t(c.g.num_tiles) * 8 : 0;
    scatter_kernel<<<grid, 256, smem, c.stream>>>(a);
    return cudaGetLastError();
}

cudaError_t launch_preprocess_backward(const ChunkCtx& c, const SgrBackwardArgs& b) {
    PreBwdArgs a;
    a.g = c.g; a.render_base = c.render_base; a.num_renders = c.num_renders;
    a.means = c.p->means3D; a.cov = c.p->cov3D; a.view = c.p->viewmatrix; a.proj = c.p->projmatrix;
    a.radii = b.radii; a.accum = c.accum;
    a.plane = size_t(c.num_renders) * c.g.N;
    a.d_means3D = b.dL_dmeans3D; a.d_cov3D = b.dL_dcov3D; a.d_colors = b.dL_dcolors; a.d_opac = b.dL_dopacities;
    a.d_means2D = b.dL_dmeans2D;
    const int b0 = c.render_base / c.g.V, b1 = (c.render_base + c.num_renders - 1) / c.g.V;
    dim3 grid((c.g.N + 255) / 256, b1 - b0 + 1);
    preprocess_backward_kernel<<<grid, 256, 0, c.stream>>>(a);
    return cudaGetLastError();
}

cudaError_t launch_mark_visible(const float* means3D, int N, const float* view, uint8_t* visible, cudaStream_t s) {
    if (N <= 0) return cudaSuccess;
    mark_visible_kernel<<<(N + 255) / 256, 256, 0, s>>>(means3D, N, view, visible);
    return cudaGetLastError();
}

cudaError_t launch_cov3d_from_scale_rot(const float* scales, const float* rots, float mod, int N, float* cov6,
                                        cudaStream_t s) {
    if (N <= 0) return cudaSuccess;
    cov3d_kernel<<<(N + 255) / 256, 256, 0, s>>>(scales, rots, mod, N, cov6);
    return cudaGetLastError();
}

cudaError_t launch_cov3d_from_scale_rot_backward(const float* scales, const float* rots, float mod, int N,
                                                 const float* dcov, float* dscales, float* drots, cudaStream_t s) {
    if (N <= 0) return cudaSuccess;
    cov3d_backward_kernel<<<(N + 255) / 256, 256, 0, s>>>(scales, rots, mod, N, dcov, dscales, drots);
    return cudaGetLastError();
}

}  // namespace sgr
