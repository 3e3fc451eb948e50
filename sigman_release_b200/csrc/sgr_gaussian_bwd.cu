// sgr_gaussian_bwd.cu — per-Gaussian kernels that are NOT index-determining: the backward of the projection / conic
// (upstream computeCov2DCUDA + preprocessCUDA backward), mark_visible, and the optional scale/rotation
// -> cov3D path with its backward (upstream computeCov3D).  Compiled WITH FMA contraction (results are compared with
// a tolerance), unlike sgr_preprocess.cu.
#include "sgr_common.cuh"
#include "sgr_cov2d.cuh"

namespace sgr {
namespace {

// ------------------------------------------------------------------------------------------------ backward
struct PreBwdArgs {
    RenderGeom g;
    int render_base, num_renders;
    const float *means, *cov, *view, *proj;
    const int32_t* radii;
    const float* accum;          // [Rc*N][kAccumStride]
    size_t plane;                // Rc*N
    float *d_means3D, *d_cov3D, *d_colors, *d_opac, *d_means2D;
};

// oracle: preprocess_backward.  A CTA is 32 Gaussians (lanes: coalesced attribute / accumulator reads) x kPreBwdSlots
// view slots (warps) of one subject: slot y handles the subject's views r_lo + y, r_lo + y + slots, ... inside the
// chunk, the slots are then summed in slot order in shared memory (deterministic) and slot 0 writes the per-subject
// gradients — a plain store in the chunk that holds the subject's first view, read-modify-write in later chunks
// (chunks of one call run back to back on one stream), so the outputs need no zero fill.
constexpr int kPreBwdSlots = 8;
constexpr int kPreBwdVals = 13;              // means3D 3, cov3D 6, colors 3, opacity 1

#ifndef SGR_PREBWD_MIN_CTAS
#define SGR_PREBWD_MIN_CTAS 3
#endif
__global__ void __launch_bounds__(32 * kPreBwdSlots, SGR_PREBWD_MIN_CTAS) preprocess_backward_kernel(PreBwdArgs a) {
    __shared__ float s_part[kPreBwdSlots][kPreBwdVals][32];
    const int N = a.g.N, V = a.g.V;
    const int lane = threadIdx.x, slot = threadIdx.y;
    const int i_raw = blockIdx.x * 32 + lane;
    const bool live = i_raw < N;
    const int i = live ? i_raw : N - 1;
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
    for (int r = r_lo + slot; r < r_hi; r += kPreBwdSlots) {
        const size_t oi = size_t(r - a.render_base) * N + i;
        // all loads of the iteration are issued together (one exposed memory latency): nearly every Gaussian of a
        // framed subject is visible, so waiting for the radius before fetching the accumulators only serialises them
        const int rad = a.radii[size_t(r) * N + i];
        float acc[kAccumStride];
        {
            const float4* row = reinterpret_cast<const float4*>(a.accum + oi * kAccumStride);
            const float4 v0 = row[0], v1 = row[1], v2 = row[2];
            acc[0] = v0.x; acc[1] = v0.y; acc[2] = v0.z; acc[3] = v0.w; acc[4] = v1.x; acc[5] = v1.y; acc[6] = v1.z;
            acc[7] = v1.w; acc[8] = v2.x; acc[9] = v2.y; acc[10] = v2.z; acc[11] = v2.w;
        }
        const bool vis = rad > 0;
        if (a.d_means2D && live) {
            float* o = a.d_means2D + (size_t(r) * N + i) * 3;
            o[0] = vis ? acc[0] : 0.0f;
            o[1] = vis ? acc[1] : 0.0f;
            o[2] = 0.0f;
        }
        if (!vis) continue;
        const float g2x = acc[0], g2y = acc[1];
        const float dA = acc[2], dB = acc[3], dC = acc[4];
        gop += acc[5];
        gcol[0] += acc[6]; gcol[1] += acc[7]; gcol[2] += acc[8];
        const float dz = acc[9];
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
    float vals[kPreBwdVals] = {gm[0], gm[1], gm[2], gcov[0], gcov[1], gcov[2], gcov[3], gcov[4], gcov[5],
                               gcol[0], gcol[1], gcol[2], gop};
#pragma unroll
    for (int k = 0; k < kPreBwdVals; ++k) s_part[slot][k][lane] = vals[k];
    __syncthreads();
    if (!live) return;
    // the 13 per-Gaussian sums are spread over the CTA's warps: warp y sums values y and y + 8 over the view slots
    // (in slot order: deterministic) and stores them
    const bool first = r_lo == b * V;            // this chunk holds the subject's first view: store, else accumulate
#pragma unroll
    for (int kk = 0; kk < 2; ++kk) {
        const int k = slot + kk * kPreBwdSlots;
        if (k >= kPreBwdVals) break;
        float v = s_part[0][k][lane];
#pragma unroll
        for (int y = 1; y < kPreBwdSlots; ++y) v += s_part[y][k][lane];
        float* dst = k < 3 ? a.d_means3D + 3 * gi + k
                   : k < 9 ? a.d_cov3D + 6 * gi + (k - 3)
                   : k < 12 ? a.d_colors + 3 * gi + (k - 9)
                            : a.d_opac + gi;
        *dst = (first ? 0.0f : *dst) + v;
    }
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

cudaError_t launch_preprocess_backward(const ChunkCtx& c, const SgrBackwardArgs& b) {
    PreBwdArgs a;
    a.g = c.g; a.render_base = c.render_base; a.num_renders = c.num_renders;
    a.means = c.p->means3D; a.cov = c.p->cov3D; a.view = c.p->viewmatrix; a.proj = c.p->projmatrix;
    a.radii = b.radii; a.accum = c.accum;
    a.plane = size_t(c.num_renders) * c.g.N;
    a.d_means3D = b.dL_dmeans3D; a.d_cov3D = b.dL_dcov3D; a.d_colors = b.dL_dcolors; a.d_opac = b.dL_dopacities;
    a.d_means2D = b.dL_dmeans2D;
    const int b0 = c.render_base / c.g.V, b1 = (c.render_base + c.num_renders - 1) / c.g.V;
    dim3 grid((c.g.N + 31) / 32, b1 - b0 + 1), block(32, kPreBwdSlots);
    preprocess_backward_kernel<<<grid, block, 0, c.stream>>>(a);
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
