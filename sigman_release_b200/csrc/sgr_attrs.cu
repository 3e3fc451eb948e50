// sgr_attrs.cu — per-Gaussian attribute kernels on either side of the rasteriser (SURVEY.md 8f #2 and the optional
// colour path of the upstream API).  Compiled WITH FMA contraction (compared with a tolerance, not index-determining).
//
//  * prep_cov3d (+ backward): the per-subject preparation of /root/reference/core/gaussians/gs.py:69-73 in ONE kernel —
//    scale = (s + 1) * sqrt(max(d2, 1e-7)) with the kNN factor detached (gs.py:70-72), Sigma = Rm diag(scale^2) Rm^T
//    (get_covariance, gs.py:17-23) packed as (xx, xy, xz, yy, yz, zz) (strip_lowerdiag, gs.py:29-38).  The reference
//    runs this as zeros + 3 strided writes + pow + 2 bmm + 6 strided copies per subject with the autograd graph kept.
//    kBf16 reproduces the operand / result rounding of the two bmm's under accelerate's bf16 autocast (parity studies).
//  * sh_colors (+ backward): upstream computeColorFromSH / computeColorFromSH backward: real SH basis
//    up to degree 3 on dir = normalize(mean - campos), + 0.5, clamp at 0 with the clamp recorded.  The contraction is a
//    per-Gaussian [1 x K] . [K x 3] product with a different basis row per Gaussian — not a dense GEMM, so it stays on
//    the FMA pipes (no tensor cores).  SIGMAN itself passes colors_precomp (gs.py:91,102).
#include <cuda_bf16.h>

#include "sgr_common.cuh"

namespace sgr {
namespace {

__device__ __forceinline__ float bf16_round(float v) { return __bfloat162float(__float2bfloat16_rn(v)); }

template <bool kBf16>
__global__ void __launch_bounds__(256) prep_cov3d_kernel(const float* __restrict__ s_raw, const float* __restrict__ rot,
                                                         const float* __restrict__ dist2, long long n,
                                                         float* __restrict__ cov6) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float sg = sqrtf(fmaxf(dist2[i], 0.0000001f));
    float e[3], Rm[9];
#pragma unroll
    for (int k = 0; k < 3; ++k) { const float sc = (s_raw[3 * i + k] + 1.0f) * sg; e[k] = sc * sc; }
#pragma unroll
    for (int k = 0; k < 9; ++k) Rm[k] = rot[9 * i + k];
    if (kBf16) {
#pragma unroll
        for (int k = 0; k < 3; ++k) e[k] = bf16_round(e[k]);
#pragma unroll
        for (int k = 0; k < 9; ++k) Rm[k] = bf16_round(Rm[k]);
    }
    float A[9];                                   // Rm diag(e)
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int k = 0; k < 3; ++k) A[3 * r + k] = kBf16 ? bf16_round(Rm[3 * r + k] * e[k]) : Rm[3 * r + k] * e[k];
    const int ia[6] = {0, 0, 0, 1, 1, 2}, ib[6] = {0, 1, 2, 1, 2, 2};
#pragma unroll
    for (int q = 0; q < 6; ++q) {
        const float v = A[3 * ia[q]] * Rm[3 * ib[q]] + A[3 * ia[q] + 1] * Rm[3 * ib[q] + 1] + A[3 * ia[q] + 2] * Rm[3 * ib[q] + 2];
        cov6[6 * i + q] = kBf16 ? bf16_round(v) : v;
    }
}

// L depends on Sigma_ab (a <= b) = sum_k R_ak e_k R_bk.  With the symmetrised gradient Gs (Gs_aa = g_aa,
// Gs_ab = g_ab / 2):  dL/dR_ak = 2 e_k sum_b Gs_ab R_bk,  dL/de_k = sum_ab Gs_ab R_ak R_bk,
// de_k/ds_k = 2 (s_k + 1) sigma^2.  (bf16 mode: the same formulas on the rounded operands.)
template <bool kBf16>
__global__ void __launch_bounds__(256) prep_cov3d_backward_kernel(const float* __restrict__ s_raw,
                                                                  const float* __restrict__ rot,
                                                                  const float* __restrict__ dist2, long long n,
                                                                  const float* __restrict__ dcov,
                                                                  float* __restrict__ ds_raw, float* __restrict__ drot) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float d2 = fmaxf(dist2[i], 0.0000001f);
    float e[3], Rm[9], s1[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) { s1[k] = s_raw[3 * i + k] + 1.0f; e[k] = s1[k] * s1[k] * d2; }
#pragma unroll
    for (int k = 0; k < 9; ++k) Rm[k] = rot[9 * i + k];
    if (kBf16) {
#pragma unroll
        for (int k = 0; k < 3; ++k) e[k] = bf16_round(e[k]);
#pragma unroll
        for (int k = 0; k < 9; ++k) Rm[k] = bf16_round(Rm[k]);
    }
    const float* g = dcov + 6 * i;
    const float Gs[9] = {g[0], 0.5f * g[1], 0.5f * g[2], 0.5f * g[1], g[3], 0.5f * g[4], 0.5f * g[2], 0.5f * g[4], g[5]};
    float GR[9];                                  // Gs Rm
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int k = 0; k < 3; ++k) GR[3 * a + k] = Gs[3 * a] * Rm[k] + Gs[3 * a + 1] * Rm[3 + k] + Gs[3 * a + 2] * Rm[6 + k];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const float de = Rm[k] * GR[k] + Rm[3 + k] * GR[3 + k] + Rm[6 + k] * GR[6 + k];
        ds_raw[3 * i + k] = de * 2.0f * s1[k] * d2;
#pragma unroll
        for (int a = 0; a < 3; ++a) drot[9 * i + 3 * a + k] = 2.0f * e[k] * GR[3 * a + k];
    }
}

// ------------------------------------------------------------------------------------------------ spherical harmonics
constexpr float kSH0 = 0.28209479177387814f, kSH1 = 0.4886025119029199f;
__device__ __constant__ float kSH2[5] = {1.0925484305920792f, -1.0925484305920792f, 0.31539156525252005f,
                                         -1.0925484305920792f, 0.5462742152960396f};
__device__ __constant__ float kSH3[7] = {-0.5900435899266435f, 2.890611442640554f, -0.4570457994644658f,
                                         0.3731763325901154f, -0.4570457994644658f, 1.445305721320277f,
                                         -0.5900435899266435f};

// basis b[k] and (optionally) its gradient w.r.t. the unit direction (x, y, z)
template <bool kGrad>
__device__ __forceinline__ void sh_basis(int deg, float x, float y, float z, float* b, float* bx, float* by, float* bz) {
    b[0] = kSH0;
    if (kGrad) { bx[0] = by[0] = bz[0] = 0.0f; }
    if (deg < 1) return;
    b[1] = -kSH1 * y; b[2] = kSH1 * z; b[3] = -kSH1 * x;
    if (kGrad) {
        bx[1] = 0; by[1] = -kSH1; bz[1] = 0;
        bx[2] = 0; by[2] = 0; bz[2] = kSH1;
        bx[3] = -kSH1; by[3] = 0; bz[3] = 0;
    }
    if (deg < 2) return;
    const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
    b[4] = kSH2[0] * xy; b[5] = kSH2[1] * yz; b[6] = kSH2[2] * (2.0f * zz - xx - yy); b[7] = kSH2[3] * xz;
    b[8] = kSH2[4] * (xx - yy);
    if (kGrad) {
        bx[4] = kSH2[0] * y; by[4] = kSH2[0] * x; bz[4] = 0;
        bx[5] = 0; by[5] = kSH2[1] * z; bz[5] = kSH2[1] * y;
        bx[6] = -2.0f * kSH2[2] * x; by[6] = -2.0f * kSH2[2] * y; bz[6] = 4.0f * kSH2[2] * z;
        bx[7] = kSH2[3] * z; by[7] = 0; bz[7] = kSH2[3] * x;
        bx[8] = 2.0f * kSH2[4] * x; by[8] = -2.0f * kSH2[4] * y; bz[8] = 0;
    }
    if (deg < 3) return;
    b[9] = kSH3[0] * y * (3.0f * xx - yy); b[10] = kSH3[1] * xy * z; b[11] = kSH3[2] * y * (4.0f * zz - xx - yy);
    b[12] = kSH3[3] * z * (2.0f * zz - 3.0f * xx - 3.0f * yy); b[13] = kSH3[4] * x * (4.0f * zz - xx - yy);
    b[14] = kSH3[5] * z * (xx - yy); b[15] = kSH3[6] * x * (xx - 3.0f * yy);
    if (kGrad) {
        bx[9] = kSH3[0] * 6.0f * xy; by[9] = kSH3[0] * (3.0f * xx - 3.0f * yy); bz[9] = 0;
        bx[10] = kSH3[1] * yz; by[10] = kSH3[1] * xz; bz[10] = kSH3[1] * xy;
        bx[11] = kSH3[2] * -2.0f * xy; by[11] = kSH3[2] * (4.0f * zz - xx - 3.0f * yy); bz[11] = kSH3[2] * 8.0f * yz;
        bx[12] = kSH3[3] * -6.0f * xz; by[12] = kSH3[3] * -6.0f * yz; bz[12] = kSH3[3] * (6.0f * zz - 3.0f * xx - 3.0f * yy);
        bx[13] = kSH3[4] * (4.0f * zz - 3.0f * xx - yy); by[13] = kSH3[4] * -2.0f * xy; bz[13] = kSH3[4] * 8.0f * xz;
        bx[14] = kSH3[5] * 2.0f * xz; by[14] = kSH3[5] * -2.0f * yz; bz[14] = kSH3[5] * (xx - yy);
        bx[15] = kSH3[6] * (3.0f * xx - 3.0f * yy); by[15] = kSH3[6] * -6.0f * xy; bz[15] = 0;
    }
}

__global__ void __launch_bounds__(128) sh_colors_kernel(const float* __restrict__ means, const float* __restrict__ shs,
                                                        const float* __restrict__ campos, int N, int deg, int max_coeffs,
                                                        float* __restrict__ colors, uint8_t* __restrict__ clamped) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const float dx = means[3 * i] - campos[0], dy = means[3 * i + 1] - campos[1], dz = means[3 * i + 2] - campos[2];
    const float inv = 1.0f / sqrtf(dx * dx + dy * dy + dz * dz);
    float b[16];
    sh_basis<false>(deg, dx * inv, dy * inv, dz * inv, b, nullptr, nullptr, nullptr);
    const int K = (deg + 1) * (deg + 1);
    const float* sh = shs + size_t(i) * max_coeffs * 3;
    float c[3] = {0.0f, 0.0f, 0.0f};
    for (int k = 0; k < K; ++k) {
        c[0] = fmaf(b[k], sh[3 * k], c[0]); c[1] = fmaf(b[k], sh[3 * k + 1], c[1]); c[2] = fmaf(b[k], sh[3 * k + 2], c[2]);
    }
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
        const float v = c[ch] + 0.5f;
        clamped[3 * i + ch] = v < 0.0f ? 1 : 0;
        colors[3 * i + ch] = v < 0.0f ? 0.0f : v;
    }
}

__global__ void __launch_bounds__(128) sh_colors_backward_kernel(const float* __restrict__ means,
                                                                 const float* __restrict__ shs,
                                                                 const float* __restrict__ campos, int N, int deg,
                                                                 int max_coeffs, const uint8_t* __restrict__ clamped,
                                                                 const float* __restrict__ dcolors,
                                                                 float* __restrict__ dshs, float* __restrict__ dmeans) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const float vx = means[3 * i] - campos[0], vy = means[3 * i + 1] - campos[1], vz = means[3 * i + 2] - campos[2];
    const float inv = 1.0f / sqrtf(vx * vx + vy * vy + vz * vz);
    const float x = vx * inv, y = vy * inv, z = vz * inv;
    float b[16], bx[16], by[16], bz[16];
    sh_basis<true>(deg, x, y, z, b, bx, by, bz);
    const int K = (deg + 1) * (deg + 1);
    const float* sh = shs + size_t(i) * max_coeffs * 3;
    float* dsh = dshs + size_t(i) * max_coeffs * 3;
    float g[3];
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) g[ch] = clamped[3 * i + ch] ? 0.0f : dcolors[3 * i + ch];
    float ddx = 0.0f, ddy = 0.0f, ddz = 0.0f;        // dL / d(unit direction)
    for (int k = 0; k < K; ++k) {
        dsh[3 * k] = b[k] * g[0]; dsh[3 * k + 1] = b[k] * g[1]; dsh[3 * k + 2] = b[k] * g[2];
        const float w = sh[3 * k] * g[0] + sh[3 * k + 1] * g[1] + sh[3 * k + 2] * g[2];
        ddx = fmaf(bx[k], w, ddx); ddy = fmaf(by[k], w, ddy); ddz = fmaf(bz[k], w, ddz);
    }
    for (int k = K; k < max_coeffs; ++k) { dsh[3 * k] = 0.0f; dsh[3 * k + 1] = 0.0f; dsh[3 * k + 2] = 0.0f; }
    // through the normalisation: d dir / d v = (I - dir dir^T) / |v|
    const float dot = ddx * x + ddy * y + ddz * z;
    dmeans[3 * i] = (ddx - dot * x) * inv; dmeans[3 * i + 1] = (ddy - dot * y) * inv; dmeans[3 * i + 2] = (ddz - dot * z) * inv;
}

}  // namespace

cudaError_t launch_prep_cov3d(const float* s_raw, const float* rot, const float* dist2, long long n, bool bf16,
                              float* cov6, cudaStream_t s) {
    if (n <= 0) return cudaSuccess;
    const unsigned int grid = unsigned((n + 255) / 256);
    if (bf16) prep_cov3d_kernel<true><<<grid, 256, 0, s>>>(s_raw, rot, dist2, n, cov6);
    else prep_cov3d_kernel<false><<<grid, 256, 0, s>>>(s_raw, rot, dist2, n, cov6);
    return cudaGetLastError();
}

cudaError_t launch_prep_cov3d_backward(const float* s_raw, const float* rot, const float* dist2, long long n, bool bf16,
                                       const float* dcov, float* ds_raw, float* drot, cudaStream_t s) {
    if (n <= 0) return cudaSuccess;
    const unsigned int grid = unsigned((n + 255) / 256);
    if (bf16) prep_cov3d_backward_kernel<true><<<grid, 256, 0, s>>>(s_raw, rot, dist2, n, dcov, ds_raw, drot);
    else prep_cov3d_backward_kernel<false><<<grid, 256, 0, s>>>(s_raw, rot, dist2, n, dcov, ds_raw, drot);
    return cudaGetLastError();
}

cudaError_t launch_sh_colors(const float* means, const float* shs, const float* campos, int N, int deg, int max_coeffs,
                             float* colors, uint8_t* clamped, cudaStream_t s) {
    if (N <= 0) return cudaSuccess;
    sh_colors_kernel<<<(N + 127) / 128, 128, 0, s>>>(means, shs, campos, N, deg, max_coeffs, colors, clamped);
    return cudaGetLastError();
}

cudaError_t launch_sh_colors_backward(const float* means, const float* shs, const float* campos, int N, int deg,
                                      int max_coeffs, const uint8_t* clamped, const float* dcolors, float* dshs,
                                      float* dmeans, cudaStream_t s) {
    if (N <= 0) return cudaSuccess;
    sh_colors_backward_kernel<<<(N + 127) / 128, 128, 0, s>>>(means, shs, campos, N, deg, max_coeffs, clamped, dcolors,
                                                              dshs, dmeans);
    return cudaGetLastError();
}

}  // namespace sgr
