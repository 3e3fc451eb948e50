// sgr_cov2d.cuh — EWA projection of a 3D covariance to the image plane (oracle: compute_cov2d), shared by the forward
// preprocess kernel (compiled with --fmad=false: bit-exact radii / tile rectangles) and the per-Gaussian backward
// (compiled with FMA contraction: gradients are compared with a tolerance).
#pragma once
#include "sgr_common.cuh"

namespace sgr {

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

}  // namespace sgr
