// =====================================================================================
// sgr_oracle.cpp — CPU ORACLE for the Gaussian-splat rasteriser hot path.
//
// *** TEST INFRASTRUCTURE ONLY. *** Only tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs may load this library, and only as the checker or as
// the reported CPU baseline.  The product (sigman_release_b200/) never imports or links it.
//
// *** PARITY UNPINNED. ***  The arithmetic of this path is NOT in the reference tree: SIGMAN
// only calls the third-party `diff_gaussian_rasterization` package (ashawkey fork of
// graphdeco-inria/diff-gaussian-rasterization, unpinned, /root/reference/README.md:19-21;
// call sites /root/reference/core/gaussians/gs.py:8-11,82-106) and `simple_knn`
// (/root/reference/README.md:23, call site gs.py:70).  Neither source is available here and
// the reference has no tests or golden vectors.  This file therefore restates the PUBLISHED
// algorithm (Kerbl et al., "3D Gaussian Splatting for Real-Time Radiance Field Rendering",
// SIGGRAPH 2023, sections 4-6 + appendix; depth/alpha outputs as in the ashawkey fork) following
// SURVEY.md section 3.4 / Appendix A.  It is cross-checked in tests/ against hand-derived
// known answers, a dense PyTorch autograd re-derivation and fp64 finite differences — not
// against an upstream binary.
//
// Conventions that make GPU parity checkable bit-for-bit (documented in DESIGN.md):
//   * fp32 instantiation: every expression is evaluated in C source order with IEEE-754
//     binary32 round-to-nearest operations; a fused multiply-add is used exactly where the
//     source says fma() and nowhere else (build with -ffp-contract=off, no -ffast-math).
//     Upstream is built by nvcc with FMA contraction in a compiler-chosen pattern that cannot
//     be reproduced without its binary; the per-pixel blend expressions below fix one such
//     pattern (gauss_power / blend_forward) so that CPU and GPU agree bit for bit:
//         power  = fma(dx, (-A/2)*dx, dy * fma(-C/2, dy, (-B)*dx))
//         test_T = fma(-alpha, T, T)
//         w = alpha*T;  C = fma(c, w, C);  D = fma(z, w, D);  Wt = Wt + w
//   * exp(): upstream calls CUDA's expf (<= 2 ulp, built on MUFU.EX2, not reproducible on a CPU).
//     The oracle uses exp_spec() below — a fixed sequence of IEEE fp32 operations accurate to
//     ~1 ulp — and the CUDA kernels execute the same sequence.
//   * float -> int conversions saturate and map NaN to 0 (CUDA cvt.rzi semantics).
//   * ndc2Pix is evaluated in double precision (upstream's literals 1.0 / 0.5 are doubles).
//   * fp64 instantiation: same code with real = double and std::exp; used only for finite-
//     difference validation of the analytic backward.
// =====================================================================================
#include <algorithm>
#include <climits>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

constexpr int kBlockX = 16;
constexpr int kBlockY = 16;

// ------------------------------------------------------------------ exp specification
// exp_spec(x) for fp32: n = rint(x*log2e); r = x - n*ln2 (two-step Cody-Waite with fma);
// degree-7 Taylor polynomial of e^r in Horner form with fma; scale by 2^n through the exponent
// field.  Inputs below -87 return 0 (the result would be subnormal; callers reject it anyway),
// inputs above 88 return +inf.  NaN propagates.
// g_exp_mode = 1 swaps the fp32 exponential for the C library's expf (correctly rounded in glibc): the honest
// statement of how far the fixed exp_spec sequence — and the kernels' SFU exponential — sit from "the" exponential
// that an upstream build evaluates (tests/test_oracle.py, tests/test_gpu_fast_exp.py count the pixels whose
// alpha >= 1/255 / T >= 1e-4 branch flips).
int g_exp_mode = 0;
inline float exp_spec(float x) {
    if (g_exp_mode == 1) return std::exp(x);
    if (x != x) return x;
    if (x < -87.0f) return 0.0f;
    if (x > 88.0f) return INFINITY;
    const float t = x * 1.44269502162933349609375f;          // log2(e) rounded to fp32
    const float n = rintf(t);                                 // round-half-even
    float r = fmaf(n, -0.693145751953125f, x);                // ln2 high part (exact product)
    r = fmaf(n, -1.428606765330187045e-06f, r);               // ln2 low part
    float p = 1.98412701e-4f;                                 // 1/5040
    p = fmaf(p, r, 1.38888892e-3f);                           // 1/720
    p = fmaf(p, r, 8.33333377e-3f);                           // 1/120
    p = fmaf(p, r, 4.16666679e-2f);                           // 1/24
    p = fmaf(p, r, 1.66666672e-1f);                           // 1/6
    p = fmaf(p, r, 0.5f);
    p = fmaf(p, r, 1.0f);
    p = fmaf(p, r, 1.0f);
    int32_t bits;
    std::memcpy(&bits, &p, 4);
    bits += static_cast<int32_t>(n) << 23;                    // p in [0.70, 1.42], n in [-126, 127]
    float out;
    std::memcpy(&out, &bits, 4);
    return out;
}
inline double exp_spec(double x) { return std::exp(x); }

// CUDA cvt.rzi.s32.f32 semantics (C leaves out-of-range conversions undefined).
template <typename R>
inline int f2i_rz_sat(R v) {
    if (v != v) return 0;
    if (v >= R(2147483648.0)) return INT_MAX;
    if (v <= R(-2147483648.0)) return INT_MIN;
    return static_cast<int>(v);
}

inline float fma_r(float a, float b, float c) { return fmaf(a, b, c); }
inline double fma_r(double a, double b, double c) { return std::fma(a, b, c); }

// power = -1/2 (A dx^2 + C dy^2) - B dx dy in the fixed FMA pattern of the spec (header).  The halvings and the
// negation are exact, so hA, hC, nB carry exactly the information of the conic.
template <typename R>
inline R gauss_power(R A, R B, R C, R dx, R dy) {
    const R hA = R(-0.5f) * A, hC = R(-0.5f) * C, nB = -B;
    return fma_r(dx, hA * dx, dy * fma_r(hC, dy, nB * dx));
}

inline uint32_t float_bits(float f) {
    uint32_t u;
    std::memcpy(&u, &f, 4);
    return u;
}

template <typename R>
struct Cov2DTerms {
    R a, b, c;              // dilated 2D covariance (a = cov00 + 0.3, b = cov01, c = cov11 + 0.3)
    R M[2][3];              // M = J * R_w2c  (upper 2x3 of upstream's T = W * J, transposed)
    R tx, ty, tz;           // clamped view-space mean
    R xmul, ymul;           // 0 where the +-1.3*tanfov clamp was active (backward only)
    R fx, fy;
};

// SURVEY.md A.2 "computeCov2D": EWA projection of the 3D covariance.
template <typename R>
inline void compute_cov2d(const R* mean, const R* cov3D, const R* view, R tanfovx, R tanfovy,
                          int W, int H, Cov2DTerms<R>& o) {
    const R fx = R(W) / (R(2.0f) * tanfovx);
    const R fy = R(H) / (R(2.0f) * tanfovy);
    R tx = view[0] * mean[0] + view[4] * mean[1] + view[8] * mean[2] + view[12];
    R ty = view[1] * mean[0] + view[5] * mean[1] + view[9] * mean[2] + view[13];
    R tz = view[2] * mean[0] + view[6] * mean[1] + view[10] * mean[2] + view[14];
    const R limx = R(1.3f) * tanfovx;
    const R limy = R(1.3f) * tanfovy;
    const R txtz = tx / tz;
    const R tytz = ty / tz;
    o.xmul = (txtz < -limx || txtz > limx) ? R(0) : R(1);
    o.ymul = (tytz < -limy || tytz > limy) ? R(0) : R(1);
    tx = std::min(limx, std::max(-limx, txtz)) * tz;
    ty = std::min(limy, std::max(-limy, tytz)) * tz;
    const R j00 = fx / tz;
    const R j02 = -(fx * tx) / (tz * tz);
    const R j11 = fy / tz;
    const R j12 = -(fy * ty) / (tz * tz);
    // R_w2c[r][c] = view[4c + r].  M[0][k] = j00*R[0][k] + j02*R[2][k]; M[1][k] = j11*R[1][k] + j12*R[2][k]
    for (int k = 0; k < 3; ++k) {
        o.M[0][k] = view[4 * k + 0] * j00 + view[4 * k + 2] * j02;
        o.M[1][k] = view[4 * k + 1] * j11 + view[4 * k + 2] * j12;
    }
    const R S[3][3] = {{cov3D[0], cov3D[1], cov3D[2]}, {cov3D[1], cov3D[3], cov3D[4]}, {cov3D[2], cov3D[4], cov3D[5]}};
    R A[2][3];
    for (int i = 0; i < 2; ++i)
        for (int j = 0; j < 3; ++j) A[i][j] = o.M[i][0] * S[0][j] + o.M[i][1] * S[1][j] + o.M[i][2] * S[2][j];
    const R c00 = A[0][0] * o.M[0][0] + A[0][1] * o.M[0][1] + A[0][2] * o.M[0][2];
    const R c01 = A[1][0] * o.M[0][0] + A[1][1] * o.M[0][1] + A[1][2] * o.M[0][2];
    const R c11 = A[1][0] * o.M[1][0] + A[1][1] * o.M[1][1] + A[1][2] * o.M[1][2];
    o.a = c00 + R(0.3f);
    o.b = c01;
    o.c = c11 + R(0.3f);
    o.tx = tx; o.ty = ty; o.tz = tz; o.fx = fx; o.fy = fy;
}

template <typename R>
struct Geom {                       // per-Gaussian forward state (upstream "geomBuffer")
    std::vector<R> depth, x, y, cA, cB, cC, opac;
    std::vector<int> radii, rminx, rminy, rmaxx, rmaxy;
    std::vector<uint32_t> tiles_touched;
    void resize(int n) {
        depth.assign(n, 0); x.assign(n, 0); y.assign(n, 0); cA.assign(n, 0); cB.assign(n, 0); cC.assign(n, 0);
        opac.assign(n, 0); radii.assign(n, 0); rminx.assign(n, 0); rminy.assign(n, 0); rmaxx.assign(n, 0);
        rmaxy.assign(n, 0); tiles_touched.assign(n, 0);
    }
};

// per-Gaussian projection, conic, radius, tile rectangle.
template <typename R>
void preprocess(int N, int H, int W, const R* means, const R* cov3D, const R* opac, const R* view,
                const R* proj, R tanfovx, R tanfovy, Geom<R>& g) {
    const int gx = (W + kBlockX - 1) / kBlockX, gy = (H + kBlockY - 1) / kBlockY;
    g.resize(N);
#pragma omp parallel for schedule(static)
    for (int i = 0; i < N; ++i) {
        const R* m = means + 3 * i;
        const R pvz = view[2] * m[0] + view[6] * m[1] + view[10] * m[2] + view[14];
        if (!(pvz > R(0.2f))) continue;            // "p_view.z <= 0.2f -> cull" (NaN depth is culled too)
        const R hx = proj[0] * m[0] + proj[4] * m[1] + proj[8] * m[2] + proj[12];
        const R hy = proj[1] * m[0] + proj[5] * m[1] + proj[9] * m[2] + proj[13];
        const R hw = proj[3] * m[0] + proj[7] * m[1] + proj[11] * m[2] + proj[15];
        const R pw = R(1.0f) / (hw + R(0.0000001f));
        const R projx = hx * pw, projy = hy * pw;
        Cov2DTerms<R> c2;
        compute_cov2d(m, cov3D + 6 * i, view, tanfovx, tanfovy, W, H, c2);
        const R det = c2.a * c2.c - c2.b * c2.b;
        if (det == R(0)) continue;
        const R det_inv = R(1.0f) / det;
        const R mid = R(0.5f) * (c2.a + c2.c);
        const R disc = std::sqrt(std::max(R(0.1f), mid * mid - det));
        const R l1 = mid + disc, l2 = mid - disc;
        const R radius = std::ceil(R(3.0f) * std::sqrt(std::max(l1, l2)));
        // ndc2Pix in double (upstream literals are doubles)
        const R px = R(((double(projx) + 1.0) * double(W) - 1.0) * 0.5);
        const R py = R(((double(projy) + 1.0) * double(H) - 1.0) * 0.5);
        const int rminx = std::min(gx, std::max(0, f2i_rz_sat((px - radius) / R(kBlockX))));
        const int rminy = std::min(gy, std::max(0, f2i_rz_sat((py - radius) / R(kBlockY))));
        const int rmaxx = std::min(gx, std::max(0, f2i_rz_sat((px + radius + R(kBlockX - 1)) / R(kBlockX))));
        const int rmaxy = std::min(gy, std::max(0, f2i_rz_sat((py + radius + R(kBlockY - 1)) / R(kBlockY))));
        if ((rmaxx - rminx) * (rmaxy - rminy) == 0) continue;
        g.depth[i] = pvz;
        g.radii[i] = f2i_rz_sat(radius);
        g.x[i] = px; g.y[i] = py;
        g.cA[i] = c2.c * det_inv; g.cB[i] = -c2.b * det_inv; g.cC[i] = c2.a * det_inv;
        g.opac[i] = opac[i];
        g.rminx[i] = rminx; g.rminy[i] = rminy; g.rmaxx[i] = rmaxx; g.rmaxy[i] = rmaxy;
        g.tiles_touched[i] = uint32_t((rmaxx - rminx) * (rmaxy - rminy));
    }
}

struct Binning {                    // upstream "binningBuffer": sorted instance list + per-tile ranges
    std::vector<uint32_t> point_list;
    std::vector<uint32_t> range_start, range_end;
};

// duplicate with keys (tile << 32 | depth bits), stable sort, identify ranges.
template <typename R>
void bin_and_sort(int N, int H, int W, const Geom<R>& g, Binning& b) {
    const int gx = (W + kBlockX - 1) / kBlockX, gy = (H + kBlockY - 1) / kBlockY;
    size_t total = 0;
    std::vector<size_t> off(N + 1, 0);
    for (int i = 0; i < N; ++i) { off[i] = total; total += g.tiles_touched[i]; }
    off[N] = total;
    std::vector<std::pair<uint64_t, uint32_t>> kv(total);
#pragma omp parallel for schedule(static)
    for (int i = 0; i < N; ++i) {
        if (g.radii[i] <= 0) continue;
        size_t o = off[i];
        const uint64_t dbits = float_bits(float(g.depth[i]));
        for (int y = g.rminy[i]; y < g.rmaxy[i]; ++y)
            for (int x = g.rminx[i]; x < g.rmaxx[i]; ++x)
                kv[o++] = {(uint64_t(uint32_t(y * gx + x)) << 32) | dbits, uint32_t(i)};
    }
    std::stable_sort(kv.begin(), kv.end(), [](const auto& l, const auto& r) { return l.first < r.first; });
    b.point_list.resize(total);
    b.range_start.assign(size_t(gx) * gy, 0);
    b.range_end.assign(size_t(gx) * gy, 0);
    for (size_t k = 0; k < total; ++k) {
        b.point_list[k] = kv[k].second;
        const uint32_t t = uint32_t(kv[k].first >> 32);
        if (k == 0 || uint32_t(kv[k - 1].first >> 32) != t) b.range_start[t] = uint32_t(k);
        if (k + 1 == total || uint32_t(kv[k + 1].first >> 32) != t) b.range_end[t] = uint32_t(k + 1);
    }
}

// front-to-back alpha compositing of one pixel.
template <typename R>
void blend_forward(int H, int W, const Geom<R>& g, const Binning& b, const R* colors, const R* bg,
                   R* out_color, R* out_depth, R* out_alpha, uint32_t* n_contrib, uint64_t* eval_stats) {
    const int gx = (W + kBlockX - 1) / kBlockX, gy = (H + kBlockY - 1) / kBlockY;
    uint64_t evals = 0, blends = 0;
#pragma omp parallel for schedule(dynamic, 1) reduction(+ : evals, blends)
    for (int tile = 0; tile < gx * gy; ++tile) {
        const int tx = tile % gx, ty = tile / gx;
        const uint32_t s = b.range_start[tile], e = b.range_end[tile];
        for (int ly = 0; ly < kBlockY; ++ly)
            for (int lx = 0; lx < kBlockX; ++lx) {
                const int px = tx * kBlockX + lx, py = ty * kBlockY + ly;
                if (px >= W || py >= H) continue;
                const R pxf = R(px), pyf = R(py);
                R T = R(1.0f), C[3] = {0, 0, 0}, D = 0, Wt = 0;
                uint32_t contributor = 0, last = 0;
                for (uint32_t k = s; k < e; ++k) {
                    ++contributor;
                    ++evals;
                    const uint32_t id = b.point_list[k];
                    const R dx = g.x[id] - pxf, dy = g.y[id] - pyf;
                    const R power = gauss_power(g.cA[id], g.cB[id], g.cC[id], dx, dy);
                    if (power > R(0)) continue;
                    const R alpha = std::min(R(0.99f), g.opac[id] * exp_spec(power));
                    if (alpha < R(1.0f / 255.0f)) continue;
                    const R test_T = fma_r(-alpha, T, T);      // T * (1 - alpha)
                    if (test_T < R(0.0001f)) break;          // "done = true"
                    ++blends;
                    const R w = alpha * T;
                    for (int ch = 0; ch < 3; ++ch) C[ch] = fma_r(colors[3 * id + ch], w, C[ch]);
                    Wt += w;
                    D = fma_r(g.depth[id], w, D);
                    T = test_T;
                    last = contributor;
                }
                const size_t pix = size_t(py) * W + px;
                for (int ch = 0; ch < 3; ++ch) out_color[size_t(ch) * H * W + pix] = C[ch] + T * bg[ch];
                out_depth[pix] = D;
                out_alpha[pix] = Wt;
                n_contrib[pix] = last;
            }
    }
    if (eval_stats) { eval_stats[0] = evals; eval_stats[1] = blends; }
}

// back-to-front replay.  Per-Gaussian sums are accumulated in double so the
// oracle is the correctly-rounded sum of the fp32 per-pixel terms (upstream sums them with
// fp32 atomics in a nondeterministic order).
template <typename R>
void blend_backward(int N, int H, int W, const Geom<R>& g, const Binning& b, const R* colors, const R* bg,
                    const R* out_alpha, const uint32_t* n_contrib, const R* dL_dcolor, const R* dL_ddepth,
                    const R* dL_dalpha, std::vector<double>& acc /* [N][10] */) {
    const int gx = (W + kBlockX - 1) / kBlockX, gy = (H + kBlockY - 1) / kBlockY;
    acc.assign(size_t(N) * 10, 0.0);
    int nthreads = 1;
#ifdef _OPENMP
    nthreads = omp_get_max_threads();
#endif
    std::vector<std::vector<double>> priv(nthreads);
#pragma omp parallel
    {
        int tid = 0;
#ifdef _OPENMP
        tid = omp_get_thread_num();
#endif
        std::vector<double>& a = priv[tid];
        a.assign(size_t(N) * 10, 0.0);
#pragma omp for schedule(dynamic, 1)
        for (int tile = 0; tile < gx * gy; ++tile) {
            const int tx = tile % gx, ty = tile / gx;
            const uint32_t s = b.range_start[tile], e = b.range_end[tile];
            for (int ly = 0; ly < kBlockY; ++ly)
                for (int lx = 0; lx < kBlockX; ++lx) {
                    const int px = tx * kBlockX + lx, py = ty * kBlockY + ly;
                    if (px >= W || py >= H) continue;
                    const size_t pix = size_t(py) * W + px;
                    const R pxf = R(px), pyf = R(py);
                    const R T_final = R(1) - out_alpha[pix];
                    R T = T_final;
                    const uint32_t last = n_contrib[pix];
                    R accum_rec[3] = {0, 0, 0}, accum_depth_rec = 0, accum_alpha_rec = 0;
                    R last_alpha = 0, last_color[3] = {0, 0, 0}, last_depth = 0;
                    R dpix[3];
                    for (int ch = 0; ch < 3; ++ch) dpix[ch] = dL_dcolor[size_t(ch) * H * W + pix];
                    const R ddep = dL_ddepth[pix], dalp = dL_dalpha[pix];
                    const R ddelx_dx = R(0.5f) * R(W), ddely_dy = R(0.5f) * R(H);
                    R bg_dot = 0;
                    for (int ch = 0; ch < 3; ++ch) bg_dot += bg[ch] * dpix[ch];
                    uint32_t contributor = e - s;
                    for (uint32_t k = e; k-- > s;) {
                        --contributor;
                        if (contributor >= last) continue;
                        const uint32_t id = b.point_list[k];
                        const R dx = g.x[id] - pxf, dy = g.y[id] - pyf;
                        const R cA = g.cA[id], cB = g.cB[id], cC = g.cC[id], op = g.opac[id];
                        const R power = gauss_power(cA, cB, cC, dx, dy);
                        if (power > R(0)) continue;
                        const R G = exp_spec(power);
                        const R alpha = std::min(R(0.99f), op * G);
                        if (alpha < R(1.0f / 255.0f)) continue;
                        T = T / (R(1) - alpha);
                        const R w = alpha * T;
                        R dL_dal = 0;
                        double* ga = &a[size_t(id) * 10];
                        for (int ch = 0; ch < 3; ++ch) {
                            const R c = colors[3 * id + ch];
                            accum_rec[ch] = last_alpha * last_color[ch] + (R(1) - last_alpha) * accum_rec[ch];
                            last_color[ch] = c;
                            dL_dal += (c - accum_rec[ch]) * dpix[ch];
                            ga[6 + ch] += double(w * dpix[ch]);                // dL/drgb
                        }
                        const R z = g.depth[id];
                        accum_depth_rec = last_alpha * last_depth + (R(1) - last_alpha) * accum_depth_rec;
                        last_depth = z;
                        dL_dal += (z - accum_depth_rec) * ddep;
                        ga[9] += double(w * ddep);                              // dL/dz
                        accum_alpha_rec = last_alpha + (R(1) - last_alpha) * accum_alpha_rec;
                        dL_dal += (R(1) - accum_alpha_rec) * dalp;
                        dL_dal *= T;
                        last_alpha = alpha;
                        dL_dal += (-T_final / (R(1) - alpha)) * bg_dot;
                        const R dL_dG = op * dL_dal;
                        const R gdx = G * dx, gdy = G * dy;
                        const R dG_ddelx = -gdx * cA - gdy * cB;
                        const R dG_ddely = -gdy * cC - gdx * cB;
                        ga[0] += double(dL_dG * dG_ddelx * ddelx_dx);          // dL/dmean2D.x
                        ga[1] += double(dL_dG * dG_ddely * ddely_dy);          // dL/dmean2D.y
                        ga[2] += double(R(-0.5f) * gdx * dx * dL_dG);          // dL/dconic A
                        ga[3] += double(R(-0.5f) * gdx * dy * dL_dG);          // dL/dconic B (half convention)
                        ga[4] += double(R(-0.5f) * gdy * dy * dL_dG);          // dL/dconic C
                        ga[5] += double(G * dL_dal);                           // dL/dopacity
                    }
                }
        }
    }
    // fixed-order (thread 0, 1, ...) sum of the per-thread accumulators, parallel over the entries
    const long long nacc = static_cast<long long>(acc.size());
#pragma omp parallel for schedule(static)
    for (long long k = 0; k < nacc; ++k) {
        double sum = 0.0;
        for (int t = 0; t < nthreads; ++t)
            if (!priv[t].empty()) sum += priv[t][size_t(k)];
        acc[size_t(k)] = sum;
    }
}

// per-Gaussian backward (conic -> cov2D -> cov3D, mean3D through J, the projected
// mean and the depth).
template <typename R>
void preprocess_backward(int N, int H, int W, const R* means, const R* cov3D, const R* view, const R* proj,
                         R tanfovx, R tanfovy, const int* radii, const std::vector<double>& acc,
                         R* dL_dmeans3D, R* dL_dmeans2D, R* dL_dcov3D, R* dL_dcolors, R* dL_dopac) {
#pragma omp parallel for schedule(static)
    for (int i = 0; i < N; ++i) {
        for (int k = 0; k < 3; ++k) { dL_dmeans3D[3 * i + k] = 0; dL_dmeans2D[3 * i + k] = 0; dL_dcolors[3 * i + k] = 0; }
        for (int k = 0; k < 6; ++k) dL_dcov3D[6 * i + k] = 0;
        dL_dopac[i] = 0;
        if (!(radii[i] > 0)) continue;
        const double* ga = &acc[size_t(i) * 10];
        const R g2x = R(ga[0]), g2y = R(ga[1]);
        const R dA = R(ga[2]), dB = R(ga[3]), dC = R(ga[4]);
        dL_dopac[i] = R(ga[5]);
        for (int ch = 0; ch < 3; ++ch) dL_dcolors[3 * i + ch] = R(ga[6 + ch]);
        const R dz = R(ga[9]);
        dL_dmeans2D[3 * i + 0] = g2x;
        dL_dmeans2D[3 * i + 1] = g2y;

        const R* m = means + 3 * i;
        const R* S6 = cov3D + 6 * i;
        Cov2DTerms<R> c2;
        compute_cov2d(m, S6, view, tanfovx, tanfovy, W, H, c2);
        const R a = c2.a, b = c2.b, c = c2.c;
        const R denom = a * c - b * b;
        const R denom2inv = R(1.0f) / ((denom * denom) + R(0.0000001f));
        R dL_da = 0, dL_db = 0, dL_dc = 0;
        const R(*M)[3] = c2.M;
        if (denom2inv != R(0)) {
            dL_da = denom2inv * (-c * c * dA + R(2) * b * c * dB + (denom - a * c) * dC);
            dL_dc = denom2inv * (-a * a * dC + R(2) * a * b * dB + (denom - a * c) * dA);
            dL_db = denom2inv * R(2) * (b * c * dA - (denom + R(2) * b * b) * dB + a * b * dC);
            R* o = dL_dcov3D + 6 * i;
            o[0] = M[0][0] * M[0][0] * dL_da + M[0][0] * M[1][0] * dL_db + M[1][0] * M[1][0] * dL_dc;
            o[3] = M[0][1] * M[0][1] * dL_da + M[0][1] * M[1][1] * dL_db + M[1][1] * M[1][1] * dL_dc;
            o[5] = M[0][2] * M[0][2] * dL_da + M[0][2] * M[1][2] * dL_db + M[1][2] * M[1][2] * dL_dc;
            o[1] = R(2) * M[0][0] * M[0][1] * dL_da + (M[0][0] * M[1][1] + M[0][1] * M[1][0]) * dL_db + R(2) * M[1][0] * M[1][1] * dL_dc;
            o[2] = R(2) * M[0][0] * M[0][2] * dL_da + (M[0][0] * M[1][2] + M[0][2] * M[1][0]) * dL_db + R(2) * M[1][0] * M[1][2] * dL_dc;
            o[4] = R(2) * M[0][2] * M[0][1] * dL_da + (M[0][1] * M[1][2] + M[0][2] * M[1][1]) * dL_db + R(2) * M[1][1] * M[1][2] * dL_dc;
        }
        const R S[3][3] = {{S6[0], S6[1], S6[2]}, {S6[1], S6[3], S6[4]}, {S6[2], S6[4], S6[5]}};
        // dL/dM (2x3): row0 = 2*(M0.S)*dL_da + (M1.S)*dL_db ; row1 = 2*(M1.S)*dL_dc + (M0.S)*dL_db
        R dM[2][3];
        for (int k = 0; k < 3; ++k) {
            const R m0s = M[0][0] * S[k][0] + M[0][1] * S[k][1] + M[0][2] * S[k][2];
            const R m1s = M[1][0] * S[k][0] + M[1][1] * S[k][1] + M[1][2] * S[k][2];
            dM[0][k] = R(2) * m0s * dL_da + m1s * dL_db;
            dM[1][k] = R(2) * m1s * dL_dc + m0s * dL_db;
        }
        // dL/dJ = dL/dM . R_w2c^T ; R_w2c[r][k] = view[4k + r]
        const R dJ00 = view[0] * dM[0][0] + view[4] * dM[0][1] + view[8] * dM[0][2];
        const R dJ02 = view[2] * dM[0][0] + view[6] * dM[0][1] + view[10] * dM[0][2];
        const R dJ11 = view[1] * dM[1][0] + view[5] * dM[1][1] + view[9] * dM[1][2];
        const R dJ12 = view[2] * dM[1][0] + view[6] * dM[1][1] + view[10] * dM[1][2];
        const R tz = R(1.0f) / c2.tz, tz2 = tz * tz, tz3 = tz2 * tz;
        const R dtx = c2.xmul * -c2.fx * tz2 * dJ02;
        const R dty = c2.ymul * -c2.fy * tz2 * dJ12;
        const R dtz = -c2.fx * tz2 * dJ00 - c2.fy * tz2 * dJ11 + (R(2) * c2.fx * c2.tx) * tz3 * dJ02 +
                      (R(2) * c2.fy * c2.ty) * tz3 * dJ12;
        // transformVec4x3Transpose
        R gm[3] = {view[0] * dtx + view[1] * dty + view[2] * dtz, view[4] * dtx + view[5] * dty + view[6] * dtz,
                   view[8] * dtx + view[9] * dty + view[10] * dtz};
        // projected mean -> mean
        const R hx = proj[0] * m[0] + proj[4] * m[1] + proj[8] * m[2] + proj[12];
        const R hy = proj[1] * m[0] + proj[5] * m[1] + proj[9] * m[2] + proj[13];
        const R hw = proj[3] * m[0] + proj[7] * m[1] + proj[11] * m[2] + proj[15];
        const R mw = R(1.0f) / (hw + R(0.0000001f));
        const R mul1 = hx * mw * mw, mul2 = hy * mw * mw;
        gm[0] += (proj[0] * mw - proj[3] * mul1) * g2x + (proj[1] * mw - proj[3] * mul2) * g2y;
        gm[1] += (proj[4] * mw - proj[7] * mul1) * g2x + (proj[5] * mw - proj[7] * mul2) * g2y;
        gm[2] += (proj[8] * mw - proj[11] * mul1) * g2x + (proj[9] * mw - proj[11] * mul2) * g2y;
        // depth -> mean (depth/alpha fork)
        const R mul3 = view[2] * m[0] + view[6] * m[1] + view[10] * m[2] + view[14];
        gm[0] += (view[2] - view[3] * mul3) * dz;
        gm[1] += (view[6] - view[7] * mul3) * dz;
        gm[2] += (view[10] - view[11] * mul3) * dz;
        for (int k = 0; k < 3; ++k) dL_dmeans3D[3 * i + k] = gm[k];
    }
}

template <typename R>
struct Ctx {                        // state kept between forward and backward (one render)
    int N = 0, H = 0, W = 0;
    Geom<R> g;
    Binning b;
    std::vector<uint32_t> n_contrib;
};

template <typename R>
int forward_impl(Ctx<R>& ctx, int N, int H, int W, const R* means, const R* cov3D, const R* colors, const R* opac,
                 const R* view, const R* proj, R tanfovx, R tanfovy, const R* bg, R* out_color, R* out_depth,
                 R* out_alpha, int* radii, uint64_t* stats) {
    ctx.N = N; ctx.H = H; ctx.W = W;
    preprocess(N, H, W, means, cov3D, opac, view, proj, tanfovx, tanfovy, ctx.g);
    bin_and_sort(N, H, W, ctx.g, ctx.b);
    ctx.n_contrib.assign(size_t(H) * W, 0);
    blend_forward(H, W, ctx.g, ctx.b, colors, bg, out_color, out_depth, out_alpha, ctx.n_contrib.data(),
                  stats ? stats + 1 : nullptr);
    for (int i = 0; i < N; ++i) radii[i] = ctx.g.radii[i];
    if (stats) stats[0] = ctx.b.point_list.size();
    return 0;
}

template <typename R>
int backward_impl(Ctx<R>& ctx, const R* means, const R* cov3D, const R* colors, const R* view, const R* proj,
                  R tanfovx, R tanfovy, const R* bg, const R* out_alpha, const R* dL_dcolor, const R* dL_ddepth,
                  const R* dL_dalpha, R* dL_dmeans3D, R* dL_dmeans2D, R* dL_dcov3D, R* dL_dcolors, R* dL_dopac) {
    std::vector<double> acc;
    blend_backward(ctx.N, ctx.H, ctx.W, ctx.g, ctx.b, colors, bg, out_alpha, ctx.n_contrib.data(), dL_dcolor,
                   dL_ddepth, dL_dalpha, acc);
    preprocess_backward(ctx.N, ctx.H, ctx.W, means, cov3D, view, proj, tanfovx, tanfovy, ctx.g.radii.data(), acc,
                        dL_dmeans3D, dL_dmeans2D, dL_dcov3D, dL_dcolors, dL_dopac);
    return 0;
}

// --- optional input paths of the rasteriser API (unused by SIGMAN, kept for API completeness) ----
// computeCov3D: Sigma = (S R)^T (S R) with S = diag(mod * scale), R from the (unnormalised) quaternion (r,x,y,z).
template <typename R>
void cov3d_from_scale_rot(int N, const R* scales, const R* rots, R mod, R* cov6) {
    for (int i = 0; i < N; ++i) {
        const R r = rots[4 * i], x = rots[4 * i + 1], y = rots[4 * i + 2], z = rots[4 * i + 3];
        // Rm[row][col] — the rotation matrix of the quaternion
        const R Rm[3][3] = {{R(1) - R(2) * (y * y + z * z), R(2) * (x * y - r * z), R(2) * (x * z + r * y)},
                            {R(2) * (x * y + r * z), R(1) - R(2) * (x * x + z * z), R(2) * (y * z - r * x)},
                            {R(2) * (x * z - r * y), R(2) * (y * z + r * x), R(1) - R(2) * (x * x + y * y)}};
        const R s[3] = {mod * scales[3 * i], mod * scales[3 * i + 1], mod * scales[3 * i + 2]};
        // Sigma = Rm diag(s^2) Rm^T ;  L[j][k] = Rm[j][k] * s[k]
        R L[3][3];
        for (int j = 0; j < 3; ++j)
            for (int k = 0; k < 3; ++k) L[j][k] = Rm[j][k] * s[k];
        auto dot = [&](int a, int b) { return L[a][0] * L[b][0] + L[a][1] * L[b][1] + L[a][2] * L[b][2]; };
        R* o = cov6 + 6 * i;
        o[0] = dot(0, 0); o[1] = dot(0, 1); o[2] = dot(0, 2); o[3] = dot(1, 1); o[4] = dot(1, 2); o[5] = dot(2, 2);
    }
}

// distCUDA2: mean squared distance to the 3 nearest OTHER points; brute force.
void knn_mean_dist2(int N, const float* pts, float* out) {
#pragma omp parallel for schedule(static)
    for (int i = 0; i < N; ++i) {
        float best[3] = {INFINITY, INFINITY, INFINITY};
        const float px = pts[3 * i], py = pts[3 * i + 1], pz = pts[3 * i + 2];
        for (int j = 0; j < N; ++j) {
            if (j == i) continue;
            const float dx = pts[3 * j] - px, dy = pts[3 * j + 1] - py, dz = pts[3 * j + 2] - pz;
            const float d = dx * dx + dy * dy + dz * dz;
            if (d < best[2]) {
                if (d < best[1]) {
                    best[2] = best[1];
                    if (d < best[0]) { best[1] = best[0]; best[0] = d; } else best[1] = d;
                } else best[2] = d;
            }
        }
        out[i] = (best[0] + best[1] + best[2]) / 3.0f;
    }
}


// ------------------------------------------------------------------ optional colour path: spherical harmonics
// upstream computeColorFromSH (SURVEY.md A.2 "SH path"; unused by SIGMAN, which passes colors_precomp with
// sh_degree = 0, /root/reference/core/gaussians/gs.py:91,102): real SH basis of Kerbl et al. 2023 up to degree 3,
// evaluated on dir = normalize(mean - campos); colour = sum_k basis_k(dir) * sh[k] + 0.5, clamped at 0 with the clamp
// recorded so that the backward zeroes the gradient of clamped channels.
constexpr double kSH0 = 0.28209479177387814, kSH1 = 0.4886025119029199;
constexpr double kSH2[5] = {1.0925484305920792, -1.0925484305920792, 0.31539156525252005, -1.0925484305920792,
                            0.5462742152960396};
constexpr double kSH3[7] = {-0.5900435899266435, 2.890611442640554, -0.4570457994644658, 0.3731763325901154,
                            -0.4570457994644658, 1.445305721320277, -0.5900435899266435};

template <typename R>
void sh_basis(int deg, R x, R y, R z, R* b) {
    b[0] = R(kSH0);
    if (deg < 1) return;
    b[1] = -R(kSH1) * y; b[2] = R(kSH1) * z; b[3] = -R(kSH1) * x;
    if (deg < 2) return;
    const R xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
    b[4] = R(kSH2[0]) * xy; b[5] = R(kSH2[1]) * yz; b[6] = R(kSH2[2]) * (R(2) * zz - xx - yy);
    b[7] = R(kSH2[3]) * xz; b[8] = R(kSH2[4]) * (xx - yy);
    if (deg < 3) return;
    b[9] = R(kSH3[0]) * y * (R(3) * xx - yy); b[10] = R(kSH3[1]) * xy * z;
    b[11] = R(kSH3[2]) * y * (R(4) * zz - xx - yy); b[12] = R(kSH3[3]) * z * (R(2) * zz - R(3) * xx - R(3) * yy);
    b[13] = R(kSH3[4]) * x * (R(4) * zz - xx - yy); b[14] = R(kSH3[5]) * z * (xx - yy);
    b[15] = R(kSH3[6]) * x * (xx - R(3) * yy);
}

template <typename R>
void sh_colors(int N, int deg, int max_coeffs, const R* means, const R* shs, const R* campos, R* colors,
               unsigned char* clamped) {
    const int K = (deg + 1) * (deg + 1);
    for (int i = 0; i < N; ++i) {
        const R dx = means[3 * i] - campos[0], dy = means[3 * i + 1] - campos[1], dz = means[3 * i + 2] - campos[2];
        const R inv = R(1) / std::sqrt(dx * dx + dy * dy + dz * dz);
        R b[16];
        sh_basis<R>(deg, dx * inv, dy * inv, dz * inv, b);
        for (int c = 0; c < 3; ++c) {
            R v = 0;
            for (int k = 0; k < K; ++k) v += b[k] * shs[(size_t(i) * max_coeffs + k) * 3 + c];
            v += R(0.5);
            clamped[3 * i + c] = v < R(0) ? 1 : 0;
            colors[3 * i + c] = v < R(0) ? R(0) : v;
        }
    }
}

// ------------------------------------------------------------------ per-subject preparation of gs.py:69-73
// scale = (s + 1) * sqrt(max(d2, 1e-7)) (kNN factor detached), Sigma = Rm diag(scale^2) Rm^T, packed (xx,xy,xz,yy,yz,zz)
// (get_covariance + strip_lowerdiag, gs.py:17-38).  Rm is the "rotation-like" 3x3 matrix SIGMAN feeds (row-major).
template <typename R>
void prep_cov3d(int64_t n, const R* s_raw, const R* rot, const R* dist2, R* cov6) {
    for (int64_t i = 0; i < n; ++i) {
        const R d2 = dist2[i] < R(1e-7) ? R(1e-7) : dist2[i];
        const R sg = std::sqrt(d2);
        R e[3];
        for (int k = 0; k < 3; ++k) { const R sc = (s_raw[3 * i + k] + R(1)) * sg; e[k] = sc * sc; }
        const R* Rm = rot + 9 * i;
        const int ia[6] = {0, 0, 0, 1, 1, 2}, ib[6] = {0, 1, 2, 1, 2, 2};
        for (int q = 0; q < 6; ++q) {
            R v = 0;
            for (int k = 0; k < 3; ++k) v += Rm[3 * ia[q] + k] * e[k] * Rm[3 * ib[q] + k];
            cov6[6 * i + q] = v;
        }
    }
}

}  // namespace

// ------------------------------------------------------------------------------- C ABI
extern "C" {

struct OracleCtxF32 { Ctx<float> c; };
struct OracleCtxF64 { Ctx<double> c; };

float oracle_expf(float x) { return exp_spec(x); }
// 0 = exp_spec (the bit-exact specification shared with the kernels' SGR_FLAG_EXACT_EXP mode), 1 = libm expf
void oracle_set_exp_mode(int mode) { g_exp_mode = mode; }
int oracle_get_exp_mode(void) { return g_exp_mode; }
int oracle_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
void oracle_set_num_threads(int n) {
#ifdef _OPENMP
    omp_set_num_threads(n);
#else
    (void)n;
#endif
}

void* oracle_create_f32(void) { return new OracleCtxF32(); }
void oracle_destroy_f32(void* h) { delete static_cast<OracleCtxF32*>(h); }
void* oracle_create_f64(void) { return new OracleCtxF64(); }
void oracle_destroy_f64(void* h) { delete static_cast<OracleCtxF64*>(h); }

// stats[0] = number of (Gaussian, tile) instances, stats[1] = pixel*Gaussian evaluations, stats[2] = accepted blends
int oracle_forward_f32(void* h, int N, int H, int W, const float* means, const float* cov3D, const float* colors,
                       const float* opac, const float* view, const float* proj, float tanfovx, float tanfovy,
                       const float* bg, float* out_color, float* out_depth, float* out_alpha, int* radii,
                       uint64_t* stats) {
    return forward_impl(static_cast<OracleCtxF32*>(h)->c, N, H, W, means, cov3D, colors, opac, view, proj, tanfovx,
                        tanfovy, bg, out_color, out_depth, out_alpha, radii, stats);
}
int oracle_forward_f64(void* h, int N, int H, int W, const double* means, const double* cov3D, const double* colors,
                       const double* opac, const double* view, const double* proj, double tanfovx, double tanfovy,
                       const double* bg, double* out_color, double* out_depth, double* out_alpha, int* radii,
                       uint64_t* stats) {
    return forward_impl(static_cast<OracleCtxF64*>(h)->c, N, H, W, means, cov3D, colors, opac, view, proj, tanfovx,
                        tanfovy, bg, out_color, out_depth, out_alpha, radii, stats);
}
int oracle_backward_f32(void* h, const float* means, const float* cov3D, const float* colors, const float* view,
                        const float* proj, float tanfovx, float tanfovy, const float* bg, const float* out_alpha,
                        const float* dL_dcolor, const float* dL_ddepth, const float* dL_dalpha, float* dL_dmeans3D,
                        float* dL_dmeans2D, float* dL_dcov3D, float* dL_dcolors, float* dL_dopac) {
    return backward_impl(static_cast<OracleCtxF32*>(h)->c, means, cov3D, colors, view, proj, tanfovx, tanfovy, bg,
                         out_alpha, dL_dcolor, dL_ddepth, dL_dalpha, dL_dmeans3D, dL_dmeans2D, dL_dcov3D, dL_dcolors,
                         dL_dopac);
}
int oracle_backward_f64(void* h, const double* means, const double* cov3D, const double* colors, const double* view,
                        const double* proj, double tanfovx, double tanfovy, const double* bg, const double* out_alpha,
                        const double* dL_dcolor, const double* dL_ddepth, const double* dL_dalpha,
                        double* dL_dmeans3D, double* dL_dmeans2D, double* dL_dcov3D, double* dL_dcolors,
                        double* dL_dopac) {
    return backward_impl(static_cast<OracleCtxF64*>(h)->c, means, cov3D, colors, view, proj, tanfovx, tanfovy, bg,
                         out_alpha, dL_dcolor, dL_ddepth, dL_dalpha, dL_dmeans3D, dL_dmeans2D, dL_dcov3D, dL_dcolors,
                         dL_dopac);
}

// Intermediate state of the last forward (fp32 context) for stage-by-stage GPU parity checks.
// geom: float[N][7] = depth, x, y, conicA, conicB, conicC, opacity ; rect: int[N][4] = minx,miny,maxx,maxy.
void oracle_get_geom_f32(void* h, float* geom, int* rect, uint32_t* tiles_touched) {
    const Ctx<float>& c = static_cast<OracleCtxF32*>(h)->c;
    for (int i = 0; i < c.N; ++i) {
        float* o = geom + 7 * i;
        o[0] = c.g.depth[i]; o[1] = c.g.x[i]; o[2] = c.g.y[i]; o[3] = c.g.cA[i]; o[4] = c.g.cB[i]; o[5] = c.g.cC[i];
        o[6] = c.g.opac[i];
        int* r = rect + 4 * i;
        r[0] = c.g.rminx[i]; r[1] = c.g.rminy[i]; r[2] = c.g.rmaxx[i]; r[3] = c.g.rmaxy[i];
        tiles_touched[i] = c.g.tiles_touched[i];
    }
}
uint64_t oracle_num_instances_f32(void* h) { return static_cast<OracleCtxF32*>(h)->c.b.point_list.size(); }
// point_list: uint32[I]; ranges: uint32[tiles][2]; n_contrib: uint32[H*W]
void oracle_get_binning_f32(void* h, uint32_t* point_list, uint32_t* ranges, uint32_t* n_contrib) {
    const Ctx<float>& c = static_cast<OracleCtxF32*>(h)->c;
    std::copy(c.b.point_list.begin(), c.b.point_list.end(), point_list);
    for (size_t t = 0; t < c.b.range_start.size(); ++t) { ranges[2 * t] = c.b.range_start[t]; ranges[2 * t + 1] = c.b.range_end[t]; }
    std::copy(c.n_contrib.begin(), c.n_contrib.end(), n_contrib);
}

void oracle_cov3d_from_scale_rot_f32(int N, const float* scales, const float* rots, float mod, float* cov6) {
    cov3d_from_scale_rot(N, scales, rots, mod, cov6);
}
void oracle_cov3d_from_scale_rot_f64(int N, const double* scales, const double* rots, double mod, double* cov6) {
    cov3d_from_scale_rot(N, scales, rots, mod, cov6);
}
void oracle_knn_mean_dist2(int N, const float* pts, float* out) { knn_mean_dist2(N, pts, out); }

void oracle_sh_colors_f32(int N, int deg, int max_coeffs, const float* means, const float* shs, const float* campos,
                          float* colors, unsigned char* clamped) {
    sh_colors<float>(N, deg, max_coeffs, means, shs, campos, colors, clamped);
}
void oracle_sh_colors_f64(int N, int deg, int max_coeffs, const double* means, const double* shs, const double* campos,
                          double* colors, unsigned char* clamped) {
    sh_colors<double>(N, deg, max_coeffs, means, shs, campos, colors, clamped);
}
void oracle_prep_cov3d_f32(int64_t n, const float* s_raw, const float* rot, const float* dist2, float* cov6) {
    prep_cov3d<float>(n, s_raw, rot, dist2, cov6);
}
void oracle_prep_cov3d_f64(int64_t n, const double* s_raw, const double* rot, const double* dist2, double* cov6) {
    prep_cov3d<double>(n, s_raw, rot, dist2, cov6);
}

}  // extern "C"
