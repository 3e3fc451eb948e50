// sgr_blend_simple.cu — straightforward blend kernels (SGR_FLAG_SIMPLE_BLEND): one thread per pixel, one CTA per
// tile, the whole CTA walks the tile's depth-ordered list in lock step.  Shaped like the published algorithm's
// renderCUDA; kept as a debugging aid and as the "recompiled generic SIMT" comparison point for the warp-culling
// TMA kernels in sgr_blend.cu.  Same arithmetic (and therefore the same bits) as oracle/sgr_oracle.cpp::blend_forward.
//
// Compiled with --fmad=false (see sgr_preprocess.cu).
#include "sgr_common.cuh"

namespace sgr {
namespace {

struct SimpleFwdArgs {
    RenderGeom g;
    int render_base;
    const unsigned int* tile_off;
    const unsigned int* tile_cnt;
    const float4 *rec0, *rec1, *rec2;
    const float* bg;
    unsigned int* n_contrib;
    float *out_color, *out_depth, *out_alpha;
    int clamp_color;
};

__global__ void __launch_bounds__(kTilePixels) blend_forward_simple_kernel(SimpleFwdArgs a) {
    __shared__ float4 s0[kTilePixels], s1[kTilePixels], s2[kTilePixels];
    const int rl = blockIdx.x / a.g.num_tiles;
    const int tile = blockIdx.x - rl * a.g.num_tiles;
    const int r = a.render_base + rl;
    const int tx = tile % a.g.tiles_x, ty = tile / a.g.tiles_x;
    const int lx = threadIdx.x & (kTile - 1), ly = threadIdx.x >> 4;
    const int px = tx * kTile + lx, py = ty * kTile + ly;
    const bool inside = px < a.g.W && py < a.g.H;
    const size_t tg = size_t(r) * a.g.num_tiles + tile;
    const unsigned int n = a.tile_cnt[tg];
    const size_t off = a.tile_off[tg];
    const float pxf = float(px), pyf = float(py);

    float T = 1.0f, C0 = 0.0f, C1 = 0.0f, C2 = 0.0f, D = 0.0f, Wt = 0.0f;
    unsigned int contributor = 0, last = 0;
    bool done = !inside;
    for (unsigned int base = 0; base < n; base += kTilePixels) {
        if (__syncthreads_count(done) == kTilePixels) break;
        const unsigned int m = min(unsigned(kTilePixels), n - base);
        if (threadIdx.x < m) {
            s0[threadIdx.x] = __ldg(a.rec0 + off + base + threadIdx.x);
            s1[threadIdx.x] = __ldg(a.rec1 + off + base + threadIdx.x);
            s2[threadIdx.x] = __ldg(a.rec2 + off + base + threadIdx.x);
        }
        __syncthreads();
        for (unsigned int j = 0; !done && j < m; ++j) {
            ++contributor;
            const float4 q0 = s0[j], q1 = s1[j];
            const float dx = q0.x - pxf, dy = q0.y - pyf;
            const float power = gauss_power(q1.x, q1.y, q1.z, dx, dy);
            if (power > 0.0f) continue;
            const float alpha = fminf(kAlphaMax, q1.w * exp_spec(power));
            if (alpha < kAlphaMin) continue;
            const float test_T = __fmaf_rn(-alpha, T, T);
            if (test_T < kTMin) { done = true; continue; }
            const float4 q2 = s2[j];
            const float w = alpha * T;
            C0 = __fmaf_rn(q2.x, w, C0);
            C1 = __fmaf_rn(q2.y, w, C1);
            C2 = __fmaf_rn(q2.z, w, C2);
            Wt += w;
            D = __fmaf_rn(q2.w, w, D);
            T = test_T;
            last = contributor;
        }
    }
    if (inside) {
        const size_t P = size_t(a.g.H) * a.g.W;
        const size_t pix = size_t(py) * a.g.W + px;
        float c0 = C0 + T * a.bg[0], c1 = C1 + T * a.bg[1], c2 = C2 + T * a.bg[2];
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

struct SimpleBwdArgs {
    RenderGeom g;
    int render_base;
    const unsigned int* tile_off;
    const unsigned int* tile_cnt;
    const unsigned int* sorted_ids;
    const float4 *rec0, *rec1, *rec2;
    const float* bg;
    const unsigned int* n_contrib;
    const float *out_alpha, *dL_dcolor, *dL_ddepth, *dL_dalpha;
    float* accum;             // [plane][kAccumStride]
    size_t plane;
};

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
    return v;
}

// oracle: blend_backward.  Per-pixel terms are the oracle's fp32 expressions; they are summed over the 32 pixels of a
// warp with shuffles and added to the per-(render, Gaussian) accumulators with one atomic per warp and component.
__global__ void __launch_bounds__(kTilePixels) blend_backward_simple_kernel(SimpleBwdArgs a) {
    __shared__ float4 s0[kTilePixels], s1[kTilePixels], s2[kTilePixels];
    __shared__ unsigned int sid[kTilePixels];
    __shared__ unsigned int s_max_last;
    const int rl = blockIdx.x / a.g.num_tiles;
    const int tile = blockIdx.x - rl * a.g.num_tiles;
    const int r = a.render_base + rl;
    const int tx = tile % a.g.tiles_x, ty = tile / a.g.tiles_x;
    const int lx = threadIdx.x & (kTile - 1), ly = threadIdx.x >> 4;
    const int px = tx * kTile + lx, py = ty * kTile + ly;
    const bool inside = px < a.g.W && py < a.g.H;
    const size_t tg = size_t(r) * a.g.num_tiles + tile;
    const unsigned int n = a.tile_cnt[tg];
    if (n == 0) return;
    const size_t off = a.tile_off[tg];
    const float pxf = float(px), pyf = float(py);
    const size_t P = size_t(a.g.H) * a.g.W;
    const size_t pix = size_t(py) * a.g.W + px;

    unsigned int last = 0;
    float T_final = 1.0f, dp0 = 0, dp1 = 0, dp2 = 0, ddep = 0, dalp = 0;
    if (inside) {
        last = a.n_contrib[size_t(r) * P + pix];
        T_final = 1.0f - a.out_alpha[size_t(r) * P + pix];
        const float* dc = a.dL_dcolor + size_t(r) * 3 * P;
        dp0 = dc[pix]; dp1 = dc[P + pix]; dp2 = dc[2 * P + pix];
        if (a.dL_ddepth) ddep = a.dL_ddepth[size_t(r) * P + pix];
        if (a.dL_dalpha) dalp = a.dL_dalpha[size_t(r) * P + pix];
    }
    if (threadIdx.x == 0) s_max_last = 0;
    __syncthreads();
    atomicMax(&s_max_last, last);
    __syncthreads();
    const unsigned int n_eff = min(n, s_max_last);     // entries at or beyond every pixel's n_contrib are never replayed
    float T = T_final;
    float ar0 = 0, ar1 = 0, ar2 = 0, adr = 0, aar = 0, last_alpha = 0, lc0 = 0, lc1 = 0, lc2 = 0, last_depth = 0;
    const float bg_dot = (a.bg[0] * dp0 + a.bg[1] * dp1) + a.bg[2] * dp2;
    const float ddelx_dx = 0.5f * float(a.g.W), ddely_dy = 0.5f * float(a.g.H);
    float* acc = a.accum + size_t(rl) * a.g.N * kAccumStride;

    const unsigned int rounds = (n_eff + kTilePixels - 1) / kTilePixels;
    for (unsigned int rd = rounds; rd-- > 0;) {
        const unsigned int base = rd * kTilePixels;
        const unsigned int m = min(unsigned(kTilePixels), n_eff - base);
        __syncthreads();
        if (threadIdx.x < m) {
            s0[threadIdx.x] = __ldg(a.rec0 + off + base + threadIdx.x);
            s1[threadIdx.x] = __ldg(a.rec1 + off + base + threadIdx.x);
            s2[threadIdx.x] = __ldg(a.rec2 + off + base + threadIdx.x);
            sid[threadIdx.x] = a.sorted_ids[off + base + threadIdx.x];
        }
        __syncthreads();
        for (unsigned int j = m; j-- > 0;) {
            const unsigned int contributor = base + j;
            float v[kAccumPlanes];
#pragma unroll
            for (int k = 0; k < kAccumPlanes; ++k) v[k] = 0.0f;
            bool active = false;
            if (inside && contributor < last) {
                const float4 q0 = s0[j], q1 = s1[j];
                const float dx = q0.x - pxf, dy = q0.y - pyf;
                const float power = gauss_power(q1.x, q1.y, q1.z, dx, dy);
                if (!(power > 0.0f)) {
                    const float G = exp_spec(power);
                    const float alpha = fminf(kAlphaMax, q1.w * G);
                    if (!(alpha < kAlphaMin)) {
                        active = true;
                        const float4 q2 = s2[j];
                        T = T / (1.0f - alpha);
                        const float w = alpha * T;
                        float dL_dal = 0.0f;
                        ar0 = last_alpha * lc0 + (1.0f - last_alpha) * ar0; lc0 = q2.x;
                        dL_dal += (q2.x - ar0) * dp0; v[6] = w * dp0;
                        ar1 = last_alpha * lc1 + (1.0f - last_alpha) * ar1; lc1 = q2.y;
                        dL_dal += (q2.y - ar1) * dp1; v[7] = w * dp1;
                        ar2 = last_alpha * lc2 + (1.0f - last_alpha) * ar2; lc2 = q2.z;
                        dL_dal += (q2.z - ar2) * dp2; v[8] = w * dp2;
                        adr = last_alpha * last_depth + (1.0f - last_alpha) * adr; last_depth = q2.w;
                        dL_dal += (q2.w - adr) * ddep; v[9] = w * ddep;
                        aar = last_alpha + (1.0f - last_alpha) * aar;
                        dL_dal += (1.0f - aar) * dalp;
                        dL_dal *= T;
                        last_alpha = alpha;
                        dL_dal += (-T_final / (1.0f - alpha)) * bg_dot;
                        const float dL_dG = q1.w * dL_dal;
                        const float gdx = G * dx, gdy = G * dy;
                        // rec1 = (-A/2, -B, -C/2, o):  -gdx*A - gdy*B = 2*gdx*hA + gdy*nB
                        const float dG_ddelx = 2.0f * gdx * q1.x + gdy * q1.y;
                        const float dG_ddely = 2.0f * gdy * q1.z + gdx * q1.y;
                        v[0] = dL_dG * dG_ddelx * ddelx_dx;
                        v[1] = dL_dG * dG_ddely * ddely_dy;
                        v[2] = -0.5f * gdx * dx * dL_dG;
                        v[3] = -0.5f * gdx * dy * dL_dG;
                        v[4] = -0.5f * gdy * dy * dL_dG;
                        v[5] = G * dL_dal;
                    }
                }
            }
            if (__any_sync(0xffffffffu, active)) {
                const unsigned int id = sid[j];
#pragma unroll
                for (int k = 0; k < kAccumPlanes; ++k) {
                    const float s = warp_sum(v[k]);
                    if ((threadIdx.x & 31) == 0 && s != 0.0f) atomicAdd(acc + size_t(id) * kAccumStride + k, s);
                }
            }
        }
    }
}

}  // namespace

cudaError_t launch_blend_forward_simple(const ChunkCtx& c, float* out_color, float* out_depth, float* out_alpha) {
    SimpleFwdArgs a;
    a.g = c.g; a.render_base = c.render_base; a.tile_off = c.tile_off; a.tile_cnt = c.tile_cnt;
    a.rec0 = c.rec0; a.rec1 = c.rec1; a.rec2 = c.rec2; a.bg = c.p->bg; a.n_contrib = c.n_contrib;
    a.out_color = out_color; a.out_depth = out_depth; a.out_alpha = out_alpha;
    a.clamp_color = (c.p->flags & SGR_FLAG_CLAMP_COLOR) ? 1 : 0;
    blend_forward_simple_kernel<<<c.num_renders * c.g.num_tiles, kTilePixels, 0, c.stream>>>(a);
    return cudaGetLastError();
}

cudaError_t launch_blend_backward_simple(const ChunkCtx& c, const float* out_alpha, const float* dL_dcolor,
                                         const float* dL_ddepth, const float* dL_dalpha) {
    SimpleBwdArgs a;
    a.g = c.g; a.render_base = c.render_base; a.tile_off = c.tile_off; a.tile_cnt = c.tile_cnt;
    a.sorted_ids = c.sorted_ids; a.rec0 = c.rec0; a.rec1 = c.rec1; a.rec2 = c.rec2; a.bg = c.p->bg;
    a.n_contrib = c.n_contrib; a.out_alpha = out_alpha; a.dL_dcolor = dL_dcolor; a.dL_ddepth = dL_ddepth;
    a.dL_dalpha = dL_dalpha; a.accum = c.accum; a.plane = size_t(c.num_renders) * c.g.N;
    blend_backward_simple_kernel<<<c.num_renders * c.g.num_tiles, kTilePixels, 0, c.stream>>>(a);
    return cudaGetLastError();
}

}  // namespace sgr
