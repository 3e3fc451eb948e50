// sgr_wire.cu — wire formats of the multi-GPU image gather (BASELINE config 4; mirrors
// `self.accelerator.gather(out['images_pred'])`, /root/reference/core/loss/eval.py:81-82).
//
// A rank's chunk of rendered views (RGB, depth, alpha planes) travels as ONE byte buffer
//     [ RGB block: n*3*P | depth block: n*P | alpha block: n*P ]
// either exact (float32 everywhere: the kernels of sgr_blend.cu write straight into it, nothing is packed) or compact
// (uint8 RGB = trunc(clamp(c, 0, 1) * 255 + 0.5), fp16 depth / alpha: 7 instead of 20 bytes per pixel).  sgr_wire_pack
// converts a chunk to the compact layout; sgr_wire_unpack scatters the all-gathered chunks of all ranks into the
// view-ordered float32 result [world][views_per_rank][5][P] in one pass (HBM-bound: 4 pixels per thread, 128-bit
// stores).
#include <cuda_fp16.h>

#include "sgr_common.cuh"

namespace sgr {
namespace {

__device__ __forceinline__ unsigned int to_u8(float x) {
    return static_cast<unsigned int>(fminf(fmaxf(x, 0.0f), 1.0f) * 255.0f + 0.5f);
}

// one thread = 4 consecutive pixels of one plane of one view
__global__ void __launch_bounds__(256) wire_pack_kernel(const float4* __restrict__ color, const float4* __restrict__ depth,
                                                        const float4* __restrict__ alpha, long long n, long long P4,
                                                        unsigned int* __restrict__ rgb_out, uint2* __restrict__ depth_out,
                                                        uint2* __restrict__ alpha_out) {
    const long long total = n * 5 * P4;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const long long view = t / (5 * P4);
        const long long rem = t - view * 5 * P4;
        const int ch = int(rem / P4);
        const long long p = rem - ch * P4;
        if (ch < 3) {
            const float4 v = __ldg(color + (view * 3 + ch) * P4 + p);
            rgb_out[(view * 3 + ch) * P4 + p] = to_u8(v.x) | (to_u8(v.y) << 8) | (to_u8(v.z) << 16) | (to_u8(v.w) << 24);
        } else {
            const float4 v = __ldg((ch == 3 ? depth : alpha) + view * P4 + p);
            const __half2 lo = __floats2half2_rn(v.x, v.y), hi = __floats2half2_rn(v.z, v.w);
            uint2 o;
            o.x = *reinterpret_cast<const unsigned int*>(&lo);
            o.y = *reinterpret_cast<const unsigned int*>(&hi);
            (ch == 3 ? depth_out : alpha_out)[view * P4 + p] = o;
        }
    }
}

template <bool kCompact>
__global__ void __launch_bounds__(256) wire_unpack_kernel(const unsigned char* __restrict__ recv, long long rank_stride,
                                                          int world, long long c, long long P4, float4* __restrict__ final_,
                                                          long long per, long long k0) {
    const long long per_rank = c * 5 * P4;
    const long long total = world * per_rank;
    const long long es_rgb = kCompact ? 1 : 4, es_da = kCompact ? 2 : 4;
    const long long off_depth = c * 3 * P4 * 4 * es_rgb, off_alpha = off_depth + c * P4 * 4 * es_da;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const long long r = t / per_rank;
        long long rem = t - r * per_rank;
        const long long view = rem / (5 * P4);
        rem -= view * 5 * P4;
        const int ch = int(rem / P4);
        const long long p = rem - ch * P4;
        const unsigned char* src = recv + r * rank_stride;
        float4 v;
        if (ch < 3) {
            if (kCompact) {
                const unsigned int u = __ldg(reinterpret_cast<const unsigned int*>(src) + (view * 3 + ch) * P4 + p);
                v = make_float4(float(u & 0xffu) / 255.0f, float((u >> 8) & 0xffu) / 255.0f, float((u >> 16) & 0xffu) / 255.0f,
                                float(u >> 24) / 255.0f);
            } else {
                v = __ldg(reinterpret_cast<const float4*>(src) + (view * 3 + ch) * P4 + p);
            }
        } else {
            const unsigned char* blk = src + (ch == 3 ? off_depth : off_alpha);
            if (kCompact) {
                const uint2 u = __ldg(reinterpret_cast<const uint2*>(blk) + view * P4 + p);
                const float2 lo = __half22float2(*reinterpret_cast<const __half2*>(&u.x));
                const float2 hi = __half22float2(*reinterpret_cast<const __half2*>(&u.y));
                v = make_float4(lo.x, lo.y, hi.x, hi.y);
            } else {
                v = __ldg(reinterpret_cast<const float4*>(blk) + view * P4 + p);
            }
        }
        final_[((r * per + k0 + view) * 5 + ch) * P4 + p] = v;
    }
}

int launch_grid(long long threads) {
    const long long blocks = (threads + 255) / 256;
    return int(blocks < 148 * 16 ? (blocks > 0 ? blocks : 1) : 148 * 16);
}

}  // namespace
}  // namespace sgr

extern "C" {
int sgr_set_error_(int code, const char* msg);
void sgr_count_launches_(unsigned int n);

uint64_t sgr_wire_chunk_bytes(int32_t num_views, int64_t pixels, int32_t compact) {
    if (num_views < 0 || pixels < 0) return 0;
    return uint64_t(num_views) * uint64_t(pixels) * (compact ? 7u : 20u);
}

int sgr_wire_pack(const float* color, const float* depth, const float* alpha, int32_t num_views, int64_t pixels,
                  void* out_bytes, void* stream) {
    if (num_views < 0 || pixels < 0 || (pixels & 3) || (num_views > 0 && pixels > 0 && (!color || !depth || !alpha || !out_bytes)))
        return sgr_set_error_(SGR_E_INVALID_ARGUMENT, "bad wire_pack arguments (pixels per image must be a multiple of 4)");
    if (num_views == 0 || pixels == 0) return SGR_OK;
    const long long n = num_views, P4 = pixels / 4;
    unsigned char* o = static_cast<unsigned char*>(out_bytes);
    sgr::wire_pack_kernel<<<sgr::launch_grid(n * 5 * P4), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        reinterpret_cast<const float4*>(color), reinterpret_cast<const float4*>(depth), reinterpret_cast<const float4*>(alpha),
        n, P4, reinterpret_cast<unsigned int*>(o), reinterpret_cast<uint2*>(o + n * 3 * pixels),
        reinterpret_cast<uint2*>(o + n * 3 * pixels + n * pixels * 2));
    sgr_count_launches_(1);
    return cudaGetLastError() == cudaSuccess ? SGR_OK : sgr_set_error_(SGR_E_CUDA, "wire_pack launch failed");
}

int sgr_wire_unpack(const void* recv, int32_t world, int64_t rank_stride_bytes, int32_t num_views, int64_t pixels,
                    int32_t compact, float* final_stack, int64_t views_per_rank, int64_t first_view, void* stream) {
    if (world <= 0 || num_views < 0 || pixels < 0 || (pixels & 3) || views_per_rank < first_view + num_views ||
        (num_views > 0 && pixels > 0 && (!recv || !final_stack)))
        return sgr_set_error_(SGR_E_INVALID_ARGUMENT, "bad wire_unpack arguments (pixels per image must be a multiple of 4)");
    if (num_views == 0 || pixels == 0) return SGR_OK;
    const long long c = num_views, P4 = pixels / 4;
    const int grid = sgr::launch_grid((long long)world * c * 5 * P4);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (compact)
        sgr::wire_unpack_kernel<true><<<grid, 256, 0, st>>>(static_cast<const unsigned char*>(recv), rank_stride_bytes, world, c,
                                                            P4, reinterpret_cast<float4*>(final_stack), views_per_rank, first_view);
    else
        sgr::wire_unpack_kernel<false><<<grid, 256, 0, st>>>(static_cast<const unsigned char*>(recv), rank_stride_bytes, world,
                                                             c, P4, reinterpret_cast<float4*>(final_stack), views_per_rank, first_view);
    sgr_count_launches_(1);
    return cudaGetLastError() == cudaSuccess ? SGR_OK : sgr_set_error_(SGR_E_CUDA, "wire_unpack launch failed");
}
}  // extern "C"
