// sgr_api.cu — the C ABI of libsgr_b200.so (include/sgr.h): argument validation, buffer layouts and the launch
// sequence of one batched forward / backward.  No device synchronisation and no device->host copy on the launch
// path; errors are returned as codes with a thread-local message (never abort/exit — the reference's caller wraps the
// render in a bare try/except, /root/reference/core/modules/autoencoder.py:349-361).
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <vector>

#include "sgr_common.cuh"

namespace sgr {

cudaError_t launch_blend_forward_simple(const ChunkCtx& c, float* out_color, float* out_depth, float* out_alpha);
cudaError_t launch_blend_backward_simple(const ChunkCtx& c, const float* out_alpha, const float* dL_dcolor,
                                         const float* dL_ddepth, const float* dL_dalpha);

namespace {

thread_local char g_err[512] = "";

int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

#define SGR_CUDA(expr)                                                                               \
    do {                                                                                             \
        cudaError_t e__ = (expr);                                                                    \
        if (e__ != cudaSuccess)                                                                      \
            return fail(SGR_E_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__); \
    } while (0)

// ---- optional per-stage timing (bench.py roofline leg) and the launch counter -------------------------------------
enum Stage { kStPreprocess = 0, kStPlan, kStScatter, kStSort, kStBlendFwd, kStBlendBwd, kStPreBwd, kNumStages };
static_assert(kNumStages == SGR_NUM_STAGES, "stage list and SGR_NUM_STAGES differ");
const int kStageLaunches[kNumStages] = {1, 1, 1, 2, 1, 1, 1};

struct StageEvent { int stage; cudaEvent_t beg, end; };
bool g_profile = false;
std::vector<StageEvent> g_events;
unsigned long long g_launches = 0;

struct StageTimer {
    int stage; cudaStream_t stream; cudaEvent_t beg = nullptr, end = nullptr;
    StageTimer(int st, cudaStream_t s) : stage(st), stream(s) {
        g_launches += kStageLaunches[st];
        if (g_profile && cudaEventCreate(&beg) == cudaSuccess && cudaEventCreate(&end) == cudaSuccess)
            cudaEventRecord(beg, stream);
    }
    ~StageTimer() {
        if (beg && end) { cudaEventRecord(end, stream); g_events.push_back({stage, beg, end}); }
    }
};
#define SGR_STAGE(stage, expr)                  \
    do {                                        \
        StageTimer timer__(stage, stream);      \
        SGR_CUDA(expr);                         \
    } while (0)

int chunk_size(const SgrProblem& p) {
    const long long R = (long long)p.num_subjects * p.views_per_subject;
    long long rpc = p.renders_per_chunk > 0 ? p.renders_per_chunk : kDefaultRendersPerChunk;
    return int(rpc < R ? rpc : R);
}

int check_problem(const SgrProblem& p) {
    if (p.num_subjects <= 0 || p.views_per_subject <= 0) return fail(SGR_E_INVALID_ARGUMENT, "num_subjects and views_per_subject must be positive");
    if (p.num_gaussians < 0) return fail(SGR_E_INVALID_ARGUMENT, "num_gaussians must be >= 0");
    if (p.image_height <= 0 || p.image_width <= 0) return fail(SGR_E_INVALID_ARGUMENT, "image size must be positive");
    if (tiles_x(p.image_width) > 0xffff || tiles_y(p.image_height) > 0xffff) return fail(SGR_E_INVALID_ARGUMENT, "image too large (more than 65535 tiles per axis)");
    if (!(p.tanfovx > 0.0f) || !(p.tanfovy > 0.0f)) return fail(SGR_E_INVALID_ARGUMENT, "tanfovx/tanfovy must be positive");
    if (p.num_gaussians > 0 && (!p.means3D || !p.cov3D || !p.colors || !p.opacities)) return fail(SGR_E_INVALID_ARGUMENT, "null Gaussian attribute pointer");
    if (!p.viewmatrix || !p.projmatrix || !p.bg) return fail(SGR_E_INVALID_ARGUMENT, "null camera / background pointer");
    if (p.renders_per_chunk < 0) return fail(SGR_E_INVALID_ARGUMENT, "renders_per_chunk must be >= 0");
    if (p.max_instances >= (1ull << 32)) return fail(SGR_E_INVALID_ARGUMENT, "max_instances must be below 2^32");
    if (p.max_block_records >= (1ull << 32)) return fail(SGR_E_INVALID_ARGUMENT, "max_block_records must be below 2^32");
    // plan_kernel scans (instances << 24 | empty tiles) in one 64-bit word: a chunk must hold fewer than 2^24 tiles
    const long long chunk_tiles = (long long)chunk_size(p) * tiles_x(p.image_width) * tiles_y(p.image_height);
    if (chunk_tiles >= (1ll << 24))
        return fail(SGR_E_INVALID_ARGUMENT, "renders_per_chunk * tiles per render must stay below 2^24 (got %lld); lower renders_per_chunk", chunk_tiles);
    return SGR_OK;
}

StateLayout state_layout(const SgrProblem& p) {
    return make_state_layout(p.num_subjects, p.views_per_subject, p.num_gaussians, p.image_height, p.image_width,
                             p.max_instances, p.max_block_records, (p.flags & SGR_FLAG_SIMPLE_BLEND) != 0);
}

void fill_ctx(ChunkCtx& c, const SgrProblem& p, void* state, void* scratch, cudaStream_t stream) {
    const int rpc = chunk_size(p);
    const StateLayout S = state_layout(p);
    const ScratchLayout X = make_scratch_layout(p.num_subjects, p.views_per_subject, p.num_gaussians, p.image_height,
                                                p.image_width, p.max_instances, rpc);
    char* s = static_cast<char*>(state);
    char* x = static_cast<char*>(scratch);
    c.g.N = p.num_gaussians; c.g.H = p.image_height; c.g.W = p.image_width;
    c.g.tiles_x = tiles_x(p.image_width); c.g.tiles_y = tiles_y(p.image_height);
    c.g.num_tiles = c.g.tiles_x * c.g.tiles_y; c.g.V = p.views_per_subject;
    c.g.tanfovx = p.tanfovx; c.g.tanfovy = p.tanfovy;
    c.p = &p;
    c.header = reinterpret_cast<StateHeader*>(s + S.header);
    c.tile_off = reinterpret_cast<unsigned int*>(s + S.tile_off);
    c.tile_cnt = reinterpret_cast<unsigned int*>(s + S.tile_cnt);
    c.tile_time = reinterpret_cast<uint2*>(s + S.tile_time);
    c.n_contrib = reinterpret_cast<unsigned int*>(s + S.n_contrib);
    c.sorted_ids = reinterpret_cast<unsigned int*>(s + S.sorted_ids);
    c.rec0 = reinterpret_cast<float4*>(s + S.rec0);
    c.rec1 = reinterpret_cast<float4*>(s + S.rec1);
    c.rec2 = reinterpret_cast<float4*>(s + S.rec2);
    c.clamp_mask = reinterpret_cast<unsigned char*>(s + S.clamp_mask);
    c.blk_off = reinterpret_cast<unsigned int*>(s + S.blk_off);
    c.blk_cnt = reinterpret_cast<unsigned int*>(s + S.blk_cnt);
    c.blk_eff = reinterpret_cast<unsigned int*>(s + S.blk_eff);
    c.bidx = reinterpret_cast<uint2*>(s + S.bidx);
    c.blk_capacity = (p.flags & SGR_FLAG_SIMPLE_BLEND) ? 0ull : p.max_block_records;
    c.ck0 = reinterpret_cast<float4*>(s + S.ck0);
    c.ck1 = reinterpret_cast<float*>(s + S.ck1);
    c.plan = reinterpret_cast<ChunkPlan*>(s + S.plan);
    c.bwd_items = reinterpret_cast<uint2*>(s + S.bwd_items);
    c.bwd_items_stride = S.bwd_items_stride;
    c.chunk_index = 0;
    c.keys = reinterpret_cast<unsigned long long*>(x + X.keys);
    c.keys_tmp = reinterpret_cast<unsigned long long*>(x + X.keys_tmp);
    c.g0 = reinterpret_cast<float4*>(s + S.g0);       // set_chunk() moves the three to the chunk's first render
    c.g1 = reinterpret_cast<float4*>(s + S.g1);
    c.g2 = reinterpret_cast<float4*>(s + S.g2);
    c.rect = reinterpret_cast<uint2*>(x + X.rect);
    c.cursor = reinterpret_cast<unsigned int*>(x + X.cursor);
    c.work_blend = reinterpret_cast<unsigned int*>(x + X.work_blend);
    c.work_empty = reinterpret_cast<unsigned int*>(x + X.work_empty);
    c.work_counts = reinterpret_cast<WorkCounts*>(x + X.work_counts);
    c.dense_items = reinterpret_cast<unsigned int*>(x + X.dense_items);
    c.loss_part = reinterpret_cast<float*>(x + X.loss_part);
    c.accum = reinterpret_cast<float*>(x + X.accum);
    c.loss_target = nullptr; c.loss_mask = nullptr; c.loss_dL_dcolor = nullptr; c.loss_scale = 0.0f;
    c.stream = stream;
}

// Points the context at the chunk of renders [r0, r0 + rpc): plan slot and slice of the per-Gaussian records.
void set_chunk(ChunkCtx& c, const SgrProblem& p, void* state, int r0, int rpc, int R) {
    const StateLayout S = state_layout(p);
    char* s = static_cast<char*>(state);
    c.render_base = r0;
    c.num_renders = (R - r0 < rpc) ? (R - r0) : rpc;
    c.chunk_index = r0 / rpc;
    c.plan = reinterpret_cast<ChunkPlan*>(s + S.plan) + c.chunk_index;
    const size_t first = size_t(r0) * p.num_gaussians;
    c.g0 = reinterpret_cast<float4*>(s + S.g0) + first;
    c.g1 = reinterpret_cast<float4*>(s + S.g1) + first;
    c.g2 = reinterpret_cast<float4*>(s + S.g2) + first;
}

__global__ void debug_ranges_kernel(const unsigned int* tile_off, const unsigned int* tile_cnt, int num_tiles,
                                    unsigned int* ranges) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= num_tiles) return;
    const unsigned int base = tile_off[0];
    const unsigned int c = tile_cnt[t];
    ranges[2 * t] = c ? tile_off[t] - base : 0u;
    ranges[2 * t + 1] = c ? tile_off[t] - base + c : 0u;
}

__global__ void debug_point_list_kernel(const unsigned int* tile_off, const unsigned int* tile_cnt, int num_tiles,
                                        const unsigned int* sorted_ids, unsigned int* out, unsigned long long cap) {
    const unsigned long long base = tile_off[0];
    const unsigned long long n = (unsigned long long)tile_off[num_tiles - 1] + tile_cnt[num_tiles - 1] - base;
    for (unsigned long long k = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; k < n && k < cap;
         k += (unsigned long long)gridDim.x * blockDim.x)
        out[k] = sorted_ids[base + k];
}

}  // namespace
}  // namespace sgr

using namespace sgr;

extern "C" {

int sgr_set_error_(int code, const char* msg) { return fail(code, "%s", msg); }   // used by sgr_knn.cu
void sgr_count_launches_(unsigned int n) { g_launches += n; }

int sgr_abi_version(void) { return SGR_ABI_VERSION; }
const char* sgr_last_error(void) { return g_err; }

uint64_t sgr_state_bytes(int32_t B, int32_t V, int32_t N, int32_t H, int32_t W, uint64_t max_instances,
                         uint64_t max_block_records, int32_t flags) {
    if (B <= 0 || V <= 0 || N < 0 || H <= 0 || W <= 0) return 0;
    return make_state_layout(B, V, N, H, W, max_instances, max_block_records, (flags & SGR_FLAG_SIMPLE_BLEND) != 0).total;
}

uint64_t sgr_scratch_bytes(int32_t B, int32_t V, int32_t N, int32_t H, int32_t W, uint64_t max_instances,
                           int32_t renders_per_chunk) {
    if (B <= 0 || V <= 0 || N < 0 || H <= 0 || W <= 0 || renders_per_chunk < 0) return 0;
    SgrProblem p;
    memset(&p, 0, sizeof(p));
    p.num_subjects = B; p.views_per_subject = V; p.renders_per_chunk = renders_per_chunk;
    return make_scratch_layout(B, V, N, H, W, max_instances, chunk_size(p)).total;
}

int sgr_forward(const SgrForwardArgs* args) {
    if (!args) return fail(SGR_E_INVALID_ARGUMENT, "null args");
    const SgrProblem& p = args->p;
    if (int rc = check_problem(p)) return rc;
    if (!args->out_color || !args->out_depth || !args->out_alpha || (p.num_gaussians > 0 && !args->radii))
        return fail(SGR_E_INVALID_ARGUMENT, "null output pointer");
    if (!args->state || !args->scratch) return fail(SGR_E_INVALID_ARGUMENT, "null state / scratch");
    const bool fused_loss = args->loss_target != nullptr;
    if (fused_loss && (!args->loss_dL_dcolor || !args->loss_out))
        return fail(SGR_E_INVALID_ARGUMENT, "fused loss needs loss_dL_dcolor and loss_out");
    if (fused_loss && (p.flags & SGR_FLAG_SIMPLE_BLEND))
        return fail(SGR_E_INVALID_ARGUMENT, "the fused loss is not available with SGR_FLAG_SIMPLE_BLEND");
    if (args->out_lpips_feed) {
        if (p.flags & SGR_FLAG_SIMPLE_BLEND) return fail(SGR_E_INVALID_ARGUMENT, "out_lpips_feed is not available with SGR_FLAG_SIMPLE_BLEND");
        if ((p.image_height | p.image_width) & 1) return fail(SGR_E_INVALID_ARGUMENT, "out_lpips_feed needs even image_height / image_width");
        if (!fused_loss && !(p.flags & SGR_FLAG_CLAMP_COLOR)) return fail(SGR_E_INVALID_ARGUMENT, "out_lpips_feed needs SGR_FLAG_CLAMP_COLOR or the fused loss (the reference resizes the clamped image)");
    }
    const int rpc = chunk_size(p);
    const uint64_t need_state = state_layout(p).total;
    const uint64_t need_scratch = make_scratch_layout(p.num_subjects, p.views_per_subject, p.num_gaussians,
                                                      p.image_height, p.image_width, p.max_instances, rpc).total;
    if (args->state_bytes < need_state)
        return fail(SGR_E_BUFFER_TOO_SMALL, "state buffer too small: %llu < %llu bytes", (unsigned long long)args->state_bytes, (unsigned long long)need_state);
    if (args->scratch_bytes < need_scratch)
        return fail(SGR_E_BUFFER_TOO_SMALL, "scratch buffer too small: %llu < %llu bytes", (unsigned long long)args->scratch_bytes, (unsigned long long)need_scratch);

    cudaStream_t stream = static_cast<cudaStream_t>(args->stream);
    ChunkCtx c;
    fill_ctx(c, p, args->state, args->scratch, stream);
    if (fused_loss) {
        c.loss_target = args->loss_target; c.loss_mask = args->loss_mask; c.loss_dL_dcolor = args->loss_dL_dcolor;
        c.loss_scale = args->loss_scale;
    }
    const int R = p.num_subjects * p.views_per_subject;
    // tile_cnt and tile_time are adjacent in `state`: one memset (the status header is initialised by the plan kernel)
    SGR_CUDA(cudaMemsetAsync(c.tile_cnt, 0, (reinterpret_cast<char*>(c.tile_time) - reinterpret_cast<char*>(c.tile_cnt)) +
                                                size_t(R) * c.g.num_tiles * ((p.flags & SGR_FLAG_TILE_TIMING) ? 8 : 0), stream));
    for (int r0 = 0; r0 < R; r0 += rpc) {
        set_chunk(c, p, args->state, r0, rpc, R);
        if (p.num_gaussians > 0) SGR_STAGE(kStPreprocess, launch_preprocess(c, args->radii));
        SGR_STAGE(kStPlan, launch_plan(c));
        if (p.num_gaussians > 0) {
            SGR_STAGE(kStScatter, launch_scatter(c));
            SGR_STAGE(kStSort, launch_sort_tiles(c));
        }
        if (p.flags & SGR_FLAG_SIMPLE_BLEND) {
            SGR_STAGE(kStBlendFwd, launch_blend_forward_simple(c, args->out_color, args->out_depth, args->out_alpha));
        } else {
            SGR_STAGE(kStBlendFwd, launch_blend_forward(c, args->out_color, args->out_depth, args->out_alpha, args->out_lpips_feed));
            if (fused_loss) {
                SGR_CUDA(launch_loss_reduce(c, args->loss_out));
                g_launches += 1;
            }
        }
    }
    return SGR_OK;
}

int sgr_backward(const SgrBackwardArgs* args) {
    if (!args) return fail(SGR_E_INVALID_ARGUMENT, "null args");
    const SgrProblem& p = args->p;
    if (int rc = check_problem(p)) return rc;
    if (!args->out_alpha || (p.num_gaussians > 0 && !args->radii))
        return fail(SGR_E_INVALID_ARGUMENT, "null forward-result pointer");
    if (!args->dL_dcolor && !args->loss_dL_dcolor && !args->dL_dlpips_feed && !args->dL_ddepth && !args->dL_dalpha)
        return fail(SGR_E_INVALID_ARGUMENT, "no image-space gradient given (dL_dcolor, loss_dL_dcolor, dL_dlpips_feed, dL_ddepth, dL_dalpha all NULL)");
    if (p.num_gaussians > 0 && (!args->dL_dmeans3D || !args->dL_dcov3D || !args->dL_dcolors || !args->dL_dopacities))
        return fail(SGR_E_INVALID_ARGUMENT, "null gradient output pointer");
    if (!args->state || !args->scratch) return fail(SGR_E_INVALID_ARGUMENT, "null state / scratch");
    const int rpc = chunk_size(p);
    const uint64_t need_state = state_layout(p).total;
    const uint64_t need_scratch = make_scratch_layout(p.num_subjects, p.views_per_subject, p.num_gaussians,
                                                      p.image_height, p.image_width, p.max_instances, rpc).total;
    if (args->state_bytes < need_state)
        return fail(SGR_E_BUFFER_TOO_SMALL, "state buffer too small: %llu < %llu bytes", (unsigned long long)args->state_bytes, (unsigned long long)need_state);
    if (args->scratch_bytes < need_scratch)
        return fail(SGR_E_BUFFER_TOO_SMALL, "scratch buffer too small: %llu < %llu bytes", (unsigned long long)args->scratch_bytes, (unsigned long long)need_scratch);
    if (p.num_gaussians == 0) return SGR_OK;
    if ((p.flags & SGR_FLAG_SIMPLE_BLEND) && (!args->dL_dcolor || args->loss_dL_dcolor || args->dL_dlpips_feed || args->fused_clamp))
        return fail(SGR_E_INVALID_ARGUMENT, "SGR_FLAG_SIMPLE_BLEND takes dL_dcolor only (no fused loss / LPIPS feed gradients)");
    if (args->dL_dlpips_feed && ((p.image_height | p.image_width) & 1))
        return fail(SGR_E_INVALID_ARGUMENT, "dL_dlpips_feed needs even image_height / image_width");

    cudaStream_t stream = static_cast<cudaStream_t>(args->stream);
    ChunkCtx c;
    fill_ctx(c, p, args->state, args->scratch, stream);
    const int R = p.num_subjects * p.views_per_subject;
    const size_t BN = size_t(p.num_subjects) * p.num_gaussians;
    (void)BN;   // the per-subject gradients are stored (not accumulated) by the chunk holding the subject's first view
    for (int r0 = 0; r0 < R; r0 += rpc) {
        set_chunk(c, p, args->state, r0, rpc, R);
        SGR_CUDA(cudaMemsetAsync(c.accum, 0, size_t(c.num_renders) * p.num_gaussians * 4 * kAccumStride, stream));
        if (p.flags & SGR_FLAG_SIMPLE_BLEND) {
            SGR_STAGE(kStBlendBwd, launch_blend_backward_simple(c, args->out_alpha, args->dL_dcolor, args->dL_ddepth, args->dL_dalpha));
        } else {
            // the (block, segment) work lists were filled by the forward; only the queue head is reset here
            SGR_CUDA(cudaMemsetAsync(&c.plan->cursor, 0, 4, stream));
            SGR_STAGE(kStBlendBwd, launch_blend_backward(c, *args));
        }
        SGR_STAGE(kStPreBwd, launch_preprocess_backward(c, *args));
    }
    return SGR_OK;
}

int sgr_read_status(const void* state, void* stream, SgrStatus* host_status) {
    if (!state || !host_status) return fail(SGR_E_INVALID_ARGUMENT, "null state / status");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    SGR_CUDA(cudaMemcpyAsync(host_status, state, sizeof(SgrStatus), cudaMemcpyDeviceToHost, s));
    SGR_CUDA(cudaStreamSynchronize(s));
    if (host_status->overflow)
        return fail(SGR_E_INSTANCE_OVERFLOW, "instance overflow: %llu (Gaussian, tile) instances needed, capacity %llu; "
                    "%llu block records needed, capacity %llu",
                    (unsigned long long)host_status->instances_required, (unsigned long long)host_status->instances_capacity,
                    (unsigned long long)host_status->block_records_required, (unsigned long long)host_status->block_records_capacity);
    return SGR_OK;
}

void sgr_profile_enable(int on) { g_profile = on != 0; }

int sgr_profile_collect(double* stage_ms, uint32_t* stage_launches) {
    if (!stage_ms || !stage_launches) return fail(SGR_E_INVALID_ARGUMENT, "null profile output");
    for (int k = 0; k < kNumStages; ++k) { stage_ms[k] = 0.0; stage_launches[k] = 0; }
    for (StageEvent& e : g_events) {
        SGR_CUDA(cudaEventSynchronize(e.end));
        float ms = 0.0f;
        SGR_CUDA(cudaEventElapsedTime(&ms, e.beg, e.end));
        stage_ms[e.stage] += ms;
        stage_launches[e.stage] += 1;
        cudaEventDestroy(e.beg);
        cudaEventDestroy(e.end);
    }
    g_events.clear();
    return SGR_OK;
}

uint64_t sgr_launch_count(void) { return g_launches; }

int sgr_mark_visible(const float* means3D, int32_t N, const float* viewmatrix, const float* projmatrix,
                     uint8_t* visible, void* stream) {
    (void)projmatrix;
    if (N < 0 || (N > 0 && (!means3D || !visible)) || !viewmatrix) return fail(SGR_E_INVALID_ARGUMENT, "bad mark_visible arguments");
    SGR_CUDA(launch_mark_visible(means3D, N, viewmatrix, visible, static_cast<cudaStream_t>(stream)));
    return SGR_OK;
}

int sgr_cov3d_from_scale_rot(const float* scales, const float* rotations, float mod, int32_t N, float* cov3D,
                             void* stream) {
    if (N < 0 || (N > 0 && (!scales || !rotations || !cov3D))) return fail(SGR_E_INVALID_ARGUMENT, "bad cov3d arguments");
    SGR_CUDA(launch_cov3d_from_scale_rot(scales, rotations, mod, N, cov3D, static_cast<cudaStream_t>(stream)));
    return SGR_OK;
}

int sgr_cov3d_from_scale_rot_backward(const float* scales, const float* rotations, float mod, int32_t N,
                                      const float* dL_dcov3D, float* dL_dscales, float* dL_drotations, void* stream) {
    if (N < 0 || (N > 0 && (!scales || !rotations || !dL_dcov3D || !dL_dscales || !dL_drotations)))
        return fail(SGR_E_INVALID_ARGUMENT, "bad cov3d backward arguments");
    SGR_CUDA(launch_cov3d_from_scale_rot_backward(scales, rotations, mod, N, dL_dcov3D, dL_dscales, dL_drotations,
                                                  static_cast<cudaStream_t>(stream)));
    return SGR_OK;
}

int sgr_prep_cov3d(const float* scale_raw, const float* rotation, const float* dist2, int64_t n, int32_t bf16_autocast,
                   float* cov3D, void* stream) {
    if (n < 0 || (n > 0 && (!scale_raw || !rotation || !dist2 || !cov3D))) return fail(SGR_E_INVALID_ARGUMENT, "bad prep_cov3d arguments");
    SGR_CUDA(launch_prep_cov3d(scale_raw, rotation, dist2, n, bf16_autocast != 0, cov3D, static_cast<cudaStream_t>(stream)));
    g_launches += n > 0;
    return SGR_OK;
}

int sgr_prep_cov3d_backward(const float* scale_raw, const float* rotation, const float* dist2, int64_t n,
                            int32_t bf16_autocast, const float* dL_dcov3D, float* dL_dscale_raw, float* dL_drotation,
                            void* stream) {
    if (n < 0 || (n > 0 && (!scale_raw || !rotation || !dist2 || !dL_dcov3D || !dL_dscale_raw || !dL_drotation)))
        return fail(SGR_E_INVALID_ARGUMENT, "bad prep_cov3d backward arguments");
    SGR_CUDA(launch_prep_cov3d_backward(scale_raw, rotation, dist2, n, bf16_autocast != 0, dL_dcov3D, dL_dscale_raw,
                                        dL_drotation, static_cast<cudaStream_t>(stream)));
    g_launches += n > 0;
    return SGR_OK;
}

int sgr_sh_colors(const float* means3D, const float* shs, const float* campos, int32_t N, int32_t degree,
                  int32_t max_coeffs, float* colors, uint8_t* clamped, void* stream) {
    if (N < 0 || degree < 0 || degree > 3 || max_coeffs < (degree + 1) * (degree + 1) ||
        (N > 0 && (!means3D || !shs || !campos || !colors || !clamped)))
        return fail(SGR_E_INVALID_ARGUMENT, "bad sh_colors arguments (degree 0..3, max_coeffs >= (degree+1)^2)");
    SGR_CUDA(launch_sh_colors(means3D, shs, campos, N, degree, max_coeffs, colors, clamped, static_cast<cudaStream_t>(stream)));
    g_launches += N > 0;
    return SGR_OK;
}

int sgr_sh_colors_backward(const float* means3D, const float* shs, const float* campos, int32_t N, int32_t degree,
                           int32_t max_coeffs, const uint8_t* clamped, const float* dL_dcolors, float* dL_dshs,
                           float* dL_dmeans3D, void* stream) {
    if (N < 0 || degree < 0 || degree > 3 || max_coeffs < (degree + 1) * (degree + 1) ||
        (N > 0 && (!means3D || !shs || !campos || !clamped || !dL_dcolors || !dL_dshs || !dL_dmeans3D)))
        return fail(SGR_E_INVALID_ARGUMENT, "bad sh_colors backward arguments");
    SGR_CUDA(launch_sh_colors_backward(means3D, shs, campos, N, degree, max_coeffs, clamped, dL_dcolors, dL_dshs,
                                       dL_dmeans3D, static_cast<cudaStream_t>(stream)));
    g_launches += N > 0;
    return SGR_OK;
}

int sgr_debug_copy_state(const void* state, int32_t B, int32_t V, int32_t N, int32_t H, int32_t W,
                         uint64_t max_instances, uint64_t max_block_records, int32_t flags, int32_t render,
                         uint32_t* tile_ranges, uint32_t* n_contrib,
                         uint32_t* point_list, uint64_t point_list_capacity, uint32_t* tile_timing, void* stream) {
    if (!state || B <= 0 || V <= 0 || N < 0 || H <= 0 || W <= 0 || render < 0 || render >= B * V)
        return fail(SGR_E_INVALID_ARGUMENT, "bad debug_copy_state arguments");
    const StateLayout S = make_state_layout(B, V, N, H, W, max_instances, max_block_records, (flags & SGR_FLAG_SIMPLE_BLEND) != 0);
    const char* s = static_cast<const char*>(state);
    const int T = tiles_x(W) * tiles_y(H);
    const unsigned int* off = reinterpret_cast<const unsigned int*>(s + S.tile_off) + size_t(render) * T;
    const unsigned int* cnt = reinterpret_cast<const unsigned int*>(s + S.tile_cnt) + size_t(render) * T;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (tile_ranges) {
        debug_ranges_kernel<<<(T + 255) / 256, 256, 0, st>>>(off, cnt, T, tile_ranges);
        SGR_CUDA(cudaGetLastError());
    }
    if (n_contrib)
        SGR_CUDA(cudaMemcpyAsync(n_contrib, reinterpret_cast<const unsigned int*>(s + S.n_contrib) + size_t(render) * H * W,
                                 size_t(H) * W * 4, cudaMemcpyDeviceToDevice, st));
    if (tile_timing)
        SGR_CUDA(cudaMemcpyAsync(tile_timing, reinterpret_cast<const uint2*>(s + S.tile_time) + size_t(render) * T,
                                 size_t(T) * 8, cudaMemcpyDeviceToDevice, st));
    if (point_list && point_list_capacity) {
        debug_point_list_kernel<<<256, 256, 0, st>>>(off, cnt, T, reinterpret_cast<const unsigned int*>(s + S.sorted_ids),
                                                     point_list, point_list_capacity);
        SGR_CUDA(cudaGetLastError());
    }
    return SGR_OK;
}

}  // extern "C"
