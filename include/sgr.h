/* sgr.h — C ABI of the B200-native Gaussian-splat rasteriser (libsgr_b200.so).
 *
 * Drop-in boundary for the rasteriser hot path of yyvhang/SIGMAN_release.  The reference reaches this
 * path only through the third-party pybind module `diff_gaussian_rasterization._C`
 * (`rasterize_gaussians`, `rasterize_gaussians_backward`, `mark_visible`), called from
 * /root/reference/core/gaussians/gs.py:82-106, and through `simple_knn._C.distCUDA2`
 * (gs.py:6,70).  Each entry point below names the reference interface it replaces.
 *
 * Conventions
 *  - plain C: device pointers + sizes, no torch / C++ types.  All tensors fp32, contiguous.
 *  - the library never allocates user-visible memory: outputs, `state` (kept between forward and
 *    backward; replaces upstream's geomBuffer/binningBuffer/imgBuffer) and `scratch` are caller-owned
 *    device buffers sized with sgr_state_bytes() / sgr_scratch_bytes().
 *  - everything is enqueued on `stream` (a cudaStream_t passed as void*); no device synchronisation,
 *    no device->host copy on the launch path.  The number of (Gaussian, tile) instances is counted on
 *    the device; if it exceeds `max_instances` the affected renders are emitted as background only and
 *    the overflow is reported by sgr_read_status() so the caller can grow `state` and retry.
 *  - renders are batched: B subjects x V views per subject in one call (B = V = 1 reproduces one
 *    upstream rasterizer call).  Render index r = b * V + v.
 *  - return value: 0 on success, negative SgrError otherwise; sgr_last_error() gives the message
 *    (thread-local).  The library never calls abort()/exit().
 *  - threading: one host thread drives the library per process (one process per GPU, like the reference under
 *    accelerate).  The error string is thread-local, but the profiling accumulators (sgr_profile_*), the launch
 *    counter and the per-device launch caches are unsynchronised process globals.
 */
#ifndef SGR_H_
#define SGR_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SGR_ABI_VERSION 3

typedef enum SgrError {
    SGR_OK = 0,
    SGR_E_INVALID_ARGUMENT = -1,
    SGR_E_BUFFER_TOO_SMALL = -2,   /* state / scratch smaller than sgr_*_bytes() */
    SGR_E_CUDA = -3,               /* a CUDA runtime call failed; message in sgr_last_error() */
    SGR_E_INSTANCE_OVERFLOW = -4   /* reported by sgr_read_status(): grow max_instances and retry */
} SgrError;

enum {
    SGR_FLAG_SIMPLE_BLEND = 1,     /* use the straightforward (upstream-shaped) blend kernels: debugging aid */
    SGR_FLAG_CLAMP_COLOR = 2,      /* fuse gs.py:107 `rendered_image.clamp(0, 1)` into the blend epilogue */
    SGR_FLAG_FORWARD_ONLY = 4,     /* no backward will follow: skip the forward's bookkeeping for it (the refinement
                                      of the per-instance cull masks to the quarters that actually blended) */
    SGR_FLAG_TILE_TIMING = 8,      /* record per-tile forward blend timings (sgr_debug_copy_state: tile_timing) */
    SGR_FLAG_EXACT_EXP = 16        /* evaluate exp(power) with the oracle's fixed IEEE fp32 sequence (exp_spec) instead of
                                      the SFU's ex2: colour / depth / alpha / n_contrib become bit-identical to
                                      oracle/sgr_oracle.cpp at ~1.4x the blend time.  Default (flag clear): exp(power) =
                                      ex2.approx(power * log2 e) like upstream's own `exp()`; results agree with the oracle
                                      to ~1e-6 (the north-star tolerance is 1e-4), index outputs that do not depend on
                                      exp (radii, tile ranges, sorted lists) stay bit-exact */
};

/* Problem description shared by forward and backward.
 * Replaces the argument list of upstream `_C.rasterize_gaussians(...)` as assembled from
 * GaussianRasterizationSettings at gs.py:82-95 and the tensors at gs.py:99-106. */
typedef struct SgrProblem {
    int32_t num_subjects;        /* B */
    int32_t views_per_subject;   /* V */
    int32_t num_gaussians;       /* N (per subject) */
    int32_t image_height;        /* gs.py:83 */
    int32_t image_width;         /* gs.py:84 */
    float tanfovx;               /* gs.py:85 */
    float tanfovy;               /* gs.py:86 */
    const float* means3D;        /* [B,N,3]  gs.py:100 */
    const float* cov3D;          /* [B,N,6]  gs.py:105 (xx,xy,xz,yy,yz,zz; order of gs.py:32-37) */
    const float* colors;         /* [B,N,3]  gs.py:103 colors_precomp */
    const float* opacities;      /* [B,N]    gs.py:104 */
    const float* viewmatrix;     /* [B,V,16] gs.py:89  (W2C transposed, i.e. flat column-major W2C) */
    const float* projmatrix;     /* [B,V,16] gs.py:90  (flat column-major P*W2C) */
    const float* bg;             /* [3]      gs.py:87 */
    uint64_t max_instances;      /* capacity of the instance arrays inside `state` (all renders together) */
    uint64_t max_block_records;  /* capacity of the per-8x4-pixel-block record lists inside `state` (an instance appears
                                    once in the list of every block of its tile that its extent touches: typically
                                    1.3-2x max_instances, at most 8x); ignored with SGR_FLAG_SIMPLE_BLEND */
    int32_t renders_per_chunk;   /* 0 = library default; renders processed per launch set */
    int32_t flags;               /* SGR_FLAG_* */
    uint32_t max_tile_instances_hint; /* longest per-tile list seen for this kind of scene (SgrStatus of an earlier
                                    call), 0 = unknown: only sizes the shared memory of the long-list sort; any value
                                    is correct */
    uint32_t reserved;
} SgrProblem;

/* Replaces `_C.rasterize_gaussians` (forward).  Outputs match the tuple unpacked at gs.py:99:
 * color [B,V,3,H,W], radii [B,V,N] int32, depth [B,V,1,H,W], alpha [B,V,1,H,W]. */
typedef struct SgrForwardArgs {
    SgrProblem p;
    float* out_color;
    float* out_depth;
    float* out_alpha;
    int32_t* radii;
    void* state;    uint64_t state_bytes;
    void* scratch;  uint64_t scratch_bytes;
    void* stream;
    /* Optional fused reconstruction loss: replaces gs.py:107 `clamp(0, 1)` followed by
     * /root/reference/core/loss/whole_loss.py:130 `loss_l1 = l1(pred * mask, gt * mask)` (an UNREDUCED |x - y|) and its
     * reduction `torch.sum(loss_l1) / loss_l1.shape[0]` (whole_loss.py:139; shape[0] = B*V) and autograd backward.
     * When loss_target != NULL the blend epilogue writes the CLAMPED colour, accumulates
     * sum |clamp(c) * m - t * m| * loss_scale into *loss_out (device scalar, deterministic summation order) and
     * writes d loss / d colour (w.r.t. the unclamped colour: zero where the clamp saturated) to loss_dL_dcolor,
     * which sgr_backward takes as its loss_dL_dcolor.  Not available with SGR_FLAG_SIMPLE_BLEND. */
    const float* loss_target;    /* [B,V,3,H,W] or NULL (no fused loss) */
    const float* loss_mask;      /* [B,V,1,H,W] or NULL (mask of ones) */
    float* loss_dL_dcolor;       /* [B,V,3,H,W] */
    float* loss_out;             /* device scalar */
    float loss_scale;            /* 1 / (B*V) for the reference's reduction; 1 / (B*V*3*H*W) for a mean */
    /* Optional LPIPS feed: replaces `F.interpolate(pred * 2 - 1, (H/2, W/2), mode='bilinear',
     * align_corners=False)` of whole_loss.py:132-136 for even H, W (a factor-2 bilinear resize is the 2x2 box mean):
     * out_lpips_feed[b,v,c,y,x] = mean of clamp(colour, 0, 1) over the 2x2 pixels * 2 - 1.  Needs SGR_FLAG_CLAMP_COLOR
     * or the fused loss (the clamped colour is what the reference resizes).  NULL = not wanted. */
    float* out_lpips_feed;       /* [B,V,3,H/2,W/2] or NULL */
} SgrForwardArgs;

/* Replaces `_C.rasterize_gaussians_backward`.  Gradient slots follow the tuple returned by upstream's
 * autograd Function: means3D, means2D, colors_precomp, opacities, cov3D_precomp.
 * dL_ddepth / dL_dalpha may be NULL (treated as zeros - SIGMAN's case, gs.py:107-112 uses colour only).
 * The gradient w.r.t. the (unclamped) colour is assembled per pixel from up to three sources:
 *   dL_dcolor        the caller's gradient w.r.t. the colour image the forward returned (NULL = zeros);
 *   dL_dlpips_feed   the caller's gradient w.r.t. out_lpips_feed (NULL = none): adds 0.5 * g[y/2, x/2];
 *   loss_dL_dcolor   what the forward's fused loss wrote, times the device scalar *dL_dcolor_scale (NULL = none).
 * With SGR_FLAG_CLAMP_COLOR (or the fused loss) the first two are gradients w.r.t. the CLAMPED image and are zeroed
 * where the forward's clamp saturated (torch's clamp rule; the mask is kept in `state`) - gs.py:107 needs no
 * separate clamp kernel or saved tensor.
 * Gradients of one subject are summed over its V views; dL_dmeans2D (optional, may be NULL) is per
 * render: [B,V,N,3] with z = 0, in NDC units like upstream's viewspace-point gradient. */
typedef struct SgrBackwardArgs {
    SgrProblem p;
    const float* out_alpha;      /* [B,V,1,H,W] as produced by the forward */
    const int32_t* radii;        /* [B,V,N] as produced by the forward */
    const float* dL_dcolor;      /* [B,V,3,H,W] or NULL */
    const float* dL_ddepth;      /* [B,V,1,H,W] or NULL */
    const float* dL_dalpha;      /* [B,V,1,H,W] or NULL */
    float* dL_dmeans3D;          /* [B,N,3] */
    float* dL_dcov3D;            /* [B,N,6] */
    float* dL_dcolors;           /* [B,N,3] */
    float* dL_dopacities;        /* [B,N] */
    float* dL_dmeans2D;          /* [B,V,N,3] or NULL */
    void* state;    uint64_t state_bytes;
    void* scratch;  uint64_t scratch_bytes;
    void* stream;
    const float* loss_dL_dcolor;  /* [B,V,3,H,W] written by the forward's fused loss, or NULL */
    const float* dL_dcolor_scale; /* device scalar multiplying loss_dL_dcolor (the upstream gradient of the fused loss),
                                     or NULL (= 1) */
    const float* dL_dlpips_feed;  /* [B,V,3,H/2,W/2] or NULL */
    int32_t fused_clamp;          /* != 0: the forward ran with a fused loss (its colour output is clamped) */
    int32_t reserved;
} SgrBackwardArgs;

/* Device-side status of the last forward that used `state` (read with sgr_read_status).  It is stored in the first
 * sizeof(SgrStatus) bytes of `state`, so a caller that must not synchronise can also fetch it with its own
 * asynchronous device->host copy of the head of `state`. */
typedef struct SgrStatus {
    uint64_t instances_required;   /* total (Gaussian, tile) instances of all renders */
    uint64_t instances_capacity;   /* max_instances the forward ran with */
    uint32_t overflow;             /* != 0: some renders / tiles were dropped (background only);
                                      bit 0: max_instances too small, bit 1: max_block_records too small */
    uint32_t max_tile_instances;   /* longest per-tile list */
    uint32_t nonempty_tiles;       /* tiles with at least one instance (all renders) */
    uint32_t reserved;
    uint64_t block_records_required; /* block records of all renders (0 with SGR_FLAG_SIMPLE_BLEND) */
    uint64_t block_records_capacity; /* max_block_records the forward ran with */
} SgrStatus;

int sgr_abi_version(void);
const char* sgr_last_error(void);

/* Buffer sizes for a problem shape (bytes; 256-byte aligned). */
uint64_t sgr_state_bytes(int32_t num_subjects, int32_t views_per_subject, int32_t num_gaussians,
                         int32_t image_height, int32_t image_width, uint64_t max_instances,
                         uint64_t max_block_records, int32_t flags);
uint64_t sgr_scratch_bytes(int32_t num_subjects, int32_t views_per_subject, int32_t num_gaussians,
                           int32_t image_height, int32_t image_width, uint64_t max_instances,
                           int32_t renders_per_chunk);

int sgr_forward(const SgrForwardArgs* args);
int sgr_backward(const SgrBackwardArgs* args);

/* Copies the status block of `state` to the host.  Synchronises `stream` (the only synchronising call).
 * Returns SGR_E_INSTANCE_OVERFLOW if the forward overflowed max_instances (status is still filled in). */
int sgr_read_status(const void* state, void* stream, SgrStatus* host_status);

/* Replaces `_C.mark_visible` (upstream GaussianRasterizer.markVisible): visible[i] = view-space z > 0.2. */
int sgr_mark_visible(const float* means3D, int32_t num_gaussians, const float* viewmatrix,
                     const float* projmatrix, uint8_t* visible, void* stream);

/* Optional input paths of the upstream API (unused by SIGMAN, which passes cov3D_precomp and
 * colors_precomp): cov3D from scale * modifier and a (r,x,y,z) quaternion, and its backward. */
int sgr_cov3d_from_scale_rot(const float* scales, const float* rotations, float scale_modifier,
                             int32_t num_gaussians, float* cov3D, void* stream);
int sgr_cov3d_from_scale_rot_backward(const float* scales, const float* rotations, float scale_modifier,
                                      int32_t num_gaussians, const float* dL_dcov3D, float* dL_dscales,
                                      float* dL_drotations, void* stream);

/* Replaces the per-subject preparation of gs.py:69-73 for n = B*N Gaussians in one kernel:
 * scale = (scale_raw + 1) * sqrt(max(dist2, 1e-7)) with the kNN factor treated as a constant (gs.py:70-72),
 * Sigma = R diag(scale^2) R^T (get_covariance, gs.py:17-23), packed (xx,xy,xz,yy,yz,zz) (strip_lowerdiag, gs.py:29-38).
 * rotation is the row-major 3x3 "rotation-like" matrix SIGMAN feeds as gaussians['cov3d'].  bf16_autocast != 0
 * reproduces the operand / result rounding of the reference's two bmm's under accelerate's bf16 autocast. */
int sgr_prep_cov3d(const float* scale_raw, const float* rotation, const float* dist2, int64_t n, int32_t bf16_autocast,
                   float* cov3D, void* stream);
int sgr_prep_cov3d_backward(const float* scale_raw, const float* rotation, const float* dist2, int64_t n,
                            int32_t bf16_autocast, const float* dL_dcov3D, float* dL_dscale_raw, float* dL_drotation,
                            void* stream);

/* Optional colour path of the upstream API (`shs` instead of `colors_precomp`; unused by SIGMAN, gs.py:91,102):
 * upstream computeColorFromSH and its backward.  shs [N, max_coeffs, 3], degree 0..3 uses the first (degree+1)^2
 * coefficients; colors = max(0, SH(normalize(mean - campos)) + 0.5), clamped[N,3] records the clamp.
 * The backward writes dL_dshs [N, max_coeffs, 3] and dL_dmeans3D [N,3] (the view-direction term only). */
int sgr_sh_colors(const float* means3D, const float* shs, const float* campos, int32_t num_gaussians, int32_t degree,
                  int32_t max_coeffs, float* colors, uint8_t* clamped, void* stream);
int sgr_sh_colors_backward(const float* means3D, const float* shs, const float* campos, int32_t num_gaussians,
                           int32_t degree, int32_t max_coeffs, const uint8_t* clamped, const float* dL_dcolors,
                           float* dL_dshs, float* dL_dmeans3D, void* stream);

/* Replaces `simple_knn._C.distCUDA2` (gs.py:70): mean squared distance to the 3 nearest other points.
 * `scratch` must hold sgr_knn_scratch_bytes(num_points) bytes. */
uint64_t sgr_knn_scratch_bytes(int32_t num_points);
int sgr_knn_mean_dist2(const float* points, int32_t num_points, float* out_mean_dist2, void* scratch,
                       uint64_t scratch_bytes, void* stream);
/* The same for num_subjects independent point sets [B,N,3] -> [B,N] in one launch set (the reference calls
 * distCUDA2 once per subject inside its Python loop, gs.py:62-70). */
uint64_t sgr_knn_scratch_bytes_batched(int32_t num_subjects, int32_t num_points);
int sgr_knn_mean_dist2_batched(const float* points, int32_t num_subjects, int32_t num_points, float* out_mean_dist2,
                               void* scratch, uint64_t scratch_bytes, void* stream);

/* Wire formats of the multi-GPU image gather (BASELINE config 4; replaces the fp32
 * `self.accelerator.gather(out['images_pred'])` of /root/reference/core/loss/eval.py:81-82).  A chunk of n rendered views
 * travels as one byte buffer [RGB n*3*P | depth n*P | alpha n*P]: exact = float32 (the forward can write its outputs
 * straight into it: 20 bytes per pixel), compact = uint8 RGB (trunc(clamp(c, 0, 1) * 255 + 0.5)) + fp16 depth / alpha
 * (7 bytes per pixel).  `pixels` (H*W) must be a multiple of 4.
 * sgr_wire_pack: float32 planes -> the compact layout.
 * sgr_wire_unpack: the all-gathered buffers of `world` ranks (rank r at recv + r * rank_stride_bytes, either layout) ->
 * float32 final_stack[world][views_per_rank][5][pixels] at views first_view .. first_view + num_views of every rank. */
uint64_t sgr_wire_chunk_bytes(int32_t num_views, int64_t pixels, int32_t compact);
int sgr_wire_pack(const float* color, const float* depth, const float* alpha, int32_t num_views, int64_t pixels,
                  void* out_bytes, void* stream);
int sgr_wire_unpack(const void* recv, int32_t world, int64_t rank_stride_bytes, int32_t num_views, int64_t pixels,
                    int32_t compact, float* final_stack, int64_t views_per_rank, int64_t first_view, void* stream);

/* Measurement hooks (bench.py): per-stage device time with CUDA events recorded on the launch stream around every
 * stage of sgr_forward / sgr_backward while enabled, and a counter of the kernels this library has launched.
 * Stage order: preprocess, plan, scatter, sort, blend_forward, blend_backward, preprocess_backward.
 * sgr_profile_collect synchronises on the recorded events, returns summed milliseconds and the number of timed
 * stage invocations, and resets the accumulators. */
#define SGR_NUM_STAGES 7
void sgr_profile_enable(int on);
int sgr_profile_collect(double* stage_ms, uint32_t* stage_invocations);
uint64_t sgr_launch_count(void);

/* Debug/inspection (used by the parity tests): copies intermediate state of render `render` of the last forward
 * to caller-provided DEVICE buffers (any may be NULL), enqueued on `stream`:
 *   tile_ranges  uint32[tiles][2]  (start, end) offsets of each tile's depth-ordered list, relative to the render's
 *                                  first instance — upstream's `ranges` (identifyTileRanges)
 *   n_contrib    uint32[H*W]       upstream's per-pixel n_contrib
 *   point_list   uint32[point_list_capacity]  Gaussian index of every instance of the render in sorted order —
 *                                  upstream's binningState.point_list restricted to the render
 *   tile_timing  uint32[tiles][2]  (start, duration) of the tile's forward blend in ns of the GPU global timer
 *                                  (low 32 bits; load-balance diagnostics; needs SGR_FLAG_TILE_TIMING)
 * The shape arguments must be those of the forward that filled `state`. */
int sgr_debug_copy_state(const void* state, int32_t num_subjects, int32_t views_per_subject, int32_t num_gaussians,
                         int32_t image_height, int32_t image_width, uint64_t max_instances,
                         uint64_t max_block_records, int32_t flags, int32_t render,
                         uint32_t* tile_ranges, uint32_t* n_contrib, uint32_t* point_list,
                         uint64_t point_list_capacity, uint32_t* tile_timing, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SGR_H_ */
