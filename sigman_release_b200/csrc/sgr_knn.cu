// sgr_knn.cu — mean squared distance to the 3 nearest other points (replaces `simple_knn._C.distCUDA2`,
// /root/reference/core/gaussians/gs.py:70; oracle: knn_mean_dist2).
//
// Exact.  Points are binned into a uniform grid (counting sort by cell), and every point scans the cells around it
// shell by shell until the shell's distance bound exceeds its current third-best distance.  Distances are evaluated
// as (dx*dx + dy*dy) + dz*dz in fp32 without contraction (file built with --fmad=false), so the three smallest values
// — and therefore the result — are bit-identical to the brute-force oracle regardless of visiting order.
#include <cfloat>

#include "sgr_common.cuh"

namespace sgr {
namespace {

struct KnnHeader {
    float lo[3], hi[3];
    float cell, inv_cell;
    int dim[3];
    unsigned int num_cells;
};

struct KnnLayout {
    uint64_t header, cell_of, cell_start, cell_fill, block_sum, sorted, total;
};

constexpr unsigned int kMaxCells = 1u << 21;
constexpr int kScanThreads = 1024, kScanPer = 8;
constexpr unsigned int kScanBlock = kScanThreads * kScanPer;            // cells per scan CTA
constexpr unsigned int kScanBlocks = kMaxCells / kScanBlock;            // 256

// Per-subject regions, subject b at index b of every array.
KnnLayout knn_layout(int B, int N) {
    KnnLayout L;
    uint64_t o = 0;
    L.header = o;     o = align_up(o + uint64_t(B) * sizeof(KnnHeader));
    L.cell_of = o;    o = align_up(o + uint64_t(B) * N * 4);
    L.cell_start = o; o = align_up(o + uint64_t(B) * (kMaxCells + 1) * 4);
    L.cell_fill = o;  o = align_up(o + uint64_t(B) * kMaxCells * 4);
    L.block_sum = o;  o = align_up(o + uint64_t(B) * kScanBlocks * 4);
    L.sorted = o;     o = align_up(o + uint64_t(B) * N * 16);
    L.total = o;
    return L;
}

__device__ __forceinline__ float atomic_min_float(float* addr, float v) {   // valid for any sign via int/uint trick
    return (v >= 0.0f) ? __int_as_float(atomicMin(reinterpret_cast<int*>(addr), __float_as_int(v)))
                       : __uint_as_float(atomicMax(reinterpret_cast<unsigned int*>(addr), __float_as_uint(v)));
}
__device__ __forceinline__ float atomic_max_float(float* addr, float v) {
    return (v >= 0.0f) ? __int_as_float(atomicMax(reinterpret_cast<int*>(addr), __float_as_int(v)))
                       : __uint_as_float(atomicMin(reinterpret_cast<unsigned int*>(addr), __float_as_uint(v)));
}

__global__ void knn_init_kernel(KnnHeader* h, int B) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    for (int k = 0; k < 3; ++k) { h[b].lo[k] = FLT_MAX; h[b].hi[k] = -FLT_MAX; }
}

// grid (x, B): bounding box of every subject
__global__ void __launch_bounds__(256) knn_bbox_kernel(const float* __restrict__ pts_all, int N, KnnHeader* h_all) {
    const float* pts = pts_all + size_t(blockIdx.y) * N * 3;
    KnnHeader* h = h_all + blockIdx.y;
    float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < N; i += gridDim.x * blockDim.x)
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const float v = pts[3 * i + k];
            lo[k] = fminf(lo[k], v); hi[k] = fmaxf(hi[k], v);
        }
#pragma unroll
    for (int k = 0; k < 3; ++k) {
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
            lo[k] = fminf(lo[k], __shfl_xor_sync(0xffffffffu, lo[k], d));
            hi[k] = fmaxf(hi[k], __shfl_xor_sync(0xffffffffu, hi[k], d));
        }
        if ((threadIdx.x & 31) == 0) { atomic_min_float(&h->lo[k], lo[k]); atomic_max_float(&h->hi[k], hi[k]); }
    }
}

// Cell edge so that an average occupied neighbourhood holds a handful of points: the points are a surface sample, so
// the edge is derived from the bounding-box surface scale rather than its volume.
__global__ void knn_grid_kernel(KnnHeader* h_all, int N, int B) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    KnnHeader* h = h_all + b;
    float ext[3];
    for (int k = 0; k < 3; ++k) ext[k] = fmaxf(h->hi[k] - h->lo[k], 1e-12f);
    const float area = 2.0f * (ext[0] * ext[1] + ext[1] * ext[2] + ext[0] * ext[2]);
    float cell = sqrtf(area * 2.0f / float(N > 0 ? N : 1));
    const float emax = fmaxf(ext[0], fmaxf(ext[1], ext[2]));
    cell = fmaxf(cell, emax / 1024.0f);
    for (;;) {
        unsigned long long cells = 1;
        for (int k = 0; k < 3; ++k) {
            int d = int(ext[k] / cell) + 1;
            d = d < 1 ? 1 : d;
            h->dim[k] = d;
            cells *= (unsigned long long)d;
        }
        if (cells <= kMaxCells) { h->num_cells = (unsigned int)cells; break; }
        cell *= 1.26f;
    }
    h->cell = cell;
    h->inv_cell = 1.0f / cell;
}

__device__ __forceinline__ int3 cell_coord(const KnnHeader* h, float x, float y, float z) {
    int3 c;
    c.x = min(h->dim[0] - 1, max(0, int((x - h->lo[0]) * h->inv_cell)));
    c.y = min(h->dim[1] - 1, max(0, int((y - h->lo[1]) * h->inv_cell)));
    c.z = min(h->dim[2] - 1, max(0, int((z - h->lo[2]) * h->inv_cell)));
    return c;
}

__global__ void __launch_bounds__(256) knn_count_kernel(const float* __restrict__ pts_all, int N, const KnnHeader* h_all,
                                                        unsigned int* cell_of_all, unsigned int* cell_count_all) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const int b = blockIdx.y;
    const float* pts = pts_all + size_t(b) * N * 3;
    const KnnHeader* h = h_all + b;
    const int3 c = cell_coord(h, pts[3 * i], pts[3 * i + 1], pts[3 * i + 2]);
    const unsigned int id = (unsigned(c.z) * h->dim[1] + c.y) * h->dim[0] + c.x;
    cell_of_all[size_t(b) * N + i] = id;
    atomicAdd(cell_count_all + size_t(b) * (kMaxCells + 1) + id, 1u);
}

// Exclusive scan of the cell counts in three grid-wide steps (a single CTA walking up to 2^21 cells was the longest
// kernel of the whole kNN): per-CTA scans of kScanBlock cells + block totals, a scan of the <= 256 totals, offset add.
__global__ void __launch_bounds__(kScanThreads) knn_scan_blocks_kernel(const KnnHeader* h_all, unsigned int* cell_start_all,
                                                                       unsigned int* block_sum_all) {
    __shared__ unsigned int s_warp[32];
    const int b = blockIdx.y;
    const unsigned int n = h_all[b].num_cells;
    const unsigned int base = blockIdx.x * kScanBlock;
    if (base >= n) return;
    unsigned int* cell_start = cell_start_all + size_t(b) * (kMaxCells + 1);
    const int t = threadIdx.x;
    unsigned int v[kScanPer], sum = 0;
#pragma unroll
    for (int j = 0; j < kScanPer; ++j) {
        const unsigned int k = base + t * kScanPer + j;
        v[j] = k < n ? cell_start[k] : 0u;
        sum += v[j];
    }
    unsigned int inc = sum;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const unsigned int u = __shfl_up_sync(0xffffffffu, inc, d);
        if ((t & 31) >= d) inc += u;
    }
    if ((t & 31) == 31) s_warp[t >> 5] = inc;
    __syncthreads();
    if (t < 32) {
        unsigned int w = s_warp[t], wi = w;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const unsigned int u = __shfl_up_sync(0xffffffffu, wi, d);
            if (t >= d) wi += u;
        }
        s_warp[t] = wi - w;
        if (t == 31) block_sum_all[size_t(b) * kScanBlocks + blockIdx.x] = wi;
    }
    __syncthreads();
    unsigned int run = s_warp[t >> 5] + inc - sum;
#pragma unroll
    for (int j = 0; j < kScanPer; ++j) {
        const unsigned int k = base + t * kScanPer + j;
        if (k < n) cell_start[k] = run;
        run += v[j];
    }
}

__global__ void __launch_bounds__(kScanBlocks) knn_scan_tops_kernel(const KnnHeader* h_all, unsigned int* block_sum_all,
                                                                    unsigned int* cell_start_all, int N) {
    __shared__ unsigned int s_warp[kScanBlocks / 32];
    const int b = blockIdx.x, t = threadIdx.x;
    const unsigned int n = h_all[b].num_cells;
    const unsigned int nblk = (n + kScanBlock - 1) / kScanBlock;
    unsigned int* bs = block_sum_all + size_t(b) * kScanBlocks;
    const unsigned int v = unsigned(t) < nblk ? bs[t] : 0u;
    unsigned int inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const unsigned int u = __shfl_up_sync(0xffffffffu, inc, d);
        if ((t & 31) >= d) inc += u;
    }
    if ((t & 31) == 31) s_warp[t >> 5] = inc;
    __syncthreads();
    unsigned int off = 0;
    for (int w = 0; w < (t >> 5); ++w) off += s_warp[w];
    bs[t] = off + inc - v;
    if (t == 0) cell_start_all[size_t(b) * (kMaxCells + 1) + n] = unsigned(N);
}

__global__ void __launch_bounds__(256) knn_scan_add_kernel(const KnnHeader* h_all, const unsigned int* block_sum_all,
                                                           unsigned int* cell_start_all) {
    const int b = blockIdx.y;
    const unsigned int n = h_all[b].num_cells;
    for (unsigned int k = kScanBlock + blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x)
        cell_start_all[size_t(b) * (kMaxCells + 1) + k] += block_sum_all[size_t(b) * kScanBlocks + k / kScanBlock];
}

__global__ void __launch_bounds__(256) knn_fill_kernel(const float* __restrict__ pts_all, int N,
                                                       const unsigned int* __restrict__ cell_of_all,
                                                       const unsigned int* __restrict__ cell_start_all,
                                                       unsigned int* cell_fill_all, float4* sorted_all) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const int b = blockIdx.y;
    const float* pts = pts_all + size_t(b) * N * 3;
    const unsigned int c = cell_of_all[size_t(b) * N + i];
    const unsigned int pos = cell_start_all[size_t(b) * (kMaxCells + 1) + c] + atomicAdd(cell_fill_all + size_t(b) * kMaxCells + c, 1u);
    sorted_all[size_t(b) * N + pos] = make_float4(pts[3 * i], pts[3 * i + 1], pts[3 * i + 2], __int_as_float(i));
}

// Keeps the three smallest of {b0 <= b1 <= b2, d}: a branch-free min / max chain (five FMNMX).
__device__ __forceinline__ void insert3(float d, float& b0, float& b1, float& b2) {
    const float m0 = fmaxf(b0, d);
    b0 = fminf(b0, d);
    const float m1 = fmaxf(b1, m0);
    b1 = fminf(b1, m0);
    b2 = fminf(b2, m1);
}

// kLanes lanes per point, points IN CELL ORDER (lane group g takes sorted[g]): the lanes of a point stride through the
// candidates of every cell row of the point's ring (private three-best lists, merged per ring with shuffles), and the
// points of a warp — same or adjacent cells — walk rings of the same shape over the same cache lines.
//   kLanes = 1: the fewest instructions per point — what a launch that fills the machine wants (8 x 100 K points:
//               502 us per call against 578 us with eight lanes; tools/knn_bench.py);
//   kLanes = 8: a row of up to eight candidates costs one trip instead of eight — the latency of one subject's call
//               (100 K points: 121 us per call against 182 us with one lane).
// The three-best insertion is a branch-free min / max chain (with the branchy form: 591 / 248 us).

// Merges the three-best lists of a point's eight lanes: every lane returns the group's three smallest values.
__device__ __forceinline__ void merge3_group(float& b0, float& b1, float& b2, int lane) {
    const unsigned int gmask = 0xffu << (lane & 24);
    float m[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        float v = b0;
        v = fminf(v, __shfl_xor_sync(0xffffffffu, v, 1));
        v = fminf(v, __shfl_xor_sync(0xffffffffu, v, 2));
        v = fminf(v, __shfl_xor_sync(0xffffffffu, v, 4));
        m[r] = v;
        // exactly one lane holding the minimum pops it (equal distances of different candidates stay counted)
        const unsigned int owners = __ballot_sync(0xffffffffu, b0 == v) & gmask;
        if (owners && lane == __ffs(owners) - 1) { b0 = b1; b1 = b2; b2 = __int_as_float(0x7f800000); }
    }
    b0 = m[0]; b1 = m[1]; b2 = m[2];
}

template <int kKnnLanes>
__global__ void __launch_bounds__(128) knn_query_kernel(int N, const KnnHeader* h_all,
                                                        const unsigned int* __restrict__ cell_start_all,
                                                        const float4* __restrict__ sorted_all, float* __restrict__ out_all) {
    const int lane = threadIdx.x & 31, sub = lane & (kKnnLanes - 1);
    const int j_raw = (blockIdx.x * blockDim.x + threadIdx.x) / kKnnLanes;
    const bool live = j_raw < N;
    const int j = live ? j_raw : N - 1;          // surplus groups shadow the last point (the warp shuffles together)
    const int b = blockIdx.y;
    const KnnHeader* h = h_all + b;
    const unsigned int* cell_start = cell_start_all + size_t(b) * (kMaxCells + 1);
    const float4* sorted = sorted_all + size_t(b) * N;
    const float4 me = sorted[j];
    const float px = me.x, py = me.y, pz = me.z;
    const int3 c = cell_coord(h, px, py, pz);
    const int dx = h->dim[0], dy = h->dim[1], dz = h->dim[2];
    const float cell = h->cell;
    const float inf = __int_as_float(0x7f800000);
    float b0 = inf, b1 = inf, b2 = inf;          // this lane's three best (fewer than 3 other points -> +inf, like the oracle)
    float third = inf;                           // the point's third-best distance after the last merge (group-uniform)
    bool done = false;
    const int max_ring = max(dx, max(dy, dz));
    for (int ring = 0; ring <= max_ring; ++ring) {
        if (ring >= 2) {
            // every unvisited point lies at least (ring - 1) * cell + (distance to the nearest face of the own cell)
            // away; use the weaker bound (ring - 1) * cell, shrunk slightly for rounding of the cell assignment
            const float bound = float(ring - 1) * cell * 0.999f;
            if (bound * bound > third) done = true;
        }
        if (__all_sync(0xffffffffu, done)) break;
        if (!done) {
            const int z0 = c.z - ring, z1 = c.z + ring, y0 = c.y - ring, y1 = c.y + ring, x0 = c.x - ring, x1 = c.x + ring;
            for (int z = max(z0, 0); z <= min(z1, dz - 1); ++z)
                for (int y = max(y0, 0); y <= min(y1, dy - 1); ++y) {
                    const bool shell_row = (z == z0 || z == z1 || y == y0 || y == y1);
                    const unsigned int id = (unsigned(z) * dy + y) * dx;
                    // shell rows: the row's cells are consecutive in the table, one [start, end) range for the whole
                    // row; interior rows: only the two end cells
                    const int nseg = shell_row ? 1 : 2;
                    for (int sg = 0; sg < nseg; ++sg) {
                        int xa, xb;
                        if (shell_row) { xa = max(x0, 0); xb = min(x1, dx - 1); }
                        else { xa = xb = sg == 0 ? x0 : x1; if (xa < 0 || xa >= dx) continue; }
                        const unsigned int s = cell_start[id + xa], e = cell_start[id + xb + 1];
                        for (unsigned int k = s + sub; k < e; k += kKnnLanes) {
                            const float4 q = __ldg(sorted + k);
                            const float ddx = q.x - px, ddy = q.y - py, ddz = q.z - pz;
                            const float d = ddx * ddx + ddy * ddy + ddz * ddz;
                            insert3(k != unsigned(j) ? d : inf, b0, b1, b2);
                        }
                    }
                }
        }
        if (ring >= 1) {                          // (the ring-1 shell is always visited: nothing to decide after ring 0)
            if (kKnnLanes > 1) merge3_group(b0, b1, b2, lane);
            third = b2;
            if (kKnnLanes > 1 && sub != 0) b0 = b1 = b2 = inf;     // the merged list lives in the group's first lane
        }
    }
    if (kKnnLanes > 1 && max_ring == 0) merge3_group(b0, b1, b2, lane);
    if (live && sub == 0) out_all[size_t(b) * N + __float_as_int(me.w)] = (b0 + b1 + b2) / 3.0f;
}

}  // namespace
}  // namespace sgr

using namespace sgr;

extern "C" {

uint64_t sgr_knn_scratch_bytes_batched(int32_t num_subjects, int32_t num_points) {
    if (num_points < 0 || num_subjects <= 0) return 0;
    return knn_layout(num_subjects, num_points).total;
}
uint64_t sgr_knn_scratch_bytes(int32_t num_points) { return sgr_knn_scratch_bytes_batched(1, num_points); }

// defined in sgr_api.cu
int sgr_set_error_(int code, const char* msg);
void sgr_count_launches_(unsigned int n);

int sgr_knn_mean_dist2_batched(const float* points, int32_t B, int32_t N, float* out, void* scratch,
                               uint64_t scratch_bytes, void* stream) {
    if (B <= 0 || N < 0 || (N > 0 && (!points || !out))) return sgr_set_error_(SGR_E_INVALID_ARGUMENT, "bad knn arguments");
    if (N == 0) return SGR_OK;
    const KnnLayout L = knn_layout(B, N);
    if (!scratch || scratch_bytes < L.total) return sgr_set_error_(SGR_E_BUFFER_TOO_SMALL, "knn scratch buffer too small");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    char* x = static_cast<char*>(scratch);
    KnnHeader* h = reinterpret_cast<KnnHeader*>(x + L.header);
    unsigned int* cell_of = reinterpret_cast<unsigned int*>(x + L.cell_of);
    unsigned int* cell_start = reinterpret_cast<unsigned int*>(x + L.cell_start);
    unsigned int* cell_fill = reinterpret_cast<unsigned int*>(x + L.cell_fill);
    unsigned int* block_sum = reinterpret_cast<unsigned int*>(x + L.block_sum);
    float4* sorted = reinterpret_cast<float4*>(x + L.sorted);
    cudaError_t e;
    // cell_start and cell_fill are adjacent: one memset
    if ((e = cudaMemsetAsync(cell_start, 0, L.block_sum - L.cell_start, s)) != cudaSuccess)
        return sgr_set_error_(SGR_E_CUDA, cudaGetErrorString(e));
    const int gb = (N + 255) / 256;
    knn_init_kernel<<<(B + 31) / 32, 32, 0, s>>>(h, B);
    knn_bbox_kernel<<<dim3(min(gb, 148), B), 256, 0, s>>>(points, N, h);
    knn_grid_kernel<<<(B + 31) / 32, 32, 0, s>>>(h, N, B);
    knn_count_kernel<<<dim3(gb, B), 256, 0, s>>>(points, N, h, cell_of, cell_start);
    knn_scan_blocks_kernel<<<dim3(kScanBlocks, B), kScanThreads, 0, s>>>(h, cell_start, block_sum);
    knn_scan_tops_kernel<<<B, kScanBlocks, 0, s>>>(h, block_sum, cell_start, N);
    knn_scan_add_kernel<<<dim3(296, B), 256, 0, s>>>(h, block_sum, cell_start);
    knn_fill_kernel<<<dim3(gb, B), 256, 0, s>>>(points, N, cell_of, cell_start, cell_fill, sorted);
    // one lane per point when the launch fills the machine anyway, eight when a single small set leaves it idle
    if ((long long)B * N > 262144)
        knn_query_kernel<1><<<dim3((N + 127) / 128, B), 128, 0, s>>>(N, h, cell_start, sorted, out);
    else
        knn_query_kernel<8><<<dim3((N * 8 + 127) / 128, B), 128, 0, s>>>(N, h, cell_start, sorted, out);
    if ((e = cudaGetLastError()) != cudaSuccess) return sgr_set_error_(SGR_E_CUDA, cudaGetErrorString(e));
    sgr_count_launches_(9);
    return SGR_OK;
}

int sgr_knn_mean_dist2(const float* points, int32_t N, float* out, void* scratch, uint64_t scratch_bytes,
                       void* stream) {
    return sgr_knn_mean_dist2_batched(points, 1, N, out, scratch, scratch_bytes, stream);
}

}  // extern "C"
