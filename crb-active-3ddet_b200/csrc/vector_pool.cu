// Vector-pool aggregation ops of PV-RCNN++ (the last entry points of pointnet2_stack_cuda):
//   pcdet/ops/pointnet2/pointnet2_stack/src/vector_pool_gpu.cu:19-91    query_three_nn_by_stacked_local_idxs_kernel
//   .../vector_pool_gpu.cu:124-205                                      query_stacked_local_neighbor_idxs_kernel
//   .../vector_pool_gpu.cu:240-371                                      vector_pool_kernel_stack
//   .../vector_pool_gpu.cu:417-446                                      vector_pool_grad_kernel_stack
// The reference runs ONE THREAD per query point that walks all support points of its frame serially (O(M*N) per thread, 4 KB of
// thread-local index scratch) and places each point's results with a global atomicAdd, so the layout of its outputs depends on
// thread scheduling. Here a WARP owns a query point: 32 support points are tested per step, the hits are consumed in ascending
// index order (the order the reference's loop sees them - sums are accumulated one term at a time in that order, so features
// are bit-identical), and the stacked neighbour lists are laid out in query order by a count -> scan -> fill pass (one of the
// layouts the reference can produce; `start_len` tells the consumer where each list is, exactly as there).
#include "common.cuh"
#include <cmath>

namespace {

constexpr int VP_WARPS = 8;
constexpr int VP_MAX_LOCAL = 1000;      // the reference's temp_idxs[1000]

struct Frame { int b, start, n; };

// frame of query `pt` and the support range of that frame (the reference's two linear scans over the batch counts)
__device__ __forceinline__ Frame find_frame(int pt, const int* __restrict__ new_cnt, const int* __restrict__ xyz_cnt, int B) {
    Frame f;
    f.b = 0;
    int pt_cnt = new_cnt[0];
    for (int k = 1; k < B; ++k) {
        if (pt < pt_cnt) break;
        pt_cnt += new_cnt[k];
        f.b = k;
    }
    f.start = 0;
    for (int k = 0; k < f.b; ++k) f.start += xyz_cnt[k];
    f.n = xyz_cnt[f.b];
    return f;
}

// neighbourhood test of the reference (ball: squared distance > r^2 rejects; cube: any |local| > r rejects), same contraction
// as nvcc gives the reference's expression (common.cuh: sqdist3)
__device__ __forceinline__ bool in_neighbourhood(float lx, float ly, float lz, float r, float r2, int neighbor_type) {
    if (neighbor_type == 1) return !(__fmaf_rn(lz, lz, __fmaf_rn(lx, lx, __fmul_rn(ly, ly))) > r2);
    return !((fabsf(lx) > r) | (fabsf(ly) > r) | (fabsf(lz) > r));
}

// MODE 0: count the neighbours of every query (capped like the reference: 1000 local slots, nsample); MODE 1: write them
template <int MODE>
__global__ void __launch_bounds__(VP_WARPS * 32) local_neighbors_kernel(const float* __restrict__ xyz, const int* __restrict__ xyz_cnt,
                                                                        const float* __restrict__ new_xyz, const int* __restrict__ new_cnt,
                                                                        int B, int M, float r, int nsample, int neighbor_type,
                                                                        int* __restrict__ start_len, int* __restrict__ stack, int max_thresh) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int pt = blockIdx.x * VP_WARPS + warp;
    if (pt >= M) return;
    const Frame f = find_frame(pt, new_cnt, xyz_cnt, B);
    const float nx = new_xyz[(size_t)pt * 3], ny = new_xyz[(size_t)pt * 3 + 1], nz = new_xyz[(size_t)pt * 3 + 2];
    const float r2 = r * r;
    const float* src = xyz + (size_t)f.start * 3;
    int offset = 0, room = 0;
    if (MODE == 1) {
        offset = start_len[pt * 2];
        if (offset >= max_thresh) return;
        room = min(start_len[pt * 2 + 1], max_thresh - offset);       // start + cnt >= max_thresh: truncated to max_thresh - start
    }
    int cnt = 0;
    const int limit = nsample > 0 ? min(nsample, VP_MAX_LOCAL) : VP_MAX_LOCAL;
    for (int k0 = 0; k0 < f.n && cnt < limit; k0 += 32) {
        const int k = k0 + lane;
        bool hit = false;
        if (k < f.n) hit = in_neighbourhood(src[(size_t)k * 3] - nx, src[(size_t)k * 3 + 1] - ny, src[(size_t)k * 3 + 2] - nz, r, r2, neighbor_type);
        const unsigned int m = __ballot_sync(0xffffffffu, hit);
        const int pos = cnt + __popc(m & ((1u << lane) - 1u));
        if (MODE == 1 && hit && pos < room) stack[offset + pos] = k + f.start;
        cnt += __popc(m);
    }
    if (MODE == 0 && lane == 0) start_len[pt * 2 + 1] = min(cnt, limit);
}

// offsets = exclusive scan of the counts + the caller's running total (the reference's atomicAdd on `cumsum`); the total is
// advanced by a launch of its own so that every block of this one reads the old value
__global__ void local_neighbors_offsets(const int* __restrict__ scanned, int M, int* __restrict__ start_len, const int* __restrict__ cumsum) {
    const int pt = blockIdx.x * blockDim.x + threadIdx.x;
    if (pt < M) start_len[pt * 2] = *cumsum + scanned[pt];
}
__global__ void local_neighbors_total(const int* __restrict__ total, int* __restrict__ cumsum) { *cumsum += *total; }
__global__ void gather_counts(const int* __restrict__ start_len, int M, int* __restrict__ counts) {
    const int pt = blockIdx.x * blockDim.x + threadIdx.x;
    if (pt < M) counts[pt] = start_len[pt * 2 + 1];
}

__global__ void __launch_bounds__(256) three_nn_local_kernel(const float* __restrict__ xyz, const float* __restrict__ centers,
                                                             int* __restrict__ idxs, float* __restrict__ dist2, const int* __restrict__ stack,
                                                             const int* __restrict__ start_len, int M, int G) {
    const int grid_idx = blockIdx.y;
    const int pt = blockIdx.x * blockDim.x + threadIdx.x;
    if (pt >= M) return;
    const size_t o = ((size_t)pt * G + grid_idx) * 3;
    const float cx = centers[o], cy = centers[o + 1], cz = centers[o + 2];
    const int* list = stack + start_len[pt * 2];
    const int len = start_len[pt * 2 + 1];
    float b1 = INFINITY, b2 = INFINITY, b3 = INFINITY;     // the reference's 1e40 double sentinels are +inf once stored as float
    int i1 = -1, i2 = -1, i3 = -1;
    for (int k = 0; k < len; ++k) {
        const int j = __ldg(&list[k]);
        const float d = sqdist3(cx, cy, cz, __ldg(&xyz[(size_t)j * 3]), __ldg(&xyz[(size_t)j * 3 + 1]), __ldg(&xyz[(size_t)j * 3 + 2]));
        if (d < b1) { b3 = b2; i3 = i2; b2 = b1; i2 = i1; b1 = d; i1 = j; }
        else if (d < b2) { b3 = b2; i3 = i2; b2 = d; i2 = j; }
        else if (d < b3) { b3 = d; i3 = j; }
    }
    if (i2 == -1) { i2 = i1; b2 = b1; }
    if (i3 == -1) { i3 = i1; b3 = b1; }
    dist2[o] = b1; dist2[o + 1] = b2; dist2[o + 2] = b3;
    idxs[o] = i1; idxs[o + 1] = i2; idxs[o + 2] = i3;
}

__global__ void __launch_bounds__(VP_WARPS * 32) vector_pool_kernel(const float* __restrict__ xyz, const float* __restrict__ feat,
                                                                    const int* __restrict__ xyz_cnt, const float* __restrict__ new_xyz,
                                                                    float* __restrict__ new_feat, float* __restrict__ new_local_xyz,
                                                                    const int* __restrict__ new_cnt, int gx, int gy, int gz, float r, int B, int M,
                                                                    int c_in, int c_out, int ce, int G, int* __restrict__ point_cnt,
                                                                    int* __restrict__ grouped, int use_xyz, float sx, float sy, float sz,
                                                                    int* __restrict__ cum_sum, int max_sum, int nsample, int neighbor_type,
                                                                    int pooling_type) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int pt = blockIdx.x * VP_WARPS + warp;
    if (pt >= M) return;
    const Frame f = find_frame(pt, new_cnt, xyz_cnt, B);
    const float nx = new_xyz[(size_t)pt * 3], ny = new_xyz[(size_t)pt * 3 + 1], nz = new_xyz[(size_t)pt * 3 + 2];
    const float r2 = r * r;
    const float* src = xyz + (size_t)f.start * 3;
    const float* fsrc = feat + (size_t)f.start * c_in;
    float* out = new_feat + (size_t)pt * c_out;
    float* oxyz = new_local_xyz + (size_t)pt * 3 * G;
    int* pcnt = point_cnt + (size_t)pt * G;
    const int reps = c_in / ce;          // input channels i and i + ce land on the same output channel (i % ce)
    int sample_cnt = 0;
    bool done = false;
    for (int k0 = 0; k0 < f.n && !done; k0 += 32) {
        const int k = k0 + lane;
        float lx = 0.f, ly = 0.f, lz = 0.f;
        bool hit = false;
        if (k < f.n) {
            lx = src[(size_t)k * 3] - nx; ly = src[(size_t)k * 3 + 1] - ny; lz = src[(size_t)k * 3 + 2] - nz;
            hit = in_neighbourhood(lx, ly, lz, r, r2, neighbor_type);
        }
        int g = 0;
        if (hit) {
            const int ix = (int)floorf((lx + r) / sx), iy = (int)floorf((ly + r) / sy), iz = (int)floorf((lz + r) / sz);
            g = min(max(ix * gy * gz + iy * gz + iz, 0), G - 1);
        }
        unsigned int m = __ballot_sync(0xffffffffu, hit);
        while (m && !done) {             // the hits of this step in ascending index order, the whole warp on one hit at a time
            const int h = __ffs(m) - 1;
            m &= m - 1;
            const int hk = k0 + h, hg = __shfl_sync(0xffffffffu, g, h);
            const float hx = __shfl_sync(0xffffffffu, lx, h), hy = __shfl_sync(0xffffffffu, ly, h), hz = __shfl_sync(0xffffffffu, lz, h);
            int first = 1;
            if (pooling_type == 1) {     // "random choice" = the first point that falls into the sub-voxel
                first = reinterpret_cast<volatile int*>(pcnt)[hg] == 0;   // written by lane 0 below, read by all lanes after the __syncwarp
                if (!first) continue;
            }
            const float* fk = fsrc + (size_t)hk * c_in;
            for (int j = lane; j < ce; j += 32) {
                float acc = pooling_type == 0 ? out[hg * ce + j] : 0.0f;
                if (pooling_type == 0) for (int q = 0; q < reps; ++q) acc += fk[j + q * ce];      // one term at a time, reference order
                else acc = fk[j + (reps - 1) * ce];                                             // plain assignment: the last i wins
                out[hg * ce + j] = acc;
            }
            if (use_xyz && lane < 3) {
                const float v = lane == 0 ? hx : (lane == 1 ? hy : hz);
                oxyz[hg * 3 + lane] = pooling_type == 0 ? oxyz[hg * 3 + lane] + v : v;
            }
            int stop = 0;
            if (lane == 0) {
                pcnt[hg] += 1;
                const int cnt = atomicAdd(cum_sum, 1);
                if (cnt < max_sum) {       // beyond the capacity the reference only keeps counting (and skips its sample counter)
                    grouped[(size_t)cnt * 3] = f.start + hk;
                    grouped[(size_t)cnt * 3 + 1] = pt;
                    grouped[(size_t)cnt * 3 + 2] = hg;
                    ++sample_cnt;
                    if (nsample > 0 && sample_cnt >= nsample) stop = 1;
                    if (pooling_type == 1 && sample_cnt >= G) stop = 1;
                }
            }
            __syncwarp();
            done = __shfl_sync(0xffffffffu, stop, 0) != 0;
        }
    }
}

__global__ void __launch_bounds__(256) vector_pool_grad_kernel(const float* __restrict__ grad_new, const int* __restrict__ point_cnt,
                                                               const int* __restrict__ grouped, float* __restrict__ grad_feat, int c_out,
                                                               int c_in, int ce, int G, int n_grouped) {
    const int ch = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_grouped || ch >= c_in) return;
    const int src = grouped[(size_t)i * 3], pt = grouped[(size_t)i * 3 + 1], g = grouped[(size_t)i * 3 + 2];
    const int total = point_cnt[(size_t)pt * G + g];
    const float w = 1 / fmaxf((float)total, 1.0f);
    atomicAdd(&grad_feat[(size_t)src * c_in + ch], grad_new[(size_t)pt * c_out + g * ce + ch % ce] * w);
}

}  // namespace

extern "C" int crb3d_query_stacked_local_neighbor_idxs_workspace_bytes(int M, size_t* bytes) {
    if (!bytes || M < 0) return CRB3D_ERR_ARG;
    const size_t m = (size_t)(M > 0 ? M : 1);
    *bytes = crb3d_align(sizeof(int) * m) * 2 + crb3d_align(sizeof(int) * crb3d_scan_ws_ints(M)) + 256;
    return CRB3D_OK;
}

// query_stacked_local_neighbor_idxs_wrapper_stack (vector_pool.cpp:35-75): per query the support points of its frame within
// max_neighbour_distance (ball: neighbor_type 1, else cube), at most nsample (> 0) and at most 1000, as one stacked int32 list:
// start_len (M,2) = [offset, length]; cumsum (1) in/out running total; lists beyond avg_length * M entries are truncated / dropped.
extern "C" int crb3d_query_stacked_local_neighbor_idxs(const float* support_xyz, const int* xyz_batch_cnt, const float* new_xyz,
                                                       const int* new_xyz_batch_cnt, int batch_size, int M, int* stack_neighbor_idxs,
                                                       int* start_len, int* cumsum, int avg_length_of_neighbor_idxs,
                                                       float max_neighbour_distance, int nsample, int neighbor_type, void* ws,
                                                       size_t ws_bytes, cudaStream_t stream) {
    if (M < 0 || batch_size <= 0 || avg_length_of_neighbor_idxs < 0) return CRB3D_ERR_ARG;
    if (M == 0) return CRB3D_OK;
    if (!support_xyz || !xyz_batch_cnt || !new_xyz || !new_xyz_batch_cnt || !start_len || !cumsum) return CRB3D_ERR_ARG;
    WsCursor c(ws, ws_bytes);
    int* counts = c.take<int>(M);
    int* scanned = c.take<int>(M);
    int* scan_ws = c.take<int>(crb3d_scan_ws_ints(M));
    int* total = c.take<int>(1);
    if (!c.ok) return CRB3D_ERR_WORKSPACE;
    const long long max_thresh_ll = (long long)avg_length_of_neighbor_idxs * M;
    const int max_thresh = (int)(max_thresh_ll > 0x7fffffffLL ? 0x7fffffffLL : max_thresh_ll);
    const unsigned nb = (unsigned)crb3d_divup(M, VP_WARPS);
    local_neighbors_kernel<0><<<nb, VP_WARPS * 32, 0, stream>>>(support_xyz, xyz_batch_cnt, new_xyz, new_xyz_batch_cnt, batch_size, M,
                                                               max_neighbour_distance, nsample, neighbor_type, start_len, nullptr, 0);
    gather_counts<<<(unsigned)crb3d_divup(M, 256), 256, 0, stream>>>(start_len, M, counts);
    int rc = crb3d_scan_exclusive_i32(counts, scanned, M, scan_ws, total, stream);
    if (rc) return rc;
    local_neighbors_offsets<<<(unsigned)crb3d_divup(M, 256), 256, 0, stream>>>(scanned, M, start_len, cumsum);
    local_neighbors_total<<<1, 1, 0, stream>>>(total, cumsum);
    if (stack_neighbor_idxs && max_thresh > 0)
        local_neighbors_kernel<1><<<nb, VP_WARPS * 32, 0, stream>>>(support_xyz, xyz_batch_cnt, new_xyz, new_xyz_batch_cnt, batch_size, M,
                                                                   max_neighbour_distance, nsample, neighbor_type, start_len,
                                                                   stack_neighbor_idxs, max_thresh);
    CRB3D_CHECK_LAUNCH();
    return CRB3D_OK;
}

// query_three_nn_by_stacked_local_idxs_wrapper_stack (vector_pool.cpp:78-113): the three nearest of each query's local list to
// each of its num_total_grids grid centres; a list shorter than three repeats its nearest entry, an empty one gives -1.
extern "C" int crb3d_query_three_nn_by_stacked_local_idxs(const float* support_xyz, const float* new_xyz_grid_centers,
                                                          int* new_xyz_grid_idxs, float* new_xyz_grid_dist2,
                                                          const int* stack_neighbor_idxs, const int* start_len, int M,
                                                          int num_total_grids, cudaStream_t stream) {
    if (M < 0 || num_total_grids <= 0) return CRB3D_ERR_ARG;
    if (M == 0) return CRB3D_OK;
    if (!support_xyz || !new_xyz_grid_centers || !new_xyz_grid_idxs || !new_xyz_grid_dist2 || !stack_neighbor_idxs || !start_len)
        return CRB3D_ERR_ARG;
    three_nn_local_kernel<<<dim3((unsigned)crb3d_divup(M, 256), (unsigned)num_total_grids), 256, 0, stream>>>(
        support_xyz, new_xyz_grid_centers, new_xyz_grid_idxs, new_xyz_grid_dist2, stack_neighbor_idxs, start_len, M, num_total_grids);
    CRB3D_CHECK_LAUNCH();
    return CRB3D_OK;
}

// vector_pool_wrapper_stack (vector_pool.cpp:116-170). new_features (M, c_out), new_local_xyz (M, 3 * G), point_cnt_of_grid (M, G)
// and grouped_idxs (num_max_sum_points, 3) zero-filled by the caller, cum_sum (DEVICE int, zeroed here) receives the number of
// grouped entries the launch WANTED (the reference returns it to the host: the caller re-runs with more room when it exceeds
// num_max_sum_points). pooling_type 0: per sub-voxel sums (the caller divides by point_cnt), 1: first point of the sub-voxel.
extern "C" int crb3d_vector_pool_stack(const float* support_xyz, const float* support_features, const int* xyz_batch_cnt,
                                       const float* new_xyz, const int* new_xyz_batch_cnt, int batch_size, int M, int num_c_in,
                                       int num_c_out, int num_grid_x, int num_grid_y, int num_grid_z, float max_neighbour_distance,
                                       int use_xyz, int num_max_sum_points, int nsample, int neighbor_type, int pooling_type,
                                       float* new_features, float* new_local_xyz, int* point_cnt_of_grid, int* grouped_idxs,
                                       int* cum_sum, cudaStream_t stream) {
    const int G = num_grid_x * num_grid_y * num_grid_z;
    if (M < 0 || batch_size <= 0 || G <= 0 || num_c_in <= 0 || num_c_out <= 0 || num_c_out % G != 0 || !cum_sum) return CRB3D_ERR_ARG;
    const int ce = num_c_out / G;
    if (num_c_in % ce != 0 || (pooling_type != 0 && pooling_type != 1)) return CRB3D_ERR_ARG;
    CRB3D_CUDA(cudaMemsetAsync(cum_sum, 0, sizeof(int), stream));
    if (M == 0) return CRB3D_OK;
    if (!support_xyz || !support_features || !xyz_batch_cnt || !new_xyz || !new_xyz_batch_cnt || !new_features || !new_local_xyz ||
        !point_cnt_of_grid || (!grouped_idxs && num_max_sum_points > 0))
        return CRB3D_ERR_ARG;
    const float sx = max_neighbour_distance * 2 / num_grid_x, sy = max_neighbour_distance * 2 / num_grid_y,
                sz = max_neighbour_distance * 2 / num_grid_z;
    vector_pool_kernel<<<(unsigned)crb3d_divup(M, VP_WARPS), VP_WARPS * 32, 0, stream>>>(
        support_xyz, support_features, xyz_batch_cnt, new_xyz, new_features, new_local_xyz, new_xyz_batch_cnt, num_grid_x, num_grid_y,
        num_grid_z, max_neighbour_distance, batch_size, M, num_c_in, num_c_out, ce, G, point_cnt_of_grid, grouped_idxs, use_xyz, sx, sy, sz,
        cum_sum, num_max_sum_points, nsample, neighbor_type, pooling_type);
    CRB3D_CHECK_LAUNCH();
    return CRB3D_OK;
}

// vector_pool_grad_wrapper_stack (vector_pool.cpp:173-204): grad_support_features (N, c_in), zero-filled by the caller, += the
// gradient of every grouped (support point, query, sub-voxel) entry divided by the sub-voxel's point count.
extern "C" int crb3d_vector_pool_grad_stack(const float* grad_new_features, const int* point_cnt_of_grid, const int* grouped_idxs,
                                            float* grad_support_features, int num_c_out, int num_c_in, int num_total_grids,
                                            int num_grouped, cudaStream_t stream) {
    if (num_grouped < 0 || num_c_in <= 0 || num_c_out <= 0 || num_total_grids <= 0 || num_c_out % num_total_grids != 0) return CRB3D_ERR_ARG;
    if (num_grouped == 0) return CRB3D_OK;
    if (!grad_new_features || !point_cnt_of_grid || !grouped_idxs || !grad_support_features) return CRB3D_ERR_ARG;
    vector_pool_grad_kernel<<<dim3((unsigned)crb3d_divup(num_grouped, 256), (unsigned)num_c_in), 256, 0, stream>>>(
        grad_new_features, point_cnt_of_grid, grouped_idxs, grad_support_features, num_c_out, num_c_in, num_c_out / num_total_grids,
        num_total_grids, num_grouped);
    CRB3D_CHECK_LAUNCH();
    return CRB3D_OK;
}
