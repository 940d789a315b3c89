// Device-side data path upstream of the voxelizer and selection primitives of the other acquisition strategies
// (SURVEY.md 8f rows 1 and 3).
//
// crb3d_mask_collate_points replaces, for a whole batch at once and on the device,
//   pcdet/datasets/processor/data_processor.py:78-90   mask_points_and_boxes_outside_range (points part):
//       common_utils.mask_points_by_range (common_utils.py:60-63): keep x0 <= x <= x1 and y0 <= y <= y1 (z is NOT tested)
//   pcdet/datasets/dataset.py:173-178                   collate_batch: pad a batch-index column in front, concatenate
// so that raw clouds go host -> device once and the DataLoader-side numpy passes disappear. The compaction is stable (the
// kept points stay in input order, as numpy boolean indexing leaves them). shuffle_points (data_processor.py:92-103) is a
// gather with a caller-supplied permutation (crb3d_gather_rows_f32); it is off in test mode (kitti_dataset.yaml:58-62).
//
// crb3d_furthest_first replaces the greedy k-centre loop of pcdet/query_strategies/coreset_sampling.py:31-52 (an argmax,
// a (m x 1) distance GEMM and a PYTHON loop over all m elements per pick) by two launches per pick and no host round trip:
// the distance to the new centre, the running minimum and the next argmax are one pass over the embeddings.
#include "common.cuh"

namespace {

__global__ void __launch_bounds__(256) range_flags_kernel(const float* __restrict__ pts, int64_t n, int stride, int xcol,
                                                          float x0, float y0, float x1, float y1, int* __restrict__ flags) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    const float x = pts[p * stride + xcol], y = pts[p * stride + xcol + 1];
    flags[p] = (x >= x0 && x <= x1 && y >= y0 && y <= y1) ? 1 : 0;   // NaN fails every comparison, as in numpy
}

__global__ void __launch_bounds__(256) collate_scatter_kernel(const float* __restrict__ pts, int64_t n, int stride,
                                                              const int* __restrict__ flags, const int* __restrict__ rank,
                                                              const int* __restrict__ frame_off, int B,
                                                              float* __restrict__ out) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n || !flags[p]) return;
    int lo = 0, hi = B;                                  // frame of point p: frame_off[lo] <= p < frame_off[hi]
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (p >= frame_off[mid]) lo = mid; else hi = mid;
    }
    float* dst = out + (size_t)rank[p] * (stride + 1);
    dst[0] = (float)lo;
    for (int c = 0; c < stride; ++c) dst[1 + c] = pts[p * stride + c];
}

__global__ void collate_offsets_kernel(const int* __restrict__ rank, const int* __restrict__ total,
                                       const int* __restrict__ frame_off, int B, int64_t n, int* __restrict__ out_off) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b > B) return;
    const int s = frame_off[b];
    out_off[b] = (b < B && s < n) ? rank[s] : *total;
}

// ---- furthest-first
constexpr int FF_THREADS = 256;

// one warp per row: d = max(0, |x|^2 + |c|^2 - 2 x.c) (NaN -> 0), min_dist[row] = min(min_dist[row], d); block-level argmax of
// the updated min_dist (ties: lowest row) -> partial[blockIdx]
__global__ void __launch_bounds__(FF_THREADS) ff_update_kernel(const float* __restrict__ X, int m, int d,
                                                               const float* __restrict__ norms, const int* __restrict__ centre,
                                                               int do_update, float* __restrict__ min_dist,
                                                               unsigned long long* __restrict__ partial) {
    __shared__ unsigned long long best_s[FF_THREADS / 32];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int row = blockIdx.x * (FF_THREADS / 32) + warp;
    unsigned long long key = 0ull;
    if (row < m) {
        float md = min_dist[row];
        if (do_update) {
            const int c = *centre;
            const float* xr = X + (size_t)row * d;
            const float* xc = X + (size_t)c * d;
            float dot = 0.0f;
            for (int k = lane; k < d; k += 32) dot = fmaf(xr[k], __ldg(&xc[k]), dot);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
            float dist = norms[row] + norms[c] - 2.0f * dot;
            if (dist != dist) dist = 0.0f;
            dist = fmaxf(dist, 0.0f);
            md = fminf(md, dist);
            if (lane == 0) min_dist[row] = md;
        }
        // order-preserving key of a float (negative values included), lowest row wins ties
        unsigned int u = __float_as_uint(md);
        u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
        key = ((unsigned long long)u << 32) | (unsigned long long)(0xFFFFFFFFu - (unsigned int)row);
    }
    if (lane == 0) best_s[warp] = key;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long b = 0ull;
        for (int w = 0; w < FF_THREADS / 32; ++w) b = best_s[w] > b ? best_s[w] : b;
        partial[blockIdx.x] = b;
    }
}

__global__ void __launch_bounds__(256) ff_pick_kernel(const unsigned long long* __restrict__ partial, int n_partial,
                                                      int* __restrict__ centre, long long* __restrict__ out, int pick) {
    __shared__ unsigned long long red[256];
    unsigned long long b = 0ull;
    for (int t = threadIdx.x; t < n_partial; t += 256) b = partial[t] > b ? partial[t] : b;
    red[threadIdx.x] = b;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if (threadIdx.x < s) red[threadIdx.x] = red[threadIdx.x + s] > red[threadIdx.x] ? red[threadIdx.x + s] : red[threadIdx.x];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        const int idx = (int)(0xFFFFFFFFu - (unsigned int)(red[0] & 0xFFFFFFFFull));
        *centre = idx;
        out[pick] = idx;
    }
}

__global__ void __launch_bounds__(256) row_norms_kernel(const float* __restrict__ X, int m, int d, float* __restrict__ norms) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int row = blockIdx.x * 8 + warp;
    if (row >= m) return;
    float s = 0.0f;
    for (int k = lane; k < d; k += 32) { const float v = X[(size_t)row * d + k]; s = fmaf(v, v, s); }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) norms[row] = s;
}

}  // namespace

extern "C" int crb3d_mask_collate_points_workspace_bytes(int64_t n, size_t* bytes) {
    if (!bytes || n < 0) return CRB3D_ERR_ARG;
    *bytes = crb3d_align(sizeof(int) * (size_t)(n > 0 ? n : 1)) * 2 + crb3d_align(sizeof(int) * crb3d_scan_ws_ints(n)) + 256;
    return CRB3D_OK;
}

// points (n, stride) f32 with x at column xcol, y at xcol+1; frame_offsets (B+1) int32; range4 (HOST) = x0, y0, x1, y1.
// out (n, 1+stride): the kept points in input order with the batch index in column 0 (only the first out_offsets[B] rows are
// written); out_offsets (B+1) int32: per-frame row offsets of the kept points (device-side counts: no host read needed).
extern "C" int crb3d_mask_collate_points(const float* points, int64_t n, int stride, int xcol, const int* frame_offsets, int B,
                                         const float* range4, float* out, int* out_offsets, void* ws, size_t ws_bytes,
                                         cudaStream_t stream) {
    if (n < 0 || stride <= 0 || xcol < 0 || xcol + 1 >= stride || !frame_offsets || B <= 0 || !range4 || !out_offsets) return CRB3D_ERR_ARG;
    if (n >= 0x7fffffffLL) return CRB3D_ERR_UNSUPPORTED;
    if (n == 0) { CRB3D_CUDA(cudaMemsetAsync(out_offsets, 0, sizeof(int) * (B + 1), stream)); return CRB3D_OK; }
    if (!points || !out) return CRB3D_ERR_ARG;
    WsCursor c(ws, ws_bytes);
    int* flags = c.take<int>(n);
    int* rank = c.take<int>(n);
    int* scan_ws = c.take<int>(crb3d_scan_ws_ints(n));
    int* total = c.take<int>(1);
    if (!c.ok) return CRB3D_ERR_WORKSPACE;
    const unsigned nb = (unsigned)crb3d_divup(n, 256);
    range_flags_kernel<<<nb, 256, 0, stream>>>(points, n, stride, xcol, range4[0], range4[1], range4[2], range4[3], flags);
    int rc = crb3d_scan_exclusive_i32(flags, rank, n, scan_ws, total, stream);
    if (rc) return rc;
    collate_scatter_kernel<<<nb, 256, 0, stream>>>(points, n, stride, flags, rank, frame_offsets, B, out);
    collate_offsets_kernel<<<(unsigned)crb3d_divup(B + 1, 64), 64, 0, stream>>>(rank, total, frame_offsets, B, n, out_offsets);
    CRB3D_CHECK_LAUNCH();
    return CRB3D_OK;
}

extern "C" int crb3d_furthest_first_workspace_bytes(int m, size_t* bytes) {
    if (!bytes || m < 0) return CRB3D_ERR_ARG;
    *bytes = crb3d_align(sizeof(float) * (size_t)(m > 0 ? m : 1)) + crb3d_align(sizeof(unsigned long long) * (size_t)(crb3d_divup(m > 0 ? m : 1, 8))) + 512;
    return CRB3D_OK;
}

// coreset_sampling.py:31-52: X (m, d) f32 embeddings, min_dist (m) f32 = the start distances (the reference starts from the
// MEAN squared distance to the labelled set; the caller computes it), updated in place; out_idx (n_pick) int64 on the device.
// Per pick: arg-max of min_dist (ties: lowest row), then min_dist = min(min_dist, max(0, |x|^2 + |c|^2 - 2 x.c)).
extern "C" int crb3d_furthest_first(const float* X, int m, int d, float* min_dist, int n_pick, long long* out_idx, void* ws,
                                    size_t ws_bytes, cudaStream_t stream) {
    if (m < 0 || d <= 0 || n_pick < 0 || (n_pick > 0 && !out_idx)) return CRB3D_ERR_ARG;
    if (n_pick == 0) return CRB3D_OK;
    if (m == 0 || !X || !min_dist) return CRB3D_ERR_ARG;
    WsCursor c(ws, ws_bytes);
    float* norms = c.take<float>(m);
    const int n_blocks = (int)crb3d_divup(m, FF_THREADS / 32);
    unsigned long long* partial = c.take<unsigned long long>(n_blocks);
    int* centre = c.take<int>(1);
    if (!c.ok) return CRB3D_ERR_WORKSPACE;
    row_norms_kernel<<<(unsigned)crb3d_divup(m, 8), 256, 0, stream>>>(X, m, d, norms);
    for (int i = 0; i < n_pick; ++i) {
        ff_update_kernel<<<n_blocks, FF_THREADS, 0, stream>>>(X, m, d, norms, centre, i > 0 ? 1 : 0, min_dist, partial);
        ff_pick_kernel<<<1, 256, 0, stream>>>(partial, n_blocks, centre, out_idx, i);
    }
    CRB3D_CHECK_LAUNCH();
    return CRB3D_OK;
}
