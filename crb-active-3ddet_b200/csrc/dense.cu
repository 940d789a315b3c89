// SparseConvTensor.dense(): scatter (N,C) voxel features into a zero-filled dense grid, and its gradient (gather).
//
// Replaces spconv `SparseConvTensor.dense()` as called from
//   pcdet/models/backbones_2d/map_to_bev/height_compression.py:21-24   (then .view(B, C*D, H, W))
// layout 0: NCDHW  out[b][c][z][y][x]                      (the reference contract)
// layout 1: BEV channels-last  out[b][y][x][c*D + z]       (same values as view(B,C*D,H,W) in NHWC memory order;
//           what the tcgen05 BEV conv consumes - every voxel writes one contiguous strip per z)
#include "common.cuh"

namespace {

template <bool GATHER>
__global__ void __launch_bounds__(256) dense_kernel(const int* __restrict__ coords, int n, int C, int D, int H, int W,
                                                    int layout, float* __restrict__ feat, float* __restrict__ dense,
                                                    const int* __restrict__ n_dev) {
    // warp = 32 consecutive rows of one channel for layout 0 (writes land on neighbouring x);
    // for layout 1 a warp covers 32 consecutive channels of one row.
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (int64_t)n * C) return;
    int row, c;
    if (layout == 0) { row = (int)(t % n); c = (int)(t / n); }
    else { c = (int)(t % C); row = (int)(t / C); }
    if (n_dev && row >= *n_dev) return;   // n is only the capacity of a static buffer
    const int4 q = __ldg(reinterpret_cast<const int4*>(coords) + row);
    size_t off;
    if (layout == 0) off = ((((size_t)q.x * C + c) * D + q.y) * H + q.z) * W + q.w;
    else off = (((size_t)q.x * H + q.z) * W + q.w) * ((size_t)C * D) + (size_t)c * D + q.y;
    if (GATHER) feat[(size_t)row * C + c] = dense[off];
    else dense[off] = feat[(size_t)row * C + c];
}

}  // namespace

extern "C" int crb3d_sparse_to_dense(const float* feat, const int* coords, int n, int C, int B, int D, int H, int W,
                                     int layout, int zero_fill, float* dense, const int* n_dev, cudaStream_t stream) {
    if (n < 0 || C <= 0 || B <= 0 || D <= 0 || H <= 0 || W <= 0 || !dense || (layout != 0 && layout != 1)) return CRB3D_ERR_ARG;
    if (zero_fill) CRB3D_CUDA(cudaMemsetAsync(dense, 0, sizeof(float) * (size_t)B * C * D * H * W, stream));
    if (n == 0) return CRB3D_OK;
    dense_kernel<false><<<(unsigned)crb3d_divup((int64_t)n * C, 256), 256, 0, stream>>>(coords, n, C, D, H, W, layout,
                                                                                       const_cast<float*>(feat), dense, n_dev);
    CRB3D_CHECK_LAUNCH();
    return CRB3D_OK;
}

extern "C" int crb3d_dense_to_sparse(const float* dense, const int* coords, int n, int C, int B, int D, int H, int W,
                                     int layout, float* feat, cudaStream_t stream) {
    if (n < 0 || C <= 0 || B <= 0 || !feat || (layout != 0 && layout != 1)) return CRB3D_ERR_ARG;
    if (n == 0) return CRB3D_OK;
    dense_kernel<true><<<(unsigned)crb3d_divup((int64_t)n * C, 256), 256, 0, stream>>>(coords, n, C, D, H, W, layout, feat,
                                                                                      const_cast<float*>(dense), nullptr);
    CRB3D_CHECK_LAUNCH();
    return CRB3D_OK;
}
