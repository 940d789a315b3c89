// Device-wide exclusive int32 scan used by the voxelizer (leader ranks) and the rulebook builder
// (bitmap popcount ranks). Three launches: tile sums -> one-block scan of sums -> tile downsweep.
#include "common.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kItems = CRB3D_SCAN_TILE / kThreads;  // 8 per thread

__global__ void __launch_bounds__(kThreads) scan_tile_sums(const int* __restrict__ in, int64_t n, int* __restrict__ sums) {
    __shared__ int sm[33];
    const int64_t base = (int64_t)blockIdx.x * CRB3D_SCAN_TILE + (int64_t)threadIdx.x * kItems;
    int s = 0;
#pragma unroll
    for (int i = 0; i < kItems; ++i) {
        int64_t p = base + i;
        if (p < n) s += in[p];
    }
    int tot;
    block_excl_scan(s, sm, &tot);
    if (threadIdx.x == 0) sums[blockIdx.x] = tot;
}

__global__ void __launch_bounds__(1024) scan_sums_one_block(int* __restrict__ sums, int64_t nb, int* __restrict__ total) {
    __shared__ int sm[33];
    __shared__ int carry_s;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (int64_t start = 0; start < nb; start += 1024) {
        int64_t p = start + threadIdx.x;
        int v = (p < nb) ? sums[p] : 0;
        int tot;
        int ex = block_excl_scan(v, sm, &tot);
        int carry = carry_s;
        if (p < nb) sums[p] = ex + carry;
        __syncthreads();
        if (threadIdx.x == 0) carry_s = carry + tot;
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        sums[nb] = carry_s;
        if (total) *total = carry_s;
    }
}

__global__ void __launch_bounds__(kThreads) scan_downsweep(const int* __restrict__ in, int* __restrict__ out, int64_t n,
                                                            const int* __restrict__ sums) {
    __shared__ int sm[33];
    const int64_t base = (int64_t)blockIdx.x * CRB3D_SCAN_TILE + (int64_t)threadIdx.x * kItems;
    int v[kItems];
    int s = 0;
#pragma unroll
    for (int i = 0; i < kItems; ++i) {
        int64_t p = base + i;
        v[i] = (p < n) ? in[p] : 0;
        s += v[i];
    }
    int tot;
    int ex = block_excl_scan(s, sm, &tot) + sums[blockIdx.x];
#pragma unroll
    for (int i = 0; i < kItems; ++i) {
        int64_t p = base + i;
        if (p < n) out[p] = ex;
        ex += v[i];
    }
}

}  // namespace

size_t crb3d_scan_ws_ints(int64_t n) { return (size_t)crb3d_divup(n > 0 ? n : 1, CRB3D_SCAN_TILE) + 1; }

int crb3d_scan_exclusive_i32(const int* in, int* out, int64_t n, int* block_sums, int* total, cudaStream_t stream) {
    if (n <= 0) {
        if (total) CRB3D_CUDA(cudaMemsetAsync(total, 0, sizeof(int), stream));
        return CRB3D_OK;
    }
    const int64_t nb = crb3d_divup(n, CRB3D_SCAN_TILE);
    scan_tile_sums<<<(unsigned)nb, kThreads, 0, stream>>>(in, n, block_sums);
    scan_sums_one_block<<<1, 1024, 0, stream>>>(block_sums, nb, total);
    scan_downsweep<<<(unsigned)nb, kThreads, 0, stream>>>(in, out, n, block_sums);
    CRB3D_CHECK_LAUNCH();
    return CRB3D_OK;
}

// exclusive scan of per-tile sums in place (sums[nb] and *total receive the grand total): the middle pass of a tiled scan whose
// outer passes live elsewhere (csrc/rulebook.cu: popcount scan of the output-cell bitmap)
int crb3d_scan_block_sums(int* sums, int64_t nb, int* total, cudaStream_t stream) {
    if (nb <= 0) {
        if (total) CRB3D_CUDA(cudaMemsetAsync(total, 0, sizeof(int), stream));
        return CRB3D_OK;
    }
    scan_sums_one_block<<<1, 1024, 0, stream>>>(sums, nb, total);
    CRB3D_CHECK_LAUNCH();
    return CRB3D_OK;
}

namespace {
template <typename T>
__global__ void __launch_bounds__(256) fill_kernel(T* __restrict__ p, size_t n, T v) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) p[i] = v;
}
}  // namespace

int crb3d_fill_i32(int* p, size_t n, int v, cudaStream_t stream) {
    if (n == 0) return CRB3D_OK;
    unsigned nb = (unsigned)(crb3d_divup((int64_t)n, 256) < 148 * 16 ? crb3d_divup((int64_t)n, 256) : 148 * 16);
    fill_kernel<int><<<nb, 256, 0, stream>>>(p, n, v);
    CRB3D_CHECK_LAUNCH();
    return CRB3D_OK;
}
int crb3d_fill_f32(float* p, size_t n, float v, cudaStream_t stream) {
    if (n == 0) return CRB3D_OK;
    unsigned nb = (unsigned)(crb3d_divup((int64_t)n, 256) < 148 * 16 ? crb3d_divup((int64_t)n, 256) : 148 * 16);
    fill_kernel<float><<<nb, 256, 0, stream>>>(p, n, v);
    CRB3D_CHECK_LAUNCH();
    return CRB3D_OK;
}
