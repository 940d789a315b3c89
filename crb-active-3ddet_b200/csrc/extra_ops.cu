// Remaining op families of pcdet/ops, same results as the reference kernels:
//   pointnet2_stack/src/voxel_query_gpu.cu:13-98            voxel_query_kernel_stack      (Voxel-RCNN neighbour search)
//   pointnet2_batch/src/ball_query_gpu.cu:15-46             ball_query_kernel_fast        (B, N, 3) / (B, M, 3) layout
//   pointnet2_batch/src/group_points_gpu.cu:15-70           group_points[_grad]_kernel_fast   (B, C, N) features
//   pointnet2_batch/src/sampling_gpu.cu:15-70               gather_points[_grad]_kernel_fast
//   pointnet2_batch/src/interpolate_gpu.cu:17-140           three_nn / three_interpolate[_grad]_kernel_fast
//   roipoint_pool3d/src/roipoint_pool3d_kernel.cu:14-165    assign_pts_to_box3d + get_pooled_idx + roipool3d_forward
// (farthest_point_sampling of pointnet2_batch is the kernel of csrc/pointnet2.cu: same launcher in the reference.)
// Index outputs are bit-exact targets: distance expressions keep the contraction the reference compiles to (common.cuh:
// sqdist3), scans run in ascending index order, the box test keeps the reference's float / double mix.
// What changed: the batch ball query is one warp per query with a ballot-ordered append; roipoint pooling is one warp per
// box that walks the points once (the reference materialises a (B, N, M) assignment matrix in freshly cudaMalloc'ed memory
// and then scans it serially per box) and needs no scratch at all.
#include "common.cuh"
#include <cmath>

namespace {

// ------------------------------------------------------------------ voxel query (stack)
__global__ void __launch_bounds__(256) voxel_query_kernel(int M, int R1, int R2, int R3, int nsample, float radius, int z_range,
                                                          int y_range, int x_range, const float* __restrict__ new_xyz,
                                                          const float* __restrict__ xyz, const int* __restrict__ new_coords,
                                                          const int* __restrict__ point_indices, int* __restrict__ idx) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= M) return;
    const float nx = new_xyz[(size_t)q * 3], ny = new_xyz[(size_t)q * 3 + 1], nz = new_xyz[(size_t)q * 3 + 2];
    const int4 c = reinterpret_cast<const int4*>(new_coords)[q];   // b, z, y, x
    int* out = idx + (size_t)q * nsample;
    const float r2 = radius * radius;
    int cnt = 0;
    for (int dz = -z_range; dz <= z_range && cnt < nsample; ++dz) {
        const int z = c.y + dz;
        if (z < 0 || z >= R1) continue;
        for (int dy = -y_range; dy <= y_range && cnt < nsample; ++dy) {
            const int y = c.z + dy;
            if (y < 0 || y >= R2) continue;
            for (int dx = -x_range; dx <= x_range && cnt < nsample; ++dx) {
                const int x = c.w + dx;
                if (x < 0 || x >= R3) continue;
                const int nb = __ldg(&point_indices[(((size_t)c.x * R1 + z) * R2 + y) * R3 + x]);
                if (nb < 0) continue;
                // reference: (x_per - new_x)^2 + (y_per - new_y)^2 + (z_per - new_z)^2, kept when dist2 <= radius2
                const float d2 = sqdist3(xyz[(size_t)nb * 3], xyz[(size_t)nb * 3 + 1], xyz[(size_t)nb * 3 + 2], nx, ny, nz);
                if (d2 > r2) continue;
                if (cnt == 0)
                    for (int l = 0; l < nsample; ++l) out[l] = nb;
                out[cnt++] = nb;
            }
        }
    }
    if (cnt == 0) out[0] = -1;
}

// ------------------------------------------------------------------ pointnet2_batch: ball query
constexpr int BQ_WARPS = 8, BQ_TILE = 1024;
__global__ void __launch_bounds__(BQ_WARPS * 32) ball_query_batch_kernel(int n, int m, float radius, int nsample,
                                                                         const float* __restrict__ new_xyz,
                                                                         const float* __restrict__ xyz, int* __restrict__ idx) {
    __shared__ float sx[BQ_TILE], sy[BQ_TILE], sz[BQ_TILE];
    const int b = blockIdx.y, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q = blockIdx.x * BQ_WARPS + warp;
    const bool live = q < m;
    float qx = 0.f, qy = 0.f, qz = 0.f;
    if (live) {
        const float* p = new_xyz + ((size_t)b * m + q) * 3;
        qx = p[0]; qy = p[1]; qz = p[2];
    }
    const float r2 = radius * radius;
    int* out = idx + ((size_t)b * m + (live ? q : 0)) * nsample;
    const float* src = xyz + (size_t)b * n * 3;
    int cnt = 0;
    bool done = !live;
    for (int t0 = 0; t0 < n; t0 += BQ_TILE) {
        const int tn = min(BQ_TILE, n - t0);
        __syncthreads();
        for (int t = threadIdx.x; t < tn; t += blockDim.x) {
            sx[t] = src[(size_t)(t0 + t) * 3]; sy[t] = src[(size_t)(t0 + t) * 3 + 1]; sz[t] = src[(size_t)(t0 + t) * 3 + 2];
        }
        __syncthreads();
        if (!done) {
            for (int k0 = 0; k0 < tn; k0 += 32) {
                const int k = k0 + lane;
                const bool hit = k < tn && sqdist3(qx, qy, qz, sx[k], sy[k], sz[k]) < r2;
                const unsigned int msk = __ballot_sync(0xffffffffu, hit);
                if (msk) {
                    if (cnt == 0) {      // the first hit pads the whole row (ball_query_gpu.cu:36-40)
                        const int first = t0 + k0 + __ffs(msk) - 1;
                        for (int l = lane; l < nsample; l += 32) out[l] = first;
                        __syncwarp();
                    }
                    const int pos = cnt + __popc(msk & ((1u << lane) - 1u));
                    if (hit && pos < nsample) out[pos] = t0 + k;
                    cnt += __popc(msk);
                    if (cnt >= nsample) { done = true; break; }
                }
            }
        }
        if (__syncthreads_and(done)) break;
    }   // an empty ball leaves the caller's zero-filled row untouched (the batch kernel has no -1 marker)
}

// ------------------------------------------------------------------ pointnet2_batch: group / gather (+grad)
__global__ void __launch_bounds__(256) group_points_batch_kernel(int c, int n, int npoints, int nsample,
                                                                 const float* __restrict__ points, const int* __restrict__ idx,
                                                                 float* __restrict__ out) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;     // (point, sample) flattened
    const int ci = blockIdx.y, b = blockIdx.z;
    if (t >= npoints * nsample) return;
    const int src = __ldg(&idx[(size_t)b * npoints * nsample + t]);
    out[((size_t)b * c + ci) * npoints * nsample + t] = points[((size_t)b * c + ci) * n + src];
}
__global__ void __launch_bounds__(256) group_points_batch_grad_kernel(int c, int n, int npoints, int nsample,
                                                                      const float* __restrict__ grad_out,
                                                                      const int* __restrict__ idx, float* __restrict__ grad_points) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int ci = blockIdx.y, b = blockIdx.z;
    if (t >= npoints * nsample) return;
    const int src = __ldg(&idx[(size_t)b * npoints * nsample + t]);
    atomicAdd(&grad_points[((size_t)b * c + ci) * n + src], grad_out[((size_t)b * c + ci) * npoints * nsample + t]);
}

// ------------------------------------------------------------------ pointnet2_batch: 3-NN + interpolation
__global__ void __launch_bounds__(256) three_nn_batch_kernel(int n, int m, const float* __restrict__ unknown,
                                                             const float* __restrict__ known, float* __restrict__ dist2,
                                                             int* __restrict__ idx) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x, b = blockIdx.y;
    if (p >= n) return;
    const float* u = unknown + ((size_t)b * n + p) * 3;
    const float* kn = known + (size_t)b * m * 3;
    const float ux = u[0], uy = u[1], uz = u[2];
    float b1 = INFINITY, b2 = INFINITY, b3 = INFINITY;      // the reference's 1e40 double sentinel is +inf once stored as float
    int i1 = 0, i2 = 0, i3 = 0;
    for (int k = 0; k < m; ++k) {
        const float d = sqdist3(ux, uy, uz, __ldg(kn + (size_t)k * 3), __ldg(kn + (size_t)k * 3 + 1), __ldg(kn + (size_t)k * 3 + 2));
        if (d < b1) { b3 = b2; i3 = i2; b2 = b1; i2 = i1; b1 = d; i1 = k; }
        else if (d < b2) { b3 = b2; i3 = i2; b2 = d; i2 = k; }
        else if (d < b3) { b3 = d; i3 = k; }
    }
    float* dd = dist2 + ((size_t)b * n + p) * 3;
    int* ii = idx + ((size_t)b * n + p) * 3;
    dd[0] = b1; dd[1] = b2; dd[2] = b3;
    ii[0] = i1; ii[1] = i2; ii[2] = i3;
}
__global__ void __launch_bounds__(256) three_interp_batch_kernel(int c, int m, int n, const float* __restrict__ points,
                                                                 const int* __restrict__ idx, const float* __restrict__ w,
                                                                 float* __restrict__ out) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x, ci = blockIdx.y, b = blockIdx.z;
    if (p >= n) return;
    const float* ww = w + ((size_t)b * n + p) * 3;
    const int* ii = idx + ((size_t)b * n + p) * 3;
    const float* src = points + ((size_t)b * c + ci) * m;
    // the reference's expression as written, so that nvcc contracts it the same way (SASS of both: w1*p1, fma(w0,p0,.), fma(w2,p2,.))
    out[((size_t)b * c + ci) * n + p] = ww[0] * src[ii[0]] + ww[1] * src[ii[1]] + ww[2] * src[ii[2]];
}
__global__ void __launch_bounds__(256) three_interp_batch_grad_kernel(int c, int n, int m, const float* __restrict__ grad_out,
                                                                      const int* __restrict__ idx, const float* __restrict__ w,
                                                                      float* __restrict__ grad_points) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x, ci = blockIdx.y, b = blockIdx.z;
    if (p >= n) return;
    const float g = grad_out[((size_t)b * c + ci) * n + p];
    const float* ww = w + ((size_t)b * n + p) * 3;
    const int* ii = idx + ((size_t)b * n + p) * 3;
    float* dst = grad_points + ((size_t)b * c + ci) * m;
    atomicAdd(&dst[ii[0]], g * ww[0]);
    atomicAdd(&dst[ii[1]], g * ww[1]);
    atomicAdd(&dst[ii[2]], g * ww[2]);
}

// ------------------------------------------------------------------ roipoint_pool3d
// the reference's predicate, expression for expression (roipoint_pool3d_kernel.cu:14-36): float cos/sin of -rz, the height
// test against the double dz / 2.0, MARGIN = 1e-5f added in double
__device__ __forceinline__ int roipoint_in_box(const float* pt, const float* box) {
    const float MARGIN = 1e-5;
    const float x = pt[0], y = pt[1], z = pt[2];
    const float cx = box[0], cy = box[1], cz = box[2], dx = box[3], dy = box[4], dz = box[5], rz = box[6];
    if (fabsf(z - cz) > dz / 2.0) return 0;
    const float sx = x - cx, sy = y - cy;
    const float cosa = cos(-rz), sina = sin(-rz);
    const float lx = sx * cosa + sy * (-sina);
    const float ly = sx * sina + sy * cosa;
    return (fabs(lx) < dx / 2.0 + MARGIN) & (fabs(ly) < dy / 2.0 + MARGIN);
}

// one warp per (frame, box): first `S` inside points in index order, cyclic duplication when fewer, empty flag when none
__global__ void __launch_bounds__(256) roipoint_pool3d_kernel(int N, int M, int C, int S, const float* __restrict__ xyz,
                                                              const float* __restrict__ boxes, const float* __restrict__ feat,
                                                              float* __restrict__ pooled, int* __restrict__ empty_flag,
                                                              int* __restrict__ pts_idx) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int box = blockIdx.x * 8 + warp, b = blockIdx.y;
    if (box >= M) return;
    const float* bx = boxes + ((size_t)b * M + box) * 7;
    const float* pts = xyz + (size_t)b * N * 3;
    int* list = pts_idx + ((size_t)b * M + box) * S;
    int cnt = 0;
    for (int k0 = 0; k0 < N && cnt < S; k0 += 32) {
        const int k = k0 + lane;
        const bool in = k < N && roipoint_in_box(pts + (size_t)k * 3, bx);
        const unsigned int msk = __ballot_sync(0xffffffffu, in);
        const int pos = cnt + __popc(msk & ((1u << lane) - 1u));
        if (in && pos < S) list[pos] = k;
        cnt += __popc(msk);
    }
    cnt = min(cnt, S);
    __syncwarp();
    if (cnt == 0) {
        if (lane == 0) empty_flag[(size_t)b * M + box] = 1;
        return;                                       // pooled row stays as the caller allocated it (zeros), as in the reference
    }
    if (lane == 0) empty_flag[(size_t)b * M + box] = 0;
    float* dst = pooled + ((size_t)b * M + box) * S * (3 + C);
    for (int s = 0; s < S; ++s) {
        const int src = list[s < cnt ? s : (s % cnt)];   // k % cnt duplication (get_pooled_idx)
        float* d = dst + (size_t)s * (3 + C);
        if (lane < 3) d[lane] = pts[(size_t)src * 3 + lane];
        const float* f = feat + ((size_t)b * N + src) * C;
        for (int j = lane; j < C; j += 32) d[3 + j] = f[j];
    }
}

}  // namespace

extern "C" int crb3d_voxel_query_stack(int M, int R1, int R2, int R3, int nsample, float radius, int z_range, int y_range,
                                       int x_range, const float* new_xyz, const float* xyz, const int* new_coords,
                                       const int* point_indices, int* idx, cudaStream_t stream) {
    if (M < 0 || nsample <= 0 || R1 <= 0 || R2 <= 0 || R3 <= 0) return CRB3D_ERR_ARG;
    if (M == 0) return CRB3D_OK;
    if (!new_xyz || !xyz || !new_coords || !point_indices || !idx) return CRB3D_ERR_ARG;
    voxel_query_kernel<<<(unsigned)crb3d_divup(M, 256), 256, 0, stream>>>(M, R1, R2, R3, nsample, radius, z_range, y_range, x_range,
                                                                         new_xyz, xyz, new_coords, point_indices, idx);
    CRB3D_CHECK_LAUNCH();
    return CRB3D_OK;
}

// idx (b, m, nsample) int32 zero-filled by the caller (reference contract: an empty ball keeps its zeros)
extern "C" int crb3d_ball_query_batch(int b, int n, int m, float radius, int nsample, const float* new_xyz, const float* xyz,
                                      int* idx, cudaStream_t stream) {
    if (b < 0 || n < 0 || m < 0 || nsample <= 0) return CRB3D_ERR_ARG;
    if (b == 0 || m == 0 || n == 0) return CRB3D_OK;
    if (!new_xyz || !xyz || !idx) return CRB3D_ERR_ARG;
    ball_query_batch_kernel<<<dim3((unsigned)crb3d_divup(m, BQ_WARPS), (unsigned)b), BQ_WARPS * 32, 0, stream>>>(n, m, radius, nsample,
                                                                                                              new_xyz, xyz, idx);
    CRB3D_CHECK_LAUNCH();
    return CRB3D_OK;
}

// points (b, c, n), idx (b, npoints, nsample) -> out (b, c, npoints, nsample); gather_points is the nsample = 1 case
extern "C" int crb3d_group_points_batch(int b, int c, int n, int npoints, int nsample, const float* points, const int* idx,
                                        float* out, cudaStream_t stream) {
    if (b < 0 || c < 0 || n < 0 || npoints < 0 || nsample <= 0) return CRB3D_ERR_ARG;
    if (b == 0 || c == 0 || npoints == 0) return CRB3D_OK;
    if (!points || !idx || !out) return CRB3D_ERR_ARG;
    group_points_batch_kernel<<<dim3((unsigned)crb3d_divup((int64_t)npoints * nsample, 256), (unsigned)c, (unsigned)b), 256, 0, stream>>>(
        c, n, npoints, nsample, points, idx, out);
    CRB3D_CHECK_LAUNCH();
    return CRB3D_OK;
}
extern "C" int crb3d_group_points_grad_batch(int b, int c, int n, int npoints, int nsample, const float* grad_out, const int* idx,
                                             float* grad_points, cudaStream_t stream) {
    if (b < 0 || c < 0 || n < 0 || npoints < 0 || nsample <= 0) return CRB3D_ERR_ARG;
    if (b == 0 || c == 0 || npoints == 0) return CRB3D_OK;
    if (!grad_out || !idx || !grad_points) return CRB3D_ERR_ARG;
    group_points_batch_grad_kernel<<<dim3((unsigned)crb3d_divup((int64_t)npoints * nsample, 256), (unsigned)c, (unsigned)b), 256, 0, stream>>>(
        c, n, npoints, nsample, grad_out, idx, grad_points);
    CRB3D_CHECK_LAUNCH();
    return CRB3D_OK;
}

extern "C" int crb3d_three_nn_batch(int b, int n, int m, const float* unknown, const float* known, float* dist2, int* idx,
                                    cudaStream_t stream) {
    if (b < 0 || n < 0 || m < 0) return CRB3D_ERR_ARG;
    if (b == 0 || n == 0) return CRB3D_OK;
    if (!unknown || !known || !dist2 || !idx) return CRB3D_ERR_ARG;
    three_nn_batch_kernel<<<dim3((unsigned)crb3d_divup(n, 256), (unsigned)b), 256, 0, stream>>>(n, m, unknown, known, dist2, idx);
    CRB3D_CHECK_LAUNCH();
    return CRB3D_OK;
}
extern "C" int crb3d_three_interpolate_batch(int b, int c, int m, int n, const float* points, const int* idx, const float* weight,
                                             float* out, cudaStream_t stream) {
    if (b < 0 || c < 0 || m < 0 || n < 0) return CRB3D_ERR_ARG;
    if (b == 0 || c == 0 || n == 0) return CRB3D_OK;
    if (!points || !idx || !weight || !out) return CRB3D_ERR_ARG;
    three_interp_batch_kernel<<<dim3((unsigned)crb3d_divup(n, 256), (unsigned)c, (unsigned)b), 256, 0, stream>>>(c, m, n, points, idx, weight, out);
    CRB3D_CHECK_LAUNCH();
    return CRB3D_OK;
}
extern "C" int crb3d_three_interpolate_grad_batch(int b, int c, int n, int m, const float* grad_out, const int* idx,
                                                  const float* weight, float* grad_points, cudaStream_t stream) {
    if (b < 0 || c < 0 || m < 0 || n < 0) return CRB3D_ERR_ARG;
    if (b == 0 || c == 0 || n == 0) return CRB3D_OK;
    if (!grad_out || !idx || !weight || !grad_points) return CRB3D_ERR_ARG;
    three_interp_batch_grad_kernel<<<dim3((unsigned)crb3d_divup(n, 256), (unsigned)c, (unsigned)b), 256, 0, stream>>>(c, n, m, grad_out, idx,
                                                                                                                 weight, grad_points);
    CRB3D_CHECK_LAUNCH();
    return CRB3D_OK;
}

// xyz (B, N, 3), boxes3d (B, M, 7), pts_feature (B, N, C) -> pooled (B, M, S, 3 + C) (caller zero-filled), empty_flag (B, M) int32.
// ws: B*M*S int32 (the per-box point lists; the reference cudaMallocs a (B, N, M) matrix + this list inside the call)
extern "C" int crb3d_roipoint_pool3d_workspace_bytes(int B, int M, int S, size_t* bytes) {
    if (!bytes || B < 0 || M < 0 || S < 0) return CRB3D_ERR_ARG;
    *bytes = crb3d_align(sizeof(int) * (size_t)(B > 0 ? B : 1) * (size_t)(M > 0 ? M : 1) * (size_t)(S > 0 ? S : 1));
    return CRB3D_OK;
}
extern "C" int crb3d_roipoint_pool3d_forward(int B, int N, int M, int C, int S, const float* xyz, const float* boxes3d,
                                             const float* pts_feature, float* pooled, int* empty_flag, void* ws, size_t ws_bytes,
                                             cudaStream_t stream) {
    if (B < 0 || N < 0 || M < 0 || C < 0 || S <= 0) return CRB3D_ERR_ARG;
    if (B == 0 || M == 0) return CRB3D_OK;
    if (!empty_flag) return CRB3D_ERR_ARG;
    if (N == 0) return crb3d_fill_i32(empty_flag, (size_t)B * M, 1, stream);      // frames without points: every box is empty
    if (!xyz || !boxes3d || (C > 0 && !pts_feature) || !pooled) return CRB3D_ERR_ARG;
    WsCursor c(ws, ws_bytes);
    int* pts_idx = c.take<int>((size_t)B * M * S);
    if (!c.ok) return CRB3D_ERR_WORKSPACE;
    roipoint_pool3d_kernel<<<dim3((unsigned)crb3d_divup(M, 8), (unsigned)B), 256, 0, stream>>>(N, M, C, S, xyz, boxes3d, pts_feature, pooled,
                                                                                            empty_flag, pts_idx);
    CRB3D_CHECK_LAUNCH();
    return CRB3D_OK;
}
