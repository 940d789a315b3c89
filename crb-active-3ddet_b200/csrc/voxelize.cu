// Hard voxelization + fused MeanVFE on the device.
//
// Replaces (reference, /root/reference):
//   pcdet/datasets/processor/data_processor.py:15-60,115-143  VoxelGeneratorWrapper -> spconv.utils.Point2VoxelCPU3d
//   pcdet/models/backbones_3d/vfe/mean_vfe.py:14-31            MeanVFE.forward
//
// Semantics reproduced (spconv 2.1 Point2VoxelCPU): points are visited in input order per
// frame; c = floor((p - lo) / vsize) per axis in fp32; a point outside the grid is dropped; voxels are numbered
// in first-seen order; a voxel keeps its first `max_pts` points; once `max_voxels` voxels exist in a frame,
// points falling in unseen voxels are dropped.
//
// The serial first-seen rule is made parallel without a sort:
//   1. hash insert of the linear voxel key, atomicMin of the point index      -> the voxel "leader" (first point)
//   2. max_pts-1 further atomicMin rounds (only points above the previous min) -> 2nd..P-th smallest index
//   3. exclusive scan of the leader flags in point order                       -> first-seen voxel number
//   4. one thread per leader writes coords / count / (optional) padded points / mean in slot order
// All integer outputs are bit-identical to the serial algorithm; the mean is summed in slot order.
#include "common.cuh"
#include <limits.h>

namespace {

struct VoxParams {
    float lo[3];
    float vs[3];
    int grid[3];  // x, y, z
    int pt_stride, xyz_col, feat_col, n_feat;
    int max_pts, max_voxels, batch_size;
};

__device__ __forceinline__ int frame_of(const int* __restrict__ off, int B, int64_t p) {
    int lo = 0, hi = B;  // off[lo] <= p < off[hi]
    while (hi - lo > 1) {
        int mid = (lo + hi) >> 1;
        if (p >= off[mid]) lo = mid; else hi = mid;
    }
    return lo;
}

__global__ void __launch_bounds__(256) vox_insert(const float* __restrict__ pts, int64_t n, VoxParams P,
                                                  const int* __restrict__ frame_off,
                                                  unsigned long long* __restrict__ keys, uint32_t cap_mask,
                                                  int* __restrict__ mins, int* __restrict__ pt_slot) {
    int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    if (p >= frame_off[P.batch_size]) { pt_slot[p] = -1; return; }   // rows of a capacity-sized buffer beyond the last frame
    const float* q = pts + p * P.pt_stride + P.xyz_col;
    int c[3];
    bool ok = true;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        float f = floorf(__fdiv_rn(__fsub_rn(q[j], P.lo[j]), P.vs[j]));
        // NaN coordinates are rejected (the serial reference leaves them undefined).
        ok = ok && (f >= 0.0f) && (f < (float)P.grid[j]);
        c[j] = ok ? (int)f : 0;
    }
    if (!ok) { pt_slot[p] = -1; return; }
    const int b = frame_of(frame_off, P.batch_size, p);
    unsigned long long key = (((unsigned long long)b * P.grid[2] + c[2]) * P.grid[1] + c[1]) * P.grid[0] + c[0];
    uint32_t s = hash_insert(keys, cap_mask, key);
    pt_slot[p] = (int)s;
    atomicMin(&mins[(size_t)s * P.max_pts], (int)p);
}

__global__ void __launch_bounds__(256) vox_round(int64_t n, int round, int max_pts, const int* __restrict__ pt_slot,
                                                 int* __restrict__ mins) {
    int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    int s = pt_slot[p];
    if (s < 0) return;
    int* m = mins + (size_t)s * max_pts;
    if ((int)p > m[round - 1]) atomicMin(&m[round], (int)p);
}

__global__ void __launch_bounds__(256) vox_leader_flags(int64_t n, int max_pts, const int* __restrict__ pt_slot,
                                                        const int* __restrict__ mins, int* __restrict__ flags) {
    int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    int s = pt_slot[p];
    flags[p] = (s >= 0 && mins[(size_t)s * max_pts] == (int)p) ? 1 : 0;
}

// One thread: per-frame voxel counts (clamped to max_voxels) -> output row offsets.
__global__ void vox_frame_offsets(const int* __restrict__ rank, const int* __restrict__ total_leaders,
                                  const int* __restrict__ frame_off, int B, int64_t n, int max_voxels,
                                  int* __restrict__ frame_rank0, int* __restrict__ out_off) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    int acc = 0;
    for (int b = 0; b < B; ++b) {
        int s0 = frame_off[b], s1 = frame_off[b + 1];
        int r0 = (s0 < n) ? rank[s0] : *total_leaders;
        int r1 = (s1 < n) ? rank[s1] : *total_leaders;
        frame_rank0[b] = r0;
        out_off[b] = acc;
        int c = r1 - r0;
        acc += (c < max_voxels) ? c : max_voxels;
    }
    out_off[B] = acc;
}

__global__ void __launch_bounds__(128) vox_finalize(const float* __restrict__ pts, int64_t n, VoxParams P,
                                                    const int* __restrict__ frame_off, const int* __restrict__ pt_slot,
                                                    const int* __restrict__ mins, const int* __restrict__ flags,
                                                    const int* __restrict__ rank, const int* __restrict__ frame_rank0,
                                                    const int* __restrict__ out_off, float* __restrict__ mean,
                                                    float* __restrict__ voxels, int* __restrict__ coords,
                                                    int* __restrict__ num_points) {
    int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n || !flags[p]) return;
    const int b = frame_of(frame_off, P.batch_size, p);
    const int id = rank[p] - frame_rank0[b];
    if (id >= P.max_voxels) return;
    const int64_t row = (int64_t)out_off[b] + id;
    const float* q = pts + p * P.pt_stride + P.xyz_col;
    int c[3];
#pragma unroll
    for (int j = 0; j < 3; ++j) c[j] = (int)floorf(__fdiv_rn(__fsub_rn(q[j], P.lo[j]), P.vs[j]));
    coords[row * 4 + 0] = b;
    coords[row * 4 + 1] = c[2];
    coords[row * 4 + 2] = c[1];
    coords[row * 4 + 3] = c[0];
    const int* m = mins + (size_t)pt_slot[p] * P.max_pts;
    int cnt = 0;
    for (int r = 0; r < P.max_pts; ++r) cnt += (m[r] != INT_MAX);
    num_points[row] = cnt;
    for (int f = 0; f < P.n_feat; ++f) {
        float s = 0.0f;
        for (int r = 0; r < cnt; ++r) {
            float v = pts[(int64_t)m[r] * P.pt_stride + P.feat_col + f];
            s = __fadd_rn(s, v);
            if (voxels) voxels[(row * P.max_pts + r) * P.n_feat + f] = v;
        }
        if (voxels)
            for (int r = cnt; r < P.max_pts; ++r) voxels[(row * P.max_pts + r) * P.n_feat + f] = 0.0f;
        // mean_vfe.py:26-29: sum / clamp_min(num, 1)
        if (mean) mean[row * P.n_feat + f] = __fdiv_rn(s, (float)(cnt > 0 ? cnt : 1));
    }
}

struct VoxWs {
    unsigned long long* keys;
    int *mins, *pt_slot, *flags, *rank, *scan_ws, *total, *frame_rank0;
    uint32_t cap;
};

bool carve(WsCursor& c, int64_t n, int B, int max_pts, VoxWs& w) {
    w.cap = crb3d_next_pow2((uint64_t)(n > 0 ? n : 1) * 2);
    w.keys = c.take<unsigned long long>(w.cap);
    w.mins = c.take<int>((size_t)w.cap * max_pts);
    w.pt_slot = c.take<int>(n);
    w.flags = c.take<int>(n);
    w.rank = c.take<int>(n);
    w.scan_ws = c.take<int>(crb3d_scan_ws_ints(n));
    w.total = c.take<int>(1);
    w.frame_rank0 = c.take<int>(B + 1);
    return c.ok;
}

}  // namespace

extern "C" int crb3d_voxelize_workspace_bytes(int64_t n_points, int batch_size, int max_pts, size_t* bytes) {
    if (!bytes || n_points < 0 || batch_size <= 0 || max_pts <= 0) return CRB3D_ERR_ARG;
    WsCursor c(nullptr, 0);
    VoxWs w;
    carve(c, n_points, batch_size, max_pts, w);
    *bytes = c.off;
    return CRB3D_OK;
}

extern "C" int crb3d_voxelize(const float* points, int64_t n_points, int pt_stride, int xyz_col, int feat_col,
                              int n_feat, const int* frame_offsets, int batch_size, const float* range6,
                              const float* vsize3, const int* grid3, int max_pts, int max_voxels, float* mean_feats,
                              float* voxels, int* coords, int* num_points, int* frame_voxel_offsets, void* ws,
                              size_t ws_bytes, cudaStream_t stream) {
    if (n_points < 0 || batch_size <= 0 || max_pts <= 0 || max_voxels <= 0 || !frame_offsets || !range6 ||
        !vsize3 || !grid3 || !coords || !num_points || !frame_voxel_offsets)
        return CRB3D_ERR_ARG;
    if (n_points >= INT_MAX) return CRB3D_ERR_UNSUPPORTED;
    WsCursor c(ws, ws_bytes);
    VoxWs w;
    if (!carve(c, n_points, batch_size, max_pts, w)) return CRB3D_ERR_WORKSPACE;
    if (n_points == 0) {
        CRB3D_CUDA(cudaMemsetAsync(frame_voxel_offsets, 0, sizeof(int) * (batch_size + 1), stream));
        return CRB3D_OK;
    }
    VoxParams P;
    for (int j = 0; j < 3; ++j) { P.lo[j] = range6[j]; P.vs[j] = vsize3[j]; P.grid[j] = grid3[j]; }
    P.pt_stride = pt_stride; P.xyz_col = xyz_col; P.feat_col = feat_col; P.n_feat = n_feat;
    P.max_pts = max_pts; P.max_voxels = max_voxels; P.batch_size = batch_size;

    CRB3D_CUDA(cudaMemsetAsync(w.keys, 0xFF, sizeof(unsigned long long) * w.cap, stream));
    { int rc0 = crb3d_fill_i32(w.mins, (size_t)w.cap * max_pts, INT_MAX, stream); if (rc0) return rc0; }
    const unsigned nb = (unsigned)crb3d_divup(n_points, 256);
    vox_insert<<<nb, 256, 0, stream>>>(points, n_points, P, frame_offsets, w.keys, w.cap - 1, w.mins, w.pt_slot);
    for (int r = 1; r < max_pts; ++r) vox_round<<<nb, 256, 0, stream>>>(n_points, r, max_pts, w.pt_slot, w.mins);
    vox_leader_flags<<<nb, 256, 0, stream>>>(n_points, max_pts, w.pt_slot, w.mins, w.flags);
    int rc = crb3d_scan_exclusive_i32(w.flags, w.rank, n_points, w.scan_ws, w.total, stream);
    if (rc) return rc;
    vox_frame_offsets<<<1, 32, 0, stream>>>(w.rank, w.total, frame_offsets, batch_size, n_points, max_voxels,
                                            w.frame_rank0, frame_voxel_offsets);
    vox_finalize<<<(unsigned)crb3d_divup(n_points, 128), 128, 0, stream>>>(
        points, n_points, P, frame_offsets, w.pt_slot, w.mins, w.flags, w.rank, w.frame_rank0, frame_voxel_offsets,
        mean_feats, voxels, coords, num_points);
    CRB3D_CHECK_LAUNCH();
    return CRB3D_OK;
}

CRB3D_DIAG_DEFINE_SETTER(voxelize)
