// Fused set-abstraction layer: QueryAndGroup + shared MLP (1x1 Conv2d + folded BatchNorm + ReLU, up to 3 layers) +
// max-pool over the samples, one pass, nothing materialised.
//
// Replaces, on the inference path, the chain behind one scale of
//   pcdet/ops/pointnet2/pointnet2_stack/pointnet2_modules.py:78-112   StackSAModuleMSG.forward
//     pointnet2_utils.py:107-155  QueryAndGroup (ball query -> group xyz, subtract the centre, group features, cat)
//     nn.Sequential(Conv2d 1x1, BatchNorm2d, ReLU, ...) on the (1, C+3, M, nsample) tensor, F.max_pool2d over nsample
// used by VoxelSetAbstraction (voxel_set_abstraction.py:334-411: 2048 keypoints x 5 sources x 2 radii) and by
// PVRCNNHead.roi_grid_pool (pvrcnn_head.py:68-114: 128 x 216 grid points per frame). The reference materialises the
// grouped (M, C+3, nsample) tensor and every MLP activation in HBM (for the RoI grid: 27 648 x 131 x 16 floats = 232 MB
// per frame and layer); here a warp owns one query point, gathers its samples' rows straight from the source features,
// runs the MLP out of shared-memory weights with the activations in registers / per-warp shared scratch and writes only
// the pooled (M, C_out) row - directly into its column slice of the concatenated multi-scale output.
// Exact fp32 FFMA (the tensor cores are reserved for the RoI-head FC stack, csrc/fc_gemm_tc.cu): per sample the MLP is
// (C+3)*H1 + H1*H2 <= 12.5 k FMAs, the work is bound by shared-memory operand reads, not by HBM.
// idx comes from crb3d_ball_query_stack (first nsample hits in index order, padded with the first hit, idx[m][0] = -1 for
// an empty ball - whose output row is then relu(bias) pushed through the MLP, exactly what the reference computes from its
// zeroed group).
#include "common.cuh"

namespace {

constexpr int SA_WARPS = 8;
constexpr int SB = 4;            // samples processed together (register blocking)
constexpr int MAX_LAYERS = 3;
constexpr int MAX_U = 4;         // output channels per lane: widths up to 128

struct SaMlp {
    int n_layers;
    int cin[MAX_LAYERS];         // input width of layer l (layer 0: 3 + C)
    int cout[MAX_LAYERS];
    int w_off[MAX_LAYERS];       // float offset of layer l's transposed weight [cin][cout] in the packed buffer
    int b_off[MAX_LAYERS];       // ... of its bias (folded BatchNorm shift)
    int total_floats;
};

__global__ void __launch_bounds__(SA_WARPS * 32) sa_group_mlp_maxpool_kernel(
    int B, const float* __restrict__ xyz, const int* __restrict__ xyz_cnt, const float* __restrict__ feat, int C,
    const float* __restrict__ new_xyz, const int* __restrict__ new_cnt, const int* __restrict__ idx, int nsample,
    const float* __restrict__ packed, const __grid_constant__ SaMlp mlp, float* __restrict__ out, int out_stride, int M,
    int in_pitch, int h_pitch) {
    extern __shared__ float sm[];
    float* w_s = sm;                                              // all layers' weights + biases
    float* scratch = sm + ((mlp.total_floats + 3) & ~3);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* in_s = scratch + warp * (SB * in_pitch + 2 * SB * h_pitch);   // [SB][in_pitch]
    float* h_a = in_s + SB * in_pitch;                            // [SB][h_pitch] ping
    float* h_b = h_a + SB * h_pitch;                              // pong
    for (int t = threadIdx.x; t < mlp.total_floats; t += blockDim.x) w_s[t] = __ldg(&packed[t]);
    __syncthreads();
    const int cin0 = mlp.cin[0];
    const int last = mlp.n_layers - 1;
    const int n_u_last = (mlp.cout[last] + 31) >> 5;

    for (int m = blockIdx.x * SA_WARPS + warp; m < M; m += gridDim.x * SA_WARPS) {
        // which frame does query m belong to, and where do that frame's source rows start
        int b = 0, q0 = 0, s0 = 0;
        while (b < B - 1 && m >= q0 + __ldg(&new_cnt[b])) { q0 += __ldg(&new_cnt[b]); s0 += __ldg(&xyz_cnt[b]); ++b; }
        const float cx = __ldg(&new_xyz[(size_t)m * 3]), cy = __ldg(&new_xyz[(size_t)m * 3 + 1]), cz = __ldg(&new_xyz[(size_t)m * 3 + 2]);
        const int my_i = lane < nsample ? __ldg(&idx[(size_t)m * nsample + lane]) : 0;
        const bool empty = __shfl_sync(0xffffffffu, my_i, 0) < 0;
        float mx[MAX_U];
#pragma unroll
        for (int u = 0; u < MAX_U; ++u) mx[u] = -3.0e38f;
        for (int sb = 0; sb < nsample; sb += SB) {
            // ---- gather SB samples: [dx, dy, dz, features...] (zeros for an empty ball, pointnet2_utils.py:141-146)
#pragma unroll
            for (int j = 0; j < SB; ++j) {
                const int s = min(sb + j, nsample - 1);           // nsample not a multiple of SB: repeat the last (max is idempotent)
                const int i = __shfl_sync(0xffffffffu, my_i, s);
                float* dst = in_s + j * in_pitch;
                if (empty) {
                    for (int c = lane; c < cin0; c += 32) dst[c] = 0.0f;
                } else {
                    const size_t row = (size_t)(s0 + i);
                    if (lane < 3) dst[lane] = __fsub_rn(__ldg(&xyz[row * 3 + lane]), lane == 0 ? cx : (lane == 1 ? cy : cz));
                    for (int c = lane; c < C; c += 32) dst[3 + c] = __ldg(&feat[row * C + c]);
                }
            }
            __syncwarp();
            // ---- the MLP: lane owns output channels lane, lane + 32, ... of every layer
            const float* src = in_s;
            int src_pitch = in_pitch;
            float acc[SB][MAX_U];
            for (int l = 0; l < mlp.n_layers; ++l) {
                const int ci = mlp.cin[l], co = mlp.cout[l];
                const float* wt = w_s + mlp.w_off[l];              // [ci][co]
                const float* bs = w_s + mlp.b_off[l];
#pragma unroll
                for (int u = 0; u < MAX_U; ++u) {
                    const int h = lane + 32 * u;
                    const float bias = h < co ? bs[h] : 0.0f;
#pragma unroll
                    for (int j = 0; j < SB; ++j) acc[j][u] = bias;
                }
                const int n_u = (co + 31) >> 5;
                for (int c = 0; c < ci; ++c) {
                    float x[SB];
#pragma unroll
                    for (int j = 0; j < SB; ++j) x[j] = src[j * src_pitch + c];
#pragma unroll
                    for (int u = 0; u < MAX_U; ++u) {
                        if (u < n_u) {
                            const int h = lane + 32 * u;
                            const float w = h < co ? wt[c * co + h] : 0.0f;
#pragma unroll
                            for (int j = 0; j < SB; ++j) acc[j][u] = fmaf(w, x[j], acc[j][u]);
                        }
                    }
                }
                if (l < last) {            // ReLU, hand the activations to the next layer through the warp's scratch
                    float* dsth = (l & 1) ? h_b : h_a;
                    __syncwarp();
#pragma unroll
                    for (int u = 0; u < MAX_U; ++u) {
                        const int h = lane + 32 * u;
                        if (h < co) {
#pragma unroll
                            for (int j = 0; j < SB; ++j) dsth[j * h_pitch + h] = fmaxf(acc[j][u], 0.0f);
                        }
                    }
                    __syncwarp();
                    src = dsth;
                    src_pitch = h_pitch;
                }
            }
#pragma unroll
            for (int u = 0; u < MAX_U; ++u)
#pragma unroll
                for (int j = 0; j < SB; ++j) mx[u] = fmaxf(mx[u], fmaxf(acc[j][u], 0.0f));
            __syncwarp();
        }
#pragma unroll
        for (int u = 0; u < MAX_U; ++u) {
            const int h = lane + 32 * u;
            if (u < n_u_last && h < mlp.cout[last]) out[(size_t)m * out_stride + h] = mx[u];
        }
    }
}

}  // namespace

// One scale of a StackSAModuleMSG, fused (see the header comment).
//   xyz (N,3), xyz_cnt (B) int32, feat (N,C) or null (C = 0), new_xyz (M,3), new_cnt (B) int32,
//   idx (M,nsample) int32 from crb3d_ball_query_stack, nsample <= 32
//   n_layers <= 3, widths[n_layers+1] (HOST): widths[0] = 3 + C, widths[l+1] = output width of layer l (<= 128)
//   packed (DEVICE): per layer the TRANSPOSED weight [widths[l]][widths[l+1]] with the BatchNorm scale folded in, followed by
//   the bias [widths[l+1]] (folded BatchNorm shift); layers back to back
//   out + column offset: (M, out_stride) row-major, the pooled features land in columns [0, widths[n_layers]) of `out`
extern "C" int crb3d_sa_group_mlp_maxpool(int B, const float* xyz, const int* xyz_cnt, const float* feat, int C,
                                          const float* new_xyz, const int* new_cnt, int M, const int* idx, int nsample,
                                          int n_layers, const int* widths, const float* packed, float* out, int out_stride,
                                          cudaStream_t stream) {
    if (B <= 0 || M < 0 || C < 0 || nsample <= 0 || n_layers <= 0 || !widths || !packed || !out) return CRB3D_ERR_ARG;
    if (M == 0) return CRB3D_OK;
    if (!xyz || !xyz_cnt || !new_xyz || !new_cnt || !idx || (C > 0 && !feat)) return CRB3D_ERR_ARG;
    if (n_layers > MAX_LAYERS || nsample > 32 || widths[0] != 3 + C) return CRB3D_ERR_UNSUPPORTED;
    SaMlp mlp;
    mlp.n_layers = n_layers;
    int off = 0, hmax = 4;
    for (int l = 0; l < n_layers; ++l) {
        if (widths[l + 1] <= 0 || widths[l + 1] > 32 * MAX_U) return CRB3D_ERR_UNSUPPORTED;
        mlp.cin[l] = widths[l];
        mlp.cout[l] = widths[l + 1];
        mlp.w_off[l] = off;
        off += widths[l] * widths[l + 1];
        mlp.b_off[l] = off;
        off += widths[l + 1];
        if (widths[l + 1] > hmax) hmax = widths[l + 1];
    }
    mlp.total_floats = off;
    const int in_pitch = (widths[0] + 3) & ~3, h_pitch = (hmax + 3) & ~3;
    const size_t smem = sizeof(float) * (((size_t)off + 3) / 4 * 4 + (size_t)SA_WARPS * (SB * in_pitch + 2 * SB * h_pitch));
    if (smem > 200 * 1024) return CRB3D_ERR_UNSUPPORTED;
    static size_t smem_set[CRB3D_MAX_DEVICES] = {};
    const int dev = crb3d_current_device();
    if (smem > 48 * 1024 && smem > smem_set[dev]) {
        CRB3D_CUDA(cudaFuncSetAttribute(sa_group_mlp_maxpool_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        smem_set[dev] = smem;
    }
    int grid = (int)crb3d_divup(M, SA_WARPS);
    const int cap = crb3d_num_sms() * 4;
    if (grid > cap) grid = cap;
    sa_group_mlp_maxpool_kernel<<<grid, SA_WARPS * 32, smem, stream>>>(B, xyz, xyz_cnt, feat, C, new_xyz, new_cnt, idx, nsample,
                                                                      packed, mlp, out, out_stride, M, in_pitch, h_pitch);
    CRB3D_CHECK_LAUNCH();
    return CRB3D_OK;
}
