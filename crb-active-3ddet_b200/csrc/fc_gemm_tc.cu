// Split-K GEMM of the RoI-head FC stack on the 5th-gen tensor cores (tcgen05, TF32 in, fp32 accumulate in TMEM).
//
// Replaces the cuBLAS/cuDNN call behind the first shared FC of
//   pcdet/models/roi_heads/pvrcnn_head.py:21-33,181-202   shared_fc_layer[0] = Conv1d(216 * 128 = 27 648 -> 256, k = 1)
// - 1.8 GFLOP per frame with M = 128 RoIs per frame: a handful of 128-row tiles with a K of 27 648. The reference runs it once
// per Monte-Carlo dropout round (SAMPLING_ROUND = 5, pvrcnn_head.py:187-196) although its input does not change between
// rounds (the dropout sits BEHIND it); crb3d.pvrcnn computes it once and replays only the small layers.
// D[m, n] = sum_k A[m, k] * W[n, k]; epilogue y = relu?(D * scale[n] + shift[n]) (folded eval BatchNorm / bias).
//
// Grid = (M tiles) x (N / 128 column slices) x (K splits): every CTA streams its K range through a 5-stage TMA ring
// (128 x 32 box of A + 128 x 32 box of W per stage, 128B swizzle), accumulates one 128 x 128 tile in TMEM and writes the raw
// fp32 partial to the workspace; a second kernel sums the splits in a fixed order (deterministic) and applies the epilogue.
// With M = 512, N = 256, K = 27 648: 4 x 2 x 18 = 144 CTAs, each reading 2 x 196 KB per stage-round from L2.
#include "tc_common.cuh"

namespace {

using namespace tc;

constexpr int TILE = 128;
constexpr int BK = 32;
constexpr int STAGES = 5;
constexpr int A_BYTES = TILE * 128, B_BYTES = TILE * 128, STAGE_BYTES = A_BYTES + B_BYTES;

__global__ void __launch_bounds__(192, 1) fc_gemm_tc(const __grid_constant__ CUtensorMap amap,
                                                     const __grid_constant__ CUtensorMap wmap, int M, int N, int kb_per_split,
                                                     int nkb, float* __restrict__ partial) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // pointer arithmetic keeps the shared address space (LDS/STS, not generic LD/ST)
    __shared__ uint64_t full_bar[STAGES], empty_bar[STAGES], acc_bar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int m0 = blockIdx.x * TILE, n0 = blockIdx.y * TILE, split = blockIdx.z;
    const int kb0 = split * kb_per_split, kb1 = min(nkb, kb0 + kb_per_split);
    const int n_it = kb1 - kb0;
    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        mbar_init(&acc_bar, 1);
        mbar_fence_init();
        tma_prefetch_desc(&amap);
        tma_prefetch_desc(&wmap);
    }
    if (warp == 1) tmem_alloc<128>(&tmem_base_s);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_s;
    const uint32_t ring = smem_u32(smem);
    if (warp == 0) {
        if (lane == 0) {
            for (int it = 0; it < n_it; ++it) {
                const int stage = it % STAGES;
                if (it >= STAGES) mbar_wait(&empty_bar[stage], ((it / STAGES) - 1) & 1, (CRB3D_K_FC_GEMM << 8) | 9, it);
                const uint32_t dst = ring + stage * STAGE_BYTES;
                mbar_expect_tx(&full_bar[stage], STAGE_BYTES);
                tma_load_2d(dst, &amap, (kb0 + it) * BK, m0, &full_bar[stage]);
                tma_load_2d(dst + A_BYTES, &wmap, (kb0 + it) * BK, n0, &full_bar[stage]);
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        const uint32_t idesc = idesc_tf32(TILE, TILE);
        for (int it = 0; it < n_it; ++it) {
            const int stage = it % STAGES;
            mbar_wait(&full_bar[stage], (it / STAGES) & 1, (CRB3D_K_FC_GEMM << 8) | 8, it);
            tc_fence_after();
            const uint32_t a_base = ring + stage * STAGE_BYTES, b_base = a_base + A_BYTES;
            if (elect_one()) {
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    umma_tf32(tmem_base, desc_sw128(a_base + j * 32), desc_sw128(b_base + j * 32), idesc, (it > 0 || j > 0) ? 1u : 0u);
                umma_commit(&empty_bar[stage]);
                if (it == n_it - 1) umma_commit(&acc_bar);
            }
            __syncwarp();
        }
    } else {
        // epilogue: warps 2..5, TMEM lane quarter = warp % 4
        const int q = warp & 3;
        const int r = q * 32 + lane;
        if (n_it > 0) {
            mbar_wait(&acc_bar, 0, (CRB3D_K_FC_GEMM << 8) | 7);
            tc_fence_after();
        }
        const long long m = (long long)m0 + r;
        float* dst = partial + ((size_t)split * M + (size_t)(m < M ? m : 0)) * N + n0;
#pragma unroll 1
        for (int c0 = 0; c0 < TILE; c0 += 32) {
            uint32_t v[32];
            if (n_it > 0) tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + c0, v);
            else {
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = 0u;
            }
            if (m < M) {
#pragma unroll
                for (int j = 0; j < 32; j += 4)
                    *reinterpret_cast<float4*>(dst + c0 + j) = make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]),
                                                                           __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc<128>(tmem_base);
}

__global__ void __launch_bounds__(256) fc_reduce_kernel(const float* __restrict__ partial, int splits, long long MN, int N,
                                                        const float* __restrict__ scale, const float* __restrict__ shift, int relu,
                                                        float* __restrict__ out) {
    const long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (i >= MN) return;
    float4 acc = *reinterpret_cast<const float4*>(partial + i);
    for (int s = 1; s < splits; ++s) {      // fixed order: deterministic
        const float4 p = *reinterpret_cast<const float4*>(partial + (size_t)s * MN + i);
        acc.x += p.x; acc.y += p.y; acc.z += p.z; acc.w += p.w;
    }
    const int n = (int)(i % N);
    float* a = reinterpret_cast<float*>(&acc);
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        float x = a[u];
        if (scale) x = fmaf(x, __ldg(&scale[n + u]), shift ? __ldg(&shift[n + u]) : 0.0f);
        else if (shift) x += __ldg(&shift[n + u]);
        if (relu & 1) x = fmaxf(x, 0.0f);
        a[u] = x;
    }
    *reinterpret_cast<float4*>(out + i) = acc;
}

int pick_splits(int m_tiles, int n_slices, int nkb) {
    int splits = crb3d_num_sms() / (m_tiles * n_slices);
    if (splits < 1) splits = 1;
    if (splits > nkb) splits = nkb;
    return splits;
}

}  // namespace

extern "C" int crb3d_fc_gemm_workspace_bytes(long long M, int N, int K, size_t* bytes) {
    if (!bytes || M < 0 || N <= 0 || K <= 0) return CRB3D_ERR_ARG;
    if (N % TILE != 0 || K % BK != 0) return CRB3D_ERR_UNSUPPORTED;
    const int m_tiles = (int)crb3d_divup(M > 0 ? M : 1, TILE);
    const int kb_per = (int)crb3d_divup(K / BK, pick_splits(m_tiles, N / TILE, K / BK));
    const int splits = (int)crb3d_divup(K / BK, kb_per);
    *bytes = crb3d_align(sizeof(float) * (size_t)splits * (size_t)(M > 0 ? M : 1) * N);
    return CRB3D_OK;
}

// out (M, N) = relu?((A (M, K; rows lda floats apart) @ W (N, K)^T) * scale + shift). N % 128 == 0, K % 32 == 0, lda % 4 == 0.
// A and W are read as TF32 (the caller rounds W once; the tensor core truncates whatever is left).
extern "C" int crb3d_fc_gemm_tf32(const float* A, long long M, int K, long long lda, const float* W, int N, const float* scale,
                                  const float* shift, int relu, float* out, void* ws, size_t ws_bytes, cudaStream_t stream) {
    if (!A || !W || !out || M < 0 || K <= 0 || N <= 0) return CRB3D_ERR_ARG;
    if (M == 0) return CRB3D_OK;
    if (N % TILE != 0 || K % BK != 0 || lda % 4 != 0 || M > 0x7fffffffLL) return CRB3D_ERR_UNSUPPORTED;
    const int m_tiles = (int)crb3d_divup(M, TILE), n_slices = N / TILE, nkb = K / BK;
    const int kb_per = (int)crb3d_divup(nkb, pick_splits(m_tiles, n_slices, nkb));
    const int splits = (int)crb3d_divup(nkb, kb_per);
    WsCursor c(ws, ws_bytes);
    float* partial = c.take<float>((size_t)splits * M * N);
    if (!c.ok) return CRB3D_ERR_WORKSPACE;
    CUtensorMap amap, wmap;
    {
        const uint64_t dims[2] = {(uint64_t)K, (uint64_t)M}, strides[1] = {(uint64_t)lda * 4};
        const uint32_t box[2] = {BK, TILE};
        int rc = make_map_f32(&amap, A, 2, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
        if (rc) return rc;
    }
    {
        const uint64_t dims[2] = {(uint64_t)K, (uint64_t)N}, strides[1] = {(uint64_t)K * 4};
        const uint32_t box[2] = {BK, TILE};
        int rc = make_map_f32(&wmap, W, 2, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
        if (rc) return rc;
    }
    const size_t smem = 1024 + (size_t)STAGES * STAGE_BYTES;
    static bool attr_set[CRB3D_MAX_DEVICES] = {};
    const int dev = crb3d_current_device();
    if (!attr_set[dev]) {
        CRB3D_CUDA(cudaFuncSetAttribute(fc_gemm_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set[dev] = true;
    }
    fc_gemm_tc<<<dim3((unsigned)m_tiles, (unsigned)n_slices, (unsigned)splits), 192, smem, stream>>>(amap, wmap, (int)M, N, kb_per, nkb, partial);
    CRB3D_CHECK_LAUNCH();
    const long long MN = M * N;
    fc_reduce_kernel<<<(unsigned)crb3d_divup(MN / 4, 256), 256, 0, stream>>>(partial, splits, MN, N, scale, shift, relu, out);
    CRB3D_CHECK_LAUNCH();
    return CRB3D_OK;
}

CRB3D_DIAG_DEFINE_SETTER(fc_gemm)
