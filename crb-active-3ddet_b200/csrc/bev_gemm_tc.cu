// Dense 1x1-conv / deconv / head GEMMs of the BEV stack on the 5th-gen tensor cores (tcgen05, TF32 in, fp32 accumulate
// in TMEM), operands staged by TMA, with the layer's bias + ReLU + output placement fused into the epilogue.
//
// Replaces, on the inference path (reference: cuDNN/cuBLAS calls + separate elementwise kernels),
//   pcdet/models/backbones_2d/base_bev_backbone.py:60-78,100-108  deblocks: ConvTranspose2d(k = stride) + BN + ReLU and
//                                                                 torch.cat(ups, dim=1) - the concat is never
//                                                                 materialised by a copy: each deblock's GEMM writes
//                                                                 its channel slice of the (B,H,W,sum C) map directly
//   pcdet/models/dense_heads/anchor_head_single.py:18-32,41-58     conv_cls / conv_box / conv_dir_cls (three 1x1 convs
//                                                                 + permute + contiguous) as ONE GEMM with N = 18+42+12
//                                                                 whose epilogue writes the three (B, A, *) tensors
// D[m, n] = sum_k A[m, k] * W[n, k] (+ bias[n]) (ReLU);  A = channels-last activations (pixels x C_in), W = [N][K].
// A ConvTranspose2d with kernel == stride == 2 is four such GEMMs (one per output sub-position dy,dx), selected by
// blockIdx.x, whose rows land on the interleaved output pixels (2y+dy, 2x+dx).
//
// One CTA = one 128-row tile x all N columns. Warp roles (mbarrier pipelined, STAGES deep):
//   warp 0, one lane : TMA producer - per 32-channel k-block one 128x32 box of A and one Nx32 box of W, 128B swizzle
//   warp 1, one lane : tcgen05.mma issuer (4 x M128 x N x K8 per k-block), tcgen05.commit frees the stage
//   warps 2-5        : epilogue - tcgen05.ld, bias/ReLU, rows staged in shared memory (aliasing the drained pipeline
//                      buffers) and written out as whole 128-byte lines
// Two CTAs are resident per SM (TMEM 2 x 256 columns), so one CTA's epilogue overlaps the other's main loop; the
// kernels are HBM-bound (K <= 512): bytes/row = 4*(K + N).
#include "tc_common.cuh"

namespace {

using namespace tc;

constexpr int TILE_M = 128;
constexpr int BK = 32;  // floats per k-block = one 128-byte swizzle row

struct GemmOut {
    float* ptr[3];            // up to three column segments, each to its own tensor
    int col_begin[3];
    int width[3];
    long long row_stride[3];  // floats between consecutive output rows of the segment
    int n_seg;
    int up;                   // 0: output row = GEMM row; 2: 2x2 transposed conv - row (b,y,x) -> (b, 2y+dy, 2x+dx)
    int in_w, in_h;           // input spatial size when up == 2
};

template <int N>
struct Cfg {
    static constexpr int CH = N < 128 ? N : 128;             // columns staged per epilogue pass
    static constexpr int PITCH = CH + 4;                     // floats; +4 keeps float4 rows conflict-free
    static constexpr int TMEM_COLS = N <= 32 ? 32 : (N <= 64 ? 64 : (N <= 128 ? 128 : 256));
    static constexpr int A_BYTES = TILE_M * 128;
    static constexpr int B_BYTES = N * 128;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
};

template <int N, int STAGES>
__global__ void __launch_bounds__(192) bev_gemm_tc(const __grid_constant__ CUtensorMap amap,
                                                   const __grid_constant__ CUtensorMap wmap, int M, int K,
                                                   const float* __restrict__ bias, int relu,
                                                   const __grid_constant__ GemmOut out) {
    using C = Cfg<N>;
    static_assert(STAGES * C::STAGE_BYTES >= TILE_M * C::PITCH * 4, "staging must fit in the pipeline buffers");
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t full_bar[STAGES];
    __shared__ uint64_t empty_bar[STAGES];
    __shared__ uint64_t acc_bar;
    __shared__ uint32_t tmem_base_s;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int sub = blockIdx.x;                 // output sub-position of a 2x2 transposed conv (0 otherwise)
    const int m0 = blockIdx.y * TILE_M;
    const int nkb = K / BK;

    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        mbar_init(&acc_bar, 1);
        mbar_fence_init();
        tma_prefetch_desc(&amap);
        tma_prefetch_desc(&wmap);
    }
    if (warp == 1) tmem_alloc<C::TMEM_COLS>(&tmem_base_s);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_s;
    const uint32_t smem_base = smem_u32(smem);

    if (warp == 0) {
        if (lane == 0) {
            for (int kb = 0; kb < nkb; ++kb) {
                const int stage = kb % STAGES;
                if (kb >= STAGES) mbar_wait(&empty_bar[stage], ((kb / STAGES) - 1) & 1);
                const uint32_t a_dst = smem_base + stage * C::STAGE_BYTES, b_dst = a_dst + C::A_BYTES;
                mbar_expect_tx(&full_bar[stage], C::STAGE_BYTES);
                tma_load_2d(a_dst, &amap, kb * BK, m0, &full_bar[stage]);
                tma_load_2d(b_dst, &wmap, kb * BK, sub * N, &full_bar[stage]);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc = idesc_tf32(TILE_M, N);
            for (int kb = 0; kb < nkb; ++kb) {
                const int stage = kb % STAGES;
                mbar_wait(&full_bar[stage], (kb / STAGES) & 1);
                tc_fence_after();
                const uint32_t a_base = smem_base + stage * C::STAGE_BYTES, b_base = a_base + C::A_BYTES;
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    umma_tf32(tmem_base, desc_sw128(a_base + j * 32), desc_sw128(b_base + j * 32), idesc, (kb > 0 || j > 0) ? 1u : 0u);
                umma_commit(&empty_bar[stage]);
                if (kb == nkb - 1) umma_commit(&acc_bar);
            }
        }
    } else {
        // ================================ epilogue (warps 2..5 -> TMEM lane quarters 2,3,0,1) =========================
        const int q = warp & 3;
        const int r = q * 32 + lane;              // tile row owned by this thread in the TMEM read
        mbar_wait(&acc_bar, 0);
        tc_fence_after();
        float* stage_f = reinterpret_cast<float*>(smem) + (size_t)q * 32 * C::PITCH;  // this warp's 32 staged rows
        // output row offsets of this warp's rows (lane l holds row q*32 + l); -1 = beyond M
        long long orow;
        {
            const long long m = (long long)m0 + r;
            if (m >= M) orow = -1;
            else if (out.up == 2) {
                const int hw = out.in_w * out.in_h;
                const int b = (int)(m / hw), rem = (int)(m - (long long)b * hw);
                const int y = rem / out.in_w, x = rem - y * out.in_w;
                orow = ((long long)b * (2 * out.in_h) + 2 * y + (sub >> 1)) * (2 * out.in_w) + 2 * x + (sub & 1);
            } else orow = m;
        }
#pragma unroll 1
        for (int c_begin = 0; c_begin < N; c_begin += C::CH) {
#pragma unroll 1
            for (int c0 = 0; c0 < C::CH; c0 += 32) {
                uint32_t v[32];
                tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(c_begin + c0), v);
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                    if (c0 + j >= C::CH) break;
                    float4 w;
                    float* wp = reinterpret_cast<float*>(&w);
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        float x = __uint_as_float(v[j + u]);
                        if (bias) x += __ldg(&bias[c_begin + c0 + j + u]);
                        if (relu) x = fmaxf(x, 0.0f);
                        wp[u] = x;
                    }
                    *reinterpret_cast<float4*>(stage_f + (size_t)lane * C::PITCH + c0 + j) = w;
                }
            }
            __syncwarp();
            for (int s = 0; s < out.n_seg; ++s) {
                const int lo = max(out.col_begin[s], c_begin), hi = min(out.col_begin[s] + out.width[s], c_begin + C::CH);
                const int w = hi - lo;
                if (w <= 0) continue;
                float* base = out.ptr[s] + (lo - out.col_begin[s]);
                const long long stride = out.row_stride[s];
                const int soff = lo - c_begin;
                if (((w | soff | (lo - out.col_begin[s])) & 3) == 0 && (stride & 3) == 0 && ((uintptr_t)out.ptr[s] & 15) == 0) {
                    const int w4 = w >> 2;                      // float4 per row
                    for (int e = lane; e < 32 * w4; e += 32) {
                        const int rr = e / w4, c4 = e - rr * w4;
                        const long long orr = __shfl_sync(0xffffffffu, orow, rr);
                        if (orr >= 0)
                            *reinterpret_cast<float4*>(base + orr * stride + c4 * 4) =
                                *reinterpret_cast<const float4*>(stage_f + (size_t)rr * C::PITCH + soff + c4 * 4);
                    }
                } else {
                    for (int e = lane; e < 32 * w; e += 32) {
                        const int rr = e / w, c = e - rr * w;
                        const long long orr = __shfl_sync(0xffffffffu, orow, rr);
                        if (orr >= 0) base[orr * stride + c] = stage_f[(size_t)rr * C::PITCH + soff + c];
                    }
                }
            }
            __syncwarp();
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc<C::TMEM_COLS>(tmem_base);
}

template <int N, int STAGES>
int launch_gemm(const float* A, long long M, int K, long long lda, const float* W, int n_sub, const float* bias, int relu,
                const GemmOut& out, cudaStream_t stream) {
    using C = Cfg<N>;
    CUtensorMap amap, wmap;
    {
        const uint64_t dims[2] = {(uint64_t)K, (uint64_t)M}, strides[1] = {(uint64_t)lda * 4};
        const uint32_t box[2] = {BK, TILE_M};
        int rc = make_map_f32(&amap, A, 2, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
        if (rc) return rc;
    }
    {
        const uint64_t dims[2] = {(uint64_t)K, (uint64_t)N * n_sub}, strides[1] = {(uint64_t)K * 4};
        const uint32_t box[2] = {BK, (uint32_t)N};
        int rc = make_map_f32(&wmap, W, 2, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
        if (rc) return rc;
    }
    constexpr size_t smem = (size_t)STAGES * C::STAGE_BYTES + 1024;
    auto kern = bev_gemm_tc<N, STAGES>;
    static bool attr_set = false;
    if (!attr_set) {
        CRB3D_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set = true;
    }
    kern<<<dim3((unsigned)n_sub, (unsigned)crb3d_divup(M, TILE_M)), 192, smem, stream>>>(amap, wmap, (int)M, K, bias, relu, out);
    CRB3D_CHECK_LAUNCH();
    return CRB3D_OK;
}

}  // namespace

// A: (M, K) fp32 rows `lda` floats apart (channels-last activations); W: contiguous [n_sub][N][K]; bias: N floats or null.
// Output: n_seg column segments (seg s = columns [col_begin[s], col_begin[s]+width[s]) -> out_ptr[s] + row*row_stride[s]).
// up = 0: output row = GEMM row, n_sub must be 1. up = 2: ConvTranspose2d(kernel = stride = 2): n_sub = 4 weight slices
// ordered (dy, dx), GEMM row (b, y, x) of an in_h x in_w map lands on output pixel (b, 2y+dy, 2x+dx).
// Supported: K % 32 == 0, N in {80, 128, 256} (pad weights/bias with zero rows to reach a supported N).
extern "C" int crb3d_bev_gemm_tf32(const float* A, long long M, int K, long long lda, const float* W, int N, int n_sub,
                                   const float* bias, int relu, int n_seg, float* const* out_ptr, const int* col_begin,
                                   const int* width, const long long* row_stride, int up, int in_h, int in_w,
                                   cudaStream_t stream) {
    if (!A || !W || M < 0 || K <= 0 || N <= 0 || n_seg < 1 || n_seg > 3 || !out_ptr || !col_begin || !width || !row_stride)
        return CRB3D_ERR_ARG;
    if (M == 0) return CRB3D_OK;
    if (K % BK != 0 || lda % 4 != 0 || M > 0x7fffffffLL) return CRB3D_ERR_UNSUPPORTED;
    if ((up == 0 && n_sub != 1) || (up == 2 && (n_sub != 4 || in_h <= 0 || in_w <= 0 || M % ((long long)in_h * in_w) != 0)) ||
        (up != 0 && up != 2))
        return CRB3D_ERR_ARG;
    GemmOut o{};
    o.n_seg = n_seg; o.up = up; o.in_h = in_h; o.in_w = in_w;
    for (int s = 0; s < n_seg; ++s) {
        if (!out_ptr[s] || col_begin[s] < 0 || width[s] <= 0 || col_begin[s] + width[s] > N) return CRB3D_ERR_ARG;
        o.ptr[s] = out_ptr[s]; o.col_begin[s] = col_begin[s]; o.width[s] = width[s]; o.row_stride[s] = row_stride[s];
    }
    if (N == 256) return launch_gemm<256, 2>(A, M, K, lda, W, n_sub, bias, relu, o, stream);   // 2 x 48 KB
    if (N == 128) return launch_gemm<128, 3>(A, M, K, lda, W, n_sub, bias, relu, o, stream);   // 3 x 32 KB
    if (N == 80) return launch_gemm<80, 3>(A, M, K, lda, W, n_sub, bias, relu, o, stream);     // 3 x 26 KB
    return CRB3D_ERR_UNSUPPORTED;
}
