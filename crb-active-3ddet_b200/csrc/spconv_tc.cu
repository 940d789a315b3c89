// Sparse 3D convolution on the 5th-gen tensor cores (tcgen05, TF32 inputs, fp32 accumulate in TMEM).
//
// Replaces the external spconv-cu113==2.1.21 implicit-GEMM forward reached from
//   pcdet/models/backbones_3d/spconv_backbone.py:77-117   (C_in >= 16 layers of VoxelBackBone8x)
// Same contract as csrc/spconv_simt.cu:  out[o,:] = sum_k in[nbr[k][o],:] @ W[:,k,:]^T, W = [C_out, K, C_in].
//
// One CTA owns 128 consecutive output rows (the UMMA M). Warp-specialised, mbarrier pipelined:
//   producers (warps 0-6): for every kernel offset k that has a neighbour in the tile, gather the 128 input rows
//       (cp.async 16 B, zero-fill for missing neighbours) and the C_out x C_in weight slice into a 128B-swizzled
//       K-major shared-memory stage; `cp.async.mbarrier.arrive.noinc` signals full[stage] when the copies land.
//   MMA issuer (warp 7, one lane): waits full[stage], issues C_in/8 tcgen05.mma (M=128, N=C_out, K=8, kind::tf32)
//       accumulating ALL offsets into one TMEM tile, tcgen05.commit -> empty[stage] (and -> acc_full at the end).
//   epilogue (warps 0-3): tcgen05.ld 32 lanes x 32 columns, optional scale/shift/ReLU, one store per output row.
// The output is written once, there are no atomics and the summation order is fixed (k ascending).
#include "common.cuh"
#include <cuda.h>

namespace {

constexpr int TILE_M = 128;
constexpr int THREADS = 256;
constexpr int PRODUCERS = 224;  // warps 0..6
constexpr int MAX_K = 27;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "DONE:\n\t}" ::"r"(smem_u32(bar)), "r"(parity));
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, uint32_t src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes));
}
// arrive on `bar` once all cp.async issued so far by this thread have completed (does not bump the pending count)
__device__ __forceinline__ void cp_async_arrive(uint64_t* bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)));
}

// K-major, 128B-swizzle shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start address >> 4 in bits
// [0,14), LBO (ignored for swizzled K-major) = 1 in [16,30), SBO = 1024 B (8 rows x 128 B) >> 4 in [32,46),
// version = 1 in [46,48), layout type SWIZZLE_128B = 2 in [61,64).
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t smem_addr) {
    return (uint64_t)((smem_addr & 0x3FFFF) >> 4) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}

// kind::tf32 instruction descriptor (cute::UMMA::InstrDescriptor): c_format F32 = 1 @4, a/b_format TF32 = 2 @7/@10,
// a/b major K = 0, n_dim = N>>3 @17, m_dim = M>>4 @24.
__device__ __forceinline__ uint32_t make_idesc_tf32(int M, int N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc),
        "r"(accumulate));
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)));
}

// NKB = ceil(C_in / 32) k-blocks (one 128-byte swizzle row holds 32 floats; C_in = 16 is zero-padded to 32).
template <int NKB, int COUT, int STAGES, int MIN_CTAS>
__global__ void __launch_bounds__(THREADS, MIN_CTAS) spconv_fwd_tc(const float* __restrict__ feat, const int* __restrict__ nbr,
                                                                   const __grid_constant__ CUtensorMap wmap, int n_out, int K, int cin,
                                                                   const int* __restrict__ kmap, const float* __restrict__ scale,
                                                                   const float* __restrict__ shift, int relu,
                                                                   float* __restrict__ out) {
    constexpr int A_BYTES = NKB * TILE_M * 128;
    constexpr int B_BYTES = NKB * COUT * 128;
    constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    constexpr int TMEM_COLS = COUT <= 32 ? 32 : (COUT <= 64 ? 64 : (COUT <= 128 ? 128 : 256));
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    __shared__ int rows[MAX_K][TILE_M];
    __shared__ int act[MAX_K];
    __shared__ int n_act_s;
    __shared__ uint64_t full_bar[STAGES];
    __shared__ uint64_t empty_bar[STAGES];
    __shared__ uint64_t acc_bar;
    __shared__ uint32_t tmem_base_s;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int row0 = blockIdx.x * TILE_M;

    // ---- setup: barriers, TMEM, neighbour rows of this tile for every offset
    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full_bar[s], PRODUCERS + 1);  // 224 cp.async arrivals + the TMA expect_tx arrive
            mbar_init(&empty_bar[s], 1);
        }
        mbar_init(&acc_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    if (warp == 7) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "n"(TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    {   // stage A buffers start out zeroed (absent neighbours are never written, see the producer loop)
        float4* z = reinterpret_cast<float4*>(smem);
        for (int t = tid; t < STAGES * STAGE_BYTES / 16; t += THREADS) z[t] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    for (int t = tid; t < K * TILE_M; t += THREADS) {
        const int k = t / TILE_M, r = t - k * TILE_M;
        const int o = row0 + r;
        rows[k][r] = (o < n_out) ? __ldg(&nbr[(size_t)k * n_out + o]) : -1;
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem_base = tmem_base_s;
    // active offsets (warp 0 builds the compact list in ascending k: fixed summation order)
    if (warp == 0) {
        int n = 0;
        for (int k = 0; k < K; ++k) {
            bool any = false;
            for (int r = lane; r < TILE_M; r += 32) any |= rows[k][r] >= 0;
            if (__any_sync(0xffffffffu, any)) { if (lane == 0) act[n] = k; ++n; }
        }
        if (lane == 0) n_act_s = n;
    }
    __syncthreads();
    const int n_act = n_act_s;
    const uint32_t smem_base = smem_u32(smem);

    if (warp < 7) {
        // ================================ producers ================================
        const int vec_per_row = cin >> 2;  // real 16-byte chunks per row (cin % 4 == 0)
        for (int it = 0; it < n_act; ++it) {
            const int stage = it % STAGES;
            if (it >= STAGES) mbar_wait(&empty_bar[stage], ((it / STAGES) - 1) & 1);
            const int k = act[it];
            const int kw = kmap ? kmap[k] : k;
            const uint32_t a_base = smem_base + stage * STAGE_BYTES, b_base = a_base + A_BYTES;
            // B: the C_out x C_in weight slice of offset kw = one TMA box per k-block (3-D map {ci, k, co}, 128B swizzle)
            if (tid == 0) {
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&full_bar[stage])), "r"((uint32_t)B_BYTES));
#pragma unroll
                for (int kb = 0; kb < NKB; ++kb)
                    asm volatile(
                        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                        ::"r"(b_base + kb * (COUT * 128)), "l"(reinterpret_cast<uint64_t>(&wmap)), "r"(kb * 32), "r"(kw), "r"(0),
                        "r"(smem_u32(&full_bar[stage])) : "memory");
            }
            // A: 128 rows x (NKB*8) chunks of 16 B; consecutive threads copy consecutive chunks of one row. ~85 % of the
            // (offset, row) slots have no neighbour: such a row is only re-zeroed if the stage's previous tenant left
            // data there (stage buffers start out zeroed), which cuts the copy instructions ~3.5x.
            const int k_prev = (it >= STAGES) ? act[it - STAGES] : -1;
            for (int q = tid; q < TILE_M * NKB * 8; q += PRODUCERS) {
                const int r = q / (NKB * 8), c16 = q - r * (NKB * 8);
                const int kb = c16 >> 3, c = c16 & 7;
                const int src = rows[k][r];
                const bool ok = src >= 0 && c16 < vec_per_row;
                if (!ok && (k_prev < 0 || rows[k_prev][r] < 0)) continue;   // already zero
                const float* g = ok ? feat + (size_t)src * cin + c16 * 4 : feat;
                cp_async16(a_base + kb * (TILE_M * 128) + (r >> 3) * 1024 + (r & 7) * 128 + ((c ^ (r & 7)) << 4), g, ok ? 16u : 0u);
            }
            cp_async_arrive(&full_bar[stage]);
        }
    } else if (lane == 0) {
        // ================================ MMA issuer ================================
        const uint32_t idesc = make_idesc_tf32(TILE_M, COUT);
        for (int it = 0; it < n_act; ++it) {
            const int stage = it % STAGES;
            mbar_wait(&full_bar[stage], (it / STAGES) & 1);
            asm volatile("fence.proxy.async.shared::cta;");   // cp.async (generic proxy) writes -> tensor-core (async proxy) reads
            asm volatile("tcgen05.fence::after_thread_sync;");
            const uint32_t a_base = smem_base + stage * STAGE_BYTES, b_base = a_base + A_BYTES;
#pragma unroll
            for (int kb = 0; kb < NKB; ++kb) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {  // 4 x (K = 8 floats = 32 bytes) per 128-byte swizzle row
                    const uint64_t ad = make_desc_sw128(a_base + kb * (TILE_M * 128) + j * 32);
                    const uint64_t bd = make_desc_sw128(b_base + kb * (COUT * 128) + j * 32);
                    umma_tf32(tmem_base, ad, bd, idesc, (it > 0 || kb > 0 || j > 0) ? 1u : 0u);
                }
            }
            umma_commit(&empty_bar[stage]);                   // frees the stage when these MMAs retire
            if (it == n_act - 1) umma_commit(&acc_bar);       // accumulator complete
        }
    }

    // ---- epilogue: TMEM -> registers -> global (warps 0..3 own TMEM lanes 32w..32w+31 = tile rows)
    if (warp < 4) {
        if (n_act > 0) {
            mbar_wait(&acc_bar, 0);
            asm volatile("tcgen05.fence::after_thread_sync;");
        }
        const int r = warp * 32 + lane;
        const int o = row0 + r;
#pragma unroll
        for (int c0 = 0; c0 < COUT; c0 += 32) {
            uint32_t v[32];
            if (n_act > 0) {
                const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + c0;
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                    "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                    "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                    : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                      "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                      "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                      "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                    : "r"(taddr));
                asm volatile("tcgen05.wait::ld.sync.aligned;");
            } else {
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = 0u;
            }
            if (o < n_out) {
                float* dst = out + (size_t)o * COUT + c0;
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                    if (c0 + j >= COUT) break;  // C_out = 16: only half of the 32 loaded columns exist
                    float4 w;
                    float* wp = reinterpret_cast<float*>(&w);
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        float x = __uint_as_float(v[j + u]);
                        const int c = c0 + j + u;
                        if (scale) x = fmaf(x, __ldg(&scale[c]), shift ? __ldg(&shift[c]) : 0.0f);
                        else if (shift) x += __ldg(&shift[c]);
                        if (relu) x = fmaxf(x, 0.0f);
                        wp[u] = x;
                    }
                    *reinterpret_cast<float4*>(dst + j) = w;
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 7) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS));
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// cuTensorMapEncodeTiled through the runtime's driver entry point (the library does not link libcuda)
EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess) return nullptr;
        fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

// 3-D view {ci, k, co} of the contiguous [C_out, K, C_in] weight; box {32, 1, cout}, 128-byte swizzle, zero OOB fill
int make_weight_map(const float* weight, int K, int cin, int cout, CUtensorMap* map) {
    EncodeTiledFn enc = get_encode_fn();
    if (!enc) return CRB3D_ERR_CUDA;
    cuuint64_t dims[3] = {(cuuint64_t)cin, (cuuint64_t)K, (cuuint64_t)cout};
    cuuint64_t strides[2] = {(cuuint64_t)cin * 4, (cuuint64_t)K * cin * 4};
    cuuint32_t box[3] = {32, 1, (cuuint32_t)cout};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(weight), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? CRB3D_OK : CRB3D_ERR_CUDA;
}

template <int NKB, int COUT, int STAGES, int MIN_CTAS>
int launch_tc(const float* feat, const int* nbr, const float* weight, int n_out, int K, int cin, const int* kmap,
              const float* scale, const float* shift, int relu, float* out, cudaStream_t stream) {
    constexpr size_t smem = (size_t)STAGES * (NKB * TILE_M * 128 + NKB * COUT * 128) + 1024;
    CUtensorMap wmap;
    int rc = make_weight_map(weight, K, cin, COUT, &wmap);
    if (rc) return rc;
    auto kern = spconv_fwd_tc<NKB, COUT, STAGES, MIN_CTAS>;
    static bool attr_set = false;  // per instantiation; the attribute is per function (and per device - single-GPU processes)
    if (!attr_set) {
        CRB3D_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set = true;
    }
    kern<<<(unsigned)crb3d_divup(n_out, TILE_M), THREADS, smem, stream>>>(feat, nbr, wmap, n_out, K, cin, kmap, scale, shift,
                                                                         relu, out);
    CRB3D_CHECK_LAUNCH();
    return CRB3D_OK;
}

}  // namespace

// TF32 tensor-core forward. weight must be contiguous [C_out, K, C_in] (for the input gradient pass the transposed
// weight [C_in, K, C_out] and the transposed table). Supported: C_in a multiple of 4 up to 64, C_out in
// {16, 32, 64, 128}, K <= 27; anything else returns CRB3D_ERR_UNSUPPORTED (callers use crb3d_spconv_forward_f32).
// Stage counts are sized so that two CTAs fit one SM (<= ~110 KB each) wherever possible: the second CTA hides the
// first one's gather latency and a 150-tile layer fits one wave of 296 CTA slots.
extern "C" int crb3d_spconv_forward_tf32(const float* feat, const int* nbr, const float* weight, int n_out, int K, int cin,
                                         int cout, const int* kmap, const float* scale, const float* shift, int relu,
                                         float* out, cudaStream_t stream) {
    if (n_out < 0 || K <= 0 || cin <= 0 || cout <= 0 || !weight || !out) return CRB3D_ERR_ARG;
    if (n_out == 0) return CRB3D_OK;
    if (!feat || !nbr) return CRB3D_ERR_ARG;
    if (K > MAX_K || (cin & 3)) return CRB3D_ERR_UNSUPPORTED;
#define TC_ARGS feat, nbr, weight, n_out, K, cin, kmap, scale, shift, relu, out, stream
    const int nkb = (cin + 31) / 32;
    if (nkb == 1) {                                        // stage = 16 KB + C_out*128 B
        if (cout == 16) return launch_tc<1, 16, 4, 2>(TC_ARGS);
        if (cout == 32) return launch_tc<1, 32, 4, 2>(TC_ARGS);
        if (cout == 64) return launch_tc<1, 64, 4, 2>(TC_ARGS);
        if (cout == 128) return launch_tc<1, 128, 3, 2>(TC_ARGS);
    } else if (nkb == 2) {                                 // stage = 32 KB + C_out*256 B
        if (cout == 16) return launch_tc<2, 16, 2, 2>(TC_ARGS);
        if (cout == 32) return launch_tc<2, 32, 2, 2>(TC_ARGS);
        if (cout == 64) return launch_tc<2, 64, 2, 2>(TC_ARGS);
        if (cout == 128) return launch_tc<2, 128, 3, 1>(TC_ARGS);
    }
#undef TC_ARGS
    return CRB3D_ERR_UNSUPPORTED;
}
