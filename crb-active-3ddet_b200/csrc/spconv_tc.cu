// Sparse 3D convolution on the 5th-gen tensor cores (tcgen05, TF32 inputs, fp32 accumulate in TMEM) fed by TMA.
//
// Replaces the external spconv-cu113==2.1.21 implicit-GEMM forward reached from
//   pcdet/models/backbones_3d/spconv_backbone.py:77-117   (C_in >= 16 layers of VoxelBackBone8x)
// Same contract as csrc/spconv_simt.cu:  out[o,:] = sum_k in[nbr[k][o],:] @ W[:,k,:]^T, W = [C_out, K, C_in].
//
// One CTA owns 128 consecutive output rows (the UMMA M). Six warps, mbarrier pipelined:
//   warp 0 (producer): for every kernel offset k that has a neighbour in the tile, lane g gathers tile rows 4g..4g+3
//       with ONE `cp.async.bulk.tensor.2d ... tile::gather4` per 32-channel block - the four row coordinates come
//       straight from the neighbour table, an absent neighbour (-1) is out of bounds and is zero-filled by the TMA
//       unit, the 128-byte swizzle of the UMMA K-major layout is applied by the hardware. The C_out x C_in weight
//       slice is one 3-D TMA box. A 4-row group that has no neighbour now and had none when the stage was last used is
//       skipped (the buffers start out zeroed), so ~85 % of the padded rows cost nothing.
//   warp 1 (one lane): waits full[stage], issues C_in/8 tcgen05.mma (M=128, N=C_out, K=8, kind::tf32) accumulating ALL
//       offsets into one TMEM tile, tcgen05.commit -> empty[stage] (and -> acc_full at the end).
//   warps 2-5 (epilogue): tcgen05.ld 32 lanes x 32 columns, optional scale/shift/ReLU, one store per output row.
// The output is written once, there are no atomics and the summation order is fixed (k ascending). No LSU gather at
// all: an earlier cp.async version spent ~3000 cycles per stage just ISSUING 16-byte copies (tools/trace_spconv.py), and
// two later LSU variants (one thread per half row; cooperative 2-8 rows per LDGSTS) measured 65 us / 89 us on conv3.1
// against 65 us for this TMA producer (profiles/r01_spconv_variants.txt) - the stage time is set by the latency of the
// scattered 256-byte row fetches, not by who issues them.
#include "common.cuh"
#include <cuda.h>

namespace {

constexpr int TILE_M = 128;
constexpr int PW = 4;            // producer warps (also the epilogue warps)
constexpr int THREADS = (PW + 1) * 32;
constexpr int MAX_K = 27;

// optional timeline trace of one CTA (clock64 stamps per offset: producer after empty-wait / after issuing its copies,
// MMA thread after full-wait / after issuing + committing); set through crb3d_debug_set_tc_trace, null in production
__device__ long long* g_tc_trace = nullptr;

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

// NKB = ceil(C_in / 32) k-blocks (one 128-byte swizzle row holds 32 floats; C_in = 16 is zero-padded by the TMA unit).
template <int NKB, int COUT, int STAGES, int MIN_CTAS>
__global__ void __launch_bounds__(THREADS, MIN_CTAS) spconv_fwd_tc(const __grid_constant__ CUtensorMap fmap,
                                                                   const __grid_constant__ CUtensorMap wmap,
                                                                   const int* __restrict__ nbr, int n_out, int K,
                                                                   const int* __restrict__ kmap, const float* __restrict__ scale,
                                                                   const float* __restrict__ shift, int relu,
                                                                   float* __restrict__ out, const int* __restrict__ n_dev) {
    // n_dev (nullable): device-side row count; n_out is then only the capacity / row stride of the neighbour table
    const int nv = n_dev ? min(n_out, *n_dev) : n_out;
    if ((int)blockIdx.x * TILE_M >= nv) return;   // uniform per CTA, before any barrier / TMEM allocation
    constexpr int A_BYTES = NKB * TILE_M * 128;
    constexpr int B_BYTES = NKB * COUT * 128;
    constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    constexpr int TMEM_COLS = COUT <= 32 ? 32 : (COUT <= 64 ? 64 : (COUT <= 128 ? 128 : 256));
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    __shared__ __align__(16) int rows[MAX_K][TILE_M];
    __shared__ int act[MAX_K];
    __shared__ int n_act_s;
    __shared__ uint64_t full_bar[STAGES];
    __shared__ uint64_t empty_bar[STAGES];
    __shared__ uint64_t acc_bar;
    __shared__ uint32_t tmem_base_s;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int row0 = blockIdx.x * TILE_M;

    // ---- setup: barriers, TMEM, zeroed stages, neighbour rows of this tile for every offset
    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full_bar[s], PW);  // one arrive.expect_tx per producer warp; the TMA unit completes the byte count
            mbar_init(&empty_bar[s], 1);  // tcgen05.commit
        }
        mbar_init(&acc_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    if (warp == PW) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "n"(TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    {
        float4* z = reinterpret_cast<float4*>(smem);
        for (int t = tid; t < STAGES * STAGE_BYTES / 16; t += THREADS) z[t] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    for (int t = tid; t < K * TILE_M; t += THREADS) {
        const int k = t / TILE_M, r = t - k * TILE_M;
        const int o = row0 + r;
        rows[k][r] = (o < nv) ? __ldg(&nbr[(size_t)k * n_out + o]) : -1;
    }
    asm volatile("fence.proxy.async.shared::cta;");  // the zero fill (generic proxy) precedes TMA writes / MMA reads
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem_base = tmem_base_s;
    // active offsets (warp 0 builds the compact list in ascending k: fixed summation order)
    if (warp == 0) {
        int n = 0;
        for (int k = 0; k < K; ++k) {
            const int4 v = reinterpret_cast<const int4*>(rows[k])[lane];
            const bool any = (v.x & v.y & v.z & v.w) >= 0;  // some entry is non-negative <=> the AND has a clear sign bit
            if (__any_sync(0xffffffffu, any)) { if (lane == 0) act[n] = k; ++n; }
        }
        if (lane == 0) n_act_s = n;
    }
    __syncthreads();
    const int n_act = n_act_s;
    const uint32_t smem_base = smem_u32(smem);

    if (warp < PW) {
        // ================================ producer warps: TMA gather4 + weight box ================================
        // warp w, lane l < 32/PW owns the 4-row group g = l*PW + w (the per-thread TMA issue cost is spread over PW warps)
        const int g = lane * PW + warp;
        const bool owner = lane < 32 / PW;
        for (int it = 0; it < n_act; ++it) {
            const int stage = it % STAGES;
            if (it >= STAGES) mbar_wait(&empty_bar[stage], ((it / STAGES) - 1) & 1);
            long long* trace = (g_tc_trace && blockIdx.x == gridDim.x / 2 && tid == 0) ? g_tc_trace + 128 : nullptr;
            if (trace) trace[it * 4 + 0] = clock64();
            const int k = act[it];
            const uint32_t a_base = smem_base + stage * STAGE_BYTES, b_base = a_base + A_BYTES;
            const uint32_t bar = smem_u32(&full_bar[stage]);
            int4 cur = make_int4(-1, -1, -1, -1);
            bool need = false;
            if (owner) {
                cur = reinterpret_cast<const int4*>(rows[k])[g];
                need = (cur.x & cur.y & cur.z & cur.w) >= 0;
                if (!need && it >= STAGES) {  // stale data from the stage's previous tenant must be cleared
                    const int4 prev = reinterpret_cast<const int4*>(rows[act[it - STAGES]])[g];
                    need = (prev.x & prev.y & prev.z & prev.w) >= 0;
                }
            }
            const unsigned int m = __ballot_sync(0xffffffffu, need);
            if (lane == 0) {
                uint32_t bytes = (uint32_t)__popc(m) * (uint32_t)(NKB * 512);
                if (warp == 0) bytes += (uint32_t)B_BYTES;
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes));
            }
            __syncwarp();
            if (warp == PW - 1 && lane == 31) {   // an otherwise idle lane fetches the weight slice
                const int kw = kmap ? kmap[k] : k;
#pragma unroll
                for (int kb = 0; kb < NKB; ++kb)
                    asm volatile(
                        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                        ::"r"(b_base + kb * (COUT * 128)), "l"(reinterpret_cast<uint64_t>(&wmap)), "r"(kb * 32), "r"(kw), "r"(0),
                        "r"(bar) : "memory");
            }
            if (need) {
#pragma unroll
                for (int kb = 0; kb < NKB; ++kb)
                    asm volatile(
                        "cp.async.bulk.tensor.2d.shared::cluster.global.tile::gather4.mbarrier::complete_tx::bytes "
                        "[%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
                        ::"r"(a_base + kb * (TILE_M * 128) + g * 512), "l"(reinterpret_cast<uint64_t>(&fmap)), "r"(kb * 32),
                        "r"(cur.x), "r"(cur.y), "r"(cur.z), "r"(cur.w), "r"(bar) : "memory");
            }
            if (trace) trace[it * 4 + 1] = clock64();
        }
    } else if (warp == PW) {
        // ================================ MMA issuer ================================
        // the whole warp runs the loop converged and only the tcgen05 instructions are predicated on one elected lane:
        // under `if (lane == 0)` every descriptor is a per-thread value that ptxas moves to the uniform register file
        // before each UTCHMMA (issue + commit of a stage: 800 -> 500 cycles in tools/trace_spconv.py)
        const uint32_t idesc = make_idesc_tf32(TILE_M, COUT);
        const uint64_t desc0 = make_desc_sw128(smem_base);
        for (int it = 0; it < n_act; ++it) {
            const int stage = it % STAGES;
            mbar_wait(&full_bar[stage], (it / STAGES) & 1);
            long long* trace = (g_tc_trace && blockIdx.x == gridDim.x / 2 && lane == 0) ? g_tc_trace : nullptr;
            if (trace) trace[it * 4 + 2] = clock64();
            asm volatile("tcgen05.fence::after_thread_sync;");
            const uint64_t da = desc0 + (uint64_t)((stage * STAGE_BYTES) >> 4), db = da + (uint64_t)(A_BYTES >> 4);
            uint32_t elected;
            asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(elected));
            if (elected) {
#pragma unroll
                for (int kb = 0; kb < NKB; ++kb) {
#pragma unroll
                    for (int j = 0; j < 4; ++j)   // 4 x (K = 8 floats = 32 bytes) per 128-byte swizzle row
                        umma_tf32(tmem_base, da + (uint64_t)((kb * (TILE_M * 128) + j * 32) >> 4),
                                  db + (uint64_t)((kb * (COUT * 128) + j * 32) >> 4), idesc, (it > 0 || kb > 0 || j > 0) ? 1u : 0u);
                }
                umma_commit(&empty_bar[stage]);                   // frees the stage when these MMAs retire
                if (it == n_act - 1) umma_commit(&acc_bar);       // accumulator complete
            }
            __syncwarp();
            if (trace) trace[it * 4 + 3] = clock64();
        }
    }

    // ---- epilogue: TMEM -> registers -> global (warps 0..3; a warp may only touch TMEM lanes 32*(warp%4)..+31)
    if (warp < 4) {
        if (n_act > 0) {
            mbar_wait(&acc_bar, 0);
            asm volatile("tcgen05.fence::after_thread_sync;");
        }
        const int quarter = warp & 3;
        const int r = quarter * 32 + lane;
        const int o = row0 + r;
#pragma unroll
        for (int c0 = 0; c0 < COUT; c0 += 32) {
            uint32_t v[32];
            if (n_act > 0) {
                const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + c0;
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
            if (o < nv) {
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
    if (warp == PW) {
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

// 2-D view {ci, row} of the (n_in, C_in) feature matrix for tile::gather4: box {32, 1} (four rows are named per
// instruction), 128-byte swizzle, out-of-bounds rows / columns read as zero (verified by tools/probe/gather4_probe.cu)
int make_feature_map(const float* feat, int n_in, int cin, CUtensorMap* map) {
    EncodeTiledFn enc = get_encode_fn();
    if (!enc) return CRB3D_ERR_CUDA;
    cuuint64_t dims[2] = {(cuuint64_t)cin, (cuuint64_t)n_in};
    cuuint64_t strides[1] = {(cuuint64_t)cin * 4};
    cuuint32_t box[2] = {32, 1};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(feat), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? CRB3D_OK : CRB3D_ERR_CUDA;
}

template <int NKB, int COUT, int STAGES, int MIN_CTAS>
int launch_tc(const float* feat, int n_in, const int* nbr, const float* weight, int n_out, int K, int cin, const int* kmap,
              const float* scale, const float* shift, int relu, float* out, const int* n_dev, cudaStream_t stream) {
    constexpr size_t smem = (size_t)STAGES * (NKB * TILE_M * 128 + NKB * COUT * 128) + 1024;
    CUtensorMap wmap, fmap;
    int rc = make_weight_map(weight, K, cin, COUT, &wmap);
    if (rc) return rc;
    rc = make_feature_map(feat, n_in, cin, &fmap);
    if (rc) return rc;
    auto kern = spconv_fwd_tc<NKB, COUT, STAGES, MIN_CTAS>;
    static bool attr_set = false;  // per instantiation; the attribute is per function (single-GPU processes)
    if (!attr_set) {
        CRB3D_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set = true;
    }
    kern<<<(unsigned)crb3d_divup(n_out, TILE_M), THREADS, smem, stream>>>(fmap, wmap, nbr, n_out, K, kmap, scale, shift, relu, out, n_dev);
    CRB3D_CHECK_LAUNCH();
    return CRB3D_OK;
}

}  // namespace

// TF32 tensor-core forward. feat: (n_in, C_in) contiguous; weight: contiguous [C_out, K, C_in] (for the input gradient
// pass the transposed weight [C_in, K, C_out] and the transposed table). Supported: C_in in {16, 32, 64}, C_out in
// {16, 32, 64, 128}, K <= 27; anything else returns CRB3D_ERR_UNSUPPORTED (callers use crb3d_spconv_forward_f32).
// Stage counts are sized so that two or three CTAs fit one SM: the co-resident CTAs hide each other's TMA round trips.
extern "C" int crb3d_spconv_forward_tf32(const float* feat, int n_in, const int* nbr, const float* weight, int n_out, int K,
                                         int cin, int cout, const int* kmap, const float* scale, const float* shift,
                                         int relu, float* out, const int* n_dev, cudaStream_t stream) {
    if (n_out < 0 || n_in < 0 || K <= 0 || cin <= 0 || cout <= 0 || !weight || !out) return CRB3D_ERR_ARG;
    if (n_out == 0) return CRB3D_OK;
    if (!feat || !nbr || n_in == 0) return CRB3D_ERR_ARG;
    if (K > MAX_K || (cin != 16 && cin != 32 && cin != 64)) return CRB3D_ERR_UNSUPPORTED;
#define TC_ARGS feat, n_in, nbr, weight, n_out, K, cin, kmap, scale, shift, relu, out, n_dev, stream
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

// Debug hook (not part of include/crb3d.h): buf = device array of >= 256 int64, or null to switch tracing off.
extern "C" int crb3d_debug_set_tc_trace(long long* buf) {
    return cudaMemcpyToSymbol(g_tc_trace, &buf, sizeof(buf)) == cudaSuccess ? CRB3D_OK : CRB3D_ERR_CUDA;
}
