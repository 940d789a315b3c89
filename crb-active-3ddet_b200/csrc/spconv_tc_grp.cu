// Sparse 3D convolution forward for the NARROWEST layers (C_in = 4, 8: the input layer of VoxelBackBone8x) on the 5th-gen tensor
// cores: several kernel offsets share one pipeline stage. (The kernel template takes C_in = 16 / 32 too; they are not dispatched,
// see the measurements at the bottom.)
//
// Same contract as csrc/spconv_tc.cu (spconv-cu113 implicit-GEMM forward behind pcdet/models/backbones_3d/spconv_backbone.py:77-117):
//   out[o,:] = sum_k in[nbr[k][o],:] @ W[:,k,:]^T,  W = [C_out, K, C_in] contiguous.
// There one stage = one offset: a C_in = 16 layer runs 27 barrier round trips of 1.1-2.5 k cycles per 128-row tile to move 64-byte
// rows, and the tile's latency - not bandwidth - sets the time (profiles/r02_spconv.txt). Here the K dimension of a stage is the
// CONCATENATION of G = 32*NKB/C_in offsets: the A row of output o is [in[nbr[g*G][o]] | in[nbr[g*G+1][o]] | ...] (a missing
// neighbour is a zero block) and the B operand is simply columns [g*G*C_in, (g+1)*G*C_in) of the weight read as the 2-D matrix
// [C_out][K*C_in] - it is K-major already, so one 2-D TMA box per 32-float k-block, zero-filled past K*C_in. C_in = 4 (the input
// layer: 27 x 4 = 108 floats) becomes ONE stage per tile, C_in = 8 four.
// Everything else follows spconv_tc.cu: 128 output rows per CTA, row owners append (row, slot, source) to shared-memory lists that
// all producer threads of the group walk with one 16-byte cp.async.ca per (entry, chunk) into the 128B-swizzled stage, two producer
// groups alternating over the stages (single-stage layers: both groups split the slots of the one stage), tcgen05.mma kind::tf32
// accumulating every stage in TMEM, one store per output row, fixed summation order, no atomics on the output.
#include "tc_common.cuh"

namespace {

using namespace tc;

constexpr int TILE_M = 128;
constexpr int PW = 8;                    // producer warps = 2 groups x 128 row owners
constexpr int THREADS = (PW + 1) * 32;
constexpr int MAX_K = 27;

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_arrive(uint64_t* bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

template <int CIN, int COUT, int NKB, int STAGES, int MIN_CTAS>
__global__ void __launch_bounds__(THREADS, MIN_CTAS) spconv_fwd_tc_grp(const float* __restrict__ feat, const __grid_constant__ CUtensorMap wmap,
                                                                       const int* __restrict__ nbr, int n_out, int K,
                                                                       const float* __restrict__ scale, const float* __restrict__ shift,
                                                                       int relu, float* __restrict__ out, const int* __restrict__ n_dev) {
    constexpr int G = 32 * NKB / CIN;                // offsets per stage
    constexpr int CPR = CIN / 4;                     // 16-byte chunks per (row, slot)
    constexpr int CSH = CPR == 1 ? 0 : (CPR == 2 ? 1 : (CPR == 4 ? 2 : 3));
    constexpr bool SPLIT = STAGES == 1;              // one stage per tile: both producer groups work on it, slots s = grp, grp+2, ..
    constexpr int SPT = SPLIT ? (G + 1) / 2 : G;     // slots per owner thread and stage
    constexpr int NBUF = SPLIT ? 1 : 2;
    constexpr int LIST_CAP = TILE_M * SPT;
    constexpr int A_BYTES = NKB * TILE_M * 128, B_BYTES = NKB * COUT * 128, STAGE_BYTES = A_BYTES + B_BYTES;
    constexpr int TMEM_COLS = COUT <= 32 ? 32 : (COUT <= 64 ? 64 : (COUT <= 128 ? 128 : 256));
    static_assert(G >= 1 && G <= 32 && (SPLIT || STAGES % 2 == 0), "stage geometry");
    static_assert(SPLIT || STAGES * G <= 32, "dirty bits live in one 32-bit register");

    const int nv = n_dev ? min(n_out, *n_dev) : n_out;
    if ((int)blockIdx.x * TILE_M >= nv) return;     // uniform per CTA, before any barrier / TMEM allocation
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // pointer arithmetic keeps the shared address space (LDS/STS, not generic LD/ST)
    __shared__ int act[MAX_K];                       // active STAGE ids, ascending
    __shared__ int act_flag[MAX_K + 5];              // per offset: some row of the tile has this neighbour
    __shared__ int list_src[2][NBUF][LIST_CAP];
    __shared__ unsigned short list_pos[2][NBUF][LIST_CAP];          // row << 5 | slot
    __shared__ unsigned short list_z[2][NBUF][SPLIT ? 1 : LIST_CAP];
    __shared__ int cnt_v[2][4], cnt_z[2][4];
    __shared__ int n_act_s;
    __shared__ __align__(16) float scale_s[128], shift_s[128];   // BatchNorm affine of the epilogue (identity where absent)
    __shared__ uint64_t full_bar[STAGES], empty_bar[STAGES], acc_bar;
    __shared__ uint32_t tmem_base_s;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int row0 = blockIdx.x * TILE_M;
    const int n_stage_total = (K + G - 1) / G;

    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full_bar[s], (SPLIT ? 2 : 1) * TILE_M + 1);   // one arrival per producer thread + the weight boxes' expect_tx
            mbar_init(&empty_bar[s], 1);
        }
        mbar_init(&acc_bar, 1);
        mbar_fence_init();
        tma_prefetch_desc(&wmap);
    }
    if (warp == PW) tmem_alloc<TMEM_COLS>(&tmem_base_s);
    {
        float4* z = reinterpret_cast<float4*>(smem);
        for (int t = tid; t < STAGES * STAGE_BYTES / 16; t += THREADS) z[t] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    if (tid < 5) act_flag[MAX_K + tid] = 0;
    if (warp < PW) {   // warp w scans offsets w, w + PW, ...: 128 table cells = one int4 per lane
        for (int k = warp; k < MAX_K; k += PW) {
            bool any = false;
            if (k < K) {
                const int o = row0 + lane * 4;
                int4 v = make_int4(-1, -1, -1, -1);
                if (o + 3 < nv && ((((size_t)k * n_out + o) & 3) == 0)) v = __ldg(reinterpret_cast<const int4*>(nbr + (size_t)k * n_out + o));
                else {
                    if (o < nv) v.x = __ldg(&nbr[(size_t)k * n_out + o]);
                    if (o + 1 < nv) v.y = __ldg(&nbr[(size_t)k * n_out + o + 1]);
                    if (o + 2 < nv) v.z = __ldg(&nbr[(size_t)k * n_out + o + 2]);
                    if (o + 3 < nv) v.w = __ldg(&nbr[(size_t)k * n_out + o + 3]);
                }
                any = (v.x & v.y & v.z & v.w) >= 0;
            }
            const bool warp_any = __any_sync(0xffffffffu, any);
            if (lane == 0) act_flag[k] = warp_any ? 1 : 0;
        }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before();
    for (int c = threadIdx.x; c < COUT; c += blockDim.x) { scale_s[c] = scale ? __ldg(&scale[c]) : 1.0f; shift_s[c] = shift ? __ldg(&shift[c]) : 0.0f; }
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_s;
    if (warp == 0) {   // active stages in ascending order (fixed summation order): one ballot
        for (int u = lane; u < 8; u += 32) { cnt_v[u >> 2][u & 3] = 0; cnt_z[u >> 2][u & 3] = 0; }
        bool f = false;
        if (lane < n_stage_total)
            for (int s = 0; s < G; ++s) f = f || act_flag[min(lane * G + s, MAX_K + 4)] != 0;
        const unsigned int m = __ballot_sync(0xffffffffu, f);
        if (f) act[__popc(m & ((1u << lane) - 1u))] = lane;
        if (lane == 0) n_act_s = __popc(m);
    }
    __syncthreads();
    const int n_act = n_act_s;
    const uint32_t smem_base = smem_u32(smem);

    if (warp < PW) {
        // ================================ producers ================================
        constexpr int GT = TILE_M;
        const int grp = warp >> 2, gtid = tid & (GT - 1);
        const int r = gtid, o = row0 + r;
        uint32_t dirty = 0u;
        int src_next[SPT];
        // neighbour indices of this owner's slots of stage-list position `pos`
        auto load_srcs = [&](int pos, int (&dst)[SPT]) {
            const int sg = act[pos];
#pragma unroll
            for (int j = 0; j < SPT; ++j) {
                const int s = SPLIT ? grp + 2 * j : j;
                const int k = sg * G + s;
                dst[j] = (s < G && k < K && act_flag[k] && o < nv) ? __ldg(&nbr[(size_t)k * n_out + o]) : -1;
            }
        };
        const int it0 = SPLIT ? 0 : grp, it_step = SPLIT ? 1 : 2;
        if (it0 < n_act) load_srcs(it0, src_next);
        for (int it = it0, li = 0; it < n_act; it += it_step, ++li) {
            const int stage = it % STAGES, lb = SPLIT ? 0 : (li & 1);
            int src[SPT];
#pragma unroll
            for (int j = 0; j < SPT; ++j) src[j] = src_next[j];
            if (it + it_step < n_act) load_srcs(it + it_step, src_next);
            if (gtid == 0) { cnt_v[grp][(li + 2) & 3] = 0; cnt_z[grp][(li + 2) & 3] = 0; }
            if (it >= STAGES) mbar_wait(&empty_bar[stage], ((it / STAGES) - 1) & 1, (CRB3D_K_SPCONV_TC << 8) | 9, it);
            const uint32_t a_base = smem_base + stage * STAGE_BYTES, b_base = a_base + A_BYTES;
            if (gtid == GT - 1 && (!SPLIT || grp == 0)) {   // the weight columns of this stage: one 2-D TMA box per k-block
                mbar_expect_tx(&full_bar[stage], B_BYTES);
#pragma unroll
                for (int kb = 0; kb < NKB; ++kb)
                    tma_load_2d(b_base + kb * (COUT * 128), &wmap, (act[it] * NKB + kb) * 32, 0, &full_bar[stage]);
            }
            // ---- list entries of this row: valid (row, slot, source) and stale (row, slot) - one packed warp scan + one atomic
            unsigned int vm = 0u, zm = 0u;
#pragma unroll
            for (int j = 0; j < SPT; ++j) {
                const bool valid = src[j] >= 0;
                const uint32_t bit = 1u << (SPLIT ? j : stage * G + j);
                if (valid) vm |= 1u << j;
                if (!SPLIT) {
                    if (!valid && (dirty & bit)) zm |= 1u << j;
                    if (valid) dirty |= bit; else dirty &= ~bit;
                }
            }
            int packed = __popc(vm) | (__popc(zm) << 16), incl = packed;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int n = __shfl_up_sync(0xffffffffu, incl, d);
                if (lane >= d) incl += n;
            }
            int base = 0;
            if (lane == 31) {
                const int tv = incl & 0xFFFF, tz = incl >> 16;
                const int bv = tv ? atomicAdd(&cnt_v[grp][li & 3], tv) : 0;
                const int bz = tz ? atomicAdd(&cnt_z[grp][li & 3], tz) : 0;
                base = bv | (bz << 16);
            }
            base = __shfl_sync(0xffffffffu, base, 31);
            int iv = (base & 0xFFFF) + ((incl - packed) & 0xFFFF), iz = (base >> 16) + ((incl - packed) >> 16);
#pragma unroll
            for (int j = 0; j < SPT; ++j) {
                const int s = SPLIT ? grp + 2 * j : j;
                if (vm & (1u << j)) {
                    list_pos[grp][lb][iv] = (unsigned short)((r << 5) | s);
                    list_src[grp][lb][iv] = src[j];
                    ++iv;
                }
                if (!SPLIT && (zm & (1u << j))) list_z[grp][lb][iz++] = (unsigned short)((r << 5) | s);
            }
            asm volatile("bar.sync %0, %1;" ::"r"(1 + grp), "n"(GT) : "memory");   // this group's lists are complete
            const int n_v = cnt_v[grp][li & 3] << CSH, n_z = SPLIT ? 0 : (cnt_z[grp][li & 3] << CSH);
            for (int i = gtid; i < n_v; i += GT) {
                const int e = i >> CSH, chunk = i & (CPR - 1);
                const unsigned int p = list_pos[grp][lb][e];
                const int row = (int)(p >> 5), col16 = (int)(p & 31u) * CPR + chunk;
                cp_async16(a_base + (col16 >> 3) * (TILE_M * 128) + (row >> 3) * 1024 + (row & 7) * 128 + (((col16 & 7) ^ (row & 7)) << 4),
                           feat + (size_t)list_src[grp][lb][e] * CIN + chunk * 4);
            }
            if (n_z > 0) {
                for (int i = gtid; i < n_z; i += GT) {
                    const int e = i >> CSH, chunk = i & (CPR - 1);
                    const unsigned int p = list_z[grp][lb][e];
                    const int row = (int)(p >> 5), col16 = (int)(p & 31u) * CPR + chunk;
                    asm volatile("st.shared.v4.f32 [%0], {%1, %1, %1, %1};" ::"r"(a_base + (col16 >> 3) * (TILE_M * 128) + (row >> 3) * 1024 + (row & 7) * 128 + (((col16 & 7) ^ (row & 7)) << 4)), "f"(0.0f) : "memory");
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            }
            cp_async_arrive(&full_bar[stage]);
        }
    } else if (warp == PW) {
        // ================================ MMA issuer (converged warp, tcgen05 predicated on one elected lane) ================
        const uint32_t idesc = idesc_tf32(TILE_M, COUT);
        const uint64_t desc0 = desc_sw128(smem_base);
        for (int it = 0; it < n_act; ++it) {
            const int stage = it % STAGES;
            mbar_wait(&full_bar[stage], (it / STAGES) & 1, (CRB3D_K_SPCONV_TC << 8) | 8, it);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // cp.async (generic proxy) writes -> tensor-core reads
            tc_fence_after();
            const uint64_t da = desc0 + (uint64_t)((stage * STAGE_BYTES) >> 4), db = da + (uint64_t)(A_BYTES >> 4);
            if (elect_one()) {
#pragma unroll
                for (int kb = 0; kb < NKB; ++kb) {
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        umma_tf32(tmem_base, da + (uint64_t)((kb * (TILE_M * 128) + j * 32) >> 4),
                                  db + (uint64_t)((kb * (COUT * 128) + j * 32) >> 4), idesc, (it > 0 || kb > 0 || j > 0) ? 1u : 0u);
                }
                umma_commit(&empty_bar[stage]);
                if (it == n_act - 1) umma_commit(&acc_bar);
            }
            __syncwarp();
        }
    }

    // ---- epilogue: TMEM -> registers -> global (warps 0..3; a warp may only touch TMEM lanes 32*(warp%4)..+31)
    if (warp < 4) {
        if (n_act > 0) {
            mbar_wait(&acc_bar, 0, (CRB3D_K_SPCONV_TC << 8) | 7);
            tc_fence_after();
        }
        const int quarter = warp & 3;
        const int o = row0 + quarter * 32 + lane;
#pragma unroll
        for (int c0 = 0; c0 < COUT; c0 += 32) {
            uint32_t v[32];
            if (n_act > 0) tmem_ld32(tmem_base + ((uint32_t)(quarter * 32) << 16) + c0, v);
            else {
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = 0u;
            }
            if (o < nv) {
                float* dst = out + (size_t)o * COUT + c0;
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                    if (c0 + j >= COUT) break;
                    // affine from shared memory (as in spconv_tc.cu); fmaf(x, 1, shift) and fmaf(x, scale, 0) are exact
                    const float4 sc = *reinterpret_cast<const float4*>(scale_s + c0 + j), sh = *reinterpret_cast<const float4*>(shift_s + c0 + j);
                    float4 w = make_float4(fmaf(__uint_as_float(v[j]), sc.x, sh.x), fmaf(__uint_as_float(v[j + 1]), sc.y, sh.y),
                                           fmaf(__uint_as_float(v[j + 2]), sc.z, sh.z), fmaf(__uint_as_float(v[j + 3]), sc.w, sh.w));
                    if (relu & 1) { w.x = fmaxf(w.x, 0.0f); w.y = fmaxf(w.y, 0.0f); w.z = fmaxf(w.z, 0.0f); w.w = fmaxf(w.w, 0.0f); }
                    if (relu & 2) { w.x = tf32_rn(w.x); w.y = tf32_rn(w.y); w.z = tf32_rn(w.z); w.w = tf32_rn(w.w); }
                    *reinterpret_cast<float4*>(dst + j) = w;
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == PW) tmem_dealloc<TMEM_COLS>(tmem_base);
}

template <int CIN, int COUT, int NKB, int STAGES, int MIN_CTAS>
int launch_grp(const float* feat, const int* nbr, const float* weight, int n_out, int K, const float* scale, const float* shift, int relu,
               float* out, const int* n_dev, cudaStream_t stream) {
    constexpr size_t smem = (size_t)STAGES * (NKB * TILE_M * 128 + NKB * COUT * 128) + 1024;
    CUtensorMap wmap;
    {   // the weight as the 2-D K-major matrix [C_out][K*C_in]; box = 32 floats x C_out rows, zero fill past K*C_in
        const uint64_t dims[2] = {(uint64_t)K * CIN, (uint64_t)COUT}, strides[1] = {(uint64_t)K * CIN * 4};
        const uint32_t box[2] = {32, (uint32_t)COUT};
        int rc = make_map_f32(&wmap, weight, 2, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
        if (rc) return rc;
    }
    auto kern = spconv_fwd_tc_grp<CIN, COUT, NKB, STAGES, MIN_CTAS>;
    static bool attr_set[CRB3D_MAX_DEVICES] = {};
    const int dev = crb3d_current_device();
    if (!attr_set[dev]) {
        CRB3D_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set[dev] = true;
    }
    kern<<<(unsigned)crb3d_divup(n_out, TILE_M), THREADS, smem, stream>>>(feat, wmap, nbr, n_out, K, scale, shift, relu, out, n_dev);
    CRB3D_CHECK_LAUNCH();
    return CRB3D_OK;
}

}  // namespace

CRB3D_DIAG_DEFINE_SETTER(spconv_grp)

// Forward of a narrow layer with grouped stages (no kmap: the forward direction only). Returns CRB3D_ERR_UNSUPPORTED for shapes it
// does not take; crb3d_spconv_forward_tf32 then uses the one-offset-per-stage kernel. Not part of include/crb3d.h.
int crb3d_spconv_forward_tf32_grouped(const float* feat, const int* nbr, const float* weight, int n_out, int K, int cin, int cout,
                                      const float* scale, const float* shift, int relu, float* out, const int* n_dev,
                                      cudaStream_t stream) {
    if (K > MAX_K || (K * cin) % 4 != 0) return CRB3D_ERR_UNSUPPORTED;
#define GRP_ARGS feat, nbr, weight, n_out, K, scale, shift, relu, out, n_dev, stream
#define GRP_COUTS(CIN, NKB, STAGES, M16, M32, M64, M128)                                   \
    if (cout == 16) return launch_grp<CIN, 16, NKB, STAGES, M16>(GRP_ARGS);                \
    if (cout == 32) return launch_grp<CIN, 32, NKB, STAGES, M32>(GRP_ARGS);                \
    if (cout == 64) return launch_grp<CIN, 64, NKB, STAGES, M64>(GRP_ARGS);                \
    if (cout == 128) return launch_grp<CIN, 128, NKB, STAGES, M128>(GRP_ARGS);
    // Measured at batch 16 (tools/bench_spconv.py, profiles/r02_spconv.txt): the input layer (C_in = 4) 77.8 -> 60.3 us; C_in = 16
    // (4 offsets per stage) 82.8 -> 79.8 / 88.9 -> 92.9 us and C_in = 32 (2 per stage) 112 -> 126 / 80 -> 92 us: with thousands of
    // tiles per launch those layers are bound by the LDGSTS rate, not by the per-stage round trips, and the larger stages cost a
    // resident CTA. Only the layers whose rows are a single 16/32-byte chunk are routed here.
    if (cin == 4) { GRP_COUTS(4, 4, 1, 2, 2, 2, 1) }          // 27 offsets x 4 = 108 floats: ONE stage (64 KB + weights)
    if (cin == 8) { GRP_COUTS(8, 2, 2, 2, 2, 2, 1) }          // 8 offsets per stage
#undef GRP_COUTS
#undef GRP_ARGS
    return CRB3D_ERR_UNSUPPORTED;
}
