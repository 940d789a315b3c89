// CTA-pair variant of the BEV GEMM (csrc/bev_gemm_tc.cu) for the long-K deblock: ConvTranspose2d(256 -> 256, k = s = 2) + BN + ReLU
// of pcdet/models/backbones_2d/base_bev_backbone.py:60-78 as four GEMMs D[m, n] = sum_k A[m, k] W[sub][n, k] whose rows land on the
// interleaved output pixels (2y+dy, 2x+dx) of the channel slice of the concatenated map -
// and, in CONV mode, for the stride-2 3x3 conv that opens BEV block 2 (base_bev_backbone.py:33-40) as an implicit GEMM.
//
// Why a second kernel: with K = 256 the single-CTA kernel keeps a 128-column weight block resident (128 KB), which leaves room for
// 3 activation stages - the ring, not L2 bytes or the tensor pipe, sets its rate (274 us at 16 x 100 x 88, 270 TFLOP/s) - and it
// reads every activation tile once per 128-column block (8 times). Here two CTAs of a cluster issue ONE M256 x N256 x K8
// tcgen05.mma.cta_group::2: each CTA holds its own 128 rows of A and HALF of the sub-position's weight rows (128 of 256), so all
// 256 output columns come from one pass over the activations (4 passes instead of 8). The weight half either stays resident
// (128 KB, 4 activation stages of 16 KB: 184 us) or - the default - streams with the activations through 6 stages of 32 KB
// (169 us: twice the L2 -> SM bytes, but 1.5x the latency cover, and the ring is what bounds the kernel). A 16-column epilogue
// staging (20 KB instead of 36) pays for the extra stage.
//   warp 0 (both CTAs) : TMA producer - per k-block one 128 x 32 box of A (2-D, or the 4-D strided box of a conv tap) and this CTA's
//                        half of the weight k-block (.cta_group::2 loads complete on the LEADER's barriers)
//   warp 1 (leader)    : MMA issuer, two TMEM accumulators of 256 columns; commits are multicast to both CTAs
//   warps 2-9 (both)   : epilogue of the other accumulator (tcgen05.ld, bias / ReLU / TF32 rounding, rows staged in shared memory,
//                        64-byte row pieces out); they hand the accumulator back on the leader's barrier
// The pair allocation of tensor memory comes AFTER a cluster barrier (root cause of the round-1 hang).
#include "tc_common.cuh"

namespace {

using namespace tc;

constexpr int TILE_M = 128;
constexpr int BK = 32;
constexpr int NP = 256;                         // output columns per pair (one sub-position)
constexpr int NH = NP / 2;                      // weight rows held by each CTA
constexpr int A_BYTES = TILE_M * 128;
constexpr int B_BYTES = NH * 128;               // this CTA's half of one weight k-block
constexpr int CH = 16;                          // columns per epilogue pass
constexpr int PITCH = CH + 4;
constexpr int EPW = 8;
constexpr int STAGING = EPW * 32 * PITCH * 4;   // 20480 B
constexpr int MAX_STAGES = 8;
constexpr int NTHREADS = 32 * (2 + EPW);

__device__ __forceinline__ uint32_t cluster_rank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ uint32_t leader_addr(const void* local) {        // the same variable in CTA 0 of the cluster
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, 0;" : "=r"(r) : "r"(smem_u32(local)));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// Default semantics (release at CTA scope), as CUTLASS's ClusterBarrier::arrive(cta_id): what is handed over is tensor-memory state,
// ordered by the tcgen05 fences on both sides. `.release.cluster` makes every epilogue warp drain its global stores first
// (MEMBAR.ALL.CTA + ERRBAR in SASS: 12 % of this kernel's stall samples).
__device__ __forceinline__ void remote_arrive(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void tma2_load_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(bar) : "memory");
}
__device__ __forceinline__ void tma2_load_4d(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2, int c3, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
                 ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(bar) : "memory");
}
__device__ __forceinline__ void umma2_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc),
        "r"(accumulate));
}
__device__ __forceinline__ void umma2_commit(uint64_t* bar) {               // arrives on `bar` of BOTH CTAs
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}

__device__ __forceinline__ void tmem_ld16_issue(uint32_t taddr, uint32_t* v) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

struct PairOut {
    float* ptr;               // first column of the segment in the output map
    long long row_stride;     // floats between consecutive output rows
    int up;                   // 0: output row = GEMM row; 2: row (b, y, x) -> (b, 2y+dy, 2x+dx), (dy, dx) = sub-position
    int in_w, in_h;
};

// CONV mode (as in bev_gemm_tc.cu): the GEMM rows are the output pixels of a k x k convolution (stride s, zero padding p) over a
// channels-last map, a 128-row tile = 8 x 16 output pixels, k-block kb = (tap, 32-channel block): its A operand is ONE 4-D TMA box
// {32 channels, 16 x, 8 y, 1 image} whose traversal stride along x and y is the conv stride and whose out-of-bounds pixels are
// zero-filled by the TMA unit (the padding). Weights [256][tap][C_in] stream with the activations; M counts tiles.
struct PairConv {
    int on, tiles_x, tiles_y, h_out, w_out, stride, pad, cblocks, ksize, n_img;
};

// gridDim.x = 2 * n_sub * clusters_per_slice. A cluster serves one sub-position and walks the 256-row tile pairs
// p = crank, crank + clusters_per_slice, ...; CTA `rank` of the pair owns rows p * 256 + rank * 128 ...
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NTHREADS, 1)
bev_gemm_pair_tc(const __grid_constant__ CUtensorMap amap, const __grid_constant__ CUtensorMap wmap, int M, int K, int stages,
                 int bres, int clusters_per_slice, const float* __restrict__ bias, int relu, const __grid_constant__ PairOut out,
                 const __grid_constant__ PairConv cv) {
    extern __shared__ uint8_t smem_raw[];
    // the dynamic window starts at the same offset in both CTAs; descriptors address both CTAs with one offset
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // pointer arithmetic keeps the shared address space
    __shared__ uint64_t full_bar[MAX_STAGES], empty_bar[MAX_STAGES], acc_full[2], acc_empty[2], b_full;
    __shared__ uint32_t tmem_base_s;
    __shared__ long long rowoff_s[EPW][32];
    __shared__ __align__(16) float bias_s[NP];

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t rank = cluster_rank();
    const bool leader = rank == 0;
    const int cluster_id = blockIdx.x >> 1;
    const int sub = cluster_id / clusters_per_slice, crank = cluster_id - sub * clusters_per_slice;
    const int nkb = K / BK;
    const int n_pairs = cv.on ? (M + 1) / 2 : (M + 2 * TILE_M - 1) / (2 * TILE_M);
    const int my_tiles = crank < n_pairs ? (n_pairs - crank + clusters_per_slice - 1) / clusters_per_slice : 0;

    if (tid == 0) {
        for (int s = 0; s < MAX_STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(&acc_full[s], 1); mbar_init(&acc_empty[s], 2 * EPW); }
        mbar_init(&b_full, 1);
        mbar_fence_init();
        tma_prefetch_desc(&amap);
        tma_prefetch_desc(&wmap);
    }
    cluster_sync_all();                          // both CTAs are running before the pair allocation touches the peer SM's tensor memory
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "n"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
    }
    for (int i = tid; i < NP; i += NTHREADS) bias_s[i] = bias ? __ldg(&bias[i]) : 0.0f;
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                          // the peer's barriers are initialised before anything signals them
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_s;
    // layout: [staging][resident weight half][stage ring]
    const uint32_t smem_base = smem_u32(smem);
    const uint32_t bres_base = smem_base + STAGING;
    const uint32_t ring_base = bres_base + (bres ? (uint32_t)nkb * B_BYTES : 0u);
    const uint32_t stage_bytes = bres ? A_BYTES : A_BYTES + B_BYTES;   // streamed weights ride with the activations

    if (warp == 0) {
        if (lane == 0 && my_tiles > 0) {
            if (bres) {
                if (leader) mbar_expect_tx(&b_full, 2u * (uint32_t)nkb * B_BYTES);
                const uint32_t bbar = leader_addr(&b_full);
                for (int kb = 0; kb < nkb; ++kb) tma2_load_2d(bres_base + kb * B_BYTES, &wmap, kb * BK, sub * NP + (int)rank * NH, bbar);
            }
            int it = 0;
            for (int i = 0; i < my_tiles; ++i) {
                const int m0 = (crank + i * clusters_per_slice) * 2 * TILE_M + (int)rank * TILE_M;   // rows beyond M: the TMA unit writes zeros
                int img = 0, cx0 = 0, cy0 = 0;
                if (cv.on) {                              // tile index beyond the last image: every pixel out of bounds, zeros
                    const int t = (crank + i * clusters_per_slice) * 2 + (int)rank, per_img = cv.tiles_x * cv.tiles_y;
                    img = t / per_img;
                    const int rem = t - img * per_img;
                    cy0 = (rem / cv.tiles_x) * 8 * cv.stride - cv.pad;
                    cx0 = (rem % cv.tiles_x) * 16 * cv.stride - cv.pad;
                }
                for (int kb = 0; kb < nkb; ++kb, ++it) {
                    const int stage = it % stages;
                    if (it >= stages) mbar_wait(&empty_bar[stage], ((it / stages) - 1) & 1, (CRB3D_K_BEV_GEMM << 8) | 21, it);
                    if (leader) mbar_expect_tx(&full_bar[stage], 2 * stage_bytes);
                    const uint32_t fbar = leader_addr(&full_bar[stage]);
                    if (cv.on) {
                        const int tap = kb / cv.cblocks, cb = kb - tap * cv.cblocks;
                        tma2_load_4d(ring_base + stage * stage_bytes, &amap, cb * BK, cx0 + tap % cv.ksize, cy0 + tap / cv.ksize, img, fbar);
                    } else
                        tma2_load_2d(ring_base + stage * stage_bytes, &amap, kb * BK, m0, fbar);
                    if (!bres) tma2_load_2d(ring_base + stage * stage_bytes + A_BYTES, &wmap, kb * BK, sub * NP + (int)rank * NH, fbar);
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        if (leader && my_tiles > 0) {
            const uint32_t idesc = idesc_tf32(2 * TILE_M, NP);
            if (bres) mbar_wait(&b_full, 0, (CRB3D_K_BEV_GEMM << 8) | 22);
            int it = 0;
            for (int i = 0; i < my_tiles; ++i) {
                const int acc = i & 1;
                if (i >= 2) mbar_wait(&acc_empty[acc], ((i >> 1) - 1) & 1, (CRB3D_K_BEV_GEMM << 8) | 23, i);   // both CTAs have drained it
                tc_fence_after();
                for (int kb = 0; kb < nkb; ++kb, ++it) {
                    const int stage = it % stages;
                    mbar_wait(&full_bar[stage], (it / stages) & 1, (CRB3D_K_BEV_GEMM << 8) | 24, it);
                    tc_fence_after();
                    if (elect_one()) {
                        const uint32_t a_base = ring_base + stage * stage_bytes;
                        const uint32_t b_base = bres ? bres_base + kb * B_BYTES : a_base + A_BYTES;
#pragma unroll
                        for (int j = 0; j < 4; ++j)
                            umma2_tf32(tmem_base + acc * NP, desc_sw128(a_base + j * 32), desc_sw128(b_base + j * 32), idesc,
                                       (kb > 0 || j > 0) ? 1u : 0u);
                        umma2_commit(&empty_bar[stage]);
                        if (kb == nkb - 1) umma2_commit(&acc_full[acc]);
                    }
                    __syncwarp();
                }
            }
        }
    } else {
        // ================================ epilogue: 8 warps, two per TMEM lane quarter ================================
        const int q = warp & 3, h = (warp - 2) >> 2;      // quarter (TMEM lanes 32q..32q+31), which half of every 32-column group
        const int r = q * 32 + lane;                      // tile row owned by this thread in the TMEM read
        float* stage_w = reinterpret_cast<float*>(smem) + (size_t)(warp - 2) * 32 * PITCH;
        long long* rowoff = rowoff_s[warp - 2];
        const uint32_t acc_empty_leader[2] = {leader_addr(&acc_empty[0]), leader_addr(&acc_empty[1])};
        for (int i = 0; i < my_tiles; ++i) {
            const int acc = i & 1;
            const long long m = (long long)(crank + i * clusters_per_slice) * 2 * TILE_M + (long long)rank * TILE_M + r;
            long long orow = -1;                          // output row of this lane's tile row; -1 = beyond M
            if (cv.on) {                                  // tile row r = pixel (y0 + r / 16, x0 + r % 16) of image b
                const int t = (crank + i * clusters_per_slice) * 2 + (int)rank, per_img = cv.tiles_x * cv.tiles_y;
                const int b = t / per_img, rem = t - b * per_img;
                const int y = (rem / cv.tiles_x) * 8 + (r >> 4), x = (rem % cv.tiles_x) * 16 + (r & 15);
                if (b < cv.n_img && y < cv.h_out && x < cv.w_out) orow = ((long long)b * cv.h_out + y) * cv.w_out + x;
            } else if (m < M) {
                if (out.up == 2) {
                    const int hw = out.in_w * out.in_h;
                    const int b = (int)(m / hw), rem = (int)(m - (long long)b * hw);
                    const int y = rem / out.in_w, x = rem - y * out.in_w;
                    orow = ((long long)b * (2 * out.in_h) + 2 * y + (sub >> 1)) * (2 * out.in_w) + 2 * x + (sub & 1);
                } else orow = m;
            }
            __syncwarp();
            rowoff[lane] = orow < 0 ? -1 : orow * out.row_stride;
            mbar_wait(&acc_full[acc], (i >> 1) & 1, (CRB3D_K_BEV_GEMM << 8) | 25, i);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * NP + h * CH);
            auto emit = [&](const uint32_t* v, int c0) {      // 32 rows x 16 columns: bias / ReLU / rounding, staged, 64-byte row pieces out
                const float4* b4 = reinterpret_cast<const float4*>(bias_s + c0);
#pragma unroll
                for (int j = 0; j < CH; j += 4) {
                    const float4 bq = b4[j >> 2];
                    float4 w = make_float4(__uint_as_float(v[j]) + bq.x, __uint_as_float(v[j + 1]) + bq.y,
                                           __uint_as_float(v[j + 2]) + bq.z, __uint_as_float(v[j + 3]) + bq.w);
                    if (relu & 1) { w.x = fmaxf(w.x, 0.0f); w.y = fmaxf(w.y, 0.0f); w.z = fmaxf(w.z, 0.0f); w.w = fmaxf(w.w, 0.0f); }
                    if (relu & 2) { w.x = tf32_rn(w.x); w.y = tf32_rn(w.y); w.z = tf32_rn(w.z); w.w = tf32_rn(w.w); }
                    *reinterpret_cast<float4*>(stage_w + (size_t)lane * PITCH + j) = w;
                }
                __syncwarp();
                float4 val[4];
                long long off[4];
#pragma unroll
                for (int t = 0; t < 4; ++t) {         // 8 rows x 64 bytes per warp instruction
                    const int rr = t * 8 + (lane >> 2);
                    off[t] = rowoff[rr];
                    val[t] = *reinterpret_cast<const float4*>(stage_w + (size_t)rr * PITCH + (lane & 3) * 4);
                }
#pragma unroll
                for (int t = 0; t < 4; ++t)
                    if (off[t] >= 0) *reinterpret_cast<float4*>(out.ptr + off[t] + c0 + (lane & 3) * 4) = val[t];
                __syncwarp();
            };
            // the TMEM read of pass p + 1 is in flight while pass p is converted and stored
            uint32_t va[CH], vb[CH];
            tmem_ld16_issue(taddr, va);
#pragma unroll
            for (int p = 0; p < NP / (2 * CH); p += 2) {
                tmem_ld_wait();
                tmem_ld16_issue(taddr + (p + 1) * 2 * CH, vb);
                emit(va, h * CH + p * 2 * CH);
                tmem_ld_wait();
                if (p + 2 < NP / (2 * CH)) tmem_ld16_issue(taddr + (p + 2) * 2 * CH, va);
                else {                               // every column of this warp's share is in registers: hand the accumulator back
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) remote_arrive(acc_empty_leader[acc]);
                }
                emit(vb, h * CH + (p + 1) * 2 * CH);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                          // the peer may still signal this CTA's barriers
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512));
}

constexpr size_t PAIR_SMEM_BUDGET = 227 * 1024 - 8192;   // dynamic window any configuration may ask for; the kernel's static shared memory
                                                         // (barriers, row offsets, bias: 4272 B) has to fit beside it in the 227 KB opt-in limit

// The dynamic shared-memory opt-in is per function and per device, and both launchers below launch the same kernel with sizes that
// depend on K: opt in ONCE per device for the whole budget instead of tracking a high-water mark per launcher.
int pair_smem_opt_in() {
    static bool done[CRB3D_MAX_DEVICES] = {};
    const int dev = crb3d_current_device();
    if (!done[dev]) {
        CRB3D_CUDA(cudaFuncSetAttribute(bev_gemm_pair_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PAIR_SMEM_BUDGET));
        done[dev] = true;
    }
    return CRB3D_OK;
}

}  // namespace

CRB3D_DIAG_DEFINE_SETTER(bev_gemm_pair)

// D = A @ W[sub]^T (+ bias) (ReLU) for n_sub slices of N = 256 output columns, K % 32 == 0, K <= 256; one output segment holding all
// 256 columns (rows row_stride floats apart, 16-byte aligned). CRB3D_ERR_UNSUPPORTED for anything else (the caller then uses
// bev_gemm_tc). relu bits as crb3d_bev_gemm_tf32. Not part of include/crb3d.h: reached through crb3d_bev_gemm_tf32.
int crb3d_bev_gemm_pair_tf32(const float* A, long long M, int K, long long lda, const float* W, int n_sub, const float* bias, int relu,
                             float* out_ptr, long long row_stride, int up, int in_h, int in_w, cudaStream_t stream) {
    if (K % BK != 0 || K > 256 || K <= 0 || M > 0x7fffffffLL || row_stride % 4 != 0 || ((uintptr_t)out_ptr & 15) != 0)
        return CRB3D_ERR_UNSUPPORTED;
    const int nkb = K / BK;
    CUtensorMap amap, wmap;
    {
        const uint64_t dims[2] = {(uint64_t)K, (uint64_t)M}, strides[1] = {(uint64_t)lda * 4};
        const uint32_t box[2] = {BK, TILE_M};
        int rc = make_map_f32(&amap, A, 2, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
        if (rc) return rc;
    }
    {
        const uint64_t dims[2] = {(uint64_t)K, (uint64_t)NP * n_sub}, strides[1] = {(uint64_t)K * 4};
        const uint32_t box[2] = {BK, (uint32_t)NH};
        int rc = make_map_f32(&wmap, W, 2, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
        if (rc) return rc;
    }
    const size_t budget = PAIR_SMEM_BUDGET;
    const int bres = (relu & 16) ? 1 : 0;        // relu bit 4 (A/B runs): weights resident instead of streamed with the activations
    const size_t stage_bytes = bres ? A_BYTES : A_BYTES + B_BYTES, fixed = 1024 + STAGING + (bres ? (size_t)nkb * B_BYTES : 0);
    int stages = (int)((budget - fixed) / stage_bytes);
    if (stages > MAX_STAGES) stages = MAX_STAGES;
    if (stages < 2) return CRB3D_ERR_UNSUPPORTED;
    const size_t smem = fixed + (size_t)stages * stage_bytes;
    { int rc = pair_smem_opt_in(); if (rc) return rc; }
    const int n_pairs = (int)crb3d_divup(M, 2 * TILE_M);
    int cps = crb3d_num_sms() / 2 / n_sub;      // clusters per sub-position
    if (cps < 1) cps = 1;
    if (cps > n_pairs) cps = n_pairs;
    PairOut o;
    o.ptr = out_ptr; o.row_stride = row_stride; o.up = up; o.in_w = in_w; o.in_h = in_h;
    bev_gemm_pair_tc<<<(unsigned)(2 * n_sub * cps), NTHREADS, smem, stream>>>(amap, wmap, (int)M, K, stages, bres, cps, bias, relu, o, PairConv{});
    CRB3D_CHECK_LAUNCH();
    return CRB3D_OK;
}

// k x k convolution (stride, zero padding) + bias + ReLU with 256 output channels as an implicit GEMM on CTA pairs: both CTAs' 8 x 16
// pixel tiles against all 256 channels in one M256 x N256 MMA, so an activation box is fetched once (not once per 128-channel block)
// and each CTA streams half of every weight k-block. in: (B, H, W, C_in); w2: [256][ksize*ksize][C_in]; out: (B, H_out, W_out, 256).
// Reached through crb3d_bev_conv_gemm_tf32 (base_bev_backbone.py:33-40: the stride-2 first conv of BEV block 2).
int crb3d_bev_conv_gemm_pair_tf32(const float* in, int B, int H, int W, int cin, const float* w2, int ksize, int stride, int pad,
                                  const float* bias, int relu, float* out_ptr, cudaStream_t stream) {
    if (cin % BK != 0 || ((uintptr_t)out_ptr & 15) != 0) return CRB3D_ERR_UNSUPPORTED;
    PairConv cv;
    cv.on = 1;
    cv.h_out = (H + 2 * pad - ksize) / stride + 1;
    cv.w_out = (W + 2 * pad - ksize) / stride + 1;
    cv.tiles_x = (int)crb3d_divup(cv.w_out, 16);
    cv.tiles_y = (int)crb3d_divup(cv.h_out, 8);
    cv.stride = stride; cv.pad = pad; cv.cblocks = cin / BK; cv.ksize = ksize; cv.n_img = B;
    const int K = ksize * ksize * cin;
    CUtensorMap amap, wmap;
    {
        const uint64_t dims[4] = {(uint64_t)cin, (uint64_t)W, (uint64_t)H, (uint64_t)B};
        const uint64_t strides[3] = {(uint64_t)cin * 4, (uint64_t)W * cin * 4, (uint64_t)H * W * cin * 4};
        // with a traversal stride the box is the EXTENT walked in the tensor: ceil(box / stride) elements are copied
        const uint32_t box[4] = {BK, (uint32_t)(16 * stride), (uint32_t)(8 * stride), 1};
        const uint32_t es[4] = {1, (uint32_t)stride, (uint32_t)stride, 1};
        int rc = make_map_f32(&amap, in, 4, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B, es);
        if (rc) return rc;
    }
    {
        const uint64_t dims[2] = {(uint64_t)K, (uint64_t)NP}, strides[1] = {(uint64_t)K * 4};
        const uint32_t box[2] = {BK, (uint32_t)NH};
        int rc = make_map_f32(&wmap, w2, 2, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
        if (rc) return rc;
    }
    const size_t budget = PAIR_SMEM_BUDGET, stage_bytes = A_BYTES + B_BYTES, fixed = 1024 + STAGING;
    int stages = (int)((budget - fixed) / stage_bytes);
    if (stages > MAX_STAGES) stages = MAX_STAGES;
    const size_t smem = fixed + (size_t)stages * stage_bytes;
    { int rc = pair_smem_opt_in(); if (rc) return rc; }
    const int n_tiles = B * cv.tiles_x * cv.tiles_y, n_pairs = (n_tiles + 1) / 2;
    int cps = crb3d_num_sms() / 2;
    if (cps < 1) cps = 1;
    if (cps > n_pairs) cps = n_pairs;
    PairOut o;
    o.ptr = out_ptr; o.row_stride = NP; o.up = 0; o.in_w = 0; o.in_h = 0;
    bev_gemm_pair_tc<<<(unsigned)(2 * cps), NTHREADS, smem, stream>>>(amap, wmap, n_tiles, K, stages, 0, cps, bias, relu, o, cv);
    CRB3D_CHECK_LAUNCH();
    return CRB3D_OK;
}
