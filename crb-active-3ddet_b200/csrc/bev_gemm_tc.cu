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
// Persistent CTAs (one per SM), each walking 128-row tiles of one weight slice. Warp roles (mbarrier pipelined):
//   warp 0, one lane : TMA producer - per 32-channel k-block one 128x32 box of A (and, when the weights are not
//                      resident, one Nx32 box of W), 128B swizzle, STAGES deep across tile boundaries
//   warp 1, one lane : tcgen05.mma issuer (4 x M128 x N x K8 per k-block) into one of TWO TMEM accumulators
//   warps 2-5        : epilogue of the other accumulator - tcgen05.ld, bias/ReLU, rows staged in shared memory and written
//                      out as whole 128-byte lines
// The kernels are HBM-bound (K <= 512): bytes/row = 4*(K + N); a first one-tile-per-CTA version spent most of its time
// in per-CTA prologue/epilogue latency (24 % of the DRAM peak in ncu, profiles/r01_bev_gemm_v1.txt).
#include "tc_common.cuh"

namespace {

using namespace tc;

constexpr int TILE_M = 128;
constexpr int BK = 32;  // floats per k-block = one 128-byte swizzle row

// ---- thread-block clusters: CTAs that consume the same operand k-block fetch it ONCE from L2 (TMA multicast) ----
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// the box lands at the same shared-memory offset of every CTA in `mask` and completes bytes on the same barrier offset there
__device__ __forceinline__ void tma_load_2d_mc(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar, uint16_t mask) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%2, %3}], [%4], %5;"
                 ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(smem_u32(bar)), "h"(mask) : "memory");
}
__device__ __forceinline__ void tma_load_4d_mc(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2, int c3, uint64_t* bar,
                                               uint16_t mask) {
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%2, %3, %4, %5}], [%6], %7;"
                 ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(smem_u32(bar)), "h"(mask) : "memory");
}
// arrives on the barrier at this offset in every CTA of `mask` once the MMAs issued so far have retired
__device__ __forceinline__ void umma_commit_mc(uint64_t* bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"(mask) : "memory");
}

struct GemmOut {
    float* ptr[3];            // up to three column segments, each to its own tensor
    int col_begin[3];
    int width[3];
    long long row_stride[3];  // floats between consecutive output rows of the segment
    int n_seg;
    int up;                   // 0: output row = GEMM row; 2: 2x2 transposed conv - row (b,y,x) -> (b, 2y+dy, 2x+dx)
    int in_w, in_h;           // input spatial size when up == 2
};

// CONV mode: the GEMM rows are the output pixels of a k x k convolution (stride s, zero padding p) over a channels-last map,
// a 128-row tile = 8 x 16 output pixels, and k-block kb = (tap, 32-channel block): its A operand is ONE 4-D TMA box
// {32 channels, 16 x, 8 y, 1 image} whose traversal stride along x and y is the conv stride and whose out-of-bounds
// pixels are zero-filled by the TMA unit (that is the padding). Weights: [N][tap][C_in] = K index tap * C_in + c.
struct ConvArgs {
    int tiles_x, tiles_y, h_out, w_out, stride, pad, cblocks, ksize, n_img;
};

template <int N>
struct Cfg {
    static constexpr int TMEM_N = N <= 32 ? 32 : (N <= 64 ? 64 : (N <= 128 ? 128 : 256));   // columns per accumulator
    static constexpr int A_BYTES = TILE_M * 128;
    static constexpr int B_BYTES = N * 128;           // one 32-channel k-block of the weights
    static constexpr int CH = 32;                     // columns staged per epilogue pass
    static constexpr int PITCH = CH + 4;              // floats; +4 keeps float4 rows conflict-free
    // ROWS mode: 8 warps x 32 rows x PITCH; DENSE mode: 4 quarters x 32 rows x N
    static constexpr int staging_bytes(bool dense) { return ((dense ? 4 * 32 * N * 4 : 8 * 32 * PITCH * 4) + 1023) & ~1023; }
};

// PERSISTENT: gridDim.x = n_slices * ctas_per_slice. A CTA serves ONE weight slice (N output columns of one
// sub-position) and walks the 128-row tiles  t = rank, rank + ctas_per_slice, ...  of the activation matrix:
//   BRES = true : the slice's whole [N][K] weight block is loaded once and stays in shared memory; only activation
//                 k-blocks stream through the STAGES-deep ring;
//   BRES = false: weight k-blocks stream with the activations (L2 hits) - used when N*K*4 does not fit.
// Two TMEM accumulators: the MMA warp fills one while the epilogue warps drain the other, so loads, MMAs and stores of
// consecutive tiles overlap and the kernel runs at the HBM rate of its activation read + output write.
// CLUSTERS (CSA x CSW CTAs, launched with a cluster dimension; OPT-IN, see the measurements at the dispatch sites): the CSA CTAs
// that hold different weight slices but walk the SAME activation tile each load 1/CSA of every A k-block and multicast it to the
// group; with streamed weights the CSW CTAs that hold the SAME slice but different tiles do the same for the weight k-blocks.
// L2 -> SM operand traffic per CTA drops from A + W to A / CSA + W / CSW. A stage is refilled only after every CTA that receives a
// piece from this one has consumed it: tcgen05.commit is multicast to the writers' empty barriers. Results are bit-identical to the
// plain launch; it is slower on B200 at the BEV shapes because the ring is latency-bound, not L2-byte-bound.
template <int N, int STAGES, bool BRES, bool DENSE, bool CONV = false, int CSA = 1, int CSW = 1>
__global__ void __launch_bounds__(320, 1) bev_gemm_tc(const __grid_constant__ CUtensorMap amap,
                                                      const __grid_constant__ CUtensorMap wmap, int M, int K,
                                                      int ctas_per_slice, int halves, const float* __restrict__ bias, int relu,
                                                      const __grid_constant__ GemmOut out, const __grid_constant__ ConvArgs cv) {
    using C = Cfg<N>;
    constexpr int STAGE_BYTES = C::A_BYTES + (BRES ? 0 : C::B_BYTES);
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // pointer arithmetic keeps the shared address space (LDS/STS, not generic LD/ST)
    __shared__ uint64_t full_bar[STAGES], empty_bar[STAGES], acc_full[2], acc_empty[2], b_full;
    __shared__ uint32_t tmem_base_s;
    __shared__ long long rowoff_s[8][32];          // ROWS mode: output offset of each staged row, per epilogue warp
    __shared__ int colbase_s[DENSE ? N : 1], colw_s[DENSE ? N : 1];   // DENSE mode: column -> staging offset / segment width
    __shared__ __align__(16) float bias_s[N];      // this slice's bias (zeros when absent): no global load in the epilogue's inner loop

    constexpr int CS = CSA * CSW;
    static_assert(CS == 1 || (!DENSE && (CSW == 1 || !BRES) && TILE_M / CSA % 8 == 0 && N / CSW % 8 == 0 && 8 % CSA == 0), "cluster geometry");
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    // cluster rank -> (a_idx: which weight slice of the group, w_idx: which tile of the cluster's tile group)
    const int cr = CS > 1 ? (int)cluster_ctarank() : 0, a_idx = cr % CSA, w_idx = cr / CSA;
    const int cluster_id = blockIdx.x / CS;
    const int group = cluster_id / ctas_per_slice, crank = cluster_id - group * ctas_per_slice;   // ctas_per_slice = clusters per slice group
    const int slice = group * CSA + a_idx;
    const int sub = slice / halves, half = slice - sub * halves;   // transposed-conv sub-position, column block
    const int nkb = K / BK;
    const int n_tiles = CONV ? M : (M + TILE_M - 1) / TILE_M;    // CONV: M counts 8 x 16 pixel tiles
    // tile of iteration i: the same trip count for every CTA of a cluster (a tile index beyond n_tiles loads zeros, stores nothing)
    const int tile0 = crank * CSW + w_idx, tile_step = ctas_per_slice * CSW;
    const int my_tiles = crank * CSW < n_tiles ? (n_tiles - crank * CSW + tile_step - 1) / tile_step : 0;
    const uint16_t mask_a = (uint16_t)(((1u << CSA) - 1u) << (w_idx * CSA));
    uint16_t mask_w = 0;
#pragma unroll
    for (int w = 0; w < CSW; ++w) mask_w |= (uint16_t)(1u << (w * CSA + a_idx));

    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], CSA + CSW - 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(&acc_full[s], 1); mbar_init(&acc_empty[s], 8); }
        mbar_init(&b_full, 1);
        mbar_fence_init();
        tma_prefetch_desc(&amap);
        tma_prefetch_desc(&wmap);
    }
    if (warp == 1) tmem_alloc<2 * C::TMEM_N>(&tmem_base_s);
    for (int c = tid; c < N; c += blockDim.x) bias_s[c] = bias ? __ldg(&bias[half * N + c]) : 0.0f;
    if (DENSE) {
        for (int c = tid; c < N; c += blockDim.x) {
            int base = 0, cb = -1, w = 0;
            for (int sgi = 0; sgi < out.n_seg; ++sgi) {
                if (c >= out.col_begin[sgi] && c < out.col_begin[sgi] + out.width[sgi]) { cb = base + (c - out.col_begin[sgi]); w = out.width[sgi]; }
                base += 32 * out.width[sgi];
            }
            colbase_s[c] = cb;
            colw_s[c] = w;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (CS > 1) cluster_sync_all();              // every CTA's barriers exist before a peer's multicast / commit can reach them
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_s;
    // layout: [staging][resident weights (BRES)][stage ring]
    const uint32_t smem_base = smem_u32(smem);
    const uint32_t bres_base = smem_base + C::staging_bytes(DENSE);
    const uint32_t ring_base = bres_base + (BRES ? (uint32_t)nkb * C::B_BYTES : 0u);
    const int wrow = slice * N;                  // first weight row of this slice

    if (warp == 0) {
        if (lane == 0 && my_tiles > 0) {
            if (BRES) {
                mbar_expect_tx(&b_full, (uint32_t)nkb * C::B_BYTES);
                for (int kb = 0; kb < nkb; ++kb) tma_load_2d(bres_base + kb * C::B_BYTES, &wmap, kb * BK, wrow, &b_full);
            }
            int it = 0;
            for (int i = 0; i < my_tiles; ++i) {
                const int t = tile0 + i * tile_step;
                const int m0 = t * TILE_M;
                int cb_img = 0, cx0 = 0, cy0 = 0;
                if (CONV) {
                    const int per_img = cv.tiles_x * cv.tiles_y;
                    cb_img = t / per_img;
                    const int rem = t - cb_img * per_img;
                    cy0 = (rem / cv.tiles_x) * 8 * cv.stride - cv.pad;
                    cx0 = (rem % cv.tiles_x) * 16 * cv.stride - cv.pad;
                }
                for (int kb = 0; kb < nkb; ++kb, ++it) {
                    const int stage = it % STAGES;
                    if (it >= STAGES) mbar_wait(&empty_bar[stage], ((it / STAGES) - 1) & 1, (CRB3D_K_BEV_GEMM << 8) | 9);
                    const uint32_t a_dst = ring_base + stage * STAGE_BYTES;
                    mbar_expect_tx(&full_bar[stage], STAGE_BYTES);
                    if (CS == 1) {
                        if (CONV) {
                            const int tap = kb / cv.cblocks, cb = kb - tap * cv.cblocks;
                            tma_load_4d(a_dst, &amap, cb * BK, cx0 + tap % cv.ksize, cy0 + tap / cv.ksize, cb_img, &full_bar[stage]);
                        } else
                            tma_load_2d(a_dst, &amap, kb * BK, m0, &full_bar[stage]);
                        if (!BRES) tma_load_2d(a_dst + C::A_BYTES, &wmap, kb * BK, wrow, &full_bar[stage]);
                    } else {
                        // this CTA's piece of the A k-block (rows a_idx * 128 / CSA ..) to every CTA that walks this tile ...
                        constexpr int RA = TILE_M / CSA;
                        if (CONV) {
                            const int tap = kb / cv.cblocks, cb = kb - tap * cv.cblocks;
                            tma_load_4d_mc(a_dst + a_idx * RA * 128, &amap, cb * BK, cx0 + tap % cv.ksize,
                                           cy0 + tap / cv.ksize + a_idx * (8 / CSA) * cv.stride, cb_img, &full_bar[stage], mask_a);
                        } else
                            tma_load_2d_mc(a_dst + a_idx * RA * 128, &amap, kb * BK, m0 + a_idx * RA, &full_bar[stage], mask_a);
                        // ... and its piece of the weight k-block (rows w_idx * N / CSW ..) to every CTA that holds this slice
                        if (!BRES) {
                            constexpr int RW = N / CSW;
                            tma_load_2d_mc(a_dst + C::A_BYTES + w_idx * RW * 128, &wmap, kb * BK, wrow + w_idx * RW, &full_bar[stage], mask_w);
                        }
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0 && my_tiles > 0) {
            const uint32_t idesc = idesc_tf32(TILE_M, N);
            if (BRES) mbar_wait(&b_full, 0, (CRB3D_K_BEV_GEMM << 8) | 10);
            int it = 0;
            for (int i = 0; i < my_tiles; ++i) {
                const int acc = i & 1;
                if (i >= 2) mbar_wait(&acc_empty[acc], ((i >> 1) - 1) & 1, (CRB3D_K_BEV_GEMM << 8) | 6);   // the epilogue has drained this accumulator
                tc_fence_after();
                for (int kb = 0; kb < nkb; ++kb, ++it) {
                    const int stage = it % STAGES;
                    mbar_wait(&full_bar[stage], (it / STAGES) & 1, (CRB3D_K_BEV_GEMM << 8) | 8);
                    tc_fence_after();
                    const uint32_t a_base = ring_base + stage * STAGE_BYTES;
                    const uint32_t b_base = BRES ? bres_base + kb * C::B_BYTES : a_base + C::A_BYTES;
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        umma_tf32(tmem_base + acc * C::TMEM_N, desc_sw128(a_base + j * 32), desc_sw128(b_base + j * 32), idesc,
                                  (kb > 0 || j > 0) ? 1u : 0u);
                    if (CS == 1) umma_commit(&empty_bar[stage]);
                    else umma_commit_mc(&empty_bar[stage], (uint16_t)(mask_a | mask_w));   // frees the stage for every CTA that writes into it
                }
                umma_commit(&acc_full[acc]);
            }
        }
    } else {
        // ================================ epilogue: 8 warps, two per TMEM lane quarter ================================
        const int q = warp & 3, h = (warp - 2) >> 2;      // quarter (TMEM lanes 32q..32q+31), which half of the columns
        const int r = q * 32 + lane;                      // tile row owned by this thread in the TMEM read
        const int col0 = half * N;                        // first output column of this slice (within the logical N_total)
        float* stage_base = reinterpret_cast<float*>(smem);
        if (!DENSE) {
            // ---- one segment of full rows: each warp stages 32 rows x 32 columns and stores whole 128-byte lines
            float* stage_w = stage_base + (size_t)(warp - 2) * 32 * C::PITCH;
            long long* rowoff = rowoff_s[warp - 2];
            const long long stride = out.row_stride[0];
            float* obase = out.ptr[0] + col0;
            for (int i = 0; i < my_tiles; ++i) {
                const int acc = i & 1;
                const long long m = (long long)(tile0 + i * tile_step) * TILE_M + r;
                long long orow = -1;                      // output row of this lane's tile row; -1 = beyond M
                if (CONV) {                               // tile row r = pixel (y0 + r / 16, x0 + r % 16) of image b
                    const int t = tile0 + i * tile_step, per_img = cv.tiles_x * cv.tiles_y;
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
                rowoff[lane] = orow < 0 ? -1 : orow * stride;
                mbar_wait(&acc_full[acc], (i >> 1) & 1, (CRB3D_K_BEV_GEMM << 8) | 5);
                tc_fence_after();
#pragma unroll 1
                for (int c0 = h * 32; c0 < N; c0 += 64) {
                    uint32_t v[32];
                    tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * C::TMEM_N + c0), v);
                    if (c0 + 64 >= N) {           // this warp's last TMEM read of the tile: hand the accumulator back
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&acc_empty[acc]);
                    }
#pragma unroll
                    for (int j = 0; j < 32; j += 4) {
                        float4 w;
                        float* wp = reinterpret_cast<float*>(&w);
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            float x = __uint_as_float(v[j + u]);
                            x += bias_s[c0 + j + u];
                            if (relu & 1) x = fmaxf(x, 0.0f);
                            if (relu & 2) x = tf32_rn(x);
                            wp[u] = x;
                        }
                        *reinterpret_cast<float4*>(stage_w + (size_t)lane * C::PITCH + j) = w;
                    }
                    __syncwarp();
                    float4 val[8];
                    long long off[8];
#pragma unroll
                    for (int it = 0; it < 8; ++it) {      // 4 rows x 128 bytes per warp instruction
                        const int rr = it * 4 + (lane >> 3);
                        off[it] = rowoff[rr];
                        val[it] = *reinterpret_cast<const float4*>(stage_w + (size_t)rr * C::PITCH + (lane & 7) * 4);
                    }
#pragma unroll
                    for (int it = 0; it < 8; ++it)
                        if (off[it] >= 0) *reinterpret_cast<float4*>(obase + off[it] + c0 + (lane & 7) * 4) = val[it];
                    __syncwarp();
                }
            }
        } else {
            // ---- contiguous-row segments (row_stride == width): the quarter's two warps fill one dense [segment][32 rows][w]
            // staging block, then copy each segment out as one contiguous run of float4
            float* stage_q = stage_base + (size_t)q * 32 * N;
            const int tq = h * 32 + lane;                 // thread index within the quarter (0..63)
            for (int i = 0; i < my_tiles; ++i) {
                const int acc = i & 1;
                const long long mrow0 = (long long)(tile0 + i * tile_step) * TILE_M + q * 32;   // first row of the quarter
                mbar_wait(&acc_full[acc], (i >> 1) & 1, (CRB3D_K_BEV_GEMM << 8) | 5);
                tc_fence_after();
#pragma unroll 1
                for (int c0 = h * 16; c0 < N; c0 += 32) {
                    uint32_t v[16];
                    tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * C::TMEM_N + c0), v);
                    if (c0 + 32 >= N) {
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&acc_empty[acc]);
                    }
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const int c = c0 + j;
                        const int cbse = colbase_s[c];
                        if (cbse < 0) continue;            // padding column
                        float x = __uint_as_float(v[j]);
                        x += bias_s[c];
                        if (relu & 1) x = fmaxf(x, 0.0f);
                        if (relu & 2) x = tf32_rn(x);
                        stage_q[cbse + lane * colw_s[c]] = x;
                    }
                }
                asm volatile("bar.sync %0, 64;" ::"r"(1 + q));
                const int rows_valid = (int)max(0ll, min(32ll, (long long)M - mrow0));
                int sbase = 0;
                for (int sgi = 0; sgi < out.n_seg; ++sgi) {
                    const int w = out.width[sgi];
                    float* dst = out.ptr[sgi] + mrow0 * w;
                    const float* src = stage_q + sbase;
                    if (rows_valid == 32) {
                        for (int e = tq; e < 8 * w; e += 64)
                            reinterpret_cast<float4*>(dst)[e] = reinterpret_cast<const float4*>(src)[e];
                    } else {
                        for (int e = tq; e < rows_valid * w; e += 64) dst[e] = src[e];
                    }
                    sbase += 32 * w;
                }
                asm volatile("bar.sync %0, 64;" ::"r"(1 + q));
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (CS > 1) cluster_sync_all();              // a peer may still multicast into this CTA's stages / arrive on its barriers
    if (warp == 1) tmem_dealloc<2 * C::TMEM_N>(tmem_base);
}

// plain launch, or a cluster launch of CS consecutive CTAs
template <typename Kern, typename... Args>
int launch_maybe_cluster(Kern kern, unsigned grid, size_t smem, int cs, cudaStream_t stream, Args... args) {
    if (cs <= 1) {
        kern<<<grid, 320, smem, stream>>>(args...);
    } else {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(grid);
        cfg.blockDim = dim3(320);
        cfg.dynamicSmemBytes = smem;
        cfg.stream = stream;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = (unsigned)cs;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        CRB3D_CUDA(cudaLaunchKernelEx(&cfg, kern, args...));
    }
    CRB3D_CHECK_LAUNCH();
    return CRB3D_OK;
}

template <int N, int STAGES, bool BRES, bool DENSE, int CSA = 1>
int launch_gemm(const float* A, long long M, int K, long long lda, const float* W, int n_slices, int halves, const float* bias,
                int relu, const GemmOut& out, cudaStream_t stream) {
    using C = Cfg<N>;
    CUtensorMap amap, wmap;
    {
        const uint64_t dims[2] = {(uint64_t)K, (uint64_t)M}, strides[1] = {(uint64_t)lda * 4};
        const uint32_t box[2] = {BK, TILE_M / CSA};        // clusters: every CTA loads (and multicasts) its 1/CSA of the rows
        int rc = make_map_f32(&amap, A, 2, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
        if (rc) return rc;
    }
    {
        const uint64_t dims[2] = {(uint64_t)K, (uint64_t)N * n_slices}, strides[1] = {(uint64_t)K * 4};
        const uint32_t box[2] = {BK, (uint32_t)N};
        int rc = make_map_f32(&wmap, W, 2, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
        if (rc) return rc;
    }
    const size_t smem = 1024 + C::staging_bytes(DENSE) + (BRES ? (size_t)(K / BK) * C::B_BYTES : 0) +
                        (size_t)STAGES * (C::A_BYTES + (BRES ? 0 : C::B_BYTES));
    if (smem > 227 * 1024) return CRB3D_ERR_UNSUPPORTED;
    if (n_slices % CSA != 0) return CRB3D_ERR_UNSUPPORTED;
    auto kern = bev_gemm_tc<N, STAGES, BRES, DENSE, false, CSA, 1>;
    static size_t smem_set[CRB3D_MAX_DEVICES] = {};   // the attribute is per function per device
    const int dev = crb3d_current_device();
    if (smem > smem_set[dev]) {
        CRB3D_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        smem_set[dev] = smem;
    }
    const int n_tiles = (int)crb3d_divup(M, TILE_M);
    int per_slice = crb3d_num_sms() / n_slices;        // CTAs per slice = clusters per group of CSA slices
    if (per_slice < 1) per_slice = 1;
    if (per_slice > n_tiles) per_slice = n_tiles;
    return launch_maybe_cluster(kern, (unsigned)(n_slices * per_slice), smem, CSA, stream, amap, wmap, (int)M, K, per_slice, halves, bias,
                                relu, out, ConvArgs{});
}

// k x k conv (stride, zero padding) as an implicit GEMM on the same persistent kernel (CONV mode): N = 128 output channels
// per CTA (two slices for 256), weights streamed with the activations. Clusters of CSA x CSW: the CSA output-channel slices share
// each activation box, CSW neighbouring tiles share each weight box.
template <int N, int STAGES, int CSA, int CSW>
int launch_conv_gemm(const float* in, int B, int H, int W, int cin, const float* w2, int cout, int ksize, int stride, int pad,
                     const float* bias, int relu, float* out_ptr, cudaStream_t stream) {
    using C = Cfg<N>;
    ConvArgs cv;
    cv.h_out = (H + 2 * pad - ksize) / stride + 1;
    cv.w_out = (W + 2 * pad - ksize) / stride + 1;
    cv.tiles_x = (int)crb3d_divup(cv.w_out, 16);
    cv.tiles_y = (int)crb3d_divup(cv.h_out, 8);
    cv.stride = stride; cv.pad = pad; cv.cblocks = cin / BK; cv.ksize = ksize; cv.n_img = B;
    const int K = ksize * ksize * cin;
    const int n_slices = cout / N;
    if (n_slices % CSA != 0) return CRB3D_ERR_UNSUPPORTED;
    CUtensorMap amap, wmap;
    {
        const uint64_t dims[4] = {(uint64_t)cin, (uint64_t)W, (uint64_t)H, (uint64_t)B};
        const uint64_t strides[3] = {(uint64_t)cin * 4, (uint64_t)W * cin * 4, (uint64_t)H * W * cin * 4};
        // with a traversal stride the box is the EXTENT walked in the tensor: ceil(box / stride) elements are copied.
        // clusters: a CTA loads 8 / CSA of the tile's 8 pixel rows
        const uint32_t box[4] = {BK, (uint32_t)(16 * stride), (uint32_t)((8 / CSA) * stride), 1};
        const uint32_t es[4] = {1, (uint32_t)stride, (uint32_t)stride, 1};
        int rc = make_map_f32(&amap, in, 4, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B, es);
        if (rc) return rc;
    }
    {
        const uint64_t dims[2] = {(uint64_t)K, (uint64_t)cout}, strides[1] = {(uint64_t)K * 4};
        const uint32_t box[2] = {BK, (uint32_t)(N / CSW)};
        int rc = make_map_f32(&wmap, w2, 2, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
        if (rc) return rc;
    }
    const size_t smem = 1024 + C::staging_bytes(false) + (size_t)STAGES * (C::A_BYTES + C::B_BYTES);
    auto kern = bev_gemm_tc<N, STAGES, false, false, true, CSA, CSW>;
    static size_t smem_set[CRB3D_MAX_DEVICES] = {};
    const int dev = crb3d_current_device();
    if (smem > smem_set[dev]) {
        CRB3D_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        smem_set[dev] = smem;
    }
    GemmOut o{};
    o.n_seg = 1; o.up = 0; o.ptr[0] = out_ptr; o.col_begin[0] = 0; o.width[0] = cout; o.row_stride[0] = cout;
    const int n_tiles = B * cv.tiles_x * cv.tiles_y;
    constexpr int CS = CSA * CSW;
    int clusters_per_group = crb3d_num_sms() / CS / (n_slices / CSA);      // a group = CSA slices; a cluster walks CSW tiles at a time
    if (clusters_per_group < 1) clusters_per_group = 1;
    if (clusters_per_group > (int)crb3d_divup(n_tiles, CSW)) clusters_per_group = (int)crb3d_divup(n_tiles, CSW);
    return launch_maybe_cluster(kern, (unsigned)((n_slices / CSA) * clusters_per_group * CS), smem, CS, stream, amap, wmap, n_tiles, K,
                                clusters_per_group, n_slices, bias, relu, o, cv);
}

}  // namespace

int crb3d_bev_gemm_pair_tf32(const float* A, long long M, int K, long long lda, const float* W, int n_sub, const float* bias, int relu,
                             float* out_ptr, long long row_stride, int up, int in_h, int in_w, cudaStream_t stream);   // bev_gemm_pair.cu
int crb3d_bev_conv_gemm_pair_tf32(const float* in, int B, int H, int W, int cin, const float* w2, int ksize, int stride, int pad,
                                  const float* bias, int relu, float* out_ptr, cudaStream_t stream);                           // bev_gemm_pair.cu

// A: (M, K) fp32 rows `lda` floats apart (channels-last activations); W: contiguous [n_sub][N][K]; bias: N floats or null.
// Output: n_seg column segments (seg s = columns [col_begin[s], col_begin[s]+width[s]) -> out_ptr[s] + row*row_stride[s]).
// up = 0: output row = GEMM row, n_sub must be 1. up = 2: ConvTranspose2d(kernel = stride = 2): n_sub = 4 weight slices
// ordered (dy, dx), GEMM row (b, y, x) of an in_h x in_w map lands on output pixel (b, 2y+dy, 2x+dx).
// relu: bit 0 = ReLU, bit 1 = round the stored values to TF32 (round-to-nearest) so that a following tensor-core layer
// reads them exactly. Supported: K % 32 == 0, N in {80, 128, 256} (pad weights/bias with zero rows to reach a supported N).
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
    // output placement modes: ROWS = one segment holding all N columns (rows row_stride apart); DENSE = consecutive
    // segments of contiguous rows (row_stride == width), N = 80 only
    bool rows = n_seg == 1 && col_begin[0] == 0 && width[0] == N && row_stride[0] % 4 == 0 && ((uintptr_t)out_ptr[0] & 15) == 0;
    bool dense = up == 0;
    for (int s = 0, c = 0; s < n_seg; ++s) {
        dense = dense && col_begin[s] == c && row_stride[s] == width[s] && ((uintptr_t)out_ptr[s] & 15) == 0;
        c += width[s];
    }
    // slices = (sub-positions) x (column blocks of the per-CTA width); resident weights when N_cta * K * 4 <= 128 KB
    if (rows) {
        if (N == 256 && K <= 128 && !(relu & 32)) return launch_gemm<256, 3, true, false>(A, M, K, lda, W, n_sub, 1, bias, relu, o, stream);
        if (N == 256 && K <= 256 && !(relu & 12)) {
            // CTA pairs: all 256 columns from one pass over the activations (csrc/bev_gemm_pair.cu); relu bit 3 = the single-CTA kernel
            int rc = crb3d_bev_gemm_pair_tf32(A, M, K, lda, W, n_sub, bias, relu, out_ptr[0], row_stride[0], up, in_h, in_w, stream);
            if (rc != CRB3D_ERR_UNSUPPORTED) return rc;
        }
        if (N == 256 && K <= 256) {
            // relu bit 2 (opt-in, A/B measurements): the 2 column blocks x n_sub sub-positions walk the same activation tiles, so
            // clusters of 4 (or 2) slices can share every A box by TMA multicast. Measured at 16 x 100 x 88 x 256 -> 4 x 256:
            // 508 us against 274 us without clusters - the 3-stage ring is latency-bound (one k-block per ~latency / 3) and the
            // cross-CTA round trip (multicast commit -> refill -> multicast landing, slowest of 4 CTAs) is longer than the local one
            if ((relu & 4) && (n_sub * 2) % 4 == 0) return launch_gemm<128, 3, true, false, 4>(A, M, K, lda, W, n_sub * 2, 2, bias, relu, o, stream);
            if (relu & 4) return launch_gemm<128, 3, true, false, 2>(A, M, K, lda, W, n_sub * 2, 2, bias, relu, o, stream);
            return launch_gemm<128, 3, true, false>(A, M, K, lda, W, n_sub * 2, 2, bias, relu, o, stream);
        }
        if (N == 256) return launch_gemm<128, 5, false, false>(A, M, K, lda, W, n_sub * 2, 2, bias, relu, o, stream);
        if (N == 128 && K <= 256) return launch_gemm<128, 3, true, false>(A, M, K, lda, W, n_sub, 1, bias, relu, o, stream);
        if (N == 128) return launch_gemm<128, 5, false, false>(A, M, K, lda, W, n_sub, 1, bias, relu, o, stream);
    }
    if (dense && N == 80) return launch_gemm<80, 6, false, true>(A, M, K, lda, W, n_sub, 1, bias, relu, o, stream);
    return CRB3D_ERR_UNSUPPORTED;
}

CRB3D_DIAG_DEFINE_SETTER(bev_gemm)

// k x k convolution (ksize in {1, 3}, stride in {1, 2}, zero padding pad) + bias + ReLU over a channels-last map as an
// implicit GEMM whose A tiles are strided 4-D TMA boxes (CONV mode above). Replaces the cuDNN call behind the stride-2
// first conv of a BEV block (pcdet/models/backbones_2d/base_bev_backbone.py:33-40: ZeroPad2d(1) + Conv2d(3, stride 2)); also
// the fallback for 3x3 stride-1 layers the halo-tile kernel does not take.
// in: (B, H, W, C_in) fp32; w2: [C_out][ksize*ksize][C_in] (= weight.permute(0,2,3,1)), TF32-rounded by the caller;
// out: (B, H_out, W_out, C_out). Supported: C_in % 32 == 0, C_out in {128, 256}. relu bits as crb3d_bev_gemm_tf32.
extern "C" int crb3d_bev_conv_gemm_tf32(const float* in, int B, int H, int W, int cin, const float* w2, int cout, int ksize,
                                        int stride, int pad, const float* bias, int relu, float* out, cudaStream_t stream) {
    if (!in || !w2 || !out || B <= 0 || H <= 0 || W <= 0 || cin <= 0 || cout <= 0) return CRB3D_ERR_ARG;
    if (cin % BK != 0 || (cout != 128 && cout != 256) || (ksize != 1 && ksize != 3) || (stride != 1 && stride != 2) || pad < 0 ||
        pad >= ksize || H + 2 * pad < ksize || W + 2 * pad < ksize)
        return CRB3D_ERR_UNSUPPORTED;
    // relu bit 2 (opt-in, A/B measurements): 2 x 2 clusters for cout = 256 (both channel halves share the activation boxes, two
    // neighbouring tiles share the weight boxes), pairs of tiles for cout = 128. Measured at 16 x 200 x 176 x 128 -> 256, stride 2:
    // 461 us against 247 us without clusters (same reason as the deblock above: ring depth x latency, not L2 bytes, sets the rate)
    if (cout == 256 && !(relu & 12)) {   // CTA pairs, all 256 channels per activation box (csrc/bev_gemm_pair.cu); relu bit 3 = this file's kernel
        int rc = crb3d_bev_conv_gemm_pair_tf32(in, B, H, W, cin, w2, ksize, stride, pad, bias, relu, out, stream);
        if (rc != CRB3D_ERR_UNSUPPORTED) return rc;
    }
    if (!(relu & 4)) return launch_conv_gemm<128, 5, 1, 1>(in, B, H, W, cin, w2, cout, ksize, stride, pad, bias, relu, out, stream);
    if (cout == 256) return launch_conv_gemm<128, 5, 2, 2>(in, B, H, W, cin, w2, cout, ksize, stride, pad, bias, relu, out, stream);
    return launch_conv_gemm<128, 5, 1, 2>(in, B, H, W, cin, w2, cout, ksize, stride, pad, bias, relu, out, stream);
}
