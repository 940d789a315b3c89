// 3x3 / stride-1 / pad-1 convolution of the BEV backbone on the 5th-gen tensor cores (tcgen05, TF32 in, fp32 accumulate
// in TMEM) as a HALO-TILE implicit GEMM, with the folded BatchNorm shift + ReLU fused into the epilogue.
//
// Replaces, on the inference path, the cuDNN calls behind
//   pcdet/models/backbones_2d/base_bev_backbone.py:33-50,96-99   blocks[i]: ZeroPad2d + Conv2d(3x3) + BN + ReLU, then
//                                                                 LAYER_NUMS x (Conv2d(3x3, pad 1) + BN + ReLU)
// out[b, y, x, co] = relu(bias[co] + sum_{ky,kx,ci} in[b, y+ky-1, x+kx-1, ci] * W[co, ky, kx, ci]), channels-last.
//
// Why not im2col-style loads: at TF32 a 128 x 128 x 32 k-block needs 32 KB of operands per 270 tensor-pipe cycles,
// ~120 B/clk/SM against an L2 feed of ~42 B/clk/SM - the nine shifted copies of the activation tile alone would make
// the kernel L2-bound at a third of the tensor peak. Instead:
//   * the (8+2) x (16+2) input halo of a 128-pixel output tile is loaded ONCE per 16-channel chunk, by one 4-D TMA box
//     whose out-of-bounds rows/columns are zero-filled by the TMA unit (that IS the conv padding), as 64-byte pixel
//     rows with the 64B swizzle (full 32-byte sectors from L2). The swizzle is a function of the absolute shared-memory
//     address, so the operand of tap (kv,ku) is the same tile read through a descriptor whose start address is advanced
//     by (kv*10 + ku)*64 bytes (SBO = one v step). Activation traffic drops 9 x 128 / 180 = 6.4x;
//   * every weight slice (one tap, 16 input channels, 128 output channels, pre-packed on the host into the no-swizzle
//     slab layout) feeds the MMAs of several tiles.
// Two kernels: bev_conv3x3_pair_tc (default, further down: CTA pairs with cta_group::2 MMAs, persistent, accumulators
// double-buffered in TMEM) and its single-CTA predecessor bev_conv3x3_tc (flag bit 8; 4 tiles per CTA = the whole TMEM,
// warp 0 = weight producer, warp 1 = MMA issuer, warp 2 lane 0 = activation producer, warps 2-9 = epilogue staged in the
// drained pipeline buffers). profiles/r01_conv_pair_ncu.txt holds the measurements behind both designs.
#include "tc_common.cuh"

namespace {

using namespace tc;

constexpr int TU = 8, TV = 16;                 // output tile: 8 pixels along u (fast), 16 along v
constexpr int PU = TU + 2, PV = TV + 2;        // halo tile
constexpr int KC = 16;                         // input channels per chunk = one 64-byte swizzle row = 2 MMAs (K = 8)
constexpr int NT = 4;                          // output tiles per CTA
constexpr int N = 128;                         // output channels per CTA
constexpr int TILE_A_BYTES = PU * PV * KC * 4; // 11520 B land per (tile, chunk)
constexpr int TILE_A = 12288;                  // ... in a slot padded to the 1024 B alignment of the swizzle pattern
constexpr int CHUNK_A = NT * TILE_A;           // 49152 B
constexpr int NSTAGE_A = 3;
constexpr int SLAB_B = N * 16;                 // 2048 B
constexpr int STAGE_B = (KC / 4) * SLAB_B;     // 8192 B per (tap, chunk)
constexpr int NSTAGE_B = 6;
constexpr int SMEM_BYTES = NSTAGE_A * CHUNK_A + NSTAGE_B * STAGE_B;   // 196608
constexpr int PITCH = N + 4;                   // staging row pitch (floats)
constexpr int EPI_WARPS = 8;                   // two warps per TMEM lane quarter, two tiles each
constexpr int THREADS = 64 + 32 * EPI_WARPS;
static_assert(SMEM_BYTES >= EPI_WARPS * 32 * PITCH * 4, "staging must fit in the pipeline buffers");

struct ConvGeom {
    int U, V, B;                 // extent of the fast / slow tile dimension and the batch
    int tiles_u, tiles_v, n_tiles;
    long long su, sv, sb;        // output strides (floats) of u, v, batch; channels are contiguous
    int ku_is_ky;                // 1: u = y (tap row offset moves along u); 0: u = x
    // sparse-tile mode (CTA-pair kernel): only the tiles tile_list[0 .. *n_active) are computed (device-side list and count,
    // built by crb3d_bev_tile_plan); null = all n_tiles tiles in order
    const int* tile_list;
    const int* n_active;
};

// 64B swizzle K-major descriptor: rows of 64 bytes (16 channels of one pixel), 8-row groups `sbo` bytes apart. The
// swizzle is a function of the absolute shared-memory address (16-byte chunk index ^= address bits 7-8), the same the
// TMA unit applied when it wrote the tile, so the start address may sit on ANY row (measured: parity holds for all taps).
__device__ __forceinline__ uint64_t desc_sw64(uint32_t smem_addr, uint32_t sbo_bytes) {
    return (uint64_t)((smem_addr & 0x3FFFF) >> 4) | (1ull << 16) | ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) | (1ull << 46) |
           (4ull << 61);
}

__device__ long long g_conv_trace[1024 * 16];    // experiment bit 2: per-CTA phase stamps (tools/bench_bev.py trace)
__device__ __forceinline__ long long gtime() {
    long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

// tcgen05.ld split into issue and wait, so that the next 32 accumulator columns are in flight while the previous 32
// are converted; the wait names the registers as in/out operands to keep their uses behind it
__device__ __forceinline__ void tmem_ld32_issue(uint32_t taddr, uint32_t* v) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
          "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
          "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld32_wait(uint32_t* v) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]),
                   "+r"(v[8]), "+r"(v[9]), "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]), "+r"(v[15]),
                   "+r"(v[16]), "+r"(v[17]), "+r"(v[18]), "+r"(v[19]), "+r"(v[20]), "+r"(v[21]), "+r"(v[22]), "+r"(v[23]),
                   "+r"(v[24]), "+r"(v[25]), "+r"(v[26]), "+r"(v[27]), "+r"(v[28]), "+r"(v[29]), "+r"(v[30]), "+r"(v[31]));
}

__global__ void __launch_bounds__(THREADS, 1) bev_conv3x3_tc(const __grid_constant__ CUtensorMap amap,
                                                              const float* __restrict__ wpack, int n_chunks,
                                                              const float* __restrict__ bias, int relu, float* __restrict__ out,
                                                              const __grid_constant__ ConvGeom g) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // pointer arithmetic keeps the shared address space (LDS/STS, not generic LD/ST)
    __shared__ uint64_t full_a[NSTAGE_A], empty_a[NSTAGE_A], full_b[NSTAGE_B], empty_b[NSTAGE_B], acc_bar;
    __shared__ uint32_t tmem_base_s;
    __shared__ __align__(16) float bias_s[N];

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
#ifdef CRB3D_CONV_TRACE   // phase trace (variant bit 1): compiled in on request only - the clock reads sit in the MMA issue loop
    const bool tracing = ((relu >> 8) & 2) && blockIdx.x < 1024;
#else
    constexpr bool tracing = false;
#endif
    long long* tr = g_conv_trace + blockIdx.x * 16;
    if (tracing && tid == 0) { tr[0] = gtime(); uint32_t sm; asm("mov.u32 %0, %smid;" : "=r"(sm)); tr[6] = sm; }
    const int tile0 = blockIdx.x * NT;
    const int nh = blockIdx.y;                                  // which 128 output channels
    const float* wsrc = wpack + (size_t)nh * 9 * n_chunks * (STAGE_B / 4);

    if (tid == 0) {
        for (int s = 0; s < NSTAGE_A; ++s) { mbar_init(&full_a[s], 1); mbar_init(&empty_a[s], 1); }
        for (int s = 0; s < NSTAGE_B; ++s) { mbar_init(&full_b[s], 1); mbar_init(&empty_b[s], 1); }
        mbar_init(&acc_bar, 1);
        mbar_fence_init();
        tma_prefetch_desc(&amap);
    }
    if (warp == 1) tmem_alloc<512>(&tmem_base_s);
    if (warp == 3)
        for (int i = lane; i < N; i += 32) bias_s[i] = bias ? __ldg(&bias[nh * N + i]) : 0.0f;
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_s;
    if (tracing && tid == 0) tr[1] = gtime();
    const uint32_t a_smem = smem_u32(smem), b_smem = a_smem + NSTAGE_A * CHUNK_A;

    if (warp == 0) {
        // ================================ weight producer: one 8 KB slice per (chunk, tap), 6 deep ====================
        if (lane == 0) {
            int sb = 0;
            for (int kc = 0; kc < n_chunks; ++kc)
                for (int tap = 0; tap < 9; ++tap, ++sb) {
                    const int stage = sb % NSTAGE_B;
                    if (sb >= NSTAGE_B) {
                        const long long t0 = tracing ? clock64() : 0;
                        mbar_wait(&empty_b[stage], ((sb / NSTAGE_B) - 1) & 1, (CRB3D_K_BEV_CONV << 8) | 4);
                        if (tracing) tr[7] += clock64() - t0;
                    }
                    mbar_expect_tx(&full_b[stage], STAGE_B);
                    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                                 ::"r"(b_smem + stage * STAGE_B), "l"(wsrc + ((size_t)tap * n_chunks + kc) * (STAGE_B / 4)),
                                   "r"(STAGE_B), "r"(smem_u32(&full_b[stage])) : "memory");
                }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ================================ MMA issuer ====================================================================
        // whole warp, converged: only the tcgen05 instructions are predicated on one elected lane (see elect_one)
        const uint32_t idesc = idesc_tf32(128, N);
        const uint64_t desc_a0 = desc_sw64(a_smem, PU * 64), desc_b0 = desc_nosw(b_smem, SLAB_B, 128);
        int sb = 0;
        for (int kc = 0; kc < n_chunks; ++kc) {
            const int abuf = kc % NSTAGE_A;
            long long t0 = tracing ? clock64() : 0;
            mbar_wait(&full_a[abuf], (kc / NSTAGE_A) & 1, (CRB3D_K_BEV_CONV << 8) | 1);
            if (tracing && lane == 0) { tr[9] += clock64() - t0; if (kc == 0) tr[2] = gtime(); }
            for (int tap = 0; tap < 9; ++tap, ++sb) {
                const int stage = sb % NSTAGE_B;
                t0 = tracing ? clock64() : 0;
                mbar_wait(&full_b[stage], (sb / NSTAGE_B) & 1, (CRB3D_K_BEV_CONV << 8) | 3);
                if (tracing && lane == 0) tr[8] += clock64() - t0;
                tc_fence_after();
                const int ky = tap / 3, kx = tap - ky * 3;
                const int ku = g.ku_is_ky ? ky : kx, kv = g.ku_is_ky ? kx : ky;
                // descriptors differ from the base ones only in the 14-bit start-address field (bytes >> 4): the operand of
                // tap (kv, ku) is the halo tile read from pixel row kv * PU + ku on
                const uint64_t da = desc_a0 + (uint64_t)((abuf * CHUNK_A + (kv * PU + ku) * 64) >> 4);
                const uint64_t db = desc_b0 + (uint64_t)((stage * STAGE_B) >> 4);
                const uint32_t first = (kc > 0 || tap > 0) ? 1u : 0u;
                if (elect_one()) {
#pragma unroll
                    for (int t = 0; t < NT; ++t) {
#pragma unroll
                        for (int j = 0; j < KC / 8; ++j)
                            umma_tf32(tmem_base + t * N, da + (uint64_t)((t * TILE_A + j * 32) >> 4),
                                      db + (uint64_t)((j * 2 * SLAB_B) >> 4), idesc, (j > 0) ? 1u : first);
                    }
                    umma_commit(&empty_b[stage]);
                    if (tap == 8) umma_commit(&empty_a[abuf]);
                    if (tap == 8 && kc == n_chunks - 1) umma_commit(&acc_bar);
                }
                __syncwarp();
            }
        }
        if (tracing && lane == 0) tr[3] = gtime();
    } else {
        if (warp == 2) {
            // ============================ activation producer (one lane; its own thread so that the next chunks are requested
            // as soon as their buffer drains, not after the weight slices in between) =====================================
            if (lane == 0) {
                int tu0[NT], tv0[NT], tb[NT];
#pragma unroll
                for (int t = 0; t < NT; ++t) {
                    const int ti = tile0 + t;
                    if (ti < g.n_tiles) {
                        const int per_img = g.tiles_u * g.tiles_v;
                        tb[t] = ti / per_img;
                        const int rem = ti - tb[t] * per_img;
                        tv0[t] = (rem / g.tiles_u) * TV;
                        tu0[t] = (rem % g.tiles_u) * TU;
                    } else { tb[t] = g.B; tu0[t] = 0; tv0[t] = 0; }   // fully out of bounds: the TMA unit writes zeros
                }
                for (int kc = 0; kc < n_chunks; ++kc) {
                    const int abuf = kc % NSTAGE_A;
                    if (kc >= NSTAGE_A) mbar_wait(&empty_a[abuf], ((kc / NSTAGE_A) - 1) & 1, (CRB3D_K_BEV_CONV << 8) | 2);
                    mbar_expect_tx(&full_a[abuf], NT * TILE_A_BYTES);
#pragma unroll
                    for (int t = 0; t < NT; ++t)
                        tma_load_4d(a_smem + abuf * CHUNK_A + t * TILE_A, &amap, kc * KC, tu0[t] - 1, tv0[t] - 1, tb[t], &full_a[abuf]);
                }
            }
            __syncwarp();
        }
        // ================================ epilogue (warps 2..9 -> TMEM lane quarters 2,3,0,1,2,3,0,1) =================
        const int q = warp & 3, half = (warp - 2) >> 2;
        mbar_wait(&acc_bar, 0, (CRB3D_K_BEV_CONV << 8) | 7);
        tc_fence_after();
        if (tracing && tid == 64) tr[4] = gtime();
        float* stage_f = reinterpret_cast<float*>(smem) + (size_t)(warp - 2) * 32 * PITCH;
        const bool no_store = (relu >> 8) & 4;
        const int do_relu = relu & 1, do_round = relu & 2;
        const float4* bias4 = reinterpret_cast<const float4*>(bias_s);
#pragma unroll 1
        for (int t = half * (NT / 2); t < (half + 1) * (NT / 2); ++t) {
            const int ti = tile0 + t;
            if (ti >= g.n_tiles) break;              // uniform per warp
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(t * N);
            uint32_t va[32], vb[32];
            auto convert = [&](const uint32_t* vv, int c0) {
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                    const float4 bq = bias4[(c0 + j) >> 2];
                    float4 w = make_float4(__uint_as_float(vv[j]) + bq.x, __uint_as_float(vv[j + 1]) + bq.y,
                                           __uint_as_float(vv[j + 2]) + bq.z, __uint_as_float(vv[j + 3]) + bq.w);
                    if (do_relu) { w.x = fmaxf(w.x, 0.0f); w.y = fmaxf(w.y, 0.0f); w.z = fmaxf(w.z, 0.0f); w.w = fmaxf(w.w, 0.0f); }
                    if (do_round) { w.x = tf32_rn(w.x); w.y = tf32_rn(w.y); w.z = tf32_rn(w.z); w.w = tf32_rn(w.w); }
                    *reinterpret_cast<float4*>(stage_f + (size_t)lane * PITCH + c0 + j) = w;
                }
            };
            tmem_ld32_issue(taddr, va);
            tmem_ld32_wait(va);
            tmem_ld32_issue(taddr + 32, vb);
            convert(va, 0);
            tmem_ld32_wait(vb);
            tmem_ld32_issue(taddr + 64, va);
            convert(vb, 32);
            tmem_ld32_wait(va);
            tmem_ld32_issue(taddr + 96, vb);
            convert(va, 64);
            tmem_ld32_wait(vb);
            convert(vb, 96);
            __syncwarp();
            // tile row r = q*32 + rr = v_local * 8 + u_local: one 512-byte pixel row per iteration, 16 bytes per lane
            const int per_img = g.tiles_u * g.tiles_v;
            const int b = ti / per_img, rem = ti - b * per_img;
            const int tv = (rem / g.tiles_u) * TV + q * 4, tu = (rem % g.tiles_u) * TU;
            float* obase = out + (long long)b * g.sb + nh * N + lane * 4;
#pragma unroll 8
            for (int rr = 0; rr < 32; ++rr) {
                const int v = tv + (rr >> 3), u = tu + (rr & 7);
                if (u < g.U && v < g.V && !no_store)
                    *reinterpret_cast<float4*>(obase + (long long)u * g.su + (long long)v * g.sv) =
                        *reinterpret_cast<const float4*>(stage_f + (size_t)rr * PITCH + lane * 4);
            }
            __syncwarp();
        }
        if (tracing && tid == 64) tr[10] = gtime();
    }
    tc_fence_before();
    __syncthreads();
    if (tracing && tid == 0) tr[5] = gtime();
    if (warp == 1) tmem_dealloc<512>(tmem_base);
}


// =====================================================================================================================
// CTA-pair variant (cta_group::2, persistent, accumulators double-buffered in TMEM).
// The phase trace of the kernel above shows where a 2-wave launch loses time: ~5 us until the first operands arrive and
// ~8 us of epilogue per wave (the 148 x 256 KB store burst runs at the HBM write rate) with the tensor pipe idle, and a
// B operand that is re-read from shared memory by every M128 MMA (128 B/clk, the whole shared-memory bandwidth). Here
//   * two CTAs of a cluster issue ONE M256 x N128 x K8 MMA: each holds its own 128 pixels (A) and HALF of the weight
//     slice (64 output channels, B) - shared-memory operand traffic per SM drops to 96 B/clk and the weight traffic
//     from L2 per SM halves, which is what makes a 2-tile work item (instead of 4) affordable;
//   * a work item is 2 tiles per CTA = 256 TMEM columns, so the other 256 columns take the next item while 8 epilogue
//     warps drain this one: the stores are spread over the main loop of the next item, the operand pipeline never
//     empties between items, and setup / first-data latency is paid once per CTA.
// Both CTAs run the producers (TMA .cta_group::2 completes on the LEADER's barriers), the leader's warp 1 issues the
// MMAs and multicasts the commits to both CTAs, the epilogue warps of both CTAs arrive on the leader's acc_empty.
namespace pair {

// IT = tiles per CTA per work item (template parameter): 2 by default (a weight stage feeds 2 x 2 tiles); 1 when the layer
// has so few tiles that whole rounds of 2-tile items would leave most of the 74 pairs idle in the last round (block 2 of the
// KITTI BEV backbone: 312 tiles of 100 x 88 maps -> 5 rounds of 1-tile items instead of 3 rounds of 2-tile items)
constexpr int NSA = 4;
constexpr int HALF_B = STAGE_B / 2;             // 4096 B: 64 output channels x 16 input channels of one tap
constexpr int ROW_B = 3 * HALF_B;               // a weight stage = the three taps of one kernel row: 12 MMAs per barrier round trip
constexpr int NSB = 4;
constexpr int EPW = 8;                          // epilogue warps: (tile, TMEM lane quarter)
constexpr int SPITCH = 64 + 4;                  // staging row pitch (floats): 64-column halves
constexpr int STAGING = EPW * 32 * SPITCH * 4;  // 69632 B
constexpr int smem_bytes(int it) { return NSA * it * TILE_A + NSB * ROW_B + STAGING; }   // 217088 for IT = 2
constexpr int NTHREADS = 32 * (3 + EPW);

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
// Default semantics (release at CTA scope), as CUTLASS's ClusterBarrier::arrive(cta_id): what the epilogue hands over is tensor-memory
// state, ordered by the tcgen05 fences on both sides; `.release.cluster` would make the warp drain its global stores first
__device__ __forceinline__ void remote_arrive(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void remote_arrive_release(uint32_t cluster_addr) {      // A/B only (variant bit 4)
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void tma2_load_4d(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2, int c3, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
                 ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(bar) : "memory");
}
__device__ __forceinline__ void tma2_load_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(bar) : "memory");
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

template <int IT>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NTHREADS, 1)
bev_conv3x3_pair_tc(const __grid_constant__ CUtensorMap amap, const __grid_constant__ CUtensorMap wmap, int n_chunks,
                    const float* __restrict__ bias, int relu, float* __restrict__ out, const __grid_constant__ ConvGeom g) {
    constexpr int ITEM_A = IT * TILE_A;             // bytes per (item, chunk) and CTA
    extern __shared__ uint8_t smem_raw[];
    // the dynamic window starts at the same offset in both CTAs; descriptors address both CTAs with one offset
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // pointer arithmetic keeps the shared address space (LDS/STS, not generic LD/ST)
    __shared__ uint64_t full_a[NSA], empty_a[NSA], full_b[NSB], empty_b[NSB], acc_full[2], acc_empty[2];
    __shared__ uint32_t tmem_base_s;
    __shared__ __align__(16) float bias_s[N];

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t rank = cluster_rank();
    const bool leader = rank == 0;
    const int nh = blockIdx.y;
    const int n_clusters = gridDim.x >> 1, cluster_id = blockIdx.x >> 1;
    const int n_tiles = g.n_active ? min(max(__ldg(g.n_active), 0), g.n_tiles) : g.n_tiles;    // the same value in both CTAs of the pair
    const int n_items = (n_tiles + 2 * IT - 1) / (2 * IT);
#ifdef CRB3D_CONV_TRACE
    const bool tracing = ((relu >> 8) & 2) && blockIdx.x < 1024 && blockIdx.y == 0;
#else
    constexpr bool tracing = false;
#endif
    long long* tr = g_conv_trace + blockIdx.x * 16;
    if (tracing && tid == 0) { tr[0] = gtime(); uint32_t sm; asm("mov.u32 %0, %smid;" : "=r"(sm)); tr[6] = sm; }

    if (tid == 0) {
        for (int s = 0; s < NSA; ++s) { mbar_init(&full_a[s], 1); mbar_init(&empty_a[s], 1); }
        for (int s = 0; s < NSB; ++s) { mbar_init(&full_b[s], 1); mbar_init(&empty_b[s], 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(&acc_full[s], 1); mbar_init(&acc_empty[s], 2 * EPW); }
        mbar_fence_init();
        tma_prefetch_desc(&amap);
        tma_prefetch_desc(&wmap);
    }
    // Both CTAs of the pair must be running before the PAIR allocation touches the peer SM's tensor memory: round 1
    // allocated first and synchronised the cluster afterwards, and about one launch in 10^4 (with other graph copies'
    // kernels in flight) never came back from tcgen05.alloc - the only unbounded wait of this kernel (hang
    // root cause; tools/stress_hang.py reproduces it within ~50 repetitions and is clean with this barrier).
    cluster_sync_all();
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "n"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
    }
    if (warp == 3)
        for (int i = lane; i < N; i += 32) bias_s[i] = bias ? __ldg(&bias[nh * N + i]) : 0.0f;
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                          // the peer's barriers are initialised before anything signals them
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_s;
    if (tracing && tid == 0) tr[1] = gtime();
    const uint32_t a_smem = smem_u32(smem), b_smem = a_smem + NSA * ITEM_A;

    if (warp == 0) {
        // ================================ weight producer: this CTA's 64 output channels of every (chunk, tap) =========
        if (lane == 0) {
            int sb = 0;
            long long w_empty = 0;
            const int row0 = nh * 9 * n_chunks * 2;
            for (int item = cluster_id; item < n_items; item += n_clusters)
                for (int kc = 0; kc < n_chunks; ++kc)
                    for (int row = 0; row < 3; ++row, ++sb) {
                        const int stage = sb % NSB;
                        if (sb >= NSB) {
                            const long long t0 = tracing ? clock64() : 0;
                            mbar_wait(&empty_b[stage], ((sb / NSB) - 1) & 1, (CRB3D_K_BEV_CONV_PAIR << 8) | 4);
                            if (tracing) w_empty += clock64() - t0;
                        }
                        if (leader) mbar_expect_tx(&full_b[stage], 2 * ROW_B);
                        const uint32_t bar = leader_addr(&full_b[stage]);
#pragma unroll
                        for (int k = 0; k < 3; ++k)
                            tma2_load_2d(b_smem + stage * ROW_B + k * HALF_B, &wmap, 0,
                                         (row0 + ((row * 3 + k) * n_chunks + kc) * 2 + (int)rank) * 4, bar);
                    }
            if (tracing) tr[7] = w_empty;
        }
        __syncwarp();
    } else if (warp == 1) {
        // ================================ MMA issuer (leader CTA only) ===================================================
        if (leader) {
            const uint32_t idesc = idesc_tf32(256, N);
            const uint64_t desc_a0 = desc_sw64(a_smem, PU * 64), desc_b0 = desc_nosw(b_smem, 64 * 16, 128);
            int sb = 0, ca = 0, it = 0;
            const long long tloop = tracing ? clock64() : 0;
            long long w_a = 0, w_b = 0, w_acc = 0;
            for (int item = cluster_id; item < n_items; item += n_clusters, ++it) {
                const int set = it & 1;
                long long t0 = tracing ? clock64() : 0;
                if (it >= 2) mbar_wait(&acc_empty[set], ((it >> 1) - 1) & 1, (CRB3D_K_BEV_CONV_PAIR << 8) | 6);
                if (tracing) w_acc += clock64() - t0;
                tc_fence_after();
                for (int kc = 0; kc < n_chunks; ++kc, ++ca) {
                    const int abuf = ca % NSA;
                    t0 = tracing ? clock64() : 0;
                    mbar_wait(&full_a[abuf], (ca / NSA) & 1, (CRB3D_K_BEV_CONV_PAIR << 8) | 1);
                    if (tracing) { w_a += clock64() - t0; if (ca == 0 && lane == 0) tr[2] = gtime(); }
                    for (int row = 0; row < 3; ++row, ++sb) {
                        const int stage = sb % NSB;
                        t0 = tracing ? clock64() : 0;
                        mbar_wait(&full_b[stage], (sb / NSB) & 1, (CRB3D_K_BEV_CONV_PAIR << 8) | 3);
                        if (tracing) w_b += clock64() - t0;
                        tc_fence_after();
                        if (elect_one()) {
#pragma unroll
                            for (int k = 0; k < 3; ++k) {
                                // tap (ky, kx) = (row, k): the halo tile read from pixel row kv * PU + ku on
                                const int ku = g.ku_is_ky ? row : k, kv = g.ku_is_ky ? k : row;
                                const uint64_t da = desc_a0 + (uint64_t)((abuf * ITEM_A + (kv * PU + ku) * 64) >> 4);
                                const uint64_t db = desc_b0 + (uint64_t)((stage * ROW_B + k * HALF_B) >> 4);
                                const uint32_t first = (kc > 0 || row > 0 || k > 0) ? 1u : 0u;
#pragma unroll
                                for (int t = 0; t < IT; ++t) {
#pragma unroll
                                    for (int j = 0; j < KC / 8; ++j)
                                        umma2_tf32(tmem_base + set * (IT * N) + t * N, da + (uint64_t)((t * TILE_A + j * 32) >> 4),
                                                   db + (uint64_t)((j * 2 * 64 * 16) >> 4), idesc, (j > 0) ? 1u : first);
                                }
                            }
                            umma2_commit(&empty_b[stage]);
                            if (row == 2) umma2_commit(&empty_a[abuf]);
                            if (row == 2 && kc == n_chunks - 1) umma2_commit(&acc_full[set]);
                        }
                        __syncwarp();
                    }
                }
            }
            if (tracing && lane == 0) { tr[3] = gtime(); tr[13] = clock64() - tloop; tr[14] = it; tr[9] = w_a; tr[8] = w_b; tr[11] = w_acc; }
        }
    } else if (warp == 2) {
        // ================================ activation producer: this CTA's two tiles of every item ========================
        if (lane == 0) {
            int ca = 0;
            const int per_img = g.tiles_u * g.tiles_v;
            for (int item = cluster_id; item < n_items; item += n_clusters) {
                int tu0[IT], tv0[IT], tb[IT];
#pragma unroll
                for (int t = 0; t < IT; ++t) {
                    const int ti = (item * 2 + (int)rank) * IT + t;
                    if (ti < n_tiles) {
                        const int tt = g.tile_list ? __ldg(g.tile_list + ti) : ti;
                        tb[t] = tt / per_img;
                        const int rem = tt - tb[t] * per_img;
                        tv0[t] = (rem / g.tiles_u) * TV;
                        tu0[t] = (rem % g.tiles_u) * TU;
                    } else { tb[t] = g.B; tu0[t] = 0; tv0[t] = 0; }   // fully out of bounds: the TMA unit writes zeros
                }
                for (int kc = 0; kc < n_chunks; ++kc, ++ca) {
                    const int abuf = ca % NSA;
                    if (ca >= NSA) mbar_wait(&empty_a[abuf], ((ca / NSA) - 1) & 1, (CRB3D_K_BEV_CONV_PAIR << 8) | 2);
                    if (leader) mbar_expect_tx(&full_a[abuf], 2 * IT * TILE_A_BYTES);
                    const uint32_t bar = leader_addr(&full_a[abuf]);
#pragma unroll
                    for (int t = 0; t < IT; ++t)
                        tma2_load_4d(a_smem + abuf * ITEM_A + t * TILE_A, &amap, kc * KC, tu0[t] - 1, tv0[t] - 1, tb[t], bar);
                }
            }
        }
        __syncwarp();
    } else {
        // ================================ epilogue: warp -> (tile = (warp-3)/4, TMEM lane quarter = warp%4) ==============
        const int q = warp & 3, t = (warp - 3) >> 2;
        float* stage_f = reinterpret_cast<float*>(smem + NSA * ITEM_A + NSB * ROW_B) + (size_t)(warp - 3) * 32 * SPITCH;
        const int do_relu = relu & 1, do_round = relu & 2;
        const float4* bias4 = reinterpret_cast<const float4*>(bias_s);
        const int per_img = g.tiles_u * g.tiles_v;
        const uint32_t acc_empty_leader[2] = {leader_addr(&acc_empty[0]), leader_addr(&acc_empty[1])};
        int it = 0;
        long long w_full = 0;
        for (int item = cluster_id; item < n_items; item += n_clusters, ++it) {
            const int set = it & 1;
            const long long t0 = tracing ? clock64() : 0;
            mbar_wait(&acc_full[set], (it >> 1) & 1, (CRB3D_K_BEV_CONV_PAIR << 8) | 5);
            if (tracing) w_full += clock64() - t0;
            tc_fence_after();
            if (t >= IT) {                       // 1-tile items: the second tile's four warps only take part in the handshake
                tc_fence_before();
                __syncwarp();
                if (lane == 0) remote_arrive(acc_empty_leader[set]);
                continue;
            }
            const int ti = (item * 2 + (int)rank) * IT + t;
            const bool live = ti < n_tiles;          // uniform per warp
            const int tt = live ? (g.tile_list ? __ldg(g.tile_list + ti) : ti) : 0;
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(set * (IT * N) + t * N);
            const int b = tt / per_img, rem = tt - b * per_img;
            const int tv = (rem / g.tiles_u) * TV + q * 4, tu = (rem % g.tiles_u) * TU;
            float* obase = out + (long long)b * g.sb + nh * N + (lane & 15) * 4;
            uint32_t va[32], vb[32];
            auto convert = [&](const uint32_t* vv, int c0, int s0) {
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                    const float4 bq = bias4[(c0 + j) >> 2];
                    float4 w = make_float4(__uint_as_float(vv[j]) + bq.x, __uint_as_float(vv[j + 1]) + bq.y,
                                           __uint_as_float(vv[j + 2]) + bq.z, __uint_as_float(vv[j + 3]) + bq.w);
                    if (do_relu) { w.x = fmaxf(w.x, 0.0f); w.y = fmaxf(w.y, 0.0f); w.z = fmaxf(w.z, 0.0f); w.w = fmaxf(w.w, 0.0f); }
                    if (do_round) { w.x = tf32_rn(w.x); w.y = tf32_rn(w.y); w.z = tf32_rn(w.z); w.w = tf32_rn(w.w); }
                    *reinterpret_cast<float4*>(stage_f + (size_t)lane * SPITCH + s0 + j) = w;
                }
            };
            auto store_half = [&](int c0) {          // 32 rows x 64 columns: two 256-byte row pieces per instruction
                __syncwarp();
#pragma unroll 8
                for (int r2 = 0; r2 < 32; r2 += 2) {
                    const int rr = r2 + (lane >> 4);
                    const int v = tv + (rr >> 3), u = tu + (rr & 7);
                    if (live && u < g.U && v < g.V)
                        *reinterpret_cast<float4*>(obase + (long long)u * g.su + (long long)v * g.sv + c0) =
                            *reinterpret_cast<const float4*>(stage_f + (size_t)rr * SPITCH + (lane & 15) * 4);
                }
                __syncwarp();
            };
            tmem_ld32_issue(taddr, va);
            tmem_ld32_wait(va);
            tmem_ld32_issue(taddr + 32, vb);
            convert(va, 0, 0);
            tmem_ld32_wait(vb);
            tmem_ld32_issue(taddr + 64, va);
            convert(vb, 32, 32);
            store_half(0);
            tmem_ld32_wait(va);
            tmem_ld32_issue(taddr + 96, vb);
            convert(va, 64, 0);
            tmem_ld32_wait(vb);
            // every column of this warp's accumulator slice is in registers: hand the TMEM set back to the MMA issuer
            tc_fence_before();
            __syncwarp();
            if (lane == 0) { if ((relu >> 8) & 16) remote_arrive_release(acc_empty_leader[set]); else remote_arrive(acc_empty_leader[set]); }
            convert(vb, 96, 32);
            store_half(64);
        }
        if (tracing && warp == 3 && lane == 0) { tr[10] = gtime(); tr[12] = w_full; }
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                          // the peer may still read this CTA's shared memory / signal its barriers
    if (tracing && tid == 0) tr[5] = gtime();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512));
}

}  // namespace pair

}  // namespace

// debug: copies the phase stamps of the last variant-4 launch (16 int64 per CTA) to the host
extern "C" int crb3d_bev_conv3x3_trace(long long* host_out, int n_ctas) {
    if (!host_out || n_ctas <= 0 || n_ctas > 1024) return CRB3D_ERR_ARG;
    CRB3D_CUDA(cudaMemcpyFromSymbol(host_out, g_conv_trace, sizeof(long long) * 16 * n_ctas));
    long long* zero = new long long[16 * 1024]();
    cudaMemcpyToSymbol(g_conv_trace, zero, sizeof(long long) * 16 * 1024);
    delete[] zero;
    return CRB3D_OK;
}

namespace {

// tile orientation of an H x W map: u (8 pixels per tile) along y or along x, whichever pads less
bool tile_u_is_y(int H, int W) {
    return (H % TU == 0) || (W % TU != 0 && (crb3d_divup(H, TU) * TU - H) * W <= (crb3d_divup(W, TU) * TU - W) * H);
}

// ---------------------------------------------------------------------------------------------------------------------
// Sparse-tile plan of a BEV block. The input of the 2-D backbone is the dense() of a sparse tensor: zero except at the occupied
// cells (13-17 % of a KITTI-synthetic map). Far from every occupied cell and from the image border a stack of 3x3 convs
// produces, layer after layer, ONE constant vector per layer (conv of a constant field + bias + ReLU). A 128-pixel output tile
// of layer l (0-based) is constant iff the tile grown by l + 1 pixels lies inside the image and holds no occupied cell; such
// tiles are filled with the layer's constant (computed once per plan by running the very kernel on a constant image, so the
// bits are the ones the dense computation would produce) and only the others go through the tensor cores.
__global__ void __launch_bounds__(256) occ_scatter_kernel(const int* __restrict__ coords, int n, const int* __restrict__ n_dev, int B, int H, int W,
                                                          unsigned char* __restrict__ occ) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || (n_dev && i >= *n_dev)) return;
    const int4 c = __ldg(reinterpret_cast<const int4*>(coords) + i);     // b, z, y, x
    if (c.x >= 0 && c.x < B && c.z >= 0 && c.z < H && c.w >= 0 && c.w < W) occ[((size_t)c.x * H + c.z) * W + c.w] = 1;
}
// summed-area table sat[b][y + 1][x + 1] = number of occupied cells in [0, y] x [0, x]; row 0 / column 0 stay zero
__global__ void __launch_bounds__(256) sat_rows_kernel(const unsigned char* __restrict__ occ, int BH, int H, int W, int* __restrict__ sat) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= BH) return;
    const int b = warp / H, y = warp - b * H;
    const unsigned char* row = occ + (size_t)warp * W;
    int* dst = sat + ((size_t)b * (H + 1) + y + 1) * (W + 1) + 1;
    int carry = 0;
    for (int x0 = 0; x0 < W; x0 += 32) {
        const int x = x0 + lane;
        int v = x < W ? row[x] : 0;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, v, o);
            if (lane >= o) v += t;
        }
        if (x < W) dst[x] = v + carry;
        carry += __shfl_sync(0xffffffffu, v, 31);
    }
}
// column pass: one warp per (image, column), 32 rows per step (warp scan + carry) - a thread per column walking all H rows is a
// chain of H dependent loads (92 us for a 16 x 200 x 176 map)
__global__ void __launch_bounds__(256) sat_cols_kernel(int B, int H, int W, int* __restrict__ sat) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= B * W) return;
    const int b = warp / W, x = warp - b * W;
    int* col = sat + (size_t)b * (H + 1) * (W + 1) + x + 1;
    int carry = 0;
    for (int y0 = 1; y0 <= H; y0 += 32) {
        const int y = y0 + lane;
        int v = y <= H ? col[(size_t)y * (W + 1)] : 0;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, v, o);
            if (lane >= o) v += t;
        }
        if (y <= H) col[(size_t)y * (W + 1)] = v + carry;
        carry += __shfl_sync(0xffffffffu, v, 31);
    }
}
// one block per level: flags[level][tile] (1 = compute) and the ascending list of the tiles to compute + their count
__global__ void __launch_bounds__(1024) tile_classify_kernel(const int* __restrict__ sat, int B, int H, int W, int u_is_y, int tiles_u, int tiles_v,
                                                             int n_tiles, int* __restrict__ lists, int* __restrict__ counts,
                                                             unsigned char* __restrict__ flags) {
    __shared__ int sm[33];
    const int level = blockIdx.x, grow = level + 1;
    const int per_img = tiles_u * tiles_v;
    int base = 0;
    for (int t0 = 0; t0 < n_tiles; t0 += 1024) {
        const int t = t0 + threadIdx.x;
        int active = 0;
        if (t < n_tiles) {
            const int b = t / per_img, rem = t - b * per_img;
            const int u0 = (rem % tiles_u) * TU, v0 = (rem / tiles_u) * TV;
            const int y0 = u_is_y ? u0 : v0, x0 = u_is_y ? v0 : u0;
            const int y1 = min(y0 + (u_is_y ? TU : TV), H), x1 = min(x0 + (u_is_y ? TV : TU), W);   // exclusive ends, clipped to the image
            const int gy0 = y0 - grow, gx0 = x0 - grow, gy1 = y1 + grow, gx1 = x1 + grow;
            if (gy0 < 0 || gx0 < 0 || gy1 > H || gx1 > W || y0 + (u_is_y ? TU : TV) > H || x0 + (u_is_y ? TV : TU) > W) active = 1;   // the padding is in reach
            else {
                const int* s = sat + (size_t)b * (H + 1) * (W + 1);
                const int cnt = s[(size_t)gy1 * (W + 1) + gx1] - s[(size_t)gy0 * (W + 1) + gx1] - s[(size_t)gy1 * (W + 1) + gx0] +
                                s[(size_t)gy0 * (W + 1) + gx0];
                active = cnt > 0;
            }
            flags[(size_t)level * n_tiles + t] = (unsigned char)active;
        }
        int tot;
        const int pos = block_excl_scan(active, sm, &tot);
        if (active) lists[(size_t)level * n_tiles + base + pos] = t;
        base += tot;
    }
    if (threadIdx.x == 0) counts[level] = base;
}
// which constant tiles actually have to be written: all of them at the last level (the deblock reads the whole map), at the
// levels before only those a computed tile of the NEXT level can see through its 1-pixel halo (its 8 neighbours)
__global__ void __launch_bounds__(256) tile_fill_flags_kernel(const unsigned char* __restrict__ flags, int n_levels, int tiles_u, int tiles_v,
                                                              int n_tiles, unsigned char* __restrict__ fill_flags) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x, level = blockIdx.y;
    if (t >= n_tiles) return;
    unsigned char f = 0;
    if (!flags[(size_t)level * n_tiles + t]) {
        if (level == n_levels - 1) f = 1;
        else {
            const int per_img = tiles_u * tiles_v, b = t / per_img, rem = t - b * per_img, iv = rem / tiles_u, iu = rem % tiles_u;
            const unsigned char* nxt = flags + (size_t)(level + 1) * n_tiles + (size_t)b * per_img;
            for (int dv = -1; dv <= 1 && !f; ++dv)
                for (int du = -1; du <= 1; ++du) {
                    const int v = iv + dv, u = iu + du;
                    if (v >= 0 && v < tiles_v && u >= 0 && u < tiles_u && nxt[v * tiles_u + u]) { f = 1; break; }
                }
        }
    }
    fill_flags[(size_t)level * n_tiles + t] = f;
}
// constant tiles of a layer: every pixel gets the layer's constant vector
__global__ void __launch_bounds__(256) tile_fill_kernel(const unsigned char* __restrict__ fill_flags, const float* __restrict__ fill, int cout, ConvGeom g,
                                                        float* __restrict__ out) {
    const int t = blockIdx.x;
    if (!fill_flags[t]) return;
    const int per_img = g.tiles_u * g.tiles_v;
    const int b = t / per_img, rem = t - b * per_img;
    const int u0 = (rem % g.tiles_u) * TU, v0 = (rem / g.tiles_u) * TV;
    const int c4 = cout >> 2;
    for (int e = threadIdx.x; e < TU * TV * c4; e += blockDim.x) {
        const int pix = e / c4, c = (e - pix * c4) * 4;
        const int u = u0 + (pix % TU), v = v0 + (pix / TU);
        if (u >= g.U || v >= g.V) continue;
        *reinterpret_cast<float4*>(out + (long long)b * g.sb + (long long)u * g.su + (long long)v * g.sv + c) =
            __ldg(reinterpret_cast<const float4*>(fill + c));
    }
}

int conv3x3_impl(const float* in, int B, int H, int W, int cin, const float* wpack, int cout, const float* bias, int relu, float* out,
                 const int* tile_list, const int* n_active, const unsigned char* tile_flags, const float* fill, cudaStream_t stream);

}  // namespace

extern "C" int crb3d_bev_conv3x3_tf32(const float* in, int B, int H, int W, int cin, const float* wpack, int cout,
                                      const float* bias, int relu, float* out, cudaStream_t stream) {
    return conv3x3_impl(in, B, H, W, cin, wpack, cout, bias, relu, out, nullptr, nullptr, nullptr, nullptr, stream);
}

// number of 128-pixel tiles the halo-tile kernel cuts a (B, H, W) map into (the row stride of the plan arrays below)
extern "C" int crb3d_bev_conv3x3_num_tiles(int B, int H, int W, int* n_tiles) {
    if (!n_tiles || B <= 0 || H <= 0 || W <= 0) return CRB3D_ERR_ARG;
    const bool u_is_y = tile_u_is_y(H, W);
    *n_tiles = B * (int)crb3d_divup(u_is_y ? H : W, TU) * (int)crb3d_divup(u_is_y ? W : H, TV);
    return CRB3D_OK;
}

extern "C" int crb3d_bev_tile_plan_workspace_bytes(int B, int H, int W, size_t* bytes) {
    if (!bytes || B <= 0 || H <= 0 || W <= 0) return CRB3D_ERR_ARG;
    *bytes = crb3d_align((size_t)B * H * W) + crb3d_align(sizeof(int) * (size_t)B * (H + 1) * (W + 1));
    return CRB3D_OK;
}

// coords (n,4) int32 [b,z,y,x] = the rows of the sparse tensor whose dense() is the block's input (n_dev nullable device count);
// n_levels = number of stacked 3x3 stride-1 layers. Outputs (device): lists [n_levels][n_tiles] int32, counts [n_levels] int32,
// flags [n_levels][n_tiles] uint8 (1 = the tile must be computed at that level), fill_flags [n_levels][n_tiles] uint8 (1 = the tile
// is constant AND somebody reads it: the next level's computed tiles through their halo, or - last level - the consumer of the map).
extern "C" int crb3d_bev_tile_plan(const int* coords, int n, const int* n_dev, int B, int H, int W, int n_levels, int* lists, int* counts,
                                   unsigned char* flags, unsigned char* fill_flags, void* ws, size_t ws_bytes, cudaStream_t stream) {
    if (n < 0 || B <= 0 || H <= 0 || W <= 0 || n_levels <= 0 || !lists || !counts || !flags || !fill_flags || (n > 0 && !coords))
        return CRB3D_ERR_ARG;
    WsCursor c(ws, ws_bytes);
    unsigned char* occ = c.take<unsigned char>((size_t)B * H * W);
    int* sat = c.take<int>((size_t)B * (H + 1) * (W + 1));
    if (!c.ok) return CRB3D_ERR_WORKSPACE;
    CRB3D_CUDA(cudaMemsetAsync(occ, 0, (size_t)B * H * W, stream));
    CRB3D_CUDA(cudaMemsetAsync(sat, 0, sizeof(int) * (size_t)B * (H + 1) * (W + 1), stream));
    if (n > 0) occ_scatter_kernel<<<(unsigned)crb3d_divup(n, 256), 256, 0, stream>>>(coords, n, n_dev, B, H, W, occ);
    sat_rows_kernel<<<(unsigned)crb3d_divup((int64_t)B * H * 32, 256), 256, 0, stream>>>(occ, B * H, H, W, sat);
    sat_cols_kernel<<<(unsigned)crb3d_divup((int64_t)B * W * 32, 256), 256, 0, stream>>>(B, H, W, sat);
    const bool u_is_y = tile_u_is_y(H, W);
    const int tiles_u = (int)crb3d_divup(u_is_y ? H : W, TU), tiles_v = (int)crb3d_divup(u_is_y ? W : H, TV);
    const int n_tiles = B * tiles_u * tiles_v;
    tile_classify_kernel<<<(unsigned)n_levels, 1024, 0, stream>>>(sat, B, H, W, u_is_y ? 1 : 0, tiles_u, tiles_v, n_tiles, lists, counts, flags);
    tile_fill_flags_kernel<<<dim3((unsigned)crb3d_divup(n_tiles, 256), (unsigned)n_levels), 256, 0, stream>>>(flags, n_levels, tiles_u, tiles_v,
                                                                                                           n_tiles, fill_flags);
    CRB3D_CHECK_LAUNCH();
    return CRB3D_OK;
}

// crb3d_bev_conv3x3_tf32 over the tiles tile_list[0 .. *n_active) only (CTA-pair kernel); the tiles with tile_fill != 0 are filled
// with `fill` (C_out floats, device): one level of a crb3d_bev_tile_plan (lists / counts / fill_flags rows of that level).
extern "C" int crb3d_bev_conv3x3_tf32_tiles(const float* in, int B, int H, int W, int cin, const float* wpack, int cout, const float* bias,
                                            int relu, float* out, const int* tile_list, const int* n_active,
                                            const unsigned char* tile_flags, const float* fill, cudaStream_t stream) {
    if (!tile_list || !n_active || !tile_flags || !fill || ((relu >> 8) & 1)) return CRB3D_ERR_ARG;
    return conv3x3_impl(in, B, H, W, cin, wpack, cout, bias, relu, out, tile_list, n_active, tile_flags, fill, stream);
}

namespace {
int conv3x3_impl(const float* in, int B, int H, int W, int cin, const float* wpack, int cout, const float* bias, int relu, float* out,
                 const int* tile_list, const int* n_active, const unsigned char* tile_flags, const float* fill, cudaStream_t stream) {
    if (!in || !wpack || !out || B <= 0 || H <= 0 || W <= 0 || cin <= 0 || cout <= 0) return CRB3D_ERR_ARG;
    if (cin % KC != 0 || cout % N != 0) return CRB3D_ERR_UNSUPPORTED;
    ConvGeom g;
    g.tile_list = tile_list;
    g.n_active = n_active;
    const bool u_is_y = tile_u_is_y(H, W);
    g.ku_is_ky = u_is_y ? 1 : 0;
    g.U = u_is_y ? H : W;
    g.V = u_is_y ? W : H;
    g.B = B;
    g.tiles_u = (int)crb3d_divup(g.U, TU);
    g.tiles_v = (int)crb3d_divup(g.V, TV);
    g.n_tiles = B * g.tiles_u * g.tiles_v;
    const long long sy = (long long)W * cout, sx = cout;
    g.su = u_is_y ? sy : sx;
    g.sv = u_is_y ? sx : sy;
    g.sb = (long long)H * W * cout;
    CUtensorMap amap;
    {
        const uint64_t ysb = (uint64_t)W * cin * 4, xsb = (uint64_t)cin * 4;
        const uint64_t dims[4] = {(uint64_t)cin, (uint64_t)g.U, (uint64_t)g.V, (uint64_t)B};
        const uint64_t strides[3] = {u_is_y ? ysb : xsb, u_is_y ? xsb : ysb, (uint64_t)H * W * cin * 4};
        const uint32_t box[4] = {KC, PU, PV, 1};
        int rc = make_map_f32(&amap, in, 4, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_64B);
        if (rc) return rc;
    }
    static bool attr_set[CRB3D_MAX_DEVICES] = {};   // the attribute is per function per device
    const int dev = crb3d_current_device();
    if (!attr_set[dev]) {
        CRB3D_CUDA(cudaFuncSetAttribute(bev_conv3x3_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES + 1024));
        CRB3D_CUDA(cudaFuncSetAttribute(pair::bev_conv3x3_pair_tc<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, pair::smem_bytes(2) + 1024));
        CRB3D_CUDA(cudaFuncSetAttribute(pair::bev_conv3x3_pair_tc<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, pair::smem_bytes(1) + 1024));
        attr_set[dev] = true;
    }
    if (!((relu >> 8) & 1)) {
        // CTA-pair kernel (default); wpack is the SPLIT layout [C_out/128][tap][C_in/16][half][4 slabs][64 co][4 ci]
        CUtensorMap wmap;
        const uint64_t wrows = (uint64_t)(cout / N) * 9 * (cin / KC) * 2 * 4;
        const uint64_t wdims[2] = {256, wrows};
        const uint64_t wstr[1] = {1024};
        const uint32_t wbox[2] = {256, 4};
        int rc = make_map_f32(&wmap, wpack, 2, wdims, wstr, wbox, CU_TENSOR_MAP_SWIZZLE_NONE);
        if (rc) return rc;
        const int ny = cout / N;
        int n_clusters = crb3d_num_sms() / 2 / ny;
        if (n_clusters < 1) n_clusters = 1;
        // item size: rounds(IT) * IT = time in units of one tile per CTA. A tile of a 2-tile item costs ~0.88 of a tile of a 1-tile
        // item (the weight stage feeds two tiles: 217 vs 232 us at 16 x 100 x 88, 256 -> 256, i.e. 12.1 vs 13.7 us per tile round),
        // so 1-tile items only when their finer rounds save more than that
        const long long items2 = crb3d_divup(g.n_tiles, 4), items1 = crb3d_divup(g.n_tiles, 2);
        const long long cost2 = crb3d_divup(items2, n_clusters) * 2 * 88, cost1 = crb3d_divup(items1, n_clusters) * 100;
        const int it_sel = (((relu >> 8) & 4) || cost1 < cost2) && !((relu >> 8) & 8) ? 1 : 2;
        const int n_items = (int)(it_sel == 1 ? items1 : items2);
        if (n_clusters > n_items) n_clusters = n_items;
        if (it_sel == 1)
            pair::bev_conv3x3_pair_tc<1><<<dim3((unsigned)(2 * n_clusters), (unsigned)ny), pair::NTHREADS, pair::smem_bytes(1) + 1024, stream>>>(
                amap, wmap, cin / KC, bias, relu, out, g);
        else
            pair::bev_conv3x3_pair_tc<2><<<dim3((unsigned)(2 * n_clusters), (unsigned)ny), pair::NTHREADS, pair::smem_bytes(2) + 1024, stream>>>(
                amap, wmap, cin / KC, bias, relu, out, g);
        if (tile_flags) tile_fill_kernel<<<(unsigned)g.n_tiles, 256, 0, stream>>>(tile_flags, fill, cout, g, out);
        CRB3D_CHECK_LAUNCH();
        return CRB3D_OK;
    }
    bev_conv3x3_tc<<<dim3((unsigned)crb3d_divup(g.n_tiles, NT), (unsigned)(cout / N)), THREADS, SMEM_BYTES + 1024, stream>>>(
        amap, wpack, cin / KC, bias, relu, out, g);
    CRB3D_CHECK_LAUNCH();
    return CRB3D_OK;
}
}  // namespace

CRB3D_DIAG_DEFINE_SETTER(bev_conv)
