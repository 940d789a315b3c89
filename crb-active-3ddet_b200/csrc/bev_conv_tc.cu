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
//   * the (8+2) x (16+2) input halo of a 128-pixel output tile is loaded ONCE per 16-channel chunk, by one 5-D TMA box
//     whose out-of-bounds rows/columns are zero-filled by the TMA unit (that IS the conv padding), into the NO-SWIZZLE
//     K-major layout [4-channel slab][v][u][4 floats]: every pixel row is 16 bytes, so the operand of tap (kv,ku) is the
//     same tile read through a descriptor whose start address is advanced by (kv*10 + ku)*16 bytes (SBO = one v step,
//     LBO = one slab). Activation traffic drops 9 x 128 / 180 = 6.4x;
//   * a CTA owns FOUR output tiles (4 x 128 TMEM columns = the whole TMEM), so each 8 KB weight slice (one tap, 16 input
//     channels, 128 output channels - pre-packed on the host into the same slab layout, fetched with one bulk copy) feeds
//     eight M128 x N128 x K8 MMAs: weight traffic per flop drops 4x. Total operand feed ~24 B/clk/SM.
// Warp roles: warp 0 / one lane = TMA producer (activation chunks double-buffered, weight slices 6 deep), warp 1 / one
// lane = MMA issuer, warps 2-5 = epilogue (tcgen05.ld, bias/ReLU, 512-byte pixel rows staged in the drained pipeline
// buffers and stored as whole lines).
#include "tc_common.cuh"

namespace {

using namespace tc;

constexpr int TU = 8, TV = 16;                 // output tile: 8 pixels along u (fast), 16 along v
constexpr int PU = TU + 2, PV = TV + 2;        // halo tile
constexpr int KC = 16;                         // input channels per chunk = 4 slabs of 4 floats = 2 MMAs (K = 8)
constexpr int NT = 4;                          // output tiles per CTA
constexpr int N = 128;                         // output channels per CTA
constexpr int SLAB_A = PU * PV * 16;           // 2880 B
constexpr int TILE_A = (KC / 4) * SLAB_A;      // 11520 B per (tile, chunk)
constexpr int CHUNK_A = NT * TILE_A;           // 46080 B
constexpr int SLAB_B = N * 16;                 // 2048 B
constexpr int STAGE_B = (KC / 4) * SLAB_B;     // 8192 B per (tap, chunk)
constexpr int NSTAGE_B = 6;
constexpr int SMEM_BYTES = 2 * CHUNK_A + NSTAGE_B * STAGE_B;   // 141312
constexpr int PITCH = N + 4;                   // staging row pitch (floats)
static_assert(SMEM_BYTES >= 128 * PITCH * 4, "staging must fit in the pipeline buffers");

struct ConvGeom {
    int U, V, B;                 // extent of the fast / slow tile dimension and the batch
    int tiles_u, tiles_v, n_tiles;
    long long su, sv, sb;        // output strides (floats) of u, v, batch; channels are contiguous
    int ku_is_ky;                // 1: u = y (tap row offset moves along u); 0: u = x
};

__global__ void __launch_bounds__(192, 1) bev_conv3x3_tc(const __grid_constant__ CUtensorMap amap,
                                                          const float* __restrict__ wpack, int n_chunks,
                                                          const float* __restrict__ bias, int relu, float* __restrict__ out,
                                                          const __grid_constant__ ConvGeom g) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t full_a[2], empty_a[2], full_b[NSTAGE_B], empty_b[NSTAGE_B], acc_bar;
    __shared__ uint32_t tmem_base_s;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int tile0 = blockIdx.x * NT;
    const int nh = blockIdx.y;                                  // which 128 output channels
    const float* wsrc = wpack + (size_t)nh * 9 * n_chunks * (STAGE_B / 4);

    if (tid == 0) {
        for (int s = 0; s < 2; ++s) { mbar_init(&full_a[s], 1); mbar_init(&empty_a[s], 1); }
        for (int s = 0; s < NSTAGE_B; ++s) { mbar_init(&full_b[s], 1); mbar_init(&empty_b[s], 1); }
        mbar_init(&acc_bar, 1);
        mbar_fence_init();
        tma_prefetch_desc(&amap);
    }
    if (warp == 1) tmem_alloc<512>(&tmem_base_s);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_s;
    const uint32_t a_smem = smem_u32(smem), b_smem = a_smem + 2 * CHUNK_A;

    if (warp == 0) {
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
            int sb = 0;
            for (int kc = 0; kc < n_chunks; ++kc) {
                const int buf = kc & 1;
                if (kc >= 2) mbar_wait(&empty_a[buf], ((kc >> 1) - 1) & 1);
                mbar_expect_tx(&full_a[buf], CHUNK_A);
#pragma unroll
                for (int t = 0; t < NT; ++t)
                    tma_load_5d(a_smem + buf * CHUNK_A + t * TILE_A, &amap, 0, tu0[t] - 1, tv0[t] - 1, kc * (KC / 4), tb[t],
                                &full_a[buf]);
                for (int tap = 0; tap < 9; ++tap, ++sb) {
                    const int stage = sb % NSTAGE_B;
                    if (sb >= NSTAGE_B) mbar_wait(&empty_b[stage], ((sb / NSTAGE_B) - 1) & 1);
                    mbar_expect_tx(&full_b[stage], STAGE_B);
                    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                                 ::"r"(b_smem + stage * STAGE_B), "l"(wsrc + ((size_t)tap * n_chunks + kc) * (STAGE_B / 4)),
                                   "r"(STAGE_B), "r"(smem_u32(&full_b[stage])) : "memory");
                }
            }
        }
    } else if (warp == 1) {
        // whole warp, converged: only the tcgen05 instructions are predicated on one elected lane (see elect_one)
        const uint32_t idesc = idesc_tf32(128, N);
        const uint64_t desc_a0 = desc_nosw(a_smem, SLAB_A, PU * 16), desc_b0 = desc_nosw(b_smem, SLAB_B, 128);
        int sb = 0;
        for (int kc = 0; kc < n_chunks; ++kc) {
            const int buf = kc & 1;
            mbar_wait(&full_a[buf], (kc >> 1) & 1);
            for (int tap = 0; tap < 9; ++tap, ++sb) {
                const int stage = sb % NSTAGE_B;
                mbar_wait(&full_b[stage], (sb / NSTAGE_B) & 1);
                tc_fence_after();
                const int ky = tap / 3, kx = tap - ky * 3;
                const int ku = g.ku_is_ky ? ky : kx, kv = g.ku_is_ky ? kx : ky;
                // descriptors differ from the base ones only in the 14-bit start-address field (bytes >> 4)
                const uint64_t da = desc_a0 + (uint64_t)((buf * CHUNK_A + (kv * PU + ku) * 16) >> 4);
                const uint64_t db = desc_b0 + (uint64_t)((stage * STAGE_B) >> 4);
                const uint32_t first = (kc > 0 || tap > 0) ? 1u : 0u;
                if (elect_one()) {
#pragma unroll
                    for (int t = 0; t < NT; ++t) {
#pragma unroll
                        for (int j = 0; j < KC / 8; ++j)
                            umma_tf32(tmem_base + t * N, da + (uint64_t)((t * TILE_A + j * 2 * SLAB_A) >> 4),
                                      db + (uint64_t)((j * 2 * SLAB_B) >> 4), idesc, (j > 0) ? 1u : first);
                    }
                    umma_commit(&empty_b[stage]);
                    if (tap == 8) umma_commit(&empty_a[buf]);
                    if (tap == 8 && kc == n_chunks - 1) umma_commit(&acc_bar);
                }
                __syncwarp();
            }
        }
    } else {
        // ================================ epilogue (warps 2..5 -> TMEM lane quarters 2,3,0,1) =========================
        const int q = warp & 3;
        mbar_wait(&acc_bar, 0);
        tc_fence_after();
        float* stage_f = reinterpret_cast<float*>(smem) + (size_t)q * 32 * PITCH;
        const int r = q * 32 + lane;                 // tile row of this thread: r = v_local * 8 + u_local
        const int vl = r >> 3, ul = r & 7;
        const float* bias_n = bias ? bias + nh * N : nullptr;
#pragma unroll 1
        for (int t = 0; t < NT; ++t) {
            const int ti = tile0 + t;
            if (ti >= g.n_tiles) break;              // uniform per CTA
            const int per_img = g.tiles_u * g.tiles_v;
            const int b = ti / per_img, rem = ti - b * per_img;
            const int v = (rem / g.tiles_u) * TV + vl, u = (rem % g.tiles_u) * TU + ul;
            const long long orow = (u < g.U && v < g.V) ? (long long)b * g.sb + (long long)u * g.su + (long long)v * g.sv + nh * N : -1;
#pragma unroll 1
            for (int c0 = 0; c0 < N; c0 += 32) {
                uint32_t vv[32];
                tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(t * N + c0), vv);
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                    float4 w;
                    float* wp = reinterpret_cast<float*>(&w);
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        float x = __uint_as_float(vv[j + e]);
                        if (bias_n) x += __ldg(&bias_n[c0 + j + e]);
                        if (relu & 1) x = fmaxf(x, 0.0f);
                        if (relu & 2) x = tf32_rn(x);
                        wp[e] = x;
                    }
                    *reinterpret_cast<float4*>(stage_f + (size_t)lane * PITCH + c0 + j) = w;
                }
            }
            __syncwarp();
#pragma unroll 4
            for (int rr = 0; rr < 32; ++rr) {        // one 512-byte pixel row per iteration, 16 bytes per lane
                const long long orr = __shfl_sync(0xffffffffu, orow, rr);
                if (orr >= 0)
                    *reinterpret_cast<float4*>(out + orr + lane * 4) = *reinterpret_cast<const float4*>(stage_f + (size_t)rr * PITCH + lane * 4);
            }
            __syncwarp();
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc<512>(tmem_base);
}

}  // namespace

// in: (B, H, W, C_in) channels-last fp32; wpack: weights packed by crb3d.ops.pack_conv3x3_weight into
// [C_out/128][tap = ky*3+kx][C_in/16][slab 4][128 co][4 ci]; bias: C_out or null; out: (B, H, W, C_out) channels-last.
// relu: bit 0 = ReLU, bit 1 = round the stored values to TF32 (round-to-nearest). Supported: C_in % 16 == 0, C_out % 128 == 0. The 8-pixel tile edge runs along H when H % 8 == 0 (else along W).
extern "C" int crb3d_bev_conv3x3_tf32(const float* in, int B, int H, int W, int cin, const float* wpack, int cout,
                                      const float* bias, int relu, float* out, cudaStream_t stream) {
    if (!in || !wpack || !out || B <= 0 || H <= 0 || W <= 0 || cin <= 0 || cout <= 0) return CRB3D_ERR_ARG;
    if (cin % KC != 0 || cout % N != 0) return CRB3D_ERR_UNSUPPORTED;
    ConvGeom g;
    const bool u_is_y = (H % TU == 0) || (W % TU != 0 && (crb3d_divup(H, TU) * TU - H) * W <= (crb3d_divup(W, TU) * TU - W) * H);
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
        const uint64_t dims[5] = {4, (uint64_t)g.U, (uint64_t)g.V, (uint64_t)cin / 4, (uint64_t)B};
        const uint64_t strides[4] = {u_is_y ? ysb : xsb, u_is_y ? xsb : ysb, 16, (uint64_t)H * W * cin * 4};
        const uint32_t box[5] = {4, PU, PV, KC / 4, 1};
        int rc = make_map_f32(&amap, in, 5, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_NONE);
        if (rc) return rc;
    }
    static bool attr_set = false;
    if (!attr_set) {
        CRB3D_CUDA(cudaFuncSetAttribute(bev_conv3x3_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES + 1024));
        attr_set = true;
    }
    bev_conv3x3_tc<<<dim3((unsigned)crb3d_divup(g.n_tiles, NT), (unsigned)(cout / N)), 192, SMEM_BYTES + 1024, stream>>>(
        amap, wpack, cin / KC, bias, relu, out, g);
    CRB3D_CHECK_LAUNCH();
    return CRB3D_OK;
}
