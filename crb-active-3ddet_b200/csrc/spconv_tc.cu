// Sparse 3D convolution on the 5th-gen tensor cores (tcgen05, TF32 inputs, fp32 accumulate in TMEM) fed by TMA.
//
// Replaces the external spconv-cu113==2.1.21 implicit-GEMM forward reached from
//   pcdet/models/backbones_3d/spconv_backbone.py:77-117   (C_in >= 16 layers of VoxelBackBone8x)
// Same contract as csrc/spconv_simt.cu:  out[o,:] = sum_k in[nbr[k][o],:] @ W[:,k,:]^T, W = [C_out, K, C_in].
//
// One CTA owns 128 consecutive output rows (the UMMA M). Nine warps, mbarrier pipelined:
//   warps 0-7 (producers): threads 0..127 own one tile row each. For every kernel offset k that has a neighbour in the
//       tile the owner reads its row's neighbour index from the table (prefetched one stage ahead) and appends the row to
//       a shared-memory work list; all 256 producer threads then walk the list, one 16-byte cp.async per (row, chunk)
//       into the 128B-swizzled K-major stage (`cp.async.mbarrier.arrive.noinc` signals full[stage]). A row without a
//       neighbour costs nothing unless the stage's previous tenant left data there (one dirty bit per stage slot in the
//       owner's register), in which case it is re-zeroed. One thread issues the C_out x C_in weight slice as a 3-D TMA box.
//   warp 8 (converged, tcgen05 predicated on one elected lane): waits full[stage], issues C_in/8 tcgen05.mma (M=128,
//       N=C_out, K=8, kind::tf32) accumulating ALL offsets into one TMEM tile, tcgen05.commit -> empty[stage].
//   warps 0-3 again (epilogue): tcgen05.ld 32 lanes x 32 columns, optional scale/shift/ReLU, one store per output row.
// The output is written once, there are no atomics and the summation order is fixed (k ascending). History of the row
// path: one-warp cp.async (3000 cycles/stage of issue) -> TMA tile::gather4 (1400-2100 cycles/stage of issue + ~1500 in
// the TMA unit, tools/trace_spconv.py) -> one thread per half row (half-empty sectors) -> per-warp cooperative loop
// (serial latency) -> list-driven LSU gather -> two list-driven producer groups alternating over the stages (this file;
// 553 -> 435 us over the 11 tensor-core layers of a batch-4 step, profiles/r01_spconv_variants.txt).
#include "tc_common.cuh"

namespace {

constexpr int TILE_M = 128;
constexpr int PW = 8;            // producer warps; threads 0..127 additionally own one tile row each (warps 0..3 = epilogue)
constexpr int THREADS = (PW + 1) * 32;
constexpr int MAX_K = 27;

// optional timeline trace of one CTA (clock64 stamps per offset: producer after empty-wait / after issuing its copies,
// MMA thread after full-wait / after issuing + committing); set through crb3d_debug_set_tc_trace. The stamps are compiled in only with
// -DCRB3D_TC_TRACE (tools/trace_spconv.py): even predicated off, four clock reads + stores per stage cost the producers ~4 % of their
// stall samples (ncu source view)
__device__ long long* g_tc_trace = nullptr;
#ifdef CRB3D_TC_TRACE
#define TC_TRACE(ptr, idx) do { if (ptr) (ptr)[idx] = clock64(); } while (0)
#else
#define TC_TRACE(ptr, idx) do { } while (0)
#endif

using tc::smem_u32;
using tc::mbar_init;
using tc::mbar_wait;

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
    // .ca: the copy goes through L1, which merges the 16 lanes that read one 256-byte row into two line requests; with .cg
    // every 16-byte chunk is its own L2 request and the stage time is set by the request rate (~3 cycles per chunk per SM)
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
// arrive on `bar` once all cp.async issued so far by this thread have completed (does not bump the pending count)
__device__ __forceinline__ void cp_async_arrive(uint64_t* bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
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
// STAGES must be even: stage slot it % STAGES is always written by producer group it % 2 (its dirty bits live there).
template <int NKB, int COUT, int STAGES, int MIN_CTAS>
__global__ void __launch_bounds__(THREADS, MIN_CTAS) spconv_fwd_tc(const float* __restrict__ feat, int cin,
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
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // pointer arithmetic keeps the shared address space (LDS/STS, not generic LD/ST)
    __shared__ int act[MAX_K];
    __shared__ int act_flag[MAX_K];
    __shared__ unsigned int list_v[2][2][TILE_M], list_z[2][2][TILE_M];   // [group][double buffer][entries]: per-stage work lists
    __shared__ int cnt_v[2][4], cnt_z[2][4];
    __shared__ int n_act_s;
    __shared__ __align__(16) float scale_s[128], shift_s[128];   // BatchNorm affine of the epilogue (identity where absent)
    __shared__ uint64_t full_bar[STAGES];
    __shared__ uint64_t empty_bar[STAGES];
    __shared__ uint64_t acc_bar;
    __shared__ uint32_t tmem_base_s;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int row0 = blockIdx.x * TILE_M;

    // ---- setup: barriers, TMEM, zeroed stages, which offsets have a neighbour in this tile
    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full_bar[s], TILE_M + 1);     // one arrival per producer thread + the weight box's expect_tx arrive
            mbar_init(&empty_bar[s], 1);           // tcgen05.commit
        }
        mbar_init(&acc_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    if (warp == PW) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "n"(TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    // warp w scans offsets w, w + PW, ...: 128 table cells = one int4 per lane. All of a warp's loads are issued before the stage
    // memory is zeroed and reduced afterwards: one global-load latency per tile instead of one per offset (the scan was ~10 % of
    // the kernel's stall samples in the ncu source view)
    constexpr int SCAN = (MAX_K + PW - 1) / PW;
    int4 scan_v[SCAN];
    int src_spec = -1;      // this thread's table entry for offset `group`: almost always the group's first active offset
    if (warp < PW) {
        const int o1 = row0 + (tid & (TILE_M - 1)), k1 = warp >> 2;
        if (k1 < K && o1 < nv) src_spec = __ldg(&nbr[(size_t)k1 * n_out + o1]);
#pragma unroll
        for (int u = 0; u < SCAN; ++u) {
            const int k = warp + u * PW;
            const int o = row0 + lane * 4;
            int4 v = make_int4(-1, -1, -1, -1);
            if (k < K) {
                if (o + 3 < nv && ((((size_t)k * n_out + o) & 3) == 0)) v = __ldg(reinterpret_cast<const int4*>(nbr + (size_t)k * n_out + o));
                else {
                    if (o < nv) v.x = __ldg(&nbr[(size_t)k * n_out + o]);
                    if (o + 1 < nv) v.y = __ldg(&nbr[(size_t)k * n_out + o + 1]);
                    if (o + 2 < nv) v.z = __ldg(&nbr[(size_t)k * n_out + o + 2]);
                    if (o + 3 < nv) v.w = __ldg(&nbr[(size_t)k * n_out + o + 3]);
                }
            }
            scan_v[u] = v;
        }
    }
    {
        float4* z = reinterpret_cast<float4*>(smem);
        for (int t = tid; t < STAGES * STAGE_BYTES / 16; t += THREADS) z[t] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    for (int c = tid; c < COUT; c += THREADS) { scale_s[c] = scale ? __ldg(&scale[c]) : 1.0f; shift_s[c] = shift ? __ldg(&shift[c]) : 0.0f; }
    if (warp < PW) {
#pragma unroll
        for (int u = 0; u < SCAN; ++u) {
            const int k = warp + u * PW;
            const int4 v = scan_v[u];
            const bool any = (v.x & v.y & v.z & v.w) >= 0;  // some entry is non-negative <=> the AND has a clear sign bit
            const bool warp_any = __any_sync(0xffffffffu, any);
            if (lane == 0 && k < K) act_flag[k] = warp_any ? 1 : 0;
        }
    }
    asm volatile("fence.proxy.async.shared::cta;");  // the zero fill (generic proxy) precedes the tensor core's reads
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem_base = tmem_base_s;
    if (warp == 0) {   // compact list in ascending k (fixed summation order): one ballot
        for (int u = lane; u < 8; u += 32) { cnt_v[u >> 2][u & 3] = 0; cnt_z[u >> 2][u & 3] = 0; }
        const bool f = lane < K && act_flag[lane] != 0;
        const unsigned int m = __ballot_sync(0xffffffffu, f);
        if (f) act[__popc(m & ((1u << lane) - 1u))] = lane;
        if (lane == 0) n_act_s = __popc(m);
    }
    __syncthreads();
    const int n_act = n_act_s;
    const uint32_t smem_base = smem_u32(smem);

    if (warp < PW) {
        // ================================ producers: LSU gather (cp.async), no TMA on the row path =====================
        // Threads 0..127 own one tile row each: per offset they read THEIR neighbour index straight from the table
        // (prefetched one stage ahead) and append (row, source) to a shared-memory list when the row has a neighbour, or
        // the row to a second list when it has none now but the stage's previous tenant left data there (one dirty bit
        // per stage slot in the owner's register; the buffers start out zeroed). After one named barrier ALL 256 producer
        // threads walk the lists: item i = (list entry i / chunks-per-row, 16-byte chunk i % chunks-per-row), one
        // cp.async each (src-size 0 = zero fill for the second list). Consecutive lanes move consecutive chunks of a row,
        // so every 32-byte sector that is fetched is used, the iterations are independent, and only the ~15-30 % of the
        // (offset, row) slots that are occupied cost anything. Measured alternatives (profiles/r01_spconv_variants.txt):
        // TMA tile::gather4 is bound by the TMA unit (~80 cycles per 4-row gather), one thread per half row issues
        // half-empty sectors, a per-warp cooperative loop is bound by its own serial instruction latency.
        // TWO producer groups (warps 0-3 / 4-7) alternate over the stages: `cp.async.mbarrier.arrive.noinc` holds the
        // issuing thread until its copies have landed (measured: ~1500 cycles per stage), so one group would serialise
        // issue and landing; with two, group g issues stage it+1 while group 1-g waits for stage it.
        constexpr int NG = 2, GT = TILE_M;                // groups, threads per group (one row per thread)
        const int grp = warp >> 2, gtid = tid & (GT - 1);
        const int r = gtid;                               // tile row owned by this thread within its group
        const int o = row0 + r;
        const int cpr = cin >> 2;                         // 16-byte chunks per row (1, 2, 4, 8 or 16)
        const int cshift = 31 - __clz(cpr);
#ifdef CRB3D_TC_TRACE
        long long* const trace_base = (blockIdx.x == gridDim.x / 2 && gtid == 0) ? g_tc_trace : nullptr;   // read once
#endif
        uint32_t dirty = 0u;
        int src_next = -1;
        if (grp < n_act && o < nv) src_next = act[grp] == grp ? src_spec : __ldg(&nbr[(size_t)act[grp] * n_out + o]);
        for (int it = grp, li = 0; it < n_act; it += NG, ++li) {
            const int stage = it % STAGES, lb = li & 1;
            const int src = src_next;
            if (it + NG < n_act) src_next = (o < nv) ? __ldg(&nbr[(size_t)act[it + NG] * n_out + o]) : -1;
            if (gtid == 0) { cnt_v[grp][(li + 2) & 3] = 0; cnt_z[grp][(li + 2) & 3] = 0; }
            if (it >= STAGES) mbar_wait(&empty_bar[stage], ((it / STAGES) - 1) & 1, (CRB3D_K_SPCONV_TC << 8) | 9, it);
#ifdef CRB3D_TC_TRACE
            long long* trace = (trace_base && it < 32) ? trace_base + 128 : nullptr;
#endif
            TC_TRACE(trace, it * 4 + 0);
            const uint32_t a_base = smem_base + stage * STAGE_BYTES, b_base = a_base + A_BYTES;
            if (gtid == GT - 1) {   // one thread of the group fetches the weight slice: one 3-D TMA box per k-block
                const int kw = kmap ? kmap[act[it]] : act[it];
                const uint32_t bar = smem_u32(&full_bar[stage]);
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((uint32_t)B_BYTES));
#pragma unroll
                for (int kb = 0; kb < NKB; ++kb)
                    asm volatile(
                        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                        ::"r"(b_base + kb * (COUT * 128)), "l"(reinterpret_cast<uint64_t>(&wmap)), "r"(kb * 32), "r"(kw), "r"(0),
                        "r"(bar) : "memory");
            }
            {
                const bool valid = src >= 0;
                const bool stale = !valid && (dirty & (1u << stage));
                const unsigned int mv = __ballot_sync(0xffffffffu, valid), mz = __ballot_sync(0xffffffffu, stale);
                if (valid) dirty |= 1u << stage; else dirty &= ~(1u << stage);
                int bv = 0, bz = 0;
                if (lane == 0) {
                    if (mv) bv = atomicAdd(&cnt_v[grp][li & 3], __popc(mv));
                    if (mz) bz = atomicAdd(&cnt_z[grp][li & 3], __popc(mz));
                }
                bv = __shfl_sync(0xffffffffu, bv, 0);
                bz = __shfl_sync(0xffffffffu, bz, 0);
                const unsigned int lt = (1u << lane) - 1u;
                if (valid) list_v[grp][lb][bv + __popc(mv & lt)] = ((unsigned int)r << 25) | (unsigned int)src;
                if (stale) list_z[grp][lb][bz + __popc(mz & lt)] = (unsigned int)r;
            }
            asm volatile("bar.sync %0, %1;" ::"r"(1 + grp), "n"(GT) : "memory");   // this group's lists are complete
            TC_TRACE(trace, it * 4 + 2);
            const int n_v = cnt_v[grp][li & 3] << cshift, n_z = cnt_z[grp][li & 3] << cshift;
            if constexpr (NKB == 2) {
                // 64-channel rows (16 chunks, 4-5 list passes per stage): thread -> (list entry j0 + k * entries-per-pass, its fixed
                // 16-byte chunk) with four list reads in flight - the loop is a chain of shared-memory load -> address arithmetic ->
                // LDGSTS whose latency, not its issue rate, is what the stage waits for (conv3.1 137 -> 131 us, conv4.1 83 -> 78;
                // the narrower layers have 1-2 passes and lose 5-10 % to the extra loop structure, so they keep the plain loop)
                const int epp = GT >> cshift, c16 = gtid & (cpr - 1);
                const int n_rows = n_v >> cshift;
                const uint32_t col = a_base + (c16 >> 3) * (TILE_M * 128), cx = (uint32_t)(c16 & 7);
                const float* fsrc = feat + c16 * 4;
                const unsigned int* lst = list_v[grp][lb];
                int j = gtid >> cshift;
                for (; j + 3 * epp < n_rows; j += 4 * epp) {
                    unsigned int e[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) e[u] = lst[j + u * epp];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const uint32_t row = e[u] >> 25;
                        cp_async16(col + (row >> 3) * 1024 + (row & 7) * 128 + ((cx ^ (row & 7)) << 4), fsrc + (size_t)(e[u] & 0x1FFFFFFu) * cin);
                    }
                }
                for (; j < n_rows; j += epp) {
                    const unsigned int e = lst[j];
                    const uint32_t row = e >> 25;
                    cp_async16(col + (row >> 3) * 1024 + (row & 7) * 128 + ((cx ^ (row & 7)) << 4), fsrc + (size_t)(e & 0x1FFFFFFu) * cin);
                }
            } else {
                for (int i = gtid; i < n_v; i += GT) {
                    const unsigned int e = list_v[grp][lb][i >> cshift];
                    const int row = (int)(e >> 25), c16 = i & (cpr - 1);
                    cp_async16(a_base + (row >> 3) * 1024 + (row & 7) * 128 + (c16 >> 3) * (TILE_M * 128) + (((c16 & 7) ^ (row & 7)) << 4),
                               feat + (size_t)(e & 0x1FFFFFFu) * cin + c16 * 4);
                }
            }
            if (n_z > 0) {   // stale rows: plain 16-byte zero stores (the LDGSTS path is the scarce resource) + proxy fence
                for (int i = gtid; i < n_z; i += GT) {
                    const int row = (int)list_z[grp][lb][i >> cshift], c16 = i & (cpr - 1);
                    asm volatile("st.shared.v4.f32 [%0], {%1, %1, %1, %1};" ::"r"(a_base + (row >> 3) * 1024 + (row & 7) * 128 + (c16 >> 3) * (TILE_M * 128) + (((c16 & 7) ^ (row & 7)) << 4)), "f"(0.0f));
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            }
            TC_TRACE(trace, it * 4 + 3);
            cp_async_arrive(&full_bar[stage]);            // arrives when this thread's copies (if any) have landed
            TC_TRACE(trace, it * 4 + 1);
        }
    } else if (warp == PW) {
        // ================================ MMA issuer ================================
        // the whole warp runs the loop converged and only the tcgen05 instructions are predicated on one elected lane:
        // under `if (lane == 0)` every descriptor is a per-thread value that ptxas moves to the uniform register file
        // before each UTCHMMA (~100 issue cycles per MMA, 800 cycles per stage in tools/trace_spconv.py)
        const uint32_t idesc = make_idesc_tf32(TILE_M, COUT);
        const uint64_t desc0 = make_desc_sw128(smem_base);
#ifdef CRB3D_TC_TRACE
        long long* const trace_mma = (blockIdx.x == gridDim.x / 2 && lane == 0) ? g_tc_trace : nullptr;
#endif
        for (int it = 0; it < n_act; ++it) {
            const int stage = it % STAGES;
            mbar_wait(&full_bar[stage], (it / STAGES) & 1, (CRB3D_K_SPCONV_TC << 8) | 8, it);
#ifdef CRB3D_TC_TRACE
            long long* trace = (trace_mma && it < 32) ? trace_mma : nullptr;
#endif
            TC_TRACE(trace, it * 4 + 2);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // cp.async (generic proxy) writes -> tensor-core (async proxy) reads
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
            TC_TRACE(trace, it * 4 + 3);
        }
    }

    // ---- epilogue: TMEM -> registers -> global. A warp may only touch TMEM lanes 32*(warp%4)..+31: warps 0-3 take the first half of
    // the columns, warps 4-7 the second (C_out >= 64); the affine comes from shared memory (two global loads per element stalled
    // every FMA of the old epilogue: sparse backbone 1.23 -> 1.01 ms per batch of 16). fmaf(x, 1, shift) and fmaf(x, scale, 0) are exact, so absent terms change no bit.
    constexpr int EP_SPLIT = COUT >= 64 ? 2 : 1, EP_COLS = COUT / EP_SPLIT;
    if (warp < 4 * EP_SPLIT) {
        if (n_act > 0) {
            mbar_wait(&acc_bar, 0, (CRB3D_K_SPCONV_TC << 8) | 7);
            asm volatile("tcgen05.fence::after_thread_sync;");
        }
        const int quarter = warp & 3, cbase = (warp >> 2) * EP_COLS;
        const int r = quarter * 32 + lane;
        const int o = row0 + r;
#pragma unroll
        for (int cc = 0; cc < EP_COLS; cc += 32) {
            const int c0 = cbase + cc;
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
                    const float4 sc = *reinterpret_cast<const float4*>(scale_s + c0 + j), sh = *reinterpret_cast<const float4*>(shift_s + c0 + j);
                    float4 w = make_float4(fmaf(__uint_as_float(v[j]), sc.x, sh.x), fmaf(__uint_as_float(v[j + 1]), sc.y, sh.y),
                                           fmaf(__uint_as_float(v[j + 2]), sc.z, sh.z), fmaf(__uint_as_float(v[j + 3]), sc.w, sh.w));
                    if (relu & 1) { w.x = fmaxf(w.x, 0.0f); w.y = fmaxf(w.y, 0.0f); w.z = fmaxf(w.z, 0.0f); w.w = fmaxf(w.w, 0.0f); }
                    if (relu & 2) { w.x = tc::tf32_rn(w.x); w.y = tc::tf32_rn(w.y); w.z = tc::tf32_rn(w.z); w.w = tc::tf32_rn(w.w); }   // the next tensor-core layer reads exactly what was stored
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

// 3-D view {ci, k, co} of the contiguous [C_out, K, C_in] weight; box {32, 1, cout}, 128-byte swizzle, zero OOB fill
int make_weight_map(const float* weight, int K, int cin, int cout, CUtensorMap* map) {
    tc::EncodeTiledFn enc = tc::get_encode_fn();
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
int launch_tc(const float* feat, int n_in, const int* nbr, const float* weight, int n_out, int K, int cin, const int* kmap,
              const float* scale, const float* shift, int relu, float* out, const int* n_dev, cudaStream_t stream) {
    constexpr size_t smem = (size_t)STAGES * (NKB * TILE_M * 128 + NKB * COUT * 128) + 1024;
    CUtensorMap wmap;
    int rc = make_weight_map(weight, K, cin, COUT, &wmap);
    if (rc) return rc;
    auto kern = spconv_fwd_tc<NKB, COUT, STAGES, MIN_CTAS>;
    static bool attr_set[CRB3D_MAX_DEVICES] = {};  // per instantiation and per device (the attribute is per function per device)
    const int dev = crb3d_current_device();
    if (!attr_set[dev]) {
        CRB3D_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set[dev] = true;
    }
    kern<<<(unsigned)crb3d_divup(n_out, TILE_M), THREADS, smem, stream>>>(feat, cin, wmap, nbr, n_out, K, kmap, scale, shift, relu, out, n_dev);
    CRB3D_CHECK_LAUNCH();
    return CRB3D_OK;
}

}  // namespace

int crb3d_spconv_forward_tf32_grouped(const float* feat, const int* nbr, const float* weight, int n_out, int K, int cin, int cout,
                                      const float* scale, const float* shift, int relu, float* out, const int* n_dev,
                                      cudaStream_t stream);

// TF32 tensor-core forward. feat: (n_in, C_in) contiguous; weight: contiguous [C_out, K, C_in] (for the input gradient
// pass the transposed weight [C_in, K, C_out] and the transposed table). Supported: C_in in {16, 32, 64}, C_out in
// {16, 32, 64, 128}, K <= 27; anything else returns CRB3D_ERR_UNSUPPORTED (callers use crb3d_spconv_forward_f32).
// relu: bit 0 = ReLU, bit 1 = round the stored values to TF32 (round-to-nearest; the tensor core truncates fp32 operands).
// Stage counts are sized so that two (C_in = 64) or three (C_in <= 32) CTAs fit one SM: the gather is latency-bound and
// the co-resident CTAs hide each other's round trips (profiles/r01_spconv_variants.txt, variant G).
extern "C" int crb3d_spconv_forward_tf32(const float* feat, int n_in, const int* nbr, const float* weight, int n_out, int K,
                                         int cin, int cout, const int* kmap, const float* scale, const float* shift,
                                         int relu, float* out, const int* n_dev, cudaStream_t stream) {
    if (n_out < 0 || n_in < 0 || K <= 0 || cin <= 0 || cout <= 0 || !weight || !out) return CRB3D_ERR_ARG;
    if (n_out == 0) return CRB3D_OK;
    if (!feat || !nbr || n_in == 0) return CRB3D_ERR_ARG;
    // C_in = 4 / 8 (the first layer: raw voxel features) ride on the 32-channel path: the TMA weight box and the untouched
    // tail of every 128-byte stage row are zero
    if (K > MAX_K || (cin != 4 && cin != 8 && cin != 16 && cin != 32 && cin != 64) || n_in >= (1 << 25)) return CRB3D_ERR_UNSUPPORTED;
    // narrow layers, forward direction: several offsets per pipeline stage (csrc/spconv_tc_grp.cu); relu bit 2 keeps them on the
    // one-offset-per-stage kernel below (A/B measurements)
    if (!kmap && cin <= 8 && !(relu & 4)) {
        const int rc = crb3d_spconv_forward_tf32_grouped(feat, nbr, weight, n_out, K, cin, cout, scale, shift, relu, out, n_dev, stream);
        if (rc != CRB3D_ERR_UNSUPPORTED) return rc;
    }
#define TC_ARGS feat, n_in, nbr, weight, n_out, K, cin, kmap, scale, shift, relu, out, n_dev, stream
    const int nkb = (cin + 31) / 32;
    if (nkb == 1) {                                        // stage = 16 KB + C_out*128 B
        if (cout == 16) return launch_tc<1, 16, 2, 3>(TC_ARGS);
        if (cout == 32) return launch_tc<1, 32, 2, 3>(TC_ARGS);
        if (cout == 64) return launch_tc<1, 64, 2, 3>(TC_ARGS);
        if (cout == 128) return launch_tc<1, 128, 2, 2>(TC_ARGS);
    } else if (nkb == 2) {                                 // stage = 32 KB + C_out*256 B
        if (cout == 16) return launch_tc<2, 16, 2, 2>(TC_ARGS);
        if (cout == 32) return launch_tc<2, 32, 2, 2>(TC_ARGS);
        // fewer tiles than SMs: one CTA per SM anyway, so spend the shared memory on four stages (each producer group then
        // runs two stages ahead of the tensor core instead of waiting for its previous stage's MMAs)
        if (cout == 64 && crb3d_divup(n_out, TILE_M) <= crb3d_num_sms()) return launch_tc<2, 64, 4, 1>(TC_ARGS);
        if (cout == 64) return launch_tc<2, 64, 2, 2>(TC_ARGS);
        if (cout == 128) return launch_tc<2, 128, 2, 1>(TC_ARGS);
    }
#undef TC_ARGS
    return CRB3D_ERR_UNSUPPORTED;
}

// Debug hook (not part of include/crb3d.h): buf = device array of >= 256 int64, or null to switch tracing off.
extern "C" int crb3d_debug_set_tc_trace(long long* buf) {
    return cudaMemcpyToSymbol(g_tc_trace, &buf, sizeof(buf)) == cudaSuccess ? CRB3D_OK : CRB3D_ERR_CUDA;
}

CRB3D_DIAG_DEFINE_SETTER(spconv_tc)
