// PointNet++ set-abstraction ops on stacked batches: ball query, grouping (+grad), farthest point sampling,
// 3-NN and 3-point interpolation (+grad).
//
// Replaces (reference, /root/reference/pcdet/ops/pointnet2/pointnet2_stack/src):
//   ball_query_gpu.cu:16-66        ball_query_kernel_stack        (first `nsample` hits in index order, pad with first)
//   group_points_gpu.cu:15-102     group_points[_grad]_kernel_stack
//   sampling_gpu.cu:25-176,188-340 farthest_point_sampling_kernel / stack_farthest_point_sampling_kernel
//   interpolate_gpu.cu:16-172      three_nn_kernel_stack / three_interpolate[_grad]_kernel_stack
// Index outputs are bit-exact targets, so distance expressions keep the reference's form and tie rules:
//   ball query - ascending source index; FPS - max distance, ties to the lowest (k mod block, k) exactly as the
//   reference's strided scan + tree reduction resolves them; 3-NN - strict '<' (lowest index first among equals).
// What changed: one WARP per query walks 32 candidates per step from a shared-memory tile (ballot + popc gives the
// ordered append), grouping goes through a shared-memory transpose so both sides coalesce, FPS keeps distances in
// registers and reduces packed 64-bit (distance, tie-key) words with warp shuffles (2 barriers per round, not 11).
#include "common.cuh"
#include <cmath>

namespace {

// ------------------------------------------------------------------ ball query
constexpr int BQ_WARPS = 8;
constexpr int BQ_TILE = 1024;

__global__ void __launch_bounds__(BQ_WARPS * 32) ball_query_kernel(int B, int M, float radius, int nsample,
                                                                   const float* __restrict__ new_xyz,
                                                                   const int* __restrict__ new_cnt,
                                                                   const float* __restrict__ xyz,
                                                                   const int* __restrict__ xyz_cnt, int* __restrict__ idx) {
    __shared__ float sx[BQ_TILE], sy[BQ_TILE], sz[BQ_TILE];
    const int b = blockIdx.y;
    int q_begin = 0, s_begin = 0;
    for (int k = 0; k < b; ++k) { q_begin += new_cnt[k]; s_begin += xyz_cnt[k]; }
    const int q_cnt = new_cnt[b], n = xyz_cnt[b];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q_local = blockIdx.x * BQ_WARPS + warp;
    if (blockIdx.x * BQ_WARPS >= q_cnt) return;  // whole CTA idle for this frame
    const bool live = q_local < q_cnt;
    const int q = q_begin + q_local;
    float qx = 0.f, qy = 0.f, qz = 0.f;
    if (live) { qx = new_xyz[(size_t)q * 3]; qy = new_xyz[(size_t)q * 3 + 1]; qz = new_xyz[(size_t)q * 3 + 2]; }
    const float r2 = radius * radius;
    int* out = idx + (size_t)q * nsample;
    int cnt = 0;
    bool done = !live;
    for (int t0 = 0; t0 < n; t0 += BQ_TILE) {
        const int tn = min(BQ_TILE, n - t0);
        __syncthreads();
        for (int t = threadIdx.x; t < tn; t += blockDim.x) {
            const float* p = xyz + (size_t)(s_begin + t0 + t) * 3;
            sx[t] = p[0]; sy[t] = p[1]; sz[t] = p[2];
        }
        __syncthreads();
        if (!done) {
            for (int k0 = 0; k0 < tn; k0 += 32) {
                const int k = k0 + lane;
                bool hit = false;
                if (k < tn) {
                    const float x = sx[k], y = sy[k], z = sz[k];
                    const float d2 = sqdist3(qx, qy, qz, x, y, z);
                    hit = d2 < r2;
                }
                const unsigned int m = __ballot_sync(0xffffffffu, hit);
                if (m) {
                    if (cnt == 0) {  // first hit pads the whole row (ball_query_gpu.cu:54-58)
                        const int first = t0 + k0 + __ffs(m) - 1;
                        for (int l = lane; l < nsample; l += 32) out[l] = first;
                        __syncwarp();
                    }
                    const int pos = cnt + __popc(m & ((1u << lane) - 1u));
                    if (hit && pos < nsample) out[pos] = t0 + k;
                    cnt += __popc(m);
                    if (cnt >= nsample) { done = true; break; }
                }
            }
        }
        if (__syncthreads_and(done)) break;
    }
    if (live && cnt == 0 && lane == 0) out[0] = -1;
}

// ------------------------------------------------------------------ grouping
// out[m][c][s] = feat[start_b + idx[m][s]][c]; one CTA per query point, smem tile [ns][C+1].
__global__ void __launch_bounds__(256) group_points_kernel(int B, int M, int C, int ns, const float* __restrict__ feat,
                                                           const int* __restrict__ feat_cnt, const int* __restrict__ idx,
                                                           const int* __restrict__ idx_cnt, float* __restrict__ out) {
    extern __shared__ float tile[];  // [ns][C+1]
    const int m = blockIdx.x;
    int b = 0, acc = idx_cnt[0];
    for (int k = 1; k < B; ++k) { if (m < acc) break; acc += idx_cnt[k]; b = k; }
    int start = 0;
    for (int k = 0; k < b; ++k) start += feat_cnt[k];
    const int* id = idx + (size_t)m * ns;
    for (int t = threadIdx.x; t < ns * C; t += blockDim.x) {
        const int s = t / C, c = t - s * C;
        tile[s * (C + 1) + c] = __ldg(feat + (size_t)(start + id[s]) * C + c);
    }
    __syncthreads();
    float* o = out + (size_t)m * C * ns;
    for (int t = threadIdx.x; t < ns * C; t += blockDim.x) {
        const int c = t / ns, s = t - c * ns;
        o[t] = tile[s * (C + 1) + c];
    }
}

__global__ void __launch_bounds__(256) group_points_grad_kernel(int B, int M, int C, int ns,
                                                                const float* __restrict__ grad_out,
                                                                const int* __restrict__ idx, const int* __restrict__ idx_cnt,
                                                                const int* __restrict__ feat_cnt,
                                                                float* __restrict__ grad_feat) {
    extern __shared__ float tile[];  // [ns][C+1]
    const int m = blockIdx.x;
    int b = 0, acc = idx_cnt[0];
    for (int k = 1; k < B; ++k) { if (m < acc) break; acc += idx_cnt[k]; b = k; }
    int start = 0;
    for (int k = 0; k < b; ++k) start += feat_cnt[k];
    const float* g = grad_out + (size_t)m * C * ns;
    for (int t = threadIdx.x; t < ns * C; t += blockDim.x) {
        const int c = t / ns, s = t - c * ns;
        tile[s * (C + 1) + c] = g[t];
    }
    __syncthreads();
    const int* id = idx + (size_t)m * ns;
    for (int t = threadIdx.x; t < ns * C; t += blockDim.x) {
        const int s = t / C, c = t - s * C;
        atomicAdd(grad_feat + (size_t)(start + id[s]) * C + c, tile[s * (C + 1) + c]);
    }
}

// ------------------------------------------------------------------ farthest point sampling
// key = (distance bits << 32) | ~tie: the maximum key is the farthest point. Ties follow the reference exactly: its
// strided scan keeps the first maximum per thread (lowest k / block) and its pairwise tree (slot t absorbs slot t+s,
// left wins ties, s = block/2 ... 1) lets the thread whose id has a 0 at the LOWEST differing bit win, i.e. threads are
// ordered by their bit-reversed id. tie = bitrev(k mod block) << 21 | (k / block).
__device__ __forceinline__ unsigned long long fps_key(float d, int k, int ref_block_log2) {
    const unsigned int t = (unsigned int)(k & ((1 << ref_block_log2) - 1));
    const unsigned int rev = ref_block_log2 ? (__brev(t) >> (32 - ref_block_log2)) : 0u;
    const unsigned int tie = (rev << 21) | (unsigned int)(k >> ref_block_log2);
    return ((unsigned long long)__float_as_uint(d) << 32) | (unsigned long long)(~tie);
}

template <int PPT>  // points per thread held in registers (n <= 1024 * PPT); PPT == 0: distances in global `temp`
__global__ void __launch_bounds__(1024) fps_kernel(int n_fixed, int m_fixed, const float* __restrict__ dataset,
                                                   float* __restrict__ temp, const int* __restrict__ xyz_cnt,
                                                   const int* __restrict__ m_cnt, int* __restrict__ idxs,
                                                   int ref_block_log2_fixed) {
    __shared__ unsigned long long wbest[32];
    __shared__ int cur_s;
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int n, m, start = 0, ostart = 0, rbl;
    if (xyz_cnt) {  // stacked layout (sampling_gpu.cu:188-319): block size fixed at 1024 in the reference
        for (int k = 0; k < b; ++k) { start += xyz_cnt[k]; ostart += m_cnt[k]; }
        n = xyz_cnt[b]; m = m_cnt[b]; rbl = 10;
    } else {
        n = n_fixed; m = m_fixed; start = b * n; ostart = b * m; rbl = ref_block_log2_fixed;
    }
    if (m <= 0) return;
    const float* pts = dataset + (size_t)start * 3;
    float* tmp = temp + start;
    int* out = idxs + ostart;
    const int out_base = xyz_cnt ? start : 0;

    float px[PPT > 0 ? PPT : 1], py[PPT > 0 ? PPT : 1], pz[PPT > 0 ? PPT : 1], pd[PPT > 0 ? PPT : 1];
    if (PPT > 0) {
#pragma unroll
        for (int i = 0; i < PPT; ++i) {
            const int k = tid + i * 1024;
            if (k < n) { px[i] = pts[(size_t)k * 3]; py[i] = pts[(size_t)k * 3 + 1]; pz[i] = pts[(size_t)k * 3 + 2]; pd[i] = tmp[k]; }
            else { px[i] = py[i] = pz[i] = 0.f; pd[i] = 0.f; }
        }
    }
    int old = 0;
    if (tid == 0) out[0] = out_base;
    for (int j = 1; j < m; ++j) {
        const float x1 = pts[(size_t)old * 3], y1 = pts[(size_t)old * 3 + 1], z1 = pts[(size_t)old * 3 + 2];
        // threads with no point contribute (best=-1, besti=0) in the reference: they can never win
        bool have = false;
        float bd = -1.0f; int bi = 0;
        if (PPT > 0) {
#pragma unroll
            for (int i = 0; i < PPT; ++i) {
                const int k = tid + i * 1024;
                if (k < n) {
                    const float x2 = px[i], y2 = py[i], z2 = pz[i];
                    const float d = sqdist3(x2, y2, z2, x1, y1, z1);
                    const float d2 = fminf(d, pd[i]);
                    pd[i] = d2;
                    if (d2 > bd) { bd = d2; bi = k; have = true; }
                }
            }
        } else {
            for (int k = tid; k < n; k += 1024) {
                const float x2 = pts[(size_t)k * 3], y2 = pts[(size_t)k * 3 + 1], z2 = pts[(size_t)k * 3 + 2];
                const float d = sqdist3(x2, y2, z2, x1, y1, z1);
                const float d2 = fminf(d, tmp[k]);
                tmp[k] = d2;
                if (d2 > bd) { bd = d2; bi = k; have = true; }
            }
        }
        // Per-thread winner -> packed key. A thread whose every d2 <= -1 (or that owns no point) reports k=0, d=-1,
        // which can only win if all distances are negative/NaN; then the reference also returns index 0.
        unsigned long long key = have ? fps_key(bd, bi, rbl) : 0ull;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            unsigned long long other = __shfl_xor_sync(0xffffffffu, key, o);
            key = other > key ? other : key;
        }
        if (lane == 0) wbest[warp] = key;
        __syncthreads();
        if (warp == 0) {
            unsigned long long k2 = wbest[lane];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                unsigned long long other = __shfl_xor_sync(0xffffffffu, k2, o);
                k2 = other > k2 ? other : k2;
            }
            if (lane == 0) {
                int win = 0;
                if (k2 != 0ull) {
                    const unsigned int tie = ~(unsigned int)(k2 & 0xFFFFFFFFull);
                    const unsigned int rev = tie >> 21;
                    const unsigned int t = rbl ? (__brev(rev) >> (32 - rbl)) : 0u;
                    win = (int)((tie & ((1u << 21) - 1u)) << rbl) | (int)t;
                }
                cur_s = win;
                out[j] = win + out_base;
            }
        }
        __syncthreads();
        old = cur_s;
    }
    if (PPT > 0) {  // hand the final running distances back (the reference mutates `temp` in place)
#pragma unroll
        for (int i = 0; i < PPT; ++i) {
            const int k = tid + i * 1024;
            if (k < n) tmp[k] = pd[i];
        }
    }
}

// ---- the same selection on a thread-block CLUSTER: FPS_CS CTAs x 1024 threads share one frame, so that clouds of up to
// FPS_CS * 1024 * 8 points keep every coordinate and running distance in registers (a 20 k-point KITTI cloud is 3 points per
// thread; the single-CTA kernel above falls back to distances in global memory beyond 8 k points: 5.9 us per round, 12 ms per
// frame, 64 % of a PV-RCNN forward in profiles/r02_pvrcnn.txt). Per round: per-thread best -> warp shuffles -> CTA best ->
// every CTA writes its packed (distance, tie) key into every peer's shared memory (DSMEM) -> ONE cluster barrier -> all take
// the maximum. The key is a total order that does not depend on how the points are split over threads, so the indices are
// the reference's, bit for bit. Slots are double-buffered by round parity (a fast CTA may already publish round j+1 while a
// slow one still reads round j).
constexpr int FPS_CS = 8;

template <int PPT>
__global__ void __cluster_dims__(FPS_CS, 1, 1) __launch_bounds__(1024) fps_cluster_kernel(
    int n_fixed, int m_fixed, const float* __restrict__ dataset, float* __restrict__ temp, const int* __restrict__ xyz_cnt,
    const int* __restrict__ m_cnt, int* __restrict__ idxs, int ref_block_log2_fixed) {
    __shared__ unsigned long long wbest[32];
    __shared__ unsigned long long slots[2][FPS_CS];
    uint32_t rank;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    const int b = blockIdx.x / FPS_CS, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int n, m, start = 0, ostart = 0, rbl;
    if (xyz_cnt) {
        for (int k = 0; k < b; ++k) { start += xyz_cnt[k]; ostart += m_cnt[k]; }
        n = xyz_cnt[b]; m = m_cnt[b]; rbl = 10;
    } else {
        n = n_fixed; m = m_fixed; start = b * n; ostart = b * m; rbl = ref_block_log2_fixed;
    }
    if (m <= 0) return;                               // uniform over the cluster
    const float* pts = dataset + (size_t)start * 3;
    float* tmp = temp + start;
    int* out = idxs + ostart;
    const int out_base = xyz_cnt ? start : 0;
    const int g = (int)rank * 1024 + tid;
    float px[PPT], py[PPT], pz[PPT], pd[PPT];
#pragma unroll
    for (int i = 0; i < PPT; ++i) {
        const int k = g + i * (FPS_CS * 1024);
        if (k < n) { px[i] = pts[(size_t)k * 3]; py[i] = pts[(size_t)k * 3 + 1]; pz[i] = pts[(size_t)k * 3 + 2]; pd[i] = tmp[k]; }
        else { px[i] = py[i] = pz[i] = 0.f; pd[i] = 0.f; }
    }
    // this CTA's slot [parity][rank] in every CTA of the cluster (lane r of warp 0 addresses CTA r)
    uint32_t remote[2] = {0u, 0u};
    if (warp == 0 && lane < FPS_CS) {
#pragma unroll
        for (int p = 0; p < 2; ++p) {
            const uint32_t local = (uint32_t)__cvta_generic_to_shared(&slots[p][rank]);
            asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote[p]) : "r"(local), "r"(lane));
        }
    }
    int old = 0;
    if (rank == 0 && tid == 0) out[0] = out_base;
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");   // peers are running
    for (int j = 1; j < m; ++j) {
        const float x1 = __ldg(&pts[(size_t)old * 3]), y1 = __ldg(&pts[(size_t)old * 3 + 1]), z1 = __ldg(&pts[(size_t)old * 3 + 2]);
        unsigned long long key = 0ull;
#pragma unroll
        for (int i = 0; i < PPT; ++i) {
            const int k = g + i * (FPS_CS * 1024);
            if (k < n) {
                const float d2 = fminf(sqdist3(px[i], py[i], pz[i], x1, y1, z1), pd[i]);
                pd[i] = d2;
                if (d2 > -1.0f) {                     // the reference's running best starts at -1 with a strict '>'
                    const unsigned long long kk = fps_key(d2, k, rbl);
                    key = kk > key ? kk : key;
                }
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const unsigned long long other = __shfl_xor_sync(0xffffffffu, key, o);
            key = other > key ? other : key;
        }
        if (lane == 0) wbest[warp] = key;
        __syncthreads();
        if (warp == 0) {
            unsigned long long k2 = wbest[lane];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const unsigned long long other = __shfl_xor_sync(0xffffffffu, k2, o);
                k2 = other > k2 ? other : k2;
            }
            if (lane < FPS_CS) asm volatile("st.shared::cluster.u64 [%0], %1;" ::"r"(remote[j & 1]), "l"(k2) : "memory");
        }
        asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
        unsigned long long best = 0ull;
#pragma unroll
        for (int r = 0; r < FPS_CS; ++r) {
            const unsigned long long v = slots[j & 1][r];
            best = v > best ? v : best;
        }
        int win = 0;
        if (best != 0ull) {
            const unsigned int tie = ~(unsigned int)(best & 0xFFFFFFFFull);
            const unsigned int rev = tie >> 21;
            const unsigned int t = rbl ? (__brev(rev) >> (32 - rbl)) : 0u;
            win = (int)((tie & ((1u << 21) - 1u)) << rbl) | (int)t;
        }
        if (rank == 0 && tid == 0) out[j] = win + out_base;
        old = win;
    }
#pragma unroll
    for (int i = 0; i < PPT; ++i) {
        const int k = g + i * (FPS_CS * 1024);
        if (k < n) tmp[k] = pd[i];
    }
    // no CTA may leave while a peer can still write into its shared memory
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// ------------------------------------------------------------------ 3-NN + interpolation
__global__ void __launch_bounds__(256) three_nn_kernel(int B, int N, const float* __restrict__ unknown,
                                                       const int* __restrict__ unknown_cnt, const float* __restrict__ known,
                                                       const int* __restrict__ known_cnt, float* __restrict__ dist2,
                                                       int* __restrict__ idx) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= N) return;
    int b = 0, acc = unknown_cnt[0];
    for (int k = 1; k < B; ++k) { if (p < acc) break; acc += unknown_cnt[k]; b = k; }
    int kstart = 0;
    for (int k = 0; k < b; ++k) kstart += known_cnt[k];
    const int nk = known_cnt[b];
    const float* kn = known + (size_t)kstart * 3;
    const float ux = unknown[(size_t)p * 3], uy = unknown[(size_t)p * 3 + 1], uz = unknown[(size_t)p * 3 + 2];
    float b1 = INFINITY, b2 = INFINITY, b3 = INFINITY;  // the reference's 1e40 double sentinel == +inf once stored as float
    int i1 = 0, i2 = 0, i3 = 0;
    for (int k = 0; k < nk; ++k) {
        const float x = __ldg(kn + (size_t)k * 3), y = __ldg(kn + (size_t)k * 3 + 1), z = __ldg(kn + (size_t)k * 3 + 2);
        const float d = sqdist3(ux, uy, uz, x, y, z);
        if (d < b1) { b3 = b2; i3 = i2; b2 = b1; i2 = i1; b1 = d; i1 = k; }
        else if (d < b2) { b3 = b2; i3 = i2; b2 = d; i2 = k; }
        else if (d < b3) { b3 = d; i3 = k; }
    }
    dist2[(size_t)p * 3] = b1; dist2[(size_t)p * 3 + 1] = b2; dist2[(size_t)p * 3 + 2] = b3;
    idx[(size_t)p * 3] = i1 + kstart; idx[(size_t)p * 3 + 1] = i2 + kstart; idx[(size_t)p * 3 + 2] = i3 + kstart;
}

__global__ void __launch_bounds__(256) three_interp_kernel(int N, int C, const float* __restrict__ feat,
                                                           const int* __restrict__ idx, const float* __restrict__ w,
                                                           float* __restrict__ out) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (int64_t)N * C) return;
    const int c = (int)(t % C);
    const int64_t p = t / C;
    const int* id = idx + p * 3;
    const float* ww = w + p * 3;
    out[t] = ww[0] * __ldg(feat + (size_t)id[0] * C + c) + ww[1] * __ldg(feat + (size_t)id[1] * C + c) +
             ww[2] * __ldg(feat + (size_t)id[2] * C + c);
}

__global__ void __launch_bounds__(256) three_interp_grad_kernel(int N, int C, const float* __restrict__ grad_out,
                                                                const int* __restrict__ idx, const float* __restrict__ w,
                                                                float* __restrict__ grad_feat) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (int64_t)N * C) return;
    const int c = (int)(t % C);
    const int64_t p = t / C;
    const int* id = idx + p * 3;
    const float* ww = w + p * 3;
    const float g = grad_out[t];
    atomicAdd(grad_feat + (size_t)id[0] * C + c, g * ww[0]);
    atomicAdd(grad_feat + (size_t)id[1] * C + c, g * ww[1]);
    atomicAdd(grad_feat + (size_t)id[2] * C + c, g * ww[2]);
}

// floor(log2(n)) evaluated the way sampling_gpu.cu:9-13 does (double log ratio), capped at 1024 threads
int ref_block_log2(int n) {
    const int pow_2 = (int)(std::log(static_cast<double>(n)) / std::log(2.0));
    int t = 1 << pow_2;
    if (t > 1024) t = 1024;
    if (t < 1) t = 1;
    int l = 0;
    while ((1 << (l + 1)) <= t) ++l;
    return l;
}

template <int PPT>
void launch_fps(int grid, int n, int m, const float* d, float* t, const int* xc, const int* mc, int* idx, int rbl,
                cudaStream_t s) {
    fps_kernel<PPT><<<grid, 1024, 0, s>>>(n, m, d, t, xc, mc, idx, rbl);
}

int fps_dispatch(int grid, int n_max, int n, int m, const float* d, float* t, const int* xc, const int* mc, int* idx,
                 int rbl, cudaStream_t s) {
    if (n_max <= 1024) launch_fps<1>(grid, n, m, d, t, xc, mc, idx, rbl, s);
    else if (n_max <= 2048) launch_fps<2>(grid, n, m, d, t, xc, mc, idx, rbl, s);
    else if (n_max <= 4096) launch_fps<4>(grid, n, m, d, t, xc, mc, idx, rbl, s);
    else if (n_max <= FPS_CS * 1024 * 1) fps_cluster_kernel<1><<<grid * FPS_CS, 1024, 0, s>>>(n, m, d, t, xc, mc, idx, rbl);
    else if (n_max <= FPS_CS * 1024 * 2) fps_cluster_kernel<2><<<grid * FPS_CS, 1024, 0, s>>>(n, m, d, t, xc, mc, idx, rbl);
    else if (n_max <= FPS_CS * 1024 * 3) fps_cluster_kernel<3><<<grid * FPS_CS, 1024, 0, s>>>(n, m, d, t, xc, mc, idx, rbl);
    else if (n_max <= FPS_CS * 1024 * 4) fps_cluster_kernel<4><<<grid * FPS_CS, 1024, 0, s>>>(n, m, d, t, xc, mc, idx, rbl);
    else if (n_max <= FPS_CS * 1024 * 8) fps_cluster_kernel<8><<<grid * FPS_CS, 1024, 0, s>>>(n, m, d, t, xc, mc, idx, rbl);
    else launch_fps<0>(grid, n, m, d, t, xc, mc, idx, rbl, s);
    CRB3D_CHECK_LAUNCH();
    return CRB3D_OK;
}

}  // namespace

// idx (M, nsample) int32, caller-zeroed (reference contract: rows without a hit keep zeros except idx[0] = -1).
// max_queries_per_frame bounds the grid (<= M).
extern "C" int crb3d_ball_query_stack(int B, int M, float radius, int nsample, const float* new_xyz,
                                      const int* new_xyz_batch_cnt, const float* xyz, const int* xyz_batch_cnt,
                                      int* idx, int max_queries_per_frame, cudaStream_t stream) {
    if (B <= 0 || M < 0 || nsample <= 0 || !idx) return CRB3D_ERR_ARG;
    if (M == 0) return CRB3D_OK;
    const int mq = (max_queries_per_frame > 0 && max_queries_per_frame < M) ? max_queries_per_frame : M;
    dim3 grid((unsigned)crb3d_divup(mq, BQ_WARPS), B);
    ball_query_kernel<<<grid, BQ_WARPS * 32, 0, stream>>>(B, M, radius, nsample, new_xyz, new_xyz_batch_cnt, xyz,
                                                         xyz_batch_cnt, idx);
    CRB3D_CHECK_LAUNCH();
    return CRB3D_OK;
}

extern "C" int crb3d_group_points_stack(int B, int M, int C, int nsample, const float* features,
                                        const int* features_batch_cnt, const int* idx, const int* idx_batch_cnt,
                                        float* out, cudaStream_t stream) {
    if (B <= 0 || M < 0 || C <= 0 || nsample <= 0) return CRB3D_ERR_ARG;
    if (M == 0) return CRB3D_OK;
    size_t smem = sizeof(float) * nsample * (C + 1);
    if (smem > 200 * 1024) return CRB3D_ERR_UNSUPPORTED;
    if (smem > 48 * 1024) CRB3D_CUDA(cudaFuncSetAttribute(group_points_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    group_points_kernel<<<M, 256, smem, stream>>>(B, M, C, nsample, features, features_batch_cnt, idx, idx_batch_cnt, out);
    CRB3D_CHECK_LAUNCH();
    return CRB3D_OK;
}

extern "C" int crb3d_group_points_grad_stack(int B, int M, int C, int N, int nsample, const float* grad_out,
                                             const int* idx, const int* idx_batch_cnt, const int* features_batch_cnt,
                                             float* grad_features, cudaStream_t stream) {
    if (B <= 0 || M < 0 || C <= 0 || nsample <= 0) return CRB3D_ERR_ARG;
    (void)N;
    if (M == 0) return CRB3D_OK;
    size_t smem = sizeof(float) * nsample * (C + 1);
    if (smem > 200 * 1024) return CRB3D_ERR_UNSUPPORTED;
    if (smem > 48 * 1024) CRB3D_CUDA(cudaFuncSetAttribute(group_points_grad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    group_points_grad_kernel<<<M, 256, smem, stream>>>(B, M, C, nsample, grad_out, idx, idx_batch_cnt, features_batch_cnt,
                                                      grad_features);
    CRB3D_CHECK_LAUNCH();
    return CRB3D_OK;
}

// dataset (b,n,3), temp (b,n) pre-filled 1e10 by the caller, idx (b,m).
extern "C" int crb3d_farthest_point_sampling(int b, int n, int m, const float* dataset, float* temp, int* idx,
                                             cudaStream_t stream) {
    if (b < 0 || n <= 0 || m < 0 || !dataset || !temp || !idx) return CRB3D_ERR_ARG;
    if (b == 0 || m == 0) return CRB3D_OK;
    if (n >= (1 << 21) * 1) return CRB3D_ERR_UNSUPPORTED;
    return fps_dispatch(b, n, n, m, dataset, temp, nullptr, nullptr, idx, ref_block_log2(n), stream);
}

// stacked: dataset (N1+N2+...,3), temp (N1+...), xyz_batch_cnt (B), num_sampled (B), idx (sum M) GLOBAL indices.
extern "C" int crb3d_stack_farthest_point_sampling(int B, int n_max, const float* dataset, float* temp,
                                                   const int* xyz_batch_cnt, int* idx, const int* num_sampled_points,
                                                   cudaStream_t stream) {
    if (B <= 0 || !dataset || !temp || !idx || !xyz_batch_cnt || !num_sampled_points) return CRB3D_ERR_ARG;
    if (n_max >= (1 << 21)) return CRB3D_ERR_UNSUPPORTED;
    return fps_dispatch(B, n_max, 0, 0, dataset, temp, xyz_batch_cnt, num_sampled_points, idx, 10, stream);
}

extern "C" int crb3d_three_nn_stack(int B, int N, int M, const float* unknown, const int* unknown_batch_cnt,
                                    const float* known, const int* known_batch_cnt, float* dist2, int* idx,
                                    cudaStream_t stream) {
    if (B <= 0 || N < 0 || M < 0) return CRB3D_ERR_ARG;
    if (N == 0) return CRB3D_OK;
    three_nn_kernel<<<(unsigned)crb3d_divup(N, 256), 256, 0, stream>>>(B, N, unknown, unknown_batch_cnt, known,
                                                                      known_batch_cnt, dist2, idx);
    CRB3D_CHECK_LAUNCH();
    return CRB3D_OK;
}

extern "C" int crb3d_three_interpolate_stack(int N, int C, const float* features, const int* idx, const float* weight,
                                             float* out, cudaStream_t stream) {
    if (N < 0 || C <= 0) return CRB3D_ERR_ARG;
    if (N == 0) return CRB3D_OK;
    three_interp_kernel<<<(unsigned)crb3d_divup((int64_t)N * C, 256), 256, 0, stream>>>(N, C, features, idx, weight, out);
    CRB3D_CHECK_LAUNCH();
    return CRB3D_OK;
}

extern "C" int crb3d_three_interpolate_grad_stack(int N, int C, const float* grad_out, const int* idx,
                                                  const float* weight, float* grad_features, cudaStream_t stream) {
    if (N < 0 || C <= 0) return CRB3D_ERR_ARG;
    if (N == 0) return CRB3D_OK;
    three_interp_grad_kernel<<<(unsigned)crb3d_divup((int64_t)N * C, 256), 256, 0, stream>>>(N, C, grad_out, idx, weight,
                                                                                            grad_features);
    CRB3D_CHECK_LAUNCH();
    return CRB3D_OK;
}
