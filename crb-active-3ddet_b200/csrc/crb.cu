// CRB acquisition scoring kernels: stage-1 label entropy, stage-2 pairwise squared distances of gradient
// embeddings (input of k-means++), stage-3 greedy KDE/KL density balancing.
//
// Replaces (reference, /root/reference/pcdet/query_strategies/crb_sampling.py):
//   :86-100   torch.unique + Categorical(probs).entropy() per frame (absent classes get pseudo-count 1)
//   :219-226  sklearn kmeans_plusplus on the (K1*N_r, 65536) gradient matrix  -> we provide the distance matrix
//   :276-338  greedy loop: per remaining candidate and class, sklearn KernelDensity(gaussian, h).fit/score_samples on
//             400 axis points, scipy.stats.entropy(uniform, exp(logprob)), 2/pi*atan(pi/2*KL), mean_c(1-.), first max
// The reference refits a KDE from scratch for every (round, candidate, class) with a GPU->CPU copy in the innermost
// loop. Here every candidate's per-class log-kernel sums over the 400-point axis are computed ONCE (log-sum-exp
// form, fp64, identical to sklearn's log-density up to rounding); each greedy round then only combines the running
// "selected" sums with the candidate sums - O(N_r * K2N_r * C * 400) flops total, no host round trip.
#include "common.cuh"
#include <math.h>

namespace {

constexpr int AXIS = 400;  // crb_sampling.py:259 np.linspace(..., 400)

// ------------------------------------------------------------------ stage 1: label entropy
__global__ void __launch_bounds__(128) label_entropy_kernel(const int* __restrict__ labels, const int* __restrict__ box_off,
                                                            const int* __restrict__ box_end, int B, int num_class,
                                                            float* __restrict__ entropy,
                                                            int* __restrict__ class_counts /*B x num_class, optional*/) {
    extern __shared__ int cnt[];  // num_class
    const int b = blockIdx.x;
    for (int c = threadIdx.x; c < num_class; c += blockDim.x) cnt[c] = 0;
    __syncthreads();
    const int s = box_off[b], e = box_end[b];
    for (int i = s + threadIdx.x; i < e; i += blockDim.x) {
        int l = labels[i] - 1;
        if (l >= 0 && l < num_class) atomicAdd(&cnt[l], 1);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        float h = 0.0f;
        if (e > s) {
            long long total = 0;
            for (int c = 0; c < num_class; ++c) total += cnt[c];
            // unique_proportions = ones; [value-1] = counts; probs = unique_proportions / sum(counts)
            // Categorical(probs): probs /= probs.sum(-1); logits = log(clamp(probs, eps, 1-eps)); H = -sum(p * logit)
            float psum = 0.0f;
            for (int c = 0; c < num_class; ++c) psum += __fdiv_rn(cnt[c] > 0 ? (float)cnt[c] : 1.0f, (float)total);
            const float eps = 1.1920928955078125e-07f;
            for (int c = 0; c < num_class; ++c) {
                float p = __fdiv_rn(__fdiv_rn(cnt[c] > 0 ? (float)cnt[c] : 1.0f, (float)total), psum);
                float pc = fminf(fmaxf(p, eps), 1.0f - eps);
                h += __fmul_rn(logf(pc), p);
            }
            h = -h;
        }
        entropy[b] = h;
        if (class_counts)
            for (int c = 0; c < num_class; ++c) class_counts[b * num_class + c] = cnt[c];
    }
}

// ------------------------------------------------------------------ stage 2: pairwise squared distances
// D[i][j] = |x_i|^2 + |x_j|^2 - 2 x_i.x_j accumulated in fp64 from fp32 inputs (sklearn upcasts float32 chunks to
// float64 in euclidean_distances). 32x32 output tile per CTA, K streamed in slabs of 32.
__global__ void __launch_bounds__(256) sqdist_kernel(const float* __restrict__ X, int n, int d, double* __restrict__ D) {
    __shared__ float As[32][33], Bs[32][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // ty 0..7 -> rows ty*4..ty*4+3
    const int i0 = blockIdx.y * 32, j0 = blockIdx.x * 32;
    if (j0 < i0) return;  // symmetric: compute upper tiles, mirror on store
    double acc[4] = {0, 0, 0, 0}, na[4] = {0, 0, 0, 0}, nb = 0;
    for (int k0 = 0; k0 < d; k0 += 32) {
        for (int r = ty; r < 32; r += 8) {
            As[r][tx] = (i0 + r < n && k0 + tx < d) ? X[(size_t)(i0 + r) * d + k0 + tx] : 0.f;
            Bs[r][tx] = (j0 + r < n && k0 + tx < d) ? X[(size_t)(j0 + r) * d + k0 + tx] : 0.f;
        }
        __syncthreads();
#pragma unroll 8
        for (int k = 0; k < 32; ++k) {
            const double bv = (double)Bs[tx][k];
            nb += bv * bv;
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const double av = (double)As[ty * 4 + r][k];
                acc[r] += av * bv;
                na[r] += av * av;
            }
        }
        __syncthreads();
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const int i = i0 + ty * 4 + r, j = j0 + tx;
        if (i < n && j < n) {
            double v = na[r] + nb - 2.0 * acc[r];
            if (v < 0) v = 0;
            if (i == j) v = 0;
            D[(size_t)i * n + j] = v;
            D[(size_t)j * n + i] = v;
        }
    }
}

// ------------------------------------------------------------------ stage 3: KDE / KL greedy
struct LSE { double m, s; };  // sum = exp(m) * s ; empty = (-inf, 0)

__device__ __forceinline__ LSE lse_add(LSE a, LSE b) {
    if (b.s == 0.0) return a;
    if (a.s == 0.0) return b;
    LSE r;
    if (a.m >= b.m) { r.m = a.m; r.s = a.s + b.s * exp(b.m - a.m); }
    else { r.m = b.m; r.s = b.s + a.s * exp(a.m - b.m); }
    return r;
}

// candidate sums: grid (n_cand, n_class), block 128
__global__ void __launch_bounds__(128) kde_candidate_kernel(const float* __restrict__ dens, const int* __restrict__ labels,
                                                            const int* __restrict__ cand_off, int n_class,
                                                            const double* __restrict__ axis, double bandwidth,
                                                            double* __restrict__ Bm, double* __restrict__ Bs,
                                                            int* __restrict__ cand_cnt) {
    const int i = blockIdx.x, c = blockIdx.y;
    const int s = cand_off[i], e = cand_off[i + 1];
    __shared__ int n_s;
    if (threadIdx.x == 0) n_s = 0;
    __syncthreads();
    int local = 0;
    for (int t = s + threadIdx.x; t < e; t += blockDim.x) local += (labels[t] == c + 1);
    if (local) atomicAdd(&n_s, local);
    __syncthreads();
    if (threadIdx.x == 0) cand_cnt[i * n_class + c] = n_s;
    const double inv_h2 = 1.0 / (bandwidth * bandwidth);
    for (int x = threadIdx.x; x < AXIS; x += blockDim.x) {
        const double ax = axis[c * AXIS + x];
        LSE acc; acc.m = -INFINITY; acc.s = 0.0;
        for (int t = s; t < e; ++t) {
            if (labels[t] != c + 1) continue;
            const double dd = ax - (double)dens[t];
            LSE one; one.m = -0.5 * (dd * dd) * inv_h2; one.s = 1.0;
            acc = lse_add(acc, one);
        }
        const size_t o = ((size_t)i * n_class + c) * AXIS + x;
        Bm[o] = acc.m; Bs[o] = acc.s;
    }
}

__device__ __forceinline__ double block_sum(double v, double* sm) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) sm[warp] = v;
    __syncthreads();
    double t = 0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += sm[w];
    return t;
}

// score of every alive candidate against the current selected sums: grid n_cand, block 128
__global__ void __launch_bounds__(128) kde_score_kernel(int n_class, const double* __restrict__ Am,
                                                        const double* __restrict__ As, const int* __restrict__ n_sel,
                                                        const double* __restrict__ Bm, const double* __restrict__ Bs,
                                                        const int* __restrict__ cand_cnt, const int* __restrict__ alive,
                                                        const double* __restrict__ prior_n /*normalised pk*/,
                                                        double bandwidth, double* __restrict__ score) {
    __shared__ double sm[4];
    const int i = blockIdx.x;
    if (!alive[i]) { if (threadIdx.x == 0) score[i] = -INFINITY; return; }
    double mean_acc = 0.0;
    for (int c = 0; c < n_class; ++c) {
        const int nc = cand_cnt[i * n_class + c];
        double prop;
        if (nc == 0) {
            prop = 1.0;  // crb_sampling.py:298-300
        } else {
            const double log_norm = log((double)(n_sel[c] + nc) * bandwidth * sqrt(2.0 * M_PI));
            double q[(AXIS + 127) / 128];
            double qs = 0.0;
            int u = 0;
            for (int x = threadIdx.x; x < AXIS; x += blockDim.x, ++u) {
                LSE a; a.m = Am[c * AXIS + x]; a.s = As[c * AXIS + x];
                LSE b; const size_t o = ((size_t)i * n_class + c) * AXIS + x; b.m = Bm[o]; b.s = Bs[o];
                LSE r = lse_add(a, b);
                q[u] = exp(r.m + log(r.s) - log_norm);
                qs += q[u];
            }
            qs = block_sum(qs, sm);
            double kl = 0.0;
            u = 0;
            for (int x = threadIdx.x; x < AXIS; x += blockDim.x, ++u) {
                const double pk = prior_n[c * AXIS + x];
                const double qk = q[u] / qs;
                double v;
                if (pk > 0.0 && qk > 0.0) v = pk * log(pk / qk);
                else if (pk == 0.0 && qk >= 0.0) v = 0.0;
                else v = (pk != pk || qk != qk) ? NAN : INFINITY;  // scipy rel_entr
                kl += v;
            }
            kl = block_sum(kl, sm);
            prop = 2.0 / M_PI * atan(M_PI / 2.0 * kl);
        }
        mean_acc += 1.0 - prop;
    }
    if (threadIdx.x == 0) score[i] = mean_acc / (double)n_class;
}

// pick (first strictly-greater maximum, best initialised to -1) and fold the winner into the selected sums.
// round 0 (force_first = 1) takes candidate 0 unconditionally (crb_sampling.py:277-286).
__global__ void __launch_bounds__(256) kde_pick_kernel(int n_cand, int n_class, int round, int force_first,
                                                       const double* __restrict__ score, int* __restrict__ alive,
                                                       double* __restrict__ Am, double* __restrict__ As,
                                                       int* __restrict__ n_sel, const double* __restrict__ Bm,
                                                       const double* __restrict__ Bs, const int* __restrict__ cand_cnt,
                                                       int* __restrict__ order, double* __restrict__ picked_score) {
    __shared__ double bv[256];
    __shared__ int bi[256];
    __shared__ int win_s;
    const int tid = threadIdx.x;
    if (force_first) {
        if (tid == 0) win_s = 0;
    } else {
        double best = -1.0; int besti = -1;
        for (int i = tid; i < n_cand; i += blockDim.x)
            if (alive[i] && score[i] > best) { best = score[i]; besti = i; }
        bv[tid] = best; bi[tid] = besti;
        __syncthreads();
        for (int o = 128; o > 0; o >>= 1) {
            if (tid < o) {
                const double v2 = bv[tid + o]; const int i2 = bi[tid + o];
                const bool take = (i2 >= 0) && (bi[tid] < 0 || v2 > bv[tid] || (v2 == bv[tid] && i2 < bi[tid]));
                if (take) { bv[tid] = v2; bi[tid] = i2; }
            }
            __syncthreads();
        }
        if (tid == 0) win_s = bi[0];
    }
    __syncthreads();
    const int w = win_s;
    if (tid == 0) {
        order[round] = w;
        if (picked_score) picked_score[round] = (force_first || w < 0) ? NAN : score[w];
    }
    if (w < 0) return;  // no candidate beat -1 (all NaN): the reference raises here; host checks order[round] < 0
    for (int t = tid; t < n_class * AXIS; t += blockDim.x) {
        const int c = t / AXIS, x = t - c * AXIS;
        LSE a; a.m = Am[t]; a.s = As[t];
        LSE b; const size_t o = ((size_t)w * n_class + c) * AXIS + x; b.m = Bm[o]; b.s = Bs[o];
        LSE r = lse_add(a, b);
        Am[t] = r.m; As[t] = r.s;
    }
    if (tid < n_class) n_sel[tid] += cand_cnt[w * n_class + tid];
    if (tid == 0) alive[w] = 0;
}

}  // namespace

// labels: int32 (1-based) stacked over frames, box_off (B+1). entropy: float[B]; class_counts optional (B*num_class).
extern "C" int crb3d_label_entropy(const int* labels, const int* box_off, int B, int num_class, float* entropy,
                                   int* class_counts, cudaStream_t stream) {
    if (B < 0 || num_class <= 0 || num_class > 4096 || !box_off || !entropy) return CRB3D_ERR_ARG;
    if (B == 0) return CRB3D_OK;
    label_entropy_kernel<<<B, 128, sizeof(int) * num_class, stream>>>(labels, box_off, box_off + 1, B, num_class, entropy, class_counts);
    CRB3D_CHECK_LAUNCH();
    return CRB3D_OK;
}

// explicit [begin, end) ranges per frame (padded label tensors with per-frame counts)
extern "C" int crb3d_label_entropy_ranges(const int* labels, const int* box_begin, const int* box_end, int B,
                                          int num_class, float* entropy, int* class_counts, cudaStream_t stream) {
    if (B < 0 || num_class <= 0 || num_class > 4096 || !box_begin || !box_end || !entropy) return CRB3D_ERR_ARG;
    if (B == 0) return CRB3D_OK;
    label_entropy_kernel<<<B, 128, sizeof(int) * num_class, stream>>>(labels, box_begin, box_end, B, num_class, entropy, class_counts);
    CRB3D_CHECK_LAUNCH();
    return CRB3D_OK;
}

extern "C" int crb3d_pairwise_sqdist_f64(const float* X, int n, int d, double* D, cudaStream_t stream) {
    if (n < 0 || d <= 0 || !X || !D) return CRB3D_ERR_ARG;
    if (n == 0) return CRB3D_OK;
    dim3 grid((unsigned)crb3d_divup(n, 32), (unsigned)crb3d_divup(n, 32));
    sqdist_kernel<<<grid, 256, 0, stream>>>(X, n, d, D);
    CRB3D_CHECK_LAUNCH();
    return CRB3D_OK;
}

extern "C" int crb3d_kde_greedy_workspace_bytes(int n_cand, int n_class, size_t* bytes) {
    if (!bytes || n_cand < 0 || n_class <= 0) return CRB3D_ERR_ARG;
    size_t per = (size_t)n_class * AXIS;
    *bytes = 2 * crb3d_align(sizeof(double) * per * (size_t)(n_cand > 0 ? n_cand : 1)) + 2 * crb3d_align(sizeof(double) * per) +
             crb3d_align(sizeof(int) * (size_t)(n_cand > 0 ? n_cand : 1) * n_class) + crb3d_align(sizeof(int) * n_class) +
             crb3d_align(sizeof(int) * (size_t)(n_cand > 0 ? n_cand : 1)) + crb3d_align(sizeof(double) * (size_t)(n_cand > 0 ? n_cand : 1));
    return CRB3D_OK;
}

// Greedy density balancing over n_cand candidate frames (stacked densities/labels, cand_off n_cand+1).
// axis: (n_class, 400) fp64 evaluation points; prior_n: (n_class, 400) fp64 uniform pdf ALREADY normalised to sum 1
// per class (scipy.stats.entropy normalises pk). Selects n_select frames; order[j] = candidate index picked in round j
// (-1 = no candidate beat the reference's initial best of -1); picked_score optional (n_select doubles).
extern "C" int crb3d_kde_greedy(const float* dens, const int* labels, const int* cand_off, int n_cand, int n_class,
                                const double* axis, const double* prior_n, double bandwidth, int n_select, int* order,
                                double* picked_score, void* ws, size_t ws_bytes, cudaStream_t stream) {
    if (n_cand <= 0 || n_class <= 0 || n_select <= 0 || n_select > n_cand || !cand_off || !axis || !prior_n || !order)
        return CRB3D_ERR_ARG;
    WsCursor c(ws, ws_bytes);
    const size_t per = (size_t)n_class * AXIS;
    double* Bm = c.take<double>(per * n_cand);
    double* Bs = c.take<double>(per * n_cand);
    double* Am = c.take<double>(per);
    double* As = c.take<double>(per);
    int* cand_cnt = c.take<int>((size_t)n_cand * n_class);
    int* n_sel = c.take<int>(n_class);
    int* alive = c.take<int>(n_cand);
    double* score = c.take<double>(n_cand);
    if (!c.ok) return CRB3D_ERR_WORKSPACE;
    kde_candidate_kernel<<<dim3(n_cand, n_class), 128, 0, stream>>>(dens, labels, cand_off, n_class, axis, bandwidth, Bm, Bs, cand_cnt);
    // empty selected set: m = -inf (bit pattern 0xFFF0...) , s = 0
    {
        int rc = crb3d_fill_i32(alive, n_cand, 1, stream); if (rc) return rc;
        CRB3D_CUDA(cudaMemsetAsync(As, 0, sizeof(double) * per, stream));
        CRB3D_CUDA(cudaMemsetAsync(Am, 0, sizeof(double) * per, stream));  // value irrelevant while s == 0
        CRB3D_CUDA(cudaMemsetAsync(n_sel, 0, sizeof(int) * n_class, stream));
    }
    for (int j = 0; j < n_select; ++j) {
        if (j > 0) kde_score_kernel<<<n_cand, 128, 0, stream>>>(n_class, Am, As, n_sel, Bm, Bs, cand_cnt, alive, prior_n, bandwidth, score);
        kde_pick_kernel<<<1, 256, 0, stream>>>(n_cand, n_class, j, j == 0, score, alive, Am, As, n_sel, Bm, Bs, cand_cnt, order, picked_score);
    }
    CRB3D_CHECK_LAUNCH();
    return CRB3D_OK;
}
