// Anchor-head post-processing kernels: per-anchor max-class score/label and lazy box decoding.
//
// Replaces, on the scoring path, the dense torch ops of
//   pcdet/models/dense_heads/anchor_head_template.py:238-285  generate_predicted_boxes (all 211 200 anchors/frame)
//   pcdet/utils/box_coder_utils.py:45-77                      ResidualCoder.decode_torch
//   pcdet/models/detectors/detector3d_template.py:281-311     sigmoid + max over classes (+1 for the label)
// The reference decodes every anchor and then keeps <= 4096 of them; here the score/label pass reads the class logits
// once, and boxes are decoded only for the anchors that survive the score filter + top-k (same values, 50x less work).
// Anchor layout (anchor_generator.py:18-62 + torch.cat(dim=-3) in generate_predicted_boxes): anchor index
// a = ((y*nx + x)*n_class_sets + set)*n_rot + rot ; logits / box codes / dir bins are channels-last per location.
#include "common.cuh"
#include <string.h>

#define CRB3D_MAX_ANCHOR_TYPES 16

struct AnchorSpec {
    int nx, ny, n_types;          // feature-map size and anchors per location
    float x0, y0;                 // first anchor centre
    double x_stride, y_stride;    // centre spacing (float64 like torch.arange's accumulator)
    float size[CRB3D_MAX_ANCHOR_TYPES][3];  // dx, dy, dz
    float rot[CRB3D_MAX_ANCHOR_TYPES];
    float zc[CRB3D_MAX_ANCHOR_TYPES];       // bottom height + dz/2
    float dir_offset, dir_limit_offset;
    int num_dir_bins;             // 0 = no direction classifier
};

namespace {

__global__ void __launch_bounds__(256) head_scores_kernel(const float* __restrict__ cls, int64_t n_anchor_total,
                                                          int n_class, float* __restrict__ score, int* __restrict__ label) {
    const int64_t a = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= n_anchor_total) return;
    const float* p = cls + a * n_class;
    float best = -INFINITY;
    int bi = 0;
    for (int c = 0; c < n_class; ++c) {
        const float v = p[c];
        if (v > best) { best = v; bi = c; }  // torch.max: first maximum
    }
    // sigmoid is monotonic: max(sigmoid(x)) == sigmoid(max(x))
    score[a] = 1.0f / (1.0f + expf(-best));
    label[a] = bi + 1;
}

// ------------------------------------------------------------------ score >= thresh candidates + sorted top-k
// class_agnostic_nms (model_nms_utils.py:6-25) keeps the anchors with score >= SCORE_THRESH and takes topk(NMS_PRE_MAXSIZE)
// of them. The reference does a full torch.topk over all 211 200 anchors per frame (multi-pass radix select + sort, ~20
// launches); here the score pass itself appends the few thousand candidates as 64-bit keys (score bits << 32 | ~index:
// descending key order = descending score, ties by ascending anchor index), and one CTA per frame selects the K largest
// (only when more than K pass the threshold: MSD radix select, 11-bit digits) and bitonic-sorts them in shared memory.
constexpr int TOPK_BINS = 2048;

__global__ void __launch_bounds__(256) head_scores_cand_kernel(const float* __restrict__ cls, int64_t n_anchor_total, int n_class,
                                                               int64_t n_per_frame, float thresh, float* __restrict__ score,
                                                               int* __restrict__ label, unsigned long long* __restrict__ cand,
                                                               int* __restrict__ cand_count, int* __restrict__ hist,
                                                               unsigned int thresh_bits, int hshift) {
    // candidates are appended with ONE global atomic per (block, frame): a per-candidate atomicAdd on the frame counter
    // serialises at its L2 slice (measured 530 us for 4 x 60 k candidates)
    __shared__ int cnt[2], base[2];
    const int64_t a0 = (int64_t)blockIdx.x * blockDim.x;
    const int64_t a = a0 + threadIdx.x;
    const int f0 = (int)(a0 / n_per_frame);          // a block spans at most two frames (n_per_frame >= blockDim.x)
    if (threadIdx.x < 2) cnt[threadIdx.x] = 0;
    __syncthreads();
    bool pred = false;
    int slot = 0, mypos = 0;
    unsigned long long key = 0ull;
    int f = f0;
    if (a < n_anchor_total) {
        const float* p = cls + a * n_class;
        float best = -INFINITY;
        int bi = 0;
        for (int c = 0; c < n_class; ++c) {
            const float v = p[c];
            if (v > best) { best = v; bi = c; }  // torch.max: first maximum
        }
        const float sc = 1.0f / (1.0f + expf(-best));
        score[a] = sc;
        label[a] = bi + 1;
        if (sc >= thresh) {
            f = (int)(a / n_per_frame);
            slot = f - f0;
            const unsigned int li = (unsigned int)(a - (int64_t)f * n_per_frame);
            key = ((unsigned long long)__float_as_uint(sc) << 32) | (unsigned long long)(0xFFFFFFFFu - li);
            // coarse score histogram (monotone bins of the float bits above the threshold): lets the top-k kernel find its
            // cut with ONE pass over the candidates instead of a multi-pass radix select
            atomicAdd(&hist[(size_t)f * TOPK_BINS + min((int)((__float_as_uint(sc) - thresh_bits) >> hshift), TOPK_BINS - 1)], 1);
            if (n_per_frame < (int64_t)blockDim.x) {   // tiny frames: a block may span many of them - append directly
                cand[(size_t)f * n_per_frame + atomicAdd(&cand_count[f], 1)] = key;
            } else {
                pred = true;
                mypos = atomicAdd(&cnt[slot], 1);
            }
        }
    }
    __syncthreads();
    if (threadIdx.x < 2 && cnt[threadIdx.x] > 0) base[threadIdx.x] = atomicAdd(&cand_count[f0 + threadIdx.x], cnt[threadIdx.x]);
    __syncthreads();
    if (pred) cand[(size_t)f * n_per_frame + base[slot] + mypos] = key;
}

constexpr int TOPK_MAX = 4096;
constexpr int TOPK_THREADS = 1024;

__global__ void __launch_bounds__(TOPK_THREADS) topk_sort_kernel(const unsigned long long* __restrict__ cand,
                                                                 const int* __restrict__ cand_count, int64_t n_per_frame, int K,
                                                                 float* __restrict__ top_score, long long* __restrict__ top_idx,
                                                                 int* __restrict__ counts, const int* __restrict__ ghist,
                                                                 unsigned int thresh_bits, int hshift) {
    __shared__ unsigned long long keys[TOPK_MAX];
    __shared__ unsigned int hist[2048];               // generic path: digit histogram; fast path: boundary list (u64 x 1024)
    __shared__ int scan_s[33];
    __shared__ int d_s, need_s, cnt_s, nsel;
    const int f = blockIdx.x, tid = threadIdx.x;
    const unsigned long long* src = cand + (size_t)f * n_per_frame;
    const int nc = cand_count[f];
    int KP = 1;                                   // sort size: power of two >= min(nc, K)
    while (KP < min(nc, K)) KP <<= 1;
    for (int t = tid; t < TOPK_MAX; t += TOPK_THREADS) keys[t] = 0ull;
    if (tid == 0) nsel = 0;
    __syncthreads();
    bool done_fast = false;
    if (nc <= K) {
        for (int t = tid; t < nc; t += TOPK_THREADS) keys[t] = src[t];
        done_fast = true;
    } else {
        // fast path: the score histogram filled by the score pass gives the boundary bin b (bins above it hold fewer
        // than K candidates, together with b at least K); ONE pass over the candidates takes every key above b and
        // collects the keys of b into a small list whose `need` largest complete the selection
        const int* gh = ghist + (size_t)f * TOPK_BINS;
        const int h0 = gh[2 * tid], h1 = gh[2 * tid + 1];
        int total;
        const int ex = block_excl_scan(h0 + h1, scan_s, &total);
        const int above = total - ex - (h0 + h1);
        if (above < K && above + h1 >= K) { d_s = 2 * tid + 1; need_s = K - above; cnt_s = h1; }
        else if (above + h1 < K && above + h1 + h0 >= K) { d_s = 2 * tid; need_s = K - above - h1; cnt_s = h0; }
        __syncthreads();
        const int bbin = d_s, need = need_s, cnt_b = cnt_s;
        constexpr int BCAP = 1024;
        if (cnt_b <= BCAP) {
            unsigned long long* bl = reinterpret_cast<unsigned long long*>(hist);
            __shared__ int nb;
            if (tid == 0) nb = 0;
            for (int t = tid; t < BCAP; t += TOPK_THREADS) bl[t] = 0ull;
            __syncthreads();
            for (int t0 = 0; t0 < nc; t0 += 4 * TOPK_THREADS) {      // 4 independent loads in flight per thread
                unsigned long long kk[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) { const int t = t0 + u * TOPK_THREADS + tid; kk[u] = t < nc ? src[t] : 0ull; }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    if (t0 + u * TOPK_THREADS + tid >= nc) continue;
                    const int bin = min((int)(((unsigned int)(kk[u] >> 32) - thresh_bits) >> hshift), TOPK_BINS - 1);
                    if (bin > bbin) { const int sl = atomicAdd(&nsel, 1); if (sl < TOPK_MAX) keys[sl] = kk[u]; }
                    else if (bin == bbin) { const int sl = atomicAdd(&nb, 1); if (sl < BCAP) bl[sl] = kk[u]; }
                }
            }
            __syncthreads();
            int BP = 2;
            while (BP < cnt_b) BP <<= 1;
            for (int k = 2; k <= BP; k <<= 1) {
                for (int j = k >> 1; j > 0; j >>= 1) {
                    for (int t = tid; t < BP; t += TOPK_THREADS) {
                        const int u = t ^ j;
                        if (u > t) {
                            const unsigned long long a = bl[t], b = bl[u];
                            if ((a < b) == ((t & k) == 0)) { bl[t] = b; bl[u] = a; }
                        }
                    }
                    __syncthreads();
                }
            }
            const int base = nsel;                     // == K - need
            for (int t = tid; t < need; t += TOPK_THREADS) if (base + t < TOPK_MAX) keys[base + t] = bl[t];
            done_fast = true;
        }
        __syncthreads();
    }
    if (!done_fast) {
        if (tid == 0) nsel = 0;
        for (int t = tid; t < TOPK_MAX; t += TOPK_THREADS) keys[t] = 0ull;
        __syncthreads();
        // MSD radix select of the K largest keys (keys are unique): fix 11 (last pass: 9) bits per pass
        // bits shared by ALL keys are fixed up front (scores in [thresh, 1) share their exponent bits: without this the
        // first passes would pile every key into one or two histogram bins)
        unsigned long long k_and = ~0ull, k_or = 0ull;
        for (int t = tid; t < nc; t += TOPK_THREADS) { const unsigned long long k = src[t]; k_and &= k; k_or |= k; }
        for (int o = 16; o > 0; o >>= 1) {
            k_and &= __shfl_xor_sync(0xffffffffu, k_and, o);
            k_or |= __shfl_xor_sync(0xffffffffu, k_or, o);
        }
        __shared__ unsigned long long red_and[32], red_or[32];
        if ((tid & 31) == 0) { red_and[tid >> 5] = k_and; red_or[tid >> 5] = k_or; }
        __syncthreads();
        k_and = red_and[0]; k_or = red_or[0];
        for (int wv = 1; wv < TOPK_THREADS / 32; ++wv) { k_and &= red_and[wv]; k_or |= red_or[wv]; }
        const unsigned long long diff = k_and ^ k_or;             // bit set = keys differ there
        int fixed = diff ? __clzll((long long)diff) : 63;          // length of the common prefix (keys are unique: diff != 0)
        unsigned long long prefix = fixed ? (k_and >> (64 - fixed)) : 0ull;   // value of the bits fixed so far
        int need = K;
        while (fixed < 64) {
            const int w = min(11, 64 - fixed), shift = 64 - fixed - w;
            for (int t = tid; t < 2048; t += TOPK_THREADS) hist[t] = 0u;
            __syncthreads();
            for (int t = tid; t < nc; t += TOPK_THREADS) {
                const unsigned long long k = src[t];
                if (fixed == 0 || (k >> (64 - fixed)) == prefix) atomicAdd(&hist[(unsigned int)(k >> shift) & ((1u << w) - 1u)], 1u);
            }
            __syncthreads();
            const int h0 = (int)hist[2 * tid], h1 = (int)hist[2 * tid + 1];   // bins ascending; (1 << w) <= 2048
            int total;
            const int ex = block_excl_scan(h0 + h1, scan_s, &total);
            const int above = total - ex - (h0 + h1);     // keys (with the fixed prefix) in bins above this thread's pair
            if (above < need && above + h1 >= need) { d_s = 2 * tid + 1; need_s = need - above; cnt_s = h1; }
            else if (above + h1 < need && above + h1 + h0 >= need) { d_s = 2 * tid; need_s = need - above - h1; cnt_s = h0; }
            __syncthreads();
            prefix = (prefix << w) | (unsigned long long)d_s;
            fixed += w;
            need = need_s;
            const bool done = cnt_s == need;       // the boundary bin is taken whole: threshold found
            __syncthreads();
            if (done) break;
        }
        for (int t = tid; t < nc; t += TOPK_THREADS) {
            const unsigned long long k = src[t];
            if ((k >> (64 - fixed)) >= prefix) { const int s = atomicAdd(&nsel, 1); if (s < TOPK_MAX) keys[s] = k; }
        }
    }
    __syncthreads();
    for (int k = 2; k <= KP; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int t = tid; t < KP; t += TOPK_THREADS) {
                const int u = t ^ j;
                if (u > t) {
                    const unsigned long long a = keys[t], b = keys[u];
                    const bool desc = (t & k) == 0 || k == KP;
                    if ((a < b) == desc) { keys[t] = b; keys[u] = a; }
                }
            }
            __syncthreads();
        }
    }
    const int nv = min(nc, K);
    for (int t = tid; t < K; t += TOPK_THREADS) {
        const unsigned long long k = t < nv ? keys[t] : 0ull;
        top_score[(size_t)f * K + t] = t < nv ? __uint_as_float((unsigned int)(k >> 32)) : 0.0f;
        top_idx[(size_t)f * K + t] = t < nv ? (long long)(0xFFFFFFFFu - (unsigned int)k) : 0ll;
    }
    if (tid == 0) counts[f] = nv;
}

__global__ void __launch_bounds__(128) head_decode_kernel(const float* __restrict__ box, const float* __restrict__ dir,
                                                          const long long* __restrict__ sel, int B, int K,
                                                          int64_t n_anchor_per_frame, AnchorSpec S,
                                                          float* __restrict__ out) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= B * K) return;
    const int b = t / K;
    long long a = sel[t];
    if (a < 0 || a >= n_anchor_per_frame) a = 0;
    const int type = (int)(a % S.n_types);
    const long long loc = a / S.n_types;
    const int x = (int)(loc % S.nx), y = (int)(loc / S.nx);
    const float xa = (float)((double)S.x0 + S.x_stride * x), ya = (float)((double)S.y0 + S.y_stride * y);
    const float za = S.zc[type], dxa = S.size[type][0], dya = S.size[type][1], dza = S.size[type][2], ra = S.rot[type];
    const float* e = box + ((size_t)b * n_anchor_per_frame + a) * 7;
    const float diag = sqrtf(dxa * dxa + dya * dya);
    float* o = out + (size_t)t * 7;
    o[0] = e[0] * diag + xa;
    o[1] = e[1] * diag + ya;
    o[2] = e[2] * dza + za;
    o[3] = expf(e[3]) * dxa;
    o[4] = expf(e[4]) * dya;
    o[5] = expf(e[5]) * dza;
    float rg = e[6] + ra;
    if (S.num_dir_bins > 0 && dir) {
        const float* d = dir + ((size_t)b * n_anchor_per_frame + a) * S.num_dir_bins;
        int dl = 0;
        float bd = d[0];
        for (int q = 1; q < S.num_dir_bins; ++q)
            if (d[q] > bd) { bd = d[q]; dl = q; }
        const float period = (float)(2.0 * 3.14159265358979323846 / S.num_dir_bins);
        const float v = rg - S.dir_offset;
        // common_utils.limit_period: val - floor(val / period + offset) * period
        const float dir_rot = v - floorf(v / period + S.dir_limit_offset) * period;
        rg = dir_rot + S.dir_offset + period * (float)dl;
    }
    o[6] = rg;
}

// gather helper: out[b][k][:] = src[b][idx[b][k]][:] for int/float rows (used to pick labels / boxes by keep lists)
template <typename T>
__global__ void __launch_bounds__(256) gather_rows_kernel(const T* __restrict__ src, const long long* __restrict__ idx,
                                                          const int* __restrict__ valid, int B, int K, int64_t n_src,
                                                          int width, T fill, T* __restrict__ out) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (int64_t)B * K * width) return;
    const int w = (int)(t % width);
    const int64_t bk = t / width;
    const int b = (int)(bk / K), k = (int)(bk % K);
    T v = fill;
    if (!valid || k < valid[b]) {
        long long i = idx[bk];
        if (i >= 0 && i < n_src) v = src[((size_t)b * n_src + i) * width + w];
    }
    out[t] = v;
}

}  // namespace

extern "C" int crb3d_anchor_head_scores(const float* cls_preds, int64_t n_anchor_total, int n_class, float* score,
                                        int* label, cudaStream_t stream) {
    if (n_anchor_total < 0 || n_class <= 0 || !score || !label) return CRB3D_ERR_ARG;
    if (n_anchor_total == 0) return CRB3D_OK;
    head_scores_kernel<<<(unsigned)crb3d_divup(n_anchor_total, 256), 256, 0, stream>>>(cls_preds, n_anchor_total, n_class,
                                                                                      score, label);
    CRB3D_CHECK_LAUNCH();
    return CRB3D_OK;
}

// Scores + labels of every anchor AND the sorted top-K (K <= 4096) of the anchors with score >= thresh, per frame.
// cand: (B, n_per_frame) uint64 scratch; cand_count: B * 2049 int scratch (zeroed here: counts + score histograms); top_score/top_idx: (B, K), valid prefix
// = counts[b] = min(#candidates, K), the rest is 0.
extern "C" int crb3d_anchor_head_scores_topk(const float* cls_preds, int B, int64_t n_per_frame, int n_class, float thresh, int K,
                                             float* score, int* label, unsigned long long* cand, int* cand_count,
                                             float* top_score, long long* top_idx, int* counts, cudaStream_t stream) {
    if (B < 0 || n_per_frame <= 0 || n_class <= 0 || K <= 0 || !score || !label || !cand || !cand_count || !top_score ||
        !top_idx || !counts)
        return CRB3D_ERR_ARG;
    if (K > TOPK_MAX || n_per_frame >= 0xFFFFFFFFll) return CRB3D_ERR_UNSUPPORTED;
    if (B == 0) return CRB3D_OK;
    // cand_count scratch layout: [B] candidate counts, then [B][TOPK_BINS] score histograms
    CRB3D_CUDA(cudaMemsetAsync(cand_count, 0, sizeof(int) * (size_t)B * (1 + TOPK_BINS), stream));
    int* hist = cand_count + B;
    unsigned int tb, one;
    { float t = thresh > 0.f ? thresh : 0.f, o = 1.0f; memcpy(&tb, &t, 4); memcpy(&one, &o, 4); }
    int hshift = 0;
    while (((one - tb) >> hshift) >= (unsigned int)TOPK_BINS) ++hshift;
    head_scores_cand_kernel<<<(unsigned)crb3d_divup((int64_t)B * n_per_frame, 256), 256, 0, stream>>>(
        cls_preds, (int64_t)B * n_per_frame, n_class, n_per_frame, thresh, score, label, cand, cand_count, hist, tb, hshift);
    topk_sort_kernel<<<B, TOPK_THREADS, 0, stream>>>(cand, cand_count, n_per_frame, K, top_score, top_idx, counts, hist, tb, hshift);
    CRB3D_CHECK_LAUNCH();
    return CRB3D_OK;
}

// spec: HOST pointer to an AnchorSpec (layout above, mirrored by crb3d/ops.py). sel: (B,K) int64 anchor indices.
extern "C" int crb3d_anchor_decode_select(const float* box_preds, const float* dir_preds, const long long* sel, int B,
                                          int K, int64_t n_anchor_per_frame, const void* spec, float* out,
                                          cudaStream_t stream) {
    if (B < 0 || K < 0 || !spec || !out) return CRB3D_ERR_ARG;
    if (B * K == 0) return CRB3D_OK;
    AnchorSpec S = *reinterpret_cast<const AnchorSpec*>(spec);
    if (S.n_types <= 0 || S.n_types > CRB3D_MAX_ANCHOR_TYPES) return CRB3D_ERR_UNSUPPORTED;
    head_decode_kernel<<<(unsigned)crb3d_divup(B * K, 128), 128, 0, stream>>>(box_preds, dir_preds, sel, B, K,
                                                                             n_anchor_per_frame, S, out);
    CRB3D_CHECK_LAUNCH();
    return CRB3D_OK;
}

// out[b][k][0..width) = src[b][idx[b][k]][0..width) for k < valid[b] (valid nullable), else `fill`.
extern "C" int crb3d_gather_rows_f32(const float* src, const long long* idx, const int* valid, int B, int K,
                                     int64_t n_src, int width, float fill, float* out, cudaStream_t stream) {
    if (B < 0 || K < 0 || width <= 0 || !out) return CRB3D_ERR_ARG;
    if ((int64_t)B * K == 0) return CRB3D_OK;
    gather_rows_kernel<float><<<(unsigned)crb3d_divup((int64_t)B * K * width, 256), 256, 0, stream>>>(src, idx, valid, B, K,
                                                                                                     n_src, width, fill, out);
    CRB3D_CHECK_LAUNCH();
    return CRB3D_OK;
}

extern "C" int crb3d_gather_rows_i32(const int* src, const long long* idx, const int* valid, int B, int K, int64_t n_src,
                                     int width, int fill, int* out, cudaStream_t stream) {
    if (B < 0 || K < 0 || width <= 0 || !out) return CRB3D_ERR_ARG;
    if ((int64_t)B * K == 0) return CRB3D_OK;
    gather_rows_kernel<int><<<(unsigned)crb3d_divup((int64_t)B * K * width, 256), 256, 0, stream>>>(src, idx, valid, B, K, n_src,
                                                                                                   width, fill, out);
    CRB3D_CHECK_LAUNCH();
    return CRB3D_OK;
}
