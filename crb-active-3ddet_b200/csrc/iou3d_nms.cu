// Rotated BEV overlap / IoU and NMS (rotated + axis-aligned) with an ON-DEVICE greedy reduce.
//
// Replaces (reference, /root/reference):
//   pcdet/ops/iou3d_nms/src/iou3d_nms_kernel.cu:236-265  boxes_overlap_kernel / boxes_iou_bev_kernel
//   pcdet/ops/iou3d_nms/src/iou3d_nms_kernel.cu:267-372  nms_kernel / nms_normal_kernel (64x64 suppression bitmask)
//   pcdet/ops/iou3d_nms/src/iou3d_nms.cpp:90-188         nms_gpu / nms_normal_gpu host side: D2H of the mask +
//                                                        serial CPU greedy loop (iou3d_nms.cpp:121-132)
// The polygon arithmetic (edge crossings, corner containment with MARGIN=1e-2, angular sort about the centroid,
// shoelace sum) follows iou3d_nms_kernel.cu:36-234 operation for operation because the kept-index list is a
// bit-exact target. What is different:
//   * per-box data (corners, sin/cos) is computed once per tile in shared memory instead of once per pair;
//   * pairs whose centres are further apart than the two half-diagonals (+slack) are rejected before any polygon
//     work (their overlap is exactly 0 in the reference as well);
//   * only the upper-triangular tiles of the bitmask are produced (the reference's host loop never reads the rest);
//   * the greedy pass runs on the GPU in 64-box chunks and writes `keep` / `num_keep` in device memory, so the
//     2 MB mask never crosses PCIe and there is no host synchronisation.
#include "common.cuh"
#include "rbox.cuh"

namespace {

// ------------------------------------------------------------------ pairwise (N x M) overlap / IoU
template <bool IOU>
__global__ void __launch_bounds__(256) pairwise_kernel(int na, const float* __restrict__ a, int nb,
                                                       const float* __restrict__ b, float* __restrict__ out) {
    __shared__ RBox sa[16], sb[16];
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const int a0 = blockIdx.y * 16, b0 = blockIdx.x * 16;
    if (threadIdx.x < 16) {
        if (a0 + threadIdx.x < na) make_rbox(a + (size_t)(a0 + threadIdx.x) * 7, sa[threadIdx.x]);
    } else if (threadIdx.x < 32) {
        int t = threadIdx.x - 16;
        if (b0 + t < nb) make_rbox(b + (size_t)(b0 + t) * 7, sb[t]);
    }
    __syncthreads();
    const int ia = a0 + ty, ib = b0 + tx;
    if (ia >= na || ib >= nb) return;
    out[(size_t)ia * nb + ib] = IOU ? rbox_iou(sa[ty], sb[tx]) : rbox_overlap(sa[ty], sb[tx]);
}

// ------------------------------------------------------------------ NMS bitmask (upper-triangular tiles)
template <bool ROTATED>
__global__ void __launch_bounds__(64) nms_mask_kernel(int n, float thresh, const float* __restrict__ boxes,
                                                      unsigned long long* __restrict__ mask, int col_blocks,
                                                      const int* __restrict__ counts, int n_max) {
    // blockIdx.y = frame (batched NMS: frame b owns boxes[b*n_max ...] and mask[b*n_max*col_blocks ...], n = counts[b])
    if (counts) {
        n = min(counts[blockIdx.y], n_max);
        boxes += (size_t)blockIdx.y * n_max * 7;
        mask += (size_t)blockIdx.y * n_max * col_blocks;
    }
    // blockIdx.x enumerates tiles (r, c) with c >= r
    int t = blockIdx.x, r = 0;
    while (t >= col_blocks - r) { t -= col_blocks - r; ++r; }
    const int c = r + t;
    if (r * 64 >= n || c * 64 >= n) return;  // tile beyond this frame's boxes (uniform per CTA)
    const int row_size = min(n - r * 64, 64), col_size = min(n - c * 64, 64);
    __shared__ RBox cb[64];
    __shared__ float craw[64 * 7];
    const int tid = threadIdx.x;
    if (tid < col_size) {
        const float* src = boxes + (size_t)(c * 64 + tid) * 7;
        if (ROTATED) make_rbox(src, cb[tid]);
        else
            for (int q = 0; q < 7; ++q) craw[tid * 7 + q] = src[q];
    }
    __syncthreads();
    if (tid >= row_size) return;
    const int i = r * 64 + tid;
    unsigned long long bits = 0;
    const int start = (r == c) ? tid + 1 : 0;
    if (ROTATED) {
        RBox me;
        make_rbox(boxes + (size_t)i * 7, me);
        for (int j = start; j < col_size; ++j)
            if (rbox_iou(me, cb[j]) > thresh) bits |= 1ull << j;
    } else {
        float me[7];
        for (int q = 0; q < 7; ++q) me[q] = boxes[(size_t)i * 7 + q];
        for (int j = start; j < col_size; ++j)
            if (aabb_iou(me, craw + j * 7) > thresh) bits |= 1ull << j;
    }
    mask[(size_t)i * col_blocks + c] = bits;
}

// ------------------------------------------------------------------ greedy pass on the device
// Equivalent to iou3d_nms.cpp:116-132 (remv bitset, keep[] in ascending box index).
__global__ void __launch_bounds__(256) nms_reduce_kernel(int n, int col_blocks, const unsigned long long* __restrict__ mask,
                                                         int max_keep, long long* __restrict__ keep,
                                                         int* __restrict__ num_keep, const int* __restrict__ counts,
                                                         int n_max, int keep_stride) {
    extern __shared__ unsigned long long remv[];  // col_blocks
    if (counts) {
        n = min(counts[blockIdx.x], n_max);
        mask += (size_t)blockIdx.x * n_max * col_blocks;
        keep += (size_t)blockIdx.x * keep_stride;
        num_keep += blockIdx.x;
        col_blocks = (n_max + 63) / 64;
    }
    const int row_stride = col_blocks;
    col_blocks = (n + 63) / 64;
    __shared__ unsigned long long diag[64];
    __shared__ unsigned long long keepbits_s;
    __shared__ int kept_s;
    const int tid = threadIdx.x;
    for (int j = tid; j < col_blocks; j += blockDim.x) remv[j] = 0ull;
    if (tid == 0) kept_s = 0;
    __syncthreads();
    for (int c = 0; c < col_blocks; ++c) {
        const int base = c * 64, sz = min(64, n - base);
        if (tid < 64) diag[tid] = (tid < sz) ? mask[(size_t)(base + tid) * row_stride + c] : 0ull;
        __syncthreads();
        if (tid == 0) {
            unsigned long long cur = remv[c], kb = 0ull;
            int kept = kept_s;
            for (int b = 0; b < sz; ++b) {
                if (!((cur >> b) & 1ull)) {
                    if (max_keep > 0 && kept >= max_keep) break;
                    kb |= 1ull << b;
                    cur |= diag[b];
                    ++kept;
                }
            }
            keepbits_s = kb;
        }
        __syncthreads();
        const unsigned long long kb = keepbits_s;
        const int kept0 = kept_s;
        if (tid < 64 && ((kb >> tid) & 1ull))
            keep[kept0 + __popcll(kb & ((1ull << tid) - 1ull))] = base + tid;
        for (int j = c + 1 + tid; j < col_blocks; j += blockDim.x) {
            unsigned long long acc = 0ull, bits = kb;
            while (bits) {
                int b = __ffsll((long long)bits) - 1;
                bits &= bits - 1;
                acc |= mask[(size_t)(base + b) * row_stride + j];
            }
            remv[j] |= acc;
        }
        __syncthreads();
        if (tid == 0) kept_s = kept0 + __popcll(kb);
        __syncthreads();
        if (max_keep > 0 && kept_s >= max_keep) break;
    }
    if (tid == 0) *num_keep = kept_s;
}

// ------------------------------------------------------------------ fused greedy NMS (no bitmask)
// Greedy NMS only ever needs IoU(kept box, candidate): a candidate survives iff no PREVIOUSLY KEPT box overlaps it
// by more than the threshold - identical to the reference's mask + host loop, but with <= n*max_keep pair tests
// instead of n^2/2 and no n x n/64 mask in memory. One CTA per frame walks the score-sorted candidates in chunks of 64:
//   A) chunk vs kept list (all threads, quick centre-distance reject first)
//   B) chunk vs chunk upper triangle -> 64 suppression words in shared memory
//   C) one thread resolves the 64 candidates in order and appends the survivors to the kept list.
// IoU argument order is (earlier box, later box) like nms_kernel's iou_bev(cur_box, block_boxes + i*7).
constexpr int GREEDY_MAX_KEEP = 512;
constexpr int GREEDY_THREADS = 512;

struct RawBox { float v[7]; };
__device__ __forceinline__ void load_box(const float* b, RBox& r) { make_rbox(b, r); }
__device__ __forceinline__ void load_box(const float* b, RawBox& r) {
#pragma unroll
    for (int q = 0; q < 7; ++q) r.v[q] = b[q];
}
__device__ __forceinline__ float pair_iou(const RBox& a, const RBox& b) { return rbox_iou(a, b); }
__device__ __forceinline__ float pair_iou(const RawBox& a, const RawBox& b) { return aabb_iou(a.v, b.v); }
// false only when the overlap is exactly 0 (same centre-distance early-out as rbox_overlap)
__device__ __forceinline__ bool pair_near(const RBox& a, const RBox& b) {
    const float dx = a.cx - b.cx, dy = a.cy - b.cy, reach = a.rad + b.rad + 0.1f;
    return !(dx * dx + dy * dy > reach * reach);
}
__device__ __forceinline__ bool pair_near(const RawBox&, const RawBox&) { return true; }
constexpr int GREEDY_QUEUE = 8192;  // 128 kept boxes x 64 candidates per filtering round

template <typename BOX>
__global__ void __launch_bounds__(GREEDY_THREADS) nms_greedy_kernel(int n_fixed, const int* __restrict__ counts, int n_max,
                                                                    float thresh, const float* __restrict__ boxes,
                                                                    int max_keep, long long* __restrict__ keep,
                                                                    int keep_stride, int* __restrict__ num_keep) {
    // dynamic smem: kept[GREEDY_MAX_KEEP] | cand[64] | queue[GREEDY_QUEUE] (u16: kept/first index << 6 | candidate)
    extern __shared__ __align__(16) unsigned char greedy_smem[];
    BOX* kept = reinterpret_cast<BOX*>(greedy_smem);
    BOX* cand = kept + GREEDY_MAX_KEEP;
    unsigned short* queue = reinterpret_cast<unsigned short*>(cand + 64);
    __shared__ int supp[64];
    __shared__ unsigned long long cmask[64];
    __shared__ int nk_s, qn_s;
    const int b = blockIdx.x, tid = threadIdx.x;
    int n = n_fixed;
    if (counts) {
        n = min(counts[b], n_max);
        boxes += (size_t)b * n_max * 7;
        keep += (size_t)b * keep_stride;
        num_keep += b;
    }
    if (tid == 0) nk_s = 0;
    __syncthreads();
    for (int c0 = 0; c0 < n; c0 += 64) {
        const int sz = min(64, n - c0);
        const int nk = nk_s;
        if (tid < 64) {
            supp[tid] = 0;
            cmask[tid] = 0ull;
            if (tid < sz) load_box(boxes + (size_t)(c0 + tid) * 7, cand[tid]);
        }
        if (tid == 0) qn_s = 0;
        __syncthreads();
        // A) kept (earlier) vs candidates (later), 128 kept boxes at a time. A1 filters with the cheap exact-zero
        //    centre-distance test into a queue, A2 runs the polygon clipping one pair per thread: the expensive,
        //    divergent work is compacted first, so a warp never idles 31 lanes behind one overlapping pair.
        for (int j0 = 0; j0 < nk; j0 += GREEDY_QUEUE / 64) {
            const int jn = min(GREEDY_QUEUE / 64, nk - j0);
            for (int p = tid; p < jn * 64; p += GREEDY_THREADS) {
                const int j = j0 + (p >> 6), i = p & 63;
                if (i < sz && !supp[i] && pair_near(kept[j], cand[i])) queue[atomicAdd(&qn_s, 1)] = (unsigned short)((p >> 6) << 6 | i);
            }
            __syncthreads();
            const int qn = qn_s;
            for (int t = tid; t < qn; t += GREEDY_THREADS) {
                const int j = j0 + (queue[t] >> 6), i = queue[t] & 63;
                if (!supp[i] && pair_iou(kept[j], cand[i]) > thresh) supp[i] = 1;
            }
            __syncthreads();
            if (tid == 0) qn_s = 0;
            __syncthreads();
        }
        // B) within the chunk: i earlier, j later (same filter-then-evaluate split)
        for (int p = tid; p < 64 * 64; p += GREEDY_THREADS) {
            const int i = p >> 6, j = p & 63;
            if (j > i && j < sz && !supp[i] && !supp[j] && pair_near(cand[i], cand[j])) queue[atomicAdd(&qn_s, 1)] = (unsigned short)(p);
        }
        __syncthreads();
        {
            const int qn = qn_s;
            for (int t = tid; t < qn; t += GREEDY_THREADS) {
                const int i = queue[t] >> 6, j = queue[t] & 63;
                if (pair_iou(cand[i], cand[j]) > thresh) atomicOr(&cmask[i], 1ull << j);
            }
        }
        __syncthreads();
        // C) serial resolve of the chunk
        if (tid == 0) {
            unsigned long long removed = 0ull;
            int k = nk;
            for (int i = 0; i < sz && k < max_keep; ++i) {
                if (supp[i] || ((removed >> i) & 1ull)) continue;
                kept[k] = cand[i];
                keep[k] = c0 + i;
                ++k;
                removed |= cmask[i];
            }
            nk_s = k;
        }
        __syncthreads();
        if (nk_s >= max_keep) break;
    }
    if (tid == 0) *num_keep = nk_s;
}

template <typename BOX>
int launch_greedy_t(int grid, int n_fixed, const int* counts, int n_max, float thresh, const float* boxes, int max_keep,
                    long long* keep, int keep_stride, int* num_keep, cudaStream_t stream) {
    const size_t smem = sizeof(BOX) * (GREEDY_MAX_KEEP + 64) + sizeof(unsigned short) * GREEDY_QUEUE;
    auto kern = nms_greedy_kernel<BOX>;
    if (smem > 48 * 1024) CRB3D_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid, GREEDY_THREADS, smem, stream>>>(n_fixed, counts, n_max, thresh, boxes, max_keep, keep, keep_stride, num_keep);
    CRB3D_CHECK_LAUNCH();
    return CRB3D_OK;
}

int launch_greedy(int rotated, int grid, int n_fixed, const int* counts, int n_max, float thresh, const float* boxes,
                  int max_keep, long long* keep, int keep_stride, int* num_keep, cudaStream_t stream) {
    return rotated ? launch_greedy_t<RBox>(grid, n_fixed, counts, n_max, thresh, boxes, max_keep, keep, keep_stride, num_keep, stream)
                   : launch_greedy_t<RawBox>(grid, n_fixed, counts, n_max, thresh, boxes, max_keep, keep, keep_stride, num_keep, stream);
}

}  // namespace

extern "C" int crb3d_boxes_overlap_bev(const float* boxes_a, int na, const float* boxes_b, int nb, float* out,
                                       cudaStream_t stream) {
    if (na < 0 || nb < 0 || !out) return CRB3D_ERR_ARG;
    if (na == 0 || nb == 0) return CRB3D_OK;
    dim3 grid((unsigned)crb3d_divup(nb, 16), (unsigned)crb3d_divup(na, 16));
    pairwise_kernel<false><<<grid, 256, 0, stream>>>(na, boxes_a, nb, boxes_b, out);
    CRB3D_CHECK_LAUNCH();
    return CRB3D_OK;
}

extern "C" int crb3d_boxes_iou_bev(const float* boxes_a, int na, const float* boxes_b, int nb, float* out,
                                   cudaStream_t stream) {
    if (na < 0 || nb < 0 || !out) return CRB3D_ERR_ARG;
    if (na == 0 || nb == 0) return CRB3D_OK;
    dim3 grid((unsigned)crb3d_divup(nb, 16), (unsigned)crb3d_divup(na, 16));
    pairwise_kernel<true><<<grid, 256, 0, stream>>>(na, boxes_a, nb, boxes_b, out);
    CRB3D_CHECK_LAUNCH();
    return CRB3D_OK;
}

extern "C" int crb3d_nms_workspace_bytes(int n, size_t* bytes) {
    if (!bytes || n < 0) return CRB3D_ERR_ARG;
    size_t cb = (size_t)crb3d_divup(n > 0 ? n : 1, 64);
    *bytes = crb3d_align(sizeof(unsigned long long) * (size_t)(n > 0 ? n : 1) * cb);
    return CRB3D_OK;
}

// boxes: (n,7) sorted by descending score. keep: device int64[n] (first *num_keep entries valid, ascending).
extern "C" int crb3d_nms(const float* boxes, int n, float thresh, int rotated, int max_keep, long long* keep,
                         int* num_keep, void* ws, size_t ws_bytes, cudaStream_t stream) {
    if (n < 0 || !keep || !num_keep) return CRB3D_ERR_ARG;
    if (n == 0) { CRB3D_CUDA(cudaMemsetAsync(num_keep, 0, sizeof(int), stream)); return CRB3D_OK; }
    if (max_keep > 0 && max_keep <= GREEDY_MAX_KEEP) {  // bounded keep list: fused greedy kernel, no mask
        return launch_greedy(rotated, 1, n, nullptr, n, thresh, boxes, max_keep, keep, 0, num_keep, stream);
    }
    const int cb = (int)crb3d_divup(n, 64);
    WsCursor c(ws, ws_bytes);
    unsigned long long* mask = c.take<unsigned long long>((size_t)n * cb);
    if (!c.ok) return CRB3D_ERR_WORKSPACE;
    const unsigned tiles = (unsigned)((int64_t)cb * (cb + 1) / 2);
    if (rotated) nms_mask_kernel<true><<<tiles, 64, 0, stream>>>(n, thresh, boxes, mask, cb, nullptr, 0);
    else nms_mask_kernel<false><<<tiles, 64, 0, stream>>>(n, thresh, boxes, mask, cb, nullptr, 0);
    size_t smem = sizeof(unsigned long long) * cb;
    if (smem > 48 * 1024) {
        if (smem > 200 * 1024) return CRB3D_ERR_UNSUPPORTED;
        CRB3D_CUDA(cudaFuncSetAttribute(nms_reduce_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    }
    nms_reduce_kernel<<<1, 256, smem, stream>>>(n, cb, mask, max_keep, keep, num_keep, nullptr, 0, 0);
    CRB3D_CHECK_LAUNCH();
    return CRB3D_OK;
}

// Raw suppression bitmask (full upper triangle), for parity checks against the reference nms_kernel.
extern "C" int crb3d_nms_mask(const float* boxes, int n, float thresh, int rotated, unsigned long long* mask,
                              cudaStream_t stream) {
    if (n <= 0 || !mask) return CRB3D_ERR_ARG;
    const int cb = (int)crb3d_divup(n, 64);
    const unsigned tiles = (unsigned)((int64_t)cb * (cb + 1) / 2);
    if (rotated) nms_mask_kernel<true><<<tiles, 64, 0, stream>>>(n, thresh, boxes, mask, cb, nullptr, 0);
    else nms_mask_kernel<false><<<tiles, 64, 0, stream>>>(n, thresh, boxes, mask, cb, nullptr, 0);
    CRB3D_CHECK_LAUNCH();
    return CRB3D_OK;
}

// Batched NMS for the scoring path: frame b owns boxes[b][0..counts[b]) of a padded (B, n_max, 7) tensor (each frame
// sorted by descending score). keep: (B, keep_stride) int64, num_keep: (B). One mask launch + one reduce launch for
// the whole batch, no host synchronisation. ws: B * n_max * ceil(n_max/64) * 8 bytes.
extern "C" int crb3d_nms_batched_workspace_bytes(int B, int n_max, size_t* bytes) {
    if (!bytes || B < 0 || n_max < 0) return CRB3D_ERR_ARG;
    size_t cb = (size_t)crb3d_divup(n_max > 0 ? n_max : 1, 64);
    *bytes = crb3d_align(sizeof(unsigned long long) * (size_t)(B > 0 ? B : 1) * (size_t)(n_max > 0 ? n_max : 1) * cb);
    return CRB3D_OK;
}

extern "C" int crb3d_nms_batched(const float* boxes, const int* counts, int B, int n_max, float thresh, int rotated,
                                 int max_keep, long long* keep, int keep_stride, int* num_keep, void* ws,
                                 size_t ws_bytes, cudaStream_t stream) {
    if (B < 0 || n_max < 0 || !counts || !keep || !num_keep || keep_stride <= 0) return CRB3D_ERR_ARG;
    if (max_keep <= 0 && keep_stride < n_max) return CRB3D_ERR_ARG;
    if (max_keep > 0 && keep_stride < max_keep) return CRB3D_ERR_ARG;
    if (B == 0) return CRB3D_OK;
    if (n_max == 0) { CRB3D_CUDA(cudaMemsetAsync(num_keep, 0, sizeof(int) * B, stream)); return CRB3D_OK; }
    if (max_keep > 0 && max_keep <= GREEDY_MAX_KEEP) {  // the scoring path (NMS_POST_MAXSIZE = 500): one CTA per frame
        return launch_greedy(rotated, B, 0, counts, n_max, thresh, boxes, max_keep, keep, keep_stride, num_keep, stream);
    }
    const int cb = (int)crb3d_divup(n_max, 64);
    WsCursor c(ws, ws_bytes);
    unsigned long long* mask = c.take<unsigned long long>((size_t)B * n_max * cb);
    if (!c.ok) return CRB3D_ERR_WORKSPACE;
    const unsigned tiles = (unsigned)((int64_t)cb * (cb + 1) / 2);
    if (rotated) nms_mask_kernel<true><<<dim3(tiles, B), 64, 0, stream>>>(0, thresh, boxes, mask, cb, counts, n_max);
    else nms_mask_kernel<false><<<dim3(tiles, B), 64, 0, stream>>>(0, thresh, boxes, mask, cb, counts, n_max);
    size_t smem = sizeof(unsigned long long) * cb;
    if (smem > 48 * 1024) {
        if (smem > 200 * 1024) return CRB3D_ERR_UNSUPPORTED;
        CRB3D_CUDA(cudaFuncSetAttribute(nms_reduce_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    }
    nms_reduce_kernel<<<B, 256, smem, stream>>>(0, cb, mask, max_keep, keep, num_keep, counts, n_max, keep_stride);
    CRB3D_CHECK_LAUNCH();
    return CRB3D_OK;
}
