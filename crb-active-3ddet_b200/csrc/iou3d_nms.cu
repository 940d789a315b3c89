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
