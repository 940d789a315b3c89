// Rotated BEV overlap / IoU and NMS (rotated + axis-aligned) with an ON-DEVICE greedy pass.
//
// Replaces (reference, /root/reference):
//   pcdet/ops/iou3d_nms/src/iou3d_nms_kernel.cu:236-265  boxes_overlap_kernel / boxes_iou_bev_kernel
//   pcdet/ops/iou3d_nms/src/iou3d_nms_kernel.cu:267-372  nms_kernel / nms_normal_kernel (64x64 suppression bitmask)
//   pcdet/ops/iou3d_nms/src/iou3d_nms.cpp:90-188         nms_gpu / nms_normal_gpu host side: D2H of the mask +
//                                                        serial CPU greedy loop (iou3d_nms.cpp:121-132)
// The polygon arithmetic (rbox.cuh) follows iou3d_nms_kernel.cu:36-234 operation for operation because the kept-index
// list is a bit-exact target. What is different:
//   * per-box data (corners, sin/cos) is computed once per tile in shared memory instead of once per pair;
//   * FILTER THEN EVALUATE: every tile first runs the cheap exact-zero centre-distance test on its 64x64 pairs and
//     compacts the survivors into a shared-memory queue; only then is the divergent polygon clipping run, one queued
//     pair per thread, so a warp never idles 31 lanes behind one overlapping pair (the reference evaluates 64 pairs
//     serially per thread);
//   * only the upper-triangular tiles of the bitmask are produced (the reference's host loop never reads the rest);
//   * the greedy pass runs on the GPU: per 64-box chunk it PULLS the suppression word of every already-kept box
//     (<= keep-count loads spread over the CTA + one OR-reduction) and resolves the chunk against its diagonal tile,
//     writing `keep` / `num_keep` in device memory - the mask never crosses PCIe, there is no host synchronisation,
//     and a bounded keep list (NMS_POST_MAXSIZE) stops the pass early;
//   * a batched entry point runs all frames of a batch in one mask launch + one reduce launch.
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

// false only when the overlap is exactly 0 (same centre-distance early-out as rbox_overlap)
__device__ __forceinline__ bool rbox_near(const RBox& a, const RBox& b) {
    const float dx = a.cx - b.cx, dy = a.cy - b.cy, reach = a.rad + b.rad + 0.1f;
    return !(dx * dx + dy * dy > reach * reach);
}

// ------------------------------------------------------------------ NMS bitmask (upper-triangular 64x64 tiles)
constexpr int MASK_THREADS = 256;

template <bool ROTATED>
__global__ void __launch_bounds__(MASK_THREADS) nms_mask_kernel(int n, float thresh, const float* __restrict__ boxes,
                                                                unsigned long long* __restrict__ mask, int col_blocks,
                                                                const int* __restrict__ counts, int n_max) {
    // blockIdx.y = frame (batched: frame b owns boxes[b*n_max ...] and mask[b*n_max*col_blocks ...], n = counts[b])
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
    __shared__ RBox rb[64], cb[64];
    __shared__ float rraw[64 * 7], craw[64 * 7];
    __shared__ unsigned short queue[4096];
    __shared__ unsigned long long bits[64];
    __shared__ int qn;
    const int tid = threadIdx.x, lane = tid & 31;
    if (tid < 64) {
        bits[tid] = 0ull;
        if (tid < row_size) {
            const float* src = boxes + (size_t)(r * 64 + tid) * 7;
            if (ROTATED) make_rbox(src, rb[tid]);
            else
                for (int q = 0; q < 7; ++q) rraw[tid * 7 + q] = src[q];
        }
    } else if (tid < 128) {
        const int u = tid - 64;
        if (u < col_size) {
            const float* src = boxes + (size_t)(c * 64 + u) * 7;
            if (ROTATED) make_rbox(src, cb[u]);
            else
                for (int q = 0; q < 7; ++q) craw[u * 7 + q] = src[q];
        }
    }
    if (tid == 0) qn = 0;
    __syncthreads();
    // filter: row i (earlier box) vs column j (later box); diagonal tile keeps j > i only
    for (int p = tid; p < 64 * 64; p += MASK_THREADS) {
        const int i = p >> 6, j = p & 63;
        bool near = i < row_size && j < col_size && (r != c || j > i);
        if (ROTATED) near = near && rbox_near(rb[i], cb[j]);
        const unsigned int m = __ballot_sync(0xffffffffu, near);
        if (m) {
            int base = 0;
            if (lane == 0) base = atomicAdd(&qn, __popc(m));
            base = __shfl_sync(0xffffffffu, base, 0);
            if (near) queue[base + __popc(m & ((1u << lane) - 1u))] = (unsigned short)p;
        }
    }
    __syncthreads();
    const int nq = qn;
    for (int q = tid; q < nq; q += MASK_THREADS) {
        const int i = queue[q] >> 6, j = queue[q] & 63;
        const float v = ROTATED ? rbox_iou(rb[i], cb[j]) : aabb_iou(rraw + i * 7, craw + j * 7);
        if (v > thresh) atomicOr(&bits[i], 1ull << j);
    }
    __syncthreads();
    if (tid < row_size) mask[(size_t)(r * 64 + tid) * col_blocks + c] = bits[tid];
}

// ------------------------------------------------------------------ greedy pass on the device
// Equivalent to iou3d_nms.cpp:116-132 (remv bitset, keep[] in ascending box index). Box i of chunk c is kept iff no
// kept box j < i suppresses it: kept boxes of EARLIER chunks are pulled (mask[j][c], one word each, all in flight at
// once), kept boxes of the same chunk come from the diagonal tile.
constexpr int REDUCE_THREADS = 256;

__global__ void __launch_bounds__(REDUCE_THREADS) nms_reduce_kernel(int n, int col_blocks, const unsigned long long* __restrict__ mask,
                                                                    int max_keep, long long* __restrict__ keep,
                                                                    int* __restrict__ num_keep, const int* __restrict__ counts,
                                                                    int n_max, int keep_stride) {
    extern __shared__ int kept_list[];  // up to n entries
    if (counts) {
        n = min(counts[blockIdx.x], n_max);
        mask += (size_t)blockIdx.x * n_max * col_blocks;
        keep += (size_t)blockIdx.x * keep_stride;
        num_keep += blockIdx.x;
    }
    const int row_stride = col_blocks;
    const int chunks = (n + 63) / 64;
    __shared__ unsigned long long diag[64];
    __shared__ unsigned long long wor[REDUCE_THREADS / 32];
    __shared__ int kept_s;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) kept_s = 0;
    __syncthreads();
    for (int c = 0; c < chunks; ++c) {
        const int base = c * 64, sz = min(64, n - base);
        const int nk = kept_s;
        unsigned long long part = 0ull;
        for (int t = tid; t < nk; t += REDUCE_THREADS) part |= __ldg(&mask[(size_t)kept_list[t] * row_stride + c]);
        if (tid < 64) diag[tid] = (tid < sz) ? __ldg(&mask[(size_t)(base + tid) * row_stride + c]) : 0ull;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) part |= __shfl_xor_sync(0xffffffffu, part, o);
        if (lane == 0) wor[warp] = part;
        __syncthreads();
        if (tid == 0) {
            unsigned long long cur = 0ull;
            for (int w = 0; w < REDUCE_THREADS / 32; ++w) cur |= wor[w];
            // kept boxes that belong to THIS chunk were appended below in earlier iterations only, so `cur` holds the
            // earlier chunks' verdict; the diagonal tile adds the within-chunk suppressions in index order
            int k = nk;
            for (int b = 0; b < sz; ++b) {
                if ((cur >> b) & 1ull) continue;
                if (max_keep > 0 && k >= max_keep) break;
                kept_list[k] = base + b;
                keep[k] = base + b;
                ++k;
                cur |= diag[b];
            }
            kept_s = k;
        }
        __syncthreads();
        if (max_keep > 0 && kept_s >= max_keep) break;
    }
    if (tid == 0) *num_keep = kept_s;
}

int launch_nms(const float* boxes, const int* counts, int B, int n, float thresh, int rotated, int max_keep, long long* keep,
               int keep_stride, int* num_keep, unsigned long long* mask, cudaStream_t stream) {
    const int cb = (int)crb3d_divup(n, 64);
    const unsigned tiles = (unsigned)((int64_t)cb * (cb + 1) / 2);
    if (rotated) nms_mask_kernel<true><<<dim3(tiles, B), MASK_THREADS, 0, stream>>>(n, thresh, boxes, mask, cb, counts, n);
    else nms_mask_kernel<false><<<dim3(tiles, B), MASK_THREADS, 0, stream>>>(n, thresh, boxes, mask, cb, counts, n);
    const size_t smem = sizeof(int) * (size_t)(max_keep > 0 ? (max_keep < n ? max_keep : n) : n);
    if (smem > 200 * 1024) return CRB3D_ERR_UNSUPPORTED;
    if (smem > 40 * 1024) CRB3D_CUDA(cudaFuncSetAttribute(nms_reduce_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    nms_reduce_kernel<<<B, REDUCE_THREADS, smem, stream>>>(n, cb, mask, max_keep, keep, num_keep, counts, n, keep_stride);
    CRB3D_CHECK_LAUNCH();
    return CRB3D_OK;
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
    WsCursor c(ws, ws_bytes);
    unsigned long long* mask = c.take<unsigned long long>((size_t)n * crb3d_divup(n, 64));
    if (!c.ok) return CRB3D_ERR_WORKSPACE;
    return launch_nms(boxes, nullptr, 1, n, thresh, rotated, max_keep, keep, 0, num_keep, mask, stream);
}

// Raw suppression bitmask (upper-triangular tiles; the rest is left untouched), for parity checks against the
// reference nms_kernel.
extern "C" int crb3d_nms_mask(const float* boxes, int n, float thresh, int rotated, unsigned long long* mask,
                              cudaStream_t stream) {
    if (n <= 0 || !mask) return CRB3D_ERR_ARG;
    const int cb = (int)crb3d_divup(n, 64);
    const unsigned tiles = (unsigned)((int64_t)cb * (cb + 1) / 2);
    if (rotated) nms_mask_kernel<true><<<tiles, MASK_THREADS, 0, stream>>>(n, thresh, boxes, mask, cb, nullptr, 0);
    else nms_mask_kernel<false><<<tiles, MASK_THREADS, 0, stream>>>(n, thresh, boxes, mask, cb, nullptr, 0);
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
    WsCursor c(ws, ws_bytes);
    unsigned long long* mask = c.take<unsigned long long>((size_t)B * n_max * crb3d_divup(n_max, 64));
    if (!c.ok) return CRB3D_ERR_WORKSPACE;
    return launch_nms(boxes, counts, B, n_max, thresh, rotated, max_keep, keep, keep_stride, num_keep, mask, stream);
}
