// Rotated BEV overlap / IoU and NMS (rotated + axis-aligned) with an ON-DEVICE greedy pass.
//
// Replaces (reference, /root/reference):
//   pcdet/ops/iou3d_nms/src/iou3d_nms_kernel.cu:236-265  boxes_overlap_kernel / boxes_iou_bev_kernel
//   pcdet/ops/iou3d_nms/src/iou3d_nms_kernel.cu:267-372  nms_kernel / nms_normal_kernel (64x64 suppression bitmask)
//   pcdet/ops/iou3d_nms/src/iou3d_nms.cpp:90-188         nms_gpu / nms_normal_gpu host side: D2H of the mask +
//                                                        serial CPU greedy loop (iou3d_nms.cpp:121-132)
// The polygon arithmetic (rbox.cuh) follows iou3d_nms_kernel.cu:36-234 operation for operation because the kept-index
// list is a bit-exact target. What is different:
//   * per-box data (corners, sin/cos) is computed once per box by a prep kernel instead of once per pair;
//   * FILTER THEN EVALUATE: a CTA owns 64 rows x 512 columns; warps sweep the cheap exact-zero centre-distance test
//     over 16-byte filter records in shared memory and compact the survivors into a queue; only then is the divergent
//     polygon clipping run, one queued pair per thread, so a warp never idles 31 lanes behind one overlapping pair
//     (the reference evaluates 64 pairs serially per thread);
//   * only the blocks on/above the diagonal of the bitmask are produced (the reference's host loop never reads the
//     rest), stored column-block major so the greedy pass streams them with contiguous copies;
//   * the greedy pass runs on the GPU, one CTA per frame: column block c+1 is prefetched (cp.async, double buffer)
//     while chunk c is resolved from shared memory, writing `keep` / `num_keep` in device memory - the mask never
//     crosses PCIe, there is no host synchronisation, and a bounded keep list (NMS_POST_MAXSIZE) stops the pass early;
//   * a batched entry point runs all frames of a batch in one prep + one mask + one reduce launch.
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
__device__ __forceinline__ bool circ_near(const float4& a, const float4& b) {
    const float dx = a.x - b.x, dy = a.y - b.y, reach = a.z + b.z + 0.1f;
    return !(dx * dx + dy * dy > reach * reach);
}
// axis-aligned: (left, right, top, bottom) with the expressions of aabb_iou; false only when inter == 0 exactly
__device__ __forceinline__ bool aabb_near(const float4& a, const float4& b) {
    return fminf(a.y, b.y) - fmaxf(a.x, b.x) > 0.f && fminf(a.w, b.w) - fmaxf(a.z, b.z) > 0.f;
}

template <bool ROTATED>
__device__ __forceinline__ float pair_iou(const float* __restrict__ boxes, const RBox* __restrict__ rb, int j, int i) {
    if (ROTATED) {
        RBox a, b;
        const float4* pa = reinterpret_cast<const float4*>(rb + j);
        const float4* pb = reinterpret_cast<const float4*>(rb + i);
        float4* da = reinterpret_cast<float4*>(&a);
        float4* db = reinterpret_cast<float4*>(&b);
#pragma unroll
        for (int u = 0; u < 4; ++u) { da[u] = __ldg(pa + u); db[u] = __ldg(pb + u); }
        return rbox_iou(a, b);
    }
    return aabb_iou(boxes + (size_t)j * 7, boxes + (size_t)i * 7);
}

// ------------------------------------------------------------------ NMS step 1: per-box precompute
// One thread per box: the rotated-box record (corners, sin/cos: 64 B) once per box instead of once per tile, and a
// 16-byte filter record (rotated: centre + half diagonal; axis-aligned: the four edges).
template <bool ROTATED>
__global__ void __launch_bounds__(256) nms_prep_kernel(int n, const float* __restrict__ boxes, const int* __restrict__ counts,
                                                       int n_max, RBox* __restrict__ rb, float4* __restrict__ circ) {
    const int f = blockIdx.y;
    if (counts) n = min(counts[f], n_max);
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float* src = boxes + ((size_t)f * n_max + i) * 7;
    const size_t o = (size_t)f * n_max + i;
    if (ROTATED) {
        RBox r;
        make_rbox(src, r);
        const float4* p = reinterpret_cast<const float4*>(&r);
        float4* d = reinterpret_cast<float4*>(rb + o);
        d[0] = p[0]; d[1] = p[1]; d[2] = p[2]; d[3] = p[3];
        circ[o] = make_float4(r.cx, r.cy, r.rad, 0.f);
    } else {
        circ[o] = make_float4(src[0] - src[3] / 2, src[0] + src[3] / 2, src[1] - src[4] / 2, src[1] + src[4] / 2);
    }
}

// ------------------------------------------------------------------ NMS step 2: suppression bitmask
// CTA = 64 rows (earlier boxes) x up to MASK_CW*64 columns (later boxes). FILTER THEN EVALUATE: a warp sweeps one row's
// columns with the 16-byte filter records (shared memory) and compacts the surviving pairs into a queue; then the CTA
// runs the divergent polygon clipping one queued pair per thread. Words are stored column-block major
// (maskT[c][row]) so that the greedy pass can stream a whole column block with contiguous 16-byte copies.
constexpr int MASK_THREADS = 256;
constexpr int MASK_CW = 8;

template <bool ROTATED>
__global__ void __launch_bounds__(MASK_THREADS) nms_mask_kernel(int n, float thresh, const float* __restrict__ boxes,
                                                                const RBox* __restrict__ rb, const float4* __restrict__ circ,
                                                                unsigned long long* __restrict__ maskT, int rows_pad,
                                                                const int* __restrict__ counts, int n_max, int prefix,
                                                                const int* __restrict__ idx, const int* __restrict__ m_dev,
                                                                unsigned int* __restrict__ gq, int* __restrict__ gq_count, int gq_cap) {
    // gq != null: surviving pairs are appended to a GLOBAL queue (frame:4 | row:14 | column:14 bits) that nms_eval_kernel
    // drains with every SM of the chip - a tile whose pairs pile up (the top-scored boxes of a frame sit on a few objects)
    // would otherwise serialise thousands of ~15 us polygon clippings on its eight warps. A reservation that does not fit
    // is evaluated here instead.
    // level selection: prefix > 0 -> only the first `prefix` boxes; idx/m_dev -> the m_dev[f] boxes listed in idx
    const int f = blockIdx.z;
    if (counts) n = min(counts[f], n_max);
    if (prefix > 0) n = min(n, prefix);
    if (m_dev) { n = m_dev[f]; idx += (size_t)f * n_max; }
    const int cbn = (n + 63) >> 6;
    const int r = blockIdx.y, c0 = max((int)blockIdx.x * MASK_CW, r), c1 = min(((int)blockIdx.x + 1) * MASK_CW, cbn);
    if (r >= cbn || c1 <= c0) return;  // uniform per CTA
    boxes += (size_t)f * n_max * 7;
    rb += (size_t)f * n_max;
    circ += (size_t)f * n_max;
    maskT += (size_t)f * rows_pad * (rows_pad >> 6);
    __shared__ RBox rrow[64];
    __shared__ float4 rcirc[64];
    __shared__ float4 ccirc[MASK_CW * 64];
    __shared__ unsigned long long bits[MASK_CW][64];
    __shared__ unsigned int queue[(MASK_THREADS / 32) * MASK_CW * 64];
    __shared__ int qn, gbase_s;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ncols = min(n - c0 * 64, (c1 - c0) * 64);
    const bool all_near = thresh < 0.f;  // then even a zero IoU suppresses: nothing may be filtered out
    if (tid < 64) {
        const int i = r * 64 + tid;
        if (i < n) {
            const int bi = idx ? idx[i] : i;
            rcirc[tid] = circ[bi];
            if (ROTATED) {
                const float4* p = reinterpret_cast<const float4*>(rb + bi);
                float4* d = reinterpret_cast<float4*>(&rrow[tid]);
                d[0] = p[0]; d[1] = p[1]; d[2] = p[2]; d[3] = p[3];
            }
        }
    }
    for (int t = tid; t < ncols; t += MASK_THREADS) ccirc[t] = circ[idx ? idx[c0 * 64 + t] : c0 * 64 + t];
    for (int t = tid; t < MASK_CW * 64; t += MASK_THREADS) (&bits[0][0])[t] = 0ull;
    __syncthreads();
    for (int i0 = 0; i0 < 64; i0 += MASK_THREADS / 32) {
        if (tid == 0) qn = 0;
        __syncthreads();
        const int il = i0 + warp, ig = r * 64 + il;  // one row per warp
        if (ig < n) {
            const float4 me = rcirc[il];
            for (int jj = lane; jj < ((ncols + 31) & ~31); jj += 32) {
                const int jg = c0 * 64 + jj;
                bool near = jj < ncols && jg > ig;
                if (near && !all_near) near = ROTATED ? circ_near(me, ccirc[jj]) : aabb_near(me, ccirc[jj]);
                const unsigned int m = __ballot_sync(0xffffffffu, near);
                if (m) {
                    int base = 0;
                    if (lane == 0) base = atomicAdd(&qn, __popc(m));
                    base = __shfl_sync(0xffffffffu, base, 0);
                    if (near) queue[base + __popc(m & ((1u << lane) - 1u))] = ((unsigned int)il << 16) | (unsigned int)jj;
                }
            }
        }
        __syncthreads();
        const int nq = qn;
        if (gq) {
            if (tid == 0) {
                gbase_s = -1;
                if (nq > 0) {
                    const int b = atomicAdd(gq_count, nq);
                    if (b + nq <= gq_cap) gbase_s = b;
                    else if (b < gq_cap) gbase_s = -2 - b;   // straddles the end: pad the tail with skip markers
                }
            }
            __syncthreads();
            const int gb = gbase_s;
            if (gb >= 0) {
                for (int q = tid; q < nq; q += MASK_THREADS)
                    gq[gb + q] = ((unsigned int)f << 28) | ((unsigned int)(r * 64 + (queue[q] >> 16)) << 14) |
                                 (unsigned int)(c0 * 64 + (queue[q] & 0xffff));
                __syncthreads();
                continue;
            }
            if (gb <= -2) for (int t = -2 - gb + tid; t < gq_cap; t += MASK_THREADS) gq[t] = 0xFFFFFFFFu;
        }
        for (int q = tid; q < nq; q += MASK_THREADS) {
            const int il2 = queue[q] >> 16, jj = queue[q] & 0xffff;
            const int jg = idx ? idx[c0 * 64 + jj] : c0 * 64 + jj;
            float v;
            if (ROTATED) {
                RBox cbx;
                const float4* p = reinterpret_cast<const float4*>(rb + jg);
                float4* d = reinterpret_cast<float4*>(&cbx);
                d[0] = __ldg(p); d[1] = __ldg(p + 1); d[2] = __ldg(p + 2); d[3] = __ldg(p + 3);
                v = rbox_iou(rrow[il2], cbx);
            } else {
                v = aabb_iou(boxes + (size_t)(idx ? idx[r * 64 + il2] : r * 64 + il2) * 7, boxes + (size_t)jg * 7);
            }
            if (v > thresh) atomicOr(&bits[jj >> 6][il2], 1ull << (jj & 63));
        }
        __syncthreads();
    }
    for (int t = tid; t < (c1 - c0) * 64; t += MASK_THREADS)
        maskT[(size_t)(c0 + (t >> 6)) * rows_pad + r * 64 + (t & 63)] = bits[t >> 6][t & 63];
}

// drains the global pair queue of nms_mask_kernel: one pair per thread, the whole chip
template <bool ROTATED>
__global__ void __launch_bounds__(256) nms_eval_kernel(const unsigned int* __restrict__ gq, const int* __restrict__ gq_count,
                                                       int gq_cap, float thresh, const float* __restrict__ boxes,
                                                       const RBox* __restrict__ rb, const int* __restrict__ idx, int n_max,
                                                       unsigned long long* __restrict__ maskT, int rows_pad) {
    const int total = min(*gq_count, gq_cap);
    for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < total; q += gridDim.x * blockDim.x) {
        const unsigned int e = gq[q];
        if (e == 0xFFFFFFFFu) continue;
        const int f = e >> 28, i = (e >> 14) & 0x3FFF, j = e & 0x3FFF;
        const int bi = idx ? idx[(size_t)f * n_max + i] : i, bj = idx ? idx[(size_t)f * n_max + j] : j;
        const float v = pair_iou<ROTATED>(boxes + (size_t)f * n_max * 7, rb + (size_t)f * n_max, bi, bj);
        if (v > thresh)
            atomicOr(&maskT[(size_t)f * rows_pad * (rows_pad >> 6) + (size_t)(j >> 6) * rows_pad + i], 1ull << (j & 63));
    }
}

// maskT[c][row] -> the reference's row-major mask[row][c] (parity checks only); cells below the diagonal are not written
__global__ void nms_mask_transpose_kernel(int n, int cb, int rows_pad, const unsigned long long* __restrict__ maskT,
                                          unsigned long long* __restrict__ mask) {
    const int row = blockIdx.x * blockDim.x + threadIdx.x, c = blockIdx.y;
    if (row < n && c >= (row >> 6)) mask[(size_t)row * cb + c] = maskT[(size_t)c * rows_pad + row];
}

// ------------------------------------------------------------------ NMS step 3: greedy pass on the device
// Equivalent to iou3d_nms.cpp:116-132 (remv bitset, keep[] in ascending box index). Box i of chunk c is kept iff no
// kept box j < i suppresses it. One CTA per frame walks the chunks; column block c+1 of maskT (the words of ALL
// earlier rows, <= 32 KB) is prefetched with cp.async into the other half of a double buffer while chunk c is being
// resolved, so no global-memory latency sits on the serial chain: the verdict of the kept boxes of earlier chunks is
// an OR over <= keep-count shared-memory words, and the within-chunk resolve visits only the boxes that survive
// (find-first-set over the not-yet-suppressed bits) instead of all 64.
constexpr int REDUCE_THREADS = 256;

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc));
}

// STREAM = false (very large n: the double buffer does not fit): words are pulled from global memory instead.
template <bool STREAM>
__global__ void __launch_bounds__(REDUCE_THREADS) nms_reduce_kernel(int n, int rows_pad, const unsigned long long* __restrict__ maskT,
                                                                    int max_keep, long long* __restrict__ keep,
                                                                    int* __restrict__ num_keep, const int* __restrict__ counts,
                                                                    int n_max, int keep_stride, int prefix,
                                                                    const int* __restrict__ idx, const int* __restrict__ m_dev) {
    // idx/m_dev: second level - positions map to box indices through idx and the kept boxes are APPENDED after the
    // *num_keep boxes the first level kept (their indices are all smaller, so the list stays ascending)
    extern __shared__ __align__(16) unsigned char dyn[];
    // layout: two column-block buffers of rows_pad words each, then the kept list (up to n ints)
    unsigned long long* buf0 = reinterpret_cast<unsigned long long*>(dyn);
    unsigned long long* buf1 = buf0 + (STREAM ? rows_pad : 0);
    int* kept_list = reinterpret_cast<int*>(buf1 + (STREAM ? rows_pad : 0));
    const int f = blockIdx.x;
    if (counts) {
        n = min(counts[f], n_max);
        maskT += (size_t)f * rows_pad * (rows_pad >> 6);
        keep += (size_t)f * keep_stride;
        num_keep += f;
    }
    if (prefix > 0) n = min(n, prefix);
    int kbase = 0;
    if (m_dev) { n = m_dev[f]; idx += (size_t)f * n_max; kbase = *num_keep; }
    const int chunks = (n + 63) / 64;
    __shared__ unsigned long long wor[REDUCE_THREADS / 32];
    __shared__ int kept_s;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) kept_s = 0;
    auto prefetch = [&](int c, unsigned long long* dst) {  // rows [0, (c+1)*64) of column block c: (c+1)*512 bytes
        if (!STREAM) return;
        const unsigned long long* src = maskT + (size_t)c * rows_pad;
        for (int t = tid; t < (c + 1) * 32; t += REDUCE_THREADS) cp_async16(dst + 2 * t, src + 2 * t);
    };
    if (chunks > 0) prefetch(0, buf0);
    asm volatile("cp.async.commit_group;");
    __syncthreads();
    for (int c = 0; c < chunks; ++c) {
        const unsigned long long* cur_buf = STREAM ? ((c & 1) ? buf1 : buf0) : maskT + (size_t)c * rows_pad;
        if (c + 1 < chunks) prefetch(c + 1, (c & 1) ? buf0 : buf1);
        asm volatile("cp.async.commit_group;");
        asm volatile("cp.async.wait_group 1;");
        __syncthreads();
        const int base = c * 64, sz = min(64, n - base);
        const int nk = kept_s;
        unsigned long long part = 0ull;
        for (int t = tid; t < nk; t += REDUCE_THREADS) part |= cur_buf[kept_list[t]];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) part |= __shfl_xor_sync(0xffffffffu, part, o);
        if (lane == 0) wor[warp] = part;
        __syncthreads();
        if (tid == 0) {
            unsigned long long cur = 0ull;
            for (int w = 0; w < REDUCE_THREADS / 32; ++w) cur |= wor[w];
            const unsigned long long valid = sz == 64 ? ~0ull : ((1ull << sz) - 1ull);
            int k = nk;
            unsigned long long avail = ~cur & valid;
            while (avail) {
                if (max_keep > 0 && kbase + k >= max_keep) break;
                const int b = __ffsll((long long)avail) - 1;
                kept_list[k] = base + b;
                keep[kbase + k] = idx ? idx[base + b] : base + b;
                ++k;
                cur |= cur_buf[base + b];                               // diagonal tile: bits above b only
                avail = ~cur & valid & ~((2ull << b) - 1ull);          // (2<<63) wraps to 0 -> mask 0xFF..FF, avail 0
            }
            kept_s = k;
        }
        __syncthreads();
        if (max_keep > 0 && kbase + kept_s >= max_keep) break;
    }
    asm volatile("cp.async.wait_group 0;");
    if (tid == 0) *num_keep = kbase + kept_s;
}

// ------------------------------------------------------------------ two-level NMS: prefix, wide suppression, survivors
// The bitmask costs n^2/2 pair tests although the greedy rule only ever reads the rows of KEPT boxes. Greedy NMS over a
// score-sorted list has the prefix property (the kept boxes of the first P boxes are exactly the kept boxes < P of the
// whole list), so: (1) mask + reduce over the first P boxes; (2) every later box is tested against those few kept boxes
// only (k-major rounds, already-suppressed boxes are skipped) - typically most of the list dies here; (3) the survivors
// are compacted in index order and (4) mask + reduce run over the survivors alone, appending to the keep list. The keep
// list is identical to the full-mask result: a survivor can only be suppressed by a kept survivor.
constexpr int SUP_THREADS = 1024;
constexpr int SUP_CAND = 512;     // candidates per CTA
constexpr int SUP_KCH = 16;       // kept boxes per round: SUP_CAND * SUP_KCH pairs never overflow the queue

template <bool ROTATED>
__global__ void __launch_bounds__(SUP_THREADS) nms_suppress_kernel(int n, float thresh, const float* __restrict__ boxes,
                                                                   const RBox* __restrict__ rb, const float4* __restrict__ circ,
                                                                   const long long* __restrict__ keep, const int* __restrict__ num_keep,
                                                                   const int* __restrict__ counts, int n_max, int keep_stride,
                                                                   int prefix, int* __restrict__ alive) {
    const int f = blockIdx.y;
    if (counts) n = min(counts[f], n_max);
    const int i0 = prefix + blockIdx.x * SUP_CAND;
    boxes += (size_t)f * n_max * 7;
    rb += (size_t)f * n_max;
    circ += (size_t)f * n_max;
    alive += (size_t)f * n_max;
    if (counts) { keep += (size_t)f * keep_stride; num_keep += f; }
    const int tid = threadIdx.x, lane = tid & 31;
    if (i0 >= n) {   // nothing (left) in this slice: still clear the flags the compaction reads
        for (int t = tid; t < SUP_CAND; t += SUP_THREADS) if (i0 + t < n_max) alive[i0 + t] = 0;
        return;
    }
    __shared__ float4 ccirc[SUP_CAND];
    __shared__ float4 kcirc[SUP_KCH];
    __shared__ int kidx[SUP_KCH];
    __shared__ int supp[SUP_CAND];
    __shared__ unsigned int queue[SUP_CAND * SUP_KCH];
    __shared__ int qn;
    const int gsz = min(SUP_CAND, n - i0), nk = *num_keep;
    const bool all_near = thresh < 0.f;
    for (int t = tid; t < SUP_CAND; t += SUP_THREADS) { supp[t] = 0; if (t < gsz) ccirc[t] = circ[i0 + t]; }
    for (int k0 = 0; k0 < nk; k0 += SUP_KCH) {
        const int kn = min(SUP_KCH, nk - k0);
        __syncthreads();
        if (tid < kn) { const int b = (int)keep[k0 + tid]; kidx[tid] = b; kcirc[tid] = circ[b]; }
        if (tid == 0) qn = 0;
        __syncthreads();
        for (int p0 = 0; p0 < kn * gsz; p0 += SUP_THREADS) {
            const int p = p0 + tid;
            bool near = false;
            int k = 0, t = 0;
            if (p < kn * gsz) {
                k = p / gsz; t = p - k * gsz;
                near = supp[t] == 0 && (all_near || (ROTATED ? circ_near(kcirc[k], ccirc[t]) : aabb_near(kcirc[k], ccirc[t])));
            }
            const unsigned int m = __ballot_sync(0xffffffffu, near);
            if (m) {
                int base = 0;
                if (lane == 0) base = atomicAdd(&qn, __popc(m));
                base = __shfl_sync(0xffffffffu, base, 0);
                if (near) queue[base + __popc(m & ((1u << lane) - 1u))] = ((unsigned int)k << 9) | (unsigned int)t;
            }
        }
        __syncthreads();
        const int nq = qn;
        for (int q = tid; q < nq; q += SUP_THREADS) {
            const int k = queue[q] >> 9, t = queue[q] & 511;
            if (supp[t]) continue;                 // benign race: a stale 0 only costs one evaluation
            if (pair_iou<ROTATED>(boxes, rb, kidx[k], i0 + t) > thresh) supp[t] = 1;
        }
    }
    __syncthreads();
    for (int t = tid; t < SUP_CAND; t += SUP_THREADS)
        if (i0 + t < n_max) alive[i0 + t] = (t < gsz && supp[t] == 0) ? 1 : 0;
}

// ordered compaction of the alive flags (positions >= prefix) into idx; m = 0 when the keep list is already full
__global__ void __launch_bounds__(1024) nms_compact_kernel(int n, const int* __restrict__ counts, int n_max, int prefix,
                                                           int max_keep, const int* __restrict__ num_keep,
                                                           const int* __restrict__ alive, int* __restrict__ idx,
                                                           int* __restrict__ m_dev) {
    const int f = blockIdx.x;
    if (counts) { n = min(counts[f], n_max); num_keep += f; }
    alive += (size_t)f * n_max;
    idx += (size_t)f * n_max;
    __shared__ int scan_s[33];
    __shared__ int run;
    if (threadIdx.x == 0) run = 0;
    __syncthreads();
    const bool full = max_keep > 0 && *num_keep >= max_keep;
    if (!full) {
        for (int i0 = prefix; i0 < n; i0 += 1024) {
            const int i = i0 + threadIdx.x;
            const int a = (i < n) ? alive[i] : 0;
            int total;
            const int ex = block_excl_scan(a, scan_s, &total);
            if (a) idx[run + ex] = i;
            __syncthreads();
            if (threadIdx.x == 0) run += total;
            __syncthreads();
        }
    }
    if (threadIdx.x == 0) m_dev[f] = run;
}

struct NmsWs {
    unsigned long long* maskT;
    RBox* rb;
    float4* circ;
    int* alive;   // (B, rows_pad) survivor flags of the two-level path
    int* idx;     // (B, rows_pad) compacted survivor indices
    int* m_dev;   // (B) survivor counts
    unsigned int* gq;   // global pair queue (null when n or B exceed its bit fields)
    int* gq_count;      // two counters: level 1 / level 2
    int gq_cap;
    int rows_pad;
};

size_t nms_gq_cap(int B, int n) {   // pairs; 0 = no global queue
    if (B > 16 || n > 16384) return 0;
    const size_t all = (size_t)B * n * n / 2;
    return all < ((size_t)4 << 20) ? (all < 1024 ? 1024 : all) : ((size_t)4 << 20);
}

size_t nms_ws_bytes(int B, int n_max) {
    const size_t cb = (size_t)crb3d_divup(n_max > 0 ? n_max : 1, 64), nb = (size_t)(B > 0 ? B : 1);
    return crb3d_align(8 * nb * cb * cb * 64) + crb3d_align(sizeof(RBox) * nb * cb * 64) + crb3d_align(16 * nb * cb * 64) +
           2 * crb3d_align(4 * nb * cb * 64) + crb3d_align(4 * nb) + crb3d_align(4 * nms_gq_cap((int)nb, (int)cb * 64)) + crb3d_align(8);
}

bool nms_ws_take(void* ws, size_t ws_bytes, int B, int n_max, NmsWs& w) {
    WsCursor c(ws, ws_bytes);
    const size_t cb = (size_t)crb3d_divup(n_max, 64);
    w.rows_pad = (int)(cb * 64);
    w.maskT = c.take<unsigned long long>((size_t)B * cb * cb * 64);
    w.rb = c.take<RBox>((size_t)B * cb * 64);
    w.circ = c.take<float4>((size_t)B * cb * 64);
    w.alive = c.take<int>((size_t)B * cb * 64);
    w.idx = c.take<int>((size_t)B * cb * 64);
    w.m_dev = c.take<int>((size_t)B);
    w.gq_cap = (int)nms_gq_cap(B, w.rows_pad);
    w.gq = c.take<unsigned int>((size_t)w.gq_cap);
    w.gq_count = c.take<int>(2);
    if (w.gq_cap == 0) w.gq = nullptr;
    return c.ok;
}

// level: prefix > 0 -> first `prefix` boxes; idx/m_dev -> survivors (second level); both 0/null -> all boxes
int launch_mask(const float* boxes, const int* counts, int B, int n, float thresh, int rotated, const NmsWs& w, int prefix,
                const int* idx, const int* m_dev, int level, cudaStream_t stream) {
    const int cb = (int)crb3d_divup(prefix > 0 ? (prefix < n ? prefix : n) : n, 64);
    const dim3 mg((unsigned)crb3d_divup(cb, MASK_CW), cb, B);
    int* cnt = w.gq ? w.gq_count + level : nullptr;
    if (rotated)
        nms_mask_kernel<true><<<mg, MASK_THREADS, 0, stream>>>(n, thresh, boxes, w.rb, w.circ, w.maskT, w.rows_pad, counts, n, prefix, idx, m_dev, w.gq, cnt, w.gq_cap);
    else
        nms_mask_kernel<false><<<mg, MASK_THREADS, 0, stream>>>(n, thresh, boxes, w.rb, w.circ, w.maskT, w.rows_pad, counts, n, prefix, idx, m_dev, w.gq, cnt, w.gq_cap);
    if (w.gq) {
        if (rotated) nms_eval_kernel<true><<<crb3d_num_sms() * 4, 256, 0, stream>>>(w.gq, cnt, w.gq_cap, thresh, boxes, w.rb, idx, n, w.maskT, w.rows_pad);
        else nms_eval_kernel<false><<<crb3d_num_sms() * 4, 256, 0, stream>>>(w.gq, cnt, w.gq_cap, thresh, boxes, w.rb, idx, n, w.maskT, w.rows_pad);
    }
    CRB3D_CHECK_LAUNCH();
    return CRB3D_OK;
}

int launch_prep(const float* boxes, const int* counts, int B, int n, int rotated, const NmsWs& w, cudaStream_t stream) {
    if (w.gq) CRB3D_CUDA(cudaMemsetAsync(w.gq_count, 0, 2 * sizeof(int), stream));
    const dim3 pg((unsigned)crb3d_divup(n, 256), B);
    if (rotated) nms_prep_kernel<true><<<pg, 256, 0, stream>>>(n, boxes, counts, n, w.rb, w.circ);
    else nms_prep_kernel<false><<<pg, 256, 0, stream>>>(n, boxes, counts, n, w.rb, w.circ);
    CRB3D_CHECK_LAUNCH();
    return CRB3D_OK;
}

int launch_reduce(const int* counts, int B, int n, int max_keep, long long* keep, int keep_stride, int* num_keep, const NmsWs& w,
                  int prefix, const int* idx, const int* m_dev, cudaStream_t stream) {
    const int kcap = max_keep > 0 ? (max_keep < n ? max_keep : n) : n;
    const size_t list = sizeof(int) * (size_t)kcap + 16;
    const bool stream_mode = 16 * (size_t)w.rows_pad + list <= 200 * 1024;
    const size_t smem = (stream_mode ? 16 * (size_t)w.rows_pad : 0) + list;
    if (smem > 200 * 1024) return CRB3D_ERR_UNSUPPORTED;
    static size_t smem_set_dev[CRB3D_MAX_DEVICES][2] = {};   // the attribute is per function per device
    size_t* smem_set = smem_set_dev[crb3d_current_device()];
    if (smem > 40 * 1024 && smem > smem_set[stream_mode]) {
        if (stream_mode) CRB3D_CUDA(cudaFuncSetAttribute(nms_reduce_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        else CRB3D_CUDA(cudaFuncSetAttribute(nms_reduce_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        smem_set[stream_mode] = smem;
    }
    if (stream_mode)
        nms_reduce_kernel<true><<<B, REDUCE_THREADS, smem, stream>>>(n, w.rows_pad, w.maskT, max_keep, keep, num_keep, counts, n, keep_stride, prefix, idx, m_dev);
    else
        nms_reduce_kernel<false><<<B, REDUCE_THREADS, smem, stream>>>(n, w.rows_pad, w.maskT, max_keep, keep, num_keep, counts, n, keep_stride, prefix, idx, m_dev);
    CRB3D_CHECK_LAUNCH();
    return CRB3D_OK;
}

constexpr int NMS_PREFIX = 256;

int launch_nms(const float* boxes, const int* counts, int B, int n, float thresh, int rotated, int max_keep, long long* keep,
               int keep_stride, int* num_keep, const NmsWs& w, cudaStream_t stream) {
    int rc = launch_prep(boxes, counts, B, n, rotated, w, stream);
    if (rc) return rc;
    if (n <= 2 * NMS_PREFIX) {   // short lists: one level
        rc = launch_mask(boxes, counts, B, n, thresh, rotated, w, 0, nullptr, nullptr, 0, stream);
        if (rc) return rc;
        return launch_reduce(counts, B, n, max_keep, keep, keep_stride, num_keep, w, 0, nullptr, nullptr, stream);
    }
    // level 1: the first NMS_PREFIX boxes
    rc = launch_mask(boxes, counts, B, n, thresh, rotated, w, NMS_PREFIX, nullptr, nullptr, 0, stream);
    if (rc) return rc;
    rc = launch_reduce(counts, B, n, max_keep, keep, keep_stride, num_keep, w, NMS_PREFIX, nullptr, nullptr, stream);
    if (rc) return rc;
    // every later box against the boxes kept so far, then the ordered list of survivors
    const dim3 sg((unsigned)crb3d_divup(n - NMS_PREFIX, SUP_CAND), B);
    if (rotated)
        nms_suppress_kernel<true><<<sg, SUP_THREADS, 0, stream>>>(n, thresh, boxes, w.rb, w.circ, keep, num_keep, counts, n, keep_stride, NMS_PREFIX, w.alive);
    else
        nms_suppress_kernel<false><<<sg, SUP_THREADS, 0, stream>>>(n, thresh, boxes, w.rb, w.circ, keep, num_keep, counts, n, keep_stride, NMS_PREFIX, w.alive);
    nms_compact_kernel<<<B, 1024, 0, stream>>>(n, counts, n, NMS_PREFIX, max_keep, num_keep, w.alive, w.idx, w.m_dev);
    CRB3D_CHECK_LAUNCH();
    // level 2: the survivors among themselves, appended to the keep list
    rc = launch_mask(boxes, counts, B, n, thresh, rotated, w, 0, w.idx, w.m_dev, 1, stream);
    if (rc) return rc;
    return launch_reduce(counts, B, n, max_keep, keep, keep_stride, num_keep, w, 0, w.idx, w.m_dev, stream);
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
    *bytes = nms_ws_bytes(1, n);
    return CRB3D_OK;
}

// boxes: (n,7) sorted by descending score. keep: device int64[n] (first *num_keep entries valid, ascending).
extern "C" int crb3d_nms(const float* boxes, int n, float thresh, int rotated, int max_keep, long long* keep,
                         int* num_keep, void* ws, size_t ws_bytes, cudaStream_t stream) {
    if (n < 0 || !keep || !num_keep) return CRB3D_ERR_ARG;
    if (n == 0) { CRB3D_CUDA(cudaMemsetAsync(num_keep, 0, sizeof(int), stream)); return CRB3D_OK; }
    NmsWs w;
    if (!nms_ws_take(ws, ws_bytes, 1, n, w)) return CRB3D_ERR_WORKSPACE;
    return launch_nms(boxes, nullptr, 1, n, thresh, rotated, max_keep, keep, n, num_keep, w, stream);
}

// Raw suppression bitmask in the reference's row-major layout mask[n][ceil(n/64)] (cells on and above the diagonal
// block; the rest is left untouched), for parity checks against the reference nms_kernel. ws: crb3d_nms_workspace_bytes(n).
extern "C" int crb3d_nms_mask(const float* boxes, int n, float thresh, int rotated, unsigned long long* mask, void* ws,
                              size_t ws_bytes, cudaStream_t stream) {
    if (n <= 0 || !mask) return CRB3D_ERR_ARG;
    NmsWs w;
    if (!nms_ws_take(ws, ws_bytes, 1, n, w)) return CRB3D_ERR_WORKSPACE;
    int rc = launch_prep(boxes, nullptr, 1, n, rotated, w, stream);
    if (rc) return rc;
    rc = launch_mask(boxes, nullptr, 1, n, thresh, rotated, w, 0, nullptr, nullptr, 0, stream);
    if (rc) return rc;
    const int cb = (int)crb3d_divup(n, 64);
    nms_mask_transpose_kernel<<<dim3((unsigned)crb3d_divup(n, 256), cb), 256, 0, stream>>>(n, cb, w.rows_pad, w.maskT, mask);
    CRB3D_CHECK_LAUNCH();
    return CRB3D_OK;
}

// Batched NMS for the scoring path: frame b owns boxes[b][0..counts[b]) of a padded (B, n_max, 7) tensor (each frame
// sorted by descending score). keep: (B, keep_stride) int64, num_keep: (B). prep + mask + reduce launches for the whole
// batch, no host synchronisation.
extern "C" int crb3d_nms_batched_workspace_bytes(int B, int n_max, size_t* bytes) {
    if (!bytes || B < 0 || n_max < 0) return CRB3D_ERR_ARG;
    *bytes = nms_ws_bytes(B, n_max);
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
    NmsWs w;
    if (!nms_ws_take(ws, ws_bytes, B, n_max, w)) return CRB3D_ERR_WORKSPACE;
    return launch_nms(boxes, counts, B, n_max, thresh, rotated, max_keep, keep, keep_stride, num_keep, w, stream);
}
