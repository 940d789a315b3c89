// Rulebook (indice-pair) construction for SubMConv3d / SparseConv3d.
//
// Replaces the external spconv-cu113==2.1.21 `get_indice_pairs` reached from the reference at
//   pcdet/models/backbones_3d/spconv_backbone.py:12-17,77-117 (SubMConv3d / SparseConv3d constructors + forward)
// Contract: kernel offset id k = (kz*KY + ky)*KX + kx; a pair (i -> o) exists for offset k iff
//   p_in = p_out * stride - pad + k_vec * dilation   (cross-correlation, identical to torch conv3d).
//
// B200-first layout: instead of spconv's atomically-compacted `indice_pairs[2,K,N]` we emit the
// OUTPUT-STATIONARY neighbour table  nbr[k][o] = input row or -1  (and its transpose nbr_t[k][i] = output row
// or -1 for the backward pass).  Every (k,o) has at most one input, so the table is written without atomics,
// is deterministic, coalesces (k-major, rows contiguous) and lets the conv kernel accumulate over k in TMEM /
// registers with a single store per output row.  `crb3d_rulebook_compact_pairs` derives the spconv-format
// pair lists from it (ascending output row per offset) for API parity.
//
// Strided conv output rows are in ascending linear (b,z,y,x) key order (the spconv-GPU order): an output-cell
// bitmap is filled with atomicOr, a popcount scan ranks the set bits, no sort is needed.
#include "common.cuh"

namespace {

struct ConvGeom {
    int in_shape[3];   // D,H,W
    int out_shape[3];  // D,H,W
    int k[3], s[3], p[3], d[3];
    int K;  // k[0]*k[1]*k[2]
};

__device__ __forceinline__ unsigned long long lin_key(int b, int z, int y, int x, const int* shp) {
    return (((unsigned long long)b * shp[0] + z) * shp[1] + y) * shp[2] + x;
}

// ---------------------------------------------------------------- SubM ------------------------
// n_dev (nullable): device-side row count - the host-side n is then only the capacity / row stride of the tables, so a
// whole step can be recorded in a CUDA graph without reading any count back
__global__ void __launch_bounds__(256) subm_insert(const int* __restrict__ coords, int n, ConvGeom g,
                                                   unsigned long long* __restrict__ keys, int* __restrict__ vals,
                                                   uint32_t cap_mask, const int* __restrict__ n_dev) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || (n_dev && i >= *n_dev)) return;
    int4 c = reinterpret_cast<const int4*>(coords)[i];
    uint32_t s = hash_insert(keys, cap_mask, lin_key(c.x, c.y, c.z, c.w, g.in_shape));
    vals[s] = i;  // coordinates are unique (voxelizer / previous rulebook guarantee it)
}

// One thread per (k, o): k-major so a warp reads 32 consecutive rows and writes 32 consecutive table cells.
__global__ void __launch_bounds__(256) subm_lookup(const int* __restrict__ coords, int n, ConvGeom g,
                                                   const unsigned long long* __restrict__ keys,
                                                   const int* __restrict__ vals, uint32_t cap_mask,
                                                   int* __restrict__ nbr, const int* __restrict__ n_dev) {
    int o = blockIdx.x * blockDim.x + threadIdx.x;
    int k = blockIdx.y;
    if (o >= n || (n_dev && o >= *n_dev)) return;
    int4 c = __ldg(reinterpret_cast<const int4*>(coords) + o);
    int kx = k % g.k[2], ky = (k / g.k[2]) % g.k[1], kz = k / (g.k[2] * g.k[1]);
    int z = c.y - g.p[0] + kz * g.d[0];
    int y = c.z - g.p[1] + ky * g.d[1];
    int x = c.w - g.p[2] + kx * g.d[2];
    int r = -1;
    if (z >= 0 && z < g.in_shape[0] && y >= 0 && y < g.in_shape[1] && x >= 0 && x < g.in_shape[2]) {
        uint32_t s = hash_find(keys, cap_mask, lin_key(c.x, z, y, x, g.in_shape));
        if (s != 0xFFFFFFFFu) r = vals[s];
    }
    nbr[(size_t)k * n + o] = r;
}

// ---------------------------------------------------------------- strided ---------------------
// For input coordinate p and offset k the output coordinate (if any): (p + pad - k*dil) / stride, exact.
__device__ __forceinline__ bool out_coord(const ConvGeom& g, int4 c, int k, int& oz, int& oy, int& ox) {
    int kx = k % g.k[2], ky = (k / g.k[2]) % g.k[1], kz = k / (g.k[2] * g.k[1]);
    int z = c.y + g.p[0] - kz * g.d[0];
    int y = c.z + g.p[1] - ky * g.d[1];
    int x = c.w + g.p[2] - kx * g.d[2];
    if (z < 0 || y < 0 || x < 0) return false;
    if (z % g.s[0] || y % g.s[1] || x % g.s[2]) return false;
    oz = z / g.s[0]; oy = y / g.s[1]; ox = x / g.s[2];
    return oz < g.out_shape[0] && oy < g.out_shape[1] && ox < g.out_shape[2];
}

// The valid (offset, output cell) pairs of ONE input row, enumerated with per-axis pruning: a k=3, s=2 conv reaches 1..8 of
// its 27 offsets per input (odd coordinates pair with two taps per axis, even ones with one), so one thread per row does the
// work the first version spread over 27 mostly idle threads.
template <typename F>
__device__ __forceinline__ void for_each_output(const ConvGeom& g, int4 c, F f) {
    for (int kz = 0; kz < g.k[0]; ++kz) {
        const int z = c.y + g.p[0] - kz * g.d[0];
        if (z < 0 || z % g.s[0]) continue;
        const int oz = z / g.s[0];
        if (oz >= g.out_shape[0]) continue;
        for (int ky = 0; ky < g.k[1]; ++ky) {
            const int y = c.z + g.p[1] - ky * g.d[1];
            if (y < 0 || y % g.s[1]) continue;
            const int oy = y / g.s[1];
            if (oy >= g.out_shape[1]) continue;
            for (int kx = 0; kx < g.k[2]; ++kx) {
                const int x = c.w + g.p[2] - kx * g.d[2];
                if (x < 0 || x % g.s[2]) continue;
                const int ox = x / g.s[2];
                if (ox >= g.out_shape[2]) continue;
                f((kz * g.k[1] + ky) * g.k[2] + kx, lin_key(c.x, oz, oy, ox, g.out_shape));
            }
        }
    }
}

__global__ void __launch_bounds__(256) sparse_mark(const int* __restrict__ coords, int n, ConvGeom g,
                                                   unsigned int* __restrict__ bitmap, const int* __restrict__ n_dev) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || (n_dev && i >= *n_dev)) return;
    const int4 c = __ldg(reinterpret_cast<const int4*>(coords) + i);
    for_each_output(g, c, [&](int, unsigned long long key) {
        unsigned int* w = bitmap + (key >> 5);
        const unsigned int bit = 1u << (key & 31);
        // neighbouring inputs hit the same cells: test first (plain load, the bit only ever goes 0 -> 1), atomics for the rest
        if (!(*reinterpret_cast<volatile unsigned int*>(w) & bit)) atomicOr(w, bit);
    });
}

// popcount + exclusive scan of the bitmap words + emission of the active output coordinates, in two passes over the bitmap
// (tile sums, then ranks + coordinates) around the one-block scan of the tile sums; a tile = CRB3D_SCAN_TILE words
constexpr int BM_ITEMS = CRB3D_SCAN_TILE / 256;   // 8 words per thread = two 16-byte loads
static_assert(BM_ITEMS == 8, "bitmap kernels read two uint4 per thread");

__device__ __forceinline__ void load_words8(const unsigned int* __restrict__ bitmap, int64_t base, int64_t nwords, unsigned int (&w)[8]) {
    if (base + 8 <= nwords) {
        const uint4 a = __ldg(reinterpret_cast<const uint4*>(bitmap + base)), b = __ldg(reinterpret_cast<const uint4*>(bitmap + base) + 1);
        w[0] = a.x; w[1] = a.y; w[2] = a.z; w[3] = a.w; w[4] = b.x; w[5] = b.y; w[6] = b.z; w[7] = b.w;
    } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) w[i] = (base + i < nwords) ? __ldg(bitmap + base + i) : 0u;
    }
}

__global__ void __launch_bounds__(256) bitmap_tile_sums(const unsigned int* __restrict__ bitmap, int64_t nwords, int* __restrict__ sums) {
    __shared__ int sm[33];
    unsigned int w[8];
    load_words8(bitmap, (int64_t)blockIdx.x * CRB3D_SCAN_TILE + (int64_t)threadIdx.x * 8, nwords, w);
    int s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += __popc(w[i]);
    int tot;
    block_excl_scan(s, sm, &tot);
    if (threadIdx.x == 0) sums[blockIdx.x] = tot;
}

__global__ void __launch_bounds__(256) bitmap_rank(const unsigned int* __restrict__ bitmap, int64_t nwords, const int* __restrict__ sums,
                                                   int* __restrict__ rank) {
    __shared__ int sm[33];
    const int64_t base = (int64_t)blockIdx.x * CRB3D_SCAN_TILE + (int64_t)threadIdx.x * 8;
    unsigned int w[8];
    load_words8(bitmap, base, nwords, w);
    int s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += __popc(w[i]);
    int tot;
    int ex = block_excl_scan(s, sm, &tot) + sums[blockIdx.x];
    int r[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { r[i] = ex; ex += __popc(w[i]); }
    if (base + 8 <= nwords) {
        reinterpret_cast<int4*>(rank + base)[0] = make_int4(r[0], r[1], r[2], r[3]);
        reinterpret_cast<int4*>(rank + base)[1] = make_int4(r[4], r[5], r[6], r[7]);
    } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) if (base + i < nwords) rank[base + i] = r[i];
    }
}

// Active output coordinates in rank (= ascending key) order. A warp owns 32 consecutive bitmap words: every lane decodes the
// coordinates of ITS word's first cell once (the only divisions), then the warp walks the non-empty words with lane = bit, so
// a dense word emits its 32 rows in one step and an empty region costs one ballot.
__global__ void __launch_bounds__(256) bitmap_emit_coords(const unsigned int* __restrict__ bitmap, int64_t nwords, const int* __restrict__ rank,
                                                          ConvGeom g, int* __restrict__ coords_out, int cap_out) {
    const int64_t wi = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    const unsigned int word = wi < nwords ? __ldg(&bitmap[wi]) : 0u;
    unsigned int nz = __ballot_sync(0xffffffffu, word != 0u);
    if (!nz) return;
    const int r0 = word ? __ldg(&rank[wi]) : 0;
    unsigned long long key = (unsigned long long)wi << 5;
    const int W = g.out_shape[2], H = g.out_shape[1], D = g.out_shape[0];
    const int x0 = (int)(key % W); key /= W;
    const int y0 = (int)(key % H); key /= H;
    const int z0 = (int)(key % D);
    const int b0 = (int)(key / D);
    while (nz) {
        const int src = __ffs(nz) - 1;
        nz &= nz - 1;
        const unsigned int w = __shfl_sync(0xffffffffu, word, src);
        const int r = __shfl_sync(0xffffffffu, r0, src);
        int x = __shfl_sync(0xffffffffu, x0, src) + lane, y = __shfl_sync(0xffffffffu, y0, src);
        int z = __shfl_sync(0xffffffffu, z0, src), bb = __shfl_sync(0xffffffffu, b0, src);
        if ((w >> lane) & 1u) {
            const int o = r + __popc(w & ((1u << lane) - 1u));
            if (o < cap_out) {
                while (x >= W) {           // the word straddles a row end (rows shorter than 32 cells wrap more than once)
                    x -= W;
                    if (++y == H) { y = 0; if (++z == D) { z = 0; ++bb; } }
                }
                reinterpret_cast<int4*>(coords_out)[o] = make_int4(bb, z, y, x);
            }
        }
    }
}

__global__ void __launch_bounds__(256) sparse_fill_pairs(const int* __restrict__ coords, int n, ConvGeom g,
                                                         const unsigned int* __restrict__ bitmap,
                                                         const int* __restrict__ rank, int n_out,
                                                         int* __restrict__ nbr, int* __restrict__ nbr_t,
                                                         const int* __restrict__ n_dev) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || (n_dev && i >= *n_dev)) return;
    const int4 c = __ldg(reinterpret_cast<const int4*>(coords) + i);
    if (nbr_t)
        for (int k = 0; k < g.K; ++k) nbr_t[(size_t)k * n + i] = -1;      // coalesced per offset plane; the hits below overwrite
    for_each_output(g, c, [&](int k, unsigned long long key) {
        const unsigned int word = __ldg(&bitmap[key >> 5]);
        const int o = __ldg(&rank[key >> 5]) + __popc(word & ((1u << (key & 31)) - 1u));
        if (o < n_out) {
            nbr[(size_t)k * n_out + o] = i;
            if (nbr_t) nbr_t[(size_t)k * n + i] = o;
        }
    });
}

// SubM neighbour table of a level whose rows were ranked by a strided rulebook (its output-cell bitmap + word ranks =
// cell -> row map): two loads per (z, y) line instead of a hash probe chain per neighbour, no table to build. One thread per
// row walks its k^3 neighbourhood; consecutive kx share a bitmap word. A rank at or beyond the valid row count (truncated
// level: the capacity was exceeded) reads as "no neighbour".
__global__ void __launch_bounds__(256) subm_lookup_cellmap(const int* __restrict__ coords, int n, ConvGeom g,
                                                           const unsigned int* __restrict__ bitmap, const int* __restrict__ rank,
                                                           int* __restrict__ nbr, const int* __restrict__ n_dev) {
    const int o = blockIdx.x * blockDim.x + threadIdx.x;
    const int nv = n_dev ? min(n, *n_dev) : n;
    if (o >= nv) return;
    const int4 c = __ldg(reinterpret_cast<const int4*>(coords) + o);
    unsigned long long last_w = ~0ull;
    unsigned int word = 0;
    int base = 0, k = 0;
    for (int kz = 0; kz < g.k[0]; ++kz) {
        const int z = c.y - g.p[0] + kz * g.d[0];
        for (int ky = 0; ky < g.k[1]; ++ky) {
            const int y = c.z - g.p[1] + ky * g.d[1];
            const bool line_ok = z >= 0 && z < g.in_shape[0] && y >= 0 && y < g.in_shape[1];
            for (int kx = 0; kx < g.k[2]; ++kx, ++k) {
                const int x = c.w - g.p[2] + kx * g.d[2];
                int r = -1;
                if (line_ok && x >= 0 && x < g.in_shape[2]) {
                    const unsigned long long key = lin_key(c.x, z, y, x, g.in_shape);
                    if ((key >> 5) != last_w) {
                        last_w = key >> 5;
                        word = __ldg(&bitmap[last_w]);
                        base = word ? __ldg(&rank[last_w]) : 0;
                    }
                    const unsigned int bit = (unsigned int)(key & 31);
                    if ((word >> bit) & 1u) {
                        r = base + __popc(word & ((1u << bit) - 1u));
                        if (r >= nv) r = -1;
                    }
                }
                nbr[(size_t)k * n + o] = r;
            }
        }
    }
}

// ---------------------------------------------------------------- compaction to spconv pair lists
__global__ void __launch_bounds__(256) pair_flags(const int* __restrict__ nbr, int64_t total, int* __restrict__ flags) {
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < total) flags[t] = nbr[t] >= 0;
}

__global__ void __launch_bounds__(256) pair_write(const int* __restrict__ nbr, int K, int n_out,
                                                  const int* __restrict__ rank, int pair_cap,
                                                  int* __restrict__ pairs_in, int* __restrict__ pairs_out) {
    int o = blockIdx.x * blockDim.x + threadIdx.x;
    int k = blockIdx.y;
    if (o >= n_out) return;
    size_t t = (size_t)k * n_out + o;
    int i = nbr[t];
    if (i < 0) return;
    int pos = rank[t] - rank[(size_t)k * n_out];
    if (pos < pair_cap) {
        pairs_in[(size_t)k * pair_cap + pos] = i;
        pairs_out[(size_t)k * pair_cap + pos] = o;
    }
}

__global__ void pair_counts(const int* __restrict__ rank, const int* __restrict__ total, int K, int n_out,
                            int* __restrict__ pair_num) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= K) return;
    int a = rank[(size_t)k * n_out];
    int b = (k + 1 < K) ? rank[(size_t)(k + 1) * n_out] : *total;
    pair_num[k] = b - a;
}

bool make_geom(const int* in_shape, const int* out_shape, const int* k, const int* s, const int* p, const int* d,
               ConvGeom& g) {
    for (int j = 0; j < 3; ++j) {
        g.in_shape[j] = in_shape[j]; g.out_shape[j] = out_shape[j];
        g.k[j] = k[j]; g.s[j] = s ? s[j] : 1; g.p[j] = p[j]; g.d[j] = d ? d[j] : 1;
        if (g.k[j] <= 0 || g.s[j] <= 0 || g.d[j] <= 0 || g.in_shape[j] <= 0 || g.out_shape[j] <= 0) return false;
    }
    g.K = g.k[0] * g.k[1] * g.k[2];
    return true;
}

int64_t bitmap_words(int batch_size, const int* out_shape) {
    unsigned long long cells = (unsigned long long)batch_size * out_shape[0] * out_shape[1] * out_shape[2];
    return (int64_t)((cells + 31) / 32);
}

}  // namespace

// ------------------------------------------------------------------------------------------------ C ABI
extern "C" int crb3d_subm_rulebook_workspace_bytes(int n, size_t* bytes) {
    if (!bytes || n < 0) return CRB3D_ERR_ARG;
    uint32_t cap = crb3d_next_pow2((uint64_t)(n > 0 ? n : 1) * 2);
    *bytes = crb3d_align(sizeof(unsigned long long) * cap) + crb3d_align(sizeof(int) * cap);
    return CRB3D_OK;
}

extern "C" int crb3d_subm_rulebook(const int* coords, int n, const int* n_dev, const int* spatial_shape3, const int* ksize3,
                                   const int* dilation3, int* nbr, void* ws, size_t ws_bytes, cudaStream_t stream) {
    if (n < 0 || !spatial_shape3 || !ksize3 || (!nbr && n > 0)) return CRB3D_ERR_ARG;
    int pad[3];
    for (int j = 0; j < 3; ++j) {
        if (ksize3[j] % 2 == 0) return CRB3D_ERR_UNSUPPORTED;  // SubM needs odd kernels (centre = identity)
        pad[j] = (ksize3[j] / 2) * (dilation3 ? dilation3[j] : 1);
    }
    ConvGeom g;
    if (!make_geom(spatial_shape3, spatial_shape3, ksize3, nullptr, pad, dilation3, g)) return CRB3D_ERR_ARG;
    if (n == 0) return CRB3D_OK;
    WsCursor c(ws, ws_bytes);
    uint32_t cap = crb3d_next_pow2((uint64_t)n * 2);
    unsigned long long* keys = c.take<unsigned long long>(cap);
    int* vals = c.take<int>(cap);
    if (!c.ok) return CRB3D_ERR_WORKSPACE;
    CRB3D_CUDA(cudaMemsetAsync(keys, 0xFF, sizeof(unsigned long long) * cap, stream));
    const unsigned nb = (unsigned)crb3d_divup(n, 256);
    subm_insert<<<nb, 256, 0, stream>>>(coords, n, g, keys, vals, cap - 1, n_dev);
    subm_lookup<<<dim3(nb, g.K), 256, 0, stream>>>(coords, n, g, keys, vals, cap - 1, nbr, n_dev);
    CRB3D_CHECK_LAUNCH();
    return CRB3D_OK;
}

// SubM table of a level produced by a strided rulebook, looked up through that rulebook's cell -> row map (the bitmap + word
// ranks that crb3d_sparse_rulebook_coords leaves in its workspace; crb3d_sparse_rulebook_cellmap locates them). `coords` must be
// the out_coords of that call (rows in rank order) and spatial_shape3 its out_shape3. Same table as crb3d_subm_rulebook.
extern "C" int crb3d_subm_rulebook_cellmap(const int* coords, int n, const int* n_dev, const int* spatial_shape3, const int* ksize3,
                                           const int* dilation3, const unsigned int* cell_bitmap, const int* cell_rank, int* nbr,
                                           cudaStream_t stream) {
    if (n < 0 || !spatial_shape3 || !ksize3 || (!nbr && n > 0) || !cell_bitmap || !cell_rank) return CRB3D_ERR_ARG;
    int pad[3];
    for (int j = 0; j < 3; ++j) {
        if (ksize3[j] % 2 == 0) return CRB3D_ERR_UNSUPPORTED;
        pad[j] = (ksize3[j] / 2) * (dilation3 ? dilation3[j] : 1);
    }
    ConvGeom g;
    if (!make_geom(spatial_shape3, spatial_shape3, ksize3, nullptr, pad, dilation3, g)) return CRB3D_ERR_ARG;
    if (n == 0) return CRB3D_OK;
    subm_lookup_cellmap<<<(unsigned)crb3d_divup(n, 256), 256, 0, stream>>>(coords, n, g, cell_bitmap, cell_rank, nbr, n_dev);
    CRB3D_CHECK_LAUNCH();
    return CRB3D_OK;
}

// byte offsets of the cell bitmap and its word ranks inside the workspace of crb3d_sparse_rulebook_coords, and the word count
extern "C" int crb3d_sparse_rulebook_cellmap(int batch_size, const int* out_shape3, size_t* bitmap_offset, size_t* rank_offset,
                                             long long* n_words) {
    if (!out_shape3 || batch_size <= 0 || !bitmap_offset || !rank_offset || !n_words) return CRB3D_ERR_ARG;
    const int64_t nw = bitmap_words(batch_size, out_shape3);
    *bitmap_offset = 0;
    *rank_offset = crb3d_align(sizeof(unsigned int) * nw);
    *n_words = nw;
    return CRB3D_OK;
}

extern "C" int crb3d_conv_out_shape(const int* in_shape3, const int* ksize3, const int* stride3, const int* pad3,
                                    const int* dilation3, int* out_shape3) {
    if (!in_shape3 || !ksize3 || !stride3 || !pad3 || !out_shape3) return CRB3D_ERR_ARG;
    for (int j = 0; j < 3; ++j) {
        int d = dilation3 ? dilation3[j] : 1;
        out_shape3[j] = (in_shape3[j] + 2 * pad3[j] - d * (ksize3[j] - 1) - 1) / stride3[j] + 1;
    }
    return CRB3D_OK;
}

extern "C" int crb3d_sparse_rulebook_workspace_bytes(int batch_size, const int* out_shape3, size_t* bytes) {
    if (!bytes || !out_shape3 || batch_size <= 0) return CRB3D_ERR_ARG;
    int64_t nw = bitmap_words(batch_size, out_shape3);
    *bytes = crb3d_align(sizeof(unsigned int) * nw) + crb3d_align(sizeof(int) * nw) +
             crb3d_align(sizeof(int) * crb3d_scan_ws_ints(nw));
    return CRB3D_OK;
}

// n_in_dev (nullable, both phases): device-side input row count; n_in is then the capacity of coords_in / stride of nbr_t,
// and n_out of phase 2 the capacity of coords_out / stride of nbr (rows beyond the true counts are never written).
// Phase 1: active output coordinates in ascending key order. Writes min(count, cap_out) rows of coords_out and the
// true count to n_out_dev. The bitmap + ranks stay in `ws` for phase 2 (same ws, untouched in between).
extern "C" int crb3d_sparse_rulebook_coords(const int* coords_in, int n_in, const int* n_in_dev, int batch_size, const int* in_shape3,
                                            const int* out_shape3, const int* ksize3, const int* stride3,
                                            const int* pad3, const int* dilation3, int* coords_out, int cap_out,
                                            int* n_out_dev, void* ws, size_t ws_bytes, cudaStream_t stream) {
    ConvGeom g;
    if (n_in < 0 || batch_size <= 0 || !in_shape3 || !out_shape3 || !ksize3 || !stride3 || !pad3 || !n_out_dev ||
        !make_geom(in_shape3, out_shape3, ksize3, stride3, pad3, dilation3, g))
        return CRB3D_ERR_ARG;
    int64_t nw = bitmap_words(batch_size, out_shape3);
    WsCursor c(ws, ws_bytes);
    unsigned int* bitmap = c.take<unsigned int>(nw);
    int* rank = c.take<int>(nw);
    int* scan_ws = c.take<int>(crb3d_scan_ws_ints(nw));
    if (!c.ok) return CRB3D_ERR_WORKSPACE;
    CRB3D_CUDA(cudaMemsetAsync(bitmap, 0, sizeof(unsigned int) * nw, stream));
    if (n_in > 0) sparse_mark<<<(unsigned)crb3d_divup(n_in, 256), 256, 0, stream>>>(coords_in, n_in, g, bitmap, n_in_dev);
    const int64_t nt = crb3d_divup(nw, CRB3D_SCAN_TILE);
    bitmap_tile_sums<<<(unsigned)nt, 256, 0, stream>>>(bitmap, nw, scan_ws);
    int rc = crb3d_scan_block_sums(scan_ws, nt, n_out_dev, stream);
    if (rc) return rc;
    bitmap_rank<<<(unsigned)nt, 256, 0, stream>>>(bitmap, nw, scan_ws, rank);
    if (coords_out && cap_out > 0)
        bitmap_emit_coords<<<(unsigned)crb3d_divup(nw, 256), 256, 0, stream>>>(bitmap, nw, rank, g, coords_out, cap_out);
    CRB3D_CHECK_LAUNCH();
    return CRB3D_OK;
}

// Phase 2: neighbour tables. nbr is [K][n_out], nbr_t (optional) is [K][n_in].
extern "C" int crb3d_sparse_rulebook_pairs(const int* coords_in, int n_in, const int* n_in_dev, int batch_size, const int* in_shape3,
                                           const int* out_shape3, const int* ksize3, const int* stride3,
                                           const int* pad3, const int* dilation3, int n_out, int* nbr, int* nbr_t,
                                           void* ws, size_t ws_bytes, cudaStream_t stream) {
    ConvGeom g;
    if (n_in < 0 || n_out < 0 || (!nbr && n_out > 0) || !make_geom(in_shape3, out_shape3, ksize3, stride3, pad3, dilation3, g))
        return CRB3D_ERR_ARG;
    if (n_in == 0 || n_out == 0) {  // nothing can pair up
        if (nbr && n_out > 0) CRB3D_CUDA(cudaMemsetAsync(nbr, 0xFF, sizeof(int) * (size_t)g.K * n_out, stream));
        if (nbr_t && n_in > 0) CRB3D_CUDA(cudaMemsetAsync(nbr_t, 0xFF, sizeof(int) * (size_t)g.K * n_in, stream));
        return CRB3D_OK;
    }
    int64_t nw = bitmap_words(batch_size, out_shape3);
    WsCursor c(ws, ws_bytes);
    unsigned int* bitmap = c.take<unsigned int>(nw);
    int* rank = c.take<int>(nw);
    if (!c.ok) return CRB3D_ERR_WORKSPACE;
    CRB3D_CUDA(cudaMemsetAsync(nbr, 0xFF, sizeof(int) * (size_t)g.K * n_out, stream));
    if (n_in > 0)
        sparse_fill_pairs<<<(unsigned)crb3d_divup(n_in, 256), 256, 0, stream>>>(coords_in, n_in, g, bitmap, rank, n_out, nbr, nbr_t,
                                                                              n_in_dev);
    CRB3D_CHECK_LAUNCH();
    return CRB3D_OK;
}

// spconv-format pair lists from a neighbour table: pairs_in/out are [K][pair_cap] (-1 padded by the caller),
// pair_num[K]. Within an offset the pairs are in ascending output row. ws: K*n_out*2 ints + scan scratch.
extern "C" int crb3d_rulebook_compact_pairs_workspace_bytes(int K, int n_out, size_t* bytes) {
    if (!bytes || K <= 0 || n_out < 0) return CRB3D_ERR_ARG;
    int64_t t = (int64_t)K * n_out;
    *bytes = crb3d_align(sizeof(int) * (t > 0 ? t : 1)) + crb3d_align(sizeof(int) * crb3d_scan_ws_ints(t)) + 256;
    return CRB3D_OK;
}

extern "C" int crb3d_rulebook_compact_pairs(const int* nbr, int K, int n_out, int pair_cap, int* pairs_in,
                                            int* pairs_out, int* pair_num, void* ws, size_t ws_bytes,
                                            cudaStream_t stream) {
    if (!nbr || K <= 0 || n_out < 0 || !pairs_in || !pairs_out || !pair_num) return CRB3D_ERR_ARG;
    int64_t t = (int64_t)K * n_out;
    if (t == 0) { CRB3D_CUDA(cudaMemsetAsync(pair_num, 0, sizeof(int) * K, stream)); return CRB3D_OK; }
    WsCursor c(ws, ws_bytes);
    int* rank = c.take<int>(t);
    int* scan_ws = c.take<int>(crb3d_scan_ws_ints(t));
    int* total = c.take<int>(1);
    if (!c.ok) return CRB3D_ERR_WORKSPACE;
    pair_flags<<<(unsigned)crb3d_divup(t, 256), 256, 0, stream>>>(nbr, t, rank);
    int rc = crb3d_scan_exclusive_i32(rank, rank, t, scan_ws, total, stream);
    if (rc) return rc;
    pair_write<<<dim3((unsigned)crb3d_divup(n_out, 256), K), 256, 0, stream>>>(nbr, K, n_out, rank, pair_cap, pairs_in,
                                                                              pairs_out);
    pair_counts<<<(unsigned)crb3d_divup(K, 64), 64, 0, stream>>>(rank, total, K, n_out, pair_num);
    CRB3D_CHECK_LAUNCH();
    return CRB3D_OK;
}

CRB3D_DIAG_DEFINE_SETTER(rulebook)
