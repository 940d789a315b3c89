// Shared device/host helpers for libcrb3d_sm100 (sm_100a only).
// Everything here is internal; the public surface is include/crb3d.h.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>

#define CRB3D_OK 0
#define CRB3D_ERR_ARG (-1)
#define CRB3D_ERR_CUDA (-2)
#define CRB3D_ERR_WORKSPACE (-3)
#define CRB3D_ERR_UNSUPPORTED (-4)

#define CRB3D_NUM_SMS 148  // B200: 2 dies x 74 SMs

#define CRB3D_CHECK_LAUNCH()                                   \
    do {                                                       \
        cudaError_t e__ = cudaGetLastError();                  \
        if (e__ != cudaSuccess) return CRB3D_ERR_CUDA;         \
    } while (0)

#define CRB3D_CUDA(call)                                       \
    do {                                                       \
        cudaError_t e__ = (call);                              \
        if (e__ != cudaSuccess) return CRB3D_ERR_CUDA;         \
    } while (0)

static inline int64_t crb3d_divup(int64_t a, int64_t b) { return (a + b - 1) / b; }
static inline size_t crb3d_align(size_t x, size_t a = 256) { return (x + a - 1) / a * a; }

// Bump allocator over a caller-provided workspace (no cudaMalloc inside the library).
struct WsCursor {
    char* base;
    size_t off;
    size_t cap;
    bool ok;
    WsCursor(void* p, size_t bytes) : base((char*)p), off(0), cap(bytes), ok(true) {}
    template <typename T>
    T* take(size_t n) {
        size_t need = crb3d_align(n * sizeof(T));
        if (base == nullptr || off + need > cap) { ok = false; off += need; return nullptr; }
        T* r = (T*)(base + off);
        off += need;
        return r;
    }
};

static inline uint32_t crb3d_next_pow2(uint64_t v) {
    uint64_t p = 1;
    while (p < v) p <<= 1;
    return (uint32_t)p;
}

// ---------------------------------------------------------------------------------------------
// 64-bit open-addressing hash set/map helpers (linear probing, power-of-two capacity).
// ---------------------------------------------------------------------------------------------
#define CRB3D_EMPTY_KEY 0xFFFFFFFFFFFFFFFFull

__device__ __forceinline__ uint32_t hash_u64(unsigned long long k) {
    // murmur3 fmix64
    k ^= k >> 33;
    k *= 0xff51afd7ed558ccdull;
    k ^= k >> 33;
    k *= 0xc4ceb9fe1a85ec53ull;
    k ^= k >> 33;
    return (uint32_t)k;
}

// Returns the slot that holds `key`, inserting it if absent.
__device__ __forceinline__ uint32_t hash_insert(unsigned long long* keys, uint32_t cap_mask, unsigned long long key) {
    uint32_t s = hash_u64(key) & cap_mask;
    while (true) {
        unsigned long long cur = keys[s];
        if (cur == key) return s;
        if (cur == CRB3D_EMPTY_KEY) {
            unsigned long long prev = atomicCAS(&keys[s], CRB3D_EMPTY_KEY, key);
            if (prev == CRB3D_EMPTY_KEY || prev == key) return s;
        }
        s = (s + 1) & cap_mask;
    }
}

// Returns slot or 0xFFFFFFFF when absent.
__device__ __forceinline__ uint32_t hash_find(const unsigned long long* __restrict__ keys, uint32_t cap_mask,
                                              unsigned long long key) {
    uint32_t s = hash_u64(key) & cap_mask;
    while (true) {
        unsigned long long cur = __ldg(&keys[s]);
        if (cur == key) return s;
        if (cur == CRB3D_EMPTY_KEY) return 0xFFFFFFFFu;
        s = (s + 1) & cap_mask;
    }
}

// ---------------------------------------------------------------------------------------------
// Warp / block primitives
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int warp_incl_scan(int v) {
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int n = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += n;
    }
    return v;
}

// Block-wide exclusive scan for blockDim.x <= 1024 (multiple of 32). `smem` needs 33 ints.
// Returns exclusive prefix of v; *total receives the block sum.
__device__ __forceinline__ int block_excl_scan(int v, int* smem, int* total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    int incl = warp_incl_scan(v);
    if (lane == 31) smem[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        int w = (lane < nwarp) ? smem[lane] : 0;
        int wi = warp_incl_scan(w);
        smem[lane] = wi - w;
        if (lane == 31) smem[32] = wi;
    }
    __syncthreads();
    int res = incl - v + smem[warp];
    *total = smem[32];
    __syncthreads();
    return res;
}

// Squared distance with the contraction the reference kernels compile to (nvcc -fmad default, verified in SASS of
// ball_query_gpu.cu / sampling_gpu.cu / interpolate_gpu.cu): d = fma(dz,dz, fma(dx,dx, dy*dy)). Index outputs depend on it.
__device__ __forceinline__ float sqdist3(float ax, float ay, float az, float bx, float by, float bz) {
    const float dx = __fsub_rn(ax, bx), dy = __fsub_rn(ay, by), dz = __fsub_rn(az, bz);
    return __fmaf_rn(dz, dz, __fmaf_rn(dx, dx, __fmul_rn(dy, dy)));
}

// Device-wide exclusive scan of int32 (three launches; n up to 2^31).
// out may alias in. block_sums needs divup(n, SCAN_TILE) + 1 ints. total (device int) may be null.
#define CRB3D_SCAN_TILE 2048
int crb3d_scan_exclusive_i32(const int* in, int* out, int64_t n, int* block_sums, int* total, cudaStream_t stream);
size_t crb3d_scan_ws_ints(int64_t n);
int crb3d_fill_i32(int* p, size_t n, int v, cudaStream_t stream);
int crb3d_fill_f32(float* p, size_t n, float v, cudaStream_t stream);
