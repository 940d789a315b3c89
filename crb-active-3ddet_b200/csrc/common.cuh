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

#define CRB3D_ERR_DEVICE (-5)   // a kernel gave up on a bounded wait / probe (see crb3d_last_device_error)

#define CRB3D_NUM_SMS 148  // B200: 2 dies x 74 SMs (default; launchers size grids with crb3d_num_sms())
#define CRB3D_MAX_DEVICES 16

// SM count of the current device (cudaDeviceGetAttribute, cached per device) and the current device's index
int crb3d_num_sms();
int crb3d_current_device();

// ---------------------------------------------------------------------------------------------
// Device-side diagnostics: every wait / probe loop in this library is BOUNDED. A kernel that exceeds its budget claims
// the process-wide record (zero-copy pinned host memory, so it stays readable after the context dies), fills it in and
// traps: a hang becomes a launch failure at the next synchronisation and crb3d_last_device_error() says where.
// Without relocatable device code every translation unit owns a copy of the pointer; csrc/diag.cu sets all of them.
// ---------------------------------------------------------------------------------------------
struct Crb3dDiagRec {
    unsigned int flag;      // 0 = clear, 1 = being written, 2 = complete
    unsigned int kernel;    // CRB3D_K_* id
    unsigned int site;      // barrier / loop id inside the kernel
    unsigned int parity;    // mbarrier parity waited for (or probe count)
    unsigned int block_x, block_y, thread;
    unsigned int extra;     // iteration / stage / whatever the site documents
    unsigned long long waited_ns;
    unsigned int device;
    unsigned int pad;
};

enum {
    CRB3D_K_SPCONV_TC = 1, CRB3D_K_BEV_CONV = 2, CRB3D_K_BEV_CONV_PAIR = 3, CRB3D_K_BEV_GEMM = 4, CRB3D_K_HASH_INSERT = 5,
    CRB3D_K_HASH_FIND = 6, CRB3D_K_BEV_CONV_S2 = 7, CRB3D_K_FC_GEMM = 8
};

#define CRB3D_WAIT_BUDGET_NS 4000000000ull   // 4 s: three orders of magnitude above the longest kernel of this library

#ifdef __CUDACC__
static __device__ Crb3dDiagRec* g_crb3d_diag = nullptr;   // per translation unit (set by CRB3D_DIAG_DEFINE_SETTER's function)
static __device__ unsigned int g_crb3d_diag_device = 0;

#define CRB3D_DIAG_DEFINE_SETTER(name)                                                                        \
    extern "C" int crb3d_diag_set_##name(void* host_mapped, unsigned int device) {                            \
        Crb3dDiagRec* p = (Crb3dDiagRec*)host_mapped;                                                         \
        if (cudaMemcpyToSymbol(g_crb3d_diag, &p, sizeof(p)) != cudaSuccess) return CRB3D_ERR_CUDA;             \
        if (cudaMemcpyToSymbol(g_crb3d_diag_device, &device, sizeof(device)) != cudaSuccess) return CRB3D_ERR_CUDA; \
        return CRB3D_OK;                                                                                       \
    }

__device__ __forceinline__ unsigned long long crb3d_globaltimer() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// cold path: record + trap (never returns)
static __device__ __noinline__ void crb3d_diag_fail(unsigned int kernel, unsigned int site, unsigned int parity, unsigned int extra,
                                             unsigned long long waited_ns) {
    Crb3dDiagRec* r = g_crb3d_diag;
    if (r && atomicCAS_system(&r->flag, 0u, 1u) == 0u) {
        r->kernel = kernel; r->site = site; r->parity = parity; r->extra = extra;
        r->block_x = blockIdx.x; r->block_y = blockIdx.y; r->thread = threadIdx.x;
        r->waited_ns = waited_ns; r->device = g_crb3d_diag_device;
        __threadfence_system();
        r->flag = 2u;
        __threadfence_system();
    }
    __trap();
}
#endif

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
// The probe sequence is bounded by the capacity: a full table (a caller sized it wrong) is a device error, not a hang.
__device__ __forceinline__ uint32_t hash_insert(unsigned long long* keys, uint32_t cap_mask, unsigned long long key) {
    uint32_t s = hash_u64(key) & cap_mask;
    for (uint32_t probes = 0; probes <= cap_mask; ++probes) {
        unsigned long long cur = keys[s];
        if (cur == key) return s;
        if (cur == CRB3D_EMPTY_KEY) {
            unsigned long long prev = atomicCAS(&keys[s], CRB3D_EMPTY_KEY, key);
            if (prev == CRB3D_EMPTY_KEY || prev == key) return s;
        }
        s = (s + 1) & cap_mask;
    }
    crb3d_diag_fail(CRB3D_K_HASH_INSERT, 0, cap_mask, (unsigned int)key, 0);
    return 0;
}

// Returns slot or 0xFFFFFFFF when absent.
__device__ __forceinline__ uint32_t hash_find(const unsigned long long* __restrict__ keys, uint32_t cap_mask,
                                              unsigned long long key) {
    uint32_t s = hash_u64(key) & cap_mask;
    for (uint32_t probes = 0; probes <= cap_mask; ++probes) {
        unsigned long long cur = __ldg(&keys[s]);
        if (cur == key) return s;
        if (cur == CRB3D_EMPTY_KEY) return 0xFFFFFFFFu;
        s = (s + 1) & cap_mask;
    }
    return 0xFFFFFFFFu;   // a table without an empty slot and without the key: absent
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
int crb3d_scan_block_sums(int* sums, int64_t nb, int* total, cudaStream_t stream);
int crb3d_fill_i32(int* p, size_t n, int v, cudaStream_t stream);
int crb3d_fill_f32(float* p, size_t n, float v, cudaStream_t stream);
