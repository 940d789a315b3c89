// Device-side diagnostics plumbing + per-device properties (see common.cuh: Crb3dDiagRec, crb3d_diag_fail).
//
// Every bounded wait / probe of this library reports into ONE zero-copy pinned host record per process and traps; the host
// reads it with crb3d_last_device_error() even after the context died. No reference counterpart: the reference's native
// code has no device-side waits (its errors are fprintf + exit(-1), pcdet/ops/iou3d_nms/src/iou3d_nms.cpp:14-26).
#include "common.cuh"
#include <string.h>

extern "C" int crb3d_diag_set_spconv_tc(void*, unsigned int);
extern "C" int crb3d_diag_set_spconv_grp(void*, unsigned int);
extern "C" int crb3d_diag_set_bev_conv(void*, unsigned int);
extern "C" int crb3d_diag_set_bev_gemm(void*, unsigned int);
extern "C" int crb3d_diag_set_bev_gemm_pair(void*, unsigned int);
extern "C" int crb3d_diag_set_rulebook(void*, unsigned int);
extern "C" int crb3d_diag_set_voxelize(void*, unsigned int);
extern "C" int crb3d_diag_set_fc_gemm(void*, unsigned int);

namespace {
// debug markers (tools/stress_hang.py): 64 ints of zero-copy host memory behind the record; a one-thread kernel stores a
// stage id into its slot, so that after a hang the host can read how far every concurrently replayed graph copy got
__global__ void mark_kernel(volatile int* p, int v) { *p = v; __threadfence_system(); }
constexpr int N_MARKERS = 64;
Crb3dDiagRec* g_host_rec = nullptr;          // cudaHostAllocMapped | Portable: one per process
bool g_dev_init[CRB3D_MAX_DEVICES] = {};
int g_sms[CRB3D_MAX_DEVICES] = {};
}  // namespace

int crb3d_current_device() {
    int d = 0;
    if (cudaGetDevice(&d) != cudaSuccess || d < 0 || d >= CRB3D_MAX_DEVICES) return 0;
    return d;
}

int crb3d_num_sms() {
    const int d = crb3d_current_device();
    if (g_sms[d] == 0) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, d) != cudaSuccess || n <= 0) n = CRB3D_NUM_SMS;
        g_sms[d] = n;
    }
    return g_sms[d];
}

// Call once per device before the first kernel of this library on it (the Python front end does so lazily; it must not
// run while a stream of the device is capturing). Idempotent.
extern "C" int crb3d_diag_init(void) {
    const int d = crb3d_current_device();
    if (g_dev_init[d]) return CRB3D_OK;
    if (!g_host_rec) {
        void* p = nullptr;
        CRB3D_CUDA(cudaHostAlloc(&p, sizeof(Crb3dDiagRec) + sizeof(int) * N_MARKERS, cudaHostAllocMapped | cudaHostAllocPortable));
        memset(p, 0, sizeof(Crb3dDiagRec) + sizeof(int) * N_MARKERS);
        g_host_rec = (Crb3dDiagRec*)p;
    }
    void* dptr = nullptr;
    CRB3D_CUDA(cudaHostGetDevicePointer(&dptr, g_host_rec, 0));
    int rc;
    if ((rc = crb3d_diag_set_spconv_tc(dptr, (unsigned)d))) return rc;
    if ((rc = crb3d_diag_set_spconv_grp(dptr, (unsigned)d))) return rc;
    if ((rc = crb3d_diag_set_bev_conv(dptr, (unsigned)d))) return rc;
    if ((rc = crb3d_diag_set_bev_gemm(dptr, (unsigned)d))) return rc;
    if ((rc = crb3d_diag_set_bev_gemm_pair(dptr, (unsigned)d))) return rc;
    if ((rc = crb3d_diag_set_rulebook(dptr, (unsigned)d))) return rc;
    if ((rc = crb3d_diag_set_voxelize(dptr, (unsigned)d))) return rc;
    if ((rc = crb3d_diag_set_fc_gemm(dptr, (unsigned)d))) return rc;
    g_dev_init[d] = true;
    return CRB3D_OK;
}

// HOST out[12] (uint32): flag, kernel, site, parity, block_x, block_y, thread, extra, waited_ns lo, hi, device, 0.
// Returns CRB3D_OK when no kernel reported, CRB3D_ERR_DEVICE when out[] holds a record. Reads host memory only, so it
// works after a trap killed the context.
extern "C" int crb3d_last_device_error(unsigned int* out) {
    if (!out) return CRB3D_ERR_ARG;
    memset(out, 0, sizeof(unsigned int) * 12);
    if (!g_host_rec) return CRB3D_OK;
    volatile Crb3dDiagRec* r = g_host_rec;
    if (r->flag == 0) return CRB3D_OK;
    out[0] = r->flag; out[1] = r->kernel; out[2] = r->site; out[3] = r->parity; out[4] = r->block_x; out[5] = r->block_y;
    out[6] = r->thread; out[7] = r->extra; out[8] = (unsigned int)(r->waited_ns & 0xFFFFFFFFull);
    out[9] = (unsigned int)(r->waited_ns >> 32); out[10] = r->device;
    return CRB3D_ERR_DEVICE;
}

extern "C" int crb3d_diag_clear(void) {
    if (g_host_rec) memset(g_host_rec, 0, sizeof(Crb3dDiagRec));
    return CRB3D_OK;
}

extern "C" int crb3d_device_sm_count(int* n) {
    if (!n) return CRB3D_ERR_ARG;
    *n = crb3d_num_sms();
    return CRB3D_OK;
}

// Debug (not in include/crb3d.h): queue a one-thread kernel on `stream` that stores `value` into host-visible marker `slot`.
extern "C" int crb3d_debug_mark(int slot, int value, cudaStream_t stream) {
    if (!g_host_rec || slot < 0 || slot >= N_MARKERS) return CRB3D_ERR_ARG;
    void* dptr = nullptr;
    CRB3D_CUDA(cudaHostGetDevicePointer(&dptr, g_host_rec, 0));
    mark_kernel<<<1, 1, 0, stream>>>((volatile int*)((char*)dptr + sizeof(Crb3dDiagRec)) + slot, value);
    CRB3D_CHECK_LAUNCH();
    return CRB3D_OK;
}
extern "C" int crb3d_debug_read_markers(int* out, int n) {
    if (!g_host_rec || !out || n < 0 || n > N_MARKERS) return CRB3D_ERR_ARG;
    const volatile int* m = (const volatile int*)((const char*)g_host_rec + sizeof(Crb3dDiagRec));
    for (int i = 0; i < n; ++i) out[i] = m[i];
    return CRB3D_OK;
}
