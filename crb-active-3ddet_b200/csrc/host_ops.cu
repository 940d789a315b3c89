// Host-side entry points of the C ABI: version/error strings and the two CPU ops the reference exports next to
// its CUDA ops (used by dataset preparation / gt-sampling augmentation inside DataLoader workers):
//   pcdet/ops/iou3d_nms/src/iou3d_cpu.cpp:232-252           boxes_iou_bev_cpu
//   pcdet/ops/roiaware_pool3d/src/roiaware_pool3d.cpp:119-168 points_in_boxes_cpu  (MARGIN = 1e-2, not the GPU's 1e-5)
#include "common.cuh"
#include "rbox.cuh"

extern "C" const char* crb3d_version(void) { return "crb3d-b200 0.1.0 (sm_100a)"; }

extern "C" const char* crb3d_strerror(int code) {
    switch (code) {
        case CRB3D_OK: return "ok";
        case CRB3D_ERR_ARG: return "invalid argument";
        case CRB3D_ERR_CUDA: return "CUDA call or kernel launch failed";
        case CRB3D_ERR_WORKSPACE: return "workspace missing or too small";
        case CRB3D_ERR_UNSUPPORTED: return "shape not supported by this kernel";
        case CRB3D_ERR_DEVICE: return "a kernel exceeded a bounded wait / probe (crb3d_last_device_error has the record)";
        default: return "unknown error";
    }
}

extern "C" int crb3d_boxes_iou_bev_cpu(const float* boxes_a, int na, const float* boxes_b, int nb, float* out) {
    if (na < 0 || nb < 0 || (na > 0 && nb > 0 && (!boxes_a || !boxes_b || !out))) return CRB3D_ERR_ARG;
    for (int j = 0; j < nb; ++j) {
        RBox B;
        make_rbox(boxes_b + (size_t)j * 7, B);
        for (int i = 0; i < na; ++i) {
            RBox A;
            make_rbox(boxes_a + (size_t)i * 7, A);
            out[(size_t)i * nb + j] = rbox_iou(A, B);
        }
    }
    return CRB3D_OK;
}

extern "C" int crb3d_points_in_boxes_cpu(const float* boxes, int n_boxes, const float* pts, int n_pts, int* out) {
    if (n_boxes < 0 || n_pts < 0 || (n_boxes > 0 && n_pts > 0 && (!boxes || !pts || !out))) return CRB3D_ERR_ARG;
    const float margin = 1e-2f;
    for (int i = 0; i < n_boxes; ++i) {
        const float* b = boxes + (size_t)i * 7;
        const float c = cosf(-b[6]), s = sinf(-b[6]);
        const double tx = (double)b[3] / 2.0 + (double)margin, ty = (double)b[4] / 2.0 + (double)margin;
        for (int j = 0; j < n_pts; ++j) {
            const float* p = pts + (size_t)j * 3;
            int flag = 0;
            if (!((double)fabsf(p[2] - b[2]) > (double)b[5] / 2.0)) {
                const float sx = p[0] - b[0], sy = p[1] - b[1];
                const float lx = sx * c + sy * (-s), ly = sx * s + sy * c;
                flag = ((double)fabsf(lx) < tx) & ((double)fabsf(ly) < ty);
            }
            out[(size_t)i * n_pts + j] = flag;
        }
    }
    return CRB3D_OK;
}

// Host-side hard voxelizer with the Point2VoxelCPU3d contract (runs inside DataLoader worker processes, where no CUDA
// context exists): pcdet/datasets/processor/data_processor.py:25,36-42,54-59. HOST pointers. One frame.
// voxels (max_voxels,max_pts,n_feat) zero-filled here, coords (max_voxels,3) zyx, num (max_voxels). *n_voxels = count.
#include <vector>
#include <cstring>
#include <cmath>
extern "C" int crb3d_point_to_voxel_cpu(const float* pts, int64_t n, int stride, int n_feat, const float* range6,
                                        const float* vsize3, const int* grid3, int max_pts, int max_voxels,
                                        float* voxels, int* coords, int* num, int* n_voxels) {
    if (n < 0 || stride < 3 || n_feat <= 0 || n_feat > stride || max_pts <= 0 || max_voxels <= 0 || !range6 || !vsize3 ||
        !grid3 || !voxels || !coords || !num || !n_voxels)
        return CRB3D_ERR_ARG;
    size_t cap = 16;
    while (cap < (size_t)n * 2 + 2) cap <<= 1;
    std::vector<int64_t> keys(cap, -1);
    std::vector<int> vals(cap, -1);
    std::memset(voxels, 0, sizeof(float) * (size_t)max_voxels * max_pts * n_feat);
    std::memset(num, 0, sizeof(int) * (size_t)max_voxels);
    int count = 0;
    for (int64_t i = 0; i < n; ++i) {
        const float* p = pts + i * stride;
        int c[3];
        bool ok = true;
        for (int j = 0; j < 3 && ok; ++j) {
            const float f = std::floor((p[j] - range6[j]) / vsize3[j]);
            ok = (f >= 0.0f) && (f < (float)grid3[j]);
            c[j] = ok ? (int)f : 0;
        }
        if (!ok) continue;
        const int64_t key = ((int64_t)c[2] * grid3[1] + c[1]) * grid3[0] + c[0];
        size_t s = (size_t)(((uint64_t)key * 0x9E3779B97F4A7C15ull) >> 17) & (cap - 1);
        int vid = -1;
        while (keys[s] != -1) {
            if (keys[s] == key) { vid = vals[s]; break; }
            s = (s + 1) & (cap - 1);
        }
        if (vid < 0) {
            if (count >= max_voxels) continue;
            vid = count++;
            keys[s] = key; vals[s] = vid;
            coords[vid * 3] = c[2]; coords[vid * 3 + 1] = c[1]; coords[vid * 3 + 2] = c[0];
        }
        if (num[vid] < max_pts) {
            std::memcpy(voxels + ((size_t)vid * max_pts + num[vid]) * n_feat, p, sizeof(float) * n_feat);
            ++num[vid];
        }
    }
    *n_voxels = count;
    return CRB3D_OK;
}
