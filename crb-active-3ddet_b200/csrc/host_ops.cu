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
