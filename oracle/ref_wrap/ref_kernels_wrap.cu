// extern "C" doors onto the REFERENCE's own CUDA launchers (compiled from /root/reference by oracle/build.py into
// oracle/_ref/libpcdet_ref_kernels.so; no reference source is copied into this repo). Used by the -m gpu parity tests
// as a second oracle: the unmodified reference kernels run on the same B200 next to ours.
// TEST INFRASTRUCTURE ONLY.
#include <cuda_runtime.h>

// prototypes as declared by the reference (iou3d_nms.cpp:30-34, roiaware_pool3d.cpp:22-32, *_gpu.h)
void boxesoverlapLauncher(const int num_a, const float* boxes_a, const int num_b, const float* boxes_b, float* ans_overlap);
void boxesioubevLauncher(const int num_a, const float* boxes_a, const int num_b, const float* boxes_b, float* ans_iou);
void nmsLauncher(const float* boxes, unsigned long long* mask, int boxes_num, float nms_overlap_thresh);
void nmsNormalLauncher(const float* boxes, unsigned long long* mask, int boxes_num, float nms_overlap_thresh);
void roiaware_pool3d_launcher(int boxes_num, int pts_num, int channels, int max_pts_each_voxel, int out_x, int out_y,
                              int out_z, const float* rois, const float* pts, const float* pts_feature, int* argmax,
                              int* pts_idx_of_voxels, float* pooled_features, int pool_method);
void roiaware_pool3d_backward_launcher(int boxes_num, int out_x, int out_y, int out_z, int channels,
                                       int max_pts_each_voxel, const int* pts_idx_of_voxels, const int* argmax,
                                       const float* grad_out, float* grad_in, int pool_method);
void points_in_boxes_launcher(int batch_size, int boxes_num, int pts_num, const float* boxes, const float* pts,
                              int* box_idx_of_points);
void ball_query_kernel_launcher_stack(int B, int M, float radius, int nsample, const float* new_xyz,
                                      const int* new_xyz_batch_cnt, const float* xyz, const int* xyz_batch_cnt, int* idx);
void group_points_kernel_launcher_stack(int B, int M, int C, int nsample, const float* features,
                                        const int* features_batch_cnt, const int* idx, const int* idx_batch_cnt, float* out);
void group_points_grad_kernel_launcher_stack(int B, int M, int C, int N, int nsample, const float* grad_out,
                                             const int* idx, const int* idx_batch_cnt, const int* features_batch_cnt,
                                             float* grad_features);
void farthest_point_sampling_kernel_launcher(int b, int n, int m, const float* dataset, float* temp, int* idxs);
void stack_farthest_point_sampling_kernel_launcher(int N, int batch_size, const float* dataset, float* temp,
                                                   int* xyz_batch_cnt, int* idxs, int* num_sampled_points);
void three_nn_kernel_launcher_stack(int batch_size, int N, int M, const float* unknown, const int* unknown_batch_cnt,
                                    const float* known, const int* known_batch_cnt, float* dist2, int* idx);
void three_interpolate_kernel_launcher_stack(int N, int channels, const float* features, const int* idx,
                                             const float* weight, float* out);
void three_interpolate_grad_kernel_launcher_stack(int N, int channels, const float* grad_out, const int* idx,
                                                  const float* weight, float* grad_features);

extern "C" {
int ref_sync() { return (int)cudaDeviceSynchronize(); }
void ref_boxes_overlap(int na, const float* a, int nb, const float* b, float* out) { boxesoverlapLauncher(na, a, nb, b, out); }
void ref_boxes_iou_bev(int na, const float* a, int nb, const float* b, float* out) { boxesioubevLauncher(na, a, nb, b, out); }
void ref_nms_mask(const float* boxes, unsigned long long* mask, int n, float thr) { nmsLauncher(boxes, mask, n, thr); }
void ref_nms_normal_mask(const float* boxes, unsigned long long* mask, int n, float thr) { nmsNormalLauncher(boxes, mask, n, thr); }
void ref_roiaware_pool3d(int boxes_num, int pts_num, int channels, int max_pts_each_voxel, int ox, int oy, int oz,
                         const float* rois, const float* pts, const float* feat, int* argmax, int* pidx, float* pooled, int method) {
    roiaware_pool3d_launcher(boxes_num, pts_num, channels, max_pts_each_voxel, ox, oy, oz, rois, pts, feat, argmax, pidx, pooled, method);
}
void ref_roiaware_pool3d_backward(int boxes_num, int ox, int oy, int oz, int channels, int max_pts_each_voxel,
                                  const int* pidx, const int* argmax, const float* grad_out, float* grad_in, int method) {
    roiaware_pool3d_backward_launcher(boxes_num, ox, oy, oz, channels, max_pts_each_voxel, pidx, argmax, grad_out, grad_in, method);
}
void ref_points_in_boxes(int B, int T, int M, const float* boxes, const float* pts, int* out) { points_in_boxes_launcher(B, T, M, boxes, pts, out); }
void ref_ball_query(int B, int M, float radius, int nsample, const float* new_xyz, const int* new_cnt, const float* xyz,
                    const int* xyz_cnt, int* idx) {
    ball_query_kernel_launcher_stack(B, M, radius, nsample, new_xyz, new_cnt, xyz, xyz_cnt, idx);
}
void ref_group_points(int B, int M, int C, int ns, const float* f, const int* fc, const int* idx, const int* ic, float* out) {
    group_points_kernel_launcher_stack(B, M, C, ns, f, fc, idx, ic, out);
}
void ref_group_points_grad(int B, int M, int C, int N, int ns, const float* go, const int* idx, const int* ic, const int* fc, float* gf) {
    group_points_grad_kernel_launcher_stack(B, M, C, N, ns, go, idx, ic, fc, gf);
}
void ref_fps(int b, int n, int m, const float* d, float* temp, int* idx) { farthest_point_sampling_kernel_launcher(b, n, m, d, temp, idx); }
void ref_stack_fps(int N, int B, const float* d, float* temp, int* xyz_cnt, int* idx, int* num_sampled) {
    stack_farthest_point_sampling_kernel_launcher(N, B, d, temp, xyz_cnt, idx, num_sampled);
}
void ref_three_nn(int B, int N, int M, const float* u, const int* uc, const float* k, const int* kc, float* d2, int* idx) {
    three_nn_kernel_launcher_stack(B, N, M, u, uc, k, kc, d2, idx);
}
void ref_three_interpolate(int N, int C, const float* f, const int* idx, const float* w, float* out) {
    three_interpolate_kernel_launcher_stack(N, C, f, idx, w, out);
}
void ref_three_interpolate_grad(int N, int C, const float* go, const int* idx, const float* w, float* gf) {
    three_interpolate_grad_kernel_launcher_stack(N, C, go, idx, w, gf);
}
}
