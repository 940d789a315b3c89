// extern "C" doors onto the REFERENCE's pointnet2_batch, voxel_query and roipoint_pool3d launchers (compiled from
// /root/reference by oracle/build.py into oracle/_ref/libpcdet_ref_kernels_batch.so - a second library because the batch
// and stack sampling_gpu.cu define the same launcher name). No reference source is copied into this repo.
// TEST INFRASTRUCTURE ONLY.
#include <cuda_runtime.h>

// prototypes as declared by the reference (pointnet2_batch/src/*_gpu.h, pointnet2_stack/src/voxel_query_gpu.h:14,
// roipoint_pool3d/src/roipoint_pool3d.cpp:19)
void ball_query_kernel_launcher_fast(int b, int n, int m, float radius, int nsample, const float* new_xyz, const float* xyz, int* idx);
void group_points_kernel_launcher_fast(int b, int c, int n, int npoints, int nsample, const float* points, const int* idx, float* out);
void group_points_grad_kernel_launcher_fast(int b, int c, int n, int npoints, int nsample, const float* grad_out, const int* idx,
                                            float* grad_points);
void gather_points_kernel_launcher_fast(int b, int c, int n, int npoints, const float* points, const int* idx, float* out);
void gather_points_grad_kernel_launcher_fast(int b, int c, int n, int npoints, const float* grad_out, const int* idx, float* grad_points);
void farthest_point_sampling_kernel_launcher(int b, int n, int m, const float* dataset, float* temp, int* idxs);
void three_nn_kernel_launcher_fast(int b, int n, int m, const float* unknown, const float* known, float* dist2, int* idx);
void three_interpolate_kernel_launcher_fast(int b, int c, int m, int n, const float* points, const int* idx, const float* weight, float* out);
void three_interpolate_grad_kernel_launcher_fast(int b, int c, int n, int m, const float* grad_out, const int* idx, const float* weight,
                                                 float* grad_points);
void voxel_query_kernel_launcher_stack(int M, int R1, int R2, int R3, int nsample, float radius, int z_range, int y_range, int x_range,
                                       const float* new_xyz, const float* xyz, const int* new_coords, const int* point_indices, int* idx);
void roipool3dLauncher(int batch_size, int pts_num, int boxes_num, int feature_in_len, int sampled_pts_num, const float* xyz,
                       const float* boxes3d, const float* pts_feature, float* pooled_features, int* pooled_empty_flag);

// pointnet2_stack/src/vector_pool_gpu.h
int query_stacked_local_neighbor_idxs_kernel_launcher_stack(const float* support_xyz, const int* xyz_batch_cnt, const float* new_xyz,
                                                            const int* new_xyz_batch_cnt, int* stack_neighbor_idxs, int* start_len, int* cumsum,
                                                            int avg_length_of_neighbor_idxs, float max_neighbour_distance, int batch_size, int M,
                                                            int nsample, int neighbor_type);
int query_three_nn_by_stacked_local_idxs_kernel_launcher_stack(const float* support_xyz, const float* new_xyz, const float* new_xyz_grid_centers,
                                                               int* new_xyz_grid_idxs, float* new_xyz_grid_dist2, const int* stack_neighbor_idxs,
                                                               const int* start_len, int M, int num_total_grids);
int vector_pool_kernel_launcher_stack(const float* support_xyz, const float* support_features, const int* xyz_batch_cnt, const float* new_xyz,
                                      float* new_features, float* new_local_xyz, const int* new_xyz_batch_cnt, int* point_cnt_of_grid,
                                      int* grouped_idxs, int num_grid_x, int num_grid_y, int num_grid_z, float max_neighbour_distance,
                                      int batch_size, int N, int M, int num_c_in, int num_c_out, int num_total_grids, int use_xyz,
                                      int num_max_sum_points, int nsample, int neighbor_type, int pooling_type);
void vector_pool_grad_kernel_launcher_stack(const float* grad_new_features, const int* point_cnt_of_grid, const int* grouped_idxs,
                                            float* grad_support_features, int N, int M, int num_c_out, int num_c_in, int num_total_grids,
                                            int num_max_sum_points);

extern "C" {
int refb_local_neighbors(const float* sxyz, const int* xc, const float* nxyz, const int* nc, int* stack, int* start_len, int* cumsum, int avg,
                         float dist, int B, int M, int nsample, int type) {
    return query_stacked_local_neighbor_idxs_kernel_launcher_stack(sxyz, xc, nxyz, nc, stack, start_len, cumsum, avg, dist, B, M, nsample, type);
}
int refb_three_nn_local(const float* sxyz, const float* nxyz, const float* centers, int* idxs, float* d2, const int* stack, const int* start_len,
                        int M, int G) {
    return query_three_nn_by_stacked_local_idxs_kernel_launcher_stack(sxyz, nxyz, centers, idxs, d2, stack, start_len, M, G);
}
int refb_vector_pool(const float* sxyz, const float* sfeat, const int* xc, const float* nxyz, float* nfeat, float* nlocal, const int* nc,
                     int* pcnt, int* grouped, int gx, int gy, int gz, float dist, int B, int N, int M, int cin, int cout, int G, int use_xyz,
                     int max_sum, int nsample, int type, int pooling) {
    return vector_pool_kernel_launcher_stack(sxyz, sfeat, xc, nxyz, nfeat, nlocal, nc, pcnt, grouped, gx, gy, gz, dist, B, N, M, cin, cout, G,
                                             use_xyz, max_sum, nsample, type, pooling);
}
void refb_vector_pool_grad(const float* g, const int* pcnt, const int* grouped, float* gs, int N, int M, int cout, int cin, int G, int n) {
    vector_pool_grad_kernel_launcher_stack(g, pcnt, grouped, gs, N, M, cout, cin, G, n);
}
int refb_sync() { return (int)cudaDeviceSynchronize(); }
void refb_ball_query(int b, int n, int m, float radius, int nsample, const float* new_xyz, const float* xyz, int* idx) {
    ball_query_kernel_launcher_fast(b, n, m, radius, nsample, new_xyz, xyz, idx);
}
void refb_group_points(int b, int c, int n, int np, int ns, const float* p, const int* idx, float* out) {
    group_points_kernel_launcher_fast(b, c, n, np, ns, p, idx, out);
}
void refb_group_points_grad(int b, int c, int n, int np, int ns, const float* go, const int* idx, float* gp) {
    group_points_grad_kernel_launcher_fast(b, c, n, np, ns, go, idx, gp);
}
void refb_gather_points(int b, int c, int n, int np, const float* p, const int* idx, float* out) {
    gather_points_kernel_launcher_fast(b, c, n, np, p, idx, out);
}
void refb_gather_points_grad(int b, int c, int n, int np, const float* go, const int* idx, float* gp) {
    gather_points_grad_kernel_launcher_fast(b, c, n, np, go, idx, gp);
}
void refb_fps(int b, int n, int m, const float* d, float* temp, int* idx) { farthest_point_sampling_kernel_launcher(b, n, m, d, temp, idx); }
void refb_three_nn(int b, int n, int m, const float* u, const float* k, float* d2, int* idx) {
    three_nn_kernel_launcher_fast(b, n, m, u, k, d2, idx);
}
void refb_three_interpolate(int b, int c, int m, int n, const float* p, const int* idx, const float* w, float* out) {
    three_interpolate_kernel_launcher_fast(b, c, m, n, p, idx, w, out);
}
void refb_three_interpolate_grad(int b, int c, int n, int m, const float* go, const int* idx, const float* w, float* gp) {
    three_interpolate_grad_kernel_launcher_fast(b, c, n, m, go, idx, w, gp);
}
void refb_voxel_query(int M, int R1, int R2, int R3, int nsample, float radius, int zr, int yr, int xr, const float* new_xyz,
                      const float* xyz, const int* new_coords, const int* point_indices, int* idx) {
    voxel_query_kernel_launcher_stack(M, R1, R2, R3, nsample, radius, zr, yr, xr, new_xyz, xyz, new_coords, point_indices, idx);
}
void refb_roipoint_pool3d(int B, int N, int M, int C, int S, const float* xyz, const float* boxes, const float* feat, float* pooled,
                          int* empty_flag) {
    roipool3dLauncher(B, N, M, C, S, xyz, boxes, feat, pooled, empty_flag);
}
}
