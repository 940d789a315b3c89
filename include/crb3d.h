/*
 * libcrb3d_sm100 - C ABI of the B200-native (sm_100a) hot path of CRB-active-3Ddet.
 *
 * Every entry point takes plain device pointers + sizes + a cudaStream_t, returns 0 on success or a negative
 * CRB3D_ERR_* code, never allocates (scratch comes in through `ws`, sized by the matching *_workspace_bytes call),
 * never synchronises the host and keeps no global state. All pointers are DEVICE pointers unless a parameter is
 * documented as "host". Citations name the reference interface (under /root/reference) each call replaces.
 */
#ifndef CRB3D_H_
#define CRB3D_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#ifndef __CUDA_RUNTIME_H__
typedef struct CUstream_st* cudaStream_t;
#endif

#define CRB3D_OK 0
#define CRB3D_ERR_ARG (-1)         /* invalid argument */
#define CRB3D_ERR_CUDA (-2)        /* a CUDA call / launch failed */
#define CRB3D_ERR_WORKSPACE (-3)   /* workspace missing or too small */
#define CRB3D_ERR_UNSUPPORTED (-4) /* shape outside what the kernels cover */
#define CRB3D_ERR_DEVICE (-5)      /* a kernel exceeded a bounded wait / probe: see crb3d_last_device_error */

const char* crb3d_version(void);
const char* crb3d_strerror(int code);

/* ---- device-side diagnostics -----------------------------------------------------------------------------------
 * No reference counterpart (the reference's native code never waits on the device; its host errors are
 * fprintf + exit(-1), pcdet/ops/iou3d_nms/src/iou3d_nms.cpp:14-26). Every mbarrier wait and hash probe of this library
 * is bounded (4 s / table capacity). A kernel that runs out of budget writes {kernel id, site, parity, block, thread,
 * waited ns} into one zero-copy pinned HOST record and traps, so a lost arrival surfaces as a launch failure with a
 * location at the next synchronisation instead of a GPU that spins until a watchdog kills the process.
 * crb3d_diag_init: once per device (current device), before the first kernel and outside stream capture; this is the one
 *   call of the library that allocates (64 bytes of pinned host memory per process).
 * crb3d_last_device_error: HOST out[12] = {flag, kernel, site, parity, block_x, block_y, thread, extra, ns_lo, ns_hi,
 *   device, 0}; returns CRB3D_OK (nothing recorded) or CRB3D_ERR_DEVICE. Reads host memory only: valid after a trap. */
int crb3d_diag_init(void);
int crb3d_last_device_error(unsigned int* out12);
int crb3d_diag_clear(void);
int crb3d_device_sm_count(int* n);

/* ---- voxelization + MeanVFE ---------------------------------------------------------------------------------
 * replaces spconv.utils.Point2VoxelCPU3d.point_to_voxel (pcdet/datasets/processor/data_processor.py:15-60,115-143)
 * fused with MeanVFE.forward (pcdet/models/backbones_3d/vfe/mean_vfe.py:14-31).
 * points: (n_points, pt_stride) f32, xyz at column xyz_col, the n_feat feature columns start at feat_col.
 * frame_offsets: (B+1) int32 row offsets of each frame inside `points`. range6/vsize3/grid3: HOST arrays
 * (xmin,ymin,zmin,xmax,ymax,zmax), (vx,vy,vz), grid (x,y,z).
 * outputs sized for cap = min(n_points, B*max_voxels) rows: mean_feats (cap,n_feat) [nullable], voxels
 * (cap,max_pts,n_feat) zero padded [nullable], coords (cap,4) int32 (b,z,y,x), num_points (cap),
 * frame_voxel_offsets (B+1) int32 (exclusive prefix of per-frame voxel counts; [B] = total rows). */
int crb3d_voxelize_workspace_bytes(int64_t n_points, int batch_size, int max_pts, size_t* bytes);
int crb3d_voxelize(const float* points, int64_t n_points, int pt_stride, int xyz_col, int feat_col, int n_feat,
                   const int* frame_offsets, int batch_size, const float* range6, const float* vsize3,
                   const int* grid3, int max_pts, int max_voxels, float* mean_feats, float* voxels, int* coords,
                   int* num_points, int* frame_voxel_offsets, void* ws, size_t ws_bytes, cudaStream_t stream);

/* Host-side voxelizer with the same contract for DataLoader workers (no CUDA context there): HOST pointers, one frame.
 * voxels (max_voxels,max_pts,n_feat), coords (max_voxels,3) zyx, num (max_voxels); *n_voxels = number of voxels. */
int crb3d_point_to_voxel_cpu(const float* pts, int64_t n, int stride, int n_feat, const float* range6,
                             const float* vsize3, const int* grid3, int max_pts, int max_voxels, float* voxels,
                             int* coords, int* num, int* n_voxels);

/* ---- rulebook ------------------------------------------------------------------------------------------------
 * replaces spconv 2.1 get_indice_pairs behind SubMConv3d / SparseConv3d
 * (pcdet/models/backbones_3d/spconv_backbone.py:12-17,77-117). coords: (n,4) int32 (b,z,y,x), unique rows.
 * Neighbour table nbr[k][o] = input row feeding output row o through kernel offset k = (kz*KY+ky)*KX+kx, or -1.
 * shape / ksize / stride / pad / dilation arguments are HOST int[3] in (z,y,x) order. */
int crb3d_subm_rulebook_workspace_bytes(int n, size_t* bytes);
/* n_dev / n_in_dev / n_out_dev style arguments (all nullable): DEVICE-side row counts. When given, the host-side n is only
 * the capacity / row stride of the buffers, rows beyond the device count are neither read nor written, and a whole step
 * can be recorded into one CUDA graph without reading any count back to the host. */
int crb3d_subm_rulebook(const int* coords, int n, const int* n_dev, const int* spatial_shape3, const int* ksize3,
                        const int* dilation3, int* nbr, void* ws, size_t ws_bytes, cudaStream_t stream);
/* SubM table of a level that a strided rulebook produced, through that rulebook's cell -> row map (output-cell bitmap + word
 * ranks left in the workspace of crb3d_sparse_rulebook_coords; crb3d_sparse_rulebook_cellmap gives their byte offsets): two
 * loads per (z,y) line instead of a hash probe chain per neighbour. coords = the out_coords of that call, spatial_shape3 = its
 * out_shape3. Writes the same table as crb3d_subm_rulebook (spconv get_indice_pairs, subm=True). */
int crb3d_subm_rulebook_cellmap(const int* coords, int n, const int* n_dev, const int* spatial_shape3, const int* ksize3,
                                const int* dilation3, const unsigned int* cell_bitmap, const int* cell_rank, int* nbr,
                                cudaStream_t stream);
int crb3d_sparse_rulebook_cellmap(int batch_size, const int* out_shape3, size_t* bitmap_offset, size_t* rank_offset,
                                  long long* n_words);
int crb3d_conv_out_shape(const int* in_shape3, const int* ksize3, const int* stride3, const int* pad3,
                         const int* dilation3, int* out_shape3);
int crb3d_sparse_rulebook_workspace_bytes(int batch_size, const int* out_shape3, size_t* bytes);
/* phase 1: active output coords in ascending linear (b,z,y,x) order; true count -> *n_out_dev (device int). */
int crb3d_sparse_rulebook_coords(const int* coords_in, int n_in, const int* n_in_dev, int batch_size, const int* in_shape3,
                                 const int* out_shape3, const int* ksize3, const int* stride3, const int* pad3,
                                 const int* dilation3, int* coords_out, int cap_out, int* n_out_dev, void* ws,
                                 size_t ws_bytes, cudaStream_t stream);
/* phase 2 (same ws, untouched since phase 1): nbr [K][n_out], nbr_t [K][n_in] (nullable). */
int crb3d_sparse_rulebook_pairs(const int* coords_in, int n_in, const int* n_in_dev, int batch_size, const int* in_shape3,
                                const int* out_shape3, const int* ksize3, const int* stride3, const int* pad3,
                                const int* dilation3, int n_out, int* nbr, int* nbr_t, void* ws, size_t ws_bytes,
                                cudaStream_t stream);
/* spconv-format pair lists derived from a table: pairs_in/out [K][pair_cap] (caller pre-fills -1), pair_num [K]. */
int crb3d_rulebook_compact_pairs_workspace_bytes(int K, int n_out, size_t* bytes);
int crb3d_rulebook_compact_pairs(const int* nbr, int K, int n_out, int pair_cap, int* pairs_in, int* pairs_out,
                                 int* pair_num, void* ws, size_t ws_bytes, cudaStream_t stream);

/* ---- sparse convolution --------------------------------------------------------------------------------------
 * replaces spconv 2.1 SparseConvFunction / SubMConvFunction forward+backward (spconv_backbone.py:77-117;
 * backward reached from crb_sampling.py:205 and train_active_utils.py:50).
 * out[o,:] = sum_k feat[nbr[k][o],:] @ W_k ; weight element (co,k,ci) at co*w_co_stride + k*w_k_stride +
 * ci*w_ci_stride (spconv layout [C_out,K,C_in] = strides (K*cin, cin, 1)). Optional fused epilogue
 * y = relu(acc*scale[c] + shift[c]) (scale/shift nullable). kmap (nullable, device int[K]) remaps weight slices.
 * Input gradient = the same call with cin<->cout swapped, the transposed table and strides (1, cin, K*cin). */
int crb3d_spconv_forward_f32(const float* feat, const int* nbr, const float* weight, int n_out, int K, int cin,
                             int cout, int64_t w_co_stride, int64_t w_k_stride, int64_t w_ci_stride, const int* kmap,
                             const float* scale, const float* shift, int relu, float* out, const int* n_dev,
                             cudaStream_t stream);
/* tcgen05 path: TF32 inputs, fp32 accumulation in TMEM; feat (n_in, C_in) contiguous and 16-byte aligned (rows are
 * gathered with cp.async); weight contiguous [C_out,K,C_in]; C_in in {4,8,16,32,64}, C_out in {16,32,64,128}, K <= 27, else
 * CRB3D_ERR_UNSUPPORTED (use the f32 entry point). relu: bit 0 = ReLU, bit 1 = store TF32-rounded values (round to nearest:
 * a following tensor-core layer then reads exactly what was stored instead of truncating). */
int crb3d_spconv_forward_tf32(const float* feat, int n_in, const int* nbr, const float* weight, int n_out, int K, int cin,
                              int cout, const int* kmap, const float* scale, const float* shift, int relu, float* out,
                              const int* n_dev, cudaStream_t stream);
int crb3d_spconv_wgrad_workspace_bytes(int n_out, int K, int cin, int cout, size_t* bytes);
int crb3d_spconv_wgrad_f32(const float* feat, const float* dout, const int* nbr, int n_out, int K, int cin, int cout,
                           int accumulate, float* dw, void* ws, size_t ws_bytes, cudaStream_t stream);

/* ---- SparseConvTensor.dense() (height_compression.py:21-24) ------------------------------------------------
 * layout 0: (B,C,D,H,W); layout 1: (B,H,W,C*D) channels-last view of the BEV map. */
int crb3d_sparse_to_dense(const float* feat, const int* coords, int n, int C, int B, int D, int H, int W, int layout,
                          int zero_fill, float* dense, const int* n_dev, cudaStream_t stream);
int crb3d_dense_to_sparse(const float* dense, const int* coords, int n, int C, int B, int D, int H, int W, int layout,
                          float* feat, cudaStream_t stream);

/* ---- iou3d_nms (pcdet/ops/iou3d_nms/src/iou3d_nms_api.cpp:12-16, iou3d_nms.cpp:49-188) ---------------------
 * boxes: (n,7) f32 [x,y,z,dx,dy,dz,heading]. */
int crb3d_boxes_overlap_bev(const float* boxes_a, int na, const float* boxes_b, int nb, float* out, cudaStream_t stream);
int crb3d_boxes_iou_bev(const float* boxes_a, int na, const float* boxes_b, int nb, float* out, cudaStream_t stream);
int crb3d_nms_workspace_bytes(int n, size_t* bytes);
/* boxes sorted by descending score; keep: device int64[n]; num_keep: device int; max_keep <= 0 keeps all. */
int crb3d_nms(const float* boxes, int n, float thresh, int rotated, int max_keep, long long* keep, int* num_keep,
              void* ws, size_t ws_bytes, cudaStream_t stream);
/* batched: frame b owns boxes[b][0..counts[b]) of a padded (B,n_max,7) tensor; keep (B,keep_stride) int64, num_keep (B). */
int crb3d_nms_batched_workspace_bytes(int B, int n_max, size_t* bytes);
int crb3d_nms_batched(const float* boxes, const int* counts, int B, int n_max, float thresh, int rotated, int max_keep,
                      long long* keep, int keep_stride, int* num_keep, void* ws, size_t ws_bytes, cudaStream_t stream);
/* raw suppression bitmask, row-major mask[n][ceil(n/64)] like the reference nms_kernel (blocks on/above the diagonal only);
 * ws: crb3d_nms_workspace_bytes(n). */
int crb3d_nms_mask(const float* boxes, int n, float thresh, int rotated, unsigned long long* mask, void* ws, size_t ws_bytes,
                   cudaStream_t stream);
/* host-side (CPU tensors) BEV IoU, replaces boxes_iou_bev_cpu (iou3d_cpu.cpp:232-252): HOST pointers. */
int crb3d_boxes_iou_bev_cpu(const float* boxes_a, int na, const float* boxes_b, int nb, float* out);

/* ---- roiaware_pool3d (pcdet/ops/roiaware_pool3d/src/roiaware_pool3d.cpp:172-177) ---------------------------- */
int crb3d_points_in_boxes(const float* boxes, const float* pts, int B, int T, int M, int* out_idx, cudaStream_t stream);
int crb3d_points_in_boxes_stack(const float* pts, int pt_stride, const int* pt_off, int max_pts_per_frame,
                                const float* boxes, const int* box_off, int B, int total_pts, int total_boxes,
                                int* out_idx, int* counts, float* density, cudaStream_t stream);
/* explicit [begin,end) row ranges per frame (padded boxes with per-frame counts; no compaction, no sync) */
int crb3d_points_in_boxes_ranges(const float* pts, int pt_stride, const int* pt_begin, const int* pt_end,
                                 int max_pts_per_frame, const float* boxes, const int* box_begin, const int* box_end,
                                 int B, int n_box_slots, int* out_idx, int* counts, float* density, cudaStream_t stream);
/* HOST pointers; (n_boxes, n_pts) 0/1 matrix with the CPU op's MARGIN=1e-2 (roiaware_pool3d.cpp:119-168). */
int crb3d_points_in_boxes_cpu(const float* boxes, int n_boxes, const float* pts, int n_pts, int* out);
int crb3d_roiaware_pool3d_forward(const float* rois, const float* pts, const float* pts_feature, int n_boxes,
                                  int n_pts, int C, int max_pts_each_voxel, int ox, int oy, int oz, int* argmax,
                                  int* pts_idx_of_voxels, float* pooled, int pool_method, cudaStream_t stream);
int crb3d_roiaware_pool3d_backward(const int* pts_idx_of_voxels, const int* argmax, const float* grad_out,
                                   float* grad_in, int n_boxes, int ox, int oy, int oz, int C, int max_pts_each_voxel,
                                   int pool_method, cudaStream_t stream);

/* ---- pointnet2_stack (pcdet/ops/pointnet2/pointnet2_stack/src/pointnet2_api.cpp:13-30) --------------------- */
int crb3d_ball_query_stack(int B, int M, float radius, int nsample, const float* new_xyz, const int* new_xyz_batch_cnt,
                           const float* xyz, const int* xyz_batch_cnt, int* idx, int max_queries_per_frame,
                           cudaStream_t stream);
int crb3d_group_points_stack(int B, int M, int C, int nsample, const float* features, const int* features_batch_cnt,
                             const int* idx, const int* idx_batch_cnt, float* out, cudaStream_t stream);
int crb3d_group_points_grad_stack(int B, int M, int C, int N, int nsample, const float* grad_out, const int* idx,
                                  const int* idx_batch_cnt, const int* features_batch_cnt, float* grad_features,
                                  cudaStream_t stream);
int crb3d_farthest_point_sampling(int b, int n, int m, const float* dataset, float* temp, int* idx, cudaStream_t stream);
int crb3d_stack_farthest_point_sampling(int B, int n_max, const float* dataset, float* temp, const int* xyz_batch_cnt,
                                        int* idx, const int* num_sampled_points, cudaStream_t stream);
int crb3d_three_nn_stack(int B, int N, int M, const float* unknown, const int* unknown_batch_cnt, const float* known,
                         const int* known_batch_cnt, float* dist2, int* idx, cudaStream_t stream);
int crb3d_three_interpolate_stack(int N, int C, const float* features, const int* idx, const float* weight, float* out,
                                  cudaStream_t stream);
int crb3d_three_interpolate_grad_stack(int N, int C, const float* grad_out, const int* idx, const float* weight,
                                       float* grad_features, cudaStream_t stream);

/* ---- dense BEV GEMMs on tcgen05 (TF32 in, fp32 accumulate): the deblocks of BaseBEVBackbone
 *      (base_bev_backbone.py:60-78,100-108: ConvTranspose2d(k = stride) + BN + ReLU + torch.cat) and the three 1x1
 *      head convs of AnchorHeadSingle (anchor_head_single.py:18-32,41-58) with bias/ReLU/placement fused.
 *      D[m,n] = sum_k A[m,k] W[n,k]; A: (M,K) rows lda floats apart; W: contiguous [n_sub][N][K]; bias: N or null.
 *      Column segment s = [col_begin[s], +width[s]) is written to out_ptr[s] + row * row_stride[s] (1 <= n_seg <= 3).
 *      up = 0: n_sub = 1, output row = GEMM row. up = 2: n_sub = 4 (dy,dx) slices of a kernel=stride=2 transposed
 *      conv; GEMM row (b,y,x) of an in_h x in_w map lands on output pixel (b, 2y+dy, 2x+dx).
 *      Supported: K % 32 == 0, N in {80, 128, 256}, lda % 4 == 0.
 *      relu: bit 0 = ReLU, bit 1 = round the stored values to TF32 (a following tensor-core layer then reads them exactly).
 *      Bits 2-5 select kernel variants for A/B measurements and never change a result bit: 2 = thread-block clusters with TMA
 *      multicast, 3 = the single-CTA kernel where the CTA-pair kernel (N = 256, 128 < K <= 256) would run, 4 = the pair kernel with
 *      resident instead of streamed weights, 5 = the pair kernel for K <= 128 too. */
int crb3d_bev_gemm_tf32(const float* A, long long M, int K, long long lda, const float* W, int N, int n_sub,
                        const float* bias, int relu, int n_seg, float* const* out_ptr, const int* col_begin,
                        const int* width, const long long* row_stride, int up, int in_h, int in_w, cudaStream_t stream);

/* ---- 3x3 / stride 1 / pad 1 BEV convolution on tcgen05 (halo-tile implicit GEMM, TF32 in, fp32 accumulate) with
 *      the folded BatchNorm shift + ReLU fused (base_bev_backbone.py:33-50,96-99).
 *      in: (B,H,W,C_in) channels-last; wpack: [C_out/128][ky*3+kx][C_in/16][2][4][64][4] (crb3d.ops.pack_conv3x3_weight:
 *      each CTA of a cta_group::2 pair holds 64 output channels of a slice); bias: C_out or null; out: (B,H,W,C_out)
 *      channels-last. relu: bit 0 ReLU, bit 1 round the stored values to TF32, bit 8 single-CTA kernel (wpack then
 *      [C_out/128][ky*3+kx][C_in/16][4][128][4]). Supported: C_in % 16 == 0, C_out % 128 == 0. */
int crb3d_bev_conv3x3_tf32(const float* in, int B, int H, int W, int cin, const float* wpack, int cout, const float* bias,
                           int relu, float* out, cudaStream_t stream);
/* Sparse-tile mode of the halo-tile conv for a stack of 3x3 stride-1 layers fed by the dense() of a sparse tensor (block 1 of
 * BaseBEVBackbone): far from every occupied cell and from the border each layer's output is ONE constant vector, so a 128-pixel
 * tile whose neighbourhood (grown by level + 1 pixels) is empty and inside the image is filled with that constant instead of
 * going through the tensor cores. crb3d_bev_tile_plan builds, from the sparse rows (n,4) [b,z,y,x] (n_dev nullable), per level
 * the list / count / flags of the tiles to compute ([n_levels][n_tiles], n_tiles from crb3d_bev_conv3x3_num_tiles) and fill_flags
 * (constant tiles that somebody reads: the next level's computed tiles through their halo; every constant tile at the last level);
 * crb3d_bev_conv3x3_tf32_tiles runs one level: tiles tile_list[0 .. *n_active) on the CTA-pair kernel, `fill` (C_out floats, the
 * layer's constant: computed by the caller by running the layer on a constant image) in the tiles with tile_flags (= the level's
 * fill_flags row) != 0.
 * Bit-identical to crb3d_bev_conv3x3_tf32 on the whole map. */
int crb3d_bev_conv3x3_num_tiles(int B, int H, int W, int* n_tiles);
int crb3d_bev_tile_plan_workspace_bytes(int B, int H, int W, size_t* bytes);
int crb3d_bev_tile_plan(const int* coords, int n, const int* n_dev, int B, int H, int W, int n_levels, int* lists, int* counts,
                        unsigned char* flags, unsigned char* fill_flags, void* ws, size_t ws_bytes, cudaStream_t stream);
int crb3d_bev_conv3x3_tf32_tiles(const float* in, int B, int H, int W, int cin, const float* wpack, int cout, const float* bias,
                                 int relu, float* out, const int* tile_list, const int* n_active, const unsigned char* tile_flags,
                                 const float* fill, cudaStream_t stream);
/* k x k convolution (ksize in {1,3}, stride in {1,2}, zero padding) + bias + ReLU over a channels-last map as an implicit
 * GEMM whose A tiles are strided 4-D TMA boxes: the stride-2 first conv of a BEV block (base_bev_backbone.py:33-40) and
 * the fallback for 3x3 stride-1 layers the halo-tile kernel does not take - no cuDNN on the inference path.
 * in (B,H,W,C_in); w2 [C_out][ksize*ksize][C_in] (= weight.permute(0,2,3,1)); out (B,H_out,W_out,C_out).
 * Supported: C_in % 32 == 0, C_out in {128, 256}. relu bits as crb3d_bev_gemm_tf32. */
int crb3d_bev_conv_gemm_tf32(const float* in, int B, int H, int W, int cin, const float* w2, int cout, int ksize, int stride,
                             int pad, const float* bias, int relu, float* out, cudaStream_t stream);

/* ---- device-side data path upstream of the voxelizer + selection primitives of the other strategies ----
 * crb3d_mask_collate_points: mask_points_by_range (pcdet/utils/common_utils.py:60-63 via data_processor.py:78-90: keep
 * x0 <= x <= x1 and y0 <= y <= y1, z untested) + collate_batch's batch-index column (pcdet/datasets/dataset.py:173-178) for a
 * whole batch: stable compaction, out (n, 1+stride) [b, cols...], out_offsets (B+1) device-side row offsets. range4 is HOST.
 * crb3d_furthest_first: the greedy k-centre loop of pcdet/query_strategies/coreset_sampling.py:31-52 without host round trips:
 * min_dist (m) holds the start distances and is updated in place, out_idx (n_pick) int64 on the device (ties: lowest row). */
int crb3d_mask_collate_points_workspace_bytes(int64_t n, size_t* bytes);
int crb3d_mask_collate_points(const float* points, int64_t n, int stride, int xcol, const int* frame_offsets, int B,
                              const float* range4, float* out, int* out_offsets, void* ws, size_t ws_bytes, cudaStream_t stream);
int crb3d_furthest_first_workspace_bytes(int m, size_t* bytes);
int crb3d_furthest_first(const float* X, int m, int d, float* min_dist, int n_pick, long long* out_idx, void* ws, size_t ws_bytes,
                         cudaStream_t stream);

/* ---- remaining op families of pcdet/ops (csrc/extra_ops.cu) ----------------------------------------
 * crb3d_voxel_query_stack: pointnet2_stack/src/voxel_query_gpu.cu:13-98 + voxel_query.cpp:28-45 (Voxel-RCNN neighbour
 *   search). new_coords (M,4) int32 (b,z,y,x); point_indices (B,R1,R2,R3) int32 (-1: empty voxel); idx (M,nsample) int32:
 *   neighbours within `radius` (dist2 <= r2) in (dz,dy,dx) scan order, padded with the first, idx[m][0] = -1 when none.
 * pointnet2_batch layout (pointnet2_batch/src/{ball_query,group_points,sampling,interpolate}_gpu.cu; pointnet2_api.cpp:10-24),
 *   xyz (b,n,3), features (b,c,n):
 *   crb3d_ball_query_batch: idx (b,m,nsample) int32, ZERO-FILLED by the caller (an empty ball keeps its zeros; this layout has
 *     no -1 marker), hits with dist2 < r2 in ascending index order, padded with the first hit.
 *   crb3d_group_points_batch / _grad_batch: out (b,c,npoints,nsample) = points[b,c,idx]; the gradient is accumulated into
 *     grad_points (b,c,n) (caller zero-fills). gather_points[_grad] of sampling_gpu.cu:15-70 is the nsample = 1 case.
 *   crb3d_three_nn_batch / crb3d_three_interpolate_batch / _grad_batch: dist2/idx (b,n,3) of the three nearest of known
 *     (b,m,3) (strict <, ascending scan: the lowest index wins a tie); out (b,c,n) = sum_k weight[b,n,k] * points[b,c,idx].
 *   farthest point sampling of this layout is crb3d_farthest_point_sampling (same launcher in the reference).
 * crb3d_roipoint_pool3d_forward: roipoint_pool3d/src/roipoint_pool3d_kernel.cu:38-165 + roipoint_pool3d.cpp:23-46: for every
 *   box (already enlarged by the caller) the first S inside points in index order, repeated cyclically when fewer, xyz +
 *   features copied to pooled (B,M,S,3+C); empty_flag (B,M) int32 = 1 for a box without points (its pooled rows are not
 *   written). Workspace instead of the reference's cudaMalloc of a (B,N,M) assignment matrix inside the call. */
int crb3d_voxel_query_stack(int M, int R1, int R2, int R3, int nsample, float radius, int z_range, int y_range, int x_range,
                            const float* new_xyz, const float* xyz, const int* new_coords, const int* point_indices, int* idx,
                            cudaStream_t stream);
int crb3d_ball_query_batch(int b, int n, int m, float radius, int nsample, const float* new_xyz, const float* xyz, int* idx,
                           cudaStream_t stream);
int crb3d_group_points_batch(int b, int c, int n, int npoints, int nsample, const float* points, const int* idx, float* out,
                             cudaStream_t stream);
int crb3d_group_points_grad_batch(int b, int c, int n, int npoints, int nsample, const float* grad_out, const int* idx,
                                  float* grad_points, cudaStream_t stream);
int crb3d_three_nn_batch(int b, int n, int m, const float* unknown, const float* known, float* dist2, int* idx, cudaStream_t stream);
int crb3d_three_interpolate_batch(int b, int c, int m, int n, const float* points, const int* idx, const float* weight, float* out,
                                  cudaStream_t stream);
int crb3d_three_interpolate_grad_batch(int b, int c, int n, int m, const float* grad_out, const int* idx, const float* weight,
                                       float* grad_points, cudaStream_t stream);
int crb3d_roipoint_pool3d_workspace_bytes(int B, int M, int S, size_t* bytes);
int crb3d_roipoint_pool3d_forward(int B, int N, int M, int C, int S, const float* xyz, const float* boxes3d,
                                  const float* pts_feature, float* pooled, int* empty_flag, void* ws, size_t ws_bytes,
                                  cudaStream_t stream);

/* ---- vector-pool aggregation of PV-RCNN++ (pointnet2_stack/src/vector_pool_gpu.cu:19-476, vector_pool.cpp:35-204;
 * csrc/vector_pool.cu). A warp per query instead of the reference's thread per query; neighbour lists are laid out in query order
 * (count -> scan -> fill; the reference's layout depends on thread scheduling, `start_len` locates each list either way), sums are
 * accumulated in the reference's order.
 *   crb3d_query_stacked_local_neighbor_idxs: start_len (M,2) int32 [offset, length], stack_neighbor_idxs (avg_length * M) int32,
 *     cumsum (1) int32 in/out running total; neighbor_type 1 = ball, else cube; nsample > 0 limits a list (always <= 1000).
 *   crb3d_query_three_nn_by_stacked_local_idxs: new_xyz_grid_centers / _idxs / _dist2 (M, num_total_grids, 3).
 *   crb3d_vector_pool_stack: new_features (M,c_out), new_local_xyz (M,3*G), point_cnt_of_grid (M,G), grouped_idxs (max,3) zero-filled
 *     by the caller; cum_sum (DEVICE int) = number of grouped entries wanted (> num_max_sum_points: re-run with more room).
 *   crb3d_vector_pool_grad_stack: grad_support_features (N,c_in) += over the first num_grouped rows of grouped_idxs. */
int crb3d_query_stacked_local_neighbor_idxs_workspace_bytes(int M, size_t* bytes);
int crb3d_query_stacked_local_neighbor_idxs(const float* support_xyz, const int* xyz_batch_cnt, const float* new_xyz,
                                            const int* new_xyz_batch_cnt, int batch_size, int M, int* stack_neighbor_idxs,
                                            int* start_len, int* cumsum, int avg_length_of_neighbor_idxs,
                                            float max_neighbour_distance, int nsample, int neighbor_type, void* ws, size_t ws_bytes,
                                            cudaStream_t stream);
int crb3d_query_three_nn_by_stacked_local_idxs(const float* support_xyz, const float* new_xyz_grid_centers, int* new_xyz_grid_idxs,
                                               float* new_xyz_grid_dist2, const int* stack_neighbor_idxs, const int* start_len, int M,
                                               int num_total_grids, cudaStream_t stream);
int crb3d_vector_pool_stack(const float* support_xyz, const float* support_features, const int* xyz_batch_cnt, const float* new_xyz,
                            const int* new_xyz_batch_cnt, int batch_size, int M, int num_c_in, int num_c_out, int num_grid_x,
                            int num_grid_y, int num_grid_z, float max_neighbour_distance, int use_xyz, int num_max_sum_points,
                            int nsample, int neighbor_type, int pooling_type, float* new_features, float* new_local_xyz,
                            int* point_cnt_of_grid, int* grouped_idxs, int* cum_sum, cudaStream_t stream);
int crb3d_vector_pool_grad_stack(const float* grad_new_features, const int* point_cnt_of_grid, const int* grouped_idxs,
                                 float* grad_support_features, int num_c_out, int num_c_in, int num_total_grids, int num_grouped,
                                 cudaStream_t stream);

/* ---- PV-RCNN: fused set-abstraction layer and the RoI-head FC GEMM -------------------------------------------------------
 * crb3d_sa_group_mlp_maxpool: one scale of StackSAModuleMSG.forward (pcdet/ops/pointnet2/pointnet2_stack/
 * pointnet2_modules.py:78-112: QueryAndGroup + shared 1x1-conv MLP + BatchNorm + ReLU + max-pool over the samples) in one
 * pass, nothing materialised; used by VoxelSetAbstraction (voxel_set_abstraction.py:334-411) and PVRCNNHead.roi_grid_pool
 * (pvrcnn_head.py:68-114). idx (M,nsample) int32 comes from crb3d_ball_query_stack (idx[m][0] = -1: empty ball).
 * widths (HOST int[n_layers+1]): widths[0] = 3 + C, then the layer widths (<= 128, n_layers <= 3); packed (DEVICE): per
 * layer the TRANSPOSED weight [widths[l]][widths[l+1]] with the BatchNorm scale folded in, then the bias [widths[l+1]].
 * out: row m -> out + m*out_stride, columns [0, widths[n_layers]). Exact fp32. */
int crb3d_sa_group_mlp_maxpool(int B, const float* xyz, const int* xyz_cnt, const float* feat, int C, const float* new_xyz,
                               const int* new_cnt, int M, const int* idx, int nsample, int n_layers, const int* widths,
                               const float* packed, float* out, int out_stride, cudaStream_t stream);
/* crb3d_fc_gemm_tf32: out (M,N) = relu?((A (M,K) @ W (N,K)^T) * scale + shift), split-K tcgen05 GEMM for few rows and a long
 * K: the first shared FC of PVRCNNHead (pvrcnn_head.py:21-33: Conv1d(27648 -> 256, k = 1) + BatchNorm1d + ReLU on 128 RoIs
 * per frame). N % 128 == 0, K % 32 == 0, lda % 4 == 0; scale / shift nullable; deterministic (fixed split order). */
int crb3d_fc_gemm_workspace_bytes(long long M, int N, int K, size_t* bytes);
int crb3d_fc_gemm_tf32(const float* A, long long M, int K, long long lda, const float* W, int N, const float* scale,
                       const float* shift, int relu, float* out, void* ws, size_t ws_bytes, cudaStream_t stream);
/* debug: per-CTA phase stamps (16 int64 per CTA) of the last launch made with relu bit 9 set (tools/bench_bev.py trace). */
int crb3d_bev_conv3x3_trace(long long* host_out, int n_ctas);

/* ---- anchor-head post-processing (anchor_head_template.py:238-285, box_coder_utils.py:45-77,
 *      detector3d_template.py:281-311): max-class sigmoid score + 1-based label for every anchor, and lazy
 *      ResidualCoder decoding (+ direction-bin fix) of the selected anchors only. cls/box/dir are channels-last
 *      (B, H*W*A_loc, n_class | 7 | n_bins). spec: HOST pointer to the AnchorSpec struct of csrc/head.cu. */
int crb3d_anchor_head_scores(const float* cls_preds, int64_t n_anchor_total, int n_class, float* score, int* label,
                             cudaStream_t stream);
/* scores + labels of every anchor and, per frame, the anchors with score >= thresh sorted by descending score (ties:
 * ascending index) cut at K <= 4096 - the front half of class_agnostic_nms (model_nms_utils.py:6-25) without a full top-k.
 * cand (B, n_per_frame) uint64 and cand_count (B * 2049 ints) are scratch; counts[b] = valid prefix of top_score/top_idx (B, K). */
int crb3d_anchor_head_scores_topk(const float* cls_preds, int B, int64_t n_per_frame, int n_class, float thresh, int K,
                                  float* score, int* label, unsigned long long* cand, int* cand_count, float* top_score,
                                  long long* top_idx, int* counts, cudaStream_t stream);
int crb3d_anchor_decode_select(const float* box_preds, const float* dir_preds, const long long* sel, int B, int K,
                               int64_t n_anchor_per_frame, const void* spec, float* out, cudaStream_t stream);
int crb3d_gather_rows_f32(const float* src, const long long* idx, const int* valid, int B, int K, int64_t n_src,
                          int width, float fill, float* out, cudaStream_t stream);
int crb3d_gather_rows_i32(const int* src, const long long* idx, const int* valid, int B, int K, int64_t n_src,
                          int width, int fill, int* out, cudaStream_t stream);

/* ---- anchor-head training path (csrc/train_ops.cu) -----------------------------------------------------
 * crb3d_assign_targets_axis_aligned: AxisAlignedTargetAssigner.assign_targets / assign_targets_single
 *   (pcdet/models/dense_heads/target_assigner/axis_aligned_target_assigner.py:36-210; POS_FRACTION < 0, match_height False:
 *   box_utils.boxes3d_nearest_bev_iou, box_utils.py:249-298) + ResidualCoder.encode_torch (box_coder_utils.py:13-43) for a whole
 *   batch. anchors (A,7) in the head's (y, x, type) order; type_class / matched / unmatched: HOST arrays over the n_types anchor
 *   types of a location (1-based class, thresholds of ANCHOR_GENERATOR_CONFIG); gt_boxes (B,M,gt_stride>=8) with the class in the
 *   last column (<= 0: padding); M <= 128. labels (B,A) int32: -1 ignore / 0 background / class; reg_targets (B,A,7);
 *   reg_weights (B,A); num_pos (B) int32.
 * crb3d_anchor_head_loss: AnchorHeadTemplate.get_loss (anchor_head_template.py:101-229): sigmoid focal loss, smooth-L1 with the
 *   heading's sin difference, direction-bin cross entropy (loss_utils.py:9-135), normalised by each frame's positives, / B,
 *   x LOSS_WEIGHTS. losses (DEVICE float[3]) = {rpn_loss_cls, rpn_loss_loc, rpn_loss_dir}; g_* (nullable) = d(sum)/d(pred).
 *   code_weights (HOST float[7], null = ones), loss_weights3 (HOST {cls, loc, dir}). Deterministic (fixed-order reduction). */
int crb3d_assign_targets_workspace_bytes(int B, int64_t A, int M, size_t* bytes);
int crb3d_assign_targets_axis_aligned(const float* anchors, int64_t A, int n_types, const int* type_class, const float* matched,
                                      const float* unmatched, const float* gt_boxes, int B, int M, int gt_stride, int* labels,
                                      float* reg_targets, float* reg_weights, int* num_pos, void* ws, size_t ws_bytes,
                                      cudaStream_t stream);
int crb3d_anchor_head_loss_workspace_bytes(int B, int64_t A, size_t* bytes);
int crb3d_anchor_head_loss(const float* cls_preds, const float* box_preds, const float* dir_preds, const int* labels,
                           const float* reg_targets, const float* anchors, int B, int64_t A, int n_class, int num_dir_bins,
                           const float* code_weights, float alpha, float gamma, float beta, float dir_offset,
                           const float* loss_weights3, float* losses, float* g_cls, float* g_box, float* g_dir, void* ws,
                           size_t ws_bytes, cudaStream_t stream);

/* ---- CRB scoring (pcdet/query_strategies/crb_sampling.py:86-100, 219-226, 276-338) -------------------------- */
int crb3d_label_entropy(const int* labels, const int* box_off, int B, int num_class, float* entropy, int* class_counts,
                        cudaStream_t stream);
int crb3d_label_entropy_ranges(const int* labels, const int* box_begin, const int* box_end, int B, int num_class,
                               float* entropy, int* class_counts, cudaStream_t stream);
int crb3d_pairwise_sqdist_f64(const float* X, int n, int d, double* D, cudaStream_t stream);
int crb3d_kde_greedy_workspace_bytes(int n_cand, int n_class, size_t* bytes);
int crb3d_kde_greedy(const float* dens, const int* labels, const int* cand_off, int n_cand, int n_class,
                     const double* axis, const double* prior_n, double bandwidth, int n_select, int* order,
                     double* picked_score, void* ws, size_t ws_bytes, cudaStream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* CRB3D_H_ */
