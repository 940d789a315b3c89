/*
 * oracle.c - CPU restatement (plain C, serial) of the integer/geometry parts of the CRB-active-3Ddet hot path.
 *
 * TEST INFRASTRUCTURE ONLY. Nothing under oracle/ may be imported, linked or executed by the product path
 * (crb-active-3ddet_b200/); only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs use it, and only as the checker or the timed CPU baseline.
 *
 * Build: gcc -O2 -ffp-contract=off -fPIC -shared (see oracle/build.py). -ffp-contract=off matters: every fused
 * multiply-add below is written explicitly (fmaf) where the CUDA reference's SASS contracts one, so that index
 * results are bit-identical to the GPU kernels.
 *
 * Each function cites the reference file:line (under /root/reference) whose behaviour it restates.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------------------------------
 * Hard voxelization, spconv 2.1 Point2VoxelCPU3d semantics as used by
 * pcdet/datasets/processor/data_processor.py:15-60,115-143 (third-party spconv-cu113==2.1.21, not vendored).
 * One frame. points (n, stride); xyz at [0..2]; features = first n_feat columns.
 * Outputs: voxels (max_voxels, max_pts, n_feat) zero padded, coords (max_voxels, 3) zyx, num (max_voxels).
 * Returns the number of voxels.
 * ---------------------------------------------------------------------------------------------------------- */
typedef struct { int64_t key; int val; } vslot_t;

int oracle_voxelize(const float* pts, int n, int stride, int n_feat, const float* range6, const float* vsize3,
                    const int* grid3, int max_pts, int max_voxels, float* voxels, int* coords, int* num) {
    size_t cap = 1;
    while (cap < (size_t)n * 2 + 2) cap <<= 1;
    vslot_t* tab = (vslot_t*)malloc(cap * sizeof(vslot_t));
    for (size_t i = 0; i < cap; ++i) tab[i].key = -1;
    int voxel_num = 0;
    memset(voxels, 0, sizeof(float) * (size_t)max_voxels * max_pts * n_feat);
    memset(num, 0, sizeof(int) * (size_t)max_voxels);
    for (int i = 0; i < n; ++i) {
        const float* p = pts + (size_t)i * stride;
        int c[3];
        int failed = 0;
        for (int j = 0; j < 3; ++j) {
            float f = floorf((p[j] - range6[j]) / vsize3[j]);
            if (!(f >= 0.0f) || !(f < (float)grid3[j])) { failed = 1; break; }
            c[j] = (int)f;
        }
        if (failed) continue;
        int64_t key = ((int64_t)c[2] * grid3[1] + c[1]) * grid3[0] + c[0];
        uint64_t h = (uint64_t)key * 0x9E3779B97F4A7C15ull;
        size_t s = (size_t)(h >> 20) & (cap - 1);
        int vid = -1;
        while (1) {
            if (tab[s].key == key) { vid = tab[s].val; break; }
            if (tab[s].key == -1) break;
            s = (s + 1) & (cap - 1);
        }
        if (vid == -1) {
            if (voxel_num >= max_voxels) continue; /* unseen voxel after the cap: point dropped */
            vid = voxel_num++;
            tab[s].key = key; tab[s].val = vid;
            coords[vid * 3 + 0] = c[2]; coords[vid * 3 + 1] = c[1]; coords[vid * 3 + 2] = c[0];
        }
        if (num[vid] < max_pts) {
            memcpy(voxels + ((size_t)vid * max_pts + num[vid]) * n_feat, p, sizeof(float) * n_feat);
            num[vid]++;
        }
    }
    free(tab);
    return voxel_num;
}

/* MeanVFE (pcdet/models/backbones_3d/vfe/mean_vfe.py:14-31): sum over the point slots / clamp_min(num, 1). */
void oracle_mean_vfe(const float* voxels, const int* num, int m, int max_pts, int n_feat, float* out) {
    for (int v = 0; v < m; ++v)
        for (int f = 0; f < n_feat; ++f) {
            float s = 0.0f;
            for (int r = 0; r < max_pts; ++r) s += voxels[((size_t)v * max_pts + r) * n_feat + f];
            out[(size_t)v * n_feat + f] = s / (float)(num[v] > 1 ? num[v] : 1);
        }
}

/* ------------------------------------------------------------------------------------------------------------
 * Rotated BEV overlap. Restates pcdet/ops/iou3d_nms/src/iou3d_cpu.cpp:60-229 (same algorithm as
 * iou3d_nms_kernel.cu:36-234): rotated corners, 16 edge-edge crossings, corner containment with MARGIN 1e-2,
 * angular sort about the centroid, shoelace area.
 * ---------------------------------------------------------------------------------------------------------- */
typedef struct { float x, y; } pt_t;

static float cross3f(pt_t a, pt_t b, pt_t o) { return (a.x - o.x) * (b.y - o.y) - (b.x - o.x) * (a.y - o.y); }

static int seg_hit(pt_t p1, pt_t p0, pt_t q1, pt_t q0, pt_t* ans) {
    if (!(fminf(p0.x, p1.x) <= fmaxf(q0.x, q1.x) && fminf(q0.x, q1.x) <= fmaxf(p0.x, p1.x) &&
          fminf(p0.y, p1.y) <= fmaxf(q0.y, q1.y) && fminf(q0.y, q1.y) <= fmaxf(p0.y, p1.y)))
        return 0;
    float s1 = cross3f(q0, p1, p0), s2 = cross3f(p1, q1, p0), s3 = cross3f(p0, q1, q0), s4 = cross3f(q1, p1, q0);
    if (!(s1 * s2 > 0 && s3 * s4 > 0)) return 0;
    float s5 = cross3f(q1, p1, p0);
    if (fabsf(s5 - s1) > 1e-8f) {
        ans->x = (s5 * q0.x - s1 * q1.x) / (s5 - s1);
        ans->y = (s5 * q0.y - s1 * q1.y) / (s5 - s1);
    } else {
        float a0 = p0.y - p1.y, b0 = p1.x - p0.x, c0 = p0.x * p1.y - p1.x * p0.y;
        float a1 = q0.y - q1.y, b1 = q1.x - q0.x, c1 = q0.x * q1.y - q1.x * q0.y;
        float D = a0 * b1 - a1 * b0;
        ans->x = (b0 * c1 - b1 * c0) / D;
        ans->y = (a1 * c0 - a0 * c1) / D;
    }
    return 1;
}

static void box_corners(const float* b, pt_t* c) {
    float hx = b[3] / 2, hy = b[4] / 2, cs = cosf(b[6]), sn = sinf(b[6]);
    float x1 = b[0] - hx, y1 = b[1] - hy, x2 = b[0] + hx, y2 = b[1] + hy;
    float qx[4] = {x1, x2, x2, x1}, qy[4] = {y1, y1, y2, y2};
    for (int k = 0; k < 4; ++k) {
        c[k].x = (qx[k] - b[0]) * cs + (qy[k] - b[1]) * (-sn) + b[0];
        c[k].y = (qx[k] - b[0]) * sn + (qy[k] - b[1]) * cs + b[1];
    }
    c[4] = c[0];
}

static int corner_in(const float* b, pt_t p) {
    const float margin = 1e-2f;
    float cs = cosf(-b[6]), sn = sinf(-b[6]);
    float rx = (p.x - b[0]) * cs + (p.y - b[1]) * (-sn);
    float ry = (p.x - b[0]) * sn + (p.y - b[1]) * cs;
    return fabsf(rx) < b[3] / 2 + margin && fabsf(ry) < b[4] / 2 + margin;
}

float oracle_box_overlap(const float* a, const float* b) {
    pt_t ca[5], cb[5], v[16], ctr = {0.f, 0.f};
    box_corners(a, ca);
    box_corners(b, cb);
    int cnt = 0;
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j)
            if (seg_hit(ca[i + 1], ca[i], cb[j + 1], cb[j], &v[cnt])) { ctr.x += v[cnt].x; ctr.y += v[cnt].y; ++cnt; }
    for (int k = 0; k < 4; ++k) {
        if (corner_in(a, cb[k])) { ctr.x += cb[k].x; ctr.y += cb[k].y; v[cnt++] = cb[k]; }
        if (corner_in(b, ca[k])) { ctr.x += ca[k].x; ctr.y += ca[k].y; v[cnt++] = ca[k]; }
    }
    if (cnt == 0) return 0.0f;
    ctr.x /= cnt; ctr.y /= cnt;
    for (int j = 0; j < cnt - 1; ++j)
        for (int i = 0; i < cnt - j - 1; ++i)
            if (atan2f(v[i].y - ctr.y, v[i].x - ctr.x) > atan2f(v[i + 1].y - ctr.y, v[i + 1].x - ctr.x)) {
                pt_t t = v[i]; v[i] = v[i + 1]; v[i + 1] = t;
            }
    float area = 0.f;
    for (int k = 0; k < cnt - 1; ++k) {
        float ax = v[k].x - v[0].x, ay = v[k].y - v[0].y, bx = v[k + 1].x - v[0].x, by = v[k + 1].y - v[0].y;
        area += ax * by - ay * bx;
    }
    return fabsf(area) / 2.0f;
}

float oracle_iou_bev(const float* a, const float* b) {
    /* boxes further apart than their half diagonals (+ margin slack) cannot touch: overlap is exactly 0 */
    float ddx = a[0] - b[0], ddy = a[1] - b[1];
    float reach = 0.5f * (sqrtf(a[3] * a[3] + a[4] * a[4]) + sqrtf(b[3] * b[3] + b[4] * b[4])) + 0.1f;
    if (ddx * ddx + ddy * ddy > reach * reach) return 0.0f;
    float sa = a[3] * a[4], sb = b[3] * b[4], ov = oracle_box_overlap(a, b);
    return ov / fmaxf(sa + sb - ov, 1e-8f);
}

void oracle_pairwise(const float* a, int na, const float* b, int nb, int iou, float* out) {
    for (int i = 0; i < na; ++i)
        for (int j = 0; j < nb; ++j)
            out[(size_t)i * nb + j] = iou ? oracle_iou_bev(a + i * 7, b + j * 7) : oracle_box_overlap(a + i * 7, b + j * 7);
}

static float iou_axis_aligned(const float* a, const float* b) { /* iou3d_nms_kernel.cu:314-325 */
    float left = fmaxf(a[0] - a[3] / 2, b[0] - b[3] / 2), right = fminf(a[0] + a[3] / 2, b[0] + b[3] / 2);
    float top = fmaxf(a[1] - a[4] / 2, b[1] - b[4] / 2), bottom = fminf(a[1] + a[4] / 2, b[1] + b[4] / 2);
    float w = fmaxf(right - left, 0.f), h = fmaxf(bottom - top, 0.f), inter = w * h;
    return inter / fmaxf(a[3] * a[4] + b[3] * b[4] - inter, 1e-8f);
}

/* Greedy NMS over score-sorted boxes: the bitmask of iou3d_nms_kernel.cu:267-311 followed by the host loop of
 * iou3d_nms.cpp:116-132 collapses to "keep i unless a kept j < i has IoU(j, i) > thresh".
 * iou_out (optional, n*n) receives the IoU matrix so tests can assert a margin around the threshold. */
int oracle_nms(const float* boxes, int n, float thresh, int rotated, int64_t* keep, float* iou_out) {
    unsigned char* removed = (unsigned char*)calloc((size_t)n + 1, 1);
    int nk = 0;
    for (int i = 0; i < n; ++i) {
        if (iou_out)
            for (int j = i + 1; j < n; ++j)
                iou_out[(size_t)i * n + j] = rotated ? oracle_iou_bev(boxes + i * 7, boxes + j * 7)
                                                     : iou_axis_aligned(boxes + i * 7, boxes + j * 7);
        if (removed[i]) continue;
        keep[nk++] = i;
        for (int j = i + 1; j < n; ++j) {
            if (removed[j]) continue;
            float v = iou_out ? iou_out[(size_t)i * n + j]
                              : (rotated ? oracle_iou_bev(boxes + i * 7, boxes + j * 7)
                                         : iou_axis_aligned(boxes + i * 7, boxes + j * 7));
            if (v > thresh) removed[j] = 1;
        }
    }
    free(removed);
    return nk;
}

/* ------------------------------------------------------------------------------------------------------------
 * points in boxes, GPU predicate: pcdet/ops/roiaware_pool3d/src/roiaware_pool3d_kernel.cu:16-36,313-336.
 * margin is a parameter so the same code restates the CPU op (roiaware_pool3d.cpp:119-141, MARGIN 1e-2).
 * The rotation is written with the contraction the reference SASS uses:
 *   local_x = fma(sx, cos, sy*(-sina));  local_y = fma(sy, cos, -(sx*(-sina)))   with sina = sin(-rz).
 * ---------------------------------------------------------------------------------------------------------- */
static int pt_in_box(const float* p, const float* b, float margin, float* lx, float* ly) {
    if ((double)fabsf(p[2] - b[2]) > (double)b[5] / 2.0) return 0;
    float sx = p[0] - b[0], sy = p[1] - b[1];
    float c = cosf(b[6]), s = sinf(b[6]); /* cos(-rz) = c, -sin(-rz) = s */
    *lx = fmaf(sx, c, sy * s);
    *ly = fmaf(sy, c, -(sx * s));
    return ((double)fabsf(*lx) < (double)b[3] / 2.0 + (double)margin) & ((double)fabsf(*ly) < (double)b[4] / 2.0 + (double)margin);
}

/* first containing box per point (lowest box index wins), -1 = none */
void oracle_points_in_boxes(const float* boxes, int n_boxes, const float* pts, int n_pts, int pt_stride, int* out) {
    for (int j = 0; j < n_pts; ++j) {
        float lx, ly;
        out[j] = -1;
        for (int k = 0; k < n_boxes; ++k)
            if (pt_in_box(pts + (size_t)j * pt_stride, boxes + (size_t)k * 7, 1e-5f, &lx, &ly)) { out[j] = k; break; }
    }
}

void oracle_points_in_boxes_cpu(const float* boxes, int n_boxes, const float* pts, int n_pts, int* out) {
    for (int i = 0; i < n_boxes; ++i)
        for (int j = 0; j < n_pts; ++j) {
            float lx, ly;
            out[(size_t)i * n_pts + j] = pt_in_box(pts + (size_t)j * 3, boxes + (size_t)i * 7, 1e-2f, &lx, &ly);
        }
}

/* RoI-aware pooling forward: roiaware_pool3d_kernel.cu:39-190 (mask -> serial collect -> max/avg pool). */
void oracle_roiaware_pool(const float* rois, int n_boxes, const float* pts, const float* feat, int n_pts, int C,
                          int max_pts_each_voxel, int ox, int oy, int oz, int pool_method, int* argmax,
                          int* pts_idx_of_voxels, float* pooled) {
    size_t vox = (size_t)ox * oy * oz;
    for (int b = 0; b < n_boxes; ++b) {
        const float* roi = rois + (size_t)b * 7;
        int* lists = pts_idx_of_voxels + (size_t)b * vox * max_pts_each_voxel;
        for (int k = 0; k < n_pts; ++k) {
            float lx, ly;
            if (!pt_in_box(pts + (size_t)k * 3, roi, 1e-5f, &lx, &ly)) continue;
            float lz = pts[(size_t)k * 3 + 2] - roi[2];
            float xr = roi[3] / ox, yr = roi[4] / oy, zr = roi[5] / oz;
            unsigned int xi = (unsigned int)(int)((lx + roi[3] / 2) / xr);
            unsigned int yi = (unsigned int)(int)((ly + roi[4] / 2) / yr);
            unsigned int zi = (unsigned int)(int)((lz + roi[5] / 2) / zr);
            if (xi > (unsigned int)(ox - 1)) xi = ox - 1;
            if (yi > (unsigned int)(oy - 1)) yi = oy - 1;
            if (zi > (unsigned int)(oz - 1)) zi = oz - 1;
            int* l = lists + ((size_t)xi * oy * oz + (size_t)yi * oz + zi) * max_pts_each_voxel;
            if (l[0] < max_pts_each_voxel - 1) { l[l[0] + 1] = k; l[0]++; }
        }
        for (size_t v = 0; v < vox; ++v) {
            const int* l = lists + v * max_pts_each_voxel;
            for (int c = 0; c < C; ++c) {
                size_t o = ((size_t)b * vox + v) * C + c;
                if (pool_method == 0) {
                    int am = -1; float mx = -INFINITY;
                    for (int k = 1; k <= l[0]; ++k) {
                        float val = feat[(size_t)l[k] * C + c];
                        if (val > mx) { mx = val; am = l[k]; }
                    }
                    if (am != -1) pooled[o] = mx;
                    argmax[o] = am;
                } else {
                    float s = 0.f;
                    for (int k = 1; k <= l[0]; ++k) s += feat[(size_t)l[k] * C + c];
                    if (l[0] > 0) pooled[o] = s / l[0];
                }
            }
        }
    }
}

/* ------------------------------------------------------------------------------------------------------------
 * PointNet++ stacked ops. Squared distances use the contraction in the reference SASS:
 *   d = fma(dz, dz, fma(dx, dx, dy*dy)).
 * ---------------------------------------------------------------------------------------------------------- */
static float sqdist(float ax, float ay, float az, float bx, float by, float bz) {
    float dx = ax - bx, dy = ay - by, dz = az - bz;
    return fmaf(dz, dz, fmaf(dx, dx, dy * dy));
}

/* pointnet2_stack/src/ball_query_gpu.cu:16-66 ; idx (M, nsample) must be zero-initialised by the caller */
void oracle_ball_query(int B, float radius, int nsample, const float* new_xyz, const int* new_cnt, const float* xyz,
                       const int* xyz_cnt, int* idx) {
    float r2 = radius * radius;
    int q0 = 0, s0 = 0;
    for (int b = 0; b < B; ++b) {
        for (int q = q0; q < q0 + new_cnt[b]; ++q) {
            int cnt = 0;
            int* out = idx + (size_t)q * nsample;
            for (int k = 0; k < xyz_cnt[b]; ++k) {
                const float* p = xyz + (size_t)(s0 + k) * 3;
                float d2 = sqdist(new_xyz[(size_t)q * 3], new_xyz[(size_t)q * 3 + 1], new_xyz[(size_t)q * 3 + 2], p[0], p[1], p[2]);
                if (d2 < r2) {
                    if (cnt == 0) for (int l = 0; l < nsample; ++l) out[l] = k;
                    out[cnt++] = k;
                    if (cnt >= nsample) break;
                }
            }
            if (cnt == 0) out[0] = -1;
        }
        q0 += new_cnt[b]; s0 += xyz_cnt[b];
    }
}

/* pointnet2_stack/src/voxel_query_gpu.cu:13-98: neighbours through the voxel -> point table, (dz,dy,dx) scan order, kept when
 * dist2 <= radius^2 (not <), first hit pads the row, idx[0] = -1 for an empty neighbourhood. idx zero-initialised by the caller. */
void oracle_voxel_query(int M, int R1, int R2, int R3, int nsample, float radius, int zr, int yr, int xr, const float* new_xyz,
                        const float* xyz, const int* new_coords, const int* point_indices, int* idx) {
    float r2 = radius * radius;
    for (int q = 0; q < M; ++q) {
        const int* c = new_coords + (size_t)q * 4;
        int* out = idx + (size_t)q * nsample;
        int cnt = 0;
        for (int dz = -zr; dz <= zr; ++dz) {
            int z = c[1] + dz;
            if (z < 0 || z >= R1) continue;
            for (int dy = -yr; dy <= yr; ++dy) {
                int y = c[2] + dy;
                if (y < 0 || y >= R2) continue;
                for (int dx = -xr; dx <= xr; ++dx) {
                    int x = c[3] + dx;
                    if (x < 0 || x >= R3) continue;
                    int nb = point_indices[(((size_t)c[0] * R1 + z) * R2 + y) * R3 + x];
                    if (nb < 0) continue;
                    float d2 = sqdist(xyz[(size_t)nb * 3], xyz[(size_t)nb * 3 + 1], xyz[(size_t)nb * 3 + 2], new_xyz[(size_t)q * 3],
                                      new_xyz[(size_t)q * 3 + 1], new_xyz[(size_t)q * 3 + 2]);
                    if (d2 > r2) continue;
                    if (cnt < nsample) {
                        if (cnt == 0) for (int l = 0; l < nsample; ++l) out[l] = nb;
                        out[cnt++] = nb;
                    }
                }
            }
        }
        if (cnt == 0) out[0] = -1;
    }
}

/* roipoint_pool3d/src/roipoint_pool3d_kernel.cu:38-140 (assign_pts_to_box3d + get_pooled_idx + roipool3d_forward) for one frame:
 * first S inside points in index order, k % cnt duplication, empty flag; pooled (M,S,3+C) and empty (M) zero-initialised. */
void oracle_roipoint_pool3d(int N, int M, int C, int S, const float* xyz, const float* boxes, const float* feat, float* pooled,
                            int* empty) {
    for (int m = 0; m < M; ++m) {
        int cnt = 0;
        float* dst = pooled + (size_t)m * S * (3 + C);
        for (int k = 0; k < N && cnt < S; ++k) {
            float lx, ly;
            if (!pt_in_box(xyz + (size_t)k * 3, boxes + (size_t)m * 7, 1e-5f, &lx, &ly)) continue;
            float* d = dst + (size_t)cnt * (3 + C);
            for (int j = 0; j < 3; ++j) d[j] = xyz[(size_t)k * 3 + j];
            for (int j = 0; j < C; ++j) d[3 + j] = feat[(size_t)k * C + j];
            ++cnt;
        }
        if (cnt == 0) { empty[m] = 1; continue; }
        for (int s = cnt; s < S; ++s)
            for (int j = 0; j < 3 + C; ++j) dst[(size_t)s * (3 + C) + j] = dst[(size_t)(s % cnt) * (3 + C) + j];
    }
}

/* pointnet2_stack/src/group_points_gpu.cu:71-102 */
void oracle_group_points(int B, int C, int nsample, const float* feat, const int* feat_cnt, const int* idx,
                         const int* idx_cnt, float* out) {
    int m0 = 0, f0 = 0;
    for (int b = 0; b < B; ++b) {
        for (int m = m0; m < m0 + idx_cnt[b]; ++m)
            for (int c = 0; c < C; ++c)
                for (int s = 0; s < nsample; ++s)
                    out[((size_t)m * C + c) * nsample + s] = feat[(size_t)(f0 + idx[(size_t)m * nsample + s]) * C + c];
        m0 += idx_cnt[b]; f0 += feat_cnt[b];
    }
}

/* pointnet2_stack/src/sampling_gpu.cu:17-140. The reference scans with `block` threads (thread t owns k = t, t+block,
 * ...; strict '>' keeps the first maximum) and merges with a pairwise tree (slot t absorbs slot t+s for s = block/2..1,
 * the LEFT operand wins ties). Two tied threads first meet at the lowest bit in which their ids differ and the one
 * with a 0 there wins, so threads rank by bit-reversed id: tie order = (bitrev(k mod block), k / block).
 * block = min(1024, 2^floor(log2 n)) (sampling_gpu.cu:9-13). */
static unsigned bitrev(unsigned v, int bits) {
    unsigned r = 0;
    for (int i = 0; i < bits; ++i) r |= ((v >> i) & 1u) << (bits - 1 - i);
    return r;
}

void oracle_fps(int n, int m, int block, const float* pts, float* temp, int* idx) {
    if (m <= 0) return;
    int bits = 0;
    while ((1 << (bits + 1)) <= block) ++bits;
    int old = 0;
    idx[0] = 0;
    for (int j = 1; j < m; ++j) {
        float best = -1.0f; int besti = 0; long bestkey = -1;
        for (int k = 0; k < n; ++k) {
            float d = sqdist(pts[(size_t)k * 3], pts[(size_t)k * 3 + 1], pts[(size_t)k * 3 + 2], pts[(size_t)old * 3],
                             pts[(size_t)old * 3 + 1], pts[(size_t)old * 3 + 2]);
            float d2 = d < temp[k] ? d : temp[k];
            temp[k] = d2;
            long key = (long)bitrev((unsigned)(k % block), bits) * (1l << 32) + k / block;
            if (d2 > best || (d2 == best && bestkey >= 0 && key < bestkey)) { best = d2; besti = k; bestkey = key; }
        }
        old = besti;
        idx[j] = old;
    }
}

/* pointnet2_stack/src/interpolate_gpu.cu:16-75 */
void oracle_three_nn(int B, const float* unknown, const int* unknown_cnt, const float* known, const int* known_cnt,
                     float* dist2, int* idx) {
    int u0 = 0, k0 = 0;
    for (int b = 0; b < B; ++b) {
        for (int p = u0; p < u0 + unknown_cnt[b]; ++p) {
            double b1 = 1e40, b2 = 1e40, b3 = 1e40;
            int i1 = 0, i2 = 0, i3 = 0;
            for (int k = 0; k < known_cnt[b]; ++k) {
                const float* q = known + (size_t)(k0 + k) * 3;
                float d = sqdist(unknown[(size_t)p * 3], unknown[(size_t)p * 3 + 1], unknown[(size_t)p * 3 + 2], q[0], q[1], q[2]);
                if (d < b1) { b3 = b2; i3 = i2; b2 = b1; i2 = i1; b1 = d; i1 = k; }
                else if (d < b2) { b3 = b2; i3 = i2; b2 = d; i2 = k; }
                else if (d < b3) { b3 = d; i3 = k; }
            }
            dist2[(size_t)p * 3] = (float)b1; dist2[(size_t)p * 3 + 1] = (float)b2; dist2[(size_t)p * 3 + 2] = (float)b3;
            idx[(size_t)p * 3] = i1 + k0; idx[(size_t)p * 3 + 1] = i2 + k0; idx[(size_t)p * 3 + 2] = i3 + k0;
        }
        u0 += unknown_cnt[b]; k0 += known_cnt[b];
    }
}
