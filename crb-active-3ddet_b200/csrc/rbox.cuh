// Rotated-box BEV geometry shared by the device kernels (iou3d_nms.cu) and the host op (host_ops.cu).
// Arithmetic follows pcdet/ops/iou3d_nms/src/iou3d_nms_kernel.cu:36-234 / iou3d_cpu.cpp:60-229 step for step
// (see iou3d_nms.cu for the list of structural differences).
#pragma once
#include <math.h>
#ifdef __CUDACC__
#define CRB3D_HD __host__ __device__
#else
#define CRB3D_HD
#endif

constexpr float kEps = 1e-8f;

struct alignas(16) RBox {  // 64 bytes: moved as four float4
    float cx, cy, hx, hy, c, s;  // centre, half extents, cos/sin(heading)
    float px[4], py[4];          // rotated corners
    float area, rad;             // dx*dy, half diagonal
};

CRB3D_HD inline void make_rbox(const float* __restrict__ b, RBox& r) {
    r.cx = b[0]; r.cy = b[1];
    r.hx = b[3] / 2; r.hy = b[4] / 2;
    const float ang = b[6];
    r.c = cosf(ang); r.s = sinf(ang);
    const float x1 = r.cx - r.hx, y1 = r.cy - r.hy, x2 = r.cx + r.hx, y2 = r.cy + r.hy;
    const float qx[4] = {x1, x2, x2, x1}, qy[4] = {y1, y1, y2, y2};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        r.px[k] = (qx[k] - r.cx) * r.c + (qy[k] - r.cy) * (-r.s) + r.cx;
        r.py[k] = (qx[k] - r.cx) * r.s + (qy[k] - r.cy) * r.c + r.cy;
    }
    r.area = b[3] * b[4];
    r.rad = sqrtf(r.hx * r.hx + r.hy * r.hy);
}

CRB3D_HD inline float cross3(float ax, float ay, float bx, float by, float ox, float oy) {
    return (ax - ox) * (by - oy) - (bx - ox) * (ay - oy);
}

// segment p0->p1 against q0->q1 (iou3d_nms_kernel.cu:62-92)
CRB3D_HD inline bool seg_cross(float p1x, float p1y, float p0x, float p0y, float q1x, float q1y, float q0x,
                                          float q0y, float& ox, float& oy) {
    const bool hit = fminf(p0x, p1x) <= fmaxf(q0x, q1x) && fminf(q0x, q1x) <= fmaxf(p0x, p1x) &&
                     fminf(p0y, p1y) <= fmaxf(q0y, q1y) && fminf(q0y, q1y) <= fmaxf(p0y, p1y);
    if (!hit) return false;
    const float s1 = cross3(q0x, q0y, p1x, p1y, p0x, p0y);
    const float s2 = cross3(p1x, p1y, q1x, q1y, p0x, p0y);
    const float s3 = cross3(p0x, p0y, q1x, q1y, q0x, q0y);
    const float s4 = cross3(q1x, q1y, p1x, p1y, q0x, q0y);
    if (!(s1 * s2 > 0 && s3 * s4 > 0)) return false;
    const float s5 = cross3(q1x, q1y, p1x, p1y, p0x, p0y);
    if (fabsf(s5 - s1) > kEps) {
        ox = (s5 * q0x - s1 * q1x) / (s5 - s1);
        oy = (s5 * q0y - s1 * q1y) / (s5 - s1);
    } else {
        const float a0 = p0y - p1y, b0 = p1x - p0x, c0 = p0x * p1y - p1x * p0y;
        const float a1 = q0y - q1y, b1 = q1x - q0x, c1 = q0x * q1y - q1x * q0y;
        const float D = a0 * b1 - a1 * b0;
        ox = (b0 * c1 - b1 * c0) / D;
        oy = (a1 * c0 - a0 * c1) / D;
    }
    return true;
}

// corner containment with the reference's MARGIN (iou3d_nms_kernel.cu:51-60); cos(-t)=cos t, sin(-t)=-sin t
CRB3D_HD inline bool corner_inside(const RBox& b, float x, float y) {
    const float margin = 1e-2f;
    const float rx = (x - b.cx) * b.c + (y - b.cy) * b.s;
    const float ry = (x - b.cx) * (-b.s) + (y - b.cy) * b.c;
    return fabsf(rx) < b.hx + margin && fabsf(ry) < b.hy + margin;
}

CRB3D_HD inline float rbox_overlap(const RBox& A, const RBox& B) {
    // exact-zero early out: no edge can cross and no corner can be within MARGIN of the other box
    {
        const float dx = A.cx - B.cx, dy = A.cy - B.cy, reach = A.rad + B.rad + 0.1f;
        if (dx * dx + dy * dy > reach * reach) return 0.0f;
    }
    float vx[16], vy[16];
    int cnt = 0;
    float sx = 0.f, sy = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int i1 = (i + 1) & 3;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int j1 = (j + 1) & 3;
            float ox, oy;
            if (seg_cross(A.px[i1], A.py[i1], A.px[i], A.py[i], B.px[j1], B.py[j1], B.px[j], B.py[j], ox, oy)) {
                sx = sx + ox; sy = sy + oy;
                vx[cnt] = ox; vy[cnt] = oy; ++cnt;
            }
        }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        if (corner_inside(A, B.px[k], B.py[k])) {
            sx = sx + B.px[k]; sy = sy + B.py[k];
            vx[cnt] = B.px[k]; vy[cnt] = B.py[k]; ++cnt;
        }
        if (corner_inside(B, A.px[k], A.py[k])) {
            sx = sx + A.px[k]; sy = sy + A.py[k];
            vx[cnt] = A.px[k]; vy[cnt] = A.py[k]; ++cnt;
        }
    }
    if (cnt < 3) return 0.0f;  // fewer than 3 vertices: the shoelace sum is exactly 0
    sx /= cnt; sy /= cnt;
    // stable ascending sort by polar angle about the centroid (same permutation as the reference bubble sort)
    float ang[16];
    for (int t = 0; t < cnt; ++t) ang[t] = atan2f(vy[t] - sy, vx[t] - sx);
    for (int t = 1; t < cnt; ++t) {
        const float a = ang[t], x = vx[t], y = vy[t];
        int u = t - 1;
        while (u >= 0 && ang[u] > a) {
            ang[u + 1] = ang[u]; vx[u + 1] = vx[u]; vy[u + 1] = vy[u];
            --u;
        }
        ang[u + 1] = a; vx[u + 1] = x; vy[u + 1] = y;
    }
    float area = 0.f;
    for (int k = 0; k < cnt - 1; ++k) {
        const float ax = vx[k] - vx[0], ay = vy[k] - vy[0], bx = vx[k + 1] - vx[0], by = vy[k + 1] - vy[0];
        area += ax * by - ay * bx;
    }
    return fabsf(area) / 2.0f;
}

CRB3D_HD inline float rbox_iou(const RBox& A, const RBox& B) {
    const float ov = rbox_overlap(A, B);
    return ov / fmaxf(A.area + B.area - ov, kEps);
}

// axis-aligned IoU on [x,y,z,dx,dy,dz,heading] (iou3d_nms_kernel.cu:314-325)
CRB3D_HD inline float aabb_iou(const float* a, const float* b) {
    const float left = fmaxf(a[0] - a[3] / 2, b[0] - b[3] / 2), right = fminf(a[0] + a[3] / 2, b[0] + b[3] / 2);
    const float top = fmaxf(a[1] - a[4] / 2, b[1] - b[4] / 2), bottom = fminf(a[1] + a[4] / 2, b[1] + b[4] / 2);
    const float w = fmaxf(right - left, 0.f), h = fmaxf(bottom - top, 0.f);
    const float inter = w * h;
    return inter / fmaxf(a[3] * a[4] + b[3] * b[4] - inter, kEps);
}

