// Anchor-head post-processing kernels: per-anchor max-class score/label and lazy box decoding.
//
// Replaces, on the scoring path, the dense torch ops of
//   pcdet/models/dense_heads/anchor_head_template.py:238-285  generate_predicted_boxes (all 211 200 anchors/frame)
//   pcdet/utils/box_coder_utils.py:45-77                      ResidualCoder.decode_torch
//   pcdet/models/detectors/detector3d_template.py:281-311     sigmoid + max over classes (+1 for the label)
// The reference decodes every anchor and then keeps <= 4096 of them; here the score/label pass reads the class logits
// once, and boxes are decoded only for the anchors that survive the score filter + top-k (same values, 50x less work).
// Anchor layout (anchor_generator.py:18-62 + torch.cat(dim=-3) in generate_predicted_boxes): anchor index
// a = ((y*nx + x)*n_class_sets + set)*n_rot + rot ; logits / box codes / dir bins are channels-last per location.
#include "common.cuh"

#define CRB3D_MAX_ANCHOR_TYPES 16

struct AnchorSpec {
    int nx, ny, n_types;          // feature-map size and anchors per location
    float x0, y0;                 // first anchor centre
    double x_stride, y_stride;    // centre spacing (float64 like torch.arange's accumulator)
    float size[CRB3D_MAX_ANCHOR_TYPES][3];  // dx, dy, dz
    float rot[CRB3D_MAX_ANCHOR_TYPES];
    float zc[CRB3D_MAX_ANCHOR_TYPES];       // bottom height + dz/2
    float dir_offset, dir_limit_offset;
    int num_dir_bins;             // 0 = no direction classifier
};

namespace {

__global__ void __launch_bounds__(256) head_scores_kernel(const float* __restrict__ cls, int64_t n_anchor_total,
                                                          int n_class, float* __restrict__ score, int* __restrict__ label) {
    const int64_t a = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= n_anchor_total) return;
    const float* p = cls + a * n_class;
    float best = -INFINITY;
    int bi = 0;
    for (int c = 0; c < n_class; ++c) {
        const float v = p[c];
        if (v > best) { best = v; bi = c; }  // torch.max: first maximum
    }
    // sigmoid is monotonic: max(sigmoid(x)) == sigmoid(max(x))
    score[a] = 1.0f / (1.0f + expf(-best));
    label[a] = bi + 1;
}

__global__ void __launch_bounds__(128) head_decode_kernel(const float* __restrict__ box, const float* __restrict__ dir,
                                                          const long long* __restrict__ sel, int B, int K,
                                                          int64_t n_anchor_per_frame, AnchorSpec S,
                                                          float* __restrict__ out) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= B * K) return;
    const int b = t / K;
    long long a = sel[t];
    if (a < 0 || a >= n_anchor_per_frame) a = 0;
    const int type = (int)(a % S.n_types);
    const long long loc = a / S.n_types;
    const int x = (int)(loc % S.nx), y = (int)(loc / S.nx);
    const float xa = (float)((double)S.x0 + S.x_stride * x), ya = (float)((double)S.y0 + S.y_stride * y);
    const float za = S.zc[type], dxa = S.size[type][0], dya = S.size[type][1], dza = S.size[type][2], ra = S.rot[type];
    const float* e = box + ((size_t)b * n_anchor_per_frame + a) * 7;
    const float diag = sqrtf(dxa * dxa + dya * dya);
    float* o = out + (size_t)t * 7;
    o[0] = e[0] * diag + xa;
    o[1] = e[1] * diag + ya;
    o[2] = e[2] * dza + za;
    o[3] = expf(e[3]) * dxa;
    o[4] = expf(e[4]) * dya;
    o[5] = expf(e[5]) * dza;
    float rg = e[6] + ra;
    if (S.num_dir_bins > 0 && dir) {
        const float* d = dir + ((size_t)b * n_anchor_per_frame + a) * S.num_dir_bins;
        int dl = 0;
        float bd = d[0];
        for (int q = 1; q < S.num_dir_bins; ++q)
            if (d[q] > bd) { bd = d[q]; dl = q; }
        const float period = (float)(2.0 * 3.14159265358979323846 / S.num_dir_bins);
        const float v = rg - S.dir_offset;
        // common_utils.limit_period: val - floor(val / period + offset) * period
        const float dir_rot = v - floorf(v / period + S.dir_limit_offset) * period;
        rg = dir_rot + S.dir_offset + period * (float)dl;
    }
    o[6] = rg;
}

// gather helper: out[b][k][:] = src[b][idx[b][k]][:] for int/float rows (used to pick labels / boxes by keep lists)
template <typename T>
__global__ void __launch_bounds__(256) gather_rows_kernel(const T* __restrict__ src, const long long* __restrict__ idx,
                                                          const int* __restrict__ valid, int B, int K, int64_t n_src,
                                                          int width, T fill, T* __restrict__ out) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (int64_t)B * K * width) return;
    const int w = (int)(t % width);
    const int64_t bk = t / width;
    const int b = (int)(bk / K), k = (int)(bk % K);
    T v = fill;
    if (!valid || k < valid[b]) {
        long long i = idx[bk];
        if (i >= 0 && i < n_src) v = src[((size_t)b * n_src + i) * width + w];
    }
    out[t] = v;
}

}  // namespace

extern "C" int crb3d_anchor_head_scores(const float* cls_preds, int64_t n_anchor_total, int n_class, float* score,
                                        int* label, cudaStream_t stream) {
    if (n_anchor_total < 0 || n_class <= 0 || !score || !label) return CRB3D_ERR_ARG;
    if (n_anchor_total == 0) return CRB3D_OK;
    head_scores_kernel<<<(unsigned)crb3d_divup(n_anchor_total, 256), 256, 0, stream>>>(cls_preds, n_anchor_total, n_class,
                                                                                      score, label);
    CRB3D_CHECK_LAUNCH();
    return CRB3D_OK;
}

// spec: HOST pointer to an AnchorSpec (layout above, mirrored by crb3d/ops.py). sel: (B,K) int64 anchor indices.
extern "C" int crb3d_anchor_decode_select(const float* box_preds, const float* dir_preds, const long long* sel, int B,
                                          int K, int64_t n_anchor_per_frame, const void* spec, float* out,
                                          cudaStream_t stream) {
    if (B < 0 || K < 0 || !spec || !out) return CRB3D_ERR_ARG;
    if (B * K == 0) return CRB3D_OK;
    AnchorSpec S = *reinterpret_cast<const AnchorSpec*>(spec);
    if (S.n_types <= 0 || S.n_types > CRB3D_MAX_ANCHOR_TYPES) return CRB3D_ERR_UNSUPPORTED;
    head_decode_kernel<<<(unsigned)crb3d_divup(B * K, 128), 128, 0, stream>>>(box_preds, dir_preds, sel, B, K,
                                                                             n_anchor_per_frame, S, out);
    CRB3D_CHECK_LAUNCH();
    return CRB3D_OK;
}

// out[b][k][0..width) = src[b][idx[b][k]][0..width) for k < valid[b] (valid nullable), else `fill`.
extern "C" int crb3d_gather_rows_f32(const float* src, const long long* idx, const int* valid, int B, int K,
                                     int64_t n_src, int width, float fill, float* out, cudaStream_t stream) {
    if (B < 0 || K < 0 || width <= 0 || !out) return CRB3D_ERR_ARG;
    if ((int64_t)B * K == 0) return CRB3D_OK;
    gather_rows_kernel<float><<<(unsigned)crb3d_divup((int64_t)B * K * width, 256), 256, 0, stream>>>(src, idx, valid, B, K,
                                                                                                     n_src, width, fill, out);
    CRB3D_CHECK_LAUNCH();
    return CRB3D_OK;
}

extern "C" int crb3d_gather_rows_i32(const int* src, const long long* idx, const int* valid, int B, int K, int64_t n_src,
                                     int width, int fill, int* out, cudaStream_t stream) {
    if (B < 0 || K < 0 || width <= 0 || !out) return CRB3D_ERR_ARG;
    if ((int64_t)B * K == 0) return CRB3D_OK;
    gather_rows_kernel<int><<<(unsigned)crb3d_divup((int64_t)B * K * width, 256), 256, 0, stream>>>(src, idx, valid, B, K, n_src,
                                                                                                   width, fill, out);
    CRB3D_CHECK_LAUNCH();
    return CRB3D_OK;
}
