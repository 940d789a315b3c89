// Anchor-head training path: target assignment and the three anchor-head losses with their gradients.
//
// crb3d_assign_targets_axis_aligned restates, for a whole batch in two launches,
//   pcdet/models/dense_heads/target_assigner/axis_aligned_target_assigner.py:36-210 (assign_targets + assign_targets_single,
//   POS_FRACTION < 0 branch: no sampling - every reference config sets -1) with match_height = False:
//   box_utils.boxes3d_nearest_bev_iou (box_utils.py:249-298) between the anchors of a class and the ground-truth boxes of that
//   class, anchor -> gt arg-max, gt -> anchor max (forced matches: every anchor that reaches a gt's best IoU), matched / unmatched
//   thresholds, ResidualCoder.encode_torch (box_coder_utils.py:13-43) of the matched gt for the foreground anchors.
//   The reference loops over frames x classes in Python and materialises an (anchors x gts) IoU matrix per pair; here a thread
//   owns one anchor of one frame, the frame's boxes sit in shared memory, the gt -> anchor maxima are integer atomicMax of the
//   non-negative IoU bits (pass 1) and the equality test of pass 2 recomputes the same IoU bit for bit.
//   Labels are index-exact: the IoU arithmetic mirrors torch's op-by-op fp32 evaluation (no FMA contraction; `x / python_scalar`
//   is a multiplication by the fp32 reciprocal in torch's CUDA kernels).
// crb3d_anchor_head_loss = AnchorHeadTemplate.get_loss (anchor_head_template.py:101-229): SigmoidFocalClassificationLoss
//   (loss_utils.py:9-76), WeightedSmoothL1Loss with the sin-difference on the heading (loss_utils.py:79-135,
//   anchor_head_template.py:146-154), WeightedCrossEntropyLoss on the direction bins (anchor_head_template.py:156-171, 201-215),
//   each normalised by the frame's positives, summed, divided by the batch size and scaled by its LOSS_WEIGHT - plus the three
//   gradients in the same pass (what autograd would produce through ~60 elementwise launches). Deterministic: per-block partial
//   sums reduced in a fixed order.
#include "common.cuh"
#include <cmath>

namespace {

constexpr int MAX_GT = 128;             // ground-truth rows per frame staged in shared memory
constexpr int MAX_TYPES = 16;           // anchor types per location

struct AssignSpec {
    int n_types;
    int type_class[MAX_TYPES];          // 1-based class of each anchor type
    float matched[MAX_TYPES], unmatched[MAX_TYPES];
};

struct BevBox { float x1, y1, x2, y2, area; };

// box_utils.boxes3d_lidar_to_aligned_bev_boxes, torch op by op (fp32, no contraction)
__device__ __forceinline__ BevBox aligned_bev(const float* b) {
    const float PI_F = 3.14159274101257324f;                        // fp32(np.pi)
    const float inv_pi = __fdiv_rn(1.0f, PI_F);                     // torch: tensor / python scalar = tensor * (1 / scalar)
    const float t = floorf(__fadd_rn(__fmul_rn(b[6], inv_pi), 0.5f));
    const float rot = fabsf(__fsub_rn(b[6], __fmul_rn(t, PI_F)));
    const bool keep = rot < 0.785398185253143311f;                  // fp32(np.pi / 4)
    const float dx = keep ? b[3] : b[4], dy = keep ? b[4] : b[3];
    const float hx = __fmul_rn(dx, 0.5f), hy = __fmul_rn(dy, 0.5f);
    BevBox r;
    r.x1 = __fsub_rn(b[0], hx); r.y1 = __fsub_rn(b[1], hy);
    r.x2 = __fadd_rn(b[0], hx); r.y2 = __fadd_rn(b[1], hy);
    r.area = __fmul_rn(__fsub_rn(r.x2, r.x1), __fsub_rn(r.y2, r.y1));
    return r;
}
// box_utils.boxes_iou_normal
__device__ __forceinline__ float iou_normal(const BevBox& a, const BevBox& b) {
    const float xl = fmaxf(__fsub_rn(fminf(a.x2, b.x2), fmaxf(a.x1, b.x1)), 0.0f);
    const float yl = fmaxf(__fsub_rn(fminf(a.y2, b.y2), fmaxf(a.y1, b.y1)), 0.0f);
    const float inter = __fmul_rn(xl, yl);
    return __fdiv_rn(inter, fmaxf(__fsub_rn(__fadd_rn(a.area, b.area), inter), 1e-6f));
}

struct FrameGt {
    BevBox box[MAX_GT];
    int cls[MAX_GT];
};

__device__ __forceinline__ void stage_gt(const float* __restrict__ gt, int M, int gt_stride, FrameGt& s) {
    for (int g = threadIdx.x; g < M; g += blockDim.x) {
        const float* row = gt + (size_t)g * gt_stride;
        s.box[g] = aligned_bev(row);
        s.cls[g] = (int)row[gt_stride - 1];           // gt_boxes[..., -1]; padded rows carry class 0
    }
    __syncthreads();
}

// pass 1: per anchor the best gt of its class (first maximum, like torch.argmax over the class-filtered list) and, per gt, the
// best IoU over the anchors of its class
__global__ void __launch_bounds__(256) assign_pass1(const float* __restrict__ anchors, int64_t A, AssignSpec spec, const float* __restrict__ gt,
                                                    int M, int gt_stride, float* __restrict__ amax, int* __restrict__ aarg,
                                                    int* __restrict__ gt_best) {
    __shared__ FrameGt s;
    const int b = blockIdx.y;
    stage_gt(gt + (size_t)b * M * gt_stride, M, gt_stride, s);
    const int64_t a = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= A) return;
    const int cls = spec.type_class[a % spec.n_types];
    const BevBox ab = aligned_bev(anchors + a * 7);
    float best = -1.0f;
    int arg = -1;
    for (int g = 0; g < M; ++g) {
        if (s.cls[g] != cls) continue;
        const float v = iou_normal(ab, s.box[g]);
        if (v > best) { best = v; arg = g; }
        if (v > 0.0f) atomicMax(&gt_best[(size_t)b * M + g], __float_as_int(v));   // IoU >= 0: integer order == float order
    }
    amax[(size_t)b * A + a] = best;
    aarg[(size_t)b * A + a] = arg;
}

// pass 2: labels, regression targets, regression weights, positives per frame
__global__ void __launch_bounds__(256) assign_pass2(const float* __restrict__ anchors, int64_t A, AssignSpec spec, const float* __restrict__ gt,
                                                    int M, int gt_stride, const float* __restrict__ amax, const int* __restrict__ aarg,
                                                    const int* __restrict__ gt_best, int* __restrict__ labels, float* __restrict__ reg_targets,
                                                    float* __restrict__ reg_weights, int* __restrict__ num_pos) {
    __shared__ FrameGt s;
    __shared__ int best_s[MAX_GT];
    __shared__ int pos_s;
    const int b = blockIdx.y;
    const float* fgt = gt + (size_t)b * M * gt_stride;
    for (int g = threadIdx.x; g < M; g += blockDim.x) best_s[g] = gt_best[(size_t)b * M + g];
    if (threadIdx.x == 0) pos_s = 0;
    stage_gt(fgt, M, gt_stride, s);
    const int64_t a = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (a < A) {
        const int t = (int)(a % spec.n_types);
        const int cls = spec.type_class[t];
        const float* anc = anchors + a * 7;
        const BevBox ab = aligned_bev(anc);
        const float best = amax[(size_t)b * A + a];
        const int arg = aarg[(size_t)b * A + a];
        int label;
        if (arg < 0) label = 0;                          // no gt of this class in the frame: labels[:] = 0
        else {
            bool forced = false;                         // anchor_by_gt_overlap == gt_to_anchor_max (0 maxima were set to -1)
            for (int g = 0; g < M && !forced; ++g)
                if (s.cls[g] == cls && best_s[g] > 0 && __float_as_int(iou_normal(ab, s.box[g])) == best_s[g]) forced = true;
            if (forced || best >= spec.matched[t]) label = s.cls[arg];
            else if (best < spec.unmatched[t]) label = 0;
            else label = -1;
        }
        labels[(size_t)b * A + a] = label;
        float* rt = reg_targets + ((size_t)b * A + a) * 7;
        if (label > 0) {
            // ResidualCoder.encode_torch of the matched gt
            const float* g = fgt + (size_t)arg * gt_stride;
            const float dxa = fmaxf(anc[3], 1e-5f), dya = fmaxf(anc[4], 1e-5f), dza = fmaxf(anc[5], 1e-5f);
            const float dxg = fmaxf(g[3], 1e-5f), dyg = fmaxf(g[4], 1e-5f), dzg = fmaxf(g[5], 1e-5f);
            const float diag = sqrtf(__fadd_rn(__fmul_rn(dxa, dxa), __fmul_rn(dya, dya)));
            rt[0] = __fdiv_rn(__fsub_rn(g[0], anc[0]), diag);
            rt[1] = __fdiv_rn(__fsub_rn(g[1], anc[1]), diag);
            rt[2] = __fdiv_rn(__fsub_rn(g[2], anc[2]), dza);
            rt[3] = logf(__fdiv_rn(dxg, dxa));
            rt[4] = logf(__fdiv_rn(dyg, dya));
            rt[5] = logf(__fdiv_rn(dzg, dza));
            rt[6] = __fsub_rn(g[6], anc[6]);
            atomicAdd(&pos_s, 1);
        } else {
#pragma unroll
            for (int j = 0; j < 7; ++j) rt[j] = 0.0f;
        }
        reg_weights[(size_t)b * A + a] = label > 0 ? 1.0f : 0.0f;
    }
    __syncthreads();
    if (threadIdx.x == 0 && pos_s) atomicAdd(&num_pos[b], pos_s);
}

// ---------------------------------------------------------------------------------------------------- losses
struct LossSpec {
    int n_class, num_dir_bins;
    float alpha, gamma, beta, dir_offset;
    float w_cls, w_loc, w_dir;          // LOSS_WEIGHTS / batch size
    float code_w[7];
};

constexpr int LOSS_THREADS = 256;

__global__ void __launch_bounds__(LOSS_THREADS) head_loss_kernel(const float* __restrict__ cls_preds, const float* __restrict__ box_preds,
                                                                 const float* __restrict__ dir_preds, const int* __restrict__ labels,
                                                                 const float* __restrict__ reg_targets, const float* __restrict__ anchors,
                                                                 const int* __restrict__ num_pos, int64_t A, LossSpec sp,
                                                                 float* __restrict__ g_cls, float* __restrict__ g_box, float* __restrict__ g_dir,
                                                                 double* __restrict__ partial) {
    __shared__ double red[3][LOSS_THREADS / 32];
    const int b = blockIdx.y;
    const int64_t a = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    double l_cls = 0.0, l_loc = 0.0, l_dir = 0.0;
    if (a < A) {
        const size_t i = (size_t)b * A + a;
        const int label = labels[i];
        const float norm = 1.0f / fmaxf((float)num_pos[b], 1.0f);
        // ---- classification: focal loss over the n_class logits, weight (label >= 0) / positives
        const float wc = label >= 0 ? norm : 0.0f;
        for (int c = 0; c < sp.n_class; ++c) {
            const float x = cls_preds[i * sp.n_class + c];
            float g = 0.0f;
            if (wc > 0.0f) {
                const float t = (label == c + 1) ? 1.0f : 0.0f;
                const float p = 1.0f / (1.0f + expf(-x));
                const float alpha_w = t * sp.alpha + (1.0f - t) * (1.0f - sp.alpha);
                const float pt = t * (1.0f - p) + (1.0f - t) * p;
                const float bce = fmaxf(x, 0.0f) - x * t + log1pf(expf(-fabsf(x)));
                const float fw = powf(pt, sp.gamma);
                l_cls += (double)(alpha_w * fw * bce * wc);
                // d/dx [pt^gamma * bce]: dpt/dx = (1 - 2t) p (1 - p), dbce/dx = p - t
                const float dpt = (1.0f - 2.0f * t) * p * (1.0f - p);
                const float dfw = pt > 0.0f ? sp.gamma * powf(pt, sp.gamma - 1.0f) * dpt : 0.0f;
                g = alpha_w * (dfw * bce + fw * (p - t)) * wc * sp.w_cls;
            }
            if (g_cls) g_cls[i * sp.n_class + c] = g;
        }
        // ---- localisation: smooth L1 on the code-weighted residual differences, sin(a - b) on the heading
        const float wr = label > 0 ? norm : 0.0f;
        const float* bp = box_preds + i * 7;
        const float* rt = reg_targets + i * 7;
        for (int j = 0; j < 7; ++j) {
            float g = 0.0f;
            if (wr > 0.0f) {
                const float pj = bp[j];
                float tj = rt[j];
                if (tj != tj) tj = pj;                                   // NaN targets are ignored (loss_utils.py:123)
                float d, dd;                                             // difference and its derivative w.r.t. the prediction
                if (j == 6) { d = sinf(pj) * cosf(tj) - cosf(pj) * sinf(tj); dd = cosf(pj) * cosf(tj) + sinf(pj) * sinf(tj); }
                else { d = pj - tj; dd = 1.0f; }
                if (rt[j] != rt[j]) { d = 0.0f; dd = 0.0f; }
                d *= sp.code_w[j];
                const float n = fabsf(d);
                float l, dl;
                if (sp.beta < 1e-5f) { l = n; dl = d > 0.0f ? 1.0f : (d < 0.0f ? -1.0f : 0.0f); }
                else if (n < sp.beta) { l = 0.5f * n * n / sp.beta; dl = d / sp.beta; }
                else { l = n - 0.5f * sp.beta; dl = d > 0.0f ? 1.0f : -1.0f; }
                l_loc += (double)(l * wr);
                g = dl * sp.code_w[j] * dd * wr * sp.w_loc;
            }
            if (g_box) g_box[i * 7 + j] = g;
        }
        // ---- direction bins: cross entropy against the bin of (target heading + anchor heading - dir_offset)
        if (dir_preds) {
            const int nb = sp.num_dir_bins;
            const float* dp = dir_preds + i * nb;
            if (wr > 0.0f) {
                const float TWO_PI = 6.28318530717958647692f;
                const float rot_gt = rt[6] + anchors[a * 7 + 6];
                const float v = rot_gt - sp.dir_offset;
                const float off = v - floorf(v / TWO_PI) * TWO_PI;        // limit_period(v, 0, 2 pi)
                int bin = (int)floorf(off / (TWO_PI / nb));
                bin = min(max(bin, 0), nb - 1);
                float mx = dp[0];
                for (int k = 1; k < nb; ++k) mx = fmaxf(mx, dp[k]);
                float se = 0.0f;
                for (int k = 0; k < nb; ++k) se += expf(dp[k] - mx);
                const float lse = mx + logf(se);
                l_dir += (double)((lse - dp[bin]) * wr);
                if (g_dir)
                    for (int k = 0; k < nb; ++k) g_dir[i * nb + k] = (expf(dp[k] - lse) - (k == bin ? 1.0f : 0.0f)) * wr * sp.w_dir;
            } else if (g_dir) {
                for (int k = 0; k < nb; ++k) g_dir[i * nb + k] = 0.0f;
            }
        }
    }
    // block reduction in a fixed order
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        l_cls += __shfl_down_sync(0xffffffffu, l_cls, o);
        l_loc += __shfl_down_sync(0xffffffffu, l_loc, o);
        l_dir += __shfl_down_sync(0xffffffffu, l_dir, o);
    }
    if (lane == 0) { red[0][warp] = l_cls; red[1][warp] = l_loc; red[2][warp] = l_dir; }
    __syncthreads();
    if (threadIdx.x < 3) {
        double s = 0.0;
        for (int w = 0; w < LOSS_THREADS / 32; ++w) s += red[threadIdx.x][w];
        partial[((size_t)blockIdx.y * gridDim.x + blockIdx.x) * 3 + threadIdx.x] = s;
    }
}

__global__ void __launch_bounds__(256) head_loss_reduce(const double* __restrict__ partial, int n_blocks, LossSpec sp, float* __restrict__ losses) {
    __shared__ double red[256];
    for (int k = 0; k < 3; ++k) {
        double s = 0.0;
        for (int i = threadIdx.x; i < n_blocks; i += 256) s += partial[(size_t)i * 3 + k];
        red[threadIdx.x] = s;
        __syncthreads();
        for (int o = 128; o > 0; o >>= 1) {
            if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
            __syncthreads();
        }
        if (threadIdx.x == 0) losses[k] = (float)(red[0] * (double)(k == 0 ? sp.w_cls : (k == 1 ? sp.w_loc : sp.w_dir)));
        __syncthreads();
    }
}

__global__ void __launch_bounds__(256) count_pos_kernel(const int* __restrict__ labels, int64_t A, int* __restrict__ num_pos) {
    __shared__ int s;
    if (threadIdx.x == 0) s = 0;
    __syncthreads();
    const int64_t a = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool pos = a < A && labels[(size_t)blockIdx.y * A + a] > 0;
    const unsigned int m = __ballot_sync(0xffffffffu, pos);
    if ((threadIdx.x & 31) == 0 && m) atomicAdd(&s, __popc(m));
    __syncthreads();
    if (threadIdx.x == 0 && s) atomicAdd(&num_pos[blockIdx.y], s);
}

}  // namespace

extern "C" int crb3d_assign_targets_workspace_bytes(int B, int64_t A, int M, size_t* bytes) {
    if (!bytes || B < 0 || A < 0 || M < 0) return CRB3D_ERR_ARG;
    const size_t n = (size_t)(B > 0 ? B : 1) * (size_t)(A > 0 ? A : 1);
    *bytes = crb3d_align(sizeof(float) * n) + crb3d_align(sizeof(int) * n) + crb3d_align(sizeof(int) * (size_t)(B > 0 ? B : 1) * (M > 0 ? M : 1));
    return CRB3D_OK;
}

// anchors (A,7) in the head's order (y, x, type); type_class / matched / unmatched: HOST arrays of n_types entries (1-based class
// and thresholds of every anchor type at a location); gt_boxes (B, M, gt_stride >= 8): [x,y,z,dx,dy,dz,heading,..,class], rows with
// class <= 0 are padding. Outputs: labels (B,A) int32 (-1 ignore, 0 background, class id), reg_targets (B,A,7), reg_weights (B,A),
// num_pos (B) int32 = positives per frame.
extern "C" int crb3d_assign_targets_axis_aligned(const float* anchors, int64_t A, int n_types, const int* type_class, const float* matched,
                                                 const float* unmatched, const float* gt_boxes, int B, int M, int gt_stride, int* labels,
                                                 float* reg_targets, float* reg_weights, int* num_pos, void* ws, size_t ws_bytes,
                                                 cudaStream_t stream) {
    if (A < 0 || B < 0 || M < 0 || n_types <= 0 || !type_class || !matched || !unmatched || gt_stride < 8) return CRB3D_ERR_ARG;
    if (n_types > MAX_TYPES || M > MAX_GT) return CRB3D_ERR_UNSUPPORTED;
    if (A == 0 || B == 0) return CRB3D_OK;
    if (!anchors || (M > 0 && !gt_boxes) || !labels || !reg_targets || !reg_weights || !num_pos) return CRB3D_ERR_ARG;
    AssignSpec spec;
    spec.n_types = n_types;
    for (int t = 0; t < n_types; ++t) { spec.type_class[t] = type_class[t]; spec.matched[t] = matched[t]; spec.unmatched[t] = unmatched[t]; }
    WsCursor c(ws, ws_bytes);
    float* amax = c.take<float>((size_t)B * A);
    int* aarg = c.take<int>((size_t)B * A);
    int* gt_best = c.take<int>((size_t)B * (M > 0 ? M : 1));
    if (!c.ok) return CRB3D_ERR_WORKSPACE;
    CRB3D_CUDA(cudaMemsetAsync(gt_best, 0, sizeof(int) * (size_t)B * (M > 0 ? M : 1), stream));
    CRB3D_CUDA(cudaMemsetAsync(num_pos, 0, sizeof(int) * B, stream));
    const dim3 grid((unsigned)crb3d_divup(A, 256), (unsigned)B);
    assign_pass1<<<grid, 256, 0, stream>>>(anchors, A, spec, gt_boxes, M, gt_stride, amax, aarg, gt_best);
    assign_pass2<<<grid, 256, 0, stream>>>(anchors, A, spec, gt_boxes, M, gt_stride, amax, aarg, gt_best, labels, reg_targets, reg_weights,
                                           num_pos);
    CRB3D_CHECK_LAUNCH();
    return CRB3D_OK;
}

extern "C" int crb3d_anchor_head_loss_workspace_bytes(int B, int64_t A, size_t* bytes) {
    if (!bytes || B < 0 || A < 0) return CRB3D_ERR_ARG;
    *bytes = crb3d_align(sizeof(double) * 3 * (size_t)(B > 0 ? B : 1) * (size_t)crb3d_divup(A > 0 ? A : 1, LOSS_THREADS)) +
             crb3d_align(sizeof(int) * (size_t)(B > 0 ? B : 1));
    return CRB3D_OK;
}

// cls_preds (B,A,n_class), box_preds (B,A,7), dir_preds (B,A,num_dir_bins) or null, labels (B,A) int32, reg_targets (B,A,7),
// anchors (A,7). code_weights: HOST float[7] or null (= ones). loss_weights3: HOST {cls, loc, dir} LOSS_WEIGHTS.
// losses (DEVICE float[3]) = {rpn_loss_cls, rpn_loss_loc, rpn_loss_dir} (their sum is rpn_loss); g_* (nullable) = d rpn_loss / d pred.
extern "C" int crb3d_anchor_head_loss(const float* cls_preds, const float* box_preds, const float* dir_preds, const int* labels,
                                      const float* reg_targets, const float* anchors, int B, int64_t A, int n_class, int num_dir_bins,
                                      const float* code_weights, float alpha, float gamma, float beta, float dir_offset,
                                      const float* loss_weights3, float* losses, float* g_cls, float* g_box, float* g_dir, void* ws,
                                      size_t ws_bytes, cudaStream_t stream) {
    if (B <= 0 || A <= 0 || n_class <= 0 || !cls_preds || !box_preds || !labels || !reg_targets || !anchors || !loss_weights3 || !losses)
        return CRB3D_ERR_ARG;
    if (dir_preds && (num_dir_bins <= 0 || num_dir_bins > 8)) return CRB3D_ERR_UNSUPPORTED;
    LossSpec sp;
    sp.n_class = n_class; sp.num_dir_bins = num_dir_bins;
    sp.alpha = alpha; sp.gamma = gamma; sp.beta = beta; sp.dir_offset = dir_offset;
    sp.w_cls = loss_weights3[0] / B; sp.w_loc = loss_weights3[1] / B; sp.w_dir = loss_weights3[2] / B;
    for (int j = 0; j < 7; ++j) sp.code_w[j] = code_weights ? code_weights[j] : 1.0f;
    const unsigned nbx = (unsigned)crb3d_divup(A, LOSS_THREADS);
    WsCursor c(ws, ws_bytes);
    double* partial = c.take<double>((size_t)3 * B * nbx);
    int* num_pos = c.take<int>(B);
    if (!c.ok) return CRB3D_ERR_WORKSPACE;
    CRB3D_CUDA(cudaMemsetAsync(num_pos, 0, sizeof(int) * B, stream));
    count_pos_kernel<<<dim3((unsigned)crb3d_divup(A, 256), (unsigned)B), 256, 0, stream>>>(labels, A, num_pos);
    head_loss_kernel<<<dim3(nbx, (unsigned)B), LOSS_THREADS, 0, stream>>>(cls_preds, box_preds, dir_preds, labels, reg_targets, anchors, num_pos,
                                                                        A, sp, g_cls, g_box, g_dir, partial);
    head_loss_reduce<<<1, 256, 0, stream>>>(partial, (int)(B * nbx), sp, losses);
    CRB3D_CHECK_LAUNCH();
    return CRB3D_OK;
}
