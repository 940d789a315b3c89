// Sparse 3D convolution, exact-fp32 SIMT path (output-stationary gather -> FFMA -> single store).
//
// Replaces the external spconv-cu113==2.1.21 `indice_conv` / implicit-GEMM forward + backward reached from
//   pcdet/models/backbones_3d/spconv_backbone.py:77-117 (12 conv layers) and autograd from crb_sampling.py:205.
// Contract:  out[o,:] = sum_k  in[nbr[k][o], :] @ W[:, k, :]^T   (+ bias), W is [C_out, K, C_in]
// (the spconv-2.x checkpoint layout, detector3d_template.py:455-484).
//
// This kernel is the exact-fp32 variant (used for C_in < 16, odd channel counts and as the on-device fp32
// reference for the tcgen05 TF32 kernel in spconv_tc.cu). Sum order is fixed: k ascending, then c_in ascending.
// Optional fused epilogue: y = relu(acc * scale[c] + shift[c])  (eval-mode BatchNorm1d + ReLU folded).
#include "common.cuh"

namespace {

constexpr int TM = 64;        // output rows per CTA
constexpr int THREADS = 256;  // 16 column groups x 16 row groups
constexpr int RM = 4;         // rows per thread
constexpr int CK = 64;        // C_in chunk staged in smem

template <int RN>
__global__ void __launch_bounds__(THREADS) spconv_fwd_simt(const float* __restrict__ feat, const int* __restrict__ nbr,
                                                           const float* __restrict__ weight, int n_out, int K, int cin,
                                                           int cout, int w_k_stride, int w_co_stride, int w_ci_stride,
                                                           const int* __restrict__ kmap, const float* __restrict__ scale,
                                                           const float* __restrict__ shift, int relu,
                                                           float* __restrict__ out, const int* __restrict__ n_dev) {
    // n_dev (nullable): device-side row count; n_out is then only the capacity / row stride of the neighbour table
    const int nv = n_dev ? min(n_out, *n_dev) : n_out;
    if ((int)blockIdx.x * TM >= nv) return;
    extern __shared__ float smem[];
    constexpr int COUTP = 16 * RN + 1;
    float* As = smem;                   // [TM][CK + 4]
    float* Ws = smem + TM * (CK + 4);   // [CK][COUTP]
    __shared__ int rows[TM];

    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int row0 = blockIdx.x * TM;
    float acc[RM][RN];
#pragma unroll
    for (int r = 0; r < RM; ++r)
#pragma unroll
        for (int j = 0; j < RN; ++j) acc[r][j] = 0.0f;

    for (int k = 0; k < K; ++k) {
        int my = -1;
        if (tid < TM) {
            int o = row0 + tid;
            my = (o < nv) ? __ldg(&nbr[(size_t)k * n_out + o]) : -1;
            rows[tid] = my;
        }
        if (!__syncthreads_or(my >= 0)) continue;  // nobody in this tile has a neighbour at offset k
        const int kw = kmap ? kmap[k] : k;
        for (int c0 = 0; c0 < cin; c0 += CK) {
            const int cw = min(CK, cin - c0);
            // gather A: TM rows x cw floats (zero for missing neighbours)
            if ((cin & 3) == 0) {
                const int vec_per_row = cw >> 2;
                for (int t = tid; t < TM * vec_per_row; t += THREADS) {
                    int r = t / vec_per_row, v = t - r * vec_per_row;
                    int src = rows[r];
                    float4 val = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (src >= 0) val = __ldg(reinterpret_cast<const float4*>(feat + (size_t)src * cin + c0) + v);
                    *reinterpret_cast<float4*>(&As[r * (CK + 4) + v * 4]) = val;
                }
            } else {
                for (int t = tid; t < TM * cw; t += THREADS) {
                    int r = t / cw, v = t - r * cw;
                    int src = rows[r];
                    As[r * (CK + 4) + v] = (src >= 0) ? __ldg(feat + (size_t)src * cin + c0 + v) : 0.0f;
                }
            }
            // stage W[kw] chunk transposed to [ci][co]
            for (int t = tid; t < cw * cout; t += THREADS) {
                int co = t / cw, ci = t - co * cw;
                Ws[ci * COUTP + co] = __ldg(weight + (size_t)co * w_co_stride + (size_t)kw * w_k_stride + (size_t)(c0 + ci) * w_ci_stride);
            }
            __syncthreads();
            for (int ci = 0; ci < cw; ++ci) {
                float a[RM], w[RN];
#pragma unroll
                for (int r = 0; r < RM; ++r) a[r] = As[(ty * RM + r) * (CK + 4) + ci];
#pragma unroll
                for (int j = 0; j < RN; ++j) w[j] = (tx + 16 * j < cout) ? Ws[ci * COUTP + tx + 16 * j] : 0.0f;
#pragma unroll
                for (int r = 0; r < RM; ++r)
#pragma unroll
                    for (int j = 0; j < RN; ++j) acc[r][j] = fmaf(a[r], w[j], acc[r][j]);
            }
            __syncthreads();
        }
    }
#pragma unroll
    for (int r = 0; r < RM; ++r) {
        int o = row0 + ty * RM + r;
        if (o >= nv) continue;
#pragma unroll
        for (int j = 0; j < RN; ++j) {
            int c = tx + 16 * j;
            if (c >= cout) continue;
            float v = acc[r][j];
            if (scale) v = fmaf(v, __ldg(&scale[c]), shift ? __ldg(&shift[c]) : 0.0f);
            else if (shift) v += __ldg(&shift[c]);
            if (relu) v = fmaxf(v, 0.0f);
            out[(size_t)o * cout + c] = v;
        }
    }
}

// Weight gradient: dW[co][k][ci] = sum_o  dY[o][co] * X[nbr[k][o]][ci].
// grid (K, row chunks); each CTA reduces its chunk in registers and writes a partial; a second kernel sums the
// partials in chunk order (deterministic, no float atomics).
constexpr int WG_ROWS = 32;  // rows staged per step
template <int RN>
__global__ void __launch_bounds__(THREADS) spconv_wgrad_partial(const float* __restrict__ feat,
                                                                const float* __restrict__ dout,
                                                                const int* __restrict__ nbr, int n_out, int cin,
                                                                int cout, int rows_per_chunk,
                                                                float* __restrict__ partial /*[chunks][K][cout][cin]*/) {
    // thread tile: ci = tx + 16*a (a < RA), co = ty + 16*j (j < RN); supports cin <= 64, cout <= 16*RN
    constexpr int RA = 4;
    extern __shared__ float smem[];
    float* Xs = smem;                      // [WG_ROWS][64 + 1]
    float* Ys = smem + WG_ROWS * 65;       // [WG_ROWS][16*RN + 1]
    __shared__ int rows[WG_ROWS];
    constexpr int YP = 16 * RN + 1;
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int k = blockIdx.x, chunk = blockIdx.y, K = gridDim.x;
    const int r_begin = chunk * rows_per_chunk, r_end = min(n_out, r_begin + rows_per_chunk);
    float acc[RA][RN];
#pragma unroll
    for (int a = 0; a < RA; ++a)
#pragma unroll
        for (int j = 0; j < RN; ++j) acc[a][j] = 0.0f;
    for (int r0 = r_begin; r0 < r_end; r0 += WG_ROWS) {
        int my = -1;
        if (tid < WG_ROWS) {
            int o = r0 + tid;
            my = (o < r_end) ? __ldg(&nbr[(size_t)k * n_out + o]) : -1;
            rows[tid] = my;
        }
        if (!__syncthreads_or(my >= 0)) continue;
        for (int t = tid; t < WG_ROWS * cin; t += THREADS) {
            int r = t / cin, c = t - r * cin;
            int src = rows[r];
            Xs[r * 65 + c] = (src >= 0) ? __ldg(feat + (size_t)src * cin + c) : 0.0f;
        }
        for (int t = tid; t < WG_ROWS * cout; t += THREADS) {
            int r = t / cout, c = t - r * cout;
            Ys[r * YP + c] = (rows[r] >= 0) ? __ldg(dout + (size_t)(r0 + r) * cout + c) : 0.0f;
        }
        __syncthreads();
        for (int r = 0; r < WG_ROWS; ++r) {
            float x[RA], y[RN];
#pragma unroll
            for (int a = 0; a < RA; ++a) x[a] = (tx + 16 * a < cin) ? Xs[r * 65 + tx + 16 * a] : 0.0f;
#pragma unroll
            for (int j = 0; j < RN; ++j) y[j] = (ty + 16 * j < cout) ? Ys[r * YP + ty + 16 * j] : 0.0f;
#pragma unroll
            for (int a = 0; a < RA; ++a)
#pragma unroll
                for (int j = 0; j < RN; ++j) acc[a][j] = fmaf(x[a], y[j], acc[a][j]);
        }
        __syncthreads();
    }
    float* dst = partial + ((size_t)chunk * K + k) * cout * cin;
#pragma unroll
    for (int a = 0; a < RA; ++a)
#pragma unroll
        for (int j = 0; j < RN; ++j) {
            int ci = tx + 16 * a, co = ty + 16 * j;
            if (ci < cin && co < cout) dst[(size_t)co * cin + ci] = acc[a][j];
        }
}

__global__ void __launch_bounds__(256) spconv_wgrad_reduce(const float* __restrict__ partial, int chunks, int K, int cin,
                                                           int cout, int accumulate, float* __restrict__ dw /*[cout][K][cin]*/) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    int total = K * cout * cin;
    if (t >= total) return;
    int ci = t % cin, co = (t / cin) % cout, k = t / (cin * cout);
    float s = 0.0f;
    for (int c = 0; c < chunks; ++c) s += partial[(size_t)c * total + t];
    size_t d = ((size_t)co * K + k) * cin + ci;
    dw[d] = accumulate ? dw[d] + s : s;
}

template <int RN>
int launch_fwd(const float* feat, const int* nbr, const float* weight, int n_out, int K, int cin, int cout,
               int w_k_stride, int w_co_stride, int w_ci_stride, const int* kmap, const float* scale, const float* shift,
               int relu, float* out, const int* n_dev, cudaStream_t stream) {
    size_t smem = sizeof(float) * (TM * (CK + 4) + CK * (16 * RN + 1));
    auto kern = spconv_fwd_simt<RN>;
    if (smem > 48 * 1024) CRB3D_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<(unsigned)crb3d_divup(n_out, TM), THREADS, smem, stream>>>(feat, nbr, weight, n_out, K, cin, cout, w_k_stride,
                                                                     w_co_stride, w_ci_stride, kmap, scale, shift, relu, out, n_dev);
    CRB3D_CHECK_LAUNCH();
    return CRB3D_OK;
}

}  // namespace

// weight element (co, k, ci) lives at weight[co * w_co_stride + k * w_k_stride + ci].
// weight element (co, k, ci) at co*w_co_stride + k*w_k_stride + ci*w_ci_stride.
// forward: (K*cin, cin, 1) for the spconv layout [C_out, K, C_in]; input-gradient: call with cin<->cout swapped,
// the transposed table and strides (1, cin, K*cin).
// n_dev (device, optional): true row count when n_out is only the capacity / stride of a static neighbour table.
// kmap (device, optional): offset k of the table uses weight slice kmap[k] (used to run dX of a SubM conv on the
// forward table with flipped offsets).
extern "C" int crb3d_spconv_forward_f32(const float* feat, const int* nbr, const float* weight, int n_out, int K,
                                        int cin, int cout, int64_t w_co_stride, int64_t w_k_stride, int64_t w_ci_stride,
                                        const int* kmap,
                                        const float* scale, const float* shift, int relu, float* out, const int* n_dev,
                                        cudaStream_t stream) {
    if (n_out < 0 || K <= 0 || cin <= 0 || cout <= 0 || !weight || !out) return CRB3D_ERR_ARG;
    if (n_out == 0) return CRB3D_OK;
    if (!feat || !nbr) return CRB3D_ERR_ARG;
    if (cout > 256) return CRB3D_ERR_UNSUPPORTED;
#define ARGS feat, nbr, weight, n_out, K, cin, cout, (int)w_k_stride, (int)w_co_stride, (int)w_ci_stride, kmap, scale, shift, relu, out, n_dev, stream
    if (cout <= 16) return launch_fwd<1>(ARGS);
    if (cout <= 32) return launch_fwd<2>(ARGS);
    if (cout <= 64) return launch_fwd<4>(ARGS);
    if (cout <= 128) return launch_fwd<8>(ARGS);
    return launch_fwd<16>(ARGS);
#undef ARGS
}

extern "C" int crb3d_spconv_wgrad_workspace_bytes(int n_out, int K, int cin, int cout, size_t* bytes) {
    if (!bytes || n_out < 0 || K <= 0 || cin <= 0 || cout <= 0) return CRB3D_ERR_ARG;
    int rows_per_chunk = 4096;
    int chunks = (int)crb3d_divup(n_out > 0 ? n_out : 1, rows_per_chunk);
    *bytes = crb3d_align(sizeof(float) * (size_t)chunks * K * cin * cout);
    return CRB3D_OK;
}

// dW (spconv layout [C_out, K, C_in]); accumulate != 0 adds into dw (autograd .grad accumulation).
extern "C" int crb3d_spconv_wgrad_f32(const float* feat, const float* dout, const int* nbr, int n_out, int K, int cin,
                                      int cout, int accumulate, float* dw, void* ws, size_t ws_bytes,
                                      cudaStream_t stream) {
    if (n_out < 0 || K <= 0 || cin <= 0 || cout <= 0 || !dw) return CRB3D_ERR_ARG;
    if (cin > 64 || cout > 128) return CRB3D_ERR_UNSUPPORTED;
    const int rows_per_chunk = 4096;
    const int chunks = (int)crb3d_divup(n_out > 0 ? n_out : 1, rows_per_chunk);
    WsCursor c(ws, ws_bytes);
    float* partial = c.take<float>((size_t)chunks * K * cin * cout);
    if (!c.ok) return CRB3D_ERR_WORKSPACE;
    if (n_out == 0) {
        if (!accumulate) CRB3D_CUDA(cudaMemsetAsync(dw, 0, sizeof(float) * (size_t)K * cin * cout, stream));
        return CRB3D_OK;
    }
    dim3 grid(K, chunks);
    if (cout <= 16) {
        size_t smem = sizeof(float) * (WG_ROWS * 65 + WG_ROWS * 17);
        spconv_wgrad_partial<1><<<grid, THREADS, smem, stream>>>(feat, dout, nbr, n_out, cin, cout, rows_per_chunk, partial);
    } else if (cout <= 32) {
        size_t smem = sizeof(float) * (WG_ROWS * 65 + WG_ROWS * 33);
        spconv_wgrad_partial<2><<<grid, THREADS, smem, stream>>>(feat, dout, nbr, n_out, cin, cout, rows_per_chunk, partial);
    } else if (cout <= 64) {
        size_t smem = sizeof(float) * (WG_ROWS * 65 + WG_ROWS * 65);
        spconv_wgrad_partial<4><<<grid, THREADS, smem, stream>>>(feat, dout, nbr, n_out, cin, cout, rows_per_chunk, partial);
    } else {
        size_t smem = sizeof(float) * (WG_ROWS * 65 + WG_ROWS * 129);
        spconv_wgrad_partial<8><<<grid, THREADS, smem, stream>>>(feat, dout, nbr, n_out, cin, cout, rows_per_chunk, partial);
    }
    int total = K * cin * cout;
    spconv_wgrad_reduce<<<(unsigned)crb3d_divup(total, 256), 256, 0, stream>>>(partial, chunks, K, cin, cout, accumulate, dw);
    CRB3D_CHECK_LAUNCH();
    return CRB3D_OK;
}
