// Probe: semantics of cp.async.bulk.tensor.2d ... tile::gather4 on sm_100a (box dims, OOB rows, swizzle, tx bytes).
// nvcc -gencode arch=compute_100a,code=sm_100a -o gather4_probe gather4_probe.cu && ./gather4_probe
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <vector>

__global__ void probe(const __grid_constant__ CUtensorMap map, int r0, int r1, int r2, int r3, int col, uint32_t expect, float* out, int* status) {
    __shared__ __align__(1024) float buf[4 * 32 * 2];
    __shared__ uint64_t bar;
    uint32_t bar_a = (uint32_t)__cvta_generic_to_shared(&bar), buf_a = (uint32_t)__cvta_generic_to_shared(buf);
    if (threadIdx.x == 0) {
        for (int i = 0; i < 256; ++i) buf[i] = -777.f;
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_a));
        asm volatile("fence.mbarrier_init.release.cluster;");
        asm volatile("fence.proxy.async.shared::cta;");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_a), "r"(expect));
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
                     ::"r"(buf_a), "l"(reinterpret_cast<uint64_t>(&map)), "r"(col), "r"(r0), "r"(r1), "r"(r2), "r"(r3), "r"(bar_a) : "memory");
        int ok = 0;
        for (int spin = 0; spin < 2000000 && !ok; ++spin) {
            uint32_t p;
            asm volatile("{.reg .pred q; mbarrier.try_wait.parity.shared::cta.b64 q, [%1], 0; selp.u32 %0, 1, 0, q;}" : "=r"(p) : "r"(bar_a));
            ok = p;
        }
        *status = ok;
        for (int i = 0; i < 256; ++i) out[i] = buf[i];
    }
}

int main() {
    const int n = 1000, C = 64;
    std::vector<float> h(n * C);
    for (int i = 0; i < n; ++i) for (int c = 0; c < C; ++c) h[i * C + c] = i + c * 0.001f;
    float* d; cudaMalloc(&d, h.size() * 4); cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
    float* out; cudaMalloc(&out, 256 * 4); int* st; cudaMalloc(&st, 4);
    void* fn = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    auto enc = (CUresult(*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                            const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill))fn;
    for (int boxrows : {1, 4}) for (int sw : {0, 1}) {
        CUtensorMap map; cuuint64_t dims[2] = {(cuuint64_t)C, (cuuint64_t)n}; cuuint64_t strides[1] = {(cuuint64_t)C * 4};
        cuuint32_t box[2] = {32, (cuuint32_t)boxrows}; cuuint32_t es[2] = {1, 1};
        CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         sw ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        printf("boxrows=%d swizzle=%d encode=%d\n", boxrows, sw, (int)r);
        if (r) continue;
        for (uint32_t expect : {512u}) {
            cudaMemset(st, 0, 4);
            probe<<<1, 32>>>(map, 5, -1, 7, 100000, 32, expect, out, st);
            cudaError_t e = cudaDeviceSynchronize();
            float ho[256]; int hs = -1; cudaMemcpy(ho, out, 1024, cudaMemcpyDeviceToHost); cudaMemcpy(&hs, st, 4, cudaMemcpyDeviceToHost);
            printf("  expect=%u err=%s barrier_done=%d\n", expect, cudaGetErrorString(e), hs);
            for (int row = 0; row < 4; ++row) {
                printf("   smem row %d:", row);
                for (int c = 0; c < 32; c += 4) printf(" %.3f", ho[row * 32 + c]);
                printf("\n");
            }
            if (e != cudaSuccess) return 0;
        }
    }
    return 0;
}
