// Microbenchmark: FFMA vs packed FFMA2 issue throughput on sm_100a (is f32x2 worth restructuring for?).
#include <cuda_runtime.h>
#include <cstdio>
template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, int iters, float s)
{
    float2 a[8];
    for (int j = 0; j < 8; ++j) a[j] = make_float2(threadIdx.x * 1e-3f + j, threadIdx.x * 2e-3f - j);
    const float2 m = make_float2(s, s * 0.5f), c = make_float2(1e-3f, 2e-3f);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            if (MODE == 0) { a[j].x = fmaf(a[j].x, m.x, c.x); a[j].y = fmaf(a[j].y, m.y, c.y); }
            else a[j] = __ffma2_rn(a[j], m, c);
        }
    }
    float acc = 0;
    for (int j = 0; j < 8; ++j) acc += a[j].x + a[j].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
int main()
{
    float* out; cudaMalloc(&out, 148 * 8 * 256 * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 20000;
    for (int mode = 0; mode < 2; ++mode)
        for (int rep = 0; rep < 3; ++rep) {
            cudaEventRecord(e0);
            if (mode == 0) k<0><<<148 * 8, 256>>>(out, iters, 0.999f); else k<1><<<148 * 8, 256>>>(out, iters, 0.999f);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            double fma = 148.0 * 8 * 256 * 16.0 * iters;
            printf("%s: %.3f ms  %.2f TFMA/s  (%.1f FMA/clk/SM at 1.965 GHz)\n", mode ? "FFMA2" : "FFMA ", ms, fma / ms / 1e9,
                   fma / (ms * 1e-3) / 148 / 1.965e9);
        }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
