// FP32 FMA issue-rate probe for sm_100a: scalar FFMA vs packed FFMA2 (fma.rn.f32x2), 16 independent chains per thread.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/fma_peak scripts/fma_peak.cu && ./scripts/fma_peak
#include <cstdio>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(256) scalar_fma(float* out, int iters, float a, float b) {
    float acc[16];
    for (int i = 0; i < 16; ++i) acc[i] = threadIdx.x * 0.001f + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) acc[i] = fmaf(acc[i], a, b);
    }
    float s = 0.f;
    for (int i = 0; i < 16; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void __launch_bounds__(256) packed_fma(float* out, int iters, float a, float b) {
    unsigned long long acc[16], a2, b2;
    asm("mov.b64 %0, {%1, %1};" : "=l"(a2) : "f"(a));
    asm("mov.b64 %0, {%1, %1};" : "=l"(b2) : "f"(b));
    for (int i = 0; i < 16; ++i) { float v = threadIdx.x * 0.001f + i; asm("mov.b64 %0, {%1, %1};" : "=l"(acc[i]) : "f"(v)); }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(acc[i]) : "l"(a2), "l"(b2));
    }
    unsigned long long s = 0;
    for (int i = 0; i < 16; ++i) s ^= acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = __uint_as_float((unsigned)s);
}
int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    const int sms = p.multiProcessorCount, grid = sms * 8, iters = 20000;
    float* out; cudaMalloc(&out, grid * 256 * sizeof(float));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int mode = 0; mode < 2; ++mode) {
        for (int rep = 0; rep < 2; ++rep) {
            cudaEventRecord(e0);
            if (mode == 0) scalar_fma<<<grid, 256>>>(out, iters, 1.0001f, 0.5f); else packed_fma<<<grid, 256>>>(out, iters, 1.0001f, 0.5f);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
        }
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        const double fma = (double)grid * 256 * iters * 16 * (mode == 0 ? 1 : 2);
        printf("%s: %.3f ms, %.1f TFLOP/s (2 flop per fma), %.1f fma/clk/SM at %d MHz\n", mode == 0 ? "scalar FFMA " : "packed FFMA2", ms,
               2 * fma / ms / 1e9, fma / (ms * 1e-3) / sms / (p.clockRate * 1e3), p.clockRate / 1000);
    }
    return 0;
}
