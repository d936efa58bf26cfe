// Issue-rate microbenchmark: scalar FFMA vs packed FFMA2 (fma.rn.f32x2, sm_100a), and both mixed with ALU work.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench/ffma2 tools/microbench/ffma2.cu
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ unsigned long long pk(float lo, float hi) { unsigned long long r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void upk(unsigned long long v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c) {
    unsigned long long r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r;
}
template <int MODE>   // 0 scalar FFMA (R,UR,R), 1 FFMA2 (pair x broadcast weight), 2 scalar + LOP3 mix 1:1, 3 FFMA2 + LOP3 mix (1 FFMA2 : 2 LOP3)
__global__ void __launch_bounds__(256) k(float* out, int iters, float a, float b) {
    float v[16];
    unsigned x[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) x[j] = threadIdx.x * 17 + j;
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = (float)(threadIdx.x + j);
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            if (MODE == 0 || MODE == 2) {
#pragma unroll
                for (int j = 0; j < 16; ++j) v[j] = fmaf(v[j], a, b);
            } else {
#pragma unroll
                for (int j = 0; j < 16; j += 2) { unsigned long long p = fma2(pk(v[j], v[j + 1]), pk(a, a), pk(b, b)); upk(p, v[j], v[j + 1]); }
            }
            if (MODE >= 2) {
#pragma unroll
                for (int j = 0; j < 16; ++j) x[j] = (x[j] ^ (x[j] << 1)) ^ 0x9E3779B9u;      // 16 independent LOP3/SHF chains
            }
        }
    }
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < 16; ++j) s += (float)x[j];
#pragma unroll
    for (int j = 0; j < 16; ++j) s += v[j];
    if (s == 123.456f) out[0] = s;
}
template <int MODE> double run(int sms, int iters) {
    float* out; cudaMalloc(&out, 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<sms * 8, 256>>>(out, 10, 0.999f, 0.001f);
    cudaEventRecord(e0); k<MODE><<<sms * 8, 256>>>(out, iters, 0.999f, 0.001f); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    return 2.0 * 16 * 8 * (double)iters * 256 * sms * 8 / (ms * 1e-3) / 1e12;
}
int main() {
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    printf("scalar FFMA            %.1f TFLOP/s\n", run<0>(sms, 4000));
    printf("packed FFMA2           %.1f TFLOP/s\n", run<1>(sms, 4000));
    printf("scalar FFMA + ALU mix  %.1f TFLOP/s (FMA flops only)\n", run<2>(sms, 2000));
    printf("packed FFMA2 + ALU mix %.1f TFLOP/s (FMA flops only)\n", run<3>(sms, 2000));
    return 0;
}
