// Pipe-cost microbenchmark for the instruction kinds the fused simulator is made of (sm_100a): warp-instructions per cycle per
// scheduler (SMSP) for each kind alone and for the mixes that matter (packed FMA with wide integer multiplies, with LOP3, with MUFU).
// 8 independent chains per thread, 1024 threads per SM (8 warps per scheduler): throughput, not latency.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench/pipes tools/microbench/pipes.cu && tools/microbench/pipes
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 pk(float lo, float hi) { u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 r; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ float ffma(float a, float b, float c) { float r; asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r; }
__device__ __forceinline__ float fadd(float a, float b) { float r; asm volatile("add.rn.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ float fmnmx(float a, float b) { float r; asm volatile("max.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ float ex2(float a) { float r; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(a)); return r; }
__device__ __forceinline__ unsigned lop(unsigned a, unsigned b, unsigned c) { unsigned r; asm volatile("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
__device__ __forceinline__ void mulwide(unsigned x, unsigned& hi, unsigned& lo) {
    asm volatile("{\n\t.reg .b64 t;\n\tmul.wide.u32 t, %2, 0xD2511F53;\n\tmov.b64 {%1, %0}, t;\n\t}" : "=r"(hi), "=r"(lo) : "r"(x));
}
__device__ __forceinline__ unsigned mulhi(unsigned x) { unsigned r; asm volatile("mul.hi.u32 %0, %1, 0xD2511F53;" : "=r"(r) : "r"(x)); return r; }
__device__ __forceinline__ unsigned mullo(unsigned x, unsigned y) { unsigned r; asm volatile("mad.lo.u32 %0, %1, %2, %2;" : "=r"(r) : "r"(x), "r"(y)); return r; }

enum { FFMA, FFMA2_UR, FFMA2_R, IMADW, IMADHI, IMADLO, LOP3, MUFU, FADD, FMNMX, MIX_F2_IW, MIX_F2_LOP, MIX_F2_MUFU, MIX_F2_FFMA, MIX_IW_LOP, MIX_F_LOP, MIX_2F2_LOP, MIX_SIM, MIX_CRITIC, N_MODES };
const char* NAMES[] = {"FFMA R,R,UR,R", "FFMA2 pair x (broadcast R) x UR pair", "FFMA2 pair x R x R pair", "IMAD.WIDE.U32 R,R,imm", "IMAD.HI.U32", "IMAD (lo)",
                       "LOP3", "MUFU.EX2", "FADD", "FMNMX", "mix 1 FFMA2 : 1 IMAD.WIDE", "mix 1 FFMA2 : 1 LOP3", "mix 4 FFMA2 : 1 MUFU", "mix 1 FFMA2 : 1 FFMA",
                       "mix 1 IMAD.WIDE : 1 LOP3 (Philox round)", "mix 1 FFMA : 1 LOP3", "mix 2 FFMA2 : 1 LOP3",
                       "mix of the fused kernel (15 FFMA2 : 22 scalar FP : 6 IMAD.WIDE : 13 ALU : 7 MUFU = 612 : 908 : 252 : 560 : 295)",
                       "mix of the critic kernel without its SHFL / LDC (8 FFMA2 : 2 scalar FP : 5 ALU)"};
// instructions issued per inner iteration (per thread) for each mode
__host__ __device__ constexpr int per_iter(int m) {
    return m == MIX_F2_MUFU ? 8 + 2 : m == MIX_SIM ? 63 : m == MIX_CRITIC ? 15 : m == MIX_2F2_LOP ? 24 : m >= MIX_F2_IW ? 16 : 8;
}
template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, int iters, float a, float b, u64 w2) {
    float v[8];
    u64 p[8];
    unsigned x[8], y[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { v[j] = (float)(threadIdx.x + j); p[j] = pk(v[j], v[j] + 1.f); x[j] = threadIdx.x * 17 + j; y[j] = x[j] * 3; }
    const u64 wr = pk(v[0] * 1e-9f + 0.999f, 0.998f);      // a weight pair in vector registers
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                if (MODE == FFMA) v[j] = ffma(v[j], a, b);
                if (MODE == FFMA2_UR) p[j] = fma2(pk(v[j], v[j]), w2, p[j]);
                if (MODE == FFMA2_R) p[j] = fma2(pk(v[j], v[j]), wr, p[j]);
                if (MODE == IMADW) { unsigned h, l; mulwide(x[j], h, l); x[j] = h ^ l; }
                if (MODE == IMADHI) x[j] = mulhi(x[j]);
                if (MODE == IMADLO) x[j] = mullo(x[j], y[j]);
                if (MODE == LOP3) x[j] = lop(x[j], y[j], 0x9E3779B9u);
                if (MODE == MUFU) v[j] = ex2(v[j]);
                if (MODE == FADD) v[j] = fadd(v[j], a);
                if (MODE == FMNMX) v[j] = fmnmx(v[j], a);
                if (MODE == MIX_F2_IW) { p[j] = fma2(pk(v[j], v[j]), w2, p[j]); unsigned h, l; mulwide(x[j], h, l); x[j] = h; y[j] = l; }
                if (MODE == MIX_F2_LOP) { p[j] = fma2(pk(v[j], v[j]), w2, p[j]); x[j] = lop(x[j], y[j], 0x9E3779B9u); }
                if (MODE == MIX_F2_FFMA) { p[j] = fma2(pk(v[j], v[j]), w2, p[j]); v[j] = ffma(v[j], a, b); }
                if (MODE == MIX_F_LOP) { v[j] = ffma(v[j], a, b); x[j] = lop(x[j], y[j], 0x9E3779B9u); }
                if (MODE == MIX_2F2_LOP) { p[j] = fma2(pk(v[j], v[j]), w2, p[j]); p[j] = fma2(pk(v[j], v[j]), w2, p[j]); x[j] = lop(x[j], y[j], 0x9E3779B9u); }
                if (MODE == MIX_IW_LOP) { unsigned h, l; mulwide(x[j], h, l); x[j] = lop(h, l, y[j]); }
            }
            if (MODE == MIX_F2_MUFU) {
#pragma unroll
                for (int j = 0; j < 8; ++j) p[j] = fma2(pk(v[j], v[j]), w2, p[j]);
                v[r & 7] = ex2(v[r & 7]);
                v[(r + 4) & 7] = ex2(v[(r + 4) & 7]);
            }
            if (MODE == MIX_CRITIC) {
#pragma unroll
                for (int j = 0; j < 8; ++j) p[j] = fma2(pk(v[j], v[j]), w2, p[j]);
                v[0] = ffma(v[0], a, b); v[1] = fadd(v[1], a);
#pragma unroll
                for (int j = 0; j < 5; ++j) x[j] = lop(x[j], y[j], 0x9E3779B9u);
            }
            if (MODE == MIX_SIM) {       // 63 instructions in the kernel's proportions, kinds interleaved
#pragma unroll
                for (int j = 0; j < 8; ++j) { p[j] = fma2(pk(v[j], v[j]), w2, p[j]); v[j] = ffma(v[j], a, b); }
#pragma unroll
                for (int j = 0; j < 6; ++j) { unsigned h, l; mulwide(x[j], h, l); x[j] = lop(h, l, y[j]); v[j] = fadd(v[j], a); }
#pragma unroll
                for (int j = 0; j < 7; ++j) { p[j] = fma2(pk(v[j], v[j]), w2, p[j]); v[j] = ex2(v[j]); x[j] = lop(x[j], y[j], 0x9E3779B9u); }
#pragma unroll
                for (int j = 0; j < 8; ++j) v[j] = ffma(v[j], a, b);
            }
        }
    }
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) { float lo, hi; asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(p[j])); s += v[j] + lo + hi + (float)x[j] + (float)y[j]; }
    if (s == 123.456f) out[0] = s;
}
template <int MODE>
void run(int sms, double ghz, int wps) {
    float* out; cudaMalloc(&out, 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 2000, blocks = sms * wps / 2;       // 256-thread blocks: 2 warps per scheduler each
    const u64 w2 = 0x3F7FBE773F7FDF3BULL;
    k<MODE><<<blocks, 256>>>(out, 10, 0.999f, 0.001f, w2);
    cudaEventRecord(e0); k<MODE><<<blocks, 256>>>(out, iters, 0.999f, 0.001f, w2); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double inst_per_warp = (double)iters * 8 * per_iter(MODE);
    const double warps_per_smsp = wps, cycles = ms * 1e-3 * ghz * 1e9;
    printf("%-120s %6.3f warp-inst / cycle / scheduler   (%.3f ms)\n", NAMES[MODE], inst_per_warp * warps_per_smsp / cycles, ms);
    cudaFree(out);
}
template <int M> struct All { static void go(int sms, double ghz, int wps) { run<M>(sms, ghz, wps); All<M + 1>::go(sms, ghz, wps); } };
template <> struct All<N_MODES> { static void go(int, double, int) {} };
int main() {
    int sms, khz; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0); cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    printf("SMs %d, clock %.0f MHz (rates assume the maximum clock)\n", sms, khz / 1e3);
    for (int wps : {8, 4}) { printf("-- %d warps per scheduler\n", wps); All<0>::go(sms, khz / 1e6, wps); }
    return 0;
}
