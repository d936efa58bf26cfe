// libofdmgan internals shared by the translation units: error plumbing, launch geometry, the per-stream
// constant-memory weight slots, Philox4x32-10 and the warp transpose-reduce used by every gradient kernel.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/ofdmgan.h"

#define OG_CHECK(expr)                                   \
    do {                                                 \
        cudaError_t _e = (expr);                         \
        if (_e != cudaSuccess) return (int)_e;           \
    } while (0)

#define OG_THREADS 128          // frames per CTA tile: one frame per thread
#define OG_NSLOT 1              // one scratch set per device (calls are serialised by CallGuard)

namespace og {

// ---- launch geometry: persistent grids sized as a multiple of the SM count (148 on B200) ---------------------
struct DeviceInfo {
    int device = -1;
    int sms = 0;
};
const DeviceInfo& device_info(int* err);
// CTAs for a frame-parallel kernel over B frames with `per_sm` resident CTAs per SM
int grid_for(int64_t B, int threads, int per_sm);

// ---- call serialisation ------------------------------------------------------------------------------------
// Every kernel reads its weights as immediate constant-bank operands from one fixed __constant__ image per network
// (weights.cuh); the image is refreshed stream-ordered before each launch (prep kernel -> staging ->
// cudaMemcpyToSymbolAsync D2D), so CUDA graphs replay with the current weights.  The images and the library's scratch
// buffers are therefore shared state: a CallGuard is held for the duration of every entry point that touches them.
// It (a) serialises host threads and (b) when the stream differs from the previous call's stream, makes the new stream
// wait for everything enqueued on the previous one (event record + stream wait), so concurrent streams stay correct -
// they simply do not overlap inside libofdmgan.
struct CallGuard {
    int rc;
    explicit CallGuard(cudaStream_t s);
    ~CallGuard();
    CallGuard(const CallGuard&) = delete;
    CallGuard& operator=(const CallGuard&) = delete;
};
// device scratch owned by the library (per slot), `bytes` each; returns the slot's pointer
int scratch_for_slot(int slot, size_t bytes, int which, void** ptr);
// copy `n` floats that may live on the host or the device into device scratch (no-op if already on device)
int to_device_f32(const float* src, int n, int slot, int which, cudaStream_t s, const float** dev);

// ---- Philox4x32-10 (counter-based RNG; definition in oracle/channel.c header and DESIGN.md) -----------------
struct PhiloxKeys {
    uint32_t k0[10], k1[10];   // the 10 round keys are the same for every thread: precomputed on the host
};
inline PhiloxKeys philox_keys(uint64_t seed) {
    PhiloxKeys k;
    uint32_t a = (uint32_t)seed, b = (uint32_t)(seed >> 32);
    for (int r = 0; r < 10; ++r) {
        k.k0[r] = a;
        k.k1[r] = b;
        a += 0x9E3779B9u;
        b += 0xBB67AE85u;
    }
    return k;
}

#ifdef __CUDACC__
__device__ __forceinline__ void philox4x32_10(const PhiloxKeys& k, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                              uint32_t (&out)[4]) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        uint32_t n0 = hi1 ^ c1 ^ k.k0[r];
        uint32_t n2 = hi0 ^ c3 ^ k.k1[r];
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// u1 = ((x>>9)+0.5)*2^-23 in (0,1): exact in fp32.   u2 = (x>>8)*2^-24 in [0,1): exact.
__device__ __forceinline__ float u_open(uint32_t x) { return __uint_as_float(0x3F800000u | (x >> 9)) - (1.0f - 5.9604644775390625e-08f); }
__device__ __forceinline__ float u_half(uint32_t x) { return (float)(x >> 8) * 5.9604644775390625e-08f; }

// ---- packed fp32x2 (sm_100 FFMA2): two FMAs per issue slot -----------------------------------------------------
// A 64-bit value holds two floats (lo, hi) in an aligned register pair.  ptxas folds a pair built from the same scalar
// twice into the broadcast operand form (FFMA2 Rd, Ra.F32, URb.F32x2, Rc), and a pair loaded from constant memory into a
// uniform-register pair, so no packing instructions are executed.  Same FLOP/s peak as scalar FFMA (the FMA pipe is
// busy two cycles), but half the issue slots - which is what these issue-bound kernels are short of
// (tools/microbench/ffma2.cu: 72 vs 74 TFLOP/s alone, 17.8 vs 25.9 TFLOP/s when mixed 1:2 with ALU work).
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pk2(float lo, float hi) { f32x2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void upk2(f32x2 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) { f32x2 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ f32x2 fma2_rd(f32x2 a, f32x2 b, f32x2 c) { f32x2 r; asm("fma.rm.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) { f32x2 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
// a pair of adjacent floats of a constant image (8-byte aligned index)
__device__ __forceinline__ f32x2 ldc2(const float* p) { return *reinterpret_cast<const f32x2*>(p); }

// ---- single-instruction MUFU forms (no range fix-up code, no slow-path calls) ---------------------------------
__device__ __forceinline__ float fast_sqrt(float x) { float y; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float fast_rcp(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float fast_ex2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float fast_lg2(float x) { float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float fast_sin(float x) { float y; asm("sin.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float fast_cos(float x) { float y; asm("cos.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

// Box-Muller on the word pairs (x0,x1), (x2,x3): 4 normals per Philox block.
//   r = sqrt(-2 ln u1) = sqrt(-2 ln2 * lg2 u1), u1 in (0,1) so the radicand is > 0; angle 2 pi u2 in [0, 2 pi)
//   `var` scales the variance (the caller folds any output scale s as var = s^2: the multiply disappears into the
//   constant under the square root); the angle is (x >> 8) * (2 pi 2^-24), one multiply.
__device__ __forceinline__ void normals_from_block(const uint32_t (&x)[4], float (&n)[4], float var = 1.0f) {
    const float k = -1.3862943611198906f * var;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const float r = fast_sqrt(k * fast_lg2(u_open(x[2 * h])));
        const float th = (float)(x[2 * h + 1] >> 8) * 3.7450702829239286e-07f;
        n[2 * h] = r * fast_cos(th);
        n[2 * h + 1] = r * fast_sin(th);
    }
}

// ---- warp transpose-reduce ---------------------------------------------------------------------------------
// in: every lane holds 32 per-lane partials v[0..31].  out: lane l returns sum over lanes of v[l].
// 31 shuffles + 31 adds instead of 32 x 5: the split-K reduction every weight-gradient kernel ends with.
__device__ __forceinline__ float warp_transpose_reduce(float (&v)[32], int lane) {
#pragma unroll
    for (int s = 16; s >= 1; s >>= 1) {
        const bool upper = (lane & s) != 0;
#pragma unroll
        for (int i = 0; i < s; ++i) {
            float send = upper ? v[i] : v[i + s];
            float keep = upper ? v[i + s] : v[i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, s);
        }
    }
    return v[0];
}

// per-lane register accumulators of transpose-reduced gradient groups: NG registers per lane = 32*NG slots per warp
template <int NG>
struct GradAcc {
    float g[NG];
    __device__ __forceinline__ void zero() {
#pragma unroll
        for (int i = 0; i < NG; ++i) g[i] = 0.f;
    }
};

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int s = 16; s >= 1; s >>= 1) v += __shfl_xor_sync(0xffffffffu, v, s);
    return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int s = 16; s >= 1; s >>= 1) v += __shfl_xor_sync(0xffffffffu, v, s);
    return v;
}

// LeakyReLU for 0 <= slope <= 1 (the API rejects anything else): max(v, slope*v), FMUL + FMNMX
__device__ __forceinline__ float lrelu(float v, float slope) { return fmaxf(v, slope * v); }
inline bool slope_ok(float s) { return s >= 0.f && s <= 1.f; }
#endif  // __CUDACC__

}  // namespace og
