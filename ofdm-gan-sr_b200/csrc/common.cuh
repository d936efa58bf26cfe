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
// device scratch owned by the library PER STREAM (and device): the entry points that carry their weights by value hold no lock, so
// two streams must never share a scratch buffer
int scratch_for_stream(cudaStream_t s, size_t bytes, int which, void** ptr);
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
// one IMAD.WIDE.U32 for both halves of a 32 x 32 -> 64 product (separate mul.hi / mul.lo are not always re-fused by ptxas, and a
// wide multiply holds the heavy FMA pipe for 4 cycles: it is the most expensive instruction of the simulator kernels)
__device__ __forceinline__ void mulwide(uint32_t m, uint32_t x, uint32_t& hi, uint32_t& lo) {
    asm("{\n\t.reg .b64 t;\n\tmul.wide.u32 t, %2, %3;\n\tmov.b64 {%1, %0}, t;\n\t}" : "=r"(hi), "=r"(lo) : "r"(x), "r"(m));
}
// R rounds (10: the default everywhere; 7: the "fast RNG" workload of ofdmgan_chan_cfg.rng_rounds)
template <int R>
__device__ __forceinline__ void philox4x32(const PhiloxKeys& k, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t (&out)[4]);
__device__ __forceinline__ void philox4x32_10(const PhiloxKeys& k, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                              uint32_t (&out)[4]) {
    philox4x32<10>(k, c0, c1, c2, c3, out);
}
template <int R>
__device__ __forceinline__ void philox4x32(const PhiloxKeys& k, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t (&out)[4]) {
#pragma unroll
    for (int r = 0; r < R; ++r) {
        uint32_t hi0, lo0, hi1, lo1;
        mulwide(0xD2511F53u, c0, hi0, lo0);
        mulwide(0xCD9E8D57u, c2, hi1, lo1);
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
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) { f32x2 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
// a pair of adjacent floats of a constant image (8-byte aligned index)
__device__ __forceinline__ f32x2 ldc2(const float* p) { return *reinterpret_cast<const f32x2*>(p); }

// ---- single-instruction MUFU forms (no range fix-up code, no slow-path calls) ---------------------------------
__device__ __forceinline__ float fast_sqrt(float x) { float y; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float fast_rcp(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float fast_ex2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float fast_lg2(float x) { float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float fast_sin(float x) { float y; asm("sin.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float fast_cos(float x) { float y; asm("cos.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

// ---- normals: Box-Muller, THREE pairs per Philox block -----------------------------------------------------------------
// A pair needs a radius uniform u1 (its resolution sets the tail: 23 bits reach 5.8 sigma) and an angle; 16 angle bits (65,536
// directions) are plenty, so one 128-bit block feeds three pairs instead of two - a quarter fewer Philox blocks per frame,
// and the 32 x 32 -> 64 multiplies of Philox are the most expensive instructions of the fused kernel (4 cycles of the FMA-heavy
// pipe each).  Pair `slot` of block (x0, x1, x2, x3):
//     radius bits m = x[slot] & 0x7FFFFF,          u1 = (m + 0.5) * 2^-23 in (0,1)
//     angle bits  a = x3 & 0xFFFF | x3 >> 16 | (x0 >> 24) | (x1 >> 24) << 8   for slot 0 | 1 | 2,   theta = 2 pi a / 65536
//     normals (r cos theta, r sin theta), r = sqrt(-2 ln u1 * var) = sqrt(k * lg2 u1), k = -2 ln2 * var
// (any output scale s enters as var = s^2 under the square root: free).  No integer-to-float conversion anywhere: the radius bits
// are OR-ed into the mantissa of [1,2) (one LOP3) and shifted down by an exact subtraction; the angle bits are byte-permuted
// into the mantissa of 2^23 (value 2^23 + a exactly) and ONE fma forms a * (2 pi / 65536) with a single rounding.
// A "section" of the stream is NP consecutive pairs starting at block blk0: pair p lives in block blk0 + p / 3, slot p % 3.
// Definition shared with oracle/channel.c (section_normals).
#define OG_BM_K (-1.3862943611198906f)                     /* -2 ln 2 */
// (x & m) | c as ONE LOP3 (the compiler splits it in two when both constants are immediates)
__device__ __forceinline__ uint32_t and_or(uint32_t x, uint32_t m, uint32_t c) {
    uint32_t r;
    asm("lop3.b32 %0, %1, %2, %3, 0xEA;" : "=r"(r) : "r"(x), "r"(m), "r"(c));
    return r;
}
// polar pieces of pair `slot` (compile-time after unrolling) so that callers can fold the product into an FMA
__device__ __forceinline__ void bm_polar(const uint32_t (&x)[4], int slot, float k, float& r, float& c, float& s) {
    const float u1 = __uint_as_float(and_or(x[slot], 0x007FFFFFu, 0x3F800000u)) - (1.0f - 5.9604644775390625e-08f);
    r = fast_sqrt(k * fast_lg2(u1));
    const uint32_t tb = slot == 0   ? __byte_perm(x[3], 0x4B000000u, 0x7610)
                        : slot == 1 ? __byte_perm(x[3], 0x4B000000u, 0x7632)
                                    : __byte_perm(__byte_perm(x[0], x[1], 0x3373), 0x4B000000u, 0x7610);
    const float th = fmaf(__uint_as_float(tb), 9.587379924285257e-05f, -804.247719318987f);   // (2^23 + a) 2 pi 2^-16 - 2^23 2 pi 2^-16
    c = fast_cos(th);
    s = fast_sin(th);
}
// the first NP (<= 3) pairs of one block as 2 NP normals of variance `var`
template <int NP>
__device__ __forceinline__ void normals_from_block(const uint32_t (&x)[4], float (&n)[2 * NP], float var = 1.0f) {
    const float k = OG_BM_K * var;
#pragma unroll
    for (int q = 0; q < NP; ++q) {
        float r, c, s;
        bm_polar(x, q, k, r, c, s);
        n[2 * q] = r * c;
        n[2 * q + 1] = r * s;
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
