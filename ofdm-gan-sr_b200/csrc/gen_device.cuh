// Per-thread (one OFDM frame per thread, everything in registers) generator forward passes.
//   gen_fwd_f32   : MiniGenerator.forward, models/generator.py:180-208, from the folded G image (weights.cuh)
//   gen_fwd_q     : Q1.7/Q8.8 integer generator, rtl/ofdmGAN/generator_mini.v:326-649, modes spec / rtl_literal,
//                   evaluated bit-exactly on the FP32 pipe (see fxacc below)
// Weights come from a __constant__ image: ptxas turns every W[i] into an LDCU'd uniform register, so one tap is
// one FFMA R,R,UR,R.  All loops are compile-time unrolled; there is no indexing at run time.
#pragma once
#include "weights.cuh"

namespace og {

// tanh(x) = 1 - 2/(1+e^{2x}) = 1 - 2/(1 + 2^(x * 2 log2 e)): FMUL, MUFU.EX2, FADD, MUFU.RCP, FFMA.  Saturates correctly
// (e -> inf gives 1, e -> 0 gives -1); absolute error <= ~2e-7 over the whole range.
__device__ __forceinline__ float tanh_fast(float x) {
    const float e = fast_ex2(x * 2.8853900817779268f);
    return fmaf(-2.0f, fast_rcp(e + 1.0f), 1.0f);
}

__device__ __forceinline__ float lrelu_sel(float v, float slope) { return lrelu(v, slope); }

// ---------------------------------------------------------------------------------------------------- fp32
// x[2][16] -> y[2][16], models/generator.py:180-208.  Two output channels per packed FFMA2: the input sample is the broadcast
// scalar operand, the weight pair (w[oc], w[oc+1]) comes from the pair-interleaved half of the G image as one uniform load.
__device__ __forceinline__ void gen_fwd_f32_infer(const float* __restrict__ W, float slope, const float (&x)[2][16],
                                                  float (&y)[2][16]) {
    float a1[4][8], a2[8][4], sk[4][8];
    // enc1: Conv1d(2->4, k3, s2, p1) + LeakyReLU
#pragma unroll
    for (int o2 = 0; o2 < 2; ++o2)
#pragma unroll
        for (int p = 0; p < 8; ++p) {
            f32x2 acc = ldc2(W + GI_ENC_B + 2 * o2);
#pragma unroll
            for (int ic = 0; ic < 2; ++ic)
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    const int i = 2 * p + k - 1;
                    if (i >= 0) acc = fma2(pk2(x[ic][i], x[ic][i]), ldc2(W + GI2_ENC + ((o2 * 2 + ic) * 3 + k) * 2), acc);
                }
            float lo, hi;
            upk2(acc, lo, hi);
            a1[2 * o2][p] = lrelu(lo, slope);
            a1[2 * o2 + 1][p] = lrelu(hi, slope);
        }
    // bottleneck: Conv1d(4->8, k3, s2, p1) + LeakyReLU
#pragma unroll
    for (int o2 = 0; o2 < 4; ++o2)
#pragma unroll
        for (int p = 0; p < 4; ++p) {
            f32x2 acc = ldc2(W + GI_BN_B + 2 * o2);
#pragma unroll
            for (int ic = 0; ic < 4; ++ic)
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    const int i = 2 * p + k - 1;
                    if (i >= 0) acc = fma2(pk2(a1[ic][i], a1[ic][i]), ldc2(W + GI2_BN + ((o2 * 4 + ic) * 3 + k) * 2), acc);
                }
            float lo, hi;
            upk2(acc, lo, hi);
            a2[2 * o2][p] = lrelu(lo, slope);
            a2[2 * o2 + 1][p] = lrelu(hi, slope);
        }
    // upsample x2 + dec1 Conv1d(8->4, k3, s1, p1) + LeakyReLU, folded; + additive skip
#pragma unroll
    for (int o2 = 0; o2 < 2; ++o2)
#pragma unroll
        for (int p = 0; p < 4; ++p) {
            f32x2 e = ldc2(W + GI_DEC_B + 2 * o2), o = e;
#pragma unroll
            for (int ic = 0; ic < 8; ++ic) {
                const float* F = W + GI2_DEC + (o2 * 8 + ic) * 8;
                if (p > 0) e = fma2(pk2(a2[ic][p - 1], a2[ic][p - 1]), ldc2(F + 0), e);
                e = fma2(pk2(a2[ic][p], a2[ic][p]), ldc2(F + 2), e);
                o = fma2(pk2(a2[ic][p], a2[ic][p]), ldc2(F + 4), o);
                if (p < 3) o = fma2(pk2(a2[ic][p + 1], a2[ic][p + 1]), ldc2(F + 6), o);
            }
            float e0, e1, o0, o1;
            upk2(e, e0, e1);
            upk2(o, o0, o1);
            sk[2 * o2][2 * p] = lrelu(e0, slope) + a1[2 * o2][2 * p];
            sk[2 * o2][2 * p + 1] = lrelu(o0, slope) + a1[2 * o2][2 * p + 1];
            sk[2 * o2 + 1][2 * p] = lrelu(e1, slope) + a1[2 * o2 + 1][2 * p];
            sk[2 * o2 + 1][2 * p + 1] = lrelu(o1, slope) + a1[2 * o2 + 1][2 * p + 1];
        }
    // upsample x2 + out_conv Conv1d(4->2, k3, s1, p1), folded; tanh
#pragma unroll
    for (int p = 0; p < 8; ++p) {
        f32x2 e = ldc2(W + GI_OUT_B), o = e;
#pragma unroll
        for (int ic = 0; ic < 4; ++ic) {
            const float* F = W + GI2_OUT + ic * 8;
            if (p > 0) e = fma2(pk2(sk[ic][p - 1], sk[ic][p - 1]), ldc2(F + 0), e);
            e = fma2(pk2(sk[ic][p], sk[ic][p]), ldc2(F + 2), e);
            o = fma2(pk2(sk[ic][p], sk[ic][p]), ldc2(F + 4), o);
            if (p < 7) o = fma2(pk2(sk[ic][p + 1], sk[ic][p + 1]), ldc2(F + 6), o);
        }
        float e0, e1, o0, o1;
        upk2(e, e0, e1);
        upk2(o, o0, o1);
        y[0][2 * p] = tanh_fast(e0);
        y[0][2 * p + 1] = tanh_fast(o0);
        y[1][2 * p] = tanh_fast(e1);
        y[1][2 * p + 1] = tanh_fast(o1);
    }
}

// ---------------------------------------------------------------------------------------------------- fixed point
// Integer semantics on the FP32 pipe.  Activations are int16 values held exactly in floats; weights are k/128
// (exact).  An accumulator carries MAGIC + n with MAGIC = 1.5*2^23, where one ulp is exactly 1, so
//     acc = fma_rd(a, w, acc)  ==  acc + floor(a*k/128)          (round toward -inf, a*w formed exactly inside the FMA)
// which is the RTL's per-tap `(a*w) >>> 7` followed by the 32-bit add (generator_mini.v:141-146): one FFMA.RM per tap.
// Valid while |n| < 2^22; a layer's worst case is 24 taps * 2^15 < 2^20.
constexpr float FX_MAGIC = 12582912.0f;
__device__ __forceinline__ float fxacc(float a, float w, float acc) { return __fmaf_rd(a, w, acc); }
__device__ __forceinline__ float fx_sat16(float v) { return fminf(fmaxf(v, -32768.0f), 32767.0f); }
// LeakyReLU of the RTL: r < 0 -> (r>>>2) + (r>>>4)   (generator_mini.v:359-360)
__device__ __forceinline__ float fx_lrelu(float r) {
    float t = __fmaf_rd(r, 0.25f, FX_MAGIC);
    t = __fmaf_rd(r, 0.0625f, t);
    return r < 0.f ? t - FX_MAGIC : r;
}
__device__ __forceinline__ float fx_finish(float acc, float bias, bool act) {
    float v = fx_sat16((acc - FX_MAGIC) + bias);
    return act ? fx_lrelu(v) : v;
}
__device__ __forceinline__ float fx_clip(float v) { return v > 256.f ? 255.f : (v < -256.f ? -255.f : v); }

// mode spec: RTL primitives on the textbook dataflow (all channels, aligned weights, 1x1 output conv, clip both).
// Two output channels per FFMA2.RM: acc2 = fma.rm.f32x2((a, a), (w[oc], w[oc+1]), acc2) with the weight pair read from the
// interleaved half of the Q image.
__device__ __forceinline__ f32x2 fxacc2(float a, f32x2 w2, f32x2 acc2) { return fma2_rd(pk2(a, a), w2, acc2); }
// Finish a channel pair whose accumulator started at MAGIC + bias (exact): leave the magic domain, saturate, LeakyReLU - the
// subtracts and the two floor-FMAs of the activation packed (6 instead of 9 instructions per value).
__device__ __forceinline__ void fx_finish2(f32x2 acc, bool act, float& lo, float& hi) {
    const f32x2 nmagic2 = pk2(-FX_MAGIC, -FX_MAGIC);
    upk2(add2(acc, nmagic2), lo, hi);
    lo = fx_sat16(lo);
    hi = fx_sat16(hi);
    if (act) {
        const f32x2 r2 = pk2(lo, hi);
        f32x2 t2 = fma2_rd(r2, pk2(0.25f, 0.25f), pk2(FX_MAGIC, FX_MAGIC));
        t2 = fma2_rd(r2, pk2(0.0625f, 0.0625f), t2);
        float tl, th;
        upk2(add2(t2, nmagic2), tl, th);
        lo = lo < 0.f ? tl : lo;
        hi = hi < 0.f ? th : hi;
    }
}

__device__ __forceinline__ void gen_fwd_q_spec(const float* __restrict__ Q, const float (&x)[2][16], float (&y)[2][16]) {
    float a1[4][8], a2[8][4], sk[4][8];
    const f32x2 magic2 = pk2(FX_MAGIC, FX_MAGIC);
#pragma unroll
    for (int o2 = 0; o2 < 2; ++o2)
#pragma unroll
        for (int p = 0; p < 8; ++p) {
            f32x2 acc = ldc2(Q + QI_BIASM + 0 + 2 * o2);
#pragma unroll
            for (int ic = 0; ic < 2; ++ic)
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    const int i = 2 * p + k - 1;
                    if (i >= 0) acc = fxacc2(x[ic][i], ldc2(Q + QI2_ENC + ((o2 * 2 + ic) * 3 + k) * 2), acc);
                }
            fx_finish2(acc, true, a1[2 * o2][p], a1[2 * o2 + 1][p]);
        }
#pragma unroll
    for (int o2 = 0; o2 < 4; ++o2)
#pragma unroll
        for (int p = 0; p < 4; ++p) {
            f32x2 acc = ldc2(Q + QI_BIASM + 4 + 2 * o2);
#pragma unroll
            for (int ic = 0; ic < 4; ++ic)
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    const int i = 2 * p + k - 1;
                    if (i >= 0) acc = fxacc2(a1[ic][i], ldc2(Q + QI2_BN + ((o2 * 4 + ic) * 3 + k) * 2), acc);
                }
            fx_finish2(acc, true, a2[2 * o2][p], a2[2 * o2 + 1][p]);
        }
    // per-tap floor does not commute with weight folding: evaluate the 3 taps on the upsampled signal
#pragma unroll
    for (int o2 = 0; o2 < 2; ++o2)
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            f32x2 acc = ldc2(Q + QI_BIASM + 12 + 2 * o2);
#pragma unroll
            for (int ic = 0; ic < 8; ++ic)
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    const int i = q + k - 1;
                    if (i >= 0 && i < 8) acc = fxacc2(a2[ic][i >> 1], ldc2(Q + QI2_DEC + ((o2 * 8 + ic) * 3 + k) * 2), acc);
                }
            float lo, hi;
            fx_finish2(acc, true, lo, hi);
            sk[2 * o2][q] = fx_sat16(lo + a1[2 * o2][q]);
            sk[2 * o2 + 1][q] = fx_sat16(hi + a1[2 * o2 + 1][q]);
        }
#pragma unroll
    for (int p = 0; p < 8; ++p) {
        f32x2 acc = ldc2(Q + QI_BIASM + 16);
#pragma unroll
        for (int ic = 0; ic < 4; ++ic) acc = fxacc2(sk[ic][p], ldc2(Q + QI2_OUT + ic * 2), acc);
        float lo, hi;
        fx_finish2(acc, false, lo, hi);
        const float v0 = fx_clip(lo), v1 = fx_clip(hi);
        y[0][2 * p] = v0; y[0][2 * p + 1] = v0;
        y[1][2 * p] = v1; y[1][2 * p + 1] = v1;
    }
}

// mode rtl_literal: what the committed RTL computes (oracle/fixed_point.c skewed_conv).
// The weight triple used by loop iteration (oc,op,ic) is the one addressed by the previous iteration.
template <int IN, int OC, int K, int WA, int OC_FIRST, int STALE>
__device__ __forceinline__ constexpr int skew_addr(int oc, int op, int ic) {
    return ic > 0 ? WA + oc * IN * K + (ic - 1) * K
                  : (op > 0 || oc == OC - 1) ? WA + oc * IN * K + (IN - 1) * K
                                             : (oc > OC_FIRST ? WA + (oc - 1) * IN * K + (IN - 1) * K : STALE);
}

__device__ __forceinline__ void gen_fwd_q_rtl(const float* __restrict__ Q, const float (&x)[2][16], float (&y)[2][16]) {
    float a1[4][8], bn7[4], dec[4][8];
    // enc1: all 4 channels, stale address 223 (steady state)
#pragma unroll
    for (int oc = 0; oc < 4; ++oc)
#pragma unroll
        for (int p = 0; p < 8; ++p) {
            float acc = FX_MAGIC;
#pragma unroll
            for (int ic = 0; ic < 2; ++ic) {
                const int a = skew_addr<2, 4, 3, 0, 0, 223>(oc, p, ic);
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    const int i = 2 * p + k - 1;
                    if (i >= 0) acc = fxacc(x[ic][i], Q[a + k], acc);
                }
            }
            a1[oc][p] = fx_finish(acc, Q[QI_BIAS + 0 + oc], true);
        }
    // bottleneck: only out-channels 3..7 are computed and only channel 7 is ever consumed (UPSAMPLE1 copies ch 7 only)
#pragma unroll
    for (int p = 0; p < 4; ++p) {
        float acc = FX_MAGIC;
#pragma unroll
        for (int ic = 0; ic < 4; ++ic) {
            const int a = skew_addr<4, 8, 3, 24, 3, 21>(7, p, ic);
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const int i = 2 * p + k - 1;
                if (i >= 0) acc = fxacc(a1[ic][i], Q[a + k], acc);
            }
        }
        bn7[p] = fx_finish(acc, Q[QI_BIAS + 4 + 7], true);
    }
    // dec1 over up1 where only channel 7 is non-zero: a zero activation contributes floor(0)=0 per tap
#pragma unroll
    for (int oc = 0; oc < 4; ++oc)
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            float acc = FX_MAGIC;
            const int a = skew_addr<8, 4, 3, 120, 0, 117>(oc, q, 7);
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const int i = q + k - 1;
                if (i >= 0 && i < 8) acc = fxacc(bn7[i >> 1], Q[a + k], acc);
            }
            dec[oc][q] = fx_finish(acc, Q[QI_BIAS + 12 + oc], true);
        }
#pragma unroll
    for (int q = 0; q < 8; ++q) dec[3][q] = fx_sat16(dec[3][q] + a1[3][q]);           // skip add only on channel 3
    // 1x1 output conv on the upsampled signal, skewed; clip only channel 1
#pragma unroll
    for (int oc = 0; oc < 2; ++oc)
#pragma unroll
        for (int q = 0; q < 16; ++q) {
            float acc = FX_MAGIC;
#pragma unroll
            for (int ic = 0; ic < 4; ++ic) acc = fxacc(dec[ic][q >> 1], Q[skew_addr<4, 2, 1, 216, 0, 213>(oc, q, ic)], acc);
            float v = fx_finish(acc, Q[QI_BIAS + 16 + oc], false);
            y[oc][q] = oc == 1 ? fx_clip(v) : v;
        }
}

}  // namespace og
