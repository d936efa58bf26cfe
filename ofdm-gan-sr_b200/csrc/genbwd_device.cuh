// Per-thread MiniGenerator backward (what autograd does for train.py:295 g_loss.backward() through
// models/generator.py:180-208), on the folded weight image of weights.cuh.
//
// Generator slot map (10 groups = 320 slots).  The two upsample+conv layers accumulate gradients of their FOLDED taps
// {F0=w0, F1=w1+w2, F2=w0+w1, F3=w2}; the finalize kernel maps them back linearly: dw0 = dF0+dF2, dw1 = dF1+dF2,
// dw2 = dF1+dF3.
//   G0      out_conv folded dF[2][4][4]     slots   0..31
//   G1-G4   dec1 folded dF[4][8][4]         slots  32..159
//   G5-G7   bottleneck.weight[8][4][3]      slots 160..255
//   G8      enc1.weight[4][2][3] 256..279, enc1.bias[4] 280..283
//   G9      bottleneck.bias[8] 288..295, dec1.bias[4] 296..299, out_conv.bias[2] 300..301 (stats 302..304 added by kernels)
#pragma once
#include "critic_device.cuh"
#include "gen_device.cuh"

namespace og {

constexpr int G_NG = 10;
constexpr int GS_OUTF = 0, GS_DECF = 32, GS_BNW = 160, GS_ENCW = 256, GS_ENCB = 280, GS_BNB = 288, GS_DECB = 296, GS_OUTB = 300,
              GS_S0 = 302, GS_SLOTS = 320;

// x: generator input; (a1, a2, sk, z3pos, y): tape from gen_fwd_f32<true>; dy: upstream gradient.
// Accumulates parameter-gradient groups; if NEED_DX writes dx.
template <bool NEED_DX>
__device__ __forceinline__ void gen_bwd(const float* __restrict__ W, float slope, const float (&x)[2][16], const float (&a1)[4][8],
                                        const float (&a2)[8][4], const float (&sk)[4][8], uint32_t z3pos, const float (&y)[2][16],
                                        const float (&dy)[2][16], GradAcc<G_NG>& acc, int lane, float (&dx)[2][16]) {
    float bias_g[14];   // bn.b[8], dec.b[4], out.b[2] - flushed together as G9
    // ---- tanh + out_conv (folded over upsample2)
    float dz4[2][16];
#pragma unroll
    for (int oc = 0; oc < 2; ++oc) {
        float s = 0.f;
#pragma unroll
        for (int q = 0; q < 16; ++q) { dz4[oc][q] = dy[oc][q] * (1.0f - y[oc][q] * y[oc][q]); s += dz4[oc][q]; }
        bias_g[12 + oc] = s;
    }
    {
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            const int oc = j / 16, ic = (j / 4) % 4, t = j % 4;
            float a = 0.f;
#pragma unroll
            for (int p = 0; p < 8; ++p) {
                if (t == 0 && p > 0) a = fmaf(dz4[oc][2 * p], sk[ic][p - 1], a);
                if (t == 1) a = fmaf(dz4[oc][2 * p], sk[ic][p], a);
                if (t == 2) a = fmaf(dz4[oc][2 * p + 1], sk[ic][p], a);
                if (t == 3 && p < 7) a = fmaf(dz4[oc][2 * p + 1], sk[ic][p + 1], a);
            }
            v[j] = a;
        }
        acc.g[0] += warp_transpose_reduce(v, lane);
    }
    float dsk[4][8];
#pragma unroll
    for (int ic = 0; ic < 4; ++ic)
#pragma unroll
        for (int p = 0; p < 8; ++p) {
            float a = 0.f;
#pragma unroll
            for (int oc = 0; oc < 2; ++oc) {
                const float* F = W + GI_OUT_F + (oc * 4 + ic) * 4;
                a = fmaf(F[1], dz4[oc][2 * p], a);
                a = fmaf(F[2], dz4[oc][2 * p + 1], a);
                if (p < 7) a = fmaf(F[0], dz4[oc][2 * p + 2], a);
                if (p > 0) a = fmaf(F[3], dz4[oc][2 * p - 1], a);
            }
            dsk[ic][p] = a;
        }
    // ---- dec1 (folded over upsample1) ; dz3 = dsk . lrelu'(z3)
    float dz3[4][8];
#pragma unroll
    for (int oc = 0; oc < 4; ++oc) {
        float s = 0.f;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            dz3[oc][q] = dsk[oc][q] * (((z3pos >> (oc * 8 + q)) & 1u) ? 1.0f : slope);
            s += dz3[oc][q];
        }
        bias_g[8 + oc] = s;
    }
#pragma unroll
    for (int grp = 0; grp < 4; ++grp) {
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            const int s = grp * 32 + j, oc = s / 32, ic = (s / 4) % 8, t = s % 4;
            float a = 0.f;
#pragma unroll
            for (int p = 0; p < 4; ++p) {
                if (t == 0 && p > 0) a = fmaf(dz3[oc][2 * p], a2[ic][p - 1], a);
                if (t == 1) a = fmaf(dz3[oc][2 * p], a2[ic][p], a);
                if (t == 2) a = fmaf(dz3[oc][2 * p + 1], a2[ic][p], a);
                if (t == 3 && p < 3) a = fmaf(dz3[oc][2 * p + 1], a2[ic][p + 1], a);
            }
            v[j] = a;
        }
        acc.g[1 + grp] += warp_transpose_reduce(v, lane);
    }
    float dz2[8][4];
#pragma unroll
    for (int ic = 0; ic < 8; ++ic) {
        float s = 0.f;
#pragma unroll
        for (int p = 0; p < 4; ++p) {
            float a = 0.f;
#pragma unroll
            for (int oc = 0; oc < 4; ++oc) {
                const float* F = W + GI_DEC_F + (oc * 8 + ic) * 4;
                a = fmaf(F[1], dz3[oc][2 * p], a);
                a = fmaf(F[2], dz3[oc][2 * p + 1], a);
                if (p < 3) a = fmaf(F[0], dz3[oc][2 * p + 2], a);
                if (p > 0) a = fmaf(F[3], dz3[oc][2 * p - 1], a);
            }
            dz2[ic][p] = a * (a2[ic][p] > 0.f ? 1.0f : slope);
            s += dz2[ic][p];
        }
        bias_g[ic] = s;
    }
    // ---- bottleneck
#pragma unroll
    for (int grp = 0; grp < 3; ++grp) {
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            const int s = grp * 32 + j, oc = s / 12, ic = (s / 3) % 4, k = s % 3;
            float a = 0.f;
#pragma unroll
            for (int p = 0; p < 4; ++p) {
                const int i = 2 * p + k - 1;
                if (i >= 0) a = fmaf(dz2[oc][p], a1[ic][i], a);
            }
            v[j] = a;
        }
        acc.g[5 + grp] += warp_transpose_reduce(v, lane);
    }
    float dz1[4][8];
#pragma unroll
    for (int ic = 0; ic < 4; ++ic)
#pragma unroll
        for (int i = 0; i < 8; ++i) dz1[ic][i] = dsk[ic][i];                     // skip branch
#pragma unroll
    for (int oc = 0; oc < 8; ++oc)
#pragma unroll
        for (int p = 0; p < 4; ++p)
#pragma unroll
            for (int ic = 0; ic < 4; ++ic)
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    const int i = 2 * p + k - 1;
                    if (i >= 0) dz1[ic][i] = fmaf(W[GI_BN_W + (oc * 4 + ic) * 3 + k], dz2[oc][p], dz1[ic][i]);
                }
#pragma unroll
    for (int ic = 0; ic < 4; ++ic)
#pragma unroll
        for (int i = 0; i < 8; ++i) dz1[ic][i] *= a1[ic][i] > 0.f ? 1.0f : slope;
    // ---- enc1
    {
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            float a = 0.f;
            if (j < 24) {
                const int oc = j / 6, ic = (j / 3) % 2, k = j % 3;
#pragma unroll
                for (int p = 0; p < 8; ++p) {
                    const int i = 2 * p + k - 1;
                    if (i >= 0) a = fmaf(dz1[oc][p], x[ic][i], a);
                }
            } else if (j < 28) {
#pragma unroll
                for (int p = 0; p < 8; ++p) a += dz1[j - 24][p];
            }
            v[j] = a;
        }
        acc.g[8] += warp_transpose_reduce(v, lane);
    }
    {
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = j < 14 ? bias_g[j] : 0.f;
        acc.g[9] += warp_transpose_reduce(v, lane);
    }
    if (NEED_DX) {
#pragma unroll
        for (int ic = 0; ic < 2; ++ic)
#pragma unroll
            for (int i = 0; i < 16; ++i) dx[ic][i] = 0.f;
#pragma unroll
        for (int oc = 0; oc < 4; ++oc)
#pragma unroll
            for (int p = 0; p < 8; ++p)
#pragma unroll
                for (int ic = 0; ic < 2; ++ic)
#pragma unroll
                    for (int k = 0; k < 3; ++k) {
                        const int i = 2 * p + k - 1;
                        if (i >= 0) dx[ic][i] = fmaf(W[GI_ENC_W + (oc * 2 + ic) * 3 + k], dz1[oc][p], dx[ic][i]);
                    }
    }
}

// slot of the accumulator -> flat torch parameter index handling (finalize kernels)
// raw dec1 / out_conv taps from folded gradients: (w0,w1,w2) <- (F0+F2, F1+F2, F1+F3)
__device__ __forceinline__ float unfold_tap(const float* F4, int k) {
    return k == 0 ? F4[0] + F4[2] : (k == 1 ? F4[1] + F4[2] : F4[1] + F4[3]);
}

}  // namespace og
