// Register-lean critic passes for the fused training kernels (<= 128 registers per thread, 16 warps per SM).
//
// One sample per thread.  What keeps the live set, the code and the issue-slot count small:
//   * the sample's frames stay in the warp's shared-memory tile (swizzled, io_tile.cuh) and are re-read ONE 16-float
//     row at a time wherever a pass needs them;
//   * the per-lane parameter-gradient accumulators live in shared memory ([group][thread], conflict free): a group's
//     transpose-reduced sum is added with one LDS + FADD + STS;
//   * every layer is a ROLLED loop over the index that only addresses weights and shared memory (conv2: output-channel pair,
//     conv1: input channel).  All register arrays keep compile-time indices, but a loop iteration is a real basic block:
//     ptxas can no longer hoist the constant loads of a whole layer to the top of a 10k-instruction block and spill them
//     (the fully unrolled version carried 3 KB of stack per thread, essentially the weight image);
//   * activations are held as CHANNEL PAIRS in 64-bit registers, (x[2c][pos], x[2c+1][pos]), and every multiply-accumulate
//     is a packed FFMA2 (common.cuh): forward convolutions pair two output channels (scalar input x weight pair),
//     transposed convolutions pair two input channels, weight gradients pair two channels of the upstream gradient.
//     Weight pairs come from pair-interleaved copies in the constant image (weights.cuh DI2_*).
//
// Accumulator slot map (20 groups of 32 = 640 slots):
//   group ic (0..3)       conv1.weight[oc][ic][k] at j = oc*3 + k (24 slots);
//                         spare j = 24..31: group 0 -> conv1.bias[8]; group 1 -> dense.bias (24), sum D(real) (25),
//                         sum D(fake) (26), sum penalty (27)
//   group 4 + 2*o2 + q    (o2 = output-channel pair 0..7, q = input-channel half 0..1)
//                         conv2.weight[2*o2+h][4*q+c][k] at j = h*16 + c*3 + k (24 slots); in the q = 0 group additionally
//                         conv2.bias[2*o2+h] at j = h*16 + 12 and dense.weight[2*o2+h] at j = h*16 + 13.
//                         Channel h of the pair sits in half h of the group, so the first (16-lane) stage of the transpose-reduce
//                         is an exchange of the two halves of an FFMA2 result: lanes 16-31 build their pairs swapped and the
//                         stage needs no selects (warp_transpose_reduce_pairs below).
// Maths: models/discriminator.py:112-152 (forward), :172-236 (penalty), closed-form double backward as in
// oracle/fp32_models.c gp_sample.
#pragma once
#include "io_tile.cuh"
#include "weights.cuh"

namespace og {

constexpr int CS_NG = 20;
constexpr int CS_SLOTS = CS_NG * 32;
constexpr int CS_C1B = 24;                 // group 0
constexpr int CS_FCB = 32 + 24, CS_SREAL = 32 + 25, CS_SFAKE = 32 + 26, CS_SGP = 32 + 27;

// accumulator slot (group, j) -> parameter index (torch order), or -1 for a spare / statistics slot
__host__ __device__ constexpr int cs_param_of(int grp, int j) {
    if (grp < 4) return j < 24 ? ((j / 3) * 4 + grp) * 3 + j % 3 : (grp == 0 ? DP_C1_B + (j - 24) : (grp == 1 && j == 24 ? DP_FC_B : -1));
    const int o2 = (grp - 4) >> 1, q = (grp - 4) & 1;
    const int h = j >> 4, r = j & 15;
    if (r < 12) return DP_C2_W + ((2 * o2 + h) * 8 + 4 * q + r / 3) * 3 + r % 3;
    if (q == 0 && r == 12) return DP_C2_B + 2 * o2 + h;
    if (q == 0 && r == 13) return DP_FC_W + 2 * o2 + h;
    return -1;
}

// (lo, hi) -> (hi, lo) in the lanes that `up` selects
__device__ __forceinline__ f32x2 cs_swap_if(bool up, f32x2 v) {
    float lo, hi;
    upk2(v, lo, hi);
    return pk2(up ? hi : lo, up ? lo : hi);
}

// Transpose-reduce of a 32-slot group given as N <= 16 PAIRS: pv[i] = (slot i, slot 16 + i) in lanes 0-15 and (slot 16 + i, slot i)
// in lanes 16-31, slots N..15 and 16+N..31 empty.  Every lane keeps the first half and sends the second: the 16-lane stage costs
// SHFL + FADD per pair and no selects; the remaining four stages are the generic ones on 16 values.  Lane l returns slot l.
template <int N, int M>
__device__ __forceinline__ float warp_transpose_reduce_pairs(const f32x2 (&pv)[M], int lane) {
    static_assert(N <= M && N <= 16, "pairs in use");
    float v[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        if (i < N) {
            float keep, send;
            upk2(pv[i], keep, send);
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
        } else {
            v[i] = 0.f;
        }
    }
#pragma unroll
    for (int s = 8; s >= 1; s >>= 1) {
        const bool upper = (lane & s) != 0;
#pragma unroll
        for (int i = 0; i < s; ++i) {
            const float send = upper ? v[i] : v[i + s];
            const float keep = upper ? v[i + s] : v[i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, s);
        }
    }
    return v[0];
}

// shared-memory gradient accumulator of one CTA: [CS_NG][threads]
struct SAcc {
    float* base;       // + threadIdx.x already applied
    int stride;        // threads per CTA
    __device__ __forceinline__ void add(int grp, float r) { base[grp * stride] += r; }
};

// one 16-float row (0 = I, 1 = Q) of the thread's own frame in a resident tile
__device__ __forceinline__ void row_read(const float4* wsm, int lane, int row, float (&x)[16]) {
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        const float4 v = wsm[lane * 8 + ((row * 4 + c) ^ (lane & 7))];
        x[4 * c] = v.x; x[4 * c + 1] = v.y; x[4 * c + 2] = v.z; x[4 * c + 3] = v.w;
    }
}
__device__ __forceinline__ void row_write(float4* wsm, int lane, int row, const float (&x)[16]) {
#pragma unroll
    for (int c = 0; c < 4; ++c)
        wsm[lane * 8 + ((row * 4 + c) ^ (lane & 7))] = make_float4(x[4 * c], x[4 * c + 1], x[4 * c + 2], x[4 * c + 3]);
}

__device__ __forceinline__ float2 f2(f32x2 v) { float2 r; upk2(v, r.x, r.y); return r; }
// channel c of a channel-pair array
#define CS_CH(arr, c, pos) ((c) & 1 ? f2(arr[(c) >> 1][pos]).y : f2(arr[(c) >> 1][pos]).x)

// LeakyReLU derivative pairs (1 or slope) of output-channel pair o2 of conv2: channel 2*o2+h owns bits [4*(2*o2+h), +4) of m2
__device__ __forceinline__ void cs_masks4x2(uint64_t m2, int o2, float slope, f32x2 (&mk)[4]) {
    const uint32_t byte = (uint32_t)(m2 >> (o2 * 8)) & 255u;
#pragma unroll
    for (int p = 0; p < 4; ++p) mk[p] = pk2(((byte >> p) & 1u) ? 1.0f : slope, ((byte >> (4 + p)) & 1u) ? 1.0f : slope);
}

// a1 = LeakyReLU(conv1([cand; cond]) + b1) as channel pairs; input rows streamed from the tiles.
// Returns the sign mask m1 (bit oc*8+p).  BIAS = false: no bias, no activation (the penalty's conv1x(h)).
template <bool BIAS>
__device__ __forceinline__ uint64_t cs_conv1_fwd(const float* W, float slope, const float4* t_cand, const float4* t_cond, int n_rows,
                                                 int lane, f32x2 (&a1)[4][8]) {
#pragma unroll
    for (int o2 = 0; o2 < 4; ++o2)
#pragma unroll
        for (int p = 0; p < 8; ++p) a1[o2][p] = BIAS ? ldc2(W + DP_C1_B + 2 * o2) : pk2(0.f, 0.f);
#pragma unroll 1
    for (int ic = 0; ic < n_rows; ++ic) {
        float row[16];
        row_read(ic < 2 ? t_cand : t_cond, lane, ic & 1, row);
        const float* w = W + DI2_C1 + ic * 6;                   // pair (o2, ic, k) at o2*24 + ic*6 + k*2
#pragma unroll
        for (int o2 = 0; o2 < 4; ++o2)
#pragma unroll
            for (int p = 0; p < 8; ++p)
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    const int i = 2 * p + k - 1;
                    if (i >= 0) a1[o2][p] = fma2(pk2(row[i], row[i]), ldc2(w + o2 * 24 + k * 2), a1[o2][p]);
                }
    }
    uint64_t m1 = 0;
    if (BIAS) {
#pragma unroll
        for (int o2 = 0; o2 < 4; ++o2) {
            uint32_t blo = 0, bhi = 0;
#pragma unroll
            for (int p = 0; p < 8; ++p) {
                float lo, hi;
                upk2(a1[o2][p], lo, hi);
                if (lo > 0.f) blo |= 1u << p; else lo *= slope;
                if (hi > 0.f) bhi |= 1u << p; else hi *= slope;
                a1[o2][p] = pk2(lo, hi);
            }
            m1 |= (uint64_t)(blo | (bhi << 8)) << (o2 * 16);
        }
    }
    return m1;
}

// conv2 + LeakyReLU + sum pool + dense, one output-channel PAIR per loop iteration.
//   MODE 0: forward only (m2, score).
//   MODE 1: + gradient groups of that pair for upstream g = dL/dscore (conv2.weight, conv2.bias, dense.weight).
//   MODE 2: the penalty's second-order terms with `in` = u1: dW2 += dz2 (x) u1, d wd = sum_p m2 * conv2(u1) (no bias gradient);
//           m2 is an INPUT here (the masks of the first-order pass).
template <int MODE>
__device__ __forceinline__ uint64_t cs_conv2_pass(const float* W, float slope, float g, const f32x2 (&in)[4][8], uint64_t m2_in, SAcc& acc,
                                                  int lane, float& score) {
    uint64_t m2 = MODE == 2 ? m2_in : 0;
    score = W[DP_FC_B];
#pragma unroll 1
    for (int o2 = 0; o2 < 8; ++o2) {
        const float* w = W + DI2_C2 + o2 * 48;                  // pair (o2, ic, k) at (ic*3 + k)*2
        const float2 wd = f2(ldc2(W + DP_FC_W + 2 * o2));
        f32x2 mk[4], z[4];
#pragma unroll
        for (int p = 0; p < 4; ++p) {
            z[p] = MODE == 2 ? pk2(0.f, 0.f) : ldc2(W + DP_C2_B + 2 * o2);
#pragma unroll
            for (int ic = 0; ic < 8; ++ic)
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    const int i = 2 * p + k - 1;
                    if (i >= 0) {
                        const float a = CS_CH(in, ic, i);
                        z[p] = fma2(pk2(a, a), ldc2(w + (ic * 3 + k) * 2), z[p]);
                    }
                }
        }
        float pl_lo = 0.f, pl_hi = 0.f;
        if (MODE == 2) {
            cs_masks4x2(m2, o2, slope, mk);
            f32x2 s = pk2(0.f, 0.f);
#pragma unroll
            for (int p = 0; p < 4; ++p) s = fma2(mk[p], z[p], s);
            upk2(s, pl_lo, pl_hi);                              // d wd[2*o2+h]
        } else {
            uint32_t blo = 0, bhi = 0;
            f32x2 pl = pk2(0.f, 0.f);                           // pooled LeakyReLU of the channel pair: sum_p mk[p] * z[p], packed
#pragma unroll
            for (int p = 0; p < 4; ++p) {
                float lo, hi;
                upk2(z[p], lo, hi);
                const bool plo = lo > 0.f, phi = hi > 0.f;
                blo |= (plo ? 1u : 0u) << p;
                bhi |= (phi ? 1u : 0u) << p;
                mk[p] = pk2(plo ? 1.0f : slope, phi ? 1.0f : slope);
                pl = fma2(mk[p], z[p], pl);
            }
            upk2(pl, pl_lo, pl_hi);
            m2 |= (uint64_t)(blo | (bhi << 4)) << (o2 * 8);
            score = fmaf(wd.x, pl_lo, fmaf(wd.y, pl_hi, score));
        }
        if (MODE >= 1) {
            // lanes 16-31 carry the pair's two output channels swapped (see warp_transpose_reduce_pairs)
            const bool up = (lane & 16) != 0;
            const f32x2 gw = cs_swap_if(up, pk2(g * wd.x, g * wd.y));
            f32x2 mg[4];
#pragma unroll
            for (int p = 0; p < 4; ++p) mg[p] = cs_swap_if(up, mk[p]);
#pragma unroll
            for (int q = 0; q < 2; ++q) {                       // input-channel half: one 32-slot group each
                f32x2 pv[14];
#pragma unroll
                for (int c = 0; c < 4; ++c)
#pragma unroll
                    for (int k = 0; k < 3; ++k) {
                        f32x2 a2 = pk2(0.f, 0.f);
#pragma unroll
                        for (int p = 0; p < 4; ++p) {
                            const int i = 2 * p + k - 1;
                            if (i >= 0) {
                                const float a = CS_CH(in, 4 * q + c, i);
                                a2 = fma2(mg[p], pk2(a, a), a2);
                            }
                        }
                        pv[c * 3 + k] = fma2(a2, gw, pk2(0.f, 0.f));
                    }
                if (q == 0) {
                    const f32x2 pl = cs_swap_if(up, pk2(pl_lo, pl_hi));
                    if (MODE == 1) {
                        pv[12] = mul2(gw, add2(add2(mg[0], mg[1]), add2(mg[2], mg[3])));     // conv2.bias = sum_p dz2
                        pv[13] = mul2(pk2(g, g), pl);                                        // dense.weight = g * pool
                    } else {
                        pv[12] = pk2(0.f, 0.f);                                   // penalty: no bias gradient
                        pv[13] = pl;                                              // d wd
                    }
                    acc.add(4 + 2 * o2 + q, warp_transpose_reduce_pairs<14>(pv, lane));
                } else {
                    acc.add(4 + 2 * o2 + q, warp_transpose_reduce_pairs<12>(pv, lane));
                }
            }
        }
    }
    return m2;
}

// dz1 (channel pairs) = m1 . conv2^T(dz2),  dz2[oc][p] = g * wd[oc] * m2[oc][p]; one output channel of conv2 per iteration,
// two of its input channels per FFMA2
__device__ __forceinline__ void cs_bwd_to_z1(const float* W, float slope, float g, uint64_t m1, uint64_t m2, f32x2 (&dz1)[4][8]) {
#pragma unroll
    for (int c = 0; c < 4; ++c)
#pragma unroll
        for (int i = 0; i < 8; ++i) dz1[c][i] = pk2(0.f, 0.f);
#pragma unroll 1
    for (int oc = 0; oc < 16; ++oc) {
        const float* w = W + DI2_C2T + oc * 24;                 // pair (oc, i2, k) at (i2*3 + k)*2
        const float gw = g * W[DP_FC_W + oc];
        const uint32_t nib = (uint32_t)(m2 >> (oc * 4)) & 15u;
        const float gws = gw * slope;
        float d[4];
#pragma unroll
        for (int p = 0; p < 4; ++p) d[p] = ((nib >> p) & 1u) ? gw : gws;
#pragma unroll
        for (int i2 = 0; i2 < 4; ++i2)
#pragma unroll
            for (int k = 0; k < 3; ++k)
#pragma unroll
                for (int p = 0; p < 4; ++p) {
                    const int i = 2 * p + k - 1;
                    if (i >= 0) dz1[i2][i] = fma2(pk2(d[p], d[p]), ldc2(w + (i2 * 3 + k) * 2), dz1[i2][i]);
                }
    }
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        const uint32_t half = (uint32_t)(m1 >> (c * 16)) & 65535u;
#pragma unroll
        for (int i = 0; i < 8; ++i)
            dz1[c][i] = fma2(dz1[c][i], pk2(((half >> i) & 1u) ? 1.0f : slope, ((half >> (8 + i)) & 1u) ? 1.0f : slope), pk2(0.f, 0.f));
    }
}

// conv1.weight group `ic` from one input row: slot j = oc*3+k -> sum_p dz1[oc][p] * row[2p+k-1]; `extra` fills j = 24..31
__device__ __forceinline__ void cs_grads_conv1_row(const f32x2 (&dz1)[4][8], const float (&row)[16], const float (&extra)[8], int ic,
                                                   SAcc& acc, int lane) {
    float v[32];
#pragma unroll
    for (int o2 = 0; o2 < 4; ++o2)
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            f32x2 a = pk2(0.f, 0.f);
#pragma unroll
            for (int p = 0; p < 8; ++p) {
                const int i = 2 * p + k - 1;
                if (i >= 0) a = fma2(dz1[o2][p], pk2(row[i], row[i]), a);
            }
            upk2(a, v[(2 * o2) * 3 + k], v[(2 * o2 + 1) * 3 + k]);
        }
#pragma unroll
    for (int j = 0; j < 8; ++j) v[24 + j] = extra[j];
    acc.add(ic, warp_transpose_reduce(v, lane));
}

// one row (input channel ic) of conv1^T(dz1): reduction over output-channel pairs, the two halves summed at the end
__device__ __forceinline__ void cs_conv1T_row(const float* W, const f32x2 (&dz1)[4][8], int ic, float (&row)[16]) {
    const float* w = W + DI2_C1 + ic * 6;
    f32x2 r2[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) r2[i] = pk2(0.f, 0.f);
#pragma unroll
    for (int o2 = 0; o2 < 4; ++o2)
#pragma unroll
        for (int p = 0; p < 8; ++p)
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const int i = 2 * p + k - 1;
                if (i >= 0) r2[i] = fma2(dz1[o2][p], ldc2(w + o2 * 24 + k * 2), r2[i]);
            }
#pragma unroll
    for (int i = 0; i < 16; ++i) { const float2 t = f2(r2[i]); row[i] = t.x + t.y; }
}

// ---- one Wasserstein term: L += g * D(cand, cond).  Returns the score; accumulates dL/dtheta.
// NEED_DU: additionally overwrites the thread's rows of t_cand / t_cond with dL/d cand and dL/d cond.
template <bool NEED_DU>
__device__ __forceinline__ float cs_score_pass(const float* W, float slope, float g, float4* t_cand, float4* t_cond, SAcc& acc,
                                               int lane) {
    uint64_t m1, m2;
    float score;
    {
        f32x2 a1[4][8];
        m1 = cs_conv1_fwd<true>(W, slope, t_cand, t_cond, 4, lane, a1);
        m2 = cs_conv2_pass<1>(W, slope, g, a1, 0, acc, lane, score);
    }
    f32x2 dz1[4][8];
    cs_bwd_to_z1(W, slope, g, m1, m2, dz1);
    float c1b[8];                                               // conv1.bias = sum_p dz1
#pragma unroll
    for (int o2 = 0; o2 < 4; ++o2) {
        f32x2 a = pk2(0.f, 0.f);
#pragma unroll
        for (int p = 0; p < 8; ++p) a = fma2(dz1[o2][p], pk2(1.0f, 1.0f), a);
        upk2(a, c1b[2 * o2], c1b[2 * o2 + 1]);
    }
#pragma unroll 1
    for (int ic = 0; ic < 4; ++ic) {
        float row[16], extra[8];
        row_read(ic < 2 ? t_cand : t_cond, lane, ic & 1, row);
#pragma unroll
        for (int j = 0; j < 8; ++j) extra[j] = ic == 0 ? c1b[j] : 0.f;
        if (ic == 1) extra[0] = g;                              // dense.bias
        cs_grads_conv1_row(dz1, row, extra, ic, acc, lane);
    }
    if (NEED_DU) {                                              // du = conv1^T(dz1), one input row per iteration
#pragma unroll 1
        for (int ic = 0; ic < 4; ++ic) {
            float row[16];
            cs_conv1T_row(W, dz1, ic, row);
            row_write(ic < 2 ? t_cand : t_cond, lane, ic & 1, row);
        }
    }
    return score;
}

// forward only: the score of (cand, cond)
__device__ __forceinline__ float cs_score_only(const float* W, float slope, const float4* t_cand, const float4* t_cond, SAcc& acc, int lane,
                                               uint64_t& m1, uint64_t& m2) {
    f32x2 a1[4][8];
    float score;
    m1 = cs_conv1_fwd<true>(W, slope, t_cand, t_cond, 4, lane, a1);
    m2 = cs_conv2_pass<0>(W, slope, 0.f, a1, 0, acc, lane, score);
    return score;
}

// ---- gradient-penalty term of one sample.  t_xh holds x_hat = alpha*real + (1-alpha)*fake (the thread's own slots are
// overwritten with v = d D / d x_hat on the way).  Returns (||v||-1)^2, accumulates scale * d pen / d theta.
__device__ __forceinline__ float cs_gp_pass(const float* W, float slope, float scale, float4* t_xh, const float4* t_cond, SAcc& acc,
                                            int lane, float& norm_out) {
    uint64_t m1, m2;
    cs_score_only(W, slope, t_xh, t_cond, acc, lane, m1, m2);
    f32x2 dz1[4][8];
    cs_bwd_to_z1(W, slope, 1.0f, m1, m2, dz1);
    // v = conv1^T(dz1) restricted to the two candidate rows, one row per iteration, parked in the thread's tile slots
    float n2 = 0.f;
#pragma unroll 1
    for (int ic = 0; ic < 2; ++ic) {
        float row[16];
        cs_conv1T_row(W, dz1, ic, row);
#pragma unroll
        for (int i = 0; i < 16; ++i) n2 = fmaf(row[i], row[i], n2);
        row_write(t_xh, lane, ic, row);
    }
    const float n = sqrtf(n2);
    norm_out = n;
    const float pen = (n - 1.0f) * (n - 1.0f);
    // h = scale * d pen / d v = coef * v, coef = scale * 2 (n-1) / n   (torch norm backward; 0 when n == 0)
    const float coef = n > 0.f ? scale * 2.0f * (n - 1.0f) / n : 0.f;
    // dW1[:, 0:2] += dz1 (x) h
#pragma unroll 1
    for (int ic = 0; ic < 2; ++ic) {
        float row[16], extra[8];
        row_read(t_xh, lane, ic, row);
#pragma unroll
        for (int i = 0; i < 16; ++i) row[i] *= coef;
#pragma unroll
        for (int j = 0; j < 8; ++j) extra[j] = 0.f;
        cs_grads_conv1_row(dz1, row, extra, ic, acc, lane);
    }
    // u1 = m1 . conv1x(h) (no bias): conv1x(v) from the parked rows, then coef and the masks in one packed multiply
    f32x2 u1[4][8];
    const float coef_s = coef * slope;
    cs_conv1_fwd<false>(W, slope, t_xh, t_xh, 2, lane, u1);
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        const uint32_t half = (uint32_t)(m1 >> (c * 16)) & 65535u;
#pragma unroll
        for (int p = 0; p < 8; ++p)
            u1[c][p] = fma2(u1[c][p], pk2(((half >> p) & 1u) ? coef : coef_s, ((half >> (8 + p)) & 1u) ? coef : coef_s), pk2(0.f, 0.f));
    }
    // dW2 += dz2 (x) u1 ; d wd[oc] = sum_p m2 * conv2(u1) ; no bias gradient
    float unused;
    cs_conv2_pass<2>(W, slope, 1.0f, u1, m2, acc, lane, unused);
    return pen;
}

}  // namespace og
