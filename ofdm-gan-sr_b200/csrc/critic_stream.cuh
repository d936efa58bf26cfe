// Register-lean critic passes for the fused training kernels (<= 128 registers per thread, 16 warps per SM).
//
// One sample per thread.  Three things keep the live set (and the code) small:
//   * the sample's frames stay in the warp's shared-memory tile (swizzled, io_tile.cuh) and are re-read ONE 16-float
//     row at a time wherever a pass needs them;
//   * the per-lane parameter-gradient accumulators live in shared memory ([group][thread], conflict free): a group's
//     transpose-reduced sum is added with one LDS + FADD + STS;
//   * every layer is a ROLLED loop over the index that only addresses weights and shared memory (conv2: output
//     channel, conv1: input channel).  All register arrays keep compile-time indices, but a loop iteration is a real
//     basic block: ptxas can no longer hoist the constant loads of a whole layer to the top of a 10k-instruction block
//     and spill them (the fully unrolled version carried 3 KB of stack per thread, essentially the weight image), and
//     the kernel shrinks from ~50k to a few thousand instructions, which the instruction cache likes.
//
// Accumulator slot map (20 groups of 32 = 640 slots):
//   group ic (0..3)    conv1.weight[oc][ic][k] at j = oc*3 + k (24 slots);
//                      spare j = 24..31: group 0 -> conv1.bias[8]; group 1 -> dense.bias (24), sum D(real) (25),
//                      sum D(fake) (26), sum penalty (27)
//   group 4+oc (4..19) conv2.weight[oc][ic][k] at j = ic*3 + k (24 slots), conv2.bias[oc] at j = 24, dense.weight[oc] at 25
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

// parameter index (torch order) -> accumulator slot
__host__ __device__ constexpr int cs_slot_of(int i) {
    return i < DP_C1_B   ? ((i / 3) % 4) * 32 + (i / 12) * 3 + i % 3
           : i < DP_C2_W ? CS_C1B + (i - DP_C1_B)
           : i < DP_C2_B ? (4 + (i - DP_C2_W) / 24) * 32 + (i - DP_C2_W) % 24
           : i < DP_FC_W ? (4 + (i - DP_C2_B)) * 32 + 24
           : i < DP_FC_B ? (4 + (i - DP_FC_W)) * 32 + 25
                         : CS_FCB;
}

// accumulator slot (group, j) -> parameter index (torch order), or -1 for a spare / statistics slot
__host__ __device__ constexpr int cs_param_of(int grp, int j) {
    return grp < 4 ? (j < 24 ? ((j / 3) * 4 + grp) * 3 + j % 3 : (grp == 0 ? DP_C1_B + (j - 24) : (grp == 1 && j == 24 ? DP_FC_B : -1)))
                   : (j < 24 ? DP_C2_W + (grp - 4) * 24 + j : (j == 24 ? DP_C2_B + (grp - 4) : (j == 25 ? DP_FC_W + (grp - 4) : -1)));
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

// LeakyReLU derivatives (1 or slope) of the 4 activations whose sign bits are bits [4*oc, 4*oc+4) of m2
__device__ __forceinline__ void cs_masks4(uint64_t m2, int oc, float slope, float (&mk)[4]) {
    const uint32_t nib = (uint32_t)(m2 >> (oc * 4)) & 15u;
#pragma unroll
    for (int p = 0; p < 4; ++p) mk[p] = ((nib >> p) & 1u) ? 1.0f : slope;
}

// a1 = LeakyReLU(conv1([cand; cond]) + b1); input rows streamed from the tiles; returns the sign mask m1 (bit oc*8+p)
__device__ __forceinline__ uint64_t cs_conv1_fwd(const float* W, float slope, const float4* t_cand, const float4* t_cond, int lane,
                                                 float (&a1)[8][8]) {
#pragma unroll
    for (int oc = 0; oc < 8; ++oc)
#pragma unroll
        for (int p = 0; p < 8; ++p) a1[oc][p] = W[DP_C1_B + oc];
#pragma unroll 1
    for (int ic = 0; ic < 4; ++ic) {
        float row[16];
        row_read(ic < 2 ? t_cand : t_cond, lane, ic & 1, row);
        const float* w = W + DP_C1_W + ic * 3;
#pragma unroll
        for (int oc = 0; oc < 8; ++oc)
#pragma unroll
            for (int p = 0; p < 8; ++p)
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    const int i = 2 * p + k - 1;
                    if (i >= 0) a1[oc][p] = fmaf(w[oc * 12 + k], row[i], a1[oc][p]);
                }
    }
    uint64_t m1 = 0;
#pragma unroll
    for (int oc = 0; oc < 8; ++oc) {
        uint32_t bits = 0;
#pragma unroll
        for (int p = 0; p < 8; ++p) {
            if (a1[oc][p] > 0.f) bits |= 1u << p;
            else a1[oc][p] *= slope;
        }
        m1 |= (uint64_t)bits << (oc * 8);
    }
    return m1;
}

// conv2 + LeakyReLU + sum pool + dense, one output channel per loop iteration.
//   GRADS: also accumulate, for upstream g = dL/dscore, the gradient group of that channel (conv2.weight[oc], conv2.bias[oc],
//   dense.weight[oc]).  Returns m2 (bit oc*4+p: pre-activation > 0) and the score.
template <bool GRADS>
__device__ __forceinline__ uint64_t cs_conv2_fwd(const float* W, float slope, float g, const float (&a1)[8][8], SAcc& acc, int lane,
                                                 float& score) {
    uint64_t m2 = 0;
    score = W[DP_FC_B];
#pragma unroll 1
    for (int oc = 0; oc < 16; ++oc) {
        const float* w = W + DP_C2_W + oc * 24;
        const float bias = W[DP_C2_B + oc], wd = W[DP_FC_W + oc];
        float mk[4], pl = 0.f;
        uint32_t bits = 0;
#pragma unroll
        for (int p = 0; p < 4; ++p) {
            float z = bias;
#pragma unroll
            for (int ic = 0; ic < 8; ++ic)
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    const int i = 2 * p + k - 1;
                    if (i >= 0) z = fmaf(w[ic * 3 + k], a1[ic][i], z);
                }
            const bool pos = z > 0.f;
            bits |= (pos ? 1u : 0u) << p;
            mk[p] = pos ? 1.0f : slope;
            pl += pos ? z : slope * z;
        }
        m2 |= (uint64_t)bits << (oc * 4);
        score = fmaf(wd, pl, score);
        if (GRADS) {
            const float gw = g * wd;
            float v[32];
#pragma unroll
            for (int ic = 0; ic < 8; ++ic)
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    float a = 0.f;
#pragma unroll
                    for (int p = 0; p < 4; ++p) {
                        const int i = 2 * p + k - 1;
                        if (i >= 0) a = fmaf(mk[p], a1[ic][i], a);
                    }
                    v[ic * 3 + k] = gw * a;
                }
            v[24] = gw * ((mk[0] + mk[1]) + (mk[2] + mk[3]));       // conv2.bias[oc] = sum_p dz2
            v[25] = g * pl;                                          // dense.weight[oc]
#pragma unroll
            for (int j = 26; j < 32; ++j) v[j] = 0.f;
            acc.add(4 + oc, warp_transpose_reduce(v, lane));
        }
    }
    return m2;
}

// dz1[8][8] = m1 . conv2^T(dz2),  dz2[oc][p] = g * wd[oc] * m2[oc][p]
__device__ __forceinline__ void cs_bwd_to_z1(const float* W, float slope, float g, uint64_t m1, uint64_t m2, float (&dz1)[8][8]) {
#pragma unroll
    for (int ic = 0; ic < 8; ++ic)
#pragma unroll
        for (int i = 0; i < 8; ++i) dz1[ic][i] = 0.f;
#pragma unroll 1
    for (int oc = 0; oc < 16; ++oc) {
        const float* w = W + DP_C2_W + oc * 24;
        const float gw = g * W[DP_FC_W + oc];
        float d[4];
        cs_masks4(m2, oc, slope, d);
#pragma unroll
        for (int p = 0; p < 4; ++p) d[p] *= gw;
#pragma unroll
        for (int ic = 0; ic < 8; ++ic)
#pragma unroll
            for (int k = 0; k < 3; ++k)
#pragma unroll
                for (int p = 0; p < 4; ++p) {
                    const int i = 2 * p + k - 1;
                    if (i >= 0) dz1[ic][i] = fmaf(w[ic * 3 + k], d[p], dz1[ic][i]);
                }
    }
#pragma unroll
    for (int ic = 0; ic < 8; ++ic) {
        const uint32_t byte = (uint32_t)(m1 >> (ic * 8)) & 255u;
#pragma unroll
        for (int i = 0; i < 8; ++i) dz1[ic][i] *= ((byte >> i) & 1u) ? 1.0f : slope;
    }
}

// conv1.weight group `ic` from one input row: slot j = oc*3+k -> sum_p dz1[oc][p] * row[2p+k-1]; `extra` fills j = 24..31
__device__ __forceinline__ void cs_grads_conv1_row(const float (&dz1)[8][8], const float (&row)[16], const float (&extra)[8], int ic,
                                                   SAcc& acc, int lane) {
    float v[32];
#pragma unroll
    for (int oc = 0; oc < 8; ++oc)
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            float a = 0.f;
#pragma unroll
            for (int p = 0; p < 8; ++p) {
                const int i = 2 * p + k - 1;
                if (i >= 0) a = fmaf(dz1[oc][p], row[i], a);
            }
            v[oc * 3 + k] = a;
        }
#pragma unroll
    for (int j = 0; j < 8; ++j) v[24 + j] = extra[j];
    acc.add(ic, warp_transpose_reduce(v, lane));
}

// ---- one Wasserstein term: L += g * D(cand, cond).  Returns the score; accumulates dL/dtheta.
// NEED_DU: additionally overwrites the thread's rows of t_cand / t_cond with dL/d cand and dL/d cond.
template <bool NEED_DU>
__device__ __forceinline__ float cs_score_pass(const float* W, float slope, float g, float4* t_cand, float4* t_cond, SAcc& acc,
                                               int lane) {
    uint64_t m1, m2;
    float score;
    {
        float a1[8][8];
        m1 = cs_conv1_fwd(W, slope, t_cand, t_cond, lane, a1);
        m2 = cs_conv2_fwd<true>(W, slope, g, a1, acc, lane, score);
    }
    float dz1[8][8];
    cs_bwd_to_z1(W, slope, g, m1, m2, dz1);
    float c1b[8];                                               // conv1.bias = sum_p dz1
#pragma unroll
    for (int oc = 0; oc < 8; ++oc) {
        float a = 0.f;
#pragma unroll
        for (int p = 0; p < 8; ++p) a += dz1[oc][p];
        c1b[oc] = a;
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
            const float* w = W + DP_C1_W + ic * 3;
            float row[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) row[i] = 0.f;
#pragma unroll
            for (int oc = 0; oc < 8; ++oc)
#pragma unroll
                for (int p = 0; p < 8; ++p)
#pragma unroll
                    for (int k = 0; k < 3; ++k) {
                        const int i = 2 * p + k - 1;
                        if (i >= 0) row[i] = fmaf(w[oc * 12 + k], dz1[oc][p], row[i]);
                    }
            row_write(ic < 2 ? t_cand : t_cond, lane, ic & 1, row);
        }
    }
    return score;
}

// ---- gradient-penalty term of one sample.  t_xh holds x_hat = alpha*real + (1-alpha)*fake (the thread's own slots are
// overwritten with v = d D / d x_hat on the way).  Returns (||v||-1)^2, accumulates scale * d pen / d theta.
__device__ __forceinline__ float cs_gp_pass(const float* W, float slope, float scale, float4* t_xh, const float4* t_cond, SAcc& acc,
                                            int lane, float& norm_out) {
    uint64_t m1, m2;
    {
        float a1[8][8], score;
        m1 = cs_conv1_fwd(W, slope, t_xh, t_cond, lane, a1);
        m2 = cs_conv2_fwd<false>(W, slope, 0.f, a1, acc, lane, score);
    }
    float dz1[8][8];
    cs_bwd_to_z1(W, slope, 1.0f, m1, m2, dz1);
    // v = conv1^T(dz1) restricted to the two candidate rows, one row per iteration, parked in the thread's tile slots
    float n2 = 0.f;
#pragma unroll 1
    for (int ic = 0; ic < 2; ++ic) {
        const float* w = W + DP_C1_W + ic * 3;
        float row[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) row[i] = 0.f;
#pragma unroll
        for (int oc = 0; oc < 8; ++oc)
#pragma unroll
            for (int p = 0; p < 8; ++p)
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    const int i = 2 * p + k - 1;
                    if (i >= 0) row[i] = fmaf(w[oc * 12 + k], dz1[oc][p], row[i]);
                }
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
    // u1 = m1 . conv1x(h) (no bias)
    float u1[8][8];
#pragma unroll
    for (int oc = 0; oc < 8; ++oc)
#pragma unroll
        for (int p = 0; p < 8; ++p) u1[oc][p] = 0.f;
#pragma unroll 1
    for (int ic = 0; ic < 2; ++ic) {
        const float* w = W + DP_C1_W + ic * 3;
        float row[16];
        row_read(t_xh, lane, ic, row);
#pragma unroll
        for (int oc = 0; oc < 8; ++oc)
#pragma unroll
            for (int p = 0; p < 8; ++p)
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    const int i = 2 * p + k - 1;
                    if (i >= 0) u1[oc][p] = fmaf(w[oc * 12 + k], row[i], u1[oc][p]);
                }
    }
#pragma unroll
    for (int oc = 0; oc < 8; ++oc) {
        const uint32_t byte = (uint32_t)(m1 >> (oc * 8)) & 255u;
#pragma unroll
        for (int p = 0; p < 8; ++p) u1[oc][p] *= coef * (((byte >> p) & 1u) ? 1.0f : slope);
    }
    // dW2 += dz2 (x) u1 ; d wd[oc] = sum_p m2 * conv2(u1) ; no bias gradient
#pragma unroll 1
    for (int oc = 0; oc < 16; ++oc) {
        const float* w = W + DP_C2_W + oc * 24;
        const float wd = W[DP_FC_W + oc];
        float mk[4], v[32], s = 0.f;
        cs_masks4(m2, oc, slope, mk);
#pragma unroll
        for (int p = 0; p < 4; ++p) {
            float a = 0.f;
#pragma unroll
            for (int ic = 0; ic < 8; ++ic)
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    const int i = 2 * p + k - 1;
                    if (i >= 0) a = fmaf(w[ic * 3 + k], u1[ic][i], a);
                }
            s = fmaf(mk[p], a, s);
        }
#pragma unroll
        for (int ic = 0; ic < 8; ++ic)
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                float a = 0.f;
#pragma unroll
                for (int p = 0; p < 4; ++p) {
                    const int i = 2 * p + k - 1;
                    if (i >= 0) a = fmaf(mk[p], u1[ic][i], a);
                }
                v[ic * 3 + k] = wd * a;
            }
        v[24] = 0.f;
        v[25] = s;
#pragma unroll
        for (int j = 26; j < 32; ++j) v[j] = 0.f;
        acc.add(4 + oc, warp_transpose_reduce(v, lane));
    }
    return pen;
}

}  // namespace og
