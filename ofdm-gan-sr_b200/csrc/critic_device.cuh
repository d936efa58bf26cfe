// Per-thread (one sample per thread) critic maths: MiniDiscriminator forward, its backward, and the closed-form
// gradient-penalty double backward.  models/discriminator.py:112-152 (forward), :172-236 (penalty),
// train.py:228-250 (loss).  D = conv1(4->8,k3,s2,p1) -> LReLU -> conv2(8->16,k3,s2,p1) -> LReLU -> sum_L -> dense.
//
// Gradient accumulation.  A thread produces, for its own sample, one partial per parameter.  Partials are produced in
// groups of 32; each group is summed over the warp's 32 samples with the transpose-reduce of common.cuh (lane l ends
// up with slot l of the group) and added to a per-lane register accumulator that lives across all the tiles the warp
// processes: GradAcc<NG> = NG registers per lane = 32*NG slots per warp.
//
// Critic slot map (17 groups = 544 slots):
//   G0-G2   conv1.weight[8][4][3]   slots   0..95    (flat torch order)
//   G3      conv1.bias[8] 96..103, dense.bias 104, stats: sum D(real) 105, sum D(fake) 106, sum gp 107
//   G4-G15  conv2.weight[16][8][3]  slots 128..511   (flat torch order)
//   G16     conv2.bias[16] 512..527, dense.weight[16] 528..543
#pragma once
#include "weights.cuh"

namespace og {

constexpr int D_NG = 17;
constexpr int DS_C1W = 0, DS_C1B = 96, DS_FCB = 104, DS_SREAL = 105, DS_SFAKE = 106, DS_SGP = 107, DS_C2W = 128,
              DS_C2B = 512, DS_FCW = 528, DS_SLOTS = 544;

__device__ __forceinline__ float mask_of(uint64_t m, int bit, float slope) { return ((m >> bit) & 1ull) ? 1.0f : slope; }

// ---- forward ---------------------------------------------------------------------------------------------
// cand/cond [2][16] -> a1[8][8] (post-LReLU), m2 (bit oc*4+p: conv2 pre-activation > 0), pool[16], score
__device__ __forceinline__ void disc_fwd(const float* __restrict__ W, float slope, const float (&cand)[2][16],
                                         const float (&cond)[2][16], float (&a1)[8][8], uint64_t& m2, float (&pool)[16],
                                         float& score) {
#pragma unroll
    for (int oc = 0; oc < 8; ++oc)
#pragma unroll
        for (int p = 0; p < 8; ++p) {
            float acc = W[DP_C1_B + oc];
#pragma unroll
            for (int ic = 0; ic < 4; ++ic)
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    const int i = 2 * p + k - 1;
                    if (i >= 0) acc = fmaf(W[DP_C1_W + (oc * 4 + ic) * 3 + k], ic < 2 ? cand[ic][i] : cond[ic - 2][i], acc);
                }
            a1[oc][p] = acc > 0.f ? acc : slope * acc;
        }
    m2 = 0;
    score = W[DP_FC_B];
#pragma unroll
    for (int oc = 0; oc < 16; ++oc) {
        float pl = 0.f;
#pragma unroll
        for (int p = 0; p < 4; ++p) {
            float acc = W[DP_C2_B + oc];
#pragma unroll
            for (int ic = 0; ic < 8; ++ic)
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    const int i = 2 * p + k - 1;
                    if (i >= 0) acc = fmaf(W[DP_C2_W + (oc * 8 + ic) * 3 + k], a1[ic][i], acc);
                }
            if (acc > 0.f) m2 |= 1ull << (oc * 4 + p);
            pl += acc > 0.f ? acc : slope * acc;
        }
        pool[oc] = pl;
        score = fmaf(W[DP_FC_W + oc], pl, score);
    }
}

__device__ __forceinline__ uint64_t sign_mask(const float (&a)[8][8]) {
    uint64_t m = 0;
#pragma unroll
    for (int c = 0; c < 8; ++c)
#pragma unroll
        for (int p = 0; p < 8; ++p)
            if (a[c][p] > 0.f) m |= 1ull << (c * 8 + p);
    return m;
}

// dz1[8][8] = m1 . conv2^T(dz2),  dz2[oc][p] = g * wd[oc] * m2[oc][p]
__device__ __forceinline__ void disc_bwd_to_z1(const float* __restrict__ W, float slope, float g, uint64_t m1, uint64_t m2,
                                               float (&dz1)[8][8]) {
#pragma unroll
    for (int ic = 0; ic < 8; ++ic)
#pragma unroll
        for (int i = 0; i < 8; ++i) dz1[ic][i] = 0.f;
#pragma unroll
    for (int oc = 0; oc < 16; ++oc) {
        const float gw = g * W[DP_FC_W + oc];
#pragma unroll
        for (int p = 0; p < 4; ++p) {
            const float d = gw * mask_of(m2, oc * 4 + p, slope);
#pragma unroll
            for (int ic = 0; ic < 8; ++ic)
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    const int i = 2 * p + k - 1;
                    if (i >= 0) dz1[ic][i] = fmaf(W[DP_C2_W + (oc * 8 + ic) * 3 + k], d, dz1[ic][i]);
                }
        }
    }
#pragma unroll
    for (int ic = 0; ic < 8; ++ic)
#pragma unroll
        for (int i = 0; i < 8; ++i) dz1[ic][i] *= mask_of(m1, ic * 8 + i, slope);
}

// conv1^T(dz1) restricted to in-channels [IC0, IC0+NIC)
template <int IC0, int NIC>
__device__ __forceinline__ void disc_bwd_to_input(const float* __restrict__ W, const float (&dz1)[8][8], float (&du)[NIC][16]) {
#pragma unroll
    for (int ic = 0; ic < NIC; ++ic)
#pragma unroll
        for (int i = 0; i < 16; ++i) du[ic][i] = 0.f;
#pragma unroll
    for (int oc = 0; oc < 8; ++oc)
#pragma unroll
        for (int p = 0; p < 8; ++p)
#pragma unroll
            for (int ic = 0; ic < NIC; ++ic)
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    const int i = 2 * p + k - 1;
                    if (i >= 0) du[ic][i] = fmaf(W[DP_C1_W + (oc * 4 + IC0 + ic) * 3 + k], dz1[oc][p], du[ic][i]);
                }
}

// ---- parameter-gradient groups -----------------------------------------------------------------------------
// conv2.weight groups G4..G15: slot s = (oc*8+ic)*3+k ; partial = sum_p dz2[oc][p] * in[ic][2p+k-1]
// dz2[oc][p] = gw[oc] * mask(m2): rebuilt per group from the mask bits, so only `in` (64 regs) stays live.
__device__ __forceinline__ void grads_conv2_w(const float* __restrict__ W, float slope, float g, uint64_t m2,
                                              const float (&in)[8][8], GradAcc<D_NG>& acc, int lane) {
#pragma unroll
    for (int grp = 0; grp < 12; ++grp) {
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            const int s = grp * 32 + j, oc = s / 24, ic = (s / 3) % 8, k = s % 3;
            const float gw = g * W[DP_FC_W + oc];
            float a = 0.f;
#pragma unroll
            for (int p = 0; p < 4; ++p) {
                const int i = 2 * p + k - 1;
                if (i >= 0) a = fmaf(mask_of(m2, oc * 4 + p, slope), in[ic][i], a);
            }
            v[j] = gw * a;
        }
        acc.g[4 + grp] += warp_transpose_reduce(v, lane);
    }
}

// conv1.weight groups G0..G2: slot s = (oc*4+ic)*3+k ; partial = sum_p dz1[oc][p] * u[ic][2p+k-1]
// NIC = 4: u rows 0,1 = candidate, 2,3 = condition.  NIC = 2: only the candidate rows get a gradient (the penalty).
template <int NIC>
__device__ __forceinline__ void grads_conv1_w(const float (&dz1)[8][8], const float (&u)[NIC][16], GradAcc<D_NG>& acc, int lane) {
#pragma unroll
    for (int grp = 0; grp < 3; ++grp) {
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            const int s = grp * 32 + j, oc = s / 12, ic = (s / 3) % 4, k = s % 3;
            float a = 0.f;
            if (ic < NIC) {
#pragma unroll
                for (int p = 0; p < 8; ++p) {
                    const int i = 2 * p + k - 1;
                    if (i >= 0) a = fmaf(dz1[oc][p], u[ic < NIC ? ic : 0][i], a);
                }
            }
            v[j] = a;
        }
        acc.g[grp] += warp_transpose_reduce(v, lane);
    }
}

// G16: conv2.bias = sum_p dz2 ; dense.weight = g * pool
__device__ __forceinline__ void grads_c2b_fcw(const float* __restrict__ W, float slope, float g, uint64_t m2,
                                              const float (&pool)[16], GradAcc<D_NG>& acc, int lane) {
    float v[32];
#pragma unroll
    for (int oc = 0; oc < 16; ++oc) {
        float a = 0.f;
#pragma unroll
        for (int p = 0; p < 4; ++p) a += mask_of(m2, oc * 4 + p, slope);
        v[oc] = g * W[DP_FC_W + oc] * a;
        v[16 + oc] = g * pool[oc];
    }
    acc.g[16] += warp_transpose_reduce(v, lane);
}

// G3: conv1.bias = sum_p dz1 ; dense.bias = g.  (The statistic slots of G3 are filled once, at the end of the kernel.)
__device__ __forceinline__ void grads_c1b_fcb(const float (&dz1)[8][8], float g, GradAcc<D_NG>& acc, int lane) {
    float v[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = 0.f;
#pragma unroll
    for (int oc = 0; oc < 8; ++oc) {
        float a = 0.f;
#pragma unroll
        for (int p = 0; p < 8; ++p) a += dz1[oc][p];
        v[oc] = a;
    }
    v[8] = g;
    acc.g[3] += warp_transpose_reduce(v, lane);
}

// One score pass: L += g * D(cand, cond).  Accumulates dL/dtheta; returns the score.  If NEED_DU, also returns
// dL/d[cand;cond] (4 rows).
template <bool NEED_DU>
__device__ __forceinline__ float critic_score_pass(const float* __restrict__ W, float slope, float g, const float (&cand)[2][16],
                                                   const float (&cond)[2][16], bool want_grads, GradAcc<D_NG>& acc, int lane,
                                                   float (&du)[4][16]) {
    float a1[8][8], pool[16], score;
    uint64_t m2;
    disc_fwd(W, slope, cand, cond, a1, m2, pool, score);
    if (want_grads) {
        grads_conv2_w(W, slope, g, m2, a1, acc, lane);
        grads_c2b_fcw(W, slope, g, m2, pool, acc, lane);
    }
    const uint64_t m1 = sign_mask(a1);
    float dz1[8][8];
    disc_bwd_to_z1(W, slope, g, m1, m2, dz1);
    if (want_grads) {
        float u[4][16];
#pragma unroll
        for (int i = 0; i < 16; ++i) { u[0][i] = cand[0][i]; u[1][i] = cand[1][i]; u[2][i] = cond[0][i]; u[3][i] = cond[1][i]; }
        grads_conv1_w<4>(dz1, u, acc, lane);
        grads_c1b_fcb(dz1, g, acc, lane);
    }
    if (NEED_DU) disc_bwd_to_input<0, 4>(W, dz1, du);
    return score;
}

// Gradient-penalty pass for one sample: xh = alpha*real + (1-alpha)*fake.  Returns (||grad||-1)^2 and accumulates
// scale * d(penalty)/dtheta (the double backward, closed form - see oracle/fp32_models.c gp_sample and SURVEY 3.4).
__device__ __forceinline__ float critic_gp_pass(const float* __restrict__ W, float slope, float scale, const float (&xh)[2][16],
                                                const float (&cond)[2][16], bool want_grads, GradAcc<D_NG>& acc, int lane,
                                                float& norm_out) {
    uint64_t m1, m2;
    {
        float a1[8][8], pool[16], score;
        disc_fwd(W, slope, xh, cond, a1, m2, pool, score);
        m1 = sign_mask(a1);
    }
    float dz1[8][8];
    disc_bwd_to_z1(W, slope, 1.0f, m1, m2, dz1);
    float h[2][16];
    disc_bwd_to_input<0, 2>(W, dz1, h);                      // h = v = dD/dxh for now
    float n2 = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) n2 = fmaf(h[0][i], h[0][i], fmaf(h[1][i], h[1][i], n2));
    const float n = sqrtf(n2);
    norm_out = n;
    const float pen = (n - 1.0f) * (n - 1.0f);
    if (!want_grads) return pen;
    const float coef = n > 0.f ? scale * 2.0f * (n - 1.0f) / n : 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) { h[0][i] *= coef; h[1][i] *= coef; }
    // v = conv1x^T(dz1): dW1[:, 0:2] += dz1 (x) h
    grads_conv1_w<2>(dz1, h, acc, lane);
    // d(dz1) = conv1x(h) (no bias); u1 = m1 . d(dz1)
    float u1[8][8];
#pragma unroll
    for (int oc = 0; oc < 8; ++oc)
#pragma unroll
        for (int p = 0; p < 8; ++p) {
            float a = 0.f;
#pragma unroll
            for (int ic = 0; ic < 2; ++ic)
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    const int i = 2 * p + k - 1;
                    if (i >= 0) a = fmaf(W[DP_C1_W + (oc * 4 + ic) * 3 + k], h[ic][i], a);
                }
            u1[oc][p] = a * mask_of(m1, oc * 8 + p, slope);
        }
    // da1 = conv2^T(dz2), dz2 = wd (x) m2: dW2 += dz2 (x) u1 ; d(dz2) = conv2(u1) ; d wd[oc] = sum_p m2 * d(dz2)
    grads_conv2_w(W, slope, 1.0f, m2, u1, acc, lane);
    float v[32];
#pragma unroll
    for (int oc = 0; oc < 16; ++oc) {
        float s = 0.f;
#pragma unroll
        for (int p = 0; p < 4; ++p) {
            float a = 0.f;
#pragma unroll
            for (int ic = 0; ic < 8; ++ic)
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    const int i = 2 * p + k - 1;
                    if (i >= 0) a = fmaf(W[DP_C2_W + (oc * 8 + ic) * 3 + k], u1[ic][i], a);
                }
            s = fmaf(mask_of(m2, oc * 4 + p, slope), a, s);
        }
        v[oc] = 0.f;             // conv2.bias: the penalty has no bias gradient
        v[16 + oc] = s;
    }
    acc.g[16] += warp_transpose_reduce(v, lane);
    return pen;
}

}  // namespace og
