// Register-lean MiniGenerator forward-with-tape and backward for the fused generator step (models/generator.py:180-208
// and its autograd graph), in the same style as critic_stream.cuh: rolled loops over the weight-only index of each layer,
// rows parked in the thread's own slots of the warp's shared-memory tiles, gradient accumulators in shared memory.
//
// Works on the folded weight image of weights.cuh: the two convolutions that follow a nearest x2 upsample use 4 folded taps
// {F0=w0, F1=w1+w2, F2=w0+w1, F3=w2} per (oc, ic):   z[2p] = F0 a[p-1] + F1 a[p],   z[2p+1] = F2 a[p] + F3 a[p+1].
//
// Gradient slot map (10 groups of 32 = 320 slots, shared with the finalize kernel in gen_train.cu):
//   G0      out_conv folded dF[2][4][4]     slots   0..31
//   G1-G4   dec1 folded dF[4][8][4]         slots  32..159
//   G5-G7   bottleneck.weight[8][4][3]      slots 160..255
//   G8      enc1.weight[4][2][3] 256..279, enc1.bias[4] 280..283
//   G9      bottleneck.bias[8] 288..295, dec1.bias[4] 296..299, out_conv.bias[2] 300..301, statistics 302..303
#pragma once
#include "critic_stream.cuh"
#include "gen_device.cuh"

namespace og {

constexpr int GSX_NG = 10;
constexpr int GS_OUTF = 0, GS_DECF = 32, GS_BNW = 160, GS_ENCW = 256, GS_ENCB = 280, GS_BNB = 288, GS_DECB = 296, GS_OUTB = 300,
              GS_S0 = 302, GS_SLOTS = 320;

// LOCKSTEP: a CTA-wide barrier between the stages of these passes.  They are ~100 KB of mostly straight-line code that every warp
// runs once per work item: warps that drift apart fetch it from L2 over and over (`no_inst` was 40-45 % of the stall samples in
// front of the FP instructions); kept within a stage of each other they share the SM's 32 KB instruction cache.  Only for kernels
// whose warps ALL run the same sequence (k_gen_step); the control flow around every call site is CTA-uniform.
template <bool LOCKSTEP>
__device__ __forceinline__ void gs_stage_barrier() {
    if (LOCKSTEP) __syncthreads();
}

// the thread's own 32-float slot of a resident tile, addressed in 16-byte chunks c = 0..7
__device__ __forceinline__ float4 slot_ld(const float4* wsm, int lane, int c) { return wsm[lane * 8 + (c ^ (lane & 7))]; }
__device__ __forceinline__ void slot_st(float4* wsm, int lane, int c, float4 v) { wsm[lane * 8 + (c ^ (lane & 7))] = v; }

// [4][8] register array <-> slot (row oc = chunks 2oc, 2oc+1)
__device__ __forceinline__ void park48(float4* wsm, int lane, const float (&a)[4][8]) {
#pragma unroll
    for (int oc = 0; oc < 4; ++oc) {
        slot_st(wsm, lane, 2 * oc, make_float4(a[oc][0], a[oc][1], a[oc][2], a[oc][3]));
        slot_st(wsm, lane, 2 * oc + 1, make_float4(a[oc][4], a[oc][5], a[oc][6], a[oc][7]));
    }
}
__device__ __forceinline__ void unpark48(const float4* wsm, int lane, float (&a)[4][8]) {
#pragma unroll
    for (int oc = 0; oc < 4; ++oc) {
        const float4 u = slot_ld(wsm, lane, 2 * oc), v = slot_ld(wsm, lane, 2 * oc + 1);
        a[oc][0] = u.x; a[oc][1] = u.y; a[oc][2] = u.z; a[oc][3] = u.w;
        a[oc][4] = v.x; a[oc][5] = v.y; a[oc][6] = v.z; a[oc][7] = v.w;
    }
}

// ---- forward.  x rows come from t_x; y rows are written to t_y; t_s is a scratch slot (free on return).
// TAPE: returns a1 = lrelu(enc1), a2 = lrelu(bottleneck), sk = skip sum, z3pos (bit oc*8+q: dec1 pre-activation > 0).
// WRITE_Y = false: the output layer is skipped (the caller already holds what it needs of y; t_y is left alone).
template <bool TAPE, bool WRITE_Y = true, bool LOCKSTEP = false>
__device__ __forceinline__ void gs_fwd(const float* W, float slope, const float4* t_x, float4* t_y, float4* t_s, int lane,
                                       float (&a1)[4][8], float (&a2)[8][4], float (&sk)[4][8], uint32_t& z3pos) {
    // enc1: Conv1d(2->4, k3, s2, p1) + LeakyReLU, one input row per iteration
#pragma unroll
    for (int oc = 0; oc < 4; ++oc)
#pragma unroll
        for (int p = 0; p < 8; ++p) a1[oc][p] = W[GI_ENC_B + oc];
#pragma unroll 1
    for (int ic = 0; ic < 2; ++ic) {
        float row[16];
        row_read(t_x, lane, ic, row);
        const float* w = W + GI_ENC_W + ic * 3;
#pragma unroll
        for (int oc = 0; oc < 4; ++oc)
#pragma unroll
            for (int p = 0; p < 8; ++p)
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    const int i = 2 * p + k - 1;
                    if (i >= 0) a1[oc][p] = fmaf(w[oc * 6 + k], row[i], a1[oc][p]);
                }
    }
#pragma unroll
    for (int oc = 0; oc < 4; ++oc)
#pragma unroll
        for (int p = 0; p < 8; ++p) a1[oc][p] = lrelu(a1[oc][p], slope);
    // bottleneck: Conv1d(4->8, k3, s2, p1) + LeakyReLU, one output channel per iteration (row -> scratch chunk oc)
#pragma unroll 1
    for (int oc = 0; oc < 8; ++oc) {
        const float* w = W + GI_BN_W + oc * 12;
        float z[4];
#pragma unroll
        for (int p = 0; p < 4; ++p) {
            z[p] = W[GI_BN_B + oc];
#pragma unroll
            for (int ic = 0; ic < 4; ++ic)
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    const int i = 2 * p + k - 1;
                    if (i >= 0) z[p] = fmaf(w[ic * 3 + k], a1[ic][i], z[p]);
                }
            z[p] = lrelu(z[p], slope);
        }
        slot_st(t_s, lane, oc, make_float4(z[0], z[1], z[2], z[3]));
    }
#pragma unroll
    for (int oc = 0; oc < 8; ++oc) {
        const float4 v = slot_ld(t_s, lane, oc);
        a2[oc][0] = v.x; a2[oc][1] = v.y; a2[oc][2] = v.z; a2[oc][3] = v.w;
    }
    gs_stage_barrier<LOCKSTEP>();
    // upsample x2 + dec1 Conv1d(8->4, k3, s1, p1) + LeakyReLU, folded; one output channel per iteration
    z3pos = 0;
#pragma unroll 1
    for (int oc = 0; oc < 4; ++oc) {
        const float* F = W + GI_DEC_F + oc * 32;
        float z[8];
#pragma unroll
        for (int p = 0; p < 4; ++p) {
            float e = W[GI_DEC_B + oc], o = e;
#pragma unroll
            for (int ic = 0; ic < 8; ++ic) {
                if (p > 0) e = fmaf(F[ic * 4 + 0], a2[ic][p - 1], e);
                e = fmaf(F[ic * 4 + 1], a2[ic][p], e);
                o = fmaf(F[ic * 4 + 2], a2[ic][p], o);
                if (p < 3) o = fmaf(F[ic * 4 + 3], a2[ic][p + 1], o);
            }
            z[2 * p] = e;
            z[2 * p + 1] = o;
        }
        uint32_t bits = 0;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            if (TAPE) bits |= (z[q] > 0.f ? 1u : 0u) << q;
            z[q] = lrelu(z[q], slope);
        }
        if (TAPE) z3pos |= bits << (oc * 8);
        slot_st(t_s, lane, 2 * oc, make_float4(z[0], z[1], z[2], z[3]));
        slot_st(t_s, lane, 2 * oc + 1, make_float4(z[4], z[5], z[6], z[7]));
    }
    unpark48(t_s, lane, sk);
#pragma unroll
    for (int oc = 0; oc < 4; ++oc)
#pragma unroll
        for (int q = 0; q < 8; ++q) sk[oc][q] += a1[oc][q];                      // additive skip (models/generator.py:199)
    gs_stage_barrier<LOCKSTEP>();
    // upsample x2 + out_conv Conv1d(4->2, k3, s1, p1), folded; tanh; one output row per iteration, written to t_y
#pragma unroll 1
    for (int oc = 0; oc < (WRITE_Y ? 2 : 0); ++oc) {
        const float* F = W + GI_OUT_F + oc * 16;
        float y[16];
#pragma unroll
        for (int p = 0; p < 8; ++p) {
            float e = W[GI_OUT_B + oc], o = e;
#pragma unroll
            for (int ic = 0; ic < 4; ++ic) {
                if (p > 0) e = fmaf(F[ic * 4 + 0], sk[ic][p - 1], e);
                e = fmaf(F[ic * 4 + 1], sk[ic][p], e);
                o = fmaf(F[ic * 4 + 2], sk[ic][p], o);
                if (p < 7) o = fmaf(F[ic * 4 + 3], sk[ic][p + 1], o);
            }
            y[2 * p] = tanh_fast(e);
            y[2 * p + 1] = tanh_fast(o);
        }
        row_write(t_y, lane, oc, y);
    }
}

// ---- backward.  On entry: t_x = input rows, t_y = y rows, t_dy = upstream gradient rows, (a1, a2, sk, z3pos) = tape.
// t_y and t_dy are used as parking space once their contents are consumed; t_p is a further free slot.
// Accumulates the 10 parameter-gradient groups into acc.  NEED_DX: the input gradient rows are left in t_dy.
// THREE_SLOTS (the fused generator step): on entry t_y already holds dz4 = upstream gradient * tanh' (the caller had y in registers
// when it formed the upstream gradient), and t_dy IS the slot of the input rows: it is used as parking space like any other, and the
// input rows are fetched again from x_glob (frames base .. of B) for the last gradient group.  One 4 KB tile per warp less.
template <bool NEED_DX, bool THREE_SLOTS = false, bool LOCKSTEP = false>
__device__ __forceinline__ void gs_bwd(const float* W, float slope, const float4* t_x, float4* t_y, float4* t_dy, float4* t_p, int lane,
                                       float (&a1)[4][8], const float (&a2)[8][4], const float (&sk)[4][8], uint32_t z3pos, SAcc& acc,
                                       const float* x_glob = nullptr, int64_t base = 0, int64_t B = 0) {
    float bias_g[14];                                            // bn.b[8], dec.b[4], out.b[2] -> group 9
    park48(t_p, lane, a1);                                       // a1 is next needed at the bottleneck weight gradient
    // ---- tanh' and the out_conv weight gradient
    float dz4[2][16];
#pragma unroll
    for (int oc = 0; oc < 2; ++oc) {
        float s = 0.f;
        if (THREE_SLOTS) {
            row_read(t_y, lane, oc, dz4[oc]);
#pragma unroll
            for (int q = 0; q < 16; ++q) s += dz4[oc][q];
        } else {
            float y[16];
            row_read(t_dy, lane, oc, dz4[oc]);
            row_read(t_y, lane, oc, y);
#pragma unroll
            for (int q = 0; q < 16; ++q) {
                dz4[oc][q] *= fmaf(-y[q], y[q], 1.0f);
                s += dz4[oc][q];
            }
        }
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
        acc.add(0, warp_transpose_reduce(v, lane));
    }
    gs_stage_barrier<LOCKSTEP>();
    // ---- dsk = out_conv^T(dz4) (one input channel per iteration, row -> t_y chunks 2ic, 2ic+1), dz3 = dsk . lrelu'(z3)
#pragma unroll 1
    for (int ic = 0; ic < 4; ++ic) {
        const float* F = W + GI_OUT_F + ic * 4;
        float r[8];
#pragma unroll
        for (int p = 0; p < 8; ++p) {
            float a = 0.f;
#pragma unroll
            for (int oc = 0; oc < 2; ++oc) {
                a = fmaf(F[oc * 16 + 1], dz4[oc][2 * p], a);
                a = fmaf(F[oc * 16 + 2], dz4[oc][2 * p + 1], a);
                if (p < 7) a = fmaf(F[oc * 16 + 0], dz4[oc][2 * p + 2], a);
                if (p > 0) a = fmaf(F[oc * 16 + 3], dz4[oc][2 * p - 1], a);
            }
            r[p] = a;
        }
        slot_st(t_y, lane, 2 * ic, make_float4(r[0], r[1], r[2], r[3]));
        slot_st(t_y, lane, 2 * ic + 1, make_float4(r[4], r[5], r[6], r[7]));
    }
    float dz3[4][8];
    unpark48(t_y, lane, dz3);                                    // = dsk, which also stays parked in t_y for the skip branch
#pragma unroll
    for (int oc = 0; oc < 4; ++oc) {
        float s = 0.f;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            dz3[oc][q] *= ((z3pos >> (oc * 8 + q)) & 1u) ? 1.0f : slope;
            s += dz3[oc][q];
        }
        bias_g[8 + oc] = s;
    }
    gs_stage_barrier<LOCKSTEP>();
    // ---- dec1 folded weight gradient (4 groups, one per output channel) - no weights involved, unrolled
#pragma unroll
    for (int grp = 0; grp < 4; ++grp) {
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            const int oc = grp, ic = j / 4, t = j % 4;
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
        acc.add(1 + grp, warp_transpose_reduce(v, lane));
        if (grp & 1) gs_stage_barrier<LOCKSTEP>();
    }
    // ---- dz2 = dec1^T(dz3) . lrelu'(a2): one input channel per iteration, raw row -> t_dy chunk ic
#pragma unroll 1
    for (int ic = 0; ic < 8; ++ic) {
        const float* F = W + GI_DEC_F + ic * 4;
        float r[4];
#pragma unroll
        for (int p = 0; p < 4; ++p) {
            float a = 0.f;
#pragma unroll
            for (int oc = 0; oc < 4; ++oc) {
                a = fmaf(F[oc * 32 + 1], dz3[oc][2 * p], a);
                a = fmaf(F[oc * 32 + 2], dz3[oc][2 * p + 1], a);
                if (p < 3) a = fmaf(F[oc * 32 + 0], dz3[oc][2 * p + 2], a);
                if (p > 0) a = fmaf(F[oc * 32 + 3], dz3[oc][2 * p - 1], a);
            }
            r[p] = a;
        }
        slot_st(t_dy, lane, ic, make_float4(r[0], r[1], r[2], r[3]));
    }
    float dz2[8][4];
#pragma unroll
    for (int ic = 0; ic < 8; ++ic) {
        const float4 v = slot_ld(t_dy, lane, ic);
        dz2[ic][0] = v.x; dz2[ic][1] = v.y; dz2[ic][2] = v.z; dz2[ic][3] = v.w;
        float s = 0.f;
#pragma unroll
        for (int p = 0; p < 4; ++p) {
            dz2[ic][p] *= a2[ic][p] > 0.f ? 1.0f : slope;
            s += dz2[ic][p];
        }
        bias_g[ic] = s;
    }
    gs_stage_barrier<LOCKSTEP>();
    // ---- bottleneck weight gradient (3 groups) - needs a1 back
    unpark48(t_p, lane, a1);
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
        acc.add(5 + grp, warp_transpose_reduce(v, lane));
    }
    gs_stage_barrier<LOCKSTEP>();
    // ---- dz1 = (dsk + bottleneck^T(dz2)) . lrelu'(a1): one input channel per iteration, raw row -> t_dy chunks 2ic, 2ic+1
    // (t_dy's dz2 rows are consumed: dz2 is in registers)
#pragma unroll 1
    for (int ic = 0; ic < 4; ++ic) {
        const float* w = W + GI_BN_W + ic * 3;
        float r[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) r[i] = 0.f;
#pragma unroll
        for (int oc = 0; oc < 8; ++oc)
#pragma unroll
            for (int p = 0; p < 4; ++p)
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    const int i = 2 * p + k - 1;
                    if (i >= 0) r[i] = fmaf(w[oc * 12 + k], dz2[oc][p], r[i]);
                }
        slot_st(t_dy, lane, 2 * ic, make_float4(r[0], r[1], r[2], r[3]));
        slot_st(t_dy, lane, 2 * ic + 1, make_float4(r[4], r[5], r[6], r[7]));
    }
    float dz1[4][8];
    {
        float dsk[4][8];
        unpark48(t_dy, lane, dz1);
        unpark48(t_y, lane, dsk);
#pragma unroll
        for (int ic = 0; ic < 4; ++ic)
#pragma unroll
            for (int i = 0; i < 8; ++i) dz1[ic][i] = (dz1[ic][i] + dsk[ic][i]) * (a1[ic][i] > 0.f ? 1.0f : slope);
    }
    gs_stage_barrier<LOCKSTEP>();
    // ---- enc1 weight + bias gradient (group 8)
    if (THREE_SLOTS) {                                           // t_dy's parked rows are consumed: bring the input rows back
        __syncwarp();
        tile_fill_f32(x_glob, base, B, t_dy, lane);
        __syncwarp();
    }
    {
        float x[2][16], v[32];
        row_read(t_x, lane, 0, x[0]);
        row_read(t_x, lane, 1, x[1]);
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
        acc.add(8, warp_transpose_reduce(v, lane));
    }
    {
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = j < 14 ? bias_g[j] : 0.f;
        acc.add(9, warp_transpose_reduce(v, lane));
    }
    if (NEED_DX) {                                               // dx = enc1^T(dz1), one input row per iteration -> t_dy
#pragma unroll 1
        for (int ic = 0; ic < 2; ++ic) {
            const float* w = W + GI_ENC_W + ic * 3;
            float row[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) row[i] = 0.f;
#pragma unroll
            for (int oc = 0; oc < 4; ++oc)
#pragma unroll
                for (int p = 0; p < 8; ++p)
#pragma unroll
                    for (int k = 0; k < 3; ++k) {
                        const int i = 2 * p + k - 1;
                        if (i >= 0) row[i] = fmaf(w[oc * 6 + k], dz1[oc][p], row[i]);
                    }
            row_write(t_dy, lane, ic, row);
        }
    }
}

}  // namespace og
