// EXPERIMENT, NOT PART OF THE BUILD (kept for the record; results in profiles/r1_notes.md, "frame per pair of threads").
// To try it again: copy next to sim_kernel.cuh, include it from sim_gauss.cu instead of sim_kernel.cuh and route
// sim_launch_gauss through pair_eligible() / sim_launch_pair().  It passed every GPU parity test unchanged.
// The headline instantiation of the fused path (Gaussian symbols, all randomness from Philox, fp32 generator, GAN + NoEQ rows)
// with ONE FRAME PER PAIR OF THREADS: lane 2i holds samples 0..7, lane 2i+1 samples 8..15 of every per-frame array.
//
// Why: k_sim keeps a frame in one thread and needs 128 registers, i.e. four warps per scheduler, and every capture shows
// issue utilisation = 4 warps / (6.5 cycles average latency per instruction) (profiles/r1_notes.md).  Halving the per-thread
// state buys more resident warps for the same work.  Same Philox counters, same operations in the same order wherever the order
// is observable; what moves is who computes what:
//   * symbols: thread h draws blocks {2h, 2h+1} (Re) and {4+2h, 4+2h+1} (Im) = bins 8h..8h+7; one exchange hands each thread
//     the even (h = 0) or odd (h = 1) bins, which is exactly the split of the radix-2 DIT network of fft_inplace<16> before its
//     last stage: both run fft_inplace<8>, thread 1 applies the last-stage twiddles, one more exchange, add / subtract;
//   * Rapp PA, IQ imbalance, AWGN (blocks 13+2h.., 17+2h..) are per sample; the phase-noise random walk is sequential, so thread 1
//     starts its eight steps from thread 0's last angle (one shuffle);
//   * frame power, maximum, squared errors: per-thread partials + one shuffle each;
//   * generator: thread h computes output positions of its half; every layer needs one halo column from the partner
//     (a zero halo stands in for the padding taps k_sim skips: fma(0, w, acc) == acc).
#pragma once
#include <cstdlib>

#include "sim_kernel.cuh"

#ifndef OG_PAIR_THREADS
#define OG_PAIR_THREADS 768
#endif
#ifndef OG_PAIR_BARRIER
#define OG_PAIR_BARRIER 1
#endif

namespace og {

constexpr int PT = OG_PAIR_THREADS, PFR = PT / 2;              // threads per CTA (one CTA per SM), frames per tile
constexpr size_t PAIR_SMEM = OFDMGAN_MAX_SNR_BINS * NM * NC * sizeof(double);

__device__ __forceinline__ float px(float v) { return __shfl_xor_sync(0xffffffffu, v, 1); }

// clean half-frame cr/ci[8] (samples 8h..8h+7)
__device__ __forceinline__ void pair_tx(const SimArgs& a, uint64_t frame, int h, float (&cr)[8], float (&ci)[8]) {
    const float sc = 0.70710678118654752f * (a.cfg.ifft_scale == OFDMGAN_SCALE_N ? 1.0f : 0.25f), var = sc * sc;
    float Xr[8], Xi[8];                                          // bins 8h + i
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        float n[4];
        draw_normals(a, frame, (uint32_t)(2 * h + j), n, var);
#pragma unroll
        for (int t = 0; t < 4; ++t) Xr[4 * j + t] = n[t];
        draw_normals(a, frame, (uint32_t)(4 + 2 * h + j), n, var);
#pragma unroll
        for (int t = 0; t < 4; ++t) Xi[4 * j + t] = n[t];
    }
    // thread 0 wants the even bins (its own 0,2,4,6 and the partner's 8,10,12,14), thread 1 the odd ones
    float er[8], ei[8];
#pragma unroll
    for (int m = 0; m < 4; ++m) {
        const float sr = h ? Xr[2 * m] : Xr[2 * m + 1], si = h ? Xi[2 * m] : Xi[2 * m + 1];
        const float rr = px(sr), ri = px(si);
        er[m] = h ? rr : Xr[2 * m];         ei[m] = h ? ri : Xi[2 * m];                 // sub-sequence index m:  bin 2m + h
        er[4 + m] = h ? Xr[2 * m + 1] : rr; ei[4 + m] = h ? Xi[2 * m + 1] : ri;
    }
    fft_inplace<8, +1>(er, ei);                                  // stages 1-3 of the 16-point network on this half
    // last stage (fft_inplace<16>, st = 4): x = W^j O[j];  out[j] = E[j] + x,  out[j + 8] = E[j] - x
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const float wr = Tw<16>::c(j), wi = Tw<16>::s(j);
        float xr, xi;
        if (j == 0) { xr = er[j]; xi = ei[j]; }
        else if (j == 4) { xr = -ei[j]; xi = er[j]; }
        else { xr = er[j] * wr - ei[j] * wi; xi = er[j] * wi + ei[j] * wr; }
        const float vr = h ? xr : er[j], vi = h ? xi : ei[j];   // thread 1 sends its rotated O[j], thread 0 its E[j]
        const float pr = px(vr), pi = px(vi);
        cr[j] = h ? pr - xr : er[j] + pr;
        ci[j] = h ? pi - xi : ei[j] + pi;
    }
}

// impairments + AWGN on a copy of the clean half-frame (impair_channel<false> of chan_device.cuh, split over the pair)
__device__ __forceinline__ void pair_impair(const SimArgs& a, uint64_t frame, int h, float snr_db, const float (&cr)[8],
                                            const float (&ci)[8], float (&nr)[8], float (&ni)[8]) {
    const ofdmgan_chan_cfg& c = a.cfg;
#pragma unroll
    for (int i = 0; i < 8; ++i) { nr[i] = cr[i]; ni[i] = ci[i]; }
    if (c.impair & OFDMGAN_IMPAIR_PA) {
        const float invA2 = 1.0f / (c.pa_saturation * c.pa_saturation);
        const float p = c.pa_smoothness, ninv2p = -0.5f / p;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float t = (nr[i] * nr[i] + ni[i] * ni[i]) * invA2;
            const float yp = p == 3.0f ? t * t * t : fast_ex2(p * fast_lg2(t));
            const float gain = fast_ex2(ninv2p * fast_lg2(1.0f + yp));
            nr[i] *= gain; ni[i] *= gain;
        }
    }
    if (c.impair & OFDMGAN_IMPAIR_IQ) {
#pragma unroll
        for (int i = 0; i < 8; ++i) ni[i] = c.iq_gain * (c.iq_cos * ni[i] + c.iq_sin * nr[i]);
    }
    if (c.impair & OFDMGAN_IMPAIR_PN) {
        float inc[8];
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            float n[4];
            draw_normals(a, frame, (uint32_t)(8 + 2 * h + j), n);
#pragma unroll
            for (int t = 0; t < 4; ++t) inc[4 * j + t] = n[t];
        }
        float th = 0.f;                                          // thread 0's walk; its end point is where thread 1 starts
#pragma unroll
        for (int i = 0; i < 8; ++i) th = fmaf(c.pn_sigma, inc[i], th);
        const float th7 = px(th);
        th = h ? th7 : 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            th = fmaf(c.pn_sigma, inc[i], th);
            const float red = fmaf(-6.283185307179586f, rintf(th * 0.15915494309189535f), th);
            const float s = fast_sin(red), co = fast_cos(red);
            const float xr = nr[i], xi = ni[i];
            nr[i] = xr * co - xi * s;
            ni[i] = xr * s + xi * co;
        }
    }
    float P = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) P = fmaf(nr[i], nr[i], fmaf(ni[i], ni[i], P));
    P = (P + px(P)) * 0.0625f;
    const float sd = fast_sqrt(0.5f * P * fast_ex2(-0.33219280948873623f * snr_db));
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        float n[4], m[4];
        draw_normals(a, frame, (uint32_t)(13 + 2 * h + j), n);
        draw_normals(a, frame, (uint32_t)(17 + 2 * h + j), m);
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            nr[4 * j + t] = fmaf(sd, n[t], nr[4 * j + t]);
            ni[4 * j + t] = fmaf(sd, m[t], ni[4 * j + t]);
        }
    }
}

__device__ __forceinline__ float max_abs8(const float (&r)[8], const float (&i)[8]) {
    float m = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) m = fmaxf(m, fmaxf(fabsf(r[k]), fabsf(i[k])));
    return m;
}

__device__ __forceinline__ void pair_normalise(int mode, float (&cr)[8], float (&ci)[8], float (&nr)[8], float (&ni)[8]) {
    if (mode == OFDMGAN_NORM_NONE) return;
    float mc = max_abs8(cr, ci), mn = max_abs8(nr, ni);
    mc = fmaxf(mc, px(mc));
    mn = fmaxf(mn, px(mn));
    if (mode == OFDMGAN_NORM_JOINT) mc = mn = fmaxf(mc, mn);
    const float sc = mc > 0.f ? 1.0f / mc : 1.0f, sn = mn > 0.f ? 1.0f / mn : 1.0f;
#pragma unroll
    for (int k = 0; k < 8; ++k) { cr[k] *= sc; ci[k] *= sc; nr[k] *= sn; ni[k] *= sn; }
}

// MiniGenerator forward on half a frame per thread (gen_fwd_f32_infer of gen_device.cuh with halo exchanges).
// x[ic][i] <-> sample 8h + i;  y likewise.
__device__ __forceinline__ void pair_gen_fwd(const float* __restrict__ W, float slope, int h, const float (&x)[2][8], float (&y)[2][8]) {
    float a1[4][4], a2[8][2], sk[4][4];
    // enc1: output positions 4h + p, taps at local samples 2p - 1 .. 2p + 1; the left halo is the partner's sample 7
    float xl[2];
#pragma unroll
    for (int ic = 0; ic < 2; ++ic) { const float v = px(x[ic][7]); xl[ic] = h ? v : 0.f; }
#pragma unroll
    for (int o2 = 0; o2 < 2; ++o2)
#pragma unroll
        for (int p = 0; p < 4; ++p) {
            f32x2 acc = ldc2(W + GI_ENC_B + 2 * o2);
#pragma unroll
            for (int ic = 0; ic < 2; ++ic)
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    const int i = 2 * p + k - 1;
                    const float v = i >= 0 ? x[ic][i] : xl[ic];
                    acc = fma2(pk2(v, v), ldc2(W + GI2_ENC + ((o2 * 2 + ic) * 3 + k) * 2), acc);
                }
            float lo, hi;
            upk2(acc, lo, hi);
            a1[2 * o2][p] = lrelu(lo, slope);
            a1[2 * o2 + 1][p] = lrelu(hi, slope);
        }
    // bottleneck: output positions 2h + p
    float al[4];
#pragma unroll
    for (int ic = 0; ic < 4; ++ic) { const float v = px(a1[ic][3]); al[ic] = h ? v : 0.f; }
#pragma unroll
    for (int o2 = 0; o2 < 4; ++o2)
#pragma unroll
        for (int p = 0; p < 2; ++p) {
            f32x2 acc = ldc2(W + GI_BN_B + 2 * o2);
#pragma unroll
            for (int ic = 0; ic < 4; ++ic)
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    const int i = 2 * p + k - 1;
                    const float v = i >= 0 ? a1[ic][i] : al[ic];
                    acc = fma2(pk2(v, v), ldc2(W + GI2_BN + ((o2 * 4 + ic) * 3 + k) * 2), acc);
                }
            float lo, hi;
            upk2(acc, lo, hi);
            a2[2 * o2][p] = lrelu(lo, slope);
            a2[2 * o2 + 1][p] = lrelu(hi, slope);
        }
    // upsample x2 + dec1 (folded taps) + skip: bottleneck positions 2h + p; halos: left = partner's column 1 (thread 1),
    // right = partner's column 0 (thread 0); the frame edges get zeros
    float bl[8], br[8];
#pragma unroll
    for (int ic = 0; ic < 8; ++ic) {
        const float v = px(h ? a2[ic][0] : a2[ic][1]);
        bl[ic] = h ? v : 0.f;
        br[ic] = h ? 0.f : v;
    }
#pragma unroll
    for (int o2 = 0; o2 < 2; ++o2)
#pragma unroll
        for (int p = 0; p < 2; ++p) {
            f32x2 e = ldc2(W + GI_DEC_B + 2 * o2), o = e;
#pragma unroll
            for (int ic = 0; ic < 8; ++ic) {
                const float* F = W + GI2_DEC + (o2 * 8 + ic) * 8;
                const float prev = p > 0 ? a2[ic][p - 1] : bl[ic], next = p < 1 ? a2[ic][p + 1] : br[ic];
                e = fma2(pk2(prev, prev), ldc2(F + 0), e);
                e = fma2(pk2(a2[ic][p], a2[ic][p]), ldc2(F + 2), e);
                o = fma2(pk2(a2[ic][p], a2[ic][p]), ldc2(F + 4), o);
                o = fma2(pk2(next, next), ldc2(F + 6), o);
            }
            float e0, e1, o0, o1;
            upk2(e, e0, e1);
            upk2(o, o0, o1);
            sk[2 * o2][2 * p] = lrelu(e0, slope) + a1[2 * o2][2 * p];
            sk[2 * o2][2 * p + 1] = lrelu(o0, slope) + a1[2 * o2][2 * p + 1];
            sk[2 * o2 + 1][2 * p] = lrelu(e1, slope) + a1[2 * o2 + 1][2 * p];
            sk[2 * o2 + 1][2 * p + 1] = lrelu(o1, slope) + a1[2 * o2 + 1][2 * p + 1];
        }
    // upsample x2 + out_conv (folded) + tanh: skip positions 4h + p
    float sl[4], sr[4];
#pragma unroll
    for (int ic = 0; ic < 4; ++ic) {
        const float v = px(h ? sk[ic][0] : sk[ic][3]);
        sl[ic] = h ? v : 0.f;
        sr[ic] = h ? 0.f : v;
    }
#pragma unroll
    for (int p = 0; p < 4; ++p) {
        f32x2 e = ldc2(W + GI_OUT_B), o = e;
#pragma unroll
        for (int ic = 0; ic < 4; ++ic) {
            const float* F = W + GI2_OUT + ic * 8;
            const float prev = p > 0 ? sk[ic][p - 1] : sl[ic], next = p < 3 ? sk[ic][p + 1] : sr[ic];
            e = fma2(pk2(prev, prev), ldc2(F + 0), e);
            e = fma2(pk2(sk[ic][p], sk[ic][p]), ldc2(F + 2), e);
            o = fma2(pk2(sk[ic][p], sk[ic][p]), ldc2(F + 4), o);
            o = fma2(pk2(next, next), ldc2(F + 6), o);
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

__device__ __forceinline__ float pair_sqerr(const float (&er)[8], const float (&ei)[8], const float (&cr)[8], const float (&ci)[8]) {
    float se = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const float a = er[i] - cr[i], b = ei[i] - ci[i];
        se = fmaf(a, a, fmaf(b, b, se));
    }
    return se + px(se);
}

__global__ void __launch_bounds__(PT, 1) k_sim_pair(const __grid_constant__ SimArgs a) {
    constexpr int NMETH = 2;
    extern __shared__ double table[];
    const int lane = threadIdx.x & 31, h = threadIdx.x & 1, slot = threadIdx.x >> 1;
    for (int i = threadIdx.x; i < a.n_snr * NM * NC; i += blockDim.x) table[i] = 0.0;
    __syncthreads();
    Acc<NMETH> acc;
    acc_reset(acc, -1);

    const int64_t ntiles = (a.B + PFR - 1) / PFR;
    const int64_t per_cta = (ntiles + gridDim.x - 1) / gridDim.x;
    const int64_t t0 = (int64_t)blockIdx.x * per_cta;
    const int64_t t1 = t0 + per_cta < ntiles ? t0 + per_cta : ntiles;

    const bool grid_mode = a.cfg.snr_mode == OFDMGAN_SNR_GRID;
    const uint64_t fps = grid_mode ? (uint64_t)a.cfg.frames_per_snr : 1;
    const bool incremental = grid_mode && fps >= (uint64_t)PFR;
    int bin = 0;
    uint64_t in_bin = 0;
    if (grid_mode && t0 < t1) {
        const uint64_t f0 = a.frame0 + (uint64_t)(t0 * PFR + slot);
        const uint64_t q = f0 / fps;
        in_bin = f0 - q * fps;
        bin = (int)(q % (uint64_t)a.cfg.n_snr);
    }

    for (int64_t t = t0; t < t1; ++t) {
        const int64_t b = t * PFR + slot;
        const bool live = b < a.B;
        const int64_t bb = live ? b : a.B - 1;
        const uint64_t frame = a.frame0 + (uint64_t)bb;
        int fbin = bin;
        if (grid_mode && !incremental) fbin = snr_bin_of(a.cfg, frame);
        else if (grid_mode && !live) fbin = acc.bin >= 0 ? acc.bin : bin;
        if (incremental) {
            in_bin += PFR;
            if (in_bin >= fps) { in_bin -= fps; bin = bin + 1 == a.cfg.n_snr ? 0 : bin + 1; }
        }
        float snr_db;
        if (grid_mode) {
            snr_db = fmaf(a.cfg.snr_step, (float)(live || !incremental ? fbin : snr_bin_of(a.cfg, frame)), a.cfg.snr_lo);
        } else {
            uint32_t x12[4];
            philox4x32_10(a.keys, (uint32_t)frame, (uint32_t)(frame >> 32), 12u, 0u, x12);
            snr_db = fmaf(a.cfg.snr_hi - a.cfg.snr_lo, u_half(x12[0]), a.cfg.snr_lo);
        }
        float cr[8], ci[8], nr[8], ni[8];
        pair_tx(a, frame, h, cr, ci);
        pair_impair(a, frame, h, snr_db, cr, ci, nr, ni);
        pair_normalise(a.cfg.normalize, cr, ci, nr, ni);
        if (OG_PAIR_BARRIER) __syncthreads();                    // keeps the warps of a scheduler on the same instruction-cache lines

        if (__any_sync(0xffffffffu, fbin != acc.bin || acc.count >= FLUSH_EVERY)) {
            acc_flush<false, NMETH>(acc, table, lane);
            acc.bin = fbin;
        }
        float sr = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) sr = fmaf(cr[i], cr[i], fmaf(ci[i], ci[i], sr));
        const float inv_energy = fast_rcp(sr + px(sr));
        {
            const float se = pair_sqerr(nr, ni, cr, ci);
            float mse, evm, ratio;
            err_to_metrics(se, inv_energy, mse, evm, ratio);
            if (live && h == 0) {
                acc_add<false, NMETH>(acc, OFDMGAN_METHOD_NOEQ, mse, evm, ratio, 0);
                acc.count++;
            }
        }
        float xin[2][8], yo[2][8];
#pragma unroll
        for (int i = 0; i < 8; ++i) { xin[0][i] = nr[i]; xin[1][i] = ni[i]; }
        pair_gen_fwd(c_g, a.slope, h, xin, yo);
        {
            const float se = pair_sqerr(yo[0], yo[1], cr, ci);
            float mse, evm, ratio;
            err_to_metrics(se, inv_energy, mse, evm, ratio);
            if (live && h == 0) acc_add<false, NMETH>(acc, OFDMGAN_METHOD_GAN, mse, evm, ratio, 0);
        }
    }
    acc_flush<false, NMETH>(acc, table, lane);
    __syncthreads();
    double* out = a.partials + (size_t)blockIdx.x * a.n_snr * NM * NC;
    for (int i = threadIdx.x; i < a.n_snr * NM * NC; i += blockDim.x) out[i] = table[i];
}

// OFDMGAN_SIM_PAIR=0 in the environment keeps the one-thread-per-frame kernel (A/B measurements)
static bool pair_disabled() {
    static const bool off = [] { const char* e = getenv("OFDMGAN_SIM_PAIR"); return e && e[0] == '0'; }();
    return off;
}

// what k_sim_pair is built for: metrics only, every draw from Philox, AWGN after the three standard impairments
static bool pair_eligible(const SimCall& c) {
    return !pair_disabled() && c.gen_kind == OFDMGAN_GEN_F32 && c.metrics && !c.rand && !c.clean && !c.noisy && !c.snr &&
           c.cfg->equalizers == 0 && c.cfg->channel_type == OFDMGAN_CHAN_AWGN && c.cfg->snr_mode != OFDMGAN_SNR_NONE &&
           (c.cfg->impair & ~(OFDMGAN_IMPAIR_PA | OFDMGAN_IMPAIR_IQ | OFDMGAN_IMPAIR_PN)) == 0;
}

static int sim_launch_pair(const SimCall& c) {
    cudaStream_t s = c.stream;
    int rc;
    CallGuard guard(s);
    if ((rc = guard.rc)) return rc;
    if ((rc = upload_g(c.gparams258, 0, s))) return rc;
    const int grid = grid_for(c.B, PFR, 1);
    const int n = c.n_snr * NM * NC;
    void* partials = nullptr;
    if ((rc = scratch_for_slot(0, (size_t)grid * n * sizeof(double), 4, &partials))) return rc;
    SimArgs a{};
    a.cfg = *c.cfg;
    a.keys = philox_keys(c.seed);
    a.frame0 = c.frame0;
    a.B = c.B;
    a.slope = c.slope;
    a.partials = (double*)partials;
    a.n_snr = c.n_snr;
    k_sim_pair<<<grid, PT, PAIR_SMEM, s>>>(a);
    OG_CHECK(cudaGetLastError());
    k_reduce_partials<<<(n + 63) / 64, 64, 0, s>>>((const double*)partials, grid, n, c.metrics);
    OG_CHECK(cudaGetLastError());
    return 0;
}

}  // namespace og
