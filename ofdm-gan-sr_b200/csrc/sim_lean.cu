// The headline path as its own kernel: Gaussian-symbol OFDM frames -> non-linear chain -> AWGN -> normalisation
// [-> fp32 MiniGenerator -> per-SNR MSE / EVM rows], one frame per thread, everything in registers.
//   utils/dataset.py:236-293 (SyntheticOFDMDataset.__getitem__), utils/ofdm_utils.py:394-421 (Rapp), :458-488 (IQ),
//   :491-521 (phase noise), :675-708 (AWGN), benchmark_comparison.py:129-146,179-250, models/generator.py:180-208
//
// What distinguishes it from the general kernel (sim_kernel.cuh k_sim, which keeps every option of the API):
//  * The frame is carried as z = x / A_sat (the PA's normalised input).  The Rapp gain needs |x / A|^2, and every later use of the
//    clean frame (normalisation, both error sums) is linear in it, so the factor A and the normalisation scale are folded into the
//    FMA that consumes z; the normalised clean and received frames are only materialised when the caller asks for them.  The
//    noise variance goes under Box-Muller's square root, the received frame's scale onto the first convolution's accumulators, and
//    tanh's 2 log2(e) into the output convolution's taps (weights.cuh GI2_OUT_T).
//  * Randomness: three Box-Muller pairs per Philox block (common.cuh), rounds 1-2 of Philox partly hoisted per frame, every
//    32 x 32 -> 64 product one IMAD.WIDE.  Philox's wide multiplies and the generator's packed FMAs share the FMA-heavy pipe, which
//    is the unit this kernel is bound by (profiles/r2_notes.md).
//  * No CTA barrier in the loop and no tile structure: every warp owns a CONTIGUOUS range of 32-frame groups (its successive
//    frames are 32 apart, so the SNR bin changes rarely and the fp32 running sums are folded into the CTA's double table only
//    every FLUSH_EVERY frames).
//  * Injected draws (parity runs against the reference's own np.random draws) are a separate instantiation: the production one
//    carries no loads - a predicated-off LDG still costs its issue slot.
// A warp-specialised variant (RNG producer warps feeding consumer warps through shared memory, setmaxnreg) was built and
// measured slower; it is kept under tools/experiments/ with its numbers in profiles/r2_notes.md.
#include <type_traits>

#include "chan_device.cuh"
#include "gen_device.cuh"
#include "io_tile.cuh"
#include "sim_metrics.cuh"

#ifndef OG_LEAN_WARPS
#define OG_LEAN_WARPS 16     // warps per CTA, one CTA per SM (128 registers per thread)
#endif

namespace og {

constexpr int LN_W = OG_LEAN_WARPS;
constexpr int LN_THREADS = 32 * LN_W;
constexpr int LN_TBL_NM = 2;                                          // methods in the CTA table: GAN, NoEQ
constexpr size_t LN_SMEM = (size_t)LN_W * 32 * 8 * sizeof(float4) + (size_t)OFDMGAN_MAX_SNR_BINS * LN_TBL_NM * NC * sizeof(double);

// ---- Philox4x32-10 with the frame-invariant parts of rounds 1 and 2 hoisted ------------------------------------------
// counter = (frame lo, frame hi, block, 0): round 1 multiplies frame lo (the same for all blocks of a frame) and the block
// index; round 2's second multiply sees only frame-level values.  Per block: 2 + 8 x 2 wide multiplies instead of 20.
struct PhiloxFrame {
    uint32_t a;          // frame hi ^ k0[0]
    uint32_t lo1p;       // low word of M1 * n2
    uint32_t bc;         // high word of M1 * n2, ^ k0[1]
    uint32_t cc;         // low word of M0 * frame lo, ^ k1[1]
};
__device__ __forceinline__ PhiloxFrame philox_frame(const PhiloxKeys& k, uint64_t frame) {
    const uint32_t flo = (uint32_t)frame, fhi = (uint32_t)(frame >> 32);
    uint32_t hi0, lo0, hi1, lo1;
    mulwide(0xD2511F53u, flo, hi0, lo0);
    const uint32_t n2 = hi0 ^ k.k1[0];
    mulwide(0xCD9E8D57u, n2, hi1, lo1);
    PhiloxFrame f;
    f.a = fhi ^ k.k0[0];
    f.lo1p = lo1;
    f.bc = hi1 ^ k.k0[1];
    f.cc = lo0 ^ k.k1[1];
    return f;
}
template <int R>
__device__ __forceinline__ void philox_block(const PhiloxKeys& k, const PhiloxFrame& f, uint32_t blk, uint32_t (&out)[4]) {
    uint32_t hi1, lo1, c2, c3;
    mulwide(0xCD9E8D57u, blk, hi1, lo1);
    const uint32_t n0 = hi1 ^ f.a;                                    // round 1
    uint32_t c0 = f.bc ^ lo1, c1 = f.lo1p;                            // round 2
    mulwide(0xD2511F53u, n0, c2, c3);
    c2 ^= f.cc;
#pragma unroll
    for (int r = 2; r < R; ++r) {
        uint32_t h0, l0, h1, l1;
        mulwide(0xD2511F53u, c0, h0, l0);
        mulwide(0xCD9E8D57u, c2, h1, l1);
        c0 = h1 ^ c1 ^ k.k0[r];
        c1 = l1;
        c2 = h0 ^ c3 ^ k.k1[r];
        c3 = l0;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
// NP pairs of the section that starts at block blk0, in polar pieces: normal 2p = r[p] c[p], normal 2p+1 = r[p] s[p]
template <int NP, int R>
__device__ __forceinline__ void section_polar(const PhiloxKeys& keys, const PhiloxFrame& f, uint32_t blk0, float k, float (&r)[NP],
                                              float (&c)[NP], float (&s)[NP]) {
#pragma unroll
    for (int b = 0; b < (NP + 2) / 3; ++b) {
        uint32_t x[4];
        philox_block<R>(keys, f, blk0 + b, x);
#pragma unroll
        for (int q = 0; q < 3; ++q)
            if (3 * b + q < NP) bm_polar(x, q, k, r[3 * b + q], c[3 * b + q], s[3 * b + q]);
    }
}

// ---- 16-point inverse FFT (unscaled), radix-2 decimation in time on bit-reversed input -------------------------------
// General butterfly in 6 instructions: a' = a + W b as two chained FMAs per component, b' = 2a - a' as one.
template <int TW>
__device__ __forceinline__ void bfly(float& ar, float& ai, float& br, float& bi) {
    if (TW == 0) {
        const float tr = ar - br, ti = ai - bi;
        ar += br; ai += bi; br = tr; bi = ti;
    } else if (TW == 4) {                                             // W = +j: W b = (-bi, br)
        const float tr = ar + bi, ti = ai - br;
        ar -= bi; ai += br; br = tr; bi = ti;
    } else {
        const float wr = Tw<16>::c(TW), wi = Tw<16>::s(TW);
        const float xr = fmaf(br, wr, fmaf(-bi, wi, ar)), xi = fmaf(br, wi, fmaf(bi, wr, ai));
        br = fmaf(2.0f, ar, -xr); bi = fmaf(2.0f, ai, -xi);
        ar = xr; ai = xi;
    }
}
template <int ST_, int Q>
__device__ __forceinline__ void ifft16_bf(float (&tr)[16], float (&ti)[16]) {
    constexpr int h = 1 << (ST_ - 1), j = Q & (h - 1), a = ((Q >> (ST_ - 1)) << ST_) + j, b = a + h, tw = j * (8 >> (ST_ - 1));
    bfly<tw>(tr[a], ti[a], tr[b], ti[b]);
}
template <int ST_, int... Q>
__device__ __forceinline__ void ifft16_stage(float (&tr)[16], float (&ti)[16]) { (ifft16_bf<ST_, Q>(tr, ti), ...); }
__device__ __forceinline__ void ifft16(float (&re)[16], float (&im)[16]) {
    float tr[16], ti[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) { tr[i] = re[bitrev(i, 4)]; ti[i] = im[bitrev(i, 4)]; }
    ifft16_stage<1, 0, 1, 2, 3, 4, 5, 6, 7>(tr, ti);
    ifft16_stage<2, 0, 1, 2, 3, 4, 5, 6, 7>(tr, ti);
    ifft16_stage<3, 0, 1, 2, 3, 4, 5, 6, 7>(tr, ti);
    ifft16_stage<4, 0, 1, 2, 3, 4, 5, 6, 7>(tr, ti);
#pragma unroll
    for (int i = 0; i < 16; ++i) { re[i] = tr[i]; im[i] = ti[i]; }
}

// The same transform with the symbols still in Box-Muller's polar pieces (Re X[k] = normal k, Im X[k] = normal 16 + k, normal
// 2p | 2p+1 = r[p] c[p] | r[p] s[p]): the first stage a +- b forms one product and two FMAs instead of two products and two adds.
__device__ __forceinline__ void ifft16_polar(const float (&r)[16], const float (&c)[16], const float (&s)[16], float (&re)[16],
                                             float (&im)[16]) {
    float tr[16], ti[16];
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        const int k = bitrev(2 * q, 4), pa = k >> 1, pb = pa + 4;        // X[k] and X[k + 8]
        const float ta = r[pb] * ((k & 1) ? s[pb] : c[pb]);
        tr[2 * q] = fmaf(r[pa], (k & 1) ? s[pa] : c[pa], ta);
        tr[2 * q + 1] = fmaf(r[pa], (k & 1) ? s[pa] : c[pa], -ta);
        const float tb = r[8 + pb] * ((k & 1) ? s[8 + pb] : c[8 + pb]);
        ti[2 * q] = fmaf(r[8 + pa], (k & 1) ? s[8 + pa] : c[8 + pa], tb);
        ti[2 * q + 1] = fmaf(r[8 + pa], (k & 1) ? s[8 + pa] : c[8 + pa], -tb);
    }
    ifft16_stage<2, 0, 1, 2, 3, 4, 5, 6, 7>(tr, ti);
    ifft16_stage<3, 0, 1, 2, 3, 4, 5, 6, 7>(tr, ti);
    ifft16_stage<4, 0, 1, 2, 3, 4, 5, 6, 7>(tr, ti);
#pragma unroll
    for (int i = 0; i < 16; ++i) { re[i] = tr[i]; im[i] = ti[i]; }
}

// ---- generator forward, inference form ---------------------------------------------------------------------------------
// As gen_fwd_f32_infer (gen_device.cuh) with three more folds: the input scale (the normalisation of the received frame) is
// applied to enc1's accumulators, LeakyReLU's multiply is packed over the channel pair, and the output convolution uses taps
// pre-multiplied by 2 log2 e so that tanh(v) = 1 - 2 / (1 + 2^v') needs no multiply.  out(p, e, o) receives the two channels'
// scaled pre-activations at positions 2p (e) and 2p+1 (o).
__device__ __forceinline__ void lrelu2(f32x2 v, f32x2 slope2, float& lo, float& hi) {
    float a, b, c, d;
    upk2(v, a, b);
    upk2(mul2(v, slope2), c, d);
    lo = fmaxf(a, c);
    hi = fmaxf(b, d);
}
template <class F>
__device__ __forceinline__ void gen_fwd_f32_scaled(const float* __restrict__ W, float slope, float s_in, const float (&x0)[16],
                                                   const float (&x1)[16], F&& out) {
    const f32x2 sl2 = pk2(slope, slope), s2 = pk2(s_in, s_in);
    float a1[4][8], a2[8][4], sk[4][8];
#pragma unroll
    for (int o2 = 0; o2 < 2; ++o2)
#pragma unroll
        for (int p = 0; p < 8; ++p) {
            f32x2 acc = 0;
            bool first = true;
#pragma unroll
            for (int ic = 0; ic < 2; ++ic)
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    const int i = 2 * p + k - 1;
                    if (i < 0) continue;
                    const float x = ic ? x1[i] : x0[i];
                    const f32x2 w = ldc2(W + GI2_ENC + ((o2 * 2 + ic) * 3 + k) * 2);
                    acc = first ? mul2(pk2(x, x), w) : fma2(pk2(x, x), w, acc);
                    first = false;
                }
            lrelu2(fma2(acc, s2, ldc2(W + GI_ENC_B + 2 * o2)), sl2, a1[2 * o2][p], a1[2 * o2 + 1][p]);
        }
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
            lrelu2(acc, sl2, a2[2 * o2][p], a2[2 * o2 + 1][p]);
        }
#pragma unroll
    for (int o2 = 0; o2 < 2; ++o2)
#pragma unroll
        for (int p = 0; p < 4; ++p) {
            f32x2 e = ldc2(W + GI_DEC_B + 2 * o2), o = e;
#pragma unroll
            for (int ic = 0; ic < 8; ++ic) {
                const float* F4 = W + GI2_DEC + (o2 * 8 + ic) * 8;
                if (p > 0) e = fma2(pk2(a2[ic][p - 1], a2[ic][p - 1]), ldc2(F4 + 0), e);
                e = fma2(pk2(a2[ic][p], a2[ic][p]), ldc2(F4 + 2), e);
                o = fma2(pk2(a2[ic][p], a2[ic][p]), ldc2(F4 + 4), o);
                if (p < 3) o = fma2(pk2(a2[ic][p + 1], a2[ic][p + 1]), ldc2(F4 + 6), o);
            }
            float e0, e1, o0, o1;
            lrelu2(e, sl2, e0, e1);
            lrelu2(o, sl2, o0, o1);
            upk2(add2(pk2(e0, e1), pk2(a1[2 * o2][2 * p], a1[2 * o2 + 1][2 * p])), sk[2 * o2][2 * p], sk[2 * o2 + 1][2 * p]);
            upk2(add2(pk2(o0, o1), pk2(a1[2 * o2][2 * p + 1], a1[2 * o2 + 1][2 * p + 1])), sk[2 * o2][2 * p + 1], sk[2 * o2 + 1][2 * p + 1]);
        }
#pragma unroll
    for (int p = 0; p < 8; ++p) {
        f32x2 e = ldc2(W + GI_OUT_BT), o = e;
#pragma unroll
        for (int ic = 0; ic < 4; ++ic) {
            const float* F4 = W + GI2_OUT_T + ic * 8;
            if (p > 0) e = fma2(pk2(sk[ic][p - 1], sk[ic][p - 1]), ldc2(F4 + 0), e);
            e = fma2(pk2(sk[ic][p], sk[ic][p]), ldc2(F4 + 2), e);
            o = fma2(pk2(sk[ic][p], sk[ic][p]), ldc2(F4 + 4), o);
            if (p < 7) o = fma2(pk2(sk[ic][p + 1], sk[ic][p + 1]), ldc2(F4 + 6), o);
        }
        out(p, e, o);
    }
}

// ---- work split ---------------------------------------------------------------------------------------------------------
// warp `cw` of CTA `cta` owns the 32-frame groups [g0, g1)
struct LnChunk {
    int64_t g0, g1;
};
__device__ __forceinline__ LnChunk ln_chunk(int64_t B, int cta, int ncta, int cw) {
    const int64_t ng = (B + 31) >> 5, nw = (int64_t)ncta * LN_W, per = (ng + nw - 1) / nw;
    LnChunk c;
    c.g0 = ((int64_t)cta * LN_W + cw) * per;
    c.g1 = c.g0 + per < ng ? c.g0 + per : ng;
    if (c.g0 > ng) c.g0 = ng;
    return c;
}
// per-launch scalars every role derives from the configuration (uniform registers)
struct LnScal {
    bool pa_on, p3, pn_on, pn_fast, awgn, grid_mode;
    float A, invA, log2A, p, ninv2p, gc, gs, k_sym, sc_in, k_pn, pn_sigma;
};
__device__ __forceinline__ LnScal ln_scalars(const ofdmgan_chan_cfg& c) {
    LnScal s;
    s.pa_on = (c.impair & OFDMGAN_IMPAIR_PA) != 0;
    s.A = s.pa_on ? c.pa_saturation : 1.0f;
    s.invA = 1.0f / s.A;
    s.log2A = log2f(s.A);
    s.p = c.pa_smoothness;
    s.p3 = c.pa_smoothness == 3.0f;
    s.ninv2p = -0.5f / c.pa_smoothness;
    const bool iq = (c.impair & OFDMGAN_IMPAIR_IQ) != 0;
    s.gc = iq ? c.iq_gain * c.iq_cos : 1.0f;
    s.gs = iq ? c.iq_gain * c.iq_sin : 0.0f;
    // (1/sqrt2 per bin) * (ifft 1/N) * (sqrt(N) or N), and 1/A: the IFFT is linear, so the scale goes onto the symbols - inside
    // Box-Muller's square root for Philox symbols (free), one multiply each for injected ones
    s.sc_in = 0.70710678118654752f * (c.ifft_scale == OFDMGAN_SCALE_N ? 1.0f : 0.25f) * s.invA;
    s.k_sym = OG_BM_K * s.sc_in * s.sc_in;
    s.pn_on = (c.impair & OFDMGAN_IMPAIR_PN) != 0;
    s.pn_sigma = c.pn_sigma;
    s.k_pn = OG_BM_K * c.pn_sigma * c.pn_sigma;
    // the accumulated phase of 16 steps stays far inside MUFU.SIN's accurate range (|theta| < 2 pi is a 7.8 sigma event at
    // sigma = 0.2): skip the explicit reduction there
    s.pn_fast = c.pn_sigma <= 0.2f;
    s.awgn = c.snr_mode != OFDMGAN_SNR_NONE;
    s.grid_mode = c.snr_mode == OFDMGAN_SNR_GRID;
    return s;
}

// ---- one frame per thread -------------------------------------------------------------------------------------------------
// INJ: the caller may inject host-generated draws (parity runs).  OUT: frames go to HBM (the dataset path).
// CHAIN: which stages are compiled in.  0: whatever the configuration says, as uniform branches; 1: the linear chain (AWGN only:
// SyntheticOFDMDataset's default); 2: the reference's non-linear chain (Rapp with p = 3, IQ imbalance, phase noise with sigma <=
// 0.2, AWGN: --nonlinear).  The specialised instantiations are straight-line code: no branch joins, no register shuffling.
enum { CHAIN_ANY = 0, CHAIN_LINEAR = 1, CHAIN_NONLINEAR = 2 };
// BYVAL: the generator's weight image arrives in the parameter block (host-resident inference weights) instead of the __constant__ one
struct LeanArgs {
    SimArgs a;
    GImage g;
};
// R: Philox rounds (10; 7 = the separately named fast-RNG workload, ofdmgan_chan_cfg.rng_rounds)
template <int GEN, bool INJ, bool OUT, int CHAIN, bool BYVAL, int R>
__global__ void __launch_bounds__(LN_THREADS, 1) k_sim_lean(const __grid_constant__ LeanArgs la) {
    const SimArgs& a = la.a;
    extern __shared__ float4 sm[];
    double* table = reinterpret_cast<double*>(sm + LN_W * 32 * 8);    // [n_snr][LN_TBL_NM][NC]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float4* park = sm + warp * 32 * 8;                                 // the warp's tile: z parked per lane / staging of frame stores
    const unsigned full_mask = 0xffffffffu;
    const bool want_metrics = GEN >= 0 && a.partials != nullptr;
    if (want_metrics) {
        for (int i = threadIdx.x; i < a.n_snr * LN_TBL_NM * NC; i += blockDim.x) table[i] = 0.0;
        __syncthreads();
    }
    const LnScal sc = ln_scalars(a.cfg);
    const bool pa_on = CHAIN == CHAIN_ANY ? sc.pa_on : CHAIN == CHAIN_NONLINEAR, p3 = CHAIN == CHAIN_ANY ? sc.p3 : true;
    const bool pn_on = CHAIN == CHAIN_ANY ? sc.pn_on : CHAIN == CHAIN_NONLINEAR, pn_fast = CHAIN == CHAIN_ANY ? sc.pn_fast : true;
    const bool awgn = CHAIN == CHAIN_ANY ? sc.awgn : true;
    const LnChunk ch = ln_chunk(a.B, blockIdx.x, gridDim.x, warp);
    const float* const inj_sym = INJ ? a.sym : nullptr;
    const float* const inj_pn = INJ ? a.pn : nullptr;
    const float* const inj_snr = INJ ? a.snr_db : nullptr;
    const float* const inj_noise = INJ ? a.noise : nullptr;
    Acc<2> acc;
    acc_reset(acc, -1);

    // SNR grid position of this lane's first frame; afterwards advanced by 32 frames per group without divisions
    const uint64_t fps = sc.grid_mode ? (uint64_t)a.cfg.frames_per_snr : 1;
    int bin = 0;
    uint64_t in_bin = 0;
    if (sc.grid_mode && ch.g0 < ch.g1) {
        const uint64_t f0 = a.frame0 + (uint64_t)(ch.g0 * 32 + lane);
        const uint64_t q = f0 / fps;
        in_bin = f0 - q * fps;
        bin = (int)(q % (uint64_t)a.cfg.n_snr);
    }

    for (int64_t g = ch.g0; g < ch.g1; ++g) {
        const int64_t gbase = g * 32, b = gbase + lane;
        const bool live = b < a.B;
        const int64_t bb = live ? b : a.B - 1;                       // dead lanes recompute the last frame, results dropped
        const uint64_t frame = a.frame0 + (uint64_t)bb;
        const int fbin = bin;
        if (sc.grid_mode) {                                          // advance to the next group's frame
            in_bin += 32;
            while (in_bin >= fps) { in_bin -= fps; bin = bin + 1 == a.cfg.n_snr ? 0 : bin + 1; }
        }
        const PhiloxFrame pf = philox_frame(a.keys, frame);

        float snr_db;
        if (sc.grid_mode) {
            snr_db = fmaf(a.cfg.snr_step, (float)fbin, a.cfg.snr_lo);
        } else if (!awgn) {
            snr_db = __int_as_float(0x7f800000);                     // +inf: reported as "no noise"
        } else if (inj_snr) {
            snr_db = inj_snr[bb];
        } else {
            uint32_t x12[4];
            philox4x32<R>(a.keys, (uint32_t)frame, (uint32_t)(frame >> 32), 12u, 0u, x12);
            snr_db = fmaf(a.cfg.snr_hi - a.cfg.snr_lo, u_half(x12[0]), a.cfg.snr_lo);
        }

        // ---- symbols -> z = x / A (time domain)
        float zr[16], zi[16];
        if (inj_sym) {
#pragma unroll
            for (int k = 0; k < 16; ++k) { zr[k] = inj_sym[bb * 32 + k] * sc.sc_in; zi[k] = inj_sym[bb * 32 + 16 + k] * sc.sc_in; }
        } else {
            float r[16], c[16], s[16];
            section_polar<16, R>(a.keys, pf, 0u, sc.k_sym, r, c, s);
            ifft16_polar(r, c, s, zr, zi);
        }
        if (inj_sym) ifft16(zr, zi);
        {
            float f[2][16];
#pragma unroll
            for (int i = 0; i < 16; ++i) { f[0][i] = zr[i]; f[1][i] = zi[i]; }
            tile_write_f32(park, lane, f);
        }
        float mz = 0.f;
#pragma unroll
        for (int i = 0; i < 16; ++i) mz = fmaxf(mz, fmaxf(fabsf(zr[i]), fabsf(zi[i])));

        // ---- Rapp PA (on z: gain = A (1 + |z|^2p)^(-1/2p)) and IQ imbalance; Ez = sum |z|^2
        float nr[16], ni[16], Ez = 0.f;
        if (pa_on) {
            if (p3) {
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const float t = fmaf(zi[i], zi[i], zr[i] * zr[i]);
                    Ez += t;
                    const float G = fast_ex2(fmaf(sc.ninv2p, fast_lg2(fmaf(t * t, t, 1.0f)), sc.log2A));
                    const float w = fmaf(sc.gs, zr[i], sc.gc * zi[i]);
                    nr[i] = zr[i] * G;
                    ni[i] = w * G;
                }
            } else {
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const float t = fmaf(zi[i], zi[i], zr[i] * zr[i]);
                    Ez += t;
                    const float G = fast_ex2(fmaf(sc.ninv2p, fast_lg2(1.0f + fast_ex2(sc.p * fast_lg2(t))), sc.log2A));
                    const float w = fmaf(sc.gs, zr[i], sc.gc * zi[i]);
                    nr[i] = zr[i] * G;
                    ni[i] = w * G;
                }
            }
        } else {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                Ez = fmaf(zi[i], zi[i], fmaf(zr[i], zr[i], Ez));
                nr[i] = zr[i];
                ni[i] = CHAIN == CHAIN_LINEAR ? zi[i] : fmaf(sc.gs, zr[i], sc.gc * zi[i]);
            }
        }

        // ---- Wiener phase noise: theta_i = theta_{i-1} + sigma n_i, x_i *= e^{j theta_i}
        if (pn_on) {
            auto steps = [&](auto fast) {
                float th = 0.f, r[8], c[8], s[8];
                if (!inj_pn) section_polar<8, R>(a.keys, pf, 8u, sc.k_pn, r, c, s);
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    if (inj_pn) th = fmaf(inj_pn[bb * 16 + i], sc.pn_sigma, th);
                    else th = fmaf(r[i >> 1], (i & 1) ? s[i >> 1] : c[i >> 1], th);
                    const float red = decltype(fast)::value ? th : fmaf(-6.283185307179586f, rintf(th * 0.15915494309189535f), th);
                    const float sn = fast_sin(red), co = fast_cos(red);
                    const float xr = nr[i], xi = ni[i];
                    nr[i] = fmaf(xr, co, -xi * sn);
                    ni[i] = fmaf(xr, sn, xi * co);
                }
            };
            if (pn_fast) steps(std::true_type{}); else steps(std::false_type{});
        }

        // ---- AWGN at the measured power: sigma^2 = P / 10^(snr/10) / 2, P = mean |x|^2; the variance goes under Box-Muller's root
        if (awgn) {
            float P = 0.f;
#pragma unroll
            for (int i = 0; i < 16; ++i) P = fmaf(nr[i], nr[i], fmaf(ni[i], ni[i], P));
            const float nv = 0.03125f * P * fast_ex2(-0.33219280948873623f * snr_db);
            if (inj_noise) {
                const float sd = fast_sqrt(nv);
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    nr[i] = fmaf(sd, inj_noise[bb * 32 + i], nr[i]);
                    ni[i] = fmaf(sd, inj_noise[bb * 32 + 16 + i], ni[i]);
                }
            } else {
                float r[16], c[16], s[16];
                section_polar<16, R>(a.keys, pf, 13u, OG_BM_K * nv, r, c, s);
#pragma unroll
                for (int p = 0; p < 8; ++p) {
                    nr[2 * p] = fmaf(r[p], c[p], nr[2 * p]); nr[2 * p + 1] = fmaf(r[p], s[p], nr[2 * p + 1]);
                    ni[2 * p] = fmaf(r[8 + p], c[8 + p], ni[2 * p]); ni[2 * p + 1] = fmaf(r[8 + p], s[8 + p], ni[2 * p + 1]);
                }
            }
        }

        // ---- normalisation factors (utils/dataset.py:284-287 joint; benchmark_comparison.py:129-134 separate).
        // normalised clean = kappa z, normalised received = s_n n; rho = kappa / s_n
        float mn = 0.f;
#pragma unroll
        for (int i = 0; i < 16; ++i) mn = fmaxf(mn, fmaxf(fabsf(nr[i]), fabsf(ni[i])));
        const float mc = mz * sc.A;
        float s_n = 1.0f, kappa = sc.A, rho = sc.A;
        if (a.cfg.normalize == OFDMGAN_NORM_JOINT) {
            const float m = fmaxf(mc, mn);
            s_n = m > 0.f ? __frcp_rn(m) : 1.0f;
            kappa = s_n * sc.A;
        } else if (a.cfg.normalize == OFDMGAN_NORM_SEPARATE) {
            const float s_c = mc > 0.f ? __frcp_rn(mc) : 1.0f;
            s_n = mn > 0.f ? __frcp_rn(mn) : 1.0f;
            kappa = s_c * sc.A;
            rho = kappa * (mn > 0.f ? mn : 1.0f);
        }

        // ---- frames to HBM when asked for (the dataset path): materialise, stage through the warp's tile, store coalesced
        if (OUT && (a.clean || a.noisy)) {
            float zz[2][16];
            tile_read_f32(park, lane, zz);
            __syncwarp();
            if (a.clean) {
                float f[2][16];
#pragma unroll
                for (int i = 0; i < 16; ++i) { f[0][i] = zz[0][i] * kappa; f[1][i] = zz[1][i] * kappa; }
                tile_store_f32(a.clean, gbase, a.B, park, lane, f);
            }
            if (a.noisy) {
                float f[2][16];
#pragma unroll
                for (int i = 0; i < 16; ++i) { f[0][i] = nr[i] * s_n; f[1][i] = ni[i] * s_n; }
                tile_store_f32(a.noisy, gbase, a.B, park, lane, f);
            }
            if (GEN >= 0) tile_write_f32(park, lane, zz);
        }
        if (OUT && a.snr_out && live) a.snr_out[b] = snr_db;
        if (GEN < 0) continue;

        // ---- metrics without equalisation: |s_n n - kappa z|^2 = s_n^2 |n - rho z|^2
        float inv_energy = 0.f;
        if (want_metrics) {
            if (__any_sync(full_mask, fbin != acc.bin || acc.count >= FLUSH_EVERY)) {
                acc_flush<false, 2, LN_TBL_NM>(acc, table, lane);
                acc.bin = fbin;
            }
            inv_energy = fast_rcp(kappa * kappa * Ez);
            float se = 0.f;
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const float4 v = park[lane * 8 + (c ^ (lane & 7))];
                const float* n4 = c < 4 ? &nr[4 * c] : &ni[4 * (c - 4)];
                const float d0 = fmaf(-rho, v.x, n4[0]), d1 = fmaf(-rho, v.y, n4[1]), d2 = fmaf(-rho, v.z, n4[2]), d3 = fmaf(-rho, v.w, n4[3]);
                se = fmaf(d0, d0, fmaf(d1, d1, fmaf(d2, d2, fmaf(d3, d3, se))));
            }
            if (live) {
                float mse, evm, ratio;
                err_to_metrics(se * s_n * s_n, inv_energy, mse, evm, ratio);
                acc_add<false, 2>(acc, OFDMGAN_METHOD_NOEQ, mse, evm, ratio, 0);
                acc.count++;
            }
        }

        // ---- reconstruct and compare: tanh(v) - kappa z = (1 - kappa z) - 2 / (1 + 2^v')
        float se_a = 0.f, se_b = 0.f;
        float zc[2][4];
        gen_fwd_f32_scaled(BYVAL ? la.g.w : c_g, a.slope, s_n, nr, ni, [&](int p, f32x2 e, f32x2 o) {
            if ((p & 1) == 0) {
                const int c = p >> 1;
                const float4 vr = park[lane * 8 + (c ^ (lane & 7))], vi = park[lane * 8 + ((c + 4) ^ (lane & 7))];
                zc[0][0] = vr.x; zc[0][1] = vr.y; zc[0][2] = vr.z; zc[0][3] = vr.w;
                zc[1][0] = vi.x; zc[1][1] = vi.y; zc[1][2] = vi.z; zc[1][3] = vi.w;
            }
            float e0, e1, o0, o1;
            upk2(e, e0, e1);
            upk2(o, o0, o1);
            const int q = (2 * p) & 3;
            const float de0 = fmaf(-2.0f, fast_rcp(fast_ex2(e0) + 1.0f), fmaf(-kappa, zc[0][q], 1.0f));
            const float de1 = fmaf(-2.0f, fast_rcp(fast_ex2(e1) + 1.0f), fmaf(-kappa, zc[1][q], 1.0f));
            const float do0 = fmaf(-2.0f, fast_rcp(fast_ex2(o0) + 1.0f), fmaf(-kappa, zc[0][q + 1], 1.0f));
            const float do1 = fmaf(-2.0f, fast_rcp(fast_ex2(o1) + 1.0f), fmaf(-kappa, zc[1][q + 1], 1.0f));
            se_a = fmaf(de0, de0, fmaf(do0, do0, se_a));
            se_b = fmaf(de1, de1, fmaf(do1, do1, se_b));
        });
        if (want_metrics && live) {
            float mse, evm, ratio;
            err_to_metrics(se_a + se_b, inv_energy, mse, evm, ratio);
            acc_add<false, 2>(acc, OFDMGAN_METHOD_GAN, mse, evm, ratio, 0);
        }
    }
    if (want_metrics) {
        acc_flush<false, 2, LN_TBL_NM>(acc, table, lane);
        __syncthreads();
        // the CTA's rows in the caller's layout [n_snr][OFDMGAN_N_METHODS][cols]: only the GAN and NoEQ rows are produced here
        double* outp = a.partials + (size_t)blockIdx.x * a.n_snr * NM * NC;
        for (int i = threadIdx.x; i < a.n_snr * NM * NC; i += blockDim.x) {
            const int c = i % NC, m = (i / NC) % NM, bn = i / (NC * NM);
            outp[i] = m < LN_TBL_NM ? table[(bn * LN_TBL_NM + m) * NC + c] : 0.0;
        }
    }
}

template <int GEN, bool INJ, bool OUT, int CHAIN, bool BYVAL, int R>
static int sim_lean_launch_impl(const SimCall& c, LeanArgs& la) {
    cudaStream_t s = c.stream;
    int rc, err = 0;
    const DeviceInfo& di = device_info(&err);
    const int sms = err ? 148 : di.sms;
    const int64_t ng = (c.B + 31) / 32;
    int64_t want = (ng + LN_W - 1) / LN_W;
    if (want < 1) want = 1;
    const int grid = (int)(want < sms ? want : sms);                 // persistent: one CTA per SM
    OG_CHECK(cudaFuncSetAttribute(k_sim_lean<GEN, INJ, OUT, CHAIN, BYVAL, R>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)LN_SMEM));
    const int n = c.n_snr * NM * NC;
    void* partials = nullptr;
    // per-stream scratch for the per-CTA partial tables: by-value calls hold no lock, so streams must not share it
    if (GEN >= 0 && c.metrics && (rc = scratch_for_stream(s, (size_t)grid * n * sizeof(double), 0, &partials))) return rc;
    SimArgs& a = la.a;
    a = SimArgs{};
    a.cfg = *c.cfg;
    a.keys = philox_keys(c.seed);
    a.frame0 = c.frame0;
    a.B = c.B;
    if (c.rand) { a.sym = c.rand->sym; a.pn = c.rand->pn; a.snr_db = c.rand->snr_db; a.noise = c.rand->noise; }
    a.clean = c.clean; a.noisy = c.noisy; a.snr_out = c.snr;
    a.wslot = 0;
    a.slope = c.slope;
    a.partials = (double*)partials;
    a.n_snr = c.n_snr;
    k_sim_lean<GEN, INJ, OUT, CHAIN, BYVAL, R><<<grid, LN_THREADS, LN_SMEM, s>>>(la);
    OG_CHECK(cudaGetLastError());
    if (partials) {
        reduce_partials_launch((const double*)partials, grid, n, c.metrics, s);
        OG_CHECK(cudaGetLastError());
    }
    return 0;
}
template <int GEN, bool INJ, bool OUT, int CHAIN, int R = 10>
static int sim_lean_launch_one(const SimCall& c) {
    static thread_local LeanArgs la;                                 // (5 KB: not on the stack of every caller)
    if constexpr (GEN == OFDMGAN_GEN_F32 && !INJ) {
        if (is_host_pointer(c.gparams258)) {
            g_image_host(c.gparams258, la.g);                        // inference weights by value: no shared state, no lock
            return sim_lean_launch_impl<GEN, INJ, OUT, CHAIN, true, R>(c, la);
        }
    }
    int rc;
    CallGuard guard(c.stream);                                       // device-resident weights: the __constant__ image
    if ((rc = guard.rc)) return rc;
    if (GEN == OFDMGAN_GEN_F32 && (rc = upload_g(c.gparams258, 0, c.stream))) return rc;
    return sim_lean_launch_impl<GEN, INJ, OUT, CHAIN, false, R>(c, la);
}

// Is this call the headline shape?  Gaussian source, no injected time-domain frames / fading draws, no late stages, no
// equaliser rows, fp32 generator or none.  Everything else runs on the general kernel (sim_kernel.cuh).
bool sim_lean_eligible(const SimCall& c) {
    if (c.src != SRC_GAUSS || c.B < 1) return false;
    if (c.gen_kind != -1 && c.gen_kind != OFDMGAN_GEN_F32) return false;
    if (c.cfg->equalizers != 0 || c.cfg->channel_type != OFDMGAN_CHAN_AWGN) return false;
    if (c.cfg->impair & (OFDMGAN_IMPAIR_SALEH | OFDMGAN_IMPAIR_DC | OFDMGAN_IMPAIR_CFO)) return false;
    if ((c.cfg->impair & OFDMGAN_IMPAIR_PA) && !(c.cfg->pa_saturation > 0.f)) return false;
    if (c.rand && (c.rand->tx || c.rand->fade)) return false;
    if (c.gen_kind == OFDMGAN_GEN_F32 && (c.clean || c.noisy || c.snr)) return false;   // (not reachable through the C ABI)
    if (c.cfg->rng_rounds == 7 && c.rand && (c.rand->sym || c.rand->pn || c.rand->snr_db || c.rand->noise)) return false;
    return true;
}
// the stage set of a configuration (CHAIN_*)
static int chain_of(const ofdmgan_chan_cfg& c) {
    if (c.snr_mode == OFDMGAN_SNR_NONE) return CHAIN_ANY;
    const int nl = OFDMGAN_IMPAIR_PA | OFDMGAN_IMPAIR_IQ | OFDMGAN_IMPAIR_PN;
    if ((c.impair & nl) == 0) return CHAIN_LINEAR;
    if ((c.impair & nl) == nl && c.pa_smoothness == 3.0f && c.pn_sigma <= 0.2f) return CHAIN_NONLINEAR;
    return CHAIN_ANY;
}
// The C ABI writes frames only from the simulate-only call (ofdmgan_chan_sim) and metrics only from the fused call
// (ofdmgan_sim_gen_metrics), so OUT == (GEN < 0).
template <int GEN>
static int sim_lean_launch_chain(const SimCall& c, bool inj) {
    constexpr bool OUT = GEN < 0;
    if (inj) return sim_lean_launch_one<GEN, true, OUT, CHAIN_ANY>(c);     // parity runs: one instantiation with every branch
    if (c.cfg->rng_rounds == 7) {                                          // the fast-RNG workload: the two reference chains only
        switch (chain_of(*c.cfg)) {
            case CHAIN_LINEAR: return sim_lean_launch_one<GEN, false, OUT, CHAIN_LINEAR, 7>(c);
            case CHAIN_NONLINEAR: return sim_lean_launch_one<GEN, false, OUT, CHAIN_NONLINEAR, 7>(c);
            default: return OFDMGAN_E_UNSUPPORTED;
        }
    }
    switch (chain_of(*c.cfg)) {
        case CHAIN_LINEAR: return sim_lean_launch_one<GEN, false, OUT, CHAIN_LINEAR>(c);
        case CHAIN_NONLINEAR: return sim_lean_launch_one<GEN, false, OUT, CHAIN_NONLINEAR>(c);
        default: return sim_lean_launch_one<GEN, false, OUT, CHAIN_ANY>(c);
    }
}
int sim_launch_lean(const SimCall& c) {
    const bool inj = c.rand && (c.rand->sym || c.rand->pn || c.rand->snr_db || c.rand->noise);
    if (c.gen_kind == -1) return sim_lean_launch_chain<-1>(c, inj);
    return sim_lean_launch_chain<OFDMGAN_GEN_F32>(c, inj);
}

}  // namespace og
