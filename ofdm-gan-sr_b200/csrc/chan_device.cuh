// Per-thread channel simulator: symbols -> IFFT in registers -> (CP) -> Rapp PA -> IQ imbalance -> Wiener phase noise
// -> AWGN -> normalisation.  One frame per thread, every array index is a compile-time constant.
//   utils/dataset.py:236-293 (SyntheticOFDMDataset.__getitem__), utils/ofdm_utils.py:105-109,163-193 (QPSK map),
//   :281-329 (OFDMModulator.modulate), :394-421 (Rapp), :458-488 (IQ), :491-521 (phase noise), :675-708 (AWGN),
//   benchmark_comparison.py:129-134 (separate normalisation)
#pragma once
#include "common.cuh"

namespace og {

// source layouts compiled into the simulator (anything else: OFDMGAN_E_UNSUPPORTED)
enum { SRC_GAUSS = 0, SRC_Q16_CP0 = 1, SRC_Q16_CP2 = 2, SRC_Q8_CP0 = 3, SRC_Q8_CP2 = 4, SRC_COUNT = 5 };

struct SimArgs {
    ofdmgan_chan_cfg cfg;
    PhiloxKeys keys;
    uint64_t frame0;
    int64_t B;
    const float* sym;        // injected draws (device, nullable)
    const uint32_t* bits;
    const float* pn;
    const float* snr_db;
    const float* noise;
    const float* tx;         // injected time-domain frames (device, nullable)
    const float* fade;       // injected fading draws [B][8] (device, nullable)
    const float* tx_gain;    // [B] with tx: the channel sees tx * tx_gain[b], the clean output stays tx (device, nullable)
    float* clean;            // outputs (device, nullable)
    float* noisy;
    float* snr_out;
    int wslot;               // weight slot of the generator image
    float slope;
    double* partials;        // [grid][n_snr][N_METHODS][COLS] per-CTA metric partials (nullable)
    int n_snr;
};

// what the API layer hands to a simulator translation unit (host struct)
struct SimCall {
    const ofdmgan_chan_cfg* cfg;
    const ofdmgan_chan_rand* rand;     // nullable
    uint64_t seed, frame0;
    int64_t B;
    float* clean;                      // device outputs, nullable
    float* noisy;
    float* snr;
    int gen_kind;                      // -1: simulate only, else OFDMGAN_GEN_*
    const float* gparams258;           // host or device (OFDMGAN_GEN_F32)
    const int8_t* wrom;                // host ROMs (Q kinds)
    const int16_t* brom;
    float slope;
    double* metrics;                   // device accumulator, nullable
    int n_snr;
    int src;                           // SRC_*
    cudaStream_t stream;
};
int sim_launch_gauss(const SimCall& c);    // sim_gauss.cu
int sim_launch_qpsk(const SimCall& c);     // sim_qpsk.cu
int sim_launch_gauss_ext(const SimCall& c); // sim_gauss_ext.cu: the less common (generator, equaliser, late-stage) combinations
int sim_launch_lean(const SimCall& c);     // sim_lean.cu: the headline kernel
bool sim_lean_eligible(const SimCall& c);

// ---- radix-2 DIT inverse FFT, fully unrolled, unscaled: x[n] = sum_k X[k] e^{+j 2 pi k n / N} -------------
// twiddles e^{+j 2 pi k / 16}, k = 0..7, as constant-folded selects (a constexpr table indexed after unrolling is
// materialised in local memory by nvcc 12.9 and costs an LDL per use)
template <int N>
struct Tw;
template <>
struct Tw<16> {
    static __device__ __forceinline__ constexpr float c(int k) {
        return k == 0 ? 1.f : k == 1 ? 0.92387953251128674f : k == 2 ? 0.70710678118654752f : k == 3 ? 0.38268343236508977f
             : k == 4 ? 0.f : k == 5 ? -0.38268343236508977f : k == 6 ? -0.70710678118654752f : -0.92387953251128674f;
    }
    static __device__ __forceinline__ constexpr float s(int k) {
        return k == 0 ? 0.f : k == 1 ? 0.38268343236508977f : k == 2 ? 0.70710678118654752f : k == 3 ? 0.92387953251128674f
             : k == 4 ? 1.f : k == 5 ? 0.92387953251128674f : k == 6 ? 0.70710678118654752f : 0.38268343236508977f;
    }
};

__host__ __device__ constexpr int bitrev(int v, int bits) {
    int r = 0;
    for (int i = 0; i < bits; ++i) r |= ((v >> i) & 1) << (bits - 1 - i);
    return r;
}

// SIGN=+1: inverse (e^{+j}), SIGN=-1: forward (e^{-j}).  N in {8,16}; twiddle index scaled to the N=16 table.
template <int N, int SIGN>
__device__ __forceinline__ void fft_inplace(float (&re)[N], float (&im)[N]) {
    constexpr int LOG = N == 16 ? 4 : 3;
    float tr[N], ti[N];
#pragma unroll
    for (int i = 0; i < N; ++i) { tr[i] = re[bitrev(i, LOG)]; ti[i] = im[bitrev(i, LOG)]; }
#pragma unroll
    for (int st = 1; st <= LOG; ++st) {
        const int h = 1 << (st - 1);
#pragma unroll
        for (int q = 0; q < N / 2; ++q) {                        // fixed trip count: every index below folds to a constant
            const int j = q & (h - 1), a = ((q >> (st - 1)) << st) + j, b = a + h;
            const int tw = j * (8 >> (st - 1));                  // index into the 16-point table
            const float wr = Tw<16>::c(tw), wi = (float)SIGN * Tw<16>::s(tw);
            float xr, xi;
            if (tw == 0) { xr = tr[b]; xi = ti[b]; }
            else if (tw == 4) { xr = -(float)SIGN * ti[b]; xi = (float)SIGN * tr[b]; }
            else { xr = tr[b] * wr - ti[b] * wi; xi = tr[b] * wi + ti[b] * wr; }
            tr[b] = tr[a] - xr; ti[b] = ti[a] - xi;
            tr[a] = tr[a] + xr; ti[a] = ti[a] + xi;
        }
    }
#pragma unroll
    for (int i = 0; i < N; ++i) { re[i] = tr[i]; im[i] = ti[i]; }
}

// ---- draws ------------------------------------------------------------------------------------------------
// Philox block layout per frame (purpose 0), three Box-Muller pairs per block (common.cuh): blocks 0-5 the 16 symbol pairs
// (normals 0-15 Re, 16-31 Im), 8-10 the 8 phase-increment pairs, 12 {snr uniform, payload bits, LOS phase}, 13-18 the 16 noise
// pairs (Re then Im), 21 / 22 two fading pairs each.  See oracle/channel.c header.
// the section of NP pairs that starts at block blk0, as 2 NP normals of variance var
template <int NP, int R = 10>
__device__ __forceinline__ void draw_section(const SimArgs& a, uint64_t frame, uint32_t blk0, float (&n)[2 * NP], float var = 1.0f) {
#pragma unroll
    for (int b = 0; b < (NP + 2) / 3; ++b) {
        uint32_t x[4];
        philox4x32<R>(a.keys, (uint32_t)frame, (uint32_t)(frame >> 32), blk0 + b, 0u, x);
        constexpr int last = NP - 3 * ((NP + 2) / 3 - 1);         // pairs in the last block
        const float k = OG_BM_K * var;
#pragma unroll
        for (int q = 0; q < 3; ++q) {
            if (3 * b + q >= NP) break;
            float r, c, s;
            bm_polar(x, q, k, r, c, s);
            n[2 * (3 * b + q)] = r * c;
            n[2 * (3 * b + q) + 1] = r * s;
        }
        (void)last;
    }
}

__device__ __forceinline__ int snr_bin_of(const ofdmgan_chan_cfg& c, uint64_t frame) {
    if (c.snr_mode != OFDMGAN_SNR_GRID) return 0;
    return (int)((frame / (uint64_t)c.frames_per_snr) % (uint64_t)c.n_snr);
}

// QPSK symbol k of an OFDM symbol: pilot or the next two payload bits (MSB -> Re sign, LSB -> Im sign)
__device__ __forceinline__ void qpsk_fill(const ofdmgan_chan_cfg& c, uint32_t bits, int& bitpos, int k, float& xr, float& xi) {
    const bool pilot = c.pilot_spacing > 0 && (k % c.pilot_spacing) == 0;
    if (pilot) { xr = c.pilot_re; xi = c.pilot_im; return; }
    const uint32_t b1 = bitpos < 32 ? (bits >> (31 - bitpos)) & 1u : 0u;
    const uint32_t b0 = bitpos < 31 ? (bits >> (30 - bitpos)) & 1u : 0u;
    bitpos += 2;
    xr = b1 ? -0.70710678118654752f : 0.70710678118654752f;
    xi = b0 ? -0.70710678118654752f : 0.70710678118654752f;
}

// clean time-domain frame cr/ci[16]
template <int SRC>
__device__ __forceinline__ void tx_frame(const SimArgs& a, int64_t b, uint64_t frame, uint32_t bits, float (&cr)[16],
                                         float (&ci)[16]) {
    const ofdmgan_chan_cfg& c = a.cfg;
    if (a.tx) {                                                  // caller-supplied signal: no symbol source, no IFFT
#pragma unroll
        for (int k = 0; k < 16; ++k) { cr[k] = a.tx[b * 32 + k]; ci[k] = a.tx[b * 32 + 16 + k]; }
        return;
    }
    if (SRC == SRC_GAUSS) {
        // (1/sqrt2 per bin) * (ifft 1/N) * (sqrt(N) or N): the IFFT is linear, so the scale is applied to the symbols -
        // for Philox symbols inside Box-Muller's square root (free), for injected symbols by one multiply each
        const float sc = 0.70710678118654752f * (c.ifft_scale == OFDMGAN_SCALE_N ? 1.0f : 0.25f);
        if (a.sym) {
#pragma unroll
            for (int k = 0; k < 16; ++k) { cr[k] = a.sym[b * 32 + k] * sc; ci[k] = a.sym[b * 32 + 16 + k] * sc; }
        } else {
            float n[32];
            draw_section<16>(a, frame, 0u, n, sc * sc);
#pragma unroll
            for (int k = 0; k < 16; ++k) { cr[k] = n[k]; ci[k] = n[16 + k]; }
        }
        fft_inplace<16, +1>(cr, ci);
        return;
    }
    constexpr int N = (SRC == SRC_Q16_CP0 || SRC == SRC_Q16_CP2) ? 16 : 8;
    constexpr int CP = (SRC == SRC_Q16_CP2 || SRC == SRC_Q8_CP2) ? 2 : 0;
    constexpr int NSYM = (16 + N + CP - 1) / (N + CP);             // OFDM symbols started inside the frame
    const float sc = c.ifft_scale == OFDMGAN_SCALE_N ? 1.0f : (N == 16 ? 0.25f : 0.35355339059327376f);
    int bitpos = 0;
#pragma unroll
    for (int s = 0; s < NSYM; ++s) {
        float Xr[N], Xi[N];
#pragma unroll
        for (int k = 0; k < N; ++k) qpsk_fill(c, bits, bitpos, k, Xr[k], Xi[k]);
        fft_inplace<N, +1>(Xr, Xi);
#pragma unroll
        for (int i = 0; i < N + CP; ++i) {
            const int pos = s * (N + CP) + i;
            const int srcI = i < CP ? N - CP + i : i - CP;
            if (pos < 16) { cr[pos] = Xr[srcI] * sc; ci[pos] = Xi[srcI] * sc; }
        }
    }
}

// fading draws of a frame: Philox block 21 (4 normals), block 22 (4 more, multipath), block 12 word 2 (Rician LOS phase)
__device__ __forceinline__ void fade_draws(const SimArgs& a, int64_t b, uint64_t frame, float (&f)[8]) {
    if (a.fade) {
#pragma unroll
        for (int j = 0; j < 8; ++j) f[j] = a.fade[b * 8 + j];
        return;
    }
    float n0[4], n1[4];
    draw_section<2>(a, frame, 21u, n0);
    draw_section<2>(a, frame, 22u, n1);
    if (a.cfg.channel_type == OFDMGAN_CHAN_RICIAN) {
        uint32_t x12[4];
        philox4x32_10(a.keys, (uint32_t)frame, (uint32_t)(frame >> 32), 12u, 0u, x12);
        f[0] = 6.283185307179586f * u_half(x12[2]);
        f[1] = n0[0]; f[2] = n0[1]; f[3] = n0[2];
#pragma unroll
        for (int j = 0; j < 4; ++j) f[4 + j] = n1[j];
    } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) { f[j] = n0[j]; f[4 + j] = n1[j]; }
    }
}

// DC offset, CFO (the tail of apply_all, utils/ofdm_utils.py:524-568) and the fading channels (:710-832), in place
__device__ __forceinline__ void late_stages(const SimArgs& a, int64_t b, uint64_t frame, float (&nr)[16], float (&ni)[16]) {
    const ofdmgan_chan_cfg& c = a.cfg;
    if (c.impair & OFDMGAN_IMPAIR_DC) {
        float P = 0.f;
#pragma unroll
        for (int i = 0; i < 16; ++i) P = fmaf(nr[i], nr[i], fmaf(ni[i], ni[i], P));
        const float mag = sqrtf(P * 0.0625f);
#pragma unroll
        for (int i = 0; i < 16; ++i) { nr[i] = fmaf(mag, c.dc_i, nr[i]); ni[i] = fmaf(mag, c.dc_q, ni[i]); }
    }
    if (c.impair & OFDMGAN_IMPAIR_CFO) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            float s, co;
            sincosf(c.cfo_step * (float)i, &s, &co);
            const float xr = nr[i], xi = ni[i];
            nr[i] = xr * co - xi * s;
            ni[i] = xr * s + xi * co;
        }
    }
    if (c.channel_type == OFDMGAN_CHAN_AWGN) return;
    float f[8];
    fade_draws(a, b, frame, f);
    const float r2 = 0.70710678118654752f;
    if (c.channel_type == OFDMGAN_CHAN_MULTIPATH) {
        // np.convolve(x, h, 'same'): y[n] = sum_t h_t x[n + off - d_t], off = (len(h) - 1) / 2 = max_delay / 2
        int maxd = 0;
        for (int t = 0; t < c.n_taps; ++t) maxd = c.tap_delay[t] > maxd ? c.tap_delay[t] : maxd;
        const int off = maxd / 2;
        float yr[16], yi[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) { yr[i] = 0.f; yi[i] = 0.f; }
        for (int t = 0; t < c.n_taps; ++t) {
            const float hr = c.tap_amp[t] * f[2 * t] * r2, hi = c.tap_amp[t] * f[2 * t + 1] * r2;
            const int sh = off - c.tap_delay[t];
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                float xr = 0.f, xi = 0.f;
#pragma unroll
                for (int j = 0; j < 16; ++j) if (j == i + sh) { xr = nr[j]; xi = ni[j]; }
                yr[i] = fmaf(hr, xr, fmaf(-hi, xi, yr[i]));
                yi[i] = fmaf(hr, xi, fmaf(hi, xr, yi[i]));
            }
        }
#pragma unroll
        for (int i = 0; i < 16; ++i) { nr[i] = yr[i]; ni[i] = yi[i]; }
        return;
    }
    float hr, hi;
    if (c.channel_type == OFDMGAN_CHAN_RAYLEIGH) {
        hr = f[0] * r2; hi = f[1] * r2;
    } else {                                                     // Rician: sqrt(K/(K+1)) e^{j theta} + sqrt(1/(K+1)) CN(0,1)
        const float los = sqrtf(c.rician_k / (c.rician_k + 1.0f)), nl = sqrtf(1.0f / (c.rician_k + 1.0f)) * r2;
        float s, co;
        sincosf(f[0], &s, &co);
        hr = fmaf(nl, f[1], los * co);
        hi = fmaf(nl, f[2], los * s);
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        const float xr = nr[i], xi = ni[i];
        nr[i] = hr * xr - hi * xi;
        ni[i] = hr * xi + hi * xr;
    }
}

// impairments + AWGN on a copy of the clean frame -> nr/ni
// LATE (compile time): also the DC-offset / CFO stages and the fading channels - kept out of the headline instantiations
template <bool LATE>
__device__ __forceinline__ void impair_channel(const SimArgs& a, int64_t b, uint64_t frame, float snr_db,
                                               const float (&cr)[16], const float (&ci)[16], float (&nr)[16],
                                               float (&ni)[16]) {
    const ofdmgan_chan_cfg& c = a.cfg;
    const float g_in = a.tx && a.tx_gain ? a.tx_gain[b] : 1.0f;   // OFDMDataset: the cached clean frame back at signal scale
#pragma unroll
    for (int i = 0; i < 16; ++i) { nr[i] = cr[i] * g_in; ni[i] = ci[i] * g_in; }
    if (c.impair & OFDMGAN_IMPAIR_PA) {
        const float invA2 = 1.0f / (c.pa_saturation * c.pa_saturation);
        const float p = c.pa_smoothness, ninv2p = -0.5f / p;
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const float t = (nr[i] * nr[i] + ni[i] * ni[i]) * invA2;             // (|x|/A)^2
            const float yp = p == 3.0f ? t * t * t : fast_ex2(p * fast_lg2(t));     // (|x|/A)^(2p)
            const float gain = fast_ex2(ninv2p * fast_lg2(1.0f + yp));            // (1+.)^(-1/2p); phase preserved
            nr[i] *= gain; ni[i] *= gain;
        }
    }
    if (LATE && (c.impair & OFDMGAN_IMPAIR_SALEH)) {             // AM/AM + AM/PM: x * alpha_a/(1+beta_a r^2) * e^{j alpha_p r^2/(1+beta_p r^2)}
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const float r2 = nr[i] * nr[i] + ni[i] * ni[i];
            const float gain = c.saleh_alpha_a / (1.0f + c.saleh_beta_a * r2);
            const float phi = c.saleh_alpha_p * r2 / (1.0f + c.saleh_beta_p * r2);
            float s, co;
            __sincosf(phi, &s, &co);
            const float xr = nr[i] * gain, xi = ni[i] * gain;
            nr[i] = xr * co - xi * s;
            ni[i] = xr * s + xi * co;
        }
    }
    if (c.impair & OFDMGAN_IMPAIR_IQ) {
#pragma unroll
        for (int i = 0; i < 16; ++i) ni[i] = c.iq_gain * (c.iq_cos * ni[i] + c.iq_sin * nr[i]);
    }
    if (c.impair & OFDMGAN_IMPAIR_PN) {
        float th = 0.f, n16[16];
        if (a.pn) {
#pragma unroll
            for (int t = 0; t < 16; ++t) n16[t] = a.pn[b * 16 + t];
        } else {
            draw_section<8>(a, frame, 8u, n16);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                const int i = 4 * j + t;
                th = fmaf(c.pn_sigma, n16[i], th);
                const float red = fmaf(-6.283185307179586f, rintf(th * 0.15915494309189535f), th);
                const float s = fast_sin(red), co = fast_cos(red);
                const float xr = nr[i], xi = ni[i];
                nr[i] = xr * co - xi * s;
                ni[i] = xr * s + xi * co;
            }
        }
    }
    if (LATE) late_stages(a, b, frame, nr, ni);
    if (c.snr_mode == OFDMGAN_SNR_NONE) return;
    float P = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) P = fmaf(nr[i], nr[i], fmaf(ni[i], ni[i], P));
    P *= 0.0625f;
    // sigma = sqrt(P / 10^(snr/10) / 2)
    const float sd = fast_sqrt(0.5f * P * fast_ex2(-0.33219280948873623f * snr_db));
    float n32[32];
    if (a.noise) {
#pragma unroll
        for (int t = 0; t < 32; ++t) n32[t] = a.noise[b * 32 + t];
    } else {
        draw_section<16>(a, frame, 13u, n32);
    }
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        nr[k] = fmaf(sd, n32[k], nr[k]);
        ni[k] = fmaf(sd, n32[16 + k], ni[k]);
    }
}

__device__ __forceinline__ float max_abs16(const float (&r)[16], const float (&i)[16]) {
    float m = 0.f;
#pragma unroll
    for (int k = 0; k < 16; ++k) m = fmaxf(m, fmaxf(fabsf(r[k]), fabsf(i[k])));
    return m;
}

__device__ __forceinline__ void normalise(int mode, float (&cr)[16], float (&ci)[16], float (&nr)[16], float (&ni)[16]) {
    if (mode == OFDMGAN_NORM_NONE) return;
    float mc = max_abs16(cr, ci), mn = max_abs16(nr, ni);
    if (mode == OFDMGAN_NORM_JOINT) mc = mn = fmaxf(mc, mn);
    const float sc = mc > 0.f ? 1.0f / mc : 1.0f, sn = mn > 0.f ? 1.0f / mn : 1.0f;
#pragma unroll
    for (int k = 0; k < 16; ++k) { cr[k] *= sc; ci[k] *= sc; nr[k] *= sn; ni[k] *= sn; }
}

// hard QPSK decisions on the complete OFDM symbols inside a frame; returns payload bits compared
template <int SRC>
__device__ __forceinline__ int qpsk_errors(const ofdmgan_chan_cfg& c, const float (&fr)[16], const float (&fi)[16],
                                           uint32_t bits, int& errs) {
    errs = 0;
    if (SRC == SRC_GAUSS || SRC == SRC_Q16_CP2) return 0;
    constexpr int N = SRC == SRC_Q16_CP0 ? 16 : 8;
    constexpr int CP = SRC == SRC_Q8_CP2 ? 2 : 0;
    constexpr int NSYM = 16 / (N + CP);
    int bitpos = 0, nb = 0;
#pragma unroll
    for (int s = 0; s < NSYM; ++s) {
        float Xr[N], Xi[N];
#pragma unroll
        for (int i = 0; i < N; ++i) { Xr[i] = fr[s * (N + CP) + CP + i]; Xi[i] = fi[s * (N + CP) + CP + i]; }
        fft_inplace<N, -1>(Xr, Xi);
#pragma unroll
        for (int k = 0; k < N; ++k) {
            const bool pilot = c.pilot_spacing > 0 && (k % c.pilot_spacing) == 0;
            if (pilot) continue;
            const uint32_t t1 = bitpos < 32 ? (bits >> (31 - bitpos)) & 1u : 0u;
            const uint32_t t0 = bitpos < 31 ? (bits >> (30 - bitpos)) & 1u : 0u;
            bitpos += 2;
            errs += ((Xr[k] < 0.f) != (t1 != 0u)) + ((Xi[k] < 0.f) != (t0 != 0u));
            nb += 2;
        }
    }
    return nb;
}

}  // namespace og
