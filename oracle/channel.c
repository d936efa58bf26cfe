/* TEST INFRASTRUCTURE (oracle) - CPU restatement of the reference's synthetic OFDM channel simulator, of the
 * benchmark metrics, and of the counter-based RNG the CUDA path defines.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference leg may call this; the product never links or loads it.
 *
 * float64 throughout, like the reference's NumPy path, then cast to float32 exactly where the reference casts.
 * PINNED against fixtures recorded from the reference itself with its np.random draws captured
 * (tests/golden/make_reference_fixtures.py -> tests/golden/ref_channel.npz).
 *
 * Follows:
 *   utils/dataset.py:243-247      Gaussian symbols /sqrt2, np.fft.ifft * sqrt(N)
 *   utils/ofdm_utils.py:105-109   QPSK table [1+1j, 1-1j, -1+1j, -1-1j]/sqrt2, bits MSB first (:163-193)
 *   utils/ofdm_utils.py:281-329   OFDMModulator.modulate: data/pilot placement, ifft*N, cyclic prefix, flatten
 *   utils/ofdm_utils.py:923-929   truncate / zero-pad the stream to frame_length
 *   utils/ofdm_utils.py:394-421   Rapp PA      :458-488 IQ imbalance      :491-521 Wiener phase noise
 *   utils/ofdm_utils.py:675-708   AWGN with per-frame measured power
 *   utils/dataset.py:273-287      float32 cast, joint max-abs normalisation
 *   benchmark_comparison.py:129-146,196-197   separate normalisation, MSE, EVM(dB)
 *   utils/ofdm_utils.py:195-222,331-371       demodulate: FFT/N, nearest constellation point (sign decisions)
 *
 * RNG (defined by this project, not by the reference, whose draws come from NumPy's global MT19937):
 *   Philox4x32-10, key = seed, counter = (frame lo, frame hi, block, purpose).  purpose 0 = channel:
 *   Normals come in Box-Muller pairs, THREE per block: pair `slot` uses radius bits x[slot] & 0x7FFFFF (u1 = (m+0.5)*2^-23) and the
 *   16 angle bits x3 & 0xFFFF | x3 >> 16 | (x0>>24) | (x1>>24)<<8 (theta = 2 pi a / 65536); (r cos theta, r sin theta), r = sqrt(-2 ln u1).
 *   A section of n pairs starting at block b0 puts pair p in block b0 + p/3, slot p%3.  Sections: blocks 0-5 the 16 symbol pairs
 *   (normals 0-15 Re, 16-31 Im), 8-10 the 8 phase-noise pairs, 12 = {snr uniform, payload bits, LOS phase, -}, 13-18 the 16 noise
 *   pairs, 21 / 22 two fading pairs each.  purpose 1 = gradient-penalty alpha (block = critic iteration).
 *   Plain uniforms (snr, alpha, LOS phase): u = (x>>8)*2^-24.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../include/ofdmgan.h"

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

/* ------------------------------------------------------------------ Philox4x32-10 ------------------------ */
static int g_rounds = 10;   /* set per call from cfg->rng_rounds by the entry points below (the oracle is test infrastructure: not re-entrant across
                               different round counts at the same time; every OpenMP thread of one call sees the same value) */
static void set_rounds(const ofdmgan_chan_cfg* cfg) { g_rounds = (cfg && cfg->rng_rounds == 7) ? 7 : 10; }
void oracle_philox4x32_r(uint32_t k0, uint32_t k1, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, int rounds, uint32_t out[4]);
void oracle_philox4x32_10(uint32_t k0, uint32_t k1, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t out[4]) {
    oracle_philox4x32_r(k0, k1, c0, c1, c2, c3, 10, out);
}
void oracle_philox4x32_r(uint32_t k0, uint32_t k1, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, int rounds, uint32_t out[4]) {
    for (int r = 0; r < rounds; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

void oracle_philox_blocks(uint64_t seed, uint64_t ctr0, uint32_t c2, uint32_t c3, uint32_t* out, int64_t n) {
    for (int64_t i = 0; i < n; ++i) {
        uint64_t c = ctr0 + (uint64_t)i;
        oracle_philox4x32_10((uint32_t)seed, (uint32_t)(seed >> 32), (uint32_t)c, (uint32_t)(c >> 32), c2, c3, out + 4 * i);
    }
}

static inline double u_open(uint32_t x) { return ((double)(x >> 9) + 0.5) * (1.0 / 8388608.0); }
static inline double u_half(uint32_t x) { return (double)(x >> 8) * (1.0 / 16777216.0); }

/* 2*npairs normals of the section that starts at block blk0: pair p = block blk0 + p/3, slot p%3 (csrc/common.cuh) */
static void section_normals(uint64_t seed, uint64_t frame, uint32_t blk0, uint32_t purpose, int npairs, double* out) {
    uint32_t x[4] = {0, 0, 0, 0};
    for (int p = 0; p < npairs; ++p) {
        int slot = p % 3;
        if (slot == 0)
            oracle_philox4x32_r((uint32_t)seed, (uint32_t)(seed >> 32), (uint32_t)frame, (uint32_t)(frame >> 32), blk0 + (uint32_t)(p / 3),
                                purpose, purpose == 0 ? g_rounds : 10, x);
        uint32_t a16 = slot == 0 ? (x[3] & 0xFFFFu) : slot == 1 ? (x[3] >> 16) : ((x[0] >> 24) | ((x[1] >> 24) << 8));
        double u1 = ((double)(x[slot] & 0x7FFFFFu) + 0.5) * (1.0 / 8388608.0);
        double r = sqrt(-2.0 * log(u1)), th = 2.0 * M_PI * (double)a16 * (1.0 / 65536.0);
        out[2 * p] = r * cos(th);
        out[2 * p + 1] = r * sin(th);
    }
}

/* the draws frame `frame` consumes; any output may be NULL */
static void frame_draws(const ofdmgan_chan_cfg* cfg, uint64_t seed, uint64_t frame, double* sym32, uint32_t* bits,
                        double* pn16, double* snr_db, double* noise32);
void oracle_frame_draws(const ofdmgan_chan_cfg* cfg, uint64_t seed, uint64_t frame, double* sym32, uint32_t* bits,
                        double* pn16, double* snr_db, double* noise32) {
    set_rounds(cfg);
    frame_draws(cfg, seed, frame, sym32, bits, pn16, snr_db, noise32);
}
static void frame_draws(const ofdmgan_chan_cfg* cfg, uint64_t seed, uint64_t frame, double* sym32, uint32_t* bits,
                        double* pn16, double* snr_db, double* noise32) {
    if (sym32) section_normals(seed, frame, 0, 0, 16, sym32);
    if (pn16) section_normals(seed, frame, 8, 0, 8, pn16);
    if (noise32) section_normals(seed, frame, 13, 0, 16, noise32);
    if (bits || snr_db) {
        uint32_t x[4];
        oracle_philox4x32_r((uint32_t)seed, (uint32_t)(seed >> 32), (uint32_t)frame, (uint32_t)(frame >> 32), 12, 0, g_rounds, x);
        if (bits) *bits = x[1];
        if (snr_db) *snr_db = (double)cfg->snr_lo + ((double)cfg->snr_hi - (double)cfg->snr_lo) * u_half(x[0]);
    }
}

double oracle_snr_of_frame(const ofdmgan_chan_cfg* cfg, uint64_t frame, double uniform_draw_db) {
    if (cfg->snr_mode == OFDMGAN_SNR_GRID) {
        uint64_t fps = cfg->frames_per_snr > 0 ? (uint64_t)cfg->frames_per_snr : 1;
        uint64_t bin = (frame / fps) % (uint64_t)(cfg->n_snr > 0 ? cfg->n_snr : 1);
        return (double)cfg->snr_lo + (double)cfg->snr_step * (double)bin;
    }
    return uniform_draw_db;
}

int oracle_snr_bin(const ofdmgan_chan_cfg* cfg, uint64_t frame) {
    if (cfg->snr_mode != OFDMGAN_SNR_GRID) return 0;
    uint64_t fps = cfg->frames_per_snr > 0 ? (uint64_t)cfg->frames_per_snr : 1;
    return (int)((frame / fps) % (uint64_t)(cfg->n_snr > 0 ? cfg->n_snr : 1));
}

/* ------------------------------------------------------------------ transmit side ------------------------ */
/* naive inverse DFT, x[n] = (scale/N) sum_k X[k] e^{+j 2 pi k n / N} */
static void idft(const double* Xr, const double* Xi, int N, double scale, double* xr, double* xi) {
    for (int n = 0; n < N; ++n) {
        double sr = 0, si = 0;
        for (int k = 0; k < N; ++k) {
            double a = 2.0 * M_PI * (double)((k * n) % N) / (double)N;
            double c = cos(a), s = sin(a);
            sr += Xr[k] * c - Xi[k] * s;
            si += Xr[k] * s + Xi[k] * c;
        }
        xr[n] = sr * scale / (double)N;
        xi[n] = si * scale / (double)N;
    }
}

static int is_pilot(const ofdmgan_chan_cfg* cfg, int k) { return cfg->pilot_spacing > 0 && (k % cfg->pilot_spacing) == 0; }

/* clean time-domain frame (16 complex samples) from the draws */
static void tx_frame(const ofdmgan_chan_cfg* cfg, const double* sym32, uint32_t bits, double* xr, double* xi) {
    int N = cfg->n_fft;
    double scale = cfg->ifft_scale == OFDMGAN_SCALE_N ? (double)N : sqrt((double)N);
    if (cfg->symbol_source == OFDMGAN_SYM_GAUSSIAN) {
        /* utils/dataset.py:243-247 - one N=16 symbol, no CP, no pilots */
        double Xr[16], Xi[16];
        for (int k = 0; k < 16; ++k) { Xr[k] = sym32[k] / sqrt(2.0); Xi[k] = sym32[16 + k] / sqrt(2.0); }
        idft(Xr, Xi, 16, cfg->ifft_scale == OFDMGAN_SCALE_N ? 16.0 : 4.0, xr, xi);
        return;
    }
    /* QPSK: stream of OFDM symbols (each cp + N samples), truncated to 16 */
    int bitpos = 0, outpos = 0;
    for (int i = 0; i < 16; ++i) xr[i] = xi[i] = 0.0;
    while (outpos < 16) {
        double Xr[16], Xi[16], tr[16], ti[16];
        for (int k = 0; k < N; ++k) {
            if (is_pilot(cfg, k)) { Xr[k] = cfg->pilot_re; Xi[k] = cfg->pilot_im; continue; }
            int b1 = bitpos < 32 ? (int)((bits >> (31 - bitpos)) & 1u) : 0; ++bitpos;     /* MSB: Re sign */
            int b0 = bitpos < 32 ? (int)((bits >> (31 - bitpos)) & 1u) : 0; ++bitpos;     /* LSB: Im sign */
            Xr[k] = (b1 ? -1.0 : 1.0) / sqrt(2.0);
            Xi[k] = (b0 ? -1.0 : 1.0) / sqrt(2.0);
        }
        idft(Xr, Xi, N, scale, tr, ti);
        for (int i = 0; i < cfg->cp_len && outpos < 16; ++i, ++outpos) { xr[outpos] = tr[N - cfg->cp_len + i]; xi[outpos] = ti[N - cfg->cp_len + i]; }
        for (int i = 0; i < N && outpos < 16; ++i, ++outpos) { xr[outpos] = tr[i]; xi[outpos] = ti[i]; }
    }
}

/* ------------------------------------------------------------------ impairments + channel ---------------- */
static void impair_and_channel(const ofdmgan_chan_cfg* cfg, const double* pn16, double snr_db, const double* noise32,
                               const double* xr, const double* xi, double* yr, double* yi) {
    double r[16], q[16];
    for (int i = 0; i < 16; ++i) { r[i] = xr[i]; q[i] = xi[i]; }
    if (cfg->impair & OFDMGAN_IMPAIR_PA) {
        double p = cfg->pa_smoothness, A = cfg->pa_saturation;
        for (int i = 0; i < 16; ++i) {
            double amp = hypot(r[i], q[i]), ph = atan2(q[i], r[i]);
            double gain = 1.0 / pow(1.0 + pow(amp / A, 2.0 * p), 1.0 / (2.0 * p));
            double oa = amp * gain;
            r[i] = oa * cos(ph); q[i] = oa * sin(ph);
        }
    }
    if (cfg->impair & OFDMGAN_IMPAIR_IQ) {
        for (int i = 0; i < 16; ++i) q[i] = (double)cfg->iq_gain * ((double)cfg->iq_cos * q[i] + (double)cfg->iq_sin * r[i]);
    }
    if (cfg->impair & OFDMGAN_IMPAIR_PN) {
        double th = 0;
        for (int i = 0; i < 16; ++i) {
            th += (double)cfg->pn_sigma * pn16[i];
            double c = cos(th), s = sin(th), a = r[i], b = q[i];
            r[i] = a * c - b * s; q[i] = a * s + b * c;
        }
    }
    double P = 0;
    for (int i = 0; i < 16; ++i) P += r[i] * r[i] + q[i] * q[i];
    P /= 16.0;
    /* OFDMGAN_SNR_NONE: the impairments on their own (NonLinearImpairments.apply_* without ChannelModel.apply) */
    double sd = cfg->snr_mode == OFDMGAN_SNR_NONE ? 0.0 : sqrt(P / pow(10.0, snr_db / 10.0) / 2.0);
    for (int i = 0; i < 16; ++i) { yr[i] = r[i] + sd * noise32[i]; yi[i] = q[i] + sd * noise32[16 + i]; }
}

static void normalise(int mode, float* clean, float* noisy) {
    float mc = 0, mn = 0;
    for (int i = 0; i < 32; ++i) { mc = fmaxf(mc, fabsf(clean[i])); mn = fmaxf(mn, fabsf(noisy[i])); }
    if (mode == OFDMGAN_NORM_JOINT) {
        float m = fmaxf(mc, mn);
        if (m > 0) for (int i = 0; i < 32; ++i) { clean[i] = clean[i] / m; noisy[i] = noisy[i] / m; }
    } else if (mode == OFDMGAN_NORM_SEPARATE) {
        if (mc > 0) for (int i = 0; i < 32; ++i) clean[i] = clean[i] / mc;
        if (mn > 0) for (int i = 0; i < 32; ++i) noisy[i] = noisy[i] / mn;
    }
}

/* One frame: draws (injected when the pointer is non-NULL, else Philox) -> clean, noisy (float32, normalised) */
static void sim_frame(const ofdmgan_chan_cfg* cfg, const double* sym, const uint32_t* bits, const double* pn,
                      const double* snr_db, const double* noise, uint64_t seed, uint64_t frame, float* clean,
                      float* noisy, float* snr_out, uint32_t* bits_out) {
    double s32[32], p16[16], n32[32], snr_u = 0, xr[16], xi[16], yr[16], yi[16];
    uint32_t bw = 0;
    frame_draws(cfg, seed, frame, sym ? NULL : s32, bits ? NULL : &bw, pn ? NULL : p16, snr_db ? NULL : &snr_u,
                       noise ? NULL : n32);
    if (sym) memcpy(s32, sym, sizeof s32);
    if (pn) memcpy(p16, pn, sizeof p16);
    if (noise) memcpy(n32, noise, sizeof n32);
    if (bits) bw = *bits;
    if (snr_db) snr_u = *snr_db;
    double snr = oracle_snr_of_frame(cfg, frame, snr_u);
    tx_frame(cfg, s32, bw, xr, xi);
    impair_and_channel(cfg, p16, snr, n32, xr, xi, yr, yi);
    for (int i = 0; i < 16; ++i) {
        clean[i] = (float)xr[i]; clean[16 + i] = (float)xi[i];
        noisy[i] = (float)yr[i]; noisy[16 + i] = (float)yi[i];
    }
    normalise(cfg->normalize, clean, noisy);
    if (snr_out) *snr_out = (float)snr;
    if (bits_out) *bits_out = bw;
}

int oracle_chan_sim(const ofdmgan_chan_cfg* cfg, const double* sym, const uint32_t* bits, const double* pn,
                    const double* snr_db, const double* noise, uint64_t seed, uint64_t frame0, float* clean,
                    float* noisy, float* snr_out, int64_t B) {
    if (!cfg || (cfg->n_fft != 8 && cfg->n_fft != 16) || cfg->cp_len < 0 || cfg->cp_len > cfg->n_fft) return -1;
    set_rounds(cfg);
#pragma omp parallel for schedule(static)
    for (int64_t b = 0; b < B; ++b) {
        float c[32], n[32], s;
        sim_frame(cfg, sym ? sym + 32 * b : NULL, bits ? bits + b : NULL, pn ? pn + 16 * b : NULL,
                  snr_db ? snr_db + b : NULL, noise ? noise + 32 * b : NULL, seed, frame0 + (uint64_t)b, c, n, &s, NULL);
        if (clean) memcpy(clean + 32 * b, c, sizeof c);
        if (noisy) memcpy(noisy + 32 * b, n, sizeof n);
        if (snr_out) snr_out[b] = s;
    }
    return 0;
}

/* ------------------------------------------------------------------ metrics ------------------------------ */
/* benchmark_comparison.py:137-146 (float32 arrays in, python floats out) */
static void frame_mse_evm(const float* est, const float* ref, double* mse, double* evm_db, double* ratio) {
    double se = 0, sr = 0;
    for (int i = 0; i < 32; ++i) { double e = (double)est[i] - (double)ref[i]; se += e * e; sr += (double)ref[i] * (double)ref[i]; }
    *mse = se / 32.0;
    *ratio = se / sr;
    *evm_db = 20.0 * log10(sqrt((se / 32.0) / (sr / 32.0)) + 1e-10);
}

/* hard QPSK decisions on the complete OFDM symbols contained in a 16-sample frame; returns number of payload
 * bits compared, *errs = mismatches against the transmitted word */
static int qpsk_bit_errors(const ofdmgan_chan_cfg* cfg, const float* frame, uint32_t bits, int* errs) {
    int N = cfg->n_fft, per = N + cfg->cp_len, nsym = 16 / per, bitpos = 0, nb = 0, e = 0;
    for (int s = 0; s < nsym; ++s) {
        const float* fr = frame + s * per + cfg->cp_len;
        const float* fi = frame + 16 + s * per + cfg->cp_len;
        for (int k = 0; k < N; ++k) {
            if (is_pilot(cfg, k)) continue;
            double sr = 0, si = 0;
            for (int n = 0; n < N; ++n) {
                double a = -2.0 * M_PI * (double)((k * n) % N) / (double)N, c = cos(a), sn = sin(a);
                sr += fr[n] * c - fi[n] * sn;
                si += fr[n] * sn + fi[n] * c;
            }
            int b1 = sr < 0, b0 = si < 0;                         /* argmin ties -> lowest index -> bit 0 */
            int t1 = bitpos < 32 ? (int)((bits >> (31 - bitpos)) & 1u) : 0; ++bitpos;
            int t0 = bitpos < 32 ? (int)((bits >> (31 - bitpos)) & 1u) : 0; ++bitpos;
            e += (b1 != t1) + (b0 != t0);
            nb += 2;
        }
    }
    *errs = e;
    return nb;
}

void oracle_metrics_add(double* row, double mse, double evm, double ratio, int errs, int nbits) {
    row[0] += 1.0; row[1] += mse; row[2] += mse * mse; row[3] += evm; row[4] += evm * evm;
    row[5] += errs; row[6] += nbits; row[7] += ratio;
}

/* metrics for frames that already exist (mirror of ofdmgan_frame_metrics) */
int oracle_frame_metrics(const float* est, const float* ref, const int32_t* bin, int method, int n_snr, int64_t B,
                         double* metrics) {
    for (int64_t b = 0; b < B; ++b) {
        double mse, evm, ratio;
        frame_mse_evm(est + 32 * b, ref + 32 * b, &mse, &evm, &ratio);
        int s = bin ? bin[b] : 0;
        if (s < 0 || s >= n_snr) return -1;
        oracle_metrics_add(metrics + ((size_t)s * OFDMGAN_N_METHODS + method) * OFDMGAN_METRIC_COLS, mse, evm, ratio, 0, 0);
    }
    return 0;
}

/* ------------------------------------------------------------------ fused restatement --------------------- */
int oracle_gen_fwd_f32(const float* x, const float* gp, float* y, int64_t B, float slope);
int oracle_gen_fwd_q(const int16_t* x, const int8_t* W, const int16_t* Bq, int16_t* y, int64_t B, int mode);

void oracle_equalize_frame(const float* noisy, const float* clean, double snr_db, int method, float* est);

/* mirror of ofdmgan_sim_gen_metrics: simulate -> reconstruct -> accumulate.  Also the CPU baseline that bench.py
 * times (OpenMP over frames, per-thread accumulators merged in thread order). */
int oracle_sim_gen_metrics(const ofdmgan_chan_cfg* cfg, int gen_kind, const float* gparams, const int8_t* wrom,
                           const int16_t* brom, float slope, uint64_t seed, uint64_t frame0, int64_t B, double* metrics) {
    int n_snr = cfg->snr_mode == OFDMGAN_SNR_GRID ? cfg->n_snr : 1;
    if (n_snr < 1 || n_snr > OFDMGAN_MAX_SNR_BINS) return -1;
    size_t rows = (size_t)n_snr * OFDMGAN_N_METHODS * OFDMGAN_METRIC_COLS;
    int rc = 0;
    set_rounds(cfg);
#pragma omp parallel
    {
        double* local = (double*)calloc(rows, sizeof(double));
#pragma omp for schedule(static)
        for (int64_t b = 0; b < B; ++b) {
            float clean[32], noisy[32], y[32];
            uint32_t bw;
            uint64_t frame = frame0 + (uint64_t)b;
            sim_frame(cfg, NULL, NULL, NULL, NULL, NULL, seed, frame, clean, noisy, NULL, &bw);
            if (gen_kind == OFDMGAN_GEN_F32) {
                oracle_gen_fwd_f32(noisy, gparams, y, 1, slope);
            } else {
                int16_t xq[32], yq[32];
                for (int i = 0; i < 32; ++i) xq[i] = (int16_t)(noisy[i] * 256.0f);        /* truncation toward zero */
                oracle_gen_fwd_q(xq, wrom, brom, yq, 1, gen_kind == OFDMGAN_GEN_Q_SPEC ? 0 : 1);
                for (int i = 0; i < 32; ++i) y[i] = (float)yq[i] * (1.0f / 256.0f);
            }
            int bin = oracle_snr_bin(cfg, frame);
            double mse, evm, ratio;
            int errs = 0, nb = 0;
            frame_mse_evm(y, clean, &mse, &evm, &ratio);
            if (cfg->symbol_source == OFDMGAN_SYM_QPSK) nb = qpsk_bit_errors(cfg, y, bw, &errs);
            oracle_metrics_add(local + ((size_t)bin * OFDMGAN_N_METHODS + OFDMGAN_METHOD_GAN) * OFDMGAN_METRIC_COLS, mse, evm, ratio, errs, nb);
            frame_mse_evm(noisy, clean, &mse, &evm, &ratio);
            errs = nb = 0;
            if (cfg->symbol_source == OFDMGAN_SYM_QPSK) nb = qpsk_bit_errors(cfg, noisy, bw, &errs);
            oracle_metrics_add(local + ((size_t)bin * OFDMGAN_N_METHODS + OFDMGAN_METHOD_NOEQ) * OFDMGAN_METRIC_COLS, mse, evm, ratio, errs, nb);
            if (cfg->equalizers) {
                double snr_db = oracle_snr_of_frame(cfg, frame, 0.0);
                if (cfg->snr_mode != OFDMGAN_SNR_GRID) {
                    uint32_t x12[4];
                    oracle_philox4x32_r((uint32_t)seed, (uint32_t)(seed >> 32), (uint32_t)frame, (uint32_t)(frame >> 32), 12, 0, g_rounds, x12);
                    snr_db = (double)(float)((double)cfg->snr_lo + ((double)cfg->snr_hi - (double)cfg->snr_lo) * u_half(x12[0]));
                }
                for (int method = OFDMGAN_METHOD_ZF; method <= OFDMGAN_METHOD_MMSE; ++method) {
                    float est[32];
                    oracle_equalize_frame(noisy, clean, snr_db, method, est);
                    frame_mse_evm(est, clean, &mse, &evm, &ratio);
                    errs = nb = 0;
                    if (cfg->symbol_source == OFDMGAN_SYM_QPSK) nb = qpsk_bit_errors(cfg, est, bw, &errs);
                    oracle_metrics_add(local + ((size_t)bin * OFDMGAN_N_METHODS + method) * OFDMGAN_METRIC_COLS, mse, evm, ratio, errs, nb);
                }
            }
        }
#pragma omp critical
        for (size_t i = 0; i < rows; ++i) metrics[i] += local[i];
        free(local);
    }
    return rc;
}

/* QPSK helpers exposed for the API-level parity tests (QAMModulator.modulate / demodulate) */
int oracle_qpsk_bit_errors(const ofdmgan_chan_cfg* cfg, const float* frames, const uint32_t* bits, int64_t B,
                           int64_t* errs_out, int64_t* nbits_out) {
    int64_t e = 0, n = 0;
    for (int64_t b = 0; b < B; ++b) { int eb; n += qpsk_bit_errors(cfg, frames + 32 * b, bits[b], &eb); e += eb; }
    *errs_out = e; *nbits_out = n;
    return 0;
}

/* ------------------------------------------------------------------ classical equalisers ----------------- */
/* Genie-aided ZF and MMSE as benchmark_comparison.py:218-226 calls them (utils/classical_equalizers.py:33-230), in the
 * complex64 arithmetic NumPy 2 gives the reference: CFLOAT_divide (Smith's algorithm, separately rounded float
 * operations), eps = float32(1e-10) added to the real part.  Compiled without FP contraction (-std=c11), so the float
 * sequence below is exactly NumPy's; ZF frames are bit-identical to the reference's (tests/golden/ref_eq.npz). */
static void cdiv_np(float ar, float ai, float br, float bi, float* outr, float* outi) {
    float abr = fabsf(br), abi = fabsf(bi);
    if (abr >= abi) {
        if (abr == 0.f && abi == 0.f) { *outr = ar / abr; *outi = ai / abr; return; }
        float rat = bi / br;
        float scl = 1.0f / (br + bi * rat);
        *outr = (ar + ai * rat) * scl;
        *outi = (ai - ar * rat) * scl;
    } else {
        float rat = br / bi;
        float scl = 1.0f / (bi + br * rat);
        *outr = (ar * rat + ai) * scl;
        *outi = (ai * rat - ar) * scl;
    }
}

/* method: OFDMGAN_METHOD_ZF or OFDMGAN_METHOD_MMSE; est[32] = equalised frame */
void oracle_equalize_frame(const float* noisy, const float* clean, double snr_db, int method, float* est) {
    const float eps = 1e-10f;
    float inv_snr = (float)(1.0 / pow(10.0, snr_db / 10.0));
    for (int i = 0; i < 16; ++i) {
        float yr = noisy[i], yi = noisy[16 + i], hr, hi;
        cdiv_np(yr, yi, clean[i] + eps, clean[16 + i], &hr, &hi);
        if (method == OFDMGAN_METHOD_ZF) {
            cdiv_np(yr, yi, hr + eps, hi, &est[i], &est[16 + i]);
        } else {
            float a = (float)sqrt((double)hr * (double)hr + (double)hi * (double)hi);
            float den = a * a + inv_snr;
            float scl = 1.0f / den;
            float fr = hr * scl, fi = -hi * scl;
            float p0 = fr * yr, p1 = fi * yi, p2 = fr * yi, p3 = fi * yr;
            est[i] = p0 - p1;
            est[16 + i] = p2 + p3;
        }
    }
}

int oracle_equalize(const float* noisy, const float* clean, const float* snr_db, int method, float* est, int64_t B) {
    if (method != OFDMGAN_METHOD_ZF && method != OFDMGAN_METHOD_MMSE) return -1;
    for (int64_t b = 0; b < B; ++b) oracle_equalize_frame(noisy + 32 * b, clean + 32 * b, snr_db ? (double)snr_db[b] : 20.0, method, est + 32 * b);
    return 0;
}

/* Thread count of this library's OpenMP regions (bench.py's CPU-baseline legs: torchrun exports OMP_NUM_THREADS=1 to its workers).
 * Lives here so that it reaches whichever OpenMP runtime this shared object is bound to. */
#ifdef _OPENMP
#include <omp.h>
int oracle_set_threads(int n) { if (n > 0) omp_set_num_threads(n); return omp_get_max_threads(); }
#else
int oracle_set_threads(int n) { (void)n; return 1; }
#endif
