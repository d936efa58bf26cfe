/* libofdmgan - C ABI of the B200 (sm_100a) hot path of ofdm-gan-sr.
 *
 * The reference (orpheus016/ofdm-gan-sr) is pure PyTorch/NumPy and has NO plugin / custom-op / FFI layer
 * (SURVEY.md section 8b).  Its drop-in boundary is the Python call surface; the entry points below are what a
 * ctypes binding underneath those unchanged Python signatures calls.  Each declaration cites the reference
 * interface it replaces.  INTEGRATION.md shows the reference-side stub.
 *
 * Conventions
 *  - plain pointers and sizes only; frames are [B][2][16] (I row then Q row), contiguous.
 *  - pointers named *_dev are device pointers owned by the caller (e.g. torch's caching allocator); pointers
 *    named *_host are host pointers; `params` pointers may be either (cudaMemcpyDefault is used).
 *  - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).  Every call is asynchronous
 *    on that stream unless it says "host" in its name; *_host entry points take host buffers, do their own
 *    H2D / D2H copies and synchronise the stream before returning.
 *  - returns 0 on success, a positive cudaError_t, or a negative OFDMGAN_E_* argument error.  Never throws,
 *    never calls exit().  Safe to call from several host threads and on several streams.
 *  - weights: a `params` / ROM pointer in HOST memory (inference) is turned into the kernel's weight image on the host and passed BY
 *    VALUE in the kernel's parameter block: ofdmgan_gen_fwd_f32, ofdmgan_gen_fwd_q and the headline configuration of
 *    ofdmgan_sim_gen_metrics(_host) (ofdmgan_sim_impl_for == 1) then touch no shared state, take no lock, and calls on different
 *    streams run independently (scratch is per stream).  A pointer in DEVICE memory (training: the optimiser updates the weights on
 *    the device) goes through one constant-memory image per network and device, refreshed stream-ordered before the launch; those
 *    calls, and the general simulator kernel, are serialised inside the library (a call on a new stream is ordered after the
 *    previous such call's stream).
 *  - there is NO CPU fallback: without a CUDA device every compute entry point returns a cudaError_t.
 */
#ifndef OFDMGAN_H
#define OFDMGAN_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define OFDMGAN_ABI_VERSION 16
#define OFDMGAN_FRAME_LEN 16
#define OFDMGAN_FRAME_ELEMS 32            /* 2 x 16 */
#define OFDMGAN_G_NPARAMS 258             /* models/generator.py:102 */
#define OFDMGAN_D_NPARAMS 521             /* models/discriminator.py:61 */
#define OFDMGAN_WROM_DEPTH 2048           /* rtl/ofdmGAN/weight_rom.v:14 */
#define OFDMGAN_BROM_DEPTH 64             /* rtl/ofdmGAN/weight_rom.v:186 */
#define OFDMGAN_MAX_SNR_BINS 16

#define OFDMGAN_E_ARG (-1)                /* null pointer / bad enum / bad size */
#define OFDMGAN_E_STREAMS (-2)            /* reserved (stream bookkeeping failure) */
#define OFDMGAN_E_UNSUPPORTED (-3)        /* valid in the reference but not built here (named in DESIGN.md) */
#define OFDMGAN_E_COMM (-4)               /* a data-parallel peer did not arrive within the wait limit (see ofdmgan_comm_create) */

/* Parameter packing = torch parameters()/state_dict order, flattened (models/generator.py:129-164,
 * models/discriminator.py:78-100):
 *   G: enc1.conv.weight[4][2][3]@0 .bias[4]@24 bottleneck.conv.weight[8][4][3]@28 .bias[8]@124
 *      dec1.conv.weight[4][8][3]@132 .bias[4]@228 out_conv.weight[2][4][3]@232 .bias[2]@256
 *   D: conv1.weight[8][4][3]@0 .bias[8]@96 conv2.weight[16][8][3]@104 .bias[16]@488 dense.weight[16]@504 .bias@520 */

/* ---- channel simulator configuration (utils/dataset.py:195-293, utils/ofdm_utils.py:229-329,378-605,675-708) */
enum { OFDMGAN_SYM_GAUSSIAN = 0,          /* (randn+j randn)/sqrt2 per bin            utils/dataset.py:243-244 */
       OFDMGAN_SYM_QPSK = 1 };            /* QAMModulator('QPSK') + OFDMModulator     utils/ofdm_utils.py:105-109,281-329 */
enum { OFDMGAN_SCALE_SQRT_N = 0,          /* ifft * sqrt(N)                           utils/dataset.py:247 */
       OFDMGAN_SCALE_N = 1 };             /* ifft * N                                 utils/ofdm_utils.py:320 */
enum { OFDMGAN_IMPAIR_PA = 1, OFDMGAN_IMPAIR_IQ = 2, OFDMGAN_IMPAIR_PN = 4,     /* apply_all order :571-605 */
       OFDMGAN_IMPAIR_SALEH = 8,          /* apply_pa_saleh :424-455 (in the PA position)      */
       OFDMGAN_IMPAIR_DC = 16,            /* apply_dc_offset :524-543                           */
       OFDMGAN_IMPAIR_CFO = 32 };         /* apply_cfo :546-568                                 */
enum { OFDMGAN_CHAN_AWGN = 0,             /* ChannelModel._apply_awgn      :675-708 */
       OFDMGAN_CHAN_RAYLEIGH = 1,         /* ChannelModel._apply_rayleigh  :710-740 : y = h x + n, h ~ CN(0,1) per frame */
       OFDMGAN_CHAN_RICIAN = 2,           /* ChannelModel._apply_rician    :742-786 */
       OFDMGAN_CHAN_MULTIPATH = 3 };      /* ChannelModel._apply_multipath :788-832 : np.convolve(x, h, 'same'), Rayleigh taps */
#define OFDMGAN_MAX_TAPS 4
enum { OFDMGAN_SNR_UNIFORM = 0,           /* np.random.uniform(lo, hi) per frame      utils/dataset.py:267 */
       OFDMGAN_SNR_GRID = 1,              /* snr = lo + step*((frame/frames_per_snr) % n_snr)  benchmark_comparison.py:179-182 */
       OFDMGAN_SNR_NONE = 2 };            /* no AWGN stage (NonLinearImpairments.apply_* on their own) */
enum { OFDMGAN_NORM_NONE = 0,
       OFDMGAN_NORM_JOINT = 1,            /* one max over noisy U clean               utils/dataset.py:284-287 */
       OFDMGAN_NORM_SEPARATE = 2 };       /* each by its own max                      benchmark_comparison.py:129-134,196-197 */
enum { OFDMGAN_GEN_F32 = 0,               /* MiniGenerator.forward fp32               models/generator.py:180-208 */
       OFDMGAN_GEN_Q_SPEC = 1,            /* Q1.7/Q8.8 integer generator, clean dataflow */
       OFDMGAN_GEN_Q_RTL = 2 };           /* Q1.7/Q8.8 integer generator, literal RTL behaviour (generator_mini.v) */

typedef struct ofdmgan_chan_cfg {
    int32_t symbol_source;     /* OFDMGAN_SYM_* */
    int32_t n_fft;             /* 16 or 8 subcarriers */
    int32_t cp_len;            /* cyclic prefix length, 0..n_fft (OFDMModulator.cp_length) */
    int32_t pilot_spacing;     /* 0 = no pilots, else pilots at arange(0, n_fft, spacing) (QPSK source only) */
    float   pilot_re, pilot_im;
    int32_t ifft_scale;        /* OFDMGAN_SCALE_* */
    int32_t impair;            /* OR of OFDMGAN_IMPAIR_* */
    float   pa_saturation;     /* A_sat            apply_pa_rapp :394 */
    float   pa_smoothness;     /* p (3.0)          apply_pa_rapp :394 */
    float   iq_gain;           /* 10^(dB/20)       apply_iq_imbalance :478 */
    float   iq_cos, iq_sin;    /* cos/sin(deg2rad) apply_iq_imbalance :479 */
    float   pn_sigma;          /* sqrt(10^(dBc/10) * sample_rate)   apply_phase_noise :514-515 */
    int32_t snr_mode;          /* OFDMGAN_SNR_* */
    float   snr_lo, snr_hi;    /* dB */
    float   snr_step;          /* dB, grid mode */
    int32_t n_snr;             /* grid mode: number of grid points (<= OFDMGAN_MAX_SNR_BINS); uniform mode: 1 */
    int64_t frames_per_snr;    /* grid mode: consecutive frames sharing one grid point (n_trials) */
    int32_t normalize;         /* OFDMGAN_NORM_* */
    int32_t equalizers;        /* != 0: the fused sweep also fills the OFDMGAN_METHOD_ZF / _MMSE rows (genie-aided equalisers of
                                  benchmark_comparison.py:218-226, utils/classical_equalizers.py:33-230) */
    int32_t channel_type;      /* OFDMGAN_CHAN_* (fading is applied after the impairments, before the AWGN) */
    float   rician_k;          /* K factor                                                   :745 */
    int32_t n_taps;            /* multipath: number of taps (<= OFDMGAN_MAX_TAPS), distinct delays */
    int32_t tap_delay[OFDMGAN_MAX_TAPS];
    float   tap_amp[OFDMGAN_MAX_TAPS];     /* sqrt(power / sum(powers))                       :806-812 */
    float   saleh_alpha_a, saleh_beta_a, saleh_alpha_p, saleh_beta_p;
    float   dc_i, dc_q;        /* DC offset relative to sqrt(mean |x|^2)                     :541-543 */
    float   cfo_step;          /* 2 pi cfo_hz / sample_rate (radians per sample)             :565-566 */
    int32_t rng_rounds;        /* rounds of the Philox4x32 counter RNG behind the on-chip draws: 0 or 10 = Philox4x32-10 (the default and
                                  the parity workload); 7 = Philox4x32-7, the smallest round count Random123 documents as passing
                                  BigCrush - a separately named "fast RNG" workload, built for the headline configurations only
                                  (ofdmgan_sim_impl_for == 1 without injected draws; otherwise OFDMGAN_E_UNSUPPORTED) */
} ofdmgan_chan_cfg;

/* Host-generated randomness for parity runs, in the reference's np.random draw order per frame
 * (utils/dataset.py:243, utils/ofdm_utils.py:518, utils/dataset.py:267, utils/ofdm_utils.py:696-697).
 * All device pointers; any member may be NULL to take that draw from Philox instead. */
typedef struct ofdmgan_chan_rand {
    const float*    sym;       /* [B][32]  randn Re[16], randn Im[16]   (gaussian source) */
    const uint32_t* bits;      /* [B]      32 payload bits, stream bit i = (word >> (31-i)) & 1 (QPSK source) */
    const float*    pn;        /* [B][16]  randn phase increments */
    const float*    snr_db;    /* [B]      the uniform(lo,hi) draw itself */
    const float*    noise;     /* [B][32]  randn Re[16], randn Im[16] */
    const float*    tx;        /* [B][32]  time-domain frame Re[16], Im[16]: replaces symbol generation + IFFT altogether
                                  (NonLinearImpairments.apply_* / ChannelModel.apply on caller-supplied signals) */
    const float*    fade;      /* [B][8]   fading draws in the reference's order: rayleigh {randn, randn}; rician {uniform(0, 2 pi),
                                  randn, randn}; multipath {randn, randn} per tap */
    const float*    tx_gain;   /* [B]      with tx: the channel sees tx * tx_gain[b] while the clean output stays tx, and the joint
                                  normalisation divides both by max(max|noisy|, max|tx|) - OFDMDataset.__getitem__,
                                  utils/dataset.py:131-147 (clean_iq * normalization_factor through the channel, then
                                  max(|noisy_iq|, |clean_iq|)) */
} ofdmgan_chan_rand;

/* per-SNR-bin, per-method accumulator row produced by ofdmgan_sim_gen_metrics (doubles):
 *   [0] n frames  [1] sum mse  [2] sum mse^2  [3] sum evm_dB  [4] sum evm_dB^2  [5] bit errors  [6] bits  [7] sum |err|^2/|ref|^2 */
#define OFDMGAN_METRIC_COLS 8
enum { OFDMGAN_METHOD_GAN = 0, OFDMGAN_METHOD_NOEQ = 1, OFDMGAN_METHOD_ZF = 2, OFDMGAN_METHOD_MMSE = 3,
       OFDMGAN_N_METHODS = 4 };
/* metrics_dev layout: double [n_snr][OFDMGAN_N_METHODS][OFDMGAN_METRIC_COLS]; the call ADDS into it. */

/* ---- library ------------------------------------------------------------------------------------------ */
int ofdmgan_abi_version(void);
/* number of SMs of the current device (grid sizing is a multiple of it); <0 on error */
int ofdmgan_device_sms(void);
const char* ofdmgan_error_string(int code);

/* ---- kernel (2): fp32 generator ----------------------------------------------------------------------- */
/* replaces MiniGenerator.forward, models/generator.py:180-208.  x,y: [B][2][16] f32 device. */
int ofdmgan_gen_fwd_f32(const float* x_dev, const float* gparams258, float* y_dev, int64_t B, float leaky_slope,
                        void* stream);
/* replaces autograd through MiniGenerator.forward (train.py:295 g_loss.backward()).  dparams258_dev receives
 * sum_b backward(dy_b) (overwritten, deterministic two-stage reduction); dx_dev may be NULL. */
int ofdmgan_gen_bwd_f32(const float* x_dev, const float* gparams258, const float* dy_dev, float* dx_dev,
                        float* dparams258_dev, int64_t B, float leaky_slope, void* stream);

/* ---- kernel (3): Q1.7-weight / Q8.8-activation integer generator --------------------------------------- */
/* replaces rtl/ofdmGAN/generator_mini.v:326-649 (+ weight_rom.v) - the only executable definition of the integer
 * forward.  x,y: [B][2][16] int16 device; ROMs are HOST pointers (static inference weights, weight_rom.v layout).
 * mode: OFDMGAN_GEN_Q_SPEC or OFDMGAN_GEN_Q_RTL.  digest_dev (optional, 2 x uint64, ADDED into) receives the
 * order-independent (sum, xor) digest of the output words used by the full-size parity test. */
int ofdmgan_gen_fwd_q(const int16_t* x_dev, const int8_t* wrom_host, const int16_t* brom_host, int16_t* y_dev,
                      int64_t B, int mode, uint64_t* digest_dev, void* stream);
/* Integer critic: replaces rtl/ofdmGAN/discriminator_mini.v:261-500 (+ weight_rom.v: weights 256..751, biases 32..56).
 * cand, cond: [B][2][16] int16 device; score: [B] int16 device (Q8.8); ROMs are HOST pointers (2048 int8 / 64 int16).
 * mode OFDMGAN_GEN_Q_SPEC = the RTL's arithmetic primitives on the dataflow of models/discriminator.py:112-152;
 * OFDMGAN_GEN_Q_RTL = what the committed RTL computes in steady state (lagging weight reads, counters that are not reset
 * between states: only conv2 channel 15 reaches the pool). */
int ofdmgan_disc_fwd_q(const int16_t* cand_dev, const int16_t* cond_dev, const int8_t* wrom_host,
                       const int16_t* brom_host, int16_t* score_dev, int64_t B, int mode, void* stream);
/* replace compute_scale / quantize_tensor / dequantize_tensor, utils/quantization.py:73-161, on device tensors laid out
 * [C][inner] with one scale per leading index (C = 1: per tensor): scale = max(amax|x|, 1e-8) / (2^(n-1) - 1);
 * q = clamp(round_half_even(x / scale), -2^(n-1), 2^(n-1) - 1) as float; x = q * scale. */
int ofdmgan_compute_scale(const float* x_dev, int64_t C, int64_t inner, int n_bits, float* scale_dev, void* stream);
int ofdmgan_quantize_tensor(const float* x_dev, int64_t C, int64_t inner, const float* scale_dev, int n_bits, float* q_dev,
                            void* stream);
int ofdmgan_dequantize_tensor(const float* q_dev, int64_t C, int64_t inner, const float* scale_dev, float* x_dev, void* stream);
/* float -> Q8.8 by truncation toward zero, (x*256).astype(int16): proof/verification.py:297-298 */
int ofdmgan_quantize_q88(const float* x_dev, int16_t* q_dev, int64_t n, void* stream);
int ofdmgan_dequantize_q88(const int16_t* q_dev, float* x_dev, int64_t n, void* stream);

/* ---- kernel (1): channel simulator -------------------------------------------------------------------- */
/* replaces SyntheticOFDMDataset.__getitem__ (utils/dataset.py:236-293) for frames frame0..frame0+B-1.
 * clean/noisy: [B][2][16] f32 device, snr: [B] f32 device (any may be NULL).  rand may be NULL (all Philox). */
int ofdmgan_chan_sim(const ofdmgan_chan_cfg* cfg_host, const ofdmgan_chan_rand* rand_host, uint64_t seed,
                     uint64_t frame0, float* clean_dev, float* noisy_dev, float* snr_dev, int64_t B, void* stream);
/* the raw draws kernel (1) consumes for those frames (test hook: lets the oracle replay exactly the same
 * randomness).  sym/noise [B][32], pn [B][16], snr_db [B], bits [B]; any may be NULL. */
int ofdmgan_chan_draws(const ofdmgan_chan_cfg* cfg_host, uint64_t seed, uint64_t frame0, float* sym_dev,
                       uint32_t* bits_dev, float* pn_dev, float* snr_db_dev, float* noise_dev, int64_t B, void* stream);
/* the fading draws of those frames in the ofdmgan_chan_rand.fade layout, [B][8] (ChannelModel.apply's channel_response is
 * formed from them) */
int ofdmgan_chan_fade_draws(const ofdmgan_chan_cfg* cfg_host, uint64_t seed, uint64_t frame0, float* fade_dev, int64_t B, void* stream);
/* Philox4x32-10 block (test hook): out[n][4] = philox(key=seed, ctr=(c0[i] lo/hi, c2, c3)) */
int ofdmgan_philox_blocks(uint64_t seed, uint64_t ctr0, uint32_t c2, uint32_t c3, uint32_t* out_dev, int64_t n,
                          void* stream);

/* ---- modulator API (utils/ofdm_utils.py QAMModulator / OFDMModulator as batched entry points) ------------------------ */
/* complex values are interleaved (re, im) float pairs = torch.complex64 / numpy complex64 memory layout. */
/* replaces QAMModulator(QPSK | QAM16 | QAM64).modulate / .demodulate, utils/ofdm_utils.py:137-222: bits_per_symbol in {2, 4, 6};
 * constellation index = bits MSB first; QAM16/64 point = (levels[idx % sqrtM] + j levels[idx / sqrtM]) / norm (np.meshgrid order);
 * hard decisions pick the nearest point, ties to the lowest index. */
int ofdmgan_qam_modulate(const uint8_t* bits_dev, float* sym_dev, int64_t n_symbols, int bits_per_symbol, void* stream);
/* np.unpackbits / np.packbits (MSB first): the pixel-byte <-> bit step of ImageOFDMConverter, utils/ofdm_utils.py:910,986.
 * bits are one 0/1 value per byte, 8 * n_bytes of them. */
int ofdmgan_unpackbits(const uint8_t* bytes_dev, int64_t n_bytes, uint8_t* bits_dev, void* stream);
int ofdmgan_packbits(const uint8_t* bits_dev, int64_t n_bytes, uint8_t* bytes_dev, void* stream);
int ofdmgan_qam_demodulate(const float* sym_dev, uint8_t* bits_dev, int64_t n_symbols, int bits_per_symbol, void* stream);
/* replaces QAMModulator('QPSK').modulate, utils/ofdm_utils.py:163-193: bits_dev[2n] (0/1 bytes, MSB first) -> sym_dev[n] */
int ofdmgan_qpsk_modulate(const uint8_t* bits_dev, float* sym_dev, int64_t n_symbols, void* stream);
/* replaces QAMModulator('QPSK').demodulate, :195-222: nearest constellation point, ties to the lowest index */
int ofdmgan_qpsk_demodulate(const float* sym_dev, uint8_t* bits_dev, int64_t n_symbols, void* stream);
/* replaces OFDMModulator.modulate, :281-329: n_symbols QAM symbols -> n_ofdm = ceil(n_symbols / n_data) OFDM symbols of
 * n_fft + cp_len samples (data on the non-pilot bins in order, zero padded; pilots at arange(0, n_fft, pilot_spacing);
 * ifft * n_fft; cyclic prefix).  n_fft in {8, 16}; 0 <= cp_len <= n_fft; pilot_spacing 0 = no pilots.  out_dev: n_ofdm*(n_fft+cp_len)
 * samples - except that cp_len == 0 behaves as the reference does (its slice [-0:] is the whole symbol): every symbol is emitted
 * twice, n_ofdm * 2 * n_fft samples. */
int ofdmgan_ofdm_modulate(const float* sym_dev, int64_t n_symbols, int n_fft, int cp_len, int pilot_spacing, float pilot_re,
                          float pilot_im, float* out_dev, void* stream);
/* replaces OFDMModulator.demodulate, :331-371: n_ofdm symbols of n_fft+cp_len samples -> data_dev[n_ofdm*n_data] (fft / n_fft on
 * the data bins) and chan_dev[n_ofdm*n_pilots] = pilot bins / pilot_value (may be NULL) */
int ofdmgan_ofdm_demodulate(const float* sig_dev, int64_t n_ofdm, int n_fft, int cp_len, int pilot_spacing, float pilot_re,
                            float pilot_im, float* data_dev, float* chan_dev, void* stream);

/* ---- fused (1)+(2|3)+metrics: the headline path -------------------------------------------------------- */
/* replaces the inner loops of run_benchmark (benchmark_comparison.py:179-250) and of
 * DataLoader(SyntheticOFDMDataset) -> MiniGenerator (train.py:327-329,394): simulate frames frame0..frame0+B-1
 * on chip, reconstruct them with the generator selected by gen_kind, and ADD per-SNR-bin metric rows into
 * metrics_dev.  gen_params: 258 floats (host or device) for OFDMGAN_GEN_F32; for the Q kinds pass wrom_host /
 * brom_host instead.  Nothing but the accumulator touches HBM. */
int ofdmgan_sim_gen_metrics(const ofdmgan_chan_cfg* cfg_host, int gen_kind, const float* gparams258,
                            const int8_t* wrom_host, const int16_t* brom_host, float leaky_slope, uint64_t seed,
                            uint64_t frame0, int64_t B, double* metrics_dev, void* stream);
/* same, host buffers in and out (cfg, weights in; metrics table out, ADDED into metrics_host); synchronous. */
int ofdmgan_sim_gen_metrics_host(const ofdmgan_chan_cfg* cfg_host, int gen_kind, const float* gparams258_host,
                                 const int8_t* wrom_host, const int16_t* brom_host, float leaky_slope,
                                 uint64_t seed, uint64_t frame0, int64_t B, double* metrics_host, void* stream);
/* Which simulator kernel a configuration runs on: 1 = the lean headline kernel (csrc/sim_lean.cu: Gaussian source, AWGN channel,
 * Rapp / IQ / phase noise, fp32 generator or none), 0 = the general kernel (every other option of ofdmgan_chan_cfg).  gen_kind -1 =
 * simulate only.  has_tx_or_fade != 0: the call injects time-domain frames or fading draws.  The environment variable
 * OFDMGAN_SIM_IMPL=general forces 0 (A/B runs, cross-check tests).  Pure host function. */
int ofdmgan_sim_impl_for(const ofdmgan_chan_cfg* cfg_host, int gen_kind, int has_tx_or_fade);
/* metric rows for frames that already exist in HBM (benchmark_comparison.py:137-146,205-214).
 * est/ref: [B][2][16] f32 device; bin_dev: [B] int32 SNR bin per frame or NULL (all bin 0). */
int ofdmgan_frame_metrics(const float* est_dev, const float* ref_dev, const int32_t* bin_dev, int method, int n_snr,
                          int64_t B, double* metrics_dev, void* stream);

/* replaces ZeroForcingEqualizer.equalize_iq / MMSEEqualizer.equalize_iq with the genie channel estimate (clean given),
 * utils/classical_equalizers.py:88-126,203-230.  method: OFDMGAN_METHOD_ZF | OFDMGAN_METHOD_MMSE.  snr_db_dev: [B] per-frame SNR
 * for MMSE (NULL: the reference's default 20 dB).  est_dev: [B][2][16] equalised frames. */
int ofdmgan_equalize(const float* noisy_dev, const float* clean_dev, const float* snr_db_dev, int method, float* est_dev, int64_t B,
                     void* stream);

/* ---- kernel (4): critic ------------------------------------------------------------------------------- */
/* replaces MiniDiscriminator.forward, models/discriminator.py:112-152.  score: [B] f32 device. */
int ofdmgan_disc_fwd_f32(const float* cand_dev, const float* cond_dev, const float* dparams521, float* score_dev,
                         int64_t B, float leaky_slope, void* stream);
/* replaces autograd through MiniDiscriminator.forward: given upstream g[B] = dL/dscore, writes
 * dcand/dcond [B][2][16] (each may be NULL) and dparams521_dev = sum_b g_b dD_b/dtheta (may be NULL). */
int ofdmgan_disc_bwd_f32(const float* cand_dev, const float* cond_dev, const float* dparams521, const float* g_dev,
                         float* dcand_dev, float* dcond_dev, float* dparams521_dev, int64_t B, float leaky_slope,
                         void* stream);
/* replaces compute_gradient_penalty, models/discriminator.py:172-236 (+ its backward w.r.t. the critic's
 * parameters).  alpha_dev: [B] the torch.rand(B,1,1) draw (NULL: Philox(seed, sample index, alpha_iter)).
 * gp_dev: 1 float = mean (||grad||-1)^2.  dparams521_dev (may be NULL) = d gp / d theta. */
int ofdmgan_gradient_penalty(const float* real_dev, const float* fake_dev, const float* cond_dev,
                             const float* alpha_dev, uint64_t seed, uint64_t sample0, uint32_t alpha_iter,
                             const float* dparams521, float* gp_dev, float* dparams521_dev, int64_t B,
                             float leaky_slope, void* stream);
/* replaces the loss + backward of CWGANGPTrainer.train_discriminator, train.py:228-250 (everything between
 * zero_grad and optimizer_D.step()).  fake_dev = G(noisy) computed by the caller once per batch (train.py:225-226).
 * out_dev: 528 floats = grad[521] of d_loss w.r.t. theta_D summed over the LOCAL batch and scaled by 1/B_global,
 * then stats[5] = partial sums (already scaled by 1/B_global) of d_loss, wasserstein_distance, gradient_penalty,
 * d_real_mean, d_fake_mean (train.py:255-261), then 2 pad floats.  A data-parallel caller allreduces (sum) the
 * 528 floats and feeds them to ofdmgan_adam. */
#define OFDMGAN_CRITIC_OUT 528
int ofdmgan_critic_step(const float* clean_dev, const float* noisy_dev, const float* fake_dev, const float* alpha_dev,
                        uint64_t seed, uint64_t sample0, uint32_t alpha_iter, const float* dparams521,
                        float gp_weight, float leaky_slope, int64_t B_local, int64_t B_global, float* out_dev,
                        void* stream);
/* The same with the alpha counter read from device memory (*alpha_iter_dev) at run time instead of being baked into the launch:
 * together with ofdmgan_adam_ctr it makes the whole iteration of train.py:327-344 replayable as ONE CUDA graph
 * (no per-step host scalars).  Always draws alpha from Philox. */
int ofdmgan_critic_step_ctr(const float* clean_dev, const float* noisy_dev, const float* fake_dev, uint64_t seed,
                            uint64_t sample0, const int32_t* alpha_iter_dev, const float* dparams521, float gp_weight,
                            float leaky_slope, int64_t B_local, int64_t B_global, float* out_dev, void* stream);
/* One WHOLE critic iteration of train.py:201-261 (loss, backward, optimizer_D.step()) in two launches: the kernel of
 * ofdmgan_critic_step_ctr, then ONE tail launch that reduces the gradients, sums them over the data-parallel ranks through `comm`
 * (NULL: single GPU), applies Adam to dparams521_dev / m_dev / v_dev in place (step count t = *step_dev + 1, stored back;
 * *step_dev before the call is also the Philox alpha counter) and refreshes the kernel's weight image.  out_dev as for
 * ofdmgan_critic_step (global sums when comm is given).  image_is_current != 0: the previous call on this stream was this
 * function on the same parameters (its tail already installed their image), so the image refresh at entry is skipped; pass 0
 * whenever anything else may have written the parameters. */
typedef struct ofdmgan_comm ofdmgan_comm;
int ofdmgan_critic_train_ctr(const float* clean_dev, const float* noisy_dev, const float* fake_dev, uint64_t seed,
                             uint64_t sample0, int32_t* step_dev, float* dparams521_dev, float* m_dev, float* v_dev, double lr,
                             double beta1, double beta2, double eps, float gp_weight, float leaky_slope, int64_t B_local,
                             int64_t B_global, float* out_dev, int image_is_current, ofdmgan_comm* comm, void* stream);
/* replaces the loss + backward of CWGANGPTrainer.train_generator, train.py:285-298.
 * out_dev: 264 floats = grad[258] of g_loss w.r.t. theta_G (local sum, scaled for the global batch), stats[3] =
 * g_loss, adv_loss, rec_loss partial sums, 3 pad.  fake_out_dev (may be NULL) receives G(noisy). */
#define OFDMGAN_GEN_OUT 264
int ofdmgan_gen_step(const float* clean_dev, const float* noisy_dev, const float* dparams521, const float* gparams258,
                     float adv_weight, float rec_weight, float leaky_slope, int64_t B_local, int64_t B_global,
                     float* out_dev, float* fake_out_dev, void* stream);
/* The same step when G(noisy) already exists: fake_dev = the output of ofdmgan_gen_fwd_f32(noisy_dev, gparams258) - a training
 * iteration computes it once for its critic updates (train.py:228-232) and train.py:285 is the same forward with the same
 * generator, so the step skips its first forward pass.  The backward pass still recomputes the forward with its tape from
 * gparams258: fake_dev only feeds the critic term and the L1 term.  Passing anything but G(noisy) is the caller's error. */
int ofdmgan_gen_step_fake(const float* clean_dev, const float* noisy_dev, const float* fake_dev, const float* dparams521,
                          const float* gparams258, float adv_weight, float rec_weight, float leaky_slope, int64_t B_local,
                          int64_t B_global, float* out_dev, void* stream);
/* One whole generator update of a training iteration in two launches (the generator-side twin of ofdmgan_critic_train_ctr,
 * train.py:282-299): ofdmgan_gen_step_fake's loss + backward, then ONE tail launch = fixed-order reduction of the gradient, sum
 * over the ranks of `comm` through peer memory (comm may be NULL), Adam in place on gparams258_dev / m_dev / v_dev with the step
 * count in *step_dev.  out_dev: OFDMGAN_GEN_OUT floats as ofdmgan_gen_step (global sums when comm is given).
 * d_image_staged / g_image_staged != 0: the caller states that this device's staging buffer already holds the weight image of
 * dparams521_dev (true right after ofdmgan_critic_train_ctr on them) / of gparams258_dev (true right after ofdmgan_gen_fwd_f32 with
 * that device pointer) and nothing else has run on the library in between: the image is copied instead of rebuilt.  Pass 0
 * whenever in doubt. */
int ofdmgan_gen_train_ctr(const float* clean_dev, const float* noisy_dev, const float* fake_dev, int32_t* step_dev,
                          float* gparams258_dev, float* m_dev, float* v_dev, double lr, double beta1, double beta2, double eps,
                          const float* dparams521_dev, float adv_weight, float rec_weight, float leaky_slope, int64_t B_local,
                          int64_t B_global, float* out_dev, int d_image_staged, int g_image_staged, ofdmgan_comm* comm,
                          void* stream);
/* replaces torch.optim.Adam.step for one flat parameter vector (train.py:114-127,253,299): fp32 state,
 * no weight decay, no amsgrad.  g = grad_scale * g_dev[i].  All device pointers; lr/betas/eps are doubles like the
 * python floats the optimizer holds (1-beta is formed in double before narrowing, as ATen does). */
int ofdmgan_adam(float* p_dev, float* m_dev, float* v_dev, const float* g_dev, int n, double lr, double beta1,
                 double beta2, double eps, int step, float grad_scale, void* stream);

/* ofdmgan_adam with the step count in device memory: uses t = *step_dev + 1 for the bias corrections and stores t back.  n <= 1024. */
int ofdmgan_adam_ctr(float* p_dev, float* m_dev, float* v_dev, const float* g_dev, int n, double lr, double beta1,
                     double beta2, double eps, int32_t* step_dev, float grad_scale, void* stream);

/* ---- data-parallel exchange fused with the optimiser ---------------------------------------------------- */
/* replaces, for one-process-per-GPU replicas on one node, the pair  all_reduce(grad buffer) ; optimizer.step()  that ends
 * train.py:253 / train.py:297 under data parallelism: ONE launch stores this rank's n floats into every peer's exchange block
 * over NVLink peer memory, waits for all ranks, sums them in rank order (bit-identical on every rank), writes the total back
 * to g_dev and applies ofdmgan_adam's update to the first n_params entries (n_params may be 0: plain all-reduce).
 * Setup: every rank calls ofdmgan_comm_create (allocates its exchange block, returns a 64-byte CUDA IPC handle), the host
 * side all-gathers the handles (torch.distributed), every rank calls ofdmgan_comm_connect with the world x 64 bytes.
 * All ranks must issue the same sequence of ofdmgan_allreduce_adam calls.  n <= 1024. */
int ofdmgan_comm_create(int rank, int world, ofdmgan_comm** out, void* ipc_handle64);
int ofdmgan_comm_connect(ofdmgan_comm* comm, const void* all_handles);
int ofdmgan_comm_destroy(ofdmgan_comm* comm);
/* synchronises; non-zero once a wait has run out: a peer that does not arrive within the wait limit (environment variable
 * OFDMGAN_COMM_TIMEOUT_S at ofdmgan_comm_create, default 120 s) makes the waiting kernel set the communicator's error word and TRAP, so
 * that launch and every later call on the context fail with a CUDA error - no rank continues on a partial sum, replicas cannot
 * diverge silently.  Put a barrier after rank-asymmetric work (rank-0 checkpointing, validation) that may exceed the limit. */
int ofdmgan_comm_check(ofdmgan_comm* comm, void* stream);
int ofdmgan_allreduce_adam(ofdmgan_comm* comm, float* g_dev, int n, float* p_dev, float* m_dev, float* v_dev, int n_params,
                           double lr, double beta1, double beta2, double eps, int step, float grad_scale, void* stream);

/* the same with Adam's step count in device memory (t = *step_dev + 1, stored back): graph-replayable, like ofdmgan_adam_ctr */
int ofdmgan_allreduce_adam_ctr(ofdmgan_comm* comm, float* g_dev, int n, float* p_dev, float* m_dev, float* v_dev, int n_params,
                               double lr, double beta1, double beta2, double eps, int32_t* step_dev, float grad_scale,
                               void* stream);

/* ---- measurement helper -------------------------------------------------------------------------------- */
/* FP32 FFMA issue-rate microbenchmark (denominator of the fp32 roofline): runs `iters`
 * dependent-chain FFMA loops on every SM, returns achieved TFLOP/s in *tflops_host.  Synchronous. */
int ofdmgan_ffma_peak(int iters, double* tflops_host, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* OFDMGAN_H */
