/* TEST INFRASTRUCTURE (oracle) - CPU restatement of the reference's Q1.7-weight / Q8.8-activation generator.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may call this.
 * The product path (ofdm-gan-sr_b200/csrc) never links or loads it.
 *
 * The reference has no integer forward pass in Python (utils/quantization.py only quantises weights); the
 * only executable definition is the Verilog:
 *   rtl/ofdmGAN/generator_mini.v:141-146   per-tap product, `>>> 7` (floor) BEFORE the tap sum
 *   rtl/ofdmGAN/generator_mini.v:352-362   + sign-extended bias, saturate to int16, LeakyReLU (r>>>2)+(r>>>4)
 *   rtl/ofdmGAN/generator_mini.v:534-546   skip add with saturation
 *   rtl/ofdmGAN/generator_mini.v:457-458,561-562  nearest-neighbour x2
 *   rtl/ofdmGAN/generator_mini.v:576-590   1x1 output convolution, saturation, no activation
 *   rtl/ofdmGAN/generator_mini.v:630-636   "tanh" = hard clip: v>256 -> 255, v<-256 -> -255
 *   rtl/ofdmGAN/generator_mini.v:70-79     ROM bases: weights enc1 0, bneck 24, dec1 120, out 216 ([oc][ic][k]);
 *                                          biases 0 / 4 / 12 / 16
 *
 * mode 0 ("spec"): those primitives on the textbook U-Net dataflow of models/generator.py:180-208
 *   (all channels, aligned weights, clip on both output channels).  No reference fixture pins this mode.
 * mode 1 ("rtl_literal"): what the committed RTL computes, including its scheduling artefacts (weights skewed by
 *   one loop iteration because of the synchronous ROM read, counters not reset between states).  Closed form per
 *   SURVEY.md Appendix C; PINNED bit-for-bit by the 10 known-answer frames of rtl/ofdmGAN/tb_generator_mini.vcd
 *   (tests/golden/rtl_generator_vectors.json) and by oracle/rtl_cycle_emulator.py on random ROMs and frames.
 */
#include <stdint.h>
#include <string.h>

static inline int32_t sat16(int32_t v) { return v > 32767 ? 32767 : (v < -32768 ? -32768 : v); }

static inline int32_t lrelu_q(int32_t r) {
    if (r >= 0) return r;
    return (int32_t)(int16_t)((r >> 2) + (r >> 4)); /* 16-bit wrap as in the RTL (cannot actually wrap) */
}

static inline int32_t tap(int32_t a, int32_t w) { return (a * w) >> 7; } /* arithmetic shift == floor */

/* ------------------------------------------------------------------ mode 0: spec ------------------------- */
static void gen_q_spec(const int16_t* x, const int8_t* W, const int16_t* Bq, int16_t* y) {
    int32_t xp[2][18], enc[4][10], bn[8][4], up1[8][10], dec[4][8], out[2][16];
    memset(xp, 0, sizeof xp); memset(enc, 0, sizeof enc); memset(up1, 0, sizeof up1);
    for (int c = 0; c < 2; ++c) for (int i = 0; i < 16; ++i) xp[c][i + 1] = x[c * 16 + i];
    for (int oc = 0; oc < 4; ++oc) for (int p = 0; p < 8; ++p) {            /* enc1: 2->4, k3, s2 */
        int32_t acc = 0;
        for (int ic = 0; ic < 2; ++ic) for (int k = 0; k < 3; ++k)
            acc += tap(xp[ic][2 * p + k], W[0 + (oc * 2 + ic) * 3 + k]);
        enc[oc][p + 1] = lrelu_q(sat16(acc + Bq[0 + oc]));
    }
    for (int oc = 0; oc < 8; ++oc) for (int p = 0; p < 4; ++p) {            /* bottleneck: 4->8, k3, s2 */
        int32_t acc = 0;
        for (int ic = 0; ic < 4; ++ic) for (int k = 0; k < 3; ++k)
            acc += tap(enc[ic][2 * p + k], W[24 + (oc * 4 + ic) * 3 + k]);
        bn[oc][p] = lrelu_q(sat16(acc + Bq[4 + oc]));
    }
    for (int c = 0; c < 8; ++c) for (int p = 0; p < 4; ++p) up1[c][1 + 2 * p] = up1[c][2 + 2 * p] = bn[c][p];
    for (int oc = 0; oc < 4; ++oc) for (int p = 0; p < 8; ++p) {            /* dec1: 8->4, k3, s1 */
        int32_t acc = 0;
        for (int ic = 0; ic < 8; ++ic) for (int k = 0; k < 3; ++k)
            acc += tap(up1[ic][p + k], W[120 + (oc * 8 + ic) * 3 + k]);
        dec[oc][p] = sat16(lrelu_q(sat16(acc + Bq[12 + oc])) + enc[oc][p + 1]);   /* + skip, saturating */
    }
    for (int oc = 0; oc < 2; ++oc) for (int q = 0; q < 16; ++q) {           /* out: 4->2, 1x1 on the x2 upsample */
        int32_t acc = 0;
        for (int ic = 0; ic < 4; ++ic) acc += tap(dec[ic][q >> 1], W[216 + oc * 4 + ic]);
        int32_t v = sat16(acc + Bq[16 + oc]);
        out[oc][q] = v > 256 ? 255 : (v < -256 ? -255 : v);
    }
    for (int c = 0; c < 2; ++c) for (int q = 0; q < 16; ++q) y[c * 16 + q] = (int16_t)out[c][q];
}

/* ------------------------------------------------------------------ mode 1: rtl_literal ------------------ */
/* One RTL conv state.  src is [IN][srcw] (zero padded by the caller where the RTL buffer is padded).
 * The loop iteration (oc,op,ic) multiplies its data window by the weight triple addressed by the PREVIOUS
 * iteration (synchronous ROM, weight_rom.v:164-166 vs generator_mini.v:328-346). */
static void skewed_conv(const int32_t* src, int srcw, int IN, int OC, int OL, int stride, int WA, int BA,
                        int oc_first, int stale, int K, int act, const int8_t* W, const int16_t* Bq,
                        int32_t* out /* [OC][OL] */) {
    for (int oc = oc_first; oc < OC; ++oc)
        for (int op = 0; op < OL; ++op) {
            int32_t acc = 0;
            for (int ic = 0; ic < IN; ++ic) {
                int a;
                if (ic > 0) a = WA + oc * IN * K + (ic - 1) * K;
                else if (op > 0 || oc == OC - 1) a = WA + oc * IN * K + (IN - 1) * K;
                else if (oc > oc_first) a = WA + (oc - 1) * IN * K + (IN - 1) * K;
                else a = stale;
                for (int k = 0; k < K; ++k) acc += tap(src[ic * srcw + op * stride + k], W[(a + k) & 2047]);
            }
            int32_t v = sat16(acc + Bq[BA + oc]);
            out[oc * OL + op] = act ? lrelu_q(v) : v;
        }
}

static void gen_q_rtl(const int16_t* x, const int8_t* W, const int16_t* Bq, int16_t* y, int stale_enc1) {
    int32_t xp[2][18], enc[4][8], encp[4][10], bn[8][4], up1[8][10], dec[4][8], up2[4][16], out[2][16];
    memset(xp, 0, sizeof xp); memset(encp, 0, sizeof encp); memset(bn, 0, sizeof bn); memset(up1, 0, sizeof up1);
    for (int c = 0; c < 2; ++c) for (int i = 0; i < 16; ++i) xp[c][i + 1] = x[c * 16 + i];
    skewed_conv(&xp[0][0], 18, 2, 4, 8, 2, 0, 0, 0, stale_enc1, 3, 1, W, Bq, &enc[0][0]);
    for (int c = 0; c < 4; ++c) for (int p = 0; p < 8; ++p) encp[c][p + 1] = enc[c][p];
    skewed_conv(&encp[0][0], 10, 4, 8, 4, 2, 24, 4, 3, 21, 3, 1, W, Bq, &bn[0][0]);   /* only out-channels 3..7 */
    for (int p = 0; p < 4; ++p) up1[7][1 + 2 * p] = up1[7][2 + 2 * p] = bn[7][p];       /* only channel 7 */
    skewed_conv(&up1[0][0], 10, 8, 4, 8, 1, 120, 12, 0, 117, 3, 1, W, Bq, &dec[0][0]);
    for (int p = 0; p < 8; ++p) dec[3][p] = sat16(dec[3][p] + enc[3][p]);               /* skip add only ch 3 */
    for (int c = 0; c < 4; ++c) for (int q = 0; q < 16; ++q) up2[c][q] = dec[c][q >> 1];
    skewed_conv(&up2[0][0], 16, 4, 2, 16, 1, 216, 16, 0, 213, 1, 0, W, Bq, &out[0][0]);
    for (int q = 0; q < 16; ++q) {                                                      /* clip only channel 1 */
        int32_t v = out[1][q];
        out[1][q] = v > 256 ? 255 : (v < -256 ? -255 : v);
    }
    for (int c = 0; c < 2; ++c) for (int q = 0; q < 16; ++q) y[c * 16 + q] = (int16_t)out[c][q];
}

/* x: [B][2][16] int16 Q8.8; W: 2048 int8 Q1.7 (weight_rom.v layout); Bq: 64 int16 Q8.8; y: [B][2][16].
 * mode 0 spec, 1 rtl_literal steady state, 2 rtl_literal first frame after reset (enc1 stale address 0). */
int oracle_gen_fwd_q(const int16_t* x, const int8_t* W, const int16_t* Bq, int16_t* y, int64_t B, int mode) {
    if (mode < 0 || mode > 2) return -1;
#pragma omp parallel for schedule(static)
    for (int64_t b = 0; b < B; ++b) {
        if (mode == 0) gen_q_spec(x + b * 32, W, Bq, y + b * 32);
        else gen_q_rtl(x + b * 32, W, Bq, y + b * 32, mode == 1 ? 223 : 0);
    }
    return 0;
}

/* Order-independent digest of an int16 output buffer (sum and xor-fold of position-salted words), used by the
 * full-size (2^24 frame) parity test: the CUDA kernel computes the same digest on the device. */
void oracle_digest_i16(const int16_t* y, int64_t n, uint64_t* sum_out, uint64_t* xor_out) {
    uint64_t s = 0, xr = 0;
#pragma omp parallel for reduction(+ : s) reduction(^ : xr) schedule(static)
    for (int64_t i = 0; i < n; ++i) {
        uint64_t v = (uint16_t)y[i];
        uint64_t h = (v + 0x9E3779B97F4A7C15ull * (uint64_t)(i + 1));
        h ^= h >> 29; h *= 0xBF58476D1CE4E5B9ull; h ^= h >> 32;
        s += h; xr ^= h;
    }
    *sum_out = s; *xor_out = xr;
}

/* ================================================================== fixed-point critic ===================
 * rtl/ofdmGAN/discriminator_mini.v (the only executable definition; utils/quantization.py has no integer forward):
 *   :66-74    ROM bases: weights conv1 256 ([8][4][3]), conv2 352 ([16][8][3]), dense 736 ([16]); biases 32 / 40 / 56
 *   :126-134  three parallel taps, each `(a*w) >>> 7` before the sum        :284-291 / :365-367  stride 2, zero pad 1
 *   :311-321  + sign-extended bias, saturate to int16, LeakyReLU (r>>>2)+(r>>>4)   (same at :388-398 for conv2)
 *   :418-432  sum pool over the 4 positions into 32 bits; :447 the dense stage reads pool[15:0] (16-bit wrap)
 *   :458-466  dense taps `(pool*w) >>> 7` + bias; :488-495 saturate the accumulator into the int16 score
 * mode 0 ("spec"): those primitives on the dataflow of models/discriminator.py:112-152 (all channels, aligned weights).
 *   No reference fixture pins this mode.
 * mode 1 ("rtl_literal", steady state): what the committed RTL computes.  Weights lag the data by one loop iteration
 *   (synchronous ROM read); out_ch_cnt is not reset between states, so CONV2 computes channels 7..15 only, POOL sums
 *   channel 15 only, and the dense stage therefore sees one non-zero input, multiplied by the lagging weight ROM[750].
 *   The very first conv1 iteration multiplies by the address left by the previous frame's dense stage (ROM[751..753]);
 *   mode 2 is the first frame after reset (that address is 0).  PINNED by the five runs of
 *   rtl/ofdmGAN/tb_discriminator_mini.vcd (tests/golden/rtl_critic_vectors.json: scores and all 846 accumulate-stage
 *   entries per run, through oracle/rtl_cycle_emulator.py) and by that emulator on random ROMs and frames. */
static void load_critic_input(const int16_t* cand, const int16_t* cond, int32_t x[4][18]) {
    memset(x, 0, sizeof(int32_t) * 4 * 18);
    for (int c = 0; c < 2; ++c)
        for (int p = 0; p < 16; ++p) { x[c][p + 1] = cand[c * 16 + p]; x[c + 2][p + 1] = cond[c * 16 + p]; }
}

static int16_t critic_q_spec(const int16_t* cand, const int16_t* cond, const int8_t* W, const int16_t* Bq) {
    int32_t x[4][18], c1[8][10];
    load_critic_input(cand, cond, x);
    memset(c1, 0, sizeof c1);
    for (int oc = 0; oc < 8; ++oc)
        for (int op = 0; op < 8; ++op) {
            int32_t acc = 0;
            for (int ic = 0; ic < 4; ++ic)
                for (int k = 0; k < 3; ++k) acc += tap(x[ic][2 * op + k], W[256 + (oc * 4 + ic) * 3 + k]);
            c1[oc][op + 1] = lrelu_q(sat16(acc + Bq[32 + oc]));
        }
    int32_t dense = 0;
    for (int oc = 0; oc < 16; ++oc) {
        int32_t pool = 0;
        for (int op = 0; op < 4; ++op) {
            int32_t acc = 0;
            for (int ic = 0; ic < 8; ++ic)
                for (int k = 0; k < 3; ++k) acc += tap(c1[ic][2 * op + k], W[352 + (oc * 8 + ic) * 3 + k]);
            pool += lrelu_q(sat16(acc + Bq[40 + oc]));
        }
        dense += tap((int16_t)pool, W[736 + oc]);
    }
    return (int16_t)sat16(dense + Bq[56]);
}

static int16_t critic_q_rtl(const int16_t* cand, const int16_t* cond, const int8_t* W, const int16_t* Bq, int stale) {
    int32_t x[4][18], c1[8][10];
    load_critic_input(cand, cond, x);
    memset(c1, 0, sizeof c1);
    for (int oc = 0; oc < 8; ++oc)
        for (int op = 0; op < 8; ++op) {
            int32_t acc = 0;
            for (int it = 0; it < 4; ++it) {
                int a;                                   /* the weight triple addressed by the previous iteration */
                if (it > 0) a = 256 + oc * 12 + (it - 1) * 3;
                else if (op > 0 || oc == 7) a = 256 + oc * 12 + 9;     /* the last channel is recomputed by the flush passes */
                else if (oc > 0) a = 256 + (oc - 1) * 12 + 9;
                else a = stale;
                for (int k = 0; k < 3; ++k) acc += tap(x[it][2 * op + k], W[(a + k) & 2047]);
            }
            c1[oc][op + 1] = lrelu_q(sat16(acc + Bq[32 + oc]));
        }
    int32_t pool = 0;
    for (int op = 0; op < 4; ++op) {
        int32_t acc = 0;
        for (int it = 0; it < 8; ++it) {
            const int a = 352 + 15 * 24 + (it > 0 ? (it - 1) * 3 : 21);
            for (int k = 0; k < 3; ++k) acc += tap(c1[it][2 * op + k], W[a + k]);
        }
        pool += lrelu_q(sat16(acc + Bq[40 + 15]));
    }
    return (int16_t)sat16(tap((int16_t)pool, W[750]) + Bq[56]);
}

/* cand, cond: [B][2][16] int16 Q8.8; score: [B] int16 Q8.8.  mode as above. */
int oracle_disc_fwd_q(const int16_t* cand, const int16_t* cond, const int8_t* W, const int16_t* Bq, int16_t* score,
                      int64_t B, int mode) {
    if (mode < 0 || mode > 2) return -1;
#pragma omp parallel for schedule(static)
    for (int64_t b = 0; b < B; ++b)
        score[b] = mode == 0 ? critic_q_spec(cand + b * 32, cond + b * 32, W, Bq)
                             : critic_q_rtl(cand + b * 32, cond + b * 32, W, Bq, mode == 1 ? 751 : 0);
    return 0;
}
