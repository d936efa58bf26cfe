// libofdmgan: Q1.7-weight / Q8.8-activation integer critic (sm_100a), ofdmgan_disc_fwd_q.
// Replaces rtl/ofdmGAN/discriminator_mini.v:261-500 (+ weight_rom.v), the only executable definition of the integer critic.
// One (candidate, condition) pair per thread, all activations in registers as exact integers held in floats; every tap is
// one round-toward-minus-infinity FMA against the 1.5*2^23 accumulator (gen_device.cuh: fma_rd == acc + ((a*w) >>> 7)),
// two output channels per FFMA2.RM in spec mode.  Frames arrive through the warp-private swizzled tiles of io_tile.cuh
// (coalesced 16-byte loads); the 2-byte scores are stored lane-contiguously.
#include "common.cuh"
#include "io_tile.cuh"
#include "weights.cuh"
#include "gen_device.cuh"

namespace og {

// constant image of the critic part of the ROMs (floats; weights as k/128, exact)
constexpr int DQ_RAW = 0;        // ROM[256 .. 753]   index = address - 256 (conv1 0, conv2 96, dense 480, 3 spill-over words)
constexpr int DQ_BIAS = 512;     // bias ROM[32 .. 56] index = address - 32  (conv1 0, conv2 8, dense 24)
constexpr int DQ2_C1 = 544;      // conv1 channel pairs  [((o2*4 + ic)*3 + k)*2 + h]
constexpr int DQ2_C2 = 640;      // conv2 channel pairs  [((o2*8 + ic)*3 + k)*2 + h]
constexpr int DQ_BIASM = 1024;   // 1.5*2^23 + bias (exact): accumulator start values, same indexing as DQ_BIAS
constexpr int OG_DQ_IMG = 1056;
static __constant__ __align__(16) float c_dq[OG_DQ_IMG];

static int upload_dq(const int8_t* wrom_host, const int16_t* brom_host, cudaStream_t s) {
    float img[OG_DQ_IMG];
    for (int i = 0; i < OG_DQ_IMG; ++i) img[i] = 0.f;
    for (int a = 256; a < 754; ++a) img[DQ_RAW + a - 256] = (float)wrom_host[a] * (1.0f / 128.0f);
    for (int a = 32; a <= 56; ++a) img[DQ_BIAS + a - 32] = (float)brom_host[a];
    for (int a = 32; a <= 56; ++a) img[DQ_BIASM + a - 32] = 12582912.0f + (float)brom_host[a];
    auto pairs = [&](int dst, int wa, int OC, int IC) {
        for (int o2 = 0; o2 < OC / 2; ++o2)
            for (int ic = 0; ic < IC; ++ic)
                for (int k = 0; k < 3; ++k)
                    for (int h = 0; h < 2; ++h)
                        img[dst + ((o2 * IC + ic) * 3 + k) * 2 + h] = (float)wrom_host[wa + ((2 * o2 + h) * IC + ic) * 3 + k] * (1.0f / 128.0f);
    };
    pairs(DQ2_C1, 256, 8, 4);
    pairs(DQ2_C2, 352, 16, 8);
    OG_CHECK(cudaMemcpyToSymbolAsync(c_dq, img, sizeof img, 0, cudaMemcpyHostToDevice, s));
    return 0;
}

// pool_buf[oc][15:0] (discriminator_mini.v:447): the 32-bit pooled sum re-read as a 16-bit signed value
__device__ __forceinline__ float fx_wrap16(float v) { return (float)(short)__float2int_rn(v); }

// mode spec: the RTL primitives on the dataflow of models/discriminator.py:112-152 (all channels, aligned weights)
__device__ __forceinline__ float critic_q_spec(const float* __restrict__ Q, const float (&x)[4][16]) {
    float c1[8][8];
    const f32x2 magic2 = pk2(FX_MAGIC, FX_MAGIC);
#pragma unroll
    for (int o2 = 0; o2 < 4; ++o2)
#pragma unroll
        for (int p = 0; p < 8; ++p) {
            f32x2 acc = ldc2(Q + DQ_BIASM + 2 * o2);
#pragma unroll
            for (int ic = 0; ic < 4; ++ic)
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    const int i = 2 * p + k - 1;
                    if (i >= 0) acc = fxacc2(x[ic][i], ldc2(Q + DQ2_C1 + ((o2 * 4 + ic) * 3 + k) * 2), acc);
                }
            fx_finish2(acc, true, c1[2 * o2][p], c1[2 * o2 + 1][p]);
        }
    float dense = FX_MAGIC;
    // rolled over the channel pair: the 48 weights of a pair are loaded inside the iteration that uses them
#pragma unroll 1
    for (int o2 = 0; o2 < 8; ++o2) {
        const float* W2 = Q + DQ2_C2 + o2 * 8 * 3 * 2;
        const f32x2 b2 = ldc2(Q + DQ_BIASM + 8 + 2 * o2);
        f32x2 acc[4] = {b2, b2, b2, b2};
#pragma unroll
        for (int ic = 0; ic < 8; ++ic)
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const f32x2 w = ldc2(W2 + (ic * 3 + k) * 2);
#pragma unroll
                for (int p = 0; p < 4; ++p) {
                    const int i = 2 * p + k - 1;
                    if (i >= 0) acc[p] = fxacc2(c1[ic][i], w, acc[p]);
                }
            }
        float pool_lo = 0.f, pool_hi = 0.f;
#pragma unroll
        for (int p = 0; p < 4; ++p) {
            float lo, hi;
            fx_finish2(acc[p], true, lo, hi);
            pool_lo += lo;
            pool_hi += hi;
        }
        dense = fxacc(fx_wrap16(pool_lo), Q[DQ_RAW + 480 + 2 * o2], dense);
        dense = fxacc(fx_wrap16(pool_hi), Q[DQ_RAW + 480 + 2 * o2 + 1], dense);
    }
    return fx_sat16((dense - FX_MAGIC) + Q[DQ_BIAS + 24]);
}

// weight address (minus 256) the RTL multiplies conv1 iteration (oc, op, it) by: the triple of the previous iteration
__device__ __forceinline__ constexpr int c1_skew(int oc, int op, int it) {
    return it > 0 ? oc * 12 + (it - 1) * 3 : ((op > 0 || oc == 7) ? oc * 12 + 9 : (oc > 0 ? (oc - 1) * 12 + 9 : 751 - 256));
}

// mode rtl_literal (steady state): oracle/fixed_point.c critic_q_rtl, discriminator_mini.v as committed
__device__ __forceinline__ float critic_q_rtl(const float* __restrict__ Q, const float (&x)[4][16]) {
    float c1[8][8];
#pragma unroll
    for (int oc = 0; oc < 8; ++oc)
#pragma unroll
        for (int p = 0; p < 8; ++p) {
            float acc = FX_MAGIC;
#pragma unroll
            for (int it = 0; it < 4; ++it)
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    const int i = 2 * p + k - 1;
                    if (i >= 0) acc = fxacc(x[it][i], Q[DQ_RAW + c1_skew(oc, p, it) + k], acc);
                }
            c1[oc][p] = fx_finish(acc, Q[DQ_BIAS + oc], true);
        }
    float pool = 0.f;
#pragma unroll
    for (int p = 0; p < 4; ++p) {
        float acc = FX_MAGIC;
#pragma unroll
        for (int it = 0; it < 8; ++it)
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const int i = 2 * p + k - 1;
                if (i >= 0) acc = fxacc(c1[it][i], Q[DQ_RAW + 96 + 15 * 24 + (it > 0 ? (it - 1) * 3 : 21) + k], acc);
            }
        pool += fx_finish(acc, Q[DQ_BIAS + 8 + 15], true);
    }
    const float d = fxacc(fx_wrap16(pool), Q[DQ_RAW + 750 - 256], FX_MAGIC);
    return fx_sat16((d - FX_MAGIC) + Q[DQ_BIAS + 24]);
}

template <int MODE>
__global__ void __launch_bounds__(OG_THREADS) k_disc_fwd_q(const int16_t* __restrict__ cand, const int16_t* __restrict__ cond,
                                                           int16_t* __restrict__ score, int64_t B) {
    __shared__ uint4 sm[OG_THREADS * 4];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint4* wsm = sm + warp * 32 * 4;
    const int64_t ntiles = (B + OG_THREADS - 1) / OG_THREADS;
    for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const int64_t base = t * OG_THREADS + warp * 32;
        if (base >= B) continue;
        float a[2][16], c[2][16], x[4][16];
        tile_load_i16(cand, base, B, wsm, lane, a);
        tile_load_i16(cond, base, B, wsm, lane, c);
#pragma unroll
        for (int i = 0; i < 16; ++i) { x[0][i] = a[0][i]; x[1][i] = a[1][i]; x[2][i] = c[0][i]; x[3][i] = c[1][i]; }
        const float s = MODE == OFDMGAN_GEN_Q_SPEC ? critic_q_spec(c_dq, x) : critic_q_rtl(c_dq, x);
        if (base + lane < B) score[base + lane] = (int16_t)__float2int_rn(s);
    }
}

}  // namespace og

using namespace og;

static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

extern "C" int ofdmgan_disc_fwd_q(const int16_t* cand_dev, const int16_t* cond_dev, const int8_t* wrom_host,
                                  const int16_t* brom_host, int16_t* score_dev, int64_t B, int mode, void* stream) {
    if (mode != OFDMGAN_GEN_Q_SPEC && mode != OFDMGAN_GEN_Q_RTL) return OFDMGAN_E_ARG;
    if (B == 0 && wrom_host && brom_host) return 0;
    if (!cand_dev || !cond_dev || !score_dev || !wrom_host || !brom_host || B < 0 || !aligned16(cand_dev) || !aligned16(cond_dev))
        return OFDMGAN_E_ARG;
    cudaStream_t s = (cudaStream_t)stream;
    int rc;
    CallGuard guard(s);
    if ((rc = guard.rc)) return rc;
    if ((rc = upload_dq(wrom_host, brom_host, s))) return rc;
    const int grid = grid_for(B, OG_THREADS, 4);
    if (mode == OFDMGAN_GEN_Q_SPEC) k_disc_fwd_q<OFDMGAN_GEN_Q_SPEC><<<grid, OG_THREADS, 0, s>>>(cand_dev, cond_dev, score_dev, B);
    else k_disc_fwd_q<OFDMGAN_GEN_Q_RTL><<<grid, OG_THREADS, 0, s>>>(cand_dev, cond_dev, score_dev, B);
    return (int)cudaGetLastError();
}
