// libofdmgan training-side kernels (sm_100a):
//   (4) critic forward / backward / gradient penalty with closed-form double backward, the fused critic step
//       ofdmgan_disc_fwd_f32 / ofdmgan_disc_bwd_f32 / ofdmgan_gradient_penalty / ofdmgan_critic_step
//   generator backward and the fused generator step      ofdmgan_gen_bwd_f32 / ofdmgan_gen_step
//   fused Adam                                           ofdmgan_adam
// One sample per thread.  The warp's frame tiles stay resident in shared memory and are re-read when a pass needs
// them again; parameter gradients are summed over the warp's 32 samples with a transpose-reduce and kept in per-lane
// registers across all tiles (critic_device.cuh), then CTA partial rows -> fixed-order finalize (deterministic).
#include <cmath>

#include "genbwd_device.cuh"
#include "io_tile.cuh"

namespace og {

constexpr int NWARP = OG_THREADS / 32;
constexpr int TILE4 = 32 * 8;                     // float4 per warp tile (32 frames x 128 B)

// per-warp accumulators -> this CTA's row of the partial table.  `red` is shared scratch of NWARP*32*NG floats and
// may alias the frame tiles: the leading barrier retires every warp's last tile access first.
template <int NG>
__device__ __forceinline__ void cta_store_partials(const GradAcc<NG>& acc, float* red, float* __restrict__ row) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __syncthreads();
#pragma unroll
    for (int g = 0; g < NG; ++g) red[warp * 32 * NG + g * 32 + lane] = acc.g[g];
    __syncthreads();
    for (int s = threadIdx.x; s < 32 * NG; s += OG_THREADS) {
        float t = 0.f;
#pragma unroll
        for (int w = 0; w < NWARP; ++w) t += red[w * 32 * NG + s];
        row[s] = t;
    }
}

// ------------------------------------------------------------------------------------------------ critic forward
__global__ void __launch_bounds__(OG_THREADS) k_disc_fwd(const float* __restrict__ cand, const float* __restrict__ cond,
                                                         float* __restrict__ score, int64_t B, int slot, float slope) {
    __shared__ float4 sm[OG_THREADS * 8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float4* wsm = sm + warp * TILE4;
    const float* W = c_d[slot];
    const int64_t ntiles = (B + OG_THREADS - 1) / OG_THREADS;
    for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const int64_t base = t * OG_THREADS + warp * 32;
        if (base >= B) continue;
        float x[2][16], c[2][16], a1[8][8], pool[16], s;
        uint64_t m2;
        tile_load_f32(cand, base, B, wsm, lane, x);
        tile_load_f32(cond, base, B, wsm, lane, c);
        disc_fwd(W, slope, x, c, a1, m2, pool, s);
        if (base + lane < B) score[base + lane] = s;
    }
}

// ------------------------------------------------------------------------------------------------ critic passes
// L += g * D(cand, cond) for the thread's sample, frames re-read from the warp's resident tiles.
// Returns the score; accumulates dL/dtheta; if DU_ROWS > 0 also returns dL/d(input rows 0..DU_ROWS-1).
template <int DU_ROWS>
__device__ __forceinline__ float score_pass(const float* __restrict__ W, float slope, float g, const float4* t_cand,
                                            const float4* t_cond, bool want_grads, GradAcc<D_NG>& acc, int lane,
                                            float (&du)[DU_ROWS > 0 ? DU_ROWS : 1][16]) {
    float dz1[8][8], score;
    {
        uint64_t m1, m2;
        {
            float a1[8][8], pool[16];
            {
                float cand[2][16], cond[2][16];
                tile_read_f32(t_cand, lane, cand);
                tile_read_f32(t_cond, lane, cond);
                disc_fwd(W, slope, cand, cond, a1, m2, pool, score);
            }
            if (want_grads) {
                grads_conv2_w(W, slope, g, m2, a1, acc, lane);
                grads_c2b_fcw(W, slope, g, m2, pool, acc, lane);
            }
            m1 = sign_mask(a1);
        }
        disc_bwd_to_z1(W, slope, g, m1, m2, dz1);
    }
    if (want_grads) {
        float u[4][16];
        {
            float cand[2][16], cond[2][16];
            tile_read_f32(t_cand, lane, cand);
            tile_read_f32(t_cond, lane, cond);
#pragma unroll
            for (int i = 0; i < 16; ++i) { u[0][i] = cand[0][i]; u[1][i] = cand[1][i]; u[2][i] = cond[0][i]; u[3][i] = cond[1][i]; }
        }
        grads_conv1_w<4>(dz1, u, acc, lane);
        grads_c1b_fcb(dz1, g, acc, lane);
    }
    if (DU_ROWS > 0) disc_bwd_to_input<0, (DU_ROWS > 0 ? DU_ROWS : 1)>(W, dz1, du);
    return score;
}

struct CriticArgs {
    const float* real;        // clean
    const float* cond;        // noisy
    const float* fake;
    const float* alpha;       // nullable -> Philox(seed, sample, alpha_iter, purpose 1)
    PhiloxKeys keys;
    uint64_t sample0;
    uint32_t alpha_iter;
    int64_t B;
    int slot;
    float slope;
    float gp_scale;           // weight of the penalty term relative to the score terms
    int want_grads;
    float* partials;          // [grid][DS_SLOTS]
    float* norms;             // nullable [B]: ||grad|| per sample (test hook of ofdmgan_gradient_penalty)
};

// SCORE: the two Wasserstein terms (-D(real) + D(fake)); GP: the penalty term.  Per-sample weights are +-1 and
// gp_scale; the 1/B_global factor is applied once in the finalize kernel.
template <bool SCORE, bool GP>
__global__ void __launch_bounds__(OG_THREADS) k_critic(const __grid_constant__ CriticArgs a) {
    __shared__ float4 sm[3 * OG_THREADS * 8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float4* t_real = sm + warp * TILE4;
    float4* t_cond = sm + (NWARP + warp) * TILE4;
    float4* t_fake = sm + (2 * NWARP + warp) * TILE4;
    const float* W = c_d[a.slot];
    GradAcc<D_NG> acc;
    acc.zero();
    float s_real = 0.f, s_fake = 0.f, s_gp = 0.f;
    const int64_t ntiles = (a.B + OG_THREADS - 1) / OG_THREADS;
    for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const int64_t base = t * OG_THREADS + warp * 32;
        if (base >= a.B) continue;
        const int64_t b = base + lane;
        const bool live = b < a.B;
        __syncwarp();
        tile_fill_f32(a.real, base, a.B, t_real, lane);
        tile_fill_f32(a.cond, base, a.B, t_cond, lane);
        tile_fill_f32(a.fake, base, a.B, t_fake, lane);
        __syncwarp();
        float none[1][16];
        if (SCORE) {
            const float sr = score_pass<0>(W, a.slope, live ? -1.0f : 0.0f, t_real, t_cond, a.want_grads, acc, lane, none);
            const float sf = score_pass<0>(W, a.slope, live ? 1.0f : 0.0f, t_fake, t_cond, a.want_grads, acc, lane, none);
            if (live) { s_real += sr; s_fake += sf; }
        }
        if (GP) {
            float alpha = 0.f;
            if (live) {
                if (a.alpha) {
                    alpha = a.alpha[b];
                } else {
                    const uint64_t smp = a.sample0 + (uint64_t)b;
                    uint32_t x[4];
                    philox4x32_10(a.keys, (uint32_t)smp, (uint32_t)(smp >> 32), a.alpha_iter, 1u, x);
                    alpha = u_half(x[0]);
                }
            }
            float xh[2][16], cond[2][16];
            {
                float r[2][16], f[2][16];
                tile_read_f32(t_real, lane, r);
                tile_read_f32(t_fake, lane, f);
                const float om = 1.0f - alpha;
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    // alpha * real + (1 - alpha) * fake, two products then a sum (models/discriminator.py:211)
                    xh[0][i] = __fadd_rn(__fmul_rn(alpha, r[0][i]), __fmul_rn(om, f[0][i]));
                    xh[1][i] = __fadd_rn(__fmul_rn(alpha, r[1][i]), __fmul_rn(om, f[1][i]));
                }
            }
            tile_read_f32(t_cond, lane, cond);
            float n;
            const float pen = critic_gp_pass(W, a.slope, live ? a.gp_scale : 0.0f, xh, cond, a.want_grads != 0, acc, lane, n);
            if (live) {
                s_gp += pen;
                if (a.norms) a.norms[b] = n;
            }
        }
    }
    // statistics ride in the spare slots of group 3
    {
        const float r = warp_sum(s_real), f = warp_sum(s_fake), g = warp_sum(s_gp);
        if (lane == DS_SREAL - 96) acc.g[3] += r;
        if (lane == DS_SFAKE - 96) acc.g[3] += f;
        if (lane == DS_SGP - 96) acc.g[3] += g;
    }
    cta_store_partials<D_NG>(acc, reinterpret_cast<float*>(sm), a.partials + (size_t)blockIdx.x * DS_SLOTS);
}

// parameter index (torch order) -> accumulator slot
__device__ __forceinline__ int critic_slot_of(int i) {
    if (i < DP_C1_B) return DS_C1W + i;
    if (i < DP_C2_W) return DS_C1B + (i - DP_C1_B);
    if (i < DP_C2_B) return DS_C2W + (i - DP_C2_W);
    if (i < DP_FC_W) return DS_C2B + (i - DP_C2_B);
    if (i < DP_FC_B) return DS_FCW + (i - DP_FC_W);
    return DS_FCB;
}

// Fixed-order sum over CTA rows, slot -> parameter order, 1/B scaling, loss statistics.
//   grads (nullable): 521 floats.  stats (nullable): stat_mode 0 -> the 5 scalars of train.py:255-261,
//   stat_mode 1 -> 1 float = mean penalty.
__global__ void __launch_bounds__(DS_SLOTS) k_finalize_critic(const float* __restrict__ partials, int nblocks, double inv_b,
                                                              double gp_weight, float* __restrict__ grads,
                                                              float* __restrict__ stats, int stat_mode) {
    __shared__ double s[DS_SLOTS];
    const int t = threadIdx.x;
    double sum = 0.0;
    for (int b = 0; b < nblocks; ++b) sum += (double)partials[(size_t)b * DS_SLOTS + t];
    s[t] = sum;
    __syncthreads();
    if (grads && t < OFDMGAN_D_NPARAMS) grads[t] = (float)(s[critic_slot_of(t)] * inv_b);
    if (stats && t == 0) {
        const double dr = s[DS_SREAL] * inv_b, df = s[DS_SFAKE] * inv_b, gp = s[DS_SGP] * inv_b;
        if (stat_mode == 0) {
            stats[0] = (float)(df - dr + gp_weight * gp);
            stats[1] = (float)(dr - df);
            stats[2] = (float)gp;
            stats[3] = (float)dr;
            stats[4] = (float)df;
            stats[5] = 0.f;
            stats[6] = 0.f;
        } else {
            stats[0] = (float)gp;
        }
    }
}

// ------------------------------------------------------------------------------------------------ critic backward (API)
template <bool NEED_DU>
__global__ void __launch_bounds__(OG_THREADS) k_disc_bwd(const float* __restrict__ cand, const float* __restrict__ cond,
                                                         const float* __restrict__ g, float* __restrict__ dcand,
                                                         float* __restrict__ dcond, float* __restrict__ partials, int64_t B,
                                                         int slot, float slope) {
    __shared__ float4 sm[3 * OG_THREADS * 8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float4* t_cand = sm + warp * TILE4;
    float4* t_cond = sm + (NWARP + warp) * TILE4;
    float4* t_out = sm + (2 * NWARP + warp) * TILE4;
    const float* W = c_d[slot];
    GradAcc<D_NG> acc;
    acc.zero();
    const int64_t ntiles = (B + OG_THREADS - 1) / OG_THREADS;
    for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const int64_t base = t * OG_THREADS + warp * 32;
        if (base >= B) continue;
        const int64_t b = base + lane;
        __syncwarp();
        tile_fill_f32(cand, base, B, t_cand, lane);
        tile_fill_f32(cond, base, B, t_cond, lane);
        __syncwarp();
        const float gb = b < B ? g[b] : 0.f;
        float du[NEED_DU ? 4 : 1][16];
        score_pass<NEED_DU ? 4 : 0>(W, slope, gb, t_cand, t_cond, partials != nullptr, acc, lane, du);
        if (NEED_DU) {
            float f[2][16];
            if (dcand) {
#pragma unroll
                for (int i = 0; i < 16; ++i) { f[0][i] = du[0][i]; f[1][i] = du[NEED_DU ? 1 : 0][i]; }
                tile_store_f32(dcand, base, B, t_out, lane, f);
            }
            if (dcond) {
#pragma unroll
                for (int i = 0; i < 16; ++i) { f[0][i] = du[NEED_DU ? 2 : 0][i]; f[1][i] = du[NEED_DU ? 3 : 0][i]; }
                tile_store_f32(dcond, base, B, t_out, lane, f);
            }
        }
    }
    if (partials) cta_store_partials<D_NG>(acc, reinterpret_cast<float*>(sm), partials + (size_t)blockIdx.x * DS_SLOTS);
}

// ------------------------------------------------------------------------------------------------ generator step
struct GenStepArgs {
    const float* clean;
    const float* noisy;
    float* fake_out;          // nullable
    int64_t B;
    int slot;
    float slope;
    float adv_w, rec_w;
    float* partials;          // [grid][GS_SLOTS]
};

__global__ void __launch_bounds__(OG_THREADS) k_gen_step(const __grid_constant__ GenStepArgs a) {
    __shared__ float4 sm[3 * OG_THREADS * 8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float4* t_noisy = sm + warp * TILE4;
    float4* t_clean = sm + (NWARP + warp) * TILE4;
    float4* t_fake = sm + (2 * NWARP + warp) * TILE4;
    const float* WG = c_g[a.slot];
    const float* WD = c_d[a.slot];
    GradAcc<G_NG> acc;
    acc.zero();
    float s_d = 0.f, s_l1 = 0.f;
    const float rec_g = a.rec_w * 0.03125f;                      // rec_w / 32: l1_loss is a mean over B*32 elements
    const int64_t ntiles = (a.B + OG_THREADS - 1) / OG_THREADS;
    for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const int64_t base = t * OG_THREADS + warp * 32;
        if (base >= a.B) continue;
        const bool live = base + lane < a.B;
        __syncwarp();
        tile_fill_f32(a.noisy, base, a.B, t_noisy, lane);
        tile_fill_f32(a.clean, base, a.B, t_clean, lane);
        __syncwarp();
        // pass 1: fake = G(noisy), kept in the warp's third tile
        {
            float x[2][16], y[2][16];
            tile_read_f32(t_noisy, lane, x);
            gen_fwd_f32_infer(WG, a.slope, x, y);
            tile_write_f32(t_fake, lane, y);
        }
        __syncwarp();
        if (a.fake_out) tile_drain_f32(a.fake_out, base, a.B, t_fake, lane);
        // pass 2: adversarial term through the critic, input gradient w.r.t. the candidate rows only
        float dy[2][16];
        {
            float dz1[8][8], score;
            {
                uint64_t m1, m2;
                {
                    float a1[8][8], pool[16], cand[2][16], cond[2][16];
                    tile_read_f32(t_fake, lane, cand);
                    tile_read_f32(t_noisy, lane, cond);
                    disc_fwd(WD, a.slope, cand, cond, a1, m2, pool, score);
                    m1 = sign_mask(a1);
                }
                disc_bwd_to_z1(WD, a.slope, live ? -a.adv_w : 0.f, m1, m2, dz1);
            }
            disc_bwd_to_input<0, 2>(WD, dz1, dy);
            if (live) s_d += score;
        }
        // pass 3: + reconstruction term, then backward through G (forward recomputed with its tape: cheaper than
        // keeping 130 activations live across the critic pass)
        {
            float x[2][16], y[2][16], a1[4][8], a2[8][4], sk[4][8], none[2][16];
            uint32_t z3pos;
            tile_read_f32(t_noisy, lane, x);
            gen_fwd_f32<true>(WG, a.slope, x, y, a1, a2, sk, z3pos);
            {
                float c[2][16];
                tile_read_f32(t_clean, lane, c);
#pragma unroll
                for (int r = 0; r < 2; ++r)
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        const float e = y[r][i] - c[r][i];
                        if (live) {
                            s_l1 += fabsf(e);
                            dy[r][i] += e > 0.f ? rec_g : (e < 0.f ? -rec_g : 0.f);     // l1_loss backward: sign(e)
                        } else {
                            dy[r][i] = 0.f;
                        }
                    }
            }
            gen_bwd<false>(WG, a.slope, x, a1, a2, sk, z3pos, y, dy, acc, lane, none);
        }
    }
    {
        const float d = warp_sum(s_d), l = warp_sum(s_l1);
        if (lane == GS_S0 - 288) acc.g[9] += d;
        if (lane == GS_S0 + 1 - 288) acc.g[9] += l;
    }
    cta_store_partials<G_NG>(acc, reinterpret_cast<float*>(sm), a.partials + (size_t)blockIdx.x * GS_SLOTS);
}

// generator parameter gradient (torch order) from the summed slot table
__device__ __forceinline__ double gen_param_from_slots(const double* s, int i) {
    if (i < GP_ENC_B) return s[GS_ENCW + i];
    if (i < GP_BN_W) return s[GS_ENCB + (i - GP_ENC_B)];
    if (i < GP_BN_B) return s[GS_BNW + (i - GP_BN_W)];
    if (i < GP_DEC_W) return s[GS_BNB + (i - GP_BN_B)];
    if (i < GP_DEC_B) {
        const int j = i - GP_DEC_W, pair = j / 3, k = j % 3;
        const double* F = s + GS_DECF + pair * 4;
        return k == 0 ? F[0] + F[2] : (k == 1 ? F[1] + F[2] : F[1] + F[3]);
    }
    if (i < GP_OUT_W) return s[GS_DECB + (i - GP_DEC_B)];
    if (i < GP_OUT_B) {
        const int j = i - GP_OUT_W, pair = j / 3, k = j % 3;
        const double* F = s + GS_OUTF + pair * 4;
        return k == 0 ? F[0] + F[2] : (k == 1 ? F[1] + F[2] : F[1] + F[3]);
    }
    return s[GS_OUTB + (i - GP_OUT_B)];
}

// stats (nullable): g_loss, adv_loss, rec_loss (train.py:301-305) + 3 pad
__global__ void __launch_bounds__(GS_SLOTS) k_finalize_gen(const float* __restrict__ partials, int nblocks, double inv_b,
                                                           double adv_w, double rec_w, float* __restrict__ grads,
                                                           float* __restrict__ stats) {
    __shared__ double s[GS_SLOTS];
    const int t = threadIdx.x;
    double sum = 0.0;
    for (int b = 0; b < nblocks; ++b) sum += (double)partials[(size_t)b * GS_SLOTS + t];
    s[t] = sum;
    __syncthreads();
    if (grads && t < OFDMGAN_G_NPARAMS) grads[t] = (float)(gen_param_from_slots(s, t) * inv_b);
    if (stats && t == 0) {
        const double adv = -s[GS_S0] * inv_b, rec = s[GS_S0 + 1] * inv_b * 0.03125;
        stats[0] = (float)(adv_w * adv + rec_w * rec);
        stats[1] = (float)adv;
        stats[2] = (float)rec;
        stats[3] = 0.f;
        stats[4] = 0.f;
        stats[5] = 0.f;
    }
}

// ------------------------------------------------------------------------------------------------ generator backward (API)
// (dx is always formed: the dx-less instantiation trips a ptxas 12.9 register-allocation failure, and this entry point
// is the API-level backward, not the training hot path - ofdmgan_gen_step is.)
__global__ void __launch_bounds__(OG_THREADS) k_gen_bwd(const float* __restrict__ x, const float* __restrict__ dy,
                                                        float* __restrict__ dx, float* __restrict__ partials, int64_t B, int slot,
                                                        float slope) {
    __shared__ float4 sm[2 * OG_THREADS * 8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float4* t_x = sm + warp * TILE4;
    float4* t_dy = sm + (NWARP + warp) * TILE4;
    const float* W = c_g[slot];
    GradAcc<G_NG> acc;
    acc.zero();
    const int64_t ntiles = (B + OG_THREADS - 1) / OG_THREADS;
    for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const int64_t base = t * OG_THREADS + warp * 32;
        if (base >= B) continue;
        __syncwarp();
        tile_fill_f32(x, base, B, t_x, lane);
        tile_fill_f32(dy, base, B, t_dy, lane);            // rows beyond B are zero-filled: they contribute nothing
        __syncwarp();
        float xi[2][16], y[2][16], a1[4][8], a2[8][4], sk[4][8], g[2][16], dxo[2][16];
        uint32_t z3pos;
        tile_read_f32(t_x, lane, xi);
        gen_fwd_f32<true>(W, slope, xi, y, a1, a2, sk, z3pos);
        tile_read_f32(t_dy, lane, g);
        gen_bwd<true>(W, slope, xi, a1, a2, sk, z3pos, y, g, acc, lane, dxo);
        if (dx) {
            __syncwarp();
            tile_store_f32(dx, base, B, t_dy, lane, dxo);
        }
    }
    cta_store_partials<G_NG>(acc, reinterpret_cast<float*>(sm), partials + (size_t)blockIdx.x * GS_SLOTS);
}

// ------------------------------------------------------------------------------------------------ Adam
// torch.optim.Adam._single_tensor_adam, fp32 state, no amsgrad / weight decay (oracle/fp32_models.c oracle_adam).
// Written with explicit round-to-nearest ops so nothing is contracted into an FMA the eager reference does not have.
__global__ void k_adam(float* __restrict__ p, float* __restrict__ m, float* __restrict__ v, const float* __restrict__ g, int n,
                       float step_size, float bc2_sqrt, float w, float b2, float omb2, float eps, float grad_scale) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float gi = __fmul_rn(g[i], grad_scale);
    float mi = m[i], vi = v[i];
    mi = w < 0.5f ? __fadd_rn(mi, __fmul_rn(w, __fsub_rn(gi, mi)))
                  : __fsub_rn(gi, __fmul_rn(__fsub_rn(gi, mi), __fsub_rn(1.0f, w)));       // lerp_(grad, 1-beta1)
    vi = __fadd_rn(__fmul_rn(vi, b2), __fmul_rn(__fmul_rn(omb2, gi), gi));                 // mul_(b2).addcmul_(g,g,1-b2)
    const float denom = __fadd_rn(__fdiv_rn(__fsqrt_rn(vi), bc2_sqrt), eps);
    p[i] = __fsub_rn(p[i], __fmul_rn(step_size, __fdiv_rn(mi, denom)));                    // addcdiv_(m, denom, -step_size)
    m[i] = mi;
    v[i] = vi;
}

static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// resident CTAs per SM the training kernels are sized for (register-bound: 2 x 128 threads x <=255 registers)
constexpr int TRAIN_PER_SM = 2;

}  // namespace og

using namespace og;

extern "C" {

int ofdmgan_disc_fwd_f32(const float* cand_dev, const float* cond_dev, const float* dparams521, float* score_dev, int64_t B,
                         float leaky_slope, void* stream) {
    if (B == 0 && dparams521) return 0;
    if (!cand_dev || !cond_dev || !dparams521 || !score_dev || B < 0 || !aligned16(cand_dev) || !aligned16(cond_dev)) return OFDMGAN_E_ARG;
    cudaStream_t s = (cudaStream_t)stream;
    int slot, rc;
    if ((rc = slot_for_stream(s, &slot))) return rc;
    if ((rc = upload_d(dparams521, slot, s))) return rc;
    k_disc_fwd<<<grid_for(B, OG_THREADS, 4), OG_THREADS, 0, s>>>(cand_dev, cond_dev, score_dev, B, slot, leaky_slope);
    return (int)cudaGetLastError();
}

int ofdmgan_disc_bwd_f32(const float* cand_dev, const float* cond_dev, const float* dparams521, const float* g_dev,
                         float* dcand_dev, float* dcond_dev, float* dparams521_dev, int64_t B, float leaky_slope, void* stream) {
    cudaStream_t s = (cudaStream_t)stream;
    if (B < 0 || !dparams521) return OFDMGAN_E_ARG;
    if (B > 0 && (!cand_dev || !cond_dev || !g_dev || !aligned16(cand_dev) || !aligned16(cond_dev))) return OFDMGAN_E_ARG;
    if ((dcand_dev && !aligned16(dcand_dev)) || (dcond_dev && !aligned16(dcond_dev))) return OFDMGAN_E_ARG;
    if (B == 0) {
        if (dparams521_dev) OG_CHECK(cudaMemsetAsync(dparams521_dev, 0, OFDMGAN_D_NPARAMS * sizeof(float), s));
        return 0;
    }
    int slot, rc;
    if ((rc = slot_for_stream(s, &slot))) return rc;
    if ((rc = upload_d(dparams521, slot, s))) return rc;
    const int grid = grid_for(B, OG_THREADS, TRAIN_PER_SM);
    void* partials = nullptr;
    if (dparams521_dev && (rc = scratch_for_slot(slot, (size_t)grid * DS_SLOTS * sizeof(float), 6, &partials))) return rc;
    if (dcand_dev || dcond_dev)
        k_disc_bwd<true><<<grid, OG_THREADS, 0, s>>>(cand_dev, cond_dev, g_dev, dcand_dev, dcond_dev, (float*)partials, B, slot, leaky_slope);
    else
        k_disc_bwd<false><<<grid, OG_THREADS, 0, s>>>(cand_dev, cond_dev, g_dev, nullptr, nullptr, (float*)partials, B, slot, leaky_slope);
    OG_CHECK(cudaGetLastError());
    if (dparams521_dev) {
        k_finalize_critic<<<1, DS_SLOTS, 0, s>>>((const float*)partials, grid, 1.0, 0.0, dparams521_dev, nullptr, 0);
        OG_CHECK(cudaGetLastError());
    }
    return 0;
}

static int launch_critic(bool score, const float* real, const float* fake, const float* cond, const float* alpha, uint64_t seed,
                         uint64_t sample0, uint32_t alpha_iter, const float* dparams521, float gp_scale, float slope, int64_t B,
                         bool want_grads, float* norms, cudaStream_t s, int* slot_out, int* grid_out, void** partials_out) {
    int slot, rc;
    if ((rc = slot_for_stream(s, &slot))) return rc;
    if ((rc = upload_d(dparams521, slot, s))) return rc;
    const int grid = grid_for(B, OG_THREADS, TRAIN_PER_SM);
    void* partials = nullptr;
    if ((rc = scratch_for_slot(slot, (size_t)grid * DS_SLOTS * sizeof(float), 6, &partials))) return rc;
    CriticArgs a{};
    a.real = real; a.cond = cond; a.fake = fake; a.alpha = alpha;
    a.keys = philox_keys(seed);
    a.sample0 = sample0; a.alpha_iter = alpha_iter;
    a.B = B; a.slot = slot; a.slope = slope; a.gp_scale = gp_scale;
    a.want_grads = want_grads ? 1 : 0;
    a.partials = (float*)partials;
    a.norms = norms;
    if (score) k_critic<true, true><<<grid, OG_THREADS, 0, s>>>(a);
    else k_critic<false, true><<<grid, OG_THREADS, 0, s>>>(a);
    OG_CHECK(cudaGetLastError());
    *slot_out = slot; *grid_out = grid; *partials_out = partials;
    return 0;
}

int ofdmgan_gradient_penalty(const float* real_dev, const float* fake_dev, const float* cond_dev, const float* alpha_dev,
                             uint64_t seed, uint64_t sample0, uint32_t alpha_iter, const float* dparams521, float* gp_dev,
                             float* dparams521_dev, int64_t B, float leaky_slope, void* stream) {
    if (!real_dev || !fake_dev || !cond_dev || !dparams521 || !gp_dev || B < 1) return OFDMGAN_E_ARG;
    if (!aligned16(real_dev) || !aligned16(fake_dev) || !aligned16(cond_dev)) return OFDMGAN_E_ARG;
    cudaStream_t s = (cudaStream_t)stream;
    int slot, grid, rc;
    void* partials;
    if ((rc = launch_critic(false, real_dev, fake_dev, cond_dev, alpha_dev, seed, sample0, alpha_iter, dparams521, 1.0f, leaky_slope, B,
                            dparams521_dev != nullptr, nullptr, s, &slot, &grid, &partials))) return rc;
    k_finalize_critic<<<1, DS_SLOTS, 0, s>>>((const float*)partials, grid, 1.0 / (double)B, 1.0, dparams521_dev, gp_dev, 1);
    return (int)cudaGetLastError();
}

int ofdmgan_critic_step(const float* clean_dev, const float* noisy_dev, const float* fake_dev, const float* alpha_dev, uint64_t seed,
                        uint64_t sample0, uint32_t alpha_iter, const float* dparams521, float gp_weight, float leaky_slope,
                        int64_t B_local, int64_t B_global, float* out_dev, void* stream) {
    cudaStream_t s = (cudaStream_t)stream;
    if (!dparams521 || !out_dev || B_local < 0 || B_global < 1 || B_global < B_local) return OFDMGAN_E_ARG;
    if (B_local > 0 && (!clean_dev || !noisy_dev || !fake_dev || !aligned16(clean_dev) || !aligned16(noisy_dev) || !aligned16(fake_dev)))
        return OFDMGAN_E_ARG;
    if (B_local == 0) {
        OG_CHECK(cudaMemsetAsync(out_dev, 0, OFDMGAN_CRITIC_OUT * sizeof(float), s));
        return 0;
    }
    int slot, grid, rc;
    void* partials;
    if ((rc = launch_critic(true, clean_dev, fake_dev, noisy_dev, alpha_dev, seed, sample0, alpha_iter, dparams521, gp_weight,
                            leaky_slope, B_local, true, nullptr, s, &slot, &grid, &partials))) return rc;
    k_finalize_critic<<<1, DS_SLOTS, 0, s>>>((const float*)partials, grid, 1.0 / (double)B_global, (double)gp_weight, out_dev,
                                             out_dev + OFDMGAN_D_NPARAMS, 0);
    return (int)cudaGetLastError();
}

int ofdmgan_gen_step(const float* clean_dev, const float* noisy_dev, const float* dparams521, const float* gparams258, float adv_weight,
                     float rec_weight, float leaky_slope, int64_t B_local, int64_t B_global, float* out_dev, float* fake_out_dev,
                     void* stream) {
    cudaStream_t s = (cudaStream_t)stream;
    if (!dparams521 || !gparams258 || !out_dev || B_local < 0 || B_global < 1 || B_global < B_local) return OFDMGAN_E_ARG;
    if (B_local > 0 && (!clean_dev || !noisy_dev || !aligned16(clean_dev) || !aligned16(noisy_dev))) return OFDMGAN_E_ARG;
    if (fake_out_dev && !aligned16(fake_out_dev)) return OFDMGAN_E_ARG;
    if (B_local == 0) {
        OG_CHECK(cudaMemsetAsync(out_dev, 0, OFDMGAN_GEN_OUT * sizeof(float), s));
        return 0;
    }
    int slot, rc;
    if ((rc = slot_for_stream(s, &slot))) return rc;
    if ((rc = upload_d(dparams521, slot, s))) return rc;
    if ((rc = upload_g(gparams258, slot, s))) return rc;
    const int grid = grid_for(B_local, OG_THREADS, TRAIN_PER_SM);
    void* partials = nullptr;
    if ((rc = scratch_for_slot(slot, (size_t)grid * GS_SLOTS * sizeof(float), 7, &partials))) return rc;
    GenStepArgs a{};
    a.clean = clean_dev; a.noisy = noisy_dev; a.fake_out = fake_out_dev;
    a.B = B_local; a.slot = slot; a.slope = leaky_slope; a.adv_w = adv_weight; a.rec_w = rec_weight;
    a.partials = (float*)partials;
    k_gen_step<<<grid, OG_THREADS, 0, s>>>(a);
    OG_CHECK(cudaGetLastError());
    k_finalize_gen<<<1, GS_SLOTS, 0, s>>>((const float*)partials, grid, 1.0 / (double)B_global, (double)adv_weight, (double)rec_weight,
                                          out_dev, out_dev + OFDMGAN_G_NPARAMS);
    return (int)cudaGetLastError();
}

int ofdmgan_gen_bwd_f32(const float* x_dev, const float* gparams258, const float* dy_dev, float* dx_dev, float* dparams258_dev,
                        int64_t B, float leaky_slope, void* stream) {
    cudaStream_t s = (cudaStream_t)stream;
    if (!gparams258 || !dparams258_dev || B < 0) return OFDMGAN_E_ARG;
    if (B > 0 && (!x_dev || !dy_dev || !aligned16(x_dev) || !aligned16(dy_dev))) return OFDMGAN_E_ARG;
    if (dx_dev && !aligned16(dx_dev)) return OFDMGAN_E_ARG;
    if (B == 0) {
        OG_CHECK(cudaMemsetAsync(dparams258_dev, 0, OFDMGAN_G_NPARAMS * sizeof(float), s));
        return 0;
    }
    int slot, rc;
    if ((rc = slot_for_stream(s, &slot))) return rc;
    if ((rc = upload_g(gparams258, slot, s))) return rc;
    const int grid = grid_for(B, OG_THREADS, TRAIN_PER_SM);
    void* partials = nullptr;
    if ((rc = scratch_for_slot(slot, (size_t)grid * GS_SLOTS * sizeof(float), 7, &partials))) return rc;
    k_gen_bwd<<<grid, OG_THREADS, 0, s>>>(x_dev, dy_dev, dx_dev, (float*)partials, B, slot, leaky_slope);
    OG_CHECK(cudaGetLastError());
    k_finalize_gen<<<1, GS_SLOTS, 0, s>>>((const float*)partials, grid, 1.0, 0.0, 0.0, dparams258_dev, nullptr);
    return (int)cudaGetLastError();
}

int ofdmgan_adam(float* p_dev, float* m_dev, float* v_dev, const float* g_dev, int n, double lr, double beta1, double beta2,
                 double eps, int step, float grad_scale, void* stream) {
    if (!p_dev || !m_dev || !v_dev || !g_dev || n < 0 || step < 1) return OFDMGAN_E_ARG;
    if (n == 0) return 0;
    const double bc1 = 1.0 - pow(beta1, step), bc2 = 1.0 - pow(beta2, step);
    k_adam<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(p_dev, m_dev, v_dev, g_dev, n, (float)(lr / bc1), (float)sqrt(bc2),
                                                               (float)(1.0 - beta1), (float)beta2, (float)(1.0 - beta2), (float)eps,
                                                               grad_scale);
    return (int)cudaGetLastError();
}

}  // extern "C"
