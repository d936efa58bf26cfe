// libofdmgan generator-side training kernels: the fused generator step of CWGAN-GP (train.py:282-299), the API-level
// MiniGenerator backward, and fused Adam.   ofdmgan_gen_step / ofdmgan_gen_bwd_f32 / ofdmgan_adam
#include "train_common.cuh"
#include "genbwd_device.cuh"

namespace og {

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
    const float* WG = c_g;
    const float* WD = c_d;
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
    const float* W = c_g;
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

}  // namespace og

using namespace og;

extern "C" {

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
    CallGuard guard(s);
    if ((rc = guard.rc)) return rc;
    slot = 0;
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
    CallGuard guard(s);
    if ((rc = guard.rc)) return rc;
    slot = 0;
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
