// libofdmgan kernel (4), API-level entry points: MiniDiscriminator forward and its backward for an arbitrary upstream
// gradient (what the autograd.Function of models/discriminator.py needs).   ofdmgan_disc_fwd_f32 / ofdmgan_disc_bwd_f32
#include "train_common.cuh"
#include "critic_device.cuh"

namespace og {

// ------------------------------------------------------------------------------------------------ critic forward
__global__ void __launch_bounds__(OG_THREADS) k_disc_fwd(const float* __restrict__ cand, const float* __restrict__ cond,
                                                         float* __restrict__ score, int64_t B, int slot, float slope) {
    __shared__ float4 sm[OG_THREADS * 8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float4* wsm = sm + warp * TILE4;
    const float* W = c_d;
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

// parameter index (torch order) -> accumulator slot (critic_device.cuh slot map)
__device__ __forceinline__ int critic_slot_of(int i) {
    if (i < DP_C1_B) return DS_C1W + i;
    if (i < DP_C2_W) return DS_C1B + (i - DP_C1_B);
    if (i < DP_C2_B) return DS_C2W + (i - DP_C2_W);
    if (i < DP_FC_W) return DS_C2B + (i - DP_C2_B);
    if (i < DP_FC_B) return DS_FCW + (i - DP_FC_W);
    return DS_FCB;
}

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
    const float* W = c_d;
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

}  // namespace og

using namespace og;

extern "C" {

int ofdmgan_disc_fwd_f32(const float* cand_dev, const float* cond_dev, const float* dparams521, float* score_dev, int64_t B,
                         float leaky_slope, void* stream) {
    if (B == 0 && dparams521) return 0;
    if (!cand_dev || !cond_dev || !dparams521 || !score_dev || B < 0 || !aligned16(cand_dev) || !aligned16(cond_dev)) return OFDMGAN_E_ARG;
    cudaStream_t s = (cudaStream_t)stream;
    int slot, rc;
    CallGuard guard(s);
    if ((rc = guard.rc)) return rc;
    slot = 0;
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
    CallGuard guard(s);
    if ((rc = guard.rc)) return rc;
    slot = 0;
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

}  // extern "C"
