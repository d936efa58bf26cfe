// libofdmgan kernel (4), API-level entry points: MiniDiscriminator forward and its backward for an arbitrary upstream
// gradient (what the autograd.Function of models/discriminator.py needs).   ofdmgan_disc_fwd_f32 / ofdmgan_disc_bwd_f32
#include "train_common.cuh"
#include "critic_stream.cuh"

namespace og {

constexpr int CAPI_PER_SM = 4;

// ------------------------------------------------------------------------------------------------ critic forward
__global__ void __launch_bounds__(OG_THREADS, CAPI_PER_SM) k_disc_fwd(const float* __restrict__ cand, const float* __restrict__ cond,
                                                                      float* __restrict__ score, int64_t B, float slope) {
    __shared__ float4 sm[2 * OG_THREADS * 8];
    __shared__ float dummy[OG_THREADS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float4* t_a = sm + warp * TILE4;
    float4* t_c = sm + (NWARP + warp) * TILE4;
    SAcc acc{dummy + threadIdx.x, 0};
    const int64_t ntiles = (B + OG_THREADS - 1) / OG_THREADS;
    for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const int64_t base = t * OG_THREADS + warp * 32;
        if (base >= B) continue;
        __syncwarp();
        tile_fill_f32(cand, base, B, t_a, lane);
        tile_fill_f32(cond, base, B, t_c, lane);
        __syncwarp();
        uint64_t m1, m2;
        const float s = cs_score_only(c_d, slope, t_a, t_c, acc, lane, m1, m2);
        if (base + lane < B) score[base + lane] = s;
    }
}

// ------------------------------------------------------------------------------------------------ critic backward
// upstream g[b] = dL/dscore_b: dparams = sum_b g_b dD_b/dtheta (per-CTA partial rows), optionally dL/dcand, dL/dcond
template <bool NEED_DU>
__global__ void __launch_bounds__(OG_THREADS, CAPI_PER_SM) k_disc_bwd(const float* __restrict__ cand, const float* __restrict__ cond,
                                                                      const float* __restrict__ g, float* __restrict__ dcand,
                                                                      float* __restrict__ dcond, float* __restrict__ partials, int64_t B,
                                                                      float slope) {
    __shared__ float4 sm[2 * OG_THREADS * 8];
    __shared__ float sacc[CS_NG * OG_THREADS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float4* t_a = sm + warp * TILE4;
    float4* t_c = sm + (NWARP + warp) * TILE4;
    SAcc acc{sacc + threadIdx.x, OG_THREADS};
#pragma unroll
    for (int k = 0; k < CS_NG; ++k) sacc[k * OG_THREADS + threadIdx.x] = 0.f;
    const int64_t ntiles = (B + OG_THREADS - 1) / OG_THREADS;
    for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const int64_t base = t * OG_THREADS + warp * 32;
        if (base >= B) continue;
        const int64_t b = base + lane;
        __syncwarp();
        tile_fill_f32(cand, base, B, t_a, lane);
        tile_fill_f32(cond, base, B, t_c, lane);
        __syncwarp();
        cs_score_pass<NEED_DU>(c_d, slope, b < B ? g[b] : 0.f, t_a, t_c, acc, lane);
        if (NEED_DU) {
            __syncwarp();
            if (dcand) tile_drain_f32(dcand, base, B, t_a, lane);
            if (dcond) tile_drain_f32(dcond, base, B, t_c, lane);
        }
    }
    if (partials) {
        __syncthreads();
        float* row = partials + (size_t)blockIdx.x * CS_SLOTS;
        for (int s = threadIdx.x; s < CS_SLOTS; s += OG_THREADS) {
            const int grp = s >> 5, j = s & 31;
            float t = 0.f;
#pragma unroll
            for (int w = 0; w < NWARP; ++w) t += sacc[grp * OG_THREADS + w * 32 + j];
            row[s] = t;
        }
    }
}

// one block per accumulator group: fixed-order sum over CTA rows, slot -> parameter order
__global__ void __launch_bounds__(1024) k_finalize_dparams(const float* __restrict__ partials, int nblocks, float* __restrict__ grads) {
    __shared__ double red[32 * 32], total[32];
    reduce_group_rows(partials, nblocks, CS_SLOTS, blockIdx.x, red, total);
    if (threadIdx.x < 32) {
        const int i = cs_param_of(blockIdx.x, threadIdx.x);
        if (i >= 0) grads[i] = (float)total[threadIdx.x];
    }
}

}  // namespace og

using namespace og;

extern "C" {

int ofdmgan_disc_fwd_f32(const float* cand_dev, const float* cond_dev, const float* dparams521, float* score_dev, int64_t B,
                         float leaky_slope, void* stream) {
    if (!slope_ok(leaky_slope)) return OFDMGAN_E_ARG;
    if (B == 0 && dparams521) return 0;
    if (!cand_dev || !cond_dev || !dparams521 || !score_dev || B < 0 || !aligned16(cand_dev) || !aligned16(cond_dev)) return OFDMGAN_E_ARG;
    cudaStream_t s = (cudaStream_t)stream;
    int rc;
    CallGuard guard(s);
    if ((rc = guard.rc)) return rc;
    if ((rc = upload_d(dparams521, 0, s))) return rc;
    k_disc_fwd<<<grid_for(B, OG_THREADS, CAPI_PER_SM), OG_THREADS, 0, s>>>(cand_dev, cond_dev, score_dev, B, leaky_slope);
    return (int)cudaGetLastError();
}

int ofdmgan_disc_bwd_f32(const float* cand_dev, const float* cond_dev, const float* dparams521, const float* g_dev,
                         float* dcand_dev, float* dcond_dev, float* dparams521_dev, int64_t B, float leaky_slope, void* stream) {
    cudaStream_t s = (cudaStream_t)stream;
    if (B < 0 || !dparams521 || !slope_ok(leaky_slope)) return OFDMGAN_E_ARG;
    if (B > 0 && (!cand_dev || !cond_dev || !g_dev || !aligned16(cand_dev) || !aligned16(cond_dev))) return OFDMGAN_E_ARG;
    if ((dcand_dev && !aligned16(dcand_dev)) || (dcond_dev && !aligned16(dcond_dev))) return OFDMGAN_E_ARG;
    if (B == 0) {
        if (dparams521_dev) OG_CHECK(cudaMemsetAsync(dparams521_dev, 0, OFDMGAN_D_NPARAMS * sizeof(float), s));
        return 0;
    }
    int rc;
    CallGuard guard(s);
    if ((rc = guard.rc)) return rc;
    if ((rc = upload_d(dparams521, 0, s))) return rc;
    const int grid = grid_for(B, OG_THREADS, CAPI_PER_SM);
    void* partials = nullptr;
    if (dparams521_dev && (rc = scratch_for_slot(0, (size_t)grid * CS_SLOTS * sizeof(float), 6, &partials))) return rc;
    if (dcand_dev || dcond_dev)
        k_disc_bwd<true><<<grid, OG_THREADS, 0, s>>>(cand_dev, cond_dev, g_dev, dcand_dev, dcond_dev, (float*)partials, B, leaky_slope);
    else
        k_disc_bwd<false><<<grid, OG_THREADS, 0, s>>>(cand_dev, cond_dev, g_dev, nullptr, nullptr, (float*)partials, B, leaky_slope);
    OG_CHECK(cudaGetLastError());
    if (dparams521_dev) {
        k_finalize_dparams<<<CS_NG, 1024, 0, s>>>((const float*)partials, grid, dparams521_dev);
        OG_CHECK(cudaGetLastError());
    }
    return 0;
}

}  // extern "C"
