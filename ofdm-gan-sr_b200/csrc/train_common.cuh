// Shared by the training-side translation units: tile geometry, CTA-level reduction of per-warp gradient accumulators.
#pragma once
#include <cmath>

#include "io_tile.cuh"
#include "weights.cuh"

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


// Sum one 32-slot group over all CTA partial rows, in a fixed order (deterministic): launched with 1024 threads per
// block, one block per group.  Warp w adds rows w, w+32, ... (one coalesced 128-byte read per row), then warp 0 adds the
// 32 warp totals in order.  Returns the group's 32 totals in `total` (shared) after a barrier.
__device__ __forceinline__ void reduce_group_rows(const float* __restrict__ partials, int nblocks, int slots, int grp, double* red /*[32][32]*/,
                                                  double* total /*[32]*/) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // Every row load of a warp is issued before the first add (up to 20 in flight: a full B200 grid is 592 rows = 18.5 per warp, and
    // the tail kernels sit on the critical path of a training iteration); four chains, fixed order.
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
    const float* col = partials + grp * 32 + lane;
    for (int b0 = warp; b0 < nblocks; b0 += 32 * 20) {
        float v[20];
#pragma unroll
        for (int k = 0; k < 20; ++k) {
            const int b = b0 + 32 * k;
            v[k] = b < nblocks ? __ldcg(col + (size_t)b * slots) : 0.f;
        }
#pragma unroll
        for (int k = 0; k < 20; k += 4) {
            s0 += (double)v[k]; s1 += (double)v[k + 1]; s2 += (double)v[k + 2]; s3 += (double)v[k + 3];
        }
    }
    red[warp * 32 + lane] = (s0 + s1) + (s2 + s3);
    __syncthreads();
    if (warp == 0) {
        double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
#pragma unroll
        for (int w = 0; w < 32; w += 4) {
            a0 += red[w * 32 + lane]; a1 += red[(w + 1) * 32 + lane]; a2 += red[(w + 2) * 32 + lane]; a3 += red[(w + 3) * 32 + lane];
        }
        total[lane] = (a0 + a1) + (a2 + a3);
    }
    __syncthreads();
}

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// resident CTAs per SM of the 255-register training kernels (2 x 128 threads)
constexpr int TRAIN_PER_SM = 2;

// ---- Adam: torch.optim.Adam._single_tensor_adam, fp32 state, no amsgrad / weight decay (oracle/fp32_models.c oracle_adam).
// Explicit round-to-nearest ops so nothing is contracted into an FMA the eager reference does not have.
struct AdamCoef { float step_size, bc2_sqrt, w, b2, omb2, eps; };
static inline AdamCoef adam_coef(double lr, double beta1, double beta2, double eps, int step) {
    const double bc1 = 1.0 - pow(beta1, step), bc2 = 1.0 - pow(beta2, step);
    return AdamCoef{(float)(lr / bc1), (float)sqrt(bc2), (float)(1.0 - beta1), (float)beta2, (float)(1.0 - beta2), (float)eps};
}
#ifdef __CUDACC__
// the same coefficients from a step count that lives on the device (graph-replayable optimiser steps): beta^t by repeated squaring
__device__ __forceinline__ double ipow_dev(double b, int t) {
    double r = 1.0;
    while (t > 0) { if (t & 1) r *= b; b *= b; t >>= 1; }
    return r;
}
__device__ __forceinline__ AdamCoef adam_coef_dev(double lr, double beta1, double beta2, double eps, int step) {
    const double bc1 = 1.0 - ipow_dev(beta1, step), bc2 = 1.0 - ipow_dev(beta2, step);
    return AdamCoef{(float)(lr / bc1), (float)sqrt(bc2), (float)(1.0 - beta1), (float)beta2, (float)(1.0 - beta2), (float)eps};
}
__device__ __forceinline__ void adam_one(float& p, float& m, float& v, float gi, const AdamCoef& c) {
    float mi = m, vi = v;
    mi = c.w < 0.5f ? __fadd_rn(mi, __fmul_rn(c.w, __fsub_rn(gi, mi)))
                    : __fsub_rn(gi, __fmul_rn(__fsub_rn(gi, mi), __fsub_rn(1.0f, c.w)));     // lerp_(grad, 1-beta1)
    vi = __fadd_rn(__fmul_rn(vi, c.b2), __fmul_rn(__fmul_rn(c.omb2, gi), gi));               // mul_(b2).addcmul_(g,g,1-b2)
    const float denom = __fadd_rn(__fdiv_rn(__fsqrt_rn(vi), c.bc2_sqrt), c.eps);
    p = __fsub_rn(p, __fmul_rn(c.step_size, __fdiv_rn(mi, denom)));                          // addcdiv_(m, denom, -step_size)
    m = mi;
    v = vi;
}
#endif

}  // namespace og
