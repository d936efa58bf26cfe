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
    double s = 0.0;
    for (int b = warp; b < nblocks; b += 32) s += (double)partials[(size_t)b * slots + grp * 32 + lane];
    red[warp * 32 + lane] = s;
    __syncthreads();
    if (warp == 0) {
        double a = 0.0;
#pragma unroll 4
        for (int w = 0; w < 32; ++w) a += red[w * 32 + lane];
        total[lane] = a;
    }
    __syncthreads();
}

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// resident CTAs per SM of the 255-register training kernels (2 x 128 threads)
constexpr int TRAIN_PER_SM = 2;

}  // namespace og
