// Per-thread running metric sums of the fused simulator kernels and their reduction: MSE / EVM(dB) per trial, mean and
// standard deviation across trials via sum and sum of squares (benchmark_comparison.py:137-146,253-259).
#pragma once
#include "common.cuh"

namespace og {

constexpr int NM = OFDMGAN_N_METHODS, NC = OFDMGAN_METRIC_COLS;
constexpr int FLUSH_EVERY = 32;     // frames a thread accumulates in fp32 before folding into the double table

// per-thread running sums for the SNR bin the thread is currently in; methods [GAN, NoEQ] or [GAN, NoEQ, ZF, MMSE]
template <int NMETH>
struct Acc {
    float mse[NMETH], mse2[NMETH], evm[NMETH], evm2[NMETH], ratio[NMETH], errs[NMETH];
    float nbits;       // payload bits compared per method (same for all)
    int count;         // live frames accumulated
    int bin;
};

template <int NMETH>
__device__ __forceinline__ void acc_reset(Acc<NMETH>& a, int bin) {
#pragma unroll
    for (int m = 0; m < NMETH; ++m) { a.mse[m] = a.mse2[m] = a.evm[m] = a.evm2[m] = a.ratio[m] = a.errs[m] = 0.f; }
    a.nbits = 0.f;
    a.count = 0;
    a.bin = bin;
}

// sum |ref|^2 of a frame (the EVM denominator; computed once per frame and shared by both methods)
__device__ __forceinline__ float frame_energy(const float (&cr)[16], const float (&ci)[16]) {
    float sr = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) sr = fmaf(cr[i], cr[i], fmaf(ci[i], ci[i], sr));
    return sr;
}

__device__ __forceinline__ void frame_err(const float (&er)[16], const float (&ei)[16], const float (&cr)[16],
                                          const float (&ci)[16], float inv_energy, float& mse, float& evm, float& ratio) {
    float se = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        const float a = er[i] - cr[i], b = ei[i] - ci[i];
        se = fmaf(a, a, fmaf(b, b, se));
    }
    mse = se * 0.03125f;
    ratio = se * inv_energy;
    // 20 log10(sqrt(mean|e|^2 / mean|ref|^2) + 1e-10)     benchmark_comparison.py:142-146
    evm = 6.020599913279624f * fast_lg2(fast_sqrt(ratio) + 1e-10f);
}

// mse / evm / ratio from a frame's summed squared error
__device__ __forceinline__ void err_to_metrics(float se, float inv_energy, float& mse, float& evm, float& ratio) {
    mse = se * 0.03125f;
    ratio = se * inv_energy;
    evm = 6.020599913279624f * fast_lg2(fast_sqrt(ratio) + 1e-10f);
}

template <bool BITS, int NMETH>
__device__ __forceinline__ void acc_add(Acc<NMETH>& a, int m, float mse, float evm, float ratio, int errs) {
    a.mse[m] += mse; a.mse2[m] = fmaf(mse, mse, a.mse2[m]);
    a.evm[m] += evm; a.evm2[m] = fmaf(evm, evm, a.evm2[m]);
    a.ratio[m] += ratio;
    if (BITS) a.errs[m] += (float)errs;
}

// fold the thread's running sums into the CTA table (shared, double).  Warp-uniform bins (the common case) are
// transpose-reduced: P-1 shuffles leave column (lane % P) of the warp total in every lane, then one atomic per column.
// TBL_NM: methods per bin in `table` (the warp-specialised kernel keeps only its two rows)
template <bool BITS, int NMETH, int TBL_NM = NM>
__device__ __forceinline__ void acc_flush(Acc<NMETH>& a, double* table, int lane) {
    const unsigned full = 0xffffffffu;
    constexpr int NCOL = BITS ? NC : NC - 2;                         // columns 5, 6 (bit errors, bits) only with a payload
    constexpr int NV = NMETH * NCOL, P = NV <= 16 ? 16 : 32;
    const int bin0 = __shfl_sync(full, a.bin, 0);
    const bool uniform = __all_sync(full, a.bin == bin0);
    float v[P];
#pragma unroll
    for (int m = 0; m < NMETH; ++m) {
        const float col[NC] = {(float)a.count, a.mse[m], a.mse2[m], a.evm[m], a.evm2[m], a.errs[m], a.nbits, a.ratio[m]};
        int j = 0;
#pragma unroll
        for (int c = 0; c < NC; ++c) {
            if (!BITS && (c == 5 || c == 6)) continue;
            v[m * NCOL + j++] = col[c];
        }
    }
#pragma unroll
    for (int i = NV; i < P; ++i) v[i] = 0.f;
    if (uniform) {
#pragma unroll
        for (int s = P / 2; s >= 1; s >>= 1) {
            const bool upper = (lane & s) != 0;
#pragma unroll
            for (int i = 0; i < s; ++i) {
                const float send = upper ? v[i] : v[i + s], keep = upper ? v[i + s] : v[i];
                v[i] = keep + __shfl_xor_sync(full, send, s);
            }
        }
        if (P == 16) v[0] += __shfl_xor_sync(full, v[0], 16);
        const int q = lane & (P - 1), m = q / NCOL, j = q - m * NCOL;
        const int c = BITS ? j : (j < 5 ? j : 7);
        if (lane < NV && bin0 >= 0) atomicAdd(&table[(bin0 * TBL_NM + m) * NC + c], (double)v[0]);
    } else if (a.bin >= 0) {
#pragma unroll
        for (int q = 0; q < NV; ++q) {
            const int m = q / NCOL, j = q - m * NCOL, c = BITS ? j : (j < 5 ? j : 7);
            if (v[q] != 0.f) atomicAdd(&table[(a.bin * TBL_NM + m) * NC + c], (double)v[q]);
        }
    }
    acc_reset(a, a.bin);
}

// fixed-order sum of the per-CTA partial tables into the caller's accumulator (deterministic for a given grid): one warp per
// table entry - lane l adds rows l, l+32, ... in order, then a fixed shuffle tree.  (One thread per entry walked the ~148 rows as
// one dependent chain: 21 us per call.)
constexpr int RP_WARPS = 8;
static __global__ void __launch_bounds__(RP_WARPS * 32) k_reduce_partials(const double* __restrict__ partials, int nblocks, int n,
                                                                          double* __restrict__ metrics) {
    const int i = blockIdx.x * RP_WARPS + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (i >= n) return;
    double s = 0.0;
    for (int b = lane; b < nblocks; b += 32) s += partials[(size_t)b * n + i];
    s = warp_sum(s);
    if (lane == 0) metrics[i] += s;
}
static inline void reduce_partials_launch(const double* partials, int nblocks, int n, double* metrics, cudaStream_t s) {
    k_reduce_partials<<<(n + RP_WARPS - 1) / RP_WARPS, RP_WARPS * 32, 0, s>>>(partials, nblocks, n, metrics);
}

}  // namespace og
