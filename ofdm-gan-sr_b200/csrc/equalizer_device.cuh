// Genie-aided zero-forcing and MMSE equalisers of the reference's benchmark (utils/classical_equalizers.py:33-230 as called
// from benchmark_comparison.py:218-226), per sample, on complex64 arithmetic.
//
// Under NumPy 2 the reference's `noisy_iq[0] + 1j * noisy_iq[1]` on float32 arrays is complex64, so the whole equaliser runs
// in single precision and its residual error is rounding noise (ZF EVM ~ -140 dB).  Matching that means reproducing NumPy's
// complex64 operations exactly: CFLOAT_divide is Smith's algorithm with separately rounded operations (no FMA), the "+ eps"
// adds float32(1e-10) to the real part.  cdiv_np below is that, with explicit round-to-nearest intrinsics so nothing is
// contracted; ZF frames are bit-identical to the reference's (tests/golden/ref_eq.npz).  MMSE additionally goes through
// np.abs (a SIMD hypot whose last bit is not reproduced): equal to ~1e-7 relative, which is irrelevant at its -30 dB error.
#pragma once
#include "common.cuh"

namespace og {

// NumPy CFLOAT_divide: (ar + j ai) / (br + j bi)
__device__ __forceinline__ void cdiv_np(float ar, float ai, float br, float bi, float& outr, float& outi) {
    const float abr = fabsf(br), abi = fabsf(bi);
    if (abr >= abi) {
        if (abr == 0.f && abi == 0.f) {                          // divide by zero: complex inf / nan, as NumPy
            outr = __fdiv_rn(ar, abr);
            outi = __fdiv_rn(ai, abr);
            return;
        }
        const float rat = __fdiv_rn(bi, br);
        const float scl = __fdiv_rn(1.0f, __fadd_rn(br, __fmul_rn(bi, rat)));
        outr = __fmul_rn(__fadd_rn(ar, __fmul_rn(ai, rat)), scl);
        outi = __fmul_rn(__fsub_rn(ai, __fmul_rn(ar, rat)), scl);
    } else {
        const float rat = __fdiv_rn(br, bi);
        const float scl = __fdiv_rn(1.0f, __fadd_rn(bi, __fmul_rn(br, rat)));
        outr = __fmul_rn(__fadd_rn(__fmul_rn(ar, rat), ai), scl);
        outi = __fmul_rn(__fsub_rn(__fmul_rn(ai, rat), ar), scl);
    }
}

constexpr float EQ_EPS = 1e-10f;

// H = Y / (X + eps)                                              classical_equalizers.py:61-62 / :161-162
__device__ __forceinline__ void eq_channel(float yr, float yi, float xr, float xi, float& hr, float& hi) {
    cdiv_np(yr, yi, __fadd_rn(xr, EQ_EPS), xi, hr, hi);
}
// ZF: X_hat = Y / (H + eps)                                      :83-84
__device__ __forceinline__ void eq_zf(float yr, float yi, float hr, float hi, float& zr, float& zi) {
    cdiv_np(yr, yi, __fadd_rn(hr, EQ_EPS), hi, zr, zi);
}
// MMSE: X_hat = conj(H) / (|H|^2 + 1/snr) * Y                    :193-201   (complex / real = multiply by the reciprocal)
__device__ __forceinline__ void eq_mmse(float yr, float yi, float hr, float hi, float inv_snr, float& mr, float& mi) {
    const float a = (float)sqrt((double)hr * (double)hr + (double)hi * (double)hi);      // np.abs(H)
    const float den = __fadd_rn(__fmul_rn(a, a), inv_snr);
    const float scl = __fdiv_rn(1.0f, den);
    const float fr = __fmul_rn(hr, scl), fi = __fmul_rn(-hi, scl);
    mr = __fsub_rn(__fmul_rn(fr, yr), __fmul_rn(fi, yi));
    mi = __fadd_rn(__fmul_rn(fr, yi), __fmul_rn(fi, yr));
}
// 1 / 10^(snr_db/10) formed in double like the Python scalars, then narrowed where NumPy narrows it
__device__ __forceinline__ float inv_snr_linear(float snr_db) { return (float)(1.0 / pow(10.0, (double)snr_db / 10.0)); }

}  // namespace og
