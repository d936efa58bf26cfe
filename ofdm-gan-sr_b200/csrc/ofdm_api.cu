// libofdmgan modulator entry points: QAMModulator('QPSK') and OFDMModulator of utils/ofdm_utils.py as batched kernels
// (one constellation symbol / one OFDM symbol per thread; the N-point transforms run in registers, chan_device.cuh).
//   ofdmgan_qpsk_modulate / ofdmgan_qpsk_demodulate / ofdmgan_ofdm_modulate / ofdmgan_ofdm_demodulate
// HBM-bound element-wise work: coalesced float2 / byte accesses, grids sized as a multiple of the SM count.
#include "chan_device.cuh"

namespace og {

constexpr float INV_SQRT2 = 0.70710678118654752f;

// utils/ofdm_utils.py:105-109: index = 2*b1 + b0 into [1+1j, 1-1j, -1+1j, -1-1j]/sqrt2 (MSB -> Re sign, LSB -> Im sign)
__global__ void k_qpsk_mod(const uint8_t* __restrict__ bits, float2* __restrict__ sym, int64_t n) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const uchar2 b = reinterpret_cast<const uchar2*>(bits)[i];
        sym[i] = make_float2(b.x ? -INV_SQRT2 : INV_SQRT2, b.y ? -INV_SQRT2 : INV_SQRT2);
    }
}

// :195-222: argmin_k |s - c_k|^2, np.argmin returns the lowest index on ties -> an exact 0 decides "bit 0"
__global__ void k_qpsk_demod(const float2* __restrict__ sym, uint8_t* __restrict__ bits, int64_t n) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const float2 s = sym[i];
        reinterpret_cast<uchar2*>(bits)[i] = make_uchar2(s.x < 0.f ? 1 : 0, s.y < 0.f ? 1 : 0);
    }
}

// QAM16 / QAM64 (utils/ofdm_utils.py:137-161): levels = -sqrtM+1, -sqrtM+3, ..., sqrtM-1; I, Q = meshgrid(levels, levels);
// constellation = (I + jQ).flatten() / norm  =>  point[idx] = (levels[idx % sqrtM] + j levels[idx / sqrtM]) / norm.
template <int BPS>
__global__ void k_qam_mod(const uint8_t* __restrict__ bits, float2* __restrict__ sym, int64_t n) {
    constexpr int SQ = 1 << (BPS / 2);
    constexpr float INV_NORM = BPS == 4 ? 0.31622776601683794f : 0.1543033499620919f;     // 1/sqrt(10), 1/sqrt(42)
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        int idx = 0;
#pragma unroll
        for (int b = 0; b < BPS; ++b) idx = (idx << 1) | (bits[i * BPS + b] & 1);
        sym[i] = make_float2((float)(2 * (idx % SQ) - SQ + 1) * INV_NORM, (float)(2 * (idx / SQ) - SQ + 1) * INV_NORM);
    }
}

// nearest level per axis; a tie between two levels goes to the lower one (np.argmin returns the first minimum, and a lower
// level index on either axis is a lower constellation index)
template <int BPS>
__global__ void k_qam_demod(const float2* __restrict__ sym, uint8_t* __restrict__ bits, int64_t n) {
    constexpr int SQ = 1 << (BPS / 2);
    constexpr float NORM = BPS == 4 ? 3.1622776601683795f : 6.48074069840786f;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const float2 s = sym[i];
        int axis[2];
#pragma unroll
        for (int a = 0; a < 2; ++a) {
            const float t = ((a ? s.y : s.x) * NORM + (float)(SQ - 1)) * 0.5f;       // position in level-index units
            int k = (int)ceilf(t - 0.5f);                                            // round half DOWN: ties to the lower level
            axis[a] = k < 0 ? 0 : (k > SQ - 1 ? SQ - 1 : k);
        }
        const int idx = axis[1] * SQ + axis[0];
#pragma unroll
        for (int b = 0; b < BPS; ++b) bits[i * BPS + b] = (idx >> (BPS - 1 - b)) & 1;
    }
}

// 32- and 64-point transforms (OFDMModulator's class default is 64 sub-carriers): same radix-2 network as fft_inplace with a
// 64-point twiddle table; cos(2 pi k / 64) for k = 0..16, the rest by symmetry
__device__ __forceinline__ constexpr float tw64_q(int k) {
    return k == 0 ? 1.f : k == 1 ? 0.99518472667219693f : k == 2 ? 0.98078528040323043f : k == 3 ? 0.95694033573220882f
         : k == 4 ? 0.92387953251128674f : k == 5 ? 0.88192126434835505f : k == 6 ? 0.83146961230254524f
         : k == 7 ? 0.77301045336273699f : k == 8 ? 0.70710678118654752f : k == 9 ? 0.63439328416364549f
         : k == 10 ? 0.55557023301960218f : k == 11 ? 0.47139673682599764f : k == 12 ? 0.38268343236508977f
         : k == 13 ? 0.29028467725446233f : k == 14 ? 0.19509032201612825f : k == 15 ? 0.098017140329560604f : 0.f;
}
__device__ __forceinline__ constexpr float tw64_c(int k) { return k <= 16 ? tw64_q(k) : -tw64_q(32 - k); }   // k in 0..31
__device__ __forceinline__ constexpr float tw64_s(int k) { return k <= 16 ? tw64_q(16 - k) : tw64_q(k - 16); }

template <int N, int SIGN>
__device__ __forceinline__ void fft_any(float (&re)[N], float (&im)[N]) {
    if constexpr (N <= 16) {
        fft_inplace<N, SIGN>(re, im);
    } else {
        constexpr int LOG = N == 64 ? 6 : 5;
        float tr[N], ti[N];
#pragma unroll
        for (int i = 0; i < N; ++i) { tr[i] = re[bitrev(i, LOG)]; ti[i] = im[bitrev(i, LOG)]; }
#pragma unroll
        for (int st = 1; st <= LOG; ++st) {
            const int h = 1 << (st - 1);
#pragma unroll
            for (int q = 0; q < N / 2; ++q) {
                const int j = q & (h - 1), a = ((q >> (st - 1)) << st) + j, b = a + h;
                const int tw = j * (32 >> (st - 1));                 // index into the 64-point table
                const float wr = tw64_c(tw), wi = (float)SIGN * tw64_s(tw);
                float xr, xi;
                if (tw == 0) { xr = tr[b]; xi = ti[b]; }
                else if (tw == 16) { xr = -(float)SIGN * ti[b]; xi = (float)SIGN * tr[b]; }
                else { xr = tr[b] * wr - ti[b] * wi; xi = tr[b] * wi + ti[b] * wr; }
                tr[b] = tr[a] - xr; ti[b] = ti[a] - xi;
                tr[a] = tr[a] + xr; ti[a] = ti[a] + xi;
            }
        }
#pragma unroll
        for (int i = 0; i < N; ++i) { re[i] = tr[i]; im[i] = ti[i]; }
    }
}

__global__ void k_unpackbits(const uint8_t* __restrict__ bytes, int64_t n_bits, uint8_t* __restrict__ bits) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n_bits; i += (int64_t)gridDim.x * blockDim.x)
        bits[i] = (bytes[i >> 3] >> (7 - (int)(i & 7))) & 1u;
}
__global__ void k_packbits(const uint8_t* __restrict__ bits, int64_t n_bytes, uint8_t* __restrict__ bytes) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n_bytes; i += (int64_t)gridDim.x * blockDim.x) {
        unsigned v = 0;
#pragma unroll
        for (int b = 0; b < 8; ++b) v = (v << 1) | (bits[i * 8 + b] != 0 ? 1u : 0u);
        bytes[i] = (uint8_t)v;
    }
}

__device__ __forceinline__ bool is_pilot(int k, int spacing) { return spacing > 0 && (k % spacing) == 0; }

// :281-329.  One OFDM symbol per thread.
template <int N>
__global__ void __launch_bounds__(128) k_ofdm_mod(const float2* __restrict__ sym, int64_t n_symbols, int64_t n_ofdm, int cp, int spacing,
                                                  int n_data, float pr, float pi, float2* __restrict__ out) {
    for (int64_t s = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; s < n_ofdm; s += (int64_t)gridDim.x * blockDim.x) {
        float Xr[N], Xi[N];
        int64_t d = s * n_data;
#pragma unroll
        for (int k = 0; k < N; ++k) {
            if (is_pilot(k, spacing)) { Xr[k] = pr; Xi[k] = pi; }
            else {
                const float2 v = d < n_symbols ? sym[d] : make_float2(0.f, 0.f);      // zero padding of the last symbol
                Xr[k] = v.x; Xi[k] = v.y;
                ++d;
            }
        }
        fft_any<N, +1>(Xr, Xi);                              // unscaled inverse = np.fft.ifft * N
        float2* o = out + s * (N + cp);
        for (int i = 0; i < cp; ++i) {                           // cyclic prefix: the last cp samples first
            float r = 0.f, q = 0.f;
#pragma unroll
            for (int j = 0; j < N; ++j) if (j == N - cp + i) { r = Xr[j]; q = Xi[j]; }
            o[i] = make_float2(r, q);
        }
#pragma unroll
        for (int j = 0; j < N; ++j) o[cp + j] = make_float2(Xr[j], Xi[j]);
    }
}

// :331-371.  data = fft / N on the data bins; channel estimate = pilot bins / pilot_value
template <int N>
__global__ void __launch_bounds__(128) k_ofdm_demod(const float2* __restrict__ sig, int64_t n_ofdm, int cp, int spacing, int n_data,
                                                    int n_pilot, float pr, float pi, float2* __restrict__ data, float2* __restrict__ chan) {
    const float inv_p = 1.0f / (pr * pr + pi * pi);
    for (int64_t s = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; s < n_ofdm; s += (int64_t)gridDim.x * blockDim.x) {
        float Xr[N], Xi[N];
        const float2* in = sig + s * (N + cp) + cp;
#pragma unroll
        for (int j = 0; j < N; ++j) { const float2 v = in[j]; Xr[j] = v.x; Xi[j] = v.y; }
        fft_any<N, -1>(Xr, Xi);
        int64_t d = s * n_data, p = s * n_pilot;
#pragma unroll
        for (int k = 0; k < N; ++k) {
            const float r = Xr[k] * (1.0f / N), q = Xi[k] * (1.0f / N);
            if (is_pilot(k, spacing)) {
                if (chan) chan[p] = make_float2((r * pr + q * pi) * inv_p, (q * pr - r * pi) * inv_p);
                ++p;
            } else {
                data[d++] = make_float2(r, q);
            }
        }
    }
}

static int pilots_of(int n_fft, int spacing) { return spacing > 0 ? (n_fft + spacing - 1) / spacing : 0; }

}  // namespace og

using namespace og;

extern "C" {

int ofdmgan_qpsk_modulate(const uint8_t* bits_dev, float* sym_dev, int64_t n_symbols, void* stream) {
    if (n_symbols < 0) return OFDMGAN_E_ARG;
    if (n_symbols == 0) return 0;
    if (!bits_dev || !sym_dev || (reinterpret_cast<uintptr_t>(bits_dev) & 1u) || (reinterpret_cast<uintptr_t>(sym_dev) & 7u)) return OFDMGAN_E_ARG;
    k_qpsk_mod<<<grid_for(n_symbols, 256, 8), 256, 0, (cudaStream_t)stream>>>(bits_dev, reinterpret_cast<float2*>(sym_dev), n_symbols);
    return (int)cudaGetLastError();
}

int ofdmgan_qpsk_demodulate(const float* sym_dev, uint8_t* bits_dev, int64_t n_symbols, void* stream) {
    if (n_symbols < 0) return OFDMGAN_E_ARG;
    if (n_symbols == 0) return 0;
    if (!bits_dev || !sym_dev || (reinterpret_cast<uintptr_t>(bits_dev) & 1u) || (reinterpret_cast<uintptr_t>(sym_dev) & 7u)) return OFDMGAN_E_ARG;
    k_qpsk_demod<<<grid_for(n_symbols, 256, 8), 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const float2*>(sym_dev), bits_dev, n_symbols);
    return (int)cudaGetLastError();
}

// np.unpackbits / np.packbits (MSB first), the byte <-> bit step of ImageOFDMConverter (utils/ofdm_utils.py:910, :986)
int ofdmgan_unpackbits(const uint8_t* bytes_dev, int64_t n_bytes, uint8_t* bits_dev, void* stream) {
    if (n_bytes < 0) return OFDMGAN_E_ARG;
    if (n_bytes == 0) return 0;
    if (!bytes_dev || !bits_dev) return OFDMGAN_E_ARG;
    k_unpackbits<<<grid_for(n_bytes * 8, 256, 8), 256, 0, (cudaStream_t)stream>>>(bytes_dev, n_bytes * 8, bits_dev);
    return (int)cudaGetLastError();
}

int ofdmgan_packbits(const uint8_t* bits_dev, int64_t n_bytes, uint8_t* bytes_dev, void* stream) {
    if (n_bytes < 0) return OFDMGAN_E_ARG;
    if (n_bytes == 0) return 0;
    if (!bytes_dev || !bits_dev) return OFDMGAN_E_ARG;
    k_packbits<<<grid_for(n_bytes, 256, 8), 256, 0, (cudaStream_t)stream>>>(bits_dev, n_bytes, bytes_dev);
    return (int)cudaGetLastError();
}

int ofdmgan_qam_modulate(const uint8_t* bits_dev, float* sym_dev, int64_t n_symbols, int bits_per_symbol, void* stream) {
    if (bits_per_symbol == 2) return ofdmgan_qpsk_modulate(bits_dev, sym_dev, n_symbols, stream);
    if (n_symbols < 0 || (bits_per_symbol != 4 && bits_per_symbol != 6)) return OFDMGAN_E_ARG;
    if (n_symbols == 0) return 0;
    if (!bits_dev || !sym_dev || (reinterpret_cast<uintptr_t>(sym_dev) & 7u)) return OFDMGAN_E_ARG;
    const int grid = grid_for(n_symbols, 256, 8);
    if (bits_per_symbol == 4) k_qam_mod<4><<<grid, 256, 0, (cudaStream_t)stream>>>(bits_dev, reinterpret_cast<float2*>(sym_dev), n_symbols);
    else k_qam_mod<6><<<grid, 256, 0, (cudaStream_t)stream>>>(bits_dev, reinterpret_cast<float2*>(sym_dev), n_symbols);
    return (int)cudaGetLastError();
}

int ofdmgan_qam_demodulate(const float* sym_dev, uint8_t* bits_dev, int64_t n_symbols, int bits_per_symbol, void* stream) {
    if (bits_per_symbol == 2) return ofdmgan_qpsk_demodulate(sym_dev, bits_dev, n_symbols, stream);
    if (n_symbols < 0 || (bits_per_symbol != 4 && bits_per_symbol != 6)) return OFDMGAN_E_ARG;
    if (n_symbols == 0) return 0;
    if (!bits_dev || !sym_dev || (reinterpret_cast<uintptr_t>(sym_dev) & 7u)) return OFDMGAN_E_ARG;
    const int grid = grid_for(n_symbols, 256, 8);
    if (bits_per_symbol == 4) k_qam_demod<4><<<grid, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const float2*>(sym_dev), bits_dev, n_symbols);
    else k_qam_demod<6><<<grid, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const float2*>(sym_dev), bits_dev, n_symbols);
    return (int)cudaGetLastError();
}

int ofdmgan_ofdm_modulate(const float* sym_dev, int64_t n_symbols, int n_fft, int cp_len, int pilot_spacing, float pilot_re, float pilot_im,
                          float* out_dev, void* stream) {
    if (n_symbols < 0 || cp_len < 0 || pilot_spacing < 0) return OFDMGAN_E_ARG;
    if (n_fft != 8 && n_fft != 16 && n_fft != 32 && n_fft != 64) return n_fft > 0 ? OFDMGAN_E_UNSUPPORTED : OFDMGAN_E_ARG;
    if (cp_len > n_fft) return OFDMGAN_E_ARG;
    const int n_data = n_fft - pilots_of(n_fft, pilot_spacing);
    if (n_data < 1) return OFDMGAN_E_ARG;
    const int64_t n_ofdm = (n_symbols + n_data - 1) / n_data;
    if (n_ofdm == 0) return 0;
    if (!sym_dev || !out_dev || (reinterpret_cast<uintptr_t>(sym_dev) & 7u) || (reinterpret_cast<uintptr_t>(out_dev) & 7u)) return OFDMGAN_E_ARG;
    // reference behaviour kept on purpose: cp = time_symbols[:, -cp_length:] with cp_length == 0 is the WHOLE symbol
    // (numpy's -0 == 0), so the reference emits every symbol twice (utils/ofdm_utils.py:323-324)
    if (cp_len == 0) cp_len = n_fft;
    const int grid = grid_for(n_ofdm, 128, 8);
    cudaStream_t s = (cudaStream_t)stream;
    const float2* in = reinterpret_cast<const float2*>(sym_dev);
    float2* out = reinterpret_cast<float2*>(out_dev);
    if (n_fft == 64) k_ofdm_mod<64><<<grid, 128, 0, s>>>(in, n_symbols, n_ofdm, cp_len, pilot_spacing, n_data, pilot_re, pilot_im, out);
    else if (n_fft == 32) k_ofdm_mod<32><<<grid, 128, 0, s>>>(in, n_symbols, n_ofdm, cp_len, pilot_spacing, n_data, pilot_re, pilot_im, out);
    else if (n_fft == 16) k_ofdm_mod<16><<<grid, 128, 0, s>>>(in, n_symbols, n_ofdm, cp_len, pilot_spacing, n_data, pilot_re, pilot_im, out);
    else k_ofdm_mod<8><<<grid, 128, 0, s>>>(in, n_symbols, n_ofdm, cp_len, pilot_spacing, n_data, pilot_re, pilot_im, out);
    return (int)cudaGetLastError();
}

int ofdmgan_ofdm_demodulate(const float* sig_dev, int64_t n_ofdm, int n_fft, int cp_len, int pilot_spacing, float pilot_re, float pilot_im,
                            float* data_dev, float* chan_dev, void* stream) {
    if (n_ofdm < 0 || cp_len < 0 || pilot_spacing < 0) return OFDMGAN_E_ARG;
    if (n_fft != 8 && n_fft != 16 && n_fft != 32 && n_fft != 64) return n_fft > 0 ? OFDMGAN_E_UNSUPPORTED : OFDMGAN_E_ARG;
    if (cp_len > n_fft || (pilot_re == 0.f && pilot_im == 0.f && pilot_spacing > 0 && chan_dev)) return OFDMGAN_E_ARG;
    const int n_pilot = pilots_of(n_fft, pilot_spacing), n_data = n_fft - n_pilot;
    if (n_data < 1) return OFDMGAN_E_ARG;
    if (n_ofdm == 0) return 0;
    if (!sig_dev || !data_dev || (reinterpret_cast<uintptr_t>(sig_dev) & 7u) || (reinterpret_cast<uintptr_t>(data_dev) & 7u)) return OFDMGAN_E_ARG;
    const int grid = grid_for(n_ofdm, 128, 8);
    cudaStream_t s = (cudaStream_t)stream;
    const float2* in = reinterpret_cast<const float2*>(sig_dev);
    float2 *data = reinterpret_cast<float2*>(data_dev), *chan = reinterpret_cast<float2*>(chan_dev);
    if (n_fft == 64) k_ofdm_demod<64><<<grid, 128, 0, s>>>(in, n_ofdm, cp_len, pilot_spacing, n_data, n_pilot, pilot_re, pilot_im, data, chan);
    else if (n_fft == 32) k_ofdm_demod<32><<<grid, 128, 0, s>>>(in, n_ofdm, cp_len, pilot_spacing, n_data, n_pilot, pilot_re, pilot_im, data, chan);
    else if (n_fft == 16) k_ofdm_demod<16><<<grid, 128, 0, s>>>(in, n_ofdm, cp_len, pilot_spacing, n_data, n_pilot, pilot_re, pilot_im, data, chan);
    else k_ofdm_demod<8><<<grid, 128, 0, s>>>(in, n_ofdm, cp_len, pilot_spacing, n_data, n_pilot, pilot_re, pilot_im, data, chan);
    return (int)cudaGetLastError();
}

}  // extern "C"
