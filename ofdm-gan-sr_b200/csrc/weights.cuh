// Constant-memory weight images and their stream-ordered refresh.  Included by every TU that runs a network
// (each TU owns its own __constant__ bank, so each gets private copies of these symbols).
//
// There is ONE image per network at a FIXED constant-bank address, so after unrolling every weight is an immediate
// constant operand of its FFMA (FFMA R, R, c[0x3][imm], R): no load instruction and no register per weight - this is
// what lets the training kernels fit 128 registers.  Calls on different streams are ordered against each other by
// CallGuard (common.cuh), which is what keeps a single image safe.
//
// G image (OG_G_IMG floats): raw enc1 / bottleneck weights + biases, and the two convolutions that follow a
// nearest x2 upsample (dec1, out_conv; models/generator.py:196-205) FOLDED over the duplicated samples:
//   z[2p]   = w0*a[p-1] + (w1+w2)*a[p]        z[2p+1] = (w0+w1)*a[p] + w2*a[p+1]
// i.e. 4 folded taps {w0, w1+w2, w0+w1, w2} per (oc,ic) and 2 MACs per output instead of 3 (1344 instead of
// 1728 MACs per frame).  The raw taps are kept too for the weight-gradient kernels.
// D image (OG_D_IMG floats): the 521 raw critic parameters (models/discriminator.py:78-100).
// Q image (OG_Q_IMG floats): the fixed-point generator ROM as exact floats: weights[0..225]/128, biases at 232.
#pragma once
#include "common.cuh"

namespace og {

// ---- G image layout
constexpr int GI_ENC_W = 0;      // [4][2][3]
constexpr int GI_ENC_B = 24;     // [4]
constexpr int GI_BN_W = 28;      // [8][4][3]
constexpr int GI_BN_B = 124;     // [8]
constexpr int GI_DEC_F = 132;    // [4][8][4] folded
constexpr int GI_DEC_B = 260;    // [4]
constexpr int GI_OUT_F = 264;    // [2][4][4] folded
constexpr int GI_OUT_B = 296;    // [2]
// pair-interleaved copies of the four weight blocks, [oc/2][ic][k][2]: one 64-bit uniform load feeds a packed FFMA2
constexpr int GI2_ENC = 304, GI2_BN = 328, GI2_DEC = 424, GI2_OUT = 552;   // 24 + 96 + 128 + 32 floats
// out_conv once more, folded taps (pair-interleaved) and bias multiplied by 2 log2(e): tanh(v) = 1 - 2 / (1 + 2^(v 2 log2 e)) then
// needs no multiply in the fused simulator (sim_ws.cu)
constexpr int GI2_OUT_T = 584, GI_OUT_BT = 616;                            // 32 + 2 floats
constexpr float G_TANH_SCALE = 2.8853900817779268f;
constexpr int OG_G_IMG = 624;
// raw parameter offsets (torch order, include/ofdmgan.h)
constexpr int GP_ENC_W = 0, GP_ENC_B = 24, GP_BN_W = 28, GP_BN_B = 124, GP_DEC_W = 132, GP_DEC_B = 228, GP_OUT_W = 232,
              GP_OUT_B = 256;
constexpr int DP_C1_W = 0, DP_C1_B = 96, DP_C2_W = 104, DP_C2_B = 488, DP_FC_W = 504, DP_FC_B = 520;
// D image: the 521 raw parameters (+pad to 528), then pair-interleaved copies for packed FFMA2:
//   DI2_C1   conv1.weight as [oc/2][ic][k][2]   (two output channels per instruction: forward, d/d input reduction)
//   DI2_C2   conv2.weight as [oc/2][ic][k][2]   (forward)
//   DI2_C2T  conv2.weight as [oc][ic/2][k][2]   (transposed conv of the backward pass: two INPUT channels per instruction)
constexpr int DI2_C1 = 528, DI2_C2 = 624, DI2_C2T = 1008;                 // 96 + 384 + 384 floats
constexpr int OG_D_IMG = 1392;
constexpr int OG_Q_IMG = 512;
constexpr int QI_BIAS = 232;
constexpr int QI_BIASM = 480;    // [18] 1.5*2^23 + bias: accumulator start values of the packed spec path (pairs stay adjacent)
// second half of the Q image: the same weights with the two output channels of a pair interleaved, [oc/2][ic][k][2], so one
// 64-bit uniform load feeds a packed FFMA2 (two output channels per instruction)
constexpr int QI2_ENC = 256, QI2_BN = 280, QI2_DEC = 376, QI2_OUT = 472;     // 24 + 96 + 96 + 8 floats

static __constant__ __align__(16) float c_g[OG_G_IMG];
static __constant__ __align__(16) float c_d[OG_D_IMG];
static __constant__ __align__(16) float c_q[OG_Q_IMG];

// value of entry i (< GI2_ENC) of the G image from the raw parameters
__host__ __device__ __forceinline__ float g_img_base(const float* __restrict__ p, int i) {
    if (i < 132) return p[i];                                              // enc1 + bottleneck, verbatim
    if (i < 132 + 128) {                                                   // dec1 folded
        const int j = i - 132, pair = j >> 2, t = j & 3;
        const float* w = p + GP_DEC_W + pair * 3;
        return t == 0 ? w[0] : t == 1 ? w[1] + w[2] : t == 2 ? w[0] + w[1] : w[2];
    }
    if (i < 264) return p[GP_DEC_B + (i - 260)];
    if (i < 296) {
        const int j = i - 264, pair = j >> 2, t = j & 3;
        const float* w = p + GP_OUT_W + pair * 3;
        return t == 0 ? w[0] : t == 1 ? w[1] + w[2] : t == 2 ? w[0] + w[1] : w[2];
    }
    if (i < 298) return p[GP_OUT_B + (i - 296)];
    return 0.f;
}

// entry i of the whole G image
__host__ __device__ __forceinline__ float g_img_entry(const float* __restrict__ p, int i) {
    if (i < GI2_ENC) return g_img_base(p, i);
    if (i >= OG_G_IMG) return 0.f;
    if (i >= GI_OUT_BT) return i < GI_OUT_BT + 2 ? p[GP_OUT_B + (i - GI_OUT_BT)] * G_TANH_SCALE : 0.f;
    const float post = i >= GI2_OUT_T ? G_TANH_SCALE : 1.0f;
    if (i >= GI2_OUT_T) i -= GI2_OUT_T - GI2_OUT;
    // pair-interleaved copies: entry ((o2*IC + ic)*K + k)*2 + h  <-  base[((2*o2+h)*IC + ic)*K + k]
    int e, base, IC, K;
    if (i < GI2_BN) { e = i - GI2_ENC; base = GI_ENC_W; IC = 2; K = 3; }
    else if (i < GI2_DEC) { e = i - GI2_BN; base = GI_BN_W; IC = 4; K = 3; }
    else if (i < GI2_OUT) { e = i - GI2_DEC; base = GI_DEC_F; IC = 8; K = 4; }
    else { e = i - GI2_OUT; base = GI_OUT_F; IC = 4; K = 4; }
    const int h = e & 1, r = e >> 1, k = r % K, ic = (r / K) % IC, o2 = r / (K * IC);
    return g_img_base(p, base + ((2 * o2 + h) * IC + ic) * K + k) * post;
}

static __global__ void prep_g_image(const float* __restrict__ p, float* __restrict__ img) {
    const int i = threadIdx.x;
    if (i < OG_G_IMG) img[i] = g_img_entry(p, i);
}

// The same images BY VALUE, for weights that live in host memory (inference): the image is built on the host and travels in the
// kernel's parameter block (constant bank 0), so the call needs no preparation kernel, no copy into a __constant__ symbol and
// therefore no lock - calls on different streams are independent (include/ofdmgan.h, "Conventions").
struct GImage {
    float w[OG_G_IMG];
};
struct QImage {
    float w[OG_Q_IMG];
};
static inline void g_image_host(const float* params258_host, GImage& img) {
    for (int i = 0; i < OG_G_IMG; ++i) img.w[i] = g_img_entry(params258_host, i);
}
// is p a host pointer (anything cudaPointerGetAttributes does not call device / managed memory)
static inline bool is_host_pointer(const void* p) {
    cudaPointerAttributes at;
    const cudaError_t e = cudaPointerGetAttributes(&at, p);
    if (e != cudaSuccess) { (void)cudaGetLastError(); return true; }
    return !(at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged);
}

// entry i of the D image from the raw parameters
__device__ __forceinline__ float d_img_entry(const float* __restrict__ p, int i) {
    if (i < DI2_C1) return i < OFDMGAN_D_NPARAMS ? p[i] : 0.f;
    if (i < DI2_C2) {                                            // ((o2*4 + ic)*3 + k)*2 + h <- conv1[2*o2+h][ic][k]
        const int e = i - DI2_C1, h = e & 1, r = e >> 1, k = r % 3, ic = (r / 3) % 4, o2 = r / 12;
        return p[DP_C1_W + ((2 * o2 + h) * 4 + ic) * 3 + k];
    }
    if (i < DI2_C2T) {                                           // ((o2*8 + ic)*3 + k)*2 + h <- conv2[2*o2+h][ic][k]
        const int e = i - DI2_C2, h = e & 1, r = e >> 1, k = r % 3, ic = (r / 3) % 8, o2 = r / 24;
        return p[DP_C2_W + ((2 * o2 + h) * 8 + ic) * 3 + k];
    }
    const int e = i - DI2_C2T, h = e & 1, r = e >> 1, k = r % 3, i2 = (r / 3) % 4, oc = r / 12;   // ((oc*4 + i2)*3 + k)*2 + h <- conv2[oc][2*i2+h][k]
    return p[DP_C2_W + (oc * 8 + 2 * i2 + h) * 3 + k];
}

// the inverse: write parameter i (new value v) to every entry of the D image that holds it (1 to 3 entries)
__device__ __forceinline__ void d_img_scatter(float* __restrict__ img, int i, float v) {
    img[i] = v;
    if (i < DP_C1_B) {                                           // conv1.weight[oc][ic][k]
        const int k = i % 3, ic = (i / 3) % 4, oc = i / 12;
        img[DI2_C1 + (((oc >> 1) * 4 + ic) * 3 + k) * 2 + (oc & 1)] = v;
    } else if (i >= DP_C2_W && i < DP_C2_B) {                    // conv2.weight[oc][ic][k]
        const int e = i - DP_C2_W, k = e % 3, ic = (e / 3) % 8, oc = e / 24;
        img[DI2_C2 + (((oc >> 1) * 8 + ic) * 3 + k) * 2 + (oc & 1)] = v;
        img[DI2_C2T + ((oc * 4 + (ic >> 1)) * 3 + k) * 2 + (ic & 1)] = v;
    }
}

static __global__ void prep_d_image(const float* __restrict__ p, float* __restrict__ img) {
    const int i = threadIdx.x + blockIdx.x * blockDim.x;
    if (i < OG_D_IMG) img[i] = d_img_entry(p, i);
}

// params258 may be a host or a device pointer
static int upload_g(const float* params258, int slot, cudaStream_t s) {
    const float* dev = nullptr;
    int rc = to_device_f32(params258, OFDMGAN_G_NPARAMS, slot, 0, s, &dev);
    if (rc) return rc;
    void* img = nullptr;
    rc = scratch_for_slot(slot, OG_G_IMG * sizeof(float), 1, &img);
    if (rc) return rc;
    prep_g_image<<<1, OG_G_IMG, 0, s>>>(dev, (float*)img);
    OG_CHECK(cudaGetLastError());
    OG_CHECK(cudaMemcpyToSymbolAsync(c_g, img, OG_G_IMG * sizeof(float), 0,
                                     cudaMemcpyDeviceToDevice, s));
    return 0;
}

// copy the G image that an earlier upload_g on this device left in the staging buffer into THIS translation unit's constant bank
static int commit_g_image(int slot, cudaStream_t s) {
    void* img = nullptr;
    int rc = scratch_for_slot(slot, OG_G_IMG * sizeof(float), 1, &img);
    if (rc) return rc;
    OG_CHECK(cudaMemcpyToSymbolAsync(c_g, img, OG_G_IMG * sizeof(float), 0, cudaMemcpyDeviceToDevice, s));
    return 0;
}

// the staging buffer of the D image (device), for kernels that refresh the image themselves
static int d_image_staging(int slot, float** img) {
    void* p = nullptr;
    int rc = scratch_for_slot(slot, OG_D_IMG * sizeof(float), 3, &p);
    *img = (float*)p;
    return rc;
}
// copy an already prepared staging image into the constant bank
static int commit_d_image(int slot, cudaStream_t s) {
    float* img = nullptr;
    int rc = d_image_staging(slot, &img);
    if (rc) return rc;
    OG_CHECK(cudaMemcpyToSymbolAsync(c_d, img, OG_D_IMG * sizeof(float), 0, cudaMemcpyDeviceToDevice, s));
    return 0;
}

static int upload_d(const float* params521, int slot, cudaStream_t s) {
    const float* dev = nullptr;
    int rc = to_device_f32(params521, OFDMGAN_D_NPARAMS, slot, 2, s, &dev);
    if (rc) return rc;
    void* img = nullptr;
    rc = scratch_for_slot(slot, OG_D_IMG * sizeof(float), 3, &img);
    if (rc) return rc;
    prep_d_image<<<(OG_D_IMG + 255) / 256, 256, 0, s>>>(dev, (float*)img);
    OG_CHECK(cudaGetLastError());
    OG_CHECK(cudaMemcpyToSymbolAsync(c_d, img, OG_D_IMG * sizeof(float), 0,
                                     cudaMemcpyDeviceToDevice, s));
    return 0;
}

// ROMs are host pointers (weight_rom.v layout): converted on the host, exact (int8/128 and int16 are fp32-exact)
static inline void q_image_host(const int8_t* wrom_host, const int16_t* brom_host, QImage& qi) {
    float* img = qi.w;
    for (int i = 0; i < OG_Q_IMG; ++i) img[i] = 0.f;
    for (int i = 0; i < 226; ++i) img[i] = (float)wrom_host[i] * (1.0f / 128.0f);
    for (int i = 0; i < 18; ++i) img[QI_BIAS + i] = (float)brom_host[i];
    for (int i = 0; i < 18; ++i) img[QI_BIASM + i] = 12582912.0f + (float)brom_host[i];      // exact: |bias| < 2^15, ulp = 1
    auto pairs = [&](int dst, int wa, int OC, int IC, int K) {
        for (int o2 = 0; o2 < OC / 2; ++o2)
            for (int ic = 0; ic < IC; ++ic)
                for (int k = 0; k < K; ++k)
                    for (int h = 0; h < 2; ++h)
                        img[dst + ((o2 * IC + ic) * K + k) * 2 + h] = (float)wrom_host[wa + ((2 * o2 + h) * IC + ic) * K + k] * (1.0f / 128.0f);
    };
    pairs(QI2_ENC, 0, 4, 2, 3);
    pairs(QI2_BN, 24, 8, 4, 3);
    pairs(QI2_DEC, 120, 4, 8, 3);
    pairs(QI2_OUT, 216, 2, 4, 1);
}
static int upload_q(const int8_t* wrom_host, const int16_t* brom_host, int slot, cudaStream_t s) {
    QImage qi;
    q_image_host(wrom_host, brom_host, qi);
    // pageable source: the runtime stages it before returning, so the stack buffer may die afterwards
    OG_CHECK(cudaMemcpyToSymbolAsync(c_q, qi.w, sizeof qi.w, 0, cudaMemcpyHostToDevice, s));
    return 0;
}

}  // namespace og
