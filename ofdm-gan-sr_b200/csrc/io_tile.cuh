// Warp-private staging of frame tiles between HBM and registers.
//
// One thread owns one frame, but a frame is 128 B (fp32) or 64 B (int16): per-thread 16-byte accesses at a 128-byte
// stride would touch 32 different lines per warp instruction and bind the kernel on L1 wavefronts.  Instead each
// warp moves its 32 consecutive frames as one contiguous 4 KB (2 KB) run with fully coalesced 128-bit accesses,
// through a 16-byte-chunk XOR swizzle in shared memory so that both the row-major fill and the per-thread frame
// read are bank-conflict free.  Only __syncwarp() is needed: warps never share a tile.
#pragma once
#include "common.cuh"

namespace og {

// ---- fp32 frames: 8 chunks of 16 B per frame; chunk c of frame f lives at slot f*8 + (c ^ (f & 7))
__device__ __forceinline__ void tile_load_f32(const float* __restrict__ g, int64_t frame_base, int64_t B, float4* wsm,
                                              int lane, float (&x)[2][16]) {
    const float4* src = reinterpret_cast<const float4*>(g) + frame_base * 8;
    const int64_t limit = (B - frame_base) * 8;                 // chunks that exist in this tile
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        const int idx = r * 32 + lane, f = idx >> 3, c = idx & 7;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (idx < limit) v = __ldg(src + idx);
        wsm[f * 8 + (c ^ (f & 7))] = v;
    }
    __syncwarp();
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        float4 v = wsm[lane * 8 + (c ^ (lane & 7))];
        const int row = c >> 2, col = (c & 3) * 4;
        x[row][col] = v.x; x[row][col + 1] = v.y; x[row][col + 2] = v.z; x[row][col + 3] = v.w;
    }
    __syncwarp();
}

// Two-phase variant for kernels that keep the warp's tile resident in shared memory and re-read a frame whenever it
// is needed again (cheaper than holding 32 more registers live across a backward pass).
__device__ __forceinline__ void tile_fill_f32(const float* __restrict__ g, int64_t frame_base, int64_t B, float4* wsm, int lane) {
    const float4* src = reinterpret_cast<const float4*>(g) + frame_base * 8;
    const int64_t limit = (B - frame_base) * 8;
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        const int idx = r * 32 + lane, f = idx >> 3, c = idx & 7;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (idx < limit) v = __ldg(src + idx);
        wsm[f * 8 + (c ^ (f & 7))] = v;
    }
}
__device__ __forceinline__ void tile_read_f32(const float4* wsm, int lane, float (&x)[2][16]) {
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        float4 v = wsm[lane * 8 + (c ^ (lane & 7))];
        const int row = c >> 2, col = (c & 3) * 4;
        x[row][col] = v.x; x[row][col + 1] = v.y; x[row][col + 2] = v.z; x[row][col + 3] = v.w;
    }
}
// the thread's own frame back into its resident tile (e.g. the generator output the critic pass re-reads)
__device__ __forceinline__ void tile_write_f32(float4* wsm, int lane, const float (&y)[2][16]) {
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        const int row = c >> 2, col = (c & 3) * 4;
        wsm[lane * 8 + (c ^ (lane & 7))] = make_float4(y[row][col], y[row][col + 1], y[row][col + 2], y[row][col + 3]);
    }
}
// resident tile -> HBM, coalesced
__device__ __forceinline__ void tile_drain_f32(float* __restrict__ g, int64_t frame_base, int64_t B, const float4* wsm, int lane) {
    float4* dst = reinterpret_cast<float4*>(g) + frame_base * 8;
    const int64_t limit = (B - frame_base) * 8;
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        const int idx = r * 32 + lane, f = idx >> 3, c = idx & 7;
        if (idx < limit) dst[idx] = wsm[f * 8 + (c ^ (f & 7))];
    }
}

__device__ __forceinline__ void tile_store_f32(float* __restrict__ g, int64_t frame_base, int64_t B, float4* wsm, int lane,
                                               const float (&y)[2][16]) {
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        const int row = c >> 2, col = (c & 3) * 4;
        wsm[lane * 8 + (c ^ (lane & 7))] = make_float4(y[row][col], y[row][col + 1], y[row][col + 2], y[row][col + 3]);
    }
    __syncwarp();
    float4* dst = reinterpret_cast<float4*>(g) + frame_base * 8;
    const int64_t limit = (B - frame_base) * 8;
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        const int idx = r * 32 + lane, f = idx >> 3, c = idx & 7;
        if (idx < limit) dst[idx] = wsm[f * 8 + (c ^ (f & 7))];
    }
    __syncwarp();
}

// ---- int16 frames: 4 chunks of 16 B per frame; chunk c of frame f lives at slot f*4 + (c ^ ((f >> 1) & 3))
__device__ __forceinline__ void tile_load_i16(const int16_t* __restrict__ g, int64_t frame_base, int64_t B, uint4* wsm,
                                              int lane, float (&x)[2][16]) {
    const uint4* src = reinterpret_cast<const uint4*>(g) + frame_base * 4;
    const int64_t limit = (B - frame_base) * 4;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const int idx = r * 32 + lane, f = idx >> 2, c = idx & 3;
        uint4 v = make_uint4(0u, 0u, 0u, 0u);
        if (idx < limit) v = __ldg(src + idx);
        wsm[f * 4 + (c ^ ((f >> 1) & 3))] = v;
    }
    __syncwarp();
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        uint4 v = wsm[lane * 4 + (c ^ ((lane >> 1) & 3))];
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
        const int row = c >> 1, col = (c & 1) * 8;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            x[row][col + 2 * j] = (float)(int16_t)(w[j] & 0xffffu);
            x[row][col + 2 * j + 1] = (float)((int32_t)w[j] >> 16);
        }
    }
    __syncwarp();
}

// y holds exact integers in [-32768, 32767]
__device__ __forceinline__ uint32_t pack_i16x2(float lo, float hi) {
    // (v + 1.5*2^23) leaves v's two's-complement low bits in the mantissa: no F2I needed
    uint32_t a = __float_as_uint(lo + 12582912.0f), b = __float_as_uint(hi + 12582912.0f);
    return __byte_perm(a, b, 0x5410);
}

__device__ __forceinline__ void tile_store_i16(int16_t* __restrict__ g, int64_t frame_base, int64_t B, uint4* wsm, int lane,
                                               const float (&y)[2][16]) {
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        const int row = c >> 1, col = (c & 1) * 8;
        uint4 v;
        v.x = pack_i16x2(y[row][col + 0], y[row][col + 1]);
        v.y = pack_i16x2(y[row][col + 2], y[row][col + 3]);
        v.z = pack_i16x2(y[row][col + 4], y[row][col + 5]);
        v.w = pack_i16x2(y[row][col + 6], y[row][col + 7]);
        wsm[lane * 4 + (c ^ ((lane >> 1) & 3))] = v;
    }
    __syncwarp();
    uint4* dst = reinterpret_cast<uint4*>(g) + frame_base * 4;
    const int64_t limit = (B - frame_base) * 4;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const int idx = r * 32 + lane, f = idx >> 2, c = idx & 3;
        if (idx < limit) dst[idx] = wsm[f * 4 + (c ^ ((f >> 1) & 3))];
    }
    __syncwarp();
}

}  // namespace og
