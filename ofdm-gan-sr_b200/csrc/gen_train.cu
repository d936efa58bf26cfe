// libofdmgan generator-side training kernels: the fused generator step of CWGAN-GP (train.py:282-299), the API-level
// MiniGenerator backward, and fused Adam.   ofdmgan_gen_step / ofdmgan_gen_bwd_f32 / ofdmgan_adam
#include "train_common.cuh"
#include "gen_stream.cuh"
#include "peer_comm.cuh"

namespace og {

// ------------------------------------------------------------------------------------------------ generator step
struct GenStepArgs {
    const float* clean;
    const float* noisy;
    const float* fake_in;     // HAVE_FAKE: G(noisy) under the installed generator image, computed by an earlier launch
    float* fake_out;          // nullable
    int64_t B;
    int slot;
    float slope;
    float adv_w, rec_w;
    float* partials;          // [grid][GS_SLOTS]
};

// Per sample: (1) fake = G(noisy), or - HAVE_FAKE - read from memory (a training iteration has already computed it for its critic
// updates with the same generator: train.py:228-232 and :285 are the same forward); (2) adversarial term through the critic ->
// d L / d fake, plus the L1 term, times tanh' while y is in registers; (3) G forward again with its tape and backward (recomputing
// 1.2k FMAs is cheaper than carrying 130 activations across the critic pass).  Register-lean rolled-loop passes (gen_stream.cuh,
// critic_stream.cuh).  THREE resident 4 KB tiles per warp - noisy (later parking space; refetched for the last gradient group) |
// fake -> dz4 rows -> parked rows | clean -> scratch / parked rows - and <= 128 registers: 4 CTAs per SM, so the 512 tiles of a
// 65,536-sample shard are ONE round (592 slots) instead of 1.15 rounds of 444.
// The kernel runs as ONE 512-thread CTA per SM (16 warps = the same residency) with a barrier between the stages of an item: see
// gs_stage_barrier.  Warps past the end of the batch run the item on all-dead lanes (every warp must reach every barrier).
constexpr int GENSTEP_THREADS = 512, GENSTEP_WARPS = GENSTEP_THREADS / 32;
constexpr int GENSTEP_PER_SM = 1;
constexpr size_t GENSTEP_SMEM = (size_t)3 * GENSTEP_THREADS * 8 * sizeof(float4) + (size_t)GSX_NG * GENSTEP_THREADS * sizeof(float);

template <bool HAVE_FAKE>
__global__ void __launch_bounds__(GENSTEP_THREADS, GENSTEP_PER_SM) k_gen_step(const __grid_constant__ GenStepArgs a) {
    extern __shared__ float4 sm[];
    float* sacc = reinterpret_cast<float*>(sm + 3 * GENSTEP_THREADS * 8);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float4* t_x = sm + warp * TILE4;
    float4* t_y = sm + (GENSTEP_WARPS + warp) * TILE4;
    float4* t_p = sm + (2 * GENSTEP_WARPS + warp) * TILE4;
    const float* WG = c_g;
    const float* WD = c_d;
    SAcc acc{sacc + threadIdx.x, GENSTEP_THREADS};
#pragma unroll
    for (int g = 0; g < GSX_NG; ++g) sacc[g * GENSTEP_THREADS + threadIdx.x] = 0.f;
    float s_d = 0.f, s_l1 = 0.f;
    const float rec_g = a.rec_w * 0.03125f;                      // rec_w / 32: l1_loss is a mean over B*32 elements
    const int64_t ntiles = (a.B + GENSTEP_THREADS - 1) / GENSTEP_THREADS;
#pragma unroll 1
    for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const int64_t base = t * GENSTEP_THREADS + warp * 32;     // (>= B for the surplus warps of the last tile: all lanes dead)
        const bool live = base + lane < a.B;
        __syncwarp();
        tile_fill_f32(a.noisy, base, a.B, t_x, lane);
        // (1) fake = G(noisy) -> t_y
        if (HAVE_FAKE) {
            tile_fill_f32(a.fake_in, base, a.B, t_y, lane);
        } else {
            __syncwarp();
            float a1[4][8], a2[8][4], sk[4][8];
            uint32_t z;
            gs_fwd<false, true, true>(WG, a.slope, t_x, t_y, t_p, lane, a1, a2, sk, z);
        }
        __syncwarp();
        if (!HAVE_FAKE && a.fake_out) {
            tile_drain_f32(a.fake_out, base, a.B, t_y, lane);
            __syncwarp();
        }
        tile_fill_f32(a.clean, base, a.B, t_p, lane);
        __syncthreads();
        // (2) critic on (fake, noisy): d(-adv_w * D)/d fake, + rec_w * sign(fake - clean)/32 = upstream gradient; times tanh'(z) =
        // 1 - y^2 while the y row is in registers -> dz4 rows, which replace the y rows in t_y
        {
            f32x2 dz1[4][8];
            {
                uint64_t m1, m2;
                const float score = cs_score_only(WD, a.slope, t_y, t_x, acc, lane, m1, m2);
                if (live) s_d += score;
                __syncthreads();
                cs_bwd_to_z1(WD, a.slope, live ? -a.adv_w : 0.f, m1, m2, dz1);
            }
            __syncthreads();
#pragma unroll 1
            for (int ic = 0; ic < 2; ++ic) {
                float row[16], y[16], c[16];
                cs_conv1T_row(WD, dz1, ic, row);
                row_read(t_y, lane, ic, y);
                row_read(t_p, lane, ic, c);
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const float e = y[i] - c[i];
                    if (live) {
                        s_l1 += fabsf(e);
                        row[i] += e > 0.f ? rec_g : (e < 0.f ? -rec_g : 0.f);       // l1_loss backward: sign(e)
                        row[i] *= fmaf(-y[i], y[i], 1.0f);
                    } else {
                        row[i] = 0.f;
                    }
                }
                row_write(t_y, lane, ic, row);
            }
        }
        __syncthreads();
        // (3) forward with tape (t_p is scratch again), backward
        {
            float a1[4][8], a2[8][4], sk[4][8];
            uint32_t z3pos;
            gs_fwd<true, false, true>(WG, a.slope, t_x, t_y, t_p, lane, a1, a2, sk, z3pos);
            gs_bwd<false, true, true>(WG, a.slope, t_x, t_y, t_x, t_p, lane, a1, a2, sk, z3pos, acc, a.noisy, base, a.B);
        }
    }
    {
        const float d = warp_sum(s_d), l = warp_sum(s_l1);
        if (lane == GS_S0 - 288) acc.add(9, d);
        if (lane == GS_S0 + 1 - 288) acc.add(9, l);
    }
    __syncthreads();
    float* row = a.partials + (size_t)blockIdx.x * GS_SLOTS;
    for (int s = threadIdx.x; s < GS_SLOTS; s += GENSTEP_THREADS) {
        const int grp = s >> 5, j = s & 31;
        float t = 0.f;
#pragma unroll
        for (int w = 0; w < GENSTEP_WARPS; ++w) t += sacc[grp * GENSTEP_THREADS + w * 32 + j];
        row[s] = t;
    }
}

// One block per accumulator group: fixed-order sum over the CTA rows, then slot -> parameter order.  The two
// upsample+conv layers accumulate gradients of their FOLDED taps {F0=w0, F1=w1+w2, F2=w0+w1, F3=w2}; the raw taps follow
// linearly: dw0 = dF0+dF2, dw1 = dF1+dF2, dw2 = dF1+dF3 (all four slots of a pair sit in the same group).
// stats (nullable): g_loss, adv_loss, rec_loss (train.py:301-305) + 3 pad, written by the block of group 9.
// slot j of group grp -> (parameter index, gradient sum) from the group's 32 totals; index -1 = no parameter in this slot
__device__ __forceinline__ int gs_param_of(int grp, int j, const double* total, double& v) {
    int i = -1;
    v = total[j];
    if (grp <= 4) {                                               // folded layers: thread (pair, t<3) emits raw tap k = t
        const int pair = j >> 2, k = j & 3;
        const double* F = total + pair * 4;
        if (k < 3) {
            v = k == 0 ? F[0] + F[2] : (k == 1 ? F[1] + F[2] : F[1] + F[3]);
            i = grp == 0 ? GP_OUT_W + pair * 3 + k : GP_DEC_W + ((grp - 1) * 8 + pair) * 3 + k;
        }
    } else if (grp <= 7) {
        i = GP_BN_W + (grp - 5) * 32 + j;
    } else if (grp == 8) {
        i = j < 24 ? GP_ENC_W + j : (j < 28 ? GP_ENC_B + (j - 24) : -1);
    } else {
        i = j < 8 ? GP_BN_B + j : (j < 12 ? GP_DEC_B + (j - 8) : (j < 14 ? GP_OUT_B + (j - 12) : -1));
    }
    return i;
}
__device__ __forceinline__ void gs_write_stats(const double* total, double inv_b, double adv_w, double rec_w, float* stats) {
    const double adv = -total[GS_S0 - 288] * inv_b, rec = total[GS_S0 + 1 - 288] * inv_b * 0.03125;
    stats[0] = (float)(adv_w * adv + rec_w * rec);
    stats[1] = (float)adv;
    stats[2] = (float)rec;
    stats[3] = 0.f;
    stats[4] = 0.f;
    stats[5] = 0.f;
}

__global__ void __launch_bounds__(1024) k_finalize_gen(const float* __restrict__ partials, int nblocks, double inv_b, double adv_w,
                                                       double rec_w, float* __restrict__ grads, float* __restrict__ stats) {
    __shared__ double red[32 * 32], total[32];
    const int grp = blockIdx.x, j = threadIdx.x;
    reduce_group_rows(partials, nblocks, GS_SLOTS, grp, red, total);
    if (grads && j < 32) {
        double v;
        const int i = gs_param_of(grp, j, total, v);
        if (i >= 0) grads[i] = (float)(v * inv_b);
    }
    if (grp == 9 && stats && j == 0) gs_write_stats(total, inv_b, adv_w, rec_w, stats);
}

// Tail of a generator update in ONE launch (as k_critic_tail): the fixed-order reduction of k_finalize_gen and - in the block that
// finishes last - the sum over the ranks through peer memory, Adam on the 258 parameters with the step count in device memory.
__global__ void __launch_bounds__(1024) k_gen_tail(const float* __restrict__ partials, int nblocks, double inv_b, double adv_w, double rec_w,
                                                   float* __restrict__ out, float* __restrict__ p, float* __restrict__ m,
                                                   float* __restrict__ v, double lr, double b1, double b2, double eps,
                                                   int32_t* __restrict__ step_dev, unsigned int* __restrict__ arrivals, PeerPtrs peers,
                                                   int rank, int world) {
    __shared__ double red[32 * 32], total[32];
    __shared__ unsigned int ticket;
    const int grp = blockIdx.x, j = threadIdx.x;
    reduce_group_rows(partials, nblocks, GS_SLOTS, grp, red, total);
    if (j < 32) {
        double g;
        const int i = gs_param_of(grp, j, total, g);
        if (i >= 0) out[i] = (float)(g * inv_b);
    }
    if (grp == 9 && j == 0) gs_write_stats(total, inv_b, adv_w, rec_w, out + OFDMGAN_G_NPARAMS);
    __threadfence();                                             // this block's gradients are visible before it takes its ticket
    __syncthreads();
    if (j == 0) ticket = atomicAdd(arrivals, 1u);
    __syncthreads();
    if (ticket != gridDim.x - 1) return;
    if (j == 0) *arrivals = 0u;
    if (world > 1) {
        PeerBlock* mine = peers.p[rank];
        const unsigned int seq = mine->seq + 1u;
        peer_allreduce_block(peers, rank, world, seq, out, OFDMGAN_GEN_OUT);     // traps if a peer never arrives
        if (j == 0) mine->seq = seq;
    }
    const int t = *step_dev + 1;
    const AdamCoef c = adam_coef_dev(lr, b1, b2, eps, t);
    if (j < OFDMGAN_G_NPARAMS) {
        float pi = p[j], mi = m[j], vi = v[j];
        adam_one(pi, mi, vi, __ldcg(out + j), c);
        p[j] = pi; m[j] = mi; v[j] = vi;
    }
    __syncthreads();
    if (j == 0) *step_dev = t;
}

// The same tail on ONE GPU (as k_critic_tail1): every block applies Adam to the parameters of its own group right after its reduction;
// the ticket only decides who advances the step count.  Bit-identical to k_gen_tail.
__global__ void __launch_bounds__(1024) k_gen_tail1(const float* __restrict__ partials, int nblocks, double inv_b, double adv_w, double rec_w,
                                                    float* __restrict__ out, float* __restrict__ p, float* __restrict__ m,
                                                    float* __restrict__ v, double lr, double b1, double b2, double eps,
                                                    int32_t* __restrict__ step_dev, unsigned int* __restrict__ arrivals) {
    __shared__ double red[32 * 32], total[32];
    const int grp = blockIdx.x, j = threadIdx.x;
    const int t = *step_dev + 1;
    reduce_group_rows(partials, nblocks, GS_SLOTS, grp, red, total);
    if (j < 32) {
        double gd;
        const int i = gs_param_of(grp, j, total, gd);
        if (i >= 0) {
            const float g = (float)(gd * inv_b);
            out[i] = g;
            float pi = p[i], mi = m[i], vi = v[i];
            const AdamCoef c = adam_coef_dev(lr, b1, b2, eps, t);
            adam_one(pi, mi, vi, g, c);
            p[i] = pi; m[i] = mi; v[i] = vi;
        }
    }
    if (grp == 9 && j == 0) gs_write_stats(total, inv_b, adv_w, rec_w, out + OFDMGAN_G_NPARAMS);
    __syncthreads();
    if (j == 0 && atomicAdd(arrivals, 1u) == gridDim.x - 1) {   // every block has read the step count: advance it
        *arrivals = 0u;
        *step_dev = t;
    }
}

// ------------------------------------------------------------------------------------------------ generator backward (API)
// what autograd needs for MiniGenerator.forward: dparams = sum_b backward(dy_b) and (optionally) dx
// (four tiles per warp - x | dy | y | scratch - and up to 168 registers: 3 CTAs per SM)
constexpr int GENBWD_PER_SM = 3;
constexpr size_t GENBWD_SMEM = (size_t)4 * OG_THREADS * 8 * sizeof(float4) + (size_t)GSX_NG * OG_THREADS * sizeof(float);
template <bool NEED_DX>
__global__ void __launch_bounds__(OG_THREADS, GENBWD_PER_SM) k_gen_bwd(const float* __restrict__ x, const float* __restrict__ dy,
                                                                        float* __restrict__ dx, float* __restrict__ partials, int64_t B,
                                                                        float slope) {
    extern __shared__ float4 sm[];
    float* sacc = reinterpret_cast<float*>(sm + 4 * OG_THREADS * 8);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float4* t_x = sm + warp * TILE4;
    float4* t_dy = sm + (NWARP + warp) * TILE4;
    float4* t_y = sm + (2 * NWARP + warp) * TILE4;
    float4* t_p = sm + (3 * NWARP + warp) * TILE4;
    SAcc acc{sacc + threadIdx.x, OG_THREADS};
#pragma unroll
    for (int g = 0; g < GSX_NG; ++g) sacc[g * OG_THREADS + threadIdx.x] = 0.f;
    const int64_t ntiles = (B + OG_THREADS - 1) / OG_THREADS;
    for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const int64_t base = t * OG_THREADS + warp * 32;
        if (base >= B) continue;
        __syncwarp();
        tile_fill_f32(x, base, B, t_x, lane);
        tile_fill_f32(dy, base, B, t_dy, lane);            // rows beyond B are zero-filled: they contribute nothing
        __syncwarp();
        float a1[4][8], a2[8][4], sk[4][8];
        uint32_t z3pos;
        gs_fwd<true>(c_g, slope, t_x, t_y, t_p, lane, a1, a2, sk, z3pos);
        gs_bwd<NEED_DX>(c_g, slope, t_x, t_y, t_dy, t_p, lane, a1, a2, sk, z3pos, acc);
        if (NEED_DX) {
            __syncwarp();
            tile_drain_f32(dx, base, B, t_dy, lane);
        }
    }
    __syncthreads();
    float* row = partials + (size_t)blockIdx.x * GS_SLOTS;
    for (int s = threadIdx.x; s < GS_SLOTS; s += OG_THREADS) {
        const int grp = s >> 5, j = s & 31;
        float t = 0.f;
#pragma unroll
        for (int w = 0; w < NWARP; ++w) t += sacc[grp * OG_THREADS + w * 32 + j];
        row[s] = t;
    }
}

// ------------------------------------------------------------------------------------------------ Adam
// torch.optim.Adam._single_tensor_adam, fp32 state, no amsgrad / weight decay (oracle/fp32_models.c oracle_adam).
// Written with explicit round-to-nearest ops so nothing is contracted into an FMA the eager reference does not have.
__global__ void k_adam(float* __restrict__ p, float* __restrict__ m, float* __restrict__ v, const float* __restrict__ g, int n,
                       AdamCoef c, float grad_scale) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    adam_one(p[i], m[i], v[i], __fmul_rn(g[i], grad_scale), c);
}

// step count read from / advanced in device memory: one block, n <= 1024
__global__ void __launch_bounds__(1024) k_adam_ctr(float* __restrict__ p, float* __restrict__ m, float* __restrict__ v,
                                                   const float* __restrict__ g, int n, double lr, double b1, double b2, double eps,
                                                   int32_t* __restrict__ step_dev, float grad_scale) {
    const int t = *step_dev + 1;
    const AdamCoef c = adam_coef_dev(lr, b1, b2, eps, t);
    const int i = threadIdx.x;
    if (i < n) adam_one(p[i], m[i], v[i], __fmul_rn(g[i], grad_scale), c);
    __syncthreads();
    if (i == 0) *step_dev = t;
}

}  // namespace og

using namespace og;

extern "C" {

static int gen_step_impl(const float* clean_dev, const float* noisy_dev, const float* fake_in_dev, const float* dparams521,
                         const float* gparams258, float adv_weight, float rec_weight, float leaky_slope, int64_t B_local, int64_t B_global,
                         float* out_dev, float* fake_out_dev, void* stream) {
    cudaStream_t s = (cudaStream_t)stream;
    if (!dparams521 || !gparams258 || !out_dev || B_local < 0 || B_global < 1 || B_global < B_local) return OFDMGAN_E_ARG;
    if (B_local > 0 && (!clean_dev || !noisy_dev || !aligned16(clean_dev) || !aligned16(noisy_dev))) return OFDMGAN_E_ARG;
    if (fake_out_dev && !aligned16(fake_out_dev)) return OFDMGAN_E_ARG;
    if (fake_in_dev && !aligned16(fake_in_dev)) return OFDMGAN_E_ARG;
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
    const int grid = grid_for(B_local, GENSTEP_THREADS, GENSTEP_PER_SM);
    void* partials = nullptr;
    if ((rc = scratch_for_slot(slot, (size_t)grid * GS_SLOTS * sizeof(float), 7, &partials))) return rc;
    GenStepArgs a{};
    a.clean = clean_dev; a.noisy = noisy_dev; a.fake_in = fake_in_dev; a.fake_out = fake_out_dev;
    a.B = B_local; a.slot = slot; a.slope = leaky_slope; a.adv_w = adv_weight; a.rec_w = rec_weight;
    a.partials = (float*)partials;
    if (fake_in_dev) {
        OG_CHECK(cudaFuncSetAttribute(k_gen_step<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GENSTEP_SMEM));
        k_gen_step<true><<<grid, GENSTEP_THREADS, GENSTEP_SMEM, s>>>(a);
    } else {
        OG_CHECK(cudaFuncSetAttribute(k_gen_step<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GENSTEP_SMEM));
        k_gen_step<false><<<grid, GENSTEP_THREADS, GENSTEP_SMEM, s>>>(a);
    }
    OG_CHECK(cudaGetLastError());
    k_finalize_gen<<<GSX_NG, 1024, 0, s>>>((const float*)partials, grid, 1.0 / (double)B_global, (double)adv_weight, (double)rec_weight,
                                          out_dev, out_dev + OFDMGAN_G_NPARAMS);
    return (int)cudaGetLastError();
}

int ofdmgan_gen_step(const float* clean_dev, const float* noisy_dev, const float* dparams521, const float* gparams258, float adv_weight,
                     float rec_weight, float leaky_slope, int64_t B_local, int64_t B_global, float* out_dev, float* fake_out_dev,
                     void* stream) {
    return gen_step_impl(clean_dev, noisy_dev, nullptr, dparams521, gparams258, adv_weight, rec_weight, leaky_slope, B_local, B_global,
                         out_dev, fake_out_dev, stream);
}

int ofdmgan_gen_step_fake(const float* clean_dev, const float* noisy_dev, const float* fake_dev, const float* dparams521,
                          const float* gparams258, float adv_weight, float rec_weight, float leaky_slope, int64_t B_local,
                          int64_t B_global, float* out_dev, void* stream) {
    if (B_local > 0 && !fake_dev) return OFDMGAN_E_ARG;
    return gen_step_impl(clean_dev, noisy_dev, fake_dev, dparams521, gparams258, adv_weight, rec_weight, leaky_slope, B_local, B_global,
                         out_dev, nullptr, stream);
}

int ofdmgan_gen_train_ctr(const float* clean_dev, const float* noisy_dev, const float* fake_dev, int32_t* step_dev, float* gparams258_dev,
                          float* m_dev, float* v_dev, double lr, double beta1, double beta2, double eps, const float* dparams521_dev,
                          float adv_weight, float rec_weight, float leaky_slope, int64_t B, int64_t B_global, float* out_dev,
                          int d_image_staged, int g_image_staged, ofdmgan_comm* comm, void* stream) {
    cudaStream_t s = (cudaStream_t)stream;
    if (!clean_dev || !noisy_dev || !fake_dev || !step_dev || !gparams258_dev || !m_dev || !v_dev || !dparams521_dev || !out_dev || B < 1 ||
        B_global < B)
        return OFDMGAN_E_ARG;
    if (!aligned16(clean_dev) || !aligned16(noisy_dev) || !aligned16(fake_dev)) return OFDMGAN_E_ARG;
    PeerPtrs peers{};
    int rank = 0, world = 1;
    if (comm && !comm_view(comm, &peers, &rank, &world)) return OFDMGAN_E_ARG;
    int rc;
    CallGuard guard(s);
    if ((rc = guard.rc)) return rc;
    const int slot = 0;
    // *_staged: the staging buffer of this device already holds the image of these parameters (D: the tail of
    // ofdmgan_critic_train_ctr left it; G: ofdmgan_gen_fwd_f32 with the same device parameters built it) - copy it, skip the rebuild
    if ((rc = d_image_staged ? commit_d_image(slot, s) : upload_d(dparams521_dev, slot, s))) return rc;
    if ((rc = g_image_staged ? commit_g_image(slot, s) : upload_g(gparams258_dev, slot, s))) return rc;
    const int grid = grid_for(B, GENSTEP_THREADS, GENSTEP_PER_SM);
    void *partials = nullptr, *arrivals = nullptr;
    if ((rc = scratch_for_slot(slot, (size_t)grid * GS_SLOTS * sizeof(float), 7, &partials))) return rc;
    if ((rc = scratch_for_slot(slot, 256, 9, &arrivals))) return rc;
    static bool zeroed[64] = {false};
    int dev = 0;
    OG_CHECK(cudaGetDevice(&dev));
    if (!zeroed[dev]) {                                          // the tail leaves the counter at zero; only the very first use needs this
        OG_CHECK(cudaMemsetAsync(arrivals, 0, 256, s));
        zeroed[dev] = true;
    }
    GenStepArgs a{};
    a.clean = clean_dev; a.noisy = noisy_dev; a.fake_in = fake_dev; a.fake_out = nullptr;
    a.B = B; a.slot = slot; a.slope = leaky_slope; a.adv_w = adv_weight; a.rec_w = rec_weight;
    a.partials = (float*)partials;
    OG_CHECK(cudaFuncSetAttribute(k_gen_step<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GENSTEP_SMEM));
    k_gen_step<true><<<grid, GENSTEP_THREADS, GENSTEP_SMEM, s>>>(a);
    OG_CHECK(cudaGetLastError());
    if (world == 1)
        k_gen_tail1<<<GSX_NG, 1024, 0, s>>>((const float*)partials, grid, 1.0 / (double)B_global, (double)adv_weight, (double)rec_weight,
                                            out_dev, gparams258_dev, m_dev, v_dev, lr, beta1, beta2, eps, step_dev, (unsigned int*)arrivals);
    else
        k_gen_tail<<<GSX_NG, 1024, 0, s>>>((const float*)partials, grid, 1.0 / (double)B_global, (double)adv_weight, (double)rec_weight,
                                           out_dev, gparams258_dev, m_dev, v_dev, lr, beta1, beta2, eps, step_dev, (unsigned int*)arrivals,
                                           peers, rank, world);
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
    const int grid = grid_for(B, OG_THREADS, GENBWD_PER_SM);
    void* partials = nullptr;
    if ((rc = scratch_for_slot(slot, (size_t)grid * GS_SLOTS * sizeof(float), 7, &partials))) return rc;
    if (dx_dev) {
        OG_CHECK(cudaFuncSetAttribute(k_gen_bwd<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GENBWD_SMEM));
        k_gen_bwd<true><<<grid, OG_THREADS, GENBWD_SMEM, s>>>(x_dev, dy_dev, dx_dev, (float*)partials, B, leaky_slope);
    } else {
        OG_CHECK(cudaFuncSetAttribute(k_gen_bwd<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GENBWD_SMEM));
        k_gen_bwd<false><<<grid, OG_THREADS, GENBWD_SMEM, s>>>(x_dev, dy_dev, nullptr, (float*)partials, B, leaky_slope);
    }
    OG_CHECK(cudaGetLastError());
    k_finalize_gen<<<GSX_NG, 1024, 0, s>>>((const float*)partials, grid, 1.0, 0.0, 0.0, dparams258_dev, nullptr);
    return (int)cudaGetLastError();
}

int ofdmgan_adam(float* p_dev, float* m_dev, float* v_dev, const float* g_dev, int n, double lr, double beta1, double beta2,
                 double eps, int step, float grad_scale, void* stream) {
    if (!p_dev || !m_dev || !v_dev || !g_dev || n < 0 || step < 1) return OFDMGAN_E_ARG;
    if (n == 0) return 0;
    k_adam<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(p_dev, m_dev, v_dev, g_dev, n, adam_coef(lr, beta1, beta2, eps, step),
                                                               grad_scale);
    return (int)cudaGetLastError();
}

int ofdmgan_adam_ctr(float* p_dev, float* m_dev, float* v_dev, const float* g_dev, int n, double lr, double beta1, double beta2, double eps,
                     int32_t* step_dev, float grad_scale, void* stream) {
    if (!p_dev || !m_dev || !v_dev || !g_dev || !step_dev || n < 0 || n > 1024) return OFDMGAN_E_ARG;
    if (n == 0) return 0;
    k_adam_ctr<<<1, 1024, 0, (cudaStream_t)stream>>>(p_dev, m_dev, v_dev, g_dev, n, lr, beta1, beta2, eps, step_dev, grad_scale);
    return (int)cudaGetLastError();
}

}  // extern "C"
