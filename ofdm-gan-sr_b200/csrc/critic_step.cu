// libofdmgan kernel (4), the hot part: one critic iteration of CWGAN-GP (train.py:228-250) as one launch -
// -mean D(real) + mean D(fake) + gp_weight * penalty, its gradient w.r.t. the 521 critic parameters (closed-form double
// backward for the penalty) and the five logged statistics.   ofdmgan_critic_step / ofdmgan_gradient_penalty
#include "train_common.cuh"
#include "critic_stream.cuh"
#include "peer_comm.cuh"

namespace og {

struct CriticArgs {
    const float* real;        // clean
    const float* cond;        // noisy
    const float* fake;
    const float* alpha;       // nullable -> Philox(seed, sample, alpha_iter, purpose 1)
    PhiloxKeys keys;
    uint64_t sample0;
    uint32_t alpha_iter;
    const int32_t* alpha_iter_dev;   // nullable: device-resident counter read instead of alpha_iter (graph-replayable steps)
    int64_t B;
    int slot;
    float slope;
    float gp_scale;           // weight of the penalty term relative to the score terms
    int want_grads;
    float* partials;          // [grid][DS_SLOTS]
    float* norms;             // nullable [B]: ||grad|| per sample (test hook of ofdmgan_gradient_penalty)
};

// ------------------------------------------------------------------------------------------------ critic step, v2
// Work item = (term, 128-sample tile) with term in {penalty, -D(real), +D(fake)}: three times the parallelism of one
// thread doing all three terms of its sample, which matters at 65,536 samples per GPU (443 samples per SM).  The
// heaviest term (penalty, 1.15x a score term by instruction count) is scheduled first.  <= 128 registers: 4 CTAs (16 warps) per SM, see critic_stream.cuh.
#ifndef OG_CRITIC_PER_SM
#define OG_CRITIC_PER_SM 4
#endif
constexpr int CRITIC_PER_SM = OG_CRITIC_PER_SM;

template <bool SCORE>
__global__ void __launch_bounds__(OG_THREADS, CRITIC_PER_SM) k_critic2(const __grid_constant__ CriticArgs a) {
    __shared__ float4 sm[2 * OG_THREADS * 8];
    __shared__ float sacc[CS_NG * OG_THREADS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float4* t_a = sm + warp * TILE4;                           // candidate (real | fake | x_hat)
    float4* t_c = sm + (NWARP + warp) * TILE4;                 // condition
    const float* W = c_d;
    SAcc acc{sacc + threadIdx.x, OG_THREADS};
#pragma unroll
    for (int g = 0; g < CS_NG; ++g) sacc[g * OG_THREADS + threadIdx.x] = 0.f;     // thread-private entries: no barrier needed
    float s_real = 0.f, s_fake = 0.f, s_gp = 0.f;
    // Static round-robin of tile items over the CTAs (the summation order is fixed).  With 4 CTAs per SM placed at stride 148 (traced
    // with %smid) every scheduler of an SM holds one warp of each CTA, so 1 536 items over 592 CTAs are 10 or 11 warp items per
    // scheduler - as even as 32-sample items get.  Handing items out per warp, spreading the last round's items evenly over the CTAs,
    // or steering the (15 % dearer) penalty items to the SMs with 10 items all measured slower (profiles/r2_notes.md).
    const int64_t ntiles = (a.B + OG_THREADS - 1) / OG_THREADS;
    const int64_t ntasks = ntiles * (SCORE ? 3 : 1);
    for (int64_t task = blockIdx.x; task < ntasks; task += gridDim.x) {
        const int kind = (int)(task / ntiles);                 // 0 penalty, 1 real, 2 fake (CTA-uniform)
        const int64_t base = (task - kind * ntiles) * OG_THREADS + warp * 32;
        if (base >= a.B) continue;
        const int64_t b = base + lane;
        const bool live = b < a.B;
        __syncwarp();
        tile_fill_f32(a.cond, base, a.B, t_c, lane);
        if (kind == 0) {
            float alpha = 0.f;
            if (live) {
                if (a.alpha) {
                    alpha = a.alpha[b];
                } else {
                    const uint64_t smp = a.sample0 + (uint64_t)b;
                    uint32_t x[4];
                    philox4x32_10(a.keys, (uint32_t)smp, (uint32_t)(smp >> 32), a.alpha_iter_dev ? (uint32_t)*a.alpha_iter_dev : a.alpha_iter, 1u, x);
                    alpha = u_half(x[0]);
                }
            }
            // x_hat = alpha * real + (1 - alpha) * fake, two products then a sum (models/discriminator.py:211), formed
            // while the tile is filled: chunk idx belongs to frame idx >> 3 of this warp
            const float4* pr = reinterpret_cast<const float4*>(a.real) + base * 8;
            const float4* pf = reinterpret_cast<const float4*>(a.fake) + base * 8;
            const int64_t limit = (a.B - base) * 8;
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                const int idx = r * 32 + lane, f = idx >> 3, c = idx & 7;
                const float al = __shfl_sync(0xffffffffu, alpha, f), om = 1.0f - al;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (idx < limit) {
                    const float4 x = __ldg(pr + idx), y = __ldg(pf + idx);
                    v.x = __fadd_rn(__fmul_rn(al, x.x), __fmul_rn(om, y.x));
                    v.y = __fadd_rn(__fmul_rn(al, x.y), __fmul_rn(om, y.y));
                    v.z = __fadd_rn(__fmul_rn(al, x.z), __fmul_rn(om, y.z));
                    v.w = __fadd_rn(__fmul_rn(al, x.w), __fmul_rn(om, y.w));
                }
                t_a[f * 8 + (c ^ (f & 7))] = v;
            }
            __syncwarp();
            float n;
            const float pen = cs_gp_pass(W, a.slope, live ? a.gp_scale : 0.0f, t_a, t_c, acc, lane, n);
            if (live) {
                s_gp += pen;
                if (a.norms) a.norms[b] = n;
            }
        } else {
            tile_fill_f32(kind == 1 ? a.real : a.fake, base, a.B, t_a, lane);
            __syncwarp();
            const float g = live ? (kind == 1 ? -1.0f : 1.0f) : 0.0f;
            const float sc = cs_score_pass<false>(W, a.slope, g, t_a, t_c, acc, lane);
            if (live) {
                if (kind == 1) s_real += sc; else s_fake += sc;
            }
        }
    }
    {   // statistics ride in the spare slots of group 1
        const float r = warp_sum(s_real), f = warp_sum(s_fake), g = warp_sum(s_gp);
        if (lane == CS_SREAL - 32) acc.add(1, r);
        if (lane == CS_SFAKE - 32) acc.add(1, f);
        if (lane == CS_SGP - 32) acc.add(1, g);
    }
    __syncthreads();
    float* row = a.partials + (size_t)blockIdx.x * CS_SLOTS;
    for (int s = threadIdx.x; s < CS_SLOTS; s += OG_THREADS) {
        const int grp = s >> 5, j = s & 31;
        float t = 0.f;
#pragma unroll
        for (int w = 0; w < NWARP; ++w) t += sacc[grp * OG_THREADS + w * 32 + j];
        row[s] = t;
    }
}

// One block per accumulator group: fixed-order sum over the CTA rows, slot -> parameter order, 1/B scaling; the block of
// group 1 also writes the loss statistics.   grads (nullable): 521 floats.  stats (nullable): stat_mode 0 -> the 5 scalars
// of train.py:255-261 (+2 pad), stat_mode 1 -> 1 float = mean penalty.
__global__ void __launch_bounds__(1024) k_finalize_critic2(const float* __restrict__ partials, int nblocks, double inv_b,
                                                           double gp_weight, float* __restrict__ grads, float* __restrict__ stats,
                                                           int stat_mode) {
    __shared__ double red[32 * 32], total[32];
    const int grp = blockIdx.x;
    reduce_group_rows(partials, nblocks, CS_SLOTS, grp, red, total);
    if (threadIdx.x < 32) {
        const int i = cs_param_of(grp, threadIdx.x);
        if (grads && i >= 0) grads[i] = (float)(total[threadIdx.x] * inv_b);
    }
    if (grp == 1 && stats && threadIdx.x == 0) {
        const double dr = total[CS_SREAL - 32] * inv_b, df = total[CS_SFAKE - 32] * inv_b, gp = total[CS_SGP - 32] * inv_b;
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

// Single-GPU tail of a critic iteration: the fixed-order reduction of k_finalize_critic2, and - in the block that finishes last -
// Adam on the 521 parameters (step count in device memory, as k_adam_ctr) and the refresh of the weight-image staging buffer.
// One launch instead of finalize + Adam + prep_d_image.
__global__ void __launch_bounds__(1024) k_critic_tail(const float* __restrict__ partials, int nblocks, double inv_b, double gp_weight,
                                                      float* __restrict__ out, float* __restrict__ p, float* __restrict__ m,
                                                      float* __restrict__ v, double lr, double b1, double b2, double eps,
                                                      int32_t* __restrict__ step_dev, unsigned int* __restrict__ arrivals,
                                                      float* __restrict__ image, PeerPtrs peers, int rank, int world) {
    __shared__ double red[32 * 32], total[32];
    __shared__ unsigned int ticket;
    __shared__ float pnew[OFDMGAN_D_NPARAMS];
    const int grp = blockIdx.x;
    reduce_group_rows(partials, nblocks, CS_SLOTS, grp, red, total);
    if (threadIdx.x < 32) {
        const int i = cs_param_of(grp, threadIdx.x);
        if (i >= 0) out[i] = (float)(total[threadIdx.x] * inv_b);
    }
    if (grp == 1 && threadIdx.x == 0) {
        float* stats = out + OFDMGAN_D_NPARAMS;
        const double dr = total[CS_SREAL - 32] * inv_b, df = total[CS_SFAKE - 32] * inv_b, gp = total[CS_SGP - 32] * inv_b;
        stats[0] = (float)(df - dr + gp_weight * gp);
        stats[1] = (float)(dr - df);
        stats[2] = (float)gp;
        stats[3] = (float)dr;
        stats[4] = (float)df;
        stats[5] = 0.f;
        stats[6] = 0.f;
    }
    __threadfence();                                             // this block's gradients are visible before it takes its ticket
    __syncthreads();
    if (threadIdx.x == 0) ticket = atomicAdd(arrivals, 1u);
    __syncthreads();
    if (ticket != gridDim.x - 1) return;
    // last block: every group's gradients are in `out`
    if (threadIdx.x == 0) *arrivals = 0u;
    if (world > 1) {                                             // data parallel: sum the 528 floats over the ranks through peer memory
        PeerBlock* mine = peers.p[rank];
        const unsigned int seq = mine->seq + 1u;
        peer_allreduce_block(peers, rank, world, seq, out, OFDMGAN_CRITIC_OUT);   // traps if a peer never arrives
        if (threadIdx.x == 0) mine->seq = seq;
    }
    const int t = *step_dev + 1;
    const AdamCoef c = adam_coef_dev(lr, b1, b2, eps, t);
    const int i = threadIdx.x;
    if (i < OFDMGAN_D_NPARAMS) {
        float pi = p[i], mi = m[i], vi = v[i];
        adam_one(pi, mi, vi, __ldcg(out + i), c);               // (__fmul_rn(g, 1) of k_adam_ctr is the identity)
        p[i] = pi; m[i] = mi; v[i] = vi;
        pnew[i] = pi;
    }
    __syncthreads();
    if (i == 0) *step_dev = t;
    for (int j = i; j < OG_D_IMG; j += blockDim.x) image[j] = d_img_entry(pnew, j);
}

// The same tail on ONE GPU (no exchange): a parameter's gradient is complete inside the block of its accumulator group, so every block
// applies Adam to its own <= 28 parameters and writes their entries of the weight image right after its reduction - no hand-over to a
// last block.  The ticket only decides who advances the step count (after every block has read it).  Bit-identical to k_critic_tail.
__global__ void __launch_bounds__(1024) k_critic_tail1(const float* __restrict__ partials, int nblocks, double inv_b, double gp_weight,
                                                       float* __restrict__ out, float* __restrict__ p, float* __restrict__ m,
                                                       float* __restrict__ v, double lr, double b1, double b2, double eps,
                                                       int32_t* __restrict__ step_dev, unsigned int* __restrict__ arrivals,
                                                       float* __restrict__ image) {
    __shared__ double red[32 * 32], total[32];
    const int grp = blockIdx.x, j = threadIdx.x;
    const int t = *step_dev + 1;
    int i = -1;
    float pi = 0.f, mi = 0.f, vi = 0.f;
    if (j < 32) {
        i = cs_param_of(grp, j);
        if (i >= 0) { pi = p[i]; mi = m[i]; vi = v[i]; }
    }
    reduce_group_rows(partials, nblocks, CS_SLOTS, grp, red, total);
    if (i >= 0) {
        const float g = (float)(total[j] * inv_b);
        out[i] = g;
        const AdamCoef c = adam_coef_dev(lr, b1, b2, eps, t);
        adam_one(pi, mi, vi, g, c);
        p[i] = pi; m[i] = mi; v[i] = vi;
        d_img_scatter(image, i, pi);
    }
    if (grp == 1 && j == 0) {
        float* stats = out + OFDMGAN_D_NPARAMS;
        const double dr = total[CS_SREAL - 32] * inv_b, df = total[CS_SFAKE - 32] * inv_b, gp = total[CS_SGP - 32] * inv_b;
        stats[0] = (float)(df - dr + gp_weight * gp);
        stats[1] = (float)(dr - df);
        stats[2] = (float)gp;
        stats[3] = (float)dr;
        stats[4] = (float)df;
        stats[5] = 0.f;
        stats[6] = 0.f;
    }
    __syncthreads();
    if (j == 0 && atomicAdd(arrivals, 1u) == gridDim.x - 1) {   // every block has read the step count: advance it
        *arrivals = 0u;
        *step_dev = t;
    }
}

}  // namespace og

using namespace og;

extern "C" {

static int launch_critic(bool score, const float* real, const float* fake, const float* cond, const float* alpha, uint64_t seed,
                         uint64_t sample0, uint32_t alpha_iter, const float* dparams521, float gp_scale, float slope, int64_t B,
                         bool want_grads, float* norms, cudaStream_t s, int* slot_out, int* grid_out, void** partials_out,
                         const int32_t* alpha_iter_dev = nullptr) {
    int slot, rc;
    CallGuard guard(s);
    if ((rc = guard.rc)) return rc;
    slot = 0;
    if ((rc = upload_d(dparams521, slot, s))) return rc;
    const int grid = grid_for(B * (score ? 3 : 1), OG_THREADS, CRITIC_PER_SM);
    void* partials = nullptr;
    if ((rc = scratch_for_slot(slot, (size_t)grid * CS_SLOTS * sizeof(float), 6, &partials))) return rc;
    CriticArgs a{};
    a.real = real; a.cond = cond; a.fake = fake; a.alpha = alpha;
    a.keys = philox_keys(seed);
    a.sample0 = sample0; a.alpha_iter = alpha_iter; a.alpha_iter_dev = alpha_iter_dev;
    a.B = B; a.slot = slot; a.slope = slope; a.gp_scale = gp_scale;
    a.want_grads = want_grads ? 1 : 0;
    a.partials = (float*)partials;
    a.norms = norms;
    if (score) k_critic2<true><<<grid, OG_THREADS, 0, s>>>(a);
    else k_critic2<false><<<grid, OG_THREADS, 0, s>>>(a);
    OG_CHECK(cudaGetLastError());
    *slot_out = slot; *grid_out = grid; *partials_out = partials;
    return 0;
}

int ofdmgan_gradient_penalty(const float* real_dev, const float* fake_dev, const float* cond_dev, const float* alpha_dev,
                             uint64_t seed, uint64_t sample0, uint32_t alpha_iter, const float* dparams521, float* gp_dev,
                             float* dparams521_dev, int64_t B, float leaky_slope, void* stream) {
    if (!real_dev || !fake_dev || !cond_dev || !dparams521 || !gp_dev || B < 1) return OFDMGAN_E_ARG;
    if (!aligned16(real_dev) || !aligned16(fake_dev) || !aligned16(cond_dev)) return OFDMGAN_E_ARG;
    cudaStream_t s = (cudaStream_t)stream;
    int slot, grid, rc;
    void* partials;
    if ((rc = launch_critic(false, real_dev, fake_dev, cond_dev, alpha_dev, seed, sample0, alpha_iter, dparams521, 1.0f, leaky_slope, B,
                            dparams521_dev != nullptr, nullptr, s, &slot, &grid, &partials))) return rc;
    k_finalize_critic2<<<CS_NG, 1024, 0, s>>>((const float*)partials, grid, 1.0 / (double)B, 1.0, dparams521_dev, gp_dev, 1);
    return (int)cudaGetLastError();
}

static int critic_step_impl(const float* clean_dev, const float* noisy_dev, const float* fake_dev, const float* alpha_dev, uint64_t seed,
                            uint64_t sample0, uint32_t alpha_iter, const int32_t* alpha_iter_dev, const float* dparams521,
                            float gp_weight, float leaky_slope, int64_t B_local, int64_t B_global, float* out_dev, void* stream) {
    cudaStream_t s = (cudaStream_t)stream;
    if (!dparams521 || !out_dev || B_local < 0 || B_global < 1 || B_global < B_local) return OFDMGAN_E_ARG;
    if (B_local > 0 && (!clean_dev || !noisy_dev || !fake_dev || !aligned16(clean_dev) || !aligned16(noisy_dev) || !aligned16(fake_dev)))
        return OFDMGAN_E_ARG;
    if (B_local == 0) {
        OG_CHECK(cudaMemsetAsync(out_dev, 0, OFDMGAN_CRITIC_OUT * sizeof(float), s));
        return 0;
    }
    int slot, grid, rc;
    void* partials;
    if ((rc = launch_critic(true, clean_dev, fake_dev, noisy_dev, alpha_dev, seed, sample0, alpha_iter, dparams521, gp_weight,
                            leaky_slope, B_local, true, nullptr, s, &slot, &grid, &partials, alpha_iter_dev))) return rc;
    k_finalize_critic2<<<CS_NG, 1024, 0, s>>>((const float*)partials, grid, 1.0 / (double)B_global, (double)gp_weight, out_dev,
                                              out_dev + OFDMGAN_D_NPARAMS, 0);
    return (int)cudaGetLastError();
}

int ofdmgan_critic_step(const float* clean_dev, const float* noisy_dev, const float* fake_dev, const float* alpha_dev, uint64_t seed,
                        uint64_t sample0, uint32_t alpha_iter, const float* dparams521, float gp_weight, float leaky_slope,
                        int64_t B_local, int64_t B_global, float* out_dev, void* stream) {
    return critic_step_impl(clean_dev, noisy_dev, fake_dev, alpha_dev, seed, sample0, alpha_iter, nullptr, dparams521, gp_weight,
                            leaky_slope, B_local, B_global, out_dev, stream);
}

int ofdmgan_critic_step_ctr(const float* clean_dev, const float* noisy_dev, const float* fake_dev, uint64_t seed, uint64_t sample0,
                            const int32_t* alpha_iter_dev, const float* dparams521, float gp_weight, float leaky_slope,
                            int64_t B_local, int64_t B_global, float* out_dev, void* stream) {
    if (!alpha_iter_dev) return OFDMGAN_E_ARG;
    return critic_step_impl(clean_dev, noisy_dev, fake_dev, nullptr, seed, sample0, 0u, alpha_iter_dev, dparams521, gp_weight,
                            leaky_slope, B_local, B_global, out_dev, stream);
}

int ofdmgan_critic_train_ctr(const float* clean_dev, const float* noisy_dev, const float* fake_dev, uint64_t seed, uint64_t sample0,
                             int32_t* step_dev, float* dparams521_dev, float* m_dev, float* v_dev, double lr, double beta1, double beta2,
                             double eps, float gp_weight, float leaky_slope, int64_t B, int64_t B_global, float* out_dev,
                             int image_is_current, ofdmgan_comm* comm, void* stream) {
    cudaStream_t s = (cudaStream_t)stream;
    if (!clean_dev || !noisy_dev || !fake_dev || !step_dev || !dparams521_dev || !m_dev || !v_dev || !out_dev || B < 1 || B_global < B)
        return OFDMGAN_E_ARG;
    PeerPtrs peers{};
    int rank = 0, world = 1;
    if (comm && !comm_view(comm, &peers, &rank, &world)) return OFDMGAN_E_ARG;
    if (!aligned16(clean_dev) || !aligned16(noisy_dev) || !aligned16(fake_dev)) return OFDMGAN_E_ARG;
    int rc;
    CallGuard guard(s);
    if ((rc = guard.rc)) return rc;
    const int slot = 0;
    // image_is_current: the previous call on this stream was this function on the same parameters, so the constant bank already
    // holds the image of dparams521_dev (the tail below refreshed and committed it)
    if (!image_is_current && (rc = upload_d(dparams521_dev, slot, s))) return rc;
    const int grid = grid_for(B * 3, OG_THREADS, CRITIC_PER_SM);
    void *partials = nullptr, *arrivals = nullptr;
    if ((rc = scratch_for_slot(slot, (size_t)grid * CS_SLOTS * sizeof(float), 6, &partials))) return rc;
    if ((rc = scratch_for_slot(slot, 256, 8, &arrivals))) return rc;
    static bool zeroed[64] = {false};
    int dev = 0;
    OG_CHECK(cudaGetDevice(&dev));
    if (!zeroed[dev]) {                                          // the tail leaves the counter at zero; only the very first use needs this
        OG_CHECK(cudaMemsetAsync(arrivals, 0, 256, s));
        zeroed[dev] = true;
    }
    float* image = nullptr;
    if ((rc = d_image_staging(slot, &image))) return rc;
    CriticArgs a{};
    a.real = clean_dev; a.cond = noisy_dev; a.fake = fake_dev; a.alpha = nullptr;
    a.keys = philox_keys(seed);
    a.sample0 = sample0; a.alpha_iter = 0u; a.alpha_iter_dev = step_dev;
    a.B = B; a.slot = slot; a.slope = leaky_slope; a.gp_scale = gp_weight;
    a.want_grads = 1;
    a.partials = (float*)partials;
    a.norms = nullptr;
    k_critic2<true><<<grid, OG_THREADS, 0, s>>>(a);
    OG_CHECK(cudaGetLastError());
    if (world == 1)
        k_critic_tail1<<<CS_NG, 1024, 0, s>>>((const float*)partials, grid, 1.0 / (double)B_global, (double)gp_weight, out_dev,
                                              dparams521_dev, m_dev, v_dev, lr, beta1, beta2, eps, step_dev, (unsigned int*)arrivals, image);
    else
        k_critic_tail<<<CS_NG, 1024, 0, s>>>((const float*)partials, grid, 1.0 / (double)B_global, (double)gp_weight, out_dev,
                                             dparams521_dev, m_dev, v_dev, lr, beta1, beta2, eps, step_dev, (unsigned int*)arrivals, image,
                                             peers, rank, world);
    OG_CHECK(cudaGetLastError());
    return commit_d_image(slot, s);
}

}  // extern "C"
