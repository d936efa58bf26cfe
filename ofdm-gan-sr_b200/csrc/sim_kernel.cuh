// Kernel (1) and the fused headline path (1)+(2|3)+metrics: one OFDM frame per thread, everything in registers.
//   Philox draws -> symbols -> IFFT -> [CP] -> Rapp PA -> IQ imbalance -> phase noise -> AWGN -> normalise
//   [-> generator (fp32 | Q spec | Q rtl_literal) -> per-SNR-bin MSE / EVM / bit-error accumulation]
// Included by sim_gauss.cu and sim_qpsk.cu (each translation unit owns its constant-memory weight images); the
// instantiation list is what differs.  Template parameters: SRC = symbol source layout (chan_device.cuh),
// GEN = -1 (simulate only) or OFDMGAN_GEN_*.
//
// Work distribution: a persistent grid of (SM count x resident CTAs) blocks, each taking one CONTIGUOUS chunk of
// 128-frame tiles, so the SNR bin of a thread's successive frames changes rarely and the per-thread running sums are
// folded into the CTA's shared double table only on a bin change or every FLUSH_EVERY frames.
#pragma once
#include "chan_device.cuh"
#include "gen_device.cuh"
#include "io_tile.cuh"
#include "equalizer_device.cuh"
#include "sim_metrics.cuh"

// Launch shape.  16 warps per SM at <= 128 registers (measured best of 12/16/20/24 warps on B200, profiles/r1_notes.md)
// as ONE 512-thread CTA per SM with a few CTA barriers per tile: the loop body is ~70 KB of straight-line code, far
// beyond the instruction caches, so warps that drift apart each stream their own copy of it (stall_no_inst was 33 % of
// all stall samples with 4 independent 128-thread CTAs).  Barriers keep the 4 warps of every scheduler on the same
// cache lines.
#ifndef OG_SIM_THREADS
#define OG_SIM_THREADS 512
#endif
#ifndef OG_SIM_BARRIERS
#define OG_SIM_BARRIERS 1
#endif

namespace og {

constexpr int ST = OG_SIM_THREADS;                                   // frames per CTA tile
constexpr int SIM_PER_SM = ST >= 512 ? 1 : 512 / ST;                 // resident CTAs per SM
constexpr size_t SIM_SMEM = (size_t)ST * 8 * sizeof(float4) + OFDMGAN_MAX_SNR_BINS * NM * NC * sizeof(double);

__device__ __forceinline__ void cta_lockstep(int level) {
    if (OG_SIM_BARRIERS >= level) __syncthreads();
}

// EQ: also fill the ZF / MMSE rows (genie-aided equalisers, equalizer_device.cuh)
// LATE: Saleh PA / DC offset / CFO stages and the fading channels compiled in (chan_device.cuh late_stages)
template <int SRC, int GEN, bool EQ, bool LATE>
__global__ void __launch_bounds__(ST, SIM_PER_SM) k_sim(const __grid_constant__ SimArgs a) {
    constexpr bool BITS = SRC != SRC_GAUSS;
    constexpr int NMETH = EQ ? 4 : 2;
    extern __shared__ float4 sm[];                                   // [ST*8] warp tiles, then the CTA's metric table
    double* table = reinterpret_cast<double*>(sm + ST * 8);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float4* wsm = sm + warp * 32 * 8;
    const bool want_metrics = GEN >= 0 && a.partials != nullptr;
    if (want_metrics) {
        for (int i = threadIdx.x; i < a.n_snr * NM * NC; i += blockDim.x) table[i] = 0.0;
        __syncthreads();
    }
    Acc<NMETH> acc;
    acc_reset(acc, -1);

    const int64_t ntiles = (a.B + ST - 1) / ST;
    const int64_t per_cta = (ntiles + gridDim.x - 1) / gridDim.x;
    const int64_t t0 = (int64_t)blockIdx.x * per_cta;
    const int64_t t1 = t0 + per_cta < ntiles ? t0 + per_cta : ntiles;

    // SNR grid position of this thread's first frame; afterwards advanced by 128 frames per tile without divisions
    const bool grid_mode = a.cfg.snr_mode == OFDMGAN_SNR_GRID;
    const uint64_t fps = grid_mode ? (uint64_t)a.cfg.frames_per_snr : 1;
    const bool incremental = grid_mode && fps >= (uint64_t)ST;
    int bin = 0;
    uint64_t in_bin = 0;
    if (grid_mode && t0 < t1) {
        const uint64_t f0 = a.frame0 + (uint64_t)(t0 * ST + threadIdx.x);
        const uint64_t q = f0 / fps;
        in_bin = f0 - q * fps;
        bin = (int)(q % (uint64_t)a.cfg.n_snr);
    }

    for (int64_t t = t0; t < t1; ++t) {
        const int64_t wbase = t * ST + warp * 32;
        const int64_t b = wbase + lane;
        const bool live = b < a.B;
        const int64_t bb = live ? b : a.B - 1;                   // dead lanes recompute the last frame, results dropped
        const uint64_t frame = a.frame0 + (uint64_t)bb;
        int fbin = bin;
        if (grid_mode && !incremental) fbin = snr_bin_of(a.cfg, frame);
        else if (grid_mode && !live) fbin = acc.bin >= 0 ? acc.bin : bin;   // dead lanes must not force a flush
        if (incremental) {                                         // advance to the next tile's frame
            in_bin += ST;
            if (in_bin >= fps) { in_bin -= fps; bin = bin + 1 == a.cfg.n_snr ? 0 : bin + 1; }
        }
        // (a warp wholly beyond the batch keeps going with live == false everywhere: the CTA barriers below need it)

        // block 12: {snr uniform, payload bits}
        uint32_t bits = 0;
        float snr_db;
        {
            uint32_t x12[4] = {0u, 0u, 0u, 0u};
            const bool no_awgn = a.cfg.snr_mode == OFDMGAN_SNR_NONE;
            const bool need12 = (BITS && !a.bits && !a.tx) || (!grid_mode && !no_awgn && !a.snr_db);
            if (need12) philox4x32_10(a.keys, (uint32_t)frame, (uint32_t)(frame >> 32), 12u, 0u, x12);
            if (BITS) bits = a.bits ? a.bits[bb] : x12[1];
            if (grid_mode) snr_db = fmaf(a.cfg.snr_step, (float)(live || !incremental ? fbin : snr_bin_of(a.cfg, frame)), a.cfg.snr_lo);
            else if (no_awgn) snr_db = __int_as_float(0x7f800000);     // +inf: reported as "no noise"
            else snr_db = a.snr_db ? a.snr_db[bb] : fmaf(a.cfg.snr_hi - a.cfg.snr_lo, u_half(x12[0]), a.cfg.snr_lo);
        }
        float cr[16], ci[16], nr[16], ni[16];
        tx_frame<SRC>(a, bb, frame, bits, cr, ci);
        cta_lockstep(2);
        impair_channel<LATE>(a, bb, frame, snr_db, cr, ci, nr, ni);
        normalise(a.cfg.normalize, cr, ci, nr, ni);
        cta_lockstep(1);

        if (a.clean) {
            float f[2][16];
#pragma unroll
            for (int i = 0; i < 16; ++i) { f[0][i] = cr[i]; f[1][i] = ci[i]; }
            tile_store_f32(a.clean, wbase, a.B, wsm, lane, f);
        }
        if (a.noisy) {
            float f[2][16];
#pragma unroll
            for (int i = 0; i < 16; ++i) { f[0][i] = nr[i]; f[1][i] = ni[i]; }
            tile_store_f32(a.noisy, wbase, a.B, wsm, lane, f);
        }
        if (a.snr_out && live) a.snr_out[b] = snr_db;
        if (GEN < 0) continue;                                     // (compile-time: the simulate-only kernel ends its tile here)
        float inv_energy = 0.f;

        if (want_metrics) {
            if (__any_sync(0xffffffffu, fbin != acc.bin || acc.count >= FLUSH_EVERY)) {
                acc_flush<BITS, NMETH>(acc, table, lane);
                acc.bin = fbin;
            }
            inv_energy = fast_rcp(frame_energy(cr, ci));
            if (live) {                                            // NoEQ first: the received frame dies into G's input
                float mse, evm, ratio;
                int errs = 0;
                frame_err(nr, ni, cr, ci, inv_energy, mse, evm, ratio);
                if (BITS) acc.nbits += (float)qpsk_errors<SRC>(a.cfg, nr, ni, bits, errs);
                acc_add<BITS, NMETH>(acc, OFDMGAN_METHOD_NOEQ, mse, evm, ratio, errs);
                acc.count++;
                if (EQ) {                                          // genie-aided ZF and MMSE, sample by sample
                    const float inv_snr = inv_snr_linear(snr_db);
                    float zr[16], zi[16], se_m = 0.f;
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        float hr, hi, mr, mi;
                        eq_channel(nr[i], ni[i], cr[i], ci[i], hr, hi);
                        eq_zf(nr[i], ni[i], hr, hi, zr[i], zi[i]);
                        eq_mmse(nr[i], ni[i], hr, hi, inv_snr, mr, mi);
                        const float dr = mr - cr[i], di = mi - ci[i];
                        se_m = fmaf(dr, dr, fmaf(di, di, se_m));
                    }
                    frame_err(zr, zi, cr, ci, inv_energy, mse, evm, ratio);
                    if (BITS) qpsk_errors<SRC>(a.cfg, zr, zi, bits, errs);
                    acc_add<BITS, NMETH>(acc, OFDMGAN_METHOD_ZF, mse, evm, ratio, errs);
                    err_to_metrics(se_m, inv_energy, mse, evm, ratio);
                    acc_add<BITS, NMETH>(acc, OFDMGAN_METHOD_MMSE, mse, evm, ratio, 0);
                }
            }
        }
        // reconstruct.  The clean frame waits in the thread's own (swizzled) slots of the warp tile meanwhile.
        float xin[2][16], yo[2][16];
        {
            float f[2][16];
#pragma unroll
            for (int i = 0; i < 16; ++i) { f[0][i] = cr[i]; f[1][i] = ci[i]; }
            tile_write_f32(wsm, lane, f);
        }
        if (GEN == OFDMGAN_GEN_F32) {
#pragma unroll
            for (int i = 0; i < 16; ++i) { xin[0][i] = nr[i]; xin[1][i] = ni[i]; }
            gen_fwd_f32_infer(c_g, a.slope, xin, yo);
        } else {
            // Q8.8 by truncation toward zero (proof/verification.py:297-298); back to float by /256
#pragma unroll
            for (int i = 0; i < 16; ++i) { xin[0][i] = truncf(nr[i] * 256.0f); xin[1][i] = truncf(ni[i] * 256.0f); }
            if (GEN == OFDMGAN_GEN_Q_SPEC) gen_fwd_q_spec(c_q, xin, yo); else gen_fwd_q_rtl(c_q, xin, yo);
#pragma unroll
            for (int i = 0; i < 16; ++i) { yo[0][i] *= 0.00390625f; yo[1][i] *= 0.00390625f; }
        }
        if (want_metrics && live) {
            float f[2][16];
            tile_read_f32(wsm, lane, f);
            float mse, evm, ratio;
            int errs = 0;
            frame_err(yo[0], yo[1], f[0], f[1], inv_energy, mse, evm, ratio);
            if (BITS) qpsk_errors<SRC>(a.cfg, yo[0], yo[1], bits, errs);
            acc_add<BITS, NMETH>(acc, OFDMGAN_METHOD_GAN, mse, evm, ratio, errs);
        }
        __syncwarp();
        cta_lockstep(3);
    }
    if (want_metrics) {
        acc_flush<BITS, NMETH>(acc, table, lane);
        __syncthreads();
        double* out = a.partials + (size_t)blockIdx.x * a.n_snr * NM * NC;
        for (int i = threadIdx.x; i < a.n_snr * NM * NC; i += blockDim.x) out[i] = table[i];
    }
}

template <int SRC, int GEN, bool EQ, bool LATE>
static int sim_launch_one(const SimCall& c) {
    cudaStream_t s = c.stream;
    int slot = 0, rc;
    CallGuard guard(s);
    if ((rc = guard.rc)) return rc;
    slot = 0;
    if (GEN == OFDMGAN_GEN_F32) rc = upload_g(c.gparams258, slot, s);
    else if (GEN > 0) rc = upload_q(c.wrom, c.brom, slot, s);
    if (rc) return rc;
    const int grid = grid_for(c.B, ST, SIM_PER_SM);
    // opt in to > 48 KB of dynamic shared memory (per device; a host-side table write, no launch)
    OG_CHECK(cudaFuncSetAttribute(k_sim<SRC, GEN, EQ, LATE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SIM_SMEM));
    const int n = c.n_snr * NM * NC;
    void* partials = nullptr;
    if (GEN >= 0 && c.metrics && (rc = scratch_for_slot(slot, (size_t)grid * n * sizeof(double), 4, &partials))) return rc;
    SimArgs a{};
    a.cfg = *c.cfg;
    a.keys = philox_keys(c.seed);
    a.frame0 = c.frame0;
    a.B = c.B;
    if (c.rand) { a.sym = c.rand->sym; a.bits = c.rand->bits; a.pn = c.rand->pn; a.snr_db = c.rand->snr_db; a.noise = c.rand->noise; a.tx = c.rand->tx; a.fade = c.rand->fade; a.tx_gain = c.rand->tx_gain; }
    a.clean = c.clean; a.noisy = c.noisy; a.snr_out = c.snr;
    a.wslot = slot;
    a.slope = c.slope;
    a.partials = (double*)partials;
    a.n_snr = c.n_snr;
    k_sim<SRC, GEN, EQ, LATE><<<grid, ST, SIM_SMEM, s>>>(a);
    OG_CHECK(cudaGetLastError());
    if (partials) {
        reduce_partials_launch((const double*)partials, grid, n, c.metrics, s);
        OG_CHECK(cudaGetLastError());
    }
    return 0;
}

// Which (generator, equaliser rows, late stages) combinations a translation unit builds.  SIM_TU_BASE: the headline combinations for
// its sources; SIM_TU_EXT (sim_gauss_ext.cu, Gaussian source only - the signal run_benchmark uses): fp32 generator with equaliser
// rows AND late stages (run_benchmark(channel_type='rayleigh' | 'rician' | 'multipath'), benchmark_comparison.py:154), and the integer
// generators with equaliser rows and / or late stages.
inline bool sim_needs_late(const ofdmgan_chan_cfg& c) {
    return c.channel_type != OFDMGAN_CHAN_AWGN || (c.impair & (OFDMGAN_IMPAIR_SALEH | OFDMGAN_IMPAIR_DC | OFDMGAN_IMPAIR_CFO)) != 0;
}
inline bool sim_is_ext_combo(const SimCall& c) {
    const bool eq = c.cfg->equalizers != 0, late = sim_needs_late(*c.cfg);
    if (c.gen_kind == OFDMGAN_GEN_F32) return eq && late;
    if (c.gen_kind == OFDMGAN_GEN_Q_SPEC || c.gen_kind == OFDMGAN_GEN_Q_RTL) return eq || late;
    return false;
}

#ifndef SIM_TU_EXT
template <int SRC>
static int sim_launch_src(const SimCall& c) {
    const bool eq = c.cfg->equalizers != 0;
    // the rarely used stages live in their own instantiations so the headline kernels do not carry them
    const bool late = sim_needs_late(*c.cfg);
    if (sim_is_ext_combo(c)) return SRC == SRC_GAUSS ? sim_launch_gauss_ext(c) : OFDMGAN_E_UNSUPPORTED;
    switch (c.gen_kind) {
        case -1: return late ? sim_launch_one<SRC, -1, false, true>(c) : sim_launch_one<SRC, -1, false, false>(c);
        case OFDMGAN_GEN_F32:
            if (late) return sim_launch_one<SRC, OFDMGAN_GEN_F32, false, true>(c);
            return eq ? sim_launch_one<SRC, OFDMGAN_GEN_F32, true, false>(c) : sim_launch_one<SRC, OFDMGAN_GEN_F32, false, false>(c);
        case OFDMGAN_GEN_Q_SPEC: return sim_launch_one<SRC, OFDMGAN_GEN_Q_SPEC, false, false>(c);
        case OFDMGAN_GEN_Q_RTL: return sim_launch_one<SRC, OFDMGAN_GEN_Q_RTL, false, false>(c);
        default: return OFDMGAN_E_ARG;
    }
}
#else
template <int GEN>
static int sim_launch_ext_gen(const SimCall& c, bool eq, bool late) {
    if (eq && late) return sim_launch_one<SRC_GAUSS, GEN, true, true>(c);
    if (GEN == OFDMGAN_GEN_F32) return OFDMGAN_E_ARG;                // (the other fp32 combinations are base ones)
    return eq ? sim_launch_one<SRC_GAUSS, GEN == OFDMGAN_GEN_F32 ? OFDMGAN_GEN_Q_SPEC : GEN, true, false>(c)
              : sim_launch_one<SRC_GAUSS, GEN == OFDMGAN_GEN_F32 ? OFDMGAN_GEN_Q_SPEC : GEN, false, true>(c);
}
static int sim_launch_ext(const SimCall& c) {
    const bool eq = c.cfg->equalizers != 0, late = sim_needs_late(*c.cfg);
    switch (c.gen_kind) {
        case OFDMGAN_GEN_F32: return sim_launch_ext_gen<OFDMGAN_GEN_F32>(c, eq, late);
        case OFDMGAN_GEN_Q_SPEC: return sim_launch_ext_gen<OFDMGAN_GEN_Q_SPEC>(c, eq, late);
        case OFDMGAN_GEN_Q_RTL: return sim_launch_ext_gen<OFDMGAN_GEN_Q_RTL>(c, eq, late);
        default: return OFDMGAN_E_ARG;
    }
}
#endif

}  // namespace og
