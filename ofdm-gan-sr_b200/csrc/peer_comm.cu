// libofdmgan: gradient all-reduce fused with Adam over NVLink / NVSwitch peer memory (sm_100a).
//
// The training step of train.py:201-305 ends each of its 5 + 1 optimiser steps with a sum of a 528- / 264-float buffer over
// the data-parallel ranks followed by Adam on 521 / 258 parameters.  At 2 KB the collective is pure latency (NCCL: ~25 us per
// call at 8 ranks, a fifth of the step).  Here one 1024-thread CTA per rank does all of it in one launch (peer_comm.cuh):
//   1. every rank stores its buffer straight into slot [parity][rank] of EVERY peer's exchange block (P2P stores over NVLink), each
//      word tagged with the call's sequence number;
//   2. reads the `world` slots of its own block, polling each word until its tag is current, and sums them in rank order - the same
//      order on every rank, so replicas stay bit-identical - writes the total back to the caller's buffer (loss statistics ride
//      along) and applies Adam to the parameters.
// Exchange blocks are cudaMalloc'ed by the library and shared between the one-process-per-GPU ranks through CUDA IPC handles
// that the host side swaps over torch.distributed.  The same block-wide exchange is the middle of the critic tail kernel
// (critic_step.cu, ofdmgan_critic_train_ctr).
#include <cstdlib>
#include <cstring>

#include "peer_comm.cuh"
#include "train_common.cuh"

namespace og {

// step_dev == nullptr: Adam coefficients from the host (coef); else t = *step_dev + 1 is used and stored back (ofdmgan_adam_ctr)
__global__ void __launch_bounds__(1024) k_allreduce_adam(PeerPtrs peers, int rank, int world, float* __restrict__ g,
                                                         int n, float* __restrict__ p, float* __restrict__ m, float* __restrict__ v,
                                                         int n_params, AdamCoef coef, double lr, double b1, double b2, double eps,
                                                         int32_t* __restrict__ step_dev, float grad_scale) {
    PeerBlock* mine = peers.p[rank];
    const unsigned int seq = mine->seq + 1u;                     // written only by this rank's previous launch (stream order)
    const int tid = threadIdx.x;
    int t_adam = 0;
    if (step_dev) { t_adam = *step_dev + 1; coef = adam_coef_dev(lr, b1, b2, eps, t_adam); }
    peer_allreduce_block(peers, rank, world, seq, g, n);         // traps if a peer never arrives: nothing below runs on a partial sum
    if (tid == 0) { mine->seq = seq; if (step_dev) *step_dev = t_adam; }
    for (int i = tid; i < n_params; i += blockDim.x) adam_one(p[i], m[i], v[i], __fmul_rn(g[i], grad_scale), coef);
}

}  // namespace og

using namespace og;

struct ofdmgan_comm {
    int rank, world, device;
    PeerBlock* local;
    PeerPtrs peers;
    bool connected;
};

namespace og {
bool comm_view(const ofdmgan_comm* c, PeerPtrs* peers, int* rank, int* world) {
    if (!c || !c->connected) return false;
    *peers = c->peers; *rank = c->rank; *world = c->world;
    return true;
}
}  // namespace og

extern "C" {

int ofdmgan_comm_create(int rank, int world, ofdmgan_comm** out, void* ipc_handle64) {
    if (!out || !ipc_handle64 || world < 1 || world > PC_MAX_WORLD || rank < 0 || rank >= world) return OFDMGAN_E_ARG;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    ofdmgan_comm* c = new ofdmgan_comm();
    c->rank = rank; c->world = world; c->connected = false; c->local = nullptr;
    for (int r = 0; r < PC_MAX_WORLD; ++r) c->peers.p[r] = nullptr;
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaGetDevice(&c->device);
    {                                                            // the wait limit in SM clocks
        double secs = PC_DEFAULT_TIMEOUT_S;
        if (const char* env = getenv("OFDMGAN_COMM_TIMEOUT_S")) { const double v = atof(env); if (v > 0.0) secs = v; }
        int khz = 0;
        if (e == cudaSuccess) e = cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, c->device);
        c->peers.spin_limit = (long long)(secs * 1e3 * (double)(khz > 0 ? khz : 2000000));
    }
    if (e == cudaSuccess) e = cudaMalloc((void**)&c->local, sizeof(PeerBlock));
    if (e == cudaSuccess) e = cudaMemset(c->local, 0, sizeof(PeerBlock));
    if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, c->local);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();          // the zeroed block must be in place before any peer writes to it
    if (e != cudaSuccess) {                                      // nothing is handed out on failure
        if (c->local) cudaFree(c->local);
        delete c;
        (void)cudaGetLastError();
        return (int)e;
    }
    memcpy(ipc_handle64, &h, 64);
    c->peers.p[rank] = c->local;
    *out = c;
    return 0;
}

int ofdmgan_comm_connect(ofdmgan_comm* c, const void* all_handles) {
    if (!c || !all_handles) return OFDMGAN_E_ARG;
    for (int r = 0; r < c->world; ++r) {
        if (r == c->rank) continue;
        cudaIpcMemHandle_t h;
        memcpy(&h, (const char*)all_handles + 64 * r, 64);
        void* ptr = nullptr;
        OG_CHECK(cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess));
        c->peers.p[r] = (PeerBlock*)ptr;
    }
    c->connected = true;
    return 0;
}

int ofdmgan_comm_destroy(ofdmgan_comm* c) {
    if (!c) return 0;
    cudaDeviceSynchronize();
    for (int r = 0; r < c->world; ++r)
        if (r != c->rank && c->peers.p[r]) cudaIpcCloseMemHandle(c->peers.p[r]);
    cudaFree(c->local);
    delete c;
    return 0;
}

// 0 while every wait so far completed.  A wait that ran out traps its kernel, so afterwards this (like every call on the context)
// returns the CUDA error of the failed launch; OFDMGAN_E_COMM if the error word could still be read.  Synchronises the stream.
int ofdmgan_comm_check(ofdmgan_comm* c, void* stream) {
    if (!c) return OFDMGAN_E_ARG;
    int err = 0;
    OG_CHECK(cudaMemcpyAsync(&err, &c->local->error, sizeof(int), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    OG_CHECK(cudaStreamSynchronize((cudaStream_t)stream));
    return err ? OFDMGAN_E_COMM : 0;
}

static int allreduce_adam_impl(ofdmgan_comm* c, float* g_dev, int n, float* p_dev, float* m_dev, float* v_dev, int n_params, double lr,
                               double beta1, double beta2, double eps, int step, int32_t* step_dev, float grad_scale, void* stream) {
    if (!c || !c->connected || !g_dev || n < 1 || n > PC_MAX_N || n_params < 0 || n_params > n) return OFDMGAN_E_ARG;
    if (n_params > 0 && (!p_dev || !m_dev || !v_dev || (!step_dev && step < 1))) return OFDMGAN_E_ARG;
    k_allreduce_adam<<<1, 1024, 0, (cudaStream_t)stream>>>(c->peers, c->rank, c->world, g_dev, n, p_dev, m_dev, v_dev, n_params,
                                                           adam_coef(lr, beta1, beta2, eps, step > 0 ? step : 1), lr, beta1, beta2, eps,
                                                           n_params > 0 ? step_dev : nullptr, grad_scale);
    return (int)cudaGetLastError();
}

int ofdmgan_allreduce_adam(ofdmgan_comm* c, float* g_dev, int n, float* p_dev, float* m_dev, float* v_dev, int n_params, double lr,
                           double beta1, double beta2, double eps, int step, float grad_scale, void* stream) {
    return allreduce_adam_impl(c, g_dev, n, p_dev, m_dev, v_dev, n_params, lr, beta1, beta2, eps, step, nullptr, grad_scale, stream);
}

int ofdmgan_allreduce_adam_ctr(ofdmgan_comm* c, float* g_dev, int n, float* p_dev, float* m_dev, float* v_dev, int n_params, double lr,
                               double beta1, double beta2, double eps, int32_t* step_dev, float grad_scale, void* stream) {
    if (!step_dev) return OFDMGAN_E_ARG;
    return allreduce_adam_impl(c, g_dev, n, p_dev, m_dev, v_dev, n_params, lr, beta1, beta2, eps, 0, step_dev, grad_scale, stream);
}

}  // extern "C"
