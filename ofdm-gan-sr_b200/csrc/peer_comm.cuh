// Exchange blocks of the data-parallel ranks and the block-wide all-reduce over them (see peer_comm.cu for the protocol).
#pragma once
#include "common.cuh"

struct ofdmgan_comm;

namespace og {

constexpr int PC_MAX_WORLD = 16;
constexpr int PC_MAX_N = 1024;                                   // floats per message (critic 528, generator 264)
constexpr double PC_DEFAULT_TIMEOUT_S = 120.0;                   // a peer that has not arrived after this long is treated as dead

// Every exchanged word is self-validating (the "LL" idea): a 64-bit slot holds {sequence number, float bits}, written with ONE 8-byte
// store - single-copy atomic over NVLink - so a reader simply polls the slot until the sequence number is the one it waits for.  No
// system-scope fence, no separate flag, no barrier between sending and receiving.
struct PeerBlock {                                               // one per rank, device memory
    unsigned long long slot[2][PC_MAX_WORLD][PC_MAX_N];
    unsigned int seq;                                            // calls completed by the owning rank (device-resident: graph replays advance it)
    int error;                                                   // sticky: set when a wait timed out
};

struct PeerPtrs {
    PeerBlock* p[PC_MAX_WORLD];
    long long spin_limit;                                        // the wait limit in SM clocks (OFDMGAN_COMM_TIMEOUT_S x clock rate)
};

__device__ __forceinline__ void st_sys_u64(unsigned long long* a, unsigned long long v) {
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(a), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_sys_u64(const unsigned long long* a) {
    unsigned long long v;
    asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(a) : "memory");
    return v;
}

// Block-wide exchange: every thread of ONE block calls it.  g[0..n) of this rank goes to every peer, the rank-ordered sum comes back
// into g.  `seq` = this rank's call number (mine->seq + 1, >= 1).
// A peer that does not arrive within the wait limit (default 120 s: ranks may legitimately be late by a checkpoint or a validation
// pass - put a barrier after rank-asymmetric work that can take longer) is fatal: the sticky error word is set and the kernel TRAPS,
// so the launch, and every later call on this context, fails with a CUDA error instead of returning stale sums.  Nothing continues
// with a partial result: no rank can apply an optimiser step the others skipped.
// Slot sets alternate by sequence parity: a rank can start call s+1 while a peer still reads call s, and cannot reach call s+2
// before that peer has sent its call s+1 words, i.e. after it finished reading call s.
__device__ __forceinline__ void peer_allreduce_block(const PeerPtrs& peers, int rank, int world, unsigned int seq, float* g, int n) {
    PeerBlock* mine = peers.p[rank];
    const int par = seq & 1u, tid = threadIdx.x;
    for (int i = tid; i < n; i += blockDim.x) {
        const unsigned long long w = ((unsigned long long)seq << 32) | (unsigned long long)__float_as_uint(__ldcg(g + i));
        for (int r = 0; r < world; ++r) st_sys_u64(&peers.p[r]->slot[par][rank][i], w);
    }
    for (int i = tid; i < n; i += blockDim.x) {
        float s = 0.f;
        for (int r = 0; r < world; ++r) {                        // rank order: the same bits on every rank
            const long long t0 = clock64();
            unsigned long long w = ld_sys_u64(&mine->slot[par][r][i]);
            while ((unsigned int)(w >> 32) != seq) {
                if (clock64() - t0 > peers.spin_limit) {
                    mine->error = 1;
                    __threadfence_system();
                    __trap();
                }
                w = ld_sys_u64(&mine->slot[par][r][i]);
            }
            s += __uint_as_float((unsigned int)w);
        }
        g[i] = s;
    }
    __syncthreads();
}

// peers / rank / world of a connected communicator (peer_comm.cu); false if comm is null or not connected
bool comm_view(const ofdmgan_comm* comm, PeerPtrs* peers, int* rank, int* world);

}  // namespace og
