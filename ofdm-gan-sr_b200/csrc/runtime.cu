// libofdmgan runtime: device query, per-stream weight slots, library-owned scratch, error strings and the FFMA
// issue-rate microbenchmark that the fp32 roofline uses as its denominator.
#include <map>
#include <mutex>
#include <utility>

#include "common.cuh"

namespace og {

static std::mutex g_mu;
static DeviceInfo g_dev[64];
static void* g_scratch[64][OG_NSLOT][12];
static size_t g_scratch_bytes[64][OG_NSLOT][12];

const DeviceInfo& device_info(int* err) {
    static DeviceInfo none;
    int d = -1;
    cudaError_t e = cudaGetDevice(&d);
    if (e != cudaSuccess || d < 0 || d >= 64) {
        if (err) *err = e != cudaSuccess ? (int)e : OFDMGAN_E_ARG;
        return none;
    }
    std::lock_guard<std::mutex> lk(g_mu);
    if (g_dev[d].device != d) {
        int sms = 0;
        e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, d);
        if (e != cudaSuccess) {
            if (err) *err = (int)e;
            return none;
        }
        g_dev[d].device = d;
        g_dev[d].sms = sms;
    }
    if (err) *err = 0;
    return g_dev[d];
}

int grid_for(int64_t B, int threads, int per_sm) {
    int err = 0;
    const DeviceInfo& di = device_info(&err);
    int sms = err ? 148 : di.sms;
    int64_t tiles = (B + threads - 1) / threads;
    int64_t cap = (int64_t)sms * per_sm;           // persistent grid: a multiple of the SM count
    if (tiles < 1) tiles = 1;
    return (int)(tiles < cap ? tiles : cap);
}

static std::recursive_mutex g_call_mu;
static cudaStream_t g_last_stream[64];
static bool g_have_last[64];
static cudaEvent_t g_order_ev[64];

CallGuard::CallGuard(cudaStream_t s) : rc(0) {
    g_call_mu.lock();
    int d = -1;
    cudaError_t e = cudaGetDevice(&d);
    if (e != cudaSuccess || d < 0 || d >= 64) {
        rc = e != cudaSuccess ? (int)e : OFDMGAN_E_ARG;
        return;
    }
    if (g_have_last[d] && g_last_stream[d] != s) {
        // order this stream after everything the previous stream was given (skipped while either side is being
        // captured into a CUDA graph: a captured stream may not wait on / be waited on by work outside the capture)
        cudaStreamCaptureStatus c0 = cudaStreamCaptureStatusNone, c1 = cudaStreamCaptureStatusNone;
        cudaStreamIsCapturing(g_last_stream[d], &c0);
        cudaStreamIsCapturing(s, &c1);
        (void)cudaGetLastError();
        if (c0 == cudaStreamCaptureStatusNone && c1 == cudaStreamCaptureStatusNone) {
            if (!g_order_ev[d]) e = cudaEventCreateWithFlags(&g_order_ev[d], cudaEventDisableTiming);
            if (e == cudaSuccess) e = cudaEventRecord(g_order_ev[d], g_last_stream[d]);
            if (e == cudaSuccess) e = cudaStreamWaitEvent(s, g_order_ev[d], 0);
            if (e != cudaSuccess) (void)cudaGetLastError();     // e.g. the previous stream was destroyed: nothing left to order against
        }
    }
    g_last_stream[d] = s;
    g_have_last[d] = true;
}

CallGuard::~CallGuard() { g_call_mu.unlock(); }

int scratch_for_slot(int slot, size_t bytes, int which, void** ptr) {
    int d = -1;
    OG_CHECK(cudaGetDevice(&d));
    if (d < 0 || d >= 64 || slot < 0 || slot >= OG_NSLOT || which < 0 || which >= 12) return OFDMGAN_E_ARG;
    std::lock_guard<std::mutex> lk(g_mu);
    if (g_scratch_bytes[d][slot][which] < bytes) {
        // growth only happens on the first calls (sizes are bounded by the grid); never inside stream capture
        if (g_scratch[d][slot][which]) OG_CHECK(cudaFree(g_scratch[d][slot][which]));
        g_scratch[d][slot][which] = nullptr;
        g_scratch_bytes[d][slot][which] = 0;
        void* p = nullptr;
        OG_CHECK(cudaMalloc(&p, bytes));
        g_scratch[d][slot][which] = p;
        g_scratch_bytes[d][slot][which] = bytes;
    }
    *ptr = g_scratch[d][slot][which];
    return 0;
}

namespace {
struct StreamScratch {
    void* p[4] = {nullptr, nullptr, nullptr, nullptr};
    size_t bytes[4] = {0, 0, 0, 0};
};
std::map<std::pair<int, cudaStream_t>, StreamScratch> g_stream_scratch;
}  // namespace

int scratch_for_stream(cudaStream_t s, size_t bytes, int which, void** ptr) {
    int d = -1;
    OG_CHECK(cudaGetDevice(&d));
    if (which < 0 || which >= 4) return OFDMGAN_E_ARG;
    std::lock_guard<std::mutex> lk(g_mu);
    StreamScratch& e = g_stream_scratch[std::make_pair(d, s)];
    if (e.bytes[which] < bytes) {                                 // growth only on the first calls of a stream; never inside capture
        if (e.p[which]) OG_CHECK(cudaFree(e.p[which]));
        e.p[which] = nullptr;
        e.bytes[which] = 0;
        void* p = nullptr;
        OG_CHECK(cudaMalloc(&p, bytes));
        e.p[which] = p;
        e.bytes[which] = bytes;
    }
    *ptr = e.p[which];
    return 0;
}

int to_device_f32(const float* src, int n, int slot, int which, cudaStream_t s, const float** dev) {
    if (!src) return OFDMGAN_E_ARG;
    cudaPointerAttributes at;
    cudaError_t e = cudaPointerGetAttributes(&at, src);
    if (e == cudaSuccess && (at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged)) {
        *dev = src;
        return 0;
    }
    if (e != cudaSuccess) (void)cudaGetLastError();   // plain malloc'ed memory on old drivers: treat as host
    void* buf = nullptr;
    int rc = scratch_for_slot(slot, (size_t)n * sizeof(float), which, &buf);
    if (rc) return rc;
    OG_CHECK(cudaMemcpyAsync(buf, src, (size_t)n * sizeof(float), cudaMemcpyHostToDevice, s));
    *dev = (const float*)buf;
    return 0;
}

// ---- FFMA issue-rate microbenchmark --------------------------------------------------------------------------
// 8 independent accumulator chains per thread, register x uniform operand form (the form the network kernels
// use: FFMA R,R,UR,R), 1024 threads per SM.
__global__ void __launch_bounds__(256) k_ffma_peak(float* out, int iters, float a, float b) {
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = (float)(threadIdx.x + j);
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int r = 0; r < 16; ++r)
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = fmaf(v[j], a, b);
    }
    float s = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) s += v[j];
    if (s == 123.456f) out[0] = s;   // never true; keeps the chains alive
}

}  // namespace og

using namespace og;

extern "C" {

int ofdmgan_abi_version(void) { return OFDMGAN_ABI_VERSION; }

int ofdmgan_device_sms(void) {
    int err = 0;
    const DeviceInfo& di = device_info(&err);
    return err ? (err > 0 ? -err : err) : di.sms;
}

const char* ofdmgan_error_string(int code) {
    if (code == 0) return "ok";
    if (code == OFDMGAN_E_ARG) return "ofdmgan: invalid argument";
    if (code == OFDMGAN_E_STREAMS) return "ofdmgan: too many distinct streams in use";
    if (code == OFDMGAN_E_UNSUPPORTED) return "ofdmgan: configuration not supported by this build";
    if (code == OFDMGAN_E_COMM) return "ofdmgan: a data-parallel peer did not arrive (exchange wait timed out)";
    if (code > 0) return cudaGetErrorString((cudaError_t)code);
    return "ofdmgan: unknown error";
}

int ofdmgan_ffma_peak(int iters, double* tflops_host, void* stream) {
    if (!tflops_host || iters < 1) return OFDMGAN_E_ARG;
    cudaStream_t s = (cudaStream_t)stream;
    int err = 0;
    const DeviceInfo& di = device_info(&err);
    if (err) return err;
    float* out = nullptr;
    OG_CHECK(cudaMalloc(&out, sizeof(float)));
    int blocks = di.sms * 4;
    cudaEvent_t e0, e1;
    OG_CHECK(cudaEventCreate(&e0));
    OG_CHECK(cudaEventCreate(&e1));
    k_ffma_peak<<<blocks, 256, 0, s>>>(out, iters / 8 + 1, 0.999f, 0.001f);   // warm-up
    OG_CHECK(cudaEventRecord(e0, s));
    k_ffma_peak<<<blocks, 256, 0, s>>>(out, iters, 0.999f, 0.001f);
    OG_CHECK(cudaEventRecord(e1, s));
    OG_CHECK(cudaEventSynchronize(e1));
    float ms = 0;
    OG_CHECK(cudaEventElapsedTime(&ms, e0, e1));
    double flop = 2.0 * 16.0 * 8.0 * (double)iters * 256.0 * (double)blocks;
    *tflops_host = flop / ((double)ms * 1e-3) / 1e12;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    OG_CHECK(cudaFree(out));
    return 0;
}

}  // extern "C"
