/*
 * seqprod_cuda.cu — the C-ABI batching layer (include/b200seqprod.h) over the sm_100a kernels.
 *
 * Replaces the icp_sal / cpaDc submission loop and the USDM/SVM buffer management of the
 * reference (/root/reference/src/qatseqprod.c:685-822 buffers, :1222-1227 staging memcpy,
 * :1245-1249 submit, :1263-1272 poll): an engine owns one CUDA stream, device scratch and pinned
 * host staging; a batch of independent blocks is one kernel launch.
 */
#include "b200seqprod.h"
#include "lz77_kernels.cuh"

#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <mutex>
#include <thread>
#include <vector>

namespace {

thread_local char g_err[256] = "";

int fail(int code, const char *what, cudaError_t ce = cudaSuccess)
{
    if (ce != cudaSuccess) snprintf(g_err, sizeof g_err, "%s: %s", what, cudaGetErrorString(ce));
    else snprintf(g_err, sizeof g_err, "%s", what);
    return code;
}

#define CU_TRY(expr, what)                                            \
    do {                                                              \
        cudaError_t ce_ = (expr);                                     \
        if (ce_ != cudaSuccess) return fail(B200SP_ECUDA, what, ce_); \
    } while (0)

// Every C-ABI entry point runs on its engine's device and puts the caller's current device back on the way out:
// a host process that uses CUDA on another GPU must not find its thread switched.
struct DeviceGuard {
    int prev = -1;
    bool switched = false;
    cudaError_t enter(int dev)
    {
        if (cudaGetDevice(&prev) != cudaSuccess) { (void)cudaGetLastError(); prev = -1; }
        if (prev == dev) return cudaSuccess;
        const cudaError_t ce = cudaSetDevice(dev);
        switched = ce == cudaSuccess;
        return ce;
    }
    ~DeviceGuard() { if (switched && prev >= 0) cudaSetDevice(prev); }
};

// Completion wait with the reference's budget (QZSTD timeout, /root/reference/src/qatseqprod.c:107, :1099-1104,
// :1267-1270: 2 s, then ERROR): poll instead of blocking forever.
constexpr double kTimeoutSeconds = 2.0;

double now_seconds()
{
    timespec t;
    clock_gettime(CLOCK_MONOTONIC, &t);
    return static_cast<double>(t.tv_sec) + 1e-9 * static_cast<double>(t.tv_nsec);
}

// 0 = done, 1 = timed out, otherwise the CUDA error
// Polling discipline shared by both waits: a few dozen back-to-back queries catch short waits quickly, then the
// thread sleeps between queries with a slowly growing interval (20 us ... 200 us).  Every query takes a driver lock: with
// one or two engines per compressing thread, thousands of spinning queries slow down everybody's launches.
struct Backoff {
    double deadline = now_seconds() + kTimeoutSeconds;
    unsigned polls = 0;
    long sleepNs = 20000;
    bool expired()          // call after an unsuccessful query; true = give up
    {
        if (++polls <= 64) return false;
        if (now_seconds() > deadline) return true;
        timespec ts = {0, sleepNs};
        nanosleep(&ts, nullptr);
        if ((polls & 7u) == 0u && sleepNs < 200000) sleepNs += sleepNs / 2;     // 20 us for the first waits of a call, 200 us after ~2 ms
        return false;
    }
};

int wait_event(cudaEvent_t ev, cudaError_t *ce)
{
    Backoff b;
    for (;;) {
        const cudaError_t q = cudaEventQuery(ev);
        if (q == cudaSuccess) return 0;
        if (q != cudaErrorNotReady) { *ce = q; return 2; }
        if (b.expired()) return 1;
    }
}

int wait_stream(cudaStream_t st, cudaError_t *ce)
{
    Backoff b;
    for (;;) {
        const cudaError_t q = cudaStreamQuery(st);
        if (q == cudaSuccess) return 0;
        if (q != cudaErrorNotReady) { *ce = q; return 2; }
        if (b.expired()) return 1;
    }
}

// memcpy of a large range by a few threads (results landing in pageable caller memory)
void parallel_copy(void *dst, const void *src, size_t bytes)
{
    constexpr size_t kPerThread = 4u << 20;
    unsigned n = static_cast<unsigned>(bytes / kPerThread);
    const unsigned hw = std::thread::hardware_concurrency();
    if (n > 8) n = 8;
    if (hw && n > hw) n = hw;
    if (n <= 1) { memcpy(dst, src, bytes); return; }
    std::vector<std::thread> th;
    const size_t per = (bytes / n + 63) & ~static_cast<size_t>(63);
    for (unsigned i = 0; i < n; i++) {
        const size_t off = static_cast<size_t>(i) * per;
        if (off >= bytes) break;
        const size_t len = off + per < bytes ? per : bytes - off;
        th.emplace_back([=] { memcpy(static_cast<char *>(dst) + off, static_cast<const char *>(src) + off, len); });
    }
    for (auto &t : th) t.join();
}

// What an engine needs to know about a device, asked once per process: cudaGetDeviceProperties takes milliseconds
// and a driver-wide lock, and engines are created by the dozen (one or two per compressing thread).
struct DeviceFacts { int state; int numSMs; };        // state: 0 unknown, 1 usable, -1 not usable
static DeviceFacts g_facts[64];
static std::mutex g_factsMu;

static DeviceFacts device_facts(int dev)
{
    if (dev < 0 || dev >= 64) return DeviceFacts{-1, 0};
    std::lock_guard<std::mutex> lock(g_factsMu);
    if (g_facts[dev].state == 0) {
        int major = 0, minor = 0, optin = 0, sms = 0;
        const bool ok = cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) == cudaSuccess &&
                        cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev) == cudaSuccess &&
                        cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev) == cudaSuccess &&
                        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess;
        if (!ok) { (void)cudaGetLastError(); return DeviceFacts{-1, 0}; }       // not cached: the driver may come up later
        // the cubin is sm_100a only; the parser needs the 227 KB opt-in shared memory carve-out
        g_facts[dev].state = (major == 10 && minor == 0 && static_cast<size_t>(optin) >= static_cast<size_t>(b200sp::kSmemTotal)) ? 1 : -1;
        g_facts[dev].numSMs = sms;
    }
    return g_facts[dev];
}

bool device_usable(int dev)
{
    return device_facts(dev).state == 1;
}

// ---- wire format: offset | litLength << 17 | matchLength << 35 ----------------------------
// The post-processing kernels of chunk k are queued while the parser CTAs of chunk k+1 own every SM (a parser
// CTA takes the whole register file): they are small and sit on a high-priority stream, so they get the first
// SMs that free up.
constexpr int kPostThreads = 128;

__global__ void __launch_bounds__(32) scan_counts_kernel(const uint32_t *__restrict__ counts, uint32_t nBlocks,
                                                         unsigned long long *__restrict__ offsets)
{
    // one warp, shuffles only: no shared memory, so the CTA fits in the ~1 KB the parser leaves free on an SM
    const uint32_t lane = threadIdx.x;
    unsigned long long running = 0;
    for (uint32_t base = 0; base < nBlocks; base += 32) {
        const uint32_t i = base + lane;
        const unsigned long long v = i < nBlocks ? counts[i] : 0;
        unsigned long long incl = v;
        for (int d = 1; d < 32; d <<= 1) {
            const unsigned long long o = __shfl_up_sync(0xFFFFFFFFu, incl, d);
            if (lane >= (uint32_t)d) incl += o;
        }
        if (i < nBlocks) offsets[i] = running + incl - v;
        running += __shfl_sync(0xFFFFFFFFu, incl, 31);
    }
    if (lane == 0) offsets[nBlocks] = running;
}

__global__ void __launch_bounds__(kPostThreads) pack_kernel(const uint4 *__restrict__ seqs, uint64_t seqStride,
                            const uint32_t *__restrict__ counts,
                            const unsigned long long *__restrict__ offsets, uint32_t nBlocks,
                            unsigned long long *__restrict__ packed)
{
    for (uint32_t b = blockIdx.x; b < nBlocks; b += gridDim.x) {
        const uint4 *s = seqs + static_cast<uint64_t>(b) * seqStride;
        unsigned long long *o = packed + offsets[b];
        const uint32_t c = counts[b];
        for (uint32_t i = threadIdx.x; i < c; i += blockDim.x) {
            const uint4 q = s[i];
            o[i] = static_cast<unsigned long long>(q.x) | (static_cast<unsigned long long>(q.y) << 17) |
                   (static_cast<unsigned long long>(q.z) << 35);
        }
    }
}

// The same gather without repacking: dense ZSTD_Sequence[] (16 bytes each), the array the plugin API hands to libzstd.
__global__ void __launch_bounds__(kPostThreads) dense_kernel(const uint4 *__restrict__ seqs, uint64_t seqStride,
                             const uint32_t *__restrict__ counts,
                             const unsigned long long *__restrict__ offsets, uint32_t nBlocks,
                             uint4 *__restrict__ dense)
{
    for (uint32_t b = blockIdx.x; b < nBlocks; b += gridDim.x) {
        const uint4 *s = seqs + static_cast<uint64_t>(b) * seqStride;
        uint4 *o = dense + offsets[b];
        const uint32_t c = counts[b];
        for (uint32_t i = threadIdx.x; i < c; i += blockDim.x) o[i] = s[i];
    }
}

__global__ void count_bad_kernel(const uint32_t *__restrict__ bad, uint32_t nBlocks, uint32_t *__restrict__ total)
{
    uint32_t n = 0;
    for (uint32_t b = blockIdx.x * blockDim.x + threadIdx.x; b < nBlocks; b += gridDim.x * blockDim.x) n += bad[b] != 0u;
    if (n) atomicAdd(total, n);
}

// ---- on-device verification: one warp replays one block ------------------------------------
__global__ void verify_kernel(const uint8_t *__restrict__ src, uint64_t stride, uint64_t totalSize,
                              uint32_t blockSize, const uint32_t *__restrict__ sizes, uint32_t nBlocks,
                              const uint4 *__restrict__ seqs, uint64_t seqStride,
                              const uint32_t *__restrict__ counts, uint32_t *__restrict__ bad)
{
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t warpsPerGrid = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t b = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; b < nBlocks; b += warpsPerGrid) {
        uint32_t n;
        if (sizes) n = sizes[b];
        else {
            const uint64_t start = static_cast<uint64_t>(b) * stride;
            const uint64_t left = totalSize > start ? totalSize - start : 0;
            n = left < blockSize ? static_cast<uint32_t>(left) : blockSize;
        }
        const uint8_t *in = src + static_cast<uint64_t>(b) * stride;
        const uint4 *s = seqs + static_cast<uint64_t>(b) * seqStride;
        const uint32_t c = counts[b];
        uint32_t err = (c == 0) ? 1u : 0u;
        // pass 1: positions by warp scan, chunk of 32 sequences at a time
        uint32_t pos = 0;
        for (uint32_t base = 0; base < c && !err; base += 32) {
            const uint32_t i = base + lane;
            uint4 q = make_uint4(0, 0, 0, 0);
            if (i < c) q = s[i];
            uint32_t span = q.y + q.z, incl = span;
            for (int d = 1; d < 32; d <<= 1) {
                uint32_t o = __shfl_up_sync(0xFFFFFFFFu, incl, d);
                if (lane >= (uint32_t)d) incl += o;
            }
            const uint32_t start = pos + incl - span + q.y;     // first byte of this match
            uint32_t e = 0;
            if (i < c) {
                if (q.z == 0) { if (q.x != 0 || i + 1 != c) e = 7; }
                else if (q.z < 3) e = 2;
                else if (q.x == 0) e = 3;
                else if (q.x > start) e = 4;
                else if (static_cast<uint64_t>(start) + q.z > n) e = 6;
                else {
                    for (uint32_t k = 0; k < q.z; k++)
                        if (in[start + k] != in[start + k - q.x]) { e = 5; break; }
                }
            }
            err = __reduce_max_sync(0xFFFFFFFFu, e);
            pos += __shfl_sync(0xFFFFFFFFu, incl, 31);
        }
        if (!err && pos != n) err = 6;
        if (lane == 0) bad[b] = err;
    }
}

template <typename T>
static cudaError_t grow_dev(T *&ptr, size_t &cap, size_t want)
{
    if (want <= cap) return cudaSuccess;
    cudaFree(ptr);
    ptr = nullptr; cap = 0;
    size_t newCap = want + want / 4;
    cudaError_t ce = cudaMalloc(&ptr, newCap * sizeof(T));
    if (ce == cudaSuccess) cap = newCap;
    return ce;
}

template <typename T>
static cudaError_t grow_host(T *&ptr, size_t &cap, size_t want)
{
    if (want <= cap) return cudaSuccess;
    cudaFreeHost(ptr);
    ptr = nullptr; cap = 0;
    size_t newCap = want + want / 4;
    cudaError_t ce = cudaMallocHost(&ptr, newCap * sizeof(T));
    if (ce == cudaSuccess) cap = newCap;
    return ce;
}

}  // namespace

// --------------------------------------------------------------------------------------------
constexpr int kMaxChunks = 32;

struct b200sp_engine {
    int device;
    int numSMs;
    cudaStream_t stream;         // compute
    cudaStream_t sIn, sOut;      // host->device / device->host copies of the pipelined host path
    cudaStream_t sParse[2];      // parser launches of consecutive chunks alternate: chunk k+1's CTAs move in as chunk k's retire
    cudaStream_t sPost;          // scan + pack + small D2H of finished chunks (high priority, co-resident with the parser)
    cudaEvent_t evIn[kMaxChunks], evParsed[kMaxChunks], evDone[kMaxChunks];
    unsigned int *d_work;        // dynamic scheduler counter (+ developer role counters)
    unsigned int *d_chunkWork;   // [kMaxChunks] scheduler counters of the pipelined host path
    // the parser's sorted-table scratch (b200sp::kSortedCap entries per CTA), one area per stream that can hold a
    // launch in flight: [0] the engine's / the caller's stream, [1], [2] the two parser streams of the host path
    uint32_t *d_sorted[3]; size_t d_sortedCap[3];
    // host-path scratch (grown on demand)
    uint8_t *d_src;      size_t d_srcCap;
    uint4 *d_seqs;       size_t d_seqsCap;      // entries
    uint32_t *d_counts;  size_t d_countsCap;
    unsigned long long *d_offsets; size_t d_offsetsCap;
    unsigned long long *d_packed; size_t d_packedCap;   // entries
    uint8_t *h_stage;    size_t h_stageCap;     // pinned staging for pageable inputs of the host path
    uint8_t *h_slots;    size_t h_slotsCap;     // pinned staging of the scattered-block path (its own buffer: the address handed out stays valid)
    uint32_t stageSlots;                        // slots reserved by b200sp_stage_reserve (scattered-block path)
    uint32_t *d_bad;     size_t d_badCap;       // verify-on-return: per-block verdicts + [0] of d_badTotal
    uint32_t *d_badTotal;
    uint32_t *h_badTotal;                       // pinned
    uint32_t *h_flag;                           // pinned copy of the kernel's error flag
    int verify;                                 // QZSTD_VERIFY / b200sp_engine_set_verify: replay every block on the device before returning
    uint32_t *h_counts;  size_t h_countsCap;
    unsigned long long *h_offsets; size_t h_offsetsCap;  // chunk-local (pinned)
    unsigned long long *h_goffsets; size_t h_goffsetsCap; // global, handed to the caller
    unsigned long long *h_packed; size_t h_packedCap;
};

extern "C" {

const char *b200sp_error_string(void) { return g_err; }
const char *b200sp_version(void) { return "b200seqprod 0.2.0"; }

int b200sp_driver_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { (void)cudaGetLastError(); return 0; }
    return n;
}

int b200sp_device_count(void)
{
    int n = 0;
    cudaError_t ce = cudaGetDeviceCount(&n);
    if (ce != cudaSuccess) { (void)cudaGetLastError(); return fail(B200SP_ENODEVICE, "cudaGetDeviceCount", ce); }
    int usable = 0;
    for (int d = 0; d < n; d++) usable += device_usable(d) ? 1 : 0;
    return usable;
}

int b200sp_usable_devices(int *devices, int capacity)
{
    int n = 0;
    cudaError_t ce = cudaGetDeviceCount(&n);
    if (ce != cudaSuccess) { (void)cudaGetLastError(); return fail(B200SP_ENODEVICE, "cudaGetDeviceCount", ce); }
    int usable = 0;
    for (int d = 0; d < n; d++)
        if (device_usable(d)) { if (devices && usable < capacity) devices[usable] = d; usable++; }
    return usable;
}

int b200sp_warmup(int device)
{
    int n = 0;
    cudaError_t ce = cudaGetDeviceCount(&n);
    if (ce != cudaSuccess || n == 0) { (void)cudaGetLastError(); return fail(B200SP_ENODEVICE, "no CUDA device", ce); }
    if (device < 0 || device >= n) return fail(B200SP_EINVAL, "warmup: device index out of range");
    if (!device_usable(device)) return fail(B200SP_EUNSUPPORTED, "device is not sm_100 with 227 KB shared memory per CTA");
    DeviceGuard guard;
    CU_TRY(guard.enter(device), "cudaSetDevice");
    CU_TRY(cudaFree(nullptr), "context creation");
    CU_TRY(b200sp::configure_kernels(), "cudaFuncSetAttribute(max dynamic smem)");      // loads the module
    return B200SP_OK;
}

int b200sp_engine_create(int device, b200sp_engine **out)
{
    if (!out) return fail(B200SP_EINVAL, "engine_create: null out pointer");
    *out = nullptr;
    int n = 0;
    cudaError_t ce = cudaGetDeviceCount(&n);
    if (ce != cudaSuccess || n == 0) { (void)cudaGetLastError(); return fail(B200SP_ENODEVICE, "no CUDA device", ce); }
    if (device < 0 || device >= n) return fail(B200SP_EINVAL, "engine_create: device index out of range");
    if (!device_usable(device)) return fail(B200SP_EUNSUPPORTED, "device is not sm_100 with 227 KB shared memory per CTA");
    DeviceGuard guard;
    CU_TRY(guard.enter(device), "cudaSetDevice");
    {
        static bool configured[64];
        std::lock_guard<std::mutex> lock(g_factsMu);
        if (!configured[device]) {
            CU_TRY(b200sp::configure_kernels(), "cudaFuncSetAttribute(max dynamic smem)");
            configured[device] = true;
        }
    }
    b200sp_engine *e = static_cast<b200sp_engine *>(calloc(1, sizeof(b200sp_engine)));
    if (!e) return fail(B200SP_ENOMEM, "engine_create: out of host memory");
    e->device = device;
    { const char *v = getenv("QZSTD_VERIFY"); e->verify = (v && *v && *v != '0') ? 1 : 0; }
    e->numSMs = device_facts(device).numSMs;
    ce = cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking);
    if (ce == cudaSuccess) ce = cudaStreamCreateWithFlags(&e->sIn, cudaStreamNonBlocking);
    if (ce == cudaSuccess) ce = cudaStreamCreateWithFlags(&e->sOut, cudaStreamNonBlocking);
    if (ce == cudaSuccess) ce = cudaStreamCreateWithFlags(&e->sParse[0], cudaStreamNonBlocking);
    if (ce == cudaSuccess) ce = cudaStreamCreateWithFlags(&e->sParse[1], cudaStreamNonBlocking);
    if (ce == cudaSuccess) {
        int lo = 0, hi = 0;
        cudaDeviceGetStreamPriorityRange(&lo, &hi);
        ce = cudaStreamCreateWithPriority(&e->sPost, cudaStreamNonBlocking, hi);
    }
    for (int k = 0; k < kMaxChunks && ce == cudaSuccess; k++) {
        ce = cudaEventCreateWithFlags(&e->evIn[k], cudaEventDisableTiming);
        if (ce == cudaSuccess) ce = cudaEventCreateWithFlags(&e->evParsed[k], cudaEventDisableTiming);
        if (ce == cudaSuccess) ce = cudaEventCreateWithFlags(&e->evDone[k], cudaEventDisableTiming);
    }
    if (ce == cudaSuccess) ce = cudaMalloc(&e->d_work, 256);
    if (ce == cudaSuccess) ce = cudaMemsetAsync(e->d_work, 0, 256, e->stream);
    if (ce == cudaSuccess) ce = cudaStreamSynchronize(e->stream);
    if (ce == cudaSuccess) ce = cudaMalloc(&e->d_chunkWork, kMaxChunks * sizeof(unsigned int));
    if (ce == cudaSuccess) ce = cudaMallocHost(&e->h_flag, sizeof(uint32_t));
    if (ce == cudaSuccess) *e->h_flag = 0;
    if (ce != cudaSuccess) { b200sp_engine_destroy(e); return fail(B200SP_ECUDA, "engine_create", ce); }
    *out = e;
    return B200SP_OK;
}

void b200sp_engine_destroy(b200sp_engine *e)
{
    if (!e) return;
    DeviceGuard guard;
    guard.enter(e->device);
    if (e->stream) { cudaStreamSynchronize(e->stream); cudaStreamDestroy(e->stream); }
    if (e->sIn) cudaStreamDestroy(e->sIn);
    if (e->sOut) { cudaStreamSynchronize(e->sOut); cudaStreamDestroy(e->sOut); }
    for (int i = 0; i < 2; i++) if (e->sParse[i]) { cudaStreamSynchronize(e->sParse[i]); cudaStreamDestroy(e->sParse[i]); }
    if (e->sPost) { cudaStreamSynchronize(e->sPost); cudaStreamDestroy(e->sPost); }
    for (int k = 0; k < kMaxChunks; k++) {
        if (e->evIn[k]) cudaEventDestroy(e->evIn[k]);
        if (e->evParsed[k]) cudaEventDestroy(e->evParsed[k]);
        if (e->evDone[k]) cudaEventDestroy(e->evDone[k]);
    }
    cudaFree(e->d_chunkWork);
    for (int i = 0; i < 3; i++) cudaFree(e->d_sorted[i]);
    free(e->h_goffsets);
    cudaFree(e->d_work); cudaFree(e->d_src); cudaFree(e->d_seqs); cudaFree(e->d_counts);
    cudaFree(e->d_offsets); cudaFree(e->d_packed);
    cudaFreeHost(e->h_flag); cudaFreeHost(e->h_stage); cudaFreeHost(e->h_slots); cudaFreeHost(e->h_badTotal); cudaFree(e->d_bad); cudaFree(e->d_badTotal); cudaFreeHost(e->h_counts); cudaFreeHost(e->h_offsets); cudaFreeHost(e->h_packed);
    free(e);
}

int b200sp_engine_device(const b200sp_engine *e) { return e ? e->device : -1; }
int b200sp_engine_sm_count(const b200sp_engine *e) { return e ? e->numSMs : 0; }

static int check_batch(const void *d_src, uint32_t blockSize, uint64_t stride, const uint32_t *d_sizes,
                       uint32_t nBlocks, uint64_t seqStride)
{
    if (nBlocks == 0) return B200SP_OK;
    if (!d_src) return fail(B200SP_EINVAL, "null source");
    if (reinterpret_cast<uintptr_t>(d_src) & 15u) return fail(B200SP_EINVAL, "source must be 16-byte aligned");
    if (stride & 15u) return fail(B200SP_EINVAL, "stride must be a multiple of 16");
    if (!d_sizes && (blockSize == 0 || blockSize > B200SP_BLOCK_MAX)) return fail(B200SP_EINVAL, "blockSize must be 1..131072");
    if (!d_sizes && stride < blockSize) return fail(B200SP_EINVAL, "stride smaller than blockSize");
    const uint32_t biggest = d_sizes ? B200SP_BLOCK_MAX : blockSize;
    if (seqStride < static_cast<uint64_t>(biggest) / 4 + 2) return fail(B200SP_EINVAL, "seqStride too small");
    return B200SP_OK;
}

static int launch_batch(b200sp_engine *e, const void *d_src, uint64_t totalSize, uint32_t blockSize,
                        uint64_t stride, const uint32_t *d_sizes, uint32_t nBlocks, int level,
                        b200sp_sequence *d_seqs, uint64_t seqStride, uint32_t *d_counts, cudaStream_t st,
                        unsigned int *workCounter, int scratch)
{
    if (!e) return fail(B200SP_EINVAL, "null engine");
    b200sp::ParseParams p;
    memset(&p, 0, sizeof p);
    if (!b200sp::params_for_level(level, p)) return fail(B200SP_EINVAL, "compression level outside 1..12");
    int rc = check_batch(d_src, blockSize, stride, d_sizes, nBlocks, seqStride);
    if (rc) return rc;
    if (nBlocks == 0) return B200SP_OK;
    if (!d_seqs || !d_counts) return fail(B200SP_EINVAL, "null output");
    DeviceGuard guard;
    CU_TRY(guard.enter(e->device), "cudaSetDevice");
    {
        const uint32_t grid = nBlocks < static_cast<uint32_t>(e->numSMs) ? nBlocks : static_cast<uint32_t>(e->numSMs);
        const size_t want = static_cast<size_t>(grid) * b200sp::kSortedCap;
        if (want > e->d_sortedCap[scratch]) {       // cudaFree waits for launches still using the old area
            cudaFree(e->d_sorted[scratch]);
            e->d_sorted[scratch] = nullptr; e->d_sortedCap[scratch] = 0;
            CU_TRY(cudaMalloc(&e->d_sorted[scratch], want * sizeof(uint32_t)), "cudaMalloc(sorted-table scratch)");
            e->d_sortedCap[scratch] = want;
        }
        p.sorted = e->d_sorted[scratch];
    }
    p.src = static_cast<const uint8_t *>(d_src);
    p.stride = stride;
    p.totalSize = totalSize;
    p.blockSize = blockSize;
    p.sizes = d_sizes;
    p.nBlocks = nBlocks;
    p.seqs = reinterpret_cast<uint4 *>(d_seqs);
    p.seqStride = seqStride;
    p.counts = d_counts;
    p.workCounter = workCounter;
    p.errorFlag = e->d_work + 48;                  // one word of the 256-byte counter block, checked by the host paths
    // developer profiling: B200SP_ROLE_PROFILE=1 accumulates per-role busy cycles in d_work[8..]
    static const bool roleProfile = getenv("B200SP_ROLE_PROFILE") != nullptr;
    p.roleCycles = roleProfile ? reinterpret_cast<unsigned long long *>(e->d_work) + 1 : nullptr;
    CU_TRY(cudaMemsetAsync(workCounter, 0, sizeof(unsigned int), st), "cudaMemsetAsync(work counter)");
    CU_TRY(b200sp::launch_parse(p, e->numSMs, st), "launch lz77_parse_kernel");
    return B200SP_OK;
}

int b200sp_parse_device(b200sp_engine *e, const void *d_src, uint64_t totalSize, uint32_t blockSize,
                        uint64_t stride, const uint32_t *d_sizes, uint32_t nBlocks, int level,
                        b200sp_sequence *d_seqs, uint64_t seqStride, uint32_t *d_counts, void *cudaStream)
{
    if (!e) return fail(B200SP_EINVAL, "null engine");
    cudaStream_t st = cudaStream ? static_cast<cudaStream_t>(cudaStream) : e->stream;
    return launch_batch(e, d_src, totalSize, blockSize, stride, d_sizes, nBlocks, level, d_seqs, seqStride, d_counts, st,
                        e->d_work, 0);
}

int b200sp_verify_device(b200sp_engine *e, const void *d_src, uint64_t totalSize, uint32_t blockSize,
                         uint64_t stride, const uint32_t *d_sizes, uint32_t nBlocks,
                         const b200sp_sequence *d_seqs, uint64_t seqStride, const uint32_t *d_counts,
                         uint32_t *d_bad, void *cudaStream)
{
    if (!e) return fail(B200SP_EINVAL, "null engine");
    if (nBlocks == 0) return B200SP_OK;
    if (!d_src || !d_seqs || !d_counts || !d_bad) return fail(B200SP_EINVAL, "null argument");
    cudaStream_t st = cudaStream ? static_cast<cudaStream_t>(cudaStream) : e->stream;
    DeviceGuard guard;
    CU_TRY(guard.enter(e->device), "cudaSetDevice");
    const unsigned threads = 128, warpsPerCta = threads / 32;
    unsigned grid = (nBlocks + warpsPerCta - 1) / warpsPerCta;
    if (grid > static_cast<unsigned>(e->numSMs) * 16u) grid = e->numSMs * 16u;
    verify_kernel<<<grid, threads, 0, st>>>(static_cast<const uint8_t *>(d_src), stride, totalSize, blockSize,
                                           d_sizes, nBlocks, reinterpret_cast<const uint4 *>(d_seqs), seqStride,
                                           d_counts, d_bad);
    CU_TRY(cudaGetLastError(), "launch verify_kernel");
    return B200SP_OK;
}

/* developer profiling (not in the public header): copies the 8 role counters (busy cycles of EH, TL, TS, P1,
 * P2; block wall cycles; stage count; spare) and zeroes them */
int b200sp_debug_role_cycles(b200sp_engine *e, unsigned long long *out6)
{
    if (!e || !out6) return B200SP_EINVAL;
    cudaStreamSynchronize(e->stream);
    cudaMemcpy(out6, reinterpret_cast<unsigned long long *>(e->d_work) + 1, 10 * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
    cudaMemset(reinterpret_cast<unsigned long long *>(e->d_work) + 1, 0, 10 * sizeof(unsigned long long));
    return B200SP_OK;
}

int b200sp_sync(b200sp_engine *e)
{
    if (!e) return fail(B200SP_EINVAL, "null engine");
    DeviceGuard guard;
    CU_TRY(guard.enter(e->device), "cudaSetDevice");
    cudaError_t wce = cudaSuccess;
    const int w = wait_stream(e->stream, &wce);
    if (w == 1) return fail(B200SP_ETIMEOUT, "no completion within 2 s");
    if (w) return fail(B200SP_ECUDA, "cudaStreamQuery", wce);
    return B200SP_OK;
}

// Drains the engine's streams after a failure part-way through a pipelined call, so that nothing is still writing
// into (or reading from) the engine's buffers when the caller sees the error.
static void quiesce(b200sp_engine *e)
{
    cudaStreamSynchronize(e->sIn);
    cudaStreamSynchronize(e->sParse[0]);
    cudaStreamSynchronize(e->sParse[1]);
    cudaStreamSynchronize(e->sPost);
    cudaStreamSynchronize(e->sOut);
    (void)cudaGetLastError();
}

#define CU_TRY_Q(expr, what)                                                         \
    do {                                                                             \
        cudaError_t ce_ = (expr);                                                    \
        if (ce_ != cudaSuccess) { quiesce(e); return fail(B200SP_ECUDA, what, ce_); } \
    } while (0)

// The pipelined host path.  out16 == nullptr: results come back in the 8-byte wire format (res);
// out16 != nullptr: dense ZSTD_Sequence[] straight into the caller's array (direct DMA when it is pinned,
// through pinned staging and a threaded copy otherwise), *nOut = entries written.
static int host_pipeline(b200sp_engine *e, const void *h_src, size_t srcSize, uint32_t blockSize, int level,
                         b200sp_result *res, b200sp_sequence *out16, size_t outCap, size_t *nOut)
{
    if (blockSize == 0 || blockSize > B200SP_BLOCK_MAX) return fail(B200SP_EINVAL, "blockSize must be 1..131072");
    if (level < 1 || level > 12) return fail(B200SP_EINVAL, "compression level outside 1..12");
    if (srcSize == 0) return B200SP_OK;
    if (!h_src) return fail(B200SP_EINVAL, "null source");
    DeviceGuard guard;
    CU_TRY(guard.enter(e->device), "cudaSetDevice");
    const bool dense = out16 != nullptr;
    const size_t entryBytes = dense ? 16 : 8;

    // blocks are laid out on the device at a 16-byte-aligned stride
    const uint64_t stride = (static_cast<uint64_t>(blockSize) + 15u) & ~15ull;
    const size_t nBlocks = (srcSize + blockSize - 1) / blockSize;
    if (nBlocks > 0x7FFFFFFFu) return fail(B200SP_EINVAL, "too many blocks");
    const size_t seqStride = (static_cast<size_t>(blockSize) / 4 + 2 + 7) & ~static_cast<size_t>(7);
    const size_t perBlockWorst = static_cast<size_t>(blockSize) / 4 + 2;   // entries
    const size_t devBytes = nBlocks * stride + 16;

    // ---- chunk plan: the copies of chunk k+1 / k-1 overlap the kernels of chunk k, and the parser launches
    // of consecutive chunks sit on two streams, so the CTAs of chunk k+1 move onto the SMs as those of chunk
    // k retire (no idle tail per chunk).  The first chunk is a quarter wave (the first SMs start after a 5 MB
    // copy), the second fills the wave, later chunks are one wave each (more when there are many blocks).
    const size_t sms = static_cast<size_t>(e->numSMs);
    const size_t first = sms / 4 ? sms / 4 : 1;
    size_t big = sms;
    if (nBlocks > sms + big * (kMaxChunks - 2)) big = ((nBlocks - sms + kMaxChunks - 3) / (kMaxChunks - 2) + sms - 1) / sms * sms;
    size_t chunkStart[kMaxChunks + 1];
    int nChunks = 0;
    for (size_t b = 0; b < nBlocks;) {
        chunkStart[nChunks++] = b;
        b += nChunks == 1 ? first : nChunks == 2 ? sms - first : big;
    }
    // ... and the last one a quarter wave again: what cannot overlap anything is the fetch of the last chunk's entries
    if (nChunks >= 2 && nChunks < kMaxChunks && nBlocks - chunkStart[nChunks - 1] > 2 * first) {
        chunkStart[nChunks] = nBlocks - first;
        nChunks++;
    }
    chunkStart[nChunks] = nBlocks;

    {
        CU_TRY(grow_dev(e->d_src, e->d_srcCap, devBytes), "cudaMalloc(src)");
        CU_TRY(grow_dev(e->d_seqs, e->d_seqsCap, nBlocks * seqStride), "cudaMalloc(seqs)");
        CU_TRY(grow_dev(e->d_counts, e->d_countsCap, nBlocks), "cudaMalloc(counts)");
        CU_TRY(grow_dev(e->d_offsets, e->d_offsetsCap, nBlocks + kMaxChunks), "cudaMalloc(offsets)");
        CU_TRY(grow_dev(e->d_packed, e->d_packedCap, nBlocks * perBlockWorst * (dense ? 2 : 1)), "cudaMalloc(packed)");
        CU_TRY(grow_host(e->h_counts, e->h_countsCap, nBlocks), "cudaMallocHost(counts)");
        CU_TRY(grow_host(e->h_offsets, e->h_offsetsCap, nBlocks + kMaxChunks), "cudaMallocHost(offsets)");
        if (e->h_goffsetsCap < nBlocks + 1) {
            free(e->h_goffsets);
            e->h_goffsets = static_cast<unsigned long long *>(malloc((nBlocks + 1 + nBlocks / 4) * sizeof(unsigned long long)));
            if (!e->h_goffsets) { e->h_goffsetsCap = 0; return fail(B200SP_ENOMEM, "out of host memory"); }
            e->h_goffsetsCap = nBlocks + 1 + nBlocks / 4;
        }
        if (e->verify) {
            CU_TRY(grow_dev(e->d_bad, e->d_badCap, nBlocks), "cudaMalloc(verify)");
            if (!e->d_badTotal) CU_TRY(cudaMalloc(&e->d_badTotal, sizeof(uint32_t)), "cudaMalloc(verify total)");
            if (!e->h_badTotal) CU_TRY(cudaMallocHost(&e->h_badTotal, sizeof(uint32_t)), "cudaMallocHost(verify total)");
            CU_TRY(cudaMemsetAsync(e->d_badTotal, 0, sizeof(uint32_t), e->sPost), "memset(verify total)");
        }
    }

    cudaPointerAttributes attr;
    const bool pinned = cudaPointerGetAttributes(&attr, h_src) == cudaSuccess &&
                        (attr.type == cudaMemoryTypeHost || attr.type == cudaMemoryTypeManaged);
    (void)cudaGetLastError();
    bool outPinned = false;
    if (dense) {
        outPinned = cudaPointerGetAttributes(&attr, out16) == cudaSuccess &&
                    (attr.type == cudaMemoryTypeHost || attr.type == cudaMemoryTypeManaged);
        (void)cudaGetLastError();
    }
    // staging for results that cannot be copied straight to their destination: the wire format always lands in the
    // engine's pinned array; dense output only when the caller's array is pageable.
    // typical text yields one entry per 10-20 input bytes; grown on demand below
    if (!dense || !outPinned)
        CU_TRY(grow_host(e->h_packed, e->h_packedCap, (srcSize / 12 + 4 * nBlocks + 1024) * (dense ? 2 : 1)), "cudaMallocHost(packed)");
    const uint8_t *hs = static_cast<const uint8_t *>(h_src);
    if (!pinned) CU_TRY(grow_host(e->h_stage, e->h_stageCap, srcSize), "cudaMallocHost(stage)");

    // ---- enqueue: per chunk H2D on the copy-in stream, then parse (+ verify) + scan + pack + small D2H behind events
    for (int k = 0; k < nChunks; k++) {
        const size_t b0 = chunkStart[k], b1 = chunkStart[k + 1], nb = b1 - b0;
        const size_t byte0 = b0 * blockSize, byte1 = b1 * blockSize < srcSize ? b1 * blockSize : srcSize;
        const uint8_t *hsrc = hs + byte0;
        if (!pinned) {
            // USDM-style staging (/root/reference/src/qatseqprod.c:1222-1224): memcpy into pinned memory
            memcpy(e->h_stage + byte0, hs + byte0, byte1 - byte0);
            hsrc = e->h_stage + byte0;
        }
        uint8_t *dsrc = e->d_src + b0 * stride;
        if (stride == blockSize) {
            CU_TRY_Q(cudaMemcpyAsync(dsrc, hsrc, byte1 - byte0, cudaMemcpyHostToDevice, e->sIn), "H2D");
        } else {
            const size_t full = (byte1 - byte0) / blockSize, rem = (byte1 - byte0) % blockSize;
            if (full) CU_TRY_Q(cudaMemcpy2DAsync(dsrc, stride, hsrc, blockSize, blockSize, full, cudaMemcpyHostToDevice, e->sIn), "H2D 2D");
            if (rem) CU_TRY_Q(cudaMemcpyAsync(dsrc + full * stride, hsrc + full * blockSize, rem, cudaMemcpyHostToDevice, e->sIn), "H2D tail");
        }
        CU_TRY_Q(cudaEventRecord(e->evIn[k], e->sIn), "event record");
        cudaStream_t sp = e->sParse[k & 1];
        CU_TRY_Q(cudaStreamWaitEvent(sp, e->evIn[k], 0), "stream wait");

        // sizes in the strided layout: the last block of the chunk holds what is left of the input
        const uint64_t lastBytes = (byte1 - byte0) - (nb - 1) * static_cast<size_t>(blockSize);
        const uint64_t totalStrided = (nb - 1) * stride + lastBytes;
        uint4 *dseqs = e->d_seqs + b0 * seqStride;
        uint32_t *dcounts = e->d_counts + b0;
        unsigned long long *doffs = e->d_offsets + b0 + k;
        int rc = launch_batch(e, dsrc, totalStrided, blockSize, stride, nullptr, static_cast<uint32_t>(nb), level,
                              reinterpret_cast<b200sp_sequence *>(dseqs), seqStride, dcounts, sp, e->d_chunkWork + k, 1 + (k & 1));
        if (rc) { quiesce(e); return rc; }
        CU_TRY_Q(cudaEventRecord(e->evParsed[k], sp), "event record");
        CU_TRY_Q(cudaStreamWaitEvent(e->sPost, e->evParsed[k], 0), "stream wait");
        if (e->verify) {        // compress-and-verify (/root/reference/src/qatseqprod.c:1238): replay before anything is returned
            const unsigned vthreads = 128, warpsPerCta = vthreads / 32;
            unsigned vgrid = static_cast<unsigned>((nb + warpsPerCta - 1) / warpsPerCta);
            if (vgrid > static_cast<unsigned>(e->numSMs) * 16u) vgrid = e->numSMs * 16u;
            verify_kernel<<<vgrid, vthreads, 0, e->sPost>>>(dsrc, stride, totalStrided, blockSize, nullptr, static_cast<uint32_t>(nb),
                                                           dseqs, seqStride, dcounts, e->d_bad + b0);
            count_bad_kernel<<<8, 128, 0, e->sPost>>>(e->d_bad + b0, static_cast<uint32_t>(nb), e->d_badTotal);
            CU_TRY_Q(cudaGetLastError(), "launch verify_kernel");
        }
        scan_counts_kernel<<<1, 32, 0, e->sPost>>>(dcounts, static_cast<uint32_t>(nb), doffs);
        CU_TRY_Q(cudaGetLastError(), "launch scan_counts_kernel");
        if (dense)
            dense_kernel<<<e->numSMs * 2, kPostThreads, 0, e->sPost>>>(dseqs, seqStride, dcounts, doffs, static_cast<uint32_t>(nb),
                                                                      reinterpret_cast<uint4 *>(e->d_packed) + b0 * perBlockWorst);
        else
            pack_kernel<<<e->numSMs * 2, kPostThreads, 0, e->sPost>>>(dseqs, seqStride, dcounts, doffs, static_cast<uint32_t>(nb),
                                                                     e->d_packed + b0 * perBlockWorst);
        CU_TRY_Q(cudaGetLastError(), "launch pack_kernel");
        CU_TRY_Q(cudaMemcpyAsync(e->h_counts + b0, dcounts, nb * sizeof(uint32_t), cudaMemcpyDeviceToHost, e->sPost), "D2H counts");
        CU_TRY_Q(cudaMemcpyAsync(e->h_offsets + b0 + k, doffs, (nb + 1) * sizeof(unsigned long long), cudaMemcpyDeviceToHost, e->sPost), "D2H offsets");
        CU_TRY_Q(cudaEventRecord(e->evDone[k], e->sPost), "event record");
    }
    if (e->verify) CU_TRY_Q(cudaMemcpyAsync(e->h_badTotal, e->d_badTotal, sizeof(uint32_t), cudaMemcpyDeviceToHost, e->sPost), "D2H verify total");
    CU_TRY_Q(cudaMemcpyAsync(e->h_flag, e->d_work + 48, sizeof(uint32_t), cudaMemcpyDeviceToHost, e->sPost), "D2H error flag");

    // ---- drain: as each chunk completes, fetch exactly its entries on the copy-out stream
    size_t hostPos = 0;
    for (int k = 0; k < nChunks; k++) {
        const size_t b0 = chunkStart[k], nb = chunkStart[k + 1] - b0;
        cudaError_t wce = cudaSuccess;
        const int w = wait_event(e->evDone[k], &wce);
        if (w == 1) { return fail(B200SP_ETIMEOUT, "no completion within 2 s"); }      // like the reference: give up, report an error
        if (w) { quiesce(e); return fail(B200SP_ECUDA, "wait chunk", wce); }
        const unsigned long long *lo = e->h_offsets + b0 + k;
        const size_t total = static_cast<size_t>(lo[nb]);
        uint8_t *dst;
        if (dense && outPinned) {
            if (hostPos + total > outCap) { quiesce(e); return fail(B200SP_EINVAL, "sequence array too small"); }
            dst = reinterpret_cast<uint8_t *>(out16 + hostPos);
        } else {
            if ((hostPos + total) * entryBytes > e->h_packedCap * 8) {
                // rare: denser than one entry per 12 bytes. Finish the copies in flight, move to a bigger buffer.
                CU_TRY_Q(cudaStreamSynchronize(e->sOut), "sync before grow");
                unsigned long long *bigger = nullptr;
                const size_t cap = ((hostPos + total) * 2 + 1024) * (dense ? 2 : 1);
                CU_TRY_Q(cudaMallocHost(&bigger, cap * sizeof(unsigned long long)), "cudaMallocHost(packed grow)");
                memcpy(bigger, e->h_packed, hostPos * entryBytes);
                cudaFreeHost(e->h_packed);
                e->h_packed = bigger;
                e->h_packedCap = cap;
            }
            dst = reinterpret_cast<uint8_t *>(e->h_packed) + hostPos * entryBytes;
        }
        const uint8_t *dsrcEntries = reinterpret_cast<const uint8_t *>(e->d_packed) + b0 * perBlockWorst * entryBytes;
        CU_TRY_Q(cudaMemcpyAsync(dst, dsrcEntries, total * entryBytes, cudaMemcpyDeviceToHost, e->sOut), "D2H entries");
        for (size_t i = 0; i < nb; i++) e->h_goffsets[b0 + i] = hostPos + lo[i];
        hostPos += total;
    }
    e->h_goffsets[nBlocks] = hostPos;
    {
        cudaError_t wce = cudaSuccess;
        int w = wait_stream(e->sOut, &wce);
        if (!w) w = wait_stream(e->sPost, &wce);
        if (w == 1) return fail(B200SP_ETIMEOUT, "no completion within 2 s");
        if (w) { quiesce(e); return fail(B200SP_ECUDA, "sync after D2H", wce); }
    }
    if (*e->h_flag != 0) return fail(B200SP_ECUDA, "the parser gave up waiting for an internal hand-off");
    if (e->verify && *e->h_badTotal != 0) return fail(B200SP_EVERIFY, "on-device verification rejected a block");
    if (dense && !outPinned) {
        if (hostPos > outCap) return fail(B200SP_EINVAL, "sequence array too small");
        parallel_copy(out16, e->h_packed, hostPos * 16);
    }
    if (nOut) *nOut = hostPos;

    res->nBlocks = static_cast<uint32_t>(nBlocks);
    res->counts = e->h_counts;
    res->offsets = reinterpret_cast<const uint64_t *>(e->h_goffsets);
    res->packed = dense ? nullptr : reinterpret_cast<const uint64_t *>(e->h_packed);
    return B200SP_OK;
}

int b200sp_parse_host(b200sp_engine *e, const void *h_src, size_t srcSize, uint32_t blockSize, int level,
                      b200sp_result *res)
{
    if (!e || !res) return fail(B200SP_EINVAL, "null engine/result");
    memset(res, 0, sizeof *res);
    return host_pipeline(e, h_src, srcSize, blockSize, level, res, nullptr, 0, nullptr);
}

int b200sp_sequences_host(b200sp_engine *e, const void *h_src, size_t srcSize, uint32_t blockSize, int level,
                          b200sp_sequence *h_out, size_t outCapacity, size_t *nSeqs, b200sp_result *res)
{
    b200sp_result local;
    if (!e || !h_out || !nSeqs) return fail(B200SP_EINVAL, "null engine/output");
    *nSeqs = 0;
    if (!res) res = &local;
    memset(res, 0, sizeof *res);
    return host_pipeline(e, h_src, srcSize, blockSize, level, res, h_out, outCapacity, nSeqs);
}

int b200sp_engine_set_verify(b200sp_engine *e, int enable)
{
    if (!e) return fail(B200SP_EINVAL, "null engine");
    const int before = e->verify;
    e->verify = enable ? 1 : 0;
    return before;
}

// Staging layout of the scattered-block path: [sizes (u32 per slot, padded to 16 B)] [slot 0] [slot 1] ... at a
// 128 KiB stride, in pinned host memory and mirrored on the device.
static size_t staged_sizes_bytes(uint32_t nSlots)
{
    return (static_cast<size_t>(nSlots) * sizeof(uint32_t) + 15u) & ~static_cast<size_t>(15);
}

// Buffers of the staged path, sized for the whole reservation at once: growing in steps means cudaFree, which waits
// for every stream of the device - with many engines at work (one or two per compressing thread) each step stalls
// all of them.  The parser's table scratch is included for the same reason.
static int reserve_staged_buffers(b200sp_engine *e)
{
    const uint64_t stride = B200SP_BLOCK_MAX;
    const size_t seqStride = B200SP_SEQ_STRIDE, perBlockWorst = B200SP_BLOCK_MAX / 4 + 2;
    const size_t sizesBytes = staged_sizes_bytes(e->stageSlots);
    const size_t capBlocks = e->stageSlots;
    CU_TRY(grow_dev(e->d_src, e->d_srcCap, sizesBytes + capBlocks * stride + 16), "cudaMalloc(src)");
    CU_TRY(grow_dev(e->d_seqs, e->d_seqsCap, capBlocks * seqStride), "cudaMalloc(seqs)");
    CU_TRY(grow_dev(e->d_counts, e->d_countsCap, capBlocks), "cudaMalloc(counts)");
    CU_TRY(grow_dev(e->d_offsets, e->d_offsetsCap, capBlocks + kMaxChunks), "cudaMalloc(offsets)");
    CU_TRY(grow_dev(e->d_packed, e->d_packedCap, capBlocks * perBlockWorst), "cudaMalloc(packed)");
    CU_TRY(grow_host(e->h_counts, e->h_countsCap, capBlocks), "cudaMallocHost(counts)");
    CU_TRY(grow_host(e->h_offsets, e->h_offsetsCap, capBlocks + kMaxChunks), "cudaMallocHost(offsets)");
    CU_TRY(grow_host(e->h_packed, e->h_packedCap, capBlocks * (B200SP_BLOCK_MAX / 12) + 1024), "cudaMallocHost(packed)");
    {
        const size_t grid = capBlocks < static_cast<size_t>(e->numSMs) ? capBlocks : static_cast<size_t>(e->numSMs);
        const size_t want = grid * b200sp::kSortedCap;
        if (want > e->d_sortedCap[0]) {
            cudaFree(e->d_sorted[0]);
            e->d_sorted[0] = nullptr; e->d_sortedCap[0] = 0;
            CU_TRY(cudaMalloc(&e->d_sorted[0], want * sizeof(uint32_t)), "cudaMalloc(sorted-table scratch)");
            e->d_sortedCap[0] = want;
        }
    }
    return B200SP_OK;
}

int b200sp_stage_reserve(b200sp_engine *e, uint32_t nSlots, void **slots)
{
    if (!e || !slots || nSlots == 0) return fail(B200SP_EINVAL, "stage_reserve: bad argument");
    DeviceGuard guard;
    CU_TRY(guard.enter(e->device), "cudaSetDevice");
    // the slot area starts at the offset the LARGEST reservation needs, so it never moves when fewer slots are used
    if (nSlots > e->stageSlots) {
        CU_TRY(cudaStreamSynchronize(e->stream), "sync before growing the staging area");
        CU_TRY(grow_host(e->h_slots, e->h_slotsCap, staged_sizes_bytes(nSlots) + static_cast<size_t>(nSlots) * B200SP_BLOCK_MAX),
               "cudaMallocHost(stage)");
        e->stageSlots = nSlots;
    }
    *slots = e->h_slots + staged_sizes_bytes(e->stageSlots);
    return reserve_staged_buffers(e);       // everything the staged parse of that many blocks needs, now rather than inside it
}

int b200sp_parse_staged(b200sp_engine *e, const uint32_t *sizes, uint32_t nBlocks, int level, b200sp_result *res)
{
    if (!e || !res) return fail(B200SP_EINVAL, "null engine/result");
    memset(res, 0, sizeof *res);
    if (level < 1 || level > 12) return fail(B200SP_EINVAL, "compression level outside 1..12");
    if (nBlocks == 0) return B200SP_OK;
    if (!sizes || nBlocks > e->stageSlots) return fail(B200SP_EINVAL, "parse_staged: more blocks than reserved slots");
    bool allFull = true;
    for (uint32_t b = 0; b < nBlocks; b++) {
        if (sizes[b] == 0 || sizes[b] > B200SP_BLOCK_MAX) return fail(B200SP_EINVAL, "block size must be 1..131072");
        allFull = allFull && sizes[b] == B200SP_BLOCK_MAX;
    }
    DeviceGuard guard;
    CU_TRY(guard.enter(e->device), "cudaSetDevice");

    const uint64_t stride = B200SP_BLOCK_MAX;
    const size_t seqStride = B200SP_SEQ_STRIDE;
    const size_t sizesBytes = staged_sizes_bytes(e->stageSlots);
    { const int rcAlloc = reserve_staged_buffers(e); if (rcAlloc) return rcAlloc; }
    if (e->h_goffsetsCap < nBlocks + 1) {
        free(e->h_goffsets);
        e->h_goffsets = static_cast<unsigned long long *>(malloc((nBlocks + 1 + nBlocks / 4) * sizeof(unsigned long long)));
        if (!e->h_goffsets) { e->h_goffsetsCap = 0; return fail(B200SP_ENOMEM, "out of host memory"); }
        e->h_goffsetsCap = nBlocks + 1 + nBlocks / 4;
    }

    uint32_t *hSizes = reinterpret_cast<uint32_t *>(e->h_slots);
    uint8_t *hBlocks = e->h_slots + sizesBytes;
    uint8_t *dSizes = e->d_src, *dBlocks = e->d_src + sizesBytes;
    cudaStream_t st = e->stream;
    memcpy(hSizes, sizes, nBlocks * sizeof(uint32_t));
    if (allFull) {          // sizes and blocks are contiguous: one copy
        CU_TRY(cudaMemcpyAsync(dSizes, hSizes, sizesBytes + static_cast<size_t>(nBlocks) * stride, cudaMemcpyHostToDevice, st), "H2D batch");
    } else {                // only the bytes of each block travel
        CU_TRY(cudaMemcpyAsync(dSizes, hSizes, nBlocks * sizeof(uint32_t), cudaMemcpyHostToDevice, st), "H2D sizes");
        for (uint32_t b = 0; b < nBlocks; b++)
            CU_TRY(cudaMemcpyAsync(dBlocks + b * stride, hBlocks + b * stride, sizes[b], cudaMemcpyHostToDevice, st), "H2D block");
    }
    int rc = launch_batch(e, dBlocks, static_cast<uint64_t>(nBlocks) * stride, B200SP_BLOCK_MAX, stride,
                          reinterpret_cast<const uint32_t *>(dSizes), nBlocks, level,
                          reinterpret_cast<b200sp_sequence *>(e->d_seqs), seqStride, e->d_counts, st, e->d_work, 0);
    if (rc) return rc;
    if (e->verify) {
        CU_TRY(grow_dev(e->d_bad, e->d_badCap, nBlocks), "cudaMalloc(verify)");
        if (!e->d_badTotal) CU_TRY(cudaMalloc(&e->d_badTotal, sizeof(uint32_t)), "cudaMalloc(verify total)");
        if (!e->h_badTotal) CU_TRY(cudaMallocHost(&e->h_badTotal, sizeof(uint32_t)), "cudaMallocHost(verify total)");
        CU_TRY(cudaMemsetAsync(e->d_badTotal, 0, sizeof(uint32_t), st), "memset(verify total)");
        unsigned vgrid = (nBlocks + 3) / 4;
        if (vgrid > static_cast<unsigned>(e->numSMs) * 16u) vgrid = e->numSMs * 16u;
        verify_kernel<<<vgrid, 128, 0, st>>>(dBlocks, stride, static_cast<uint64_t>(nBlocks) * stride, B200SP_BLOCK_MAX,
                                            reinterpret_cast<const uint32_t *>(dSizes), nBlocks, e->d_seqs, seqStride, e->d_counts, e->d_bad);
        count_bad_kernel<<<8, 128, 0, st>>>(e->d_bad, nBlocks, e->d_badTotal);
        CU_TRY(cudaGetLastError(), "launch verify_kernel");
        CU_TRY(cudaMemcpyAsync(e->h_badTotal, e->d_badTotal, sizeof(uint32_t), cudaMemcpyDeviceToHost, st), "D2H verify total");
    }
    scan_counts_kernel<<<1, 32, 0, st>>>(e->d_counts, nBlocks, e->d_offsets);
    CU_TRY(cudaGetLastError(), "launch scan_counts_kernel");
    pack_kernel<<<e->numSMs * 2, kPostThreads, 0, st>>>(e->d_seqs, seqStride, e->d_counts, e->d_offsets, nBlocks, e->d_packed);
    CU_TRY(cudaGetLastError(), "launch pack_kernel");
    CU_TRY(cudaMemcpyAsync(e->h_counts, e->d_counts, nBlocks * sizeof(uint32_t), cudaMemcpyDeviceToHost, st), "D2H counts");
    CU_TRY(cudaMemcpyAsync(e->h_offsets, e->d_offsets, (nBlocks + 1) * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st), "D2H offsets");
    {
        cudaError_t wce = cudaSuccess;
        const int w = wait_stream(st, &wce);
        if (w == 1) return fail(B200SP_ETIMEOUT, "no completion within 2 s");
        if (w) return fail(B200SP_ECUDA, "sync after parse", wce);
    }
    if (e->verify && *e->h_badTotal != 0) return fail(B200SP_EVERIFY, "on-device verification rejected a block");
    const size_t total = static_cast<size_t>(e->h_offsets[nBlocks]);
    CU_TRY(grow_host(e->h_packed, e->h_packedCap, total + 1024), "cudaMallocHost(packed)");
    CU_TRY(cudaMemcpyAsync(e->h_packed, e->d_packed, total * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st), "D2H packed");
    for (uint32_t b = 0; b <= nBlocks; b++) e->h_goffsets[b] = e->h_offsets[b];
    {
        cudaError_t wce = cudaSuccess;
        const int w = wait_stream(st, &wce);
        if (w == 1) return fail(B200SP_ETIMEOUT, "no completion within 2 s");
        if (w) return fail(B200SP_ECUDA, "sync after D2H", wce);
    }

    res->nBlocks = nBlocks;
    res->counts = e->h_counts;
    res->offsets = reinterpret_cast<const uint64_t *>(e->h_goffsets);
    res->packed = reinterpret_cast<const uint64_t *>(e->h_packed);
    return B200SP_OK;
}

int b200sp_parse_blocks(b200sp_engine *e, const void *const *h_blocks, const uint32_t *sizes, uint32_t nBlocks,
                        int level, b200sp_result *res)
{
    if (!e || !res) return fail(B200SP_EINVAL, "null engine/result");
    memset(res, 0, sizeof *res);
    if (nBlocks == 0) return B200SP_OK;
    if (!h_blocks || !sizes) return fail(B200SP_EINVAL, "null block list");
    for (uint32_t b = 0; b < nBlocks; b++)
        if (!h_blocks[b] || sizes[b] == 0 || sizes[b] > B200SP_BLOCK_MAX) return fail(B200SP_EINVAL, "block size must be 1..131072");
    void *slots = nullptr;
    int rc = b200sp_stage_reserve(e, nBlocks, &slots);
    if (rc) return rc;
    for (uint32_t b = 0; b < nBlocks; b++)      // gather (callers that can copy in parallel use the two calls above directly)
        memcpy(static_cast<uint8_t *>(slots) + static_cast<size_t>(b) * B200SP_BLOCK_MAX, h_blocks[b], sizes[b]);
    return b200sp_parse_staged(e, sizes, nBlocks, level, res);
}

int b200sp_host_register(void *ptr, size_t bytes)
{
    if (!ptr || bytes == 0) return fail(B200SP_EINVAL, "host_register: bad argument");
    CU_TRY(cudaHostRegister(ptr, bytes, cudaHostRegisterPortable), "cudaHostRegister");
    return B200SP_OK;
}

int b200sp_host_unregister(void *ptr)
{
    if (!ptr) return fail(B200SP_EINVAL, "host_unregister: bad argument");
    CU_TRY(cudaHostUnregister(ptr), "cudaHostUnregister");
    return B200SP_OK;
}

void b200sp_expand(const uint64_t *packed, size_t count, b200sp_sequence *out)
{
    for (size_t i = 0; i < count; i++) {
        const uint64_t v = packed[i];
        out[i].offset = static_cast<uint32_t>(v & 0x1FFFFu);
        out[i].litLength = static_cast<uint32_t>((v >> 17) & 0x3FFFFu);
        out[i].matchLength = static_cast<uint32_t>((v >> 35) & 0x3FFFFu);
        out[i].rep = 0;
    }
}

}  // extern "C"
