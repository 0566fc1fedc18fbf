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

bool device_usable(int dev)
{
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, dev) != cudaSuccess) return false;
    // the cubin is sm_100a only; the parser needs the 227 KB opt-in shared memory carve-out
    return prop.major == 10 && prop.minor == 0 &&
           prop.sharedMemPerBlockOptin >= static_cast<size_t>(b200sp::kSmemTotal);
}

// ---- wire format: offset | litLength << 17 | matchLength << 35 ----------------------------
__global__ void scan_counts_kernel(const uint32_t *__restrict__ counts, uint32_t nBlocks,
                                   unsigned long long *__restrict__ offsets)
{
    // single CTA: chunked exclusive scan, enough for a few hundred thousand blocks
    __shared__ unsigned long long warpSums[32];
    __shared__ unsigned long long running;
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) running = 0;
    __syncthreads();
    for (uint32_t base = 0; base < nBlocks; base += blockDim.x) {
        const uint32_t i = base + tid;
        unsigned long long v = i < nBlocks ? counts[i] : 0, incl = v;
        for (int d = 1; d < 32; d <<= 1) {
            unsigned long long o = __shfl_up_sync(0xFFFFFFFFu, incl, d);
            if (lane >= (uint32_t)d) incl += o;
        }
        if (lane == 31) warpSums[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            unsigned long long w = lane < (blockDim.x >> 5) ? warpSums[lane] : 0, wi = w;
            for (int d = 1; d < 32; d <<= 1) {
                unsigned long long o = __shfl_up_sync(0xFFFFFFFFu, wi, d);
                if (lane >= (uint32_t)d) wi += o;
            }
            warpSums[lane] = wi - w;      // exclusive
        }
        __syncthreads();
        const unsigned long long excl = running + warpSums[warp] + incl - v;
        if (i < nBlocks) offsets[i] = excl;
        __syncthreads();
        if (tid == blockDim.x - 1) running = excl + v;
        __syncthreads();
    }
    if (tid == 0) offsets[nBlocks] = running;
}

__global__ void pack_kernel(const uint4 *__restrict__ seqs, uint64_t seqStride,
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
struct b200sp_engine {
    int device;
    int numSMs;
    cudaStream_t stream;
    unsigned int *d_work;        // dynamic scheduler counter
    // host-path scratch (grown on demand)
    uint8_t *d_src;      size_t d_srcCap;
    uint4 *d_seqs;       size_t d_seqsCap;      // entries
    uint32_t *d_counts;  size_t d_countsCap;
    unsigned long long *d_offsets;
    unsigned long long *d_packed; size_t d_packedCap;   // entries
    uint8_t *h_stage;    size_t h_stageCap;     // pinned staging for pageable inputs
    uint32_t *h_counts;  unsigned long long *h_offsets; size_t h_countsCap;
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

int b200sp_engine_create(int device, b200sp_engine **out)
{
    if (!out) return fail(B200SP_EINVAL, "engine_create: null out pointer");
    *out = nullptr;
    int n = 0;
    cudaError_t ce = cudaGetDeviceCount(&n);
    if (ce != cudaSuccess || n == 0) { (void)cudaGetLastError(); return fail(B200SP_ENODEVICE, "no CUDA device", ce); }
    if (device < 0 || device >= n) return fail(B200SP_EINVAL, "engine_create: device index out of range");
    if (!device_usable(device)) return fail(B200SP_EUNSUPPORTED, "device is not sm_100 with 227 KB shared memory per CTA");
    CU_TRY(cudaSetDevice(device), "cudaSetDevice");
    CU_TRY(b200sp::configure_kernels(), "cudaFuncSetAttribute(max dynamic smem)");
    b200sp_engine *e = static_cast<b200sp_engine *>(calloc(1, sizeof(b200sp_engine)));
    if (!e) return fail(B200SP_ENOMEM, "engine_create: out of host memory");
    e->device = device;
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, device);
    e->numSMs = prop.multiProcessorCount;
    ce = cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking);
    if (ce == cudaSuccess) ce = cudaMalloc(&e->d_work, 256);
    if (ce == cudaSuccess) ce = cudaMemset(e->d_work, 0, 256);
    if (ce != cudaSuccess) { b200sp_engine_destroy(e); return fail(B200SP_ECUDA, "engine_create", ce); }
    *out = e;
    return B200SP_OK;
}

void b200sp_engine_destroy(b200sp_engine *e)
{
    if (!e) return;
    cudaSetDevice(e->device);
    if (e->stream) { cudaStreamSynchronize(e->stream); cudaStreamDestroy(e->stream); }
    cudaFree(e->d_work); cudaFree(e->d_src); cudaFree(e->d_seqs); cudaFree(e->d_counts);
    cudaFree(e->d_offsets); cudaFree(e->d_packed);
    cudaFreeHost(e->h_stage); cudaFreeHost(e->h_counts); cudaFreeHost(e->h_offsets); cudaFreeHost(e->h_packed);
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

int b200sp_parse_device(b200sp_engine *e, const void *d_src, uint64_t totalSize, uint32_t blockSize,
                        uint64_t stride, const uint32_t *d_sizes, uint32_t nBlocks, int level,
                        b200sp_sequence *d_seqs, uint64_t seqStride, uint32_t *d_counts, void *cudaStream)
{
    if (!e) return fail(B200SP_EINVAL, "null engine");
    b200sp::ParseParams p;
    memset(&p, 0, sizeof p);
    if (!b200sp::params_for_level(level, p)) return fail(B200SP_EINVAL, "compression level outside 1..12");
    int rc = check_batch(d_src, blockSize, stride, d_sizes, nBlocks, seqStride);
    if (rc) return rc;
    if (nBlocks == 0) return B200SP_OK;
    if (!d_seqs || !d_counts) return fail(B200SP_EINVAL, "null output");
    cudaStream_t st = cudaStream ? static_cast<cudaStream_t>(cudaStream) : e->stream;
    CU_TRY(cudaSetDevice(e->device), "cudaSetDevice");
    p.src = static_cast<const uint8_t *>(d_src);
    p.stride = stride;
    p.totalSize = totalSize;
    p.blockSize = blockSize;
    p.sizes = d_sizes;
    p.nBlocks = nBlocks;
    p.seqs = reinterpret_cast<uint4 *>(d_seqs);
    p.seqStride = seqStride;
    p.counts = d_counts;
    p.workCounter = e->d_work;
    // developer profiling: B200SP_ROLE_PROFILE=1 accumulates per-role busy cycles in d_work[8..]
    static const bool roleProfile = getenv("B200SP_ROLE_PROFILE") != nullptr;
    p.roleCycles = roleProfile ? reinterpret_cast<unsigned long long *>(e->d_work) + 1 : nullptr;
    CU_TRY(cudaMemsetAsync(e->d_work, 0, sizeof(unsigned int), st), "cudaMemsetAsync(work counter)");
    CU_TRY(b200sp::launch_parse(p, e->numSMs, st), "launch lz77_parse_kernel");
    return B200SP_OK;
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
    CU_TRY(cudaSetDevice(e->device), "cudaSetDevice");
    const unsigned threads = 128, warpsPerCta = threads / 32;
    unsigned grid = (nBlocks + warpsPerCta - 1) / warpsPerCta;
    if (grid > static_cast<unsigned>(e->numSMs) * 16u) grid = e->numSMs * 16u;
    verify_kernel<<<grid, threads, 0, st>>>(static_cast<const uint8_t *>(d_src), stride, totalSize, blockSize,
                                           d_sizes, nBlocks, reinterpret_cast<const uint4 *>(d_seqs), seqStride,
                                           d_counts, d_bad);
    CU_TRY(cudaGetLastError(), "launch verify_kernel");
    return B200SP_OK;
}

/* developer profiling (not in the public header): copies the 6 role counters and zeroes them */
int b200sp_debug_role_cycles(b200sp_engine *e, unsigned long long *out6)
{
    if (!e || !out6) return B200SP_EINVAL;
    cudaStreamSynchronize(e->stream);
    cudaMemcpy(out6, reinterpret_cast<unsigned long long *>(e->d_work) + 1, 6 * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
    cudaMemset(reinterpret_cast<unsigned long long *>(e->d_work) + 1, 0, 6 * sizeof(unsigned long long));
    return B200SP_OK;
}

int b200sp_sync(b200sp_engine *e)
{
    if (!e) return fail(B200SP_EINVAL, "null engine");
    CU_TRY(cudaStreamSynchronize(e->stream), "cudaStreamSynchronize");
    return B200SP_OK;
}

int b200sp_parse_host(b200sp_engine *e, const void *h_src, size_t srcSize, uint32_t blockSize, int level,
                      b200sp_result *res)
{
    if (!e || !res) return fail(B200SP_EINVAL, "null engine/result");
    memset(res, 0, sizeof *res);
    if (blockSize == 0 || blockSize > B200SP_BLOCK_MAX) return fail(B200SP_EINVAL, "blockSize must be 1..131072");
    if (level < 1 || level > 12) return fail(B200SP_EINVAL, "compression level outside 1..12");
    if (srcSize == 0) return B200SP_OK;
    if (!h_src) return fail(B200SP_EINVAL, "null source");
    CU_TRY(cudaSetDevice(e->device), "cudaSetDevice");

    // blocks are laid out on the device at a 16-byte-aligned stride
    const uint64_t stride = (static_cast<uint64_t>(blockSize) + 15u) & ~15ull;
    const size_t nBlocks = (srcSize + blockSize - 1) / blockSize;
    if (nBlocks > 0x7FFFFFFFu) return fail(B200SP_EINVAL, "too many blocks");
    const size_t seqStride = (static_cast<size_t>(blockSize) / 4 + 2 + 7) & ~static_cast<size_t>(7);
    const size_t devBytes = nBlocks * stride + 16;

    {
        size_t cap = e->d_countsCap;
        CU_TRY(grow_dev(e->d_src, e->d_srcCap, devBytes), "cudaMalloc(src)");
        CU_TRY(grow_dev(e->d_seqs, e->d_seqsCap, nBlocks * seqStride), "cudaMalloc(seqs)");
        CU_TRY(grow_dev(e->d_counts, e->d_countsCap, nBlocks), "cudaMalloc(counts)");
        if (e->d_countsCap != cap) {
            cudaFree(e->d_offsets); e->d_offsets = nullptr;
            CU_TRY(cudaMalloc(&e->d_offsets, (e->d_countsCap + 1) * sizeof(unsigned long long)), "cudaMalloc(offsets)");
        }
        size_t hc = e->h_countsCap;
        CU_TRY(grow_host(e->h_counts, e->h_countsCap, nBlocks), "cudaMallocHost(counts)");
        if (e->h_countsCap != hc) {
            cudaFreeHost(e->h_offsets); e->h_offsets = nullptr;
            CU_TRY(cudaMallocHost(&e->h_offsets, (e->h_countsCap + 1) * sizeof(unsigned long long)), "cudaMallocHost(offsets)");
        }
    }

    // ---- H2D.  Contiguous when the stride equals the block size (the common 128 KiB case).
    cudaPointerAttributes attr;
    bool pinned = cudaPointerGetAttributes(&attr, h_src) == cudaSuccess &&
                  (attr.type == cudaMemoryTypeHost || attr.type == cudaMemoryTypeManaged);
    (void)cudaGetLastError();
    const uint8_t *hs = static_cast<const uint8_t *>(h_src);
    if (!pinned) {
        // USDM-style staging (/root/reference/src/qatseqprod.c:1222-1224): memcpy into pinned memory
        CU_TRY(grow_host(e->h_stage, e->h_stageCap, srcSize), "cudaMallocHost(stage)");
        memcpy(e->h_stage, h_src, srcSize);
        hs = e->h_stage;
    }
    if (stride == blockSize) {
        CU_TRY(cudaMemcpyAsync(e->d_src, hs, srcSize, cudaMemcpyHostToDevice, e->stream), "H2D");
    } else {
        CU_TRY(cudaMemcpy2DAsync(e->d_src, stride, hs, blockSize, blockSize, srcSize / blockSize,
                                 cudaMemcpyHostToDevice, e->stream), "H2D 2D");
        const size_t rem = srcSize % blockSize;
        if (rem) CU_TRY(cudaMemcpyAsync(e->d_src + (srcSize / blockSize) * stride, hs + (srcSize / blockSize) * blockSize,
                                        rem, cudaMemcpyHostToDevice, e->stream), "H2D tail");
    }

    // ---- parse + pack
    // totalSize is expressed in the strided layout: the last block holds what is left of srcSize
    const uint64_t lastBytes = srcSize - (nBlocks - 1) * static_cast<size_t>(blockSize);
    const uint64_t totalStrided = (nBlocks - 1) * stride + lastBytes;
    int rc = b200sp_parse_device(e, e->d_src, totalStrided, blockSize, stride, nullptr, static_cast<uint32_t>(nBlocks),
                                 level, reinterpret_cast<b200sp_sequence *>(e->d_seqs), seqStride, e->d_counts, nullptr);
    if (rc) return rc;
    scan_counts_kernel<<<1, 1024, 0, e->stream>>>(e->d_counts, static_cast<uint32_t>(nBlocks), e->d_offsets);
    CU_TRY(cudaGetLastError(), "launch scan_counts_kernel");
    // worst case one entry per 4 input bytes plus one per block
    const size_t packedWorst = srcSize / 4 + 2 * nBlocks;
    CU_TRY(grow_dev(e->d_packed, e->d_packedCap, packedWorst), "cudaMalloc(packed)");
    pack_kernel<<<e->numSMs * 8, 256, 0, e->stream>>>(e->d_seqs, seqStride, e->d_counts, e->d_offsets,
                                                     static_cast<uint32_t>(nBlocks), e->d_packed);
    CU_TRY(cudaGetLastError(), "launch pack_kernel");

    // ---- D2H: sizes first, then exactly the packed entries
    CU_TRY(cudaMemcpyAsync(e->h_counts, e->d_counts, nBlocks * sizeof(uint32_t), cudaMemcpyDeviceToHost, e->stream), "D2H counts");
    CU_TRY(cudaMemcpyAsync(e->h_offsets, e->d_offsets, (nBlocks + 1) * sizeof(unsigned long long), cudaMemcpyDeviceToHost, e->stream), "D2H offsets");
    CU_TRY(cudaStreamSynchronize(e->stream), "sync after parse");
    const size_t total = static_cast<size_t>(e->h_offsets[nBlocks]);
    CU_TRY(grow_host(e->h_packed, e->h_packedCap, total), "cudaMallocHost(packed)");
    CU_TRY(cudaMemcpyAsync(e->h_packed, e->d_packed, total * sizeof(unsigned long long), cudaMemcpyDeviceToHost, e->stream), "D2H packed");
    CU_TRY(cudaStreamSynchronize(e->stream), "sync after D2H");

    res->nBlocks = static_cast<uint32_t>(nBlocks);
    res->counts = e->h_counts;
    res->offsets = reinterpret_cast<const uint64_t *>(e->h_offsets);
    res->packed = reinterpret_cast<const uint64_t *>(e->h_packed);
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
