/*
 * lz77_kernels.cuh — launch interface of the sm_100a LZ77 block parser (internal to the library;
 * the public C-ABI is include/b200seqprod.h).
 *
 * Replaces the QAT LZ4s engine behind cpaDcCompressData2 plus the QZSTD_decLz4s token walk
 * (/root/reference/src/qatseqprod.c:1245-1249, :1013-1091): one launch parses a batch of
 * independent blocks (<= 128 KiB each) resident in HBM into ZSTD_Sequence arrays in HBM.
 */
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace b200sp {

constexpr uint32_t kBlockMax      = 1u << 17;   // ZSTD_BLOCKSIZE_MAX
#ifndef B200SP_GROUPS
#define B200SP_GROUPS 52
#endif
// A window is one stage's worth of work: kGroups groups of 32 positions = two tasks per pool warp, so that the
// task queue balances inside a stage (32 tasks on 28 warps made every stage last two tasks with most warps
// waiting through the second).  The parse warps own one group per lane: they take a window in two halves.
constexpr uint32_t kGroups        = B200SP_GROUPS;
constexpr uint32_t kHalf          = kGroups / 2;    // groups per parse pass (<= 32: one lane each)
constexpr uint32_t kWindow        = kGroups * 32;   // positions per pipeline window
static_assert(kGroups % 2 == 0 && kHalf >= 9 && kHalf <= 32 && kGroups <= 64, "two parse passes of at most 32 groups; carries look 8 groups back");
constexpr uint32_t kRingC         = 4;          // candidate/match ring: windows in flight between hash and entries
// The match finder's table is a stable counting sort of the block's positions by key bucket ("hash chains" laid
// out contiguously, most recent last), in global memory (L2-resident scratch, one area per CTA):
constexpr uint32_t kBucketBits    = 13;         // buckets: top bits of the key hash
constexpr uint32_t kBuckets       = 1u << kBucketBits;
constexpr uint32_t kTagBits       = 15;         // further hash bits stored with every entry: candidates are filtered without touching their bytes
constexpr uint32_t kSegAlign      = 8;          // bucket segments start at multiples of 8 entries (15-bit segment starts, 16-byte aligned loads)
constexpr uint32_t kSortedCap     = kBlockMax + (kSegAlign - 1) * kBuckets;   // entries (u32) of one CTA's scratch
constexpr uint32_t kIdxCap        = 16383;      // a position's insertion index travels in 14 bits: scan widths stay below
constexpr uint32_t kBitmapBits    = 491520;     // repeated-key detector of the incompressible shortcut (60 KiB of shared memory)
#ifndef B200SP_HASH_GROUPS
#define B200SP_HASH_GROUPS 1
#endif
constexpr uint32_t kHashGroups    = B200SP_HASH_GROUPS;   // groups per hash task (their MATCH.ANY latencies overlap)
constexpr uint32_t kProbe         = 16;         // bytes compared per candidate before a winner is picked
constexpr uint32_t kMaxExtCap     = 256;
constexpr uint32_t kInputPad      = 320;        // readable slack after the block in shared memory
constexpr uint32_t kTmaChunk      = 16384;
constexpr uint32_t kTmaChunks     = kBlockMax / kTmaChunk;

#ifndef B200SP_EH_WARPS
#define B200SP_EH_WARPS 26
#endif
constexpr int kEhWarps     = B200SP_EH_WARPS;   // hash + extension warps
constexpr int kWarpTab     = kEhWarps;          // serial owner of the bucket counters (kEhWarps + 1 is spare)
constexpr int kWarpEntries = kEhWarps + 2;      // P1 (two warps, one per half window): lazy decisions + group entries
constexpr int kWarpEmit    = kEhWarps + 4;      // P2 (two warps, one per half window): scans + ZSTD_Sequence stores
constexpr int kNumWarps    = kEhWarps + 6;
static_assert(kGroups == 2 * kEhWarps, "two tasks per pool warp and stage");
constexpr int kThreads     = kNumWarps * 32;

// Shared-memory carve-up (bytes)
constexpr uint32_t kSmemInput   = kBlockMax + kInputPad;
constexpr uint32_t kSmemTabL    = kBuckets * 4;           // per bucket {segment start / 8 : 15 | entries so far : 17}; the histogram before that
constexpr uint32_t kSmemTabS    = 8192;                   // spare; first part of the detector's bitmap
constexpr uint32_t kSmemRingH   = 2 * kWindow * 4;        // hash words, H -> T
constexpr uint32_t kSmemRingC   = kRingC * kWindow * 4;   // candidates -> packed prefix maxima, H/T -> E -> P1
constexpr uint32_t kSmemRingL   = 2 * kWindow * 4;        // memoised parse decisions, P1 -> P2
constexpr uint32_t kSmemGroup   = (2 * kRingC + 4) * 64 * 4;   // gmax, gown, hasA, entA (64 entries per window)
constexpr uint32_t kSmemMisc    = 128;          // mbarriers + work-item slot + task counters
constexpr uint32_t kSmemTotal   = kSmemInput + kSmemTabL + kSmemTabS + kSmemRingH + kSmemRingC + kSmemRingL + kSmemGroup + kSmemMisc;
static_assert(kSmemTotal <= 232448, "exceeds 227 KB of shared memory per CTA");
static_assert(kSmemTabS >= 2 * kWindow * 2, "the tag ring (u16 per position, two windows) lives in the spare table");
static_assert(kSmemTabS + kSmemRingH + kSmemRingC + kSmemRingL >= kBitmapBits / 8, "the detector's bitmap overlays the spare table and the rings");

struct ParseParams {
    const uint8_t *src;        // batch base, 16-byte aligned
    uint64_t stride;           // byte distance between consecutive blocks (multiple of 16)
    uint64_t totalSize;        // only used when sizes == nullptr: block b holds min(blockSize, totalSize - b*stride)
    uint32_t blockSize;        // nominal block size (<= 128 KiB) when sizes == nullptr
    const uint32_t *sizes;     // optional per-block sizes (device), each <= 128 KiB
    uint32_t nBlocks;
    uint4 *seqs;               // ZSTD_Sequence arrays, seqStride entries per block
    uint64_t seqStride;
    uint32_t *counts;          // entries written per block (incl. the final literals entry)
    unsigned int *workCounter; // zeroed before launch; dynamic block scheduler
    unsigned int *errorFlag;   // optional: set to 1 when a bounded in-kernel wait ran out (never cleared by the kernel)
    uint32_t *sorted;          // scratch: kSortedCap entries per CTA of the grid (the counting-sorted positions)
    uint32_t keyMask;          // mask on bytes 4..7 for the key hash: 0 (4 B), 0xFF (5 B), 0xFFFF (6 B)
    uint32_t scan;             // bucket entries examined per position, most recent first (<= kIdxCap): the level-scaled depth
    uint32_t rank16;           // 1: candidates are ranked on their first 16 bytes and only the winner is extended (scan = 4 or 8)
    uint32_t minMatch;         // >= 4
    uint32_t extCap;           // <= kMaxExtCap
    uint32_t lazyDepth;        // 0..2
    unsigned long long *roleCycles; // optional (may be null): per-role busy cycles, developer profiling
};

// Fills the per-level fields of ParseParams. Returns false for levels outside 1..12
// (same range the reference accepts, /root/reference/src/qatseqprod.c:1132-1137).
bool params_for_level(int level, ParseParams &p);

// Launches the persistent parser (grid = min(blocks, number of SMs)) on `stream`.  p.sorted must hold
// scratch_bytes(numSMs) bytes that no other launch in flight uses.
cudaError_t launch_parse(const ParseParams &p, int numSMs, cudaStream_t stream);
inline size_t scratch_bytes(int numSMs) { return static_cast<size_t>(numSMs) * kSortedCap * sizeof(uint32_t); }

// One-time per-device setup (opt-in shared memory). Returns cudaSuccess or the CUDA error.
cudaError_t configure_kernels();

}  // namespace b200sp
