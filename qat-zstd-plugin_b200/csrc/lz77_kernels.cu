/*
 * lz77_kernels.cu — sm_100a LZ77 block parser: batch of independent <=128 KiB blocks in HBM ->
 * ZSTD_Sequence arrays in HBM.
 *
 * This is the B200 replacement for the QAT LZ4s engine + QZSTD_decLz4s
 * (/root/reference/src/qatseqprod.c:1245-1249 submit, :1013-1091 token walk, :1308-1313
 * incompressible shortcut, level handed to the engine at :1154).  Output convention is the reference's:
 * matchLength >= 3 for real matches, the last entry is {offset 0, trailing literals, matchLength 0}, `rep` is 0.
 *
 * One persistent CTA per SM; a CTA owns one block at a time:
 *   - the block is staged into shared memory with 1-D TMA bulk copies (UBLKCP);
 *   - phase A (all warps): every position's key hash feeds a histogram of the 8192 key buckets and a 480 Kbit
 *     repeated-key bitmap.  A block with no more repeated keys than chance produces is emitted as one literal run
 *     (the incompressible shortcut).  Otherwise the histogram is scanned into bucket segment starts: the table is a
 *     stable counting sort of the block's positions by bucket - every bucket is the chain of all earlier
 *     positions with that hash, contiguous, most recent last - in this CTA's scratch in global memory (L2);
 *   - the block then flows through rings of windows (52 groups of 32 positions), one role per stage time t:
 *       pool        26 warps    one queue of 52 fused tasks per stage (two per warp): task g scans the last `scan`
 *                               entries of each position's bucket for group g of window t-2 (16-byte loads of the
 *                               sorted table, tag filter, most recent first; the longest match wins - `scan` is the
 *                               level-scaled search depth; levels 1-4 rank candidates on 16 bytes, extend the winner and
 *                               probe the group's dominant offset, levels 5-12 measure every candidate in full, deep
 *                               scans by the whole warp), leaves the packed prefix maximum of match ends (levels 1-4) or
 *                               the own match of every position (levels 5-12), and - hidden behind its first loads -
 *                               hashes group g of window t
 *       T  table    window t-1   1 warp   walks the window in order: bucket counter += multiplicity (exact serial
 *                               insertion index of every position), stores the position into its slot of the sorted
 *                               table, hands slot and index to the pool
 *       P1 entries  window t-3   2 warps  (one per half window) lane = group: carry of the previous 8 groups,
 *                               lazy decisions memoised as link words, group entries iterated to the serial
 *                               fixed point; the entry of the second half is handed over through shared memory
 *       P2 emit     window t-4   2 warps  (one per half window) lane = group: follow the links, scans for
 *                               anchors / output slots, 16-byte ZSTD_Sequence stores; carry handed over likewise
 *       R  rep parse window t-3  1 warp   levels 5-12 instead of P1/P2: the serial repcode-aware lazy parse
 *                               (stage_rep_parse), warp-uniform state, lane-parallel inside a step
 * L2 policy: the input (TMA) and the sequence stores are evict-first, the table stores evict-last.
 * The result is bit-identical to oracle/seqmodel.c (the serial statement); oracle/lanemodel.c states
 * the P1/P2 formulation lane by lane on the CPU.  Integer/indexing work only: no tensor cores, no TMEM.
 */
#include "lz77_kernels.cuh"

#ifndef B200SP_SPIN_NS
#define B200SP_SPIN_NS 100
#endif


namespace b200sp {

// ------------------------------------------------------------------------------------------
// small PTX helpers
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p)
{
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}

__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}

__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    uint32_t done;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
    } while (!done);
}

// 1-D TMA bulk copy global -> shared, completion counted in bytes on an mbarrier.
// The input is read once: evict-first in L2, so that it does not push out the CTAs' sorted tables (the L2-resident
// scratch every bucket scan reads).
__device__ __forceinline__ void tma_load_1d(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar)
{
    asm volatile(
        "{\n\t.reg .b64 pol;\n\t"
        "createpolicy.fractional.L2::evict_first.b64 pol, 1.0;\n\t"
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], pol;\n\t}"
        ::"r"(dst), "l"(src), "r"(bytes), "r"(bar)
        : "memory");
}

// ZSTD_Sequence stores: written once, read by a later kernel or the copy engine - streaming (evict-first) as well.
__device__ __forceinline__ void st_seq(uint4 *p, uint32_t off, uint32_t lit, uint32_t len)
{
    asm volatile("st.global.cs.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(off), "r"(lit), "r"(len), "r"(0u) : "memory");
}

__device__ __forceinline__ void fence_proxy_async()
{
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// Shared-memory accesses by 32-bit shared-window address.  All of the CTA's shared memory is one
// dynamic array; its base is converted once (and made opaque, so ptxas keeps it in a register instead
// of re-deriving it from SR_CgaCtaId at every access) and every structure is an offset from it.
__device__ __forceinline__ uint32_t lds32(uint32_t a)
{
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
    return v;
}
// Same load, but free to be scheduled (not ordered against the other asm statements): only for words that
// cannot change while the calling task runs - the staged block, ring words written in earlier stages - and
// whose address depends on the task index (so it cannot move above the pop / the stage barrier).
__device__ __forceinline__ uint32_t ldsc32(uint32_t a)
{
    uint32_t v;
    asm("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ uint32_t lds16(uint32_t a)
{
    uint32_t v;
    asm volatile("{\n\t.reg .u16 t;\n\tld.shared.u16 t, [%1];\n\tcvt.u32.u16 %0, t;\n\t}" : "=r"(v) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ void sts32(uint32_t a, uint32_t v)
{
    asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory");
}
__device__ __forceinline__ void sts16(uint32_t a, uint32_t v)
{
    asm volatile("{\n\t.reg .u16 t;\n\tcvt.u16.u32 t, %1;\n\tst.shared.u16 [%0], t;\n\t}" ::"r"(a), "r"(v) : "memory");
}
__device__ __forceinline__ void sts8(uint32_t a, uint32_t v)
{
    asm volatile("{\n\t.reg .u16 t;\n\tcvt.u16.u32 t, %1;\n\tst.shared.u8 [%0], t;\n\t}" ::"r"(a), "r"(v) : "memory");
}
__device__ __forceinline__ void sts128(uint32_t a, uint32_t v)
{
    asm volatile("st.shared.v4.u32 [%0], {%1, %1, %1, %1};" ::"r"(a), "r"(v) : "memory");
}

// Unaligned little-endian 32-bit read from the staged block (`in` = shared address of the block).
__device__ __forceinline__ uint32_t ld32u(uint32_t in, uint32_t bytePos)
{
    const uint32_t a = in + (bytePos & ~3u);
    return __funnelshift_r(ldsc32(a), ldsc32(a + 4u), (bytePos & 3u) * 8u);
}

// Warp-uniform pop from a shared-memory task counter: one ATOMS by lane 0, one broadcast.
__device__ __forceinline__ uint32_t pop_task(uint32_t ctrAddr, uint32_t lane)
{
    uint32_t id = 0;
    // atom.inc (not .add): ptxas expands a predicated atom.add into a vote/popc aggregation sequence
    if (lane == 0) asm volatile("atom.shared.inc.u32 %0, [%1], 0x7fffffff;" : "=r"(id) : "r"(ctrAddr) : "memory");
    return __shfl_sync(0xFFFFFFFFu, id, 0);
}

// Bounded wait for a hand-off word between the parse / emit warps of one CTA (written a few hundred cycles after
// it is first polled): a word that never arrives would otherwise hang the device until the watchdog.  On overrun
// the error flag is raised (the host reports B200SP_ECUDA) and the warp carries on with whatever it reads.
constexpr uint32_t kSpinLimit = 1u << 22;      // x B200SP_SPIN_NS = ~0.4 s
__device__ __forceinline__ void spin_until(uint32_t addr, uint32_t expect, unsigned int *errorFlag)
{
    uint32_t spins = 0;
    while (lds32(addr) != expect) {
        __nanosleep(B200SP_SPIN_NS);
        if (++spins > kSpinLimit) { if (errorFlag) atomicExch(errorFlag, 1u); break; }
    }
}

// XOR swizzle of the ring words: conflict-free by group (a warp writes its group) and by lane (a parse lane
// reads the same local position in 32 groups).
__device__ __forceinline__ uint32_t ring_byte(uint32_t group, uint32_t lane)   // byte offset of a 32-bit ring word
{
    return (group * 32u + (lane ^ (group & 31u))) * 4u;
}

struct Shared {         // 32-bit shared-window addresses
    uint32_t in;        // the staged block (+ pad)
    uint32_t tab;       // u32[kBuckets]        phase A: histogram; then {segment start / 8 : 15 | entries so far : 17}
    uint32_t bitmap;    // kBitmapBits bits     phase A only: overlays the spare table and the rings
    uint32_t ringH;     // u32[2][kWindow]      H -> T: {first lane of its bucket:1 | lanes in the bucket:6 | first such lane:5 | earlier such lanes:5 | bucket:13}, 0 = invalid
    uint32_t ringT;     // u16[2][kWindow]      H -> T: tag (in the spare table)
    uint32_t ringC;     // u32[kRingC][kWindow] {slot in the sorted table:18 | insertion index (capped):14} (T) -> packed prefix maxima (E)
    uint32_t ringL;     // u32[2][kWindow]      P1 -> P2: memoised decisions {end:9 | take lane:5 | offset:17}
    uint32_t gmax;      // u32[kRingC][kGroups] packed farthest-reaching match of each group
    uint32_t gown;      // u32[kRingC][kGroups] lanes whose own prefix maximum is a usable match
    uint32_t hasA;      // u32[2][kGroups]      P1 -> P2: usable-match mask incl. the carry
    uint32_t entA;      // u32[2][kGroups]      P1 -> P2: position at which the parse enters each group
    uint32_t mbar;      // u64[kTmaChunks]
    uint32_t work;      // next block index
    uint32_t task;      // u32[2] per-stage task counters of the hash/extend warps (double-buffered)
    uint32_t curVal;    // entry warps: where the parse enters the next half window ...
    uint32_t curTag;    // ... and which half that is: 2 * window + half + 1 of the publisher (0 = none yet)
    uint32_t ecVal;     // emit warps: u32[3] anchor, previous offset, sequences written, after the publisher's half ...
    uint32_t ecTag;     // ... same numbering
    uint32_t emTag;     // emit warps: window + 1 once the first half's sequences are in memory
    uint32_t hits;      // phase A: positions whose bitmap bit was already set
    uint32_t scan;      // phase A: u32[32] warp totals of the segment scan (in the spare table)
};

// ------------------------------------------------------------------------------------------
// H: key hash of one 32-position group -> ringH word {valid:1 | bucket:13 | tag:15} (the top 28 bits of the hash).
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t key_hash(uint32_t lo, uint32_t hi, uint32_t keyMask)
{
    return lo * 0x9E3779B1u + (hi & keyMask) * 0xC2B2AE3Du;
}

// In two halves: MATCH.ANY takes a long time to come back, so a task starts the hash of its group early
// (hash_begin) and writes the ring words when its own work is done (hash_finish).
struct HashState { uint32_t v, m, tag; };

// Tag stored with every table entry (15 bits).  Levels 1-4: further bits of the key hash (the key covers 5 or 6 bytes).
// Levels 5-12 (4-byte keys): a hash of the NEXT four bytes, so that a deep scan measures only candidates that agree on
// eight bytes; the kNearUnfiltered nearest entries are measured whatever their tag (short matches pay only nearby).
__device__ __forceinline__ uint32_t entry_tag(bool next4, uint32_t v, uint32_t hi)
{
    return next4 ? (hi * 0xC2B2AE3Du) >> 17 : (v >> 4) & 0x7FFFu;
}

__device__ __forceinline__ HashState hash_begin(const Shared &S, uint32_t w, uint32_t group, uint32_t lane, uint32_t nh, uint32_t keyMask)
{
    const uint32_t p = w * kWindow + group * 32u + lane;
    const uint32_t a = S.in + (min(p, kBlockMax) & ~3u), sh = (p & 3u) * 8u;       // reads stay inside the padded buffer
    const uint32_t w0 = ldsc32(a), w1 = ldsc32(a + 4u), w2 = ldsc32(a + 8u);
    HashState h;
    const uint32_t hi = __funnelshift_r(w1, w2, sh);
    h.v = key_hash(__funnelshift_r(w0, w1, sh), hi, keyMask);
    h.tag = entry_tag(keyMask == 0u, h.v, hi);          // keyMask 0 <=> 4-byte keys <=> levels 5-12
    // lanes of the group in the same bucket (invalid lanes get unique keys so they never pair up)
    h.m = __match_any_sync(0xFFFFFFFFu, p < nh ? h.v >> (32 - kBucketBits) : (0x10000u | lane));
    return h;
}

__device__ __forceinline__ void hash_finish(const Shared &S, uint32_t w, uint32_t group, uint32_t lane, uint32_t nh, const HashState &h)
{
    // everything the table warp needs, so that it has as few instructions of its own as possible:
    // {first lane of its bucket:1 (bit 29) | lanes in the bucket:6 (23..28) | their first lane:5 | earlier lanes in the
    //  bucket:5 | bucket:13}; 0 = no valid position here
    const uint32_t p = w * kWindow + group * 32u + lane;
    const uint32_t ltMask = (1u << lane) - 1u;
    const uint32_t first = __ffs(h.m) - 1;
    const uint32_t word = (h.v >> (32 - kBucketBits)) | (__popc(h.m & ltMask) << 13) | (first << 18) |
                          (__popc(h.m) << 23) | (first == lane ? 1u << 29 : 0u);
    const uint32_t rb = ring_byte(group, lane);
    sts32(S.ringH + (w & 1u) * (kWindow * 4u) + rb, p < nh ? word : 0u);
    sts16(S.ringT + (w & 1u) * (kWindow * 2u) + (rb >> 1), h.tag);
}

// ------------------------------------------------------------------------------------------
// T: one warp walks the window group by group, in position order.  A position's insertion index in its bucket is
// the bucket's counter plus the number of earlier lanes of the group with the same bucket (MATCH.ANY); the first
// such lane adds the multiplicity to the counter with one shared-memory atomic and hands the old value to the
// others.  The atomics of consecutive groups are issued back to back (the unit applies them in order; nothing
// waits for a value before the next one is issued), so the walk is a stream, not a chain of round trips.
// The position goes into its slot of the sorted table (global scratch; read by the pool one stage later); slot
// and index go to the pool through ringC.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void stage_table(const Shared &S, uint32_t w, uint32_t lane, uint32_t *sorted)
{
    // kTUnroll groups per iteration: all their ring words are loaded, then all their atomics issued, before the
    // first result is needed - the walk pays the shared-memory latency once per iteration, not per group.  A lone
    // warp issues an instruction every several cycles: the hash tasks have prepared everything they could.
#ifdef B200SP_T_UNROLL
    constexpr uint32_t kTUnroll = B200SP_T_UNROLL;
#else
    constexpr uint32_t kTUnroll = 4;
#endif
    static_assert(kGroups % kTUnroll == 0, "the table warp walks whole iterations");
    const uint32_t rh = S.ringH + (w & 1u) * (kWindow * 4u);
    const uint32_t rt = S.ringT + (w & 1u) * (kWindow * 2u);
    const uint32_t rc = S.ringC + (w & (kRingC - 1)) * (kWindow * 4u);
    const uint32_t windowBase = w * kWindow;
    uint64_t polLast;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(polLast));
#pragma unroll 1
    for (uint32_t g0 = 0; g0 < kGroups; g0 += kTUnroll) {
        uint32_t hw[kTUnroll], tg[kTUnroll], old[kTUnroll];
#pragma unroll
        for (uint32_t k = 0; k < kTUnroll; k++) {
            const uint32_t rb = ring_byte(g0 + k, lane);
            hw[k] = lds32(rh + rb);
            tg[k] = lds16(rt + (rb >> 1));
        }
#pragma unroll
        for (uint32_t k = 0; k < kTUnroll; k++) {
            // the first lane of each bucket adds the bucket's multiplicity to its counter (the 17-bit counter never
            // carries into the start); predicated, not branched: the atomics of the iteration stream back to back
            asm volatile(
                "{\n\t.reg .pred q;\n\t"
                "setp.ne.u32 q, %3, 0;\n\t"
                "mov.u32 %0, 0;\n\t"
                "@q atom.shared.add.u32 %0, [%1], %2;\n\t}"
                : "=r"(old[k])
                : "r"(S.tab + (hw[k] & (kBuckets - 1u)) * 4u), "r"((hw[k] >> 23) & 63u), "r"(hw[k] & (1u << 29))
                : "memory");
        }
#pragma unroll
        for (uint32_t k = 0; k < kTUnroll; k++) {
            const bool valid = hw[k] != 0u;
            const uint32_t e = __shfl_sync(0xFFFFFFFFu, old[k], (hw[k] >> 18) & 31u);
            const uint32_t idx = (e & 0x1FFFFu) + ((hw[k] >> 13) & 31u);
            const uint32_t slot = (e >> 17) * kSegAlign + idx;
            // evict-last: the table is this CTA's working set in L2 for the whole block
            if (valid) asm volatile("st.global.L2::cache_hint.u32 [%0], %1, %2;" ::"l"(sorted + slot), "r"((windowBase + (g0 + k) * 32u + lane) | (tg[k] << 17)), "l"(polLast) : "memory");
            sts32(rc + ring_byte(g0 + k, lane), valid ? (slot << 14) | min(idx, kIdxCap) : 0u);
        }
    }
}

// ------------------------------------------------------------------------------------------
// E: bucket scan of one group -> best match per position -> prefix-max of match ends
// ringC out: packed prefix maximum {end - groupStart:9 | 31 - lane:6 | offset:17} (0 = none so far)
// ------------------------------------------------------------------------------------------
#ifndef B200SP_ALU_DIFF
#define B200SP_ALU_DIFF 0
#endif
// index (0..3) of the lowest non-zero byte of x != 0.  FLO/BREV share the memory-instruction queue with the
// shared-memory loads this stage is made of; three compares on the isolated lowest bit stay in the ALU.
__device__ __forceinline__ uint32_t low_byte_index(uint32_t x)
{
#if B200SP_ALU_DIFF
    const uint32_t low = x & (0u - x);
    return (low > 0xFFu ? 1u : 0u) + (low > 0xFFFFu ? 1u : 0u) + (low > 0xFFFFFFu ? 1u : 0u);
#else
    return (__ffs(x) - 1) >> 3;
#endif
}

__device__ __forceinline__ uint32_t first_diff_16(uint32_t x1, uint32_t x2, uint32_t x3)
{
    // length of the common prefix of two 16-byte strings whose first 4 bytes are equal,
    // given the XOR of words 1..3 (branch-free)
    uint32_t len = 16u;
    if (x3) len = 12u + low_byte_index(x3);
    if (x2) len = 8u + low_byte_index(x2);
    if (x1) len = 4u + low_byte_index(x1);
    return len;
}

__device__ __forceinline__ uint4 ldg128_cg(const uint32_t *p)      // L2 only: the entries were stored by the table warp a stage ago
{
    uint4 v;
    asm volatile("ld.global.cg.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    return v;
}

// Tail shared by both extension variants: "catch up" by one byte, packed prefix maximum, group words.
template <bool kOwnOnly>
__device__ __forceinline__ void finish_group(const Shared &S, uint32_t idx, uint32_t slotC, uint32_t group, uint32_t lane,
                                             uint32_t p, uint32_t a0, uint32_t bestLen, uint32_t bestOff,
                                             uint32_t minMatch, uint32_t extCap)
{
    if (kOwnOnly) {         // levels 5-12: the serial repcode-aware parse reads every position's own match {len:9 | offset:17}
        sts32(idx, bestLen ? (bestLen << 17) | bestOff : 0u);
        return;
    }
    const uint32_t in = S.in;
    // zstd's "catch up" by one byte: adopt the right neighbour's match if it also holds one byte earlier
    {
        const uint32_t l1 = __shfl_down_sync(0xFFFFFFFFu, bestLen, 1), o1 = __shfl_down_sync(0xFFFFFFFFu, bestOff, 1);
        const bool cand = lane < 31u && l1 != 0u && p >= o1 && l1 + 1u > bestLen && l1 + 1u <= extCap;
        const uint32_t back = cand ? lds32(in + ((p - o1) & ~3u)) >> (((p - o1) & 3u) * 8u) : ~a0;
        if (cand && ((back ^ a0) & 0xFFu) == 0u) { bestLen = l1 + 1u; bestOff = o1; }
    }
    // Pack {end relative to the group start (9 bits), 31 - lane (6 bits), offset (17 bits)}: an unsigned
    // max over packed words picks the farthest-reaching match and, on ties, the older one.
    uint32_t pk = 0;
    if (bestLen) pk = ((lane + bestLen) << 23) | ((31u - lane) << 17) | bestOff;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t o = __shfl_up_sync(0xFFFFFFFFu, pk, d);
        if (lane >= static_cast<uint32_t>(d)) pk = max(pk, o);
    }
    sts32(idx, pk);                                      // prefix-max within the group
    const uint32_t own = __ballot_sync(0xFFFFFFFFu, (pk >> 23) >= lane + minMatch);
    if (lane == 31) { sts32(S.gmax + (slotC * 64u + group) * 4u, pk); sts32(S.gown + (slotC * 64u + group) * 4u, own); }
}

__device__ __forceinline__ uint32_t ldg32_cg(const uint32_t *p)       // L2 only: the entries were stored by the table warp a stage ago
{
    uint32_t v;
    asm volatile("ld.global.cg.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// e[j] = entry (s - 1 - j) out of the two aligned chunks around it; t = (s - 1) & 3 is the place of s-1 in `hi`.
// As one array r = {lo.x .. lo.w, hi.x .. hi.w}: e[j] = r[t + 4 - j], a window of four moved by t: two levels of selects.
__device__ __forceinline__ void pick4(const uint4 &hi, const uint4 &lo, uint32_t t, uint32_t (&e)[4])
{
    const bool b0 = (t & 1u) != 0u, b1 = (t & 2u) != 0u;
    const uint32_t a0 = b0 ? lo.z : lo.y, a1 = b0 ? lo.w : lo.z, a2 = b0 ? hi.x : lo.w;
    const uint32_t a3 = b0 ? hi.y : hi.x, a4 = b0 ? hi.z : hi.y, a5 = b0 ? hi.w : hi.z;
    e[3] = b1 ? a2 : a0;
    e[2] = b1 ? a3 : a1;
    e[1] = b1 ? a4 : a2;
    e[0] = b1 ? a5 : a3;
}

// Fast classes (levels 1-4): the `scan` (4 or 8) entries before the position are probed branch-free on their first
// 16 bytes, four at a time; the longest probe wins (the nearer one on ties) and only the winner is extended: lanes
// continuing their predecessor's match derive their length from the run head, heads are extended by the whole warp.
template <bool kFuseHash>
__device__ __forceinline__ void stage_extend_fast(const Shared &S, uint32_t w, uint32_t group, uint32_t lane,
                                                  uint32_t p, uint32_t n, uint32_t nh, uint32_t minMatch,
                                                  uint32_t extCap, uint32_t scan, uint32_t keyMask, const uint32_t *sorted,
                                                  uint32_t hashWindow = 0, uint32_t hashNh = 0)
{
    const uint32_t slotC = w & (kRingC - 1);
    const uint32_t idx = S.ringC + slotC * (kWindow * 4u) + ring_byte(group, lane);
    const uint32_t cw = lds32(idx);       // ordered load: the first task's address is known before the stage barrier, and this word was written in the stage before
    const uint32_t in = S.in;
    const bool valid = p < nh;
    const uint32_t lim = valid ? min(n - p, extCap) : 0u;
    const uint32_t probe = min(lim, kProbe);
    const uint32_t s = cw >> 14;
    const uint32_t avail = valid ? min(scan, cw & 0x3FFFu) : 0u;
    // the four entries before us, s-1 .. s-4, lie in the aligned 16-byte chunk holding s-1 and in the one before it
    // (two loads of one sector per lane instead of four)
    const uint32_t cHi = avail ? (s - 1u) >> 2 : 0u;
    uint4 hi = make_uint4(0u, 0u, 0u, 0u), lo = make_uint4(0u, 0u, 0u, 0u);
    if (avail) hi = ldg128_cg(sorted + 4u * cHi);
    if (avail > ((s - 1u) & 3u) + 1u) lo = ldg128_cg(sorted + 4u * (cHi - 1u));      // only when `hi` does not hold all of them
    // our own first 16 bytes, as four unaligned words (reads stay inside the padded buffer)
    uint32_t a0, a1, a2, a3;
    {
        const uint32_t a = in + (min(p, kBlockMax) & ~3u), sh = (p & 3u) * 8u;
        const uint32_t x0 = ldsc32(a), x1 = ldsc32(a + 4u), x2 = ldsc32(a + 8u), x3 = ldsc32(a + 12u), x4 = ldsc32(a + 16u);
        a0 = __funnelshift_r(x0, x1, sh); a1 = __funnelshift_r(x1, x2, sh);
        a2 = __funnelshift_r(x2, x3, sh); a3 = __funnelshift_r(x3, x4, sh);
    }
    HashState hs = {0u, 0u};
    if (kFuseHash) hs = hash_begin(S, hashWindow, group, lane, hashNh, keyMask);    // finished at the end of the task
    const uint32_t tag = (key_hash(a0, a1, keyMask) >> 4) & 0x7FFFu;
    uint32_t bestLen = 0, bestOff = 0;
    uint32_t e[4];
    pick4(hi, lo, (s - 1u) & 3u, e);
#pragma unroll 1
    for (uint32_t base = 0; base < scan; base += 4u) {
        uint32_t q[4];
        bool ok[4];
        uint32_t yy[4][5];
#pragma unroll
        for (int j = 0; j < 4; j++) {
            ok[j] = base + j < avail && (e[j] >> 17) == tag;
            q[j] = ok[j] ? e[j] & 0x1FFFFu : 0u;
            const uint32_t qa = in + (q[j] & ~3u);
#pragma unroll
            for (int k = 0; k < 5; k++) yy[j][k] = ldsc32(qa + 4u * k);
        }
        uint4 hi2 = make_uint4(0u, 0u, 0u, 0u), lo2 = hi2;
        const uint32_t s2 = s - 4u - base;      // next round: entries s2-1 .. s2-4
        const bool more = base + 4u < scan && base + 4u < avail;
        if (more) {                      // in flight while we measure
            const uint32_t c2 = (s2 - 1u) >> 2;
            hi2 = ldg128_cg(sorted + 4u * c2);
            if (avail - base - 4u > ((s2 - 1u) & 3u) + 1u) lo2 = ldg128_cg(sorted + 4u * (c2 - 1u));
        }
#pragma unroll
        for (int j = 0; j < 4; j++) {    // most recent first: a farther candidate must be strictly longer
            const uint32_t sh = (q[j] & 3u) * 8u;
            const uint32_t b0 = __funnelshift_r(yy[j][0], yy[j][1], sh), b1 = __funnelshift_r(yy[j][1], yy[j][2], sh);
            const uint32_t b2 = __funnelshift_r(yy[j][2], yy[j][3], sh), b3 = __funnelshift_r(yy[j][3], yy[j][4], sh);
            uint32_t ml = min(first_diff_16(a1 ^ b1, a2 ^ b2, a3 ^ b3), probe);
            if (!ok[j] || b0 != a0) ml = 0;
            if (ml > bestLen) { bestLen = ml; bestOff = p - q[j]; }
        }
        pick4(hi2, lo2, (s2 - 1u) & 3u, e);
    }

    // The dominant offset of the group - the one most of its positions chose, the smaller one on ties - is probed by
    // every position: the stand-in for zstd's repeated-offset probe.  It needs no hash, so the 4- and 5-byte matches
    // between changed fields of record-like data are found, and a parse that keeps to one offset is what libzstd turns
    // into repcodes.  It wins unless the scan's winner is longer on the 16 probe bytes.
    {
        const uint32_t m = __match_any_sync(0xFFFFFFFFu, bestLen ? bestOff : (0x20000u | lane));
        const uint32_t key = bestLen ? (__popc(m) << 17) | (0x1FFFFu - bestOff) : 0u;
        const uint32_t top = __reduce_max_sync(0xFFFFFFFFu, key);
        const uint32_t D = top ? 0x1FFFFu - (top & 0x1FFFFu) : 0u;
        const bool tryD = D != 0u && valid && p >= D && bestOff != D;
        const uint32_t q = tryD ? p - D : 0u;
        const uint32_t qa = in + (q & ~3u), sh = (q & 3u) * 8u;
        const uint32_t y0 = ldsc32(qa), y1 = ldsc32(qa + 4u), y2 = ldsc32(qa + 8u), y3 = ldsc32(qa + 12u), y4 = ldsc32(qa + 16u);
        const uint32_t b0 = __funnelshift_r(y0, y1, sh), b1 = __funnelshift_r(y1, y2, sh);
        const uint32_t b2 = __funnelshift_r(y2, y3, sh), b3 = __funnelshift_r(y3, y4, sh);
        const uint32_t ml = min(first_diff_16(a1 ^ b1, a2 ^ b2, a3 ^ b3), probe);
        if (tryD && b0 == a0 && ml >= bestLen) { bestLen = ml; bestOff = D; }
    }

    // Long extension of winners that filled the probe.  A lane continuing its predecessor's
    // match (same offset, both filled the probe) derives its length from the run head; heads are
    // extended by the whole warp, 128 bytes per step, far enough to serve all their followers.
    const bool job = (bestLen == kProbe) && (lim > kProbe);
    const uint32_t prevOff = __shfl_up_sync(0xFFFFFFFFu, bestOff, 1);
    const uint32_t jobs = __ballot_sync(0xFFFFFFFFu, job);
    const bool follower = job && lane > 0 && ((jobs >> (lane - 1)) & 1u) && prevOff == bestOff;
    uint32_t heads = jobs & ~__ballot_sync(0xFFFFFFFFu, follower);
    const uint32_t myHead = job ? 31u - __clz(heads & ((2u << lane) - 1u)) : 32u;
    while (heads) {
        const uint32_t h = __ffs(heads) - 1;
        heads &= heads - 1;
        const uint32_t ph = p - lane + h;                               // head position (uniform)
        const uint32_t offh = __shfl_sync(0xFFFFFFFFu, bestOff, h);
        const uint32_t qh = ph - offh;
        const uint32_t reach = min(n - ph, extCap + 32u);               // how far any follower may need
        uint32_t U = reach;
#pragma unroll 1
        for (uint32_t k0 = kProbe; k0 < reach; k0 += 128u) {
            const uint32_t k = k0 + lane * 4u;
            uint32_t x = 0;
            if (k < reach) x = ld32u(in, ph + k) ^ ld32u(in, qh + k);
            const uint32_t bad = __ballot_sync(0xFFFFFFFFu, x != 0u);
            if (bad) {
                const uint32_t l = __ffs(bad) - 1;
                const uint32_t xl = __shfl_sync(0xFFFFFFFFu, x, l);
                U = min(reach, k0 + l * 4u + ((__ffs(xl) - 1) >> 3));
                break;
            }
        }
        if (myHead == h) bestLen = min(lim, U - (lane - h));
    }
    if (bestLen < minMatch) { bestLen = 0; bestOff = 0; }
    finish_group<false>(S, idx, slotC, group, lane, p, a0, bestLen, bestOff, minMatch, extCap);
    if (kFuseHash) hash_finish(S, hashWindow, group, lane, hashNh, hs);
}

// Levels 5-12 measure every candidate in full.  Common prefix of src[p..] and src[q..] (q < p), capped at lim, given
// our first 16 bytes a0..a3; 0 when it cannot be longer than `beat` (the bytes around index `beat` are compared first:
// most candidates of a deep scan fail there) or when the first four bytes differ.
__device__ __forceinline__ uint32_t measure_full(uint32_t in, uint32_t p, uint32_t q, uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                                                 uint32_t lim, uint32_t beat)
{
    if (beat >= 4u && ld32u(in, p + beat - 3u) != ld32u(in, q + beat - 3u)) return 0u;
    const uint32_t qa = in + (q & ~3u), sh = (q & 3u) * 8u;
    const uint32_t y0 = ldsc32(qa), y1 = ldsc32(qa + 4u), y2 = ldsc32(qa + 8u), y3 = ldsc32(qa + 12u), y4 = ldsc32(qa + 16u);
    if (__funnelshift_r(y0, y1, sh) != a0) return 0u;
    uint32_t ml = first_diff_16(a1 ^ __funnelshift_r(y1, y2, sh), a2 ^ __funnelshift_r(y2, y3, sh), a3 ^ __funnelshift_r(y3, y4, sh));
    if (ml == kProbe) {
        // 8 bytes per step: the alignments of both sides stay what they are (k advances by multiples of 4)
        const uint32_t pa = in + (p & ~3u), psh = (p & 3u) * 8u;
        uint32_t k = kProbe;
        while (k < lim) {
            const uint32_t x0 = ldsc32(pa + k), x1 = ldsc32(pa + k + 4u), x2 = ldsc32(pa + k + 8u);
            const uint32_t z0 = ldsc32(qa + k), z1 = ldsc32(qa + k + 4u), z2 = ldsc32(qa + k + 8u);
            const uint32_t d0 = __funnelshift_r(x0, x1, psh) ^ __funnelshift_r(z0, z1, sh), d1 = __funnelshift_r(x1, x2, psh) ^ __funnelshift_r(z1, z2, sh);
            if (d0 | d1) { k += d0 ? low_byte_index(d0) : 4u + low_byte_index(d1); break; }
            k += 8u;
        }
        ml = k;
    }
    return min(ml, lim);
}

#ifndef B200SP_COOP_MIN
#define B200SP_COOP_MIN 128
#endif
constexpr uint32_t kNearUnfiltered = 16;   // levels 5-12: the nearest entries of a bucket are measured whatever their tag
constexpr uint32_t kCoopMin = B200SP_COOP_MIN;     // positions with more bucket entries to scan than this are scanned by the whole warp

template <bool kFuseHash>
__device__ __forceinline__ void stage_extend_full(const Shared &S, uint32_t w, uint32_t group, uint32_t lane,
                                             uint32_t p, uint32_t n, uint32_t nh, uint32_t minMatch,
                                             uint32_t extCap, uint32_t scan, uint32_t keyMask, const uint32_t *sorted,
                                             uint32_t hashWindow = 0, uint32_t hashNh = 0)
{
    const uint32_t slotC = w & (kRingC - 1);
    const uint32_t idx = S.ringC + slotC * (kWindow * 4u) + ring_byte(group, lane);
    const uint32_t cw = lds32(idx);       // ordered load: the first task's address is known before the stage barrier, and this word was written in the stage before
    const uint32_t in = S.in;
    const bool valid = p < nh;
    const uint32_t lim = valid ? min(n - p, extCap) : 0u;
    // our own first 16 bytes, as four unaligned words (reads stay inside the padded buffer)
    uint32_t a0, a1, a2, a3;
    {
        const uint32_t a = in + (min(p, kBlockMax) & ~3u), sh = (p & 3u) * 8u;
        const uint32_t x0 = ldsc32(a), x1 = ldsc32(a + 4u), x2 = ldsc32(a + 8u), x3 = ldsc32(a + 12u), x4 = ldsc32(a + 16u);
        a0 = __funnelshift_r(x0, x1, sh); a1 = __funnelshift_r(x1, x2, sh);
        a2 = __funnelshift_r(x2, x3, sh); a3 = __funnelshift_r(x3, x4, sh);
    }
    // the entries of our bucket before us: [s - avail, s), most recent last
    const uint32_t s = cw >> 14;
    const uint32_t avail = valid ? min(scan, cw & 0x3FFFu) : 0u;
    const uint32_t first = s - avail;
    const bool shallow = avail != 0u && avail <= kCoopMin;
    uint32_t deep = __ballot_sync(0xFFFFFFFFu, avail > kCoopMin);
    // ---- short scans: every lane walks its own entries
    int32_t A = shallow ? static_cast<int32_t>((s - 1u) >> 2) : -1;        // 16-byte chunk of the table we read next
    const int32_t Alast = shallow ? static_cast<int32_t>(first >> 2) : 0;
    uint4 chunk = make_uint4(0u, 0u, 0u, 0u);
    if (A >= Alast) chunk = ldg128_cg(sorted + 4 * A);
    HashState hs = {0u, 0u};
    if (kFuseHash) hs = hash_begin(S, hashWindow, group, lane, hashNh, keyMask);    // finished at the end of the task
    const uint32_t tag = entry_tag(true, 0u, a1);
    uint32_t bestLen = 0, bestOff = 0;
    while (__any_sync(0xFFFFFFFFu, A >= Alast)) {
        uint32_t pend = 0;
        const uint4 cur = chunk;
        if (A >= Alast) {
            const uint32_t i0 = 4u * static_cast<uint32_t>(A);
            const uint32_t nearFrom = s - min(s, kNearUnfiltered);                     // entries from here on are measured whatever their tag
            if (((cur.x >> 17) == tag || i0 >= nearFrom) && i0 >= first) pend |= 1u;                       // i0 < s always holds
            if (((cur.y >> 17) == tag || i0 + 1u >= nearFrom) && i0 + 1u >= first && i0 + 1u < s) pend |= 2u;
            if (((cur.z >> 17) == tag || i0 + 2u >= nearFrom) && i0 + 2u >= first && i0 + 2u < s) pend |= 4u;
            if (((cur.w >> 17) == tag || i0 + 3u >= nearFrom) && i0 + 3u >= first && i0 + 3u < s) pend |= 8u;
            A--;
            if (A >= Alast) chunk = ldg128_cg(sorted + 4 * A);                          // in flight while we measure
        }
        while (__any_sync(0xFFFFFFFFu, pend != 0u)) {
            if (pend) {
                // most recent first; a later (farther) candidate must be strictly longer
                const uint32_t j = 31u - __clz(pend);
                pend ^= 1u << j;
                const uint32_t e = j == 3u ? cur.w : j == 2u ? cur.z : j == 1u ? cur.y : cur.x;
                const uint32_t q = e & 0x1FFFFu;
                const uint32_t ml = measure_full(in, p, q, a0, a1, a2, a3, lim, bestLen);
                if (ml > bestLen) { bestLen = ml; bestOff = p - q; }
                if (bestLen >= lim) { pend = 0u; A = Alast - 1; }      // as long as a match can get: nothing farther can win
            }
        }
    }
    // ---- deep scans: one position at a time, the whole warp on its bucket - 128 entries per step (one 16-byte chunk
    // per lane, the nearest in lane 0), every lane measures the candidates among its four; the longest wins, the
    // nearest on ties, and whatever a later step finds must be strictly longer (the serial order of the model)
    // The chunk of the next step (of this position, else of the next deep position) is in flight while the current
    // one is measured: a lone position would otherwise pay the table's L2 latency at every step.
    {
        uint32_t i = 32u, si = 0, firsti = 0;
        int32_t A0 = 0, AlastI = 0;
        uint4 nxt = make_uint4(0u, 0u, 0u, 0u);
        auto next_position = [&]() {            // pops the next deep position and starts the load of its nearest entries
            i = 32u;
            if (deep) {
                i = __ffs(deep) - 1;
                deep &= deep - 1u;
                si = __shfl_sync(0xFFFFFFFFu, s, i); firsti = __shfl_sync(0xFFFFFFFFu, first, i);
                A0 = static_cast<int32_t>((si - 1u) >> 2); AlastI = static_cast<int32_t>(firsti >> 2);
                const int32_t Al = A0 - static_cast<int32_t>(lane);
                if (Al >= AlastI) nxt = ldg128_cg(sorted + 4 * Al);
            }
        };
        next_position();
        while (i < 32u) {
            const uint32_t ci = i, cSi = si, cFirst = firsti;
            const int32_t cAlast = AlastI;
            const uint32_t limi = __shfl_sync(0xFFFFFFFFu, lim, ci), tagi = __shfl_sync(0xFFFFFFFFu, tag, ci);
            const uint32_t c0 = __shfl_sync(0xFFFFFFFFu, a0, ci), c1 = __shfl_sync(0xFFFFFFFFu, a1, ci);
            const uint32_t c2 = __shfl_sync(0xFFFFFFFFu, a2, ci), c3 = __shfl_sync(0xFFFFFFFFu, a3, ci);
            const uint32_t pi = p - lane + ci;
            uint32_t gLen = 0, gOff = 0;
            for (;;) {
                const uint4 c = nxt;
                const int32_t Al = A0 - static_cast<int32_t>(lane);
                const bool more = A0 - 32 >= cAlast;            // this position has a farther step
                if (more) {
                    A0 -= 32;
                    const int32_t Aln = A0 - static_cast<int32_t>(lane);
                    if (Aln >= cAlast) nxt = ldg128_cg(sorted + 4 * Aln);
                } else next_position();
                uint32_t lLen = gLen, lOff = 0;
                if (Al >= cAlast) {
                    const uint32_t i0 = 4u * static_cast<uint32_t>(Al);
#pragma unroll
                    for (int j = 3; j >= 0; j--) {
                        const uint32_t e = j == 3 ? c.w : j == 2 ? c.z : j == 1 ? c.y : c.x;
                        if (((e >> 17) == tagi || i0 + j + kNearUnfiltered >= cSi) && i0 + j >= cFirst && i0 + j < cSi) {
                            const uint32_t q = e & 0x1FFFFu;
                            const uint32_t ml = measure_full(in, pi, q, c0, c1, c2, c3, limi, lLen);
                            if (ml > lLen) { lLen = ml; lOff = pi - q; }
                        }
                    }
                }
                const uint32_t top = __reduce_max_sync(0xFFFFFFFFu, lOff ? (lLen << 17) | (0x1FFFFu - lOff) : 0u);
                if (top) { gLen = top >> 17; gOff = 0x1FFFFu - (top & 0x1FFFFu); }
                if (!more) break;
                if (gLen >= limi) { next_position(); break; }      // as long as a match can get: drop the farther steps
            }
            if (lane == ci) { bestLen = gLen; bestOff = gOff; }
        }
    }
    if (bestLen < minMatch) { bestLen = 0; bestOff = 0; }
    finish_group<true>(S, idx, slotC, group, lane, p, a0, bestLen, bestOff, minMatch, extCap);
    if (kFuseHash) hash_finish(S, hashWindow, group, lane, hashNh, hs);
}

// ------------------------------------------------------------------------------------------
// P1: where the serial parse enters every group of one window (lane j = group j).
// oracle/lanemodel.c is the lane-by-lane CPU statement of this function and of stage_emit.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ int32_t gain_packed(uint32_t b, uint32_t lane)
{
    return static_cast<int32_t>(((b >> 23) - lane) * 4u) - static_cast<int32_t>(31 - __clz((b & 0x1FFFFu) + 1u));
}

// One decision: the cursor stands on local position p (a usable match starts there).  The lazy rule
// (look-ahead never leaves the group) yields either a take link {end:9 | take lane:5 | offset:17} or a hop
// link {1 << 31 | next position << 22}; the three packed words are loaded together.
__device__ __forceinline__ uint32_t eval_step(uint32_t pkRow, uint32_t group, uint32_t p, uint32_t c,
                                              uint32_t has, uint32_t lazyDepth)
{
    const uint32_t q1 = p + 1u, q2 = p + 2u;
    const uint32_t w0 = lds32(pkRow + ring_byte(group, p));
    const uint32_t w1 = lds32(pkRow + ring_byte(group, min(q1, 31u)));
    const uint32_t w2 = lds32(pkRow + ring_byte(group, min(q2, 31u)));
    const uint32_t b0 = max(w0, c), b1 = max(w1, c), b2 = max(w2, c);
    const int32_t g0 = gain_packed(b0, p), g1 = gain_packed(b1, q1), g2 = gain_packed(b2, q2);
    const bool ok1 = lazyDepth >= 1u && q1 < 32u && ((has >> (q1 & 31u)) & 1u);
    const bool ok2 = ok1 && lazyDepth >= 2u && q2 < 32u && ((has >> (q2 & 31u)) & 1u);
    uint32_t L = ((b0 >> 23) << 22) | (p << 17) | (b0 & 0x1FFFFu);
    if (ok1 && g1 > g0 + 4) L = 0x80000000u | (q1 << 22);
    else if (ok2 && g2 > g0 + 7) L = 0x80000000u | (q2 << 22);
    return L;
}

__device__ __forceinline__ void stage_entries(const Shared &S, uint32_t w, uint32_t half, uint32_t lane,
                                              uint32_t minMatch, uint32_t lazyDepth, unsigned int *errorFlag)
{
    // Lane j owns group half * kHalf + j of the window.  The two halves of a window are handled by two warps in
    // the same stage: where the parse enters a half is published by the warp of the half before it (the second
    // half of the previous window: last stage; the first half of this window: any moment now), tagged
    // 2 * window + half + 1.
    const uint32_t slot = w & (kRingC - 1), base = w * kWindow, group = half * kHalf + lane;
    const bool act = lane < kHalf;
    const uint32_t pkRow = S.ringC + slot * (kWindow * 4u);
    const uint32_t linkRow = S.ringL + (w & 1u) * (kWindow * 4u);
    // ---- carry: farthest-reaching match of the previous 8 groups (a match is at most extCap = 256
    // bytes long, so nothing older can reach in), re-based to this group; it wins ties (it is older)
    uint32_t c = 0;
#pragma unroll
    for (uint32_t k = 1; k <= 8; k++) {
        int gg = static_cast<int>(group) - static_cast<int>(k);
        uint32_t sl = slot;
        bool ok = true;
        if (gg < 0) { ok = w > 0; gg += kGroups; sl = (w - 1) & (kRingC - 1); }
        const uint32_t v = ok ? lds32(S.gmax + (sl * 64u + gg) * 4u) : 0u;
        const uint32_t rel = v >> 23;
        if (rel > 32u * k) c = max(c, ((rel - 32u * k) << 23) | ((32u + k) << 17) | (v & 0x1FFFFu));
    }
    const uint32_t cRel = c >> 23;
    uint32_t cover = 0;
    if (cRel >= minMatch) cover = (cRel - minMatch >= 31u) ? 0xFFFFFFFFu : (2u << (cRel - minMatch)) - 1u;
    const uint32_t has = act ? lds32(S.gown + (slot * 64u + group) * 4u) | cover : 0u;
    const uint32_t segStart = base + group * 32u, segEnd = segStart + 32u;

    // ---- every lane guesses that the parser enters its group at its first position; the guesses are
    // corrected from lane 0 upward until nothing changes.  A lane whose entry lies beyond its group is
    // passed over (a long match covers it): the entry of a lane is the exit of the nearest earlier lane
    // that is NOT passed over, i.e. the prefix maximum over live lanes.  A walk depends only on where it
    // starts, and every decision is memoised, so a corrected lane re-evaluates nothing it has seen.
    // First guess: the parse arrives through the carried match (the farthest-reaching one usually is the
    // one the previous groups ended with); any guess converges to the same fixed point.
    // That includes lane 0: it starts on a guess as well and takes the published cursor as soon as it is there.
    const uint32_t expect = 2u * w + half;              // tag of the half before this one
    bool known = expect == 0u;
    uint32_t cursor = 0;
    if (!known && lds32(S.curTag) == expect) { __threadfence_block(); cursor = lds32(S.curVal); known = true; }
    uint32_t entry = (lane == 0 && known) ? max(cursor, segStart) : segStart + min(cRel, 32u);
    if (!act) entry = 0xFFFFFFFFu;                      // lanes beyond the half stay inert: never live, never change
    uint32_t visited = 0, pm = 0, walked = 0xFFFFFFFFu, exitPos = 0;
    for (;;) {
        if (entry != walked) {                   // a lane whose entry did not change keeps its exit
            walked = entry;
            const bool live = entry < segEnd;
            uint32_t cur = live ? entry - segStart : 32u;
            while (cur < 32u) {
                const uint32_t m = has & (0xFFFFFFFFu << cur);
                if (!m) { cur = 32u; break; }
                const uint32_t p0 = __ffs(m) - 1;
                uint32_t L;
                if ((visited >> p0) & 1u) L = lds32(linkRow + ring_byte(group, p0));
                else {
                    L = eval_step(pkRow, group, p0, c, has, lazyDepth);
                    sts32(linkRow + ring_byte(group, p0), L);
                    visited |= 1u << p0;
                }
                cur = (L >> 22) & 0x1FFu;
            }
            exitPos = live ? segStart + cur : 0u;                    // passed-over lanes contribute nothing
        }
        __syncwarp();
        pm = lane == 0 ? max(exitPos, entry) : exitPos;              // lane 0 also carries the window's entry cursor
        // an exit lies at most 9 groups ahead (32 + extCap + 32 bytes): a maximum over the previous 15 lanes is enough
#pragma unroll
        for (int d = 1; d < 16; d <<= 1) {
            const uint32_t o = __shfl_up_sync(0xFFFFFFFFu, pm, d);
            if (lane >= static_cast<uint32_t>(d)) pm = max(pm, o);
        }
        const uint32_t prevMax = __shfl_up_sync(0xFFFFFFFFu, pm, 1);
        uint32_t want = (lane == 0 || !act) ? entry : max(prevMax, segStart);
        bool changed = want != entry;
        if (!__any_sync(0xFFFFFFFFu, changed)) {
            if (known) break;
            // converged on a guessed entry of the half: wait for the real one (warp-uniform spin)
            spin_until(S.curTag, expect, errorFlag);
            __threadfence_block();
            cursor = lds32(S.curVal);
            known = true;
            if (lane == 0) { want = max(cursor, segStart); changed = want != entry; }
            if (!__any_sync(0xFFFFFFFFu, changed)) break;
        } else if (!known && lds32(S.curTag) == expect) {       // warp-uniform: every lane reads the same word
            __threadfence_block();
            cursor = lds32(S.curVal);
            known = true;
            if (lane == 0) want = max(cursor, segStart);
        }
        entry = want;
    }
    if (act) {
        sts32(S.hasA + ((w & 1u) * 64u + group) * 4u, has);
        sts32(S.entA + ((w & 1u) * 64u + group) * 4u, entry);
    }
    // publish where the parse enters the next half
    const uint32_t next = max(__shfl_sync(0xFFFFFFFFu, pm, 31), base + (half + 1u) * (kHalf * 32u));
    if (lane == 0) {
        sts32(S.curVal, next);
        __threadfence_block();
        sts32(S.curTag, expect + 1u);
    }
}

// ------------------------------------------------------------------------------------------
// P2: emit one window.  Lane j owns positions [base + 32 j, base + 32 j + 32).
// ------------------------------------------------------------------------------------------
struct EmitCarry {           // uniform across the warp, carried from window to window
    uint32_t anchor;         // end of the last emitted match
    uint32_t prevOff;        // offset of the last emitted match
    uint32_t nOut;           // sequences written so far
};

// The two halves of a window are emitted by two warps in the same stage.  The carry (anchor, previous offset,
// sequences written) travels through shared memory like the entry cursor: the first half's warp publishes it
// right after its scans, so the second half's warp - which has done its counting walk meanwhile - waits little.
__device__ __forceinline__ void stage_emit(const Shared &S, uint32_t w, uint32_t half, uint32_t lane, EmitCarry &ec, uint4 *out, unsigned int *errorFlag)
{
    const uint32_t base = w * kWindow, group = half * kHalf + lane;
    const bool act = lane < kHalf;
    const uint32_t linkRow = S.ringL + (w & 1u) * (kWindow * 4u);
    const uint32_t has = act ? lds32(S.hasA + ((w & 1u) * 64u + group) * 4u) : 0u;
    const uint32_t entry = act ? lds32(S.entA + ((w & 1u) * 64u + group) * 4u) : 0xFFFFFFFFu;
    const uint32_t segStart = base + group * 32u, segEnd = segStart + 32u;
    const uint32_t cur0 = entry < segEnd ? entry - segStart : 32u;

    // ---- counting walk along the memoised links
    uint32_t cnt = 0, merges = 0, firstPos = 0, firstOff = 0, lastEnd = 0, lastOff = 0;
    {
        uint32_t cur = cur0;
        while (cur < 32u) {
            const uint32_t m = has & (0xFFFFFFFFu << cur);
            if (!m) break;
            const uint32_t L = lds32(linkRow + ring_byte(group, __ffs(m) - 1));
            if (L >> 31) { cur = (L >> 22) & 0x1FFu; continue; }         // hop: a later start is better (lazy)
            const uint32_t p = segStart + ((L >> 17) & 31u), end = segStart + (L >> 22), off = L & 0x1FFFFu;
            if (cnt && p == lastEnd && off == lastOff) merges++;
            if (!cnt) { firstPos = p; firstOff = off; }
            cnt++; lastEnd = end; lastOff = off;
            cur = L >> 22;
        }
    }
    __syncwarp();

    // ---- the carry of the half before this one
    {
        const uint32_t expect = 2u * w + half;
        if (expect == 0u) { ec.anchor = 0u; ec.prevOff = 0u; ec.nOut = 0u; }
        else {
            spin_until(S.ecTag, expect, errorFlag);
            __threadfence_block();
            ec.anchor = lds32(S.ecVal); ec.prevOff = lds32(S.ecVal + 4u); ec.nOut = lds32(S.ecVal + 8u);
        }
    }

    // ---- anchor / previous offset at each lane's entry: exclusive "last match" scan
    uint32_t aE = cnt ? lastEnd : 0u, aO = lastOff;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t oe = __shfl_up_sync(0xFFFFFFFFu, aE, d);
        const uint32_t oo = __shfl_up_sync(0xFFFFFFFFu, aO, d);
        if (lane >= static_cast<uint32_t>(d) && !(aE > oe)) { aE = oe; aO = oo; }
    }
    const uint32_t totE = __shfl_sync(0xFFFFFFFFu, aE, 31), totO = __shfl_sync(0xFFFFFFFFu, aO, 31);
    uint32_t anchor = __shfl_up_sync(0xFFFFFFFFu, aE, 1), prevOff = __shfl_up_sync(0xFFFFFFFFu, aO, 1);
    if (lane == 0 || anchor == 0) { anchor = ec.anchor; prevOff = ec.prevOff; }

    // ---- output slots
    const bool headMerge = cnt && firstPos == anchor && firstOff == prevOff && anchor > 0;
    const uint32_t fresh = cnt - merges - (headMerge ? 1u : 0u);
    uint32_t incl = fresh;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t v = __shfl_up_sync(0xFFFFFFFFu, incl, d);
        if (lane >= static_cast<uint32_t>(d)) incl += v;
    }
    const uint32_t firstIdx = ec.nOut + incl - fresh;
    const uint32_t total = __shfl_sync(0xFFFFFFFFu, incl, 31);
    if (totE) { ec.anchor = totE; ec.prevOff = totO; }
    ec.nOut += total;
    if (lane == 0) {            // the next half can run its scans while this one emits
        sts32(S.ecVal, ec.anchor); sts32(S.ecVal + 4u, ec.prevOff); sts32(S.ecVal + 8u, ec.nOut);
        __threadfence_block();
        sts32(S.ecTag, 2u * w + half + 1u);
    }

    // ---- emitting walk.  New sequences go to out[firstIdx...]; a leading continuation of an earlier
    // lane's sequence is added to out[firstIdx - 1].matchLength afterwards.
    uint32_t headAdd = 0;
    if (cnt) {
        uint32_t cur = cur0, outIdx = firstIdx;
        uint32_t openOff = 0, openLit = 0, openLen = 0;   // sequence being accumulated
        bool haveOpen = false;
        while (cur < 32u) {
            const uint32_t m = has & (0xFFFFFFFFu << cur);
            if (!m) break;
            const uint32_t L = lds32(linkRow + ring_byte(group, __ffs(m) - 1));
            if (L >> 31) { cur = (L >> 22) & 0x1FFu; continue; }         // hop
            const uint32_t p = segStart + ((L >> 17) & 31u), end = segStart + (L >> 22), o = L & 0x1FFFFu;
            const uint32_t lit = p - anchor, len = end - p;
            if (lit == 0 && o == prevOff && anchor > 0) {
                if (haveOpen) openLen += len; else headAdd += len;
            } else {
                if (haveOpen) st_seq(out + outIdx++, openOff, openLit, openLen);
                openOff = o; openLit = lit; openLen = len; haveOpen = true;
            }
            anchor = end; prevOff = o;
            cur = L >> 22;
        }
        if (haveOpen) st_seq(out + outIdx, openOff, openLit, openLen);
    }
    __syncwarp();
    // A continuation is added to a sequence an earlier lane wrote - for the second half possibly a lane of the
    // first half's warp, which must have stored it before (its stores of earlier windows are a stage old).
    if (half == 1u) {
        spin_until(S.emTag, w + 1u, errorFlag);
        __threadfence();
    }
    if (headAdd) atomicAdd(&out[firstIdx - 1].z, headAdd);   // continuation of an earlier lane's sequence
    __syncwarp();
    if (half == 0u && lane == 0) {
        __threadfence();
        sts32(S.emTag, w + 1u);
    }
}

// ------------------------------------------------------------------------------------------
// R: the serial repcode-aware lazy parse of levels 5-12 (oracle/seqmodel.c:rep_parse is the normative text).
// ONE warp per block walks the parse sequence by sequence; what is parallel is inside a step: 32 positions are
// tested for "a match or a repeated-offset match starts here" at once, lengths at the repeated offsets are
// measured 64 or 128 bytes per step, the catch-up to the left is one ballot.  All state is warp-uniform.
// The bucket scans of these levels take several milliseconds per block; this walk hides behind them.
// ------------------------------------------------------------------------------------------
struct RepState {
    uint32_t ip, anchor, rep1, rep2, nOut;
    uint32_t openOff, openLit, openLen;     // the sequence being accumulated (openLen == 0: none)
};

// Common prefix of src[a..] and src[a - off..] from byte `from` on (the bytes before are known equal), whole warp,
// 128 bytes per step, at most n - a.
__device__ __forceinline__ uint32_t coop_len(uint32_t in, uint32_t a, uint32_t off, uint32_t from, uint32_t n, uint32_t lane)
{
    const uint32_t lim = n - a;
    for (uint32_t k0 = from & ~3u; k0 < lim; k0 += 128u) {
        const uint32_t k = k0 + lane * 4u;
        uint32_t x = 1u;                                               // beyond the block: a difference at its first byte
        if (k < lim) x = ld32u(in, a + k) ^ ld32u(in, a - off + k);
        const uint32_t bad = __ballot_sync(0xFFFFFFFFu, x != 0u);
        if (bad) {
            const uint32_t l = __ffs(bad) - 1;
            const uint32_t xl = __shfl_sync(0xFFFFFFFFu, x, l);
            return min(lim, k0 + l * 4u + ((__ffs(xl) - 1) >> 3));
        }
    }
    return lim;
}

__device__ __forceinline__ void rep_emit(RepState &st, uint4 *out, uint32_t lane, uint32_t off, uint32_t lit, uint32_t len)
{
    if (st.openLen && lane == 0) st_seq(out + st.nOut, st.openOff, st.openLit, st.openLen);
    if (st.openLen) st.nOut++;
    st.openOff = off; st.openLit = lit; st.openLen = len;
}

__device__ __forceinline__ void stage_rep_parse(const Shared &S, uint32_t w, uint32_t lane, uint32_t n, uint32_t nh,
                                                uint32_t lazyDepth, uint32_t extCap, RepState &st, uint4 *out)
{
    const uint32_t in = S.in;
    // own matches {len:9 | offset:17} are known below `avail` (this window's extension stage ended a stage ago); the
    // last stage of a block knows them all (none exists from nh on).  The walk never reads further back than the
    // window before this one.
    const uint32_t wBase = w * kWindow;
    const uint32_t avail = min(nh, wBase + kWindow);
    const bool last = avail == nh;
    const uint32_t rowCur = S.ringC + (w & (kRingC - 1)) * (kWindow * 4u), rowPrev = S.ringC + ((w - 1u) & (kRingC - 1)) * (kWindow * 4u);
    uint32_t ip = st.ip, anchor = st.anchor, rep1 = st.rep1, rep2 = st.rep2;
    uint32_t dec0 = 0;          // where the decision in progress started (it is taken up again from there when deferred)
    bool chain = false;         // a decision is in progress and its best so far is the own match of `ip`
    // The span: 32 positions from sBase with their own matches and the masks that do not depend on the repeated
    // offset (H: has a match, M1 / M2: its own-only lazy step).  It is kept while the walk stays inside it - a take
    // usually lands a dozen positions further on - and only "the repeated offset matches here" (E) is computed again.
    uint32_t sBase = 0, ow = 0, H = 0, M1 = 0, M2 = 0;
    bool spanValid = false;
    while (ip < avail) {
        if (!spanValid || ip < sBase || ip - sBase > 20u) {
            sBase = ip;
            spanValid = true;
            const uint32_t p = sBase + lane;
            ow = 0;
            if (p < avail) {
                const uint32_t r = p >= wBase ? p - wBase : p + kWindow - wBase;
                ow = lds32((p >= wBase ? rowCur : rowPrev) + ring_byte(r >> 5, r & 31u));
            }
            const int32_t G = static_cast<int32_t>((ow >> 17) * 4u) - static_cast<int32_t>(31 - __clz((ow & 0x1FFFFu) + 1u));
            const uint32_t ow1 = __shfl_down_sync(0xFFFFFFFFu, ow, 1), ow2 = __shfl_down_sync(0xFFFFFFFFu, ow, 2);
            const int32_t G1 = __shfl_down_sync(0xFFFFFFFFu, G, 1), G2 = __shfl_down_sync(0xFFFFFFFFu, G, 2);
            // the lazy step of a position whose best so far is its own match, repeated offset not involved (lanes 0..29)
            const bool mv1 = lazyDepth >= 1u && ow1 != 0u && G1 > G + 4;
            const bool mv2 = !mv1 && lazyDepth >= 2u && ow2 != 0u && G2 > G + 7;
            H = __ballot_sync(0xFFFFFFFFu, ow != 0u);
            M1 = __ballot_sync(0xFFFFFFFFu, mv1);
            M2 = __ballot_sync(0xFFFFFFFFu, mv2);
        }
        const uint32_t shift = ip - sBase;                  // the walk stands on lane `shift` of the span
        bool eq = false;
        {
            const uint32_t p = sBase + lane;
            if (rep1 != 0u && p >= rep1 && p + 4u <= n) eq = ld32u(in, p) == ld32u(in, p - rep1);
        }
        const uint32_t E = __ballot_sync(0xFFFFFFFFu, eq);
        const uint32_t C = H & ~(E >> 1) & ~(E >> 2);       // own match here, repeated offset silent one and two bytes on
        const uint32_t lim = avail - sBase;                 // positions of the span that are known (>= 1)
        uint32_t s = shift;
        if (!chain) {
            // next position at which an own match starts or the repeated offset matches one byte further on
            uint32_t starts = (H | (E >> 1)) & 0x7FFFFFFFu & (0xFFFFFFFFu << shift);      // lane 31 cannot see the byte after the span
            if (lim < 32u) starts &= (1u << lim) - 1u;
            if (!starts) { ip = sBase + min(lim, 31u); spanValid = false; continue; }
            s = __ffs(starts) - 1;
            dec0 = sBase + s;
        }
        uint32_t ml = 0, off = 0, start = 0;
        bool isRep = false, done = false, defer = false;
        while (!done) {
            if (s > 29u) break;                                               // look-ahead leaves the span: move the span
            if (!last && s + lazyDepth >= lim) { defer = true; break; }        // ... or needs the next window: next stage
            const uint32_t base = sBase + s;
            if ((C >> s) & 1u) {
                if ((M1 >> s) & 1u) { s += 1u; chain = true; continue; }
                if ((M2 >> s) & 1u) { s += 2u; chain = true; continue; }
                const uint32_t o0 = __shfl_sync(0xFFFFFFFFu, ow, s);
                ml = o0 >> 17; off = o0 & 0x1FFFFu; start = base; isRep = false; done = true;
                break;
            }
            // ---- the repeated offset is involved (oracle/seqmodel.c:rep_parse, one pass of its look-ahead loop):
            // its lengths one and two bytes ahead, lanes 0-15 / 16-31, 64 bytes per half and step
            uint32_t R1 = 0, R2 = 0;
            {
                const uint32_t h = lane >> 4, j = lane & 15u, q = base + 1u + h;
                bool open = ((E >> (s + 1u + h)) & 1u) != 0u && (h == 0u || lazyDepth >= 2u);
                uint32_t len = 0;
                for (uint32_t k0 = 0; __any_sync(0xFFFFFFFFu, open); k0 += 64u) {
                    const uint32_t k = k0 + j * 4u;
                    uint32_t x = 1u;
                    if (open && q + k < n) x = ld32u(in, q + k) ^ ld32u(in, q - rep1 + k);
                    const uint32_t bad = (__ballot_sync(0xFFFFFFFFu, x != 0u) >> (h * 16u)) & 0xFFFFu;
                    const uint32_t l = bad ? __ffs(bad) - 1 : 0u;
                    const uint32_t xl = __shfl_sync(0xFFFFFFFFu, x, h * 16u + l);
                    if (open && bad) { len = min(n - q, k0 + l * 4u + ((__ffs(xl) - 1) >> 3)); open = false; }
                }
                R1 = __shfl_sync(0xFFFFFFFFu, len, 0);
                R2 = __shfl_sync(0xFFFFFFFFu, len, 16);
            }
            const uint32_t o0 = __shfl_sync(0xFFFFFFFFu, ow, s), o1 = __shfl_sync(0xFFFFFFFFu, ow, s + 1u), o2 = __shfl_sync(0xFFFFFFFFu, ow, s + 2u);
            if (!chain && (o0 >> 17) <= R1) { ml = R1; off = rep1; start = base + 1u; isRep = true; }
            else { ml = o0 >> 17; off = o0 & 0x1FFFFu; start = base; isRep = false; }
            bool moved = false;
#pragma unroll
            for (uint32_t d = 1; d <= 2u; d++) {
                if (d > lazyDepth || moved) break;
                const uint32_t q = base + d;
                if (q >= nh) break;
                const uint32_t rq = d == 1u ? R1 : R2, oq = d == 1u ? o1 : o2, wgt = d == 1u ? 3u : 4u;
                int32_t price = isRep ? 0 : static_cast<int32_t>(31 - __clz(off + 1u));
                if (rq >= 4u && static_cast<int32_t>(rq * wgt) > static_cast<int32_t>(ml * wgt) - price + 1) {
                    ml = rq; off = rep1; start = q; isRep = true; price = 0;
                }
                const uint32_t ol = oq >> 17, oo = oq & 0x1FFFFu;
                if (ol && static_cast<int32_t>(ol * 4u) - static_cast<int32_t>(31 - __clz(oo + 1u)) >
                              static_cast<int32_t>(ml * 4u) - price + (d == 1u ? 4 : 7)) {
                    s += d; moved = true;
                }
            }
            chain = true;
            if (!moved) done = true;
        }
        if (defer) { ip = dec0; chain = false; break; }
        if (!done) { ip = sBase + s; spanValid = false; continue; }            // same decision, span moved to its base
        chain = false;
        if (!isRep) {
            if (ml >= extCap) ml = coop_len(in, start, off, ml, n, lane);          // cut by the cap: extend
            // catch up to the left.  Usually the byte before the match already differs (the search at start - 1 would have
            // found the longer match): one uniform byte compare settles that before the warp-wide test
            bool grow = start > anchor && start > off;
            if (grow) {
                const uint32_t x = start - 1u, y = start - 1u - off;
                grow = ((ldsc32(in + (x & ~3u)) >> ((x & 3u) * 8u)) & 0xFFu) == ((ldsc32(in + (y & ~3u)) >> ((y & 3u) * 8u)) & 0xFFu);
            }
            if (grow) {
                // lane k tests byte start - k (k = 1..31)
                const bool ok = lane >= 1u && start >= anchor + lane && start >= off + lane &&
                                ((ldsc32(in + ((start - lane) & ~3u)) >> (((start - lane) & 3u) * 8u)) & 0xFFu) ==
                                ((ldsc32(in + ((start - lane - off) & ~3u)) >> (((start - lane - off) & 3u) * 8u)) & 0xFFu);
                const uint32_t k = __ffs(~(__ballot_sync(0xFFFFFFFFu, ok) >> 1)) - 1;    // leading run of lanes 1, 2, ...
                start -= k; ml += k;
            }
            if (off != rep1) { rep2 = rep1; rep1 = off; }
        }
        if (start == anchor && st.openLen && st.openOff == off) st.openLen += ml;
        else rep_emit(st, out, lane, off, start - anchor, ml);
        ip = anchor = start + ml;
        // ---- the other repeated offset, right after the match
        while (ip < nh) {
            if (rep2 == 0u || ip < rep2) break;
            if (ld32u(in, ip) != ld32u(in, ip - rep2)) break;           // uniform
            const uint32_t m2 = coop_len(in, ip, rep2, 4u, n, lane);
            { const uint32_t t = rep2; rep2 = rep1; rep1 = t; }
            rep_emit(st, out, lane, rep1, 0u, m2);
            ip = anchor = ip + m2;
        }
    }
    st.ip = ip; st.anchor = anchor; st.rep1 = rep1; st.rep2 = rep2;
}

// ------------------------------------------------------------------------------------------
// the persistent kernel
// ------------------------------------------------------------------------------------------
// Hits that chance alone produces in the repeated-key bitmap for nh insertions, plus six standard deviations:
// M (x - 1 + exp(-x)), x = nh / M, by its series in integer arithmetic (oracle/seqmodel.c:seqmodel_chance_threshold).
__device__ __forceinline__ uint32_t chance_threshold(uint32_t nh)
{
    const unsigned long long M = kBitmapBits, x = nh;
    const unsigned long long a = x * x / M;
    const uint32_t e = static_cast<uint32_t>(a / 2 - a * x / (6 * M) + a * a / (24 * M));
    uint32_t r = 0;
    for (uint32_t bit = 1u << 15; bit; bit >>= 1) { const uint32_t t = r | bit; if (static_cast<unsigned long long>(t) * t <= e) r = t; }
    return e + 6u * r + 24u;
}

template <bool kFast>
__global__ void __launch_bounds__(kThreads, 1) lz77_parse_kernel(const ParseParams P)
{
    extern __shared__ __align__(128) uint8_t smem[];
    Shared S;
    {
        uint32_t p;
        asm volatile("mov.u32 %0, %1;" : "=r"(p) : "r"(smem_u32(smem)));   // opaque: one register, never re-derived
        S.in = p;    p += kSmemInput;
        S.tab = p;   p += kSmemTabL;
        S.bitmap = p; S.ringT = p; p += kSmemTabS;   // the bitmap of phase A runs from here through the rings; afterwards the tag ring lives here
        S.ringH = p; p += kSmemRingH;
        S.ringC = p; p += kSmemRingC;
        S.ringL = p; p += kSmemRingL;
        S.gmax = p;  S.scan = p; p += kRingC * 64 * 4;      // phase A borrows the group words (rewritten before they are read)
        S.gown = p;  S.hits = p; p += kRingC * 64 * 4;
        S.hasA = p;  p += 2 * 64 * 4;
        S.entA = p;  p += 2 * 64 * 4;
        S.mbar = p;  p += kTmaChunks * 8;
        S.work = p;  p += 8;
        S.task = p;  p += 8;
        S.curVal = p; p += 4;
        S.curTag = p; p += 4;
        S.ecVal = p;  p += 12;
        S.ecTag = p;  p += 4;
        S.emTag = p;
    }
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    // Role of this warp: 0 = hash/extend pool, 1 = bucket counters (T), 2 = spare, 3 / 4 = entries of the first /
    // second half window (P1), 5 / 6 = emit of the first / second half window (P2).
    // Which warp plays which role.  Levels 5-12: the table warp and the parse warp sit on one scheduler (warps 3, 7 of
    // sub-partition 3; the idle entry/emit warps take 11, 15, 19) and share it with three pool warps instead of
    // six or seven: the parse warp is the longest role there and an issue slot lost to a pool warp is lost time
    // (parse warp 118 K -> 105 K cycles per stage).  Levels 1-4 gain nothing from it (the pool becomes the longest role).
    uint32_t vw = warp;
    if (!kFast)
        vw = warp == 3u ? 26u : warp == 26u ? 3u :
             warp == 7u ? 28u : warp == 11u ? 29u : warp == 15u ? 30u : warp == 19u ? 31u :
             warp == 28u ? 7u : warp == 29u ? 11u : warp == 30u ? 15u : warp == 31u ? 19u : warp;
    const uint32_t role = vw < kEhWarps ? 0u : vw - kEhWarps + 1u;
    const uint32_t poolIdx = vw;
    uint32_t *sorted = P.sorted + static_cast<size_t>(blockIdx.x) * kSortedCap;

    if (tid == 0) {
        for (uint32_t c = 0; c < kTmaChunks; c++) mbar_init(S.mbar + c * 8u, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    uint32_t tmaParity = 0;            // bit c = phase parity mbarrier c completes next

    for (;;) {
        __syncthreads();                       // previous block fully retired; mbarriers initialised
        if (tid == 0) sts32(S.work, atomicAdd(P.workCounter, 1u));
        __syncthreads();
        const uint32_t b = lds32(S.work);
        if (b >= P.nBlocks) break;

        uint32_t n;
        if (P.sizes) n = P.sizes[b];
        else {
            const uint64_t start = static_cast<uint64_t>(b) * P.stride;
            const uint64_t left = P.totalSize > start ? P.totalSize - start : 0;
            n = left < P.blockSize ? static_cast<uint32_t>(left) : P.blockSize;
        }
        if (n > kBlockMax) n = kBlockMax;
        const uint8_t *gsrc = P.src + static_cast<uint64_t>(b) * P.stride;
        uint4 *out = P.seqs + static_cast<uint64_t>(b) * P.seqStride;
        const uint32_t bulk = n & ~15u;
        const uint32_t nChunks = (bulk + kTmaChunk - 1) / kTmaChunk;
        const uint32_t nh = n >= 8 ? n - 7 : 0;

        // ---- stage the block: TMA bulk copies (one elected thread) + ragged tail; clear histogram and bitmap
        if (tid == 0) {
            fence_proxy_async();               // earlier generic-proxy reads of the buffer are done
            for (uint32_t c = 0; c < nChunks; c++) {
                const uint32_t bytes = min(kTmaChunk, bulk - c * kTmaChunk);
                const uint32_t bar = S.mbar + c * 8u;
                mbar_expect_tx(bar, bytes);
                tma_load_1d(S.in + c * kTmaChunk, gsrc + c * kTmaChunk, bytes, bar);
            }
            sts32(S.hits, 0u);
            sts32(S.hits + 4u, chance_threshold(nh));
        }
        if (tid >= 32 && tid < 32 + (n - bulk))
            sts8(S.in + bulk + tid - 32, gsrc[bulk + tid - 32]);
        for (uint32_t i = tid; i < (kSmemTabL + kBitmapBits / 8) / 16; i += kThreads)    // the table and the bitmap are contiguous
            sts128(S.tab + i * 16u, 0u);
        if (tid < 2) sts32(S.task + tid * 4u, kEhWarps);
        if (tid == 2) { sts32(S.curTag, 0u); sts32(S.ecTag, 0u); sts32(S.emTag, 0u); }
        for (uint32_t c = 0; c < nChunks; c++) mbar_wait(S.mbar + c * 8u, (tmaParity >> c) & 1u);
        tmaParity ^= (1u << nChunks) - 1u;     // only the barriers armed for this block changed phase
        __syncthreads();

        // ---- phase A: histogram of the key buckets + repeated-key bitmap, all warps, any order
        {
            uint32_t hits = 0;
            for (uint32_t g = warp; g * 32u < nh; g += kNumWarps) {
                const uint32_t p = g * 32u + lane;
                const bool valid = p < nh;
                const uint32_t a = S.in + (p & ~3u), sh = (p & 3u) * 8u;
                const uint32_t w0 = lds32(a), w1 = lds32(a + 4u), w2 = lds32(a + 8u);
                const uint32_t v = key_hash(__funnelshift_r(w0, w1, sh), __funnelshift_r(w1, w2, sh), P.keyMask);
                const uint32_t bkt = v >> (32 - kBucketBits);
                // one reduction per lane: lanes of the same bucket are serialised by the unit, which costs less than
                // pairing them up first (MATCH.ANY)
                if (valid) asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(S.tab + bkt * 4u), "r"(1u) : "memory");
                const uint32_t bi = __umulhi(v, kBitmapBits);
                uint32_t old = 0;
                if (valid) asm volatile("atom.shared.or.b32 %0, [%1], %2;" : "=r"(old) : "r"(S.bitmap + (bi >> 5) * 4u), "r"(1u << (bi & 31u)) : "memory");
                hits += __popc(__ballot_sync(0xFFFFFFFFu, valid && ((old >> (bi & 31u)) & 1u)));
            }
            if (lane == 0 && hits) asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(S.hits), "r"(hits) : "memory");
        }
        __syncthreads();
        if (lds32(S.hits) < lds32(S.hits + 4u)) {
            // no more repeated keys than chance: one literal run (/root/reference/src/qatseqprod.c:1308-1313)
            if (tid == 0) { out[0] = make_uint4(0u, n, 0u, 0u); P.counts[b] = 1u; }
            continue;
        }
        // ---- bucket segment starts: exclusive scan of the histogram, in units of kSegAlign entries
        {
            constexpr uint32_t kBins = (kBuckets + kThreads - 1) / kThreads;       // consecutive buckets per thread
            uint32_t seg[kBins], sum = 0;
#pragma unroll
            for (uint32_t i = 0; i < kBins; i++) {
                seg[i] = sum;
                if (tid * kBins + i < kBuckets) sum += (lds32(S.tab + (tid * kBins + i) * 4u) + kSegAlign - 1u) / kSegAlign;
            }
            uint32_t incl = sum;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t o = __shfl_up_sync(0xFFFFFFFFu, incl, d);
                if (lane >= static_cast<uint32_t>(d)) incl += o;
            }
            if (lane == 31) sts32(S.scan + warp * 4u, incl);
            __syncthreads();
            uint32_t base = lane < warp ? lds32(S.scan + lane * 4u) : 0u;
#pragma unroll
            for (int d = 16; d; d >>= 1) base += __shfl_xor_sync(0xFFFFFFFFu, base, d);
            base += incl - sum;
#pragma unroll
            for (uint32_t i = 0; i < kBins; i++)
                if (tid * kBins + i < kBuckets) sts32(S.tab + (tid * kBins + i) * 4u, (base + seg[i]) << 17);
        }
        __syncthreads();

        const uint32_t nW = (n + kWindow - 1) / kWindow;
        EmitCarry ec = {0, 0, 0};              // P2 (loaded from / published to shared memory every half window)
        RepState rs = {0, 0, 0, 0, 0, 0, 0, 0};  // R (levels 5-12)

#ifdef B200SP_ROLE_PROFILE
        unsigned long long busy = 0, blockStart = clock64();
#endif
        for (uint32_t t = 0; t < nW + 4; t++) {
#ifdef B200SP_ROLE_PROFILE
            const unsigned long long c0 = clock64();
#endif
            if (role == 0u) {
                // One task queue per stage.  Task g scans and extends group g of window t-2 and, behind its first
                // loads, hashes group g of window t.
                const uint32_t nE = (t >= 2 && t - 2 < nW) ? kGroups : 0u;
                // hash-only tasks exist only while there is no extension work yet (the first two stages)
                const uint32_t nAll = nE ? nE : (t < nW ? kGroups / kHashGroups : 0u);
                const uint32_t ctr = S.task + (t & 1u) * 4u;
                // the first task of every pool warp is its own index (the counter starts at kEhWarps); only the
                // later ones cost an atomic
                for (uint32_t id = poolIdx; id < nAll; id = pop_task(ctr, lane)) {
                    if (id < nE) {
                        const uint32_t wdx = t - 2;
                        // past the last window the fused hash runs with no valid position (harmless ring writes)
                        if (kFast) stage_extend_fast<true>(S, wdx, id, lane, wdx * kWindow + id * 32u + lane, n, nh, P.minMatch, P.extCap,
                                                           P.scan, P.keyMask, sorted, t, t < nW ? nh : 0u);
                        else stage_extend_full<true>(S, wdx, id, lane, wdx * kWindow + id * 32u + lane, n, nh, P.minMatch, P.extCap,
                                                     P.scan, P.keyMask, sorted, t, t < nW ? nh : 0u);
                    } else {
                        const HashState hs = hash_begin(S, t, id, lane, nh, P.keyMask);
                        hash_finish(S, t, id, lane, nh, hs);
                    }
                }
            } else if (role == 1u) {
                if (t >= 1 && t - 1 < nW) stage_table(S, t - 1, lane, sorted);
            } else if (role == 2u) {
                // spare warp
            } else if (role <= 4u) {
                if (role == 3u && lane == 0) sts32(S.task + ((t + 1u) & 1u) * 4u, kEhWarps);   // next stage's queue (nobody touches it now)
                if (kFast) { if (t >= 3 && t - 3 < nW) stage_entries(S, t - 3, role - 3u, lane, P.minMatch, P.lazyDepth, P.errorFlag); }
                else if (role == 3u && t >= 3 && t - 3 < nW) stage_rep_parse(S, t - 3, lane, n, nh, P.lazyDepth, P.extCap, rs, out);
            } else {
                if (kFast && t >= 4) stage_emit(S, t - 4, role - 5u, lane, ec, out, P.errorFlag);
            }
#ifdef B200SP_ROLE_PROFILE
            busy += clock64() - c0;
#endif
            __syncthreads();
        }

#ifdef B200SP_ROLE_PROFILE          // developer builds only (tools/ab_build.sh NAME -DB200SP_ROLE_PROFILE)
        if (P.roleCycles && lane == 0) {
            atomicAdd(&P.roleCycles[role], busy);
            if (role == 6u) { atomicAdd(&P.roleCycles[7], clock64() - blockStart); atomicAdd(&P.roleCycles[8], (unsigned long long)(nW + 4)); }
        }
#endif
        if (kFast && role == 6u && lane == 0) {       // the second half's emit warp holds the carry after the last window
            out[ec.nOut] = make_uint4(0u, n - ec.anchor, 0u, 0u);   // trailing literals / block delimiter
            P.counts[b] = ec.nOut + 1u;
        }
        if (!kFast && role == 3u && lane == 0) {
            if (rs.openLen) out[rs.nOut++] = make_uint4(rs.openOff, rs.openLit, rs.openLen, 0u);
            out[rs.nOut] = make_uint4(0u, n - rs.anchor, 0u, 0u);
            P.counts[b] = rs.nOut + 1u;
        }
    }
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
bool params_for_level(int level, ParseParams &p)
{
    // One parameter class per zstd strategy class; the scan width is the level-scaled search
    // depth (the reference hands the level to its engine, /root/reference/src/qatseqprod.c:1154, and rebuilds the
    // session when it changes, :1193-1201).  Must equal oracle/seqmodel.c:seqmodel_params_for_level.
    static const uint32_t scanOf[13] = { 0, 4, 4, 4, 8, 16, 24, 32, 32, 48, 64, 384, 768 };
    if (level < 1 || level > 12) return false;
    p.keyMask = level <= 2 ? 0xFFFFu : level <= 4 ? 0xFFu : 0u;   // 6-byte keys at levels 1-2, 5-byte keys at 3-4, 4-byte keys from greedy up
    p.rank16 = level <= 4 ? 1u : 0u;             // fast classes rank candidates on 16 bytes and extend the winner; the others measure every candidate in full and parse repcode-aware
    p.scan = scanOf[level];
    p.minMatch = 4;
    p.extCap = kMaxExtCap;
    p.lazyDepth = 2;
    return true;
}

cudaError_t configure_kernels()
{
    cudaError_t ce = cudaFuncSetAttribute(lz77_parse_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kSmemTotal));
    if (ce == cudaSuccess) ce = cudaFuncSetAttribute(lz77_parse_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kSmemTotal));
    return ce;
}

cudaError_t launch_parse(const ParseParams &p, int numSMs, cudaStream_t stream)
{
    if (p.nBlocks == 0) return cudaSuccess;
    const unsigned grid = static_cast<unsigned>(p.nBlocks < static_cast<uint32_t>(numSMs) ? p.nBlocks : numSMs);
    if (p.rank16) lz77_parse_kernel<true><<<grid, kThreads, kSmemTotal, stream>>>(p);
    else lz77_parse_kernel<false><<<grid, kThreads, kSmemTotal, stream>>>(p);
    return cudaGetLastError();
}

}  // namespace b200sp
