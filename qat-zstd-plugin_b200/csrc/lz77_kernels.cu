/*
 * lz77_kernels.cu — sm_100a LZ77 block parser: batch of independent <=128 KiB blocks in HBM ->
 * ZSTD_Sequence arrays in HBM.
 *
 * This is the B200 replacement for the QAT LZ4s engine + QZSTD_decLz4s
 * (/root/reference/src/qatseqprod.c:1245-1249 submit, :1013-1091 token walk, :1308-1313
 * incompressible shortcut).  Output convention is the reference's: matchLength >= 3 for real
 * matches, the last entry is {offset 0, trailing literals, matchLength 0}, `rep` is 0.
 *
 * One persistent CTA per SM; a CTA owns one block at a time:
 *   - the block is staged into shared memory with 1-D TMA bulk copies (UBLKCP), 16 KiB per
 *     mbarrier so the pipeline starts before the whole block has landed;
 *   - both hash tables (2 x 16 Ki x u16, positions stored >> 1) live in shared memory;
 *   - the block flows through a 4-deep ring of 1024-position windows, one role per stage:
 *       H  hash      16 warps   8-byte + short hash per position, intra-warp duplicate links
 *       T  table      2 warps   one warp per table walks the window in order: read slot,
 *                               overwrite with the newer position (exact serial semantics)
 *       E  extend    16 warps   probe the candidates (4-byte word compares), warp-cooperative
 *                               long extension, warp prefix-max of match ends
 *       P  parse      1 warp    lane-parallel speculative greedy/lazy parse (iterated to the
 *                               serial fixed point), sequence compaction, 16 B stores
 *     (the H and E roles share the same 16 warps).
 * The result is bit-identical to oracle/seqmodel.c, which is the serial statement of the same
 * four steps.  Integer/indexing work only: no tensor cores, no TMEM.
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
__device__ __forceinline__ void tma_load_1d(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar)
{
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
        ::"r"(dst), "l"(src), "r"(bytes), "r"(bar)
        : "memory");
}

__device__ __forceinline__ void fence_proxy_async()
{
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// Unaligned little-endian 32-bit read from the staged block.
__device__ __forceinline__ uint32_t ld32u(const uint32_t *in32, uint32_t bytePos)
{
    const uint32_t w = bytePos >> 2;
    return __funnelshift_r(in32[w], in32[w + 1], (bytePos & 3u) * 8u);
}

// Warp-uniform pop from a shared-memory task counter: one ATOMS by lane 0, one broadcast.
__device__ __forceinline__ uint32_t pop_task(uint32_t ctrAddr, uint32_t lane)
{
    uint32_t id = 0;
    if (lane == 0) asm volatile("atom.shared.add.u32 %0, [%1], 1;" : "=r"(id) : "r"(ctrAddr) : "memory");
    return __shfl_sync(0xFFFFFFFFu, id, 0);
}

__device__ __forceinline__ uint32_t ring_index(uint32_t group, uint32_t lane)
{
    return group * 32u + (lane ^ group);   // XOR swizzle: conflict-free by group and by lane
}

__device__ __forceinline__ int32_t gain_of(uint32_t len, uint32_t off)
{
    return static_cast<int32_t>(len * 4u) - static_cast<int32_t>(31 - __clz(off + 1u));
}

struct Shared {
    uint32_t *in32;
    uint16_t *tabL;
    uint16_t *tabS;
    uint32_t *ring0;    // [kRing][kWindow]
    uint32_t *ring1;    // [kRing][kWindow]
    uint32_t *gmax;     // [kRing][kGroups] packed farthest-reaching match of each group
    uint64_t *mbar;     // [kTmaChunks]
    volatile int *work; // next block index
    unsigned int *task; // [2][2] per-stage task counters of the hash/extend warps (double-buffered)
};

// ------------------------------------------------------------------------------------------
// H: hashes of one 32-position group -> ring words {hash | delta << 16 | last << 21 | valid << 22}
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void stage_hash(const Shared &S, uint32_t slot, uint32_t group, uint32_t lane,
                                           uint32_t p, uint32_t nh, uint32_t shortMask)
{
    const bool valid = p < nh;
    uint32_t hL = 0, hS = 0;
    if (valid) {
        const uint32_t w = p >> 2, sh = (p & 3u) * 8u;
        const uint32_t w0 = S.in32[w], w1 = S.in32[w + 1], w2 = S.in32[w + 2];
        const uint32_t lo = __funnelshift_r(w0, w1, sh);
        const uint32_t hi = __funnelshift_r(w1, w2, sh);
        hL = (lo * 0x9E3779B1u + hi * 0x85EBCA77u) >> (32 - kLongBits);
        hS = (lo * 0x9E3779B1u + (hi & shortMask) * 0xC2B2AE3Du) >> (32 - kShortBits);
    }
    const uint32_t ltMask = (1u << lane) - 1u;
    const uint32_t geMask = ~((2u << lane) - 1u);      // lanes strictly above
    // invalid lanes get unique keys so they never link
    const uint32_t mL = __match_any_sync(0xFFFFFFFFu, valid ? hL : (0x10000u | lane));
    const uint32_t mS = __match_any_sync(0xFFFFFFFFu, valid ? hS : (0x10000u | lane));
    const uint32_t bL = mL & ltMask, bS = mS & ltMask;
    const uint32_t dL = bL ? lane - (31u - __clz(bL)) : 0u;   // distance to the nearest earlier lane with the same hash
    const uint32_t dS = bS ? lane - (31u - __clz(bS)) : 0u;
    const uint32_t lastL = (mL & geMask) == 0u, lastS = (mS & geMask) == 0u;
    const uint32_t idx = slot * kWindow + ring_index(group, lane);
    S.ring0[idx] = hL | (dL << 16) | (lastL << 21) | (static_cast<uint32_t>(valid) << 22);
    S.ring1[idx] = hS | (dS << 16) | (lastS << 21) | (static_cast<uint32_t>(valid) << 22);
}

// ------------------------------------------------------------------------------------------
// T: one warp walks one table over a window, group by group, in position order.
// ring word in: hash/links from H; ring word out: candidate (position >> 1) or 0xFFFF.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void stage_table(uint32_t *ring, uint16_t *tab, uint32_t slot, uint32_t lane,
                                            uint32_t windowBase)
{
    uint32_t *r = ring + slot * kWindow;
#pragma unroll 1
    for (uint32_t g0 = 0; g0 < kGroups; g0 += 8) {
        uint32_t w[8], tv[8];
#pragma unroll
        for (int k = 0; k < 8; k++) w[k] = r[ring_index(g0 + k, lane)];
#pragma unroll
        for (int k = 0; k < 8; k++) {
            const uint32_t h = w[k] & 0xFFFFu;
            const bool valid = (w[k] >> 22) & 1u;
            const bool last = (w[k] >> 21) & 1u;
            const uint32_t p = windowBase + (g0 + k) * 32u + lane;
            tv[k] = tab[h];
            if (valid && last) tab[h] = static_cast<uint16_t>(p >> 1);
            __syncwarp();       // orders this group's stores before the next group's loads
        }
#pragma unroll
        for (int k = 0; k < 8; k++) {
            const uint32_t delta = (w[k] >> 16) & 31u;
            const bool valid = (w[k] >> 22) & 1u;
            const uint32_t p = windowBase + (g0 + k) * 32u + lane;
            uint32_t cand = delta ? ((p - delta) >> 1) : tv[k];
            if (!valid) cand = 0xFFFFu;
            r[ring_index(g0 + k, lane)] = cand;
        }
    }
}

// ------------------------------------------------------------------------------------------
// E: candidates of one group -> best match per position -> prefix-max of match ends
// ring out: ring0 = end (p + len, 0 if none so far in this group), ring1 = offset
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t first_diff_16(uint32_t x1, uint32_t x2, uint32_t x3)
{
    // length of the common prefix of two 16-byte strings whose first 4 bytes are equal,
    // given the XOR of words 1..3 (branch-free)
    uint32_t len = 16u;
    if (x3) len = 12u + ((__ffs(x3) - 1) >> 3);
    if (x2) len = 8u + ((__ffs(x2) - 1) >> 3);
    if (x1) len = 4u + ((__ffs(x1) - 1) >> 3);
    return len;
}

__device__ __forceinline__ void stage_extend(const Shared &S, uint32_t slot, uint32_t group, uint32_t lane,
                                             uint32_t p, uint32_t n, uint32_t nh, uint32_t minMatch,
                                             uint32_t extCap)
{
    const uint32_t idx = slot * kWindow + ring_index(group, lane);
    const uint32_t cL = S.ring0[idx], cS = S.ring1[idx];
    const uint32_t *in32 = S.in32;
    const bool valid = p < nh;
    uint32_t bestLen = 0, bestOff = 0;
    const uint32_t lim = valid ? min(n - p, extCap) : 0u;
    const uint32_t probe = min(lim, kProbe);
    // our own first 16 bytes, as four unaligned words (reads stay inside the padded buffer)
    uint32_t a0, a1, a2, a3;
    {
        const uint32_t w = min(p, kBlockMax) >> 2, sh = (p & 3u) * 8u;
        const uint32_t x0 = in32[w], x1 = in32[w + 1], x2 = in32[w + 2], x3 = in32[w + 3], x4 = in32[w + 4];
        a0 = __funnelshift_r(x0, x1, sh); a1 = __funnelshift_r(x1, x2, sh);
        a2 = __funnelshift_r(x2, x3, sh); a3 = __funnelshift_r(x3, x4, sh);
    }
#pragma unroll
    for (int t = 0; t < 2; t++) {
        const uint32_t c = t ? cS : cL;
        // the slot stands for positions 2c and 2c+1 (same 32-bit word row): pick the one whose first
        // 4 bytes equal ours, the nearer one if both do.  0xFFFF (empty) fails q0 < p by construction.
        const uint32_t q0 = 2u * c;
        const uint32_t w = min(q0, kBlockMax) >> 2, sh0 = (q0 & 3u) * 8u;       // q0 even: sh0 is 0 or 16
        const uint32_t y0 = in32[w], y1 = in32[w + 1], y2 = in32[w + 2], y3 = in32[w + 3], y4 = in32[w + 4];
        const bool c0 = valid && q0 < p && __funnelshift_r(y0, y1, sh0) == a0;
        const bool c1 = valid && q0 + 1u < p && __funnelshift_r(y0, y1, sh0 + 8u) == a0;
        const uint32_t sh = sh0 + (c1 ? 8u : 0u);
        const uint32_t b1 = __funnelshift_r(y1, y2, sh), b2 = __funnelshift_r(y2, y3, sh), b3 = __funnelshift_r(y3, y4, sh);
        uint32_t ml = min(first_diff_16(a1 ^ b1, a2 ^ b2, a3 ^ b3), probe);
        const uint32_t off = p - q0 - (c1 ? 1u : 0u);
        if (!(c0 || c1)) ml = 0;
        if (ml > bestLen || (ml == bestLen && ml > 0 && off < bestOff)) { bestLen = ml; bestOff = off; }
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
            if (k < reach) x = ld32u(in32, ph + k) ^ ld32u(in32, qh + k);
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
    // Pack {end relative to the group start (9 bits), 31 - lane (6 bits), offset (17 bits)}: an unsigned
    // max over packed words picks the farthest-reaching match and, on ties, the older one.
    uint32_t pk = 0;
    if (bestLen >= minMatch) pk = ((lane + bestLen) << 23) | ((31u - lane) << 17) | bestOff;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t o = __shfl_up_sync(0xFFFFFFFFu, pk, d);
        if (lane >= static_cast<uint32_t>(d)) pk = max(pk, o);
    }
    S.ring0[idx] = pk;                                   // prefix-max within the group
    if (lane == 31) S.gmax[slot * kGroups + group] = pk; // the group's total, for the carry of later groups
}

// ------------------------------------------------------------------------------------------
// J: jump links of one group.  Folds the carry of the previous 8 groups (a match is at most
// extCap = 256 bytes long, so nothing older can reach in) into the group's prefix-max, applies the
// lazy rule, and leaves one word per position: {target - windowBase (11 bits), kind (2), offset (17)}.
//   SKIP  no usable match here: go to the next position of the group that has one (or to its end)
//   HOP   a later start is better (lazy): go to p+1 / p+2
//   TAKE  emit the match [p, target) at `offset`
// ------------------------------------------------------------------------------------------
constexpr uint32_t kKindSkip = 0, kKindHop = 1, kKindTake = 2;

__device__ __forceinline__ void stage_jump(const Shared &S, uint32_t wdx, uint32_t group, uint32_t lane,
                                           uint32_t minMatch, uint32_t lazyDepth)
{
    const uint32_t slot = wdx & (kRing - 1);
    const uint32_t idx = slot * kWindow + ring_index(group, lane);
    // ---- carry: lane k-1 looks k groups back (possibly into the previous window's slot)
    uint32_t c = 0;
    if (lane < 8) {
        const uint32_t k = lane + 1;
        int gg = static_cast<int>(group) - static_cast<int>(k);
        uint32_t sl = slot;
        bool ok = true;
        if (gg < 0) { ok = wdx > 0; gg += kGroups; sl = (wdx - 1) & (kRing - 1); }
        if (ok) {
            const uint32_t v = S.gmax[sl * kGroups + gg];
            const uint32_t rel = v >> 23;
            if (rel > 32u * k) c = ((rel - 32u * k) << 23) | ((32u + k) << 17) | (v & 0x1FFFFu);
        }
    }
    c = __reduce_max_sync(0xFFFFFFFFu, c);
    const uint32_t bt = max(S.ring0[idx], c);              // B(p): farthest-reaching match covering p
    const uint32_t rel = bt >> 23, off = bt & 0x1FFFFu;
    const bool has = rel >= lane + minMatch;
    const int32_t gain = static_cast<int32_t>((rel - lane) * 4u) - static_cast<int32_t>(31 - __clz(off + 1u));

    // ---- B(p+1), B(p+2) by shuffle.  The look-ahead never leaves the group (model: window = 32).
    const uint32_t b1 = __shfl_down_sync(0xFFFFFFFFu, bt, 1), b2 = __shfl_down_sync(0xFFFFFFFFu, bt, 2);
    const uint32_t l1 = lane + 1, l2 = lane + 2;
    const bool ok1 = lane < 31 && (b1 >> 23) >= l1 + minMatch;
    const bool ok2 = lane < 30 && (b2 >> 23) >= l2 + minMatch;
    const int32_t gain1 = static_cast<int32_t>(((b1 >> 23) - l1) * 4u) - static_cast<int32_t>(31 - __clz((b1 & 0x1FFFFu) + 1u));
    const int32_t gain2 = static_cast<int32_t>(((b2 >> 23) - l2) * 4u) - static_cast<int32_t>(31 - __clz((b2 & 0x1FFFFu) + 1u));
    uint32_t adv = 0;
    if (has && lazyDepth >= 1 && ok1) {
        if (gain1 > gain + 4) adv = 1;
        else if (lazyDepth >= 2 && ok2 && gain2 > gain + 7) adv = 2;
    }
    const uint32_t hasMask = __ballot_sync(0xFFFFFFFFu, has);
    const uint32_t groupRel = group * 32u;                  // group start relative to the window
    uint32_t tgt, kind;
    if (has) {
        kind = adv ? kKindHop : kKindTake;
        tgt = groupRel + (adv ? lane + adv : rel);
    } else {
        const uint32_t m = hasMask & ~((2u << lane) - 1u);
        kind = kKindSkip;
        tgt = groupRel + (m ? __ffs(m) - 1 : 32u);
    }
    S.ring0[idx] = tgt | (kind << 11) | (off << 13);

    // ---- path summaries by pointer doubling.  For a cursor standing on this position: where the
    // walk leaves the group, how many matches it takes, how many of them absorb their successor
    // (zero literals, same offset), and which lanes hold the first / last taken match.
    //   word = next (9 bits, >= 32: left the group, relative to the group start) | taken (4) |
    //          absorbed (4) | first lane (6, 32 = none) | last lane (6, 32 = none)
    const uint32_t nxt = tgt - groupRel;                              // 1 .. 32 + 256
    const bool take = kind == kKindTake;
    uint32_t absorb = 0;
    {
        const uint32_t succ = __shfl_sync(0xFFFFFFFFu, S.ring0[idx], nxt & 31u);   // link word of the node at our target
        absorb = take && nxt < 32u && ((succ >> 11) & 3u) == kKindTake && (succ >> 13) == off;
    }
    uint32_t pk = nxt | ((take ? 1u : 0u) << 9) | (absorb << 13) | ((take ? lane : 32u) << 17) | ((take ? lane : 32u) << 23);
#pragma unroll 1
    for (int round = 0; round < 5; round++) {
        const uint32_t cur = pk & 0x1FFu;
        if (__all_sync(0xFFFFFFFFu, cur >= 32u)) break;
        const uint32_t o = __shfl_sync(0xFFFFFFFFu, pk, cur & 31u);
        if (cur < 32u) {
            const uint32_t ft = (pk >> 17) & 63u, olt = (o >> 23) & 63u;
            const uint32_t nft = ft < 32u ? ft : (o >> 17) & 63u;
            const uint32_t nlt = olt < 32u ? olt : (pk >> 23) & 63u;
            const uint32_t cnts = ((pk >> 9) & 0xFFu) + ((o >> 9) & 0xFFu);   // taken | absorbed << 4, no carry: <= 8 each
            pk = (o & 0x1FFu) | (cnts << 9) | (nft << 17) | (nlt << 23);
        }
    }
    S.ring1[idx] = pk;
}

// ------------------------------------------------------------------------------------------
// P: parse one window.  Lane j owns positions [base + 32 j, base + 32 j + 32).
// ------------------------------------------------------------------------------------------
struct ParseCarry {          // uniform across the warp, carried from window to window
    uint32_t cursor;         // first position the parser has not consumed yet
    uint32_t anchor;         // end of the last emitted match
    uint32_t prevOff;        // offset of the last emitted match
    uint32_t nOut;           // sequences written so far
};

// Emits the matches of the lane's segment (= one group) by chasing the jump links from `entry`.
// New sequences go to out[outIdx...]; a leading continuation of the previous sequence is returned
// (the caller adds it to out[firstIdx - 1].matchLength).
__device__ __forceinline__ uint32_t lane_emit(const uint32_t *links, uint32_t base, uint32_t group, uint32_t entry,
                                              uint32_t anchor, uint32_t prevOff, uint4 *out, uint32_t outIdx)
{
    const uint32_t segStart = base + group * 32u, segEnd = segStart + 32u;
    uint32_t c = entry, headAdd = 0;
    uint32_t openOff = 0, openLit = 0, openLen = 0;   // sequence being accumulated
    bool haveOpen = false;
    while (c < segEnd) {
        const uint32_t w = links[ring_index(group, c - segStart)];
        const uint32_t tgt = base + (w & 0x7FFu);
        if (((w >> 11) & 3u) == kKindTake) {
            const uint32_t o = w >> 13, lit = c - anchor, len = tgt - c;
            if (lit == 0 && o == prevOff && anchor > 0) {
                if (haveOpen) openLen += len; else headAdd += len;
            } else {
                if (haveOpen) out[outIdx++] = make_uint4(openOff, openLit, openLen, 0u);
                openOff = o; openLit = lit; openLen = len; haveOpen = true;
            }
            anchor = tgt; prevOff = o;
        }
        c = tgt;
    }
    if (haveOpen) out[outIdx] = make_uint4(openOff, openLit, openLen, 0u);
    return headAdd;
}

__device__ __forceinline__ void stage_parse(const Shared &S, uint32_t slot, uint32_t lane, uint32_t base,
                                            ParseCarry &pc, uint4 *out)
{
    const uint32_t *links = S.ring0 + slot * kWindow;
    const uint32_t *paths = S.ring1 + slot * kWindow;
    const uint32_t segStart = base + lane * 32u, segEnd = segStart + 32u;

    // ---- every lane guesses that the parser enters its segment at the segment start, then the
    // guesses are corrected from lane 0 upward until nothing changes: each correction is one
    // table look-up (the exit of the walk from any entry was tabulated by the J stage).
    // A lane whose entry lies beyond its segment is passed over (a long match covers it); the
    // entry of a lane is the exit of the nearest earlier lane that is NOT passed over, i.e. the
    // prefix maximum over live lanes — so a run of covered lanes costs one round, not one each.
    uint32_t entry = lane == 0 ? max(pc.cursor, base) : segStart;
    uint32_t path = 0, exitPos = 0;
    for (;;) {
        const bool live = entry < segEnd;
        path = live ? paths[ring_index(lane, entry - segStart)] : 0u;
        exitPos = live ? segStart + (path & 0x1FFu) : 0u;          // passed-over lanes contribute nothing
        uint32_t pm = lane == 0 ? max(exitPos, entry) : exitPos;   // lane 0 also carries the window's entry cursor
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t o = __shfl_up_sync(0xFFFFFFFFu, pm, d);
            if (lane >= static_cast<uint32_t>(d)) pm = max(pm, o);
        }
        const uint32_t prevMax = __shfl_up_sync(0xFFFFFFFFu, pm, 1);
        const uint32_t want = lane == 0 ? entry : max(prevMax, segStart);
        const bool changed = want != entry;
        entry = want;
        if (!__any_sync(0xFFFFFFFFu, changed)) { exitPos = pm; break; }   // pm: cursor after this lane's segment
    }
    const uint32_t cnt = (path >> 9) & 15u, merges = (path >> 13) & 15u;
    uint32_t firstPos = 0, firstOff = 0, lastEnd = 0, lastOff = 0;
    if (cnt) {
        const uint32_t fl = (path >> 17) & 63u, ll = (path >> 23) & 63u;
        const uint32_t wf = links[ring_index(lane, fl)], wl = links[ring_index(lane, ll)];
        firstPos = segStart + fl; firstOff = wf >> 13;
        lastEnd = base + (wl & 0x7FFu); lastOff = wl >> 13;
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
    if (lane == 0 || anchor == 0) { anchor = pc.anchor; prevOff = pc.prevOff; }

    // ---- output slots
    const bool headMerge = cnt && firstPos == anchor && firstOff == prevOff && anchor > 0;
    const uint32_t fresh = cnt - merges - (headMerge ? 1u : 0u);
    uint32_t incl = fresh;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t v = __shfl_up_sync(0xFFFFFFFFu, incl, d);
        if (lane >= static_cast<uint32_t>(d)) incl += v;
    }
    const uint32_t firstIdx = pc.nOut + incl - fresh;
    const uint32_t total = __shfl_sync(0xFFFFFFFFu, incl, 31);

    uint32_t headAdd = 0;
    if (cnt) headAdd = lane_emit(links, base, lane, entry, anchor, prevOff, out, firstIdx);
    __syncwarp();
    if (headAdd) atomicAdd(&out[firstIdx - 1].z, headAdd);   // continuation of an earlier lane's sequence
    __syncwarp();

    pc.cursor = max(__shfl_sync(0xFFFFFFFFu, exitPos, 31), base + kWindow);
    if (totE) { pc.anchor = totE; pc.prevOff = totO; }
    pc.nOut += total;
}

// ------------------------------------------------------------------------------------------
// the persistent kernel
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads, 1) lz77_parse_kernel(const ParseParams P)
{
    extern __shared__ __align__(128) uint8_t smem[];
    Shared S;
    {
        uint8_t *p = smem;
        S.in32 = reinterpret_cast<uint32_t *>(p);  p += kSmemInput;
        S.tabL = reinterpret_cast<uint16_t *>(p);  p += kSmemTabL;
        S.tabS = reinterpret_cast<uint16_t *>(p);  p += kSmemTabS;
        S.ring0 = reinterpret_cast<uint32_t *>(p); p += kSmemRing / 2;
        S.ring1 = reinterpret_cast<uint32_t *>(p); p += kSmemRing / 2;
        S.gmax = reinterpret_cast<uint32_t *>(p);  p += kSmemGmax;
        S.mbar = reinterpret_cast<uint64_t *>(p);  p += kTmaChunks * 8;
        S.work = reinterpret_cast<volatile int *>(p);  p += 8;
        S.task = reinterpret_cast<unsigned int *>(p);
    }
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;

    if (tid == 0) {
        for (uint32_t c = 0; c < kTmaChunks; c++) mbar_init(smem_u32(&S.mbar[c]), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    uint32_t tmaParity = 0;            // bit c = phase parity mbarrier c completes next

    for (;;) {
        __syncthreads();                       // previous block fully retired; mbarriers initialised
        if (tid == 0) *S.work = static_cast<int>(atomicAdd(P.workCounter, 1u));
        __syncthreads();
        const uint32_t b = static_cast<uint32_t>(*S.work);
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

        // ---- stage the block: TMA bulk copies (one elected thread) + ragged tail + table reset
        if (tid == 0) {
            fence_proxy_async();               // earlier generic-proxy reads of the buffer are done
            for (uint32_t c = 0; c < nChunks; c++) {
                const uint32_t bytes = min(kTmaChunk, bulk - c * kTmaChunk);
                const uint32_t bar = smem_u32(&S.mbar[c]);
                mbar_expect_tx(bar, bytes);
                tma_load_1d(smem_u32(S.in32) + c * kTmaChunk, gsrc + c * kTmaChunk, bytes, bar);
            }
        }
        if (tid >= 32 && tid < 32 + (n - bulk))
            reinterpret_cast<uint8_t *>(S.in32)[bulk + tid - 32] = gsrc[bulk + tid - 32];
        {
            uint4 *t = reinterpret_cast<uint4 *>(S.tabL);    // tabL and tabS are contiguous
            const uint4 ff = make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu);
            for (uint32_t i = tid; i < (kSmemTabL + kSmemTabS) / 16; i += kThreads) t[i] = ff;
        }
        if (tid < 4) S.task[tid] = 0u;
        __syncthreads();

        const uint32_t nh = n >= 8 ? n - 7 : 0;
        const uint32_t nW = (n + kWindow - 1) / kWindow;
        uint32_t chunksSeen = 0;
        ParseCarry pc = {0, 0, 0, 0};

        unsigned long long busy = 0, blockStart = clock64();
        for (uint32_t t = 0; t < nW + 3; t++) {
            const unsigned long long c0 = clock64();
            if (warp < kEhWarps) {
                // bytes this stage may touch: hashing window t reads < (t+1)*1024 + 11, extending
                // window t-2 reads < (t-1)*1024 + extCap + 36 + 3
                const uint32_t need = min(bulk, (t + 1) * kWindow + 16u);
                const uint32_t wantChunks = (need + kTmaChunk - 1) / kTmaChunk;
                while (chunksSeen < wantChunks) { mbar_wait(smem_u32(&S.mbar[chunksSeen]), (tmaParity >> chunksSeen) & 1u); chunksSeen++; }
                // One task queue per stage, heaviest first: extension groups of window t-2, their
                // jump-link groups, then the hash groups of window t.  Jump task g only needs extension
                // tasks g-8..g (a match is at most 256 bytes long), tracked in a completion bitmask, so
                // it almost never waits; groups of the previous window were finished a stage ago.
                const bool haveE = t >= 2 && t - 2 < nW;
                const uint32_t nE = haveE ? kGroups : 0u;
                const uint32_t nH = t < nW ? kGroups : 0u;
                const uint32_t nAll = 2u * nE + nH;
                const uint32_t ctr = smem_u32(&S.task[(t & 1u) * 2u]);
                volatile unsigned int *done = &S.task[(t & 1u) * 2u + 1u];
                for (;;) {
                    const uint32_t id = pop_task(ctr, lane);
                    if (id >= nAll) break;
#ifdef B200SP_ORDER_EHJ
                    // order E, H, J
                    const uint32_t jLo = nE + nH, hLo = nE;
#else
                    const uint32_t jLo = nE, hLo = 2u * nE;
#endif
                    if (id < nE) {
                        const uint32_t wdx = t - 2;
                        stage_extend(S, wdx & (kRing - 1), id, lane, wdx * kWindow + id * 32u + lane, n, nh, P.minMatch, P.extCap);
                        __syncwarp();
                        if (lane == 0) { __threadfence_block(); atomicOr(const_cast<unsigned int *>(done), 1u << id); }
                    } else if (id >= jLo && id < jLo + nE) {
                        const uint32_t g = id - jLo;
                        const uint32_t need = (g >= 8u ? 0x1FFu << (g - 8u) : (2u << g) - 1u);
                        while ((*done & need) != need) __nanosleep(B200SP_SPIN_NS);
                        __threadfence_block();
                        stage_jump(S, t - 2, g, lane, P.minMatch, P.lazyDepth);
                    } else {
                        const uint32_t g = id - hLo;
                        stage_hash(S, t & (kRing - 1), g, lane, t * kWindow + g * 32u + lane, nh, P.shortMask);
                    }
                }
            } else if (warp == kWarpTabL) {
                if (t >= 1 && t - 1 < nW) stage_table(S.ring0, S.tabL, (t - 1) & (kRing - 1), lane, (t - 1) * kWindow);
            } else if (warp == kWarpTabS) {
                if (t >= 1 && t - 1 < nW) stage_table(S.ring1, S.tabS, (t - 1) & (kRing - 1), lane, (t - 1) * kWindow);
            } else {
                if (lane < 2) S.task[((t + 1u) & 1u) * 2u + lane] = 0u;   // next stage's queues (nobody touches them now)
                if (t >= 3) stage_parse(S, (t - 3) & (kRing - 1), lane, (t - 3) * kWindow, pc, out);
            }
            busy += clock64() - c0;
            __syncthreads();
        }

        if (P.roleCycles && lane == 0) {
            const int role = warp < kEhWarps ? 0 : (warp - kEhWarps + 1);
            atomicAdd(&P.roleCycles[role], busy);
            if (warp == kWarpParse) { atomicAdd(&P.roleCycles[4], clock64() - blockStart); atomicAdd(&P.roleCycles[5], (unsigned long long)(nW + 3)); }
        }
        if (warp == kWarpParse && lane == 0) {
            out[pc.nOut] = make_uint4(0u, n - pc.anchor, 0u, 0u);   // trailing literals / block delimiter
            P.counts[b] = pc.nOut + 1u;
        }
        // every issued chunk must have landed before its mbarrier is re-armed for the next block
        if (warp == 0) while (chunksSeen < nChunks) { mbar_wait(smem_u32(&S.mbar[chunksSeen]), (tmaParity >> chunksSeen) & 1u); chunksSeen++; }
        tmaParity ^= (1u << nChunks) - 1u;   // only the barriers armed for this block changed phase
    }
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
bool params_for_level(int level, ParseParams &p)
{
    if (level < 1 || level > 12) return false;
    p.minMatch = 4;
    p.extCap = kMaxExtCap;
    if (level <= 2)      { p.shortMask = 0xFFFFu; p.lazyDepth = 0; }   // fast class
    else if (level <= 4) { p.shortMask = 0xFFu;   p.lazyDepth = 1; }   // dfast class
    else                 { p.shortMask = 0u;      p.lazyDepth = 2; }   // greedy/lazy/btlazy2 classes
    return true;
}

cudaError_t configure_kernels()
{
    return cudaFuncSetAttribute(lz77_parse_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                static_cast<int>(kSmemTotal));
}

cudaError_t launch_parse(const ParseParams &p, int numSMs, cudaStream_t stream)
{
    if (p.nBlocks == 0) return cudaSuccess;
    const unsigned grid = static_cast<unsigned>(p.nBlocks < static_cast<uint32_t>(numSMs) ? p.nBlocks : numSMs);
    lz77_parse_kernel<<<grid, kThreads, kSmemTotal, stream>>>(p);
    return cudaGetLastError();
}

}  // namespace b200sp
