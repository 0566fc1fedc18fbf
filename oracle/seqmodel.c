/*
 * seqmodel.c — serial CPU statement of the B200 match finder (see seqmodel.h).
 * TEST INFRASTRUCTURE ONLY: never linked into the product.
 *
 * The four steps below are what the sm_100a pipeline computes per 128 KiB block; the kernel
 * distributes them over warp roles (hash warps, two table warps, extension warps, one parse
 * warp) but must produce exactly this output.
 *
 *   1. candidates  every position p <= n-8 hashes 8 bytes (long) and shortBytes bytes (short)
 *                  and reads-then-overwrites one slot of each table: the candidate is the most
 *                  recent earlier position with the same hash.
 *   2. extension   common prefix of src[p..] and src[cand..]: candidates are ranked on their first
 *                  16 bytes, the winner is extended up to extCap (and never past n).
 *   3. propagation B(p) = the match, among all starting at q <= p, that reaches farthest right.
 *   4. parse       greedy left-to-right over B with lazy look-ahead; zero-literal sequences
 *                  repeating the previous offset are merged into their predecessor; the last
 *                  entry carries the trailing literals (convention of QZSTD_decLz4s,
 *                  /root/reference/src/qatseqprod.c:1037-1044, :1090).
 */
#include "seqmodel.h"
#include <stdlib.h>
#include <string.h>

#define MODEL_MAX_BLOCK (1u << 17)
#define MODEL_PROBE     16u          /* bytes compared per candidate before a winner is picked */

static inline uint32_t rd32(const uint8_t *p)
{
    uint32_t v;
    memcpy(&v, p, 4);
    return v;                       /* little-endian hosts only (x86-64, like the GPU) */
}

static inline uint32_t hash_long(uint32_t lo, uint32_t hi, int bits)
{
    return (lo * 0x9E3779B1u + hi * 0x85EBCA77u) >> (32 - bits);
}

static inline uint32_t hash_short(uint32_t lo, uint32_t hi, int bytes, int bits)
{
    uint32_t h;
    if (bytes <= 4)      h = lo * 0x9E3779B1u;
    else if (bytes == 5) h = lo * 0x9E3779B1u + (hi & 0xFFu) * 0xC2B2AE3Du;
    else                 h = lo * 0x9E3779B1u + (hi & 0xFFFFu) * 0xC2B2AE3Du;
    return h >> (32 - bits);
}

static inline uint32_t floorlog2(uint32_t v)   /* v >= 1 */
{
    return 31u - (uint32_t)__builtin_clz(v);
}

void seqmodel_params_for_level(int level, SeqModelParams *prm)
{
    /* One parameter class per zstd strategy class. */
    prm->longBits = 14;
    prm->shortBits = 12;         /* 8 KiB: the kernel spends its shared memory on a wider pipeline window instead */
    prm->shortBytes = 5;
    prm->minMatch = 4;
    prm->extCap = 256;
    prm->lazyDepth = 1;
    prm->window = 32;            /* lazy look-ahead stays inside the 32-position group one warp owns */
    if (level <= 2) {            /* fast class */
        prm->shortBytes = 6;
        prm->lazyDepth = 0;
    } else if (level <= 4) {     /* dfast class */
        prm->lazyDepth = 1;
    } else {                     /* greedy / lazy / lazy2 / btlazy2 classes */
        prm->shortBytes = 4;
        prm->lazyDepth = 2;
    }
}

typedef struct { uint32_t end, off; } BestMatch;   /* end = p + len (0 = none) */

static inline int32_t gain_of(uint32_t len, uint32_t off)
{
    return (int32_t)(len * 4u) - (int32_t)floorlog2(off + 1u);
}

/* Steps 1-2 for every position: own best match (len 0 = none).  Shared by seqmodel_block and by the
 * lane-level statement of the parse warps (lanemodel.c). */
int seqmodel_own_matches(const uint8_t *src, size_t n, const SeqModelParams *prm, uint32_t *ownLen, uint32_t *ownOff)
{
    const uint32_t N = (uint32_t)n;
    const uint32_t nh = N >= 8 ? N - 7 : 0;      /* positions that can hash 8 bytes */
    const size_t szL = (size_t)1 << prm->longBits, szS = (size_t)1 << prm->shortBits;
    uint16_t *tabL = (uint16_t *)malloc(szL * sizeof(uint16_t));
    uint16_t *tabS = (uint16_t *)malloc(szS * sizeof(uint16_t));
    if (!tabL || !tabS) { free(tabL); free(tabS); return -1; }
    /* A slot keeps (position >> 1) of the most recent position with that hash: 16 bits cover the
     * whole 128 KiB block.  The dropped parity bit is recovered by testing both 2v and 2v+1.
     * 0xFFFF (positions 131070/131071, never hashable) is the empty marker. */
    memset(tabL, 0xFF, szL * sizeof(uint16_t));
    memset(tabS, 0xFF, szS * sizeof(uint16_t));
    for (uint32_t p = 0; p < N; p++) {
        uint32_t bestLen = 0, bestOff = 0;
        if (p < nh) {
            const uint32_t lo = rd32(src + p), hi = rd32(src + p + 4);
            const uint32_t hL = hash_long(lo, hi, prm->longBits);
            const uint32_t hS = hash_short(lo, hi, prm->shortBytes, prm->shortBits);
            const uint32_t base[2] = { 2u * tabL[hL], 2u * tabS[hS] };
            tabL[hL] = (uint16_t)(p >> 1);
            tabS[hS] = (uint16_t)(p >> 1);
            uint32_t lim = N - p;
            if (lim > (uint32_t)prm->extCap) lim = (uint32_t)prm->extCap;
            /* Phase 1: each table contributes ONE candidate: of the two positions its slot stands
             * for (2v, 2v+1) the one whose first 4 bytes equal ours, the nearer one if both do.  It is
             * measured over its first PROBE bytes only.  The short-table candidate replaces the
             * long-table one if it is longer, or as long and nearer.
             * Phase 2: only the winner is extended, up to extCap. */
            const uint32_t probe = lim < MODEL_PROBE ? lim : MODEL_PROBE;
            const uint32_t a4 = rd32(src + p);
            for (int t = 0; t < 2; t++) {
                const uint32_t q0 = base[t];
                uint32_t q = UINT32_MAX;
                if (q0 < p && rd32(src + q0) == a4) q = q0;
                if (q0 + 1 < p && rd32(src + q0 + 1) == a4) q = q0 + 1;
                if (q == UINT32_MAX) continue;
                const uint8_t *a = src + p, *b = src + q;
                uint32_t ml = 4;
                while (ml < probe && a[ml] == b[ml]) ml++;
                const uint32_t off = p - q;
                if (ml > bestLen || (ml == bestLen && off < bestOff)) { bestLen = ml; bestOff = off; }
            }
            if (bestLen == MODEL_PROBE) {
                const uint8_t *a = src + p, *b = src + p - bestOff;
                while (bestLen < lim && a[bestLen] == b[bestLen]) bestLen++;
            }
            if (bestLen < (uint32_t)prm->minMatch) bestLen = 0;
        }
        ownLen[p] = bestLen;
        ownOff[p] = bestOff;
    }
    free(tabL); free(tabS);
    return 0;
}

size_t seqmodel_block(const uint8_t *src, size_t n, ZSTD_Sequence *out, size_t outCap,
                      const SeqModelParams *prm)
{
    if (n > MODEL_MAX_BLOCK || outCap == 0) return (size_t)-1;

    const uint32_t N = (uint32_t)n;
    uint32_t *ownLen = (uint32_t *)malloc((N + 1) * sizeof(uint32_t));
    uint32_t *ownOff = (uint32_t *)malloc((N + 1) * sizeof(uint32_t));
    BestMatch *B = (BestMatch *)calloc(N + 1, sizeof(BestMatch));
    if (!ownLen || !ownOff || !B || seqmodel_own_matches(src, n, prm, ownLen, ownOff) != 0) {
        free(ownLen); free(ownOff); free(B);
        return (size_t)-1;
    }

    /* step 3: run-max of match ends carried left to right */
    BestMatch run = {0, 0};
    for (uint32_t p = 0; p < N; p++) {
        if (ownLen[p] && p + ownLen[p] > run.end) { run.end = p + ownLen[p]; run.off = ownOff[p]; }
        B[p] = run;                 /* strictly-greater replaces: ties keep the older match */
    }
    free(ownLen); free(ownOff);

    /* step 4: parse */
    const uint32_t minMatch = (uint32_t)prm->minMatch;
    const uint32_t W = (uint32_t)prm->window;
    size_t nseq = 0;
    uint32_t anchor = 0, cursor = 0, prevOff = 0;
#define HAS(p_) ((p_) < N && B[p_].end >= (p_) + minMatch)
    while (cursor < N) {
        uint32_t p = cursor;
        while (p < N && !HAS(p)) p++;
        if (p >= N) break;
        for (;;) {                  /* lazy look-ahead; never crosses a window boundary */
            if (prm->lazyDepth < 1) break;
            const int32_t g0 = gain_of(B[p].end - p, B[p].off);
            uint32_t q = p + 1;
            if ((q % W) == 0 || !HAS(q)) break;
            if (gain_of(B[q].end - q, B[q].off) > g0 + 4) { p = q; continue; }
            if (prm->lazyDepth < 2) break;
            q = p + 2;
            if ((q % W) == 0 || !HAS(q)) break;
            if (gain_of(B[q].end - q, B[q].off) > g0 + 7) { p = q; continue; }
            break;
        }
        const uint32_t len = B[p].end - p, off = B[p].off, lit = p - anchor;
        if (lit == 0 && nseq > 0 && off == prevOff) {
            out[nseq - 1].matchLength += len;
        } else {
            if (nseq + 1 >= outCap) { nseq = (size_t)-1; goto done; }
            out[nseq].offset = off;
            out[nseq].litLength = lit;
            out[nseq].matchLength = len;
            out[nseq].rep = 0;
            nseq++;
        }
        prevOff = off;
        cursor = anchor = p + len;
    }
#undef HAS
    if (nseq >= outCap) { nseq = (size_t)-1; goto done; }
    out[nseq].offset = 0;
    out[nseq].litLength = N - anchor;
    out[nseq].matchLength = 0;
    out[nseq].rep = 0;
    nseq++;
done:
    free(B);
    return nseq;
}
