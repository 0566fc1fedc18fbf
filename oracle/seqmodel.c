/*
 * seqmodel.c — serial CPU statement of the B200 match finder (see seqmodel.h).
 * TEST INFRASTRUCTURE ONLY: never linked into the product.
 *
 * The steps below are what the sm_100a pipeline computes per 128 KiB block; the kernel
 * distributes them over warp roles (hash/extension pool, one counting warp, parse and emit
 * warps) but must produce exactly this output.
 *
 *   0. shortcut    a block whose repeated-key count stays at chance level is one literal run
 *                  (/root/reference/src/qatseqprod.c:1308-1313).
 *   1. candidates  every position p <= n-8 hashes keyBytes bytes into a 13-bit bucket and is appended to its bucket's
 *                  list (a stable counting sort of the positions by bucket: the "GPU-resident hash-chain table" -
 *                  every bucket is the chain of ALL earlier positions with that hash, most recent last), with a
 *                  15-bit tag: further bits of the key hash (levels 1-4, keys of 5-6 bytes) or a hash of the NEXT
 *                  four bytes (levels 5-12, keys of 4 bytes).  The candidates of p are the `scan` entries before it
 *                  in its bucket whose tag equals its own, most recent first - the level-scaled search depth; at
 *                  levels 5-12 the nearN nearest entries are candidates whatever their tag.
 *   2. extension   levels 5-12: every candidate is measured in full (common prefix of src[p..] and src[cand..], up
 *                  to extCap and never past n); the longest wins, the nearer one on ties.  Levels 1-4: candidates
 *                  are ranked on their first 16 bytes, the winner competes with the 32-group's dominant offset
 *                  (the stand-in for zstd's repeated-offset probe), only the final winner is extended, and a
 *                  position adopts its right neighbour's match when that match also holds one byte earlier
 *                  (zstd's "catch up" by one; never across a 32-position group).
 *   3./4. levels 1-4   propagation B(p) = the match, among all starting at q <= p, that reaches farthest right;
 *                  greedy left-to-right parse over B with lazy look-ahead (never across a 32-position group);
 *                  zero-literal sequences repeating the previous offset are merged into their predecessor.
 *   3./4. levels 5-12  rep_parse(): the serial repcode-aware lazy parse over the own matches (zstd's lazy parser in
 *                  shape: repeated-offset probe one byte ahead, cheaper price for the repeated offset in the
 *                  look-ahead, catch-up at take time, the other repeated offset tried right after every match).
 *   5. the last entry carries the trailing literals (convention of QZSTD_decLz4s,
 *      /root/reference/src/qatseqprod.c:1037-1044, :1090).
 */
#include "seqmodel.h"
#include <stdlib.h>
#include <string.h>

#define MODEL_MAX_BLOCK (1u << 17)
#define MODEL_PROBE     16u          /* bytes a fast-class candidate is ranked on */

static inline uint32_t rd32(const uint8_t *p)
{
    uint32_t v;
    memcpy(&v, p, 4);
    return v;                       /* little-endian hosts only (x86-64, like the GPU) */
}

static inline uint32_t floorlog2(uint32_t v)   /* v >= 1 */
{
    return 31u - (uint32_t)__builtin_clz(v);
}

void seqmodel_params_for_level(int level, SeqModelParams *prm)
{
    /* One parameter class per zstd strategy class; the scan width is the level-scaled
     * search depth (the reference hands the level to its engine, /root/reference/src/qatseqprod.c:1154). */
    static const int scanOf[13] = { 0, 4, 4, 4, 8, 16, 24, 32, 32, 48, 64, 384, 768 };
    if (level < 1) level = 1;
    if (level > 12) level = 12;
    prm->keyBytes = level <= 2 ? 6 : level <= 4 ? 5 : 4;
    prm->rank16 = level <= 4 ? 1 : 0;
    prm->scan = scanOf[level];
    prm->minMatch = 4;
    prm->extCap = 256;
    prm->lazyDepth = 2;
    prm->window = 32;            /* fast classes: lazy look-ahead stays inside the 32-position group one warp owns */
    prm->backExt = level <= 4 ? 1 : 0;   /* the repcode-aware parse catches up at take time instead */
    prm->repParse = level <= 4 ? 0 : 1;
    prm->domBias = level <= 4 ? 0 : -1;  /* fast classes probe the group's dominant offset; ties go to it */
    prm->nearN = level <= 4 ? 0 : 16;    /* levels 5-12: tags speak for the next four bytes; the 16 nearest entries are measured whatever their tag */
}

typedef struct { uint32_t end, off; } BestMatch;   /* end = p + len (0 = none) */

static inline int32_t gain_of(uint32_t len, uint32_t off)
{
    return (int32_t)(len * 4u) - (int32_t)floorlog2(off + 1u);
}

static inline uint32_t key_hash(uint32_t lo, uint32_t hi, int keyBytes)
{
    const uint32_t mask = keyBytes <= 4 ? 0u : keyBytes == 5 ? 0xFFu : 0xFFFFu;
    return lo * 0x9E3779B1u + (hi & mask) * 0xC2B2AE3Du;
}

/* Steps 1-2 for every position: own best match (len 0 = none).  Shared by seqmodel_block and by the
 * lane-level statement of the parse warps (lanemodel.c). */
int seqmodel_own_matches(const uint8_t *src, size_t n, const SeqModelParams *prm, uint32_t *ownLen, uint32_t *ownOff)
{
    const uint32_t N = (uint32_t)n;
    const uint32_t nh = N >= 8 ? N - 7 : 0;      /* positions that can hash 8 bytes */
    const uint32_t nB = 1u << SEQMODEL_BUCKET_BITS;
    /* counting sort of the hashable positions by bucket: hist -> segment starts -> entries {pos | tag << 17} */
    uint32_t *start = (uint32_t *)calloc(nB + 1, sizeof(uint32_t));
    uint32_t *count = (uint32_t *)calloc(nB, sizeof(uint32_t));
    uint32_t *sorted = (uint32_t *)malloc(((size_t)nh + 1) * sizeof(uint32_t));
    if (!start || !count || !sorted) { free(start); free(count); free(sorted); return -1; }
    for (uint32_t p = 0; p < nh; p++) {
        const uint32_t v = key_hash(rd32(src + p), rd32(src + p + 4), prm->keyBytes);
        const uint32_t b = v >> (32 - SEQMODEL_BUCKET_BITS);
        start[b + 1]++;
    }
    for (uint32_t b = 0; b < nB; b++) start[b + 1] += start[b];
    const uint32_t scan = (uint32_t)prm->scan < SEQMODEL_IDX_CAP ? (uint32_t)prm->scan : SEQMODEL_IDX_CAP;
    for (uint32_t p = 0; p < N; p++) {
        uint32_t bestLen = 0, bestOff = 0;
        if (p < nh) {
            const uint32_t lo = rd32(src + p), hi = rd32(src + p + 4);
            const uint32_t v = key_hash(lo, hi, prm->keyBytes);
            const uint32_t b = v >> (32 - SEQMODEL_BUCKET_BITS);
            uint32_t tag = (v >> (32 - SEQMODEL_BUCKET_BITS - SEQMODEL_TAG_BITS)) & ((1u << SEQMODEL_TAG_BITS) - 1u);
            if (prm->nearN > 0) tag = (hi * 0xC2B2AE3Du) >> (32 - SEQMODEL_TAG_BITS);     /* levels 5-12: the tag speaks for the NEXT four bytes */
            const uint32_t idx = count[b];                   /* entries of the bucket before p */
            sorted[start[b] + idx] = p | (tag << 17);
            count[b] = idx + 1;
            uint32_t lim = N - p;
            if (lim > (uint32_t)prm->extCap) lim = (uint32_t)prm->extCap;
            const uint32_t avail = idx < scan ? idx : scan;
            for (uint32_t j = 1; j <= avail; j++) {
                const uint32_t e = sorted[start[b] + idx - j];
                if (prm->nearN > 0 ? (j > (uint32_t)prm->nearN && (e >> 17) != tag) : (e >> 17) != tag) continue;
                const uint32_t q = e & 0x1FFFFu;
                const uint8_t *a = src + p, *c = src + q;
                const uint32_t stop = prm->rank16 && lim > MODEL_PROBE ? MODEL_PROBE : lim;   /* fast classes rank on 16 bytes */
                uint32_t ml = 0;
                while (ml < stop && a[ml] == c[ml]) ml++;
                if (prm->rank16 && ml < 4) ml = 0;
                if (ml > bestLen) { bestLen = ml; bestOff = p - q; }      /* ties keep the nearer candidate */
            }
        }
        ownLen[p] = bestLen;            /* fast classes: still the rank on 16 bytes */
        ownOff[p] = bestLen ? bestOff : 0;
        if (!prm->rank16) { if (ownLen[p] < (uint32_t)prm->minMatch) { ownLen[p] = 0; ownOff[p] = 0; } continue; }
        if ((p & 31u) != 31u && p + 1 < N) continue;
        /* ---- end of a 32-position group (fast classes): the dominant offset, then the long extension */
        {
            const uint32_t g0 = p & ~31u;
            uint32_t D = 0;
            if (prm->domBias >= 0) {
                /* the offset most positions of the group chose (ties: the smaller one) ... */
                uint32_t bestKey = 0;
                for (uint32_t a = g0; a <= p; a++) {
                    if (!ownLen[a]) continue;
                    uint32_t cnt = 0;
                    for (uint32_t c = g0; c <= p; c++) cnt += ownLen[c] && ownOff[c] == ownOff[a];
                    const uint32_t key = (cnt << 17) | (0x1FFFFu - ownOff[a]);
                    if (key > bestKey) bestKey = key;
                }
                if (bestKey) D = 0x1FFFFu - (bestKey & 0x1FFFFu);
            }
            for (uint32_t a = g0; a <= p; a++) {
                uint32_t lim = a < nh ? N - a : 0;
                if (lim > (uint32_t)prm->extCap) lim = (uint32_t)prm->extCap;
                const uint32_t probe = lim > MODEL_PROBE ? MODEL_PROBE : lim;
                /* ... is probed by every position of the group (the stand-in for zstd's repeated-offset probe: no
                 * hash needed, so 4-byte matches between changed fields are found) and wins unless the scan's
                 * winner is longer by more than domBias bytes */
                if (D && a >= D && a < nh && ownOff[a] != D) {
                    uint32_t ml = 0;
                    while (ml < probe && src[a + ml] == src[a - D + ml]) ml++;
                    if (ml >= 4 && ml + (uint32_t)prm->domBias >= ownLen[a]) { ownLen[a] = ml; ownOff[a] = D; }
                }
                if (ownLen[a] == MODEL_PROBE) {                            /* extend only the winner */
                    uint32_t l = ownLen[a];
                    while (l < lim && src[a + l] == src[a - ownOff[a] + l]) l++;
                    ownLen[a] = l;
                }
                if (ownLen[a] < (uint32_t)prm->minMatch) { ownLen[a] = 0; ownOff[a] = 0; }
            }
        }
    }
    free(start); free(count); free(sorted);
    if (prm->backExt) {
        /* ascending: own[p + 1] is still the searched value when p looks at it */
        for (uint32_t p = 0; p + 1 < N; p++) {
            if ((p & 31u) == 31u) continue;                  /* the neighbour belongs to another group (another warp) */
            const uint32_t l1 = ownLen[p + 1], o1 = ownOff[p + 1];
            if (l1 && p >= o1 && src[p] == src[p - o1] && l1 + 1 > ownLen[p] && l1 + 1 <= (uint32_t)prm->extCap) {
                ownLen[p] = l1 + 1;
                ownOff[p] = o1;
            }
        }
    }
    return 0;
}

/* Step 0, the incompressible shortcut (/root/reference/src/qatseqprod.c:1308-1313 returns one literal run for a
 * block the engine could not compress).  Every hashable position sets one bit of a SEQMODEL_BITMAP_BITS-bit map
 * chosen by its key hash; hits = positions whose bit was already set.  A block whose hits stay below what chance
 * alone produces (plus six standard deviations) holds too few repeated keys to be worth parsing. */
static uint32_t isqrt32(uint32_t v)
{
    uint32_t r = 0;
    for (uint32_t bit = 1u << 15; bit; bit >>= 1) { const uint32_t t = r | bit; if ((uint64_t)t * t <= v) r = t; }
    return r;
}

uint32_t seqmodel_chance_threshold(uint32_t nh)
{
    const uint64_t M = SEQMODEL_BITMAP_BITS, x = nh;
    const uint64_t a = x * x / M;                                   /* expected hits = M (x/M - 1 + exp(-x/M)), by its series */
    const uint64_t e = a / 2 - a * x / (6 * M) + a * a / (24 * M);
    return (uint32_t)e + 6u * isqrt32((uint32_t)e) + 24u;
}

int seqmodel_incompressible(const uint8_t *src, size_t n, const SeqModelParams *prm)
{
    const uint32_t N = (uint32_t)n, nh = N >= 8 ? N - 7 : 0;
    uint8_t *seen = (uint8_t *)calloc(SEQMODEL_BITMAP_BITS, 1);
    if (!seen) return 0;
    uint32_t hits = 0;
    for (uint32_t p = 0; p < nh; p++) {
        const uint32_t v = key_hash(rd32(src + p), rd32(src + p + 4), prm->keyBytes);
        const uint32_t i = (uint32_t)(((uint64_t)v * SEQMODEL_BITMAP_BITS) >> 32);
        hits += seen[i];
        seen[i] = 1;
    }
    free(seen);
    return hits < seqmodel_chance_threshold(nh);
}

/* Step 4 for levels 5-12: the serial, repcode-aware lazy parse (what one warp per block computes, see
 * stage_rep_parse in lz77_kernels.cu).  It follows zstd's lazy parser in shape - repcode probe one byte ahead,
 * look-ahead of lazyDepth positions with a cheaper price for the repeated offset, "catch up" to the left at
 * take time, a second repeated offset tried right after every match - over the own matches of steps 1-2.
 * rep1 / rep2 are the offsets of the last two emitted sequences with distinct offsets; repcode lengths are not
 * capped (a capped own match is extended when it is taken). */
static uint32_t common_len(const uint8_t *src, uint32_t a, uint32_t b, uint32_t lim)   /* b < a */
{
    uint32_t l = 0;
    while (l < lim && src[a + l] == src[b + l]) l++;
    return l;
}

static uint32_t rep_len(const uint8_t *src, uint32_t N, uint32_t p, uint32_t rep)
{
    if (!rep || p < rep || p + 4u > N) return 0;
    if (rd32(src + p) != rd32(src + p - rep)) return 0;
    return common_len(src, p, p - rep, N - p);
}

#define MODEL_CATCHUP 31u           /* bytes a taken match may grow to the left (one warp ballot) */

static size_t rep_parse(const uint8_t *src, uint32_t N, const uint32_t *ownLen, const uint32_t *ownOff,
                        ZSTD_Sequence *out, size_t outCap, const SeqModelParams *prm)
{
    const uint32_t nh = N >= 8 ? N - 7 : 0;
    const uint32_t extCap = (uint32_t)prm->extCap;
    size_t ns = 0;
    uint32_t ip = 0, anchor = 0, rep1 = 0, rep2 = 0;
    while (ip < nh) {
        uint32_t r = rep_len(src, N, ip + 1, rep1);
        uint32_t ml, off, start;
        int isRep;
        if (r == 0 && ownLen[ip] == 0) { ip++; continue; }
        if (ownLen[ip] > r) { ml = ownLen[ip]; off = ownOff[ip]; start = ip; isRep = 0; }
        else { ml = r; off = rep1; start = ip + 1; isRep = 1; }
        for (;;) {                                      /* lazy look-ahead from ip */
            int moved = 0;
            for (uint32_t d = 1; d <= (uint32_t)prm->lazyDepth && !moved; d++) {
                const uint32_t q = ip + d;
                if (q >= nh) break;
                const uint32_t rq = rep_len(src, N, q, rep1);
                int32_t price = isRep ? 0 : (int32_t)floorlog2(off + 1u);
                const uint32_t wgt = d == 1 ? 3u : 4u;                  /* zstd's weights for the repeated offset */
                if (rq >= 4u && (int32_t)(rq * wgt) > (int32_t)(ml * wgt) - price + 1) {
                    ml = rq; off = rep1; start = q; isRep = 1; price = 0;
                }
                if (ownLen[q] && gain_of(ownLen[q], ownOff[q]) > (int32_t)(ml * 4u) - price + (d == 1 ? 4 : 7)) {
                    ml = ownLen[q]; off = ownOff[q]; start = q; isRep = 0;
                    ip = q; moved = 1;
                }
            }
            if (!moved) break;
        }
        if (!isRep) {
            if (ml >= extCap) ml = common_len(src, start, start - off, N - start);      /* cut by the cap: extend */
            uint32_t k = 0;                                                            /* catch up to the left */
            while (k < MODEL_CATCHUP && start > anchor && start > off && src[start - 1] == src[start - 1 - off]) { start--; ml++; k++; }
            if (off != rep1) { rep2 = rep1; rep1 = off; }
        }
        if (start == anchor && ns > 0 && out[ns - 1].offset == off) out[ns - 1].matchLength += ml;
        else {
            if (ns + 1 >= outCap) return (size_t)-1;
            out[ns].offset = off; out[ns].litLength = start - anchor; out[ns].matchLength = ml; out[ns].rep = 0;
            ns++;
        }
        ip = anchor = start + ml;
        while (ip < nh) {                               /* the other repeated offset, right after the match */
            const uint32_t m2 = rep_len(src, N, ip, rep2);
            if (m2 < 4u) break;
            { const uint32_t t = rep2; rep2 = rep1; rep1 = t; }
            if (ns + 1 >= outCap) return (size_t)-1;
            out[ns].offset = rep1; out[ns].litLength = 0; out[ns].matchLength = m2; out[ns].rep = 0;
            ns++;
            ip = anchor = ip + m2;
        }
    }
    if (ns >= outCap) return (size_t)-1;
    out[ns].offset = 0; out[ns].litLength = N - anchor; out[ns].matchLength = 0; out[ns].rep = 0;
    return ns + 1;
}

size_t seqmodel_block(const uint8_t *src, size_t n, ZSTD_Sequence *out, size_t outCap,
                      const SeqModelParams *prm)
{
    if (n > MODEL_MAX_BLOCK || outCap == 0) return (size_t)-1;
    if (seqmodel_incompressible(src, n, prm)) {
        out[0].offset = 0; out[0].litLength = (uint32_t)n; out[0].matchLength = 0; out[0].rep = 0;
        return 1;
    }

    const uint32_t N = (uint32_t)n;
    uint32_t *ownLen = (uint32_t *)malloc((N + 1) * sizeof(uint32_t));
    uint32_t *ownOff = (uint32_t *)malloc((N + 1) * sizeof(uint32_t));
    BestMatch *B = (BestMatch *)calloc(N + 1, sizeof(BestMatch));
    if (!ownLen || !ownOff || !B || seqmodel_own_matches(src, n, prm, ownLen, ownOff) != 0) {
        free(ownLen); free(ownOff); free(B);
        return (size_t)-1;
    }
    if (prm->repParse) {
        const size_t r = rep_parse(src, N, ownLen, ownOff, out, outCap, prm);
        free(ownLen); free(ownOff); free(B);
        return r;
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
