/*
 * lanemodel.c — lane-level CPU statement of the kernel's parse stage.
 * TEST INFRASTRUCTURE ONLY: never linked into the product.
 *
 * seqmodel.c says WHAT the kernels emit (a serial greedy/lazy parse over B).  This file says HOW the
 * two parse warps of qat-zstd-plugin_b200/csrc/lz77_kernels.cu get there without a serial scan over
 * positions, with every "lane" written as a loop iteration, so that the formulation can be checked
 * against seqmodel_block on the CPU (tests/test_oracle.py) before and independently of the GPU:
 *
 *   E (extension warps, one 32-position group per task) leaves per position the packed PREFIX MAXIMUM of
 *       the matches starting in its group  {end - groupStart : 9 | 31 - lane : 6 | offset : 17},
 *       per group the maximum of the whole group (gmax) and the mask of positions whose own prefix
 *       maximum is a usable match (gown).
 *   P1 (lane j = group j of a half window of KGPW groups; one warp per half, both in the same stage):
 *       carry  c_j   = farthest-reaching match of the previous 8 groups, re-based to group j;
 *       B(p)         = max(prefix maximum at p, c_j);   has_j = gown_j | {lanes covered by c_j};
 *       walk(e)      = greedy/lazy parse of group j entered at e; every decision is memoised as a link
 *                      word (take {end : 9 | take position : 5 | offset : 17} or lazy hop {1 << 31 | next : 9})
 *                      so a position is evaluated once;
 *       entries      = every lane guesses "entered at my first position"; guesses are corrected by a
 *                      prefix maximum over the exits of live lanes until nothing changes (= the serial
 *                      parse, because a walk depends only on where it is entered).
 *   P2 (one warp, one window behind P1): follows the links from the final entries, counts, scans for
 *       anchors / previous offsets / output slots, and writes the ZSTD_Sequence array.
 */
#include "seqmodel.h"
#include <stdlib.h>
#include <string.h>

#define KGRP    32u               /* positions per group (one warp's worth) */
#define KGPW    26u               /* groups per parse pass = lanes of a parse warp in use (kHalf of the kernel) */
#define KWIN    (KGPW * KGRP)     /* positions per parse pass (half a pipeline window) */

static inline uint32_t umax(uint32_t a, uint32_t b) { return a > b ? a : b; }
static inline uint32_t floorlog2(uint32_t v) { return 31u - (uint32_t)__builtin_clz(v); }
static inline int32_t gain_packed(uint32_t b, uint32_t p)     /* b: packed match, p: lane */
{
    return (int32_t)(((b >> 23) - p) * 4u) - (int32_t)floorlog2((b & 0x1FFFFu) + 1u);
}

typedef struct {
    const uint32_t *pk;      /* packed prefix maxima of this group (32 words) */
    uint32_t *link;          /* memo of decisions of this group (32 words)    */
    uint32_t c, has, visited;
    uint32_t minMatch, lazyDepth;
} Group;

/* one decision: the cursor stands on local position p (a usable match starts there).  Either a take link
 * {end:9 | take lane:5 | offset:17} or, when the lazy rule prefers a later start, a hop link
 * {1 << 31 | next position << 22}; the look-ahead never leaves the group. */
static uint32_t eval_step(const Group *g, uint32_t p)
{
    const uint32_t b0 = umax(g->pk[p], g->c);
    if (g->lazyDepth >= 1) {
        const int32_t g0 = gain_packed(b0, p);
        const uint32_t q1 = p + 1, q2 = p + 2;
        if (q1 < KGRP && ((g->has >> q1) & 1u)) {
            if (gain_packed(umax(g->pk[q1], g->c), q1) > g0 + 4) return 0x80000000u | (q1 << 22);
            if (g->lazyDepth >= 2 && q2 < KGRP && ((g->has >> q2) & 1u) &&
                gain_packed(umax(g->pk[q2], g->c), q2) > g0 + 7) return 0x80000000u | (q2 << 22);
        }
    }
    return ((b0 >> 23) << 22) | (p << 17) | (b0 & 0x1FFFFu);
}

/* walk of one group from local cursor `cur` (< 32); returns the local exit (>= 32) */
static uint32_t walk_exit(Group *g, uint32_t cur)
{
    while (cur < KGRP) {
        const uint32_t m = g->has & (0xFFFFFFFFu << cur);
        if (!m) return KGRP;
        const uint32_t p0 = (uint32_t)__builtin_ctz(m);
        if (!((g->visited >> p0) & 1u)) { g->link[p0] = eval_step(g, p0); g->visited |= 1u << p0; }
        cur = (g->link[p0] >> 22) & 0x1FFu;
    }
    return cur;
}

size_t lanemodel_block(const uint8_t *src, size_t n, ZSTD_Sequence *out, size_t outCap,
                       const SeqModelParams *prm)
{
    if (n > (1u << 17) || outCap == 0 || prm->window != (int)KGRP) return (size_t)-1;
    /* levels 5-12: the kernel's parse IS the serial one (one warp walks it, stage_rep_parse); nothing to restate */
    if (prm->repParse) return seqmodel_block(src, n, out, outCap, prm);
    if (seqmodel_incompressible(src, n, prm)) {          /* step 0 of the model: one literal run */
        out[0].offset = 0; out[0].litLength = (uint32_t)n; out[0].matchLength = 0; out[0].rep = 0;
        return 1;
    }
    const uint32_t N = (uint32_t)n;
    const uint32_t nW = (N + KWIN - 1) / KWIN;
    const uint32_t nG = nW * KGPW;
    const uint32_t minMatch = (uint32_t)prm->minMatch;
    uint32_t *ownLen = (uint32_t *)calloc(nW * KWIN + 1, sizeof(uint32_t));
    uint32_t *ownOff = (uint32_t *)calloc(nW * KWIN + 1, sizeof(uint32_t));
    uint32_t *pk = (uint32_t *)calloc(nW * KWIN + 1, sizeof(uint32_t));
    uint32_t *link = (uint32_t *)calloc(nW * KWIN + 1, sizeof(uint32_t));
    uint32_t *gmax = (uint32_t *)calloc(nG + 1, sizeof(uint32_t));
    uint32_t *gown = (uint32_t *)calloc(nG + 1, sizeof(uint32_t));
    uint32_t *hasA = (uint32_t *)calloc(nG + 1, sizeof(uint32_t));
    uint32_t *entA = (uint32_t *)calloc(nG + 1, sizeof(uint32_t));
    size_t nseq = (size_t)-1;
    if (!ownLen || !ownOff || !pk || !link || !gmax || !gown || !hasA || !entA) goto done;
    if (N && seqmodel_own_matches(src, n, prm, ownLen, ownOff) != 0) goto done;

    /* ---- E: packed prefix maxima per group */
    for (uint32_t G = 0; G < nG; G++) {
        uint32_t run = 0, own = 0;
        for (uint32_t l = 0; l < KGRP; l++) {
            const uint32_t p = G * KGRP + l;
            uint32_t w = 0;
            if (p < N && ownLen[p]) w = ((l + ownLen[p]) << 23) | ((31u - l) << 17) | ownOff[p];
            run = umax(run, w);
            pk[p] = run;
            if ((run >> 23) >= l + minMatch) own |= 1u << l;
        }
        gmax[G] = run;
        gown[G] = own;
    }

    /* ---- P1: entries of every group, window by window */
    uint32_t cursor = 0;
    for (uint32_t w = 0; w < nW; w++) {
        const uint32_t base = w * KWIN;
        Group g[KGPW];
        uint32_t entry[KGPW], exitPos[KGPW], pm[KGPW];
        for (uint32_t j = 0; j < KGPW; j++) {
            const uint32_t G = w * KGPW + j;
            uint32_t c = 0;
            for (uint32_t k = 1; k <= 8 && k <= G; k++) {
                const uint32_t v = gmax[G - k], rel = v >> 23;
                if (rel > 32u * k) c = umax(c, ((rel - 32u * k) << 23) | ((32u + k) << 17) | (v & 0x1FFFFu));
            }
            const uint32_t cRel = c >> 23;
            uint32_t cover = 0;
            if (cRel >= minMatch) cover = (cRel - minMatch >= 31u) ? 0xFFFFFFFFu : (2u << (cRel - minMatch)) - 1u;
            g[j].pk = pk + G * KGRP; g[j].link = link + G * KGRP;
            g[j].c = c; g[j].has = gown[G] | cover; g[j].visited = 0;
            g[j].minMatch = minMatch; g[j].lazyDepth = (uint32_t)prm->lazyDepth;
            /* first guess: the parse arrives through the carried match; any guess converges to the same fixed point */
            entry[j] = j == 0 ? umax(cursor, base) : base + j * KGRP + (cRel < KGRP ? cRel : KGRP);
        }
        for (;;) {
            for (uint32_t j = 0; j < KGPW; j++) {
                const uint32_t segStart = base + j * KGRP, segEnd = segStart + KGRP;
                const int live = entry[j] < segEnd;
                exitPos[j] = live ? segStart + walk_exit(&g[j], entry[j] - segStart) : 0u;
            }
            uint32_t run = 0;
            for (uint32_t j = 0; j < KGPW; j++) {           /* inclusive prefix maximum */
                run = umax(run, j == 0 ? umax(exitPos[0], entry[0]) : exitPos[j]);
                pm[j] = run;
            }
            int changed = 0;
            for (uint32_t j = 1; j < KGPW; j++) {
                const uint32_t want = umax(pm[j - 1], base + j * KGRP);
                if (want != entry[j]) { entry[j] = want; changed = 1; }
            }
            if (!changed) break;
        }
        cursor = umax(pm[KGPW - 1], base + KWIN);
        for (uint32_t j = 0; j < KGPW; j++) { hasA[w * KGPW + j] = g[j].has; entA[w * KGPW + j] = entry[j]; }
    }

    /* ---- P2: follow the links from the final entries, scan, emit */
    {
        uint32_t anchorC = 0, prevOffC = 0;
        size_t nOut = 0;
        for (uint32_t w = 0; w < nW; w++) {
            const uint32_t base = w * KWIN;
            uint32_t cnt[KGPW], merges[KGPW], firstPos[KGPW], firstOff[KGPW], lastEnd[KGPW], lastOff[KGPW];
            for (uint32_t j = 0; j < KGPW; j++) {            /* counting walk */
                const uint32_t G = w * KGPW + j, segStart = base + j * KGRP;
                cnt[j] = merges[j] = firstPos[j] = firstOff[j] = lastEnd[j] = lastOff[j] = 0;
                uint32_t cur = entA[G] - segStart;           /* >= 32 (or wrapped huge) when passed over */
                if (entA[G] < segStart) goto done;           /* cannot happen */
                while (cur < KGRP) {
                    const uint32_t m = hasA[G] & (0xFFFFFFFFu << cur);
                    if (!m) break;
                    const uint32_t L = link[G * KGRP + (uint32_t)__builtin_ctz(m)];
                    if (L >> 31) { cur = (L >> 22) & 0x1FFu; continue; }      /* hop: a later start is better */
                    const uint32_t p = segStart + ((L >> 17) & 31u), end = segStart + (L >> 22), off = L & 0x1FFFFu;
                    if (cnt[j] && p == lastEnd[j] && off == lastOff[j]) merges[j]++;
                    if (!cnt[j]) { firstPos[j] = p; firstOff[j] = off; }
                    cnt[j]++; lastEnd[j] = end; lastOff[j] = off;
                    cur = end - segStart;
                }
            }
            /* exclusive "last match" scan, head merges, output slots */
            uint32_t runE = anchorC, runO = prevOffC;
            size_t idx = nOut;
            for (uint32_t j = 0; j < KGPW; j++) {
                const uint32_t G = w * KGPW + j, segStart = base + j * KGRP;
                uint32_t anchor = runE, prevOff = runO;
                const int headMerge = cnt[j] && firstPos[j] == anchor && firstOff[j] == prevOff && anchor > 0;
                /* emitting walk: new entries start at the slot the scan assigned to this lane */
                const size_t firstIdx = idx;
                const uint32_t fresh = cnt[j] - merges[j] - (headMerge ? 1u : 0u);
                uint32_t cur = entA[G] - segStart;
                while (cur < KGRP) {
                    const uint32_t m = hasA[G] & (0xFFFFFFFFu << cur);
                    if (!m) break;
                    const uint32_t L = link[G * KGRP + (uint32_t)__builtin_ctz(m)];
                    if (L >> 31) { cur = (L >> 22) & 0x1FFu; continue; }      /* hop: a later start is better */
                    const uint32_t p = segStart + ((L >> 17) & 31u), end = segStart + (L >> 22), off = L & 0x1FFFFu;
                    if (p == anchor && off == prevOff && anchor > 0) {
                        out[idx - 1].matchLength += end - p;     /* continuation (of this lane's or an earlier lane's sequence) */
                    } else {
                        if (idx + 1 >= outCap) goto done;
                        out[idx].offset = off; out[idx].litLength = p - anchor;
                        out[idx].matchLength = end - p; out[idx].rep = 0;
                        idx++;
                    }
                    anchor = end; prevOff = off;
                    cur = end - segStart;
                }
                if (idx - firstIdx != fresh) goto done;          /* the counting walk and the scan must agree */
                if (cnt[j]) { runE = lastEnd[j]; runO = lastOff[j]; }
            }
            nOut = idx; anchorC = runE; prevOffC = runO;
        }
        if (nOut >= outCap) goto done;
        out[nOut].offset = 0; out[nOut].litLength = N - anchorC; out[nOut].matchLength = 0; out[nOut].rep = 0;
        nseq = nOut + 1;
    }
done:
    free(ownLen); free(ownOff); free(pk); free(link); free(gmax); free(gown); free(hasA); free(entA);
    return nseq;
}
