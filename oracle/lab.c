/*
 * lab.c — developer tool (TEST INFRASTRUCTURE ONLY): algorithm experiments for the level-scaled search.
 * A serial model with many knobs (environment LAB_*), run through stock libzstd (-E1) against same-level
 * chunked stock.  Nothing here ships; the variant that wins is frozen into seqmodel.c.
 *
 * usage: lab <file> <level> [chunk]
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <sys/stat.h>
#include "zstd_oracle.h"
#include "seqmodel.h"

typedef struct {
    int cb, depth, coll, useLong, lb, full, cap, rep, repMin, repBonus, lazy, window, minMatch, hb, probe, tieNear;
    int sb, sbytes, useShort, gainMode, backExt, evenOnly, skipBonus, repTie, repLong, seqRep, follow, followMin, harm, repK0, seqAbl; unsigned long long nFollow, nHarm;
    size_t nseq, nblocks, bad;
    unsigned long long steps, positions, litBytes, mlBytes, repHits, nMatch;
} Lab;

static inline uint32_t rd32(const uint8_t *p) { uint32_t v; memcpy(&v, p, 4); return v; }
static inline uint32_t floorlog2(uint32_t v) { return 31u - (uint32_t)__builtin_clz(v); }
static inline uint32_t hash_long(uint32_t lo, uint32_t hi, int bits) { return (lo * 0x9E3779B1u + hi * 0x85EBCA77u) >> (32 - bits); }
static int g_hbytes = 4;
static inline uint32_t hashN(uint32_t lo, uint32_t hi, int bits)
{
    uint32_t h;
    if (g_hbytes <= 4) h = lo * 0x9E3779B1u;
    else if (g_hbytes == 5) h = lo * 0x9E3779B1u + (hi & 0xFFu) * 0xC2B2AE3Du;
    else if (g_hbytes == 6) h = lo * 0x9E3779B1u + (hi & 0xFFFFu) * 0xC2B2AE3Du;
    else if (g_hbytes == 7) h = lo * 0x9E3779B1u + (hi & 0xFFFFFFu) * 0xC2B2AE3Du;
    else h = lo * 0x9E3779B1u + hi * 0x85EBCA77u;
    return h >> (32 - bits);
}
static inline int sameKey(const uint8_t *a, const uint8_t *b) { return memcmp(a, b, g_hbytes) == 0; }
static inline uint32_t hash_short(uint32_t lo, uint32_t hi, int bytes, int bits)
{
    uint32_t h;
    if (bytes <= 4)      h = lo * 0x9E3779B1u;
    else if (bytes == 5) h = lo * 0x9E3779B1u + (hi & 0xFFu) * 0xC2B2AE3Du;
    else                 h = lo * 0x9E3779B1u + (hi & 0xFFFFu) * 0xC2B2AE3Du;
    return h >> (32 - bits);
}

static uint32_t common(const uint8_t *a, const uint8_t *b, uint32_t lim)
{
    uint32_t l = 0;
    while (l < lim && a[l] == b[l]) l++;
    return l;
}

typedef struct { uint32_t end, off, rep; } BM;

static size_t lab_block(Lab *L, const uint8_t *src, size_t n, ZSTD_Sequence *out, size_t outCap)
{
    const uint32_t N = (uint32_t)n;
    const uint32_t nh = N >= 8 ? N - 7 : 0;
    uint32_t *ownLen = calloc(N + 1, 4), *ownOff = calloc(N + 1, 4), *ownRep = calloc(N + 1, 4);
    BM *B = calloc(N + 1, sizeof(BM));
    const size_t szC = (size_t)1 << L->cb, szL = (size_t)1 << L->lb, szS = (size_t)1 << L->sb;
    int32_t *head = malloc(szC * 4), *prev = malloc((N + 1) * 4), *tabL = malloc(szL * 4), *tabS = malloc(szS * 4);
    memset(head, 0xFF, szC * 4); memset(tabL, 0xFF, szL * 4); memset(tabS, 0xFF, szS * 4);
    const uint32_t PROBE = (uint32_t)L->probe;
    uint32_t runHead = 0;

    for (uint32_t p = 0; p < N; p++) {
        uint32_t bestLen = 0, bestOff = 0;
        int isFollower = 0;
        if (p < nh && L->depth > 0) { const uint32_t h0 = hashN(rd32(src + p), rd32(src + p + 4), L->cb); prev[p] = head[h0]; head[h0] = (int32_t)p; }
        const int cont = p > 0 && ownLen[p - 1] > (uint32_t)L->followMin;
        if (!cont) runHead = p;
        if (L->follow && cont && p < nh && p - runHead >= (uint32_t)L->follow) {
            /* follower: continues the previous position's match, no search of its own */
            bestLen = ownLen[p - 1] - 1; bestOff = ownOff[p - 1]; isFollower = 1;
            if (bestLen == (uint32_t)L->cap - 1) { /* capped head: re-extend */
                uint32_t lim = N - p; if (lim > (uint32_t)L->cap) lim = (uint32_t)L->cap;
                bestLen = common(src + p, src + p - bestOff, lim);
            }
            L->nFollow++;
        }
        if (p < nh && !isFollower) {
            const uint32_t lo = rd32(src + p), hi = rd32(src + p + 4);
            uint32_t lim = N - p;
            if (lim > (uint32_t)L->cap) lim = (uint32_t)L->cap;
            const uint32_t probe = lim < PROBE ? lim : PROBE;
            int32_t cands[4096]; int nc = 0;
            if (L->useLong) {
                const uint32_t h = hash_long(lo, hi, L->lb);
                if (tabL[h] >= 0) cands[nc++] = tabL[h];
                tabL[h] = (int32_t)p;
            }
            if (L->useShort) {
                const uint32_t h = hash_short(lo, hi, L->sbytes, L->sb);
                if (tabS[h] >= 0) cands[nc++] = tabS[h];
                tabS[h] = (int32_t)p;
            }
            if (L->depth > 0) {
                const uint32_t h = 0;
                int32_t q = prev[p]; (void)h;
                int steps = 0;
                while (q >= 0 && steps < L->depth) {
                    if (sameKey(src + q, src + p)) { cands[nc++] = q; steps++; }
                    else if (L->coll) steps++;
                    q = prev[q];
                    L->steps++;
                }
            }
            L->positions++;
            /* rank on first `probe` bytes (or full), longer wins, ties nearer */
            uint32_t bestRank = 0;
            for (int i = 0; i < nc; i++) {
                const uint32_t q = (uint32_t)cands[i];
                if (rd32(src + q) != lo) continue;
                uint32_t ml = L->full ? common(src + p, src + q, lim) : common(src + p, src + q, probe);
                const uint32_t off = p - q;
                int better;
                if (L->gainMode == 1) {   /* compare by zstd-like gain */
                    better = bestRank == 0 || (int)(ml * 4 - floorlog2(off + 1)) > (int)(bestRank * 4 - floorlog2(bestOff + 1));
                } else better = ml > bestRank || (ml == bestRank && off < bestOff);
                if (better) { bestRank = ml; bestOff = off; }
            }
            bestLen = bestRank;
            if (!L->full && bestLen == PROBE && PROBE < lim) bestLen = common(src + p, src + p - bestOff, lim);
            if (bestLen < (uint32_t)L->minMatch) bestLen = 0;
        }
        ownLen[p] = bestLen; ownOff[p] = bestOff;
    }

    /* backward extension by one: position p may adopt the match of p+1 shifted left if the byte before matches */
    if (L->backExt) {
        for (uint32_t p = 0; p + 1 < N; p++) {
            const uint32_t l1 = ownLen[p + 1], o1 = ownOff[p + 1];
            if (l1 && p >= o1 && src[p] == src[p - o1] && l1 + 1 > ownLen[p] && l1 + 1 <= (uint32_t)L->cap) { /* uses original values of p+1 */
                ownLen[p] = l1 + 1; ownOff[p] = o1;
            }
        }
    }

    /* rep probe: endOff[e] = offset of the longest own match ending exactly at e (e = first unmatched position);
     * position p probes the offsets of matches that ended at p-1 .. p-K+... (gap of 1..K literals) */
    if (L->rep) {
        const int K = L->rep;
        uint32_t *endLen = calloc(N + 2, 4), *endOff = calloc(N + 2, 4);
        for (uint32_t p = 0; p < N; p++) {
            if (!ownLen[p] || ownLen[p] < (uint32_t)L->repLong) continue;
            const uint32_t e = p + ownLen[p];
            if (e <= N && ownLen[p] > endLen[e]) { endLen[e] = ownLen[p]; endOff[e] = ownOff[p]; }
        }
        for (uint32_t p = 1; p < N; p++) {
            uint32_t lim = N - p;
            if (lim > (uint32_t)L->cap) lim = (uint32_t)L->cap;
            uint32_t tried[16]; int nt = 0;
            uint32_t bestMl = 0, bestR = 0;
            for (int k = L->repK0; k <= K && (uint32_t)k <= p; k++) {
                const uint32_t x = endOff[p - k];       /* a match ended at p-k: k literals, then us */
                if (!x || x > p) continue;
                int dup = 0; for (int t = 0; t < nt; t++) if (tried[t] == x) dup = 1;
                if (dup) continue;
                tried[nt++] = x;
                const uint32_t ml = common(src + p, src + p - x, lim);
                if (ml >= (uint32_t)L->repMin && ml > bestMl) { bestMl = ml; bestR = x; }
            }
            if (!bestMl) continue;
            int take;
            if (!ownLen[p]) take = 1;
            else if (ownOff[p] == bestR) take = 1;
            else take = (int)(bestMl * 4 + L->repBonus) >= (int)(ownLen[p] * 4 - floorlog2(ownOff[p] + 1));
            if (take) { ownLen[p] = bestMl; ownOff[p] = bestR; ownRep[p] = 1; }
        }
        free(endLen); free(endOff);
    }

    BM run = {0, 0, 0};
    for (uint32_t p = 0; p < N; p++) {
        if (ownLen[p] && (p + ownLen[p] > run.end || (L->repTie && p + ownLen[p] == run.end && ownRep[p] && !run.rep))) { run.end = p + ownLen[p]; run.off = ownOff[p]; run.rep = ownRep[p]; }
        B[p] = run;
    }

    if (L->seqRep) {
        /* zstd-lazy-like serial parse over own[] with true rep offsets (upper bound experiment) */
        size_t ns = 0; uint32_t anchorS = 0, ip = 0, rep1 = 0, rep2 = 0;
        const uint32_t capL = (uint32_t)L->cap;
        #define LIM(p_) ((N - (p_)) < capL ? (N - (p_)) : capL)
        while (ip + 8 <= N) {
            uint32_t ml = 0, off = 0, start = ip; int isRep = 0;
            /* rep at ip+1 */
            if (!(L->seqAbl & 1) && rep1 && ip + 1 >= rep1 && ip + 1 + 4 <= N && rd32(src + ip + 1) == rd32(src + ip + 1 - rep1)) {
                ml = common(src + ip + 1, src + ip + 1 - rep1, LIM(ip + 1)); off = rep1; start = ip + 1; isRep = 1;
            }
            if (ownLen[ip] > ml) { ml = ownLen[ip]; off = ownOff[ip]; start = ip; isRep = 0; }
            if (ml < 4) { ip += 1; continue; }
            int depth = L->lazy;
            while (depth >= 1 && ip + 9 <= N) {
                int moved = 0;
                for (int d = 1; d <= depth && !moved; d++) {
                    ip++;
                    if (ip + 8 > N) break;
                    if (!(L->seqAbl & 2) && off && rep1 && ip >= rep1 && rd32(src + ip) == rd32(src + ip - rep1)) {
                        const uint32_t mlRep = common(src + ip, src + ip - rep1, LIM(ip));
                        const int gain2 = (int)(mlRep * 3);
                        const int gain1 = (int)(ml * 3) - (int)(isRep ? 0 : floorlog2(off + 1)) + 1;
                        if (mlRep >= 4 && gain2 > gain1) { ml = mlRep; off = rep1; start = ip; isRep = 1; }
                    }
                    if (ownLen[ip]) {
                        const uint32_t ml2 = ownLen[ip], off2 = ownOff[ip];
                        const int gain2 = (int)(ml2 * 4) - (int)floorlog2(off2 + 1);
                        const int gain1 = (int)(ml * 4) - (int)(isRep ? 0 : floorlog2(off + 1)) + (d == 1 ? 4 : 7);
                        if (ml2 >= 4 && gain2 > gain1) { ml = ml2; off = off2; start = ip; isRep = 0; moved = 1; }
                    }
                }
                if (!moved) break;
            }
            /* catch up */
            if (!(L->seqAbl & 4) && !isRep) while (start > anchorS && start > off && src[start - 1] == src[start - 1 - off]) { start--; ml++; }
            if (!isRep) { rep2 = rep1; rep1 = off; }
            {
                const uint32_t lit = start - anchorS;
                if (lit == 0 && ns > 0 && out[ns - 1].offset == off) out[ns - 1].matchLength += ml;
                else { out[ns].offset = off; out[ns].litLength = lit; out[ns].matchLength = ml; out[ns].rep = 0; ns++; }
            }
            ip = anchorS = start + ml;
            /* immediate rep2 */
            while (!(L->seqAbl & 8) && ip + 4 <= N && rep2 && ip >= rep2 && rd32(src + ip) == rd32(src + ip - rep2)) {
                const uint32_t m2 = common(src + ip, src + ip - rep2, LIM(ip));
                if (m2 < 4) break;
                { uint32_t t = rep2; rep2 = rep1; rep1 = t; }
                out[ns].offset = rep1; out[ns].litLength = 0; out[ns].matchLength = m2; out[ns].rep = 0; ns++;
                ip = anchorS = ip + m2;
            }
        }
        out[ns].offset = 0; out[ns].litLength = N - anchorS; out[ns].matchLength = 0; out[ns].rep = 0; ns++;
        free(ownLen); free(ownOff); free(ownRep); free(B); free(head); free(prev); free(tabL); free(tabS);
        return ns;
    }
    const uint32_t minMatch = (uint32_t)(L->rep && L->repMin < L->minMatch ? L->repMin : L->minMatch);
    const uint32_t W = (uint32_t)L->window;
    size_t nseq = 0;
    uint32_t anchor = 0, cursor = 0, prevOff = 0;
#define HAS(p_) ((p_) < N && B[p_].end >= (p_) + minMatch)
#define GAIN(p_) ((int32_t)((B[p_].end - (p_)) * 4u) - (int32_t)((B[p_].rep && L->skipBonus) ? 0 : floorlog2(B[p_].off + 1u)))
    while (cursor < N) {
        uint32_t p = cursor;
        while (p < N && !HAS(p)) p++;
        if (p >= N) break;
        for (;;) {
            if (L->lazy < 1) break;
            const int32_t g0 = GAIN(p);
            uint32_t q = p + 1;
            if ((q % W) == 0 || !HAS(q)) break;
            if (GAIN(q) > g0 + 4) { p = q; continue; }
            if (L->lazy < 2) break;
            q = p + 2;
            if ((q % W) == 0 || !HAS(q)) break;
            if (GAIN(q) > g0 + 7) { p = q; continue; }
            break;
        }
        const uint32_t len = B[p].end - p, off = B[p].off, lit = p - anchor;
        if (lit == 0 && nseq > 0 && off == prevOff) out[nseq - 1].matchLength += len;
        else {
            if (nseq + 1 >= outCap) { nseq = (size_t)-1; goto done; }
            out[nseq].offset = off; out[nseq].litLength = lit; out[nseq].matchLength = len; out[nseq].rep = 0;
            nseq++;
        }
        prevOff = off;
        cursor = anchor = p + len;
    }
    if (L->harm) {
        /* offset harmonisation: a sequence whose bytes also match at the previous sequence's (final) offset takes that offset */
        uint32_t pos = 0, pv = 0, pv2 = 0;
        for (size_t i = 0; i < nseq; i++) {
            pos += out[i].litLength;
            const uint32_t l = out[i].matchLength, o = out[i].offset;
            uint32_t cand[2] = { pv, pv2 };
            for (int c = 0; c < (L->harm >= 2 ? 2 : 1); c++) {
                const uint32_t x = cand[c];
                if (x && x != out[i].offset && x <= pos && common(src + pos, src + pos - x, l) >= l) { out[i].offset = x; L->nHarm++; break; }
            }
            (void)o;
            if (out[i].offset != pv) { pv2 = pv; pv = out[i].offset; }
            pos += l;
        }
        /* merge zero-literal same-offset neighbours created by the substitution */
        size_t w = 0;
        for (size_t i = 0; i < nseq; i++) {
            if (w > 0 && out[i].litLength == 0 && out[i].offset == out[w - 1].offset) out[w - 1].matchLength += out[i].matchLength;
            else out[w++] = out[i];
        }
        nseq = w;
    }
    if (nseq >= outCap) { nseq = (size_t)-1; goto done; }
    out[nseq].offset = 0; out[nseq].litLength = N - anchor; out[nseq].matchLength = 0; out[nseq].rep = 0;
    nseq++;
done:
    free(ownLen); free(ownOff); free(ownRep); free(B); free(head); free(prev); free(tabL); free(tabS);
    return nseq;
}

static int g_useModel = 0;
static SeqModelParams g_prm;
static size_t lab_producer(void *st, ZSTD_Sequence *out, size_t cap, const void *src, size_t n,
                           const void *dict, size_t dictSize, int level, size_t windowSize)
{
    Lab *m = (Lab *)st;
    (void)dict; (void)dictSize; (void)level; (void)windowSize;
    size_t r = g_useModel ? seqmodel_block((const uint8_t *)src, n, out, cap, &g_prm) : lab_block(m, (const uint8_t *)src, n, out, cap);
    if (r == (size_t)-1) return ZSTD_SEQUENCE_PRODUCER_ERROR;
    if (oracle_validate_sequences(src, n, out, r, NULL) != 0) m->bad++;
    if (getenv("LAB_DUMP") && (long)m->nblocks == atol(getenv("LAB_DUMP"))) {
        size_t pos = 0; long lo = atol(getenv("LAB_DUMP_LO")), hi = atol(getenv("LAB_DUMP_HI"));
        for (size_t i = 0; i < r; i++) { pos += out[i].litLength;
            if ((long)pos >= lo && (long)pos < hi) fprintf(stderr, "%6zu: lit %3u  off %6u ml %3u\n", pos, out[i].litLength, out[i].offset, out[i].matchLength);
            pos += out[i].matchLength; } }
    m->nseq += r; m->nblocks++;
    { uint32_t r0 = 1, r1 = 4, r2 = 8;
      for (size_t i = 0; i < r; i++) { m->litBytes += out[i].litLength; if (out[i].matchLength) { m->nMatch++; m->mlBytes += out[i].matchLength;
          uint32_t o = out[i].offset; if (o == r0 || o == r1 || o == r2) m->repHits++; if (o != r0) { if (o != r1) r2 = r1; r1 = r0; r0 = o; } } } }
    return r;
}

static int env_int(const char *name, int def) { const char *e = getenv(name); return (e && *e) ? atoi(e) : def; }

static size_t cached_ref(const char *path, long sz, int level, size_t chunk, const unsigned char *src)
{
    char key[512]; unsigned h = 5381; for (const char *c = path; *c; c++) h = h * 33 + (unsigned char)*c;
    mkdir("/tmp/labcache", 0755);
    snprintf(key, sizeof key, "/tmp/labcache/%08x_%ld_L%d_%zu", h, sz, level, chunk);
    FILE *f = fopen(key, "r"); size_t v = 0;
    if (f) { if (fscanf(f, "%zu", &v) == 1 && v) { fclose(f); return v; } fclose(f); }
    v = oracle_chunked_compress(src, sz, chunk, level, NULL, 0);
    f = fopen(key, "w"); if (f) { fprintf(f, "%zu\n", v); fclose(f); }
    return v;
}

int main(int argc, char **argv)
{
    if (argc < 3) { fprintf(stderr, "usage: %s <file> <level> [chunk]\n", argv[0]); return 2; }
    int level = atoi(argv[2]);
    size_t chunk = argc > 3 ? (size_t)atol(argv[3]) : 131072;
    FILE *f = fopen(argv[1], "rb");
    if (!f) { perror("open"); return 1; }
    fseek(f, 0, SEEK_END); long sz = ftell(f); fseek(f, 0, SEEK_SET);
    unsigned char *src = (unsigned char *)malloc(sz ? sz : 1);
    if (fread(src, 1, sz, f) != (size_t)sz) { perror("read"); return 1; }
    fclose(f);
    size_t ref = cached_ref(argv[1], sz, level, chunk, src);

    Lab L; memset(&L, 0, sizeof L);
    L.cb = env_int("LAB_CB", 17); L.depth = env_int("LAB_D", 8); L.coll = env_int("LAB_COLL", 1);
    L.useLong = env_int("LAB_LONG", 0); L.lb = env_int("LAB_LB", 14);
    L.useShort = env_int("LAB_SHORT", 0); L.sb = env_int("LAB_SB", 12); L.sbytes = env_int("LAB_SBYTES", 5);
    L.full = env_int("LAB_FULL", 0); L.cap = env_int("LAB_CAP", 256); L.probe = env_int("LAB_PROBE", 16);
    L.rep = env_int("LAB_REP", 0); L.repMin = env_int("LAB_REPMIN", 4); L.repBonus = env_int("LAB_REPBONUS", 0);
    L.lazy = env_int("LAB_LAZY", 2); L.window = env_int("LAB_WINDOW", 32); L.minMatch = env_int("LAB_MINMATCH", 4);
    L.gainMode = env_int("LAB_GAINMODE", 0); L.backExt = env_int("LAB_BACKEXT", 0); L.skipBonus = env_int("LAB_SKIPBONUS", 0); L.repTie = env_int("LAB_REPTIE", 0); L.repLong = env_int("LAB_REPLONG", 1); L.seqRep = env_int("LAB_SEQREP", 0); L.follow = env_int("LAB_FOLLOW", 0); g_hbytes = env_int("LAB_HBYTES", 4); L.harm = env_int("LAB_HARM", 0); L.repK0 = env_int("LAB_REPK0", 1); L.seqAbl = env_int("LAB_SEQABL", 0); L.followMin = env_int("LAB_FOLLOWMIN", 8);
    g_useModel = env_int("LAB_MODEL", 0);
    seqmodel_params_for_level(level, &g_prm);
    g_prm.keyBytes = env_int("MODEL_KEYBYTES", g_prm.keyBytes); g_prm.rank16 = env_int("MODEL_RANK16", g_prm.rank16); g_prm.scan = env_int("MODEL_SCAN", g_prm.scan);
    g_prm.lazyDepth = env_int("MODEL_LAZY", g_prm.lazyDepth); g_prm.backExt = env_int("MODEL_BACKEXT", g_prm.backExt);
    g_prm.minMatch = env_int("MODEL_MINMATCH", g_prm.minMatch); g_prm.repParse = env_int("MODEL_REPPARSE", g_prm.repParse); g_prm.domBias = env_int("MODEL_DOMBIAS", g_prm.domBias); g_prm.nearN = env_int("MODEL_NEARN", g_prm.nearN); g_prm.extCap = env_int("MODEL_EXTCAP", g_prm.extCap);
    size_t calls, errs; int ok;
    size_t c = oracle_compress_with_producer(src, sz, chunk, level, lab_producer, &L, ZSTD_ps_enable, 0, 1, &calls, &errs, &ok);
    const char *base = strrchr(argv[1], '/'); base = base ? base + 1 : argv[1];
    printf("L%-2d %-12s ref %9zu  lab %9zu  %+6.2f%%  rt=%d bad=%zu B/seq=%.1f steps/pos=%.2f lit=%.1f%% avgml=%.1f rep=%.1f%%\n", level, base, ref, c,
           100.0 * ((double)c / ref - 1), ok, L.bad, L.nseq ? (double)sz / L.nseq : 0.0,
           L.positions ? (double)L.steps / L.positions : 0.0, 100.0 * L.litBytes / (sz ? sz : 1), L.nMatch ? (double)L.mlBytes / L.nMatch : 0.0,
           L.nMatch ? 100.0 * L.repHits / L.nMatch : 0.0);
    free(src);
    return 0;
}
