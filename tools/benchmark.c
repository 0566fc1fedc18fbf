/*
 * qzstd_benchmark — multi-thread throughput / latency / ratio tool with the command line of the
 * reference's benchmark (/root/reference/test/benchmark.c:171-184):
 *     -t#  threads [1-128]          -l#  loops            -c#  chunk size (K/M suffix, default 32K)
 *     -E#  searchForExternalRepcodes 0 auto / 1 enable / 2 disable
 *     -L#  level [1-12]             -m#  0 software, 1 plugin (default 1)
 * plus  -B   hint the whole buffer to the plugin before each pass (QZSTD_hintSource; B200-only).
 *
 * Per thread, as in the reference (:222-402): private CCtx/DCtx/state, barrier start, every chunk is
 * its own ZSTD_compress2 frame timed with CLOCK_MONOTONIC, one ZSTD_decompress over the concatenated
 * frames + memcmp for PASS/FAIL, ratio = cSize/srcSize, MB = 1e6.  All threads compress the same
 * buffer; the tool prints per-thread rates, their sum, latency percentiles, and the number of chunks
 * that fell back to software (the reference cannot tell).
 */
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include "qatseqprod.h"

typedef struct {
    const unsigned char *src; size_t srcSize, chunk; int level, loops, mode, repcodes, hint, id;
    pthread_barrier_t *start;
    double compSec, decompSec; size_t cSize; int pass; unsigned long long calls, errors;
    unsigned long long *lat; size_t nLat;
} Job;

static unsigned long long ns_now(void)
{
    struct timespec t; clock_gettime(CLOCK_MONOTONIC, &t);
    return (unsigned long long)t.tv_sec * 1000000000ull + (unsigned long long)t.tv_nsec;
}

static size_t parse_size(const char *s)
{
    char *end; unsigned long v = strtoul(s, &end, 10);
    if (*end == 'K' || *end == 'k') v <<= 10; else if (*end == 'M' || *end == 'm') v <<= 20;
    return v;
}

static void *run(void *arg)
{
    Job *j = (Job *)arg;
    const size_t nChunks = (j->srcSize + j->chunk - 1) / j->chunk;
    ZSTD_CCtx *zc = ZSTD_createCCtx();
    ZSTD_DCtx *zd = ZSTD_createDCtx();
    const size_t dstCap = ZSTD_compressBound(j->chunk) * nChunks;
    unsigned char *dst = (unsigned char *)malloc(dstCap), *back = (unsigned char *)malloc(j->srcSize);
    size_t *cs = (size_t *)calloc(nChunks, sizeof(size_t));
    void *state = NULL;
    int ok = zc && zd && dst && back && cs;
    j->lat = (unsigned long long *)calloc(nChunks * (size_t)j->loops, sizeof(unsigned long long));
    if (ok && j->mode == 1) {
        QZSTD_startQatDevice();                 /* idempotent; every thread calls it, like the reference (:262) */
        state = QZSTD_createSeqProdState();
        ZSTD_registerSequenceProducer(zc, state, qatSequenceProducer);
        ok = state != NULL && !ZSTD_isError(ZSTD_CCtx_setParameter(zc, ZSTD_c_enableSeqProducerFallback, 1));
    }
    ok = ok && !ZSTD_isError(ZSTD_CCtx_setParameter(zc, ZSTD_c_searchForExternalRepcodes, j->repcodes))
            && !ZSTD_isError(ZSTD_CCtx_setParameter(zc, ZSTD_c_compressionLevel, j->level));
    pthread_barrier_wait(j->start);
    if (ok) {
        for (int l = 0; l < j->loops && ok; l++) {
            size_t out = 0;
            if (state && j->hint) QZSTD_hintSource(state, j->src, j->srcSize, j->chunk);
            for (size_t c = 0; c < nChunks; c++) {
                const size_t n = j->srcSize - c * j->chunk < j->chunk ? j->srcSize - c * j->chunk : j->chunk;
                const unsigned long long t0 = ns_now();
                const size_t r = ZSTD_compress2(zc, dst + out, dstCap - out, j->src + c * j->chunk, n);
                const unsigned long long dt = ns_now() - t0;
                if (ZSTD_isError(r)) { fprintf(stderr, "Compress failed: %s\n", ZSTD_getErrorName(r)); ok = 0; break; }
                cs[c] = r; out += r;
                j->lat[j->nLat++] = dt;
                j->compSec += dt * 1e-9;
            }
            j->cSize = out;
        }
    }
    if (ok) {
        const size_t d = ZSTD_decompress(back, j->srcSize, dst, j->cSize);      /* concatenated frames (:323-329) */
        j->pass = !ZSTD_isError(d) && d == j->srcSize && memcmp(back, j->src, j->srcSize) == 0;
        for (int l = 0; l < j->loops; l++) {
            size_t in = 0, outp = 0;
            for (size_t c = 0; c < nChunks; c++) {
                const unsigned long long t0 = ns_now();
                const size_t d2 = ZSTD_decompressDCtx(zd, back + outp, j->srcSize - outp, dst + in, cs[c]);
                j->decompSec += (ns_now() - t0) * 1e-9;
                if (ZSTD_isError(d2)) break;
                in += cs[c]; outp += d2;
            }
        }
    }
    if (state) { QZSTD_getStats(state, &j->calls, &j->errors, NULL); QZSTD_freeSeqProdState(state); }
    ZSTD_freeCCtx(zc); ZSTD_freeDCtx(zd); free(dst); free(back); free(cs);
    return NULL;
}

static int cmp_u64(const void *a, const void *b)
{
    const unsigned long long x = *(const unsigned long long *)a, y = *(const unsigned long long *)b;
    return x < y ? -1 : x > y;
}

int main(int argc, char **argv)
{
    int threads = 1, loops = 1, level = 1, mode = 1, repcodes = 0, hint = 0;
    size_t chunk = 32 * 1024;
    const char *path = NULL;
    for (int i = 1; i < argc; i++) {
        const char *a = argv[i];
        if (a[0] != '-') { path = a; continue; }
        switch (a[1]) {
        case 't': threads = atoi(a + 2); break;
        case 'l': loops = atoi(a + 2); break;
        case 'c': chunk = parse_size(a + 2); break;
        case 'E': repcodes = atoi(a + 2); break;
        case 'L': level = atoi(a + 2); break;
        case 'm': mode = atoi(a + 2); break;
        case 'B': hint = 1; break;
        default:
            fprintf(stderr, "Usage: %s [-t# -l# -c# -E# -L# -m# -B] filename\n", argv[0]);
            return a[1] == 'h' || a[1] == 'H' ? 0 : 1;
        }
    }
    if (!path || threads < 1 || threads > 128 || loops < 1 || chunk == 0 || level < 1 || level > 12) {
        fprintf(stderr, "Usage: %s [-t# -l# -c# -E# -L# -m# -B] filename\n", argv[0]);
        return 1;
    }
    FILE *f = fopen(path, "rb");
    if (!f) { fprintf(stderr, "Cannot open %s\n", path); return 1; }
    fseek(f, 0, SEEK_END); long sz = ftell(f); rewind(f);
    unsigned char *src = (unsigned char *)malloc(sz > 0 ? (size_t)sz : 1);
    if (sz <= 0 || fread(src, 1, (size_t)sz, f) != (size_t)sz) { fprintf(stderr, "Cannot read %s\n", path); return 1; }
    fclose(f);

    pthread_barrier_t start;
    pthread_barrier_init(&start, NULL, (unsigned)threads);
    Job *jobs = (Job *)calloc((size_t)threads, sizeof(Job));
    pthread_t *th = (pthread_t *)calloc((size_t)threads, sizeof(pthread_t));
    for (int t = 0; t < threads; t++) {
        Job j = { src, (size_t)sz, chunk, level, loops, mode, repcodes, hint, t, &start, 0, 0, 0, 0, 0, 0, NULL, 0 };
        jobs[t] = j;
        pthread_create(&th[t], NULL, run, &jobs[t]);
    }
    double sum = 0; size_t nLat = 0; int allPass = 1; unsigned long long fallbacks = 0;
    for (int t = 0; t < threads; t++) {
        pthread_join(th[t], NULL);
        const Job *j = &jobs[t];
        const double comp = j->compSec > 0 ? (double)sz * loops / j->compSec / 1e6 : 0;
        const double decomp = j->decompSec > 0 ? (double)sz * loops / j->decompSec / 1e6 : 0;
        fprintf(stderr, "Thread %d: Compression: %ld -> %lu, Throughput: Comp: %5.f MB/s, Decomp: %5.f MB/s, "
                        "Compression Ratio: %2.2f%%, %s\n", t + 1, sz, (unsigned long)j->cSize, comp, decomp,
                100.0 * (double)j->cSize / (double)sz, j->pass ? "PASS" : "FAIL");
        sum += comp; nLat += j->nLat; allPass &= j->pass; fallbacks += j->errors;
    }
    unsigned long long *all = (unsigned long long *)malloc((nLat ? nLat : 1) * sizeof(unsigned long long));
    size_t k = 0; double total = 0;
    for (int t = 0; t < threads; t++) for (size_t i = 0; i < jobs[t].nLat; i++) { all[k++] = jobs[t].lat[i]; total += (double)jobs[t].lat[i]; }
    qsort(all, k, sizeof(unsigned long long), cmp_u64);
    if (k) fprintf(stderr, "Latency per chunk (us): P25 %.1f  P50 %.1f  P75 %.1f  P99 %.1f  avg %.1f\n",
                   all[k / 4] / 1e3, all[k / 2] / 1e3, all[(3 * k) / 4] / 1e3, all[(99 * k) / 100] / 1e3, total / (double)k / 1e3);
    fprintf(stderr, "Total: %d thread(s), %.0f MB/s aggregate, mode %d, level %d, chunk %lu, -E%d, software fallbacks: %llu, %s\n",
            threads, sum, mode, level, (unsigned long)chunk, repcodes, fallbacks, allPass ? "PASS" : "FAIL");
    if (mode == 1) QZSTD_stopQatDevice();
    return allPass ? 0 : 1;
}
