/*
 * qzstd_producer_rate — raw input consumed by qatSequenceProducer (BASELINE.json's metric) through the stock
 * six-symbol surface: N threads, each with its own state (one state per CCtx, /root/reference/src/qatseqprod.h:139-151),
 * each walking the whole buffer block by block the way libzstd does inside ZSTD_compress2 - no hint, no additive
 * call, no entropy stage.  -m0 times the software sequence producer the same way (per-block ZSTD_generateSequences,
 * what runs when the plugin falls back).
 *
 *     qzstd_producer_rate [-t threads] [-l loops] [-L level] [-m mode] file
 */
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include "qatseqprod.h"

#define BLOCK ((size_t)ZSTD_BLOCKSIZE_MAX)

typedef struct {
    const unsigned char *src; size_t srcSize; int level, loops, mode, ok;
    pthread_barrier_t *start; unsigned long long seqs; double seconds;
} Job;

static double now_s(void)
{
    struct timespec t; clock_gettime(CLOCK_MONOTONIC, &t);
    return (double)t.tv_sec + 1e-9 * (double)t.tv_nsec;
}

static void *run(void *arg)
{
    Job *j = (Job *)arg;
    const size_t cap = ZSTD_sequenceBound(BLOCK);
    ZSTD_Sequence *out = (ZSTD_Sequence *)malloc(cap * sizeof(ZSTD_Sequence));
    void *state = NULL;
    ZSTD_CCtx *zc = NULL;
    int loop;
    j->ok = out != NULL;
    if (j->mode == 1) { state = QZSTD_createSeqProdState(); j->ok = j->ok && state != NULL; }
    else { zc = ZSTD_createCCtx(); j->ok = j->ok && zc && !ZSTD_isError(ZSTD_CCtx_setParameter(zc, ZSTD_c_compressionLevel, j->level)); }
    pthread_barrier_wait(j->start);
    {
        const double t0 = now_s();
        for (loop = 0; loop < j->loops && j->ok; loop++) {
            size_t off;
            for (off = 0; off < j->srcSize; off += BLOCK) {
                const size_t n = j->srcSize - off < BLOCK ? j->srcSize - off : BLOCK;
                const size_t r = j->mode == 1
                    ? qatSequenceProducer(state, out, cap, j->src + off, n, NULL, 0, j->level, (size_t)1 << 17)
                    : ZSTD_generateSequences(zc, out, cap, j->src + off, n);
                if (r == ZSTD_SEQUENCE_PRODUCER_ERROR || (j->mode == 0 && ZSTD_isError(r))) { j->ok = 0; break; }
                j->seqs += r;
            }
        }
        j->seconds = now_s() - t0;
    }
    if (state) QZSTD_freeSeqProdState(state);
    if (zc) ZSTD_freeCCtx(zc);
    free(out);
    return NULL;
}

int main(int argc, char **argv)
{
    int threads = 16, loops = 2, level = 3, mode = 1, a, i, ok = 1;
    const char *path = NULL;
    for (a = 1; a < argc; a++) {
        if (argv[a][0] == '-' && argv[a][1] && argv[a][2]) {
            const int v = atoi(argv[a] + 2);
            switch (argv[a][1]) {
            case 't': threads = v; break;
            case 'l': loops = v; break;
            case 'L': level = v; break;
            case 'm': mode = v; break;
            default: fprintf(stderr, "unknown option %s\n", argv[a]); return 2;
            }
        } else path = argv[a];
    }
    if (!path || threads < 1 || threads > 128 || loops < 1 || (mode != 0 && mode != 1)) {
        fprintf(stderr, "Usage: %s [-t# -l# -L# -m#] filename\n", argv[0]);
        return 2;
    }
    FILE *fp = fopen(path, "rb");
    if (!fp) { fprintf(stderr, "Cannot open %s\n", path); return 1; }
    fseek(fp, 0, SEEK_END);
    const size_t srcSize = (size_t)ftell(fp);
    rewind(fp);
    unsigned char *src = (unsigned char *)malloc(srcSize ? srcSize : 1);
    if (!src || fread(src, 1, srcSize, fp) != srcSize) { fprintf(stderr, "Cannot read %s\n", path); return 1; }
    fclose(fp);
    if (mode == 1 && QZSTD_startQatDevice() != QZSTD_OK) { fprintf(stderr, "no usable device\n"); return 1; }

    Job *jobs = (Job *)calloc((size_t)threads, sizeof(Job));
    pthread_t *th = (pthread_t *)calloc((size_t)threads, sizeof(pthread_t));
    pthread_barrier_t start;
    double slowest = 0;
    unsigned long long seqs = 0;
    pthread_barrier_init(&start, NULL, (unsigned)threads);
    for (i = 0; i < threads; i++) {
        jobs[i].src = src; jobs[i].srcSize = srcSize; jobs[i].level = level; jobs[i].loops = loops; jobs[i].mode = mode;
        jobs[i].start = &start;
        pthread_create(&th[i], NULL, run, &jobs[i]);
    }
    for (i = 0; i < threads; i++) {
        pthread_join(th[i], NULL);
        ok = ok && jobs[i].ok;
        if (jobs[i].seconds > slowest) slowest = jobs[i].seconds;
        seqs += jobs[i].seqs;
    }
    printf("Producer rate: %d thread(s) x %d pass(es) over %zu bytes, mode %d, level %d: %.0f MB/s of raw input, %llu sequences, %s\n",
           threads, loops, srcSize, mode, level, slowest > 0 ? (double)threads * (double)loops * (double)srcSize / slowest / 1e6 : 0.0,
           seqs, ok ? "PASS" : "FAIL");
    if (mode == 1) QZSTD_stopQatDevice();
    free(src); free(jobs); free(th);
    return ok ? 0 : 1;
}
