/*
 * cpu_bench.c — timed CPU baselines (TEST/BENCH INFRASTRUCTURE ONLY, never linked into the product).
 *
 * Methodology of the reference tool (/root/reference/test/benchmark.c:222-402): one private
 * CCtx per thread, barrier start, CLOCK_MONOTONIC around the work.  Differences, stated so the
 * numbers can be read correctly: (1) the blocks of the sample are PARTITIONED across threads
 * (thread t takes blocks t, t+T, ...), so the result is whole-job throughput over each input
 * byte once — directly comparable with the GPU's batch throughput — where the reference tool has
 * every thread compress the same buffer and sums per-thread rates; (2) wall time of the slowest
 * thread is used, not the sum of per-call times.
 *
 *   mode 0: ZSTD_compress2 per chunk, no producer          (benchmark -m0: full software compression)
 *   mode 1: per-block ZSTD_generateSequences               (software sequence producer only: the CPU
 *                                                           counterpart of qatSequenceProducer)
 */
#include "zstd_oracle.h"
#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

typedef struct {
    const unsigned char *src; size_t srcSize, chunk; int level, mode, tid, nThreads, iters;
    pthread_barrier_t *bar; double seconds; size_t outBytes, nSeq; int ok;
} Work;

static double now_s(void) { struct timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return t.tv_sec + 1e-9 * t.tv_nsec; }

static void *worker(void *arg)
{
    Work *w = (Work *)arg;
    const size_t nBlocks = (w->srcSize + w->chunk - 1) / w->chunk;
    const size_t cap = ZSTD_sequenceBound(w->chunk) + 1;
    ZSTD_CCtx *zc = ZSTD_createCCtx();
    ZSTD_Sequence *seqs = (ZSTD_Sequence *)malloc(cap * sizeof(ZSTD_Sequence));
    size_t dstCap = ZSTD_compressBound(w->chunk);
    void *dst = malloc(dstCap);
    w->ok = zc && seqs && dst && !ZSTD_isError(ZSTD_CCtx_setParameter(zc, ZSTD_c_compressionLevel, w->level));
    pthread_barrier_wait(w->bar);
    double t0 = now_s();
    if (w->ok) {
        for (int it = 0; it < w->iters; it++) {
            for (size_t b = (size_t)w->tid; b < nBlocks; b += (size_t)w->nThreads) {
                const unsigned char *p = w->src + b * w->chunk;
                size_t n = w->srcSize - b * w->chunk < w->chunk ? w->srcSize - b * w->chunk : w->chunk;
                size_t r = w->mode == 0 ? ZSTD_compress2(zc, dst, dstCap, p, n)
                                        : ZSTD_generateSequences(zc, seqs, cap, p, n);
                if (ZSTD_isError(r)) { w->ok = 0; break; }
                if (it == 0) { if (w->mode == 0) w->outBytes += r; else w->nSeq += r; }
            }
        }
    }
    w->seconds = now_s() - t0;
    free(dst); free(seqs); ZSTD_freeCCtx(zc);
    return NULL;
}

/* Returns throughput in bytes/second of raw input (srcSize * iters / slowest thread), or < 0.
 * outBytes: total compressed bytes (mode 0) ; nSeq: total sequences (mode 1), first iteration. */
double oracle_cpu_bench(const void *src, size_t srcSize, size_t chunk, int level, int mode,
                        int nThreads, int iters, size_t *outBytes, size_t *nSeq)
{
    if (nThreads < 1 || iters < 1 || chunk == 0 || srcSize == 0) return -1.0;
    pthread_t *th = (pthread_t *)calloc((size_t)nThreads, sizeof(pthread_t));
    Work *w = (Work *)calloc((size_t)nThreads, sizeof(Work));
    pthread_barrier_t bar;
    pthread_barrier_init(&bar, NULL, (unsigned)nThreads);
    for (int t = 0; t < nThreads; t++) {
        w[t].src = (const unsigned char *)src; w[t].srcSize = srcSize; w[t].chunk = chunk; w[t].level = level;
        w[t].mode = mode; w[t].tid = t; w[t].nThreads = nThreads; w[t].iters = iters; w[t].bar = &bar;
        pthread_create(&th[t], NULL, worker, &w[t]);
    }
    double slowest = 0; size_t ob = 0, ns = 0; int ok = 1;
    for (int t = 0; t < nThreads; t++) {
        pthread_join(th[t], NULL);
        if (w[t].seconds > slowest) slowest = w[t].seconds;
        ob += w[t].outBytes; ns += w[t].nSeq; ok &= w[t].ok;
    }
    pthread_barrier_destroy(&bar);
    free(th); free(w);
    if (outBytes) *outBytes = ob;
    if (nSeq) *nSeq = ns;
    if (!ok || slowest <= 0) return -1.0;
    return (double)srcSize * iters / slowest;
}
