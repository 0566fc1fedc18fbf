/*
 * qzstd_handoff — the step after the producer, multi-threaded: one GPU batch per part of the buffer
 * (QZSTD_generateSequencesIndexed: a dense ZSTD_Sequence[] with explicit block delimiters plus a per-block index),
 * then the host's entropy stage on N threads, one ZSTD_compressSequences() frame per range of blocks.  This is
 * step 4 of the reference's flow chart ("Compress Sequences API", docs/images/qatzstdplugin.png) done for a whole
 * buffer at once instead of one synchronous callback per block, with the caller pattern of the reference's
 * benchmark (/root/reference/test/benchmark.c:300-321: private CCtx per thread, ZSTD_c_searchForExternalRepcodes).
 *
 *     qzstd_handoff [-t threads] [-l loops] [-L level] [-f blocksPerFrame] [-p partMiB] file
 *
 * The buffer is cut into parts (default 64 MiB): while the worker threads entropy-code part k, the main thread has
 * the GPU parse part k+1 (two sequence arrays, ping-pong).  Frames are independent zstd frames written back to
 * back: the result is a valid zstd stream (checked with one ZSTD_decompress per frame + memcmp, outside the timed
 * region).  Prints throughput of the whole job (MB = 1e6 like the reference tool) and the compressed size.
 */
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include "qatseqprod.h"

#define BLOCK ((size_t)ZSTD_BLOCKSIZE_MAX)

typedef struct {
    /* the job of one part */
    const unsigned char *src; size_t srcSize;       /* the part */
    const ZSTD_Sequence *seqs; const size_t *index; /* its sequences and per-block index */
    size_t nBlocks, blocksPerFrame, nFrames;
    unsigned char *dst; size_t frameCap;            /* frame f is written at dst + f * frameCap */
    size_t *frameSize;                              /* compressed size of frame f (0 = failed) */
} Part;

typedef struct {
    pthread_mutex_t mu; pthread_cond_t cv;
    Part *part;             /* the part being encoded, NULL = none */
    size_t nextFrame, doneFrames;
    int quit, failed, level;
} Pool;

static double now_s(void)
{
    struct timespec t; clock_gettime(CLOCK_MONOTONIC, &t);
    return (double)t.tv_sec + 1e-9 * (double)t.tv_nsec;
}

static void *worker(void *arg)
{
    Pool *p = (Pool *)arg;
    ZSTD_CCtx *zc = ZSTD_createCCtx();
    int level = 0;
    pthread_mutex_lock(&p->mu);
    for (;;) {
        while (!p->quit && (p->part == NULL || p->nextFrame >= p->part->nFrames)) pthread_cond_wait(&p->cv, &p->mu);
        if (p->quit) break;
        {
            Part *pt = p->part;
            const size_t f = p->nextFrame++;
            const size_t b0 = f * pt->blocksPerFrame;
            const size_t b1 = b0 + pt->blocksPerFrame < pt->nBlocks ? b0 + pt->blocksPerFrame : pt->nBlocks;
            const size_t byte0 = b0 * BLOCK, byte1 = b1 * BLOCK < pt->srcSize ? b1 * BLOCK : pt->srcSize;
            size_t r;
            const int lv = p->level;
            pthread_mutex_unlock(&p->mu);
            if (lv != level) {          /* sticky parameters: set once per level */
                ZSTD_CCtx_reset(zc, ZSTD_reset_session_and_parameters);
                ZSTD_CCtx_setParameter(zc, ZSTD_c_compressionLevel, lv);
                ZSTD_CCtx_setParameter(zc, ZSTD_c_blockDelimiters, ZSTD_sf_explicitBlockDelimiters);
                ZSTD_CCtx_setParameter(zc, ZSTD_c_validateSequences, 0);
                ZSTD_CCtx_setParameter(zc, ZSTD_c_searchForExternalRepcodes, ZSTD_ps_enable);
                level = lv;
            }
            r = ZSTD_compressSequences(zc, pt->dst + f * pt->frameCap, pt->frameCap, pt->seqs + pt->index[b0],
                                       pt->index[b1] - pt->index[b0], pt->src + byte0, byte1 - byte0);
            pthread_mutex_lock(&p->mu);
            if (ZSTD_isError(r)) { fprintf(stderr, "frame %zu: %s\n", f, ZSTD_getErrorName(r)); p->failed = 1; r = 0; }
            pt->frameSize[f] = r;
            if (++p->doneFrames == pt->nFrames) pthread_cond_broadcast(&p->cv);
        }
    }
    pthread_mutex_unlock(&p->mu);
    ZSTD_freeCCtx(zc);
    return NULL;
}

static void pool_run(Pool *p, Part *pt)     /* posts a part; returns at once */
{
    pthread_mutex_lock(&p->mu);
    p->part = pt; p->nextFrame = 0; p->doneFrames = 0;
    pthread_cond_broadcast(&p->cv);
    pthread_mutex_unlock(&p->mu);
}

static void pool_wait(Pool *p)
{
    pthread_mutex_lock(&p->mu);
    while (p->part && p->doneFrames < p->part->nFrames) pthread_cond_wait(&p->cv, &p->mu);
    pthread_mutex_unlock(&p->mu);
}

int main(int argc, char **argv)
{
    int threads = 16, loops = 3, level = 3, a;
    size_t blocksPerFrame = 8, partBytes = (size_t)64 << 20;
    const char *path = NULL;
    for (a = 1; a < argc; a++) {
        if (argv[a][0] == '-' && argv[a][1] && argv[a][2]) {
            const long v = atol(argv[a] + 2);
            switch (argv[a][1]) {
            case 't': threads = (int)v; break;
            case 'l': loops = (int)v; break;
            case 'L': level = (int)v; break;
            case 'f': blocksPerFrame = (size_t)v; break;
            case 'p': partBytes = (size_t)v << 20; break;
            default: fprintf(stderr, "unknown option %s\n", argv[a]); return 2;
            }
        } else path = argv[a];
    }
    if (!path || threads < 1 || threads > 128 || loops < 1 || blocksPerFrame < 1 || partBytes < BLOCK) {
        fprintf(stderr, "Usage: %s [-t# -l# -L# -f# -p#] filename\n", argv[0]);
        return 2;
    }
    partBytes = partBytes / BLOCK * BLOCK;

    FILE *fp = fopen(path, "rb");
    if (!fp) { fprintf(stderr, "Cannot open %s\n", path); return 1; }
    fseek(fp, 0, SEEK_END);
    const size_t srcSize = (size_t)ftell(fp);
    rewind(fp);
    unsigned char *src = (unsigned char *)malloc(srcSize ? srcSize : 1);
    if (!src || fread(src, 1, srcSize, fp) != srcSize) { fprintf(stderr, "Cannot read %s\n", path); return 1; }
    fclose(fp);

    if (QZSTD_startQatDevice() != QZSTD_OK) { fprintf(stderr, "no usable device\n"); return 1; }
    void *state = QZSTD_createSeqProdState();

    const size_t partBlocks = partBytes / BLOCK;
    const size_t seqCap = ZSTD_sequenceBound(partBytes) + partBlocks + 16;
    const size_t frameCap = ZSTD_compressBound(blocksPerFrame * BLOCK);
    const size_t framesPerPart = (partBlocks + blocksPerFrame - 1) / blocksPerFrame;
    const size_t nParts = (srcSize + partBytes - 1) / partBytes;
    ZSTD_Sequence *seqs[2]; size_t *index[2];
    unsigned char *dst = (unsigned char *)malloc(nParts * framesPerPart * frameCap);
    size_t *frameSize = (size_t *)calloc(nParts * framesPerPart, sizeof(size_t));
    Part *parts = (Part *)calloc(nParts, sizeof(Part));
    Pool pool;
    pthread_t *th = (pthread_t *)calloc((size_t)threads, sizeof(pthread_t));
    int i, ok = 1;
    for (i = 0; i < 2; i++) {
        seqs[i] = (ZSTD_Sequence *)malloc(seqCap * sizeof(ZSTD_Sequence));
        index[i] = (size_t *)malloc((partBlocks + 1) * sizeof(size_t));
    }
    if (!state || !dst || !frameSize || !parts || !th || !seqs[0] || !seqs[1] || !index[0] || !index[1]) { fprintf(stderr, "out of memory\n"); return 1; }
    /* SVM-style: the device reads the input and writes the sequences in place (no staging copies) */
    const int pinned = QZSTD_registerBuffer(src, srcSize) == QZSTD_OK && QZSTD_registerBuffer(seqs[0], seqCap * sizeof(ZSTD_Sequence)) == QZSTD_OK &&
                       QZSTD_registerBuffer(seqs[1], seqCap * sizeof(ZSTD_Sequence)) == QZSTD_OK;
    if (!pinned) fprintf(stderr, "buffers not page-locked: sequence production goes through staging\n");
    memset(&pool, 0, sizeof pool);
    pthread_mutex_init(&pool.mu, NULL);
    pthread_cond_init(&pool.cv, NULL);
    pool.level = level;
    for (i = 0; i < threads; i++) pthread_create(&th[i], NULL, worker, &pool);

    double best = 1e30, gpuSeconds = 0;
    size_t cSize = 0;
    int loop;
    for (loop = 0; loop < loops && ok; loop++) {
        const double t0 = now_s();
        double tg = 0;
        size_t k;
        for (k = 0; k < nParts && ok; k++) {
            Part *pt = &parts[k];
            const size_t off = k * partBytes, bytes = srcSize - off < partBytes ? srcSize - off : partBytes;
            const double g0 = now_s();
            /* the GPU parses part k into the array the workers are NOT reading (they encode part k-1 meanwhile) */
            const size_t n = QZSTD_generateSequencesIndexed(state, seqs[k & 1], seqCap, src + off, bytes, 0, level,
                                                            index[k & 1], partBlocks + 1);
            tg += now_s() - g0;
            if (n == ZSTD_SEQUENCE_PRODUCER_ERROR) { fprintf(stderr, "sequence production failed\n"); ok = 0; break; }
            pool_wait(&pool);                       /* part k-1 is encoded: its array is free for part k+1 */
            pt->src = src + off; pt->srcSize = bytes; pt->seqs = seqs[k & 1]; pt->index = index[k & 1];
            pt->nBlocks = (bytes + BLOCK - 1) / BLOCK; pt->blocksPerFrame = blocksPerFrame;
            pt->nFrames = (pt->nBlocks + blocksPerFrame - 1) / blocksPerFrame;
            pt->dst = dst + k * framesPerPart * frameCap; pt->frameCap = frameCap;
            pt->frameSize = frameSize + k * framesPerPart;
            pool_run(&pool, pt);
        }
        pool_wait(&pool);
        {
            const double dt = now_s() - t0;
            if (dt < best) { best = dt; gpuSeconds = tg; }
        }
        if (pool.failed) ok = 0;
    }

    /* verification, outside the timed region: every frame decodes to its bytes */
    if (ok) {
        unsigned char *back = (unsigned char *)malloc(blocksPerFrame * BLOCK);
        size_t k, f;
        cSize = 0;
        for (k = 0; k < nParts && ok; k++)
            for (f = 0; f < parts[k].nFrames && ok; f++) {
                const size_t b0 = f * blocksPerFrame, byte0 = b0 * BLOCK;
                const size_t want = parts[k].srcSize - byte0 < blocksPerFrame * BLOCK ? parts[k].srcSize - byte0 : blocksPerFrame * BLOCK;
                const size_t r = ZSTD_decompress(back, blocksPerFrame * BLOCK, parts[k].dst + f * frameCap, parts[k].frameSize[f]);
                if (ZSTD_isError(r) || r != want || memcmp(back, parts[k].src + byte0, want) != 0) ok = 0;
                cSize += parts[k].frameSize[f];
            }
        free(back);
    }

    pthread_mutex_lock(&pool.mu);
    pool.quit = 1;
    pthread_cond_broadcast(&pool.cv);
    pthread_mutex_unlock(&pool.mu);
    for (i = 0; i < threads; i++) pthread_join(th[i], NULL);

    printf("Hand-off: %zu -> %zu (%.2f%%), %d entropy thread(s), level %d, %zu block(s) per frame, parts of %zu MiB: "
           "%.0f MB/s (best of %d; sequence production %.1f ms of %.1f ms), %s\n",
           srcSize, cSize, srcSize ? 100.0 * (double)cSize / (double)srcSize : 0.0, threads, level, blocksPerFrame,
           partBytes >> 20, (double)srcSize / best / 1e6, loops, 1e3 * gpuSeconds, 1e3 * best, ok ? "PASS" : "FAIL");

    if (pinned) { QZSTD_unregisterBuffer(src); QZSTD_unregisterBuffer(seqs[0]); QZSTD_unregisterBuffer(seqs[1]); }
    QZSTD_freeSeqProdState(state);
    QZSTD_stopQatDevice();
    free(src); free(dst); free(frameSize); free(parts); free(th);
    for (i = 0; i < 2; i++) { free(seqs[i]); free(index[i]); }
    return ok ? 0 : 1;
}
