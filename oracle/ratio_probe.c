/*
 * ratio_probe.c — developer tool (TEST INFRASTRUCTURE ONLY): compares, on one input file cut
 * into 128 KiB chunks, the compressed size through stock libzstd of
 *   ref    : ZSTD_compress2 per chunk, no producer               (benchmark -m0, the yard-stick)
 *   sw     : per-block ZSTD_generateSequences through the slot   (software seq-producer)
 *   model  : the serial model of the B200 match finder through the slot
 * and validates the model's sequences.  Model parameters can be overridden from the
 * environment (MODEL_KEYBYTES, MODEL_SCAN, MODEL_MINMATCH, MODEL_EXTCAP, MODEL_LAZY, MODEL_WINDOW,
 * MODEL_BACKEXT) for tuning experiments.
 *
 * usage: ratio_probe <file> [level=3] [chunk=131072]
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include "zstd_oracle.h"
#include "seqmodel.h"

typedef struct { SeqModelParams prm; size_t nseq, nblocks, bad; double secs; } ModelState;

static double now_s(void) { struct timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return t.tv_sec + 1e-9 * t.tv_nsec; }

static size_t model_producer(void *st, ZSTD_Sequence *out, size_t cap, const void *src, size_t n,
                             const void *dict, size_t dictSize, int level, size_t windowSize)
{
    ModelState *m = (ModelState *)st;
    (void)dict; (void)dictSize; (void)level; (void)windowSize;
    double t0 = now_s();
    size_t r = seqmodel_block((const uint8_t *)src, n, out, cap, &m->prm);
    m->secs += now_s() - t0;
    if (r == (size_t)-1) return ZSTD_SEQUENCE_PRODUCER_ERROR;
    if (oracle_validate_sequences(src, n, out, r, NULL) != 0) m->bad++;
    m->nseq += r; m->nblocks++;
    return r;
}

static void env_int(const char *name, int *v) { const char *e = getenv(name); if (e && *e) *v = atoi(e); }

int main(int argc, char **argv)
{
    if (argc < 2) { fprintf(stderr, "usage: %s <file> [level] [chunk]\n", argv[0]); return 2; }
    int level = argc > 2 ? atoi(argv[2]) : 3;
    size_t chunk = argc > 3 ? (size_t)atol(argv[3]) : 131072;
    FILE *f = fopen(argv[1], "rb");
    if (!f) { perror("open"); return 1; }
    fseek(f, 0, SEEK_END); long sz = ftell(f); fseek(f, 0, SEEK_SET);
    unsigned char *src = (unsigned char *)malloc(sz ? sz : 1);
    if (fread(src, 1, sz, f) != (size_t)sz) { perror("read"); return 1; }
    fclose(f);

    double t0 = now_s();
    size_t ref = oracle_chunked_compress(src, sz, chunk, level, NULL, 0);
    double tref = now_s() - t0;

    size_t calls, errs; int ok;
    void *sw = oracle_sw_create();
    size_t csw = oracle_compress_with_producer(src, sz, chunk, level, oracle_sw_producer, sw,
                                               ZSTD_ps_enable, 0, 1, &calls, &errs, &ok);
    oracle_sw_free(sw);
    printf("%-28s L%d  src %ld  ref %zu (%.3f%%, %.0f MB/s)  sw-slot(E1) %zu (%+.2f%%) rt=%d\n",
           argv[1], level, sz, ref, 100.0 * ref / sz, sz / tref / 1e6, csw, 100.0 * ((double)csw / ref - 1), ok);

    ModelState ms; memset(&ms, 0, sizeof ms);
    seqmodel_params_for_level(level, &ms.prm);
    env_int("MODEL_KEYBYTES", &ms.prm.keyBytes);   env_int("MODEL_SCAN", &ms.prm.scan);
    env_int("MODEL_MINMATCH", &ms.prm.minMatch);   env_int("MODEL_EXTCAP", &ms.prm.extCap);
    env_int("MODEL_LAZY", &ms.prm.lazyDepth);      env_int("MODEL_WINDOW", &ms.prm.window);
    env_int("MODEL_BACKEXT", &ms.prm.backExt);
    for (int e = 1; e >= 0; e--) {
        ms.nseq = ms.nblocks = ms.bad = 0; ms.secs = 0;
        size_t c = oracle_compress_with_producer(src, sz, chunk, level, model_producer, &ms,
                                                 e ? ZSTD_ps_enable : ZSTD_ps_auto, 0, 1, &calls, &errs, &ok);
        printf("   model(E%d) key%dB scan%d backext%d min%d cap%d lazy%d: %zu (%+.2f%% vs ref)  rt=%d errs=%zu bad=%zu  seq/blk=%.0f  B/seq=%.1f  model %.0f MB/s\n",
               e, ms.prm.keyBytes, ms.prm.scan, ms.prm.backExt, ms.prm.minMatch, ms.prm.extCap,
               ms.prm.lazyDepth, c, 100.0 * ((double)c / ref - 1), ok, errs, ms.bad,
               ms.nblocks ? (double)ms.nseq / ms.nblocks : 0.0, ms.nseq ? (double)sz / ms.nseq : 0.0,
               sz / ms.secs / 1e6);
    }
    free(src);
    return 0;
}
