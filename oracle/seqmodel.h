/*
 * seqmodel.h — serial CPU statement of the B200 match finder's semantics.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is linked into, imported by or executed
 * from the product (libqatseqprod.so / the qat-zstd-plugin_b200 package).  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs use it, and only
 * as the checker.
 *
 * What it pins: the sm_100a kernels in qat-zstd-plugin_b200/csrc/lz77_kernels.cu are specified
 * to emit, for every block, exactly the ZSTD_Sequence array this model emits (bit-exact), so a
 * GPU race or indexing bug shows up as a diff rather than as a slightly different but still
 * valid parse.  The model is NOT the reference's algorithm (the reference's match finder is
 * closed QAT hardware, /root/reference/src/qatseqprod.c:1245-1249); the reference-side oracle
 * is zstd_oracle.c.  The output convention follows QZSTD_decLz4s
 * (/root/reference/src/qatseqprod.c:1013-1091): matchLength >= 3 for real matches, the last
 * entry carries the trailing literals with offset = matchLength = 0, rep is left 0.
 */
#ifndef B200_SEQMODEL_H
#define B200_SEQMODEL_H

#include <stddef.h>
#include <stdint.h>
#include "zstd_abi.h"

#if defined(__cplusplus)
extern "C" {
#endif

#define SEQMODEL_BUCKET_BITS 13      /* buckets of the counting sort (key hash, top bits)                 */
#define SEQMODEL_TAG_BITS    15      /* further hash bits kept with every entry to filter candidates      */
#define SEQMODEL_BITMAP_BITS 491520u /* bits of the repeated-key detector (60 KiB of shared memory)               */
#define SEQMODEL_IDX_CAP     16383u  /* the insertion index travels in 14 bits: scan <= this                */

typedef struct {
    int keyBytes;     /* bytes hashed into the bucket key: 4, 5 or 6                                       */
    int scan;         /* bucket entries examined per position, most recent first (the level-scaled depth)  */
    int rank16;       /* 1: candidates are ranked on their first 16 bytes, only the winner is extended      */
    int minMatch;     /* shortest match the parser may emit (>= 3)                                         */
    int extCap;       /* per-position match length cap in bytes (<= 256)                                   */
    int lazyDepth;    /* 0 greedy, 1 lazy, 2 lazy2                                                         */
    int window;       /* lazy look-ahead never crosses a multiple of it (32 = one warp's group)            */
    int backExt;      /* 1: a position adopts the match of the next position (same 32-group) when it also holds one byte earlier */
    int repParse;     /* 0: greedy/lazy parse over the propagated matches B (fast classes, lanemodel.c);
                         1: serial repcode-aware lazy parse over the own matches (levels 5-12, one warp per block)  */
    int domBias;      /* fast classes: >= 0 enables the dominant-offset probe; the scan's winner must be longer by more than this */
    int nearN;        /* levels 5-12 (> 0): the entry tag is a hash of bytes 4..7, so a deep scan measures only candidates that
                         agree on eight bytes; the nearest nearN entries are measured whatever their tag */
} SeqModelParams;

/* Parameters the kernels use for a zstd compression level (1..12). */
void   seqmodel_params_for_level(int level, SeqModelParams *prm);

/* Parse one block.  Returns the number of entries written (>= 1, last one is the literal
 * tail), or (size_t)-1 if outCap is too small.  n <= 131072. */
size_t seqmodel_block(const uint8_t *src, size_t n, ZSTD_Sequence *out, size_t outCap,
                      const SeqModelParams *prm);

/* Step 0: 1 when the block holds no more repeated keys than chance produces; it is then emitted as one literal run. */
int    seqmodel_incompressible(const uint8_t *src, size_t n, const SeqModelParams *prm);
uint32_t seqmodel_chance_threshold(uint32_t nh);

/* Steps 1-2 of the model for every position p: ownLen[p] (0 = no match >= minMatch) and ownOff[p].
 * Arrays hold n entries.  Returns 0, or -1 on allocation failure. */
int    seqmodel_own_matches(const uint8_t *src, size_t n, const SeqModelParams *prm,
                            uint32_t *ownLen, uint32_t *ownOff);

/* Lane-level statement of the kernel's parse stage (lanemodel.c): same output as seqmodel_block, computed
 * the way the parse warps compute it (per-group packed prefix maxima, carries, memoised per-group walks
 * iterated to the serial fixed point, scans for anchors and output slots).  Test infrastructure. */
size_t lanemodel_block(const uint8_t *src, size_t n, ZSTD_Sequence *out, size_t outCap,
                       const SeqModelParams *prm);

#if defined(__cplusplus)
}
#endif
#endif
