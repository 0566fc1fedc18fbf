/*
 * zstd_abi.h — hand-declared subset of the stock libzstd (>= 1.5.4) ABI.
 *
 * The build image ships libzstd.so.1 (1.5.5) but no zstd.h.  Everything the
 * plugin, its oracle and its tools need from libzstd is declared here,
 * verified against /usr/lib/x86_64-linux-gnu/libzstd.so.1.5.5 (nm -D and
 * ZSTD_cParam_getBounds on every parameter id below).  If a real zstd.h was
 * included first (ZSTD_VERSION_MAJOR defined) this header is a no-op, so the
 * same sources build against a development install of libzstd unchanged.
 *
 * Link with:  -l:libzstd.so.1
 *
 * Reference call sites these declarations serve:
 *   /root/reference/test/test.c:66-123, /root/reference/test/benchmark.c:241-356,
 *   /root/reference/src/qatseqprod.h:42-45 (ZSTD_STATIC_LINKING_ONLY + zstd.h).
 */
#ifndef ZSTD_ABI_SUBSET_H
#define ZSTD_ABI_SUBSET_H

#ifndef ZSTD_VERSION_MAJOR   /* a real zstd.h wins when present */

#include <stddef.h>

#if defined(__cplusplus)
extern "C" {
#endif

typedef struct ZSTD_CCtx_s ZSTD_CCtx;
typedef struct ZSTD_DCtx_s ZSTD_DCtx;

/* 16-byte sequence record exchanged with ZSTD_registerSequenceProducer(). */
typedef struct {
    unsigned int offset;       /* match distance; 0 together with matchLength 0 = block delimiter */
    unsigned int litLength;    /* literals preceding the match */
    unsigned int matchLength;  /* >= 3 for a real match */
    unsigned int rep;          /* filled by ZSTD_generateSequences only; ignored on input */
} ZSTD_Sequence;

typedef struct {
    size_t error;
    int lowerBound;
    int upperBound;
} ZSTD_bounds;

typedef struct {
    unsigned windowLog, chainLog, hashLog, searchLog, minMatch, targetLength;
    int strategy;              /* 1 fast 2 dfast 3 greedy 4 lazy 5 lazy2 6 btlazy2 7 btopt 8 btultra 9 btultra2 */
} ZSTD_compressionParameters;

#define ZSTD_BLOCKSIZE_MAX            (1 << 17)
#define ZSTD_SEQUENCE_PRODUCER_ERROR  ((size_t)(-1))

typedef size_t (*ZSTD_sequenceProducer_F)(
    void *sequenceProducerState,
    ZSTD_Sequence *outSeqs, size_t outSeqsCapacity,
    const void *src, size_t srcSize,
    const void *dict, size_t dictSize,
    int compressionLevel, size_t windowSize);

/* ZSTD_cParameter ids (plain ints at the ABI). */
typedef enum {
    ZSTD_c_compressionLevel = 100,
    ZSTD_c_windowLog = 101,
    ZSTD_c_hashLog = 102,
    ZSTD_c_chainLog = 103,
    ZSTD_c_searchLog = 104,
    ZSTD_c_minMatch = 105,
    ZSTD_c_targetLength = 106,
    ZSTD_c_strategy = 107,
    ZSTD_c_enableLongDistanceMatching = 160,
    ZSTD_c_contentSizeFlag = 200,
    ZSTD_c_checksumFlag = 201,
    ZSTD_c_dictIDFlag = 202,
    ZSTD_c_nbWorkers = 400,
    ZSTD_c_blockDelimiters = 1008,            /* experimentalParam11 */
    ZSTD_c_validateSequences = 1009,          /* experimentalParam12 */
    ZSTD_c_enableSeqProducerFallback = 1014,  /* experimentalParam17 */
    ZSTD_c_maxBlockSize = 1015,               /* experimentalParam18 */
    ZSTD_c_searchForExternalRepcodes = 1016   /* experimentalParam19 */
} ZSTD_cParameter;

typedef enum { ZSTD_ps_auto = 0, ZSTD_ps_enable = 1, ZSTD_ps_disable = 2 } ZSTD_paramSwitch_e;
typedef enum { ZSTD_sf_noBlockDelimiters = 0, ZSTD_sf_explicitBlockDelimiters = 1 } ZSTD_sequenceFormat_e;
typedef enum { ZSTD_reset_session_only = 1, ZSTD_reset_parameters = 2,
               ZSTD_reset_session_and_parameters = 3 } ZSTD_ResetDirective;

const char *ZSTD_versionString(void);
unsigned    ZSTD_isError(size_t code);
const char *ZSTD_getErrorName(size_t code);
size_t      ZSTD_compressBound(size_t srcSize);

ZSTD_CCtx  *ZSTD_createCCtx(void);
size_t      ZSTD_freeCCtx(ZSTD_CCtx *cctx);
ZSTD_DCtx  *ZSTD_createDCtx(void);
size_t      ZSTD_freeDCtx(ZSTD_DCtx *dctx);

size_t      ZSTD_CCtx_setParameter(ZSTD_CCtx *cctx, ZSTD_cParameter param, int value);
size_t      ZSTD_CCtx_reset(ZSTD_CCtx *cctx, ZSTD_ResetDirective reset);
ZSTD_bounds ZSTD_cParam_getBounds(ZSTD_cParameter cParam);
ZSTD_compressionParameters ZSTD_getCParams(int compressionLevel,
                                           unsigned long long estimatedSrcSize, size_t dictSize);

size_t      ZSTD_compress2(ZSTD_CCtx *cctx, void *dst, size_t dstCapacity,
                           const void *src, size_t srcSize);
size_t      ZSTD_decompress(void *dst, size_t dstCapacity, const void *src, size_t compressedSize);
size_t      ZSTD_decompressDCtx(ZSTD_DCtx *dctx, void *dst, size_t dstCapacity,
                                const void *src, size_t srcSize);

void        ZSTD_registerSequenceProducer(ZSTD_CCtx *cctx, void *sequenceProducerState,
                                          ZSTD_sequenceProducer_F sequenceProducer);
size_t      ZSTD_sequenceBound(size_t srcSize);
size_t      ZSTD_generateSequences(ZSTD_CCtx *zc, ZSTD_Sequence *outSeqs, size_t outSeqsSize,
                                   const void *src, size_t srcSize);
size_t      ZSTD_mergeBlockDelimiters(ZSTD_Sequence *sequences, size_t seqsSize);
size_t      ZSTD_compressSequences(ZSTD_CCtx *cctx, void *dst, size_t dstSize,
                                   const ZSTD_Sequence *inSeqs, size_t inSeqsSize,
                                   const void *src, size_t srcSize);

#if defined(__cplusplus)
}
#endif

#endif /* !ZSTD_VERSION_MAJOR */
#endif /* ZSTD_ABI_SUBSET_H */
