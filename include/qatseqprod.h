/*
 * qatseqprod.h — public surface of libqatseqprod.so, B200 edition.
 *
 * Source-compatible with the header of intel/QAT-ZSTD-Plugin v0.2.0
 * (/root/reference/src/qatseqprod.h:50-65 version + status, :72 QZSTD_version, :110-116
 * qatSequenceProducer, :130 QZSTD_startQatDevice, :137 QZSTD_stopQatDevice, :145
 * QZSTD_createSeqProdState, :151 QZSTD_freeSeqProdState): an application written against the
 * reference header recompiles and relinks against this library unchanged; the "QAT device" it
 * starts is a B200 GPU.
 *
 * Usage (identical to /root/reference/test/test.c:66-123):
 *     QZSTD_startQatDevice();
 *     void *state = QZSTD_createSeqProdState();
 *     ZSTD_registerSequenceProducer(cctx, state, qatSequenceProducer);
 *     ZSTD_CCtx_setParameter(cctx, ZSTD_c_enableSeqProducerFallback, 1);
 *     ZSTD_compress2(cctx, dst, dstCap, src, srcSize);
 *     QZSTD_freeSeqProdState(state);
 *     QZSTD_stopQatDevice();
 *
 * Limitations, same as the reference (/root/reference/src/qatseqprod.h:96-108): levels 1..12
 * only; no dictionaries; no long-distance matching; each block is parsed without history from
 * earlier blocks; ZSTD_c_nbWorkers > 0 is rejected by libzstd when a producer is registered
 * (use one CCtx + one state per thread instead).
 */
#if defined (__cplusplus)
extern "C" {
#endif

#ifndef QATSEQPROD_H
#define QATSEQPROD_H

#ifndef ZSTD_STATIC_LINKING_ONLY
#define ZSTD_STATIC_LINKING_ONLY
#endif
#if defined(__has_include)
#  if __has_include(<zstd.h>)
#    include <zstd.h>
#  else
#    include "zstd_abi.h"     /* hand-declared libzstd subset for images without zstd.h */
#  endif
#else
#  include "zstd.h"
#endif

#define QZSTD_VERSION          "0.2.0"
#define QZSTD_VERSION_MAJOR    0
#define QZSTD_VERSION_MINOR    2
#define QZSTD_VERSION_RELEASE  0
#define QZSTD_VERSION_NUMBER  (QZSTD_VERSION_MAJOR *100*100 + QZSTD_VERSION_MINOR *100 \
                                + QZSTD_VERSION_RELEASE)

typedef enum {
    QZSTD_OK = 0,           /* device ready */
    QZSTD_STARTED = 1,      /* CUDA driver up, but no device meets the requirements (sm_100, 227 KB smem) */
    QZSTD_FAIL = -1,        /* no CUDA driver / device */
    QZSTD_UNSUPPORTED = -2  /* kept for source compatibility; never returned (the reference folds it into STARTED) */
} QZSTD_Status_e;

/* Version string of the plugin API this library implements ("0.2.0"). */
const char *QZSTD_version(void);

/* Block-level sequence producer: hand this to ZSTD_registerSequenceProducer() together with a
 * state from QZSTD_createSeqProdState().  Parses one block (srcSize <= ZSTD_BLOCKSIZE_MAX) on
 * the GPU and writes its LZ77 sequences to outSeqs; the last entry carries the trailing
 * literals with offset = matchLength = 0.  Returns the number of entries or
 * ZSTD_SEQUENCE_PRODUCER_ERROR: when dict/dictSize is set, when windowSize < min(srcSize, 32 KiB),
 * when compressionLevel is outside 1..12, when no device is ready, or on any device error. */
size_t qatSequenceProducer(
    void *sequenceProducerState, ZSTD_Sequence *outSeqs, size_t outSeqsCapacity,
    const void *src, size_t srcSize,
    const void *dict, size_t dictSize,
    int compressionLevel,
    size_t windowSize
);

/* Process-wide, idempotent, thread-safe device start; QZSTD_OK / QZSTD_STARTED / QZSTD_FAIL. */
int QZSTD_startQatDevice(void);

/* Releases the process-wide device state; call after all compression is finished. */
void QZSTD_stopQatDevice(void);

/* One state per ZSTD_CCtx (and per thread); reusable across compressions. */
void *QZSTD_createSeqProdState(void);
void QZSTD_freeSeqProdState(void *sequenceProducerState);

/* ---- additive, optional (not in the reference) ------------------------------------------
 * Look-ahead hint: tells the state that the next ZSTD_compress2() on its CCtx will compress
 * [src, src + srcSize) in blocks of blockSize bytes (0 = 128 KiB).  The first producer call
 * inside that range parses ALL of its blocks in one GPU batch; later calls are served from the
 * cached result.  The buffer must stay unmodified until the compression returns.  Pass src = NULL
 * to drop the hint.  Without a hint every call is a batch of one block. */
void QZSTD_hintSource(void *sequenceProducerState, const void *src, size_t srcSize, size_t blockSize);

/* Counters of one state: producer calls, calls answered with an error (software fallback if the
 * application enabled it), calls served from a look-ahead batch. */
void QZSTD_getStats(const void *sequenceProducerState, unsigned long long *calls,
                    unsigned long long *errors, unsigned long long *batched);

/* Cross-thread coalescing (additive, optional, off by default; also switched on by the environment variable
 * QZSTD_COALESCE=1 at QZSTD_startQatDevice).  libzstd hands the producer one block per call; with many
 * application threads, each with its own CCtx and state, a process-wide dispatcher gathers whatever single-block
 * calls are pending and parses them in ONE GPU batch - the role the shared instance pool plays in the reference
 * (QZSTD_grabInstance).  Calls covered by a QZSTD_hintSource batch are not affected.  Returns the previous
 * setting.  Switch it before the compression threads start. */
int QZSTD_setCoalescing(int enable);

/* Whole-buffer counterpart of stock ZSTD_generateSequences(): parses ALL blocks of [src, src + srcSize)
 * (blockSize bytes each, 0 = 128 KiB, blocks independent of each other) in one GPU batch and writes one
 * ZSTD_Sequence array in which every block ends with its {0, trailing literals, 0} entry, i.e. the
 * explicit-block-delimiter format ZSTD_compressSequences() accepts with ZSTD_c_blockDelimiters =
 * ZSTD_sf_explicitBlockDelimiters.  This is step 4 of the reference's flow chart ("Compress Sequences
 * API", docs/images/qatzstdplugin.png) without one synchronous callback per block.
 * outSeqsCapacity >= ZSTD_sequenceBound(srcSize) + number of blocks always suffices.  Returns the number
 * of entries, or ZSTD_SEQUENCE_PRODUCER_ERROR under the same conditions as qatSequenceProducer
 * (level outside 1..12, no device, capacity too small, device error). */
size_t QZSTD_generateSequences(void *sequenceProducerState, ZSTD_Sequence *outSeqs, size_t outSeqsCapacity,
                               const void *src, size_t srcSize, size_t blockSize, int compressionLevel);

/* The same with an index: blockIndex[b] = position in outSeqs of block b's first entry, blockIndex[nBlocks] = the
 * number of entries (blockIndexCapacity >= nBlocks + 1, nBlocks = ceil(srcSize / blockSize)).  With it a caller cuts
 * the array at block boundaries without scanning it - e.g. to entropy-code ranges of blocks on several host threads,
 * one ZSTD_compressSequences() frame per range (tools/handoff.c; SURVEY 8f-1 "multi-thread host entropy stage"). */
size_t QZSTD_generateSequencesIndexed(void *sequenceProducerState, ZSTD_Sequence *outSeqs, size_t outSeqsCapacity,
                                      const void *src, size_t srcSize, size_t blockSize, int compressionLevel,
                                      size_t *blockIndex, size_t blockIndexCapacity);

/* Page-locks [ptr, ptr + size) for the device (additive): QZSTD_generateSequences*() and the look-ahead hint then move
 * the input and the ZSTD_Sequence array by DMA straight from / into the caller's memory - the counterpart of the
 * reference's SVM mode (/root/reference/src/qatseqprod.c:1222-1227).  Returns QZSTD_OK or QZSTD_FAIL. */
int QZSTD_registerBuffer(void *ptr, size_t size);
int QZSTD_unregisterBuffer(void *ptr);

#endif /* QATSEQPROD_H */

#if defined (__cplusplus)
}
#endif
