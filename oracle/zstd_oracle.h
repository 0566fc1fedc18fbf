/*
 * zstd_oracle.h — the reference-side oracle for the sequence-producer hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is linked into, imported by or executed
 * from the product.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs use it, and only as the checker or as the timed CPU baseline.
 *
 * The reference (intel/QAT-ZSTD-Plugin @ 898a67c) does its match finding in closed QAT
 * hardware behind cpaDcCompressData2 (/root/reference/src/qatseqprod.c:1245-1249) and cannot be
 * compiled here (needs cpa.h, cpa_dc.h, icp_sal_*.h, qae_mem.h, libqat, libusdm and QAT 4xxx
 * silicon; /root/reference/src/Makefile:50-60).  The software path BASELINE.json names as the
 * yard-stick ("the reference plugin's software-fallback path") is stock libzstd's own match
 * finder, reached when the producer returns ZSTD_SEQUENCE_PRODUCER_ERROR with
 * ZSTD_c_enableSeqProducerFallback=1 (/root/reference/test/test.c:109) or with benchmark -m0
 * (/root/reference/test/benchmark.c:265-267).  libzstd is a third-party dependency that is not
 * vendored under /root/reference (spec: libzstd-devel, unpinned, >= 1.5.4 per README.md:35);
 * the image carries libzstd.so.1.5.5, which this oracle calls directly.
 *
 * PARITY PIN: the reference holds no golden vectors for this path (its tests pin only the
 * lossless round trip, /root/reference/test/test.c:123-136).  The oracle is pinned instead by
 *   (1) the round trip itself through stock libzstd, and
 *   (2) the 13 producer-contract cases of SURVEY.md App. B, replayed in tests/test_contract.py,
 *   (3) tests/golden/lz4s_*.bin: LZ4s streams decoded by oracle_declz4s, which restates
 *       QZSTD_decLz4s line by line.
 * Sequence-level and ratio parity are unpinned by the reference; the +-1 % ratio bar comes from
 * BASELINE.json and is measured against oracle_chunked_compress in the same run.
 */
#ifndef B200_ZSTD_ORACLE_H
#define B200_ZSTD_ORACLE_H

#include <stddef.h>
#include <stdint.h>
#include "zstd_abi.h"

#if defined(__cplusplus)
extern "C" {
#endif

/* ---- software sequence producer: per-block ZSTD_generateSequences on a private CCtx.
 * This is "the software seq-producer on CPU" of BASELINE.json config #1 and the CPU counterpart
 * of qatSequenceProducer (C1 in BASELINE.md).  Same signature as ZSTD_sequenceProducer_F. */
void  *oracle_sw_create(void);
void   oracle_sw_free(void *state);
size_t oracle_sw_producer(void *state, ZSTD_Sequence *outSeqs, size_t outSeqsCapacity,
                          const void *src, size_t srcSize, const void *dict, size_t dictSize,
                          int compressionLevel, size_t windowSize);

/* ---- chunked stock compression: each chunk its own frame, no producer registered.
 * Semantics of `benchmark -m0 -c<chunk> -L<level>` (/root/reference/test/benchmark.c:300-321).
 * Returns total compressed bytes or (size_t)-1.  If dst is NULL a scratch buffer is used. */
size_t oracle_chunked_compress(const void *src, size_t srcSize, size_t chunkSize, int level,
                               void *dst, size_t dstCapacity);

/* ---- chunked compression THROUGH a registered producer (benchmark -m1 semantics,
 * /root/reference/test/benchmark.c:261-264,300-321).  repcodeMode = ZSTD_ps_auto/enable/disable
 * (-E0/-E1/-E2).  With fallback=1 a producer error silently becomes a software parse, so the
 * wrapper counts how many producer calls returned an error (*nErrors) and how many calls were
 * made (*nCalls).  The frames are decompressed and memcmp'd: *roundTripOk = 1 on success
 * (/root/reference/test/benchmark.c:323-339). */
size_t oracle_compress_with_producer(const void *src, size_t srcSize, size_t chunkSize, int level,
                                     ZSTD_sequenceProducer_F producer, void *producerState,
                                     int repcodeMode, int fallback, int validateSequences,
                                     size_t *nCalls, size_t *nErrors, int *roundTripOk);

/* ---- sequence validator: replays the sequences against src.  libzstd does not check that
 * matches are true (SURVEY.md App. B cases 5, 11), so tests must.
 * Returns 0 when valid, otherwise a negative code:
 *  -1 count is 0 / exceeds capacity      -2 matchLength in {1,2}          -3 offset == 0 with matchLength > 0
 *  -4 offset reaches before block start  -5 match bytes differ            -6 sum(lit+ml) != srcSize
 *  -7 delimiter (ml == 0) before the last entry */
int oracle_validate_sequences(const void *src, size_t srcSize, const ZSTD_Sequence *seqs,
                              size_t nbSeqs, size_t *firstBad);

/* ---- LZ4s intermediate format.
 * oracle_declz4s restates QZSTD_decLz4s (/root/reference/src/qatseqprod.c:1013-1091) including
 * its quirks: literal-only tokens are folded into the next sequence, matchLength is truncated to
 * 16 bits, capacity guard idx >= cap-1, stream must end exactly, returns count incl. last entry.
 * oracle_enclz4s is its inverse (sequences -> LZ4s token stream; literal bytes are zero-filled
 * because the decoder skips them). */
size_t oracle_declz4s(ZSTD_Sequence *outSeqs, size_t outSeqsCapacity,
                      const unsigned char *lz4sBuff, unsigned int lz4sBufSize);
size_t oracle_enclz4s(unsigned char *dst, size_t dstCapacity,
                      const ZSTD_Sequence *seqs, size_t nbSeqs);

/* ZSTD_compressSequences over an explicit-delimiter sequence array (the hand-off format of
 * QZSTD_generateSequences) + decompress + compare.  Returns the compressed size or (size_t)-1. */
size_t oracle_compress_sequences(const void *src, size_t srcSize, const ZSTD_Sequence *seqs, size_t nbSeqs,
                                 int level, int repcodeMode, int *roundTripOk);

#if defined(__cplusplus)
}
#endif
#endif
