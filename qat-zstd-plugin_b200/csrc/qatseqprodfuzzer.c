/*
 * qatseqprodfuzzer.c — the five hooks upstream zstd's fuzzers call when they are built with a
 * third-party sequence producer (zstd tests/fuzz/fuzz_third_party_seq_prod.h), bound to this
 * library.  Same role and same mapping as /root/reference/test/fuzzing/qatseqprodfuzzer.c:41-74;
 * kept out of libqatseqprod.so (separate object, libqatseqprodfuzzer.{a,so}) like the reference
 * keeps it out of its library (test/fuzzing/Makefile: ld -r into qatseqprodfuzzer.o).
 *
 * Building the fuzz targets themselves needs clang (libFuzzer) and an upstream zstd tree; neither
 * ships in this image, so only the adapter and its symbol/behaviour test are provided.
 */
#include "qatseqprod.h"

size_t FUZZ_seqProdSetup(void)
{
    /* 0 (QZSTD_OK) when a usable B200 is present; the fuzzers assert on 0 */
    return (size_t)QZSTD_startQatDevice();
}

size_t FUZZ_seqProdTearDown(void)
{
    return 0;       /* like the reference: the device stays up for the next fuzz input */
}

void *FUZZ_createSeqProdState(void)
{
    return QZSTD_createSeqProdState();
}

size_t FUZZ_freeSeqProdState(void *state)
{
    QZSTD_freeSeqProdState(state);
    return 0;
}

size_t FUZZ_thirdPartySeqProd(void *sequenceProducerState, ZSTD_Sequence *outSeqs, size_t outSeqsCapacity,
                              const void *src, size_t srcSize, const void *dict, size_t dictSize,
                              int compressionLevel, size_t windowSize)
{
    return qatSequenceProducer(sequenceProducerState, outSeqs, outSeqsCapacity, src, srcSize, dict, dictSize,
                               compressionLevel, windowSize);
}
