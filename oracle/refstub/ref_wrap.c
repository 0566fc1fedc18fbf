/*
 * refstub/ref_wrap.c — compiles the UNMODIFIED reference translation unit where it lies
 * (REF_SRC = /root/reference/src/qatseqprod.c) against the fake driver headers, and exposes its one
 * static arithmetic routine to the tests.  TEST INFRASTRUCTURE ONLY; outputs go to oracle/_ref/.
 */
#include REF_SRC

size_t ref_decLz4s(ZSTD_Sequence *outSeqs, size_t outSeqsCapacity, unsigned char *lz4sBuff, unsigned int lz4sBufSize)
{
    return QZSTD_decLz4s(outSeqs, outSeqsCapacity, lz4sBuff, lz4sBufSize);
}

int ref_initStatus(void) { return gProcess.qzstdInitStatus; }
