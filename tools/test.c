/*
 * qzstd_test — functional round-trip check of the plugin, the B200 build's counterpart of
 * /root/reference/test/test.c:53-146: one file, ZSTD_compress2 with qatSequenceProducer registered
 * and software fallback enabled, ZSTD_decompress, memcmp.
 *
 * usage: qzstd_test <file> [level]
 * Unlike the reference program (which always exits 0, test.c:134-145) the exit status reflects the
 * result, and the number of blocks that fell back to software is printed: with fallback enabled a
 * dead device would otherwise pass silently.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "qatseqprod.h"

static unsigned char *slurp(const char *path, size_t *size)
{
    FILE *f = fopen(path, "rb");
    if (!f) return NULL;
    if (fseek(f, 0, SEEK_END) != 0) { fclose(f); return NULL; }
    long n = ftell(f);
    rewind(f);
    unsigned char *buf = (unsigned char *)malloc(n > 0 ? (size_t)n : 1);
    if (buf && n > 0 && fread(buf, 1, (size_t)n, f) != (size_t)n) { free(buf); buf = NULL; }
    fclose(f);
    *size = n > 0 ? (size_t)n : 0;
    return buf;
}

int main(int argc, char **argv)
{
    if (argc < 2) { printf("Usage: %s <file> [level]\n", argv[0]); return 1; }
    const int level = argc > 2 ? atoi(argv[2]) : 0;
    size_t srcSize = 0;
    unsigned char *src = slurp(argv[1], &srcSize);
    if (!src) { printf("Cannot read input file: %s\n", argv[1]); return 1; }

    int status = 1;
    const int dev = QZSTD_startQatDevice();
    void *state = QZSTD_createSeqProdState();
    ZSTD_CCtx *zc = ZSTD_createCCtx();
    const size_t dstCap = ZSTD_compressBound(srcSize);
    unsigned char *dst = (unsigned char *)malloc(dstCap ? dstCap : 1);
    unsigned char *back = (unsigned char *)malloc(srcSize ? srcSize : 1);
    if (!state || !zc || !dst || !back) { printf("Out of memory\n"); goto done; }

    printf("Plugin %s, device status %d (%s)\n", QZSTD_version(), dev,
           dev == QZSTD_OK ? "ready" : dev == QZSTD_STARTED ? "no capable device" : "no device");
    ZSTD_registerSequenceProducer(zc, state, qatSequenceProducer);
    if (ZSTD_isError(ZSTD_CCtx_setParameter(zc, ZSTD_c_enableSeqProducerFallback, 1))) {
        printf("Failed to set fallback\n");
        goto done;
    }
    if (level && ZSTD_isError(ZSTD_CCtx_setParameter(zc, ZSTD_c_compressionLevel, level))) {
        printf("Failed to set level\n");
        goto done;
    }

    const size_t cSize = ZSTD_compress2(zc, dst, dstCap, src, srcSize);
    if (ZSTD_isError(cSize)) { printf("Compress failed: %s\n", ZSTD_getErrorName(cSize)); goto done; }
    const size_t dSize = ZSTD_decompress(back, srcSize, dst, cSize);
    if (ZSTD_isError(dSize) || dSize != srcSize) { printf("Decompressed size is not equal to source size\n"); goto done; }
    if (memcmp(back, src, srcSize) != 0) { printf("ERROR: input and validation buffers don't match!\n"); goto done; }

    {
        unsigned long long calls = 0, errors = 0, batched = 0;
        QZSTD_getStats(state, &calls, &errors, &batched);
        printf("Compression and decompression were successful!\n");
        printf("Source size: %lu\n", (unsigned long)srcSize);
        printf("Compressed size: %lu\n", (unsigned long)cSize);
        printf("Producer calls: %llu, software fallbacks: %llu\n", calls, errors);
    }
    status = 0;
done:
    ZSTD_freeCCtx(zc);
    QZSTD_freeSeqProdState(state);
    QZSTD_stopQatDevice();
    free(src); free(dst); free(back);
    return status;
}
