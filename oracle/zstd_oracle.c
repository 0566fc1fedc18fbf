/*
 * zstd_oracle.c — reference-side oracle (see zstd_oracle.h for scope and the parity pin).
 * TEST INFRASTRUCTURE ONLY: never linked into the product.
 */
#include "zstd_oracle.h"

#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

/* ------------------------------------------------------------------------------------------
 * software sequence producer
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    ZSTD_CCtx *cctx;
    int level;
} SwState;

void *oracle_sw_create(void)
{
    SwState *s = (SwState *)calloc(1, sizeof(*s));
    if (!s) return NULL;
    s->cctx = ZSTD_createCCtx();
    s->level = 0x7fffffff;
    if (!s->cctx) { free(s); return NULL; }
    return s;
}

void oracle_sw_free(void *state)
{
    SwState *s = (SwState *)state;
    if (!s) return;
    ZSTD_freeCCtx(s->cctx);
    free(s);
}

size_t oracle_sw_producer(void *state, ZSTD_Sequence *outSeqs, size_t outSeqsCapacity,
                          const void *src, size_t srcSize, const void *dict, size_t dictSize,
                          int compressionLevel, size_t windowSize)
{
    SwState *s = (SwState *)state;
    (void)windowSize;
    if (!s || dict || dictSize) return ZSTD_SEQUENCE_PRODUCER_ERROR;
    if (s->level != compressionLevel) {
        if (ZSTD_isError(ZSTD_CCtx_setParameter(s->cctx, ZSTD_c_compressionLevel, compressionLevel)))
            return ZSTD_SEQUENCE_PRODUCER_ERROR;
        s->level = compressionLevel;
    }
    /* Each block is parsed with no history, exactly like the QAT engine (CPA_DC_STATELESS,
     * /root/reference/src/qatseqprod.c:941). */
    size_t n = ZSTD_generateSequences(s->cctx, outSeqs, outSeqsCapacity, src, srcSize);
    if (ZSTD_isError(n) || n == 0) return ZSTD_SEQUENCE_PRODUCER_ERROR;
    /* generateSequences ends every internal block with a {0, lastLits, 0} delimiter; a block
     * <= 128 KiB yields exactly one, at the end, which is the reference's output convention. */
    return n;
}

/* ------------------------------------------------------------------------------------------
 * chunked stock compression (benchmark -m0)
 * ---------------------------------------------------------------------------------------- */
size_t oracle_chunked_compress(const void *src, size_t srcSize, size_t chunkSize, int level,
                               void *dst, size_t dstCapacity)
{
    if (chunkSize == 0) return (size_t)-1;
    ZSTD_CCtx *zc = ZSTD_createCCtx();
    if (!zc) return (size_t)-1;
    size_t total = (size_t)-1;
    void *scratch = NULL;
    size_t scratchCap = 0;
    if (ZSTD_isError(ZSTD_CCtx_setParameter(zc, ZSTD_c_compressionLevel, level))) goto out;
    if (!dst) {
        scratchCap = ZSTD_compressBound(chunkSize < srcSize ? chunkSize : srcSize);
        scratch = malloc(scratchCap ? scratchCap : 1);
        if (!scratch) goto out;
    }
    total = 0;
    for (size_t pos = 0; pos < srcSize || (srcSize == 0 && pos == 0); pos += chunkSize) {
        size_t n = srcSize - pos < chunkSize ? srcSize - pos : chunkSize;
        size_t c;
        if (dst) c = ZSTD_compress2(zc, (char *)dst + total, dstCapacity - total, (const char *)src + pos, n);
        else     c = ZSTD_compress2(zc, scratch, scratchCap, (const char *)src + pos, n);
        if (ZSTD_isError(c)) { total = (size_t)-1; break; }
        total += c;
        if (srcSize == 0) break;
    }
out:
    free(scratch);
    ZSTD_freeCCtx(zc);
    return total;
}

/* ------------------------------------------------------------------------------------------
 * chunked compression through a registered producer (benchmark -m1) with error counting
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    ZSTD_sequenceProducer_F fn;
    void *state;
    size_t calls, errors;
} CountingState;

static size_t counting_producer(void *st, ZSTD_Sequence *outSeqs, size_t cap, const void *src,
                                size_t srcSize, const void *dict, size_t dictSize, int level,
                                size_t windowSize)
{
    CountingState *c = (CountingState *)st;
    size_t r = c->fn(c->state, outSeqs, cap, src, srcSize, dict, dictSize, level, windowSize);
    c->calls++;
    if (r == ZSTD_SEQUENCE_PRODUCER_ERROR) c->errors++;
    return r;
}

size_t oracle_compress_with_producer(const void *src, size_t srcSize, size_t chunkSize, int level,
                                     ZSTD_sequenceProducer_F producer, void *producerState,
                                     int repcodeMode, int fallback, int validateSequences,
                                     size_t *nCalls, size_t *nErrors, int *roundTripOk)
{
    if (nCalls) *nCalls = 0;
    if (nErrors) *nErrors = 0;
    if (roundTripOk) *roundTripOk = 0;
    if (chunkSize == 0) return (size_t)-1;

    CountingState cs = { producer, producerState, 0, 0 };
    ZSTD_CCtx *zc = ZSTD_createCCtx();
    size_t dstCap = ZSTD_compressBound(srcSize) + 64 * (srcSize / chunkSize + 1);
    unsigned char *dst = (unsigned char *)malloc(dstCap);
    unsigned char *back = (unsigned char *)malloc(srcSize ? srcSize : 1);
    size_t total = (size_t)-1;
    if (!zc || !dst || !back) goto out;

    ZSTD_registerSequenceProducer(zc, &cs, counting_producer);
    if (ZSTD_isError(ZSTD_CCtx_setParameter(zc, ZSTD_c_enableSeqProducerFallback, fallback))) goto out;
    if (ZSTD_isError(ZSTD_CCtx_setParameter(zc, ZSTD_c_searchForExternalRepcodes, repcodeMode))) goto out;
    if (ZSTD_isError(ZSTD_CCtx_setParameter(zc, ZSTD_c_validateSequences, validateSequences))) goto out;
    if (ZSTD_isError(ZSTD_CCtx_setParameter(zc, ZSTD_c_compressionLevel, level))) goto out;

    total = 0;
    for (size_t pos = 0; pos < srcSize; pos += chunkSize) {
        size_t n = srcSize - pos < chunkSize ? srcSize - pos : chunkSize;
        size_t c = ZSTD_compress2(zc, dst + total, dstCap - total, (const char *)src + pos, n);
        if (ZSTD_isError(c)) { total = (size_t)-1; goto out; }
        total += c;
    }
    /* one ZSTD_decompress over the concatenated frames, as the reference benchmark does */
    if (roundTripOk) {
        size_t d = srcSize ? ZSTD_decompress(back, srcSize, dst, total) : 0;
        *roundTripOk = (!ZSTD_isError(d) && d == srcSize && memcmp(back, src, srcSize) == 0);
    }
out:
    if (nCalls) *nCalls = cs.calls;
    if (nErrors) *nErrors = cs.errors;
    free(dst);
    free(back);
    ZSTD_freeCCtx(zc);
    return total;
}

/* ------------------------------------------------------------------------------------------
 * sequence validator
 * ---------------------------------------------------------------------------------------- */
int oracle_validate_sequences(const void *srcv, size_t srcSize, const ZSTD_Sequence *seqs,
                              size_t nbSeqs, size_t *firstBad)
{
    const unsigned char *src = (const unsigned char *)srcv;
    size_t pos = 0;
    size_t bad = 0;
    int rc = 0;
    if (nbSeqs == 0 || nbSeqs == (size_t)-1) { rc = -1; goto done; }
    for (size_t i = 0; i < nbSeqs; i++) {
        const ZSTD_Sequence s = seqs[i];
        bad = i;
        pos += s.litLength;
        if (pos > srcSize) { rc = -6; goto done; }
        if (s.matchLength == 0) {
            if (s.offset != 0) { rc = -3; goto done; }
            if (i + 1 != nbSeqs) { rc = -7; goto done; }
            continue;
        }
        if (s.matchLength < 3) { rc = -2; goto done; }
        if (s.offset == 0) { rc = -3; goto done; }
        if (s.offset > pos) { rc = -4; goto done; }
        if (pos + s.matchLength > srcSize) { rc = -6; goto done; }
        /* overlapping matches are legal: compare byte by byte against the already-known input */
        for (size_t k = 0; k < s.matchLength; k++) {
            if (src[pos + k] != src[pos + k - s.offset]) { rc = -5; goto done; }
        }
        pos += s.matchLength;
    }
    if (pos != srcSize) { rc = -6; bad = nbSeqs - 1; }
done:
    if (firstBad) *firstBad = bad;
    return rc;
}

/* ------------------------------------------------------------------------------------------
 * LZ4s token stream  <->  ZSTD_Sequence[]
 * ---------------------------------------------------------------------------------------- */
#define LZ4S_NIBBLE_MAX   15u
#define LZ4S_MATCH_BIAS   2u    /* LZ4MINMATCH, /root/reference/src/qatseqprod.c:104 */

/* Reads a nibble-escaped length: nibble, then while the nibble (or the last byte) saturates add
 * following bytes (/root/reference/src/qatseqprod.c:1027-1034, :1052-1059). */
static size_t lz4s_read_len(unsigned nibble, const unsigned char **ipp)
{
    size_t len = nibble;
    if (nibble == LZ4S_NIBBLE_MAX) {
        unsigned char b;
        do { b = *(*ipp)++; len += b; } while (b == 255);
    }
    return len;
}

size_t oracle_declz4s(ZSTD_Sequence *outSeqs, size_t outSeqsCapacity,
                      const unsigned char *lz4sBuff, unsigned int lz4sBufSize)
{
    const unsigned char *ip = lz4sBuff;
    const unsigned char *const end = lz4sBuff + lz4sBufSize;
    unsigned int pendingLits = 0;      /* literals of match-less tokens, folded forward (:1077-1084) */
    size_t idx = 0;

    while (lz4sBufSize > 0 && ip < end) {
        const unsigned token = *ip++;
        size_t lits = lz4s_read_len(token >> 4, &ip);
        ip += lits;                                    /* literal bytes are skipped (:1036) */
        if (ip == end) {                               /* final, literals-only entry (:1037-1045) */
            outSeqs[idx].litLength = (unsigned)(lits + pendingLits);
            outSeqs[idx].offset = 0;
            outSeqs[idx].matchLength = 0;
            break;
        }
        const unsigned offset = (unsigned)ip[0] | ((unsigned)ip[1] << 8);   /* LE16 (:1048) */
        ip += 2;
        size_t mcode = lz4s_read_len(token & 15u, &ip);
        if (mcode != 0) {
            outSeqs[idx].offset = offset;
            outSeqs[idx].litLength = (unsigned)(lits + pendingLits);
            outSeqs[idx].matchLength = (unsigned short)(mcode + LZ4S_MATCH_BIAS);  /* 16-bit truncation (:1062) */
            pendingLits = 0;
            idx++;
            if (idx >= outSeqsCapacity - 1) return ZSTD_SEQUENCE_PRODUCER_ERROR;   /* (:1073-1076) */
        } else if (lits > 0) {
            pendingLits += (unsigned)lits;
        }
    }
    if (ip != end) return ZSTD_SEQUENCE_PRODUCER_ERROR;                             /* (:1086-1089) */
    return idx + 1;                                                                /* (:1090) */
}

static size_t lz4s_put_len(unsigned char *dst, size_t pos, size_t cap, size_t rest)
{
    /* rest = len - 15, emitted as 255,255,...,r with r < 255 */
    for (;;) {
        if (pos >= cap) return (size_t)-1;
        if (rest >= 255) { dst[pos++] = 255; rest -= 255; }
        else { dst[pos++] = (unsigned char)rest; return pos; }
    }
}

size_t oracle_enclz4s(unsigned char *dst, size_t cap, const ZSTD_Sequence *seqs, size_t nbSeqs)
{
    size_t pos = 0;
    for (size_t i = 0; i < nbSeqs; i++) {
        const ZSTD_Sequence s = seqs[i];
        const int last = (i + 1 == nbSeqs);
        size_t mcode = 0;
        if (!last) {
            if (s.matchLength < 3 || s.offset == 0 || s.offset > 0xFFFF) return (size_t)-1;
            mcode = s.matchLength - LZ4S_MATCH_BIAS;
        } else if (s.matchLength != 0) {
            return (size_t)-1;                      /* stream must end on a literals-only token */
        }
        const unsigned ln = s.litLength >= LZ4S_NIBBLE_MAX ? LZ4S_NIBBLE_MAX : s.litLength;
        const unsigned mn = mcode >= LZ4S_NIBBLE_MAX ? LZ4S_NIBBLE_MAX : (unsigned)mcode;
        if (pos >= cap) return (size_t)-1;
        dst[pos++] = (unsigned char)((ln << 4) | mn);
        if (ln == LZ4S_NIBBLE_MAX) {
            pos = lz4s_put_len(dst, pos, cap, s.litLength - LZ4S_NIBBLE_MAX);
            if (pos == (size_t)-1) return pos;
        }
        if (pos + s.litLength > cap) return (size_t)-1;
        memset(dst + pos, 0, s.litLength);
        pos += s.litLength;
        if (last) break;
        if (pos + 2 > cap) return (size_t)-1;
        dst[pos++] = (unsigned char)(s.offset & 0xFF);
        dst[pos++] = (unsigned char)(s.offset >> 8);
        if (mn == LZ4S_NIBBLE_MAX) {
            pos = lz4s_put_len(dst, pos, cap, mcode - LZ4S_NIBBLE_MAX);
            if (pos == (size_t)-1) return pos;
        }
    }
    return pos;
}

/* Whole-buffer hand-off check: ZSTD_compressSequences over one ZSTD_Sequence array with explicit block
 * delimiters (what QZSTD_generateSequences produces), then ZSTD_decompress + memcmp.  Returns the
 * compressed size, or (size_t)-1 when libzstd rejects the sequences. */
size_t oracle_compress_sequences(const void *src, size_t srcSize, const ZSTD_Sequence *seqs, size_t nbSeqs,
                                 int level, int repcodeMode, int *roundTripOk)
{
    if (roundTripOk) *roundTripOk = 0;
    ZSTD_CCtx *zc = ZSTD_createCCtx();
    size_t dstCap = ZSTD_compressBound(srcSize) + 64 * (srcSize / (1u << 17) + 1);
    unsigned char *dst = (unsigned char *)malloc(dstCap);
    unsigned char *back = (unsigned char *)malloc(srcSize ? srcSize : 1);
    size_t c = (size_t)-1;
    if (!zc || !dst || !back) goto out;
    if (ZSTD_isError(ZSTD_CCtx_setParameter(zc, ZSTD_c_compressionLevel, level))) goto out;
    if (ZSTD_isError(ZSTD_CCtx_setParameter(zc, ZSTD_c_blockDelimiters, 1))) goto out;          /* explicit */
    if (ZSTD_isError(ZSTD_CCtx_setParameter(zc, ZSTD_c_validateSequences, 1))) goto out;
    if (ZSTD_isError(ZSTD_CCtx_setParameter(zc, ZSTD_c_searchForExternalRepcodes, repcodeMode))) goto out;
    c = ZSTD_compressSequences(zc, dst, dstCap, seqs, nbSeqs, src, srcSize);
    if (ZSTD_isError(c)) { c = (size_t)-1; goto out; }
    if (roundTripOk) {
        size_t d = ZSTD_decompress(back, srcSize, dst, c);
        *roundTripOk = (!ZSTD_isError(d) && d == srcSize && memcmp(back, src, srcSize) == 0);
    }
out:
    ZSTD_freeCCtx(zc);
    free(dst);
    free(back);
    return c;
}
