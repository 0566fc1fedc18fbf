/*
 * seqprod_host.c — the libzstd-facing plugin layer of libqatseqprod.so, plain C.
 *
 * Mirrors the public behaviour of /root/reference/src/qatseqprod.c on top of the C-ABI batching
 * layer (include/b200seqprod.h); it never includes a CUDA header.
 *
 *   reference                                        here
 *   gProcess + mutex (:180-183)                      g_process + mutex
 *   QZSTD_startQatDevice (:948-964)                  same FAIL -> STARTED -> OK state machine
 *   QZSTD_stopQatDevice (:428-449)                   back to FAIL (engines are owned by states)
 *   QZSTD_createSeqProdState/free (:992-1011)        state = engine (stream + buffers), lazily created
 *   qatSequenceProducer (:1106-1336)                 same argument checks in the same order (:1123-1137),
 *                                                    same device-down policy (:1140-1152, retry every
 *                                                    1000th block), same result check rc >= cap-1 (:1318)
 *   QZSTD_grabInstance (:905-928)                    one engine per state: no contention, no grab
 *   QZSTD_decLz4s (:1013-1091)                       b200sp_expand: 8-byte wire format -> ZSTD_Sequence
 */
#include "qatseqprod.h"
#include "b200seqprod.h"

#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define KB                            (1024)
#define COMP_LVL_MINIMUM              (1)
#define COMP_LVL_MAXIMUM              (12)
#define NUM_BLOCK_OF_RETRY_INTERVAL   (1000)     /* /root/reference/src/qatseqprod.c:88 */

typedef struct {
    int status;                 /* QZSTD_FAIL / QZSTD_STARTED / QZSTD_OK */
    pthread_mutex_t mutex;
} QZSTD_Process_T;

static QZSTD_Process_T g_process = { QZSTD_FAIL, PTHREAD_MUTEX_INITIALIZER };

static void coalesce_start_if_wanted(void);     /* cross-thread coalescing, further down */
static void coalesce_stop(void);
static void coalesce_enable_from_env(void);

typedef struct {
    b200sp_engine *engine;          /* lazily created on the first offloaded block */
    unsigned int failOffloadCnt;    /* blocks refused while the device is down (:1141) */
    /* look-ahead */
    const unsigned char *hintSrc;
    size_t hintSize, hintBlock;
    int batchLevel;                 /* level the cached batch was parsed at, 0 = no batch */
    b200sp_result batch;
    /* counters */
    unsigned long long calls, errors, batched;
} QZSTD_State_T;

/* ---- logging: same levels as the reference's QZSTD_LOG (:187-205), runtime switch ---------- */
static int log_level(void)
{
    static int level = -1;
    if (level < 0) {
        const char *e = getenv("QZSTD_DEBUGLEVEL");
        level = e ? atoi(e) : 0;
    }
    return level;
}
#define QZSTD_LOG(l, ...) do { if (log_level() >= (l)) fprintf(stderr, __VA_ARGS__); } while (0)

static int force_error(void)
{
    /* fault injection: makes every producer call fail so the application's software fallback
     * (ZSTD_c_enableSeqProducerFallback) can be exercised on a healthy device */
    const char *e = getenv("QZSTD_FORCE_ERROR");
    return e && *e && *e != '0';
}

const char *QZSTD_version(void)
{
    return QZSTD_VERSION;
}

int QZSTD_startQatDevice(void)
{
    int status;
    pthread_mutex_lock(&g_process.mutex);
    if (QZSTD_FAIL == g_process.status) {
        /* driver up?  (icp_sal_userStart in the reference, :498-527) */
        int total = b200sp_driver_device_count();
        g_process.status = total > 0 ? QZSTD_STARTED : QZSTD_FAIL;
    }
    if (QZSTD_STARTED == g_process.status) {
        /* a device with the required capability?  (instance discovery + capability filter, :529-663) */
        g_process.status = b200sp_device_count() > 0 ? QZSTD_OK : QZSTD_STARTED;
        /* context + module load now, not inside the first block (the reference starts its instances here too) */
        if (QZSTD_OK == g_process.status && b200sp_warmup(0) != B200SP_OK) {
            QZSTD_LOG(1, "Device warm-up failed: %s\n", b200sp_error_string());
            g_process.status = QZSTD_STARTED;
        }
    }
    status = g_process.status;
    QZSTD_LOG(2, "InitStatus: %d\n", status);
    pthread_mutex_unlock(&g_process.mutex);
    if (status == QZSTD_OK) {
        coalesce_enable_from_env();
        coalesce_start_if_wanted();
    }
    return status;
}

void QZSTD_stopQatDevice(void)
{
    coalesce_stop();                /* pending single-block calls are refused (ERROR -> software fallback) */
    pthread_mutex_lock(&g_process.mutex);
    g_process.status = QZSTD_FAIL;
    pthread_mutex_unlock(&g_process.mutex);
}

void *QZSTD_createSeqProdState(void)
{
    QZSTD_State_T *s = (QZSTD_State_T *)calloc(1, sizeof(QZSTD_State_T));
    return (void *)s;
}

void QZSTD_freeSeqProdState(void *sequenceProducerState)
{
    QZSTD_State_T *s = (QZSTD_State_T *)sequenceProducerState;
    if (s) {
        if (s->engine) {
            b200sp_engine_destroy(s->engine);
            s->engine = NULL;
        }
        free(s);
    }
}

void QZSTD_hintSource(void *sequenceProducerState, const void *src, size_t srcSize, size_t blockSize)
{
    QZSTD_State_T *s = (QZSTD_State_T *)sequenceProducerState;
    if (!s) return;
    s->hintSrc = (const unsigned char *)src;
    s->hintSize = src ? srcSize : 0;
    s->hintBlock = blockSize ? blockSize : ZSTD_BLOCKSIZE_MAX;
    s->batchLevel = 0;
}

void QZSTD_getStats(const void *sequenceProducerState, unsigned long long *calls,
                    unsigned long long *errors, unsigned long long *batched)
{
    const QZSTD_State_T *s = (const QZSTD_State_T *)sequenceProducerState;
    if (calls) *calls = s ? s->calls : 0;
    if (errors) *errors = s ? s->errors : 0;
    if (batched) *batched = s ? s->batched : 0;
}

static size_t producer_error(QZSTD_State_T *s)
{
    if (s) s->errors++;
    return ZSTD_SEQUENCE_PRODUCER_ERROR;
}

/* ---- cross-thread coalescing: a dispatcher parses the pending single-block calls of all threads in one batch ----
 *
 * Two batch buffers (two engines with their own pinned staging and result arrays).  Requesters take a slot in the
 * OPEN batch, copy their block into its staging slot themselves, and wait; the dispatcher closes the open batch
 * (requests arriving from then on go to the other one), waits for the copies in flight, parses the batch on the GPU,
 * and hands every requester a pointer to its packed entries, which the requester expands itself.  A buffer is
 * reused once all its requesters have expanded.  The copies in, the expansions out and the GPU work of
 * consecutive batches overlap; the dispatcher only launches and waits. */
#define COALESCE_MAX_BATCH 296        /* two waves of one-block CTAs */

typedef struct QZSTD_Request {
    uint32_t size;
    int done;
    int ok;
    size_t count;                   /* entries of this block */
    const uint64_t *packed;         /* its entries in the batch's result array (valid until the requester consumed them) */
} QZSTD_Request;

typedef struct {
    b200sp_engine *engine;
    unsigned char *slots;           /* pinned staging, slot k at k * 128 KiB */
    QZSTD_Request *reqs[COALESCE_MAX_BATCH];
    uint32_t sizes[COALESCE_MAX_BATCH];
    uint32_t taken;                 /* slots handed out */
    uint32_t copied;                /* requesters that have finished copying in */
    uint32_t toConsume;             /* requesters that still have to expand the last result */
    int level;
} QZSTD_Batch;

static struct {
    pthread_mutex_t mu;
    pthread_cond_t wake;            /* dispatcher: work arrived, copies finished, results consumed, stop requested */
    pthread_cond_t finished;        /* requesters: a batch finished or the open batch changed (broadcast) */
    QZSTD_Batch batch[2];
    int open;                       /* index of the batch that takes new requests */
    int enabled;                    /* requested by QZSTD_setCoalescing / QZSTD_COALESCE */
    int running;                    /* dispatcher thread alive */
    int stop;
    pthread_t thread;
    unsigned long long batches, blocks;
} g_co = { PTHREAD_MUTEX_INITIALIZER, PTHREAD_COND_INITIALIZER, PTHREAD_COND_INITIALIZER, {{0}}, 0, 0, 0, 0, 0, 0, 0 };

static void *coalesce_main(void *arg)
{
    int healthy = 1, i;
    (void)arg;
    for (i = 0; i < 2; i++) {
        void *slots = NULL;
        QZSTD_Batch *b = &g_co.batch[i];
        if (b200sp_engine_create(0, &b->engine) != B200SP_OK ||
            b200sp_stage_reserve(b->engine, COALESCE_MAX_BATCH, &slots) != B200SP_OK) healthy = 0;
        b->slots = (unsigned char *)slots;
    }
    pthread_mutex_lock(&g_co.mu);
    if (!healthy) {
        QZSTD_LOG(1, "Coalescing dispatcher could not start: %s\n", b200sp_error_string());
        g_co.stop = 1;
    }
    g_co.running = healthy ? 2 : 1;             /* 2: accepting requests */
    pthread_cond_broadcast(&g_co.finished);
    for (;;) {
        QZSTD_Batch *b;
        b200sp_result res;
        uint32_t n, k;
        int ok;
        while (g_co.batch[g_co.open].taken == 0 && !g_co.stop) pthread_cond_wait(&g_co.wake, &g_co.mu);
        if (g_co.batch[g_co.open].taken == 0 && g_co.stop) break;
        /* the other buffer must be free before new requests may go there */
        while (g_co.batch[g_co.open ^ 1].toConsume != 0) pthread_cond_wait(&g_co.wake, &g_co.mu);
        b = &g_co.batch[g_co.open];
        g_co.open ^= 1;                          /* close: later requests join the other batch */
        g_co.batch[g_co.open].taken = 0;
        g_co.batch[g_co.open].copied = 0;
        pthread_cond_broadcast(&g_co.finished);  /* requesters waiting for room */
        n = b->taken;
        while (b->copied < n) pthread_cond_wait(&g_co.wake, &g_co.mu);
        pthread_mutex_unlock(&g_co.mu);

        ok = healthy && b200sp_parse_staged(b->engine, b->sizes, n, b->level, &res) == B200SP_OK && res.nBlocks == n;
        if (!ok) QZSTD_LOG(1, "Coalesced parse failed: %s\n", b200sp_error_string());

        pthread_mutex_lock(&g_co.mu);
        for (k = 0; k < n; k++) {
            QZSTD_Request *r = b->reqs[k];
            r->ok = ok;
            if (ok) { r->count = res.counts[k]; r->packed = res.packed + res.offsets[k]; }
            r->done = 1;
        }
        b->toConsume = n;
        g_co.batches++; g_co.blocks += n;
        pthread_cond_broadcast(&g_co.finished);
    }
    g_co.running = 0;
    pthread_cond_broadcast(&g_co.finished);
    pthread_mutex_unlock(&g_co.mu);
    for (i = 0; i < 2; i++) if (g_co.batch[i].engine) { b200sp_engine_destroy(g_co.batch[i].engine); g_co.batch[i].engine = NULL; }
    return NULL;
}

/* Starts the dispatcher if coalescing is wanted and the device is up.  Caller holds no lock. */
static void coalesce_start_if_wanted(void)
{
    pthread_mutex_lock(&g_co.mu);
    if (g_co.enabled && !g_co.running && g_process.status == QZSTD_OK) {
        g_co.stop = 0;
        g_co.open = 0;
        memset(g_co.batch, 0, sizeof g_co.batch);
        if (pthread_create(&g_co.thread, NULL, coalesce_main, NULL) == 0) {
            g_co.running = 1;
            while (g_co.running == 1 && !g_co.stop) pthread_cond_wait(&g_co.finished, &g_co.mu);   /* engines ready */
        }
    }
    pthread_mutex_unlock(&g_co.mu);
}

static void coalesce_stop(void)
{
    pthread_t th;
    int join = 0;
    pthread_mutex_lock(&g_co.mu);
    if (g_co.running) { g_co.stop = 1; th = g_co.thread; join = 1; pthread_cond_broadcast(&g_co.wake); }
    pthread_mutex_unlock(&g_co.mu);
    if (join) pthread_join(th, NULL);
}

/* One block through the dispatcher.  Returns 0 and *rc when it was handled there, -1 when coalescing is off. */
static int coalesce_submit(const void *src, size_t srcSize, int level, ZSTD_Sequence *out, size_t cap, size_t *rc)
{
    QZSTD_Request r;
    QZSTD_Batch *b;
    uint32_t k;
    pthread_mutex_lock(&g_co.mu);
    for (;;) {
        if (g_co.running != 2 || g_co.stop) { pthread_mutex_unlock(&g_co.mu); return -1; }
        b = &g_co.batch[g_co.open];
        if (b->taken < COALESCE_MAX_BATCH && (b->taken == 0 || b->level == level)) break;
        pthread_cond_wait(&g_co.finished, &g_co.mu);        /* full, or another level: wait for the next batch */
    }
    k = b->taken++;
    if (k == 0) b->level = level;
    r.size = (uint32_t)srcSize; r.done = 0; r.ok = 0; r.count = 0; r.packed = NULL;
    b->reqs[k] = &r;
    b->sizes[k] = r.size;
    pthread_mutex_unlock(&g_co.mu);

    memcpy(b->slots + (size_t)k * B200SP_BLOCK_MAX, src, srcSize);          /* in parallel with the other requesters */

    pthread_mutex_lock(&g_co.mu);
    b->copied++;
    pthread_cond_signal(&g_co.wake);
    while (!r.done) pthread_cond_wait(&g_co.finished, &g_co.mu);
    pthread_mutex_unlock(&g_co.mu);

    *rc = ZSTD_SEQUENCE_PRODUCER_ERROR;
    if (r.ok && r.count < cap - 1) {            /* same guard as the reference (:1318-1322) */
        b200sp_expand(r.packed, r.count, (b200sp_sequence *)out);
        *rc = r.count;
    }
    pthread_mutex_lock(&g_co.mu);
    if (--b->toConsume == 0) pthread_cond_signal(&g_co.wake);               /* the buffer may be reused */
    pthread_mutex_unlock(&g_co.mu);
    return 0;
}

static void coalesce_enable_from_env(void)
{
    const char *e = getenv("QZSTD_COALESCE");
    if (e && *e && *e != '0') {
        pthread_mutex_lock(&g_co.mu);
        g_co.enabled = 1;
        pthread_mutex_unlock(&g_co.mu);
    }
}

int QZSTD_setCoalescing(int enable)
{
    int before;
    pthread_mutex_lock(&g_co.mu);
    before = g_co.enabled;
    g_co.enabled = enable ? 1 : 0;
    pthread_mutex_unlock(&g_co.mu);
    if (enable) coalesce_start_if_wanted();
    else coalesce_stop();
    return before;
}

/* device status: fail fast, retry the start every 1000th refused block (:1140-1152); then make sure the
 * state owns an engine.  Returns 0 when the block can be offloaded. */
static int device_ready(QZSTD_State_T *s)
{
    if (g_process.status != QZSTD_OK) {
        s->failOffloadCnt++;
        if (s->failOffloadCnt >= NUM_BLOCK_OF_RETRY_INTERVAL) {
            s->failOffloadCnt = 0;
            if (QZSTD_startQatDevice() != QZSTD_OK) {
                QZSTD_LOG(1, "Tried to restart the device, but failed\n");
                return -1;
            }
        } else {
            QZSTD_LOG(1, "The device was not successfully started\n");
            return -1;
        }
    }
    if (!s->engine) {
        if (b200sp_engine_create(0, &s->engine) != B200SP_OK) {
            QZSTD_LOG(1, "Failed to create engine: %s\n", b200sp_error_string());
            return -1;
        }
    }
    return 0;
}

size_t qatSequenceProducer(
    void *sequenceProducerState, ZSTD_Sequence *outSeqs, size_t outSeqsCapacity,
    const void *src, size_t srcSize,
    const void *dict, size_t dictSize,
    int compressionLevel,
    size_t windowSize)
{
    QZSTD_State_T *s = (QZSTD_State_T *)sequenceProducerState;
    size_t rc;

    if (s) s->calls++;

    if (windowSize < (srcSize < 32 * KB ? srcSize : 32 * KB) || dictSize > 0 || dict) {
        QZSTD_LOG(2, "windowSize/dictionary not supported, windowSize: %lu, srcSize: %lu, dictSize: %lu\n",
                  (unsigned long)windowSize, (unsigned long)srcSize, (unsigned long)dictSize);
        return producer_error(s);
    }

    if (compressionLevel < COMP_LVL_MINIMUM || compressionLevel > COMP_LVL_MAXIMUM) {
        QZSTD_LOG(1, "Only L1-L12 can be offloaded, current compression level: %d\n", compressionLevel);
        return producer_error(s);
    }

    if (!s || !outSeqs || !src || srcSize == 0 || srcSize > ZSTD_BLOCKSIZE_MAX || force_error()) {
        return producer_error(s);
    }

    if (device_ready(s) != 0) return producer_error(s);

    /* look-ahead: serve the block from (or first build) the batch over the hinted buffer */
    {
        const unsigned char *p = (const unsigned char *)src;
        if (s->hintSrc && p >= s->hintSrc && p + srcSize <= s->hintSrc + s->hintSize &&
            (size_t)(p - s->hintSrc) % s->hintBlock == 0) {
            const size_t idx = (size_t)(p - s->hintSrc) / s->hintBlock;
            const size_t left = s->hintSize - idx * s->hintBlock;
            const size_t expect = left < s->hintBlock ? left : s->hintBlock;
            if (expect == srcSize) {
                if (s->batchLevel != compressionLevel) {
                    if (b200sp_parse_host(s->engine, s->hintSrc, s->hintSize, (uint32_t)s->hintBlock,
                                          compressionLevel, &s->batch) != B200SP_OK) {
                        QZSTD_LOG(1, "Batch parse failed: %s\n", b200sp_error_string());
                        s->batchLevel = 0;
                        return producer_error(s);
                    }
                    s->batchLevel = compressionLevel;
                }
                if (idx < s->batch.nBlocks) {
                    rc = s->batch.counts[idx];
                    if (rc >= outSeqsCapacity - 1) return producer_error(s);
                    b200sp_expand(s->batch.packed + s->batch.offsets[idx], rc, (b200sp_sequence *)outSeqs);
                    s->batched++;
                    return rc;
                }
            }
        }
    }

    /* many threads, one block each: let the dispatcher batch them (optional) */
    if (coalesce_submit(src, srcSize, compressionLevel, outSeqs, outSeqsCapacity, &rc) == 0) {
        if (rc == ZSTD_SEQUENCE_PRODUCER_ERROR) return producer_error(s);
        s->batched++;
        return rc;
    }

    /* batch of one block */
    {
        b200sp_result one;
        if (b200sp_parse_host(s->engine, src, srcSize, (uint32_t)srcSize, compressionLevel, &one) != B200SP_OK ||
            one.nBlocks != 1) {
            QZSTD_LOG(1, "Parse failed: %s\n", b200sp_error_string());
            return producer_error(s);
        }
        s->batchLevel = 0;          /* the engine's result buffers were reused */
        rc = one.counts[0];
        if (rc >= outSeqsCapacity - 1) {        /* same guard as the reference (:1318-1322) */
            QZSTD_LOG(1, "Sequence count exceeds capacity\n");
            return producer_error(s);
        }
        b200sp_expand(one.packed, rc, (b200sp_sequence *)outSeqs);
    }
    QZSTD_LOG(2, "Produced %lu sequences\n", (unsigned long)rc);
    return rc;
}

size_t QZSTD_generateSequences(void *sequenceProducerState, ZSTD_Sequence *outSeqs, size_t outSeqsCapacity,
                               const void *src, size_t srcSize, size_t blockSize, int compressionLevel)
{
    QZSTD_State_T *s = (QZSTD_State_T *)sequenceProducerState;
    b200sp_result all;
    size_t total, b;

    if (s) s->calls++;
    if (blockSize == 0) blockSize = ZSTD_BLOCKSIZE_MAX;
    if (compressionLevel < COMP_LVL_MINIMUM || compressionLevel > COMP_LVL_MAXIMUM) {
        QZSTD_LOG(1, "Only L1-L12 can be offloaded, current compression level: %d\n", compressionLevel);
        return producer_error(s);
    }
    if (!s || !outSeqs || !src || srcSize == 0 || blockSize > ZSTD_BLOCKSIZE_MAX || force_error()) {
        return producer_error(s);
    }
    if (device_ready(s) != 0) return producer_error(s);

    if (b200sp_parse_host(s->engine, src, srcSize, (uint32_t)blockSize, compressionLevel, &all) != B200SP_OK) {
        QZSTD_LOG(1, "Batch parse failed: %s\n", b200sp_error_string());
        return producer_error(s);
    }
    s->batchLevel = 0;              /* the engine's result buffers were reused */
    total = (size_t)all.offsets[all.nBlocks];
    if (total > outSeqsCapacity) {
        QZSTD_LOG(1, "Sequence count exceeds capacity\n");
        return producer_error(s);
    }
    /* the blocks' entries are already consecutive in the wire array, each block ending with its
     * {0, trailing literals, 0} entry: one expansion pass is the whole hand-off */
    b200sp_expand(all.packed, total, (b200sp_sequence *)outSeqs);
    for (b = 0; b < all.nBlocks; b++) s->batched++;
    return total;
}
