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
 *   QZSTD_getAndShuffleInstance (:601-630)           states take the usable devices round-robin (QZSTD_DEVICES)
 *   QZSTD_grabInstance (:905-928)                    one engine per state out of a bounded pool (QZSTD_MAX_ENGINES):
 *                                                    exhaustion answers ERROR, never blocks
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

#define QZSTD_MAX_DEVICES 16

typedef struct {
    int status;                 /* QZSTD_FAIL / QZSTD_STARTED / QZSTD_OK */
    pthread_mutex_t mutex;
    int devices[QZSTD_MAX_DEVICES];     /* usable devices that warmed up, in the order states take them */
    int nDevices;
    unsigned int nextDevice;            /* round-robin cursor (:601-630 spreads instances over devices the same way) */
    int engines, maxEngines;            /* live engines / bound (the instance pool is finite too, :905-928) */
} QZSTD_Process_T;

static QZSTD_Process_T g_process = { QZSTD_FAIL, PTHREAD_MUTEX_INITIALIZER, {0}, 0, 0, 0, 256 };

static void coalesce_start_if_wanted(void);     /* cross-thread coalescing, further down */
static void coalesce_stop(void);
static void coalesce_enable_from_env(void);

typedef struct {
    b200sp_engine *engine;          /* lazily created on the first offloaded block */
    int device;                     /* the device this state was dealt, -1 before the first block */
    unsigned int failOffloadCnt;    /* blocks refused while the device is down (:1141) */
    /* look-ahead */
    const unsigned char *hintSrc;
    size_t hintSize, hintBlock;
    int batchLevel;                 /* level the cached batch was parsed at, 0 = no batch */
    size_t served;                  /* blocks of the cached batch handed out so far */
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

/* QZSTD_DEVICES="0,2,3" restricts (and orders) the devices the plugin uses; default: every usable device. */
static int wanted_device(int dev)
{
    const char *e = getenv("QZSTD_DEVICES");
    if (!e || !*e) return 1;
    while (*e) {
        char *end;
        long v = strtol(e, &end, 10);
        if (end == e) break;
        if (v == dev) return 1;
        e = *end ? end + 1 : end;
    }
    return 0;
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
        /* devices with the required capability (instance discovery + capability filter, :529-663); context and
         * module load now, not inside the first block (the reference starts its instances here too) */
        int all[QZSTD_MAX_DEVICES], n, i;
        const char *m = getenv("QZSTD_MAX_ENGINES");
        g_process.maxEngines = (m && atoi(m) > 0) ? atoi(m) : 256;
        n = b200sp_usable_devices(all, QZSTD_MAX_DEVICES);
        if (n > QZSTD_MAX_DEVICES) n = QZSTD_MAX_DEVICES;
        g_process.nDevices = 0;
        for (i = 0; i < n; i++) {
            if (!wanted_device(all[i])) continue;
            if (b200sp_warmup(all[i]) != B200SP_OK) {
                QZSTD_LOG(1, "Device %d warm-up failed: %s\n", all[i], b200sp_error_string());
                continue;
            }
            g_process.devices[g_process.nDevices++] = all[i];
        }
        g_process.status = g_process.nDevices > 0 ? QZSTD_OK : QZSTD_STARTED;
    }
    status = g_process.status;
    QZSTD_LOG(2, "InitStatus: %d (%d device(s))\n", status, g_process.nDevices);
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
    if (s) s->device = -1;
    return (void *)s;
}

void QZSTD_freeSeqProdState(void *sequenceProducerState)
{
    QZSTD_State_T *s = (QZSTD_State_T *)sequenceProducerState;
    if (s) {
        if (s->engine) {
            b200sp_engine_destroy(s->engine);
            s->engine = NULL;
            pthread_mutex_lock(&g_process.mutex);
            g_process.engines--;
            pthread_mutex_unlock(&g_process.mutex);
        }
        free(s);
    }
}

void QZSTD_hintSource(void *sequenceProducerState, const void *src, size_t srcSize, size_t blockSize)
{
    QZSTD_State_T *s = (QZSTD_State_T *)sequenceProducerState;
    if (!s) return;
    if (blockSize > ZSTD_BLOCKSIZE_MAX) src = NULL;         /* libzstd never hands out larger blocks: such a hint cannot match */
    s->hintSrc = (const unsigned char *)src;
    s->hintSize = src ? srcSize : 0;
    s->hintBlock = blockSize ? blockSize : ZSTD_BLOCKSIZE_MAX;
    s->batchLevel = 0;
    s->served = 0;
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

/* ---- cross-thread coalescing: a dispatcher per device parses the pending single-block calls of all threads in one batch ----
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

typedef struct {
    pthread_mutex_t mu;
    pthread_cond_t wake;            /* dispatcher: work arrived, copies finished, results consumed, stop requested */
    pthread_cond_t finished;        /* requesters: a batch finished or the open batch changed (broadcast) */
    QZSTD_Batch batch[2];
    int device;
    int open;                       /* index of the batch that takes new requests */
    int running;                    /* 0 no thread, 1 starting, 2 accepting requests */
    int stop;
    int threadValid;                /* `thread` was created and has not been joined yet */
    pthread_t thread;
    unsigned long long batches, blocks;
} QZSTD_Coalescer;

static QZSTD_Coalescer g_co[QZSTD_MAX_DEVICES];     /* one per entry of g_process.devices */
static int g_coInit = 0;
static int g_coEnabled = 0;                         /* requested by QZSTD_setCoalescing / QZSTD_COALESCE */
static pthread_mutex_t g_coMu = PTHREAD_MUTEX_INITIALIZER;      /* start/stop of the dispatchers */

static void *coalesce_main(void *arg)
{
    QZSTD_Coalescer *co = (QZSTD_Coalescer *)arg;
    int healthy = 1, i;
    for (i = 0; i < 2; i++) {
        void *slots = NULL;
        QZSTD_Batch *b = &co->batch[i];
        if (b200sp_engine_create(co->device, &b->engine) != B200SP_OK ||
            b200sp_stage_reserve(b->engine, COALESCE_MAX_BATCH, &slots) != B200SP_OK) healthy = 0;
        b->slots = (unsigned char *)slots;
    }
    pthread_mutex_lock(&co->mu);
    if (!healthy) {
        QZSTD_LOG(1, "Coalescing dispatcher could not start: %s\n", b200sp_error_string());
        co->stop = 1;
    }
    co->running = 2;                            /* requests are accepted only while !stop */
    pthread_cond_broadcast(&co->finished);
    while (healthy) {
        QZSTD_Batch *b;
        b200sp_result res;
        uint32_t n, k;
        int ok;
        while (co->batch[co->open].taken == 0 && !co->stop) pthread_cond_wait(&co->wake, &co->mu);
        if (co->batch[co->open].taken == 0 && co->stop) break;
        /* the other buffer must be free before new requests may go there */
        while (co->batch[co->open ^ 1].toConsume != 0) pthread_cond_wait(&co->wake, &co->mu);
        b = &co->batch[co->open];
        co->open ^= 1;                           /* close: later requests join the other batch */
        co->batch[co->open].taken = 0;
        co->batch[co->open].copied = 0;
        pthread_cond_broadcast(&co->finished);   /* requesters waiting for room */
        n = b->taken;
        while (b->copied < n) pthread_cond_wait(&co->wake, &co->mu);
        pthread_mutex_unlock(&co->mu);

        ok = b200sp_parse_staged(b->engine, b->sizes, n, b->level, &res) == B200SP_OK && res.nBlocks == n;
        if (!ok) QZSTD_LOG(1, "Coalesced parse failed: %s\n", b200sp_error_string());

        pthread_mutex_lock(&co->mu);
        for (k = 0; k < n; k++) {
            QZSTD_Request *r = b->reqs[k];
            r->ok = ok;
            if (ok) { r->count = res.counts[k]; r->packed = res.packed + res.offsets[k]; }
            r->done = 1;
        }
        b->toConsume = n;
        co->batches++; co->blocks += n;
        pthread_cond_broadcast(&co->finished);
    }
    /* requesters of the last batches are still reading the engines' result arrays (and will decrement toConsume):
     * the engines - and this dispatcher's state - stay until every one of them is done */
    while (co->batch[0].toConsume != 0 || co->batch[1].toConsume != 0) pthread_cond_wait(&co->wake, &co->mu);
    co->running = 0;
    pthread_cond_broadcast(&co->finished);
    pthread_mutex_unlock(&co->mu);
    for (i = 0; i < 2; i++) if (co->batch[i].engine) { b200sp_engine_destroy(co->batch[i].engine); co->batch[i].engine = NULL; }
    return NULL;
}

/* Starts the dispatchers (one per device) if coalescing is wanted and the device is up.  Caller holds no lock. */
static void coalesce_start_if_wanted(void)
{
    int d;
    pthread_mutex_lock(&g_coMu);
    if (!g_coInit) {
        for (d = 0; d < QZSTD_MAX_DEVICES; d++) {
            pthread_mutex_init(&g_co[d].mu, NULL);
            pthread_cond_init(&g_co[d].wake, NULL);
            pthread_cond_init(&g_co[d].finished, NULL);
        }
        g_coInit = 1;
    }
    if (g_coEnabled && g_process.status == QZSTD_OK) {
        for (d = 0; d < g_process.nDevices; d++) {
            QZSTD_Coalescer *co = &g_co[d];
            pthread_mutex_lock(&co->mu);
            if (!co->running) {
                if (co->threadValid) {           /* a dispatcher that could not start, or has stopped: reap it first */
                    pthread_t old = co->thread;
                    co->threadValid = 0;
                    pthread_mutex_unlock(&co->mu);
                    pthread_join(old, NULL);
                    pthread_mutex_lock(&co->mu);
                }
                co->stop = 0;
                co->open = 0;
                co->device = g_process.devices[d];
                memset(co->batch, 0, sizeof co->batch);
                if (pthread_create(&co->thread, NULL, coalesce_main, co) == 0) {
                    co->threadValid = 1;
                    co->running = 1;
                    while (co->running == 1) pthread_cond_wait(&co->finished, &co->mu);   /* engines ready (or refused) */
                }
            }
            pthread_mutex_unlock(&co->mu);
        }
    }
    pthread_mutex_unlock(&g_coMu);
}

static void coalesce_stop(void)
{
    int d;
    pthread_mutex_lock(&g_coMu);
    for (d = 0; g_coInit && d < QZSTD_MAX_DEVICES; d++) {
        QZSTD_Coalescer *co = &g_co[d];
        pthread_t th;
        int join = 0;
        pthread_mutex_lock(&co->mu);
        if (co->running) { co->stop = 1; pthread_cond_broadcast(&co->wake); pthread_cond_broadcast(&co->finished); }
        if (co->threadValid) { th = co->thread; join = 1; co->threadValid = 0; }
        pthread_mutex_unlock(&co->mu);
        if (join) pthread_join(th, NULL);       /* returns once the last requester has consumed its result */
    }
    pthread_mutex_unlock(&g_coMu);
}

/* One block through the dispatcher of device slot d.  Returns 0 and *rc when it was handled there, -1 when coalescing is off. */
static int coalesce_submit(int d, const void *src, size_t srcSize, int level, ZSTD_Sequence *out, size_t cap, size_t *rc)
{
    QZSTD_Coalescer *co;
    QZSTD_Request r;
    QZSTD_Batch *b;
    uint32_t k;
    if (!g_coEnabled || !g_coInit || d < 0) return -1;
    co = &g_co[d];
    pthread_mutex_lock(&co->mu);
    for (;;) {
        if (co->running != 2 || co->stop) { pthread_mutex_unlock(&co->mu); return -1; }
        b = &co->batch[co->open];
        if (b->taken < COALESCE_MAX_BATCH && (b->taken == 0 || b->level == level)) break;
        pthread_cond_wait(&co->finished, &co->mu);          /* full, or another level: wait for the next batch */
    }
    k = b->taken++;
    if (k == 0) b->level = level;
    r.size = (uint32_t)srcSize; r.done = 0; r.ok = 0; r.count = 0; r.packed = NULL;
    b->reqs[k] = &r;
    b->sizes[k] = r.size;
    pthread_mutex_unlock(&co->mu);

    memcpy(b->slots + (size_t)k * B200SP_BLOCK_MAX, src, srcSize);          /* in parallel with the other requesters */

    pthread_mutex_lock(&co->mu);
    b->copied++;
    pthread_cond_signal(&co->wake);
    while (!r.done) pthread_cond_wait(&co->finished, &co->mu);
    pthread_mutex_unlock(&co->mu);

    *rc = ZSTD_SEQUENCE_PRODUCER_ERROR;
    if (r.ok && r.count < cap - 1) {            /* same guard as the reference (:1318-1322) */
        b200sp_expand(r.packed, r.count, (b200sp_sequence *)out);
        *rc = r.count;
    }
    pthread_mutex_lock(&co->mu);
    if (--b->toConsume == 0) pthread_cond_broadcast(&co->wake);             /* the buffer may be reused / the dispatcher may leave */
    pthread_mutex_unlock(&co->mu);
    return 0;
}

static void coalesce_enable_from_env(void)
{
    const char *e = getenv("QZSTD_COALESCE");
    if (e && *e && *e != '0') {
        pthread_mutex_lock(&g_coMu);
        g_coEnabled = 1;
        pthread_mutex_unlock(&g_coMu);
    }
}

int QZSTD_setCoalescing(int enable)
{
    int before;
    pthread_mutex_lock(&g_coMu);
    before = g_coEnabled;
    g_coEnabled = enable ? 1 : 0;
    pthread_mutex_unlock(&g_coMu);
    if (enable) coalesce_start_if_wanted();
    else coalesce_stop();
    return before;
}

/* device status: fail fast, retry the start every 1000th refused block (:1140-1152); then make sure the
 * state owns an engine on the device it was dealt (round-robin over the usable devices, the spread of
 * QZSTD_getAndShuffleInstance :601-630; the pool of engines is bounded like the instance pool, :905-928:
 * exhaustion answers ERROR).  Returns 0 when the block can be offloaded. */
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
        int slot, granted;
        pthread_mutex_lock(&g_process.mutex);
        granted = g_process.nDevices > 0 && g_process.engines < g_process.maxEngines;
        slot = granted ? (int)(g_process.nextDevice++ % (unsigned int)g_process.nDevices) : -1;
        if (granted) g_process.engines++;
        pthread_mutex_unlock(&g_process.mutex);
        if (!granted) {
            QZSTD_LOG(1, "Failed to grab an engine: all %d are taken\n", g_process.maxEngines);
            return -1;
        }
        if (b200sp_engine_create(g_process.devices[slot], &s->engine) != B200SP_OK) {
            QZSTD_LOG(1, "Failed to create engine: %s\n", b200sp_error_string());
            pthread_mutex_lock(&g_process.mutex);
            g_process.engines--;
            pthread_mutex_unlock(&g_process.mutex);
            s->engine = NULL;
            return -1;
        }
        s->device = slot;
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
                /* The cached batch belongs to ONE pass over the buffer.  Block 0 coming round again means a new
                 * compression of a buffer that may hold new bytes: parse again rather than serve stale matches,
                 * which libzstd would not catch (SURVEY App. B case 11). */
                if (idx == 0 && s->served != 0) s->batchLevel = 0;
                if (s->batchLevel != compressionLevel) {
                    if (b200sp_parse_host(s->engine, s->hintSrc, s->hintSize, (uint32_t)s->hintBlock,
                                          compressionLevel, &s->batch) != B200SP_OK) {
                        QZSTD_LOG(1, "Batch parse failed: %s\n", b200sp_error_string());
                        /* do not pay for (and fail) the whole batch again on every later block */
                        s->batchLevel = 0; s->hintSrc = NULL; s->hintSize = 0;
                        return producer_error(s);
                    }
                    s->batchLevel = compressionLevel;
                    s->served = 0;
                }
                if (idx < s->batch.nBlocks) {
                    rc = s->batch.counts[idx];
                    if (rc >= outSeqsCapacity - 1) return producer_error(s);
                    b200sp_expand(s->batch.packed + s->batch.offsets[idx], rc, (b200sp_sequence *)outSeqs);
                    s->batched++;
                    s->served++;
                    if (idx + 1 == s->batch.nBlocks) s->batchLevel = 0;     /* the pass is over: the next one parses afresh */
                    return rc;
                }
            }
        }
    }

    /* many threads, one block each: let the device's dispatcher batch them (optional) */
    if (coalesce_submit(s->device, src, srcSize, compressionLevel, outSeqs, outSeqsCapacity, &rc) == 0) {
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
    size_t total = 0;

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

    /* the device gathers every block's entries into ONE dense ZSTD_Sequence array, each block ending with its
     * {0, trailing literals, 0} entry, and it lands in outSeqs directly (what QZSTD_decLz4s leaves there,
     * :1013-1091, for all blocks at once): no per-entry work on the host */
    if (b200sp_sequences_host(s->engine, src, srcSize, (uint32_t)blockSize, compressionLevel,
                              (b200sp_sequence *)outSeqs, outSeqsCapacity, &total, NULL) != B200SP_OK) {
        QZSTD_LOG(1, "Batch parse failed: %s\n", b200sp_error_string());
        return producer_error(s);
    }
    s->batchLevel = 0;              /* the engine's result buffers were reused */
    s->batched += (srcSize + blockSize - 1) / blockSize;
    return total;
}
