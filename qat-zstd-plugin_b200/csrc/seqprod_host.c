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
#ifndef _GNU_SOURCE
#define _GNU_SOURCE             /* process_vm_readv */
#endif
#include "qatseqprod.h"
#include "b200seqprod.h"

#include <errno.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/uio.h>
#include <time.h>
#include <unistd.h>

#define KB                            (1024)
#define COMP_LVL_MINIMUM              (1)
#define COMP_LVL_MAXIMUM              (12)
#define NUM_BLOCK_OF_RETRY_INTERVAL   (1000)     /* /root/reference/src/qatseqprod.c:88 */

#define QZSTD_MAX_DEVICES 16
#define QZSTD_RA_MAX      64         /* blocks read ahead at most (8 MiB) */
#define QZSTD_EAGER_STATES 64        /* states whose read-ahead buffers (~240 MB of device memory each) are stood up at creation; later ones on first use */

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

typedef struct {                    /* a window of the caller's memory read ahead and parsed as one batch */
    const unsigned char *base;      /* application address of block 0, NULL = none */
    unsigned char *slots;           /* the engine's pinned copy, one block per 128 KiB slot */
    size_t block;                   /* nominal block size */
    uint32_t blocks, want;          /* blocks parsed / asked for */
    uint32_t sizes[QZSTD_RA_MAX];
    int level;
    b200sp_result batch;
} QZSTD_Window;

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
    /* transparent read-ahead (no hint): see read_ahead() */
    const unsigned char *lastSrc;   /* the previous callback's block ... */
    size_t lastSize;                /* ... and its size */
    unsigned int streak;            /* consecutive callbacks that started where the previous one ended, same size */
    QZSTD_Window win[2];            /* win[cur] is being served, win[cur ^ 1] is what the helper thread fills next */
    int cur;
    b200sp_engine *engine2;         /* engine of win[1] (win[0] uses `engine`), created with the helper thread */
    /* helper thread: fills the next window while libzstd entropy-codes the blocks of the current one */
    pthread_t raThread;
    int raThreadUp;                 /* 1: thread running, -1: could not be started (no prefetch, windows still work) */
    pthread_mutex_t raMu;
    pthread_cond_t raCv;
    int raBusy;                     /* a request is posted or being worked on */
    volatile int raReady;           /* the helper's engine exists: requests may be posted */
    int raFailed;                   /* ... or could not be created: no prefetching for this state */
    int raQuit;
    const unsigned char *raReqSrc;  /* the request: fill win[raReqWin] from raReqSrc */
    size_t raReqSize;
    int raReqLevel, raReqWin;
    uint32_t raReqWant;
    unsigned long long prefetched;  /* windows the helper had ready when they were asked for */
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

static double ra_now(void)
{
    struct timespec t;
    clock_gettime(CLOCK_MONOTONIC, &t);
    return (double)t.tv_sec + 1e-9 * (double)t.tv_nsec;
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

static int device_ready(QZSTD_State_T *s);
static int ra_limit(void);
static int ra_helper_ready(QZSTD_State_T *s);

void *QZSTD_createSeqProdState(void)
{
    QZSTD_State_T *s = (QZSTD_State_T *)calloc(1, sizeof(QZSTD_State_T));
    if (s) {
        s->device = -1;
        /* With the device already started the state takes its engine now (a stream, events, scratch: tens of
         * milliseconds that would otherwise sit inside the first block's latency); otherwise on the first block, like the
         * reference's session setup (/root/reference/src/qatseqprod.c:1193-1201).  Failure here is not an error. */
        if (g_process.status == QZSTD_OK && device_ready(s) == 0 && ra_limit() >= 2 && g_process.engines <= QZSTD_EAGER_STATES) {
            void *slots;            /* ... and the read-ahead buffers with it: allocations stall every stream of the device */
            (void)b200sp_stage_reserve(s->engine, QZSTD_RA_MAX, &slots);
            /* the helper thread and its engine as well; wait until it stands (or has given up) */
            if (ra_helper_ready(s)) {
                pthread_mutex_lock(&s->raMu);
                while (!s->raReady && !s->raFailed) pthread_cond_wait(&s->raCv, &s->raMu);
                pthread_mutex_unlock(&s->raMu);
            }
        }
    }
    return (void *)s;
}

static void ra_shutdown(QZSTD_State_T *s);          /* read-ahead helper thread + its engine, further down */

void QZSTD_freeSeqProdState(void *sequenceProducerState)
{
    QZSTD_State_T *s = (QZSTD_State_T *)sequenceProducerState;
    if (s) {
        ra_shutdown(s);
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
        const double tCreate = log_level() >= 3 ? ra_now() : 0.0;
        const int created = b200sp_engine_create(g_process.devices[slot], &s->engine);
        QZSTD_LOG(3, "engine created in %.1f ms\n", 1e3 * (ra_now() - tCreate));
        if (created != B200SP_OK) {
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

/* ---- transparent read-ahead -------------------------------------------------------------------------------
 * libzstd hands the producer one block per call, and one block is one CTA for most of a millisecond: slower than
 * software.  Applications mostly walk a buffer front to back - ZSTD_compress2 over a large buffer, or a loop
 * over chunks like the reference's benchmark (/root/reference/test/benchmark.c:300-321) - so when a call starts
 * where the previous one ended, the plugin reads AHEAD of the application: it copies the bytes that follow the
 * block (as far as they are readable) into the engine's pinned staging, parses them as one GPU batch, and serves the
 * following calls from that batch.  Two things keep this invisible:
 *   - the copy is made with process_vm_readv on the process itself, which stops at the first unreadable page
 *     instead of faulting, so reading past the end of the caller's buffer is harmless;
 *   - a block is only served from the batch when the caller's bytes still equal the copy (memcmp): blocks are
 *     parsed independently, so equal bytes mean valid sequences, and anything else (a buffer refilled in the
 *     meantime, a shorter last block) is a miss that takes the single-block path.
 * QZSTD_LOOKAHEAD=0 turns it off; QZSTD_LOOKAHEAD=n bounds the window to n blocks (default and maximum 64). */
static int ra_limit(void)
{
    static int limit = -1;
    if (limit < 0) {
        const char *e = getenv("QZSTD_LOOKAHEAD");
        int v = (e && *e) ? atoi(e) : QZSTD_RA_MAX;
        limit = v < 0 ? 0 : v > QZSTD_RA_MAX ? QZSTD_RA_MAX : v;
    }
    return limit;
}


static int g_raBroken = 0;          /* process_vm_readv refused (seccomp / ptrace policy): never tried again */

/* Serves block `src` from window w, if it is there and unchanged.  Returns 1 and *rc on a hit, *last = it was the
 * window's last block, *k = its index. */
static int ra_serve(QZSTD_Window *w, const unsigned char *src, size_t srcSize, int level,
                    ZSTD_Sequence *out, size_t cap, size_t *rc, size_t *kOut)
{
    size_t k;
    if (!w->base || level != w->level || src < w->base) return 0;
    if ((size_t)(src - w->base) % w->block != 0) return 0;
    k = (size_t)(src - w->base) / w->block;
    if (k >= w->blocks || w->sizes[k] != srcSize) return 0;
    if (memcmp(src, w->slots + k * (size_t)B200SP_BLOCK_MAX, srcSize) != 0) return 0;
    *rc = w->batch.counts[k];
    *kOut = k;
    if (*rc >= cap - 1) { *rc = ZSTD_SEQUENCE_PRODUCER_ERROR; return 1; }   /* same guard as the reference (:1318-1322) */
    b200sp_expand(w->batch.packed + w->batch.offsets[k], *rc, (b200sp_sequence *)out);
    return 1;
}

/* Reads `want` blocks of `srcSize` bytes ahead from `src` into the engine's pinned staging (as far as they are
 * readable) and parses them.  Returns 1 when a window of at least two blocks is in place.  Runs on the caller's
 * thread (first window of a run) or on the state's helper thread (every further one). */
static int ra_fill(QZSTD_Window *w, b200sp_engine *engine, const unsigned char *src, size_t srcSize, int level, uint32_t want)
{
    struct iovec local[QZSTD_RA_MAX], remote[QZSTD_RA_MAX];
    void *slots = NULL;
    ssize_t got;
    uint32_t n, i;
    double t0 = 0.0, t1 = 0.0;
    w->base = NULL;
    if (want < 2 || g_raBroken || !engine) return 0;
    if (b200sp_stage_reserve(engine, QZSTD_RA_MAX, &slots) != B200SP_OK) return 0;
    for (i = 0; i < want; i++) {
        local[i].iov_base = (unsigned char *)slots + (size_t)i * B200SP_BLOCK_MAX;
        local[i].iov_len = srcSize;
        /* one remote element per block: a transfer that stops at an unreadable page is cut between elements at the
         * latest (process_vm_readv(2): partial transfers never split an element on some kernels) */
        remote[i].iov_base = (void *)(src + (size_t)i * srcSize);
        remote[i].iov_len = srcSize;
    }
    if (log_level() >= 3) t0 = ra_now();
    got = process_vm_readv(getpid(), local, want, remote, want, 0);
    if (log_level() >= 3) t1 = ra_now();
    if (got < (ssize_t)srcSize) {
        if (got < 0 && src != NULL && errno != EFAULT) { g_raBroken = 1; QZSTD_LOG(1, "read-ahead disabled: process_vm_readv refused\n"); }
        return 0;
    }
    n = (uint32_t)((size_t)got / srcSize);
    for (i = 0; i < n; i++) w->sizes[i] = (uint32_t)srcSize;
    if ((size_t)got % srcSize != 0 && n < want) w->sizes[n++] = (uint32_t)((size_t)got % srcSize);   /* what is readable of the last one */
    if (n < 2) return 0;
    if (b200sp_parse_staged(engine, w->sizes, n, level, &w->batch) != B200SP_OK || w->batch.nBlocks != n) {
        QZSTD_LOG(1, "Read-ahead parse failed: %s\n", b200sp_error_string());
        return 0;
    }
    QZSTD_LOG(3, "read-ahead: %u blocks, copy %.3f ms, parse %.3f ms\n", n, 1e3 * (t1 - t0), 1e3 * (ra_now() - t1));
    w->slots = (unsigned char *)slots;
    w->block = srcSize;
    w->blocks = n;
    w->want = want;
    w->level = level;
    w->base = src;
    return 1;
}

static uint32_t ra_window_blocks(unsigned int streak)
{
    /* the window grows with the length of the sequential run, like a file system's read-ahead */
    uint32_t want = streak < 2 ? 8u : streak < 12 ? 24u : (uint32_t)QZSTD_RA_MAX;
    const int limit = ra_limit();
    return want > (uint32_t)limit ? (uint32_t)limit : want;
}

static void *ra_helper(void *arg)
{
    QZSTD_State_T *s = (QZSTD_State_T *)arg;
    /* the second engine is created here, not on the caller's thread: a stream, scratch and pinned buffers take
     * tens of milliseconds; a request posted meanwhile waits, and fails cleanly (no window) without an engine */
    {
        const double t0 = ra_now();
        void *slots;
        if (b200sp_engine_create(b200sp_engine_device(s->engine), &s->engine2) != B200SP_OK) s->engine2 = NULL;
        else (void)b200sp_stage_reserve(s->engine2, QZSTD_RA_MAX, &slots);
        QZSTD_LOG(3, "helper engine created in %.1f ms\n", 1e3 * (ra_now() - t0));
    }
    pthread_mutex_lock(&s->raMu);
    s->raReady = s->engine2 != NULL;
    s->raFailed = s->engine2 == NULL;
    pthread_cond_broadcast(&s->raCv);
    for (;;) {
        while (!s->raQuit && !(s->raBusy && s->raReqSrc)) pthread_cond_wait(&s->raCv, &s->raMu);
        if (s->raQuit) break;
        {
            const unsigned char *src = s->raReqSrc;
            const size_t size = s->raReqSize;
            const int level = s->raReqLevel, wi = s->raReqWin;
            const uint32_t want = s->raReqWant;
            s->raReqSrc = NULL;
            pthread_mutex_unlock(&s->raMu);
            ra_fill(&s->win[wi], wi ? s->engine2 : s->engine, src, size, level, want);
            pthread_mutex_lock(&s->raMu);
        }
        s->raBusy = 0;
        pthread_cond_broadcast(&s->raCv);
    }
    pthread_mutex_unlock(&s->raMu);
    return NULL;
}

/* Waits until the helper thread is idle: before the caller's thread touches either engine. */
static void ra_quiesce(QZSTD_State_T *s)
{
    if (s->raThreadUp != 1) return;
    pthread_mutex_lock(&s->raMu);
    while (s->raBusy) pthread_cond_wait(&s->raCv, &s->raMu);
    pthread_mutex_unlock(&s->raMu);
}

/* Starts the helper thread and its engine on first use.  Returns 1 when prefetching is possible. */
static int ra_helper_ready(QZSTD_State_T *s)
{
    if (s->raThreadUp) return s->raThreadUp == 1;
    s->raThreadUp = -1;
    /* the helper's engine is an internal resource of its state: it does not count against the engine pool */
    pthread_mutex_init(&s->raMu, NULL);
    pthread_cond_init(&s->raCv, NULL);
    if (pthread_create(&s->raThread, NULL, ra_helper, s) != 0) return 0;
    s->raThreadUp = 1;
    return 1;
}

static void ra_post(QZSTD_State_T *s, int wi, const unsigned char *src, size_t size, int level, uint32_t want)
{
    pthread_mutex_lock(&s->raMu);
    s->win[wi].base = NULL;
    s->raReqSrc = src; s->raReqSize = size; s->raReqLevel = level; s->raReqWin = wi; s->raReqWant = want;
    s->raBusy = 1;
    pthread_cond_broadcast(&s->raCv);
    pthread_mutex_unlock(&s->raMu);
}

static void ra_shutdown(QZSTD_State_T *s)
{
    if (s->raThreadUp == 1) {
        pthread_mutex_lock(&s->raMu);
        s->raQuit = 1;
        pthread_cond_broadcast(&s->raCv);
        pthread_mutex_unlock(&s->raMu);
        pthread_join(s->raThread, NULL);
        pthread_mutex_destroy(&s->raMu);
        pthread_cond_destroy(&s->raCv);
        s->raThreadUp = 0;
    }
    if (s->engine2) {
        b200sp_engine_destroy(s->engine2);
        s->engine2 = NULL;
    }
}

/* A window has just become current: if it was filled completely (the buffer goes on), have the helper fetch the one
 * after it while this one is served - libzstd's entropy stage takes ~150 us per block, a window of 128 blocks is
 * fetched and parsed in 5 ms. */
static void ra_prefetch_next(QZSTD_State_T *s, int level)
{
    const QZSTD_Window *w = &s->win[s->cur];
    if (w->base && w->blocks == w->want && !s->raBusy && ra_helper_ready(s) && s->raReady)
        ra_post(s, s->cur ^ 1, w->base + (size_t)w->blocks * w->block, w->block, level, ra_window_blocks(s->streak + w->blocks));
}

/* The read-ahead step of the producer.  Returns 1 when the block was served (*rc), 0 when the caller goes on to the
 * single-block path.  On return 0 the helper is idle and no window is kept. */
static int read_ahead(QZSTD_State_T *s, const unsigned char *p, size_t srcSize, int level,
                      ZSTD_Sequence *out, size_t cap, size_t *rc)
{
    const int sequential = s->lastSrc && p == s->lastSrc + s->lastSize && srcSize == s->lastSize;
    QZSTD_Window *w = &s->win[s->cur];
    size_t k = 0;
    s->streak = sequential ? s->streak + 1 : 0;
    s->lastSrc = p;
    s->lastSize = srcSize;
    if (ra_limit() < 2) return 0;
    /* 1. the window being served */
    if (ra_serve(w, p, srcSize, level, out, cap, rc, &k)) {
        if (k + 1 == w->blocks) w->base = NULL;
        return 1;
    }
    /* 2. the window the helper has been filling */
    ra_quiesce(s);
    w->base = NULL;
    if (ra_serve(&s->win[s->cur ^ 1], p, srcSize, level, out, cap, rc, &k)) {
        s->cur ^= 1;
        s->prefetched++;
        ra_prefetch_next(s, level);
        if (k + 1 == s->win[s->cur].blocks) s->win[s->cur].base = NULL;
        return 1;
    }
    s->win[s->cur ^ 1].base = NULL;
    /* 3. a sequential run without a window: read ahead now, on this thread */
    s->cur = 0;
    if (sequential && srcSize >= 4 * KB && ra_fill(&s->win[0], s->engine, p, srcSize, level, ra_window_blocks(s->streak))) {
        s->batchLevel = 0;          /* the engine's result buffers were reused */
        if (ra_serve(&s->win[0], p, srcSize, level, out, cap, rc, &k)) {
            ra_prefetch_next(s, level);
            return 1;
        }
    }
    s->win[0].base = NULL;
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
    if (s->hintSrc) ra_quiesce(s);      /* the helper thread may be using the engine */
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
                    s->win[0].base = NULL;
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

    /* transparent read-ahead: a call that continues the previous one is served from (or first builds) a window of
     * the blocks that follow in the caller's memory */
    if (!g_coEnabled && read_ahead(s, (const unsigned char *)src, srcSize, compressionLevel, outSeqs, outSeqsCapacity, &rc)) {
        if (rc == ZSTD_SEQUENCE_PRODUCER_ERROR) return producer_error(s);
        s->batched++;
        return rc;
    }

    /* many threads, one block each: let the device's dispatcher batch them (optional) */
    if (coalesce_submit(s->device, src, srcSize, compressionLevel, outSeqs, outSeqsCapacity, &rc) == 0) {
        if (rc == ZSTD_SEQUENCE_PRODUCER_ERROR) return producer_error(s);
        s->batched++;
        return rc;
    }

    /* batch of one block */
    ra_quiesce(s);                  /* (the helper thread is idle here unless coalescing was switched on under a prefetch) */
    {
        b200sp_result one;
        if (b200sp_parse_host(s->engine, src, srcSize, (uint32_t)srcSize, compressionLevel, &one) != B200SP_OK ||
            one.nBlocks != 1) {
            QZSTD_LOG(1, "Parse failed: %s\n", b200sp_error_string());
            return producer_error(s);
        }
        s->batchLevel = 0;          /* the engine's result buffers were reused */
        s->win[0].base = NULL;
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

int QZSTD_registerBuffer(void *ptr, size_t size)
{
    if (g_process.status != QZSTD_OK) return QZSTD_FAIL;
    return b200sp_host_register(ptr, size) == B200SP_OK ? QZSTD_OK : QZSTD_FAIL;
}

int QZSTD_unregisterBuffer(void *ptr)
{
    return b200sp_host_unregister(ptr) == B200SP_OK ? QZSTD_OK : QZSTD_FAIL;
}

size_t QZSTD_generateSequences(void *sequenceProducerState, ZSTD_Sequence *outSeqs, size_t outSeqsCapacity,
                               const void *src, size_t srcSize, size_t blockSize, int compressionLevel)
{
    return QZSTD_generateSequencesIndexed(sequenceProducerState, outSeqs, outSeqsCapacity, src, srcSize, blockSize,
                                          compressionLevel, NULL, 0);
}

size_t QZSTD_generateSequencesIndexed(void *sequenceProducerState, ZSTD_Sequence *outSeqs, size_t outSeqsCapacity,
                                      const void *src, size_t srcSize, size_t blockSize, int compressionLevel,
                                      size_t *blockIndex, size_t blockIndexCapacity)
{
    QZSTD_State_T *s = (QZSTD_State_T *)sequenceProducerState;
    size_t total = 0;
    b200sp_result res;

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
    ra_quiesce(s);                  /* the helper thread may be using the engine */

    /* the device gathers every block's entries into ONE dense ZSTD_Sequence array, each block ending with its
     * {0, trailing literals, 0} entry, and it lands in outSeqs directly (what QZSTD_decLz4s leaves there,
     * :1013-1091, for all blocks at once): no per-entry work on the host */
    if (blockIndex && blockIndexCapacity < (srcSize + blockSize - 1) / blockSize + 1) return producer_error(s);
    if (b200sp_sequences_host(s->engine, src, srcSize, (uint32_t)blockSize, compressionLevel,
                              (b200sp_sequence *)outSeqs, outSeqsCapacity, &total, &res) != B200SP_OK) {
        QZSTD_LOG(1, "Batch parse failed: %s\n", b200sp_error_string());
        return producer_error(s);
    }
    if (blockIndex) {
        uint32_t b;
        for (b = 0; b <= res.nBlocks; b++) blockIndex[b] = (size_t)res.offsets[b];
    }
    s->batchLevel = 0;              /* the engine's result buffers were reused */
    s->win[0].base = NULL;
    s->batched += (srcSize + blockSize - 1) / blockSize;
    return total;
}
