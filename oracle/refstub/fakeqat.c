/*
 * refstub/fakeqat.c — the fake QAT device behind the stub headers. TEST INFRASTRUCTURE ONLY.
 *
 * A software stand-in for a QAT gen-4 data-compression endpoint, just faithful enough for the
 * reference plugin's control flow: devices x instances (FAKEQAT_DEVICES, FAKEQAT_INSTANCES, default
 * 1 x 2), stateless LZ4s compression requests that complete asynchronously and are reaped by
 * icp_sal_DcPollInstance, which fires the session callback exactly like the real driver does for
 * /root/reference/src/qatseqprod.c:665-680.  The LZ4s "engine" is a plain greedy hash matcher
 * (min match 3, 16-bit offsets) that emits the token layout QZSTD_decLz4s consumes
 * (/root/reference/src/qatseqprod.c:1013-1091).  It is NOT Intel's algorithm and
 * makes no claim about QAT ratios or speed.
 *
 * Knobs (environment): FAKEQAT_DEVICES=0 simulates "no QAT hardware" (icp_adf_get_numDevices -> 0);
 * FAKEQAT_NO_LZ4S=1 reports instances without LZ4s capability; FAKEQAT_INCOMPRESSIBLE=1 makes every
 * request report dataUncompressed; FAKEQAT_FAIL_SUBMIT=1 makes cpaDcCompressData2 fail.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "cpa.h"
#include "cpa_dc.h"
#include "icp_sal_user.h"
#include "icp_sal_poll.h"
#include "qae_mem.h"

#define MAX_INST 64
typedef struct { int device; int started; void *pendingTag; CpaDcCallbackFn pendingCb; int pending; } FakeInst;
typedef struct { CpaDcCallbackFn cb; CpaDcSessionSetupData sd; } FakeSession;

static FakeInst g_inst[MAX_INST];
static int g_numInst = -1;

static int env_int(const char *name, int dflt) { const char *e = getenv(name); return e && *e ? atoi(e) : dflt; }

static void discover(void)
{
    if (g_numInst >= 0) return;
    int dev = env_int("FAKEQAT_DEVICES", 1), per = env_int("FAKEQAT_INSTANCES", 2);
    if (dev < 0) dev = 0;
    if (per < 1) per = 1;
    g_numInst = dev * per > MAX_INST ? MAX_INST : dev * per;
    for (int i = 0; i < g_numInst; i++) { memset(&g_inst[i], 0, sizeof g_inst[i]); g_inst[i].device = i / per; }
}

/* ---- user-space service access layer ------------------------------------------------------ */
CpaStatus icp_adf_get_numDevices(Cpa32U *n) { *n = (Cpa32U)env_int("FAKEQAT_DEVICES", 1); return CPA_STATUS_SUCCESS; }
CpaBoolean icp_sal_userIsQatAvailable(void) { return env_int("FAKEQAT_DEVICES", 1) > 0 ? CPA_TRUE : CPA_FALSE; }
CpaStatus icp_sal_userStart(const char *name) { (void)name; discover(); return CPA_STATUS_SUCCESS; }
CpaStatus icp_sal_userStop(void) { g_numInst = -1; return CPA_STATUS_SUCCESS; }

/* ---- USDM ------------------------------------------------------------------------------------ */
void *qaeMemAllocNUMA(size_t size, int node, size_t align)
{
    (void)node;
    void *p = NULL;
    if (align < sizeof(void *)) align = sizeof(void *);
    if (posix_memalign(&p, align, size ? size : 1) != 0) return NULL;
    return p;
}
void qaeMemFreeNUMA(void **ptr) { if (ptr && *ptr) { free(*ptr); *ptr = NULL; } }
uint64_t qaeVirtToPhysNUMA(void *p) { return (uint64_t)(uintptr_t)p; }

/* ---- instances --------------------------------------------------------------------------------- */
CpaStatus cpaDcGetNumInstances(Cpa16U *n) { discover(); *n = (Cpa16U)g_numInst; return CPA_STATUS_SUCCESS; }
CpaStatus cpaDcGetInstances(Cpa16U n, CpaInstanceHandle *h)
{
    discover();
    for (int i = 0; i < n && i < g_numInst; i++) h[i] = &g_inst[i];
    return CPA_STATUS_SUCCESS;
}
CpaStatus cpaDcInstanceGetInfo2(const CpaInstanceHandle h, CpaInstanceInfo2 *info)
{
    memset(info, 0, sizeof *info);
    info->physInstId.packageId = (Cpa16U)((FakeInst *)h)->device;
    info->requiresPhysicallyContiguousMemory = env_int("FAKEQAT_SVM", 0) ? CPA_FALSE : CPA_TRUE;
    info->isPolled = CPA_TRUE;
    snprintf(info->partName, sizeof info->partName, "fake-4xxx");
    return CPA_STATUS_SUCCESS;
}
CpaStatus cpaDcQueryCapabilities(CpaInstanceHandle h, CpaDcInstanceCapabilities *cap)
{
    (void)h;
    memset(cap, 0, sizeof *cap);
    cap->statelessLZ4SCompression = env_int("FAKEQAT_NO_LZ4S", 0) ? CPA_FALSE : CPA_TRUE;
    cap->checksumXXHash32 = CPA_TRUE;
    cap->compressAndVerify = CPA_TRUE;
    return CPA_STATUS_SUCCESS;
}
CpaStatus cpaDcBufferListGetMetaSize(const CpaInstanceHandle h, Cpa32U nb, Cpa32U *sz) { (void)h; *sz = 64 * nb; return CPA_STATUS_SUCCESS; }
CpaStatus cpaDcGetNumIntermediateBuffers(CpaInstanceHandle h, Cpa16U *n) { (void)h; *n = 0; return CPA_STATUS_SUCCESS; }
CpaStatus cpaDcSetAddressTranslation(const CpaInstanceHandle h, CpaVirtualToPhysical fn) { (void)h; (void)fn; return CPA_STATUS_SUCCESS; }
CpaStatus cpaDcStartInstance(CpaInstanceHandle h, Cpa16U nb, CpaBufferList **ib) { (void)nb; (void)ib; ((FakeInst *)h)->started = 1; return CPA_STATUS_SUCCESS; }
CpaStatus cpaDcStopInstance(CpaInstanceHandle h) { ((FakeInst *)h)->started = 0; return CPA_STATUS_SUCCESS; }

/* ---- sessions ---------------------------------------------------------------------------------- */
CpaStatus cpaDcGetSessionSize(CpaInstanceHandle h, CpaDcSessionSetupData *sd, Cpa32U *ss, Cpa32U *cs)
{
    (void)h; (void)sd;
    *ss = (Cpa32U)sizeof(FakeSession);
    *cs = 0;
    return CPA_STATUS_SUCCESS;
}
CpaStatus cpaDcInitSession(CpaInstanceHandle h, CpaDcSessionHandle s, CpaDcSessionSetupData *sd, CpaBufferList *ctx, CpaDcCallbackFn cb)
{
    (void)h; (void)ctx;
    if (sd->compType != CPA_DC_LZ4S || sd->sessState != CPA_DC_STATELESS) return CPA_STATUS_UNSUPPORTED;
    FakeSession *fs = (FakeSession *)s;
    fs->cb = cb;
    fs->sd = *sd;
    return CPA_STATUS_SUCCESS;
}
CpaStatus cpaDcRemoveSession(const CpaInstanceHandle h, CpaDcSessionHandle s) { (void)h; (void)s; return CPA_STATUS_SUCCESS; }
CpaStatus cpaDcLZ4SCompressBound(const CpaInstanceHandle h, Cpa32U in, Cpa32U *out) { (void)h; *out = in + in / 8 + 64; return CPA_STATUS_SUCCESS; }   /* a 3-byte match after >= 15 literals costs one byte more than it saves */

/* ---- the LZ4s "engine" ----------------------------------------------------------------------- */
static size_t put_len(unsigned char *dst, size_t pos, size_t rest)
{
    while (rest >= 255) { dst[pos++] = 255; rest -= 255; }
    dst[pos++] = (unsigned char)rest;
    return pos;
}

static size_t emit(unsigned char *dst, size_t pos, const unsigned char *lit, size_t nlit, unsigned off, size_t ml, int last)
{
    const size_t code = last ? 0 : ml - 2;
    const unsigned ln = nlit >= 15 ? 15 : (unsigned)nlit, mn = code >= 15 ? 15 : (unsigned)code;
    dst[pos++] = (unsigned char)((ln << 4) | mn);
    if (ln == 15) pos = put_len(dst, pos, nlit - 15);
    memcpy(dst + pos, lit, nlit);
    pos += nlit;
    if (last) return pos;
    dst[pos++] = (unsigned char)(off & 0xFF);
    dst[pos++] = (unsigned char)(off >> 8);
    if (mn == 15) pos = put_len(dst, pos, code - 15);
    return pos;
}

static size_t lz4s_encode(const unsigned char *src, size_t n, unsigned char *dst, size_t cap, int level)
{
    enum { HBITS = 15 };
    static __thread int table[1 << HBITS];
    for (int i = 0; i < (1 << HBITS); i++) table[i] = -1;
    size_t pos = 0, anchor = 0, p = 0;
    (void)level; (void)cap;
    while (p + 4 <= n) {
        const unsigned v = (unsigned)src[p] | ((unsigned)src[p + 1] << 8) | ((unsigned)src[p + 2] << 16);
        const unsigned h = (v * 2654435761u) >> (32 - HBITS);
        const int cand = table[h];
        table[h] = (int)p;
        if (cand >= 0 && p - (size_t)cand <= 65535 && src[cand] == src[p] && src[cand + 1] == src[p + 1] && src[cand + 2] == src[p + 2]) {
            size_t ml = 3;
            while (p + ml < n && src[cand + ml] == src[p + ml] && ml < 65535) ml++;       /* the decoder keeps 16 bits of matchLength (:1062) */
            pos = emit(dst, pos, src + anchor, p - anchor, (unsigned)(p - (size_t)cand), ml, 0);
            p += ml;
            anchor = p;
        } else {
            p++;
        }
    }
    return emit(dst, pos, src + anchor, n - anchor, 0, 0, 1);
}

CpaStatus cpaDcCompressData2(CpaInstanceHandle h, CpaDcSessionHandle s, CpaBufferList *src, CpaBufferList *dst,
                             CpaDcOpData *op, CpaDcRqResults *res, void *tag)
{
    FakeInst *fi = (FakeInst *)h;
    FakeSession *fs = (FakeSession *)s;
    (void)op;
    if (!fi->started || env_int("FAKEQAT_FAIL_SUBMIT", 0)) return CPA_STATUS_FAIL;
    if (fi->pending) return CPA_STATUS_RETRY;
    const CpaFlatBuffer *in = src->pBuffers, *out = dst->pBuffers;
    res->status = CPA_DC_OK;
    res->consumed = in->dataLenInBytes;
    res->checksum = 0;
    res->endOfLastBlock = CPA_TRUE;
    res->dataUncompressed = env_int("FAKEQAT_INCOMPRESSIBLE", 0) ? CPA_TRUE : CPA_FALSE;
    res->produced = (Cpa32U)lz4s_encode(in->pData, in->dataLenInBytes, out->pData, out->dataLenInBytes, fs->sd.compLevel);
    fi->pendingCb = fs->cb;
    fi->pendingTag = tag;
    fi->pending = 1;                  /* completes at the next poll, like a ring response */
    return CPA_STATUS_SUCCESS;
}

CpaStatus icp_sal_DcPollInstance(CpaInstanceHandle h, Cpa32U quota)
{
    FakeInst *fi = (FakeInst *)h;
    (void)quota;
    if (!fi->pending) return CPA_STATUS_RETRY;
    fi->pending = 0;
    if (fi->pendingCb) fi->pendingCb(fi->pendingTag, CPA_STATUS_SUCCESS);
    return CPA_STATUS_SUCCESS;
}
