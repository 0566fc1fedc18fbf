/*
 * refstub/cpa.h — minimal stand-in for the QAT API base header, TEST INFRASTRUCTURE ONLY.
 *
 * Together with cpa_dc.h, icp_sal_user.h, icp_sal_poll.h, qae_mem.h and fakeqat.c this directory is a
 * FAKE user-space QAT driver: just enough types, constants and entry points for the UNMODIFIED
 * reference source (/root/reference/src/qatseqprod.c) to compile and run on a box without QAT
 * hardware or QATlib.  The "device" behind cpaDcCompressData2 is a small software LZ4s encoder
 * (fakeqat.c).  Nothing here is shipped or linked into the product; it only lets the tests execute
 * the reference's own plugin code (argument checks, state machine, submit/poll loop, QZSTD_decLz4s).
 * Written from the call sites in the reference source; it is not a copy of Intel's headers.
 */
#ifndef REFSTUB_CPA_H
#define REFSTUB_CPA_H
#include <stdint.h>

typedef uint8_t  Cpa8U;
typedef uint16_t Cpa16U;
typedef uint32_t Cpa32U;
typedef uint64_t Cpa64U;
typedef int32_t  CpaStatus;
typedef int      CpaBoolean;
typedef void    *CpaInstanceHandle;
typedef uint64_t CpaPhysicalAddr;
typedef CpaPhysicalAddr (*CpaVirtualToPhysical)(void *pVirtualAddr);

#define CPA_TRUE   1
#define CPA_FALSE  0

#define CPA_STATUS_SUCCESS        (0)
#define CPA_STATUS_FAIL           (-1)
#define CPA_STATUS_RETRY          (-2)
#define CPA_STATUS_RESOURCE       (-3)
#define CPA_STATUS_INVALID_PARAM  (-4)
#define CPA_STATUS_FATAL          (-5)
#define CPA_STATUS_UNSUPPORTED    (-6)

typedef struct { Cpa32U dataLenInBytes; Cpa8U *pData; } CpaFlatBuffer;
typedef struct { Cpa32U numBuffers; CpaFlatBuffer *pBuffers; void *pUserData; void *pPrivateMetaData; } CpaBufferList;

typedef struct { Cpa16U packageId; Cpa16U acceleratorId; Cpa16U executionEngineId; Cpa16U busAddress; Cpa32U kptAcHandle; } CpaPhysicalInstanceId;
typedef struct {
    int accelerationServiceType;
    char vendorName[64], partName[64], swVersion[64], instName[64], instID[128];
    CpaPhysicalInstanceId physInstId;
    Cpa32U coreAffinity[32];
    Cpa8U nodeAffinity;
    int operState;
    CpaBoolean requiresPhysicallyContiguousMemory;
    CpaBoolean isPolled;
    CpaBoolean isOffloaded;
} CpaInstanceInfo2;
#endif
