/* refstub/cpa_dc.h — data-compression subset of the fake QAT driver (see cpa.h). TEST INFRASTRUCTURE ONLY. */
#ifndef REFSTUB_CPA_DC_H
#define REFSTUB_CPA_DC_H
#include "cpa.h"

typedef void *CpaDcSessionHandle;
typedef int   CpaDcCompLvl;
typedef enum { CPA_DC_DEFLATE = 3, CPA_DC_LZ4 = 4, CPA_DC_LZ4S = 5 } CpaDcCompType;
typedef enum { CPA_DC_HT_STATIC = 0, CPA_DC_HT_FULL_DYNAMIC = 2 } CpaDcHuffType;
typedef enum { CPA_DC_ASB_DISABLED = 0, CPA_DC_ASB_ENABLED = 1 } CpaDcAutoSelectBest;
typedef enum { CPA_DC_DIR_COMPRESS = 0, CPA_DC_DIR_DECOMPRESS = 1 } CpaDcSessionDir;
typedef enum { CPA_DC_STATEFUL = 0, CPA_DC_STATELESS = 1 } CpaDcSessionState;
typedef enum { CPA_DC_NONE = 0, CPA_DC_CRC32 = 1, CPA_DC_ADLER32 = 2, CPA_DC_XXHASH32 = 4 } CpaDcChecksum;
typedef enum { CPA_DC_MIN_3_BYTE_MATCH = 0, CPA_DC_MIN_4_BYTE_MATCH = 1 } CpaDcCompMinMatch;
typedef enum { CPA_DC_FLUSH_NONE = 0, CPA_DC_FLUSH_FINAL = 1, CPA_DC_FLUSH_SYNC = 2, CPA_DC_FLUSH_FULL = 3 } CpaDcFlush;
typedef enum { CPA_DC_SKIP_DISABLED = 0 } CpaDcSkipMode;
typedef enum { CPA_DC_OK = 0, CPA_DC_OVERFLOW = -11, CPA_DC_VERIFY_ERROR = -18 } CpaDcReqStatus;

typedef struct {
    CpaDcCompLvl compLevel;
    CpaDcCompType compType;
    CpaDcHuffType huffType;
    CpaDcAutoSelectBest autoSelectBestHuffmanTree;
    CpaDcSessionDir sessDirection;
    CpaDcSessionState sessState;
    Cpa32U windowSize;
    CpaDcCompMinMatch minMatch;
    Cpa32U lz4BlockMaxSize;
    CpaBoolean lz4BlockChecksum, lz4BlockIndependence, accumulateXXHash;
    CpaDcChecksum checksum;
} CpaDcSessionSetupData;

typedef struct { CpaDcSkipMode skipMode; Cpa32U skipLength, strideLength, firstSkipOffset; } CpaDcSkipData;
typedef struct {
    CpaDcFlush flushFlag;
    CpaBoolean compressAndVerify, compressAndVerifyAndRecover, integrityCrcCheck, verifyHwIntegrityCrcs;
    CpaDcSkipData inputSkipData, outputSkipData;
    void *pCrcData;
} CpaDcOpData;

typedef struct {
    CpaDcReqStatus status;
    Cpa32U produced, consumed, checksum;
    CpaBoolean endOfLastBlock, dataUncompressed;
} CpaDcRqResults;

typedef struct {
    CpaBoolean statefulLZSCompression, statelessDeflateCompression, statelessLZ4Compression;
    CpaBoolean statelessLZ4SCompression, checksumCRC32, checksumAdler32, checksumXXHash32;
    CpaBoolean dynamicHuffman, compressAndVerify, compressAndVerifyAndRecover;
} CpaDcInstanceCapabilities;

typedef void (*CpaDcCallbackFn)(void *callbackTag, CpaStatus status);

CpaStatus cpaDcGetNumInstances(Cpa16U *pNumInstances);
CpaStatus cpaDcGetInstances(Cpa16U numInstances, CpaInstanceHandle *dcInstances);
CpaStatus cpaDcInstanceGetInfo2(const CpaInstanceHandle h, CpaInstanceInfo2 *info);
CpaStatus cpaDcQueryCapabilities(CpaInstanceHandle h, CpaDcInstanceCapabilities *cap);
CpaStatus cpaDcBufferListGetMetaSize(const CpaInstanceHandle h, Cpa32U numBuffers, Cpa32U *pSizeInBytes);
CpaStatus cpaDcGetNumIntermediateBuffers(CpaInstanceHandle h, Cpa16U *pNumBuffers);
CpaStatus cpaDcSetAddressTranslation(const CpaInstanceHandle h, CpaVirtualToPhysical fn);
CpaStatus cpaDcStartInstance(CpaInstanceHandle h, Cpa16U numBuffers, CpaBufferList **pIntermediateBuffers);
CpaStatus cpaDcStopInstance(CpaInstanceHandle h);
CpaStatus cpaDcGetSessionSize(CpaInstanceHandle h, CpaDcSessionSetupData *sd, Cpa32U *pSessionSize, Cpa32U *pContextSize);
CpaStatus cpaDcInitSession(CpaInstanceHandle h, CpaDcSessionHandle s, CpaDcSessionSetupData *sd,
                           CpaBufferList *pContextBuffer, CpaDcCallbackFn callbackFn);
CpaStatus cpaDcRemoveSession(const CpaInstanceHandle h, CpaDcSessionHandle s);
CpaStatus cpaDcLZ4SCompressBound(const CpaInstanceHandle h, Cpa32U inputSize, Cpa32U *outputSize);
CpaStatus cpaDcCompressData2(CpaInstanceHandle h, CpaDcSessionHandle s, CpaBufferList *pSrcBuff,
                             CpaBufferList *pDestBuff, CpaDcOpData *pOpData, CpaDcRqResults *pResults, void *callbackTag);
#endif
