/*
 * b200seqprod.h — C-ABI of the B200 batching layer under libqatseqprod.so.
 *
 * Plain C, no CUDA or torch types: pointers, sizes and an opaque engine handle.  This is the
 * boundary a foreign-language host (cgo / JNI / ctypes) binds; qatseqprod.h (the libzstd-facing
 * drop-in surface) is implemented on top of it in plain C (csrc/seqprod_host.c).
 *
 * What each entry point replaces in the reference (/root/reference/src/qatseqprod.c):
 *   b200sp_device_count / b200sp_engine_create   icp_sal_userStart + instance discovery
 *                                                (QZSTD_salUserStart :498-527, QZSTD_getAndShuffleInstance :529-663)
 *                                                and per-instance buffers (QZSTD_allocInstMem :685-822)
 *   b200sp_parse_device                          cpaDcCompressData2 submit (:1245-1249) for a whole batch of
 *                                                blocks already resident in device memory
 *   b200sp_sync                                  the icp_sal_DcPollInstance loop (:1263-1272)
 *   b200sp_parse_host                            staging memcpy (:1222-1227) + submit + poll + result fetch
 *   b200sp_expand                                QZSTD_decLz4s (:1013-1091): device wire format -> ZSTD_Sequence[]
 *   b200sp_engine_destroy                        QZSTD_cleanUpInstMem / QZSTD_stopQat (:364-426, :306-333)
 *
 * Error convention: 0 on success, a negative B200SP_E* code otherwise; b200sp_error_string()
 * describes the last failure of the calling thread.  Nothing here aborts or exits.
 */
#ifndef B200SEQPROD_H
#define B200SEQPROD_H

#include <stddef.h>
#include <stdint.h>

#if defined(__cplusplus)
extern "C" {
#endif

#define B200SP_BLOCK_MAX        (1u << 17)      /* ZSTD_BLOCKSIZE_MAX */
#define B200SP_SEQ_STRIDE       43696u          /* >= ZSTD_sequenceBound(128 KiB) = 43691, multiple of 8 */

#define B200SP_OK               0
#define B200SP_ENODEVICE       (-1)   /* no CUDA device / driver */
#define B200SP_EUNSUPPORTED    (-2)   /* device is not sm_100 or lacks 227 KB shared memory per CTA */
#define B200SP_EINVAL          (-3)   /* bad argument (alignment, block size, level outside 1..12) */
#define B200SP_ENOMEM          (-4)
#define B200SP_ECUDA           (-5)   /* CUDA runtime error, see b200sp_error_string() */
#define B200SP_ETIMEOUT        (-6)   /* completion not seen within the reference's 2 s budget (:107, :1267-1270) */
#define B200SP_EVERIFY         (-7)   /* verify-on-return (compressAndVerify, :1238) found a block whose sequences do not replay */

/* One ZSTD_Sequence as the library's own 16-byte type (identical layout to zstd.h's). */
typedef struct {
    uint32_t offset, litLength, matchLength, rep;
} b200sp_sequence;

typedef struct b200sp_engine b200sp_engine;

/* Number of CUDA devices the driver exposes, whatever their architecture; 0 if there is no
 * driver or no device ("icp_sal_userStart failed" in the reference's terms). */
int b200sp_driver_device_count(void);

/* Number of usable (sm_100, enough shared memory) devices; 0 if none; <0 on driver failure. */
int b200sp_device_count(void);

/* The usable devices' indices, in ascending order (at most `capacity` are written); returns how many there are.
 * The plugin layer spreads its states over them round-robin, as the reference spreads instances over its
 * devices (QZSTD_getAndShuffleInstance, /root/reference/src/qatseqprod.c:601-630). */
int b200sp_usable_devices(int *devices, int capacity);

/* Pays the one-time costs of a device up front (CUDA context, module load, opt-in shared memory), so that
 * the first parsed block does not: the counterpart of the instance start-up the reference does inside
 * QZSTD_startQatDevice (cpaDcStartInstance etc., /root/reference/src/qatseqprod.c:824-903).  Optional and
 * idempotent; engines work without it. */
int b200sp_warmup(int device);

/* Engine = one device + one stream + scratch buffers.  Not thread-safe: one engine per caller
 * thread, like one QZSTD state per CCtx (/root/reference/src/qatseqprod.h:139-151). */
int  b200sp_engine_create(int device, b200sp_engine **engine);
void b200sp_engine_destroy(b200sp_engine *engine);
int  b200sp_engine_device(const b200sp_engine *engine);
int  b200sp_engine_sm_count(const b200sp_engine *engine);

/* Device-resident batch: blocks b = 0..nBlocks-1 start at d_src + b*stride.
 *   d_sizes == NULL : block b holds min(blockSize, totalSize - b*stride) bytes (a buffer cut into blocks)
 *   d_sizes != NULL : block b holds d_sizes[b] bytes (device array)
 * d_src must be 16-byte aligned, stride a multiple of 16, every block <= 128 KiB.
 * Output: block b's sequences at d_seqs + b*seqStride (seqStride entries >= ZSTD_sequenceBound of
 * the largest block; B200SP_SEQ_STRIDE always suffices), d_counts[b] = entries written, the last
 * one being {0, trailing literals, 0}.
 * Asynchronous on `cudaStream` (a cudaStream_t passed as void*; NULL = the engine's stream). */
int b200sp_parse_device(b200sp_engine *engine, const void *d_src, uint64_t totalSize,
                        uint32_t blockSize, uint64_t stride, const uint32_t *d_sizes,
                        uint32_t nBlocks, int level, b200sp_sequence *d_seqs, uint64_t seqStride,
                        uint32_t *d_counts, void *cudaStream);

/* Waits for everything queued on the engine's stream. */
int b200sp_sync(b200sp_engine *engine);

/* Host-resident batch, synchronous: copies h_src to the device (through pinned staging unless
 * the buffer is already pinned), parses it in blocks of blockSize, and brings back the result in
 * the 8-byte wire format.  The result arrays are owned by the engine and stay valid until the
 * next b200sp_parse_host / b200sp_engine_destroy on it. */
typedef struct {
    uint32_t nBlocks;
    const uint32_t *counts;     /* [nBlocks] entries per block, incl. the final literals entry */
    const uint64_t *offsets;    /* [nBlocks + 1] start of each block's entries in `packed` */
    const uint64_t *packed;     /* offset | litLength << 17 | matchLength << 35 */
} b200sp_result;

int b200sp_parse_host(b200sp_engine *engine, const void *h_src, size_t srcSize, uint32_t blockSize,
                      int level, b200sp_result *result);

/* Host-resident batch, synchronous, result as the array libzstd consumes: dense ZSTD_Sequence[] (16 bytes each) in
 * h_out, every block ending with its {0, trailing literals, 0} entry - what QZSTD_decLz4s leaves in outSeqs
 * (/root/reference/src/qatseqprod.c:1013-1091), for every block of the buffer at once.  When h_out is pinned
 * (cudaHostAlloc / cudaHostRegister) the entries are copied straight into it; otherwise they go through the engine's
 * pinned staging and a threaded copy.  *nSeqs = entries written; `result` (may be NULL) receives the per-block
 * counts and offsets (its `packed` is NULL).  B200SP_EINVAL when outCapacity is too small. */
int b200sp_sequences_host(b200sp_engine *engine, const void *h_src, size_t srcSize, uint32_t blockSize, int level,
                          b200sp_sequence *h_out, size_t outCapacity, size_t *nSeqs, b200sp_result *result);

/* Verify-on-return, the analogue of the reference's compressAndVerify (/root/reference/src/qatseqprod.c:1238): when
 * enabled (or QZSTD_VERIFY=1 in the environment at engine creation) every host-path call replays each block's
 * sequences against its input on the device before returning and fails with B200SP_EVERIFY if any block does not
 * replay.  Returns the previous setting. */
int b200sp_engine_set_verify(b200sp_engine *engine, int enable);

/* Host-resident batch of SCATTERED blocks (one pointer and one size <= 128 KiB per block), synchronous:
 * what a dispatcher that coalesces the single-block calls of many threads submits - the analogue of
 * several threads sharing the QAT instances (QZSTD_grabInstance, :905-933).  Blocks are gathered into
 * pinned staging at a 128 KiB stride, parsed in one launch, and returned like b200sp_parse_host. */
int b200sp_parse_blocks(b200sp_engine *engine, const void *const *h_blocks, const uint32_t *sizes,
                        uint32_t nBlocks, int level, b200sp_result *result);

/* The two halves of b200sp_parse_blocks, for callers whose threads copy their own blocks in parallel:
 * b200sp_stage_reserve returns the engine's pinned staging area for nSlots blocks (slot k at offset
 * k * 128 KiB; the address is stable until a larger reservation is made); b200sp_parse_staged parses the first
 * nBlocks slots, sizes[k] bytes each. */
int b200sp_stage_reserve(b200sp_engine *engine, uint32_t nSlots, void **slots);
int b200sp_parse_staged(b200sp_engine *engine, const uint32_t *sizes, uint32_t nBlocks, int level,
                        b200sp_result *result);

/* Page-locks a range of the caller's memory so that the host-path calls copy to and from it by DMA directly
 * instead of through the engine's staging: the counterpart of the reference's SVM mode, in which the engine works
 * on the application's buffer rather than on a USDM copy (/root/reference/src/qatseqprod.c:1222-1227).  The range
 * must stay mapped until it is unregistered.  Returns B200SP_OK or B200SP_ECUDA. */
int b200sp_host_register(void *ptr, size_t bytes);
int b200sp_host_unregister(void *ptr);

/* Wire format -> ZSTD_Sequence[] (rep = 0). */
void b200sp_expand(const uint64_t *packed, size_t count, b200sp_sequence *out);

/* On-device verification (the analogue of compressAndVerify, :1238): replays every block's
 * sequences against its input; *d_bad (device, one uint32 per block) gets 0 for a valid block. */
int b200sp_verify_device(b200sp_engine *engine, const void *d_src, uint64_t totalSize,
                         uint32_t blockSize, uint64_t stride, const uint32_t *d_sizes,
                         uint32_t nBlocks, const b200sp_sequence *d_seqs, uint64_t seqStride,
                         const uint32_t *d_counts, uint32_t *d_bad, void *cudaStream);

const char *b200sp_error_string(void);
const char *b200sp_version(void);

#if defined(__cplusplus)
}
#endif
#endif /* B200SEQPROD_H */
