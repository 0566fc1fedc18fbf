/* refstub/qae_mem.h — fake USDM allocator (see cpa.h). TEST INFRASTRUCTURE ONLY. */
#ifndef REFSTUB_QAE_MEM_H
#define REFSTUB_QAE_MEM_H
#include <stddef.h>
#include <stdint.h>
void *qaeMemAllocNUMA(size_t size, int node, size_t phys_alignment_byte);
void qaeMemFreeNUMA(void **ptr);
uint64_t qaeVirtToPhysNUMA(void *pVirtAddr);
#endif
