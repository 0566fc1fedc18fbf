/* refstub/zstd.h — the image has no zstd.h; the reference includes "zstd.h", so route it to the
 * hand-declared ABI subset (include/zstd_abi.h). TEST INFRASTRUCTURE ONLY. */
#include "zstd_abi.h"
