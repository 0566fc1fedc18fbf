"""qat-zstd-plugin_b200 — B200-native ZSTD block-level sequence producer.

Python host-side mirror of the two C interfaces exported by the in-tree ``libqatseqprod.so``:

* ``include/qatseqprod.h`` — the drop-in surface of intel/QAT-ZSTD-Plugin
  (/root/reference/src/qatseqprod.h:72,110-116,130,137,145,151): ``QZSTD_version``,
  ``QZSTD_startQatDevice``, ``QZSTD_stopQatDevice``, ``QZSTD_createSeqProdState``,
  ``QZSTD_freeSeqProdState`` and the ``qatSequenceProducer`` callback that stock libzstd calls
  once per block after ``ZSTD_registerSequenceProducer``.
* ``include/b200seqprod.h`` — the C-ABI batching layer (engine, device-resident batch parse,
  host batch parse, wire-format expansion, on-device verify).

Everything that computes runs in the CUDA library; there is NO CPU fallback here.  Importing the
package without the built library raises ``ImportError`` with the build command, and every compute
entry point raises ``B200SeqProdError`` if the device is missing or the call fails.  PyTorch is
only used by callers for device memory and streams; this module takes raw pointers.

The directory name contains a hyphen, so load it with ``__graft_entry__.load_package()`` (alias
``qat_zstd_plugin_b200``) or ``importlib`` rather than a plain ``import`` statement.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import (POINTER, Structure, c_char_p, c_int, c_size_t, c_uint32, c_uint64, c_void_p,
                    c_ulonglong)

__all__ = [
    "LIB_PATH", "lib", "B200SeqProdError", "Sequence", "Engine", "QatSeqProd", "ZstdLib",
    "BLOCK_MAX", "SEQ_STRIDE", "QZSTD_OK", "QZSTD_STARTED", "QZSTD_FAIL", "QZSTD_UNSUPPORTED",
    "ZSTD_SEQUENCE_PRODUCER_ERROR",
]

_HERE = os.path.dirname(os.path.abspath(__file__))
# B200SP_LIB lets a developer A/B-test another build of the same library; it is still the CUDA library.
LIB_PATH = os.environ.get("B200SP_LIB") or os.path.join(_HERE, "libqatseqprod.so")

BLOCK_MAX = 1 << 17
SEQ_STRIDE = 43696
QZSTD_OK, QZSTD_STARTED, QZSTD_FAIL, QZSTD_UNSUPPORTED = 0, 1, -1, -2
ZSTD_SEQUENCE_PRODUCER_ERROR = ctypes.c_size_t(-1).value


class B200SeqProdError(RuntimeError):
    """A C-ABI call returned a negative B200SP_E* code."""


class Sequence(Structure):
    """ZSTD_Sequence / b200sp_sequence: 4 x u32."""
    _fields_ = [("offset", c_uint32), ("litLength", c_uint32), ("matchLength", c_uint32), ("rep", c_uint32)]


class _Result(Structure):
    _fields_ = [("nBlocks", c_uint32), ("counts", POINTER(c_uint32)), ("offsets", POINTER(c_uint64)),
                ("packed", POINTER(c_uint64))]


PRODUCER_F = ctypes.CFUNCTYPE(c_size_t, c_void_p, POINTER(Sequence), c_size_t, c_void_p, c_size_t,
                              c_void_p, c_size_t, c_int, c_size_t)


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is not built. Run `python -c 'import __graft_entry__ as g; g.build()'` "
            "(or `make -C qat-zstd-plugin_b200/csrc`). There is no CPU fallback.")
    l = ctypes.CDLL(LIB_PATH)
    # --- qatseqprod.h
    l.QZSTD_version.restype = c_char_p
    l.QZSTD_startQatDevice.restype = c_int
    l.QZSTD_stopQatDevice.restype = None
    l.QZSTD_createSeqProdState.restype = c_void_p
    l.QZSTD_freeSeqProdState.argtypes = [c_void_p]
    l.QZSTD_freeSeqProdState.restype = None
    l.qatSequenceProducer.argtypes = [c_void_p, c_void_p, c_size_t, c_void_p, c_size_t, c_void_p, c_size_t,
                                      c_int, c_size_t]
    l.qatSequenceProducer.restype = c_size_t
    l.QZSTD_hintSource.argtypes = [c_void_p, c_void_p, c_size_t, c_size_t]
    l.QZSTD_hintSource.restype = None
    l.QZSTD_getStats.argtypes = [c_void_p, POINTER(c_ulonglong), POINTER(c_ulonglong), POINTER(c_ulonglong)]
    l.QZSTD_getStats.restype = None
    l.QZSTD_generateSequences.argtypes = [c_void_p, c_void_p, c_size_t, c_void_p, c_size_t, c_size_t, c_int]
    l.QZSTD_generateSequences.restype = c_size_t
    l.QZSTD_generateSequencesIndexed.argtypes = [c_void_p, c_void_p, c_size_t, c_void_p, c_size_t, c_size_t, c_int, c_void_p, c_size_t]
    l.QZSTD_generateSequencesIndexed.restype = c_size_t
    l.QZSTD_registerBuffer.argtypes = [c_void_p, c_size_t]
    l.QZSTD_registerBuffer.restype = c_int
    l.QZSTD_unregisterBuffer.argtypes = [c_void_p]
    l.QZSTD_unregisterBuffer.restype = c_int
    l.QZSTD_setCoalescing.argtypes = [c_int]
    l.QZSTD_setCoalescing.restype = c_int
    # --- b200seqprod.h
    l.b200sp_driver_device_count.restype = c_int
    l.b200sp_device_count.restype = c_int
    l.b200sp_engine_create.argtypes = [c_int, POINTER(c_void_p)]
    l.b200sp_engine_create.restype = c_int
    l.b200sp_engine_destroy.argtypes = [c_void_p]
    l.b200sp_engine_destroy.restype = None
    l.b200sp_engine_device.argtypes = [c_void_p]
    l.b200sp_engine_sm_count.argtypes = [c_void_p]
    l.b200sp_parse_device.argtypes = [c_void_p, c_void_p, c_uint64, c_uint32, c_uint64, c_void_p, c_uint32, c_int,
                                      c_void_p, c_uint64, c_void_p, c_void_p]
    l.b200sp_parse_device.restype = c_int
    l.b200sp_verify_device.argtypes = [c_void_p, c_void_p, c_uint64, c_uint32, c_uint64, c_void_p, c_uint32,
                                       c_void_p, c_uint64, c_void_p, c_void_p, c_void_p]
    l.b200sp_verify_device.restype = c_int
    l.b200sp_sync.argtypes = [c_void_p]
    l.b200sp_sync.restype = c_int
    l.b200sp_parse_host.argtypes = [c_void_p, c_void_p, c_size_t, c_uint32, c_int, POINTER(_Result)]
    l.b200sp_parse_host.restype = c_int
    l.b200sp_sequences_host.argtypes = [c_void_p, c_void_p, c_size_t, c_uint32, c_int, c_void_p, c_size_t,
                                        POINTER(c_size_t), POINTER(_Result)]
    l.b200sp_sequences_host.restype = c_int
    l.b200sp_engine_set_verify.argtypes = [c_void_p, c_int]
    l.b200sp_engine_set_verify.restype = c_int
    l.b200sp_usable_devices.argtypes = [POINTER(c_int), c_int]
    l.b200sp_usable_devices.restype = c_int
    l.b200sp_expand.argtypes = [c_void_p, c_size_t, c_void_p]
    l.b200sp_expand.restype = None
    l.b200sp_error_string.restype = c_char_p
    l.b200sp_version.restype = c_char_p
    return l


lib = _load()

# Every symbol include/qatseqprod.h and include/b200seqprod.h declare (checked by the CPU tests).
EXPORTED_SYMBOLS = [
    "QZSTD_version", "QZSTD_startQatDevice", "QZSTD_stopQatDevice", "QZSTD_createSeqProdState",
    "QZSTD_freeSeqProdState", "qatSequenceProducer", "QZSTD_hintSource", "QZSTD_getStats", "QZSTD_generateSequences", "QZSTD_generateSequencesIndexed", "QZSTD_registerBuffer", "QZSTD_unregisterBuffer",
    "b200sp_host_register", "b200sp_host_unregister",
    "QZSTD_setCoalescing", "b200sp_parse_blocks", "b200sp_stage_reserve", "b200sp_parse_staged",
    "b200sp_driver_device_count", "b200sp_device_count", "b200sp_warmup", "b200sp_engine_create", "b200sp_engine_destroy",
    "b200sp_engine_device", "b200sp_engine_sm_count", "b200sp_parse_device", "b200sp_sync",
    "b200sp_parse_host", "b200sp_expand", "b200sp_verify_device", "b200sp_error_string", "b200sp_version",
    "b200sp_sequences_host", "b200sp_engine_set_verify", "b200sp_usable_devices",
]


def _check(rc: int, what: str) -> None:
    if rc != 0:
        raise B200SeqProdError(f"{what} failed ({rc}): {lib.b200sp_error_string().decode()}")


class Engine:
    """One device + one stream + scratch buffers (``b200sp_engine``)."""

    def __init__(self, device: int = 0):
        self._h = c_void_p()
        _check(lib.b200sp_engine_create(device, ctypes.byref(self._h)), "b200sp_engine_create")

    def close(self) -> None:
        if self._h:
            lib.b200sp_engine_destroy(self._h)
            self._h = c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def sm_count(self) -> int:
        return lib.b200sp_engine_sm_count(self._h)

    @property
    def device(self) -> int:
        return lib.b200sp_engine_device(self._h)

    def parse_device(self, d_src: int, total_size: int, block_size: int, n_blocks: int, level: int,
                     d_seqs: int, d_counts: int, *, stride: int | None = None, d_sizes: int = 0,
                     seq_stride: int = SEQ_STRIDE, stream: int = 0) -> None:
        """Asynchronous batch parse of blocks already resident in device memory (raw pointers)."""
        _check(lib.b200sp_parse_device(self._h, d_src, total_size, block_size,
                                       block_size if stride is None else stride, d_sizes or None, n_blocks,
                                       level, d_seqs, seq_stride, d_counts, stream or None),
               "b200sp_parse_device")

    def verify_device(self, d_src: int, total_size: int, block_size: int, n_blocks: int, d_seqs: int,
                      d_counts: int, d_bad: int, *, stride: int | None = None, d_sizes: int = 0,
                      seq_stride: int = SEQ_STRIDE, stream: int = 0) -> None:
        _check(lib.b200sp_verify_device(self._h, d_src, total_size, block_size,
                                        block_size if stride is None else stride, d_sizes or None, n_blocks,
                                        d_seqs, seq_stride, d_counts, d_bad, stream or None),
               "b200sp_verify_device")

    def sync(self) -> None:
        _check(lib.b200sp_sync(self._h), "b200sp_sync")

    def parse_host(self, h_src: int, size: int, block_size: int, level: int):
        """Synchronous host-buffer batch.  Returns (n_blocks, counts*, offsets*, packed*) as ctypes
        pointers into engine-owned pinned memory, valid until the next call."""
        r = _Result()
        _check(lib.b200sp_parse_host(self._h, h_src, size, block_size, level, ctypes.byref(r)),
               "b200sp_parse_host")
        return r.nBlocks, r.counts, r.offsets, r.packed

    def sequences_host(self, h_src: int, size: int, block_size: int, level: int, h_out: int, out_capacity: int) -> int:
        """Synchronous host-buffer batch with the result as dense ZSTD_Sequence[] written to h_out (raw pointers;
        direct DMA when h_out is pinned).  Returns the number of entries."""
        n = c_size_t(0)
        _check(lib.b200sp_sequences_host(self._h, h_src, size, block_size, level, h_out, out_capacity,
                                         ctypes.byref(n), None), "b200sp_sequences_host")
        return n.value

    def set_verify(self, enable: bool) -> bool:
        """Verify-on-return (the reference's compressAndVerify): replay every block on the device before returning."""
        return bool(lib.b200sp_engine_set_verify(self._h, 1 if enable else 0))

    def parse_host_numpy(self, data, block_size: int = BLOCK_MAX, level: int = 3):
        """Convenience for tests/tools: bytes-like in, (counts, offsets, sequences[n,4] u32) numpy out."""
        import numpy as np
        buf = np.frombuffer(data, dtype=np.uint8) if not hasattr(data, "ctypes") else data
        n, counts, offsets, packed = self.parse_host(buf.ctypes.data, buf.size, block_size, level)
        if n == 0:
            return np.zeros(0, np.uint32), np.zeros(1, np.uint64), np.zeros((0, 4), np.uint32)
        c = np.ctypeslib.as_array(counts, shape=(n,)).copy()
        o = np.ctypeslib.as_array(offsets, shape=(n + 1,)).copy()
        total = int(o[n])
        seqs = np.zeros((total, 4), np.uint32)
        lib.b200sp_expand(ctypes.cast(packed, c_void_p), total, seqs.ctypes.data)
        return c, o, seqs


class QatSeqProd:
    """The reference's plugin API (/root/reference/src/qatseqprod.h), same names and return codes."""

    producer = ctypes.cast(lib.qatSequenceProducer, c_void_p)   # function pointer for ZSTD_registerSequenceProducer

    @staticmethod
    def version() -> str:
        return lib.QZSTD_version().decode()

    @staticmethod
    def startQatDevice() -> int:
        return lib.QZSTD_startQatDevice()

    @staticmethod
    def stopQatDevice() -> None:
        lib.QZSTD_stopQatDevice()

    @staticmethod
    def createSeqProdState() -> int:
        return lib.QZSTD_createSeqProdState()

    @staticmethod
    def freeSeqProdState(state: int) -> None:
        lib.QZSTD_freeSeqProdState(state)

    @staticmethod
    def hintSource(state: int, src: int, size: int, block_size: int = 0) -> None:
        lib.QZSTD_hintSource(state, src, size, block_size)

    @staticmethod
    def getStats(state: int):
        a, b, c = c_ulonglong(), c_ulonglong(), c_ulonglong()
        lib.QZSTD_getStats(state, ctypes.byref(a), ctypes.byref(b), ctypes.byref(c))
        return {"calls": a.value, "errors": b.value, "batched": c.value}

    @staticmethod
    def setCoalescing(enable: bool) -> bool:
        """Cross-thread coalescing of single-block calls (process-wide); returns the previous setting."""
        return bool(lib.QZSTD_setCoalescing(1 if enable else 0))

    @staticmethod
    def generateSequences(state: int, data, level: int = 3, block_size: int = 0):
        """Whole-buffer hand-off (QZSTD_generateSequences): sequences[n, 4] u32 with explicit block
        delimiters, or None on ZSTD_SEQUENCE_PRODUCER_ERROR."""
        import numpy as np
        a = np.frombuffer(data, dtype=np.uint8)
        bs = block_size or (1 << 17)
        cap = a.size // 3 + 8 * ((a.size + bs - 1) // bs) + 16
        out = np.zeros((cap, 4), np.uint32)
        n = lib.QZSTD_generateSequences(state, out.ctypes.data, cap, a.ctypes.data, a.size, block_size, level)
        if n == ctypes.c_size_t(-1).value:
            return None
        return out[:n]

    @staticmethod
    def qatSequenceProducer(state, out_seqs, capacity, src, src_size, dict_=None, dict_size=0, level=3,
                            window_size=1 << 17) -> int:
        return lib.qatSequenceProducer(state, out_seqs, capacity, src, src_size, dict_, dict_size, level,
                                       window_size)


class ZstdLib:
    """Minimal ctypes view of stock libzstd (include/zstd_abi.h) — the library the plugin plugs into.
    Used by the tools and tests to drive ZSTD_compress2 with the registered producer."""

    c_compressionLevel = 100
    c_windowLog = 101
    c_nbWorkers = 400
    c_blockDelimiters = 1008
    c_validateSequences = 1009
    c_enableSeqProducerFallback = 1014
    c_maxBlockSize = 1015
    c_searchForExternalRepcodes = 1016
    ps_auto, ps_enable, ps_disable = 0, 1, 2

    def __init__(self, path: str = "libzstd.so.1"):
        z = ctypes.CDLL(path)
        z.ZSTD_createCCtx.restype = c_void_p
        z.ZSTD_freeCCtx.argtypes = [c_void_p]
        z.ZSTD_CCtx_setParameter.argtypes = [c_void_p, c_int, c_int]
        z.ZSTD_CCtx_setParameter.restype = c_size_t
        z.ZSTD_compress2.argtypes = [c_void_p, c_void_p, c_size_t, c_void_p, c_size_t]
        z.ZSTD_compress2.restype = c_size_t
        z.ZSTD_decompress.argtypes = [c_void_p, c_size_t, c_void_p, c_size_t]
        z.ZSTD_decompress.restype = c_size_t
        z.ZSTD_compressBound.argtypes = [c_size_t]
        z.ZSTD_compressBound.restype = c_size_t
        z.ZSTD_isError.argtypes = [c_size_t]
        z.ZSTD_isError.restype = ctypes.c_uint
        z.ZSTD_getErrorName.argtypes = [c_size_t]
        z.ZSTD_getErrorName.restype = c_char_p
        z.ZSTD_registerSequenceProducer.argtypes = [c_void_p, c_void_p, c_void_p]
        z.ZSTD_registerSequenceProducer.restype = None
        z.ZSTD_sequenceBound.argtypes = [c_size_t]
        z.ZSTD_sequenceBound.restype = c_size_t
        z.ZSTD_versionString.restype = c_char_p
        self.z = z

    def compress_with_producer(self, data, level: int, producer_ptr, state, *, chunk: int = 0,
                               fallback: int = 0, repcodes: int = 1, validate: int = 1) -> bytes:
        """ZSTD_compress2 with a registered producer; chunk > 0 compresses each chunk as its own
        frame (the reference benchmark's -c, /root/reference/test/benchmark.c:300-321)."""
        import numpy as np
        z = self.z
        src = np.frombuffer(data, dtype=np.uint8)
        cctx = z.ZSTD_createCCtx()
        try:
            z.ZSTD_registerSequenceProducer(cctx, state, producer_ptr)
            for k, v in ((self.c_enableSeqProducerFallback, fallback), (self.c_searchForExternalRepcodes, repcodes),
                         (self.c_validateSequences, validate), (self.c_compressionLevel, level)):
                rc = z.ZSTD_CCtx_setParameter(cctx, k, v)
                if z.ZSTD_isError(rc):
                    raise RuntimeError(z.ZSTD_getErrorName(rc).decode())
            out = bytearray()
            step = chunk or max(src.size, 1)
            dst = np.empty(z.ZSTD_compressBound(step) + 64, np.uint8)
            for pos in range(0, max(src.size, 1), step):
                part = src[pos:pos + step]
                n = z.ZSTD_compress2(cctx, dst.ctypes.data, dst.size, part.ctypes.data, part.size)
                if z.ZSTD_isError(n):
                    raise RuntimeError(z.ZSTD_getErrorName(n).decode())
                out += dst[:n].tobytes()
            return bytes(out)
        finally:
            z.ZSTD_freeCCtx(cctx)

    def decompress(self, comp: bytes, size: int) -> bytes:
        import numpy as np
        z = self.z
        c = np.frombuffer(comp, dtype=np.uint8)
        dst = np.empty(max(size, 1), np.uint8)
        n = z.ZSTD_decompress(dst.ctypes.data, size, c.ctypes.data, c.size)
        if z.ZSTD_isError(n):
            raise RuntimeError(z.ZSTD_getErrorName(n).decode())
        return dst[:n].tobytes()
