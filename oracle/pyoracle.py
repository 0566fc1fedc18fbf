"""ctypes bindings of oracle/_build/liboracle.so — TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module, and only as the checker or as the timed CPU baseline.  The product
(qat-zstd-plugin_b200/) never does.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from ctypes import POINTER, Structure, c_char_p, c_double, c_int, c_size_t, c_uint32, c_void_p

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_build", "liboracle.so")


class Sequence(Structure):
    _fields_ = [("offset", c_uint32), ("litLength", c_uint32), ("matchLength", c_uint32), ("rep", c_uint32)]


class ModelParams(Structure):
    _fields_ = [("keyBytes", c_int), ("scan", c_int), ("rank16", c_int), ("minMatch", c_int), ("extCap", c_int),
                ("lazyDepth", c_int), ("window", c_int), ("backExt", c_int), ("repParse", c_int), ("domBias", c_int), ("nearN", c_int)]


def build() -> None:
    subprocess.run(["make", "-s", "-C", _HERE, "_build/liboracle.so"], check=True)


def _load():
    if not os.path.exists(LIB_PATH):
        build()
    l = ctypes.CDLL(LIB_PATH)
    l.seqmodel_params_for_level.argtypes = [c_int, POINTER(ModelParams)]
    l.seqmodel_block.argtypes = [c_void_p, c_size_t, c_void_p, c_size_t, POINTER(ModelParams)]
    l.seqmodel_block.restype = c_size_t
    l.lanemodel_block.argtypes = [c_void_p, c_size_t, c_void_p, c_size_t, POINTER(ModelParams)]
    l.lanemodel_block.restype = c_size_t
    l.oracle_sw_create.restype = c_void_p
    l.oracle_sw_free.argtypes = [c_void_p]
    l.oracle_sw_producer.argtypes = [c_void_p, c_void_p, c_size_t, c_void_p, c_size_t, c_void_p, c_size_t, c_int, c_size_t]
    l.oracle_sw_producer.restype = c_size_t
    l.oracle_chunked_compress.argtypes = [c_void_p, c_size_t, c_size_t, c_int, c_void_p, c_size_t]
    l.oracle_chunked_compress.restype = c_size_t
    l.oracle_compress_with_producer.argtypes = [c_void_p, c_size_t, c_size_t, c_int, c_void_p, c_void_p, c_int, c_int,
                                                c_int, POINTER(c_size_t), POINTER(c_size_t), POINTER(c_int)]
    l.oracle_compress_with_producer.restype = c_size_t
    l.oracle_compress_sequences.argtypes = [c_void_p, c_size_t, c_void_p, c_size_t, c_int, c_int, POINTER(c_int)]
    l.oracle_compress_sequences.restype = c_size_t
    l.oracle_validate_sequences.argtypes = [c_void_p, c_size_t, c_void_p, c_size_t, POINTER(c_size_t)]
    l.oracle_validate_sequences.restype = c_int
    l.oracle_declz4s.argtypes = [c_void_p, c_size_t, c_void_p, ctypes.c_uint]
    l.oracle_declz4s.restype = c_size_t
    l.oracle_enclz4s.argtypes = [c_void_p, c_size_t, c_void_p, c_size_t]
    l.oracle_enclz4s.restype = c_size_t
    l.oracle_cpu_bench.argtypes = [c_void_p, c_size_t, c_size_t, c_int, c_int, c_int, c_int, POINTER(c_size_t), POINTER(c_size_t)]
    l.oracle_cpu_bench.restype = c_double
    return l


lib = _load()
ERROR = ctypes.c_size_t(-1).value
SEQ_BOUND_128K = 43691


def _u8(data) -> np.ndarray:
    if isinstance(data, np.ndarray):
        return np.ascontiguousarray(data, dtype=np.uint8)
    return np.frombuffer(data, dtype=np.uint8)


def model_params(level: int) -> ModelParams:
    p = ModelParams()
    lib.seqmodel_params_for_level(level, ctypes.byref(p))
    return p


def model_block(data, level: int = 3) -> np.ndarray:
    """Serial model of the GPU match finder on ONE block (<= 128 KiB) -> sequences[n, 4] u32."""
    a = _u8(data)
    cap = a.size // 3 + 8
    out = np.zeros((cap, 4), np.uint32)
    prm = model_params(level)
    n = lib.seqmodel_block(a.ctypes.data, a.size, out.ctypes.data, cap, ctypes.byref(prm))
    if n == ERROR:
        raise RuntimeError("seqmodel_block failed")
    return out[:n].copy()


def lane_model_block(data, level: int = 3) -> np.ndarray:
    """Lane-level statement of the kernel's parse stage (oracle/lanemodel.c) on ONE block; must equal model_block."""
    a = _u8(data)
    cap = a.size // 3 + 8
    out = np.zeros((cap, 4), np.uint32)
    prm = model_params(level)
    n = lib.lanemodel_block(a.ctypes.data, a.size, out.ctypes.data, cap, ctypes.byref(prm))
    if n == ERROR:
        raise RuntimeError("lanemodel_block failed")
    return out[:n].copy()


def sw_block(data, level: int = 3) -> np.ndarray:
    """Software sequence producer (per-block ZSTD_generateSequences) -> sequences[n, 4] u32."""
    a = _u8(data)
    cap = a.size // 3 + 8
    out = np.zeros((cap, 4), np.uint32)
    st = lib.oracle_sw_create()
    try:
        n = lib.oracle_sw_producer(st, out.ctypes.data, cap, a.ctypes.data, a.size, None, 0, level, 1 << 17)
    finally:
        lib.oracle_sw_free(st)
    if n == ERROR:
        raise RuntimeError("oracle_sw_producer failed")
    return out[:n].copy()


def validate(data, seqs: np.ndarray) -> int:
    """0 when the sequences replay to exactly `data`; negative code otherwise (see zstd_oracle.h)."""
    a = _u8(data)
    s = np.ascontiguousarray(seqs, dtype=np.uint32)
    return lib.oracle_validate_sequences(a.ctypes.data, a.size, s.ctypes.data, s.shape[0], None)


def chunked_compress(data, chunk: int = 1 << 17, level: int = 3) -> int:
    a = _u8(data)
    n = lib.oracle_chunked_compress(a.ctypes.data, a.size, chunk, level, None, 0)
    if n == ERROR:
        raise RuntimeError("oracle_chunked_compress failed")
    return n


def compress_with_producer(data, producer_ptr, state, *, chunk: int = 1 << 17, level: int = 3, repcodes: int = 1,
                           fallback: int = 0, validate_sequences: int = 1):
    """Chunked ZSTD_compress2 through a registered producer; returns dict(csize, calls, errors, round_trip)."""
    a = _u8(data)
    calls, errs, ok = c_size_t(), c_size_t(), c_int()
    n = lib.oracle_compress_with_producer(a.ctypes.data, a.size, chunk, level, producer_ptr, state, repcodes, fallback,
                                          validate_sequences, ctypes.byref(calls), ctypes.byref(errs), ctypes.byref(ok))
    return {"csize": None if n == ERROR else n, "calls": calls.value, "errors": errs.value, "round_trip": bool(ok.value)}


def compress_sequences(data, seqs: np.ndarray, level: int = 3, repcodes: int = 1):
    """ZSTD_compressSequences over an explicit-delimiter array + round trip -> {"csize", "round_trip"}."""
    a = _u8(data)
    s = np.ascontiguousarray(seqs, dtype=np.uint32)
    ok = c_int(0)
    c = lib.oracle_compress_sequences(a.ctypes.data, a.size, s.ctypes.data, s.shape[0], level, repcodes, ctypes.byref(ok))
    return {"csize": None if c == ERROR else int(c), "round_trip": bool(ok.value)}


def sw_producer_ptr():
    return ctypes.cast(lib.oracle_sw_producer, c_void_p)


def cpu_bench(data, chunk: int, level: int, mode: int, threads: int, iters: int = 1):
    """mode 0 = ZSTD_compress2 per chunk; mode 1 = per-block ZSTD_generateSequences.
    Returns (bytes_per_second, out_bytes, n_seq)."""
    a = _u8(data)
    ob, ns = c_size_t(), c_size_t()
    bps = lib.oracle_cpu_bench(a.ctypes.data, a.size, chunk, level, mode, threads, iters, ctypes.byref(ob), ctypes.byref(ns))
    return bps, ob.value, ns.value


def declz4s(stream: bytes, capacity: int = SEQ_BOUND_128K) -> np.ndarray | None:
    a = _u8(stream)
    out = np.zeros((capacity, 4), np.uint32)
    n = lib.oracle_declz4s(out.ctypes.data, capacity, a.ctypes.data, a.size)
    return None if n == ERROR else out[:n].copy()


def enclz4s(seqs: np.ndarray) -> bytes | None:
    s = np.ascontiguousarray(seqs, dtype=np.uint32)
    cap = int(s[:, 1].sum()) + 8 * s.shape[0] + 1024
    dst = np.zeros(cap, np.uint8)
    n = lib.oracle_enclz4s(dst.ctypes.data, cap, s.ctypes.data, s.shape[0])
    return None if n == ERROR else dst[:n].tobytes()
