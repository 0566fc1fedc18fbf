"""Runs ONE scenario against the reference plugin built with the fake QAT driver
(oracle/_ref/libqzstd_ref.so = unmodified /root/reference/src/qatseqprod.c + oracle/refstub/fakeqat.c)
or against our library, in a fresh process (both keep process-global device state), and prints JSON.

usage: python tests/ref_scenarios.py <ref|ours> <scenario>
TEST INFRASTRUCTURE ONLY.
"""
import ctypes
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import pyoracle as oracle          # noqa: E402
from tests import datagen                       # noqa: E402

ERR = ctypes.c_size_t(-1).value
BLOCK = 1 << 17


def load(which):
    if which == "ref":
        lib = ctypes.CDLL(os.path.join(ROOT, "oracle", "_ref", "libqzstd_ref.so"))
    else:
        import __graft_entry__ as g
        lib = g.load_package().lib
    lib.QZSTD_version.restype = ctypes.c_char_p
    lib.QZSTD_createSeqProdState.restype = ctypes.c_void_p
    lib.QZSTD_freeSeqProdState.argtypes = [ctypes.c_void_p]
    lib.qatSequenceProducer.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_size_t,
                                        ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_size_t]
    lib.qatSequenceProducer.restype = ctypes.c_size_t
    return lib


def call(lib, st, src, out, cap=43691, dict_=None, dsize=0, level=3, window=1 << 17, size=None):
    return lib.qatSequenceProducer(st, out.ctypes.data, cap, src.ctypes.data, src.size if size is None else size,
                                   dict_, dsize, level, window)


def main():
    which, scenario = sys.argv[1], sys.argv[2]
    lib = load(which)
    res = {"version": lib.QZSTD_version().decode()}
    res["start"] = lib.QZSTD_startQatDevice()
    res["start_again"] = lib.QZSTD_startQatDevice()
    st = lib.QZSTD_createSeqProdState()
    out = np.zeros((43691, 4), np.uint32)
    src = np.frombuffer(datagen.text_like(BLOCK, 5), dtype=np.uint8)

    if scenario == "rejections":
        cases = {
            "ok": call(lib, st, src, out),
            "dict_ptr": call(lib, st, src, out, dict_=src.ctypes.data),
            "dict_size": call(lib, st, src, out, dsize=16),
            "window_small": call(lib, st, src, out, window=1 << 14),
            "window_eq_small_src": call(lib, st, src, out, window=1000, size=1000),
            "level_0": call(lib, st, src, out, level=0),
            "level_13": call(lib, st, src, out, level=13),
            "level_1": call(lib, st, src, out, level=1),
            "level_12": call(lib, st, src, out, level=12),
            "tiny_capacity": call(lib, st, src, out, cap=10),
        }
        res["is_error"] = {k: v == ERR for k, v in cases.items()}
    elif scenario == "sequences":
        seqs_ok, total = True, 0
        for kind, seed in ((datagen.text_like, 1), (datagen.records, 2), (datagen.binary_like, 3), (datagen.zeros, None)):
            data = kind(BLOCK, seed) if seed is not None else kind(BLOCK)
            a = np.frombuffer(data, dtype=np.uint8)
            n = call(lib, st, a, out)
            ok = n != ERR and oracle.validate(data, out[:n]) == 0 and out[n - 1, 0] == 0 and out[n - 1, 2] == 0
            seqs_ok &= bool(ok)
            total += 0 if n == ERR else int(n)
        res["all_valid_and_last_entry_is_literals"] = seqs_ok
        res["total_sequences"] = total
    elif scenario == "roundtrip":
        data = datagen.mixed_corpus(9 * BLOCK + 333, seed=4)
        r = oracle.compress_with_producer(data, ctypes.cast(lib.qatSequenceProducer, ctypes.c_void_p), st,
                                          chunk=len(data), level=3, repcodes=1, fallback=0)
        res.update({"round_trip": r["round_trip"], "errors": r["errors"], "calls": r["calls"], "csize": r["csize"]})
    elif scenario == "uncompressible":
        a = np.frombuffer(datagen.rand_bytes(BLOCK, 9), dtype=np.uint8)
        n = call(lib, st, a, out)
        res["count"] = None if n == ERR else int(n)
        res["first"] = out[0, :3].tolist()
    elif scenario == "down":
        res["errors_in_1001_calls"] = sum(call(lib, st, src, out) == ERR for _ in range(1001))
    lib.QZSTD_freeSeqProdState(st)
    lib.QZSTD_stopQatDevice()
    res["start_after_stop"] = lib.QZSTD_startQatDevice()
    lib.QZSTD_stopQatDevice()
    print(json.dumps(res))


if __name__ == "__main__":
    main()
