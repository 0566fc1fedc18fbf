"""Generates the golden fixtures under tests/golden/ (run from the repo root):

    python tests/golden/make_golden.py

For each small input: <name>.in (bytes), <name>.seq (u32[n,4] ZSTD_Sequence array of the serial model at
L3), <name>.lz4s (the same parse as an LZ4s token stream, App. D of SURVEY.md).  The .lz4s/.seq pairs
are additionally decoded by the REFERENCE's own QZSTD_decLz4s when oracle/_ref/libqzstd_ref.so has been
built (make -C oracle ref), and the script refuses to write fixtures the reference decodes differently.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import pyoracle as oracle          # noqa: E402
from tests import datagen                       # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
CASES = {
    "text4k": datagen.text_like(4096, 101),
    "records6k": datagen.records(6000, 102),
    "binary5k": datagen.binary_like(5000, 103),
    "mixed8k": datagen.text_like(3000, 104) + datagen.zeros(700) + datagen.periodic(2000, 37) + datagen.rand_bytes(2492, 105),
    "tiny": b"abcabcabcabcabcabcabcabcabcabcabcabcXYZ",
}


def ref_decode(stream: bytes):
    import ctypes
    so = os.path.join(ROOT, "oracle", "_ref", "libqzstd_ref.so")
    if not os.path.exists(so):
        return None
    lib = ctypes.CDLL(so)
    lib.ref_decLz4s.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_uint]
    lib.ref_decLz4s.restype = ctypes.c_size_t
    a = np.frombuffer(stream, dtype=np.uint8).copy()
    out = np.zeros((43691, 4), np.uint32)
    n = lib.ref_decLz4s(out.ctypes.data, 43691, a.ctypes.data, a.size)
    return out[:n].copy()


def main():
    names = []
    for name, data in CASES.items():
        seqs = oracle.model_block(data, 3)
        assert oracle.validate(data, seqs) == 0
        lz = oracle.enclz4s(seqs)
        dec = oracle.declz4s(lz)
        assert (dec[:, :3] == seqs[:, :3]).all()
        ref = ref_decode(lz)
        if ref is not None:
            assert ref.shape == dec.shape and (ref[:, :3] == dec[:, :3]).all(), f"{name}: reference decLz4s disagrees"
        open(os.path.join(HERE, name + ".in"), "wb").write(data)
        seqs.astype(np.uint32).tofile(os.path.join(HERE, name + ".seq"))
        open(os.path.join(HERE, name + ".lz4s"), "wb").write(lz)
        names.append(name)
        print(name, len(data), "bytes,", len(seqs), "sequences,", len(lz), "lz4s bytes,",
              "checked against reference decLz4s" if ref is not None else "reference .so absent")
    open(os.path.join(HERE, "index.txt"), "w").write("\n".join(names) + "\n")


if __name__ == "__main__":
    main()
