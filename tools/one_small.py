import sys, os
sys.path.insert(0, os.getcwd())
import numpy as np, torch
import __graft_entry__ as g
from tests import datagen
from tests.gpu_util import parse_on_gpu
pkg = g.load_package(); oracle = g.load_oracle()
eng = pkg.Engine(0)
BLOCK = 1 << 17
for name, data in (("t5000", datagen.text_like(5000, seed=5000)), ("text3", datagen.text_like(3 * BLOCK + 777, 11)), ("t5000again", datagen.text_like(5000, seed=5000))):
    counts, seqs, bad = parse_on_gpu(pkg, eng, data)
    for b in range(len(counts)):
        blk = data[b * BLOCK:(b + 1) * BLOCK]
        got = seqs[b, :counts[b]]; want = oracle.model_block(blk, 3)
        eq = got.shape == want.shape and (got == want).all()
        print(name, "block", b, "bad", bad[b], "counts", counts[b], len(want), "equal", eq)
        if not eq:
            k = 0
            while k < min(len(got), len(want)) and (got[k] == want[k]).all(): k += 1
            pos = int(got[:k, 1].sum() + got[:k, 2].sum())
            print("   first diff at seq", k, "pos", pos, "window", pos // 1664, "group", (pos % 1664) // 32, "gpu", got[k:k+3].tolist(), "model", want[k:k+3].tolist())
            break
