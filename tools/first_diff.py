"""Developer: where does a build first differ from the serial model?  usage: B200SP_LIB=... python tools/first_diff.py [levels]"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
import numpy as np, torch
import __graft_entry__ as g
import corpus
from tests import datagen
from tests.gpu_util import parse_on_gpu
pkg = g.load_package(); oracle = g.load_oracle()
eng = pkg.Engine(0)
BLOCK = 1 << 17
data, label, info = corpus.load()
sample = b"".join(data[o:o + BLOCK] for o in range(0, len(data), 97 * BLOCK))
levels = [int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "1,3").split(",")]
for rep in range(2):
    for level in levels:
        counts, seqs, bad = parse_on_gpu(pkg, eng, sample, level=level)
        nbad = 0
        for b in range(len(counts)):
            blk = sample[b * BLOCK:(b + 1) * BLOCK]
            got = seqs[b, :counts[b]]; want = oracle.model_block(blk, level)
            if got.shape == want.shape and (got == want).all():
                continue
            nbad += 1
            k = 0
            while k < min(len(got), len(want)) and (got[k] == want[k]).all(): k += 1
            pos = int(got[:k, 1].sum() + got[:k, 2].sum())
            if nbad <= 6:
                print(f"rep {rep} L{level} block {b}: bad {bad[b]} counts {counts[b]} vs {len(want)}; first diff at seq {k} pos {pos} window {pos // 1664} "
                      f"group {(pos % 1664) // 32} lane {pos % 32}\n    gpu   {got[k:k+3, :3].tolist()}\n    model {want[k:k+3, :3].tolist()}", flush=True)
        print(f"rep {rep} L{level}: {nbad} of {len(counts)} blocks differ", flush=True)
