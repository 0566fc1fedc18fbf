"""Developer: one parser launch on a corpus slice, for an `ncu --set full` capture.
usage: python tools/ncu_one.py [level=3] [offset_mb=0] [blocks=296]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
import torch
import __graft_entry__ as g
import corpus
pkg = g.load_package()
BLOCK = 1 << 17
level = int(sys.argv[1]) if len(sys.argv) > 1 else 3
off = int(sys.argv[2]) * 1000000 if len(sys.argv) > 2 else 0
nb = int(sys.argv[3]) if len(sys.argv) > 3 else 296
data, label, info = corpus.load()
data = data[off: off + nb * BLOCK]
eng = pkg.Engine(0)
dev = torch.device("cuda:0")
n = len(data); nb = (n + BLOCK - 1) // BLOCK
src = torch.frombuffer(bytearray(data) + bytearray(64), dtype=torch.uint8).to(dev)
seqs = torch.empty((nb, pkg.SEQ_STRIDE, 4), dtype=torch.int32, device=dev)
counts = torch.zeros(nb, dtype=torch.int32, device=dev)
for _ in range(3):
    eng.parse_device(src.data_ptr(), n, BLOCK, nb, level, seqs.data_ptr(), counts.data_ptr())
    eng.sync()
print("done", nb, int(counts.sum().item()))
