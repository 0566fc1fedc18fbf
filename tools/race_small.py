"""Developer: two small blocks through the parser, for compute-sanitizer --tool racecheck / memcheck."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import __graft_entry__ as g
from tests import datagen
from tests.gpu_util import parse_on_gpu
pkg = g.load_package()
eng = pkg.Engine(0)
level = int(sys.argv[1]) if len(sys.argv) > 1 else 3
data = datagen.text_like(8192, 5) + datagen.records(8192, 6)
counts, seqs, bad = parse_on_gpu(pkg, eng, data, block_size=8192, level=level)
print("counts", counts, "bad", bad)
