"""Developer timing loop: kernel-only throughput per data kind (CUDA events on the launch stream)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
import numpy as np, torch
import __graft_entry__ as g
import corpus
from tests import datagen
pkg = g.load_package()
BLOCK = 1 << 17
eng = pkg.Engine(0)
dev = torch.device("cuda:0")
ts = torch.cuda.Stream(device=dev); torch.cuda.set_stream(ts)
def run(name, data, level=3, steps=5):
    n = len(data); nb = (n + BLOCK - 1) // BLOCK
    src = torch.frombuffer(bytearray(data) + bytearray(64), dtype=torch.uint8).to(dev)
    seqs = torch.empty((nb, pkg.SEQ_STRIDE, 4), dtype=torch.int32, device=dev)
    counts = torch.zeros(nb, dtype=torch.int32, device=dev)
    f = lambda: eng.parse_device(src.data_ptr(), n, BLOCK, nb, level, seqs.data_ptr(), counts.data_ptr(), stream=ts.cuda_stream)
    for _ in range(2): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps): f()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    ns = int(counts.sum().item())
    if os.environ.get("B200SP_ROLE_PROFILE"):
        import ctypes
        buf = (ctypes.c_ulonglong * 10)()
        pkg.lib.b200sp_debug_role_cycles(eng._h, buf)
        r = list(buf); st = max(r[8], 1)
        print(f"   per-stage cycles: EH-sum {r[0]/st:9.0f} (per warp {r[0]/st/26:7.0f})  TL {r[1]/st:7.0f}  TS {r[2]/st:7.0f}  P1a {r[3]/st:7.0f}  P1b {r[4]/st:7.0f}  P2a {r[5]/st:7.0f}  P2b {r[6]/st:7.0f}  wall {r[7]/st:7.0f}")
    print(f"{name:28s} L{level} {n/1e6:8.1f} MB {nb:5d} blk  {ms:8.3f} ms  {n/ms/1e6:8.1f} GB/s  seq/blk {ns/nb:8.0f}  us/blk/SM {ms*1e3/(nb/148 if nb>148 else 1):7.1f}", flush=True)
data, label, info = corpus.load()
run("image-corpus", data)
if len(sys.argv) > 1 and sys.argv[1] == "all":
    run("image-corpus L1", data, level=1)
    sub = data[:148 * 4 * BLOCK]
    for lv in (5, 6, 7, 9, 12): run("corpus-head L%d" % lv, sub, level=lv, steps=2)
    M = 148 * 4 * BLOCK
    run("text_like", datagen.text_like(8 * BLOCK, 3) * (M // (8 * BLOCK)))
    run("records", datagen.records(M, 4))
    run("binary_like", datagen.binary_like(M, 5))
    run("random", datagen.rand_bytes(M, 6))
    run("zeros", datagen.zeros(M))
    run("periodic100", datagen.periodic(M, 100))
    for off, nm in ((0, "py-source"), (90_000_000, "shared-objects"), (152_000_000, "json"), (180_000_000, "xml")):
        run("corpus@" + nm, data[off: off + 148 * 2 * BLOCK])
