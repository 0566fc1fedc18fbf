"""Developer A/B harness (not collected by pytest): for each library build given on the command line,
check a sample of blocks bit-for-bit against the serial model and time the bench corpus.

usage: python tests/dev_ab.py build/ab/libX.so [build/ab/libY.so ...]     (spawns one process per build)
       B200SP_LIB=... python tests/dev_ab.py --one
"""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))


def one():
    import ctypes
    import numpy as np, torch
    import __graft_entry__ as g
    import corpus
    from tests import datagen
    pkg = g.load_package(); oracle = g.load_oracle()
    BLOCK = 1 << 17
    eng = pkg.Engine(0); dev = torch.device("cuda:0")
    ts = torch.cuda.Stream(device=dev); torch.cuda.set_stream(ts)
    data, label, info = corpus.load()

    def parse(buf, level):
        n = len(buf); nb = (n + BLOCK - 1) // BLOCK
        src = torch.frombuffer(bytearray(buf) + bytearray(64), dtype=torch.uint8).to(dev)
        seqs = torch.empty((nb, pkg.SEQ_STRIDE, 4), dtype=torch.int32, device=dev)
        counts = torch.zeros(nb, dtype=torch.int32, device=dev)
        f = lambda: eng.parse_device(src.data_ptr(), n, BLOCK, nb, level, seqs.data_ptr(), counts.data_ptr(), stream=ts.cuda_stream)
        return f, seqs, counts, nb

    # parity sample
    bad = 0; checked = 0
    sample = b"".join(data[o:o + BLOCK] for o in range(0, len(data), 97 * BLOCK)) + datagen.zeros(BLOCK) + \
        datagen.periodic(BLOCK, 100) + datagen.text_like(BLOCK + 777, 3)
    for level in (1, 3, 6, 12):
        f, seqs, counts, nb = parse(sample, level); f(); torch.cuda.synchronize()
        hc = counts.cpu().numpy(); hs = seqs.cpu().numpy().view(np.uint32)
        for b in range(nb):
            want = oracle.model_block(sample[b * BLOCK:(b + 1) * BLOCK], level)
            got = hs[b, :hc[b]]
            checked += 1
            if got.shape != want.shape or not (got == want).all():
                bad += 1
                if bad <= 4: print('   differs: level', level, 'block', b, 'counts', hc[b], len(want), flush=True)
    # timing
    out = []
    for level in [int(x) for x in os.environ.get('DEV_AB_LEVELS', '3,6').split(',')]:
        f, seqs, counts, nb = parse(data if level <= 4 else data[:148 * 4 * BLOCK], level)
        for _ in range(3): f()
        torch.cuda.synchronize()
        if os.environ.get("B200SP_ROLE_PROFILE"):
            buf = (ctypes.c_ulonglong * 10)(); pkg.lib.b200sp_debug_role_cycles(eng._h, buf)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        reps = 10 if level <= 4 else 2
        for _ in range(reps): f()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        role = ""
        if os.environ.get("B200SP_ROLE_PROFILE"):
            buf = (ctypes.c_ulonglong * 10)(); pkg.lib.b200sp_debug_role_cycles(eng._h, buf)
            r = list(buf); st = max(r[8], 1)
            role = f" [EH {r[0]/st/26:.0f} TL {r[1]/st:.0f} TS {r[2]/st:.0f} P1a {r[3]/st:.0f} P1b {r[4]/st:.0f} P2a {r[5]/st:.0f} P2b {r[6]/st:.0f} wall {r[7]/st:.0f}]"
        out.append(f"L{level} {ms:.3f} ms {(len(data) if level <= 4 else 148 * 4 * BLOCK)/ms/1e6:.1f} GB/s{role}")
    print(f"{os.path.basename(pkg.LIB_PATH):24s} parity {checked - bad}/{checked}  " + "  ".join(out), flush=True)


if __name__ == "__main__":
    if sys.argv[1:] == ["--one"]:
        one()
    else:
        for lib in sys.argv[1:]:
            env = dict(os.environ, B200SP_LIB=os.path.abspath(lib), B200SP_ROLE_PROFILE="1")
            subprocess.run([sys.executable, os.path.abspath(__file__), "--one"], env=env, cwd=ROOT)
