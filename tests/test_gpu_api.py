"""-m gpu: the plugin API and the safety net around the kernels.

  * compressed size through the registered producer against same-level chunked stock libzstd, PER DATA KIND and per
    level (the bar of BASELINE.json: within +1 %; the kinds are slices of the files the bench corpus is made of plus the
    synthetic generators);
  * the whole-buffer hand-off writes ZSTD_Sequence[] straight into a pinned caller array;
  * the on-device verifier (the analogue of compressAndVerify, /root/reference/src/qatseqprod.c:1238) rejects corrupted
    sequences, and verify-on-return passes good ones;
  * a look-ahead batch is never served for a buffer whose bytes changed after the compression it was made for.
"""
import os
import sys

import numpy as np
import pytest
import torch

from tests import datagen
from tests.gpu_util import parse_on_gpu

pytestmark = pytest.mark.gpu
BLOCK = 1 << 17
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def kinds(nbytes):
    """name -> bytes: one slice per category of the image corpus (real files of the container image, present on
    every box) plus the synthetic generators."""
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import corpus
    out = {"synthetic-text": datagen.text_like(nbytes, 3), "synthetic-records": datagen.records(nbytes, 4),
           "synthetic-binary": datagen.binary_like(nbytes, 5)}
    for cat, root, suf, _ in corpus._PLAN:
        if cat in out or cat == "compressed-images" or not os.path.isdir(root):
            continue
        got = bytearray()
        for path in corpus._walk(root, suf):
            if len(got) >= nbytes + (1 << 20):
                break
            try:
                if os.path.islink(path) or not os.path.isfile(path):
                    continue
                got += open(path, "rb").read(corpus._FILE_CAP)
            except OSError:
                continue
        if len(got) >= nbytes // 2:
            out[cat] = bytes(got[-nbytes:]) if len(got) > nbytes else bytes(got)     # the tail: past the first files' headers
    return out


# The bar of BASELINE.json at every level and for every kind: at most +1 % against same-level chunked stock libzstd.
BAR = 0.010


@pytest.mark.parametrize("level", [1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12])
def test_ratio_per_kind_and_level(pkg, oracle, level):
    n = (6 if level <= 6 else 3) * BLOCK
    q = pkg.QatSeqProd
    assert q.startQatDevice() == pkg.QZSTD_OK
    st = q.createSeqProdState()
    worst = []
    try:
        for name, data in kinds(n).items():
            ref = oracle.chunked_compress(data, BLOCK, level)
            buf = np.frombuffer(data, dtype=np.uint8)
            q.hintSource(st, buf.ctypes.data, buf.size, 0)
            r = oracle.compress_with_producer(buf, q.producer, st, chunk=BLOCK, level=level, repcodes=1)
            q.hintSource(st, 0, 0, 0)
            assert r["round_trip"] and r["errors"] == 0, (name, r)
            delta = r["csize"] / ref - 1
            print(f"L{level} {name:18s} {len(data):8d} B  ours {r['csize']:8d}  stock {ref:8d}  {100 * delta:+.2f}%")
            worst.append((delta, name))
        bad = [(f"{100 * d:+.2f}%", k) for d, k in worst if d > BAR]
        assert not bad, f"level {level}: larger than same-level chunked stock by more than {100 * BAR:.1f}%: {bad}"
    finally:
        q.freeSeqProdState(st)
        q.stopQatDevice()


def test_hand_off_into_pinned_array(pkg, oracle, engine):
    """b200sp_sequences_host: dense ZSTD_Sequence[] straight into a pinned caller array (no host-side expansion),
    and through pageable memory (staging + threaded copy); both equal the concatenated model output."""
    data = datagen.mixed_corpus(300 * BLOCK + 4321, seed=91)
    src = torch.frombuffer(bytearray(data), dtype=torch.uint8).pin_memory()
    cap = len(data) // 3 + 8 * 301
    pinned = torch.zeros((cap, 4), dtype=torch.int32).pin_memory()
    n1 = engine.sequences_host(src.data_ptr(), len(data), BLOCK, 3, pinned.data_ptr(), cap)
    pageable = np.zeros((cap, 4), np.uint32)
    n2 = engine.sequences_host(src.data_ptr(), len(data), BLOCK, 3, pageable.ctypes.data, cap)
    assert n1 == n2 > 301
    a = pinned.numpy().view(np.uint32)[:n1]
    assert (a == pageable[:n2]).all()
    pos = 0
    for b in list(range(0, 301, 23)) + [300]:
        want = oracle.model_block(data[b * BLOCK:(b + 1) * BLOCK], 3)
        # find block b's entries: every block ends with its {0, lit, 0} entry
        ends = np.flatnonzero(a[:, 2] == 0)
        lo = 0 if b == 0 else int(ends[b - 1]) + 1
        got = a[lo:int(ends[b]) + 1]
        assert got.shape == want.shape and (got == want).all(), f"block {b}"
    assert len(np.flatnonzero(a[:, 2] == 0)) == 301
    with pytest.raises(pkg.B200SeqProdError):
        engine.sequences_host(src.data_ptr(), len(data), BLOCK, 3, pinned.data_ptr(), 1000)      # too small: refused, not overrun


def test_verifier_rejects_corrupted_sequences(pkg, engine):
    """The on-device verifier must FAIL on bad arrays, not only pass good ones: wrong offset (a false match,
    which libzstd itself would not catch: SURVEY App. B case 11), offset before the block, short sum, a delimiter in
    the middle, matchLength 2."""
    data = datagen.text_like(2 * BLOCK, seed=17)
    dev = torch.device("cuda:0")
    src = torch.frombuffer(bytearray(data) + bytearray(32), dtype=torch.uint8).to(dev)
    seqs = torch.zeros((2, pkg.SEQ_STRIDE, 4), dtype=torch.int32, device=dev)
    counts = torch.zeros(2, dtype=torch.int32, device=dev)
    bad = torch.zeros(2, dtype=torch.int32, device=dev)
    engine.parse_device(src.data_ptr(), len(data), BLOCK, 2, 3, seqs.data_ptr(), counts.data_ptr())
    engine.sync()
    good = seqs.clone()

    def verdict():
        bad.fill_(99)
        engine.verify_device(src.data_ptr(), len(data), BLOCK, 2, seqs.data_ptr(), counts.data_ptr(), bad.data_ptr())
        engine.sync()
        return bad.cpu().numpy()

    assert (verdict() == 0).all()
    host = good.cpu().numpy()
    k = int(np.flatnonzero(host[0, :, 2] >= 8)[5])                  # some real match of block 0
    for what, col, val in (("false match", 0, int(host[0, k, 0]) + 1), ("offset before the block", 0, 1 << 20),
                           ("short sum", 2, int(host[0, k, 2]) - 1), ("delimiter in the middle", 2, 0),
                           ("matchLength 2", 2, 2)):
        seqs.copy_(good)
        seqs[0, k, col] = val
        v = verdict()
        assert v[0] != 0 and v[1] == 0, f"{what}: verdict {v}"
    seqs.copy_(good)
    counts[1] = 0
    assert verdict()[1] != 0                                         # count 0 is an error (App. B case 8)


def test_verify_on_return_passes_good_batches(pkg, oracle, engine):
    data = datagen.mixed_corpus(40 * BLOCK + 99, seed=93)
    assert engine.set_verify(True) is False
    try:
        counts, offsets, seqs = engine.parse_host_numpy(data, level=3)
        want = oracle.model_block(data[5 * BLOCK:6 * BLOCK], 3)
        got = seqs[int(offsets[5]):int(offsets[5]) + int(counts[5])]
        assert got.shape == want.shape and (got == want).all()
    finally:
        assert engine.set_verify(False) is True


def test_look_ahead_cache_is_not_served_for_changed_bytes(pkg, oracle):
    """The hint survives the compression it was given for.  A second ZSTD_compress2 over the same address range with
    NEW bytes and no new hint must not be served the old batch (false matches decode to wrong bytes without any
    error: SURVEY App. B case 11)."""
    q = pkg.QatSeqProd
    assert q.startQatDevice() == pkg.QZSTD_OK
    st = q.createSeqProdState()
    try:
        buf = np.frombuffer(bytearray(datagen.mixed_corpus(20 * BLOCK + 555, seed=95)), dtype=np.uint8)
        q.hintSource(st, buf.ctypes.data, buf.size, 0)
        r1 = oracle.compress_with_producer(buf, q.producer, st, chunk=buf.size, level=3)
        assert r1["round_trip"] and r1["errors"] == 0
        buf[:] = np.frombuffer(datagen.mixed_corpus(buf.size, seed=96), dtype=np.uint8)      # same address, new content
        r2 = oracle.compress_with_producer(buf, q.producer, st, chunk=buf.size, level=3)
        assert r2["round_trip"] and r2["errors"] == 0, r2
        stats = q.getStats(st)
        assert stats["batched"] == 2 * 21                           # both passes were batched, the second one afresh
        # a pass abandoned half way (block 0 comes round again) is parsed again as well
        out = np.zeros((43691, 4), np.uint32)
        n0 = q.qatSequenceProducer(st, out.ctypes.data, 43691, buf.ctypes.data, BLOCK, None, 0, 3, 1 << 17)
        buf[:BLOCK] = np.frombuffer(datagen.text_like(BLOCK, 7), dtype=np.uint8)
        n1 = q.qatSequenceProducer(st, out.ctypes.data, 43691, buf.ctypes.data, BLOCK, None, 0, 3, 1 << 17)
        assert n0 != pkg.ZSTD_SEQUENCE_PRODUCER_ERROR and n1 != pkg.ZSTD_SEQUENCE_PRODUCER_ERROR
        assert oracle.validate(buf[:BLOCK].tobytes(), out[:n1]) == 0
        q.hintSource(st, 0, 0, 0)
    finally:
        q.freeSeqProdState(st)
        q.stopQatDevice()


def test_states_are_dealt_devices_round_robin_and_pool_is_bounded(pkg, monkeypatch):
    """States take the usable devices in turn (the spread of QZSTD_getAndShuffleInstance,
    /root/reference/src/qatseqprod.c:601-630) and the engine pool is finite: exhaustion answers ERROR
    (QZSTD_grabInstance gives up, :905-928), it never blocks."""
    q = pkg.QatSeqProd
    monkeypatch.setenv("QZSTD_MAX_ENGINES", "2")
    q.stopQatDevice()
    assert q.startQatDevice() == pkg.QZSTD_OK
    src = np.frombuffer(datagen.text_like(4096, 3), dtype=np.uint8)
    out = np.zeros((43691, 4), np.uint32)
    sts = [q.createSeqProdState() for _ in range(3)]
    try:
        rcs = [q.qatSequenceProducer(s, out.ctypes.data, 43691, src.ctypes.data, src.size, None, 0, 3, 1 << 17) for s in sts]
        E = pkg.ZSTD_SEQUENCE_PRODUCER_ERROR
        assert rcs[0] != E and rcs[1] != E and rcs[2] == E
        q.freeSeqProdState(sts.pop(0))                               # an engine is released: the third state gets one
        assert q.qatSequenceProducer(sts[-1], out.ctypes.data, 43691, src.ctypes.data, src.size, None, 0, 3, 1 << 17) != E
    finally:
        for s in sts:
            q.freeSeqProdState(s)
        q.stopQatDevice()
        monkeypatch.delenv("QZSTD_MAX_ENGINES")
        q.startQatDevice(); q.stopQatDevice()


def test_transparent_read_ahead(pkg, oracle):
    """No hint, the stock six symbols only: calls that walk a buffer front to back are served from a window the plugin
    read ahead (process_vm_readv + one GPU batch), with sequences identical to the single-block path; bytes that
    changed after they were read ahead are never served from the window (memcmp), and a buffer that ends inside the
    window (the page after it is unreadable) is handled."""
    import mmap
    q = pkg.QatSeqProd
    q.stopQatDevice()
    assert q.startQatDevice() == pkg.QZSTD_OK
    E = pkg.ZSTD_SEQUENCE_PRODUCER_ERROR
    nblk = 40
    data = datagen.mixed_corpus(nblk * BLOCK + 70001, seed=97)
    # the buffer ends exactly at the end of a mapping: whatever follows is not ours to read
    mm = mmap.mmap(-1, (len(data) + 4095) // 4096 * 4096)
    pad = len(mm) - len(data)
    mm[pad:] = data
    buf = np.frombuffer(mm, dtype=np.uint8)[pad:]
    out = np.zeros((43691, 4), np.uint32)

    def walk(st, level=3):
        seqs = []
        for b in range(nblk + 1):
            n = min(BLOCK, len(data) - b * BLOCK)
            rc = q.qatSequenceProducer(st, out.ctypes.data, 43691, buf.ctypes.data + b * BLOCK, n, None, 0, level, 1 << 17)
            assert rc != E, b
            seqs.append(out[:rc].copy())
        return seqs

    st1 = q.createSeqProdState()
    try:
        # reference run on the single-block path: the model is the judge, block by block
        want = [oracle.model_block(data[b * BLOCK:(b + 1) * BLOCK], 3) for b in range(nblk + 1)]
        got = walk(st1)
        stats = q.getStats(st1)
        assert stats["batched"] >= nblk - 2, stats           # everything but the first blocks came from windows
        for b in range(nblk + 1):
            assert got[b].shape == want[b].shape and (got[b] == want[b]).all(), f"block {b}"
        # through libzstd: one frame over the whole buffer, no hint
        r = oracle.compress_with_producer(buf, q.producer, st1, chunk=len(data), level=3)
        assert r["round_trip"] and r["errors"] == 0
        # bytes changed after they were read ahead: the window must not be served for them
        q.qatSequenceProducer(st1, out.ctypes.data, 43691, buf.ctypes.data, BLOCK, None, 0, 3, 1 << 17)
        q.qatSequenceProducer(st1, out.ctypes.data, 43691, buf.ctypes.data + BLOCK, BLOCK, None, 0, 3, 1 << 17)   # window built here
        fresh = np.frombuffer(datagen.text_like(BLOCK, 77), dtype=np.uint8)
        buf[2 * BLOCK:3 * BLOCK] = fresh
        rc = q.qatSequenceProducer(st1, out.ctypes.data, 43691, buf.ctypes.data + 2 * BLOCK, BLOCK, None, 0, 3, 1 << 17)
        assert rc != E and oracle.validate(fresh.tobytes(), out[:rc]) == 0
        w = oracle.model_block(fresh.tobytes(), 3)
        assert out[:rc].shape == w.shape and (out[:rc] == w).all()
        # smaller chunks (32 KiB each its own call), another level, a walk that jumps back in the middle
        st2 = q.createSeqProdState()
        small = 32768
        order = list(range(0, 30)) + list(range(10, 45))
        for b in order:
            blk = buf[b * small:(b + 1) * small].tobytes()        # (block 2 of the buffer was rewritten above)
            rc = q.qatSequenceProducer(st2, out.ctypes.data, 43691, buf.ctypes.data + b * small, small, None, 0, 6, 1 << 17)
            assert rc != E
            w = oracle.model_block(blk, 6)
            assert out[:rc].shape == w.shape and (out[:rc] == w).all(), f"32 KiB chunk {b}"
        assert q.getStats(st2)["batched"] >= len(order) - 6
        q.freeSeqProdState(st2)
    finally:
        q.freeSeqProdState(st1)
        q.stopQatDevice()


def test_indexed_hand_off_and_registered_buffers(pkg, oracle):
    """QZSTD_generateSequencesIndexed: the per-block index cuts the dense array exactly at the block delimiters, so
    ranges of blocks can be entropy-coded independently (one ZSTD_compressSequences frame per range, the multi-thread
    hand-off of tools/handoff.c); with the caller's buffers page-locked by QZSTD_registerBuffer (the SVM analogue,
    /root/reference/src/qatseqprod.c:1222-1227) the result is the same."""
    q = pkg.QatSeqProd
    assert q.startQatDevice() == pkg.QZSTD_OK
    st = q.createSeqProdState()
    try:
        data = datagen.mixed_corpus(37 * BLOCK + 4242, seed=99)
        src = np.frombuffer(bytearray(data), dtype=np.uint8)
        nb = 38
        cap = len(data) // 3 + 8 * nb
        runs = []
        for pinned in (False, True):
            out = np.zeros((cap, 4), np.uint32)
            index = np.zeros(nb + 1, np.uint64)
            if pinned:
                assert pkg.lib.QZSTD_registerBuffer(src.ctypes.data, src.size) == pkg.QZSTD_OK
                assert pkg.lib.QZSTD_registerBuffer(out.ctypes.data, out.nbytes) == pkg.QZSTD_OK
            n = pkg.lib.QZSTD_generateSequencesIndexed(st, out.ctypes.data, cap, src.ctypes.data, src.size, 0, 3, index.ctypes.data, nb + 1)
            if pinned:
                assert pkg.lib.QZSTD_unregisterBuffer(src.ctypes.data) == pkg.QZSTD_OK
                assert pkg.lib.QZSTD_unregisterBuffer(out.ctypes.data) == pkg.QZSTD_OK
            assert n != pkg.ZSTD_SEQUENCE_PRODUCER_ERROR and int(index[nb]) == n and int(index[0]) == 0
            for b in (0, 1, 17, 36, 37):
                got = out[int(index[b]):int(index[b + 1])]
                want = oracle.model_block(data[b * BLOCK:(b + 1) * BLOCK], 3)
                assert got.shape == want.shape and (got == want).all(), f"block {b} (pinned={pinned})"
                assert got[-1, 0] == 0 and got[-1, 2] == 0                 # every block closes with its delimiter
            runs.append(out[:n].copy())
        assert runs[0].shape == runs[1].shape and (runs[0] == runs[1]).all()
        # a range of blocks is a self-contained ZSTD_compressSequences job
        lo, hi = 8, 16
        part = data[lo * BLOCK:hi * BLOCK]
        r = oracle.compress_sequences(part, runs[0][int(index[lo]):int(index[hi])], level=3)
        assert r["round_trip"]
        # too small an index array is refused
        small = np.zeros(nb, np.uint64)
        assert pkg.lib.QZSTD_generateSequencesIndexed(st, runs[0].ctypes.data, cap, src.ctypes.data, src.size, 0, 3, small.ctypes.data, nb) == pkg.ZSTD_SEQUENCE_PRODUCER_ERROR
    finally:
        q.freeSeqProdState(st)
        q.stopQatDevice()


def test_hand_off_tool_multi_thread(pkg, tmp_path):
    """tools/qzstd_handoff: GPU sequences for the whole buffer + the entropy stage on four host threads, frames verified."""
    import subprocess
    tool = os.path.join(ROOT, "tools", "qzstd_handoff")
    if not os.path.exists(tool):
        subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "tools")], check=True)
    f = tmp_path / "in.bin"
    f.write_bytes(datagen.mixed_corpus(70 * BLOCK + 999, seed=98))
    r = subprocess.run([tool, "-t4", "-l2", "-L3", "-f4", "-p4", str(f)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "PASS" in r.stdout, r.stdout + r.stderr


def test_producer_rate_tool(pkg, tmp_path):
    """tools/qzstd_producer_rate: four threads through the stock qatSequenceProducer, every call answered (PASS)."""
    import subprocess
    tool = os.path.join(ROOT, "tools", "qzstd_producer_rate")
    if not os.path.exists(tool):
        subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "tools")], check=True)
    f = tmp_path / "in.bin"
    f.write_bytes(datagen.mixed_corpus(90 * BLOCK + 31, seed=94))
    for mode in ("-m1", "-m0"):
        r = subprocess.run([tool, mode, "-t4", "-l2", "-L3", str(f)], capture_output=True, text=True, timeout=300)
        assert r.returncode == 0 and "PASS" in r.stdout, r.stdout + r.stderr
