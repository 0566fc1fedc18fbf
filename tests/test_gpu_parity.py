"""-m gpu: the CUDA path through the C-ABI against the oracle.

Parity bars (integer work, so all exact):
  * every block's ZSTD_Sequence array is bit-identical to the serial model (oracle/seqmodel.c);
  * the sequences replay to exactly the input (oracle validator and the on-device verifier);
  * ZSTD_compress2 with qatSequenceProducer registered round-trips through stock libzstd
    (the reference's own and only test, /root/reference/test/test.c:102-136), with zero producer errors;
  * compressed size within +-1 % of chunked stock libzstd at the same level on the mixed corpus.
"""
import numpy as np
import pytest

from tests import datagen
from tests.gpu_util import check_against_model, parse_on_gpu

pytestmark = pytest.mark.gpu
BLOCK = 1 << 17


@pytest.mark.parametrize("maker,seed", [
    (datagen.text_like, 11), (datagen.records, 12), (datagen.binary_like, 13), (datagen.rand_bytes, 14),
])
def test_model_parity_by_data_kind(pkg, oracle, engine, maker, seed):
    data = maker(3 * BLOCK + 777, seed)
    assert check_against_model(pkg, oracle, engine, data) >= 4


def test_model_parity_zero_and_periodic(pkg, oracle, engine):
    for data in (datagen.zeros(2 * BLOCK), datagen.periodic(BLOCK + 5000, 1), datagen.periodic(2 * BLOCK, 3),
                 datagen.periodic(BLOCK, 100), datagen.periodic(BLOCK, 70000), b"ab" * 40000):
        check_against_model(pkg, oracle, engine, data)


@pytest.mark.parametrize("level", [1, 2, 3, 4, 5, 6, 7, 9, 11, 12])
def test_model_parity_all_level_classes(pkg, oracle, engine, level):
    data = datagen.mixed_corpus(5 * BLOCK + 4321, seed=20 + level)
    check_against_model(pkg, oracle, engine, data, level=level)


@pytest.mark.parametrize("n", [0, 1, 2, 3, 7, 8, 9, 15, 16, 17, 31, 32, 33, 63, 64, 100, 255, 256, 257, 1023, 1024, 1025,
                               2047, 2048, 4095, 4096, 5000, 16383, 16384, 16385, 65535, 65536, 65537, 131071, 131072])
def test_ragged_sizes(pkg, oracle, engine, n):
    """Empty, tiny and ragged single blocks, incl. the 16-byte TMA granule and 16 KiB chunk edges."""
    data = datagen.text_like(max(n, 1), seed=n)[:n]
    counts, seqs, bad = parse_on_gpu(pkg, engine, data)
    if n == 0:
        assert counts.shape[0] == 0
        return
    got = seqs[0, :counts[0]]
    want = oracle.model_block(data, 3)
    assert bad[0] == 0 and oracle.validate(data, got) == 0
    assert got.shape == want.shape and (got == want).all()
    assert got[-1, 0] == 0 and got[-1, 2] == 0          # last entry = trailing literals (reference convention)


@pytest.mark.parametrize("level", [5, 6, 9, 12])
def test_rep_parse_ragged_sizes_and_window_edges(pkg, oracle, engine, level):
    """Levels 5-12 (the serial repcode-aware parse on one warp): sizes around 0, the 8-byte hash limit, the 32-position
    span and the 1664-position pipeline window (a decision whose look-ahead needs the next window is deferred), in ONE
    batch with per-block sizes; repetitive and record-like content so that repeated offsets, capped matches (> 256 B)
    and rep2 matches occur right at the edges."""
    W = 1664
    sizes = [0, 1, 7, 8, 9, 12, 31, 32, 33, 40, W - 3, W - 2, W - 1, W, W + 1, W + 2, W + 9, 2 * W - 1, 2 * W, 2 * W + 1,
             3 * W + 5, 5000, 65535, 65536, 70001, 131071, 131072]
    stride = 131072
    buf = bytearray(stride * len(sizes))
    for i, n in enumerate(sizes):
        kind = i % 4
        blk = (datagen.records(n + 1, seed=300 + i) if kind == 0 else datagen.text_like(n + 1, seed=300 + i) if kind == 1
               else datagen.periodic(n + 1, 37 + i, seed=i) if kind == 2 else (datagen.text_like(700, seed=i) * (n // 700 + 1)))
        buf[i * stride:i * stride + n] = blk[:n]
    counts, seqs, bad = parse_on_gpu(pkg, engine, bytes(buf), level=level, sizes=sizes, stride=stride)
    for i, n in enumerate(sizes):
        blk = bytes(buf[i * stride:i * stride + n])
        got = seqs[i, :counts[i]]
        want = oracle.model_block(blk, level)
        assert bad[i] == 0, f"size {n}: verifier {bad[i]}"
        assert oracle.validate(blk, got) == 0, f"size {n}: does not replay"
        assert got.shape == want.shape and (got == want).all(), f"size {n} (level {level})"


def test_rep_parse_special_content(pkg, oracle, engine):
    """Zeros (one 128 KiB match through the uncapped extension), short periods, and two alternating offsets."""
    a, b = datagen.rand_bytes(48, 7), datagen.rand_bytes(80, 8)
    alt = (a + b[:5] + a + b) * 1100
    data = datagen.zeros(BLOCK) + datagen.periodic(BLOCK, 3) + datagen.periodic(BLOCK, 255) + alt[:BLOCK] + \
        datagen.rand_bytes(BLOCK // 2, 9) + datagen.zeros(BLOCK // 2)
    for level in (6, 12):
        check_against_model(pkg, oracle, engine, data, level=level)


def test_small_block_sizes_and_strides(pkg, oracle, engine):
    """blockSize < 128 KiB (the reference benchmark's -c 32K/64K chunks) and explicit per-block sizes."""
    data = datagen.mixed_corpus(700000, seed=31)
    for bs in (4096, 32768, 65536):
        check_against_model(pkg, oracle, engine, data[:10 * bs + 99], block_size=bs)
    sizes = [131072, 5, 70000, 0, 131072, 16, 9999]
    stride = 131072
    buf = bytearray(stride * len(sizes))
    rng = np.random.default_rng(5)
    for i, s in enumerate(sizes):
        buf[i * stride:i * stride + s] = datagen.records(s, seed=int(rng.integers(1 << 30)))[:s]
    counts, seqs, bad = parse_on_gpu(pkg, engine, bytes(buf), sizes=sizes, stride=stride)
    for i, s in enumerate(sizes):
        blk = bytes(buf[i * stride:i * stride + s])
        got = seqs[i, :counts[i]]
        want = oracle.model_block(blk, 3)
        assert bad[i] == 0
        assert got.shape == want.shape and (got == want).all(), f"sized block {i}"


def test_incompressible_block_is_one_literal_run(pkg, engine):
    """Config #5 / the reference's dataUncompressed shortcut (/root/reference/src/qatseqprod.c:1308-1313):
    a uniform-random block comes back as exactly one entry {0, srcSize, 0}."""
    data = datagen.rand_bytes(4 * BLOCK, seed=99)
    counts, seqs, bad = parse_on_gpu(pkg, engine, data)
    assert (bad == 0).all()
    for b in range(4):
        assert counts[b] == 1 and tuple(seqs[b, 0, :3]) == (0, BLOCK, 0)


def test_determinism_and_many_blocks(pkg, oracle, engine):
    """More blocks than SMs (dynamic scheduler, table reset between blocks): two runs are identical and
    a checksum over every block matches the model on a sample of blocks."""
    data = datagen.mixed_corpus(400 * BLOCK, seed=41)
    c1, s1, bad1 = parse_on_gpu(pkg, engine, data)
    c2, s2, _ = parse_on_gpu(pkg, engine, data)
    assert (bad1 == 0).all()
    assert (c1 == c2).all()
    for b in range(0, 400, 1):
        assert (s1[b, :c1[b]] == s2[b, :c2[b]]).all()
    for b in range(0, 400, 37):
        blk = data[b * BLOCK:(b + 1) * BLOCK]
        want = oracle.model_block(blk, 3)
        got = s1[b, :c1[b]]
        assert got.shape == want.shape and (got == want).all(), f"block {b}"


def test_round_trip_through_libzstd_and_ratio(pkg, oracle):
    """The reference's test.c flow with the real plugin entry points, plus the +-1 % ratio bar."""
    data = datagen.mixed_corpus(48 * BLOCK + 1000, seed=51)
    q = pkg.QatSeqProd
    assert q.startQatDevice() == pkg.QZSTD_OK
    st = q.createSeqProdState()
    try:
        for level in (1, 3, 6):
            ref = oracle.chunked_compress(data, BLOCK, level)
            r = oracle.compress_with_producer(data, q.producer, st, chunk=BLOCK, level=level, repcodes=1)
            assert r["round_trip"] and r["errors"] == 0 and r["calls"] == 49, r
            delta = r["csize"] / ref - 1
            print(f"L{level}: csize {r['csize']} vs chunked stock {ref}: {100 * delta:+.2f}%")
            assert delta <= 0.01, f"L{level}: ratio delta {delta:+.4f} outside the +1 % bar"
        # whole buffer as ONE frame: libzstd cuts it into 128 KiB blocks and calls the producer per block
        r = oracle.compress_with_producer(data, q.producer, st, chunk=len(data), level=3)
        assert r["round_trip"] and r["errors"] == 0 and r["calls"] == 49
        # look-ahead hint: one GPU batch serves all 49 callbacks
        buf = np.frombuffer(data, dtype=np.uint8)
        q.hintSource(st, buf.ctypes.data, buf.size, 0)
        before = q.getStats(st)
        r2 = oracle.compress_with_producer(buf, q.producer, st, chunk=len(data), level=3)
        after = q.getStats(st)
        q.hintSource(st, 0, 0, 0)
        assert r2["round_trip"] and r2["errors"] == 0 and r2["csize"] == r["csize"]
        assert after["batched"] - before["batched"] == 49
    finally:
        q.freeSeqProdState(st)
        q.stopQatDevice()


def test_producer_argument_rejection_on_gpu(pkg):
    """Same rejections as the reference with a live device (/root/reference/src/qatseqprod.c:1123-1137)."""
    q = pkg.QatSeqProd
    assert q.startQatDevice() == pkg.QZSTD_OK
    st = q.createSeqProdState()
    try:
        src = np.frombuffer(datagen.text_like(BLOCK, 3), dtype=np.uint8)
        out = np.zeros((43691, 4), np.uint32)
        ok = q.qatSequenceProducer(st, out.ctypes.data, 43691, src.ctypes.data, src.size, None, 0, 3, 1 << 17)
        assert ok != pkg.ZSTD_SEQUENCE_PRODUCER_ERROR and ok > 1
        E = pkg.ZSTD_SEQUENCE_PRODUCER_ERROR
        assert q.qatSequenceProducer(st, out.ctypes.data, 43691, src.ctypes.data, src.size, None, 0, 0, 1 << 17) == E
        assert q.qatSequenceProducer(st, out.ctypes.data, 43691, src.ctypes.data, src.size, None, 0, 13, 1 << 17) == E
        assert q.qatSequenceProducer(st, out.ctypes.data, 43691, src.ctypes.data, src.size, src.ctypes.data, 0, 3, 1 << 17) == E
        assert q.qatSequenceProducer(st, out.ctypes.data, 43691, src.ctypes.data, src.size, None, 8, 3, 1 << 17) == E
        assert q.qatSequenceProducer(st, out.ctypes.data, 43691, src.ctypes.data, src.size, None, 0, 3, 1 << 14) == E
        assert q.qatSequenceProducer(st, out.ctypes.data, 43691, src.ctypes.data, 1000, None, 0, 3, 1000) != E
        assert q.qatSequenceProducer(st, out.ctypes.data, 10, src.ctypes.data, src.size, None, 0, 3, 1 << 17) == E
    finally:
        q.freeSeqProdState(st)
        q.stopQatDevice()


def test_generate_sequences_hand_off(pkg, oracle):
    """QZSTD_generateSequences (the reference's flow-chart step 4 "Compress Sequences API"):
    one GPU batch -> one explicit-delimiter array -> ZSTD_compressSequences, no per-block callback.  The array
    is the concatenation of the per-block model outputs; the frame round-trips and is as small as the one the
    registered producer gives."""
    data = datagen.mixed_corpus(9 * BLOCK + 1234, seed=41)
    q = pkg.QatSeqProd
    assert q.startQatDevice() == pkg.QZSTD_OK
    st = q.createSeqProdState()
    try:
        for level in (1, 3, 6):
            seqs = q.generateSequences(st, data, level=level)
            assert seqs is not None
            want = np.concatenate([oracle.model_block(data[o:o + BLOCK], level) for o in range(0, len(data), BLOCK)])
            assert seqs.shape == want.shape and (seqs == want).all()
            r = oracle.compress_sequences(data, seqs, level=level)
            assert r["round_trip"]
            cb = oracle.compress_with_producer(data, q.producer, st, chunk=BLOCK, level=level)
            assert cb["round_trip"] and cb["errors"] == 0
            assert 0.95 < r["csize"] / cb["csize"] < 1.01     # one frame instead of one per chunk
        small = q.generateSequences(st, data[:70000], level=3, block_size=32768)
        assert small is not None and oracle.compress_sequences(data[:70000], small, level=3)["round_trip"]
        assert q.generateSequences(st, data, level=13) is None
    finally:
        q.freeSeqProdState(st)
        q.stopQatDevice()


def test_many_threads_each_with_its_own_state(pkg, oracle):
    """The reference is used from many threads, one CCtx + one producer state per thread
    (/root/reference/test/benchmark.c:222-402, README "multi-thread"); start/stop are process-wide and
    idempotent (/root/reference/src/qatseqprod.c:948-964).  Four threads compress different buffers at
    once through the registered producer: every frame round-trips, no producer error, same sizes as
    the single-threaded run."""
    import threading
    q = pkg.QatSeqProd
    bufs = [datagen.mixed_corpus(4 * BLOCK + 1000 * (i + 1), seed=50 + i) for i in range(4)]
    ref = []
    assert q.startQatDevice() == pkg.QZSTD_OK
    st0 = q.createSeqProdState()
    for b in bufs:
        ref.append(oracle.compress_with_producer(b, q.producer, st0, chunk=BLOCK, level=3)["csize"])
    q.freeSeqProdState(st0)
    out = [None] * 4

    def work(i):
        assert q.startQatDevice() == pkg.QZSTD_OK            # idempotent, from any thread
        st = q.createSeqProdState()
        try:
            for _ in range(3):
                out[i] = oracle.compress_with_producer(bufs[i], q.producer, st, chunk=BLOCK, level=3)
        finally:
            q.freeSeqProdState(st)

    ts = [threading.Thread(target=work, args=(i,)) for i in range(4)]
    for t in ts: t.start()
    for t in ts: t.join()
    q.stopQatDevice()
    for i in range(4):
        assert out[i] is not None and out[i]["round_trip"] and out[i]["errors"] == 0
        assert out[i]["csize"] == ref[i]


def test_benchmark_tool_four_threads(pkg, tmp_path):
    """tools/qzstd_benchmark (the mirror of the reference's test/benchmark.c) with -t4, plugin mode."""
    import os, subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "tools", "qzstd_benchmark")
    f = tmp_path / "in.bin"
    f.write_bytes(datagen.mixed_corpus(6 * BLOCK + 777, seed=61))
    r = subprocess.run([exe, "-t4", "-l2", "-c128K", "-E1", "-L3", "-m1", str(f)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "software fallbacks: 0" in r.stderr and r.stderr.strip().endswith("PASS"), r.stderr
    assert r.stderr.count("PASS") >= 5


def test_parse_host_many_blocks_and_pageable_input(pkg, oracle, engine):
    """The pipelined host path (b200sp_parse_host) with more blocks than the chunk plan has one-wave slots
    for (6 100 blocks of 4 KiB: chunks grow beyond one wave), a block size whose stride is padded (5 000 B ->
    2-D copies), pageable input (staging memcpy), and a ragged tail.  Every block equals the serial model."""
    data = datagen.mixed_corpus(6100 * 4096 - 1234, seed=71)
    for bs in (4096, 5000):
        part = data if bs == 4096 else data[:400 * 5000 + 77]
        counts, offsets, seqs = engine.parse_host_numpy(part, block_size=bs, level=3)
        nb = (len(part) + bs - 1) // bs
        assert counts.shape[0] == nb and int(offsets[nb]) == seqs.shape[0] == int(counts.sum())
        step = 1 if bs == 5000 else 37                      # every block of the small case, a sample of the big one
        for b in list(range(0, nb, step)) + [nb - 1]:
            got = seqs[int(offsets[b]):int(offsets[b]) + int(counts[b])]
            want = oracle.model_block(part[b * bs:(b + 1) * bs], 3)
            assert got.shape == want.shape and (got == want).all(), f"block {b} of {nb} (block size {bs})"
        total = sum(int(seqs[int(offsets[b]):int(offsets[b + 1]), 1:3].sum()) for b in (0, nb // 2, nb - 1))
        assert total == sum(len(part[b * bs:(b + 1) * bs]) for b in (0, nb // 2, nb - 1))


def test_cross_thread_coalescing(pkg, oracle):
    """SURVEY 7.3 "cross-thread coalescing": with QZSTD_setCoalescing(1) the single-block calls of many threads
    are gathered by a dispatcher and parsed in one GPU batch (b200sp_parse_blocks).  Eight threads compress
    different buffers; every frame equals the one the uncoalesced path produces, round-trips, no producer error,
    and the calls were served by the dispatcher."""
    import threading
    q = pkg.QatSeqProd
    bufs = [datagen.mixed_corpus(3 * BLOCK + 500 * (i + 1), seed=80 + i) for i in range(8)]
    assert q.startQatDevice() == pkg.QZSTD_OK
    st0 = q.createSeqProdState()
    ref = [oracle.compress_with_producer(b, q.producer, st0, chunk=BLOCK, level=3 + (i % 2) * 3)["csize"] for i, b in enumerate(bufs)]
    q.freeSeqProdState(st0)                 # (uncoalesced: single blocks and read-ahead windows)
    assert q.setCoalescing(True) is False
    out, stats = [None] * 8, [None] * 8

    def work(i):
        st = q.createSeqProdState()
        try:
            for _ in range(2):
                out[i] = oracle.compress_with_producer(bufs[i], q.producer, st, chunk=BLOCK, level=3 + (i % 2) * 3)
            stats[i] = q.getStats(st)
        finally:
            q.freeSeqProdState(st)

    ts = [threading.Thread(target=work, args=(i,)) for i in range(8)]
    for t in ts: t.start()
    for t in ts: t.join()
    assert q.setCoalescing(False) is True
    q.stopQatDevice()
    for i in range(8):
        assert out[i] is not None and out[i]["round_trip"] and out[i]["errors"] == 0
        assert out[i]["csize"] == ref[i]
        assert stats[i]["batched"] == stats[i]["calls"] > 0          # every block went through the dispatcher
