"""CPU: the oracle against itself, against stock libzstd, and against the committed golden vectors."""
import os

import numpy as np
import pytest

from tests import datagen

BLOCK = 1 << 17
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
KINDS = [(datagen.text_like, 1), (datagen.records, 2), (datagen.binary_like, 3), (datagen.rand_bytes, 4)]


@pytest.mark.parametrize("maker,seed", KINDS)
@pytest.mark.parametrize("level", [1, 3, 6])
def test_model_sequences_replay_to_input(oracle, maker, seed, level):
    data = maker(BLOCK, seed)
    seqs = oracle.model_block(data, level)
    assert oracle.validate(data, seqs) == 0
    assert seqs[-1, 0] == 0 and seqs[-1, 2] == 0                       # last entry: trailing literals
    assert (seqs[:-1, 2] >= 4).all() and (seqs[:-1, 0] >= 1).all()     # minMatch 4, real offsets
    assert int(seqs[:, 1].sum() + seqs[:, 2].sum()) == len(data)


@pytest.mark.parametrize("n", [0, 1, 3, 7, 8, 9, 100, 4095, 65536, 131072])
def test_model_edge_sizes(oracle, n):
    data = datagen.records(max(n, 1), 9)[:n]
    seqs = oracle.model_block(data, 3)
    assert oracle.validate(data, seqs) == 0 if n else len(seqs) == 1


def test_model_special_inputs(oracle):
    z = oracle.model_block(datagen.zeros(BLOCK), 3)
    assert len(z) == 2 and tuple(z[0, :3]) == (1, 1, BLOCK - 1)         # one literal, one run-length match
    r = oracle.model_block(datagen.rand_bytes(BLOCK, 5), 3)
    assert len(r) == 1 and tuple(r[0, :3]) == (0, BLOCK, 0)             # incompressible: one literal run
    p = oracle.model_block(datagen.periodic(BLOCK, 100), 3)
    assert len(p) == 2 and p[0, 0] == 100


def test_software_producer_matches_reference_convention(oracle):
    """Per-block ZSTD_generateSequences: ends in {0, lits, 0}, replays, sums to srcSize."""
    data = datagen.text_like(BLOCK, 7)
    s = oracle.sw_block(data, 3)
    assert oracle.validate(data, s) == 0
    assert s[-1, 0] == 0 and s[-1, 2] == 0


@pytest.mark.parametrize("level", [1, 3, 6, 12])
def test_software_producer_through_the_slot_equals_chunked_stock(oracle, level):
    """The software sequence producer fed back through ZSTD_registerSequenceProducer reproduces chunked
    stock compression to within 0.3 % (BASELINE.md probe row), with a lossless round trip."""
    data = datagen.mixed_corpus(12 * BLOCK, seed=3)
    import ctypes
    st = oracle.lib.oracle_sw_create()
    try:
        r = oracle.compress_with_producer(data, oracle.sw_producer_ptr(), st, level=level, repcodes=1)
    finally:
        oracle.lib.oracle_sw_free(st)
    ref = oracle.chunked_compress(data, BLOCK, level)
    assert r["round_trip"] and r["errors"] == 0 and r["calls"] == 12
    assert abs(r["csize"] / ref - 1) < 0.003


def test_model_ratio_close_to_stock_L3(oracle):
    """The serial model of the GPU parser, through stock libzstd, against chunked stock L3 (+-1 % bar)."""
    import ctypes
    data = datagen.mixed_corpus(24 * BLOCK, seed=11)

    @ctypes.CFUNCTYPE(ctypes.c_size_t, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p,
                      ctypes.c_size_t, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_size_t)
    def producer(state, out, cap, src, size, d, ds, level, window):
        prm = oracle.model_params(level)
        return oracle.lib.seqmodel_block(src, size, out, cap, ctypes.byref(prm))

    r = oracle.compress_with_producer(data, ctypes.cast(producer, ctypes.c_void_p), None, level=3, repcodes=1)
    ref = oracle.chunked_compress(data, BLOCK, 3)
    assert r["round_trip"] and r["errors"] == 0
    delta = r["csize"] / ref - 1
    assert delta < 0.01, f"model is {100 * delta:+.2f}% vs chunked stock"


def test_validator_catches_every_defect(oracle):
    data = b"ABCDEFGHIJKLMNOP" * 64
    n = len(data)
    good = np.array([[16, 16, n - 16, 0], [0, 0, 0, 0]], np.uint32)
    assert oracle.validate(data, good) == 0
    cases = {
        -2: [[16, 16, 2, 0], [16, 0, n - 18, 0], [0, 0, 0, 0]],
        -3: [[0, 16, n - 16, 0], [0, 0, 0, 0]],
        -4: [[32, 16, n - 16, 0], [0, 0, 0, 0]],
        -5: [[17, 17, n - 17, 0], [0, 0, 0, 0]],
        -6: [[16, 16, n - 32, 0], [0, 0, 0, 0]],
        -7: [[16, 16, 100, 0], [0, 16, 0, 0], [16, 0, n - 132, 0], [0, 0, 0, 0]],
    }
    for code, seqs in cases.items():
        assert oracle.validate(data, np.array(seqs, np.uint32)) == code


# ---- LZ4s: restatement of QZSTD_decLz4s (/root/reference/src/qatseqprod.c:1013-1091) -------------
def test_lz4s_handwritten_streams(oracle):
    # token 0x52: 5 literals, match code 2 (-> length 4), offset 0x0010; then final token 0x30 with 3 literals
    s = bytes([0x52]) + b"abcde" + bytes([0x10, 0x00]) + bytes([0x30]) + b"xyz"
    assert oracle.declz4s(s).tolist() == [[16, 5, 4, 0], [0, 3, 0, 0]]
    # literal-only token (match code 0) folds its literals into the next sequence (:1077-1084)
    s = bytes([0x20]) + b"ab" + bytes([0x05, 0x00]) + bytes([0x11]) + b"c" + bytes([0x07, 0x00]) + bytes([0x00])
    assert oracle.declz4s(s).tolist() == [[7, 3, 3, 0], [0, 0, 0, 0]]
    # 15-escapes on both nibbles: 15+255+4 = 274 literals, match code 15+10 = 25 -> length 27
    s = bytes([0xFF, 255, 4]) + bytes(274) + bytes([0x34, 0x12, 10]) + bytes([0x00])
    assert oracle.declz4s(s).tolist() == [[0x1234, 274, 27, 0], [0, 0, 0, 0]]
    # a stream ending right after a match leaves the loop with ip == endip: like the reference it is
    # NOT an error and the count includes one more (untouched) entry (:1086-1090)
    assert len(oracle.declz4s(bytes([0x11]) + b"a" + bytes([0x01, 0x00]))) == 2
    # capacity guard idx >= cap-1 (:1073-1076)
    s = (bytes([0x01, 0x01, 0x00]) * 5) + bytes([0x00])
    assert oracle.declz4s(s, capacity=6) is None
    assert len(oracle.declz4s(s, capacity=7)) == 6


def test_lz4s_encode_decode_round_trip_on_model_output(oracle):
    """oracle_enclz4s o oracle_declz4s is the identity on sequences with offsets < 64 KiB and
    matchLength < 64 KiB + 2 (the format's limits, which the QAT engine respects)."""
    data = datagen.mixed_corpus(BLOCK // 2, seed=21)[:65000]
    seqs = oracle.model_block(data, 3)
    stream = oracle.enclz4s(seqs)
    assert stream is not None
    back = oracle.declz4s(stream)
    assert back is not None and (back[:, :3] == seqs[:, :3]).all()


def test_golden_vectors(oracle):
    """Committed fixtures (tests/golden/make_golden.py): inputs, the model's sequences, LZ4s streams and
    their decoded sequences.  Any change of the oracle's arithmetic shows up here."""
    idx = os.path.join(GOLDEN, "index.txt")
    names = [l.strip() for l in open(idx) if l.strip()]
    assert len(names) >= 4
    for name in names:
        data = open(os.path.join(GOLDEN, name + ".in"), "rb").read()
        want = np.fromfile(os.path.join(GOLDEN, name + ".seq"), dtype=np.uint32).reshape(-1, 4)
        got = oracle.model_block(data, 3)
        assert got.shape == want.shape and (got == want).all(), name
        lz = open(os.path.join(GOLDEN, name + ".lz4s"), "rb").read()
        dec = oracle.declz4s(lz)
        assert dec is not None and (dec[:, :3] == want[:, :3]).all(), name


# ---- the lane-level statement of the kernel's parse stage (oracle/lanemodel.c) against the serial model ----
LANE_KINDS = KINDS + [(lambda n, s: datagen.zeros(n), 0), (lambda n, s: datagen.periodic(n, 100), 0),
                      (datagen.mixed_corpus, 5)]


@pytest.mark.parametrize("level", [1, 3, 6, 12])
@pytest.mark.parametrize("maker,seed", LANE_KINDS)
def test_lane_model_equals_serial_model(oracle, maker, seed, level):
    data = maker(BLOCK, seed)
    want = oracle.model_block(data, level)
    got = oracle.lane_model_block(data, level)
    assert got.shape == want.shape and (got == want).all()


@pytest.mark.parametrize("n", [0, 1, 7, 8, 9, 31, 32, 33, 1023, 1024, 1025, 4097, 70001, 131071])
def test_lane_model_ragged_sizes(oracle, n):
    for maker, seed in ((datagen.text_like, 11), (datagen.records, 12)):
        data = maker(max(n, 1), seed)[:n]
        want = oracle.model_block(data, 3)
        got = oracle.lane_model_block(data, 3)
        assert got.shape == want.shape and (got == want).all()


# ---- property test: random structured inputs (hypothesis) ----
try:
    from hypothesis import given, settings, strategies as st
    HAVE_HYPOTHESIS = True
except Exception:                                              # pragma: no cover
    HAVE_HYPOTHESIS = False


if HAVE_HYPOTHESIS:
    _piece = st.one_of(
        st.binary(min_size=1, max_size=40),                                          # literals
        st.tuples(st.binary(min_size=1, max_size=12), st.integers(2, 60)).map(lambda t: t[0] * t[1]),   # periodic runs
        st.integers(1, 600).map(lambda n: b"\x00" * n),                               # zero runs
    )

    @settings(max_examples=60, deadline=None)
    @given(pieces=st.lists(_piece, min_size=1, max_size=60), repeat=st.integers(1, 3), level=st.sampled_from([1, 3, 6]))
    def test_lane_model_property(oracle, pieces, repeat, level):
        """Arbitrary mixtures of literals, periodic runs and zero runs, repeated so that far matches exist: the model's
        sequences replay to the input and the lane-level formulation equals the serial one."""
        data = (b"".join(pieces) * repeat)[:BLOCK]
        want = oracle.model_block(data, level)
        assert oracle.validate(data, want) == 0
        got = oracle.lane_model_block(data, level)
        assert got.shape == want.shape and (got == want).all()
