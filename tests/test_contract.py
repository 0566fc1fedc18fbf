"""CPU: the producer-output contract stock libzstd 1.5.5 enforces, replayed with
ctypes producers.  These 13 cases are the known-answer set for the drop-in boundary: they say what
qatSequenceProducer may and may not return.  Case 11 is why every parity test replays the sequences
itself: libzstd does not check that matches are true."""
import ctypes

import numpy as np
import pytest

N = 1024
DATA = b"ABCDEFGHIJKLMNOP" * 64


def make_producer(pkg, seqs, ret=None):
    def fn(state, out, cap, src, size, dict_, dsize, level, window):
        for i, (off, lit, ml) in enumerate(seqs):
            out[i].offset, out[i].litLength, out[i].matchLength, out[i].rep = off, lit, ml, 0
        return len(seqs) if ret is None else ret
    return pkg.PRODUCER_F(fn)


def run(pkg, seqs, ret=None, validate=1, fallback=0):
    z = pkg.ZstdLib()
    cb = make_producer(pkg, seqs, ret)
    comp = z.compress_with_producer(DATA, 3, ctypes.cast(cb, ctypes.c_void_p), None, validate=validate,
                                    fallback=fallback)
    return z, comp


OK_CASES = {
    0: [(16, 16, N - 16), (0, 0, 0)],
    1: [(16, 16, N - 16)],                                  # libzstd appends the delimiter itself
    2: [(16, 16, N - 32), (0, 16, 0)],                      # reference style: trailing literals in the last entry
    6: [(16, 16, 3), (16, 0, N - 19), (0, 0, 0)],           # matchLength 3, litLength 0
    12: [(0, N, 0)],                                        # all literals
}


@pytest.mark.parametrize("case", sorted(OK_CASES))
def test_valid_outputs_round_trip(pkg, case):
    z, comp = run(pkg, OK_CASES[case])
    assert z.decompress(comp, N) == DATA


@pytest.mark.parametrize("case,seqs,ret", [
    (3, [(16, 16, N - 32)], None),                          # sum < srcSize
    (4, [(16, 16, N), (0, 0, 0)], None),                    # sum > srcSize
    (8, [], 0),                                             # count 0
    (9, [(0, N, 0)], 10 ** 6),                              # count > capacity
    (10, [(16, 16, 100), (0, 16, 0), (16, 0, N - 132), (0, 0, 0)], None),   # delimiter in the middle
])
def test_invalid_outputs_are_rejected(pkg, case, seqs, ret):
    with pytest.raises(RuntimeError):
        run(pkg, seqs, ret)


def test_invalid_outputs_fall_back_silently_when_enabled(pkg, oracle):
    """With ZSTD_c_enableSeqProducerFallback=1 an invalid answer becomes a software parse of that block:
    output size equals the no-producer size, which is why benchmarks must count producer errors."""
    z, comp = run(pkg, [], ret=pkg.ZSTD_SEQUENCE_PRODUCER_ERROR, fallback=1)
    assert z.decompress(comp, N) == DATA
    assert len(comp) == oracle.chunked_compress(DATA, N, 3)


def test_offset_before_block_start_corrupts(pkg):
    """Case 5: libzstd accepts an offset reaching before the block; decompression fails."""
    z, comp = run(pkg, [(32, 16, N - 16)])
    with pytest.raises(RuntimeError):
        z.decompress(comp, N)


def test_match_length_two(pkg):
    """Case 7: matchLength 2 is rejected only when ZSTD_c_validateSequences is on."""
    with pytest.raises(RuntimeError):
        run(pkg, [(16, 16, 2), (16, 0, N - 18), (0, 0, 0)], validate=1)


def test_false_match_is_silent_corruption(pkg, oracle):
    """Case 11: an in-range offset that is not a true match compresses and decompresses 'successfully'
    to the WRONG bytes — only our own validator sees it."""
    seqs = [(17, 17, N - 17), (0, 0, 0)]
    z, comp = run(pkg, seqs)
    back = z.decompress(comp, N)
    assert len(back) == N and back != DATA
    assert oracle.validate(DATA, np.array([s + (0,) for s in seqs], np.uint32)) == -5
