"""CPU (no GPU here): the C-ABI library loads, exports every symbol the headers declare, and the host
layer behaves like the reference when no device is present (no compute is attempted)."""
import ctypes
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions(header):
    txt = open(os.path.join(ROOT, "include", header)).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    txt = "\n".join(l for l in txt.split("\n") if not l.lstrip().startswith("#"))      # drop macros
    return sorted(set(re.findall(r"\b((?:QZSTD|b200sp)_[a-zA-Z][A-Za-z_0-9]*[a-z][A-Za-z_0-9]*|qatSequenceProducer)\s*\(", txt)))


def test_library_exports_every_declared_symbol(pkg):
    names = declared_functions("qatseqprod.h") + declared_functions("b200seqprod.h")
    assert len(names) >= 20
    for n in names:
        assert hasattr(pkg.lib, n), f"{n} is declared in include/ but not exported by libqatseqprod.so"
    assert sorted(set(names)) == sorted(pkg.EXPORTED_SYMBOLS)


def test_no_accidental_reference_exports(pkg):
    """The reference accidentally exports QZSTD_getSectionName, gProcess, debugLevel
    (/root/reference/src/qatseqprod.c:481,180,187); the drop-in must not."""
    out = subprocess.run(["nm", "-D", "--defined-only", pkg.LIB_PATH], capture_output=True, text=True).stdout
    for bad in ("QZSTD_getSectionName", "gProcess", "debugLevel"):
        assert bad not in out


def test_static_archive_and_names(pkg):
    """Artefact names are the drop-in contract (/root/reference/src/Makefile:85-87)."""
    d = os.path.dirname(pkg.LIB_PATH)
    assert os.path.basename(pkg.LIB_PATH) == "libqatseqprod.so"
    assert os.path.exists(os.path.join(d, "libqatseqprod.a"))


def test_version_matches_reference_header(pkg):
    assert pkg.QatSeqProd.version() == "0.2.0"
    assert (pkg.QZSTD_OK, pkg.QZSTD_STARTED, pkg.QZSTD_FAIL, pkg.QZSTD_UNSUPPORTED) == (0, 1, -1, -2)


def gpu_present():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.mark.skipif(gpu_present(), reason="no-device behaviour; covered by the gpu tests on a B200")
def test_no_device_behaviour(pkg):
    """Without a device: start -> QZSTD_FAIL, idempotent; the producer answers ERROR for every block
    (fail fast, /root/reference/src/qatseqprod.c:1140-1152) and counts them; engine creation fails loudly."""
    q = pkg.QatSeqProd
    assert q.startQatDevice() == pkg.QZSTD_FAIL
    assert q.startQatDevice() == pkg.QZSTD_FAIL
    assert pkg.lib.b200sp_driver_device_count() == 0
    assert pkg.lib.b200sp_device_count() <= 0
    with pytest.raises(pkg.B200SeqProdError):
        pkg.Engine(0)
    st = q.createSeqProdState()
    assert st
    src = np.frombuffer(b"hello world, hello world, hello world" * 100, dtype=np.uint8)
    out = np.zeros((2000, 4), np.uint32)
    E = pkg.ZSTD_SEQUENCE_PRODUCER_ERROR
    for _ in range(1001):       # crosses the 1000-block retry interval without a device appearing
        assert q.qatSequenceProducer(st, out.ctypes.data, 2000, src.ctypes.data, src.size, None, 0, 3, 1 << 17) == E
    s = q.getStats(st)
    assert s["calls"] == 1001 and s["errors"] == 1001 and s["batched"] == 0
    q.freeSeqProdState(st)
    q.freeSeqProdState(None)     # like free(NULL)
    q.stopQatDevice()


def test_argument_rejection_order_without_device(pkg):
    """dict / window / level rejections come before any device work (/root/reference/src/qatseqprod.c:1123-1137)."""
    q = pkg.QatSeqProd
    st = q.createSeqProdState()
    src = np.zeros(4096, np.uint8)
    out = np.zeros((2000, 4), np.uint32)
    E = pkg.ZSTD_SEQUENCE_PRODUCER_ERROR
    args = (st, out.ctypes.data, 2000, src.ctypes.data, src.size)
    assert q.qatSequenceProducer(*args, src.ctypes.data, 0, 3, 1 << 17) == E      # dict pointer
    assert q.qatSequenceProducer(*args, None, 1, 3, 1 << 17) == E                 # dictSize
    assert q.qatSequenceProducer(*args, None, 0, 3, 1024) == E                    # window < min(srcSize, 32K)
    assert q.qatSequenceProducer(*args, None, 0, 0, 1 << 17) == E                 # level < 1
    assert q.qatSequenceProducer(*args, None, 0, 13, 1 << 17) == E                # level > 12
    assert q.getStats(st)["errors"] == 5
    q.freeSeqProdState(st)


def test_fallback_through_libzstd_without_device(pkg, oracle):
    """The reference's test.c flow (/root/reference/test/test.c:102-136) on a box with no device: every block
    falls back to libzstd's own parser, output equals plain ZSTD_compress2, round trip holds."""
    if gpu_present():
        pytest.skip("needs a box without a device")
    from tests import datagen
    data = datagen.mixed_corpus(5 * (1 << 17) + 77, seed=2)
    q = pkg.QatSeqProd
    q.startQatDevice()
    st = q.createSeqProdState()
    r = oracle.compress_with_producer(data, q.producer, st, chunk=len(data), level=3, fallback=1)
    assert r["round_trip"] and r["errors"] == r["calls"] == 6
    r0 = oracle.compress_with_producer(data, q.producer, st, chunk=len(data), level=3, fallback=0)
    assert r0["csize"] is None                     # without fallback the compression fails, never crashes
    q.freeSeqProdState(st)
    q.stopQatDevice()


def test_tools_build_and_run(pkg, tmp_path):
    """tools/qzstd_test (mirror of the reference's test program) round-trips a file."""
    exe = os.path.join(ROOT, "tools", "qzstd_test")
    if not os.path.exists(exe):
        subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "tools")], check=True)
    from tests import datagen
    f = tmp_path / "in.bin"
    f.write_bytes(datagen.mixed_corpus(300000, seed=8))
    r = subprocess.run([exe, str(f), "3"], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "Compression and decompression were successful!" in r.stdout


def test_wire_format_expand(pkg):
    """b200sp_expand: offset | litLength << 17 | matchLength << 35 -> ZSTD_Sequence (host-only, no device)."""
    seqs = np.array([[1, 0, 3, 0], [131071, 131072, 4, 0], [65536, 17, 131071, 0], [0, 5, 0, 0]], np.uint64)
    packed = (seqs[:, 0] | (seqs[:, 1] << np.uint64(17)) | (seqs[:, 2] << np.uint64(35))).astype(np.uint64)
    out = np.zeros((4, 4), np.uint32)
    pkg.lib.b200sp_expand(packed.ctypes.data, 4, out.ctypes.data)
    assert (out == seqs.astype(np.uint32)).all()


def test_fuzz_adapter_exports_and_no_device_behaviour(pkg):
    """The five FUZZ_* hooks of /root/reference/test/fuzzing/qatseqprodfuzzer.c:41-74 live in a separate
    object (libqatseqprodfuzzer.{so,a}), not in libqatseqprod.so."""
    d = os.path.dirname(pkg.LIB_PATH)
    fz = ctypes.CDLL(os.path.join(d, "libqatseqprodfuzzer.so"))
    assert os.path.exists(os.path.join(d, "libqatseqprodfuzzer.a"))
    names = ["FUZZ_seqProdSetup", "FUZZ_seqProdTearDown", "FUZZ_createSeqProdState", "FUZZ_freeSeqProdState",
             "FUZZ_thirdPartySeqProd"]
    for n in names:
        assert hasattr(fz, n)
        assert not hasattr(pkg.lib, n), f"{n} must not be exported by libqatseqprod.so"
    fz.FUZZ_seqProdSetup.restype = ctypes.c_size_t
    fz.FUZZ_seqProdTearDown.restype = ctypes.c_size_t
    fz.FUZZ_createSeqProdState.restype = ctypes.c_void_p
    fz.FUZZ_freeSeqProdState.argtypes = [ctypes.c_void_p]
    fz.FUZZ_freeSeqProdState.restype = ctypes.c_size_t
    fz.FUZZ_thirdPartySeqProd.restype = ctypes.c_size_t
    fz.FUZZ_thirdPartySeqProd.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p,
                                          ctypes.c_size_t, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_size_t]
    if gpu_present():
        return
    assert fz.FUZZ_seqProdSetup() == ctypes.c_size_t(-1).value        # QZSTD_FAIL as size_t: the fuzzers assert == 0
    st = fz.FUZZ_createSeqProdState()
    assert st
    src = np.zeros(4096, np.uint8)
    out = np.zeros((2000, 4), np.uint32)
    assert fz.FUZZ_thirdPartySeqProd(st, out.ctypes.data, 2000, src.ctypes.data, src.size, None, 0, 3, 1 << 17) == \
        pkg.ZSTD_SEQUENCE_PRODUCER_ERROR
    assert fz.FUZZ_freeSeqProdState(st) == 0
    assert fz.FUZZ_seqProdTearDown() == 0


@pytest.mark.skipif(gpu_present(), reason="no-device behaviour; covered by the gpu tests on a B200")
def test_generate_sequences_without_device(pkg):
    q = pkg.QatSeqProd
    st = q.createSeqProdState()
    assert q.generateSequences(st, b"abcdabcdabcd" * 1000, level=3) is None
    assert q.generateSequences(st, b"abcdabcdabcd" * 1000, level=13) is None
    assert q.getStats(st)["errors"] == 2
    q.freeSeqProdState(st)


def test_compress_sequences_hand_off_format(oracle):
    """The hand-off format itself, on the CPU: per-block software sequences, each block ending with its
    {0, literals, 0} entry, concatenated, go through ZSTD_compressSequences (explicit delimiters,
    validation on) and decompress to the input."""
    from tests import datagen
    data = datagen.mixed_corpus(3 * (1 << 17) + 4321, seed=4)
    blocks = [data[o:o + (1 << 17)] for o in range(0, len(data), 1 << 17)]
    seqs = np.concatenate([oracle.sw_block(b, 3) for b in blocks])
    r = oracle.compress_sequences(data, seqs, level=3)
    assert r["round_trip"] and r["csize"] is not None
    ref = oracle.chunked_compress(data, 1 << 17, 3)
    assert 0.95 < r["csize"] / ref < 1.01      # one frame instead of one per chunk: a little smaller
    bad = seqs.copy()
    bad[5, 0] += 1                                # a wrong offset must not survive validation + round trip
    r2 = oracle.compress_sequences(data, bad, level=3)
    assert not r2["round_trip"]


@pytest.mark.skipif(gpu_present(), reason="no-device behaviour; covered by the gpu tests on a B200")
def test_coalescing_switch_without_device(pkg):
    """QZSTD_setCoalescing is a process-wide switch; without a device no dispatcher starts and the producer keeps
    answering ERROR (software fallback), and stop/start stay idempotent."""
    q = pkg.QatSeqProd
    assert q.setCoalescing(True) is False
    assert q.setCoalescing(True) is True
    assert q.startQatDevice() == pkg.QZSTD_FAIL
    st = q.createSeqProdState()
    src = np.zeros(4096, np.uint8)
    out = np.zeros((2000, 4), np.uint32)
    assert q.qatSequenceProducer(st, out.ctypes.data, 2000, src.ctypes.data, src.size, None, 0, 3, 1 << 17) == \
        pkg.ZSTD_SEQUENCE_PRODUCER_ERROR
    q.freeSeqProdState(st)
    q.stopQatDevice()
    assert q.setCoalescing(False) is True
    assert q.setCoalescing(False) is False


def test_c_abi_argument_checks_without_device(pkg):
    """Entry points of include/b200seqprod.h reject bad arguments with B200SP_EINVAL before touching a device."""
    lib = pkg.lib
    lib.b200sp_parse_blocks.restype = ctypes.c_int
    lib.b200sp_parse_host.restype = ctypes.c_int
    lib.b200sp_warmup.restype = ctypes.c_int
    EINVAL = -3
    assert lib.b200sp_parse_blocks(None, None, None, 1, 3, None) == EINVAL
    assert lib.b200sp_parse_host(None, None, 0, 131072, 3, None) == EINVAL
    assert lib.b200sp_sync(None) == EINVAL
    if not gpu_present():
        assert lib.b200sp_warmup(0) in (-1, -2)            # no device / unsupported device
        assert lib.b200sp_error_string()
