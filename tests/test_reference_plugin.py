"""The reference plugin ITSELF — unmodified /root/reference/src/qatseqprod.c compiled against the fake QAT
driver (oracle/refstub) into oracle/_ref/libqzstd_ref.so — run here as the behavioural oracle for the
plugin layer: lifecycle return codes, argument rejections, output convention, QZSTD_decLz4s arithmetic.
CPU tests compare the oracle's restatements with it; the gpu test compares OUR library with it case by case.
Skipped when the .so is absent (it can only be built where /root/reference is mounted)."""
import ctypes
import json
import os
import subprocess
import sys

import numpy as np
import pytest

from tests import datagen

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libqzstd_ref.so")
needs_ref = pytest.mark.skipif(not os.path.exists(REF_SO), reason="oracle/_ref/libqzstd_ref.so not built (make -C oracle ref)")


def scenario(which, name, **env):
    e = dict(os.environ)
    e.update({k: str(v) for k, v in env.items()})
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "ref_scenarios.py"), which, name],
                       capture_output=True, text=True, env=e, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    return json.loads(r.stdout.strip().splitlines()[-1])


@needs_ref
def test_declz4s_restatement_matches_reference(oracle):
    """oracle_declz4s == the reference's QZSTD_decLz4s on the golden streams and on fresh ones."""
    lib = ctypes.CDLL(REF_SO)
    lib.ref_decLz4s.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_uint]
    lib.ref_decLz4s.restype = ctypes.c_size_t
    streams = []
    gold = os.path.join(ROOT, "tests", "golden")
    for name in open(os.path.join(gold, "index.txt")).read().split():
        streams.append(open(os.path.join(gold, name + ".lz4s"), "rb").read())
    for seed in range(6):
        blk = datagen.mixed_corpus(60000, seed=100 + seed)
        streams.append(oracle.enclz4s(oracle.model_block(blk, 3)))
    streams.append(bytes([0x00]))                                   # empty block: one literals-only token
    streams.append(bytes([0xF0, 255, 255, 0]) + bytes(15 + 510))    # long literal run only
    for s in streams:
        a = np.frombuffer(s, dtype=np.uint8).copy()
        out = np.zeros((43691, 4), np.uint32)
        n = lib.ref_decLz4s(out.ctypes.data, 43691, a.ctypes.data, a.size)
        mine = oracle.declz4s(s)
        assert n != ctypes.c_size_t(-1).value and mine is not None
        assert n == len(mine) and (out[:n, :3] == mine[:, :3]).all()
    # capacity guard, identical on both sides
    s = (bytes([0x01, 0x01, 0x00]) * 5) + bytes([0x00])
    a = np.frombuffer(s, dtype=np.uint8).copy()
    out = np.zeros((16, 4), np.uint32)
    assert lib.ref_decLz4s(out.ctypes.data, 6, a.ctypes.data, a.size) == ctypes.c_size_t(-1).value
    assert oracle.declz4s(s, capacity=6) is None


@needs_ref
def test_reference_lifecycle_with_fake_device():
    r = scenario("ref", "sequences")
    assert r["version"] == "0.2.0"
    assert r["start"] == 0 and r["start_again"] == 0 and r["start_after_stop"] == 0     # QZSTD_OK, idempotent
    assert r["all_valid_and_last_entry_is_literals"] and r["total_sequences"] > 100


@needs_ref
def test_reference_round_trip_with_fake_device():
    r = scenario("ref", "roundtrip")
    assert r["round_trip"] and r["errors"] == 0 and r["calls"] == 10


@needs_ref
def test_reference_without_hardware_matches_ours_without_gpu():
    """No QAT hardware vs no GPU: same FAIL status, same ERROR for every block incl. across the
    1000-block retry interval (/root/reference/src/qatseqprod.c:1140-1152)."""
    ref = scenario("ref", "down", FAKEQAT_DEVICES=0)
    assert ref["start"] == -1 and ref["errors_in_1001_calls"] == 1001
    try:
        import torch
        if torch.cuda.is_available():
            return
    except Exception:
        pass
    ours = scenario("ours", "down")
    assert ours["start"] == ref["start"] and ours["errors_in_1001_calls"] == ref["errors_in_1001_calls"]
    assert ours["version"] == ref["version"]


@needs_ref
def test_reference_capability_missing_is_started():
    """Driver up, LZ4s capability missing -> QZSTD_STARTED (1), producer answers ERROR (:958-959, :1140)."""
    r = scenario("ref", "down", FAKEQAT_NO_LZ4S=1)
    assert r["start"] == 1 and r["errors_in_1001_calls"] == 1001


@needs_ref
def test_reference_uncompressible_shortcut():
    """dataUncompressed -> exactly one entry {0, srcSize, 0} (:1308-1313); our kernel emits the same for random data."""
    r = scenario("ref", "uncompressible", FAKEQAT_INCOMPRESSIBLE=1)
    assert r["count"] == 1 and r["first"] == [0, 131072, 0]


@needs_ref
@pytest.mark.gpu
def test_our_plugin_matches_reference_case_by_case():
    """Same scenarios through OUR libqatseqprod.so on a B200 and through the reference on its fake device."""
    for name in ("rejections", "sequences", "roundtrip", "uncompressible"):
        env = {"FAKEQAT_INCOMPRESSIBLE": 1} if name == "uncompressible" else {}
        ref, ours = scenario("ref", name, **env), scenario("ours", name)
        for k in ("version", "start", "start_again", "start_after_stop"):
            assert ours[k] == ref[k], (name, k, ours[k], ref[k])
        if name == "rejections":
            assert ours["is_error"] == ref["is_error"], (ours["is_error"], ref["is_error"])
        elif name == "sequences":
            assert ours["all_valid_and_last_entry_is_literals"] and ref["all_valid_and_last_entry_is_literals"]
        elif name == "roundtrip":
            assert ours["round_trip"] and ref["round_trip"] and ours["errors"] == ref["errors"] == 0
            assert ours["calls"] == ref["calls"]
        else:
            assert ours["count"] == ref["count"] == 1 and ours["first"] == ref["first"]
