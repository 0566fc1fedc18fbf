#!/usr/bin/env python
"""bench.py — sequence-producer GB/s (raw input) on the Silesia-sized workload at L3.

One "step" = one pass of the hot path (the sm_100a block parser) over the whole workload: every
128 KiB block of the corpus -> ZSTD_Sequence arrays.

  value      whole-job throughput with the input already resident in HBM (CUDA events, max over ranks)
  e2e        same metric through the C-ABI host call b200sp_parse_host: pinned HOST input, H2D copy,
             kernels, wire-format pack, D2H of the result, all inside the timed region
  roofline   algorithmic bytes (src + 16 B/sequence + 4 B/block) / average kernel duration vs the
             measured HBM copy peak in MEASURED_PEAKS.json
  cpu_baseline / --impl reference
             the reference's software path for this hot path — per-block ZSTD_generateSequences of
             stock libzstd 1.5.5 (what runs when the producer falls back) — on all host cores
  ratio      compressed size through stock libzstd with the plugin registered vs chunked stock L3

Launch: python bench.py [--gpus N --steps K --warmup W] ; for N > 1 under torchrun (one rank per GPU,
NCCL).  Blocks are independent, so ranks share no data-path collective: rank 0 owns the input and
broadcasts it once over NCCL (outside the timed region, reported as broadcast_ms); each rank then
parses its own replica (weak scaling: Silesia x N).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))

BLOCK = 1 << 17
METRIC = "sequence-producer GB/s (raw input) on Silesia L3; ratio delta vs ref"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--level", type=int, default=3)
    ap.add_argument("--workload", default="silesia", choices=["silesia", "random4g", "synthetic"])
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--no-ratio", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    return ap.parse_args()


def load_workload(name: str):
    import corpus
    if name == "silesia":
        return corpus.load()
    if name == "synthetic":
        return corpus.load(allow_image=False)
    import numpy as np
    data = np.random.default_rng(5).integers(0, 256, 1 << 30, dtype=np.uint8).tobytes()   # 1 GiB slice of config #5
    return data, "uniform random bytes (synthetic, 1 GiB)", {"bytes": len(data)}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "100"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, smax, reasons = [], [], set()
        for l in self.lines:
            f = [x.strip() for x in l.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); smax.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def run_reference(args, rank):
    """Reference arm: the software sequence producer of stock libzstd on all host cores."""
    if rank != 0:
        return
    import __graft_entry__ as g
    oracle = g.load_oracle()
    data, label, info = load_workload(args.workload)
    cores = os.cpu_count() or 1
    for _ in range(max(args.warmup, 0) and 1):
        oracle.cpu_bench(data[: 64 * BLOCK * cores], BLOCK, args.level, 1, cores, 1)
    t0 = time.perf_counter()
    bps, _, nseq = oracle.cpu_bench(data, BLOCK, args.level, 1, cores, args.steps)
    wall = time.perf_counter() - t0
    gbs = bps / 1e9
    line = {
        "impl": "reference", "metric": METRIC, "value": round(gbs, 4), "unit": "GB/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(1e3 * len(data) / bps, 3),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": label,
        "config": {"workload": f"{label}, {len(data)} B, {BLOCK >> 10} KiB blocks, L{args.level}", "level": args.level,
                   "block_bytes": BLOCK, "blocks": (len(data) + BLOCK - 1) // BLOCK},
        "cpu_baseline": {"value": round(gbs, 4), "unit": "GB/s", "cores": cores, "kind": "reference",
                         "sample": f"whole workload x {args.steps} passes, per-block ZSTD_generateSequences "
                                   f"(stock libzstd 1.5.5 software sequence producer), blocks partitioned over {cores} threads"},
        "e2e": {"value": round(gbs, 4), "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "sequences_per_step": nseq, "wall_s": round(wall, 2),
    }
    print(json.dumps(line), flush=True)


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return

    import numpy as np
    import torch
    import torch.distributed as dist
    import __graft_entry__ as g

    pkg = g.load_package()      # raises if libqatseqprod.so is not built: there is no fallback
    assert torch.cuda.is_available(), "bench.py needs a B200; there is no CPU path"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    # ---- input: rank 0 owns it, one NCCL broadcast hands every rank its replica
    label, info = "", {}
    if rank == 0:
        data, label, info = load_workload(args.workload)
        host = torch.frombuffer(bytearray(data), dtype=torch.uint8).pin_memory()
        n_bytes = host.numel()
    else:
        data, host, n_bytes = None, None, 0
    broadcast_ms = 0.0
    if world > 1:
        sz = torch.tensor([n_bytes], dtype=torch.int64, device=dev)
        dist.broadcast(sz, 0)
        n_bytes = int(sz.item())
    src = torch.empty(n_bytes + 64, dtype=torch.uint8, device=dev)
    if rank == 0:
        src[:n_bytes].copy_(host, non_blocking=False)
    if world > 1:
        torch.cuda.synchronize()
        dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        dist.broadcast(src, 0)
        e1.record()
        torch.cuda.synchronize()
        broadcast_ms = e0.elapsed_time(e1)
        meta = [label, info] if rank == 0 else [None, None]
        dist.broadcast_object_list(meta, 0)
        label, info = meta

    n_blocks = (n_bytes + BLOCK - 1) // BLOCK
    seqs = torch.empty((n_blocks, pkg.SEQ_STRIDE, 4), dtype=torch.int32, device=dev)
    counts = torch.zeros(n_blocks, dtype=torch.int32, device=dev)
    eng = pkg.Engine(local)
    # a non-default torch stream: the C-ABI launches on it and torch.cuda.Event times it
    tstream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(tstream)
    stream = tstream.cuda_stream
    assert stream != 0

    def step():
        eng.parse_device(src.data_ptr(), n_bytes, BLOCK, n_blocks, args.level, seqs.data_ptr(), counts.data_ptr(),
                         stream=stream)

    for _ in range(max(args.warmup, 3)):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()

    # ---- timed region: K launches, per-launch CUDA events on the launching stream
    sampler = ClockSampler(local)
    sampler.start()
    time.sleep(0.25)
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    torch.cuda.synchronize()
    evs[0].record()
    for i in range(args.steps):
        step()
        evs[i + 1].record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    total_ms = evs[0].elapsed_time(evs[-1])
    per_launch = [evs[i].elapsed_time(evs[i + 1]) for i in range(args.steps)]
    clocks = sampler.stop()
    if world > 1:
        t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())

    h_counts = counts.cpu().numpy().astype(np.int64)
    n_seq = int(h_counts.sum())
    value = world * n_bytes * args.steps / (total_ms * 1e-3) / 1e9

    # ---- end to end through the C-ABI host call (pinned host input, result brought back)
    if rank != 0:
        host = torch.empty(n_bytes, dtype=torch.uint8).pin_memory()
        host.copy_(src[:n_bytes])
    torch.cuda.synchronize()
    eng.parse_host(host.data_ptr(), n_bytes, BLOCK, args.level)       # warm-up (allocations)
    eng.parse_host(host.data_ptr(), n_bytes, BLOCK, args.level)
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(args.e2e_steps):
        nb, c_ptr, o_ptr, p_ptr = eng.parse_host(host.data_ptr(), n_bytes, BLOCK, args.level)
    e2e_s = time.perf_counter() - t0
    d2h = nb * 4 + (nb + 1) * 8 + int(o_ptr[nb]) * 8
    if world > 1:
        t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e = world * n_bytes * args.e2e_steps / e2e_s / 1e9

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant (only) kernel
    peak, peak_src = measured_peak()
    algo_bytes = n_bytes + 16 * n_seq + 4 * n_blocks
    avg_ms = sum(per_launch) / len(per_launch)
    achieved = algo_bytes / (avg_ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "achieved": round(achieved, 2), "peak": peak, "unit": "GB/s",
                "frac": round(achieved / peak, 4), "traffic": None, "peak_source": peak_src,
                "kernel": "lz77_parse_kernel", "algorithmic_bytes_per_launch": algo_bytes,
                "avg_launch_ms": round(avg_ms, 4)}
    prof = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(prof):
        try:
            roofline["traffic"] = json.load(open(prof)).get("dram_bytes_per_launch")
        except Exception:
            pass

    line = {
        "metric": METRIC, "value": round(value, 3), "unit": "GB/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": round(total_ms / args.steps, 4), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": label,
        "config": {"workload": f"{label}, {n_bytes} B per GPU, {BLOCK >> 10} KiB blocks, L{args.level}",
                   "level": args.level, "block_bytes": BLOCK, "blocks_per_gpu": n_blocks,
                   "cache": "input (212 MB) + sequence arrays exceed the 126 MB L2; no flush needed",
                   "corpus": info},
        "sequences_per_step": n_seq * world, "gpu_launches": args.steps * world,
        "e2e": {"value": round(e2e, 3), "unit": "GB/s", "h2d_bytes_per_step": n_bytes, "d2h_bytes_per_step": d2h,
                "steps": args.e2e_steps, "api": "b200sp_parse_host (pinned host input -> packed sequences on host)"},
        "roofline": roofline, "clocks": clocks,
    }
    if world > 1:
        line["broadcast_ms"] = round(broadcast_ms, 3)

    oracle = None
    if not args.no_cpu or not args.no_ratio:
        oracle = g.load_oracle()
    if world == 1 and not args.no_cpu:
        cores = os.cpu_count() or 1
        bps1, _, _ = oracle.cpu_bench(data, BLOCK, args.level, 1, cores, 1)
        iters = max(1, min(50, int(12.0 * bps1 / n_bytes)))
        bps, _, _ = oracle.cpu_bench(data, BLOCK, args.level, 1, cores, iters)
        bps_1t, _, _ = oracle.cpu_bench(data[: 256 * BLOCK], BLOCK, args.level, 1, 1, 1)
        bps_c2, csz, _ = oracle.cpu_bench(data, BLOCK, args.level, 0, cores, 1)
        line["cpu_baseline"] = {
            "value": round(bps / 1e9, 4), "unit": "GB/s", "cores": cores, "kind": "reference",
            "sample": f"whole workload x {iters} passes, per-block ZSTD_generateSequences (stock libzstd 1.5.5, the software "
                      f"path the plugin falls back to), blocks partitioned over {cores} threads",
            "one_thread_GBps": round(bps_1t / 1e9, 4), "full_compress2_all_cores_GBps": round(bps_c2 / 1e9, 4)}
    if world == 1 and not args.no_ratio:
        q = pkg.QatSeqProd
        ratio = {"E": 1, "level": args.level}
        if q.startQatDevice() == pkg.QZSTD_OK:
            st = q.createSeqProdState()
            arr = np.frombuffer(data, dtype=np.uint8)
            q.hintSource(st, arr.ctypes.data, arr.size, BLOCK)
            r = oracle.compress_with_producer(arr, q.producer, st, chunk=BLOCK, level=args.level, repcodes=1)
            stats = q.getStats(st)
            q.freeSeqProdState(st)
            q.stopQatDevice()
            ref = oracle.chunked_compress(data, BLOCK, args.level)
            ratio.update({"csize_plugin": r["csize"], "csize_ref_chunked_stock": ref,
                          "delta": round(r["csize"] / ref - 1, 5), "round_trip": r["round_trip"],
                          "fallback_blocks": r["errors"], "batched_blocks": stats["batched"]})
        line["ratio"] = ratio
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
