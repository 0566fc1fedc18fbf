"""Developer: aggregate host<->device copy bandwidth of N ranks copying at the same time (pinned memory, both
directions at once), the ceiling the end-to-end path of bench.py lives under.
usage: python -m torch.distributed.run --nproc-per-node N tools/pcie_ceiling.py"""
import os, json, time
import torch, torch.distributed as dist
rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local); dev = torch.device("cuda", local)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=dev)
n_in, n_out = 1700 << 20, 1400 << 20           # what one rank moves per bench step: 1.7 GB in, 1.4 GB out
h_in = torch.empty(n_in, dtype=torch.uint8).pin_memory(); h_out = torch.empty(n_out, dtype=torch.uint8).pin_memory()
d_in = torch.empty(n_in, dtype=torch.uint8, device=dev); d_out = torch.empty(n_out, dtype=torch.uint8, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def step():
    with torch.cuda.stream(s1): d_in.copy_(h_in, non_blocking=True)
    with torch.cuda.stream(s2): h_out.copy_(d_out, non_blocking=True)
for _ in range(2): step()
torch.cuda.synchronize()
if world > 1: dist.barrier()
t0 = time.perf_counter()
K = 6
for _ in range(K): step()
torch.cuda.synchronize()
if world > 1: dist.barrier()
dt = time.perf_counter() - t0
if world > 1:
    t = torch.tensor([dt], dtype=torch.float64, device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX); dt = float(t.item())
if rank == 0:
    print(json.dumps({"ranks": world, "h2d_GBps_total": round(world * K * n_in / dt / 1e9, 1), "d2h_GBps_total": round(world * K * n_out / dt / 1e9, 1),
                      "per_rank_h2d": round(K * n_in / dt / 1e9, 1), "per_rank_d2h": round(K * n_out / dt / 1e9, 1)}), flush=True)
if world > 1: dist.destroy_process_group()
