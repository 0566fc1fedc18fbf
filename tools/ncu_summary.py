"""Summarises an .ncu-rep of the parse kernel into a text file for profiles/ (run where ncu is installed).

usage: python tools/ncu_summary.py gpurun_out/prof_r1.ncu-rep profiles/r1_ncu_summary.txt
"""
import bisect
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "qat-zstd-plugin_b200", "csrc", "lz77_kernels.cu")
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "lts__t_bytes.sum", "sm__cycles_elapsed.max"]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    m = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
    lines = [f"ncu --set full --clock-control none summary of {os.path.basename(rep)} (kernel {m.get('Kernel Name', ('?',))[0]})", ""]
    for k in KEYS:
        if k in m:
            lines.append(f"{k:70s} {m[k][0]} {m[k][1]}")
    lines.append("")
    lines.append("warp stall reasons (warps stalled per issue-active cycle):")
    for h in hdr:
        if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio"):
            lines.append(f"  {h[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')]:24s} {float(m[h][0]):6.2f}")
    # per-function share from the source page
    both = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(both)))
    hdr2 = rows[2]
    iL, iS, iI, iA = hdr2.index("Line No"), hdr2.index("# Samples"), hdr2.index("Instructions Executed"), hdr2.index("Address")
    src = open(SRC).read().split("\n")
    marks = sorted((i, l.split("(")[0].split()[-1]) for i, l in enumerate(src, 1)
                   if (l.startswith("__device__") or l.startswith("__global__")) and "(" in l)
    # every CUDA line is followed by the SASS rows it produced; count each SASS address once and
    # attribute it to the function whose definition precedes that line
    agg, ti, ts, cur, seen = {}, 0, 0, None, set()
    for r in rows[3:]:
        if len(r) <= iI:
            continue
        if r[iL].isdigit():
            cur = int(r[iL])
            continue
        if cur is None or r[iA] in seen:
            continue
        seen.add(r[iA])
        try:
            ins, smp = int(r[iI]), int(r[iS])
        except ValueError:
            continue
        k = bisect.bisect_right([x[0] for x in marks], cur) - 1
        name = marks[k][1] if k >= 0 else "helpers"
        a = agg.setdefault(name, [0, 0])
        a[0] += ins
        a[1] += smp
        ti += ins
        ts += smp
    lines.append("")
    lines.append("share of executed warp instructions / of stall samples by source function (lineinfo):")
    for k, v in sorted(agg.items(), key=lambda x: -x[1][0]):
        lines.append(f"  {k:22s} inst {100 * v[0] / max(ti, 1):5.1f}%   samples {100 * v[1] / max(ts, 1):5.1f}%")
    open(out, "w").write("\n".join(lines) + "\n")
    try:
        rd = float(m["dram__bytes_read.sum"][0]) * {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1}[m["dram__bytes_read.sum"][1]]
        wr = float(m["dram__bytes_write.sum"][0]) * {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1}[m["dram__bytes_write.sum"][1]]
        json.dump({"dram_bytes_per_launch": int(rd + wr), "dram_bytes_read": int(rd), "dram_bytes_write": int(wr),
                   "source": os.path.basename(rep), "note": "one ncu --set full capture of lz77_parse_kernel on the bench workload"},
                  open(os.path.join(os.path.dirname(out), "traffic.json"), "w"), indent=1)
    except Exception as e:
        print("traffic.json not written:", e)
    print("\n".join(lines))


if __name__ == "__main__":
    main()
