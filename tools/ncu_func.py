"""Developer: stall-reason breakdown of the samples that fall into one source-line range of the kernel file.
usage: python tools/ncu_func.py rep <start marker substring> <end marker substring>"""
import csv, io, subprocess, sys
rep, m0, m1 = sys.argv[1], sys.argv[2], sys.argv[3]
both = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(both)))
hdr = rows[2]
iL, iS, iI, iA = hdr.index("Line No"), hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Address")
cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
src = open("qat-zstd-plugin_b200/csrc/lz77_kernels.cu").read().split("\n")
lo = [i for i, l in enumerate(src, 1) if m0 in l][0]
hi = [i for i, l in enumerate(src, 1) if m1 in l][0]
cur = None; seen = set(); tot = {h: 0 for _, h in cols}; ins = 0; smp = 0
byop = {}
for r in rows[3:]:
    if len(r) <= iI: continue
    if r[iL].isdigit(): cur = int(r[iL]); continue
    if cur is None or r[iA] in seen or not (lo <= cur < hi): continue
    seen.add(r[iA])
    try: ins += int(r[iI]); s = int(r[iS])
    except ValueError: continue
    smp += s
    op = r[3].split()
    op = (op[1] if op and op[0].startswith("@") else op[0]) if op else "?"
    d = byop.setdefault(op.split(".")[0], {})
    for i, h in cols:
        try: v = int(r[i])
        except ValueError: v = 0
        tot[h] += v; d[h] = d.get(h, 0) + v
print(f"lines {lo}-{hi}: executed {ins} samples {smp}")
print("  " + "  ".join(f"{h[6:]}={v}" for h, v in sorted(tot.items(), key=lambda x: -x[1]) if v))
for op, d in sorted(byop.items(), key=lambda x: -sum(x[1].values()))[:12]:
    print(f"  {op:10s} " + "  ".join(f"{h[6:]}={v}" for h, v in sorted(d.items(), key=lambda x: -x[1])[:4] if v))
