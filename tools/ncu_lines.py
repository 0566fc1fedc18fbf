"""Developer: per-source-line executed warp instructions / stall samples / avg active threads from an .ncu-rep.
usage: python tools/ncu_lines.py rep [top=40]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
both = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(both)))
hdr = rows[2]
iL, iS, iI, iA, iT = hdr.index("Line No"), hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Address"), hdr.index("Thread Instructions Executed")
per = {}; cur = None; seen = set(); text = {}
for r in rows[3:]:
    if len(r) <= iI: continue
    if r[iL].isdigit():
        cur = int(r[iL]); text[cur] = r[1]; continue
    if cur is None or r[iA] in seen: continue
    seen.add(r[iA])
    try: ins, smp, thr = int(r[iI]), int(r[iS]), int(r[iT])
    except ValueError: continue
    a = per.setdefault(cur, [0, 0, 0]); a[0] += ins; a[1] += smp; a[2] += thr
ti = sum(v[0] for v in per.values()); ts = sum(v[1] for v in per.values())
print(f"total warp instr {ti}  samples {ts}")
for ln, v in sorted(per.items(), key=lambda x: -x[1][0])[:top]:
    print(f"{ln:5d} inst {100*v[0]/ti:5.1f}%  smp {100*v[1]/ts:5.1f}%  act {v[2]/max(v[0],1):5.1f}  {text.get(ln,'').strip()[:110]}")
