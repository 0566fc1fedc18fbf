"""Opcode histogram of the shipped kernels (cuobjdump -sass on the built library), for profiles/.
usage: python tools/sass_histogram.py [lib] > profiles/rN_sass_opcodes.txt"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "qat-zstd-plugin_b200", "libqatseqprod.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
fn, hist = None, collections.OrderedDict()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        fn = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        hist[fn] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", line)
    if m and fn:
        hist[fn][m.group(1)] += 1
print(f"SASS opcode histogram of {os.path.relpath(lib, ROOT)} (cuobjdump -sass, sm_100a)\n")
for fn, h in hist.items():
    tot = sum(h.values())
    fam = collections.Counter()
    for op, c in h.items():
        fam[op.split(".")[0]] += c
    print(f"== {fn}: {tot} instructions")
    print("   families: " + ", ".join(f"{k} {v}" for k, v in fam.most_common(28)))
    key = [op for op in h if op.split(".")[0] in ("UBLKCP", "SYNCS", "MATCH", "REDUX", "ATOMS", "ATOMG", "RED", "LDS", "LDG", "STG", "STS", "VOTE", "SHFL", "BAR", "UTMALDG", "HMMA", "UTCHMMA", "CREATEPOLICY")]
    print("   memory / sync / warp-collective opcodes: " + ", ".join(f"{op} {h[op]}" for op in sorted(key)))
    print()
