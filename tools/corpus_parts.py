"""Developer: per-category hashes of the image corpus (to see which files differ between boxes)."""
import hashlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import corpus
for cat, root, suf, budget in corpus._PLAN:
    h = hashlib.sha256(); n = 0; got = 0; names = hashlib.sha256()
    if os.path.isdir(root):
        for path in corpus._walk(root, suf):
            if got >= budget: break
            try:
                if os.path.islink(path) or not os.path.isfile(path): continue
                b = open(path, "rb").read(min(corpus._FILE_CAP, budget - got))
            except OSError:
                continue
            if not b: continue
            h.update(b); names.update(path.encode()); got += len(b); n += 1
    print(f"{cat:18s} files {n:5d} bytes {got:9d} data {h.hexdigest()[:12]} names {names.hexdigest()[:12]}  {root}")
