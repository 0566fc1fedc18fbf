"""Benchmark corpora.

BASELINE.json's headline workload is the Silesia corpus (211 938 580 bytes) cut into 128 KiB blocks.
Silesia is not in the image and there is no network, so the workload is resolved in this order and
the label travels with every number:

  1. ``$SILESIA``  — a file (tar / concatenation) or a directory of the 12 Silesia files: label "silesia".
  2. image corpus  — a deterministic, Silesia-sized mix of REAL files that ship in this container
     image (identical on the build box and on every GPU box): source text, C headers, shared
     objects, JSON/XML, word lists and tables, in proportions close to Silesia's text / executable /
     database / markup split (no category that differs between boxes: the image files under site-packages did).
     Label "silesia-like image corpus (real files from the container image)".
  3. synthetic     — tests/datagen.mixed_corpus, label "silesia-like synthetic".

Never report 2 or 3 as "Silesia".
"""
from __future__ import annotations

import hashlib
import os
import sys

SILESIA_BYTES = 211_938_580
BLOCK = 1 << 17

_SP = "/opt/prime-rl/.venv/lib/python3.12/site-packages"

# (category, root, suffixes, byte budget) — walked in sorted order, files truncated at 4 MiB each so
# no single file dominates; budgets sum to a little over SILESIA_BYTES and the tail is trimmed.
_PLAN = [
    ("python-source", f"{_SP}/torch", (".py",), 46_000_000),
    ("python-source", f"{_SP}/transformers", (".py",), 24_000_000),
    ("c-headers", "/usr/local/cuda/targets/x86_64-linux/include", (".h", ".hpp", ".cuh"), 24_000_000),
    ("shared-objects", "/usr/lib/x86_64-linux-gnu", (".so",), 40_000_000),
    ("shared-objects", f"{_SP}/scipy", (".so",), 22_000_000),
    ("json", f"{_SP}/mistral_common/data", (".json",), 8_000_000),
    ("json-db", "/opt/prime-rl/deps/research-environments/environments/general_agent/tasks", ("db.json",), 20_000_000),
    ("xml", f"{_SP}/cv2/data", (".xml",), 8_000_000),
    ("wordlist", "/opt/prime-rl/deps/research-environments/environments/logic_env/logic_env/games/tasks/word_sorting/scripts", (".txt",), 4_300_000),
    ("unicode-tables", "/usr/share/perl/5.38.2", (".txt", ".pl", ".pm"), 6_000_000),
    ("numeric-tables", f"{_SP}/scipy", (".npy", ".npz", ".mat", ".dat", ".csv"), 6_000_000),
    ("python-source", f"{_SP}/scipy", (".py",), 40_000_000),
]
_FILE_CAP = 4 << 20


def _walk(root, suffixes):
    for d, dirs, files in os.walk(root):
        dirs.sort()
        for f in sorted(files):
            if f.endswith(suffixes) or (".so" in suffixes and ".so." in f):
                yield os.path.join(d, f)


def image_corpus(target: int = SILESIA_BYTES):
    parts, manifest = [], []
    total = 0
    for cat, root, suf, budget in _PLAN:
        got = 0
        nfiles = 0
        if os.path.isdir(root):
            for path in _walk(root, suf):
                if got >= budget or total >= target:
                    break
                try:
                    if os.path.islink(path) or not os.path.isfile(path):
                        continue
                    with open(path, "rb") as fh:
                        b = fh.read(min(_FILE_CAP, budget - got))
                except OSError:
                    continue
                if not b:
                    continue
                parts.append(b)
                got += len(b)
                total += len(b)
                nfiles += 1
        manifest.append({"category": cat, "files": nfiles, "bytes": got})
    data = b"".join(parts)
    return data[:target], manifest


def load(target: int = SILESIA_BYTES, allow_image: bool = True):
    """Returns (data: bytes, label: str, info: dict)."""
    env = os.environ.get("SILESIA")
    if env and os.path.exists(env):
        if os.path.isdir(env):
            data = b"".join(open(os.path.join(env, f), "rb").read() for f in sorted(os.listdir(env))
                            if os.path.isfile(os.path.join(env, f)))
        else:
            data = open(env, "rb").read()
        return data, "silesia", {"source": env, "bytes": len(data)}
    if allow_image:
        cache = f"/tmp/b200sp_image_corpus_v2_{target}.bin"
        if os.path.exists(cache) and os.path.getsize(cache) >= int(0.9 * target):
            data = open(cache, "rb").read()
            return data, "silesia-like image corpus (real files from the container image)", \
                {"bytes": len(data), "sha256_16": hashlib.sha256(data).hexdigest()[:16], "cached": True}
        data, manifest = image_corpus(target)
        if len(data) >= int(0.9 * target):
            try:
                with open(cache, "wb") as fh:
                    fh.write(data)
            except OSError:
                pass
            return data, "silesia-like image corpus (real files from the container image)", \
                {"bytes": len(data), "sha256_16": hashlib.sha256(data).hexdigest()[:16], "manifest": manifest}
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    if root not in sys.path:
        sys.path.insert(0, root)
    from tests.datagen import mixed_corpus
    # generate 1/8 and tile with a per-copy byte twist so copies do not match each other trivially
    import numpy as np
    unit = np.frombuffer(mixed_corpus(target // 8 + BLOCK, seed=2024), dtype=np.uint8)
    reps = []
    for i in range(8):
        u = unit.copy()
        u[:: 4099 + i] ^= np.uint8(i + 1)
        reps.append(u)
    data = np.concatenate(reps)[:target].tobytes()
    return data, "silesia-like synthetic", {"bytes": len(data), "sha256_16": hashlib.sha256(data).hexdigest()[:16]}


if __name__ == "__main__":
    d, label, info = load()
    print(label, len(d), info)
    if len(sys.argv) > 1:
        open(sys.argv[1], "wb").write(d)
