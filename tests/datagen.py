"""Deterministic synthetic inputs for the parity tests and the bench fallback corpus.

Nothing here is Silesia or enwik9 (neither exists in the image and there is no network); every
consumer labels these as synthetic.  All generators are seeded and vectorised with numpy.
"""
from __future__ import annotations

import numpy as np

BLOCK = 1 << 17


def rand_bytes(n: int, seed: int = 1) -> bytes:
    return np.random.default_rng(seed).integers(0, 256, n, dtype=np.uint8).tobytes()


def zeros(n: int) -> bytes:
    return bytes(n)


def periodic(n: int, period: int = 100, seed: int = 2) -> bytes:
    unit = np.random.default_rng(seed).integers(0, 256, period, dtype=np.uint8)
    reps = n // period + 1
    return np.tile(unit, reps)[:n].tobytes()


def text_like(n: int, seed: int = 3, vocab: int = 4000) -> bytes:
    """Zipf-distributed words from a random vocabulary, with punctuation and line breaks."""
    rng = np.random.default_rng(seed)
    letters = np.frombuffer(b"etaoinshrdlcumwfgypbvkjxqz", dtype=np.uint8)
    probs = np.arange(len(letters), 0, -1, dtype=np.float64)
    probs /= probs.sum()
    lens = rng.integers(2, 11, vocab)
    words = [rng.choice(letters, int(l), p=probs).tobytes() for l in lens]
    ranks = rng.zipf(1.25, size=n // 4 + 16)
    ranks = (ranks - 1) % vocab
    seps = rng.choice(np.frombuffer(b"     ,.\n", dtype=np.uint8), size=ranks.size)
    out = bytearray()
    for r, s in zip(ranks.tolist(), seps.tolist()):
        out += words[r]
        if s == 44 or s == 46:      # ", " / ". "
            out.append(s)
            out.append(32)
        else:
            out.append(s)
        if len(out) >= n:
            break
    while len(out) < n:
        out += b" padding"
    return bytes(out[:n])


def records(n: int, seed: int = 4, width: int = 96) -> bytes:
    """Fixed-width database-like records: a counter, a few slowly varying fields, a noisy field."""
    rng = np.random.default_rng(seed)
    rows = n // width + 1
    rec = np.zeros((rows, width), dtype=np.uint8)
    rec[:] = np.frombuffer((b"id=00000000;name=%-24s;grp=0000;val=000000.00;flag=N;pad=" % b"customer")[:width].ljust(width, b"."),
                           dtype=np.uint8)
    ids = np.arange(rows)
    for d in range(8):
        rec[:, 3 + 7 - d] = 48 + (ids // 10 ** d) % 10
    grp = rng.integers(0, 40, rows)
    for d in range(4):
        rec[:, 42 + 3 - d] = 48 + (grp // 10 ** d) % 10
    val = rng.integers(0, 10 ** 8, rows)
    for i, col in enumerate((51, 52, 53, 54, 55, 56, 58, 59)):
        rec[:, col] = 48 + (val // 10 ** (7 - i)) % 10
    names = rng.integers(0, 200, rows)
    for c in range(6):
        rec[:, 17 + c] = 97 + (names * (c + 3) + c * c) % 26
    rec[:, 66] = np.where(rng.random(rows) < 0.1, ord("Y"), ord("N"))
    return rec.tobytes()[:n]


def binary_like(n: int, seed: int = 5) -> bytes:
    """Executable-like: little-endian 32-bit words, mostly small values, repeated opcodes, some pointers."""
    rng = np.random.default_rng(seed)
    words = n // 4 + 1
    kind = rng.random(words)
    w = np.where(kind < 0.45, rng.integers(0, 64, words),
        np.where(kind < 0.7, rng.choice(np.array([0x48894C24, 0xE8000000, 0x0F1F4000, 0xC3909090, 0x488B4424], dtype=np.int64), words),
        np.where(kind < 0.9, 0x00400000 + rng.integers(0, 1 << 16, words) * 8, rng.integers(0, 1 << 32, words)))).astype(np.uint32)
    # repeat some 64-word "functions" to create medium-range matches
    blocks = w[: (words // 64) * 64].reshape(-1, 64)
    if blocks.shape[0] > 8:
        src = rng.integers(0, blocks.shape[0], blocks.shape[0] // 5)
        dst = rng.integers(0, blocks.shape[0], blocks.shape[0] // 5)
        blocks[dst] = blocks[src]
    return w.tobytes()[:n]


def mixed_corpus(n: int, seed: int = 6) -> bytes:
    """Silesia-like mixture (labelled synthetic): text, records, executable-like, incompressible,
    periodic and zero runs, in pieces that straddle 128 KiB block boundaries."""
    rng = np.random.default_rng(seed)
    makers = (
        (0.34, lambda m, s: text_like(m, s)),
        (0.22, lambda m, s: records(m, s)),
        (0.22, lambda m, s: binary_like(m, s)),
        (0.12, lambda m, s: rand_bytes(m, s)),
        (0.06, lambda m, s: periodic(m, int(7 + s % 300), s)),
        (0.04, lambda m, s: zeros(m)),
    )
    weights = np.array([w for w, _ in makers])
    out = bytearray()
    k = 0
    while len(out) < n:
        piece = int(rng.integers(BLOCK // 3, 3 * BLOCK))
        which = int(rng.choice(len(makers), p=weights / weights.sum()))
        out += makers[which][1](min(piece, n - len(out) + 64), seed * 1000 + k)
        k += 1
    return bytes(out[:n])
