"""Helpers shared by the -m gpu tests: run a batch through the C-ABI with torch-owned device memory."""
from __future__ import annotations

import numpy as np
import torch


def parse_on_gpu(pkg, engine, data: bytes, block_size: int = 1 << 17, level: int = 3, sizes=None, stride=None):
    """Returns (counts[nBlocks], seqs[nBlocks, SEQ_STRIDE, 4] u32 numpy, bad[nBlocks] from the on-device verifier)."""
    dev = torch.device("cuda:0")
    if sizes is None:
        n_blocks = (len(data) + block_size - 1) // block_size
        stride_ = block_size if stride is None else stride
    else:
        n_blocks = len(sizes)
        stride_ = stride
    assert stride_ % 16 == 0
    buf = bytearray(data) + bytearray(32)
    src = torch.frombuffer(buf, dtype=torch.uint8).to(dev)
    seqs = torch.full((max(n_blocks, 1), pkg.SEQ_STRIDE, 4), -1, dtype=torch.int32, device=dev)
    counts = torch.zeros(max(n_blocks, 1), dtype=torch.int32, device=dev)
    bad = torch.full((max(n_blocks, 1),), 99, dtype=torch.int32, device=dev)
    d_sizes = 0
    keep = None
    if sizes is not None:
        keep = torch.tensor(list(sizes), dtype=torch.int32, device=dev)
        d_sizes = keep.data_ptr()
    torch.cuda.synchronize()
    engine.parse_device(src.data_ptr(), len(data), block_size, n_blocks, level, seqs.data_ptr(), counts.data_ptr(),
                        stride=stride_, d_sizes=d_sizes)
    engine.verify_device(src.data_ptr(), len(data), block_size, n_blocks, seqs.data_ptr(), counts.data_ptr(),
                         bad.data_ptr(), stride=stride_, d_sizes=d_sizes)
    engine.sync()
    return (counts.cpu().numpy()[:n_blocks].astype(np.int64), seqs.cpu().numpy().view(np.uint32)[:n_blocks],
            bad.cpu().numpy()[:n_blocks])


def check_against_model(pkg, oracle, engine, data: bytes, block_size: int = 1 << 17, level: int = 3):
    counts, seqs, bad = parse_on_gpu(pkg, engine, data, block_size, level)
    n_blocks = counts.shape[0]
    total = 0
    for b in range(n_blocks):
        blk = data[b * block_size:(b + 1) * block_size]
        got = seqs[b, :counts[b]]
        want = oracle.model_block(blk, level)
        assert oracle.validate(blk, got) == 0, f"block {b}: GPU sequences do not replay to the input"
        assert bad[b] == 0, f"block {b}: on-device verifier reports {bad[b]}"
        if got.shape != want.shape or not (got == want).all():
            k = 0
            while k < min(len(got), len(want)) and (got[k] == want[k]).all():
                k += 1
            raise AssertionError(f"block {b} (level {level}): GPU differs from the serial model at sequence {k}: "
                                 f"gpu={got[k] if k < len(got) else None} model={want[k] if k < len(want) else None} "
                                 f"(counts {len(got)} vs {len(want)})")
        total += int(counts[b])
    return total
