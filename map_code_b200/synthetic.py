"""Synthetic Criteo- / Avazu-shaped id data.  There is no network for the real datasets; the layout mimics
the reference's preprocessing (data_preprocess/proc_criteo.py:109-161): ids 0-9 reserved (<mask> = 3), field f owns the
contiguous range [low_f, low_f + size_f), values inside a field are log-uniform (~Zipf-1) ranks."""
from __future__ import annotations

from typing import List, Tuple

import torch

CRITEO_FIELD_SIZES: List[int] = [64] * 13 + [300000, 250000, 200000, 120000, 80000, 50000, 30000, 20000, 12000, 8000, 5000, 3000,
                                             2000, 1500, 1000, 700, 500, 300, 200, 100, 60, 30, 20, 10, 5, 4]
AVAZU_FIELD_SIZES: List[int] = [7, 31, 24, 2, 7, 7, 5000, 7500, 26, 9000, 600, 36, 400000, 900000, 8500, 5, 4, 2600, 8, 9, 435, 4,
                                68, 172]
RESERVED = 10


def field_sizes(shape: str, scale_to_vocab: int = 0) -> List[int]:
    sizes = list(CRITEO_FIELD_SIZES if shape == "criteo" else AVAZU_FIELD_SIZES)
    if scale_to_vocab:
        big = [i for i, s in enumerate(sizes) if s >= 1000]
        small_total = RESERVED + sum(s for i, s in enumerate(sizes) if i not in big)
        factor = (scale_to_vocab - small_total) / float(sum(sizes[i] for i in big))
        for i in big:
            sizes[i] = max(1, int(sizes[i] * factor))
    return sizes


def vocab_size(sizes: List[int]) -> int:
    return RESERVED + sum(sizes)


def make_ids(sizes: List[int], n_rows: int, seed: int = 0, device="cpu", chunk: int = 1 << 18) -> torch.Tensor:
    """[n_rows, F] int64; rank r = floor(size^u) - 1, u ~ U[0,1) from torch.Generator().manual_seed(seed) (CPU stream, so
    every rank of a multi-GPU run and the CPU baseline see identical data)."""
    g = torch.Generator().manual_seed(seed)
    lows, acc = [], RESERVED
    for s in sizes:
        lows.append(acc)
        acc += s
    lows_t = torch.tensor(lows, dtype=torch.int64)
    sizes_f = torch.tensor(sizes, dtype=torch.float64)
    out = torch.empty(n_rows, len(sizes), dtype=torch.int64, device=device)
    for r0 in range(0, n_rows, chunk):
        n = min(chunk, n_rows - r0)
        u = torch.rand(n, len(sizes), generator=g, dtype=torch.float64)
        r = torch.floor(torch.pow(sizes_f[None, :], u)).long() - 1
        r = torch.minimum(r.clamp_(min=0), (sizes_f.long() - 1)[None, :])
        out[r0:r0 + n] = (lows_t[None, :] + r).to(device)
    return out


def feat_count(x_train: torch.Tensor, V: int) -> torch.Tensor:
    """dataset.py:49-62 semantics: occurrence count of every id in the training split (float32 [V]).  A device-resident id matrix
    is counted on the device (ops.feat_count: sort + run-length with the K2 kernels); a host matrix with torch.bincount."""
    if x_train.is_cuda:
        from . import ops
        return ops.feat_count(x_train, V)
    return torch.bincount(x_train.reshape(-1), minlength=V).float()


def field_ranges(sizes: List[int]) -> Tuple[torch.Tensor, torch.Tensor]:
    """idx_low / idx_high per field (dataset.py:64-72) for RFD_replace='Uniform'."""
    lows, acc = [], RESERVED
    for s in sizes:
        lows.append(acc)
        acc += s
    low = torch.tensor(lows, dtype=torch.int64)
    return low, low + torch.tensor(sizes, dtype=torch.int64)
