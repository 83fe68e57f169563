"""Sweeps BLOCK_N / stages / split-K of the tcgen05 GEMM over the shapes of the DCNv2 step (B=4096) and prints a table.
Run on the GPU box:  python scripts/tune_gemm.py > gpurun_out/tune.txt"""
import itertools
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from map_code_b200 import _lib, ops  # noqa: E402

B = int(os.environ.get("TUNE_B", "4096"))
SHAPES = [  # (name, M, N, K, ta, tb, epi)
    ("fwd cross", B, 624, 624, 0, 0, 3), ("fwd mlp0", B, 1000, 624, 0, 0, 2), ("fwd mlp1", B, 1000, 1000, 0, 0, 2),
    ("fwd enc", B, 1248, 1624, 0, 0, 1),
    ("dX enc->mlp", B, 1000, 1248, 0, 1, 4), ("dX enc->cross", B, 624, 1248, 0, 1, 0), ("dX mlp", B, 1000, 1000, 0, 1, 4),
    ("dX mlp0", B, 624, 1000, 0, 1, 0), ("dX cross", B, 624, 624, 0, 1, 5),
    ("dW enc", 1248, 1624, B, 1, 1, 0), ("dW mlp", 1000, 1000, B, 1, 1, 0), ("dW mlp0", 1000, 624, B, 1, 1, 0),
    ("dW cross", 624, 624, B, 1, 1, 0),
]


def run(M, N, K, ta, tb, epi, reps=20):
    dev = "cuda"
    A = torch.randn((K, M) if ta else (M, K), device=dev)
    Bm = torch.randn((K, N) if tb else (N, K), device=dev)
    C = torch.empty(M, N, device=dev)
    bias = torch.randn(N, device=dev)
    a0, a1, ao = torch.randn(M, N, device=dev), torch.randn(M, N, device=dev), torch.empty(M, N, device=dev)
    f = lambda: ops.gemm(A, Bm, C, M, N, K, trans_a=bool(ta), trans_b=bool(tb), epilogue=epi, bias=bias, aux0=a0, aux1=a1, aux_out=ao,
                         backend="tcgen05")
    # device time per launch from a replayed CUDA graph of `reps` back-to-back launches (no host launch overhead in the number)
    for _ in range(2):
        f()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps):
            f()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3  # us


for name, M, N, K, ta, tb, epi in SHAPES:
    step = 32 if tb else 16
    res = []
    for k in ("MAP_B200_BLOCK_N", "MAP_B200_STAGES", "MAP_B200_SPLITK"):
        os.environ.pop(k, None)
    base = run(M, N, K, ta, tb, epi)
    bns = [bn for bn in (48, 64, 80, 96, 112, 128, 144, 160, 192, 208, 256) if bn % step == 0]
    splits = [1] if epi != 0 or K < 2048 else [1, 2, 3, 4, 6, 8]
    for bn, sk in itertools.product(bns, splits):
        os.environ["MAP_B200_BLOCK_N"] = str(bn)
        os.environ["MAP_B200_SPLITK"] = str(sk)
        for st in (0, 3, 5):
            if st:
                os.environ["MAP_B200_STAGES"] = str(st)
            else:
                os.environ.pop("MAP_B200_STAGES", None)
            try:
                us = run(M, N, K, ta, tb, epi, reps=10)
            except Exception as e:  # noqa: BLE001
                us = float("inf")
            res.append((us, bn, sk, st))
    res.sort()
    fl = 2.0 * M * N * K
    print(f"{name:14s} M={M} N={N} K={K} ta={ta} tb={tb} epi={epi}: default {base:7.1f} us ({fl / base / 1e6:6.1f} TF) | best " +
          " | ".join(f"bn={bn} sk={sk} st={st}: {us:6.1f} us ({fl / us / 1e6:5.1f} TF)" for us, bn, sk, st in res[:6]), flush=True)
