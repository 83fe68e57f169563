"""Per-CTA phase timeline of the tcgen05 GEMM (map_gemm_set_trace) + graph-replayed device time per shape (no host launch
overhead in the number).  Run on the GPU box:  python scripts/trace_gemm.py > gpurun_out/trace_gemm.txt"""
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
if os.environ.get("TRACE_SHAPES"):
    want = set(os.environ["TRACE_SHAPES"].split(","))
    SHAPES = [s for s in SHAPES if s[0] in want]


def make(M, N, K, ta, tb, epi):
    dev = "cuda"
    A = torch.randn((K, M) if ta else (M, K), device=dev)
    Bm = torch.randn((K, N) if tb else (N, K), device=dev)
    C = torch.empty(M, N, device=dev)
    bias = torch.randn(N, device=dev)
    a0, a1, ao = torch.randn(M, N, device=dev), torch.randn(M, N, device=dev), torch.empty(M, N, device=dev)
    return lambda: ops.gemm(A, Bm, C, M, N, K, trans_a=bool(ta), trans_b=bool(tb), epilogue=epi, bias=bias, aux0=a0, aux1=a1,
                            aux_out=ao, backend="tcgen05")


def graph_time(f, reps=20):
    for _ in range(3):
        f()
    torch.cuda.synchronize()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
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
    return e0.elapsed_time(e1) / reps * 1e3


def pct(v, q):
    v = sorted(v)
    return v[min(len(v) - 1, int(q * len(v)))]


lib = _lib.load()
CAP = 4096
buf = torch.zeros(CAP * 8, dtype=torch.int64, device="cuda")
for name, M, N, K, ta, tb, epi in SHAPES:
    f = make(M, N, K, ta, tb, epi)
    us = graph_time(f)
    fl = 2.0 * M * N * K
    buf.zero_()
    lib.map_gemm_set_trace(buf.data_ptr(), CAP)
    f()
    torch.cuda.synchronize()
    lib.map_gemm_set_trace(None, 0)
    t = buf.view(CAP, 8).cpu()
    t = t[t[:, 1] != 0]
    n = t.shape[0]
    g0 = t[:, 0]
    g1 = t[:, 7] & 0xFFFFFFFFFFFF
    smid = (t[:, 7] >> 48) & 0xFFFF
    span_ns = int(g1.max() - (g0 & 0xFFFFFFFFFFFF).min())
    start_ns = ((g0 & 0xFFFFFFFFFFFF) - (g0 & 0xFFFFFFFFFFFF).min()).tolist()
    setup = (t[:, 2] - t[:, 1]).tolist()
    first = (t[:, 3] - t[:, 2]).tolist()
    main = (t[:, 4] - t[:, 3]).tolist()
    drain = (t[:, 5] - t[:, 4]).tolist()
    epi_c = (t[:, 6] - t[:, 5]).tolist()
    total = (t[:, 6] - t[:, 1]).tolist()
    per_sm = torch.bincount(smid.to(torch.int64), minlength=148)
    print(f"{name:14s} M={M} N={N} K={K} ta={ta} tb={tb} epi={epi}: graph {us:6.1f} us ({fl / us / 1e6:6.1f} TF) | ctas={n} "
          f"sm_used={(per_sm > 0).sum().item()} max_cta_per_sm={per_sm.max().item()} | kernel span {span_ns / 1e3:6.1f} us "
          f"start p50/p99 {pct(start_ns, .5) / 1e3:5.1f}/{pct(start_ns, .99) / 1e3:5.1f} us | cycles p50 (p99): setup {pct(setup, .5)} ({pct(setup, .99)}) "
          f"first_full {pct(first, .5)} ({pct(first, .99)}) mainloop {pct(main, .5)} ({pct(main, .99)}) drain {pct(drain, .5)} ({pct(drain, .99)}) "
          f"epilogue {pct(epi_c, .5)} ({pct(epi_c, .99)}) total {pct(total, .5)} ({pct(total, .99)})", flush=True)
