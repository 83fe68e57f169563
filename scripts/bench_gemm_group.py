"""Persistent grouped GEMM (map_gemm_tf32_group) vs one launch per problem, on the GEMM levels of the DCNv2 step (graph-replayed
device time) + the per-CTA wait breakdown of the persistent kernel.   python scripts/bench_gemm_group.py > gpurun_out/gemm_group.txt"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from map_code_b200 import _lib, ops  # noqa: E402

B = int(os.environ.get("TUNE_B", "4096"))
# (M, N, K, ta, tb, epi)
LEVELS = {
    "F1 cross0+mlp0": [(B, 624, 624, 0, 0, 3), (B, 1000, 624, 0, 0, 2)],
    "F2 cross+mlp": [(B, 624, 624, 0, 0, 3), (B, 1000, 1000, 0, 0, 2)],
    "F4 enc": [(B, 1248, 1624, 0, 0, 1)],
    "B1 head dX x2 + dW enc": [(B, 624, 1248, 0, 1, 7), (B, 1000, 1248, 0, 1, 4), (1248, 1624, B, 1, 1, 0)],
    "B2 dX+dW cross, dX+dW mlp": [(B, 624, 624, 0, 1, 7), (B, 1000, 1000, 0, 1, 4), (624, 624, B, 1, 1, 0), (1000, 1000, B, 1, 1, 0)],
    "B4 dX mlp0 + dW cross0 + dW mlp0": [(B, 624, 1000, 0, 1, 0), (624, 624, B, 1, 1, 0), (1000, 624, B, 1, 1, 0)],
    "B5 dE": [(B, 624, 624, 0, 1, 8)],
}
if os.environ.get("LEVELS"):
    LEVELS = {k: v for k, v in LEVELS.items() if k.split()[0] in os.environ["LEVELS"].split(",")}


def make(M, N, K, ta, tb, epi):
    dev = "cuda"
    A = torch.randn((K, M) if ta else (M, K), device=dev)
    Bm = torch.randn((K, N) if tb else (N, K), device=dev)
    Cm = torch.empty(M, N, device=dev)
    bias = torch.randn(N, device=dev)
    a0, a1, a2 = (torch.randn(M, N, device=dev) for _ in range(3))
    ao, acc = torch.empty(M, N, device=dev), torch.empty(M, N, device=dev)
    return dict(A=A, B=Bm, C_out=Cm, M=M, N=N, K=K, trans_a=bool(ta), trans_b=bool(tb), epilogue=epi, bias=bias, aux0=a0, aux1=a1, aux_out=ao,
                aux2=a2, acc_out=acc if epi == 7 else None)


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
tot_g = tot_s = 0.0
for name, shapes in LEVELS.items():
    probs = [make(*s) for s in shapes]
    fl = sum(2.0 * s[0] * s[1] * s[2] for s in shapes)
    t_group = graph_time(lambda: ops.gemm_group(probs, backend="tcgen05"))
    t_single = graph_time(lambda: [ops.gemm(backend="tcgen05", **p) for p in probs])
    tot_g += t_group
    tot_s += t_single
    buf.zero_()
    lib.map_gemm_set_trace(buf.data_ptr(), CAP)
    ops.gemm_group(probs, backend="tcgen05")
    torch.cuda.synchronize()
    lib.map_gemm_set_trace(None, 0)
    t = buf.view(CAP, 8).cpu()
    t = t[t[:, 1] != 0]
    extra = ""
    if t.shape[0]:
        span = (int((t[:, 7] & 0xFFFFFFFFFFFF).max()) - int((t[:, 0] & 0xFFFFFFFFFFFF).min())) / 1e3
        lead = t[t[:, 3] != 0]   # CTAs that issued MMAs (pair mode: the leaders)
        ph = lambda a, b, tt=t: pct((tt[:, a] - tt[:, b]).tolist(), .5)
        extra = (f" | ctas={t.shape[0]} kernel span {span:5.1f} us | cycles p50: setup {ph(2, 1)} first_full {ph(3, 2, lead)} "
                 f"mainloop {ph(4, 3, lead)} drain {ph(5, 4, lead)} epilogue {ph(6, 5)} total {ph(6, 1)}")
    print(f"{name:34s}: group {t_group:6.1f} us ({fl / t_group / 1e6:6.1f} TF) | per-problem launches {t_single:6.1f} us ({fl / t_single / 1e6:6.1f} TF)"
          + extra, flush=True)
print(f"sum over the levels listed: group {tot_g:.1f} us, per-problem launches {tot_s:.1f} us")
