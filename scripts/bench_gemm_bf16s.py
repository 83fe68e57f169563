"""Split-bf16 GEMM (map_gemm_bf16s_group) on the GEMM levels of the DCNv2 step: graph-replayed device time per level, tensor-pipe
work issued, and the per-CTA timeline of one launch (where the time of a tile goes: waiting for the accumulator buffer, for the
first stage, issuing MMAs, draining, epilogue).      python scripts/bench_gemm_bf16s.py > gpurun_out/gemm_bf16s_levels.txt"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from map_code_b200 import _lib, ops  # noqa: E402

B = int(os.environ.get("TUNE_B", "4096"))
# (M, N, K, ta, tb, epi, terms, planes of C)
LEVELS = {
    "F1 cross0+mlp0": [(B, 624, 624, 0, 0, 3, 3, 3), (B, 1000, 624, 0, 0, 2, 6, 3)],
    "F2 cross+mlp": [(B, 624, 624, 0, 0, 3, 3, 3), (B, 1000, 1000, 0, 0, 2, 6, 3)],
    "F4 enc": [(B, 1248, 1624, 0, 0, 1, 3, 0)],
    "B1 head dX x2 + dW enc": [(B, 624, 1248, 0, 1, 7, 3, 2), (B, 1000, 1248, 0, 1, 4, 3, 2), (1248, 1624, B, 1, 1, 0, 3, 0)],
    "B2 dX+dW cross, dX+dW mlp": [(B, 624, 624, 0, 1, 7, 3, 2), (B, 1000, 1000, 0, 1, 4, 3, 2), (624, 624, B, 1, 1, 0, 3, 0), (1000, 1000, B, 1, 1, 0, 3, 0)],
    "B4 dX mlp0 + dW cross0 + dW mlp0": [(B, 624, 1000, 0, 1, 0, 3, 0), (624, 624, B, 1, 1, 0, 3, 0), (1000, 624, B, 1, 1, 0, 3, 0)],
    "B5 dE": [(B, 624, 624, 0, 1, 8, 3, 0)],
    "S1 mlp fwd alone (terms 6)": [(B, 1000, 1000, 0, 0, 2, 6, 3)],
    "S2 mlp fwd alone (terms 3)": [(B, 1000, 1000, 0, 0, 2, 3, 3)],
    "S3 plain 4096x1024x1024 (terms 3, no epilogue operands)": [(B, 1024, 1024, 0, 0, 0, 3, 0)],
    "E1 relumask dgrad, no colsum, no planes": [(B, 1000, 1000, 0, 1, 4, 3, 0, False)],
    "E2 relumask dgrad, colsum, no planes": [(B, 1000, 1000, 0, 1, 4, 3, 0, True)],
    "E3 relumask dgrad, colsum, 2 planes": [(B, 1000, 1000, 0, 1, 4, 3, 2, True)],
    "E4 cross fwd, 3 planes": [(B, 624, 624, 0, 0, 3, 3, 3, False)],
    "E5 cross fwd, no planes": [(B, 624, 624, 0, 0, 3, 3, 0, False)],
    "E6 cross_bwd dgrad, colsum, 2 planes": [(B, 624, 624, 0, 1, 7, 3, 2, True)],
    "E7 cross_bwd dgrad, no colsum, no planes": [(B, 624, 624, 0, 1, 7, 3, 0, False)],
    "S4 plain 8192x2048x2048 (terms 3)": [(8192, 2048, 2048, 0, 0, 0, 3, 0)],
}
if os.environ.get("LEVELS"):
    LEVELS = {k: v for k, v in LEVELS.items() if k.split()[0] in os.environ["LEVELS"].split(",")}


def make(M, N, K, ta, tb, epi, terms, pc, colsum=None):
    dev = "cuda"
    if colsum is None:
        colsum = epi in (4, 7)
    A = torch.randn((K, M) if ta else (M, K), device=dev)
    Bm = torch.randn((K, N) if tb else (N, K), device=dev)
    Cm = torch.empty(M, N, device=dev)
    bias = torch.randn(N, device=dev)
    a0, a1, a2 = (torch.randn(M, N, device=dev) for _ in range(3))
    ao, acc = torch.empty(M, N, device=dev), torch.empty(M, N, device=dev)
    Ap = ops.split_planes(A, ops.alloc_planes(A.shape[0], A.shape[1], 3, dev))
    Bp = ops.split_planes(Bm, ops.alloc_planes(Bm.shape[0], Bm.shape[1], 3, dev))
    Cp = ops.alloc_planes(M, N, pc, dev) if pc else None
    return dict(A=A, B=Bm, C_out=Cm, M=M, N=N, K=K, trans_a=bool(ta), trans_b=bool(tb), epilogue=epi, bias=bias, aux0=a0, aux1=a1, aux_out=ao,
                aux2=a2, acc_out=acc if epi == 7 else None, Ap=Ap, Bp=Bp, Cp=Cp, terms=terms,
                colsum_out=torch.zeros(N, device=dev) if colsum else None)


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


def med(v):
    v = sorted(v)
    return v[len(v) // 2] if v else 0


lib = _lib.load()
WORDS = 64
CAP = 148 * WORDS
buf = torch.zeros(CAP, dtype=torch.int64, device="cuda")
tot = 0.0
for name, shapes in LEVELS.items():
    probs = [make(*s) for s in shapes]
    fl = sum(2.0 * s[0] * s[1] * s[2] for s in shapes)
    issued = sum(2.0 * s[0] * s[1] * s[2] * s[6] for s in shapes)
    t_group = graph_time(lambda: ops.gemm_group(probs, backend="bf16s"))
    if name[0] in "FB":
        tot += t_group
    buf.zero_()
    lib.map_gemm_bf16s_set_trace(buf.data_ptr(), CAP)
    ops.gemm_group(probs, backend="bf16s")
    torch.cuda.synchronize()
    lib.map_gemm_bf16s_set_trace(None, 0)
    t = buf.view(-1, WORDS).cpu()
    t = t[t[:, 0] != 0]
    print(f"{name:52s}: {t_group:7.1f} us  algorithmic {fl / t_group / 1e6:6.1f} TF/s  bf16 MMA work issued {issued / t_group / 1e6:6.1f} TF/s", flush=True)
    if t.shape[0]:
        span = (int(t[:, 2].max()) - int(t[:, 0].min())) / 1e3
        lead = t[0::2]   # even CTAs of the grid = cluster leaders (they issue the MMAs)
        ntile = (lead[:, 3] >> 32).tolist()
        rows = []
        for c in range(lead.shape[0]):
            for j in range(min(int(ntile[c]), 10)):
                r = lead[c, 4 + 6 * j: 10 + 6 * j].tolist()
                prev_end = lead[c, 7 + 6 * (j - 1)].item() if j > 0 else lead[c, 1].item()
                rows.append(dict(j=j, wait_acc=r[1] - prev_end, wait_stage=r[2] - r[1], mma_issue=r[3] - r[2], drain=r[4] - r[3], epi=r[5] - r[4],
                                 tma_first_vs_setup=r[0] - lead[c, 1].item()))
        cta_total = [(int(lead[c, 2]) - int(lead[c, 0])) / 1e3 for c in range(lead.shape[0])]
        print(f"    ctas={t.shape[0]} kernel span {span:6.1f} us; per-CTA lifetime us min/med/max {min(cta_total):.1f}/{med(cta_total):.1f}/{max(cta_total):.1f}; "
              f"tiles per cluster min/max {min(ntile)}/{max(ntile)}")
        for j in sorted(set(r["j"] for r in rows)):
            rr = [r for r in rows if r["j"] == j]
            print(f"    tile #{j}: n={len(rr):3d}  cycles (median)  wait_acc {med([r['wait_acc'] for r in rr]):7d}  wait_first_stage {med([r['wait_stage'] for r in rr]):7d}  "
                  f"mma_issue {med([r['mma_issue'] for r in rr]):7d}  drain_to_acc_ready {med([r['drain'] for r in rr]):7d}  epilogue {med([r['epi'] for r in rr]):7d}")
print(f"sum over the F*/B* levels: {tot:.1f} us")
