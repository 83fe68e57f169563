"""One small fused step (no graph) — run under compute-sanitizer to check the sort kernels' warp-level primitives."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from map_code_b200 import ops

g = torch.Generator().manual_seed(0)
for n, V in ((5000, 300), (40_000, 1_085_271)):
    ids = torch.randint(0, V, (n,), generator=g)
    ids[::5] = 3
    plan = ops.DedupPlan(n, V, "cuda").run(ids.cuda())
    rows = torch.randn(n, 16, device="cuda")
    G = plan.reduce_rows(rows, 16)
    torch.cuda.synchronize()
    print("ok", n, int(plan.n_unique.item()), float(G.sum()))
