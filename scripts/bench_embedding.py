"""Embedding-path micro-benchmark at config C5 scale: V ~ 1e8 rows x D = 64 (25.6 GB table, far larger than
the 126 MB L2), B = 65536 x 39 fields = 2.56 M lookups per step.  Reports the achieved algorithmic GB/s of
  K1  map_emb_gather_f32
  K2  map_dedup_ids + map_segment_reduce_rows + map_adamw_sparse_rows  (backward + fused sparse optimizer)
against the measured HBM peak.  Usage (GPU box):  python scripts/bench_embedding.py [--rows 100000000] [--dim 64] [--batch 65536]"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from map_code_b200 import ops, synthetic as S  # noqa: E402


def timeit(fn, reps=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e-3


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=100_000_000)
    ap.add_argument("--dim", type=int, default=64)
    ap.add_argument("--batch", type=int, default=65536)
    ap.add_argument("--fields", type=int, default=39)
    ap.add_argument("--dist", default="criteo", choices=["criteo", "uniform"])
    a = ap.parse_args()
    dev = "cuda"
    peak = 6445.3
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        peak = json.load(open(pk))["hbm_gbs"]
    D, n = a.dim, a.batch * a.fields
    if a.dist == "criteo":
        sizes = S.field_sizes("criteo", a.rows)
        V = S.vocab_size(sizes)
        ids = S.make_ids(sizes, a.batch, seed=0).to(dev)
        ids[:, :3] = 3  # MFP: int(39 * 0.1) = 3 masked fields per row hit the <mask> row
    else:
        V = a.rows
        ids = torch.randint(0, V, (a.batch, a.fields), device=dev)
    table = torch.empty(V, D, device=dev).normal_()
    m = torch.zeros_like(table)
    v = torch.zeros_like(table)
    out = torch.empty(n, D, device=dev)
    dY = torch.randn(n, D, device=dev)
    plan = ops.DedupPlan(n, V, dev)
    G = torch.empty(n, D, device=dev)
    hyper = torch.zeros(8, device=dev)
    ops.adamw_hyper_set(hyper, 1e-3, 0.9, 0.999, 1e-8, 1)
    res = {"V": V, "D": D, "lookups": n, "table_GB": V * D * 4 / 1e9, "peak_GBs": peak, "dist": a.dist}

    t = timeit(lambda: ops.emb_gather(table, ids, out=out.view(a.batch, a.fields, D)))
    by = n * (8 + 2 * 4 * D)
    res["K1_gather"] = {"us": t * 1e6, "GBs": by / t / 1e9, "frac_of_peak": by / t / 1e9 / peak, "alg_bytes": by}

    t_sort = timeit(lambda: plan.run(ids.view(-1)))
    U = int(plan.n_unique.item())
    t_red = timeit(lambda: plan.reduce_rows(dY, D, out=G))
    t_opt = timeit(lambda: ops.adamw_sparse_rows(table, m, v, plan, G, hyper, 5e-2))
    passes = (plan.key_bits + 7) // 8
    by_sort = n * (8 + 4 * 4 * passes)
    by_red = n * (4 + 4 * D) + U * 4 * D
    by_opt = U * (8 + 7 * 4 * D)
    res["unique_rows"] = U
    res["K2a_sort_dedup"] = {"us": t_sort * 1e6, "GBs": by_sort / t_sort / 1e9, "frac_of_peak": by_sort / t_sort / 1e9 / peak}
    res["K2b_segment_reduce"] = {"us": t_red * 1e6, "GBs": by_red / t_red / 1e9, "frac_of_peak": by_red / t_red / 1e9 / peak}
    res["K2c_sparse_adamw"] = {"us": t_opt * 1e6, "GBs": by_opt / t_opt / 1e9, "frac_of_peak": by_opt / t_opt / 1e9 / peak}
    tot = t_sort + t_red + t_opt
    res["K2_total"] = {"us": tot * 1e6, "GBs": (by_sort + by_red + by_opt) / tot / 1e9, "frac_of_peak": (by_sort + by_red + by_opt) / tot / 1e9 / peak}
    print(json.dumps(res))


if __name__ == "__main__":
    main()
