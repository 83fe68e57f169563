"""K2a alone: time map_dedup_ids_ex on the step's id streams (and the C5 one) and print the single-launch kernel's phase
timestamps.  Usage (GPU box): python scripts/bench_dedup.py"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from map_code_b200 import _lib, ops, synthetic as S  # noqa: E402


def timeit(fn, reps=50, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


def main():
    single = True
    out = {"cases": []}
    cases = []
    sizes = S.field_sizes("criteo")
    V = S.vocab_size(sizes)
    ids = S.make_ids(sizes, 4096, seed=0).cuda()
    ids[:, :3] = 3
    cases.append(("embed ids B=4096 x 39", ids.view(-1), V))
    g = torch.Generator(device="cuda").manual_seed(0)
    u = torch.rand(12288 * 26, device="cuda", generator=g)
    nce = torch.floor(torch.pow(torch.tensor(float(V), device="cuda"), u)).long().clamp_(0, V - 1)
    cases.append(("nce ids 12288 x 26 (log-uniform)", nce, V))
    cases.append(("uniform 319488 of 2^21", torch.randint(0, 1 << 21, (319488,), device="cuda"), 1 << 21))
    sizes5 = S.field_sizes("criteo", 100_000_000)
    V5 = S.vocab_size(sizes5)
    ids5 = S.make_ids(sizes5, 65536, seed=0).cuda()
    ids5[:, :3] = 3
    cases.append(("C5 embed ids B=65536 x 39", ids5.view(-1), V5))
    cases.append(("uniform 2555904 of 1e8", torch.randint(0, 100_000_000, (2555904,), device="cuda"), 100_000_000))
    for name, x, v in cases:
        plan = ops.DedupPlan(x.numel(), v, "cuda")
        us = timeit(lambda: plan.run(x))
        rec = {"case": name, "n": x.numel(), "key_bits": plan.key_bits, "us": round(us, 2), "unique": int(plan.n_unique.item())}
        if single:
            off = int(_lib.load().map_dedup_debug_offset(x.numel()))
            passes = (plan.key_bits + 7) // 8
            k = 3 + 3 * passes + 3
            st = plan.ws[off:off + 8 * k].view(torch.int64).cpu().tolist()
            d = [round((b - a) / 1e3, 2) for a, b in zip(st[:-1], st[1:])]
            names = ["H", "bar"] + sum([[f"col{p}", f"tiles{p}", f"bar{p}"] for p in range(passes)], []) + ["heads", "bar", "emit"]
            rec["phases_us"] = dict(zip(names, d))
            rec["kernel_us"] = round((st[-1] - st[0]) / 1e3, 2)
        out["cases"].append(rec)
        print(json.dumps(rec))


if __name__ == "__main__":
    main()
