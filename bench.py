#!/usr/bin/env python
"""bench.py — MFP / RFD pretrain samples/sec (DCNv2, Criteo shape) on N B200s of one node.

    python bench.py --gpus 1 --steps 200 --warmup 20
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...
    python bench.py --impl reference ...      # the reference's CPU path (oracle port) on this box's host cores

One "step" = one pretraining step (dynamic_mask -> forward -> backward -> AdamW -> LR schedule) on one batch of synthetic
Criteo-shaped ids.  Workload at N=1: BASELINE.json configs[1] — DCNv2 MFP, 39 fields, embed 16, hidden 1000x3, cross 3,
proj 32, K=25 negatives, mask ratio 0.1, batch 4096 per GPU (weak scaling).  Prints ONE JSON line on rank 0.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

F_CRITEO, D, H, NC, NH, P, K, MASK_RATIO, B_PER_GPU = 39, 16, 1000, 3, 3, 32, 25, 0.1, 4096
N_TRAIN = 1 << 20


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--task", default="MFP", choices=["MFP", "RFD"])
    ap.add_argument("--optimizer-mode", default="sparse", choices=["sparse", "dense_exact"])
    ap.add_argument("--batch", type=int, default=None, help="per-GPU batch (default: 4096; 65536 for --workload c5)")
    ap.add_argument("--workload", default="c2", choices=["c2", "c4", "c5"],
                    help="c2: DCNv2 Criteo shape (BASELINE configs[1]/[2], the headline); c4: DeepFM on the Avazu shape (configs[3]); "
                         "c5: DCNv2, ~1e8-row vocabulary x dim 64, batch 65536 (configs[4])")
    ap.add_argument("--mask-ratio", type=float, default=0.1, help="0.1 = the headline configs (L = int(F * ratio) = 3 at 39 fields); "
                                                                  "0.3 (L = 11) is what the reference's own run scripts use (SURVEY §8d secondary point)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-secondary", action="store_true", help="skip the side measurements (GPU-eager reference, RFD / dense_exact / mask 0.3 / TF32)")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--profile-steps", type=int, default=5)
    ap.add_argument("--timeline", default=None, help="re-capture the step with a timestamp marker after every call and write the "
                                                     "per-stream timeline of one graph replay here (single GPU)")
    ap.add_argument("--dump-profile", default=None, help="write the per-(kernel, shape) timing table of the eager profiling pass here")
    return ap.parse_args()


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return dict(hbm=d["hbm_gbs"], tensor=d["bf16_tflops"], tensor_sustained=d.get("bf16_tflops_sustained", d["bf16_tflops"]), src="measured")
    return dict(hbm=6650.0, tensor=1590.0, tensor_sustained=1400.0, src="fallback")


WORKLOAD = "c2"


def workload_name(task, batch, n):
    if WORKLOAD == "c4":
        return (f"DeepFM {task} pretraining, synthetic Avazu shape (24 fields, V=1334055), embed 16, hidden 1000x3, proj 32, K=25, "
                f"mask_ratio {MASK_RATIO:g} (L={int(24 * MASK_RATIO)}), batch {batch}/GPU x {n} GPU")
    if WORKLOAD == "c5":
        return (f"DCNv2 {task} pretraining, scaled Criteo shape (39 fields, V=99999992), embed 64, hidden 1000x3, cross 3, proj 32, K=25, "
                f"mask_ratio {MASK_RATIO:g} (L={int(39 * MASK_RATIO)}), batch {batch}/GPU x {n} GPU")
    return (f"DCNv2 {task} pretraining, synthetic Criteo shape (39 fields, V=1085271), embed 16, hidden 1000x3, cross 3, proj 32, "
            f"K=25, mask_ratio {MASK_RATIO:g} (L={int(39 * MASK_RATIO)}), batch {batch}/GPU x {n} GPU")


def config_dict(task):
    from map_code_b200 import synthetic as S
    if WORKLOAD == "c4":
        sizes, name, d = S.field_sizes("avazu"), "DeepFM", D
    elif WORKLOAD == "c5":
        sizes, name, d = S.field_sizes("criteo", 100_000_000), "DCNv2", 64
    else:
        sizes, name, d = S.field_sizes("criteo"), "DCNv2", D
    V = S.vocab_size(sizes)
    return sizes, V, dict(model_name=name, embed_size=d, hidden_size=H, num_hidden_layers=NH, num_cross_layers=NC, hidden_act="relu",
                          hidden_dropout_rate=0.0, embed_dropout_rate=0.0, embed_norm=False, layer_norm_eps=1e-12, pt_neg_num=K,
                          proj_size=P, input_size=V, num_fields=len(sizes), pretrain=True, pt_type=task, RFD_replace="Unigram")


# --------------------------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    """Samples SM clock and throttle reasons DURING the timed region through NVML (10 ms period)."""

    def __init__(self, index=0):
        self.samples, self.stop, self.index = [], threading.Event(), index
        self.t = threading.Thread(target=self.run, daemon=True)
        self.max_mhz = None

    def run(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            while not self.stop.is_set():
                self.samples.append((nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM), nv.nvmlDeviceGetCurrentClocksEventReasons(h),
                                     nv.nvmlDeviceGetPowerUsage(h) / 1000.0))
                self.stop.wait(0.01)
        except Exception as e:  # noqa: BLE001
            self.err = repr(e)

    def __enter__(self):
        self.t.start()
        return self

    def __exit__(self, *a):
        self.stop.set()
        self.t.join(timeout=6)

    def summary(self):
        sm = sorted(s[0] for s in self.samples)
        bits = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4,
                "hw_power_brake_slowdown": 0x80}
        reasons = sorted(n for n, b in bits.items() if any(s[1] & b for s in self.samples))
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": self.max_mhz, "reasons": reasons, "samples": len(self.samples),
                "power_w_max": max((s[2] for s in self.samples), default=None)}


# --------------------------------------------------------------------------------------------------------------- CPU arm
def build_oracle(task, batch, n_train=N_TRAIN):
    """The reference's CPU path restated (oracle/map_oracle.py): dense gradients + dense AdamW over every table.
    The dense reference algorithm cannot hold the C5 tables (3 x 25.6 GB + dense gradients): the CPU arm of c5 is timed at
    the largest configuration the reference handles, the c2 vocabulary and width (stated in `sample`)."""
    global WORKLOAD
    from map_code_b200 import synthetic as S
    from oracle import map_oracle as O
    saved, WORKLOAD = WORKLOAD, ("c2" if WORKLOAD == "c5" else WORKLOAD)
    sizes, V, cfgd = config_dict(task)
    WORKLOAD = saved
    X = S.make_ids(sizes, n_train, seed=0)
    fc = S.feat_count(X, V)
    cfg = O.OracleConfig(**cfgd)
    params = O.init_params(cfg, fc, seed=1)
    ap = aa = None
    if task == "MFP":
        renormed, _, _ = O.nce_noise_distribution(fc)
        ap, aa = O.alias_build(renormed)
    tr = O.OracleTrainer(cfg, params, alias_prob=ap, alias_alias=aa, x_train=X, lr=1e-3, weight_decay=5e-2, mask_ratio=MASK_RATIO,
                         sampling_method="randint", seed=42, lr_lambda=O.cosine_schedule_lambda(0, 100000))
    return tr, X


def run_cpu(task, batch, steps, warmup, budget_s=None):
    torch.set_num_threads(os.cpu_count() or 1)
    tr, X = build_oracle(task, batch)
    times = []
    t_start = time.perf_counter()
    i = 0
    while True:
        xb = X[(i * batch) % (X.shape[0] - batch):][:batch].contiguous()
        t0 = time.perf_counter()
        tr.step(xb)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
        i += 1
        if len(times) >= steps or (budget_s is not None and time.perf_counter() - t_start > budget_s and len(times) >= 3):
            break
    tot = sum(times)
    return dict(value=batch * len(times) / tot, ms_per_step=1e3 * tot / len(times), steps=len(times), cores=torch.get_num_threads())


def main_reference(args):
    """--impl reference: the reference's own CPU implementation of the path (oracle port; the reference is pure Python/torch
    and is restated 1:1 in oracle/map_oracle.py, pinned to it by tests/golden) with all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    r = run_cpu(args.task, args.batch, args.steps, max(args.warmup, 1), budget_s=240.0)  # bounded: stops after ~4 min of steps
    line = {"impl": "reference", "metric": f"{args.task} pretrain samples/sec (DCNv2, Criteo shape)", "value": r["value"], "unit": "samples/s",
            "n_gpus": args.gpus, "steps": r["steps"], "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": shared_config(args.task, args.batch, max(args.gpus, 1)),
            "impl_detail": {"device": "host CPU", "optimizer": "dense transformers-AdamW over every table row (the reference)"},
            "cpu_baseline": {"value": r["value"], "unit": "samples/s", "cores": r["cores"], "kind": "port",
                             "sample": f"{r['steps']} full steps at batch {args.batch} (dense grads + dense AdamW like the reference)"},
            "e2e": {"value": r["value"], "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------------------------- GPU arm
def algorithmic(tag):
    """(bytes, flops) of one launch from its shape tag — the per-unit figures of SURVEY.md §8(d) / DESIGN.md."""
    if tag is None:
        return 0.0, 0.0
    kind = tag[0]
    if kind == "gemm":
        _, M, N, Kd = tag[:4]
        return 4.0 * (M * Kd + N * Kd + M * N), 2.0 * M * N * Kd
    if kind == "gemm_group":   # one grouped launch: the sum over its problems (M, N, K[, bf16 products per fp32 product])
        return (sum(4.0 * (t[0] * t[2] + t[1] * t[2] + t[0] * t[1]) for t in tag[1:]), sum(2.0 * t[0] * t[1] * t[2] for t in tag[1:]))
    if kind == "gather":
        _, n, d = tag
        return n * (8 + 2 * 4 * d), 0.0
    if kind == "nce":
        _, N, Kn, Pn = tag
        return N * (Kn + 1) * (8 + 4 * Pn + 4 + 4) + N * 4 * Pn, 2.0 * N * (Kn + 1) * Pn * 2
    if kind == "segred":
        _, n, d = tag
        return n * (4 + 4 * d) + n * 4 * d, 0.0
    if kind == "sparse_adamw":
        _, n, d = tag
        return n * (8 + 7 * 4 * d), 0.0
    if kind in ("field_enc_fwd", "field_enc_dgrad", "field_enc_wgrad"):   # by-field encoder: N positions x P outputs x K inputs
        _, N, Pn, Kd = tag   # gathered rows of `final` (or the per-position gradient rows) dominate the traffic
        return 4.0 * N * (Kd + Pn), 2.0 * N * Pn * Kd
    if kind == "head_bwd_fold":   # L partial rows read, ~4 fp32 rows + operand planes written per sample
        _, Bn, Ln, nc = tag
        return 4.0 * Bn * nc * (Ln + 5), 0.0
    if kind == "dedup":   # ids read once + (key, value) read and written per 8-bit radix pass
        _, n, bits = tag
        return n * (8 + 4 * 4 * ((bits + 7) // 8)), 0.0
    return 0.0, 0.0


def measured_traffic(kernel_class: str, workload: str, task: str):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel class, from the committed `ncu --set full`
    capture of this workload (profiles/traffic.json, written by scripts/ncu_summary.py); None when no capture matches."""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    if not os.path.exists(path):
        return None
    try:
        for e in json.load(open(path)):
            if e["kernel_class"] == kernel_class and e["workload"] == workload and e["task"] == task:
                return e["avg_dram_bytes_per_launch"]
    except (ValueError, KeyError):
        pass
    return None


def replay_kernel_class(records, reps=20):
    """The launches of ONE kernel class of the step (same arguments, same order), back to back on one stream inside a CUDA
    graph: device time per launch without the event / host-launch overhead of the eager profiling pass and without the
    other streams' kernels competing for the SMs.  Returns (seconds per pass over the class, launches per pass)."""
    lib = _lib_mod().load()
    calls = [(getattr(lib, n), a) for n, a, _ in records]
    s = torch.cuda.Stream()
    st = s.cuda_stream

    def body():
        for fn, a in calls:
            fn(*a[:-1], st)
    with torch.cuda.stream(s):
        body()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=s):
        body()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e-3 / reps, len(calls)


DTYPE_NOTE = {"bf16s": "f32 (split-bf16 tensor-core products: fp32 operands as 2-3 bf16 planes, fp32 accumulate; fp32-level accuracy)",
              "tf32": "f32 (tf32 tensor-core multiplies, fp32 accumulate)", "simt": "f32 (CUDA cores)"}


def gemm_backend_name():
    from map_code_b200 import ops
    return ops.gemm_backend()


def _lib_mod():
    from map_code_b200 import _lib
    return _lib


if os.environ.get("MAP_B200_BENCH_VERBOSE"):   # where is the host when a run stalls?  (stack of every thread after 45 s)
    import faulthandler
    faulthandler.dump_traceback_later(45, exit=False, file=sys.stderr)


def _stage(msg):
    if os.environ.get("MAP_B200_BENCH_VERBOSE"):
        print(f"[bench rank {os.environ.get('RANK', '0')}] {msg}", file=sys.stderr, flush=True)


def write_timeline(eng, batch, path):
    """Timeline of the captured multi-stream step: the step is re-captured with a one-thread %globaltimer marker after every
    C-ABI call on that call's stream; one replay then yields the end time of every call inside the real graph schedule
    (markers add ~1-2 us per call, so the step is a little slower than the timed one)."""
    from map_code_b200 import _lib
    buf = torch.zeros(4096, dtype=torch.int64, device=eng.dev)
    t_begin = torch.zeros(1, dtype=torch.int64, device=eng.dev)
    _lib.TIMELINE = dict(buf=buf, ops=[])
    eng.graph = None
    eng.use_graph = True
    orig_body = eng._step_body

    def body():
        _lib.load().map_timestamp_ns(t_begin.data_ptr(), torch.cuda.current_stream().cuda_stream)
        _lib.TIMELINE["ops"].clear()
        orig_body()

    eng._step_body = body
    for i in range(4):
        eng.step(batch(i))
    torch.cuda.synchronize()
    ops_, _lib.TIMELINE = list(_lib.TIMELINE["ops"]), None
    eng._step_body = orig_body
    eng.graph = None
    t0 = int(t_begin.item())
    ends = buf[:len(ops_)].cpu().tolist()
    streams = {}
    last_end = {}
    rows = []
    for (name, tag, st), e in zip(ops_, ends):
        sid = streams.setdefault(st, len(streams))
        rows.append((e - t0, sid, name, tag, last_end.get(sid)))
        last_end[sid] = e - t0
    with open(path, "w") as f:
        f.write(f"# step span {max(r[0] for r in rows) / 1e3:.1f} us; columns: end_us stream since_prev_on_stream_us call tag\n")
        for end, sid, name, tag, prev in sorted(rows):
            d = "" if prev is None else f"{(end - prev) / 1e3:8.1f}"
            f.write(f"{end / 1e3:9.1f} s{sid} {d:>8s} {name:30s} {tag if tag else ''}\n")


def _shutdown(world):
    """Multi-rank exit: ranks leave together and skip the NCCL communicator teardown (destroy_process_group after CUDA-graph
    captured collectives was seen to hang at N=2); the result line is already flushed."""
    if world > 1:
        import torch.distributed as dist
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)



L2_NOTE = ("inputs larger than L2: tables + optimizer state (0.9 GB at c2/c4, >100 GB at c5) touched at random, a new batch every step; "
           "no explicit flush")


def shared_config(task, batch, world):
    """the WORKLOAD description both arms print (what is computed, not how): identical in the reference line"""
    return {"workload": workload_name(task, batch, world), "global_batch": batch * world, "l2": L2_NOTE}


class _DS:
    def __init__(self, X):
        self.X = X

    def __len__(self):
        return self.X.shape[0]


def build_run(task, optimizer_mode, mask_ratio, batch, world, rank, dev, data=None, single_gpu_global=False, graph=True):
    """model + Trainer + fused step for one configuration.  data = (sizes, V, X_train, feat_count) is shared between runs.
    single_gpu_global: a single-GPU step over the GLOBAL batch (the parity reference of a multi-GPU run)."""
    from types import SimpleNamespace
    from map_code_b200 import synthetic as S
    from map_code_b200.arguments import Config, TrainingArguments
    from map_code_b200.models import BaseModel
    from map_code_b200.trainer import Trainer
    sizes, V, cfgd = config_dict(task)
    Bg = batch * world
    if data is None:
        n_train = N_TRAIN if WORKLOAD != "c5" else max(1 << 18, 4 * Bg)
        X_train = S.make_ids(sizes, n_train, seed=0)
        data = (sizes, V, X_train, S.feat_count(X_train, V), X_train.to(dev))
    _, _, X_train, fc, X_dev = data
    cfgd.update(feat_count=fc, data_dir=None, seed=42, table_grad_mode="sparse")
    torch.manual_seed(1)
    with torch.device(dev):  # parameters are created on the GPU (the C5 tables do not fit comfortably in host memory)
        model = BaseModel.from_config(Config.from_dict(cfgd))
    model.to(dev)
    b_eng = Bg if single_gpu_global else batch
    targs = TrainingArguments(per_gpu_train_batch_size=b_eng, learning_rate=1e-3, weight_decay=5e-2, lr_sched="cosine",
                              sampling_method="randint", mask_ratio=mask_ratio, pretrain=True, pt_type=task, seed=42,
                              optimizer_mode=optimizer_mode)
    trainer = Trainer(model, model.config, targs, _DS(X_dev), _DS(X_dev))
    total_steps = 100000
    if world > 1 and not single_gpu_global:
        from map_code_b200 import dist as mdist
        eng = mdist.make_sharded_step(trainer, total_steps, 0, world, rank)
    else:
        eng = trainer.fused_step(total_steps, 0)
    eng.use_graph = graph  # NCCL collectives of the sharded step are captured in the graph as well
    n_batches = X_train.shape[0] // Bg

    def batch_fn(i):  # rank-local slice of global batch i (device resident)
        r0 = (i % n_batches) * Bg + (0 if single_gpu_global else rank * batch)
        return X_dev[r0:r0 + b_eng]

    return SimpleNamespace(trainer=trainer, eng=eng, model=model, data=data, X_train=X_train, X_dev=X_dev, batch=batch_fn,
                           n_fields=len(sizes), n_batches=n_batches, Bg=Bg, V=V)


def quick_measure(task, optimizer_mode, mask_ratio, batch, dev, data, steps=60, warmup=8, env=None):
    """device-timed samples/s of another configuration on ONE GPU (the `secondary` block: driver-timed evidence for the
    configurations north_star names beside the headline)"""
    saved = {}
    for k, v in (env or {}).items():
        saved[k] = os.environ.get(k)
        os.environ[k] = v
    try:
        run = build_run(task, optimizer_mode, mask_ratio, batch, 1, 0, dev, data=data)
        for i in range(warmup):
            run.eng.step(run.batch(i))
        torch.cuda.synchronize()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        for i in range(steps):
            run.eng.step(run.batch(warmup + i))
        ev1.record()
        torch.cuda.synchronize()
        run.eng.check_health()
        ms = ev0.elapsed_time(ev1) / steps
        return {"samples_per_s": batch / (ms * 1e-3), "ms_per_step": ms, "steps": steps, "loss_after": float(run.eng.outputs()[0])}
    finally:
        for k, v in saved.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
        torch.cuda.empty_cache()


def multi_gpu_parity(run, task, optimizer_mode, mask_ratio, batch, world, rank, dev, k_steps=2):
    """Multi-GPU parity evidence INSIDE the bench run (outside the timed region): the R-rank sharded step and a single-GPU
    FusedStep over the concatenated global batch start from the same seeded model and see the same k global batches; Philox
    counters are indexed by the global row, so both draw the same masks / noise.  Compared: the loss of every step and, after k
    steps, the UPDATE (p - p0) of every dense parameter and of every table (shards all-gathered to the reference layout).
    Every rank takes part in the collectives; rank 0 builds the single-GPU run and returns the record."""
    import torch.distributed as dist
    eng = run.eng
    sd0 = {k: v.clone() for k, v in eng.full_state_dict().items() if v.is_floating_point()}
    losses = []
    for i in range(k_steps):
        eng.step(run.batch(i))
        losses.append(float(eng.outputs()[0]))
    torch.cuda.synchronize()
    eng.check_health()
    sd_multi = eng.full_state_dict()
    rec = None
    if rank == 0:
        ref = build_run(task, optimizer_mode, mask_ratio, batch, world, 0, dev, data=run.data, single_gpu_global=True, graph=False)
        ref_losses = []
        for i in range(k_steps):
            ref.eng.step(ref.batch(i))
            ref_losses.append(float(ref.eng.outputs()[0]))
        torch.cuda.synchronize()
        sd_ref = ref.model.state_dict()
        upd = {}
        for kname, p0 in sd0.items():
            if kname not in sd_ref or "alias" in kname or "logprob" in kname:
                continue
            a, b = (sd_multi[kname].double() - p0.double()), (sd_ref[kname].double() - p0.double())
            upd[kname] = float((a - b).norm() / (b.norm() + 1e-30))
        rec = {"steps": k_steps, "reference": f"single-GPU FusedStep on the concatenated global batch ({batch * world} rows), same seed",
               "loss_multi": losses, "loss_single": ref_losses,
               "loss_rel": max(abs(a - b) / max(1.0, abs(b)) for a, b in zip(losses, ref_losses)),
               "update_rel": max(upd.values()), "update_rel_worst": max(upd, key=upd.get), "tensors_compared": len(upd)}
        del ref
        torch.cuda.empty_cache()
    dist.barrier()
    return rec


def run_eager_cuda(task, batch, data, dev, steps=20, warmup=3, mask_ratio=0.1):
    """The reference's algorithm as torch eager on the GPU ("that eager run is the GPU bar to beat"): the oracle
    port (= the reference's own torch ops: nn.functional.embedding / linear / BCE-with-logits, autograd backward, dense
    transformers-AdamW over every table) with its parameters on the device, fp32 with TF32 off (torch default, like the
    reference), dynamic_mask on the host with torch's generator followed by an H2D copy (trainer.py:217-266, 312-313) and the
    alias draw on the device with torch ops (alias_multinomial.py:81-97)."""
    from oracle import map_oracle as O
    sizes, V, X, fc, _ = data
    _, _, cfgd = config_dict(task)
    cfg = O.OracleConfig(**cfgd)
    params = {k: v.to(dev) for k, v in O.init_params(cfg, fc, seed=1).items()}
    ap = aa = None
    if task == "MFP":
        renormed, _, _ = O.nce_noise_distribution(fc)
        ap, aa = O.alias_build(renormed)
        ap, aa = ap.to(dev), aa.to(dev)
    tr = O.OracleTrainer(cfg, params, alias_prob=ap, alias_alias=aa, x_train=X, lr=1e-3, weight_decay=5e-2, mask_ratio=mask_ratio,
                         sampling_method="randint", seed=42, lr_lambda=O.cosine_schedule_lambda(0, 100000))
    F_ = len(sizes)
    L = int(F_ * mask_ratio)
    Kn = cfg.pt_neg_num
    Xn = X.numpy()
    g = torch.Generator().manual_seed(0)

    def one(i):
        xb = X[(i * batch) % (X.shape[0] - batch):][:batch]
        mi = torch.randint(0, F_, (batch, L), generator=g)                       # trainer.py:224-225
        if task == "MFP":
            ids, labels = O.dynamic_mask_mfp(xb, mi)
            b = {"input_ids": ids.to(dev), "labels": labels.to(dev), "masked_index": mi.to(dev)}
            kk = torch.randint(0, V, (batch, L, Kn), device=dev)                 # alias_multinomial.py:89-97
            keep = torch.bernoulli(ap[kk]).bool()
            b["noise"] = torch.where(keep, kk, aa[kk])
        else:
            si = torch.randint(0, X.shape[0], (batch * L,), generator=g)         # trainer.py:235-240
            rep = torch.from_numpy(Xn[si.numpy()])
            rf = torch.gather(rep, 1, mi.view(-1, 1)).view(batch, L)
            ids, labels = O.dynamic_mask_rfd(xb, mi, rf)
            b = {"input_ids": ids.to(dev), "labels": labels.to(dev)}
        outs = tr.forward_backward(b)
        tr.optimizer_step()
        return outs[0]

    for i in range(warmup):
        one(i)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(steps):
        loss = one(warmup + i)
    _ = float(loss)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    del tr, params
    torch.cuda.empty_cache()
    return {"value": batch * steps / dt, "unit": "samples/s", "ms_per_step": 1e3 * dt / steps, "steps": steps,
            "kind": "reference algorithm (oracle port) as torch eager on the same B200: fp32 (TF32 off), dense gradients + dense AdamW, "
                    "host dynamic_mask + H2D like the reference"}


def main_ours(args):
    import faulthandler
    import torch.distributed as dist
    faulthandler.enable()  # a SIGSEGV in native code prints the Python stack of every thread
    # stdout carries exactly ONE line (the JSON): libraries that write to fd 1 (NCCL prints its version banner there) are
    # sent to stderr for the whole run, the result line goes to the saved descriptor
    sys.stdout.flush()
    result_fd = os.dup(1)
    os.dup2(2, 1)
    from map_code_b200 import _lib, synthetic as S
    from map_code_b200.arguments import Config, TrainingArguments
    from map_code_b200.models import BaseModel
    from map_code_b200.trainer import Trainer

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} != WORLD_SIZE {world}")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        import datetime
        # a lost rank must not hang the box: collectives time out after 3 minutes, and every thread's stack is dumped if the
        # whole run is still alive after 12 minutes
        dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=180))
        faulthandler.dump_traceback_later(720, exit=True)
    _lib.load()  # fail loudly if the sm_100a library is missing

    run = build_run(args.task, args.optimizer_mode, MASK_RATIO, args.batch, world, rank, dev, graph=not args.no_graph)
    trainer, eng, X_train, n_fields, n_batches, Bg, batch = run.trainer, run.eng, run.X_train, run.n_fields, run.n_batches, run.Bg, run.batch
    parity = None
    if world > 1 and WORKLOAD != "c5":   # (the C5 tables do not fit twice on rank 0)
        parity = multi_gpu_parity(run, args.task, args.optimizer_mode, MASK_RATIO, args.batch, world, rank, dev)

    _stage('engine built')
    # ---- warm-up (also captures the CUDA graph)
    for i in range(max(args.warmup, 3)):
        eng.step(batch(i))
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()

    _stage('warm-up done')
    # ---- timed region: device-resident inputs, CUDA events, no host sync inside
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clocks:
        torch.cuda.synchronize()
        ev0.record()
        for i in range(args.steps):
            eng.step(batch(args.warmup + i))
        ev1.record()
        torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1)
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.barrier()
        ms = float(t.item())
    loss_after = float(eng.outputs()[0])
    eng.check_health()   # a device-side wait that gave up (sort grid barrier, peer barrier) voids the run: fail loudly
    value = Bg * args.steps / (ms / 1e3)

    _stage('timed region done')
    # ---- e2e: public API with HOST (pinned) inputs; H2D copy of the ids and D2H read of the loss every step
    host_batches = [X_train[(i % n_batches) * Bg + rank * args.batch:][:args.batch].contiguous().pin_memory() for i in range(8)]
    for i in range(3):
        float(trainer.train_step(host_batches[i % 8])[0])
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    pend = trainer.train_step_async(host_batches[0])
    _ = pend.loss()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    pend = None
    for i in range(args.steps):
        # every step: H2D copy of ITS batch from pinned memory (copy stream), the step, D2H copy of ITS loss; the host reads
        # the loss of step i-1 after submitting step i, so the transfers overlap the neighbouring steps' compute
        nxt = trainer.train_step_async(host_batches[i % 8])
        if pend is not None:
            _ = pend.loss()
        pend = nxt
    _ = pend.loss()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_s], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e = Bg * args.steps / e2e_s

    _stage('e2e done')
    # ---- per-kernel timing (eager replay of the same schedule, every C-ABI call bracketed by CUDA events on its stream)
    breakdown, roof, launches = None, None, None
    pk = peaks()
    # (every rank executes these steps — the sharded step contains collectives — but only rank 0 keeps the records)
    from map_code_b200 import ops as _ops0
    _ops0.SIMT_GEMMS.clear()
    _lib.LAUNCHES = {}
    eng.use_graph = False
    ms_flag, eng.multi_stream = eng.multi_stream, False   # serialise the branches: clean per-kernel durations
    eng.step(batch(0))
    torch.cuda.synchronize()
    launches = dict(_lib.LAUNCHES)
    simt_per_step = dict(_ops0.SIMT_GEMMS)   # GEMMs of ONE step that ran on the exact-fp32 CUDA-core kernel (skinny shapes only)
    _lib.LAUNCHES = None
    _lib.PROFILE = []
    for i in range(args.profile_steps):
        eng.step(batch(i + 1))
    torch.cuda.synchronize()
    prof_records, _lib.PROFILE = _lib.PROFILE, None
    _lib.RECORD = []
    eng.step(batch(0))
    torch.cuda.synchronize()
    call_records, _lib.RECORD = _lib.RECORD, None
    eng.multi_stream = ms_flag
    if world > 1:
        dist.barrier()
    if rank == 0:
        _lib.PROFILE = prof_records
        agg = {}
        for name, tag, a, b in _lib.PROFILE:
            key = name
            d = agg.setdefault(key, dict(ms=0.0, n=0, bytes=0.0, flops=0.0))
            d["ms"] += a.elapsed_time(b)
            d["n"] += 1
            by, fl = algorithmic(tag)
            d["bytes"] += by
            d["flops"] += fl
        if args.dump_profile:
            per = {}
            for name, tag, a, b in _lib.PROFILE:
                d = per.setdefault((name, tag), [0.0, 0])
                d[0] += a.elapsed_time(b)
                d[1] += 1
            with open(args.dump_profile, "w") as f:
                for (name, tag), (msum, n) in sorted(per.items(), key=lambda kv: -kv[1][0]):
                    by, fl = algorithmic(tag)
                    us = 1e3 * msum / n
                    f.write(f"{name:34s} {str(tag):44s} n/step={n / args.profile_steps:5.1f} us={us:9.1f} "
                            f"tflops={fl / (us * 1e-6) / 1e12 if fl else 0:7.1f} gbs={by / (us * 1e-6) / 1e9 if by else 0:8.1f}\n")
        _lib.PROFILE = None
        eng.use_graph = not args.no_graph
        tot = sum(d["ms"] for d in agg.values())
        breakdown = []
        for name, d in sorted(agg.items(), key=lambda kv: -kv[1]["ms"]):
            row = {"kernel": name, "calls_per_step": d["n"] / args.profile_steps, "us_per_step": 1e3 * d["ms"] / args.profile_steps,
                   "share": d["ms"] / tot}
            if d["flops"] > 0 and "gemm" in name:
                row["tflops"] = d["flops"] / (d["ms"] * 1e-3) / 1e12
            if d["bytes"] > 0:
                row["gbs"] = d["bytes"] / (d["ms"] * 1e-3) / 1e9
            breakdown.append(row)
        top = breakdown[0]
        d = agg[top["kernel"]]
        if "gemm" in top["kernel"]:
            # the tensor-core GEMM class (split-bf16 by default; TF32 with MAP_B200_GEMM=tf32): average launch duration from a graph
            # replay of exactly these launches (world 1; the eager event-bracketed figure stays in kernels[])
            is_gemm = lambda n: "gemm_bf16s" in n or "gemm_tf32" in n   # noqa: E731
            gemm_recs = [r for r in call_records if is_gemm(r[0])]
            fl = sum(algorithmic(r[2])[1] for r in gemm_recs)
            # tensor-pipe work actually issued: a split-bf16 GEMM with `terms` bf16 MMAs per fp32 product
            issued = sum(2.0 * t[0] * t[1] * t[2] * (t[3] if len(t) > 3 else 1) for r in gemm_recs if r[2] and r[2][0] == "gemm_group" for t in r[2][1:])
            if world == 1 and gemm_recs:
                sec, n_l = replay_kernel_class(gemm_recs)
            else:
                sec, n_l = sum(agg[k]["ms"] for k in agg if is_gemm(k)) * 1e-3 / args.profile_steps, len(gemm_recs)
            ach = fl / sec / 1e12
            backend = "bf16s" if any("bf16s" in r[0] for r in gemm_recs) else "tf32"
            roof = {"kernel": ("map_gemm_bf16s_group (tcgen05 kind::f16, split-bf16: 3 or 6 bf16 MMAs per fp32 product)" if backend == "bf16s"
                               else "map_gemm_tf32_group + map_gemm_tf32_tcgen05 (tcgen05 TF32 GEMM)"),
                    "bound": "tensor", "achieved": ach, "peak": pk["tensor_sustained"], "unit": "TFLOP/s", "frac": ach / pk["tensor_sustained"],
                    "traffic": measured_traffic("gemm_" + backend, WORKLOAD, args.task) if world == 1 else None,
                    "launches_per_step": n_l, "us_per_launch": 1e6 * sec / max(n_l, 1), "flops_per_step": fl,
                    "tensor_pipe_tflops_issued": issued / sec / 1e12 if backend == "bf16s" else None,
                    "tensor_pipe_frac_of_peak": issued / sec / 1e12 / pk["tensor_sustained"] if backend == "bf16s" else None,
                    "note": ("achieved = ALGORITHMIC fp32 flops (2MNK) of all GEMM launches of one step / their device time (launches replayed back to "
                             "back in a CUDA graph on one stream); peak = " + pk["src"] + " sustained bf16 cuBLAS.  The split-bf16 backend issues "
                             "3 (6 for ReLU-feeding forward GEMMs) bf16 MMAs per fp32 product to reach fp32-level accuracy, so frac <= 1/3 by "
                             "construction; tensor_pipe_* count the bf16 MMA work actually issued.") if backend == "bf16s" else
                            ("TF32 operands (nominal rate = half of bf16); peak = " + pk["src"] + " sustained bf16 cuBLAS; achieved = algorithmic "
                             "flops of all tcgen05 GEMM launches of one step / their device time, launches replayed back to back in a CUDA graph on one stream")}
        else:
            ach = d["bytes"] / (d["ms"] * 1e-3) / 1e9
            roof = {"kernel": top["kernel"], "bound": "hbm", "achieved": ach, "peak": pk["hbm"], "unit": "GB/s", "frac": ach / pk["hbm"],
                    "traffic": None, "note": f"peak = {pk['src']} copy bandwidth"}

    _stage('profiling pass done')
    if args.timeline and not args.no_graph:   # every rank re-captures and replays (collectives), rank 0 writes the file
        write_timeline(eng, batch, args.timeline if rank == 0 else os.devnull)
    if rank != 0:
        _shutdown(world)
        return

    cpu = None
    if not args.no_cpu_baseline and world == 1:
        r = run_cpu(args.task, args.batch, steps=30, warmup=2, budget_s=15.0)
        cpu = {"value": r["value"], "unit": "samples/s", "cores": r["cores"], "kind": "port",
               "sample": f"{r['steps']} full oracle steps at batch {args.batch} on the host (dense grads + dense AdamW like the reference), {r['ms_per_step']:.0f} ms/step"}

    # ---- beside the headline (single GPU, default run only): the reference algorithm as torch eager on this GPU, and the other
    # configurations north_star names, each device-timed for a bounded number of steps
    gpu_ref, secondary = None, None
    field_enc_used = getattr(eng, "field_enc", False)
    if world == 1 and not args.no_secondary and WORKLOAD != "c5":
        del eng, trainer
        run.eng = run.trainer = run.model = None
        torch.cuda.empty_cache()
        try:
            gpu_ref = run_eager_cuda(args.task, args.batch, run.data, dev, steps=20, mask_ratio=MASK_RATIO)
        except Exception as e:  # noqa: BLE001  (the headline must survive a failure of the side measurement)
            gpu_ref = {"error": repr(e)[:300]}
        secondary = {}
        if WORKLOAD == "c2":
            other = "RFD" if args.task == "MFP" else "MFP"
            plan = [(f"c2 DCNv2 {other} mask_ratio {MASK_RATIO:g} sparse", other, "sparse", MASK_RATIO, None),
                    (f"c2 DCNv2 {args.task} mask_ratio {MASK_RATIO:g} dense_exact (the reference's optimizer semantics on every table row)", args.task, "dense_exact", MASK_RATIO, None),
                    (f"c2 DCNv2 {args.task} mask_ratio 0.3 sparse (the reference's own run scripts)", args.task, "sparse", 0.3, None),
                    (f"c2 DCNv2 {args.task} mask_ratio {MASK_RATIO:g} sparse, TF32 single-pass GEMMs (fails the 1e-3 gradient tolerance: A/B only)", args.task, "sparse", MASK_RATIO, {"MAP_B200_GEMM": "tf32"})]
            if args.task == "MFP":
                plan.append((f"c2 DCNv2 MFP mask_ratio {MASK_RATIO:g} sparse, encoder by field (forward + weight gradient for the L masked slices only, csrc/fieldenc.cu; A/B of the dense encoder GEMM)", "MFP", "sparse", MASK_RATIO, {"MAP_B200_FIELD_ENC": "1"}))
            for name, task_, mode_, ratio_, env_ in plan:
                try:
                    secondary[name] = quick_measure(task_, mode_, ratio_, args.batch, dev, run.data, env=env_)
                except Exception as e:  # noqa: BLE001
                    secondary[name] = {"error": repr(e)[:300]}

    from map_code_b200 import ops as _ops
    n_launch = sum(launches.values()) if launches else None
    line = {
        "metric": f"{args.task} pretrain samples/sec ({'DeepFM, Avazu shape' if WORKLOAD == 'c4' else 'DCNv2, Criteo shape'})", "value": value,
        "unit": "samples/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": DTYPE_NOTE[gemm_backend_name()], "data": "synthetic",
        "config": shared_config(args.task, args.batch, world),
        "impl_detail": {"gemm_backend": gemm_backend_name(), "optimizer_mode": args.optimizer_mode, "cuda_graph": not args.no_graph,
                        "mfp_encoder": ({"hybrid": "by field (only the L masked slices, csrc/fieldenc.cu): forward + weight gradient; input gradient as dense GEMMs", "full": "by field: forward, dgrad, wgrad"}[field_enc_used] if field_enc_used else "dense [B, F*P] GEMM + gather") if args.task == "MFP" else None,
                        "parallelism": "single GPU" if world == 1 else f"tables row-sharded by id mod {world} in NVLink peer memory (remote rows read directly by the gather / NCE kernels, owners pull gradients); dense params replicated + NCCL all-reduce",
                        "simt_gemm_calls_per_step": {f"{k[0]}x{k[1]}x{k[2]}": v for k, v in simt_per_step.items()}},
        "e2e": {"value": e2e, "unit": "samples/s", "h2d_bytes_per_step": args.batch * n_fields * 8, "d2h_bytes_per_step": 4,
                "ms_per_step": 1e3 * e2e_s / args.steps},
        "gpu_launches": (n_launch * args.steps) if n_launch else None, "launches_per_step": n_launch,
        "clocks": clocks.summary(), "roofline": roof, "cpu_baseline": cpu, "gpu_eager_reference": gpu_ref, "secondary": secondary,
        "parity": parity, "kernels": breakdown, "loss_after": loss_after,
    }
    sys.stdout.flush()
    os.write(result_fd, (json.dumps(line) + "\n").encode())
    if parity is not None and (parity["loss_rel"] > 1e-3 or parity["update_rel"] > 2e-2):
        print(f"multi-GPU parity FAILED: {parity}", file=sys.stderr, flush=True)
        os._exit(3)
    _shutdown(world)


if __name__ == "__main__":
    a = parse()
    WORKLOAD = a.workload
    MASK_RATIO = a.mask_ratio
    if a.batch is None:
        a.batch = 65536 if WORKLOAD == "c5" else B_PER_GPU
    if a.impl == "reference":
        main_reference(a)
    else:
        main_ours(a)
