"""xDeepFM pretraining step at the Criteo shape on one B200: our module path (autograd over the sm_100a kernels, CIN as
pair-major tensor-core GEMMs, sparse row-wise AdamW on the tables) next to the reference algorithm as torch eager on the same GPU
(oracle port = the reference's own torch ops, fp32 with TF32 off, dense gradients + dense AdamW).  One JSON line on stdout.
Usage: python scripts/bench_xdeepfm.py [--task MFP|RFD] [--batch 4096] [--steps 20] [--warmup 3]"""
import argparse
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


class DS:
    def __init__(self, X):
        self.X = X

    def __len__(self):
        return self.X.shape[0]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--task", default="MFP")
    ap.add_argument("--batch", type=int, default=4096)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--units", default="50,50")
    a = ap.parse_args()
    from map_code_b200 import synthetic as S
    from map_code_b200.arguments import Config, TrainingArguments
    from map_code_b200.models import BaseModel
    from map_code_b200.trainer import Trainer
    from oracle import map_oracle as O
    dev = torch.device("cuda:0")
    sizes = S.field_sizes("criteo")
    V = S.vocab_size(sizes)
    X = S.make_ids(sizes, 1 << 16, seed=0)
    fc = S.feat_count(X, V)
    cfgd = dict(model_name="xDeepFM", embed_size=16, hidden_size=1000, num_hidden_layers=3, hidden_act="relu", hidden_dropout_rate=0.0,
                embed_dropout_rate=0.0, embed_norm=False, layer_norm_eps=1e-12, pt_neg_num=25, proj_size=32, input_size=V,
                num_fields=len(sizes), pretrain=True, pt_type=a.task, RFD_replace="Unigram", cin_layer_units=a.units, use_lr=False)
    torch.manual_seed(1)
    model = BaseModel.from_config(Config.from_dict(dict(cfgd, feat_count=fc, data_dir=None, seed=42, table_grad_mode="sparse"))).to(dev)
    targs = TrainingArguments(per_gpu_train_batch_size=a.batch, learning_rate=1e-3, weight_decay=5e-2, lr_sched="cosine",
                              sampling_method="randint", mask_ratio=0.1, pretrain=True, pt_type=a.task, seed=42)
    Xd = X.to(dev)
    tr = Trainer(model, model.config, targs, DS(Xd), DS(Xd))
    opt, sched = tr.get_optimizer(100000, 0)
    nb = X.shape[0] // a.batch

    def ours(i):
        xb = Xd[(i % nb) * a.batch:(i % nb + 1) * a.batch]
        inputs = tr.dynamic_mask({"input_ids": xb}, "randint", step=i)
        out = model(**inputs)
        out[0].backward()
        opt.step()
        sched.step()
        opt.zero_grad()
        return out[0]

    for i in range(a.warmup):
        ours(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(a.steps):
        loss = ours(a.warmup + i)
    e1.record()
    torch.cuda.synchronize()
    ms_ours = e0.elapsed_time(e1) / a.steps
    loss_ours = float(loss)
    from map_code_b200 import ops
    simt = dict(ops.SIMT_GEMMS)
    del model, tr, opt
    torch.cuda.empty_cache()

    # the reference algorithm, torch eager on the same GPU
    cfg = O.OracleConfig(**cfgd)
    params = {k: v.to(dev) for k, v in O.init_params(cfg, fc, seed=1).items()}
    apb = aab = None
    if a.task == "MFP":
        renormed, _, _ = O.nce_noise_distribution(fc)
        apb, aab = (t.to(dev) for t in O.alias_build(renormed))
    otr = O.OracleTrainer(cfg, params, alias_prob=apb, alias_alias=aab, x_train=X, lr=1e-3, weight_decay=5e-2, mask_ratio=0.1,
                          sampling_method="randint", seed=42, lr_lambda=O.cosine_schedule_lambda(0, 100000))
    F_, L, Kn = len(sizes), int(len(sizes) * 0.1), 25
    g = torch.Generator().manual_seed(0)

    def eager(i):
        xb = X[(i % nb) * a.batch:(i % nb + 1) * a.batch]
        mi = torch.randint(0, F_, (a.batch, L), generator=g)
        if a.task == "MFP":
            ids, labels = O.dynamic_mask_mfp(xb, mi)
            b = {"input_ids": ids.to(dev), "labels": labels.to(dev), "masked_index": mi.to(dev)}
            kk = torch.randint(0, V, (a.batch, L, Kn), device=dev)
            b["noise"] = torch.where(torch.bernoulli(apb[kk]).bool(), kk, aab[kk])
        else:
            si = torch.randint(0, X.shape[0], (a.batch * L,), generator=g)
            rf = torch.gather(X[si], 1, mi.view(-1, 1)).view(a.batch, L)
            ids, labels = O.dynamic_mask_rfd(xb, mi, rf)
            b = {"input_ids": ids.to(dev), "labels": labels.to(dev)}
        outs = otr.forward_backward(b)
        otr.optimizer_step()
        return outs[0]

    for i in range(a.warmup):
        eager(i)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(a.steps):
        loss = eager(a.warmup + i)
    _ = float(loss)
    torch.cuda.synchronize()
    ms_eager = 1e3 * (time.perf_counter() - t0) / a.steps
    print(json.dumps({
        "workload": f"xDeepFM {a.task} pretraining, synthetic Criteo shape (39 fields, V={V}), embed 16, CIN {a.units}, hidden 1000x3, proj 32, K=25, "
                    f"mask_ratio 0.1, batch {a.batch}, 1 GPU",
        "ours": {"path": "module path (autograd over our kernels; CIN = outer-product kernel + split-bf16 tcgen05 GEMM over the pairs; sparse "
                         "row-wise AdamW on the tables)", "ms_per_step": ms_ours, "samples_per_s": a.batch / ms_ours * 1e3, "loss_after": loss_ours,
                 "simt_gemm_calls": {f"{k[0]}x{k[1]}x{k[2]}": v for k, v in simt.items()}},
        "gpu_eager_reference": {"ms_per_step": ms_eager, "samples_per_s": a.batch / ms_eager * 1e3,
                                "kind": "oracle port (the reference's torch ops) on the same B200, fp32 TF32 off, dense grads + dense AdamW"},
        "steps": a.steps, "warmup": a.warmup}))


if __name__ == "__main__":
    main()
