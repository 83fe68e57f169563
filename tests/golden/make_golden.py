"""Generates the golden fixtures in this directory by RUNNING THE UNMODIFIED REFERENCE (CPU, torch eager).

Run in the build container only (needs /root/reference):   python tests/golden/make_golden.py
The GPU box never runs this; it only reads the frozen *.pt files.

What the reference consumes from us: index tensors only.
  * mask / replace indices: `trainer.torch.randint` is replaced by a feeder that pops our tensors, so the reference's own
    gather / scatter / compare code (trainer.py:229-240) runs on them;
  * NCE noise: `model.mfp_criterion.alias.draw` is replaced by a lambda returning our noise tensor;
  * weights: taken from the reference module's own `state_dict()` after construction (reference init, torch seed).
Everything stored under "out" is produced by reference code.
"""
import math
import os
import sys
import tempfile

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import ref_shim  # noqa: E402

ref = ref_shim.import_reference()
torch.set_num_threads(1)


def save(name, obj):
    path = os.path.join(HERE, name + ".pt")
    torch.save(obj, path)
    print("wrote", path, os.path.getsize(path), "bytes")


def synthetic_ids(gen, n_rows, field_sizes, reserved=10):
    """ids 0-9 reserved; field f owns [low_f, low_f+size_f) (data_preprocess/proc_criteo.py:109-161 layout)."""
    lows = np.cumsum([reserved] + list(field_sizes[:-1]))
    cols = []
    for lo, sz in zip(lows, field_sizes):
        u = torch.rand(n_rows, generator=gen)
        r = torch.floor(torch.pow(torch.tensor(float(sz)), u)).long() - 1
        cols.append(lo + r.clamp_(0, sz - 1))
    return torch.stack(cols, 1), lows, reserved + int(sum(field_sizes))


def make_config(tmp, **kw):
    d = dict(model_name="DCNv2", embed_size=4, hidden_size=16, num_hidden_layers=3, num_cross_layers=3,
             hidden_act="relu", hidden_dropout_rate=0.0, embed_dropout_rate=0.0, embed_norm=False,
             layer_norm_eps=1e-12, pt_neg_num=5, proj_size=8, pretrain=True, pt_type="MFP", RFD_replace="Unigram",
             data_dir=tmp, idx_low=None, idx_high=None)
    d.update(kw)
    return ref.arguments.Config.from_dict(d)


class Args:
    """The subset of TrainingArguments that Trainer touches in dynamic_mask/get_optimizer (arguments.py:15-47)."""

    def __init__(self, **kw):
        self.device = torch.device("cpu")
        self.mask_ratio = 0.34
        self.pt_type = "MFP"
        self.RFD_replace = "Unigram"
        self.sampling_method = "randint"
        self.weight_decay = 5e-2
        self.learning_rate = 1e-3
        self.adam_epsilon = 1e-8
        self.adam_betas = "0.9,0.999"
        self.lr_sched = "cosine"
        for k, v in kw.items():
            setattr(self, k, v)


class DS:
    def __init__(self, X):
        self.X = X

    def __len__(self):
        return len(self.X)


def feed_randint(tensors):
    feed = list(tensors)

    class _T:
        def __getattr__(self, name):
            if name == "randint":
                return lambda *a, **k: feed.pop(0)
            return getattr(torch, name)

    return _T()


# ---------------------------------------------------------------------------------------------------------------
def golden_alias():
    with tempfile.TemporaryDirectory() as tmp:
        cfg = make_config(tmp)
        # KAT of SURVEY.md §8c: feat_count = arange(1,17)
        fc = torch.arange(1, 17).float()
        probs = fc / fc.sum()
        probs = probs.clamp(min=1e-10)
        probs = probs / probs.sum()
        a = ref.nce.alias_multinomial.AliasMultinomial(probs, cfg)
        kat = dict(probs=probs, prob=a.prob.clone(), alias=a.alias.clone())
    with tempfile.TemporaryDirectory() as tmp:
        cfg = make_config(tmp)
        g = torch.Generator().manual_seed(7)
        fc = torch.floor(torch.pow(torch.tensor(50000.0), torch.rand(3000, generator=g)))  # Zipf-ish counts
        fc[::17] = 0  # never-seen ids -> BACKOFF_PROB clamp path (nce_loss.py:63)
        probs = fc / fc.sum()
        probs = probs.clamp(min=1e-10)
        probs = probs / probs.sum()
        a = ref.nce.alias_multinomial.AliasMultinomial(probs, cfg)
        big = dict(feat_count=fc, probs=probs, prob=a.prob.clone(), alias=a.alias.clone())
    save("alias", dict(kat=kat, zipf3000=big))


def golden_nce_kat():
    """SURVEY.md §8c NCE known-answer: V=16, P=4, K=3."""
    with tempfile.TemporaryDirectory() as tmp:
        cfg = make_config(tmp, input_size=16, num_fields=4, proj_size=4, pt_neg_num=3,
                          feat_count=torch.arange(1, 17).float())
        il = ref.nce.IndexLinear(cfg)
        il.emb.weight.data = torch.sin(torch.arange(64).float()).view(16, 4)
        inp = torch.cos(torch.arange(16).float()).view(2, 2, 4).requires_grad_(True)
        target = torch.tensor([[10, 11], [12, 3]])
        noise = torch.tensor([[[1, 2, 15], [4, 4, 5]], [[15, 14, 13], [0, 9, 10]]])
        il.alias.draw = lambda *s: noise
        outs = {}
        for lt in ("nce", "sampled"):
            il.zero_grad()
            inp.grad = None
            il.loss_type = lt
            loss, logits, ids = il(target, inp)
            loss.backward()
            outs[lt] = dict(loss=loss.detach().clone(), logits=logits.detach().clone(), ids=ids.clone(),
                            d_input=inp.grad.clone(), d_emb=il.emb.weight.grad.clone(), d_bias=il.bias.weight.grad.clone())
        full = il.ce_loss(target, inp.detach())
        save("nce_kat", dict(emb=il.emb.weight.detach().clone(), bias=il.bias.weight.detach().clone(),
                             logprob_noise=il.logprob_noise.clone(), norm_term=il.norm_term, input=inp.detach().clone(),
                             target=target, noise=noise, out=outs, full_ce=full.detach().clone(),
                             alias_prob=il.alias.prob.clone(), alias_alias=il.alias.alias.clone()))


def golden_nce_batched():
    """per_word = False: one noise draw shared by every position -> the reference takes the batched-logit path
    (index_linear.py:108-143: a real [N,P] x [P,K] matmul) and the full-softmax ce_loss (index_linear.py:145-151) on a case
    large enough to exercise tiling (V = 300, P = 8, N = 15 positions)."""
    g = torch.Generator().manual_seed(21)
    V, P, K, B, L = 300, 8, 5, 5, 3
    with tempfile.TemporaryDirectory() as tmp:
        fc = torch.floor(torch.pow(torch.tensor(300.0), torch.rand(V, generator=g)))
        cfg = make_config(tmp, input_size=V, num_fields=6, proj_size=P, pt_neg_num=K, feat_count=fc)
        il = ref.nce.IndexLinear(cfg)
        il.emb.weight.data = torch.randn(V, P, generator=g) * 0.3
        inp = torch.randn(B, L, P, generator=g).requires_grad_(True)
        target = torch.randint(0, V, (B, L), generator=g)
        noise1 = torch.randint(0, V, (1, 1, K), generator=g)
        il.per_word = False
        il.alias.draw = lambda *s: noise1
        outs = {}
        for lt in ("nce", "sampled"):
            il.zero_grad()
            inp.grad = None
            il.loss_type = lt
            loss, logits, ids = il(target, inp)
            loss.backward()
            outs[lt] = dict(loss=loss.detach().clone(), logits=logits.detach().clone(), ids=ids.clone(), d_input=inp.grad.clone(),
                            d_emb=il.emb.weight.grad.clone(), d_bias=il.bias.weight.grad.clone())
        full = il.ce_loss(target, inp.detach())
        save("nce_batched", dict(emb=il.emb.weight.detach().clone(), bias=il.bias.weight.detach().clone(),
                                 logprob_noise=il.logprob_noise.clone(), norm_term=il.norm_term, input=inp.detach().clone(), target=target,
                                 noise1=noise1, out=outs, full_ce=full.detach().clone(), feat_count=fc))


def golden_dynamic_mask():
    X = (np.arange(60).reshape(10, 6) + 100).astype(np.int64)
    cfg = ref.arguments.Config.from_dict(dict(num_fields=6, input_size=200,
                                              idx_low=torch.tensor([100, 101, 102, 103, 104, 105]),
                                              idx_high=torch.tensor([155, 156, 157, 158, 159, 160])))
    mi = torch.tensor([[0, 5], [2, 2]])
    out = {}
    batch = torch.from_numpy(X[:2].copy())
    # MFP
    t = ref.trainer.Trainer(None, cfg, Args(pt_type="MFP"), DS(X), DS(X))
    ref.trainer.torch = feed_randint([mi])
    r = t.dynamic_mask({"input_ids": batch.clone(), "labels": torch.zeros(2)}, "randint")
    out["MFP"] = {k: v.clone() for k, v in r.items()}
    # RFD Unigram
    t = ref.trainer.Trainer(None, cfg, Args(pt_type="RFD", RFD_replace="Unigram"), DS(X), DS(X))
    si = torch.tensor([7, 1, 9, 3])
    ref.trainer.torch = feed_randint([mi, si])
    r = t.dynamic_mask({"input_ids": batch.clone(), "labels": torch.zeros(2)}, "randint")
    out["RFD-Unigram"] = {k: v.clone() for k, v in r.items()}
    out["RFD-Unigram"]["sample_index"] = si
    # RFD Whole-Unigram
    t = ref.trainer.Trainer(None, cfg, Args(pt_type="RFD", RFD_replace="Whole-Unigram"), DS(X), DS(X))
    f2 = torch.tensor([[3, 0], [1, 4]])
    ref.trainer.torch = feed_randint([mi, si, f2])
    r = t.dynamic_mask({"input_ids": batch.clone(), "labels": torch.zeros(2)}, "randint")
    out["RFD-Whole-Unigram"] = {k: v.clone() for k, v in r.items()}
    out["RFD-Whole-Unigram"]["sample_index"] = si
    out["RFD-Whole-Unigram"]["field2"] = f2
    # RFD Uniform: reference draws one column per field then gathers column mi (trainer.py:242-243)
    t = ref.trainer.Trainer(None, cfg, Args(pt_type="RFD", RFD_replace="Uniform"), DS(X), DS(X))
    cols = [torch.tensor([100 + f + 10 * j for j in range(4)]) for f in range(6)]
    ref.trainer.torch = feed_randint([mi] + cols)
    r = t.dynamic_mask({"input_ids": batch.clone(), "labels": torch.zeros(2)}, "randint")
    out["RFD-Uniform"] = {k: v.clone() for k, v in r.items()}
    out["RFD-Uniform"]["replace_feat"] = torch.stack(cols, 1).gather(1, mi.view(-1, 1)).view(2, 2)
    # RFD Whole-Uniform
    t = ref.trainer.Trainer(None, cfg, Args(pt_type="RFD", RFD_replace="Whole-Uniform"), DS(X), DS(X))
    whole = torch.arange(24).view(4, 6) + 20
    ref.trainer.torch = feed_randint([mi, whole])
    r = t.dynamic_mask({"input_ids": batch.clone(), "labels": torch.zeros(2)}, "randint")
    out["RFD-Whole-Uniform"] = {k: v.clone() for k, v in r.items()}
    out["RFD-Whole-Uniform"]["replace_feat"] = whole.gather(1, mi.view(-1, 1)).view(2, 2)
    ref.trainer.torch = torch
    save("dynamic_mask", dict(X=torch.from_numpy(X), masked_index=mi, out=out))


def _grads(model):
    return {k: (p.grad.clone() if p.grad is not None else None) for k, p in model.named_parameters()}


def golden_model(name, model_name, pt_type, pretrain=True, F=6, D=4, H=16, P=8, K=5, B=12, mask_ratio=0.34,
                 field_sizes=(7, 40, 3, 120, 15, 60), steps=3, extra=None):
    """A small model end-to-end through the reference: forward, backward, and `steps` HF-AdamW steps driven by the
    reference Trainer.get_optimizer (param grouping + cosine schedule: trainer.py:60-85)."""
    g = torch.Generator().manual_seed(11)
    X, lows, V = synthetic_ids(g, 256, field_sizes)
    feat_count = torch.bincount(X.flatten(), minlength=V).float()
    torch.manual_seed(5)
    with tempfile.TemporaryDirectory() as tmp:
        cfg = make_config(tmp, model_name=model_name, pt_type=pt_type, pretrain=pretrain, input_size=V, num_fields=F,
                          embed_size=D, hidden_size=H, proj_size=P, pt_neg_num=K, feat_count=feat_count, **(extra or {}))
        model = ref.models.BaseModel.from_config(cfg)
    L = int(F * mask_ratio)
    args = Args(pt_type=pt_type, mask_ratio=mask_ratio)
    trainer = ref.trainer.Trainer(model, cfg, args, DS(X.numpy()), DS(X.numpy()))
    optimizer, scheduler = trainer.get_optimizer(num_training_steps=10, num_warmup_steps=2)
    sd0 = {k: v.clone() for k, v in model.state_dict().items()}
    rec = dict(config=dict(model_name=model_name, pt_type=pt_type, pretrain=pretrain, input_size=V, num_fields=F,
                           embed_size=D, hidden_size=H, proj_size=P, pt_neg_num=K, num_hidden_layers=3,
                           num_cross_layers=3, mask_ratio=mask_ratio, **(extra or {})),
               feat_count=feat_count, X_train=X, state_dict0=sd0, steps=[],
               optim=dict(lr=1e-3, weight_decay=5e-2, eps=1e-8, betas=(0.9, 0.999), sched="cosine", t_total=10, t_warmup=2))
    gi = torch.Generator().manual_seed(3)
    model.train()
    for s in range(steps):
        rows = torch.randint(0, X.shape[0], (B,), generator=gi)
        batch = X[rows].clone()
        st = dict(batch=batch.clone())
        if pretrain:
            mi = torch.randint(0, F, (B, L), generator=gi)
            st["masked_index"] = mi
            if pt_type == "MFP":
                ref.trainer.torch = feed_randint([mi])
                inputs = trainer.dynamic_mask({"input_ids": batch, "labels": torch.zeros(B)}, "randint")
                noise = torch.randint(0, V, (B, L, K), generator=gi)
                noise[0, 0, 0] = inputs["labels"][0, 0]  # a noise sample equal to the target (argmax tie)
                model.mfp_criterion.alias.draw = lambda *sz, _n=noise: _n
                st["noise"] = noise
            else:
                si = torch.randint(0, X.shape[0], (B * L,), generator=gi)
                ref.trainer.torch = feed_randint([mi, si])
                inputs = trainer.dynamic_mask({"input_ids": batch, "labels": torch.zeros(B)}, "randint")
                st["sample_index"] = si
            ref.trainer.torch = torch
            st["inputs"] = {k: v.clone() for k, v in inputs.items()}
            outputs = model(**inputs)
        else:
            y = (torch.rand(B, generator=gi) < 0.3).long()
            st["labels"] = y
            outputs = model(input_ids=batch, labels=y)
        loss = outputs[0]
        loss.backward()
        st["loss"] = loss.detach().clone()
        st["outputs"] = [o.detach().clone() if torch.is_tensor(o) else o for o in outputs[1:]]
        if pretrain and pt_type == "MFP":
            # logits/ids are not returned by get_outputs; recompute through the criterion for the record
            with torch.no_grad():
                enc = model.feat_encoder(_final(model, inputs["input_ids"])).view(B, F, P)
                sel = torch.gather(enc, 1, inputs["masked_index"].unsqueeze(-1).repeat(1, 1, P))
                _, lg, ids = model.mfp_criterion(inputs["labels"], sel)
            st["logits"], st["ids"] = lg.clone(), ids.clone()
        st["grads"] = _grads(model)
        st["lr"] = scheduler.get_last_lr()[0]
        optimizer.step()
        scheduler.step()
        model.zero_grad()
        st["state_dict_after"] = {k: v.clone() for k, v in model.state_dict().items()
                                  if k in dict(model.named_parameters())}
        rec["steps"].append(st)
    save(name, rec)


def _final(model, ids):
    """final_output of the reference backbone (DCNV2.forward models.py:308-315 / DeepFM.forward :219-225)."""
    if model.model_name == "DCNV2":
        fe = model.embed(ids).flatten(start_dim=1)
        return torch.cat([model.cross_net(fe), model.parallel_dnn(fe)], dim=-1)
    fe = model.embed(ids)
    dnn = model.dnn(fe.flatten(start_dim=1))
    if model.model_name == "DNN":   # DNN.forward models.py:183-186
        return dnn
    if model.model_name == "xDeepFM":   # xDeepFM.forward models.py:262-269
        return torch.cat([model.cin(fe), dnn], dim=1)
    return torch.cat([dnn, model.lr_layer(ids)[0] + model.ip_layer(fe)], dim=1)


if __name__ == "__main__":
    golden_alias()
    golden_nce_kat()
    golden_dynamic_mask()
    golden_nce_batched()
    golden_model("dcnv2_mfp", "DCNv2", "MFP")
    golden_model("dcnv2_rfd", "DCNv2", "RFD")
    golden_model("deepfm_mfp", "DeepFM", "MFP")
    golden_model("dcnv2_ctr", "DCNv2", "MFP", pretrain=False)
    golden_model("deepfm_ctr", "DeepFM", "MFP", pretrain=False)
    golden_model("dnn_mfp", "DNN", "MFP")          # models.py:164-193
    golden_model("dnn_rfd", "DNN", "RFD")
    golden_model("dnn_ctr", "DNN", "MFP", pretrain=False)
    xd = dict(cin_layer_units="5,4", use_lr=True)  # models.py:235-279, layers.py:696-721; use_lr adds the LR term to the CTR logits
    golden_model("xdeepfm_mfp", "xDeepFM", "MFP", extra=xd)
    golden_model("xdeepfm_rfd", "xDeepFM", "RFD", extra=xd)
    golden_model("xdeepfm_ctr", "xDeepFM", "MFP", pretrain=False, extra=xd)
