"""The reference-shaped public API end to end on the GPU: Trainer.MFP_pretrain / RFD_pretrain / train / eval / save / finetune
on a small synthetic dataset (the flow of code/run.py:66-84)."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


class DS:
    def __init__(self, X, Y):
        self.X, self.Y = X, Y

    def __len__(self):
        return len(self.Y)


def make(pt_type="MFP", pretrain=True, model_name="DCNv2", tmp="."):
    from map_code_b200 import synthetic as S
    from map_code_b200.arguments import Config, TrainingArguments
    from map_code_b200.models import BaseModel
    F = 39 if model_name != "DeepFM" else 24
    sizes = [max(2, s // 300) for s in (S.field_sizes("criteo") if F == 39 else S.field_sizes("avazu"))]
    V = S.vocab_size(sizes)
    X = S.make_ids(sizes, 1280, seed=0)
    Y = (torch.rand(1280, generator=torch.Generator().manual_seed(1)) < 0.2).long()
    cfg = Config.from_dict(dict(model_name=model_name, embed_size=16, hidden_size=64, num_hidden_layers=3, num_cross_layers=3,
                                hidden_act="relu", hidden_dropout_rate=0.0, embed_dropout_rate=0.0, embed_norm=False, layer_norm_eps=1e-12,
                                pt_neg_num=25, proj_size=32, input_size=V, num_fields=F, pretrain=pretrain, pt_type=pt_type,
                                RFD_replace="Unigram", feat_count=S.feat_count(X, V), data_dir=None, seed=42, table_grad_mode="sparse",
                                cin_layer_units="20,12", use_lr=True))
    args = TrainingArguments(output_dir=str(tmp), per_gpu_train_batch_size=256, per_gpu_eval_batch_size=512, learning_rate=1e-3,
                             weight_decay=5e-2, num_train_epochs=2, lr_sched="cosine", logging_steps=2, sampling_method="randint",
                             mask_ratio=0.1, pretrain=pretrain, pt_type=pt_type, seed=42)
    torch.manual_seed(0)
    model = BaseModel.from_config(cfg)
    train, valid = DS(X[:1024].numpy(), Y[:1024].numpy()), DS(X[1024:].numpy(), Y[1024:].numpy())
    return model, cfg, args, train, valid


@pytest.mark.parametrize("model_name", ["DCNv2", "DeepFM", "xDeepFM"])   # xDeepFM: module path (autograd over our kernels), no fused schedule
@pytest.mark.parametrize("pt_type", ["MFP", "RFD"])
def test_pretrain_loop_runs_and_learns(pt_type, model_name, tmp_path):
    from map_code_b200.trainer import Trainer
    model, cfg, args, train, valid = make(pt_type, True, model_name, tmp_path)
    tr = Trainer(model, cfg, args, train, valid)
    getattr(tr, f"{pt_type}_pretrain")()
    assert tr.global_step == 8 and len(tr.eval_metrics) == 2
    losses = [m[0] for m in tr.eval_metrics]
    assert all(np.isfinite(losses)) and losses[1] < losses[0], losses     # the eval loss goes down between epochs
    path = os.path.join(str(tmp_path), "8.model")
    assert os.path.exists(path)
    sd = torch.load(path)
    assert "embed.embedding.weight" in sd and sd["embed.embedding.weight"].shape == (cfg.input_size, 16)
    # finetune from the checkpoint: tensors that match by name+shape are loaded (models.py:97-112)
    m2, cfg2, args2, train2, valid2 = make(pt_type, False, model_name, tmp_path)
    m2.load_for_finetune(path)
    assert torch.equal(m2.embed.embedding.weight.detach().cpu(), sd["embed.embedding.weight"])


@pytest.mark.parametrize("model_name", ["DCNv2", "DeepFM", "xDeepFM"])
def test_ctr_train_eval_test(model_name, tmp_path):
    from map_code_b200.trainer import Trainer
    model, cfg, args, train, valid = make("MFP", False, model_name, tmp_path)
    tr = Trainer(model, cfg, args, train, valid)
    tr.train()
    assert len(tr.eval_metrics) >= 1 and 0.0 <= tr.eval_metrics[0][0] <= 1.0 and np.isfinite(tr.eval_metrics[0][1])
    auc, ll = tr.test(valid)
    assert 0.0 <= auc <= 1.0


def test_dynamic_mask_api_and_errors(tmp_path):
    from map_code_b200.trainer import Trainer
    model, cfg, args, train, valid = make("MFP", True, "DCNv2", tmp_path)
    model.cuda()
    tr = Trainer(model, cfg, args, train, valid)
    X = torch.from_numpy(train.X[:32])
    out = tr.dynamic_mask({"input_ids": X.clone(), "labels": torch.zeros(32)}, "randint")
    assert set(out.keys()) == {"input_ids", "labels", "masked_index"} and out["masked_index"].shape == (32, 3)
    assert (out["input_ids"].gather(1, out["masked_index"]) == 3).all()
    assert torch.equal(out["labels"].cpu(), X.gather(1, out["masked_index"].cpu()))
    with pytest.raises(NotImplementedError):
        tr.dynamic_mask({"input_ids": X.clone()}, "bogus")
    args.pt_type = "XYZ"
    with pytest.raises(NotImplementedError):
        tr.dynamic_mask({"input_ids": X.clone()}, "randint")
    args.pt_type, args.RFD_replace = "RFD", "Nope"
    with pytest.raises(NotImplementedError):
        tr.dynamic_mask({"input_ids": X.clone()}, "randint")
    from map_code_b200.models import BaseModel
    cfg.model_name = "autoint"   # backbones outside SURVEY §8 raise like an unknown name would
    with pytest.raises(NotImplementedError):
        BaseModel.from_config(cfg)


@pytest.mark.parametrize("pt_type", ["MFP", "RFD"])
def test_train_step_async_matches_train_step(pt_type, tmp_path):
    """the pipelined host-batch API (H2D on a copy stream, loss read one step late) produces the same losses as train_step"""
    from map_code_b200.trainer import Trainer
    losses = {}
    for mode in ("sync", "async"):
        model, cfg, args, train, valid = make(pt_type, True, "DCNv2", tmp_path)
        model.cuda()
        tr = Trainer(model, cfg, args, train, valid)
        tr.fused_step(100, 0)
        batches = [torch.from_numpy(train.X[i * 256:(i + 1) * 256]).pin_memory() for i in range(4)]
        if mode == "sync":
            losses[mode] = [float(tr.train_step(b)[0]) for b in batches + batches]
        else:
            pend, out = None, []
            for b in batches + batches:
                nxt = tr.train_step_async(b)
                if pend is not None:
                    out.append(pend.loss())
                pend = nxt
            out.append(pend.loss())
            losses[mode] = out
        assert tr.global_step == 8
    assert np.allclose(losses["sync"], losses["async"], rtol=1e-4), losses
    assert losses["sync"][-1] < losses["sync"][0]


def test_check_health_reports_a_failed_device_wait(tmp_path):
    """FusedStep.check_health: clean after normal steps; raises when a sort's error word is set (what the kernel does when a
    grid barrier is not met within its spin bound)"""
    from map_code_b200 import _lib
    from map_code_b200.trainer import Trainer
    model, cfg, args, train, valid = make("MFP", True, "DCNv2", tmp_path)
    model.cuda()
    tr = Trainer(model, cfg, args, train, valid)
    eng = tr.fused_step(100, 0)
    tr.train_step(torch.from_numpy(train.X[:256]).cuda())
    eng.check_health()
    plan = eng.tables["embed.embedding.weight"].plan
    off = int(_lib.load().map_dedup_debug_offset(plan.n)) - 64 + 4
    plan.ws[off:off + 4].view(torch.int32).fill_(1)
    with pytest.raises(_lib.MapB200Error):
        eng.check_health()


@pytest.mark.parametrize("model_name,pt_type,pretrain", [("DCNv2", "MFP", True), ("DeepFM", "MFP", True), ("DCNv2", "RFD", True),
                                                         ("DeepFM", "MFP", False)])
def test_ragged_last_batch_keeps_one_optimizer_state(model_name, pt_type, pretrain, tmp_path):
    """n % batch != 0 (reference DataLoader keeps the short last batch, trainer.py:51-58).  The short batch must run through the
    SAME optimizer state, step counter and LR schedule as the full batches: the run over [256, 256, 256, 232] rows is compared
    with a run that feeds the very same four batches through FusedSteps built by hand around one shared state."""
    from map_code_b200.engine import FusedStep
    from map_code_b200.trainer import Trainer
    model, cfg, args, train, valid = make(pt_type, pretrain, model_name, tmp_path)
    args.num_train_epochs = 1
    n = 1000                                    # 3 full batches of 256 + one of 232
    train = DS(train.X[:n], train.Y[:n])
    tr = Trainer(model, cfg, args, train, valid)
    # no shuffling: the batches are rows [0,256), ... in order, so the second run can replay them
    orig = tr.get_dataloader
    tr.get_dataloader = lambda ds, is_training=True: orig(ds, is_training=False) if ds is train else orig(ds, is_training)
    args.per_gpu_eval_batch_size = args.per_gpu_train_batch_size
    if pretrain:
        getattr(tr, f"{pt_type}_pretrain")()
    else:
        tr.train()
    assert tr.global_step == 4
    eng = tr._fused
    assert int(eng.step_counter.item()) == 4 and tr.optimizer is None       # no second optimizer was ever built
    assert list(tr._ragged.keys()) == [232] and tr._ragged[232].exp_avg is eng.exp_avg
    # replay by hand
    m2, cfg2, args2, _, _ = make(pt_type, pretrain, model_name, tmp_path)
    m2.cuda()
    X = torch.from_numpy(train.X).cuda()
    Y = torch.from_numpy(train.Y).cuda()
    kw = dict(mask_ratio=0.1, sampling_method="randint", lr=1e-3, weight_decay=5e-2, sched="cosine", warmup_steps=0, total_steps=4, seed=42,
              x_train=X if pt_type == "RFD" and pretrain else None)
    e_full = FusedStep(m2, batch_size=256, **kw)
    e_tail = FusedStep(m2, batch_size=232, share_state_with=e_full, **kw)
    for i in range(3):
        e_full.step(X[i * 256:(i + 1) * 256], Y[i * 256:(i + 1) * 256])
    e_tail.step(X[768:1000], Y[768:1000])
    torch.cuda.synchronize()
    sd1, sd2 = model.state_dict(), m2.state_dict()
    for k in sd1:
        assert torch.allclose(sd1[k].float(), sd2[k].float(), rtol=1e-5, atol=1e-7), k
    # and the tail step really moved the dense parameters with step-4 bias corrections (not a fresh t=1 optimizer)
    assert float(eng.hyper[0].item()) < 1e-3  # cosine LR near the end of the 4-step schedule, not the base LR


def test_device_batcher_rank_slices_partition_the_global_batch():
    """data-parallel batching: same permutation on every rank, disjoint per-rank slices of every global batch, ragged tail
    dropped; and the module path refuses to read a row-sharded table."""
    from map_code_b200.trainer import DeviceBatcher
    X = torch.arange(1000 * 3).view(1000, 3)
    single = [xb.cpu() for xb, _ in DeviceBatcher(X, None, 128, True, "cuda", seed=5)]
    r0 = [xb.cpu() for xb, _ in DeviceBatcher(X, None, 64, True, "cuda", seed=5, rank=0, world=2)]
    r1 = [xb.cpu() for xb, _ in DeviceBatcher(X, None, 64, True, "cuda", seed=5, rank=1, world=2)]
    assert len(r0) == len(r1) == 1000 // 128 == len(DeviceBatcher(X, None, 64, True, "cuda", seed=5, rank=1, world=2))
    for i in range(len(r0)):
        assert torch.equal(torch.cat([r0[i], r1[i]]), single[i])
    from map_code_b200 import _lib
    from map_code_b200.layers import TableEmbedding
    t = TableEmbedding(10, 4).cuda()
    t.weight._map_sharded = (2, 0, 10)
    with pytest.raises(_lib.MapB200Error):
        t(torch.zeros(3, dtype=torch.int64, device="cuda"))
