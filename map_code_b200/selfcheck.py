"""smoke(): one tiny DCNv2-MFP pretraining step on cuda:0 through the public API (Trainer.train_step -> FusedStep CUDA
graph -> libmap_b200.so kernels incl. the tcgen05 GEMM), checked against the CPU oracle replaying the same Philox streams."""
import torch


def smoke(verbose: bool = True):
    from oracle import map_oracle as O
    from . import synthetic as S
    from .arguments import Config, TrainingArguments
    from .models import BaseModel
    from .trainer import Trainer

    dev = torch.device("cuda", 0)
    F, Dm, H, P, K, B = 39, 16, 64, 32, 25, 256
    sizes = [max(2, s // 500) for s in S.field_sizes("criteo")]
    V = S.vocab_size(sizes)
    X = S.make_ids(sizes, 2048, seed=0)
    fc = S.feat_count(X, V)
    cfgd = dict(model_name="DCNv2", embed_size=Dm, hidden_size=H, num_hidden_layers=3, num_cross_layers=3, hidden_act="relu",
                hidden_dropout_rate=0.0, embed_dropout_rate=0.0, embed_norm=False, layer_norm_eps=1e-12, pt_neg_num=K, proj_size=P,
                input_size=V, num_fields=F, pretrain=True, pt_type="MFP", RFD_replace="Unigram")
    torch.manual_seed(1)
    model = BaseModel.from_config(Config.from_dict(dict(cfgd, feat_count=fc, data_dir=None, seed=42, table_grad_mode="sparse")))
    sd = {k: v.clone() for k, v in model.state_dict().items()}
    model.to(dev)

    class DS:
        def __init__(self, X):
            self.X = X

        def __len__(self):
            return self.X.shape[0]

    targs = TrainingArguments(per_gpu_train_batch_size=B, learning_rate=1e-3, weight_decay=5e-2, lr_sched="cosine", sampling_method="randint",
                              mask_ratio=0.1, pretrain=True, pt_type="MFP", seed=42, optimizer_mode="dense_exact")
    trainer = Trainer(model, model.config, targs, DS(X), DS(X))
    trainer.fused_step(total_steps=10, warmup_steps=0)
    tr = O.OracleTrainer(O.OracleConfig(**cfgd), sd, alias_prob=sd["mfp_criterion.alias.prob"], alias_alias=sd["mfp_criterion.alias.alias"],
                         x_train=X, lr=1e-3, weight_decay=5e-2, mask_ratio=0.1, sampling_method="randint", seed=42,
                         lr_lambda=O.cosine_schedule_lambda(0, 10))
    for s in range(2):
        xb = X[s * B:(s + 1) * B].contiguous()
        outs = trainer.train_step(xb.pin_memory())
        ob = tr.draw(xb)
        ref = tr.forward_backward(ob)
        tr.optimizer_step()
        eng = trainer._fused
        torch.cuda.synchronize()
        assert torch.equal(eng.mi.cpu(), ob["masked_index"]), "masked_index differs from the oracle"
        assert torch.equal(eng.ids_m.cpu(), ob["input_ids"]) and torch.equal(eng.labels.cpu(), ob["labels"]), "mask application differs"
        assert torch.equal(eng.noise.view(ob["noise"].shape).cpu(), ob["noise"]), "alias draw differs"
        got, want = float(outs[0]), float(ref[0])
        assert abs(got - want) < 1e-3 * max(1.0, abs(want)), f"loss {got} vs oracle {want}"
        if verbose:
            print(f"smoke step {s}: loss {got:.6f} (oracle {want:.6f}), acc {int(outs[2])}/{outs[1]} (oracle {int(ref[2])})")
    w = model.embed.embedding.weight.detach().cpu()
    ref_w = tr.params["embed.embedding.weight"].detach()
    upd = (w - sd["embed.embedding.weight"]).double()
    upd_ref = (ref_w - sd["embed.embedding.weight"]).double()
    rel = float((upd - upd_ref).norm() / upd_ref.norm())
    assert rel < 5e-2, f"embedding update differs from the oracle: rel {rel}"
    if verbose:
        print(f"smoke ok: embedding-table update rel err vs oracle {rel:.2e}")
