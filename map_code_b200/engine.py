"""FusedStep — the whole pretraining / finetuning step as one static kernel schedule, captured in a CUDA graph.

Replaces the inner-loop body of Trainer.MFP_pretrain / RFD_pretrain / train (reference code/trainer.py:306-331, 431-455,
122-143): dynamic_mask -> forward -> backward -> AdamW -> LR schedule, with
  * no host synchronisation inside the step (loss / accuracy stay on the device, the Philox offsets and the learning
    rate are derived from a device-side step counter, so the captured graph is replayed verbatim every step);
  * no autograd: the backward pass is scheduled explicitly, ReLU masks / residuals / CrossNet products are fused into
    the GEMM epilogues, `torch.cat` does not exist (producers write into column slices of the `final` buffer);
  * table gradients never materialise as [V, D]: sort-dedup -> compact [U, D] -> row-wise AdamW ("sparse", default) or
    an exact dense sweep ("dense_exact" = the reference's dense transformers.AdamW semantics).
It operates directly on the nn.Parameters of a map_code_b200.models model, so state_dict()/save/load keep working.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional

import torch

from . import _lib, ops

STREAM_MASK_FIELD, STREAM_RFD_REPLACE, STREAM_ALIAS, STREAM_RFD_FIELD2 = 0, 1, 2, 3


def is_no_decay(name: str) -> bool:
    """reference trainer.py:61: no_decay = ["bias", "LayerNorm.weight"] matched as substrings."""
    return any(nd in name for nd in ("bias", "LayerNorm.weight"))


class _Table:
    """One [V, D] table with AdamW state and the dedup plan of its per-step id stream."""

    def __init__(self, name, param, n_ids, weight_decay, device, share: Optional["_Table"] = None):
        self.name, self.p = name, param
        self.V, self.D = param.shape
        # `share`: the same table inside another FusedStep over the same model (ragged last batch): ONE optimizer state
        self.m = torch.zeros_like(param.data) if share is None else share.m
        self.v = torch.zeros_like(param.data) if share is None else share.v
        self.wd = weight_decay
        self.n_ids = n_ids
        self.plan: Optional[ops.DedupPlan] = None  # may be shared between tables fed by the same id stream
        self.grad = torch.empty(n_ids, self.D, dtype=torch.float32, device=device)


class FusedStep:
    def __init__(self, model, *, batch_size: int, mask_ratio: float = 0.1, sampling_method: str = "randint",
                 lr: float = 1e-3, weight_decay: float = 5e-2, betas=(0.9, 0.999), eps: float = 1e-8, sched: str = "cosine",
                 warmup_steps: int = 0, total_steps: int = 1000, seed: int = 42, optimizer_mode: str = "sparse",
                 x_train: Optional[torch.Tensor] = None, idx_low=None, idx_high=None, use_graph: bool = True,
                 row0: int = 0, global_batch: Optional[int] = None, gemm_backend: Optional[str] = None,
                 multi_stream: bool = True, share_state_with: Optional["FusedStep"] = None, field_encoder: Optional[bool] = None):
        """share_state_with: another FusedStep over the SAME model with a different batch size (the Trainer's ragged last batch,
        reference trainer.py:51-58 has drop_last=False).  Parameters, padded parameter storage, AdamW moments of every dense
        parameter and table, the step counter and the schedule state are the other engine's tensors; only the activations, the
        id plans and the captured graph are this engine's own — one optimizer and one scheduler for every batch, like the
        reference."""
        cfg = model.config
        self._share = share_state_with
        if share_state_with is not None and share_state_with.model is not model:
            raise ValueError("share_state_with: engines must wrap the same model")
        self.model, self.cfg = model, cfg
        self.dev = next(model.parameters()).device
        if self.dev.type != "cuda":
            raise _lib.MapB200Error("FusedStep needs the model on a CUDA device (no CPU path)")
        if optimizer_mode not in ("sparse", "dense_exact"):
            raise ValueError(optimizer_mode)
        if sched.lower() not in ("cosine", "const"):
            raise NotImplementedError(sched)  # reference trainer.py:82-83
        self.name = model.model_name.lower()
        if self.name not in ("dcnv2", "dnn", "deepfm"):
            raise NotImplementedError(f"FusedStep: backbone {model.model_name}")
        if str(getattr(cfg, "hidden_act", "relu")).lower() != "relu" or float(getattr(cfg, "hidden_dropout_rate", 0.0)) != 0.0 \
                or float(getattr(cfg, "embed_dropout_rate", 0.0)) != 0.0 or bool(getattr(cfg, "embed_norm", False)):
            # the fused schedule hard-wires the reference defaults (arguments.py:109-124): ReLU towers, no dropout, no embedding
            # LayerNorm; anything else must not silently compute a different model
            raise NotImplementedError("FusedStep supports hidden_act='relu', dropout 0 and embed_norm=False (the reference defaults)")
        self.B, self.F, self.D = batch_size, cfg.num_fields, cfg.embed_size
        self.V = cfg.input_size
        self.in_dim = self.F * self.D
        self.mode = ("MFP" if cfg.pt_type == "MFP" else "RFD") if cfg.pretrain else "CTR"
        if cfg.pretrain and cfg.pt_type not in ("MFP", "RFD"):
            raise NotImplementedError(cfg.pt_type)
        self.L = int(self.F * mask_ratio) if cfg.pretrain else 0  # trainer.py:220
        self.sampling_method = sampling_method
        self.rfd_mode = getattr(cfg, "RFD_replace", "Unigram")
        self.seed, self.row0 = seed, row0
        self.global_batch = global_batch or batch_size
        self.optimizer_mode = optimizer_mode
        self.lr, self.wd, self.betas, self.eps = lr, weight_decay, betas, eps
        self.sched = _lib.SCHED_COSINE if sched.lower() == "cosine" else _lib.SCHED_CONST
        self.warmup_steps, self.total_steps = warmup_steps, total_steps
        self.use_graph = use_graph
        self.gemm_backend = ops._norm_backend(gemm_backend)
        # split-bf16 backend: every GEMM operand is kept as bf16 planes next to its fp32 master (csrc/gemm_bf16s.cu)
        self.use_planes = self.gemm_backend == "bf16s"
        import os as _os
        # CrossNet forward: linear in its operands (no ReLU behind it in MFP / CTR; in RFD its output reaches pred_rfd's ReLU only
        # through one more GEMM) -> terms = 3 keeps every gradient within 4e-4 of fp32 (tests/test_fullshape_gpu.py)
        self.cross_terms = int(_os.environ.get("MAP_B200_CROSS_TERMS", "3"))
        self.x_train, self.idx_low, self.idx_high = x_train, idx_low, idx_high
        # MFP head: evaluate feat_encoder only for the L masked fields of every sample (csrc/fieldenc.cu, K16) instead of all F
        # fields + gather (models.py:73-78).  "hybrid" (default when the kernels take the shape, F >= 8 L and the backbone has no
        # CrossNet tower — see below): forward and
        # weight gradient by field (13x fewer flops at mask_ratio 0.1; the weight gradient leaves the critical path), the input
        # gradient stays a dense tensor-core GEMM over the expanded d_enc, whose epilogues already carry the first backward stage
        # of the towers (measured r02e: a by-field dgrad has to write one [Kd] row per POSITION and fold them afterwards — 80 MB
        # of round trip at C2 — and loses to the dense GEMM).  "full": dgrad by field as well (kept for A/B and for shapes with
        # very small L / F).  MAP_B200_FIELD_ENC=0 / 1 (hybrid) / 2 (full) overrides.
        fe_env = _os.environ.get("MAP_B200_FIELD_ENC")
        if fe_env is not None:
            field_encoder = {"0": False, "1": "hybrid", "2": "full"}[fe_env]
        if field_encoder is True:
            field_encoder = "hybrid"
        if field_encoder not in (None, False, "hybrid", "full"):
            raise ValueError(f"field_encoder={field_encoder!r}")
        fe_ok = self.mode == "MFP" and self.L >= 1 and ops.field_enc_supported(self.F, cfg.proj_size)
        if field_encoder and not fe_ok:
            raise NotImplementedError(f"by-field encoder: needs MFP, F <= 256 and proj_size in (8, 16, 32, 64); got {self.mode}, F={self.F}, P={cfg.proj_size}")
        if field_encoder is None:
            # Measured at HEAD of round 2 (scripts/rounds/r2_38.sh, same box): since the tensor-core GEMMs issue at the pipe's rate the
            # DENSE encoder (13x the flops, but on tensor cores, its weight gradient grouped with the head level's dgrads) beats the
            # SIMT by-field kernels again at C2 (DCNv2: 0.850 vs 0.876 ms / step), while DeepFM at the Avazu shape still gains from
            # them (0.695 vs 0.701 ms)
            field_encoder = "hybrid" if (fe_ok and self.F >= 8 * self.L and self.name != "dcnv2") else False
        self.field_enc = field_encoder
        self._collect_params()
        self._alloc()
        self.graph = None
        self.steps_done = 0
        self.overrides = None
        # Independent branches of the step (MLP tower vs CrossNet, weight gradients vs the dgrad chain, table sorts vs
        # GEMMs) are issued on side streams; under capture they become parallel branches of the CUDA graph.
        import os as _os2
        if _os2.environ.get("MAP_B200_SINGLE_STREAM") == "1":   # A/B switch: the whole step on one stream (no side branches)
            multi_stream = False
        self.multi_stream = multi_stream
        # The table stream ('tab') has NORMAL priority.  While the sort was 13 tiny launches, high priority let them slip in
        # between GEMM waves; with the single-launch sort (persistent CTAs that spin at grid barriers) high priority takes SM
        # slots from the GEMMs for the whole sort.  Measured (r01f, C2 MFP, ms/step, two runs each): normal 0.6209 / 0.6204,
        # high 0.6604 / 0.6607; sorting the NCE ids under the forward pass instead (MAP_B200_NCE_SORT=early): 0.632 / 0.638.
        import os
        tab_prio = int(os.environ.get("MAP_B200_TAB_PRIO", "0"))
        self.streams = ({k: torch.cuda.Stream(device=self.dev, priority=(tab_prio if k == "tab" else 0)) for k in ("tab", "mlp", "dw", "opt")}
                        if multi_stream else {})
        # where the sort of the NCE ids runs on one GPU: "late" = after the NCE kernel, under the backward GEMMs; "early" = right
        # after the embedding sort, under the forward GEMMs (the ids = [labels | noise] are known once the noise is drawn)
        self.nce_sort_early = os.environ.get("MAP_B200_NCE_SORT", "late") == "early"
        self._forked = set()
        self._early = False          # set by _step_body: optimizer work may start inside the backward
        self._tables_done = set()    # tables whose AdamW already ran inside the step

    # ------------------------------------------------------------------------------------------------ setup
    def _collect_params(self):
        m, cfg = self.model, self.cfg
        self.embed_w = m.embed.embedding.weight
        self.cross: List = list(m.cross_net.cross_layers) if self.name == "dcnv2" else []
        if self.name == "dcnv2":
            mlp = m.parallel_dnn.dnn if cfg.num_hidden_layers > 0 else []
        else:
            mlp = m.dnn.dnn
        self.mlp = [mod for i, mod in enumerate(mlp) if i % 3 == 0]
        self.H = cfg.hidden_size if self.mlp else 0
        self.has_fm = self.name == "deepfm"
        # layout of `final`: DCNv2 [cross | mlp], DNN [mlp], DeepFM-pretrain [mlp | lr_fm]
        self.cross_off, self.cross_w = 0, (self.in_dim if self.name == "dcnv2" else 0)
        self.mlp_off = self.cross_w
        self.final_dim = self.cross_w + self.H + (1 if (self.has_fm and cfg.pretrain) else 0)
        self.ld_final = (self.final_dim + 7) // 8 * 8  # TMA rows must be 16-byte multiples (fp32 masters and bf16 planes)
        self.fm_col = self.H  # column of `final` that holds lr_fm in DeepFM pretraining (models.py:224)
        if self.has_fm:
            self.lr_w = m.lr_layer.embed_w.weight
            self.lr_b = m.lr_layer.bias
        if cfg.pretrain:
            self.head0_name = "feat_encoder" if cfg.pt_type == "MFP" else "pred_rfd.0"
        else:
            self.head0_name = "dnn_fc_out" if self.has_fm else "fc_out"

    def _alloc(self):
        B, F, D, dev, cfg = self.B, self.F, self.D, self.dev, self.cfg
        f32 = dict(dtype=torch.float32, device=dev)
        i64 = dict(dtype=torch.int64, device=dev)
        E = lambda *s: torch.empty(*s, **f32)
        self.in_ids = torch.zeros(B, F, **i64)            # static graph inputs
        self.in_labels = torch.zeros(B, **f32)            # CTR labels
        sh = self._share
        self.step_counter = torch.zeros(1, **i64) if sh is None else sh.step_counter
        self.hyper = torch.zeros(8, **f32) if sh is None else sh.hyper
        self.loss = torch.zeros(1, **f32)
        self.acc_count = torch.zeros(1, dtype=torch.int32, device=dev)
        self.stats = torch.zeros(4, **f32)
        self.red_ws = torch.empty(int(_lib.load().map_reduce_workspace_bytes(0)), dtype=torch.uint8, device=dev)
        self.red_ws2 = torch.empty_like(self.red_ws)
        self.ids_m = torch.empty(B, F, **i64)
        if self.mode != "CTR":
            if self.L < 1:
                raise ValueError(f"mask_ratio too small: int(num_fields * mask_ratio) = {self.L} masked fields")
            self.mi = torch.empty(B, self.L, **i64)
        # activations
        self.X0 = E(B, self.in_dim)
        self.final = torch.zeros(B, self.ld_final, **f32)
        nc = len(self.cross)
        self.Xc = [self.X0] + [E(B, self.in_dim) for _ in range(max(nc - 1, 0))]  # inputs of each cross layer
        self.U = [E(B, self.in_dim) for _ in range(nc)]
        nh = len(self.mlp)
        self.Hs = [E(B, self.H) for _ in range(max(nh - 1, 0))]                   # outputs of mlp layers except the last
        self.cross_out = self.final[:, self.cross_off:self.cross_off + self.cross_w] if nc else None
        self.mlp_out = self.final[:, self.mlp_off:self.mlp_off + self.H] if nh else None
        self.final_v = self.final[:, :self.final_dim]
        # bf16 planes of the activations / gradients that GEMMs consume (written by the producing kernel)
        PL = (lambda n, rows, cols: ops.alloc_planes(rows, cols, n, dev)) if self.use_planes else (lambda n, rows, cols: None)
        self.X0p = PL(3, B, self.in_dim)
        self.finalp_full = PL(3, B, self.ld_final)
        fp_ = self.finalp_full
        self.finalp = fp_[:, :, :self.final_dim] if fp_ is not None else None
        self.Xcp = [self.X0p] + [PL(3, B, self.in_dim) for _ in range(max(nc - 1, 0))]
        self.Hsp = [PL(3, B, self.H) for _ in range(max(nh - 1, 0))]
        self.cross_outp = fp_[:, :, self.cross_off:self.cross_off + self.cross_w] if (fp_ is not None and nc) else None
        self.mlp_outp = fp_[:, :, self.mlp_off:self.mlp_off + self.H] if (fp_ is not None and nh) else None
        # RFD head: pred_rfd.2 is Linear(F*P -> F) with F = 39: not a multiple of 4 floats, so its rows of logits / gradients
        # would be unaligned for TMA.  Its weight and bias are re-pointed at the first F rows of zero-padded storage
        # ([Fp, F*P], Fp = roundup8(F)); the padded rows stay zero for ever (their gradients are exactly zero), the
        # nn.Parameter keeps its reference shape, and every GEMM of the head runs on tensor cores at N = K = Fp.
        self._pad_rows: Dict[str, int] = {}
        if self.mode == "RFD":
            l2 = getattr(self.model.pred_rfd, "2")
            self.Fp = (F + 7) // 8 * 8
            if sh is not None:      # the parameters already live in the other engine's padded storage
                self.W2p, self.b2p = sh.W2p, sh.b2p
            else:
                self.W2p = torch.zeros(self.Fp, l2.weight.shape[1], **f32)
                self.W2p[:F].copy_(l2.weight.data)
                l2.weight.data = self.W2p[:F]
                self.b2p = torch.zeros(self.Fp, **f32)
                self.b2p[:F].copy_(l2.bias.data)
                l2.bias.data = self.b2p[:F]
            self._pad_rows = {"pred_rfd.2.weight": self.Fp, "pred_rfd.2.bias": self.Fp}
        # DeepFM pretraining: final_dim = hidden + 1 = 1001 (models.py:211) is not a multiple of 4 floats.  The first head
        # layer's weight [n_out, 1001] is re-pointed at a strided view of zero-padded storage [n_out, 1004] (the padding
        # columns meet zero activations and receive zero gradients, so they stay zero under AdamW); GEMMs and the optimizer
        # work on the padded, TMA-aligned storage, the nn.Parameter keeps its reference shape.
        self._pad_cols: Dict[str, torch.Tensor] = {}
        if cfg.pretrain and self.ld_final != self.final_dim:
            wname = self.head0_name + ".weight"
            mod = self.model.feat_encoder if cfg.pt_type == "MFP" else getattr(self.model.pred_rfd, "0")
            if sh is not None:
                Wp = sh._pad_cols[wname]
            else:
                Wp = torch.zeros(mod.weight.shape[0], self.ld_final, **f32)
                Wp[:, :self.final_dim].copy_(mod.weight.data)
                mod.weight.data = Wp[:, :self.final_dim]
            self._pad_cols[wname] = Wp
        # dense parameter gradients + AdamW state
        self.dense: Dict[str, torch.nn.Parameter] = {}
        for n, p in self.model.named_parameters():
            if p.requires_grad and not hasattr(p, "_map_table_grad"):
                self.dense[n] = p
        # one flat gradient buffer (a single all-reduce in data-parallel runs); per-parameter views keep 16-byte alignment
        # 1-D parameters (biases) come first: their gradients are accumulated by fused column sums in GEMM epilogues
        # (fp32 reductions), so that prefix is zeroed with one small memset at the start of every step.
        # The 2-D weights follow in the REVERSE of the order in which the backward pass completes their gradients (layer 0 of
        # the towers first, the head last): at any point of the backward the finished gradients are a suffix of the buffer, which
        # the data-parallel step all-reduces in a few buckets under the remaining GEMMs (dist.ShardedFusedStep).
        self.dense = dict(sorted(self.dense.items(), key=lambda kv: (0, 0) if kv[1].dim() == 1 else (1, self._layer_depth(kv[0]))))
        offs, tot, padded, n_bias_floats = {}, 0, {}, 0
        for n, p in self.dense.items():
            offs[n] = tot
            if n in self._pad_rows:
                padded[n] = self._pad_rows[n] * (p.numel() // p.shape[0])
            elif n in self._pad_cols:
                padded[n] = self._pad_cols[n].numel()
            else:
                padded[n] = p.numel()
            tot += (padded[n] + 3) // 4 * 4
            if p.dim() == 1:
                n_bias_floats = tot
        self.grad_flat = torch.zeros(tot, **f32)
        self.grad_offsets = {n: (offs[n], offs[n] + (padded[n] + 3) // 4 * 4) for n in self.dense}   # float ranges in grad_flat
        self.bias_grad_flat = self.grad_flat[:n_bias_floats]
        self.grads, self.grads_padded, self.opt_param = {}, {}, {}
        for n, p in self.dense.items():
            flat = self.grad_flat[offs[n]:offs[n] + padded[n]]
            if n in self._pad_cols:  # optimizer sees the padded contiguous storage; .grads[n] is the reference-shaped view
                self.grads_padded[n] = flat.view(p.shape[0], self.ld_final)
                self.grads[n] = self.grads_padded[n][:, :p.shape[1]]
                self.opt_param[n] = self._pad_cols[n]
            else:
                if n in self._pad_rows:
                    self.grads_padded[n] = flat.view(self._pad_rows[n], -1)
                self.grads[n] = flat[:p.numel()].view_as(p.data)
                self.opt_param[n] = p.data
        self.exp_avg = {n: torch.zeros_like(self.opt_param[n]) for n in self.dense} if sh is None else sh.exp_avg
        self.exp_avg_sq = {n: torch.zeros_like(self.opt_param[n]) for n in self.dense} if sh is None else sh.exp_avg_sq
        # bf16 planes of every 2-D dense weight (3 planes: the forward GEMMs that feed a ReLU use all of them); the fused AdamW
        # rewrites them with every update, refresh_weight_planes() after anything else changed the parameters
        self.wplanes: Dict[str, torch.Tensor] = {}
        if self.use_planes:
            self._wplanes_full: Dict[str, torch.Tensor] = {} if sh is None else sh._wplanes_full
            for n in self.dense:
                w = self.opt_param[n]
                if w.dim() == 2 and w.shape[1] % 8 == 0 and w.shape[0] >= 2:
                    if sh is None:   # row-padded weights (pred_rfd.2): planes of the whole padded storage, AdamW sees the real rows
                        rows = self._pad_rows.get(n, w.shape[0])
                        self._wplanes_full[n] = torch.zeros(3, rows, w.shape[1], dtype=torch.bfloat16, device=dev)
                    self.wplanes[n] = self._wplanes_full[n][:, :w.shape[0]]
        entries = [(self.opt_param[n], self.grad_flat[offs[n]:offs[n] + self.opt_param[n].numel()].view_as(self.opt_param[n]),
                    self.exp_avg[n], self.exp_avg_sq[n], 0.0 if is_no_decay(n) else self.wd, None, self.wplanes.get(n)) for n in self.dense]
        self.adam_table, self.adam_n, self.adam_max = ops.make_adamw_tensor_list(entries, dev)
        # The same update in two launches for the fused single-GPU schedule: the weights of tower layers >= 1 and of the head have
        # their gradients two GEMM levels before the backward ends (nothing reads them again in this step), so their AdamW (most of
        # the 7 x 4 bytes per parameter) runs on the 'opt' stream under the last two levels, whose launches leave SMs idle; only
        # layer 0 and the 1-D parameters (column sums that finish with the last epilogues) are left for the end of the step.
        names = list(self.dense)
        early = [i for i, n in enumerate(names) if self.dense[n].dim() == 2 and self._layer_depth(n) >= 1]
        late = [i for i in range(len(names)) if i not in set(early)]
        self.adam_early = ops.make_adamw_tensor_list([entries[i] for i in early], dev) if early and late else None
        self.adam_late = ops.make_adamw_tensor_list([entries[i] for i in late], dev) if early and late else None
        self._adam_early_done = False
        self._ev_hyper = None
        self._param_versions = None
        self.refresh_weight_planes()
        # backward scratch
        md = max(self.in_dim, self.H, 1)
        self.dZ = [E(B, self.H) for _ in range(nh)]          # gradient at each MLP pre-activation (kept: dW reads it later)
        self.dUs = [E(B, self.in_dim) for _ in range(nc)]    # dU of each cross layer
        self.dZp = [PL(2, B, self.H) for _ in range(nh)]
        self.dUsp = [PL(2, B, self.in_dim) for _ in range(nc)]
        self.Gc = [E(B, self.in_dim) for _ in range(nc + 1)] # d(loss)/d(X_i) through the cross chain
        self.dX0_acc = E(B, self.in_dim) if nc else None
        self.dX0_mlp = E(B, self.in_dim) if nh else None
        self.dE = E(B, self.in_dim)
        self.colsum_ws = torch.empty(int(_lib.load().map_colsum_workspace_bytes(B, max(self.ld_final, F * cfg.proj_size if cfg.pretrain else 1, md))),
                                     dtype=torch.uint8, device=dev)
        # tables
        self.tables: Dict[str, _Table] = {}
        self._make_embed_table()
        if self.has_fm:
            self._make_lr_table()
            self.d_lrfm = E(B, 1)
            self.d_w_occ = E(B * F)
            self.lr_fm = E(B, 1)
        # heads
        if self.mode == "MFP":
            P, K, L = cfg.proj_size, cfg.pt_neg_num, self.L
            N = B * L
            self.N, self.P, self.K = N, P, K
            crit = self.model.mfp_criterion
            self.labels = torch.empty(B, L, **i64)
            self.enc = self.d_enc = self.d_encp = self.dxpos = None
            if self.field_enc:   # by-field encoder: the [B, F*P] output never exists; positions bucketed by field
                self.fe_perm = torch.zeros(N, dtype=torch.int32, device=dev)
                self.fe_fstart = torch.zeros(F + 1, dtype=torch.int32, device=dev)
            else:
                self.enc = E(B, F * P)
            if self.field_enc == "full":   # one gradient row per position, folded per sample afterwards
                self.dxpos = E(N, self.ld_final)
            else:
                self.d_enc = torch.zeros(B, F * P, **f32)
                self.d_encp = PL(2, B, F * P)
            self.sel = E(max(N, 1), P)
            self.noise = torch.empty(max(N, 1), K, **i64)
            self.logits = E(max(N, 1), K + 1)
            self.ids_all = torch.empty(max(N, 1), K + 1, **i64)
            self.loss_pos = E(max(N, 1))
            self.dz = E(max(N, 1), K + 1)
            self.d_sel = E(max(N, 1), P)
            self._make_nce_tables()
            self.norm_term = float(crit.norm_term)
            self.loss_type = crit.loss_type
            if crit.reduction != "elementwise_mean":
                raise NotImplementedError("FusedStep supports the default reduction='elementwise_mean'")
        elif self.mode == "RFD":
            P = cfg.proj_size
            self.labels = E(B, F)
            self.rfd_h = E(B, F * P)
            self.rfd_logits = E(B, F)
            self.rfd_logits_p = E(B, self.Fp)
            self.d_logits = E(B, F)
            self.d_logits_p = torch.zeros(B, self.Fp, **f32)
            self.d_h = E(B, F * P)
            self.rfd_hp, self.d_hp, self.d_logits_pp = PL(3, B, F * P), PL(2, B, F * P), PL(2, B, self.Fp)
        else:
            self.ctr_logits = E(B, 1)
            self.d_logits = E(B, 1)

    def _make_embed_table(self):
        name = "embed.embedding.weight"
        t = _Table(name, self.embed_w, self.B * self.F, 0.0 if is_no_decay(name) else self.wd, self.dev, self._shared_table(name))
        t.plan = ops.DedupPlan(self.B * self.F, self.V, self.dev)
        self.tables[name] = t

    def _shared_table(self, name):
        return None if self._share is None else self._share.tables[name]

    def _make_lr_table(self):
        """DeepFM's first-order table lr_layer.embed_w [V,1] is fed by the same id stream as the embedding: it shares its plan."""
        name = "lr_layer.embed_w.weight"
        t = _Table(name, self.lr_w, self.B * self.F, 0.0 if is_no_decay(name) else self.wd, self.dev, self._shared_table(name))
        t.plan = self.tables["embed.embedding.weight"].plan
        self.tables[name] = t

    def _make_nce_tables(self):
        crit = self.model.mfp_criterion
        n_occ = max(self.N, 1) * (self.K + 1)
        plan = ops.DedupPlan(n_occ, self.V, self.dev)
        te = _Table("mfp_criterion.emb.weight", crit.emb.weight, n_occ, 0.0 if is_no_decay("mfp_criterion.emb.weight") else self.wd, self.dev,
                    self._shared_table("mfp_criterion.emb.weight"))
        tb = _Table("mfp_criterion.bias.weight", crit.bias.weight, n_occ, 0.0 if is_no_decay("mfp_criterion.bias.weight") else self.wd, self.dev,
                    self._shared_table("mfp_criterion.bias.weight"))
        te.plan = tb.plan = plan
        self.tables[te.name], self.tables[tb.name] = te, tb

    # ------------------------------------------------------------------------------------------------ helpers
    @staticmethod
    def _layer_depth(name: str) -> int:
        """position of a weight's gradient in the backward pass, counted from the END (tower layer i -> i; heads -> 99)"""
        parts = name.split(".")
        for prefix, div in (("cross_net.cross_layers.", 1), ("parallel_dnn.dnn.", 3), ("dnn.dnn.", 3)):
            if name.startswith(prefix):
                return int(parts[prefix.count(".")]) // div
        return 99

    def refresh_weight_planes(self):
        """Re-derives the bf16 planes of the dense weights from their fp32 masters.  The fused AdamW keeps them current; call this
        after anything else wrote the parameters (load_state_dict is detected by step(); raw .data writes are not)."""
        if self.use_planes:
            for n, pl in self.wplanes.items():
                ops.split_planes(self.opt_param[n], pl)
        self._param_versions = {n: p._version for n, p in self.dense.items()}

    def _wp(self, name: str, cols=None):
        """bf16 planes of a dense weight (optionally a column slice), None when the backend does not use planes"""
        if not self.use_planes or name not in self.wplanes:
            return None
        pl = self._wplanes_full[name]
        return pl if cols is None else pl[:, :, cols[0]:cols[1]]

    def _gemm(self, *a, **k):
        return ops.gemm(*a, backend=self.gemm_backend, **k)

    def _gemm_group(self, problems):
        """independent GEMMs of one level of the step's dependency graph -> one grouped launch (ops.gemm_group)"""
        problems = [pr for pr in problems if pr is not None]
        if problems:
            ops.gemm_group(problems, backend=self.gemm_backend)

    def _wgrad_problem(self, dZ, X_in, layer_name, M, N, K, dZp=None, X_inp=None):
        """dW = dZ^T X as a problem of a grouped launch (the bias gradient comes from the colsum_out of the GEMM that wrote dZ)"""
        wname = layer_name + ".weight"
        if wname in self._pad_cols:  # column-padded head weight: compute on the aligned [N, ld_final] storage
            return dict(A=dZ, B=self.final, C_out=self.grads_padded[wname], M=N, N=self.ld_final, K=M, trans_a=True, trans_b=True,
                        Ap=dZp, Bp=self.finalp_full)
        return dict(A=dZ, B=X_in, C_out=self.grads[wname], M=N, N=K, K=M, trans_a=True, trans_b=True, Ap=dZp, Bp=X_inp)

    def _on(self, name):
        """context: issue on side stream `name` (or stay on the main stream when multi_stream is off)"""
        import contextlib
        return torch.cuda.stream(self.streams[name]) if self.multi_stream else contextlib.nullcontext()

    def _fork(self, name):
        """side stream `name` waits for everything issued so far on the CURRENT stream"""
        if self.multi_stream:
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream())
            self.streams[name].wait_event(ev)
            self._forked.add(name)

    def _join(self, name):
        """the CURRENT stream waits for everything issued so far on side stream `name` (no-op for a stream that was not forked
        in this step: under graph capture, waiting on an uncaptured stream is an error)"""
        if self.multi_stream and name in self._forked:
            ev = torch.cuda.Event()
            ev.record(self.streams[name])
            torch.cuda.current_stream().wait_event(ev)

    def _wgrad(self, dZ, X_in, layer_name, M, N, K, bias_done=False, dZp=None, X_inp=None):
        """dW = dZ^T X, db = colsum(dZ) for a Linear with input X_in [M,K] and pre-activation gradient dZ [M,N].
        Issued on the 'dw' stream: weight gradients are off the critical path (only the optimizer waits for them).
        bias_done: the GEMM that produced dZ already accumulated its column sums into the bias gradient (colsum_out)."""
        self._fork("dw")
        with self._on("dw"):
            wname = layer_name + ".weight"
            if wname in self._pad_cols:  # column-padded head weight: compute on the aligned [N, ld_final] storage
                self._gemm(dZ, self.final, self.grads_padded[wname], N, self.ld_final, M, trans_a=True, trans_b=True, Ap=dZp, Bp=self.finalp_full)
            else:
                self._gemm(dZ, X_in, self.grads[wname], N, K, M, trans_a=True, trans_b=True, Ap=dZp, Bp=X_inp)
            if not bias_done:
                ops.colsum(dZ, out=self.grads[layer_name + ".bias"], ws=self.colsum_ws)

    # ------------------------------------------------------------------------------------------------ the step
    def _draw_and_mask(self):
        B, F, L = self.B, self.F, self.L
        if self.mode == "CTR":
            return self.in_ids
        ov = self.overrides
        if ov is not None:  # parity tests feed the reference's index tensors instead of drawing
            if "masked_index" in ov:
                self.mi.copy_(ov["masked_index"])
            if "input_ids_masked" in ov:  # the reference's own dynamic_mask output
                self.ids_m.copy_(ov["input_ids_masked"])
                self.labels.copy_(ov["labels"])
                return self.ids_m
            if self.mode == "MFP":
                ops.mfp_mask_apply(self.in_ids, self.mi, 3, ids_out=self.ids_m, labels=self.labels)
                return self.ids_m
        ops.mask_index(B, L, F, self.sampling_method, self.seed, STREAM_MASK_FIELD, row0=self.row0, out=self.mi, step_dev=self.step_counter)
        if self.mode == "MFP":
            ops.mfp_mask_apply(self.in_ids, self.mi, 3, ids_out=self.ids_m, labels=self.labels)
        else:
            ops.rfd_replace(self.in_ids, self.mi, self.rfd_mode, self.seed, STREAM_RFD_REPLACE, STREAM_RFD_FIELD2, x_train=self.x_train,
                            idx_low=self.idx_low, idx_high=self.idx_high, input_size=self.V, row0=self.row0, ids_out=self.ids_m,
                            labels=self.labels, step_dev=self.step_counter)
        return self.ids_m

    def _embed_lookup(self, ids):
        """ids [B,F] -> X0 [B, F*D] (K1).  Table-side work that only needs the ids starts now on the 'tab' stream: the
        embedding dedup sort (K2a) and, for MFP, the alias draw of the NCE noise (K5)."""
        self._fork("tab")
        with self._on("tab"):
            self._draw_noise()   # first: the NCE head waits for the noise, nothing waits for the sort until the backward
            if self._early:      # every Philox consumer of this step has been issued: the step counter may advance
                self._hyper_step()
                if self.multi_stream:
                    self._ev_hyper = torch.cuda.Event()
                    self._ev_hyper.record(self.streams["tab"])
            self.tables["embed.embedding.weight"].plan.run(ids.view(-1))
            if self.nce_sort_early and self.mode == "MFP":
                ops.nce_ids_concat(self.labels.view(-1), self.noise, out=self.ids_all)
                self.tables["mfp_criterion.emb.weight"].plan.run(self.ids_all.view(-1))
        ops.emb_gather(self.embed_w.data, ids, out=self.X0)
        self._split(self.X0, self.X0p)

    def _split(self, t, planes):
        if planes is not None:
            ops.split_planes(t, planes)

    def _draw_noise(self):
        if self.mode != "MFP":
            return
        if self.field_enc:   # (needs only the mask: rides on the table stream with the noise draw, far ahead of the head)
            ops.field_bucket(self.mi, self.F, self.fe_perm, self.fe_fstart)
        elif self.P % 4 != 0:
            self.d_enc.zero_()   # (memset for the atomic form of the slice scatter; proj sizes that are multiples of 4 need none)
        crit = self.model.mfp_criterion
        if self.overrides is not None and "noise" in self.overrides:
            self.noise.copy_(self.overrides["noise"].reshape(self.N, self.K))
        else:
            ops.alias_draw(crit.alias.prob, crit.alias.alias, self.seed, STREAM_ALIAS, self.N * self.K,
                           elem0=self.row0 * self.L * self.K, out=self.noise.view(-1), step_dev=self.step_counter)

    def _embed_backward(self):
        """dE [B, F*D] -> compact table gradient: segmented row sum over the (already sorted) ids (K2b)."""
        self._join("tab")
        t = self.tables["embed.embedding.weight"]
        t.plan.reduce_rows(self.dE, self.D, out=t.grad)
        if self.has_fm:
            tl = self.tables["lr_layer.embed_w.weight"]
            tl.plan.reduce_rows(self.d_w_occ, 1, out=tl.grad)

    def _forward_backbone(self, ids):
        B, in_dim, H = self.B, self.in_dim, self.H
        self._embed_lookup(ids)
        if self.has_fm:  # lr_fm = LR(ids) + FM2(E): written straight into its column of `final` (pretrain) or kept apart (CTR)
            if self.cfg.pretrain:
                self._lr_fm_forward(ids, self.final[:, self.fm_col:], self.ld_final)
                if self.use_planes:   # the lr_fm column (+ its zero padding) of `final` as planes
                    self._split(self.final[:, self.fm_col:self.fm_col + 4], self.finalp_full[:, :, self.fm_col:self.fm_col + 4])
            else:
                self._lr_fm_forward(ids, self.lr_fm, 1)
        nh, nc = len(self.mlp), len(self.cross)
        # CrossNet layer i (layers.py:197-201, product fused in the epilogue) and MLP layer i (layers.py:187-188, ReLU fused) only
        # depend on layer i-1 of their own tower: one grouped launch per depth
        pref_c = "cross_net.cross_layers"
        pref_m = "parallel_dnn.dnn" if self.name == "dcnv2" else "dnn.dnn"
        # GEMMs whose output goes through a ReLU run with terms = 6 (fp32-level products): the gradient is a discontinuous
        # function of those pre-activations (csrc/gemm_bf16s.cu); everything linear runs with terms = 3.
        x, k, xp = self.X0, in_dim, self.X0p
        for i in range(max(nh, nc)):
            probs = []
            if i < nc:
                out, outp = (self.cross_out, self.cross_outp) if i == nc - 1 else (self.Xc[i + 1], self.Xcp[i + 1])
                layer = self.cross[i]
                probs.append(dict(A=self.Xc[i], B=layer.weight.data, C_out=out, M=B, N=in_dim, K=in_dim, epilogue=_lib.EPI_CROSS,
                                  bias=layer.bias.data, aux0=self.Xc[i], aux1=self.X0, aux_out=self.U[i],
                                  Ap=self.Xcp[i], Bp=self._wp(f"{pref_c}.{i}.weight"), Cp=outp, terms=self.cross_terms))
            if i < nh:
                out, outp = (self.mlp_out, self.mlp_outp) if i == nh - 1 else (self.Hs[i], self.Hsp[i])
                layer = self.mlp[i]
                probs.append(dict(A=x, B=layer.weight.data, C_out=out, M=B, N=H, K=k, epilogue=_lib.EPI_BIAS_RELU, bias=layer.bias.data,
                                  Ap=xp, Bp=self._wp(f"{pref_m}.{3 * i}.weight"), Cp=outp, terms=6))
                x, k, xp = out, H, outp
            self._gemm_group(probs)

    def _lr_fm_forward(self, ids, out, ld_out):
        ops.fm_lr_fwd(self.X0.view(self.B, self.F, self.D), ids, self.lr_w.data.view(-1), self.lr_b.data, out=out, ld_out=ld_out)

    def _backward_backbone(self, head_W, dHead, n_head, head_wgrad=None, head_Wname=None, dHeadp=None, from_fold=False):
        """head_W [n_head, final_dim] is the weight of the first head layer, dHead [B, n_head] the gradient at its
        pre-activation, head_wgrad the (optional) weight-gradient problem of that layer: it consumes dHead like the two
        tower dgrads and goes out in the same grouped launch.  Propagates into the towers and the embedding table.
        CrossNet (autograd of layers.py:200, every elementwise stage inside a GEMM epilogue):
          G_i = d(loss)/d(X_i);  dU_i = G_{i+1} * X0;  dX0 += G_{i+1} * U_i;  G_i = G_{i+1} + dU_i W_i;  db_i = colsum(dU_i)
        the GEMM that produces G_{i+1} also writes dU_i, accumulates dX0 and db_i (MAP_EPI_CROSS_BWD); the last one adds
        everything into dE (MAP_EPI_ADD3).  MLP: the gradient through each ReLU and the bias gradient (column sums of dZ) are
        fused into the dgrad GEMM's epilogue.  One grouped launch per depth: {dgrad, wgrad} x {CrossNet, MLP}.
        from_fold: the head level (dZ / dU / G / dX0 of the last tower layers, their bias gradients, d_lrfm) has already been
        produced by ops.head_bwd_fold (by-field MFP encoder); head_W / dHead are not used."""
        B, in_dim, H = self.B, self.in_dim, self.H
        nc, nh = len(self.cross), len(self.mlp)
        pref_c = "cross_net.cross_layers"
        pref_m = "parallel_dnn.dnn" if self.name == "dcnv2" else "dnn.dnn"
        probs = [head_wgrad]
        hw = (lambda lo, hi: self._wp(head_Wname, (lo, hi))) if head_Wname is not None else (lambda lo, hi: None)
        if from_fold:
            nh_, nc_ = 0, 0
        else:
            nh_, nc_ = nh, nc
        if nh_:
            probs.append(dict(A=dHead, B=head_W[:, self.mlp_off:self.mlp_off + H], C_out=self.dZ[nh - 1], M=B, N=H, K=n_head, trans_b=True,
                              epilogue=_lib.EPI_MUL_RELUMASK, aux0=self.mlp_out, colsum_out=self.grads[f"{pref_m}.{3 * (nh - 1)}.bias"],
                              Ap=dHeadp, Bp=hw(self.mlp_off, self.mlp_off + H), Cp=self.dZp[nh - 1]))
        if nc_:
            probs.append(dict(A=dHead, B=head_W[:, self.cross_off:self.cross_off + in_dim], C_out=self.dUs[nc - 1], M=B, N=in_dim, K=n_head,
                              trans_b=True, epilogue=_lib.EPI_CROSS_BWD, aux0=None, aux1=self.X0, aux2=self.U[nc - 1], aux_out=self.Gc[nc],
                              acc_out=self.dX0_acc, acc_accumulate=False, colsum_out=self.grads[f"{pref_c}.{nc - 1}.bias"],
                              Ap=dHeadp, Bp=hw(self.cross_off, self.cross_off + in_dim), Cp=self.dUsp[nc - 1]))
        self._gemm_group(probs)
        if self.has_fm and self.cfg.pretrain and not from_fold:
            # d(lr_fm) = dHead . W[:, fm_col]: here, with the head level, because the head weight is updated before the backward ends
            # (_early_dense_adamw); its consumers (FM backward, first-order table) run after dE below
            self._gemm(dHead, head_W[:, self.fm_col:self.fm_col + 1], self.d_lrfm, B, 1, n_head, trans_b=True)
        last = []   # the final level: dE needs dX0_mlp, which the last MLP dgrad of the loop below produces
        for s_ in range(max(nc, nh)):
            ic, im = nc - 1 - s_, nh - 1 - s_
            probs = []
            if ic >= 0:
                layer = self.cross[ic]
                wname = f"{pref_c}.{ic}.weight"
                wg = self._wgrad_problem(self.dUs[ic], self.Xc[ic], f"{pref_c}.{ic}", B, in_dim, in_dim, dZp=self.dUsp[ic], X_inp=self.Xcp[ic])
                if ic > 0:
                    probs.append(wg)
                    probs.append(dict(A=self.dUs[ic], B=layer.weight.data, C_out=self.dUs[ic - 1], M=B, N=in_dim, K=in_dim, trans_b=True,
                                      epilogue=_lib.EPI_CROSS_BWD, aux0=self.Gc[ic + 1], aux1=self.X0, aux2=self.U[ic - 1], aux_out=self.Gc[ic],
                                      acc_out=self.dX0_acc, acc_accumulate=True, colsum_out=self.grads[f"{pref_c}.{ic - 1}.bias"],
                                      Ap=self.dUsp[ic], Bp=self._wp(wname), Cp=self.dUsp[ic - 1]))
                else:
                    last.append(wg)
                    last.append(dict(A=self.dUs[0], B=layer.weight.data, C_out=self.dE, M=B, N=in_dim, K=in_dim, trans_b=True,
                                     epilogue=_lib.EPI_ADD3, aux0=self.Gc[1], aux1=self.dX0_acc, aux2=self.dX0_mlp if nh else None,
                                     Ap=self.dUsp[0], Bp=self._wp(wname)))
            if im >= 0:
                layer = self.mlp[im]
                dZ, dZp = self.dZ[im], self.dZp[im]
                x_in, x_inp = (self.X0, self.X0p) if im == 0 else (self.Hs[im - 1], self.Hsp[im - 1])
                k_in = in_dim if im == 0 else H
                wname = f"{pref_m}.{3 * im}.weight"
                probs.append(self._wgrad_problem(dZ, x_in, f"{pref_m}.{3 * im}", B, H, k_in, dZp=dZp, X_inp=x_inp))
                if im > 0:
                    probs.append(dict(A=dZ, B=layer.weight.data, C_out=self.dZ[im - 1], M=B, N=k_in, K=H, trans_b=True,
                                      epilogue=_lib.EPI_MUL_RELUMASK, aux0=self.Hs[im - 1], colsum_out=self.grads[f"{pref_m}.{3 * (im - 1)}.bias"],
                                      Ap=dZp, Bp=self._wp(wname), Cp=self.dZp[im - 1]))
                else:
                    probs.append(dict(A=dZ, B=layer.weight.data, C_out=self.dX0_mlp, M=B, N=k_in, K=H, trans_b=True, Ap=dZp, Bp=self._wp(wname)))
            self._gemm_group(probs)
            if s_ == max(nc, nh) - 2:
                self._early_dense_adamw()
        if nc:
            self._gemm_group(last)
        else:
            ops.copy2d(self.dX0_mlp, self.dE)
        if self.has_fm:  # d(lr_fm) flows into the embeddings (FM term) and into the first-order table + its bias
            ops.fm_lr_bwd(self.X0.view(B, self.F, self.D), self.d_lrfm, 1, self.dE.view(B, self.F, self.D), self.d_w_occ, accumulate=True)
            ops.reduce_sum(self.d_lrfm.view(-1), 1.0, out=self.grads["lr_layer.bias"], ws=self.red_ws2)
        self._embed_backward()

    def _nce_core(self):
        """sel [N,P], labels, noise -> loss, logits, acc, d_sel and the compact gradients of the two NCE tables (K6/K7)."""
        P, K, L = self.P, self.K, self.L
        crit = self.model.mfp_criterion
        self.acc_count.zero_()
        n_global = self.global_batch * L
        ops.nce_fwd(self.sel, self.labels.view(-1), self.noise, crit.emb.weight.data, crit.bias.weight.data.view(-1), crit.logprob_noise,
                    self.norm_term, self.loss_type, grad_scale=1.0 / n_global, logits=self.logits,
                    ids_out=None if self.nce_sort_early else self.ids_all, want_ids=False,
                    loss_pos=self.loss_pos, dz=self.dz, d_input=self.d_sel, acc_count=self.acc_count)
        # table gradients on the 'tab' stream: sort the (N, K+1) ids, reduce dz * input rows per unique id.  (On one GPU the sort
        # stays HERE, under the backward GEMMs: the SMs are saturated either way, and sorting under the forward pass instead
        # delayed the NCE kernel — r01f timeline.  The sharded step sorts early because its owners' pulls hang on the keys.)
        te, tb = self.tables["mfp_criterion.emb.weight"], self.tables["mfp_criterion.bias.weight"]
        self._fork("tab")
        with self._on("tab"):
            ops.reduce_sum(self.loss_pos, 1.0 / n_global, out=self.loss, ws=self.red_ws)
            if not self.nce_sort_early:
                te.plan.run(self.ids_all.view(-1))
            te.plan.reduce_rows(self.sel, P, scale=self.dz.view(-1), group=K + 1, out=te.grad, scalar_out=tb.grad)
            if self._early and self.optimizer_mode == "sparse":
                # nothing reads the two NCE tables again in this step: their row-wise AdamW runs here, under the backward GEMMs
                for t in (te, tb):
                    ops.adamw_sparse_rows(t.p.data, t.m, t.v, t.plan, t.grad, self.hyper, t.wd)
                    self._tables_done.add(t.name)

    def _head_mfp_by_field(self):
        """models.py:73-78 and its autograd for the L masked fields only (csrc/fieldenc.cu): forward on the bucketed positions;
        the encoder's weight / bias gradients by field on the 'dw' stream; the input gradient either as dense GEMMs over the
        expanded d_enc ("hybrid") or by field: dgrad per position -> fold over the L rows of a sample fused with the first
        backward stage of the towers ("full")."""
        B, F, P, L = self.B, self.F, self.P, self.L
        m = self.model
        wname = "feat_encoder.weight"
        # aligned [F*P, ld_final] storage of the weight (DeepFM: zero-padded columns, which meet the zero padding of `final`)
        W = self._pad_cols[wname] if wname in self._pad_cols else m.feat_encoder.weight.data
        Kd = self.ld_final
        self._join("tab")   # noise and the field buckets come from the 'tab' stream
        ops.field_enc_fwd(self.final, Kd, W, m.feat_encoder.bias.data, self.fe_perm, self.fe_fstart, L, F, P, out=self.sel)
        self._nce_core()
        if self.field_enc == "hybrid":   # input gradient: dense GEMMs over the expanded d_enc (their epilogues start the towers' backward)
            ops.expand_slices(self.d_sel, self.mi, F, P, self.d_enc, self.d_encp)
            self._field_wgrad(W, Kd)
            return self._backward_backbone(m.feat_encoder.weight.data, self.d_enc, F * P, head_Wname=wname, dHeadp=self.d_encp)
        ops.field_enc_dgrad(self.d_sel, W, Kd, self.fe_perm, self.fe_fstart, F, P, self.dxpos)
        nc, nh = len(self.cross), len(self.mlp)
        pref_m = "parallel_dnn.dnn" if self.name == "dcnv2" else "dnn.dnn"
        cross = relu = scalar = None
        if nc:
            cross = dict(col0=self.cross_off, width=self.in_dim, x0=self.X0, u=self.U[nc - 1], g_out=self.Gc[nc], du_out=self.dUs[nc - 1],
                         dx0_out=self.dX0_acc, du_planes=self.dUsp[nc - 1], bias_grad=self.grads[f"cross_net.cross_layers.{nc - 1}.bias"])
        if nh:
            relu = dict(col0=self.mlp_off, width=self.H, y=self.mlp_out, dz_out=self.dZ[nh - 1], dz_planes=self.dZp[nh - 1],
                        bias_grad=self.grads[f"{pref_m}.{3 * (nh - 1)}.bias"])
        if self.has_fm:
            scalar = dict(col=self.fm_col, out=self.d_lrfm)
        ops.head_bwd_fold(self.dxpos, B, L, Kd, cross=cross, relu=relu, scalar=scalar)
        self._field_wgrad(W, Kd)   # issued after the fold: the SIMT weight gradient then shares the SMs with the next GEMM level, not with the dgrad
        self._backward_backbone(None, None, F * P, from_fold=True)

    def _field_wgrad(self, W, Kd):
        """encoder weight + bias gradients of the by-field head on the 'dw' stream (only the optimizer waits for them)"""
        wname = "feat_encoder.weight"
        self._fork("dw")
        with self._on("dw"):
            dW = self.grads_padded[wname] if wname in self._pad_cols else self.grads[wname]
            ops.field_enc_wgrad(self.d_sel, self.final, Kd, self.fe_perm, self.fe_fstart, self.L, self.F, self.P, dW,
                                dbias=self.grads["feat_encoder.bias"])

    def _head_mfp(self):
        if self.field_enc:
            return self._head_mfp_by_field()
        cfg, B, F, P, K, N, L = self.cfg, self.B, self.F, self.P, self.K, self.N, self.L
        m = self.model
        enc_W, enc_b = m.feat_encoder.weight.data, m.feat_encoder.bias.data
        crit = m.mfp_criterion
        self._gemm_group([dict(A=self.final_v, B=enc_W, C_out=self.enc, M=B, N=F * P, K=self.final_dim, epilogue=_lib.EPI_BIAS, bias=enc_b,
                               Ap=self.finalp, Bp=self._wp("feat_encoder.weight", (0, self.final_dim)))])   # models.py:74
        ops.gather_slices(self.enc, self.mi, F, P, out=self.sel)                                                   # models.py:75
        self._join("tab")  # noise drawn on the 'tab' stream
        self._nce_core()
        # ---- backward of the encoder (d_enc was zeroed on the 'tab' stream at the start of the step)
        if P % 4 == 0:   # dense, deterministic pass: writes every element of d_enc (and its planes) once; no memset, no atomics
            ops.expand_slices(self.d_sel, self.mi, F, P, self.d_enc, self.d_encp)
        else:
            ops.scatter_add_slices(self.d_sel, self.mi, F, P, self.d_enc)
            self._split(self.d_enc, self.d_encp)
        self._fork("dw")
        with self._on("dw"):   # bias gradient of the encoder: off the critical path
            ops.colsum(self.d_enc, out=self.grads["feat_encoder.bias"], ws=self.colsum_ws)
        self._backward_backbone(enc_W, self.d_enc, F * P, head_Wname="feat_encoder.weight", dHeadp=self.d_encp,
                                head_wgrad=self._wgrad_problem(self.d_enc, self.final_v, "feat_encoder", B, F * P, self.final_dim,
                                                               dZp=self.d_encp, X_inp=self.finalp))

    def _head_rfd(self):
        cfg, B, F, Fp = self.cfg, self.B, self.F, self.Fp
        P = cfg.proj_size
        l0 = getattr(self.model.pred_rfd, "0")
        w2p = self._wp("pred_rfd.2.weight")
        self._gemm_group([dict(A=self.final_v, B=l0.weight.data, C_out=self.rfd_h, M=B, N=F * P, K=self.final_dim, epilogue=_lib.EPI_BIAS_RELU,
                               bias=l0.bias.data, Ap=self.finalp, Bp=self._wp("pred_rfd.0.weight", (0, self.final_dim)), Cp=self.rfd_hp,
                               terms=6)])   # feeds a ReLU
        self._gemm(self.rfd_h, self.W2p, self.rfd_logits_p, B, Fp, F * P, epilogue=_lib.EPI_BIAS, bias=self.b2p,
                   Ap=self.rfd_hp, Bp=w2p)   # models.py:80
        ops.copy2d(self.rfd_logits_p[:, :F], self.rfd_logits)
        ops.bce_logits(self.rfd_logits.view(-1), self.labels.view(-1), stats=self.stats, dlogits=self.d_logits.view(-1), ws=self.red_ws)
        if self.global_batch != B:  # mean over the GLOBAL batch
            ops.scale_by_scalar(self.d_logits.view(-1), self._ratio(), out=self.d_logits.view(-1))
        ops.copy2d(self.d_logits, self.d_logits_p[:, :F])      # padded columns stay 0
        self._split(self.d_logits_p, self.d_logits_pp)
        # backward
        self._fork("dw")
        with self._on("dw"):
            ops.colsum(self.d_logits_p, out=self.grads_padded["pred_rfd.2.bias"].view(-1), ws=self.colsum_ws)
        self._gemm_group([
            dict(A=self.d_logits_p, B=self.rfd_h, C_out=self.grads_padded["pred_rfd.2.weight"], M=Fp, N=F * P, K=B, trans_a=True, trans_b=True,
                 Ap=self.d_logits_pp, Bp=self.rfd_hp),
            dict(A=self.d_logits_p, B=self.W2p, C_out=self.d_h, M=B, N=F * P, K=Fp, trans_b=True, epilogue=_lib.EPI_MUL_RELUMASK,
                 aux0=self.rfd_h, colsum_out=self.grads["pred_rfd.0.bias"], Ap=self.d_logits_pp, Bp=w2p, Cp=self.d_hp)])
        self._backward_backbone(l0.weight.data, self.d_h, F * P, head_Wname="pred_rfd.0.weight", dHeadp=self.d_hp,
                                head_wgrad=self._wgrad_problem(self.d_h, self.final_v, "pred_rfd.0", B, F * P, self.final_dim,
                                                               dZp=self.d_hp, X_inp=self.finalp))

    def _head_ctr(self):
        B = self.B
        fc = self.model.dnn_fc_out if self.has_fm else self.model.fc_out
        name = self.head0_name
        self._gemm(self.final_v, fc.weight.data, self.ctr_logits, B, 1, self.final_dim, epilogue=_lib.EPI_BIAS, bias=fc.bias.data)
        if self.has_fm:
            ops.add3(self.ctr_logits, self.lr_fm, None, self.ctr_logits)
        ops.bce_logits(self.ctr_logits.view(-1), self.in_labels, stats=self.stats, dlogits=self.d_logits.view(-1), ws=self.red_ws)
        if self.global_batch != B:
            ops.scale_by_scalar(self.d_logits.view(-1), self._ratio(), out=self.d_logits.view(-1))
        if self.has_fm:
            ops.copy2d(self.d_logits, self.d_lrfm)
        self._wgrad(self.d_logits, self.final_v, name, B, 1, self.final_dim)
        self._backward_backbone(fc.weight.data, self.d_logits, 1)

    def _ratio(self):
        if not hasattr(self, "_ratio_t"):
            self._ratio_t = torch.full((1,), self.B / self.global_batch, dtype=torch.float32, device=self.dev)
        return self._ratio_t

    def forward_backward(self):
        """mask -> forward -> backward; gradients land in self.grads / self.tables[*].grad.  On return every side stream has
        been joined back into the current stream."""
        self._forked.clear()
        self._tables_done.clear()
        self._adam_early_done = False
        self._ev_hyper = None
        if self.bias_grad_flat.numel():
            self.bias_grad_flat.zero_()
        self.ids_cur = self._draw_and_mask()
        self._forward_backbone(self.ids_cur)
        if self.mode == "MFP":
            self._head_mfp()
        elif self.mode == "RFD":
            self._head_rfd()
        else:
            self._head_ctr()
        self._join_streams()

    def _join_streams(self):
        for name in self.streams:
            self._join(name)

    def reduce_gradients(self):
        """hook for data-parallel runs (dist.py installs the all-reduce / all-to-all here)."""

    def _hyper_step(self):
        b1, b2 = self.betas
        ops.adamw_hyper_step(self.hyper, self.step_counter, self.lr, b1, b2, self.eps, self.sched, self.warmup_steps, self.total_steps)

    def _early_dense_adamw(self):
        """AdamW of the dense weights whose gradients are complete once the level just issued has run (see _alloc): on the 'opt'
        stream, behind that level, the side-stream weight gradients issued so far and the hyper-parameter step."""
        import os
        if not (self._early and self.multi_stream and self.adam_early is not None and self._ev_hyper is not None) or os.environ.get("MAP_B200_EARLY_ADAM", "1") == "0":
            return
        opt = self.streams["opt"]
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream())
        opt.wait_event(ev)
        if "dw" in self._forked:
            ev_dw = torch.cuda.Event()
            ev_dw.record(self.streams["dw"])
            opt.wait_event(ev_dw)
        opt.wait_event(self._ev_hyper)
        self._forked.add("opt")
        with torch.cuda.stream(opt):
            ops.adamw_multi_tensor(*self.adam_early, self.hyper)
        self._adam_early_done = True

    def optimizer_step(self):
        if not self._early:
            self._hyper_step()
        # dense parameters on the 'dw' stream, next to the table updates on the main stream
        self._fork("dw")
        with self._on("dw"):
            if self._adam_early_done:
                ops.adamw_multi_tensor(*self.adam_late, self.hyper)
            else:
                ops.adamw_multi_tensor(self.adam_table, self.adam_n, self.adam_max, self.hyper)
        for t in self.tables.values():
            if t.name in self._tables_done:
                continue
            if self.optimizer_mode == "sparse":
                ops.adamw_sparse_rows(t.p.data, t.m, t.v, t.plan, t.grad, self.hyper, t.wd)
            else:
                ops.adamw_dense_rows_sparse_grad(t.p.data, t.m, t.v, t.plan, t.grad, self.hyper, t.wd)
        self._join("dw")

    def _step_body(self):
        # the fused single-GPU schedule advances the step counter and updates the NCE tables inside the step (see _embed_lookup,
        # _nce_core); a subclass with its own exchange schedule (dist.py) or a caller driving forward_backward() /
        # optimizer_step() separately gets the plain order
        self._early = (type(self).optimizer_step is FusedStep.optimizer_step and type(self)._nce_core is FusedStep._nce_core
                       and type(self)._embed_lookup is FusedStep._embed_lookup)
        try:
            self.forward_backward()
            self.reduce_gradients()
            self.optimizer_step()
        finally:
            self._early = False

    # ------------------------------------------------------------------------------------------------ public API
    def capture(self):
        """Warm up on a side stream (lazy allocations, TMA descriptor attribute) and capture the step in a CUDA graph."""
        if not self.use_graph or self.graph is not None:
            return
        snap = self._snapshot()
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            self._step_body()
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        self._restore(snap)
        g = torch.cuda.CUDAGraph()
        # The main branch of the step (the GEMM chain = the critical path) is captured on a stream of HIGHER priority than the side
        # branches (table sorts, SIMT weight gradients, optimizer): kernel nodes inherit the priority of the stream they were
        # captured on, and the block scheduler then hands freed SM resources to a pending GEMM launch (one persistent CTA per SM
        # needing ~200 KB of shared memory) before it places more CTAs of a side kernel.  MAP_B200_MAIN_PRIO=0 restores equal
        # priorities (A/B).
        import os
        prio = int(os.environ.get("MAP_B200_MAIN_PRIO", "-1"))
        cap_stream = torch.cuda.Stream(device=self.dev, priority=prio)
        with torch.cuda.graph(g, stream=cap_stream):
            self._step_body()
        self.graph = g
        self._restore(snap)

    def _state_tensors(self):
        ts = [self.step_counter, self.hyper]
        ts += list(self.opt_param.values()) + list(self.exp_avg.values()) + list(self.exp_avg_sq.values())
        for t in self.tables.values():
            ts += [t.p.data, t.m, t.v]
        return ts

    SNAPSHOT_LIMIT_BYTES = 8 << 30

    def _snapshot(self):
        """The warm-up / capture passes execute real optimizer steps; training state is put back afterwards.  When the state
        is too large to clone (the 100M-row tables of config C5) only the step counter is restored: the two warm-up passes
        then count as two real optimizer steps on their batch — capture before loading weights if that matters."""
        ts = self._state_tensors()
        if sum(t.numel() * t.element_size() for t in ts) > self.SNAPSHOT_LIMIT_BYTES:
            ts = ts[:2]
        return [(t, t.clone()) for t in ts]

    def _restore(self, snap):
        for t, s in snap:
            t.copy_(s)
        self.refresh_weight_planes()

    def step(self, input_ids: torch.Tensor, labels: Optional[torch.Tensor] = None):
        """One training step on a device-resident batch [B, F] (int64).  No host synchronisation."""
        if self.use_planes and any(p._version != self._param_versions[n] for n, p in self.dense.items()):
            self.refresh_weight_planes()   # load_state_dict & co. since the last step
        self.in_ids.copy_(input_ids, non_blocking=True)
        if self.mode == "CTR":
            self.in_labels.copy_(labels.to(torch.float32), non_blocking=True)
        if self.use_graph:
            if self.graph is None:
                self.capture()
            self.graph.replay()
        else:
            self._step_body()
        self.steps_done += 1
        return self.loss if self.mode == "MFP" else self.stats[0:1]

    def check_health(self):
        """Raises if any kernel of the steps so far gave up on a device-side wait (grid barrier of a sort, peer-flag barrier of
        the sharded step) instead of completing.  Synchronises; call it outside the hot loop (the Trainer does at logging steps)."""
        plans = {id(t.plan): (t.name, t.plan) for t in self.tables.values() if t.plan is not None}
        for t in self.tables.values():
            m = getattr(t, "merge", None)
            if m is not None:
                plans[id(m.plan)] = (t.name + " (owner-side merge)", m.plan)
        for name, plan in plans.values():
            if plan.error_word():
                raise _lib.MapB200Error(f"sort of the ids of {name}: a grid barrier was not met (CTAs not co-resident?)")
        bar = getattr(self, "bar_error", None)
        if bar is not None and int(bar.item()):
            raise _lib.MapB200Error("sharded step: a rank did not arrive at a peer-memory barrier")

    def outputs(self):
        """Reference-style output tuple of the last step (device tensors; reading them synchronises)."""
        if self.mode == "MFP":
            return (self.loss.view(()), self.B * self.L, self.acc_count.view(()))
        if self.mode == "RFD":
            cnt = self.B * self.F
            return (self.stats[0], cnt, self.stats[1] / cnt, self.stats[2] / cnt)
        return (self.stats[0], self.ctr_logits)

    def dense_table_grad(self, name: str) -> torch.Tensor:
        """[V, D] view of a table's compact gradient == the `.grad` the reference materialises (parity tests)."""
        t = self.tables[name]
        dense = torch.zeros(t.V, t.D, dtype=torch.float32, device=self.dev)
        t.plan.scatter_dense(t.grad, t.D, dense)
        return dense
