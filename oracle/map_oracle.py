"""TEST INFRASTRUCTURE ONLY — CPU restatement of the MAP pretraining / finetuning hot path.

This file is the *checker*: only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import it.  The product (``map_code_b200``) never does, and
raises if its CUDA library is missing rather than falling back to anything here.

Every function restates one piece of the reference (``/root/reference/code``; cited file:line) in plain
``torch`` CPU ops / ``numpy``.  The restatement is pinned against the real reference by
``tests/golden/make_golden.py`` (executed in the build container where the reference is mounted; it runs the
reference through ``oracle/ref_shim.py`` and freezes inputs+outputs under ``tests/golden/``) and by
``tests/test_oracle_golden.py`` (runs anywhere, compares this file with the frozen vectors).

Third-party arithmetic that is not under ``/root/reference``: ``transformers==4.26.1`` ``AdamW`` and the
warm-up schedules (``requirements.txt:1``; call sites ``code/trainer.py:75-81``).  They are restated below from the
published 4.26.1 algorithm.  No reference test pins them => optimizer-step parity is "unpinned" (DESIGN.md).

The counter-based RNG (Philox4x32-10) has no counterpart in the reference (it uses torch's global generator on the
host); the stream layout is defined here and in ``include/map_b200.h`` and the CUDA kernels must reproduce it
bit-for-bit.  The reference consumes *our* index tensors in the parity tests.
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Tuple

import numpy as np
import torch

# --------------------------------------------------------------------------------------------------------------
# Philox4x32-10 (Salmon et al., "Parallel random numbers: as easy as 1, 2, 3", SC'11) — numpy, vectorised
# --------------------------------------------------------------------------------------------------------------
PHILOX_M0 = np.uint64(0xD2511F53)
PHILOX_M1 = np.uint64(0xCD9E8D57)
PHILOX_W0 = 0x9E3779B9
PHILOX_W1 = 0xBB67AE85
_U32 = np.uint64(0xFFFFFFFF)

# stream ids folded into the 64-bit `offset` by the host: offset = step * STREAMS_PER_STEP + stream
STREAM_MASK_FIELD = 0   # masked_index draw          (trainer.py:222-225)
STREAM_RFD_REPLACE = 1  # replacement value draw      (trainer.py:235,242,248,254)
STREAM_ALIAS = 2        # NCE noise draw              (alias_multinomial.py:89-97)
STREAM_RFD_FIELD2 = 3   # Whole-Unigram second field  (trainer.py:256)
STREAMS_PER_STEP = 8


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """All arguments broadcastable uint32 arrays (passed as uint64 holding 32-bit values). Returns 4 uint32 arrays."""
    c0 = np.asarray(c0, dtype=np.uint64) & _U32
    c1 = np.asarray(c1, dtype=np.uint64) & _U32
    c2 = np.asarray(c2, dtype=np.uint64) & _U32
    c3 = np.asarray(c3, dtype=np.uint64) & _U32
    c0, c1, c2, c3 = np.broadcast_arrays(c0, c1, c2, c3)
    k0 = int(k0) & 0xFFFFFFFF
    k1 = int(k1) & 0xFFFFFFFF
    for _ in range(10):
        p0 = PHILOX_M0 * c0
        p1 = PHILOX_M1 * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & _U32
        hi1, lo1 = p1 >> np.uint64(32), p1 & _U32
        c0, c1, c2, c3 = (hi1 ^ c1 ^ np.uint64(k0)) & _U32, lo1, (hi0 ^ c3 ^ np.uint64(k1)) & _U32, lo0
        k0 = (k0 + PHILOX_W0) & 0xFFFFFFFF
        k1 = (k1 + PHILOX_W1) & 0xFFFFFFFF
    return c0, c1, c2, c3


def philox_elem(seed: int, offset: int, elem):
    """The repo-wide convention: key = seed (lo, hi); counter = (elem_lo, elem_hi, offset_lo, offset_hi)."""
    elem = np.asarray(elem, dtype=np.uint64)
    return philox4x32_10(
        elem & _U32, elem >> np.uint64(32), np.uint64(offset & 0xFFFFFFFF), np.uint64((offset >> 32) & 0xFFFFFFFF),
        seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF,
    )


def mulhi32(w, n: int):
    """floor(w * n / 2^32) for uint32 w, n < 2^32."""
    return (np.asarray(w, dtype=np.uint64) * np.uint64(n)) >> np.uint64(32)


def mulhi64(w_lo, w_hi, n):
    """floor(((w_hi << 32) | w_lo) * n / 2^64) for n < 2^32 (scalar or array)."""
    n = np.asarray(n, dtype=np.uint64)
    lo = np.asarray(w_lo, dtype=np.uint64)
    hi = np.asarray(w_hi, dtype=np.uint64)
    return (hi * n + ((lo * n) >> np.uint64(32))) >> np.uint64(32)


def uniform24(w):
    """float32 in [0,1): top 24 bits of w times 2^-24."""
    return ((np.asarray(w, dtype=np.uint64) >> np.uint64(8)).astype(np.float32)) * np.float32(2.0 ** -24)


# --------------------------------------------------------------------------------------------------------------
# dynamic_mask  (code/trainer.py:217-266)
# --------------------------------------------------------------------------------------------------------------
def mask_num(num_fields: int, mask_ratio: float) -> int:
    """trainer.py:220  mask_num = int(num_fields * mask_ratio)"""
    return int(num_fields * mask_ratio)


def draw_masked_index(seed: int, offset: int, batch: int, L: int, F: int, sampling_method: str = "randint",
                      row0: int = 0) -> torch.Tensor:
    """masked_index [B, L] i64.

    'randint' -> trainer.py:224-225 (i.i.d. fields, duplicates allowed);
    'normal'  -> trainer.py:222-223 (randperm(F)[:L] per row == the first L steps of a Fisher-Yates shuffle).
    Element (global_row, l) uses Philox element index global_row*L + l; `row0` = first global row of this shard so a
    row-sharded run draws the same indices as the single-GPU run.
    """
    rows = np.arange(row0, row0 + batch, dtype=np.uint64)[:, None]
    ls = np.arange(L, dtype=np.uint64)[None, :]
    w0, _, _, _ = philox_elem(seed, offset, rows * np.uint64(L) + ls)
    if sampling_method == "randint":
        mi = mulhi32(w0, F).astype(np.int64)
    elif sampling_method == "normal":
        perm = np.tile(np.arange(F, dtype=np.int64), (batch, 1))
        mi = np.empty((batch, L), dtype=np.int64)
        ar = np.arange(batch)
        for l in range(L):
            j = l + mulhi32(w0[:, l], F - l).astype(np.int64)
            pl, pj = perm[ar, l].copy(), perm[ar, j].copy()
            perm[ar, l], perm[ar, j] = pj, pl
            mi[:, l] = perm[ar, l]
    else:
        raise NotImplementedError(sampling_method)  # trainer.py:226-227
    return torch.from_numpy(mi)


def dynamic_mask_mfp(input_ids: torch.Tensor, masked_index: torch.Tensor, mask_id: int = 3):
    """trainer.py:229-233: labels = gather(ids); ids = scatter(ids, mi, 3).  Returns (input_ids, labels)."""
    labels = torch.gather(input_ids, 1, masked_index)
    mask_ids = torch.full_like(masked_index, mask_id)
    return torch.scatter(input_ids, 1, masked_index, mask_ids), labels


def _scatter_last_wins(input_ids: torch.Tensor, masked_index: torch.Tensor, values: torch.Tensor) -> torch.Tensor:
    """torch.scatter on CPU applies src sequentially along dim 1 -> for duplicate fields the LAST l wins
    (SURVEY.md §7 hard part 6).  Restated with an explicit loop so it does not depend on ATen's iteration order."""
    out = input_ids.clone()
    ar = torch.arange(input_ids.shape[0])
    for l in range(masked_index.shape[1]):
        out[ar, masked_index[:, l]] = values[:, l]
    return out


def draw_rfd_replacement(seed: int, offset_replace: int, offset_field2: int, masked_index: torch.Tensor,
                         mode: str, x_train: Optional[torch.Tensor] = None,
                         idx_low: Optional[torch.Tensor] = None, idx_high: Optional[torch.Tensor] = None,
                         input_size: Optional[int] = None, row0: int = 0) -> torch.Tensor:
    """replace_feat [B, L] i64 for the four RFD_replace modes (trainer.py:234-260)."""
    B, L = masked_index.shape
    mi = masked_index.numpy()
    rows = np.arange(row0, row0 + B, dtype=np.uint64)[:, None]
    elem = rows * np.uint64(L) + np.arange(L, dtype=np.uint64)[None, :]
    w0, w1, _, _ = philox_elem(seed, offset_replace, elem)
    if mode == "Unigram":  # trainer.py:235-237: same-field value of a uniformly drawn training row
        si = mulhi64(w0, w1, x_train.shape[0]).astype(np.int64)
        rep = x_train.numpy()[si, mi]
    elif mode == "Uniform":  # trainer.py:242-243: randint(idx_low[f], idx_high[f])
        lo = idx_low.numpy().astype(np.int64)[mi]
        span = (idx_high.numpy().astype(np.int64)[mi] - lo).astype(np.uint64)
        rep = lo + mulhi64(w0, w1, span).astype(np.int64)
    elif mode == "Whole-Uniform":  # trainer.py:248-249: randint(10, input_size)
        rep = 10 + mulhi64(w0, w1, input_size - 10).astype(np.int64)
    elif mode == "Whole-Unigram":  # trainer.py:254-257: random field of a random training row
        si = mulhi64(w0, w1, x_train.shape[0]).astype(np.int64)
        v0, _, _, _ = philox_elem(seed, offset_field2, elem)
        f2 = mulhi32(v0, x_train.shape[1]).astype(np.int64)
        rep = x_train.numpy()[si, f2]
    else:
        raise NotImplementedError(mode)  # trainer.py:261-262
    return torch.from_numpy(np.ascontiguousarray(rep))


def dynamic_mask_rfd(input_ids: torch.Tensor, masked_index: torch.Tensor, replace_feat: torch.Tensor):
    """trainer.py:238-240: ids = scatter(ids, mi, replace_feat); labels = (orig != ids).float()."""
    new_ids = _scatter_last_wins(input_ids, masked_index, replace_feat)
    return new_ids, (input_ids != new_ids).float()


# --------------------------------------------------------------------------------------------------------------
# Alias sampler  (code/nce/alias_multinomial.py)
# --------------------------------------------------------------------------------------------------------------
def alias_build(probs: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """Vose tables exactly as alias_multinomial.py:40-73: float32 arithmetic, scan order 0..V-1, LIFO stacks,
    leftovers set to 1.  Pure-python for small V; uses the C restatement (oracle/alias_vose.c) when built."""
    from . import native

    p = probs.detach().cpu().float().contiguous()
    if native.available():
        return native.alias_build(p)
    K = p.numel()
    pr = (np.float32(K) * p.numpy()).astype(np.float32)  # alias_multinomial.py:49  K*prob in float32
    alias = np.zeros(K, dtype=np.int64)
    smaller, larger = [], []
    for i in range(K):
        (smaller if pr[i] < 1.0 else larger).append(i)
    while smaller and larger:
        small, large = smaller.pop(), larger.pop()
        alias[small] = large
        pr[large] = np.float32(np.float32(pr[large] - np.float32(1.0)) + pr[small])  # :63
        (smaller if pr[large] < 1.0 else larger).append(large)
    for last in smaller + larger:  # :70-71
        pr[last] = 1.0
    return torch.from_numpy(pr), torch.from_numpy(alias)


def alias_draw(prob: torch.Tensor, alias: torch.Tensor, seed: int, offset: int, n: int, elem0: int = 0) -> torch.Tensor:
    """alias_multinomial.py:89-97: kk ~ U{0..V-1}; b ~ Bernoulli(prob[kk]); out = b ? kk : alias[kk].
    Flat output of n draws; element e uses Philox element index elem0+e."""
    V = prob.numel()
    w0, w1, w2, _ = philox_elem(seed, offset, np.arange(elem0, elem0 + n, dtype=np.uint64))
    kk = mulhi64(w0, w1, V).astype(np.int64)
    u = uniform24(w2)
    b = u < prob.numpy()[kk]
    return torch.from_numpy(np.where(b, kk, alias.numpy()[kk]))


# --------------------------------------------------------------------------------------------------------------
# NCE  (code/nce/nce_loss.py, code/nce/index_linear.py)
# --------------------------------------------------------------------------------------------------------------
BACKOFF_PROB = 1e-10  # nce_loss.py:10


def nce_noise_distribution(feat_count: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor, float]:
    """nce_loss.py:60-70: returns (renormed_probs, logprob_noise, norm_term=ln V)."""
    noise = feat_count
    probs = noise / noise.sum()
    probs = probs.clamp(min=BACKOFF_PROB)
    renormed = probs / probs.sum()
    return renormed, renormed.log(), math.log(noise.numel())


def index_linear_init_bias(logprob_noise: torch.Tensor, norm_term: float) -> torch.Tensor:
    """index_linear.py:46-48: bias.weight = (logprob_noise + norm_term)[:, None]"""
    return torch.unsqueeze(logprob_noise + norm_term, 1)


def nce_sampled_logits(emb_w, bias_w, target, noise, inp):
    """index_linear.py:68-106 (_compute_sampled_logit): returns (score [B,L,K+1], indices [B,L,K+1])."""
    idx = torch.cat([target.unsqueeze(-1), noise], dim=-1)
    rows = emb_w.index_select(0, idx.reshape(-1)).view(*idx.shape, -1)
    bias = bias_w.index_select(0, idx.reshape(-1)).view_as(idx)
    score = torch.sum(inp.unsqueeze(2) * rows, dim=3) + bias
    return score, idx


def nce_forward(emb_w, bias_w, logprob_noise, norm_term: float, target, noise, inp, loss_type: str = "nce",
                reduction: str = "elementwise_mean"):
    """NCELoss.forward (nce_loss.py:79-144) with the noise given.  Returns (loss, logits[B,L,K+1], ids[B,L,K+1])
    where logits = score - norm_term (nce_loss.py:171-172)."""
    K = noise.shape[-1]
    score, idx = nce_sampled_logits(emb_w, bias_w, target, noise, inp)
    logit_model = score - norm_term
    logit_noise = logprob_noise[idx.reshape(-1)].view_as(idx)
    if loss_type == "nce":  # nce_loss.py:201-230
        logit_true = logit_model - logit_noise - math.log(K)
        label = torch.zeros_like(logit_model)
        label[:, :, 0] = 1
        loss = torch.nn.functional.binary_cross_entropy_with_logits(logit_true, label, reduction="none").sum(dim=2)
    elif loss_type == "sampled":  # nce_loss.py:232-244
        lg = logit_model - logit_noise
        labels = torch.zeros(lg.shape[:2], dtype=torch.long)
        loss = torch.nn.functional.cross_entropy(lg.view(-1, lg.size(-1)), labels.view(-1), reduction="none").view_as(labels)
    else:
        raise NotImplementedError(loss_type)
    if reduction == "elementwise_mean":
        loss = loss.mean()
    elif reduction == "sum":
        loss = loss.sum()
    return loss, logit_model, idx


def nce_full_ce_loss(emb_w, bias_w, target, inp):
    """index_linear.py:145-151 (ce_loss): full-softmax cross entropy per position."""
    score = torch.nn.functional.linear(inp, emb_w, bias_w.squeeze(1))
    return torch.nn.functional.cross_entropy(score.view(-1, score.size(-1)), target.view(-1), reduction="none").view_as(target)


# --------------------------------------------------------------------------------------------------------------
# Layers / backbones / heads  (code/layers.py, code/models.py) — functional over a reference-keyed state dict
# --------------------------------------------------------------------------------------------------------------
def embeddings_forward(weight: torch.Tensor, input_ids: torch.Tensor) -> torch.Tensor:
    """layers.py:97-102 with embed_norm=False, dropout p=0 (arguments.py:109,124 defaults)."""
    return torch.nn.functional.embedding(input_ids, weight)


def crossnet_v2_forward(params: Dict[str, torch.Tensor], prefix: str, x0: torch.Tensor, num_layers: int):
    """layers.py:197-201: Xi = Xi + X0 * Linear_i(Xi)."""
    xi = x0
    for i in range(num_layers):
        w, b = params[f"{prefix}.cross_layers.{i}.weight"], params[f"{prefix}.cross_layers.{i}.bias"]
        xi = xi + x0 * torch.nn.functional.linear(xi, w, b)
    return xi


def mlp_block_forward(params: Dict[str, torch.Tensor], prefix: str, x: torch.Tensor, num_layers: int):
    """layers.py:177-188: [Linear -> ReLU -> Dropout(p=0)] x n; Sequential indices 0,3,6,..."""
    for i in range(num_layers):
        w, b = params[f"{prefix}.dnn.{3 * i}.weight"], params[f"{prefix}.dnn.{3 * i}.bias"]
        x = torch.relu(torch.nn.functional.linear(x, w, b))
    return x


def fm_lr_forward(params, input_ids, feat_embed):
    """models.py:137-143 (LR) + layers.py:123-131 (InnerProductLayer product_sum): [B,1]."""
    wx = torch.nn.functional.embedding(input_ids, params["lr_layer.embed_w.weight"])
    lr = wx.sum(dim=1) + params["lr_layer.bias"]
    sum_sq = torch.sum(feat_embed, dim=1) ** 2
    sq_sum = torch.sum(feat_embed ** 2, dim=1)
    return lr + ((sum_sq - sq_sum) * 0.5).sum(dim=-1, keepdim=True)


def cin_forward(params: Dict[str, torch.Tensor], prefix: str, x0: torch.Tensor, units) -> torch.Tensor:
    """CIN.forward, layers.py:709-721: hadamard[b, h*M + m, d] = X0[b,h,d] * Xi[b,m,d]; X_{i+1} = Conv1d(kernel 1)(hadamard);
    pooled_i = X_{i+1}.sum(-1); output = cat(pooled)."""
    B, _, D = x0.shape
    xi, pooled = x0, []
    for i in range(len(units)):
        had = torch.einsum("bhd,bmd->bhmd", x0, xi).reshape(B, -1, D)
        w, b = params[f"{prefix}.cin_layer.layer_{i + 1}.weight"], params[f"{prefix}.cin_layer.layer_{i + 1}.bias"]
        xi = torch.nn.functional.conv1d(had, w, b).view(B, -1, D)
        pooled.append(xi.sum(dim=-1))
    return torch.cat(pooled, dim=-1)


def cin_units(cfg):
    return [int(c) for c in str(cfg.cin_layer_units).split(",")]


class OracleConfig:
    """Attribute bag with the reference's config field names (arguments.py:103-157, run.py:49-61)."""

    def __init__(self, **kw):
        d = dict(model_name="DCNv2", embed_size=16, hidden_size=1000, num_hidden_layers=3, num_cross_layers=3,
                 hidden_act="relu", hidden_dropout_rate=0.0, embed_dropout_rate=0.0, embed_norm=False,
                 layer_norm_eps=1e-12, pt_neg_num=25, proj_size=32, input_size=None, num_fields=None,
                 pretrain=True, pt_type="MFP", RFD_replace="Unigram", cin_layer_units="50,50", use_lr=False)
        d.update(kw)
        for k, v in d.items():
            setattr(self, k, v)


def backbone_forward(cfg, params: Dict[str, torch.Tensor], input_ids: torch.Tensor) -> torch.Tensor:
    """DCNV2.forward models.py:306-315 / DeepFM.forward models.py:217-226 up to `final_output` (pretrain) or the
    pre-head vector (CTR)."""
    name = cfg.model_name.lower()
    feat = embeddings_forward(params["embed.embedding.weight"], input_ids)
    flat = feat.flatten(start_dim=1)
    if name == "dcnv2":
        cross = crossnet_v2_forward(params, "cross_net", flat, cfg.num_cross_layers)
        if cfg.num_hidden_layers > 0:
            dnn = mlp_block_forward(params, "parallel_dnn", flat, cfg.num_hidden_layers)
            return torch.cat([cross, dnn], dim=-1)
        return cross
    if name == "deepfm":
        dnn_vec = mlp_block_forward(params, "dnn", flat, cfg.num_hidden_layers)
        if cfg.pretrain:
            return torch.cat([dnn_vec, fm_lr_forward(params, input_ids, feat)], dim=1)
        return dnn_vec
    if name == "dnn":
        return mlp_block_forward(params, "dnn", flat, cfg.num_hidden_layers)
    if name == "xdeepfm":   # xDeepFM.forward models.py:262-269
        final = cin_forward(params, "cin", feat, cin_units(cfg))
        if cfg.num_hidden_layers > 0:
            final = torch.cat([final, mlp_block_forward(params, "dnn", flat, cfg.num_hidden_layers)], dim=1)
        return final
    raise NotImplementedError(cfg.model_name)


def model_forward(cfg, params, input_ids, labels=None, masked_index=None, noise=None, logprob_noise=None):
    """BaseModel.get_outputs (models.py:62-95) on top of backbone_forward.  `noise` [B,L,K] is given explicitly (the
    reference draws it inside mfp_criterion; parity tests patch alias.draw to return the same tensor).
    Returns the reference tuples, with acc as tensors (no .item())."""
    B = input_ids.shape[0]
    final = backbone_forward(cfg, params, input_ids)
    name = cfg.model_name.lower()
    if cfg.pretrain:
        if cfg.pt_type == "MFP":  # models.py:73-78
            enc = torch.nn.functional.linear(final, params["feat_encoder.weight"], params["feat_encoder.bias"])
            enc = enc.view(B, cfg.num_fields, cfg.proj_size)
            sel = torch.gather(enc, 1, masked_index.unsqueeze(-1).repeat(1, 1, cfg.proj_size))
            lpn = params["mfp_criterion.logprob_noise"] if logprob_noise is None else logprob_noise
            loss, logits, ids = nce_forward(params["mfp_criterion.emb.weight"], params["mfp_criterion.bias.weight"],
                                            lpn, math.log(cfg.input_size), labels, noise, sel)
            total_acc = (logits.argmax(dim=2) == 0).sum()
            return loss, labels.shape[0] * labels.shape[1], total_acc, logits, ids
        if cfg.pt_type == "RFD":  # models.py:79-85
            h = torch.relu(torch.nn.functional.linear(final, params["pred_rfd.0.weight"], params["pred_rfd.0.bias"]))
            logits = torch.nn.functional.linear(h, params["pred_rfd.2.weight"], params["pred_rfd.2.bias"])
            loss = torch.nn.functional.binary_cross_entropy_with_logits(logits, labels)
            count = labels.shape[0] * labels.shape[1]
            acc = ((torch.sigmoid(logits) > 0.5).float() == labels).sum() / count
            return loss, count, acc, labels.mean(), logits
        raise NotImplementedError(cfg.pt_type)
    # CTR: models.py:88-93 + the per-backbone output layer
    if name == "dcnv2":
        logits = torch.nn.functional.linear(final, params["fc_out.weight"], params["fc_out.bias"])
    elif name == "deepfm":  # models.py:227-231
        feat = embeddings_forward(params["embed.embedding.weight"], input_ids)
        logits = torch.nn.functional.linear(final, params["dnn_fc_out.weight"], params["dnn_fc_out.bias"])
        logits = logits + fm_lr_forward(params, input_ids, feat)
    elif name == "xdeepfm":  # models.py:273-276
        logits = torch.nn.functional.linear(final, params["fc.weight"], params["fc.bias"])
        if getattr(cfg, "use_lr", False):
            wx = torch.nn.functional.embedding(input_ids, params["lr_layer.embed_w.weight"])
            logits = logits + wx.sum(dim=1) + params["lr_layer.bias"]
    else:
        logits = torch.nn.functional.linear(final, params["fc_out.weight"], params["fc_out.bias"])
    if labels is None:
        return (logits,)
    loss = torch.nn.functional.binary_cross_entropy_with_logits(logits.view(-1), labels.float())
    return loss, logits


def init_params(cfg, feat_count: Optional[torch.Tensor], seed: int = 1) -> Dict[str, torch.Tensor]:
    """Random-init parameter dict with the reference's state_dict key names and init distributions
    (layers.py:86-91 embedding normal std=sqrt(2/(F+D)); nn.Linear default init; index_linear.py:40-48).
    Not bit-identical to a reference construction (different RNG consumption order) — used for synthetic runs."""
    g = torch.Generator().manual_seed(seed)
    V, F, D, H, P = cfg.input_size, cfg.num_fields, cfg.embed_size, cfg.hidden_size, cfg.proj_size
    p: Dict[str, torch.Tensor] = {}

    def linear(name, n_out, n_in):
        bound = 1.0 / math.sqrt(n_in)
        p[name + ".weight"] = (torch.rand(n_out, n_in, generator=g) * 2 - 1) * bound
        p[name + ".bias"] = (torch.rand(n_out, generator=g) * 2 - 1) * bound

    p["embed.embedding.weight"] = torch.randn(V, D, generator=g) * math.sqrt(2.0 / (F + D))
    in_dim = F * D
    name = cfg.model_name.lower()
    if name == "dcnv2":
        for i in range(cfg.num_cross_layers):
            linear(f"cross_net.cross_layers.{i}", in_dim, in_dim)
        d = in_dim
        for i in range(cfg.num_hidden_layers):
            linear(f"parallel_dnn.dnn.{3 * i}", H, d)
            d = H
        final = in_dim + (H if cfg.num_hidden_layers > 0 else 0)
    elif name in ("deepfm", "dnn"):
        d = in_dim
        for i in range(cfg.num_hidden_layers):
            linear(f"dnn.dnn.{3 * i}", H, d)
            d = H
        final = H
        if name == "deepfm":
            p["lr_layer.embed_w.weight"] = torch.randn(V, 1, generator=g)
            p["lr_layer.bias"] = torch.zeros(1)
            if cfg.pretrain:
                final = H + 1
    elif name == "xdeepfm":
        units, prev = cin_units(cfg), F
        for i, u in enumerate(units):
            c_in = F * prev
            bound = 1.0 / math.sqrt(c_in)
            p[f"cin.cin_layer.layer_{i + 1}.weight"] = (torch.rand(u, c_in, 1, generator=g) * 2 - 1) * bound
            p[f"cin.cin_layer.layer_{i + 1}.bias"] = (torch.rand(u, generator=g) * 2 - 1) * bound
            prev = u
        d = in_dim
        for i in range(cfg.num_hidden_layers):
            linear(f"dnn.dnn.{3 * i}", H, d)
            d = H
        final = sum(units) + (H if cfg.num_hidden_layers > 0 else 0)
        if not cfg.pretrain and getattr(cfg, "use_lr", False):
            p["lr_layer.embed_w.weight"] = torch.randn(V, 1, generator=g)
            p["lr_layer.bias"] = torch.zeros(1)
    else:
        raise NotImplementedError(cfg.model_name)
    if cfg.pretrain:
        if cfg.pt_type == "MFP":
            linear("feat_encoder", F * P, final)
            _, lpn, norm = nce_noise_distribution(feat_count)
            stdv = 1.0 / math.sqrt(P)
            p["mfp_criterion.emb.weight"] = (torch.rand(V, P, generator=g) * 2 - 1) * stdv
            p["mfp_criterion.bias.weight"] = index_linear_init_bias(lpn, norm)
            p["mfp_criterion.logprob_noise"] = lpn
        elif cfg.pt_type == "RFD":
            linear("pred_rfd.0", F * P, final)
            linear("pred_rfd.2", F, F * P)
        else:
            raise NotImplementedError(cfg.pt_type)
    else:
        linear({"deepfm": "dnn_fc_out", "xdeepfm": "fc"}.get(name, "fc_out"), 1, final)
    return p


TRAINABLE_EXCLUDE = ("mfp_criterion.logprob_noise", "mfp_criterion.alias.prob", "mfp_criterion.alias.alias",
                     "ip_layer.field_p", "ip_layer.field_q", "ip_layer.upper_triangle_mask")


def is_no_decay(name: str) -> bool:
    """trainer.py:61-72: no_decay = ["bias", "LayerNorm.weight"] matched as substrings of the parameter name.
    NB `mfp_criterion.bias.weight` contains "bias" -> no decay."""
    return any(nd in name for nd in ("bias", "LayerNorm.weight"))


# --------------------------------------------------------------------------------------------------------------
# transformers==4.26.1 AdamW + schedules (restated; call sites trainer.py:75-81)
# --------------------------------------------------------------------------------------------------------------
def hf_adamw_update(p, g, m, v, step: int, lr: float, beta1: float, beta2: float, eps: float, weight_decay: float,
                    correct_bias: bool = True):
    """One in-place update of a tensor. step is the 1-based count AFTER increment.
    m <- b1 m + (1-b1) g;  v <- b2 v + (1-b2) g^2;  denom = sqrt(v)+eps;
    step_size = lr*sqrt(1-b2^t)/(1-b1^t);  p <- p - step_size*m/denom;  then p <- p - lr*wd*p."""
    m.mul_(beta1).add_(g, alpha=1.0 - beta1)
    v.mul_(beta2).addcmul_(g, g, value=1.0 - beta2)
    denom = v.sqrt().add_(eps)
    step_size = lr
    if correct_bias:
        step_size = step_size * math.sqrt(1.0 - beta2 ** step) / (1.0 - beta1 ** step)
    p.addcdiv_(m, denom, value=-step_size)
    if weight_decay > 0.0:
        p.add_(p, alpha=-lr * weight_decay)


class HFAdamW(torch.optim.Optimizer):
    """Drop-in for `transformers.AdamW` (4.26.1): same constructor signature and update rule."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-6, weight_decay=0.0, correct_bias=True,
                 no_deprecation_warning=False):
        if lr < 0.0:
            raise ValueError(f"Invalid learning rate: {lr} - should be >= 0.0")
        if not 0.0 <= betas[0] < 1.0 or not 0.0 <= betas[1] < 1.0:
            raise ValueError(f"Invalid beta parameters: {betas}")
        if not 0.0 <= eps:
            raise ValueError(f"Invalid epsilon value: {eps} - should be >= 0.0")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, correct_bias=correct_bias))

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        for group in self.param_groups:
            for p in group["params"]:
                if p.grad is None:
                    continue
                if p.grad.is_sparse:
                    raise RuntimeError("Adam does not support sparse gradients, please consider SparseAdam instead")
                st = self.state[p]
                if len(st) == 0:
                    st["step"] = 0
                    st["exp_avg"] = torch.zeros_like(p)
                    st["exp_avg_sq"] = torch.zeros_like(p)
                st["step"] += 1
                b1, b2 = group["betas"]
                hf_adamw_update(p, p.grad, st["exp_avg"], st["exp_avg_sq"], st["step"], group["lr"], b1, b2,
                                group["eps"], group["weight_decay"], group["correct_bias"])
        return loss


def cosine_schedule_lambda(num_warmup_steps: int, num_training_steps: int, num_cycles: float = 0.5):
    """transformers.get_cosine_schedule_with_warmup lr_lambda."""

    def f(current_step: int) -> float:
        if current_step < num_warmup_steps:
            return float(current_step) / float(max(1, num_warmup_steps))
        progress = float(current_step - num_warmup_steps) / float(max(1, num_training_steps - num_warmup_steps))
        return max(0.0, 0.5 * (1.0 + math.cos(math.pi * float(num_cycles) * 2.0 * progress)))

    return f


def constant_schedule_lambda(num_warmup_steps: int):
    """transformers.get_constant_schedule_with_warmup lr_lambda."""

    def f(current_step: int) -> float:
        if current_step < num_warmup_steps:
            return float(current_step) / float(max(1.0, num_warmup_steps))
        return 1.0

    return f


# --------------------------------------------------------------------------------------------------------------
# A complete oracle training step (used by tests and by bench.py's cpu_baseline leg)
# --------------------------------------------------------------------------------------------------------------
class OracleTrainer:
    """Restates the inner loop body of Trainer.MFP_pretrain / RFD_pretrain / train (trainer.py:306-331, 431-455,
    122-143): dynamic_mask -> forward -> backward -> HF-AdamW -> schedule, on CPU, dense gradients and dense AdamW over
    every table exactly like the reference."""

    def __init__(self, cfg, params: Dict[str, torch.Tensor], alias_prob=None, alias_alias=None, x_train=None,
                 lr=1e-3, weight_decay=5e-2, betas=(0.9, 0.999), eps=1e-8, mask_ratio=0.1, sampling_method="randint",
                 seed=42, lr_lambda=None, idx_low=None, idx_high=None, table_update: str = "dense"):
        """table_update="dense" is the reference (transformers.AdamW sweeps every row of every [V, .] table, trainer.py:140).
        table_update="touched_rows" restates the PRODUCT's documented `optimizer_mode="sparse"` semantics for its parity tests
        (DESIGN.md §5.1): a table row takes part in the optimizer step only when its id occurs in this step's batch
        (embedding / first-order tables: the masked input ids; NCE tables: targets + noise); the update rule is unchanged."""
        if table_update not in ("dense", "touched_rows"):
            raise ValueError(table_update)
        self.table_update = table_update
        self._touched = {}
        self.cfg = cfg
        self.buffers = {k: v for k, v in params.items() if k in TRAINABLE_EXCLUDE}
        self.params = {k: v.clone().requires_grad_(True) for k, v in params.items() if k not in TRAINABLE_EXCLUDE}
        self.alias_prob, self.alias_alias, self.x_train = alias_prob, alias_alias, x_train
        self.idx_low, self.idx_high = idx_low, idx_high
        self.base_lr, self.wd, self.betas, self.eps = lr, weight_decay, betas, eps
        self.mask_ratio, self.sampling_method, self.seed = mask_ratio, sampling_method, seed
        self.lr_lambda = lr_lambda or (lambda s: 1.0)
        self.state = {k: (torch.zeros_like(v), torch.zeros_like(v)) for k, v in self.params.items()}
        self.global_step = 0

    def draw(self, input_ids, row0: int = 0):
        cfg, step = self.cfg, self.global_step
        B, F = input_ids.shape
        L = mask_num(F, self.mask_ratio)
        base = step * STREAMS_PER_STEP
        mi = draw_masked_index(self.seed, base + STREAM_MASK_FIELD, B, L, F, self.sampling_method, row0)
        out = {"masked_index": mi}
        if cfg.pt_type == "MFP":
            ids, labels = dynamic_mask_mfp(input_ids, mi)
            K = cfg.pt_neg_num
            noise = alias_draw(self.alias_prob, self.alias_alias, self.seed, base + STREAM_ALIAS, B * L * K,
                               elem0=row0 * L * K).view(B, L, K)
            out.update(input_ids=ids, labels=labels, noise=noise)
        else:
            rep = draw_rfd_replacement(self.seed, base + STREAM_RFD_REPLACE, base + STREAM_RFD_FIELD2, mi,
                                       cfg.RFD_replace, self.x_train, self.idx_low, self.idx_high, cfg.input_size, row0)
            ids, labels = dynamic_mask_rfd(input_ids, mi, rep)
            out.update(input_ids=ids, labels=labels)
        return out

    def forward_backward(self, batch):
        all_p = dict(self.params)
        all_p.update(self.buffers)
        for v in self.params.values():
            v.grad = None
        if self.cfg.pretrain and self.cfg.pt_type == "MFP":
            outs = model_forward(self.cfg, all_p, batch["input_ids"], batch["labels"], batch["masked_index"], batch["noise"])
        else:
            outs = model_forward(self.cfg, all_p, batch["input_ids"], batch["labels"])
        outs[0].backward()
        if self.table_update == "touched_rows":
            V = self.cfg.input_size
            t_in = torch.zeros(V, dtype=torch.bool)
            t_in[batch["input_ids"].reshape(-1)] = True
            self._touched = {"embed.embedding.weight": t_in, "lr_layer.embed_w.weight": t_in}
            if self.cfg.pretrain and self.cfg.pt_type == "MFP":
                t_nce = torch.zeros(V, dtype=torch.bool)
                t_nce[batch["labels"].reshape(-1)] = True
                t_nce[batch["noise"].reshape(-1)] = True
                self._touched["mfp_criterion.emb.weight"] = t_nce
                self._touched["mfp_criterion.bias.weight"] = t_nce
        return outs

    def optimizer_step(self):
        lr = self.base_lr * self.lr_lambda(self.global_step)
        self.global_step += 1
        with torch.no_grad():
            for k, p in self.params.items():
                if p.grad is None:
                    continue
                m, v = self.state[k]
                rows = self._touched.get(k) if self.table_update == "touched_rows" else None
                if rows is not None:
                    pr, mr, vr = p[rows], m[rows], v[rows]
                    hf_adamw_update(pr, p.grad[rows], mr, vr, self.global_step, lr, self.betas[0], self.betas[1], self.eps,
                                    0.0 if is_no_decay(k) else self.wd)
                    p[rows], m[rows], v[rows] = pr, mr, vr
                    continue
                hf_adamw_update(p, p.grad, m, v, self.global_step, lr, self.betas[0], self.betas[1], self.eps,
                                0.0 if is_no_decay(k) else self.wd)

    def step(self, input_ids, labels=None):
        if self.cfg.pretrain:
            batch = self.draw(input_ids)
        else:
            batch = {"input_ids": input_ids, "labels": labels}
        outs = self.forward_backward(batch)
        self.optimizer_step()
        return outs
