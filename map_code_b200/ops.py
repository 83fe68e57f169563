"""Functional wrappers: torch CUDA tensors in, C-ABI kernel launches on torch's current stream, torch tensors out.

torch is plumbing here (device memory + streams); every device operation below is one of our sm_100a kernels.
All functions are CUDA-graph capturable when the caller supplies the output / workspace tensors (`out=` arguments).
"""
from __future__ import annotations

import ctypes as C
import math
import os
from typing import Optional, Tuple

import torch

from . import _lib
from ._lib import GemmArgs, GemmSplitArgs, call


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def _check(t: torch.Tensor, dtype, name: str, contiguous: bool = True):
    if not t.is_cuda:
        raise _lib.MapB200Error(f"{name}: expected a CUDA tensor (map_code_b200 has no CPU path)")
    if t.dtype != dtype:
        raise TypeError(f"{name}: expected {dtype}, got {t.dtype}")
    if contiguous and not t.is_contiguous():
        raise ValueError(f"{name}: expected a contiguous tensor")
    return t


def key_bits(V: int) -> int:
    return max(1, int(V - 1).bit_length())


# ------------------------------------------------------------------------------------------------ K1
def emb_gather(table: torch.Tensor, ids: torch.Tensor, out: Optional[torch.Tensor] = None,
               oob_flag: Optional[torch.Tensor] = None) -> torch.Tensor:
    _check(table, torch.float32, "table")
    _check(ids, torch.int64, "ids")
    V, D = table.shape
    if out is None:
        out = torch.empty(*ids.shape, D, dtype=torch.float32, device=table.device)
    if ids.numel() == 0:
        return out
    if _lib.PROFILE is not None or _lib.TIMELINE is not None or _lib.RECORD is not None:
        _lib.CURRENT_TAG = ("gather", ids.numel(), D)
    call("map_emb_gather_f32", table.data_ptr(), V, D, ids.data_ptr(), ids.numel(), out.data_ptr(), _ptr(oob_flag), _stream())
    return out


def emb_gather_sharded(shard_ptrs, R: int, V: int, D: int, ids: torch.Tensor, out: torch.Tensor) -> torch.Tensor:
    """out[i,:] = shard[ids[i] % R][ids[i] // R, :]; shard_ptrs: ctypes array of R device base pointers (peer memory)."""
    _check(ids, torch.int64, "ids")
    if _lib.PROFILE is not None or _lib.TIMELINE is not None or _lib.RECORD is not None:
        _lib.CURRENT_TAG = ("gather", ids.numel(), D)
    call("map_emb_gather_sharded_f32", shard_ptrs, R, V, D, ids.data_ptr(), ids.numel(), out.data_ptr(), None, _stream())
    return out


def p2p_barrier(flag_ptrs, R: int, rank: int, site: int, n_sites: int, epochs: torch.Tensor, error_word: Optional[torch.Tensor] = None):
    """stream-ordered barrier over the ranks through peer-memory flags (map_p2p_barrier)"""
    call("map_p2p_barrier", flag_ptrs, R, rank, site, n_sites, epochs.data_ptr(), _ptr(error_word), _stream())


def p2p_reduce(send_ptrs, R: int, first: int, count: int, out: torch.Tensor):
    """out[i] = sum_q send[q][first + i] over the ranks in order 0..R-1 (peer loads); out: fp32 view of `count` floats"""
    if _lib.PROFILE is not None or _lib.TIMELINE is not None or _lib.RECORD is not None:
        _lib.CURRENT_TAG = ("bytes", count * 4)
    call("map_p2p_reduce_f32", send_ptrs, R, first, count, out.data_ptr(), _stream())


def p2p_gather_slices(send_ptrs, R: int, first: int, count: int, slice_len: int, out: torch.Tensor):
    """out[i] = send[i // slice_len][first + i] (peer loads): second half of the two-shot all-reduce"""
    if _lib.PROFILE is not None or _lib.TIMELINE is not None or _lib.RECORD is not None:
        _lib.CURRENT_TAG = ("bytes", count * 4)
    call("map_p2p_gather_slices_f32", send_ptrs, R, first, count, slice_len, out.data_ptr(), _stream())


def owned_compact(uniq_ptrs, n_unique_ptrs, R: int, rank: int, cap: int, keys: torch.Tensor, src: torch.Tensor, n_out: torch.Tensor):
    call("map_owned_compact", uniq_ptrs, n_unique_ptrs, R, rank, cap, keys.data_ptr(), src.data_ptr(), n_out.data_ptr(), _stream())


# ------------------------------------------------------------------------------------------------ K2
class DedupPlan:
    """Pre-allocated buffers for the dedup pipeline of one id stream of fixed length n (graph-capturable)."""

    def __init__(self, n: int, V: int, device, alloc=None, seg_shift: int = 0, n_dev: Optional[torch.Tensor] = None):
        """alloc(shape, dtype) -> tensor places `uniq` and `n_unique` (what the owner-side merge of the sharded tables reads
        from the other ranks) in peer-visible memory; seg_shift / n_dev: see map_dedup_ids_ex."""
        self.n, self.V = int(n), int(V)
        self.key_bits = key_bits(V)
        self.seg_shift, self.n_dev = int(seg_shift), n_dev
        mk = alloc if alloc is not None else (lambda shape, dtype: torch.zeros(shape, dtype=dtype, device=device))
        self.uniq = mk((self.n,), torch.int64)
        self.seg_start = torch.empty(self.n + 1, dtype=torch.int32, device=device)
        self.occ_sorted = torch.empty(self.n, dtype=torch.int32, device=device)
        self.n_unique = mk((1,), torch.int32)
        self.pos_seg = torch.empty(self.n, dtype=torch.int32, device=device)   # sorted position -> segment index
        self.ws_bytes = int(_lib.load().map_dedup_workspace_bytes(self.n))
        self.ws = torch.empty(self.ws_bytes, dtype=torch.uint8, device=device)

    def run(self, ids: torch.Tensor):
        _check(ids, torch.int64, "ids")
        assert ids.numel() == self.n
        if _lib.PROFILE is not None or _lib.TIMELINE is not None or _lib.RECORD is not None:
            _lib.CURRENT_TAG = ("dedup", self.n, self.key_bits)
        call("map_dedup_ids_ex", ids.data_ptr(), self.n, _ptr(self.n_dev), self.key_bits, self.seg_shift, self.uniq.data_ptr(),
             self.seg_start.data_ptr(), self.occ_sorted.data_ptr(), self.n_unique.data_ptr(), self.pos_seg.data_ptr(), self.ws.data_ptr(),
             self.ws_bytes, _stream())
        return self

    def error_word(self) -> int:
        """1 if a grid barrier of the last sort was not met within its spin bound (the kernel then gave up instead of hanging
        the GPU and its outputs are garbage).  Synchronises: call it outside the step (FusedStep.check_health)."""
        off = int(_lib.load().map_dedup_debug_offset(self.n)) - 64 + 4
        return int(self.ws[off:off + 4].view(torch.int32).item())

    def reduce_peer_rows(self, row_ptrs, n_peers: int, rows_per_peer: int, D: int, occ_map: torch.Tensor, out: torch.Tensor):
        """segment sums over rows that live in the R ranks' compact gradients (row_ptrs: ctypes array of R device pointers,
        peer memory): row code = occ_map[occurrence] = peer * rows_per_peer + row."""
        if _lib.PROFILE is not None or _lib.TIMELINE is not None or _lib.RECORD is not None:
            _lib.CURRENT_TAG = ("segred_peer", self.n, D)
        call("map_segment_reduce_rows_ex", None, D, D, None, 1, self.occ_sorted.data_ptr(), self.seg_start.data_ptr(),
             self.n_unique.data_ptr(), self.pos_seg.data_ptr(), self.n, _ptr(self.n_dev), occ_map.data_ptr(), row_ptrs, n_peers, rows_per_peer,
             out.data_ptr(), None, _stream())
        return out

    def reduce_rows(self, rows: torch.Tensor, D: int, ld_rows: Optional[int] = None, scale: Optional[torch.Tensor] = None,
                    group: int = 1, out: Optional[torch.Tensor] = None, scalar_out: Optional[torch.Tensor] = None):
        if out is None:
            out = torch.empty(self.n, D, dtype=torch.float32, device=rows.device)
        if _lib.PROFILE is not None or _lib.TIMELINE is not None or _lib.RECORD is not None:
            _lib.CURRENT_TAG = ("segred", self.n, D)
        call("map_segment_reduce_rows_ex", rows.data_ptr(), ld_rows if ld_rows is not None else D, D, _ptr(scale), group,
             self.occ_sorted.data_ptr(), self.seg_start.data_ptr(), self.n_unique.data_ptr(), self.pos_seg.data_ptr(), self.n, _ptr(self.n_dev),
             None, None, 0, 0, out.data_ptr(), _ptr(scalar_out), _stream())
        return out

    def scatter_dense(self, grad_compact: torch.Tensor, D: int, dense: torch.Tensor):
        call("map_scatter_rows", grad_compact.data_ptr(), self.uniq.data_ptr(), self.n_unique.data_ptr(), self.n, D,
             dense.data_ptr(), _stream())
        return dense


def feat_count(ids: torch.Tensor, V: int, chunk: int = 1 << 23) -> torch.Tensor:
    """Occurrence count of every id in `ids` (any shape, int64, on the device) -> float32 [V]: the device form of the reference's
    `feat_count` builder (code/dataset.py:49-62 runs a Python Counter over every id of the training split).  Sort + run-length
    with the K2 kernels: segment sums of ones over the sorted ids, scattered to the dense vector; chunked so that the sort
    workspace stays bounded for 4e7-id training matrices."""
    _check(ids, torch.int64, "ids", contiguous=False)
    flat = ids.reshape(-1).contiguous()
    dev = flat.device
    total = torch.zeros(V, 1, dtype=torch.float32, device=dev)
    part = torch.zeros(V, 1, dtype=torch.float32, device=dev)
    for lo in range(0, flat.numel(), chunk):
        piece = flat[lo:lo + chunk]
        n = piece.numel()
        plan = DedupPlan(n, V, dev).run(piece)
        ones = torch.ones(n, 1, dtype=torch.float32, device=dev)
        compact = plan.reduce_rows(ones, 1)
        if lo == 0 and n == flat.numel():
            return plan.scatter_dense(compact, 1, total).view(-1)
        part.zero_()
        plan.scatter_dense(compact, 1, part)
        add3(total, part, None, total)
    return total.view(-1)


# ------------------------------------------------------------------------------------------------ AdamW
def adamw_hyper_step(hyper, step_counter, base_lr, beta1, beta2, eps, sched: int, warmup: int, total: int):
    call("map_adamw_hyper_step", hyper.data_ptr(), step_counter.data_ptr(), base_lr, beta1, beta2, eps, sched, warmup, total, _stream())


def adamw_hyper_set(hyper, lr, beta1, beta2, eps, step: int):
    call("map_adamw_hyper_set", hyper.data_ptr(), lr, beta1, beta2, eps, step, _stream())


def make_adamw_tensor_list(entries, device) -> Tuple[torch.Tensor, int, int]:
    """entries: list of (p, g, m, v, weight_decay, p_t or None[, planes or None]).  `planes` [n_planes, *p.shape] bf16: the update
    rewrites the parameter's bf16 planes (operands of the split-bf16 GEMMs).  Returns (device descriptor table, n, max_elems)."""
    arr = (_lib.AdamwTensor * len(entries))()
    mx = 1
    for i, ent in enumerate(entries):
        p, g, m, v, wd, p_t = ent[:6]
        planes = ent[6] if len(ent) > 6 else None
        for t in (p, g, m, v):
            _check(t, torch.float32, "adamw tensor")
        rows, cols = (p.shape[0], p.shape[1]) if (p_t is not None) else (0, 0)
        n_pl, pl_ptr, pl_stride = 0, None, 0
        if planes is not None:
            _check(planes, torch.bfloat16, "adamw planes", contiguous=False)
            if tuple(planes.shape[1:]) != tuple(p.shape) or p.numel() % 4 != 0 or not planes[0].is_contiguous():
                raise ValueError("adamw planes must be [n_planes, *p.shape] with contiguous planes and p.numel() % 4 == 0")
            n_pl, pl_ptr, pl_stride = planes.shape[0], planes.data_ptr(), planes.stride(0)
        arr[i] = _lib.AdamwTensor(p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(), _ptr(p_t), p.numel(), rows, cols,
                                  float(wd), n_pl, pl_ptr, pl_stride)
        mx = max(mx, p.numel())
    raw = bytes(arr)
    table = torch.frombuffer(bytearray(raw), dtype=torch.uint8).to(device)
    return table, len(entries), mx


def adamw_multi_tensor(table: torch.Tensor, n: int, max_elems: int, hyper: torch.Tensor):
    call("map_adamw_multi_tensor", table.data_ptr(), n, max_elems, hyper.data_ptr(), _stream())


def adamw_sparse_rows(table, m, v, plan: DedupPlan, grad_compact, hyper, weight_decay: float):
    D = table.shape[1]
    if _lib.PROFILE is not None or _lib.TIMELINE is not None or _lib.RECORD is not None:
        _lib.CURRENT_TAG = ("sparse_adamw", plan.n, D)
    call("map_adamw_sparse_rows", table.data_ptr(), m.data_ptr(), v.data_ptr(), D, plan.uniq.data_ptr(), grad_compact.data_ptr(),
         plan.n_unique.data_ptr(), plan.n, hyper.data_ptr(), float(weight_decay), _stream())


def adamw_dense_rows_sparse_grad(table, m, v, plan: DedupPlan, grad_compact, hyper, weight_decay: float):
    V, D = table.shape
    call("map_adamw_dense_rows_sparse_grad", table.data_ptr(), m.data_ptr(), v.data_ptr(), V, D, plan.uniq.data_ptr(),
         grad_compact.data_ptr(), plan.n_unique.data_ptr(), hyper.data_ptr(), float(weight_decay), _stream())


# ------------------------------------------------------------------------------------------------ K8 / K9
def mask_index(B: int, L: int, F: int, sampling_method: str, seed: int, offset: int, row0: int = 0, device="cuda",
               out: Optional[torch.Tensor] = None, step_dev: Optional[torch.Tensor] = None) -> torch.Tensor:
    if sampling_method == "randint":
        sm = _lib.SAMPLING_RANDINT
    elif sampling_method == "normal":
        sm = _lib.SAMPLING_NORMAL
    else:
        raise NotImplementedError(sampling_method)  # reference: trainer.py:226-227
    if out is None:
        out = torch.empty(B, L, dtype=torch.int64, device=device)
    call("map_mask_index_philox", out.data_ptr(), B, L, F, sm, seed, offset, row0, _ptr(step_dev), _stream())
    return out


def mfp_mask_apply(ids: torch.Tensor, masked_index: torch.Tensor, mask_id: int = 3, ids_out=None, labels=None):
    _check(ids, torch.int64, "ids")
    _check(masked_index, torch.int64, "masked_index")
    B, F = ids.shape
    L = masked_index.shape[1]
    if ids_out is None:
        ids_out = torch.empty_like(ids)
    if labels is None:
        labels = torch.empty(B, L, dtype=torch.int64, device=ids.device)
    call("map_mfp_mask_apply", ids.data_ptr(), masked_index.data_ptr(), B, F, L, mask_id, ids_out.data_ptr(), labels.data_ptr(), _stream())
    return ids_out, labels


def rfd_replace(ids, masked_index, mode: str, seed: int, offset_replace: int, offset_field2: int = 0, x_train=None,
                idx_low=None, idx_high=None, input_size: int = 0, row0: int = 0, ids_out=None, labels=None, replace_out=None,
                step_dev: Optional[torch.Tensor] = None):
    if mode not in _lib.RFD_MODES:
        raise NotImplementedError(mode)  # reference: trainer.py:261-262
    _check(ids, torch.int64, "ids")
    _check(masked_index, torch.int64, "masked_index")
    B, F = ids.shape
    L = masked_index.shape[1]
    if ids_out is None:
        ids_out = torch.empty_like(ids)
    if labels is None:
        labels = torch.empty(B, F, dtype=torch.float32, device=ids.device)
    n_train = 0 if x_train is None else x_train.shape[0]
    call("map_rfd_replace_philox", ids.data_ptr(), masked_index.data_ptr(), B, F, L, _lib.RFD_MODES[mode], _ptr(x_train), n_train,
         _ptr(idx_low), _ptr(idx_high), input_size, seed, offset_replace, offset_field2, row0, _ptr(step_dev), ids_out.data_ptr(),
         labels.data_ptr(), _ptr(replace_out), _stream())
    return ids_out, labels


# ------------------------------------------------------------------------------------------------ K5
def alias_build(probs: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """HOST function of the library (native Vose builder, bit-identical to the reference's Python loop)."""
    p = probs.detach().to("cpu", torch.float32).contiguous()
    V = p.numel()
    prob = torch.empty(V, dtype=torch.float32, device="cpu")    # explicit: a torch.device(...) context may be active
    alias = torch.empty(V, dtype=torch.int64, device="cpu")
    call("map_alias_build", p.data_ptr(), V, prob.data_ptr(), alias.data_ptr())
    return prob, alias


def alias_draw(prob: torch.Tensor, alias: torch.Tensor, seed: int, offset: int, n: int, elem0: int = 0,
               out: Optional[torch.Tensor] = None, step_dev: Optional[torch.Tensor] = None) -> torch.Tensor:
    _check(prob, torch.float32, "prob")
    _check(alias, torch.int64, "alias")
    if out is None:
        out = torch.empty(n, dtype=torch.int64, device=prob.device)
    call("map_alias_draw_philox", prob.data_ptr(), alias.data_ptr(), prob.numel(), seed, offset, elem0, n, _ptr(step_dev), out.data_ptr(), _stream())
    return out


# ------------------------------------------------------------------------------------------------ K6 / K7
def nce_fwd(inp, target, noise, emb, bias, logq, norm_term: float, loss_type: str = "nce", grad_scale: Optional[float] = None,
            logits=None, ids_out=None, loss_pos=None, dz=None, d_input=None, acc_count=None, want_ids=True, want_d_input=True,
            shards=None):
    """inp [N,P]; target [N]; noise [N,K].  Returns (logits [N,K+1], ids [N,K+1] | None, loss_pos [N], dz [N,K+1], d_input | None)."""
    if loss_type not in _lib.NCE_LOSS:
        raise NotImplementedError(f"loss type {loss_type} not implemented")  # reference: nce_loss.py:126-132
    _check(inp, torch.float32, "input")
    _check(target, torch.int64, "target")
    _check(noise, torch.int64, "noise")
    N, P = inp.shape
    K = noise.shape[-1]
    dev = inp.device
    if logits is None:
        logits = torch.empty(N, K + 1, dtype=torch.float32, device=dev)
    if ids_out is None and want_ids:
        ids_out = torch.empty(N, K + 1, dtype=torch.int64, device=dev)
    if loss_pos is None:
        loss_pos = torch.empty(N, dtype=torch.float32, device=dev)
    if dz is None:
        dz = torch.empty(N, K + 1, dtype=torch.float32, device=dev)
    if d_input is None and want_d_input:
        d_input = torch.empty(N, P, dtype=torch.float32, device=dev)
    if grad_scale is None:
        grad_scale = 1.0 / max(N, 1)
    if _lib.PROFILE is not None or _lib.TIMELINE is not None or _lib.RECORD is not None:
        _lib.CURRENT_TAG = ("nce", N, K, P)
    if shards is not None:   # (emb_ptrs, bias_ptrs, R): row-sharded output tables in peer memory
        emb_ptrs, bias_ptrs, R = shards
        call("map_nce_fwd_sharded", inp.data_ptr(), N, P, K, target.data_ptr(), noise.data_ptr(), emb_ptrs, bias_ptrs, R,
             logq.data_ptr(), float(norm_term), _lib.NCE_LOSS[loss_type], float(grad_scale), logits.data_ptr(),
             _ptr(ids_out), loss_pos.data_ptr(), dz.data_ptr(), _ptr(d_input), _ptr(acc_count), _stream())
        return logits, ids_out, loss_pos, dz, d_input
    call("map_nce_fwd", inp.data_ptr(), N, P, K, target.data_ptr(), noise.data_ptr(), emb.data_ptr(), bias.data_ptr(),
         logq.data_ptr(), emb.shape[0], float(norm_term), _lib.NCE_LOSS[loss_type], float(grad_scale), logits.data_ptr(),
         _ptr(ids_out), loss_pos.data_ptr(), dz.data_ptr(), _ptr(d_input), _ptr(acc_count), _stream())
    return logits, ids_out, loss_pos, dz, d_input


def nce_ids_concat(target: torch.Tensor, noise: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """[target | noise] -> ids [N, K+1] (index_linear.py:79-83), the id list of the NCE tables' gradient"""
    _check(target, torch.int64, "target")
    _check(noise, torch.int64, "noise")
    N, K = noise.shape
    if out is None:
        out = torch.empty(N, K + 1, dtype=torch.int64, device=noise.device)
    call("map_nce_ids_concat", target.data_ptr(), noise.data_ptr(), N, K, out.data_ptr(), _stream())
    return out


def gather_slices(enc: torch.Tensor, masked_index: torch.Tensor, F: int, P: int, out=None) -> torch.Tensor:
    B, L = masked_index.shape
    if out is None:
        out = torch.empty(B * L, P, dtype=torch.float32, device=enc.device)
    call("map_gather_slices", enc.data_ptr(), masked_index.data_ptr(), B * L, L, F, P, out.data_ptr(), _stream())
    return out


def scatter_add_slices(d_sel: torch.Tensor, masked_index: torch.Tensor, F: int, P: int, d_enc: torch.Tensor):
    B, L = masked_index.shape
    call("map_scatter_add_slices", d_sel.data_ptr(), masked_index.data_ptr(), B * L, L, F, P, d_enc.data_ptr(), _stream())
    return d_enc


def expand_slices(d_sel: torch.Tensor, masked_index: torch.Tensor, F: int, P: int, d_enc: torch.Tensor, planes: Optional[torch.Tensor] = None):
    """d_enc [B, F*P] <- dense form of the slice scatter (every element written once; no memset, no atomics) + its bf16 planes"""
    B, L = masked_index.shape
    pl = (planes.data_ptr(), planes.stride(1), planes.stride(0), planes.shape[0]) if planes is not None else (None, 0, 0, 0)
    call("map_expand_slices", d_sel.data_ptr(), masked_index.data_ptr(), B, L, F, P, d_enc.data_ptr(), _ld(d_enc), *pl, _stream())
    return d_enc


# ------------------------------------------------------------------------------------------------ K16: by-field MFP encoder
def field_enc_supported(F: int, P: int) -> bool:
    return bool(_lib.load().map_field_enc_supported(int(F), int(P)))


def field_bucket(masked_index: torch.Tensor, F: int, perm: Optional[torch.Tensor] = None, fstart: Optional[torch.Tensor] = None):
    """stable counting sort of the masked positions n = b*L + l by field: perm [N] int32 (slot -> n), fstart [F+1] int32"""
    N = masked_index.numel()
    dev = masked_index.device
    if perm is None:
        perm = torch.empty(N, dtype=torch.int32, device=dev)
    if fstart is None:
        fstart = torch.empty(F + 1, dtype=torch.int32, device=dev)
    _check(masked_index, torch.int64, "masked_index")
    call("map_field_bucket", masked_index.data_ptr(), N, F, perm.data_ptr(), fstart.data_ptr(), _stream())
    return perm, fstart


def field_enc_fwd(X: torch.Tensor, K: int, W: torch.Tensor, bias: torch.Tensor, perm, fstart, L: int, F: int, P: int,
                  out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """sel[n, :] = X[n // L, :K] @ W[f(n)*P:(f(n)+1)*P, :K].T + bias[f(n)*P:...]   (models.py:73-78 for the masked slices only)"""
    N = perm.numel()
    if out is None:
        out = torch.empty(N, P, dtype=torch.float32, device=X.device)
    _lib.CURRENT_TAG = ("field_enc_fwd", N, P, K)
    call("map_field_enc_fwd", X.data_ptr(), _ld(X), K, W.data_ptr(), _ld(W), bias.data_ptr(), perm.data_ptr(), fstart.data_ptr(),
         N, L, F, P, out.data_ptr(), _stream())
    return out


def field_enc_dgrad(d_sel: torch.Tensor, W: torch.Tensor, K: int, perm, fstart, F: int, P: int, dxpos: torch.Tensor) -> torch.Tensor:
    """dxpos[n, :K] = d_sel[n, :] @ W[f(n)*P:(f(n)+1)*P, :K]"""
    N = perm.numel()
    _lib.CURRENT_TAG = ("field_enc_dgrad", N, P, K)
    call("map_field_enc_dgrad", d_sel.data_ptr(), W.data_ptr(), _ld(W), K, perm.data_ptr(), fstart.data_ptr(), N, F, P,
         dxpos.data_ptr(), _ld(dxpos), _stream())
    return dxpos


def field_enc_wgrad(d_sel: torch.Tensor, X: torch.Tensor, K: int, perm, fstart, L: int, F: int, P: int, dW: torch.Tensor,
                    dbias: Optional[torch.Tensor] = None) -> torch.Tensor:
    """dW[f*P + j, :K] = sum over the positions of field f of d_sel[n, j] * X[n // L, :K]; dbias[f*P + j] = sum d_sel[n, j]"""
    N = perm.numel()
    _lib.CURRENT_TAG = ("field_enc_wgrad", N, P, K)
    call("map_field_enc_wgrad", d_sel.data_ptr(), X.data_ptr(), _ld(X), K, perm.data_ptr(), fstart.data_ptr(), N, L, F, P,
         dW.data_ptr(), _ld(dW), _ptr(dbias), _stream())
    return dW


def head_bwd_fold(dxpos: torch.Tensor, B: int, L: int, ncols: int, *, cross=None, relu=None, scalar=None):
    """Sum of the L per-position gradient rows of every sample + the first backward stage of the towers (csrc/fieldenc.cu K16d).
    cross = dict(col0, width, x0, u, g_out, du_out, dx0_out, du_planes=None, bias_grad=None)
    relu  = dict(col0, width, y, dz_out, dz_planes=None, bias_grad=None);  scalar = dict(col, out)"""
    import ctypes as C
    a = _lib.HeadBwdArgs()
    a.dxpos, a.ld_dx, a.B, a.L, a.ncols = dxpos.data_ptr(), _ld(dxpos), B, L, ncols
    a.scalar_col = -1
    keep = [dxpos]
    if cross is not None:
        a.cross_col0, a.cross_w = cross["col0"], cross["width"]
        for k_, t_ in (("x0", "x0"), ("u", "u"), ("g_out", "g"), ("du_out", "du"), ("dx0_out", "dx0")):
            setattr(a, k_, cross[k_].data_ptr())
            setattr(a, "ld_" + t_, _ld(cross[k_]))
        pl = cross.get("du_planes")
        if pl is not None:
            a.du_planes, a.du_pl_ld, a.du_pl_stride, a.du_nplanes = pl.data_ptr(), pl.stride(1), pl.stride(0), pl.shape[0]
        if cross.get("bias_grad") is not None:
            a.cross_bias_grad = cross["bias_grad"].data_ptr()
        keep.append(cross)
    if relu is not None:
        a.relu_col0, a.relu_w = relu["col0"], relu["width"]
        a.y, a.ld_y = relu["y"].data_ptr(), _ld(relu["y"])
        a.dz_out, a.ld_dz = relu["dz_out"].data_ptr(), _ld(relu["dz_out"])
        pl = relu.get("dz_planes")
        if pl is not None:
            a.dz_planes, a.dz_pl_ld, a.dz_pl_stride, a.dz_nplanes = pl.data_ptr(), pl.stride(1), pl.stride(0), pl.shape[0]
        if relu.get("bias_grad") is not None:
            a.relu_bias_grad = relu["bias_grad"].data_ptr()
        keep.append(relu)
    if scalar is not None:
        a.scalar_col, a.scalar_out, a.ld_scalar = scalar["col"], scalar["out"].data_ptr(), _ld(scalar["out"])
        keep.append(scalar)
    _lib.CURRENT_TAG = ("head_bwd_fold", B, L, ncols)
    call("map_head_bwd_fold", C.byref(a), _stream())


# ------------------------------------------------------------------------------------------------ K17: CIN (xDeepFM)
def cin_relayout(src: torch.Tensor, dst: torch.Tensor, B: int, C: int, D: int, to_pairs: bool, accumulate: bool = False):
    """to_pairs: src [B, C, D] -> dst [(b*D + d), c];  else: src pairs -> dst [B, C, D] (optionally accumulating)"""
    pairs = dst if to_pairs else src
    _lib.CURRENT_TAG = ("cin_relayout", B, C, D)
    call("map_cin_relayout", src.data_ptr(), dst.data_ptr(), B, C, D, _ld(pairs), int(to_pairs), int(accumulate), _stream())
    return dst


def cin_hadamard_fwd(x0: torch.Tensor, xi: torch.Tensor, F: int, M: int, z: torch.Tensor) -> torch.Tensor:
    """z[p, h*M + m] = x0[p, h] * xi[p, m] (layers.py:714-715), zero in the padding columns of z"""
    _lib.CURRENT_TAG = ("cin_hadamard", x0.shape[0], F, M)
    call("map_cin_hadamard_fwd", x0.data_ptr(), _ld(x0), xi.data_ptr(), _ld(xi), x0.shape[0], F, M, z.data_ptr(), _ld(z), _stream())
    return z


def cin_hadamard_bwd(dz: torch.Tensor, x0: torch.Tensor, xi: torch.Tensor, F: int, M: int, dx0: torch.Tensor, dxi: torch.Tensor,
                     accumulate_dx0: bool = False):
    _lib.CURRENT_TAG = ("cin_hadamard_bwd", x0.shape[0], F, M)
    call("map_cin_hadamard_bwd", dz.data_ptr(), _ld(dz), x0.data_ptr(), _ld(x0), xi.data_ptr(), _ld(xi), x0.shape[0], F, M,
         dx0.data_ptr(), _ld(dx0), int(accumulate_dx0), dxi.data_ptr(), _ld(dxi), _stream())


def cin_pool_fwd(y: torch.Tensor, B: int, D: int, O: int, pooled: torch.Tensor) -> torch.Tensor:
    call("map_cin_pool_fwd", y.data_ptr(), _ld(y), B, D, O, pooled.data_ptr(), _ld(pooled), _stream())
    return pooled


def cin_pool_bwd(d_pooled: torch.Tensor, B: int, D: int, O: int, dy: torch.Tensor, d_next: Optional[torch.Tensor] = None) -> torch.Tensor:
    call("map_cin_pool_bwd", d_pooled.data_ptr(), _ld(d_pooled), _ptr(d_next), _ld(d_next) if d_next is not None else 0, B, D, O,
         dy.data_ptr(), _ld(dy), _stream())
    return dy


_red_ws = {}


def _reduce_ws(device):
    ws = _red_ws.get(device)
    if ws is None:
        ws = torch.empty(int(_lib.load().map_reduce_workspace_bytes(0)), dtype=torch.uint8, device=device)
        _red_ws[device] = ws
    return ws


def reduce_sum(x: torch.Tensor, scale: float = 1.0, out=None, ws=None) -> torch.Tensor:
    _check(x, torch.float32, "x")
    if out is None:
        out = torch.empty(1, dtype=torch.float32, device=x.device)
    ws = ws if ws is not None else _reduce_ws(x.device)
    call("map_reduce_sum_f32", x.data_ptr(), x.numel(), float(scale), out.data_ptr(), ws.data_ptr(), ws.numel(), _stream())
    return out


def bce_logits(logits: torch.Tensor, labels: torch.Tensor, stats=None, dlogits=None, ws=None, want_grad=True):
    """stats = [mean loss, #correct, sum(labels), n]; dlogits = (sigmoid(z)-y)/n."""
    _check(logits, torch.float32, "logits")
    _check(labels, torch.float32, "labels")
    n = logits.numel()
    if stats is None:
        stats = torch.empty(4, dtype=torch.float32, device=logits.device)
    if dlogits is None and want_grad:
        dlogits = torch.empty_like(logits)
    ws = ws if ws is not None else _reduce_ws(logits.device)
    call("map_bce_logits_fwd", logits.data_ptr(), labels.data_ptr(), n, stats.data_ptr(), _ptr(dlogits), ws.data_ptr(), ws.numel(), _stream())
    return stats, dlogits


def auc_logloss(logits: torch.Tensor, labels: torch.Tensor):
    """ROC AUC (ties share their average rank, like sklearn.metrics.roc_auc_score) and mean BCE-with-logits of n scores, on
    the device: scores -> sortable keys -> sort-dedup (K2a) -> positives per tie group (K2b) -> rank sum.  Returns two 0-d
    float64 device tensors; nothing synchronises.  (trainer.py:183-197 copies logits to the host and calls sklearn.)"""
    logits = logits.reshape(-1).contiguous().float()
    labels = labels.reshape(-1).contiguous().float()
    n = logits.numel()
    dev = logits.device
    keys = torch.empty(n, dtype=torch.int64, device=dev)
    call("map_float_sort_keys", logits.data_ptr(), n, keys.data_ptr(), _stream())
    plan = DedupPlan(n, 1 << 32, dev).run(keys)
    pos = plan.reduce_rows(labels.view(n, 1), 1)
    out = torch.empty(2, dtype=torch.float64, device=dev)
    call("map_auc_rank_sum", plan.seg_start.data_ptr(), pos.data_ptr(), plan.n_unique.data_ptr(), n, out.data_ptr(), _stream())
    P = out[1]
    auc = (out[0] - P * (P + 1) / 2) / (P * (n - P))
    stats, _ = bce_logits(logits, labels, want_grad=False)
    return auc, stats[0].double()


def nce_full_ce(inp: torch.Tensor, emb: torch.Tensor, bias: torch.Tensor, target: torch.Tensor, out: Optional[torch.Tensor] = None):
    """per-position full-softmax cross entropy (IndexLinear.ce_loss, index_linear.py:145-151): inp [N,P], emb [V,P], bias [V] or [V,1],
    target [N] -> [N].  The [N, V] scores stay on chip (online log-sum-exp over vocabulary slices)."""
    _check(inp, torch.float32, "input")
    _check(emb, torch.float32, "emb")
    _check(target, torch.int64, "target")
    N, P = inp.shape
    V = emb.shape[0]
    bias = bias.reshape(-1)
    if out is None:
        out = torch.empty(N, dtype=torch.float32, device=inp.device)
    lib = _lib.load()
    ws_bytes = int(lib.map_nce_full_ce_workspace_bytes(N, V))
    ws = torch.empty(max(ws_bytes, 8), dtype=torch.uint8, device=inp.device)
    call("map_nce_full_ce", inp.data_ptr(), N, P, emb.data_ptr(), bias.data_ptr(), V, target.data_ptr(), out.data_ptr(), ws.data_ptr(), ws_bytes,
         _stream())
    return out


def fm_lr_fwd(feat_embed, ids, lr_w, lr_bias, out=None, ld_out: int = 1):
    B, F, D = feat_embed.shape
    if out is None:
        out = torch.empty(B, 1, dtype=torch.float32, device=feat_embed.device)
    call("map_fm_lr_fwd", feat_embed.data_ptr(), _ptr(ids), lr_w.data_ptr(), lr_bias.data_ptr(), B, F, D, out.data_ptr(), ld_out, _stream())
    return out


def fm_lr_bwd(feat_embed, g, ld_g: int, d_embed, d_w_occ=None, accumulate=False):
    B, F, D = feat_embed.shape
    call("map_fm_lr_bwd", feat_embed.data_ptr(), g.data_ptr(), ld_g, B, F, D, 1 if accumulate else 0, d_embed.data_ptr(), _ptr(d_w_occ), _stream())
    return d_embed


# ------------------------------------------------------------------------------------------------ GEMM
def gemm_backend() -> str:
    """'bf16s' (default): split-bf16 tensor-core GEMM with fp32-level accuracy (gemm_bf16s.cu);
    'tf32': single-pass TF32 tensor cores (gemm_tcgen05.cu; faster, ~1e-3 biased per product — does NOT meet the gradient
    tolerance of north_star on this ReLU network, kept for A/B measurements);
    'simt': exact fp32 on CUDA cores (tests).  MAP_B200_GEMM selects; 'tcgen05' is accepted as an alias of 'tf32'."""
    b = os.environ.get("MAP_B200_GEMM", "bf16s")
    return "tf32" if b == "tcgen05" else b


def _norm_backend(backend: Optional[str]) -> str:
    backend = backend or gemm_backend()
    return "tf32" if backend == "tcgen05" else backend


def _ld(t: torch.Tensor) -> int:
    assert t.dim() == 2 and (t.stride(1) == 1 or t.shape[1] == 1), "matrix operands must be row-major with unit inner stride"
    return t.stride(0) if t.shape[0] > 1 else max(t.shape[1], 1)


def _gemm_args(A: torch.Tensor, B: torch.Tensor, C_out: torch.Tensor, M: int, N: int, K: int, trans_a=False, trans_b=False,
               epilogue: int = _lib.EPI_NONE, bias=None, aux0=None, aux1=None, aux_out=None, aux2=None, acc_out=None,
               acc_accumulate: bool = False, colsum_out=None, g: Optional[GemmArgs] = None, **_ignored) -> GemmArgs:
    g = GemmArgs() if g is None else g
    g.M, g.N, g.K = M, N, K
    g.trans_a, g.trans_b, g.epilogue = int(trans_a), int(trans_b), int(epilogue)
    g.A, g.lda = (A.data_ptr(), _ld(A)) if A is not None else (None, 0)
    g.B, g.ldb = (B.data_ptr(), _ld(B)) if B is not None else (None, 0)
    g.C, g.ldc = C_out.data_ptr(), _ld(C_out)
    g.bias = _ptr(bias)
    g.aux0, g.ld_aux0 = (aux0.data_ptr(), _ld(aux0)) if aux0 is not None else (None, 0)
    g.aux1, g.ld_aux1 = (aux1.data_ptr(), _ld(aux1)) if aux1 is not None else (None, 0)
    g.aux_out, g.ld_aux_out = (aux_out.data_ptr(), _ld(aux_out)) if aux_out is not None else (None, 0)
    g.aux2, g.ld_aux2 = (aux2.data_ptr(), _ld(aux2)) if aux2 is not None else (None, 0)
    g.acc_out, g.ld_acc_out = (acc_out.data_ptr(), _ld(acc_out)) if acc_out is not None else (None, 0)
    g.acc_accumulate = int(bool(acc_accumulate))
    g.colsum_out = _ptr(colsum_out)
    return g


# ---- bf16 planes (operand format of the split-bf16 GEMM)
def alloc_planes(rows: int, cols: int, n_planes: int, device) -> torch.Tensor:
    """[n_planes, rows, cols] view (bf16) of storage whose row stride is a multiple of 8 elements (16-byte TMA rows)"""
    ld = (cols + 7) // 8 * 8
    return torch.zeros(n_planes, rows, ld, dtype=torch.bfloat16, device=device)[:, :, :cols]


def split_planes(src: torch.Tensor, planes: torch.Tensor) -> torch.Tensor:
    """planes[i] <- i-th bf16 piece of the fp32 matrix `src` (hi, lo, lo2)"""
    rows, cols = src.shape
    assert planes.dtype == torch.bfloat16 and planes.dim() == 3 and tuple(planes.shape[1:]) == (rows, cols) and planes.stride(2) == 1
    if cols % 4 != 0 or _ld(src) % 4 != 0:
        raise ValueError("split_planes: cols and row stride must be multiples of 4")
    call("map_split_bf16", src.data_ptr(), _ld(src), rows, cols, planes.data_ptr(), planes.stride(1), planes.stride(0), planes.shape[0], _stream())
    return planes


def _planes_of(t: torch.Tensor, n_planes: int) -> torch.Tensor:
    """temporary planes of an fp32 matrix nobody keeps planes for (module path, tests): one extra pass over the operand"""
    rows, cols = t.shape
    if cols % 4 != 0 or _ld(t) % 4 != 0 or t.data_ptr() % 16 != 0:   # odd shapes: go through an aligned copy
        cp = (cols + 3) // 4 * 4
        tmp = torch.zeros(rows, cp, dtype=torch.float32, device=t.device)
        copy2d(t, tmp[:, :cols])
        pl = alloc_planes(rows, cp, n_planes, t.device)
        split_planes(tmp, pl)
        return pl[:, :, :cols]
    return split_planes(t, alloc_planes(rows, cols, n_planes, t.device))


SIMT_GEMMS = {}   # (M, N, K) -> count of GEMMs that ran on the exact-fp32 CUDA-core kernel although a tensor-core backend was selected


def bf16s_supported(pr) -> bool:
    """shapes the split-bf16 kernel takes: everything except skinny problems (fc_out: N = 1 or K = 1)"""
    return pr["N"] % 4 == 0 and pr["N"] >= 8 and pr["K"] >= 8 and pr["M"] >= 2


def _split_args(pr) -> GemmSplitArgs:
    terms = int(pr.get("terms", 3))
    need = 3 if terms == 6 else 2
    a = GemmSplitArgs()
    _gemm_args(g=a.g, **{k: v for k, v in pr.items() if k not in ("Ap", "Bp", "Cp", "terms")})
    Ap = pr.get("Ap")
    Bp = pr.get("Bp")
    if Ap is None:
        Ap = _planes_of(pr["A"], need)
    if Bp is None:
        Bp = _planes_of(pr["B"], need)
    for nm, P in (("A", Ap), ("B", Bp)):
        if P.dtype != torch.bfloat16 or P.dim() != 3 or P.stride(2) != 1 or P.shape[0] < need:
            raise ValueError(f"gemm: planes of {nm} must be a [>= {need}, rows, cols] bf16 view with unit inner stride")
    a.a_planes, a.a_ld, a.a_plane_stride, a.a_nplanes, a.terms = Ap.data_ptr(), Ap.stride(1), Ap.stride(0), Ap.shape[0], terms
    a.b_planes, a.b_ld, a.b_plane_stride, a.b_nplanes = Bp.data_ptr(), Bp.stride(1), Bp.stride(0), Bp.shape[0]
    Cp = pr.get("Cp")
    if Cp is not None:
        a.c_planes, a.c_ld, a.c_plane_stride, a.c_nplanes = Cp.data_ptr(), Cp.stride(1), Cp.stride(0), Cp.shape[0]
    a._keep = (Ap, Bp, Cp)
    return a


def gemm(A: torch.Tensor, B: torch.Tensor, C_out: torch.Tensor, M: int, N: int, K: int, trans_a=False, trans_b=False,
         epilogue: int = _lib.EPI_NONE, bias=None, aux0=None, aux1=None, aux_out=None, aux2=None, acc_out=None,
         acc_accumulate: bool = False, colsum_out=None, backend: Optional[str] = None, Ap=None, Bp=None, Cp=None, terms: int = 3):
    """acc[m,n] = sum_k A[m,k]*B[n,k] with storage transposes; see include/map_b200.h.  Operands are 2-D row-major views
    (row stride = leading dimension, so column slices of wider buffers work).  Ap / Bp: bf16 planes of A / B if the caller
    keeps them (otherwise they are split on the fly); Cp: planes of C to be written; terms: 3 or 6 (split-bf16 backend)."""
    gemm_group([dict(A=A, B=B, C_out=C_out, M=M, N=N, K=K, trans_a=trans_a, trans_b=trans_b, epilogue=epilogue, bias=bias, aux0=aux0,
                     aux1=aux1, aux_out=aux_out, aux2=aux2, acc_out=acc_out, acc_accumulate=acc_accumulate, colsum_out=colsum_out,
                     Ap=Ap, Bp=Bp, Cp=Cp, terms=terms)], backend=backend)
    return C_out


GEMM_GROUP_MAX = 4


def _tag(*t):
    if _lib.PROFILE is not None or _lib.TIMELINE is not None or _lib.RECORD is not None:
        _lib.CURRENT_TAG = t


def _gemm_simt(pr, count_fallback: bool):
    g = _gemm_args(**pr)
    if count_fallback:
        key = (pr["M"], pr["N"], pr["K"])
        SIMT_GEMMS[key] = SIMT_GEMMS.get(key, 0) + 1
    _tag("gemm", pr["M"], pr["N"], pr["K"], int(bool(pr.get("trans_a"))), int(bool(pr.get("trans_b"))), int(pr.get("epilogue", 0)))
    call("map_gemm_f32_simt", C.byref(g), _stream())
    if pr.get("Cp") is not None:   # consumers on the tensor-core path read planes
        split_planes(pr["C_out"], pr["Cp"])


def gemm_group(problems, backend: Optional[str] = None):
    """problems: list of dicts with the keyword arguments of `gemm` (A, B, C_out, M, N, K, ...), all INDEPENDENT of each other.
    Tensor-core backends put up to 4 problems into one grouped launch; skinny problems (N = 1 / K = 1: fc_out) run on the
    exact-fp32 CUDA-core kernel and are COUNTED in ops.SIMT_GEMMS (bench.py reports the count: no silent fallback)."""
    backend = _norm_backend(backend)
    lib = _lib.load()
    problems = [pr for pr in problems if pr is not None]
    if backend == "simt":
        for pr in problems:
            _gemm_simt(pr, False)
        return
    if backend == "bf16s":
        grouped = []
        for pr in problems:
            if bf16s_supported(pr):
                grouped.append(pr)
            else:
                _gemm_simt(pr, True)
        for i in range(0, len(grouped), GEMM_GROUP_MAX):
            chunk = grouped[i:i + GEMM_GROUP_MAX]
            arr = (GemmSplitArgs * len(chunk))()
            keep = []
            for j, pr in enumerate(chunk):
                a = _split_args(pr)
                keep.append(a)
                C.memmove(C.byref(arr, j * C.sizeof(GemmSplitArgs)), C.byref(a), C.sizeof(GemmSplitArgs))
            _tag("gemm_group", *[(pr["M"], pr["N"], pr["K"], int(pr.get("terms", 3))) for pr in chunk])
            call("map_gemm_bf16s_group", arr, len(chunk), _stream())
        return
    if backend != "tf32":
        raise ValueError(f"unknown GEMM backend {backend!r}")
    grouped = []
    for pr in problems:
        g = _gemm_args(**pr)
        if lib.map_gemm_tf32_supported(C.byref(g)):
            grouped.append((g, pr))
        else:
            _gemm_simt(pr, True)
    for i in range(0, len(grouped), GEMM_GROUP_MAX):
        chunk = grouped[i:i + GEMM_GROUP_MAX]
        arr = (GemmArgs * len(chunk))()
        for j, (g, _) in enumerate(chunk):
            C.memmove(C.byref(arr, j * C.sizeof(GemmArgs)), C.byref(g), C.sizeof(GemmArgs))
        _tag("gemm_group", *[(g.M, g.N, g.K, 1) for g, _ in chunk])
        call("map_gemm_tf32_group", arr, len(chunk), _stream())
        for _, pr in chunk:
            if pr.get("Cp") is not None:
                split_planes(pr["C_out"], pr["Cp"])


_colsum_ws = {}


def colsum(X: torch.Tensor, out: Optional[torch.Tensor] = None, ws=None) -> torch.Tensor:
    M, N = X.shape
    if out is None:
        out = torch.empty(N, dtype=torch.float32, device=X.device)
    need = int(_lib.load().map_colsum_workspace_bytes(M, N))
    if ws is None:
        key = (X.device, need)
        ws = _colsum_ws.get(key)
        if ws is None:
            ws = torch.empty(need, dtype=torch.uint8, device=X.device)
            _colsum_ws[key] = ws
    call("map_colsum_f32", X.data_ptr(), _ld(X), M, N, out.data_ptr(), ws.data_ptr(), ws.numel(), _stream())
    return out


def cross_bwd_pre(G, X0, U, dU, dX0_acc, accumulate: bool):
    M, N = G.shape
    call("map_cross_bwd_pre", G.data_ptr(), _ld(G), X0.data_ptr(), _ld(X0), U.data_ptr(), _ld(U), M, N, 1 if accumulate else 0,
         dU.data_ptr(), dX0_acc.data_ptr(), _stream())


def add3(a, b, c, out):
    M, N = a.shape
    call("map_add3_f32", a.data_ptr(), _ld(a), b.data_ptr(), _ld(b), _ptr(c), _ld(c) if c is not None else 0, M, N, out.data_ptr(), _ld(out), _stream())
    return out


def transpose(x: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    M, N = x.shape
    if out is None:
        out = torch.empty(N, M, dtype=torch.float32, device=x.device)
    call("map_transpose_f32", x.data_ptr(), _ld(x), M, N, out.data_ptr(), _ld(out), _stream())
    return out


def relu_bwd(dy: torch.Tensor, y: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    M, N = dy.shape
    if out is None:
        out = torch.empty(M, N, dtype=torch.float32, device=dy.device)
    call("map_relu_bwd_f32", dy.data_ptr(), _ld(dy), y.data_ptr(), _ld(y), M, N, out.data_ptr(), _ld(out), _stream())
    return out


def scale_by_scalar(x: torch.Tensor, scalar_dev: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    _check(x, torch.float32, "x")
    if out is None:
        out = torch.empty_like(x)
    call("map_scale_by_scalar_f32", x.data_ptr(), scalar_dev.data_ptr(), x.numel(), out.data_ptr(), _stream())
    return out


def copy2d(src: torch.Tensor, dst: torch.Tensor) -> torch.Tensor:
    M, N = src.shape
    call("map_copy2d_f32", src.data_ptr(), _ld(src), M, N, dst.data_ptr(), _ld(dst), _stream())
    return dst


def gather_rows_i64(X: torch.Tensor, idx: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    _check(X, torch.int64, "X")
    _check(idx, torch.int64, "idx")
    n_rows, F = X.shape
    if out is None:
        out = torch.empty(idx.numel(), F, dtype=torch.int64, device=X.device)
    call("map_gather_rows_i64", X.data_ptr(), n_rows, F, idx.data_ptr(), idx.numel(), out.data_ptr(), _stream())
    return out


# ------------------------------------------------------------------------------------------------ K13 row-sharded tables
def emb_gather_owned(shard: torch.Tensor, ids: torch.Tensor, R: int, rank: int, out: torch.Tensor) -> torch.Tensor:
    rows, D = shard.shape
    call("map_emb_gather_owned_f32", shard.data_ptr(), rows, D, ids.data_ptr(), ids.numel(), R, rank, out.data_ptr(), _stream())
    return out


def owned_keys(ids: torch.Tensor, R: int, rank: int, sentinel: int, out: torch.Tensor) -> torch.Tensor:
    call("map_owned_keys", ids.data_ptr(), ids.numel(), R, rank, sentinel, out.data_ptr(), _stream())
    return out


def nce_scores_owned(q, idx, emb_shard, bias_shard, R: int, rank: int, out):
    N, P = q.shape
    call("map_nce_scores_owned", q.data_ptr(), N, P, idx.shape[1], idx.data_ptr(), emb_shard.data_ptr(), bias_shard.data_ptr(), R, rank,
         out.data_ptr(), _stream())
    return out


def nce_loss_from_scores(scores, idx, logq, norm_term: float, loss_type: str, grad_scale: float, logits, loss_pos, dz, acc_count=None):
    N, K1 = scores.shape
    call("map_nce_loss_from_scores", scores.data_ptr(), idx.data_ptr(), N, K1, logq.data_ptr(), float(norm_term), _lib.NCE_LOSS[loss_type],
         float(grad_scale), logits.data_ptr(), loss_pos.data_ptr(), dz.data_ptr(), _ptr(acc_count), _stream())


def nce_dinput_owned(dz, idx, emb_shard, R: int, rank: int, out):
    N, K1 = dz.shape
    P = emb_shard.shape[1]
    call("map_nce_dinput_owned", dz.data_ptr(), N, P, K1, idx.data_ptr(), emb_shard.data_ptr(), R, rank, out.data_ptr(), _stream())
    return out
