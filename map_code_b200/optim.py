"""AdamW with transformers==4.26.1 semantics (the optimizer the reference constructs at code/trainer.py:75-76) as fused
kernels: one multi-tensor launch for all dense parameters, row-wise updates for [V, .] tables whose gradient arrives
deduplicated ("sparse" TableGrad mode), or the exact dense sweep.  Works with torch LR schedulers (reads group["lr"])."""
from __future__ import annotations

import torch

from . import ops
from .functional import TableGrad


class AdamW(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-6, weight_decay=0.0, correct_bias=True,
                 no_deprecation_warning=False, table_mode="sparse"):
        if lr < 0.0:
            raise ValueError(f"Invalid learning rate: {lr} - should be >= 0.0")
        if not 0.0 <= betas[0] < 1.0 or not 0.0 <= betas[1] < 1.0:
            raise ValueError(f"Invalid beta parameters: {betas}")
        if not 0.0 <= eps:
            raise ValueError(f"Invalid epsilon value: {eps} - should be >= 0.0")
        if not correct_bias:
            raise NotImplementedError("correct_bias=False is not used by the reference (trainer.py:75)")
        if table_mode not in ("sparse", "dense_exact"):
            raise ValueError(table_mode)
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        self.table_mode = table_mode
        self._t = 0
        self._hyper = {}
        self._cache = {}

    def _state(self, p):
        st = self.state[p]
        if len(st) == 0:
            st["exp_avg"] = torch.zeros_like(p.data)
            st["exp_avg_sq"] = torch.zeros_like(p.data)
        return st

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        self._t += 1
        for gi, group in enumerate(self.param_groups):
            dense, tables = [], []
            for p in group["params"]:
                tg: TableGrad = getattr(p, "_map_table_grad", None)
                if tg is not None and tg.mode == "sparse":
                    if tg.plan is not None:
                        tables.append((p, tg))
                    continue
                if p.grad is None:
                    continue
                if p.grad.is_sparse:
                    raise RuntimeError("Adam does not support sparse gradients, please consider SparseAdam instead")
                dense.append(p)
            if not dense and not tables:
                continue
            dev = (dense[0] if dense else tables[0][0]).device
            hyper = self._hyper.get((gi, dev))
            if hyper is None:
                hyper = torch.zeros(8, dtype=torch.float32, device=dev)
                self._hyper[(gi, dev)] = hyper
            b1, b2 = group["betas"]
            ops.adamw_hyper_set(hyper, group["lr"], b1, b2, group["eps"], self._t)
            if dense:
                key = tuple((p.data_ptr(), p.grad.data_ptr()) for p in dense)
                cached = self._cache.get(gi)
                if cached is None or cached[0] != key:
                    entries = []
                    for p in dense:
                        st = self._state(p)
                        g = p.grad if p.grad.is_contiguous() else p.grad.contiguous()
                        entries.append((p.data, g, st["exp_avg"], st["exp_avg_sq"], group["weight_decay"], None))
                    cached = (key,) + ops.make_adamw_tensor_list(entries, dev)
                    self._cache[gi] = cached
                ops.adamw_multi_tensor(cached[1], cached[2], cached[3], hyper)
            for p, tg in tables:
                st = self._state(p)
                if self.table_mode == "sparse":
                    ops.adamw_sparse_rows(p.data, st["exp_avg"], st["exp_avg_sq"], tg.plan, tg.grad_compact, hyper, group["weight_decay"])
                else:
                    ops.adamw_dense_rows_sparse_grad(p.data, st["exp_avg"], st["exp_avg_sq"], tg.plan, tg.grad_compact, hyper,
                                                     group["weight_decay"])
                tg.clear()
        return loss

    def zero_grad(self, set_to_none: bool = True):
        super().zero_grad(set_to_none)
        for group in self.param_groups:
            for p in group["params"]:
                tg = getattr(p, "_map_table_grad", None)
                if tg is not None:
                    tg.clear()
