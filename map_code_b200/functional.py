"""torch.autograd.Function wrappers: the module-level API (layers / nce / models) differentiates through these.

Forward and backward are both our kernels (ops.py -> C ABI).  Table parameters ([V, D] embeddings) have two gradient
modes, selected per table through a `TableGrad` handle:
  * "dense"  — `.grad` is a dense [V, D] tensor exactly like the reference's nn.Embedding(sparse=False) (parity tests,
               interop with any torch optimizer);
  * "sparse" — the deduplicated compact gradient (unique ids + [U, D] rows) is stashed on the handle and consumed by
               map_code_b200.optim.AdamW with a row-wise update; no [V, D] gradient is ever materialised.
"""
from __future__ import annotations

from typing import Optional

import torch

from . import _lib, ops


class TableGrad:
    """Per-table gradient plumbing: owns the dedup plans (one per id-stream length) and the last compact gradient."""

    def __init__(self, mode: str = "dense"):
        assert mode in ("dense", "sparse")
        self.mode = mode
        self._plans = {}
        self.plan: Optional[ops.DedupPlan] = None   # plan of the most recent backward
        self.grad_compact: Optional[torch.Tensor] = None

    def plan_for(self, n: int, V: int, device) -> ops.DedupPlan:
        key = (n, V, str(device))
        p = self._plans.get(key)
        if p is None:
            p = ops.DedupPlan(n, V, device)
            self._plans[key] = p
        return p

    def clear(self):
        self.plan, self.grad_compact = None, None

    def finish(self, plan: ops.DedupPlan, G: torch.Tensor, shape) -> Optional[torch.Tensor]:
        """Returns the tensor autograd should see for the table (dense grad or None)."""
        if self.mode == "dense":
            V, D = shape
            dense = torch.zeros(V, D, dtype=torch.float32, device=G.device)
            plan.scatter_dense(G, D, dense)
            return dense
        if self.plan is not None:
            raise _lib.MapB200Error("sparse TableGrad: a table may receive one gradient per step (call optimizer.step() / zero_grad())")
        self.plan, self.grad_compact = plan, G
        return None


def _c(t: torch.Tensor) -> torch.Tensor:
    return t if t.is_contiguous() else t.contiguous()


def _rows(t: torch.Tensor) -> torch.Tensor:
    """2-D row-major view with unit inner stride (copy only when the layout forces it)."""
    return t if (t.dim() == 2 and (t.stride(1) == 1 or t.shape[1] == 1)) else t.contiguous()


# ------------------------------------------------------------------------------------------------ embedding
class EmbeddingFn(torch.autograd.Function):
    """reference: nn.Embedding forward/backward behind code/layers.py:98 and code/models.py:139"""

    @staticmethod
    def forward(ctx, weight, ids, tg: TableGrad):
        ids = _c(ids)
        ctx.save_for_backward(ids)
        ctx.tg, ctx.wshape = tg, tuple(weight.shape)
        return ops.emb_gather(weight, ids)

    @staticmethod
    def backward(ctx, gout):
        (ids,) = ctx.saved_tensors
        V, D = ctx.wshape
        n = ids.numel()
        if n == 0:
            return (torch.zeros(V, D, device=gout.device) if ctx.tg.mode == "dense" else None), None, None
        plan = ctx.tg.plan_for(n, V, gout.device).run(ids.view(-1))
        G = plan.reduce_rows(_c(gout).view(n, D), D)
        return ctx.tg.finish(plan, G, (V, D)), None, None


# ------------------------------------------------------------------------------------------------ linear layers
class LinearFn(torch.autograd.Function):
    """y = act(x W^T + b), act in {none, relu}.  reference: nn.Linear (+ nn.ReLU) in code/layers.py:179-188,
    code/models.py:116-123,304"""

    @staticmethod
    def forward(ctx, x, weight, bias, relu: bool):
        x2 = _rows(x)
        M, K = x2.shape
        N = weight.shape[0]
        y = torch.empty(M, N, dtype=torch.float32, device=x.device)
        epi = _lib.EPI_BIAS_RELU if relu else (_lib.EPI_BIAS if bias is not None else _lib.EPI_NONE)
        ops.gemm(x2, weight, y, M, N, K, epilogue=epi, bias=bias, terms=6 if relu else 3)   # a ReLU follows: fp32-level products
        ctx.relu = relu
        ctx.has_bias = bias is not None
        ctx.save_for_backward(x2, weight, y if relu else None)
        return y

    @staticmethod
    def backward(ctx, gy):
        x, weight, y = ctx.saved_tensors
        M, K = x.shape
        N = weight.shape[0]
        gy = _rows(gy)
        dz = ops.relu_bwd(gy, y) if ctx.relu else gy
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            dx = torch.empty(M, K, dtype=torch.float32, device=x.device)
            ops.gemm(dz, weight, dx, M, K, N, trans_b=True)                 # dX = dZ . W
        if ctx.needs_input_grad[1]:
            dw = torch.empty(N, K, dtype=torch.float32, device=x.device)
            ops.gemm(dz, x, dw, N, K, M, trans_a=True, trans_b=True)        # dW = dZ^T . X
        if ctx.has_bias and ctx.needs_input_grad[2]:
            db = ops.colsum(dz)
        return dx, dw, db, None


class CrossLayerFn(torch.autograd.Function):
    """x_next = xi + x0 * (xi W^T + b).  reference: CrossNetV2.forward, code/layers.py:197-201"""

    @staticmethod
    def forward(ctx, xi, x0, weight, bias):
        xi, x0 = _rows(xi), _rows(x0)
        M, K = xi.shape
        out = torch.empty(M, K, dtype=torch.float32, device=xi.device)
        u = torch.empty(M, K, dtype=torch.float32, device=xi.device)
        ops.gemm(xi, weight, out, M, K, K, epilogue=_lib.EPI_CROSS, bias=bias, aux0=xi, aux1=x0, aux_out=u)
        ctx.save_for_backward(xi, x0, weight, u)
        return out

    @staticmethod
    def backward(ctx, g):
        xi, x0, weight, u = ctx.saved_tensors
        M, K = xi.shape
        g = _rows(g)
        du = torch.empty(M, K, dtype=torch.float32, device=g.device)
        dx0 = torch.empty(M, K, dtype=torch.float32, device=g.device)
        ops.cross_bwd_pre(g, x0, u, du, dx0, accumulate=False)              # dU = G*X0 ; dX0 = G*U
        dxi = torch.empty(M, K, dtype=torch.float32, device=g.device)
        ops.gemm(du, weight, dxi, M, K, K, trans_b=True, epilogue=_lib.EPI_ADD, aux0=g)   # dXi = G + dU . W
        dw = torch.empty(K, K, dtype=torch.float32, device=g.device)
        ops.gemm(du, xi, dw, K, K, M, trans_a=True, trans_b=True)           # dW = dU^T . Xi
        db = ops.colsum(du)
        return dxi, dx0, dw, db


# ------------------------------------------------------------------------------------------------ CIN (xDeepFM)
class CINRelayoutFn(torch.autograd.Function):
    """[B, C, D] -> pair-major [B*D, C] (p = b*D + d): the layout in which the 1x1 convolutions of CIN are GEMMs (csrc/cin.cu)"""

    @staticmethod
    def forward(ctx, x):
        x = _c(x)
        B, C_, D = x.shape
        ctx.shape = (B, C_, D)
        out = torch.empty(B * D, C_, dtype=torch.float32, device=x.device)
        return ops.cin_relayout(x, out, B, C_, D, to_pairs=True)

    @staticmethod
    def backward(ctx, g):
        B, C_, D = ctx.shape
        dx = torch.empty(B, C_, D, dtype=torch.float32, device=g.device)
        return ops.cin_relayout(_rows(g), dx, B, C_, D, to_pairs=False)


class CINHadamardFn(torch.autograd.Function):
    """z[p, h*M + m] = x0[p, h] * xi[p, m] with the row padded to `ldz` columns of zeros.  reference: torch.einsum("bhd,bmd->bhmd")
    + view, code/layers.py:714-715"""

    @staticmethod
    def forward(ctx, x0, xi, ldz: int):
        x0, xi = _rows(x0), _rows(xi)
        P, F = x0.shape
        M = xi.shape[1]
        z = torch.empty(P, ldz, dtype=torch.float32, device=x0.device)
        ops.cin_hadamard_fwd(x0, xi, F, M, z)
        ctx.save_for_backward(x0, xi)
        return z

    @staticmethod
    def backward(ctx, gz):
        x0, xi = ctx.saved_tensors
        P, F = x0.shape
        M = xi.shape[1]
        dx0 = torch.empty(P, F, dtype=torch.float32, device=gz.device)
        dxi = torch.empty(P, M, dtype=torch.float32, device=gz.device)
        ops.cin_hadamard_bwd(_rows(gz), x0, xi, F, M, dx0, dxi)
        return dx0, dxi, None


class CINPoolFn(torch.autograd.Function):
    """pooled[b, o] = sum_d y[b*D + d, o]   (X_i.sum(dim=-1), code/layers.py:719); y may carry padding columns beyond O"""

    @staticmethod
    def forward(ctx, y, B: int, D: int, O: int):
        y = _rows(y)
        ctx.dims = (B, D, O, y.shape[1])
        return ops.cin_pool_fwd(y, B, D, O, torch.empty(B, O, dtype=torch.float32, device=y.device))

    @staticmethod
    def backward(ctx, g):
        B, D, O, width = ctx.dims
        dy = torch.empty(B * D, width, dtype=torch.float32, device=g.device)
        return ops.cin_pool_bwd(_rows(g), B, D, O, dy), None, None, None


class PadMatrixFn(torch.autograd.Function):
    """zero-padded copy [rows, cols] -> [rows_p, cols_p] (TMA-aligned operand of a GEMM); backward = the top-left block"""

    @staticmethod
    def forward(ctx, w, rows_p: int, cols_p: int):
        w = _rows(w)
        ctx.shape = tuple(w.shape)
        out = torch.zeros(rows_p, cols_p, dtype=torch.float32, device=w.device)
        ops.copy2d(w, out[:w.shape[0], :w.shape[1]])
        return out

    @staticmethod
    def backward(ctx, g):
        r, c = ctx.shape
        dw = torch.empty(r, c, dtype=torch.float32, device=g.device)
        ops.copy2d(_rows(g)[:r, :c], dw)
        return dw, None, None


# ------------------------------------------------------------------------------------------------ MFP head
class GatherSlicesFn(torch.autograd.Function):
    """selected_output = gather(enc_output[B,F,P], 1, masked_index).  reference: code/models.py:75"""

    @staticmethod
    def forward(ctx, enc, masked_index, F: int, P: int):
        enc = _c(enc)
        mi = _c(masked_index)
        B, L = mi.shape
        ctx.save_for_backward(mi)
        ctx.dims = (B, F, P)
        ctx.enc_shape = tuple(enc.shape)
        return ops.gather_slices(enc, mi, F, P).view(B, L, P)

    @staticmethod
    def backward(ctx, g):
        (mi,) = ctx.saved_tensors
        B, F, P = ctx.dims
        d_enc = torch.zeros(B, F * P, dtype=torch.float32, device=g.device)
        ops.scatter_add_slices(_c(g).view(-1, P), mi, F, P, d_enc)
        return d_enc.view(ctx.enc_shape), None, None, None


class NCEFn(torch.autograd.Function):
    """NCELoss.forward with IndexLinear scores (code/nce/nce_loss.py:79-144, code/nce/index_linear.py:68-106).
    Returns (loss, logits [B,L,K+1], ids [B,L,K+1], acc_count int32[1])."""

    @staticmethod
    def forward(ctx, inp, emb_w, bias_w, target, noise, logq, norm_term, loss_type, reduction, tg_emb: TableGrad,
                tg_bias: TableGrad):
        B, L, P = inp.shape
        K = noise.shape[-1]
        N = B * L
        x = _c(inp).view(N, P)
        acc = torch.zeros(1, dtype=torch.int32, device=inp.device)
        scale = 1.0 / N if reduction == "elementwise_mean" else 1.0
        logits, ids, loss_pos, dz, d_in = ops.nce_fwd(x, _c(target).view(N), _c(noise).view(N, K), emb_w, bias_w.view(-1), logq,
                                                      norm_term, loss_type, grad_scale=scale, acc_count=acc)
        if reduction == "elementwise_mean":
            loss = ops.reduce_sum(loss_pos, 1.0 / N).view(())
        elif reduction == "sum":
            loss = ops.reduce_sum(loss_pos, 1.0).view(())
        else:
            loss = loss_pos.view(B, L)
        ctx.save_for_backward(x, ids, dz, d_in)
        ctx.meta = (B, L, P, K, tuple(emb_w.shape), reduction, tg_emb, tg_bias)
        ctx.mark_non_differentiable(logits, ids, acc)
        return loss, logits.view(B, L, K + 1), ids.view(B, L, K + 1), acc

    @staticmethod
    def backward(ctx, gl, _gl, _gi, _ga):
        x, ids, dz, d_in = ctx.saved_tensors
        B, L, P, K, (V, _), reduction, tg_emb, tg_bias = ctx.meta
        N = B * L
        if reduction in ("elementwise_mean", "sum"):
            gs = _c(gl).view(1)
            d_in_s = ops.scale_by_scalar(d_in, gs)
            dz_s = ops.scale_by_scalar(dz, gs)
        else:  # per-position upstream gradient (reduction='none'): plain torch broadcasting, not on the training path
            gp = _c(gl).view(N, 1)
            d_in_s, dz_s = d_in * gp, dz * gp
        d_emb = d_bias = None
        if ctx.needs_input_grad[1] or ctx.needs_input_grad[2]:
            n = N * (K + 1)
            plan = tg_emb.plan_for(n, V, x.device).run(ids.view(-1))
            gb = torch.empty(n, dtype=torch.float32, device=x.device)
            G = plan.reduce_rows(x, P, scale=dz_s.view(-1), group=K + 1, scalar_out=gb)
            d_emb = tg_emb.finish(plan, G, (V, P))
            d_bias = tg_bias.finish(plan, gb, (V, 1))
        return d_in_s.view(B, L, P), d_emb, d_bias, None, None, None, None, None, None, None, None


# ------------------------------------------------------------------------------------------------ BCE heads
class BCEWithLogitsFn(torch.autograd.Function):
    """BCEWithLogitsLoss(mean) + accuracy / positive-ratio statistics.  reference: code/models.py:80-84, 91-92.
    Returns (loss, stats[4] = {loss, #correct, sum(labels), n})."""

    @staticmethod
    def forward(ctx, logits, labels):
        z = _c(logits).view(-1)
        y = _c(labels).view(-1)
        stats, dz = ops.bce_logits(z, y)
        ctx.save_for_backward(dz)
        ctx.shape = logits.shape
        ctx.mark_non_differentiable(stats)
        return stats[0].view(()).clone(), stats

    @staticmethod
    def backward(ctx, gl, _gs):
        (dz,) = ctx.saved_tensors
        return ops.scale_by_scalar(dz, _c(gl).view(1)).view(ctx.shape), None


# ------------------------------------------------------------------------------------------------ DeepFM
class FMLRFn(torch.autograd.Function):
    """lr_fm = LR(ids) + FM2(feat_embed).  reference: code/models.py:137-143, code/layers.py:125-131"""

    @staticmethod
    def forward(ctx, feat_embed, ids, lr_w, lr_bias, tg: TableGrad):
        E = _c(feat_embed)
        ids = _c(ids)
        ctx.save_for_backward(E, ids)
        ctx.tg, ctx.V = tg, lr_w.shape[0]
        return ops.fm_lr_fwd(E, ids, lr_w.view(-1), lr_bias)

    @staticmethod
    def backward(ctx, g):
        E, ids = ctx.saved_tensors
        B, F, D = E.shape
        g = _c(g).view(B, 1)
        dE = torch.empty_like(E)
        d_occ = torch.empty(B * F, dtype=torch.float32, device=E.device)
        ops.fm_lr_bwd(E, g, 1, dE, d_occ)
        plan = ctx.tg.plan_for(B * F, ctx.V, E.device).run(ids.view(-1))
        G = plan.reduce_rows(d_occ, 1)
        d_w = ctx.tg.finish(plan, G, (ctx.V, 1))
        d_b = ops.reduce_sum(g.view(-1), 1.0)
        return dE, None, d_w, d_b, None
