"""Mirror of the hot-path part of the reference's `code/layers.py` (Embeddings, MLPBlock, CrossNetV2, InnerProductLayer,
get_act): same class names, constructor arguments, parameter names (=> state_dict keys) and init distributions; the
compute is our sm_100a kernels through `functional.py`."""
from __future__ import annotations

import math
from itertools import combinations

import torch
from torch import nn
from torch.nn import init

from . import functional as Fn


class NoneAct(nn.Module):
    def forward(self, x):
        return x


def get_act(act_func):
    """layers.py:55-80.  'relu' is fused into the GEMM epilogue; 'none' is the identity; the reference's other
    activations (tanh, gelu, swish, ...) belong to backbones outside the accelerated path."""
    if isinstance(act_func, str):
        name = act_func.lower()
        if name == "relu":
            return nn.ReLU(inplace=True)
        if name == "none":
            return NoneAct()
        raise NotImplementedError(act_func)
    return act_func


class Linear(nn.Module):
    """nn.Linear with the same parameters/init; forward = tcgen05 GEMM with the bias (and optionally ReLU) epilogue."""

    def __init__(self, in_features: int, out_features: int, bias: bool = True):
        super().__init__()
        self.in_features, self.out_features = in_features, out_features
        self.weight = nn.Parameter(torch.empty(out_features, in_features))
        self.bias = nn.Parameter(torch.empty(out_features)) if bias else None
        self.reset_parameters()

    def reset_parameters(self):  # identical to torch.nn.Linear.reset_parameters
        init.kaiming_uniform_(self.weight, a=math.sqrt(5))
        if self.bias is not None:
            bound = 1 / math.sqrt(self.in_features) if self.in_features > 0 else 0
            init.uniform_(self.bias, -bound, bound)

    def forward(self, x, relu: bool = False):
        lead = x.shape[:-1]
        y = Fn.LinearFn.apply(x.reshape(-1, self.in_features), self.weight, self.bias, relu)
        return y.view(*lead, self.out_features)

    def extra_repr(self):
        return f"in_features={self.in_features}, out_features={self.out_features}, bias={self.bias is not None}"


class TableEmbedding(nn.Module):
    """nn.Embedding(V, D) replacement: parameter name `weight`, gather / dedup-scatter kernels, dense or sparse grads."""

    def __init__(self, num_embeddings: int, embedding_dim: int, grad_mode: str = "dense"):
        super().__init__()
        self.num_embeddings, self.embedding_dim = num_embeddings, embedding_dim
        self.weight = nn.Parameter(torch.empty(num_embeddings, embedding_dim))
        init.normal_(self.weight)
        self.table_grad = Fn.TableGrad(grad_mode)
        self.weight._map_table_grad = self.table_grad

    def forward(self, input_ids):
        if getattr(self.weight, "_map_sharded", None) is not None:
            # a sharded engine re-pointed this parameter at ONE rank's [V/R + 1, D] shard: a local gather would silently return
            # the wrong rows (the reference would raise an IndexError here)
            from . import _lib
            raise _lib.MapB200Error("this table is row-sharded over %d ranks: the module path cannot read it; use the sharded fused step "
                                    "(and full_state_dict() to export the reference layout)" % self.weight._map_sharded[0])
        return Fn.EmbeddingFn.apply(self.weight, input_ids, self.table_grad)

    def _apply(self, fn, recurse=True):
        r = super()._apply(fn, recurse)
        self.weight._map_table_grad = self.table_grad  # survives .to(device) / .cuda()
        return r


class Embeddings(nn.Module):
    """layers.py:83-102.  One global table shared by all fields; normal init with std = sqrt(2 / (F + D))."""

    def __init__(self, config):
        super().__init__()
        self.embedding = TableEmbedding(config.input_size, config.embed_size, getattr(config, "table_grad_mode", "dense"))
        std = math.sqrt(2.0 / float(config.num_fields + config.embed_size))
        with torch.no_grad():
            init.normal_(self.embedding.weight, std=std)
        self.embed_norm = config.embed_norm
        if self.embed_norm:
            raise NotImplementedError("embed_norm=True (LayerNorm on embeddings) is off by default in the reference "
                                      "(arguments.py:124) and outside the accelerated path")
        if getattr(config, "embed_dropout_rate", 0.0) != 0.0:
            raise NotImplementedError("embed_dropout_rate > 0 is outside the accelerated path (reference default 0.0)")

    def forward(self, input_ids):
        return self.embedding(input_ids)


class InnerProductLayer(nn.Module):
    """layers.py:105-137, 'product_sum' output only (what DeepFM uses).  Keeps the reference's non-trainable
    field_p / field_q / upper_triangle_mask parameters so state_dicts interchange."""

    def __init__(self, num_fields=None, output="product_sum"):
        super().__init__()
        if output != "product_sum":
            raise ValueError(f"InnerProductLayer output={output} is not supported on the accelerated path")
        self.output_type = output
        if num_fields is not None:
            p, q = zip(*list(combinations(range(num_fields), 2)))
            self.field_p = nn.Parameter(torch.LongTensor(p), requires_grad=False)
            self.field_q = nn.Parameter(torch.LongTensor(q), requires_grad=False)
            self.interaction_units = int(num_fields * (num_fields - 1) / 2)
            self.upper_triangle_mask = nn.Parameter(torch.triu(torch.ones(num_fields, num_fields), 1).type(torch.bool),
                                                    requires_grad=False)


class MLPBlock(nn.Module):
    """layers.py:173-188: [Linear -> act -> Dropout] x n inside `self.dnn` (Sequential indices 0,3,6,... hold the Linears)."""

    def __init__(self, input_dim, hidden_size=128, num_hidden_layers=3, hidden_act="relu", hidden_dropout_rate=0.5,
                 batch_norm=False):
        super().__init__()
        if hidden_dropout_rate != 0.0:
            raise NotImplementedError("hidden_dropout_rate > 0 is outside the accelerated path (run scripts use 0.0)")
        dense_layers = []
        for _ in range(num_hidden_layers):
            dense_layers.append(Linear(input_dim, hidden_size))
            dense_layers.append(get_act(hidden_act))
            dense_layers.append(nn.Dropout(p=hidden_dropout_rate))
            input_dim = hidden_size
        self.dnn = nn.Sequential(*dense_layers)

    def forward(self, inputs):
        x = inputs
        mods = list(self.dnn)
        for i in range(0, len(mods), 3):
            x = mods[i](x, relu=isinstance(mods[i + 1], nn.ReLU))  # ReLU fused in the GEMM epilogue; Dropout(p=0) = identity
        return x


class CrossNetV2(nn.Module):
    """layers.py:191-201."""

    def __init__(self, input_dim, num_cross_layers):
        super().__init__()
        self.num_layers = num_cross_layers
        self.cross_layers = nn.ModuleList(Linear(input_dim, input_dim) for _ in range(num_cross_layers))

    def forward(self, X0):
        Xi = X0
        for i in range(self.num_layers):
            layer = self.cross_layers[i]
            Xi = Fn.CrossLayerFn.apply(Xi, X0, layer.weight, layer.bias)
        return Xi


class Conv1dK1(nn.Module):
    """Parameters and init of nn.Conv1d(in_channels, out_channels, kernel_size=1) (weight [O, C, 1], bias [O]); the compute is
    the pair-major GEMM inside CIN.forward."""

    def __init__(self, in_channels: int, out_channels: int):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.weight = nn.Parameter(torch.empty(out_channels, in_channels, 1))
        self.bias = nn.Parameter(torch.empty(out_channels))
        init.kaiming_uniform_(self.weight, a=math.sqrt(5))   # torch.nn.modules.conv._ConvNd.reset_parameters
        bound = 1 / math.sqrt(in_channels) if in_channels > 0 else 0
        init.uniform_(self.bias, -bound, bound)


class CIN(nn.Module):
    """layers.py:696-721 (xDeepFM's Compressed Interaction Network).  Same parameter names (cin_layer.layer_i.weight/bias).
    Activations are pair-major [B*D, channels]; each layer = outer product kernel -> tensor-core GEMM over the pairs (K padded to
    a multiple of 8, outputs to a multiple of 4 and at least 8: zero rows / columns) -> sum over d."""

    def __init__(self, num_fields, cin_layer_units):
        super().__init__()
        self.cin_layer_units = list(cin_layer_units)
        self.num_fields = num_fields
        self.cin_layer = nn.ModuleDict()
        for i, unit in enumerate(self.cin_layer_units):
            in_channels = num_fields * self.cin_layer_units[i - 1] if i > 0 else num_fields ** 2
            self.cin_layer["layer_" + str(i + 1)] = Conv1dK1(in_channels, unit)

    def forward(self, X_0):
        B, F, D = X_0.shape
        x0 = Fn.CINRelayoutFn.apply(X_0)   # [B*D, F]
        xi, M = x0, F
        pooled = []
        for i, O in enumerate(self.cin_layer_units):
            layer = self.cin_layer["layer_" + str(i + 1)]
            K8 = (F * M + 7) // 8 * 8
            Op = max(8, (O + 3) // 4 * 4)
            z = Fn.CINHadamardFn.apply(x0, xi, K8)                                         # [B*D, K8]
            w = Fn.PadMatrixFn.apply(layer.weight.view(O, F * M), Op, K8)
            b = Fn.PadMatrixFn.apply(layer.bias.view(1, O), 1, Op).view(Op)
            y = Fn.LinearFn.apply(z, w, b, False)                                          # [B*D, Op]; columns >= O are zero
            pooled.append(Fn.CINPoolFn.apply(y, B, D, O))
            xi, M = y[:, :O], O
        return torch.cat(pooled, dim=-1)   # layout glue only (layers.py:720)
