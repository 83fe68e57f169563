"""Mirror of `code/nce/index_linear.py` (IndexLinear): parameters `emb.weight [V,P]`, `bias.weight [V,1]` with the
reference's init (uniform(-1/sqrt(P), 1/sqrt(P)); bias = logprob_noise + ln V)."""
import math

import torch

from .. import functional as Fn
from ..layers import TableEmbedding
from .nce_loss import NCELoss


class IndexLinear(NCELoss):
    def __init__(self, config):
        super().__init__(config)
        mode = getattr(config, "table_grad_mode", "dense")
        self.emb = TableEmbedding(config.input_size, config.proj_size, mode)
        self.bias = TableEmbedding(config.input_size, 1, mode)
        self.reset_parameters()

    def reset_parameters(self):  # index_linear.py:40-48
        stdv = 1.0 / math.sqrt(self.emb.embedding_dim)
        self.emb.weight.data.uniform_(-stdv, stdv)
        self.bias.weight.data = torch.unsqueeze(self.logprob_noise + self.norm_term, 1)
        self.bias.weight._map_table_grad = self.bias.table_grad

    def ce_loss(self, target_idx, input):
        """index_linear.py:145-151: full-softmax cross entropy per position, [B, L].  Evaluation only (no autograd through it): the
        reference's [N, V] score matrix is never materialised (ops.nce_full_ce)."""
        from .. import ops
        P = input.shape[-1]
        with torch.no_grad():
            loss = ops.nce_full_ce(input.detach().reshape(-1, P).contiguous(), self.emb.weight.data, self.bias.weight.data,
                                   target_idx.reshape(-1).contiguous())
        return loss.view_as(target_idx)

    def _fused(self, target, noise, input):
        return Fn.NCEFn.apply(input, self.emb.weight, self.bias.weight, target, noise, self.logprob_noise, float(self.norm_term),
                              self.loss_type, self.reduction, self.emb.table_grad, self.bias.table_grad)
