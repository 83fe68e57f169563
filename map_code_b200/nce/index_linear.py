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

    def _fused(self, target, noise, input):
        return Fn.NCEFn.apply(input, self.emb.weight, self.bias.weight, target, noise, self.logprob_noise, float(self.norm_term),
                              self.loss_type, self.reduction, self.emb.table_grad, self.bias.table_grad)
