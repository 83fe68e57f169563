"""Mirror of `code/nce/nce_loss.py` (NCELoss): same attributes (`loss_type`, `per_word`, `noise_ratio`, `norm_term`,
`reduction`), buffers (`logprob_noise`) and return tuple; the forward is the fused NCE kernel."""
import math

import torch
import torch.nn as nn

from .. import functional as Fn
from .alias_multinomial import AliasMultinomial

BACKOFF_PROB = 1e-10


class NCELoss(nn.Module):
    def __init__(self, config, norm_term="auto", reduction="elementwise_mean"):
        super().__init__()
        noise = config.feat_count
        probs = noise / noise.sum()
        probs = probs.clamp(min=BACKOFF_PROB)
        renormed_probs = probs / probs.sum()  # nce_loss.py:60-64
        self.register_buffer("logprob_noise", renormed_probs.log())
        self.alias = AliasMultinomial(probs=renormed_probs, config=config)
        self.noise_ratio = config.pt_neg_num
        self.norm_term = math.log(noise.numel()) if norm_term == "auto" else norm_term
        self.reduction = reduction
        self.per_word = True
        self.loss_type = "nce"
        self.last_acc_count = None  # device int32[1]: #positions whose argmax over [target | noise] is the target

    def forward(self, target, *args, noise_samples=None, **kwargs):
        batch, max_len = target.size(0), target.size(1)
        if self.loss_type == "full":
            raise NotImplementedError("loss_type='full' (dense softmax over V) is not on the accelerated path")
        if self.loss_type not in ("nce", "sampled"):
            stage = "training" if self.training else "inference"
            raise NotImplementedError("loss type {} not implemented at {}".format(self.loss_type, stage))  # nce_loss.py:126-132
        if noise_samples is None:
            noise_samples = self.get_noise(batch, max_len)
        loss, logits, ids, acc = self._fused(target, noise_samples, *args, **kwargs)
        self.last_acc_count = acc
        return loss, logits, ids

    def get_noise(self, batch_size, max_len):
        noise_size = (batch_size, max_len, self.noise_ratio)
        if self.per_word:
            return self.alias.draw(*noise_size).contiguous()
        return self.alias.draw(1, 1, self.noise_ratio).expand(*noise_size).contiguous()

    def _fused(self, target, noise, *args, **kwargs):
        raise NotImplementedError()
