"""Mirror of the hot-path part of the reference's `code/models.py`: BaseModel (from_config, get_outputs,
create_pretraining_predictor, load_for_finetune, load_from_target_model, validate_model_config), DCNV2, DeepFM, DNN, LR.
Same constructor / forward signatures, output tuples and state_dict keys."""
from __future__ import annotations

import logging

import torch
from torch import nn

from . import functional as Fn
from .layers import CIN, CrossNetV2, Embeddings, InnerProductLayer, Linear, MLPBlock, TableEmbedding
from .nce import IndexLinear

logger = logging.getLogger(__name__)


class _RFDHead(nn.Module):
    """pred_rfd = Sequential(Linear(final, F*P), ReLU, Linear(F*P, F)) — models.py:119-123.  Registered under the same
    child names '0' and '2' so the state_dict keys are pred_rfd.0.* / pred_rfd.2.*."""

    def __init__(self, input_dim, hidden, out):
        super().__init__()
        self.add_module("0", Linear(input_dim, hidden))
        self.add_module("2", Linear(hidden, out))

    def forward(self, x):
        h = getattr(self, "0")(x, relu=True)
        return getattr(self, "2")(h)


class BaseModel(nn.Module):
    used_params = []

    def __init__(self, model_name="BaseModel", config=None):
        super().__init__()
        self.model_name = model_name
        self.config = config

    @classmethod
    def from_config(cls, config):
        name = config.model_name.lower()  # models.py:30-51
        if name == "dnn":
            model_class = DNN
        elif name == "deepfm":
            model_class = DeepFM
        elif name == "dcnv2":
            model_class = DCNV2
        elif name == "xdeepfm":
            model_class = xDeepFM
        elif name in ("autoint", "trans", "fignn", "fgcnn"):
            raise NotImplementedError(f"{config.model_name}: backbone outside the accelerated hot path (SURVEY.md §8f)")
        else:
            raise NotImplementedError(config.model_name)
        return model_class(config)

    def validate_model_config(self):
        logger.info(f"  model_name = {self.model_name}")
        for key in self.used_params:
            logger.info(f"  {key} = {getattr(self.config, key)}")

    def get_outputs(self, inputs, labels=None, masked_index=None, is_pretrain=None, noise_samples=None):
        """models.py:62-95.  MFP: (loss, count, total_acc) ; RFD: (loss, count, acc, pos_ratio) ; CTR: (loss, logits) | (logits,).
        `total_acc` is a 0-dim device tensor instead of the reference's `.item()` python int (no host sync in the step);
        it compares and divides like the int."""
        cfg = self.config
        batch_size = inputs.shape[0]
        if (is_pretrain is None and cfg.pretrain) or is_pretrain:
            if cfg.pt_type == "MFP":
                enc_output = self.feat_encoder(inputs)
                selected = Fn.GatherSlicesFn.apply(enc_output, masked_index, cfg.num_fields, cfg.proj_size)
                loss, logits, features = self.mfp_criterion(labels, selected, noise_samples=noise_samples)
                total_acc = self.mfp_criterion.last_acc_count.view(())
                self.last_logits, self.last_features = logits, features
                return (loss, labels.shape[0] * labels.shape[1], total_acc)
            elif cfg.pt_type == "RFD":
                logits = self.pred_rfd(inputs)
                loss, stats = Fn.BCEWithLogitsFn.apply(logits, labels)
                count = labels.shape[0] * labels.shape[1]
                acc = stats[1] / count
                input_pos_ratio = stats[2] / count
                self.last_logits = logits
                return (loss, count, acc, input_pos_ratio)
            else:
                raise NotImplementedError
        outputs = (inputs,)
        if labels is not None:
            loss, _ = Fn.BCEWithLogitsFn.apply(inputs.view(-1), labels.float())
            outputs = (loss,) + outputs
        return outputs

    def load_from_target_model(self, target_model_dict):
        model_dict = self.state_dict()  # models.py:97-107: copy every tensor whose name and shape match
        for k, v in target_model_dict.items():
            if k in model_dict and model_dict[k].shape == target_model_dict[k].shape:
                model_dict[k] = v
                print(f"Load tensor: {k}, {v.shape}")
            else:
                print(f"Unmatched tensor in the target model: {k}, {v.shape}")
        self.load_state_dict(model_dict)

    def load_for_finetune(self, model_path):
        self.load_from_target_model(torch.load(model_path, map_location="cpu"))

    def create_pretraining_predictor(self, input_dim):
        cfg = self.config
        if cfg.pt_type == "MFP":
            self.feat_encoder = Linear(input_dim, cfg.num_fields * cfg.proj_size)
            self.mfp_criterion = IndexLinear(cfg)
        elif cfg.pt_type == "RFD":
            self.pred_rfd = _RFDHead(input_dim, cfg.num_fields * cfg.proj_size, cfg.num_fields)
        else:
            raise NotImplementedError

    # ---- helpers used by the trainer / optimizer -------------------------------------------------------------------
    def table_parameters(self):
        """[(name, parameter)] of the [V, .] tables (row-wise optimizer / row sharding apply to these)."""
        return [(n, p) for n, p in self.named_parameters() if hasattr(p, "_map_table_grad")]


class LR(BaseModel):
    """models.py:129-143 (used inside DeepFM)."""

    def __init__(self, config):
        super().__init__(model_name="LR", config=config)
        self.embed_w = TableEmbedding(config.input_size, 1, getattr(config, "table_grad_mode", "dense"))
        self.bias = nn.Parameter(torch.zeros(1), requires_grad=True)


    def forward(self, input_ids, labels=None):
        """models.py:137-143: logits = sum_f w[ids[b,f]] + bias (used by xDeepFM with use_lr; DeepFM goes through the fused
        FM + LR kernel instead)."""
        wx = self.embed_w(input_ids)            # [B, F, 1]
        ones = torch.ones(wx.shape[1], 1, dtype=torch.float32, device=wx.device)
        logits = Fn.LinearFn.apply(wx.view(wx.shape[0], -1), ones.view(1, -1), self.bias, False)
        return self.get_outputs(logits, labels, is_pretrain=False)


class DNN(BaseModel):
    used_params = ["embed_size", "hidden_size", "num_hidden_layers", "hidden_dropout_rate", "hidden_act"]

    def __init__(self, config):
        super().__init__(model_name="DNN", config=config)
        self.embed = Embeddings(config)
        self.dnn = MLPBlock(input_dim=config.embed_size * config.num_fields, hidden_size=config.hidden_size,
                            num_hidden_layers=config.num_hidden_layers, hidden_dropout_rate=config.hidden_dropout_rate,
                            hidden_act=config.hidden_act)
        if config.pretrain:
            self.create_pretraining_predictor(config.hidden_size)
        else:
            self.fc_out = Linear(config.hidden_size, 1)

    def forward(self, input_ids, labels=None, masked_index=None, noise_samples=None):
        feat_embed = self.embed(input_ids)
        nn_output = self.dnn(torch.flatten(feat_embed, 1))
        if self.config.pretrain:
            return self.get_outputs(nn_output, labels, masked_index, noise_samples=noise_samples)
        return self.get_outputs(self.fc_out(nn_output), labels)


class DeepFM(BaseModel):
    used_params = ["embed_size", "hidden_size", "num_hidden_layers", "hidden_dropout_rate", "hidden_act"]

    def __init__(self, config):
        super().__init__(model_name="DeepFM", config=config)
        self.embed = Embeddings(config)
        self.lr_layer = LR(config)
        self.dnn = MLPBlock(input_dim=config.num_fields * config.embed_size, hidden_size=config.hidden_size,
                            num_hidden_layers=config.num_hidden_layers, hidden_dropout_rate=config.hidden_dropout_rate,
                            hidden_act=config.hidden_act)
        self.ip_layer = InnerProductLayer(num_fields=config.num_fields)
        if config.pretrain:
            self.create_pretraining_predictor(config.hidden_size + 1)
        else:
            self.dnn_fc_out = Linear(config.hidden_size, 1)

    def _lr_fm(self, input_ids, feat_embed):
        return Fn.FMLRFn.apply(feat_embed, input_ids, self.lr_layer.embed_w.weight, self.lr_layer.bias,
                               self.lr_layer.embed_w.table_grad)

    def forward(self, input_ids, labels=None, masked_index=None, noise_samples=None):
        feat_embed = self.embed(input_ids)
        dnn_vec = self.dnn(feat_embed.flatten(start_dim=1))
        if self.config.pretrain:
            lr_fm = self._lr_fm(input_ids, feat_embed)
            final_vec = torch.cat([dnn_vec, lr_fm], dim=1)  # layout glue only (models.py:224); the fused step writes in place
            return self.get_outputs(final_vec, labels, masked_index, noise_samples=noise_samples)
        logits = self.dnn_fc_out(dnn_vec)
        logits = logits + self._lr_fm(input_ids, feat_embed)
        return self.get_outputs(logits, labels)


class xDeepFM(BaseModel):
    """models.py:235-279.  CIN over the field embeddings (+ an MLP tower), pretraining heads on cat([cin, dnn])."""
    used_params = ["embed_size", "hidden_size", "num_hidden_layers", "hidden_dropout_rate", "hidden_act", "cin_layer_units", "use_lr"]

    def __init__(self, config):
        super().__init__(model_name="xDeepFM", config=config)
        self.embed = Embeddings(config)
        input_dim = config.num_fields * config.embed_size
        cin_layer_units = [int(c) for c in str(config.cin_layer_units).split(",")]
        self.cin = CIN(config.num_fields, cin_layer_units)
        if config.num_hidden_layers > 0:
            self.dnn = MLPBlock(input_dim=input_dim, hidden_size=config.hidden_size, num_hidden_layers=config.num_hidden_layers,
                                hidden_dropout_rate=config.hidden_dropout_rate, hidden_act=config.hidden_act)
            final_dim = sum(cin_layer_units) + config.hidden_size
        else:
            self.dnn = None
            final_dim = sum(cin_layer_units)
        if config.pretrain:
            self.create_pretraining_predictor(final_dim)
        else:
            self.lr_layer = LR(config) if getattr(config, "use_lr", False) else None
            self.fc = Linear(final_dim, 1)

    def forward(self, input_ids, labels=None, masked_index=None, noise_samples=None):
        feat_embed = self.embed(input_ids)
        final_vec = self.cin(feat_embed)
        if self.dnn is not None:
            dnn_vec = self.dnn(feat_embed.flatten(start_dim=1))
            final_vec = torch.cat([final_vec, dnn_vec], dim=1)   # layout glue only (models.py:269)
        if self.config.pretrain:
            return self.get_outputs(final_vec, labels, masked_index, noise_samples=noise_samples)
        logits = self.fc(final_vec)
        if self.lr_layer is not None:
            logits = logits + self.lr_layer(input_ids)[0]
        return self.get_outputs(logits, labels)


class DCNV2(BaseModel):
    used_params = ["embed_size", "hidden_size", "num_hidden_layers", "hidden_dropout_rate", "hidden_act", "num_cross_layers"]

    def __init__(self, config):
        super().__init__(model_name="DCNV2", config=config)
        self.embed = Embeddings(config)
        input_dim = config.num_fields * config.embed_size
        self.cross_net = CrossNetV2(input_dim, config.num_cross_layers)
        if config.num_hidden_layers > 0:
            self.parallel_dnn = MLPBlock(input_dim=input_dim, hidden_size=config.hidden_size,
                                         num_hidden_layers=config.num_hidden_layers,
                                         hidden_dropout_rate=config.hidden_dropout_rate, hidden_act=config.hidden_act)
            final_dim = input_dim + config.hidden_size
        else:
            final_dim = input_dim
        if config.pretrain:
            self.create_pretraining_predictor(final_dim)
        else:
            self.fc_out = Linear(final_dim, 1)

    def forward(self, input_ids, labels=None, masked_index=None, noise_samples=None):
        feat_embed = self.embed(input_ids).flatten(start_dim=1)
        cross_output = self.cross_net(feat_embed)
        if self.config.num_hidden_layers > 0:
            dnn_output = self.parallel_dnn(feat_embed)
            final_output = torch.cat([cross_output, dnn_output], dim=-1)  # layout glue only (models.py:312)
        else:
            final_output = cross_output
        if self.config.pretrain:
            return self.get_outputs(final_output, labels, masked_index, noise_samples=noise_samples)
        return self.get_outputs(self.fc_out(final_output), labels)
