"""Mirror of the reference's `code/arguments.py`: same dataclass field names (they are the config contract),
without the HfArgumentParser / transformers dependency.  Device selection follows arguments.py:58-78, except that the
multi-GPU branch hands out `cuda:LOCAL_RANK` and leaves process-group setup to `map_code_b200.dist` (the reference calls
init_process_group and then never issues a collective)."""
from __future__ import annotations

import copy
import dataclasses
import json
import os
from dataclasses import dataclass, field
from typing import Any, Dict, Optional

import torch


@dataclass
class TrainingArguments:
    output_dir: str = "outputs"
    dataset_name: str = "avazu"
    data_dir: str = "data/avazu"
    per_gpu_train_batch_size: int = 128
    per_gpu_eval_batch_size: int = 10000
    learning_rate: float = 1e-4
    weight_decay: float = 0.1
    adam_epsilon: float = 1e-8
    adam_betas: str = "0.9,0.999"
    max_grad_norm: float = 0.0
    patience: int = 2
    num_train_epochs: int = 20
    lr_sched: str = "cosine"
    warmup_ratio: float = 0.0
    logging_first_step: bool = False
    logging_steps: int = 1000
    save_steps: int = 1000
    save_total_limit: Optional[int] = 20
    no_cuda: bool = False
    seed: int = 42
    local_rank: int = -1
    # pretraining (arguments.py:41-47)
    sampling_method: str = "normal"
    mask_ratio: float = 0.1
    pretrain: bool = False
    pt_type: str = "MFP"
    RFD_replace: str = "Unigram"
    finetune: bool = False
    pretrained_model_path: Optional[str] = None
    # additions of this implementation (not in the reference)
    optimizer_mode: str = "sparse"   # "sparse" (touched rows only) | "dense_exact" (reference semantics: every row, every step)

    @property
    def n_gpu(self) -> int:
        return 0 if self.no_cuda else int(os.environ.get("WORLD_SIZE", "1")) if torch.cuda.is_available() else 0

    @property
    def train_batch_size(self) -> int:
        return self.per_gpu_train_batch_size * max(1, self.n_gpu)

    @property
    def eval_batch_size(self) -> int:
        return self.per_gpu_eval_batch_size * max(1, self.n_gpu)

    @property
    def device(self) -> torch.device:
        if self.no_cuda or not torch.cuda.is_available():
            # the reference falls back to CPU here (arguments.py:62-64); this implementation has no CPU path
            raise RuntimeError("map_code_b200 runs on CUDA (sm_100a) only; there is no CPU path (no_cuda is not supported)")
        return torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))

    def to_json_string(self):
        return json.dumps(dataclasses.asdict(self), indent=2)

    def to_sanitized_dict(self) -> Dict[str, Any]:
        d = dataclasses.asdict(self)
        valid = [bool, int, float, str, torch.Tensor]
        return {k: v if type(v) in valid else str(v) for k, v in d.items()}


@dataclass
class ModelArguments:
    model_name: str = "DCNv2"
    embed_size: int = 32
    embed_dropout_rate: float = 0.0
    hidden_size: int = 128
    num_hidden_layers: int = 1
    hidden_act: str = "relu"
    hidden_dropout_rate: float = 0.0
    layer_norm_eps: float = 1e-12
    embed_norm: bool = False
    num_cross_layers: int = 1
    use_lr: bool = False            # xDeepFM (arguments.py:141)
    cin_layer_units: str = "50,50"  # xDeepFM (arguments.py:144)
    pt_neg_num: int = 25
    proj_size: int = 32

    def to_dict(self):
        return copy.deepcopy(self.__dict__)


class Config:
    """Free-form attribute bag with JSON (de)serialisation — arguments.py:164-203."""

    def __init__(self, **kwargs):
        for k, v in kwargs.items():
            setattr(self, k, v)

    def save(self, save_directory):
        assert os.path.isdir(save_directory), f"not a directory: {save_directory}"
        self.to_json_file(os.path.join(save_directory, "config.json"))

    @classmethod
    def load(cls, load_directory):
        return cls.from_dict(cls.from_json_file(os.path.join(load_directory, "config.json")))

    @classmethod
    def from_dict(cls, config_dict: Dict):
        return cls(**config_dict)

    @classmethod
    def from_json_file(cls, json_file: str):
        with open(json_file, "r", encoding="utf-8") as reader:
            return json.loads(reader.read())

    def to_dict(self):
        out = {}
        for k, v in self.__dict__.items():
            out[k] = v.tolist() if torch.is_tensor(v) else copy.deepcopy(v)
        return out

    def to_json_string(self):
        d = {k: v for k, v in self.to_dict().items() if not isinstance(v, torch.device)}
        return json.dumps(d, indent=2, sort_keys=True, default=str) + "\n"

    def to_json_file(self, json_file_path):
        with open(json_file_path, "w", encoding="utf-8") as writer:
            writer.write(self.to_json_string())
