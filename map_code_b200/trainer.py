"""Mirror of the reference's `code/trainer.py` (Trainer): same constructor, `dynamic_mask`, `get_optimizer`,
`MFP_pretrain`, `RFD_pretrain`, `train`, `eval`, `MFP_pretrain_eval`, `RFD_pretrain_eval`, `save_model`, `load_model`,
`test`.  Differences, all on purpose:
  * the inner-loop body runs as one captured CUDA graph (`engine.FusedStep`) for the DCNv2 / DNN / DeepFM backbones (the module
    path — autograd over the same kernels — serves eval and ragged last batches);
  * `dynamic_mask` runs on the device with a counter-based Philox stream (seed, step) instead of torch's host generator;
  * batches come from a device-resident shuffled-index batcher instead of DataLoader + per-row __getitem__
    (reference: 12 ms per 4096-row batch on the host);
  * loss / accuracy are read from the device every `logging_steps` (and once per epoch), not every step;
  * the reference's NameError at trainer.py:341 (`str(log)`) is not reproduced.
"""
from __future__ import annotations

import logging
import math
import os
import time
from typing import Dict, Optional

import numpy as np
import torch

from . import ops
from .engine import FusedStep, is_no_decay, STREAM_MASK_FIELD, STREAM_RFD_REPLACE, STREAM_RFD_FIELD2
from .optim import AdamW

logger = logging.getLogger(__name__)
STREAMS_PER_STEP = 8


def cosine_schedule_lambda(num_warmup_steps: int, num_training_steps: int, num_cycles: float = 0.5):
    """transformers.get_cosine_schedule_with_warmup (call site trainer.py:77-79)."""

    def f(step: int) -> float:
        if step < num_warmup_steps:
            return float(step) / float(max(1, num_warmup_steps))
        progress = float(step - num_warmup_steps) / float(max(1, num_training_steps - num_warmup_steps))
        return max(0.0, 0.5 * (1.0 + math.cos(math.pi * float(num_cycles) * 2.0 * progress)))

    return f


def constant_schedule_lambda(num_warmup_steps: int):
    """transformers.get_constant_schedule_with_warmup (call site trainer.py:80-81)."""

    def f(step: int) -> float:
        return float(step) / float(max(1.0, num_warmup_steps)) if step < num_warmup_steps else 1.0

    return f


class DeviceBatcher:
    """Device-resident replacement of DataLoader(shuffle=...) over OurDataset (trainer.py:51-58, dataset.py:78-87): the id
    matrix lives in HBM, an epoch is a permutation of row indices, a batch is one gather kernel."""

    def __init__(self, X, Y, batch_size: int, shuffle: bool, device, seed: int = 0, rank: int = 0, world: int = 1):
        """rank / world: data-parallel runs see the SAME permutation on every rank (same seed) and take disjoint slices of every
        global batch of batch_size * world rows; the ragged tail of an epoch is dropped there (every rank must step together)."""
        self.rank, self.world = rank, world
        self.X = torch.as_tensor(np.ascontiguousarray(X) if isinstance(X, np.ndarray) else X).to(device=device, dtype=torch.int64).contiguous()
        self.Y = None if Y is None else torch.as_tensor(np.ascontiguousarray(Y) if isinstance(Y, np.ndarray) else Y).to(device=device)
        self.batch_size, self.shuffle = batch_size, shuffle
        self.gen = torch.Generator().manual_seed(seed)
        self.dataset = self

    def __len__(self):  # number of batches, like len(DataLoader) with drop_last=False
        if self.world > 1:
            return self.X.shape[0] // (self.batch_size * self.world)
        return (self.X.shape[0] + self.batch_size - 1) // self.batch_size

    def num_rows(self):
        return self.X.shape[0]

    def __iter__(self):
        n = self.X.shape[0]
        order = (torch.randperm(n, generator=self.gen) if self.shuffle else torch.arange(n)).to(self.X.device)
        if self.world > 1:
            gb = self.batch_size * self.world
            for i in range(0, n - gb + 1, gb):
                idx = order[i + self.rank * self.batch_size:i + (self.rank + 1) * self.batch_size]
                yield ops.gather_rows_i64(self.X, idx), (self.Y[idx] if self.Y is not None else None)
            return
        for i in range(0, n, self.batch_size):
            idx = order[i:i + self.batch_size]
            xb = ops.gather_rows_i64(self.X, idx)
            yb = self.Y[idx] if self.Y is not None else None
            yield xb, yb


class PendingStep:
    """Result handle of Trainer.train_step_async: the loss of that step, readable once its device work is done."""

    def __init__(self, done, host_buf):
        self._done, self._host = done, host_buf

    def loss(self) -> float:
        self._done.synchronize()
        return float(self._host[0])


class Trainer:
    def __init__(self, model, model_config, training_args, train_dataset, eval_dataset):
        self.model = model
        self.model_config = model_config
        self.args = training_args
        self.device = self.args.device
        self.train_dataset = train_dataset
        self.eval_dataset = eval_dataset
        self.global_step = 0
        self.eval_metrics = []
        logger.info(f"setting device {self.device}")
        self.optimizer = None
        self.scheduler = None
        self.train_dataloader = None
        self.eval_dataloader = None
        self.best_eval_auc = 0
        self.best_eval_step = -1
        self._fused: Optional[FusedStep] = None
        self._mask_calls = 0
        self._x_train_dev = None
        self._h2d = None
        self._pipe = None

    # ------------------------------------------------------------------------------------------------ data
    def get_dataloader(self, dataset, is_training=True):
        bs = self.args.per_gpu_train_batch_size if is_training else self.args.per_gpu_eval_batch_size
        eng = self._fused
        world, rank = (getattr(eng, "world", 1), getattr(eng, "rank", 0)) if (eng is not None and is_training) else (1, 0)
        return DeviceBatcher(dataset.X, getattr(dataset, "Y", None), bs, shuffle=is_training, device=self.device, seed=self.args.seed,
                             rank=rank, world=world)

    def _train_matrix(self):
        if self._x_train_dev is None and self.train_dataset is not None:
            X = self.train_dataset.X
            self._x_train_dev = torch.as_tensor(np.ascontiguousarray(X) if isinstance(X, np.ndarray) else X).to(
                device=self.device, dtype=torch.int64).contiguous()
        return self._x_train_dev

    # ------------------------------------------------------------------------------------------------ optimizer
    def get_optimizer(self, num_training_steps: int, num_warmup_steps: int):
        """trainer.py:60-85: two parameter groups (names containing "bias"/"LayerNorm.weight" are not decayed), AdamW with
        transformers 4.26.1 semantics, cosine or constant-with-warmup LambdaLR."""
        named = [(k, v) for k, v in self.model.named_parameters() if v.requires_grad]
        groups = [
            {"params": [p for n, p in named if not is_no_decay(n)], "weight_decay": self.args.weight_decay},
            {"params": [p for n, p in named if is_no_decay(n)], "weight_decay": 0.0},
        ]
        beta1, beta2 = (float(b) for b in self.args.adam_betas.split(","))
        optimizer = AdamW(groups, lr=self.args.learning_rate, eps=self.args.adam_epsilon, betas=(beta1, beta2),
                          table_mode=getattr(self.args, "optimizer_mode", "sparse"))
        sched = self.args.lr_sched.lower()
        if sched == "cosine":
            lam = cosine_schedule_lambda(num_warmup_steps, num_training_steps)
        elif sched == "const":
            lam = constant_schedule_lambda(num_warmup_steps)
        else:
            raise NotImplementedError
        return optimizer, torch.optim.lr_scheduler.LambdaLR(optimizer, lam)

    # ------------------------------------------------------------------------------------------------ masking
    def dynamic_mask(self, inputs, sampling_method="normal", step: Optional[int] = None):
        """trainer.py:217-266 on the device.  `step` selects the Philox subsequence (default: an internal call counter)."""
        ids = inputs["input_ids"].to(self.device, non_blocking=True).contiguous()
        batch_size = ids.shape[0]
        num_fields = self.model_config.num_fields
        mask_num = int(num_fields * self.args.mask_ratio)
        if sampling_method not in ("normal", "randint"):
            raise NotImplementedError(sampling_method)
        if step is None:
            step = self._mask_calls
            self._mask_calls += 1
        base = step * STREAMS_PER_STEP
        seed = self.args.seed
        masked_index = ops.mask_index(batch_size, mask_num, num_fields, sampling_method, seed, base + STREAM_MASK_FIELD, device=self.device)
        if self.args.pt_type == "MFP":
            new_ids, labels = ops.mfp_mask_apply(ids, masked_index, 3)
            inputs["labels"] = labels
            inputs["input_ids"] = new_ids
            inputs["masked_index"] = masked_index
        elif self.args.pt_type == "RFD":
            mode = self.args.RFD_replace
            if mode not in ("Unigram", "Uniform", "Whole-Uniform", "Whole-Unigram"):
                raise NotImplementedError
            cfg = self.model_config
            lo = getattr(cfg, "idx_low", None)
            hi = getattr(cfg, "idx_high", None)
            new_ids, labels = ops.rfd_replace(
                ids, masked_index, mode, seed, base + STREAM_RFD_REPLACE, base + STREAM_RFD_FIELD2,
                x_train=self._train_matrix() if mode in ("Unigram", "Whole-Unigram") else None,
                idx_low=None if lo is None else lo.to(self.device), idx_high=None if hi is None else hi.to(self.device),
                input_size=cfg.input_size)
            inputs["input_ids"] = new_ids
            inputs["labels"] = labels
        else:
            raise NotImplementedError(self.args.pt_type)
        return inputs

    # ------------------------------------------------------------------------------------------------ fused step
    def fused_step(self, total_steps: int, warmup_steps: int, batch_size: Optional[int] = None) -> FusedStep:
        if self._fused is None:
            cfg, a = self.model_config, self.args
            beta1, beta2 = (float(b) for b in a.adam_betas.split(","))
            need_x = cfg.pretrain and cfg.pt_type == "RFD" and a.RFD_replace in ("Unigram", "Whole-Unigram")
            lo, hi = getattr(cfg, "idx_low", None), getattr(cfg, "idx_high", None)
            if getattr(a, "max_grad_norm", 0) and a.max_grad_norm > 0:
                # reference trainer.py:137 clips when max_grad_norm > 0 (default 0: off); the fused step has no global-norm clip
                raise NotImplementedError("max_grad_norm > 0 is not implemented on the fused step")
            self._fused_kw = dict(
                mask_ratio=a.mask_ratio, sampling_method=a.sampling_method, lr=a.learning_rate, weight_decay=a.weight_decay,
                betas=(beta1, beta2), eps=a.adam_epsilon, sched=a.lr_sched, warmup_steps=warmup_steps, total_steps=total_steps,
                seed=a.seed, optimizer_mode=getattr(a, "optimizer_mode", "sparse"), x_train=self._train_matrix() if need_x else None,
                idx_low=None if lo is None else lo.to(self.device), idx_high=None if hi is None else hi.to(self.device))
            self._fused = FusedStep(self.model, batch_size=batch_size or a.per_gpu_train_batch_size, **self._fused_kw)
            self._ragged = {}
        return self._fused

    def _engine_for(self, n_rows: int) -> FusedStep:
        """The fused step for a batch of n_rows.  The last batch of an epoch may be short (the reference's DataLoader keeps it,
        trainer.py:51-58): it runs through a second FusedStep of that size which SHARES the optimizer state, step counter and
        schedule of the main one (engine.FusedStep share_state_with) — one optimizer for every batch, like the reference."""
        eng = self._fused
        if n_rows == eng.B:
            return eng
        if getattr(eng, "world", 1) > 1:
            raise NotImplementedError("ragged batches are not supported by the sharded step: use a dataset size that is a multiple of the global batch")
        if n_rows not in self._ragged:
            self._ragged[n_rows] = FusedStep(self.model, batch_size=n_rows, share_state_with=eng, **self._fused_kw)
        return self._ragged[n_rows]

    def supports_fused(self) -> bool:
        return self.model.model_name.lower() in ("dcnv2", "dnn", "deepfm")

    def train_step(self, X, Y=None):
        """Public single-step API: X [B, F] int64 on the host (ideally pinned) or on the device; returns the reference's
        output tuple as device tensors.  Host inputs are copied with one async H2D transfer per tensor."""
        if self._fused is None:
            raise RuntimeError("call fused_step(total_steps, warmup_steps) (or MFP_pretrain/RFD_pretrain/train) first")
        eng = self._engine_for(X.shape[0])
        if not X.is_cuda:
            if self._h2d is None or self._h2d.shape[0] != X.shape[0]:
                self._h2d = torch.empty(X.shape[0], eng.F, dtype=torch.int64, device=self.device)
            self._h2d.copy_(X, non_blocking=True)
            X = self._h2d
        if Y is not None and not Y.is_cuda:
            Y = Y.to(self.device, non_blocking=True)
        eng.step(X, Y)
        self.global_step += 1
        return eng.outputs()

    def train_step_async(self, X, Y=None) -> "PendingStep":
        """Pipelined form of train_step for HOST batches: the H2D copy of this batch runs on a copy stream (two staging buffers),
        the step is enqueued behind it, and the loss is copied to pinned host memory behind the step.  Nothing blocks here:
        `PendingStep.loss()` waits for that step only, so a loop that reads step i-1 after submitting step i overlaps every
        transfer with the neighbouring steps' compute."""
        eng = self._fused
        if eng is None:
            raise RuntimeError("call fused_step(total_steps, warmup_steps) (or MFP_pretrain/RFD_pretrain/train) first")
        if self._pipe is None:
            dev = self.device
            self._pipe = dict(copy_stream=torch.cuda.Stream(device=dev), k=0,
                              stage=[torch.empty(eng.B, eng.F, dtype=torch.int64, device=dev) for _ in range(2)],
                              stage_y=[None, None], loaded=[torch.cuda.Event(), torch.cuda.Event()],
                              consumed=[torch.cuda.Event(), torch.cuda.Event()], used=[False, False],
                              host=[torch.empty(1, dtype=torch.float32).pin_memory() for _ in range(2)])
        pp = self._pipe
        k = pp["k"]
        pp["k"] = 1 - k
        main = torch.cuda.current_stream()
        cs = pp["copy_stream"]
        if pp["used"][k]:
            cs.wait_event(pp["consumed"][k])      # the step that last read this staging buffer has copied it into the graph input
        with torch.cuda.stream(cs):
            pp["stage"][k].copy_(X, non_blocking=True)
            if Y is not None:
                pp["stage_y"][k] = Y.to(self.device, non_blocking=True)
            pp["loaded"][k].record(cs)
        main.wait_event(pp["loaded"][k])
        out = eng.step(pp["stage"][k], pp["stage_y"][k] if Y is not None else None)
        pp["consumed"][k].record(main)
        pp["used"][k] = True
        done = torch.cuda.Event()
        pp["host"][k].copy_(out.detach().view(1), non_blocking=True)
        done.record(main)
        self.global_step += 1
        return PendingStep(done, pp["host"][k])

    # ------------------------------------------------------------------------------------------------ loops
    def _log_header(self, t_total, t_warmup):
        a, c = self.args, self.model_config
        logger.info("***** running %s *****", "pretraining" if c.pretrain else "training")
        for k, v in [("dataset_name", a.dataset_name), ("input_size", c.input_size), ("num_fields", c.num_fields),
                     ("num_examples", len(self.train_dataset)), ("num_epochs", a.num_train_epochs),
                     ("batch_size", a.train_batch_size), ("total_steps", t_total), ("warmup_steps", t_warmup),
                     ("learning_rate", a.learning_rate), ("weight_decay", a.weight_decay), ("lr_sched", a.lr_sched)]:
            logger.info(f"  {k} = {v}")
        self.model.validate_model_config()

    def _pretrain(self, kind: str):
        a = self.args
        self.train_dataloader = self.get_dataloader(self.train_dataset)
        t_total = int(len(self.train_dataloader) * a.num_train_epochs)
        t_warmup = int(t_total * a.warmup_ratio)
        self._log_header(t_total, t_warmup)
        self.global_step = 0
        self.eval_metrics = []
        self.model.to(self.device)
        fused = self.supports_fused()
        if fused:
            eng = self.fused_step(t_total, t_warmup)
        else:
            self.optimizer, self.scheduler = self.get_optimizer(t_total, t_warmup)
            self.model.zero_grad()
        win_loss = torch.zeros(1, device=self.device)
        win_acc = torch.zeros(1, device=self.device)
        win_n = 0
        start_time = time.time()
        for epoch in range(a.num_train_epochs):
            logger.info(f"-------------------- epoch-{epoch} --------------------")
            self.model.train()
            for X, _ in self.train_dataloader:
                if fused:
                    outputs = self.train_step(X)
                else:  # module path (backbones without a fused schedule): autograd over the same kernels
                    if self.optimizer is None:
                        self.optimizer, self.scheduler = self.get_optimizer(t_total, t_warmup)
                    inputs = self.dynamic_mask({"input_ids": X}, a.sampling_method, step=self.global_step)
                    outputs = self.model(**inputs)
                    outputs[0].backward()
                    self.optimizer.step()
                    self.scheduler.step()
                    self.optimizer.zero_grad()
                    self.global_step += 1
                win_loss += outputs[0].detach().view(1)
                win_acc += (outputs[2].detach().float().view(1) / outputs[1]) if kind == "MFP" else outputs[2].detach().view(1)
                win_n += 1
                if self.global_step % a.logging_steps == 0:
                    if self._fused is not None:
                        self._fused.check_health()
                    _log = {f"window_{kind.lower()}_loss": win_loss.item() / win_n, f"window_{kind.lower()}_acc": win_acc.item() / win_n,
                            "time_cost": time.time() - start_time}
                    logger.info(f"step = {self.global_step}, {str(_log)}")
                    win_loss.zero_(); win_acc.zero_(); win_n = 0
                    start_time = time.time()
            sharded = self._fused is not None and getattr(self._fused, "world", 1) > 1
            if sharded:
                logger.info("per-epoch eval skipped: the tables are row-sharded (the module path cannot read a shard)")
            elif a.local_rank in [-1, 0]:
                self.MFP_pretrain_eval() if kind == "MFP" else self.RFD_pretrain_eval()
        if self._fused is not None and getattr(self._fused, "world", 1) > 1:
            self.save_model(a.output_dir)      # collective: every rank takes part, rank 0 writes
        elif a.local_rank in [-1, 0]:
            self.save_model(a.output_dir)

    def MFP_pretrain(self):
        """trainer.py:268-353"""
        self._pretrain("MFP")

    def RFD_pretrain(self):
        """trainer.py:393-477"""
        self._pretrain("RFD")

    def _pretrain_eval(self, kind: str):
        loader = self.get_dataloader(self.eval_dataset, is_training=False)
        logger.info("***** running eval *****")
        logger.info(f"  num examples = {loader.num_rows()}")
        total_loss = torch.zeros(1, device=self.device)
        total_acc = torch.zeros(1, device=self.device)
        count = 0
        self.model.eval()
        start = time.time()
        with torch.no_grad():
            for X, _ in loader:
                inputs = self.dynamic_mask({"input_ids": X}, self.args.sampling_method)
                outputs = self.model(**inputs)
                loss, n, acc = outputs[:3]
                count += n
                total_loss += loss.view(1) * n
                total_acc += acc.float().view(1) if kind == "MFP" else acc.view(1) * n
        ev_loss, ev_acc = total_loss.item() / max(count, 1), total_acc.item() / max(count, 1)
        self.eval_metrics.append([ev_loss, ev_acc])
        lr = self.scheduler.get_last_lr()[0] if self.scheduler is not None else float(self._fused.hyper[0].item()) if self._fused else 0.0
        _log = {"learning_rate": lr, f"eval_{kind.lower()}_loss": ev_loss, f"eval_{kind.lower()}_acc": ev_acc,
                "eval_time_cost": time.time() - start}
        logger.info(str(_log))
        return _log

    def MFP_pretrain_eval(self):
        """trainer.py:355-391"""
        return self._pretrain_eval("MFP")

    def RFD_pretrain_eval(self):
        """trainer.py:479-515"""
        return self._pretrain_eval("RFD")

    def train(self):
        """trainer.py:87-161 (from-scratch / finetune CTR training with per-epoch eval and early stopping on AUC)."""
        a = self.args
        self.train_dataloader = self.get_dataloader(self.train_dataset)
        t_total = int(len(self.train_dataloader) * a.num_train_epochs)
        t_warmup = int(t_total * a.warmup_ratio)
        self._log_header(t_total, t_warmup)
        self._patience, self._stop_training = 0, False
        self.global_step = 0
        self.eval_metrics = []
        self.model.to(self.device)
        fused = self.supports_fused()
        if fused:
            eng = self.fused_step(t_total, t_warmup)
        self.optimizer, self.scheduler = (None, None) if fused else self.get_optimizer(t_total, t_warmup)
        win_loss = torch.zeros(1, device=self.device)
        win_n = 0
        for epoch in range(a.num_train_epochs):
            logger.info(f"-------------------- epoch-{epoch} --------------------")
            self.model.train()
            for X, Y in self.train_dataloader:
                if fused:
                    outputs = self.train_step(X, Y)
                else:
                    if self.optimizer is None:
                        self.optimizer, self.scheduler = self.get_optimizer(t_total, t_warmup)
                    outputs = self.model(input_ids=X, labels=Y)
                    outputs[0].backward()
                    self.optimizer.step()
                    self.scheduler.step()
                    self.optimizer.zero_grad()
                    self.global_step += 1
                win_loss += outputs[0].detach().view(1)
                win_n += 1
                if self.global_step % a.logging_steps == 0:
                    if self._fused is not None:
                        self._fused.check_health()
                    logger.info(f"step = {self.global_step}, {{'window_loss': {win_loss.item() / win_n}}}")
                    win_loss.zero_(); win_n = 0
            self.eval()
            if self._stop_training:
                break

    def eval(self, eval_dataset=None, test_eval=False):
        """trainer.py:163-215: AUC / logloss over the eval split.  The reference copies every batch's logits to the host and calls
        sklearn (roc_auc_score, log_loss); here the logits stay on the device and both metrics come from ops.auc_logloss (sort of
        the scores with the K2 kernels, ties share their average rank like sklearn) — one host read at the end."""
        from . import ops
        loader = self.get_dataloader(self.eval_dataset if eval_dataset is None else eval_dataset, is_training=False)
        logger.info("***** running %s *****", "TEST" if test_eval else "eval")
        self.model.eval()
        logits_all, labels_all = [], []
        with torch.no_grad():
            for X, Y in loader:
                outputs = self.model(input_ids=X, labels=Y)
                logits_all.append(outputs[1].view(-1))
                labels_all.append(Y.view(-1))
            logits = torch.cat(logits_all).float()
            labels = torch.cat(labels_all).float()
            n_pos = int(labels.sum().item())
            if n_pos == 0 or n_pos == labels.numel():   # sklearn.metrics.roc_auc_score raises the same way
                raise ValueError("Only one class present in y_true. ROC AUC score is not defined in that case.")
            auc_t, ll_t = ops.auc_logloss(logits, labels)
            probs_mean = torch.sigmoid(logits.double()).mean()
            auc, ll, avg_logits, avg_probs = (float(v) for v in torch.stack([auc_t, ll_t, logits.double().mean(), probs_mean]).cpu())
        self.eval_metrics.append([auc, ll])
        logger.info(str({"eval_auc": auc, "eval_loss": ll, "avg_logits": avg_logits, "avg_probs": avg_probs}))
        if not test_eval:
            if auc > self.best_eval_auc:
                self.best_eval_auc, self.best_eval_step, self._patience = auc, self.global_step, 0
                self.save_model(self.args.output_dir)
            else:
                self._patience += 1
            if self._patience > self.args.patience:
                self._stop_training = True
        return auc, ll

    # ------------------------------------------------------------------------------------------------ checkpoints
    def save_model(self, model_dir):
        """trainer.py:517-519: torch.save(state_dict) to {model_dir}/{global_step}.model (reference key names)."""
        eng = self._fused
        if eng is not None and getattr(eng, "world", 1) > 1:
            # row-sharded tables: a COLLECTIVE (every rank calls save_model); shards are gathered to the reference's [V, D] layout and
            # rank 0 writes the file, so the checkpoint loads into the reference and into load_for_finetune
            sd = eng.full_state_dict()
            if eng.rank != 0:
                return
        else:
            sd = self.model.state_dict()
        os.makedirs(model_dir, exist_ok=True)
        torch.save({k: v.detach().cpu() for k, v in sd.items()}, os.path.join(model_dir, "{}.model".format(self.global_step)))

    def load_model(self, load_step, model_dir):
        state_dict = torch.load(os.path.join(model_dir, "{}.model".format(load_step)), map_location="cpu")
        self.model.load_state_dict(state_dict)

    def test(self, test_dataset, load_step=-1, model_dir=None):
        if load_step == -1:
            load_step = self.best_eval_step
        self.load_model(load_step, model_dir or self.args.output_dir)
        return self.eval(test_dataset, test_eval=True)
