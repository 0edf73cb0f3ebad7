"""The MEM pretraining step loop.

Drop-in for ``mem/engine_for_pretraining.py``: ``train_one_epoch`` (:108-287) and ``evaluate`` (:289-366),
same signatures, same returned statistics (``lr, min_lr, mlm_acc, loss, loss_scale, weight_decay,
grad_norm`` global averages).  What changed underneath:

* tokenise -> forward -> cross entropy -> backward is ``vit_engine.pretrain_step`` on libmemb kernels
  (bf16 tensor-core GEMMs / attention, fp32 LayerNorm / softmax / CE), with the dVAE tokens coming from
  ``vae_model.DiscreteVAE.get_codebook_indices`` (fp32-faithful);
* data parallelism is the bucketed NCCL all-reduce of ``parallel.GradReducer`` overlapped with backward
  (a ``DistributedDataParallel`` wrapper, if present, is unwrapped and not used);
* clip + AdamW is one fused pass (``optim_factory.FlatAdamW``);
* the four host syncs of the reference step (loss.item(), two synchronize(), mlm_acc.item()) are one
  read of a pinned 4-float buffer per step.

The wandb / matplotlib visualisation branches (:167-217) and the MAE ablation (``MAE=True``) are outside
the hot path and are not reproduced; ``run`` / ``plotting`` are accepted and ignored.
"""
from __future__ import annotations

import math
import sys
from typing import Iterable

import torch

from . import utils
from .parallel import GradReducer, bucket_ranges
from .vit_engine import engine_of, pretrain_step


def _unwrap(model):
    return model.module if hasattr(model, "module") else model


def _bucket_policy(flat):
    """MEMB_DP_BUCKET: minimum elements per overlapped bucket; "all" = one all-reduce after backward (tuning)."""
    import os
    v = os.environ.get("MEMB_DP_BUCKET")
    if not v:
        return {}
    return {"min_bucket_elems": flat.numel if v == "all" else int(float(v))}


def _reducer_for(core):
    """One GradReducer per model (rebuilt if the process group or the flat buffer changes).

    Building it also makes the replicas identical: the reference seeds every rank with ``args.seed + rank``
    (run_mem_pretraining.py:255) and relies on the DistributedDataParallel constructor's broadcast of rank 0's
    parameters (:365-367); here that broadcast is one NCCL call over the flat parameter buffer."""
    if utils.get_world_size() == 1:
        return None
    flat = engine_of(core).flat()
    red = getattr(core, "_memb_reducer", None)
    if red is None or red.grad is not flat.grad:
        red = GradReducer(flat.grad, bucket_ranges(flat, len(core.blocks), **_bucket_policy(flat)))
        red.sync_parameters(flat)
        object.__setattr__(core, "_memb_reducer", red)
    return red


class _StepStats:
    """Device -> pinned host hand-off of (loss, mlm_acc, grad_norm, masked count): one sync per step."""

    def __init__(self, device):
        self.dev = torch.zeros(4, dtype=torch.float32, device=device)
        self.host = torch.zeros(4, dtype=torch.float32).pin_memory() if device.type == "cuda" else torch.zeros(4)

    def read(self, stats, grad_norm):
        n = stats[2].clamp_min(1.0)
        self.dev[0] = stats[0] / n
        self.dev[1] = stats[1] / n
        self.dev[2] = grad_norm if grad_norm is not None else 0.0
        self.dev[3] = stats[2]
        self.host.copy_(self.dev, non_blocking=True)
        torch.cuda.current_stream(self.dev.device).synchronize()
        return self.host.tolist()


_STAGING = {}   # device index -> two slots of {position in the batch tuple: device buffer}


class _PrefetchToDevice:
    """Iterates ``data_loader`` yielding ``((samples, images, mask, n_masked), rest...)`` with the tensors already on
    ``device``: the host->device copies of batch k+1 run on a side stream while step k computes (the reference copies
    inside the step, engine_for_pretraining.py:136-138; with pinned DataLoader memory the copy time leaves the step).

    The copies land in two persistent staging slots per device (batch k in slot k % 2), allocated once and reused by
    every epoch: a fresh ``.to(device)`` per step costs a ``cudaMalloc`` (tens of ms) whenever the caching allocator
    has no free block on the side stream, which was the first one or two steps of every epoch.  The yielded tensors
    are views of a slot and stay valid until the batch after next is loaded."""

    def __init__(self, data_loader, device):
        self.loader, self.device = data_loader, device
        self.side = torch.cuda.Stream(device=device)
        self.slots = _STAGING.setdefault(torch.device(device).index or 0, [{}, {}])
        self.k = 0

    def __len__(self):
        return len(self.loader)

    def _stage(self, slot, pos, t):
        # same shape, dtype AND strides as the source, so that the copy is one plain memcpy (a layout change would
        # make ``copy_`` restage the tensor on the host, synchronously)
        buf = slot.get(pos)
        if buf is None or buf.dtype != t.dtype or buf.shape != t.shape or buf.stride() != t.stride():
            buf = slot[pos] = torch.empty_like(t, device=self.device)   # on the compute stream; a short last batch
        return buf                                                       # re-allocates once, which is harmless

    def _load(self, item):
        batch, rest = item[0], item[1:]
        samples, images, mask = batch
        # masked-patch count from the host copy of the mask: sizes the lm_head / CE kernels without a device sync
        n_masked = int(mask.ne(0).sum()) if not mask.is_cuda else None
        slot = self.slots[self.k % 2]
        self.k += 1
        cur = torch.cuda.current_stream(self.device)
        dst = tuple(self._stage(slot, i, t) for i, t in enumerate((samples, images, mask)))
        # the slot's previous batch (two loads ago) was consumed by work already enqueued on the compute stream
        free = torch.cuda.Event()
        free.record(cur)
        with torch.cuda.stream(self.side):
            self.side.wait_event(free)
            for d, t in zip(dst, (samples, images, mask)):
                d.copy_(t, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self.side)
        return dst, n_masked, ev, rest

    def __iter__(self):
        it = iter(self.loader)
        try:
            nxt = self._load(next(it))
        except StopIteration:
            return
        while nxt is not None:
            (samples, images, mask), n_masked, ev, rest = nxt
            torch.cuda.current_stream(self.device).wait_event(ev)
            try:
                nxt = self._load(next(it))
            except StopIteration:
                nxt = None
            yield ((samples, images, mask, n_masked),) + tuple(rest)


def train_one_epoch(model: torch.nn.Module, d_vae: torch.nn.Module, data_loader: Iterable,
                    optimizer: torch.optim.Optimizer, device: torch.device, epoch: int, loss_scaler, max_norm: float = 0,
                    log_writer=None, lr_scheduler=None, start_steps=None, lr_schedule_values=None,
                    wd_schedule_values=None, run=None, args=None, plotting=False, MAE=False):
    if MAE:
        raise NotImplementedError("the MAE ablation (modeling_mae.py) is not part of the MEM hot path")
    model.train()
    core = _unwrap(model)
    device = torch.device(device)
    metric_logger = utils.MetricLogger(delimiter="  ")
    metric_logger.add_meter("lr", utils.SmoothedValue(window_size=1, fmt="{value:.6f}"))
    metric_logger.add_meter("min_lr", utils.SmoothedValue(window_size=1, fmt="{value:.6f}"))
    header = "Epoch: [{}]".format(epoch)
    start_steps = start_steps or 0
    reducer = _reducer_for(core)
    if hasattr(optimizer, "grad_divisor"):
        optimizer.grad_divisor = float(utils.get_world_size())
    hand_off = _StepStats(device)

    for step, (batch, _) in enumerate(metric_logger.log_every(_PrefetchToDevice(data_loader, device), 10, header)):
        it = start_steps + step
        if lr_schedule_values is not None or wd_schedule_values is not None:
            for group in optimizer.param_groups:
                if lr_schedule_values is not None:
                    group["lr"] = lr_schedule_values[it] * group["lr_scale"]
                if wd_schedule_values is not None and group["weight_decay"] > 0:
                    group["weight_decay"] = wd_schedule_values[it]

        samples, images, bool_masked_pos, n_masked = batch   # already on the device (prefetched on a side stream)

        with torch.no_grad():
            input_ids = d_vae.get_codebook_indices(images).flatten(1)      # [B, P] int64

        optimizer.zero_grad()
        # fp16-pair tokenizer: if an activation left its calibrated range the tokens are invalid.  The check is a device
        # scalar (NaN / 0) dropped into a padding element of the flat gradient: it rides the gradient all-reduce to
        # every rank, makes the global norm non-finite, and FlatAdamW skips a step whose norm is not finite (what the
        # reference's GradScaler does with inf / NaN gradients, utils.py:357-371) -- no update from bad tokens, no
        # extra host sync, no extra collective.
        poison = d_vae.overflow_poison() if hasattr(d_vae, "overflow_poison") else None
        if poison is not None:
            engine_of(core).flat().poison_grad(poison)
        stats = pretrain_step(core, samples, bool_masked_pos.flatten(1), input_ids, cap=n_masked,
                              bucket_hook=reducer.hook if reducer is not None else None)
        if reducer is not None:
            reducer.finish()
        loss = utils.FusedStepLoss(stats[0] / stats[2].clamp_min(1.0))
        grad_norm = loss_scaler(loss, optimizer, clip_grad=max_norm, parameters=core.parameters())
        if reducer is not None and not hasattr(optimizer, "grad_divisor"):
            raise RuntimeError("multi-GPU training needs optim_factory.FlatAdamW (gradients are sum-reduced)")
        loss_scale_value = loss_scaler.state_dict()["scale"]

        loss_value, mlm_acc, grad_norm_value, _ = hand_off.read(stats, grad_norm)
        if hasattr(d_vae, "verify_range") and not d_vae.verify_range(raise_on_overflow=False):
            # this rank's tokens were invalid: the step was skipped on every rank (see above); the tokenizer
            # re-calibrates on its next call
            print("WARNING: dVAE tokenizer left its calibrated fp16 range; step {} skipped, re-calibrating".format(it))
        if not math.isfinite(grad_norm_value) and hasattr(optimizer, "step_skipped"):
            optimizer.step_skipped()          # the update did not happen: keep the bias-correction step count in line
        if not math.isfinite(loss_value):
            print("Loss is {}, stopping training".format(loss_value))
            print("INFO:", "samples", samples.shape, "bool_masked_pos", bool_masked_pos.shape, "images", images.shape)
            sys.exit(1)

        metric_logger.update(mlm_acc=mlm_acc)
        if log_writer is not None:
            log_writer.update(mlm_acc=mlm_acc, head="loss")
        metric_logger.update(loss=loss_value)
        metric_logger.update(loss_scale=loss_scale_value)
        lrs = [g["lr"] for g in optimizer.param_groups]
        min_lr, max_lr = min([10.0] + lrs), max([0.0] + lrs)
        metric_logger.update(lr=max_lr)
        metric_logger.update(min_lr=min_lr)
        weight_decay_value = None
        for g in optimizer.param_groups:
            if g["weight_decay"] > 0:
                weight_decay_value = g["weight_decay"]
        metric_logger.update(weight_decay=weight_decay_value)
        metric_logger.update(grad_norm=grad_norm_value)

        if log_writer is not None:
            log_writer.update(loss=loss_value, head="loss")
            log_writer.update(loss_scale=loss_scale_value, head="opt")
            log_writer.update(lr=max_lr, head="opt")
            log_writer.update(min_lr=min_lr, head="opt")
            log_writer.update(weight_decay=weight_decay_value, head="opt")
            log_writer.update(grad_norm=grad_norm_value, head="opt")
            log_writer.set_step()
        if lr_scheduler is not None:
            lr_scheduler.step_update(start_steps + step)

    metric_logger.synchronize_between_processes()
    print("Averaged stats:", metric_logger)
    return {k: meter.global_avg for k, meter in metric_logger.meters.items()}


@torch.no_grad()
def evaluate(data_loader, model, d_vae, device, args, plotting=False, MAE=False):
    if MAE:
        raise NotImplementedError("the MAE ablation (modeling_mae.py) is not part of the MEM hot path")
    metric_logger = utils.MetricLogger(delimiter="  ")
    model.eval()
    core = _unwrap(model)
    device = torch.device(device)
    hand_off = _StepStats(device)
    for batch in metric_logger.log_every(data_loader, 10, "Test:"):
        samples, images, bool_masked_pos = batch[0]
        n_masked = int(bool_masked_pos.ne(0).sum()) if not bool_masked_pos.is_cuda else None
        images = images.to(device, non_blocking=True)
        samples = samples.to(device, non_blocking=True)
        bool_masked_pos = bool_masked_pos.to(device, non_blocking=True)
        for attempt in range(2):
            input_ids = d_vae.get_codebook_indices(images).flatten(1)
            stats = pretrain_step(core, samples, bool_masked_pos.flatten(1), input_ids, backward=False, cap=n_masked)
            loss_value, mlm_acc, _, _ = hand_off.read(stats, None)
            # fp16-pair tokenizer left its calibrated range: tokens invalid -> tokenise again (it re-calibrates)
            if not hasattr(d_vae, "verify_range") or d_vae.verify_range(raise_on_overflow=attempt == 1):
                break
        metric_logger.update(loss=loss_value)
        metric_logger.meters["mlm_acc"].update(mlm_acc)
    metric_logger.synchronize_between_processes()
    print("* mlm_acc {mlm_acc.global_avg:.3f} loss {losses.global_avg:.3f}".format(mlm_acc=metric_logger.mlm_acc, losses=metric_logger.loss))
    return {k: meter.global_avg for k, meter in metric_logger.meters.items()}
