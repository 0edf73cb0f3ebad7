"""The finetuning / classification loop.

Drop-in for ``mem/engine_for_finetuning.py``: ``train_class_batch`` (:31-34), ``train_one_epoch`` (:42-205) and
``evaluate`` (:208-244), same signatures and returned statistics (``loss, class_acc, loss_scale, lr, min_lr,
weight_decay, grad_norm`` / ``loss, acc1, acc5`` global averages).  Underneath:

* ``model(samples)`` is ``ft_vit`` on libmemb kernels (``vit_engine.classify_forward``: bf16 tensor-core GEMMs and
  attention, fp32 LayerNorm / residual stream), its autograd backward accumulates into the flat gradient buffer, so
  gradient accumulation over ``update_freq`` micro-steps needs nothing extra;
* clip + AdamW is the fused ``optim_factory.FlatAdamW`` pass through ``utils.NativeScalerWithGradNormCount``
  (``update_grad`` as in the reference, utils.py:357-371); bf16 needs no loss scaling (``loss_scale == 1.0``);
* with ``torch.distributed`` initialised the flat gradient is sum-all-reduced (NCCL) once per optimizer step and the
  optimizer divides by the world size (the reference wraps the model in DDP, run_class_finetuning.py);
* the criterion runs on the ``[B, num_classes]`` logits (``LabelSmoothingCrossEntropy`` / ``SoftTargetCrossEntropy`` are
  restated below because timm is absent; ``torch.nn.CrossEntropyLoss`` works as is); ``mixup_fn`` and ``model_ema`` are
  duck-typed exactly as the reference uses them (``mixup_fn(samples, targets)``, ``model_ema.update(model)``).

Not reproduced: the DeepSpeed branch (``loss_scaler is None``), the wandb image logging (:136-156, :188-199) and the
``DUMB_DATA_HUMAN_CLASSIFIER`` matplotlib dump (:62-75).
"""
from __future__ import annotations

import math
import sys
from typing import Iterable, Optional

import torch
import torch.nn.functional as F

from . import utils
from .vit_engine import engine_of

__all__ = ["train_class_batch", "train_one_epoch", "evaluate", "accuracy", "LabelSmoothingCrossEntropy",
           "SoftTargetCrossEntropy", "ModelEma"]


class _SoftCE(torch.autograd.Function):
    """mean_i sum_c t_ic * (logsumexp_i - x_ic) on libmemb (``memb_soft_ce``): loss and d loss / d logits in one
    launch; backward scales the saved gradient."""

    @staticmethod
    def forward(ctx, logits, labels, soft, smoothing):
        from . import _lib
        _lib.require_cuda()
        if not logits.is_cuda:
            raise RuntimeError("mem_b200 criteria run on CUDA tensors only (no CPU path)")
        lib = _lib.load()
        x = logits.detach().contiguous().float()
        B, C = x.shape
        loss = torch.zeros((), dtype=torch.float32, device=x.device)
        dl = torch.empty_like(x) if ctx.needs_input_grad[0] else None
        if soft is not None:
            soft = soft.detach().contiguous().float()
            assert soft.shape == x.shape, "soft targets must have the shape of the logits"
        else:
            labels = labels.detach().contiguous().long()
            assert labels.shape == (B,), "integer targets must be [B]"
        _lib.check(lib.memb_soft_ce(x.data_ptr(), B, C, labels.data_ptr() if soft is None else None,
                                    soft.data_ptr() if soft is not None else None, float(smoothing), loss.data_ptr(),
                                    dl.data_ptr() if dl is not None else None, _lib.stream_ptr(torch, x.device)))
        ctx.dl, ctx.dtype = dl, logits.dtype
        return loss

    @staticmethod
    def backward(ctx, grad_loss):
        return (ctx.dl * grad_loss).to(ctx.dtype), None, None, None


class LabelSmoothingCrossEntropy(torch.nn.Module):
    """timm 0.4.12 ``timm.loss.LabelSmoothingCrossEntropy`` (what run_class_finetuning.py:555-556 picks for
    ``smoothing > 0``): ``(1 - s) * nll + s * mean_c(-log p_c)``, batch mean."""

    def __init__(self, smoothing=0.1):
        super().__init__()
        assert smoothing < 1.0
        self.smoothing, self.confidence = smoothing, 1.0 - smoothing

    def forward(self, x, target):
        return _SoftCE.apply(x, target, None, self.smoothing)


class SoftTargetCrossEntropy(torch.nn.Module):
    """timm 0.4.12 ``timm.loss.SoftTargetCrossEntropy`` (mixup / cutmix targets, run_class_finetuning.py:552-553):
    ``sum_c -t_c log p_c``, batch mean."""

    def forward(self, x, target):
        return _SoftCE.apply(x, None, target, 0.0)


def accuracy(output, target, topk=(1,)):
    """timm 0.4.12 ``timm.utils.accuracy``: top-k accuracies in percent."""
    maxk = max(topk)
    batch_size = target.size(0)
    _, pred = output.topk(maxk, 1, True, True)
    pred = pred.t()
    correct = pred.eq(target.reshape(1, -1).expand_as(pred))
    return [correct[:k].reshape(-1).float().sum(0) * 100.0 / batch_size for k in topk]


class ModelEma:
    """Exponential moving average of the weights with ``timm.utils.ModelEma``'s update rule
    (``ema = decay * ema + (1 - decay) * weight``), kept as one flat fp32 buffer next to the model's."""

    def __init__(self, model, decay=0.9999, device="", resume=""):
        self.decay = decay
        self.flat = engine_of(_unwrap(model)).flat()
        self.shadow = self.flat.data.detach().clone()

    @torch.no_grad()
    def update(self, model):
        self.shadow.lerp_(engine_of(_unwrap(model)).flat().data, 1.0 - self.decay)

    def state_dict(self):
        """Reference key names -> EMA tensors (views of the shadow buffer)."""
        return {n: self.shadow[self.flat.offsets[n]:self.flat.offsets[n] + p.numel()].view_as(p)
                for n, p in self.flat.params.items()}

    @torch.no_grad()
    def load_state_dict(self, state_dict):
        """Restore the EMA weights from a checkpoint's ``model_ema`` entry (mem/utils.py:521-522)."""
        mine = self.state_dict()
        missing = [n for n in mine if n not in state_dict]
        if missing:
            raise KeyError(f"model_ema checkpoint lacks {missing[:3]}{'...' if len(missing) > 3 else ''}")
        for n, dst in mine.items():
            dst.copy_(state_dict[n].to(dst.device, dst.dtype))


def _unwrap(model):
    return model.module if hasattr(model, "module") else model


def train_class_batch(model, samples, target, criterion):
    outputs = model(samples)
    loss = criterion(outputs, target)
    return loss, outputs


def _sync_replicas(core):
    """Once per model: broadcast rank 0's parameters.  The reference seeds each rank with ``args.seed + rank``
    (run_class_finetuning.py:355) and gets identical replicas from the DistributedDataParallel constructor (:487)."""
    if utils.get_world_size() > 1 and not getattr(core, "_memb_replicas_synced", False):
        flat = engine_of(core).flat()
        torch.distributed.broadcast(flat.data, src=0)
        if flat.data.is_cuda:
            flat.refresh_shadow(force=True)
        object.__setattr__(core, "_memb_replicas_synced", True)


def _all_reduce_grads(core, optimizer):
    if utils.get_world_size() > 1:
        if not hasattr(optimizer, "grad_divisor"):
            raise RuntimeError("multi-GPU finetuning needs optim_factory.FlatAdamW (gradients are sum-reduced)")
        # one bucket after the last micro-batch's backward (accumulated gradients are only final then); same wire format
        # as the pretraining loop (bf16 unless MEMB_DP_WIRE=fp32, parallel.default_wire_dtype)
        flat = engine_of(core).flat()
        red = getattr(core, "_memb_ft_reducer", None)
        if red is None or red.grad is not flat.grad:
            from .parallel import GradReducer
            red = GradReducer(flat.grad, [("all", 0, flat.numel)])
            object.__setattr__(core, "_memb_ft_reducer", red)
        red.hook("all")
        red.finish()
        optimizer.grad_divisor = float(utils.get_world_size())


def train_one_epoch(args, model: torch.nn.Module, criterion: torch.nn.Module, data_loader: Iterable,
                    optimizer: torch.optim.Optimizer, device: torch.device, epoch: int, loss_scaler, max_norm: float = 0,
                    model_ema: Optional[ModelEma] = None, mixup_fn=None, log_writer=None, start_steps=None,
                    lr_schedule_values=None, wd_schedule_values=None, num_training_steps_per_epoch=None,
                    update_freq=None):
    if loss_scaler is None:
        raise NotImplementedError("the DeepSpeed branch of the reference loop (loss_scaler is None) is not reproduced")
    model.train(True)
    core = _unwrap(model)
    metric_logger = utils.MetricLogger(delimiter="  ")
    metric_logger.add_meter("lr", utils.SmoothedValue(window_size=1, fmt="{value:.6f}"))
    metric_logger.add_meter("min_lr", utils.SmoothedValue(window_size=1, fmt="{value:.6f}"))
    header = "Epoch: [{}]".format(epoch)
    update_freq = update_freq or 1
    start_steps = start_steps or 0
    if num_training_steps_per_epoch is None:
        num_training_steps_per_epoch = math.inf
    _sync_replicas(core)
    optimizer.zero_grad()

    for data_iter_step, (samples, targets) in enumerate(metric_logger.log_every(data_loader, 10, header)):
        step = data_iter_step // update_freq
        if step >= num_training_steps_per_epoch:
            continue
        it = start_steps + step
        # (operator precedence kept from the reference, engine_for_finetuning.py:82)
        if lr_schedule_values is not None or wd_schedule_values is not None and data_iter_step % update_freq == 0:
            for param_group in optimizer.param_groups:
                if lr_schedule_values is not None:
                    param_group["lr"] = lr_schedule_values[it] * param_group.get("lr_scale", 1.0)
                if wd_schedule_values is not None and param_group["weight_decay"] > 0:
                    param_group["weight_decay"] = wd_schedule_values[it]

        samples = samples.to(device, non_blocking=True)
        targets = targets.to(device, non_blocking=True)
        if mixup_fn is not None:
            samples, targets = mixup_fn(samples, targets)

        loss, output = train_class_batch(model, samples, targets, criterion)
        loss_value = loss.item()
        if not math.isfinite(loss_value):
            print("Loss is {}, stopping training".format(loss_value))
            sys.exit(1)

        loss = loss / update_freq
        update = (data_iter_step + 1) % update_freq == 0
        loss.backward()                                   # kernels accumulate into the flat gradient buffer
        if update:
            _all_reduce_grads(core, optimizer)
        grad_norm = loss_scaler(utils.FusedStepLoss(loss.detach()), optimizer, clip_grad=max_norm,
                                parameters=core.parameters(), update_grad=update)
        if update:
            optimizer.zero_grad()
            if model_ema is not None:
                model_ema.update(model)
        loss_scale_value = loss_scaler.state_dict()["scale"]
        torch.cuda.synchronize()

        lrs = [g["lr"] for g in optimizer.param_groups]
        decays = [g["weight_decay"] for g in optimizer.param_groups if g["weight_decay"] > 0]
        # one record per micro-step, in the reference's meter order; None entries (accuracy under mixup, the norm of a
        # non-updating micro-step, no decayed group) are skipped by the logger exactly like the reference's
        record = {"loss": loss_value,
                  "class_acc": (output.max(-1)[-1] == targets).float().mean() if mixup_fn is None else None,
                  "loss_scale": loss_scale_value, "lr": max([0.0] + lrs), "min_lr": min([10.0] + lrs),
                  "weight_decay": decays[-1] if decays else None, "grad_norm": grad_norm}
        for name, value in record.items():
            metric_logger.update(**{name: value})
        if log_writer is not None:
            heads = {"loss": "loss", "class_acc": "loss"}
            for name, value in record.items():
                log_writer.update(head=heads.get(name, "opt"), **{name: value})
            log_writer.set_step()

    metric_logger.synchronize_between_processes()
    print("Averaged stats:", metric_logger)
    return {k: meter.global_avg for k, meter in metric_logger.meters.items()}


@torch.no_grad()
def evaluate(data_loader, model, device):
    criterion = torch.nn.CrossEntropyLoss()
    metric_logger = utils.MetricLogger(delimiter="  ")
    header = "Test:"
    model.eval()
    for batch in metric_logger.log_every(data_loader, 10, header):
        images = batch[0].to(device, non_blocking=True)
        target = batch[-1].to(device, non_blocking=True)
        output = model(images)
        loss = criterion(output.float(), target)
        n = min(5, len(output[0]))
        acc1, acc5 = accuracy(output, target, topk=(1, n))
        batch_size = images.shape[0]
        metric_logger.update(loss=loss.item())
        metric_logger.meters["acc1"].update(acc1.item(), n=batch_size)
        metric_logger.meters["acc5"].update(acc5.item(), n=batch_size)
    metric_logger.synchronize_between_processes()
    print("* Acc@1 {top1.global_avg:.3f} Acc@5 {top5.global_avg:.3f} loss {losses.global_avg:.3f}"
          .format(top1=metric_logger.acc1, top5=metric_logger.acc5, losses=metric_logger.loss))
    return {k: meter.global_avg for k, meter in metric_logger.meters.items()}
