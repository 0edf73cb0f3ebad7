"""Step utilities of the MEM pretraining engine.

Drop-in for the hot-path part of ``mem/utils.py``: ``SmoothedValue`` / ``MetricLogger`` (:34-183),
distributed helpers + ``init_distributed_mode`` (:220-299), ``NativeScalerWithGradNormCount`` and
``get_grad_norm_`` (:351-392), ``cosine_scheduler`` (:395-412), ``save_model`` / ``auto_load_model``
(:425-537) and ``create_d_vae`` / ``get_event_vae`` (:559-578).

Differences that follow from the B200 design (north_star: bf16, no GradScaler):
* the loss scale is the constant 1.0 -- ``loss_scaler.state_dict()["scale"]`` still exists;
* with the flat-buffer optimizer (``optim_factory.FlatAdamW``) unscale + global-norm clip + AdamW is one
  pass over flat fp32 buffers on libmemb kernels instead of 189-tensor foreach loops.
"""
from __future__ import annotations

import datetime
import glob
import math
import os
import time
from collections import defaultdict, deque

import numpy as np
import torch
import torch.distributed as dist

inf = math.inf


# ------------------------------------------------------------------------------------------ meters
class SmoothedValue:
    """Window + global statistics of a scalar series."""

    def __init__(self, window_size=20, fmt=None):
        self.deque = deque(maxlen=window_size)
        self.total, self.count = 0.0, 0
        self.fmt = fmt or "{median:.4f} ({global_avg:.4f})"

    def update(self, value, n=1):
        self.deque.append(value)
        self.count += n
        self.total += value * n

    def synchronize_between_processes(self):
        """Sum (count, total) over ranks; the window is left per-rank, like the reference."""
        if not is_dist_avail_and_initialized():
            return
        t = _meter_tensor([self.count, self.total])
        dist.barrier()
        dist.all_reduce(t)
        c, tot = t.tolist()
        self.count, self.total = int(c), tot

    @property
    def median(self):
        return torch.tensor(list(self.deque)).median().item()

    @property
    def avg(self):
        return torch.tensor(list(self.deque), dtype=torch.float32).mean().item()

    @property
    def global_avg(self):
        return float("nan") if self.count == 0 else self.total / self.count

    @property
    def max(self):
        return max(self.deque) if self.deque else float("nan")

    @property
    def value(self):
        return self.deque[-1] if self.deque else float("nan")

    def __str__(self):
        return self.fmt.format(median=self.median, avg=self.avg, global_avg=self.global_avg, max=self.max, value=self.value)


def _meter_tensor(values):
    dev = "cuda" if (torch.cuda.is_available() and dist.get_backend() == "nccl") else "cpu"
    return torch.tensor(values, dtype=torch.float64, device=dev)


class MetricLogger:
    def __init__(self, delimiter="\t"):
        self.meters = defaultdict(SmoothedValue)
        self.delimiter = delimiter

    def update(self, **kwargs):
        for k, v in kwargs.items():
            if v is None:
                continue
            if isinstance(v, torch.Tensor):
                v = v.item()
            assert isinstance(v, (float, int))
            self.meters[k].update(v)

    def __getattr__(self, attr):
        meters = self.__dict__.get("meters", {})
        if attr in meters:
            return meters[attr]
        raise AttributeError(f"'{type(self).__name__}' object has no attribute '{attr}'")

    def __str__(self):
        return self.delimiter.join(f"{name}: {meter}" for name, meter in self.meters.items())

    def synchronize_between_processes(self):
        """One barrier + ONE all-reduce for all meters (the reference does one pair per meter)."""
        if not is_dist_avail_and_initialized() or not self.meters:
            return
        names = list(self.meters)
        t = _meter_tensor([x for n in names for x in (self.meters[n].count, self.meters[n].total)])
        dist.barrier()
        dist.all_reduce(t)
        vals = t.tolist()
        for i, n in enumerate(names):
            self.meters[n].count, self.meters[n].total = int(vals[2 * i]), vals[2 * i + 1]

    def add_meter(self, name, meter):
        self.meters[name] = meter

    def log_every(self, iterable, print_freq, header=None):
        header = header or ""
        n = len(iterable)
        start = end = time.time()
        iter_time, data_time = SmoothedValue(fmt="{avg:.4f}"), SmoothedValue(fmt="{avg:.4f}")
        width = len(str(n))
        for i, obj in enumerate(iterable):
            data_time.update(time.time() - end)
            yield obj
            iter_time.update(time.time() - end)
            if i % print_freq == 0 or i == n - 1:
                eta = str(datetime.timedelta(seconds=int(iter_time.global_avg * (n - i))))
                parts = [header, f"[{i:{width}d}/{n}]", f"eta: {eta}", str(self), f"time: {iter_time}", f"data: {data_time}"]
                if torch.cuda.is_available():
                    parts.append(f"max mem: {torch.cuda.max_memory_allocated() / 2 ** 20:.0f}")
                print(self.delimiter.join(parts))
            end = time.time()
        total = time.time() - start
        print(f"{header} Total time: {datetime.timedelta(seconds=int(total))} ({total / max(n, 1):.4f} s / it)")


# ------------------------------------------------------------------------------------- distributed
def setup_for_distributed(is_master):
    """Silence ``print`` on non-master ranks (``force=True`` overrides)."""
    import builtins
    if getattr(builtins.print, "_memb_patched", False):
        return
    plain = builtins.print

    def rank_print(*args, **kwargs):
        if kwargs.pop("force", False) or is_master:
            plain(*args, **kwargs)

    rank_print._memb_patched = True
    builtins.print = rank_print


def is_dist_avail_and_initialized():
    return dist.is_available() and dist.is_initialized()


def get_world_size():
    return dist.get_world_size() if is_dist_avail_and_initialized() else 1


def get_rank():
    return dist.get_rank() if is_dist_avail_and_initialized() else 0


def is_main_process():
    return get_rank() == 0


def save_on_master(*args, **kwargs):
    if is_main_process():
        torch.save(*args, **kwargs)


def init_distributed_mode(args):
    """One process per GPU under torchrun (env:// rendezvous), NCCL over NVLink 5 / NVSwitch."""
    if "RANK" in os.environ and "WORLD_SIZE" in os.environ:
        args.rank, args.world_size = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
        args.gpu = int(os.environ.get("LOCAL_RANK", 0))
    elif "SLURM_PROCID" in os.environ:
        args.rank = int(os.environ["SLURM_PROCID"])
        args.gpu = args.rank % max(torch.cuda.device_count(), 1)
    else:
        print("Not using distributed mode")
        args.distributed = False
        return
    args.distributed = True
    torch.cuda.set_device(args.gpu)
    args.dist_backend = "nccl"
    args.dist_url = getattr(args, "dist_url", "env://")
    print(f"| distributed init (rank {args.rank}): {args.dist_url}, gpu {args.gpu}", flush=True)
    dist.init_process_group(backend=args.dist_backend, init_method=args.dist_url, world_size=args.world_size,
                            rank=args.rank, device_id=torch.device("cuda", args.gpu))
    dist.barrier()
    setup_for_distributed(args.rank == 0)


def cleanup_distributed_mode():
    if is_dist_avail_and_initialized():
        dist.destroy_process_group()


# ------------------------------------------------------------------------------- scaler / grad norm
class FusedStepLoss:
    """What ``vit_engine.pretrain_step`` hands to the loss scaler: backward has already run inside the fused
    step (gradients are in the flat buffer), ``value`` is the device scalar of the mean loss."""

    def __init__(self, value):
        self.value = value

    def item(self):
        return float(self.value.item())


class NativeScalerWithGradNormCount:
    """Same call shape as the reference scaler; bf16 training needs no loss scaling, so scale == 1.0."""
    state_dict_key = "amp_scaler"

    def __init__(self):
        self._scale = 1.0

    def __call__(self, loss, optimizer, clip_grad=None, parameters=None, create_graph=False, update_grad=True):
        from .optim_factory import FlatAdamW
        if not isinstance(loss, FusedStepLoss):
            loss.backward(create_graph=create_graph)
        if not update_grad:
            return None
        if isinstance(optimizer, FlatAdamW):
            return optimizer.step(max_norm=clip_grad if clip_grad else 0.0)
        if clip_grad is not None and clip_grad > 0:
            assert parameters is not None
            norm = torch.nn.utils.clip_grad_norm_(parameters, clip_grad)
        else:
            norm = get_grad_norm_(parameters)
        optimizer.step()
        return norm

    def state_dict(self):
        return {"scale": self._scale}

    def load_state_dict(self, state_dict):
        self._scale = 1.0


def get_grad_norm_(parameters, norm_type: float = 2.0) -> torch.Tensor:
    if isinstance(parameters, torch.Tensor):
        parameters = [parameters]
    grads = [p.grad.detach() for p in parameters if p.grad is not None]
    if not grads:
        return torch.tensor(0.0)
    if float(norm_type) == inf:
        return max(g.abs().max() for g in grads)
    return torch.norm(torch.stack([torch.norm(g, float(norm_type)) for g in grads]), float(norm_type))


def cosine_scheduler(base_value, final_value, epochs, niter_per_ep, warmup_epochs=0, start_warmup_value=0, warmup_steps=-1):
    """Per-iteration schedule: linear warm-up then half-cosine decay to ``final_value``."""
    warmup_iters = warmup_steps if warmup_steps > 0 else warmup_epochs * niter_per_ep
    print("Set warmup steps = %d" % warmup_iters)
    warm = np.linspace(start_warmup_value, base_value, warmup_iters) if warmup_epochs > 0 else np.array([])
    n = epochs * niter_per_ep - warmup_iters
    i = np.arange(n)
    decay = final_value + 0.5 * (base_value - final_value) * (1 + np.cos(math.pi * i / n)) if n > 0 else np.array([])
    schedule = np.concatenate((warm, decay))
    assert len(schedule) == epochs * niter_per_ep
    return schedule


# ----------------------------------------------------------------------------------- checkpoints
def save_model(args, epoch, model, model_without_ddp, optimizer, loss_scaler, model_ema=None):
    """``checkpoint-{epoch}.pth`` = {model, optimizer, epoch, scaler, args[, model_ema]} on rank 0 (reference
    format, mem/utils.py:425-442; ``epoch`` may be the string ``"best"``)."""
    path = os.path.join(args.output_dir, f"checkpoint-{epoch}.pth")
    to_save = {"model": model_without_ddp.state_dict(), "optimizer": optimizer.state_dict(), "epoch": epoch,
               "scaler": loss_scaler.state_dict() if loss_scaler is not None else None, "args": args}
    if model_ema is not None:
        to_save["model_ema"] = {k: v.detach().clone() for k, v in model_ema.state_dict().items()}
    save_on_master(to_save, path)


def auto_load_model(args, model, model_without_ddp, optimizer, loss_scaler, model_ema=None):
    """Resume from ``args.resume`` or, with ``args.auto_resume``, from the newest ``checkpoint-*.pth``."""
    if getattr(args, "auto_resume", False) and not getattr(args, "resume", ""):
        found = [int(os.path.basename(p)[len("checkpoint-"):-4]) for p in glob.glob(os.path.join(args.output_dir, "checkpoint-*.pth"))
                 if os.path.basename(p)[len("checkpoint-"):-4].isdigit()]
        if found:
            args.resume = os.path.join(args.output_dir, f"checkpoint-{max(found)}.pth")
            print("Auto resume checkpoint: %s" % args.resume)
    if not getattr(args, "resume", ""):
        return
    ckpt = torch.load(args.resume, map_location="cpu", weights_only=False)
    model_without_ddp.load_state_dict(ckpt["model"])
    print("Resume checkpoint %s" % args.resume)
    if "optimizer" in ckpt and "epoch" in ckpt:
        optimizer.load_state_dict(ckpt["optimizer"])
        epoch = ckpt["epoch"] if ckpt["epoch"] != "best" else args.epochs      # mem/utils.py:519
        args.start_epoch = epoch + 1
        if getattr(args, "model_ema", False) and model_ema is not None:        # mem/utils.py:521-522
            model_ema.load_state_dict(ckpt["model_ema"])
        if loss_scaler is not None and ckpt.get("scaler") is not None:
            loss_scaler.load_state_dict(ckpt["scaler"])
        print("With optim & sched!")


# ------------------------------------------------------------------------- pretrain -> finetune seam
def load_state_dict(model, state_dict, prefix="", ignore_missing="relative_position_index"):
    """Non-strict load with the reference's report (mem/utils.py:302-348): keys of ``state_dict`` under ``prefix``
    are copied into ``model``; missing keys containing any ``|``-separated ``ignore_missing`` fragment are reported
    separately."""
    own = model.state_dict()
    missing, unexpected, errors, used = [], [], [], set()
    with torch.no_grad():
        for name, dst in own.items():
            key = prefix + name
            if key not in state_dict:
                missing.append(name)
                continue
            used.add(key)
            src = state_dict[key]
            if tuple(src.shape) != tuple(dst.shape):
                errors.append("size mismatch for {}: copying a param with shape {} from checkpoint, the shape in current "
                              "model is {}.".format(name, tuple(src.shape), tuple(dst.shape)))
                continue
            dst.copy_(src)            # parameters stay views of the flat buffer
    unexpected = [k for k in state_dict if k.startswith(prefix) and k not in used]
    fragments = ignore_missing.split("|")
    ignored = [k for k in missing if any(f in k for f in fragments)]
    missing = [k for k in missing if k not in ignored]
    if missing:
        print("Weights of {} not initialized from pretrained model: {}".format(model.__class__.__name__, missing))
    if unexpected:
        print("Weights from pretrained model not used in {}: {}".format(model.__class__.__name__, unexpected))
    if ignored:
        print("Ignored weights of {} not initialized from pretrained model: {}".format(model.__class__.__name__, ignored))
    if errors:
        print("\n".join(errors))
    return missing, unexpected


def _interp_rel_pos_table(table, src_size, dst_size, num_extra_tokens):
    """Resize a ``[(2s-1)^2 + extra, heads]`` relative-position table to ``dst_size`` the way the reference does
    (mem/utils.py:660-706): source offsets placed on a geometric progression, bicubic spline through them
    (``scipy.interpolate.interp2d(kind="cubic")``; on a regular grid that is ``RectBivariateSpline(kx=ky=3, s=0)``,
    the replacement SciPy names for the removed function), evaluated at the integer target offsets."""
    from scipy import interpolate
    extra = table[-num_extra_tokens:, :]
    body = table[:-num_extra_tokens, :]

    def geometric_progression(a, r, n):
        return a * (1.0 - r ** n) / (1.0 - r)
    left, right = 1.01, 1.5
    while right - left > 1e-6:
        q = (left + right) / 2.0
        if geometric_progression(1, q, src_size // 2) > dst_size // 2:
            right = q
        else:
            left = q
    dis, cur = [], 1
    for i in range(src_size // 2):
        dis.append(cur)
        cur += q ** (i + 1)
    x = [-d for d in reversed(dis)] + [0] + dis
    t = dst_size // 2.0
    dx = np.arange(-t, t + 0.1, 1.0)
    print("Original positions = %s" % str(x))
    print("Target positions = %s" % str(dx))
    heads = []
    for i in range(body.shape[1]):
        z = body[:, i].view(src_size, src_size).float().cpu().numpy()
        spline = interpolate.RectBivariateSpline(x, x, z.T, kx=3, ky=3, s=0)      # interp2d(x, y, z): z[j, i] = f(x[i], y[j])
        heads.append(torch.Tensor(spline(dx, dx).T).contiguous().view(-1, 1).to(table.device))
    return torch.cat((torch.cat(heads, dim=-1), extra), dim=0)


def finetune(args, model):
    """Start a finetuning model from a pretraining checkpoint (mem/utils.py:613-732): pick the state dict by
    ``args.model_key``, drop a head of the wrong shape, expand the shared relative-position table to one table per
    block, resize tables / absolute position embeddings when the patch grid changed, then load non-strictly."""
    ckpt = torch.load(args.finetune, map_location="cpu", weights_only=False)
    print("Load ckpt from %s" % args.finetune)
    checkpoint_model = None
    for model_key in args.model_key.split("|"):
        if model_key in ckpt:
            checkpoint_model = ckpt[model_key]
            print("Load state_dict by model_key = %s" % model_key)
            break
    if checkpoint_model is None:
        checkpoint_model = ckpt
    checkpoint_model = remap_pretrain_state_dict(checkpoint_model, model)
    load_state_dict(model, checkpoint_model, prefix=getattr(args, "model_prefix", ""))


def remap_pretrain_state_dict(checkpoint_model, model):
    """The state-dict surgery of ``finetune`` (mem/utils.py:631-730) as a function of (checkpoint dict, target model)."""
    checkpoint_model = dict(checkpoint_model)
    state_dict = model.state_dict()
    for k in ["head.weight", "head.bias"]:
        if k in checkpoint_model and checkpoint_model[k].shape != state_dict[k].shape:
            print(f"Removing key {k} from pretrained checkpoint")
            del checkpoint_model[k]
    if model.use_rel_pos_bias and "rel_pos_bias.relative_position_bias_table" in checkpoint_model:
        print("Expand the shared relative position embedding to each transformer block. ")
        shared = checkpoint_model.pop("rel_pos_bias.relative_position_bias_table")
        for i in range(model.get_num_layers()):
            checkpoint_model["blocks.%d.attn.relative_position_bias_table" % i] = shared.clone()
    for key in list(checkpoint_model.keys()):
        if "relative_position_index" in key:
            checkpoint_model.pop(key)
        if "relative_position_bias_table" in key:
            table = checkpoint_model[key]
            src_num_pos, _ = table.size()
            dst_num_pos, _ = state_dict[key].size()
            dst_patch_shape = model.patch_embed.patch_shape
            num_extra_tokens = dst_num_pos - (dst_patch_shape[0] * 2 - 1) * (dst_patch_shape[1] * 2 - 1)
            src_size = int((src_num_pos - num_extra_tokens) ** 0.5)
            dst_size = int((dst_num_pos - num_extra_tokens) ** 0.5)
            print(dst_patch_shape, src_size, dst_size)
            if src_size != dst_size:
                print("Position interpolate for %s from %dx%d to %dx%d" % (key, src_size, src_size, dst_size, dst_size))
                checkpoint_model[key] = _interp_rel_pos_table(table, src_size, dst_size, num_extra_tokens)
    if "pos_embed" in checkpoint_model:
        pos = checkpoint_model["pos_embed"]
        dim = pos.shape[-1]
        num_patches = model.patch_embed.num_patches
        num_extra_tokens = model.pos_embed.shape[-2] - num_patches
        orig_size = int((pos.shape[-2] - num_extra_tokens) ** 0.5)
        new_size = int(num_patches ** 0.5)
        if orig_size != new_size:
            print("Position interpolate from %dx%d to %dx%d" % (orig_size, orig_size, new_size, new_size))
            extra = pos[:, :num_extra_tokens]
            grid = pos[:, num_extra_tokens:].reshape(-1, orig_size, orig_size, dim).permute(0, 3, 1, 2)
            grid = torch.nn.functional.interpolate(grid, size=(new_size, new_size), mode="bicubic", align_corners=False)
            checkpoint_model["pos_embed"] = torch.cat((extra, grid.permute(0, 2, 3, 1).flatten(1, 2)), dim=1)
    return checkpoint_model


# -------------------------------------------------------------------------------------- tokenizer
def create_d_vae(weight_path, d_vae_type, image_size, device):
    if d_vae_type == "event":
        return get_event_vae(weight_path, image_size, device)
    raise NotImplementedError()   # "dall-e" raises in the reference as well (utils.py:568-569)


def get_event_vae(weight_path, image_size, device):
    """``.pt`` written by the reference's dVAE trainer: {hparams, weights, ...} (train_vae.py:271-290)."""
    from .vae_model import DiscreteVAE
    obj = torch.load(weight_path, map_location="cpu", weights_only=False)
    vae = DiscreteVAE(**obj["hparams"]).to(device)
    vae.load_state_dict(obj["weights"])
    print(f"loaded event vae from {weight_path}")
    return vae
