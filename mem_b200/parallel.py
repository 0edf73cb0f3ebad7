"""Data-parallel gradient exchange for the flat gradient buffer.

Replaces ``torch.nn.parallel.DistributedDataParallel`` as used by the reference
(mem/run_mem_pretraining.py:365-367: ``DistributedDataParallel(model, device_ids=[gpu])``).  The MEM step
is pure data parallel: samples are independent up to the loss, weights / optimizer state / tokenizer are
replicated, and the ONE exchange per step is the gradient average.

Because gradients live in one flat fp32 buffer ordered by when backward finishes with them
(``VitEngine.param_order``: head, blocks last-to-first, embedding), a bucket is a contiguous slice:
``hook(tag)`` -- called by ``VitEngine.backward_pretrain`` after the head, after every block and after the
embedding -- launches an NCCL all-reduce of the slice that just became final on a side stream, so the
exchange overlaps the rest of backward over NVLink 5 / NVSwitch.  The reduction is a SUM; the optimizer
divides by the world size (``FlatAdamW.grad_divisor``), which also keeps the global-norm clip exact.
The reference weights every rank equally whatever its masked-token count (DDP mean of per-rank mean
losses); that behaviour is kept.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def bucket_ranges(flat, depth, min_bucket_elems=8 << 20):
    """[(tag, start, end)] slices of the flat buffer that become final at each backward hook.

    Tags: "head", block index (depth-1 .. 0), "embed".  Adjacent slices are merged until a bucket holds at
    least ``min_bucket_elems`` elements (launch latency, not link count, is what matters on NVSwitch)."""
    def first_offset(prefixes):
        offs = [flat.offsets[n] for n in flat.names if n.startswith(prefixes)]
        return min(offs) if offs else None

    marks = []   # (tag, end offset of what is final once `tag` fires)
    for i in reversed(range(depth)):
        start = first_offset((f"blocks.{i}.",))
        if start is not None:
            marks.append((("head" if i == depth - 1 else i + 1), start))
    first_tail = first_offset(("patch_embed.", "cls_token", "mask_token", "pos_embed", "rel_pos_bias."))
    marks.append((0, first_tail if first_tail is not None else flat.numel))
    marks.append(("embed", flat.numel))
    ranges, lo = [], 0
    for tag, hi in marks:
        if hi > lo:
            ranges.append([tag, lo, hi])
            lo = hi
    merged = []
    for tag, a, b in ranges:     # merge small buckets forward (a merged bucket fires at its LAST tag)
        if merged and merged[-1][2] - merged[-1][1] < min_bucket_elems:
            merged[-1][0], merged[-1][2] = tag, b
        else:
            merged.append([tag, a, b])
    return [tuple(r) for r in merged]


class GradReducer:
    """Bucketed, overlapped all-reduce(SUM) of ``flat.grad`` across the default process group."""

    def __init__(self, flat_grad, ranges, group=None):
        self.grad, self.ranges, self.group = flat_grad, list(ranges), group
        self.by_tag = {r[0]: r for r in self.ranges}
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.cuda = flat_grad.is_cuda
        self.stream = torch.cuda.Stream(device=flat_grad.device) if self.cuda else None
        self.pending = []
        self.launched = 0

    def sync_parameters(self, flat, src=0):
        """Broadcast rank ``src``'s parameters (what DistributedDataParallel's constructor does) and refresh the bf16
        shadow the GEMMs read.  ``flat``: the model's ``FlatParams``."""
        if self.world == 1:
            return
        dist.broadcast(flat.data, src=src, group=self.group)
        if flat.data.is_cuda:
            flat.refresh_shadow(force=True)

    def hook(self, tag):
        r = self.by_tag.get(tag)
        if r is None or self.world == 1:
            return
        view = self.grad[r[1]:r[2]]
        if self.cuda:
            self.stream.wait_stream(torch.cuda.current_stream(self.grad.device))
            with torch.cuda.stream(self.stream):
                self.pending.append(dist.all_reduce(view, op=dist.ReduceOp.SUM, group=self.group, async_op=True))
        else:
            self.pending.append(dist.all_reduce(view, op=dist.ReduceOp.SUM, group=self.group, async_op=True))
        self.launched += 1

    def finish(self):
        """Make the compute stream wait for every launched bucket (no host sync on CUDA)."""
        for w in self.pending:
            w.wait()
        self.pending.clear()
        if self.cuda:
            torch.cuda.current_stream(self.grad.device).wait_stream(self.stream)


def shard_indices(n_items, rank, world, epoch=0, seed=0, shuffle=True, drop_last=False):
    """``torch.utils.data.DistributedSampler`` partition (mem/run_mem_pretraining.py:307-309): a per-epoch
    seeded permutation padded to a multiple of ``world``; rank r takes r, r+world, ..."""
    if shuffle:
        g = torch.Generator().manual_seed(seed + epoch)
        idx = torch.randperm(n_items, generator=g).tolist()
    else:
        idx = list(range(n_items))
    if drop_last:
        idx = idx[: n_items // world * world]
    else:
        total = (n_items + world - 1) // world * world
        while len(idx) < total:                       # DistributedSampler repeats from the start
            idx += idx[: total - len(idx)]
    return idx[rank::world]
