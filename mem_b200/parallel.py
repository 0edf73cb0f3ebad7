"""Data-parallel gradient exchange for the flat gradient buffer.

Replaces ``torch.nn.parallel.DistributedDataParallel`` as used by the reference
(mem/run_mem_pretraining.py:365-367: ``DistributedDataParallel(model, device_ids=[gpu])``).  The MEM step
is pure data parallel: samples are independent up to the loss, weights / optimizer state / tokenizer are
replicated, and the ONE exchange per step is the gradient average.

Because gradients live in one flat fp32 buffer ordered by when backward finishes with them
(``VitEngine.param_order``: head, blocks last-to-first, embedding), a bucket is a contiguous slice:
``hook(tag)`` -- called by ``VitEngine.backward_pretrain`` after the head, after every block and after the
embedding -- launches an NCCL all-reduce of the slice that just became final on a side stream, so the
exchange overlaps the rest of backward over NVLink 5 / NVSwitch.  The wire format is bf16 by default (fp32 on
request, ``default_wire_dtype``).  The reduction is a SUM; the optimizer
divides by the world size (``FlatAdamW.grad_divisor``), which also keeps the global-norm clip exact.
The reference weights every rank equally whatever its masked-token count (DDP mean of per-rank mean
losses); that behaviour is kept.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def bucket_ranges(flat, depth, min_bucket_elems=32 << 20, tail_blocks=2):
    """[(tag, start, end)] slices of the flat buffer that become final at each backward hook.

    Tags: "head", block index (depth-1 .. 0), "embed".  Adjacent slices are merged until a bucket holds at
    least ``min_bucket_elems`` elements (32 M by default: every NCCL kernel that runs under backward takes SMs from
    one-CTA-per-SM persistent GEMMs, so few large collectives beat many small ones -- 8 GPUs, ViT segment: 8 M buckets
    22.87 ms, 32 M 22.69 ms, one all-reduce after backward 22.49 ms, 21.2 ms on one GPU; profiles/r02_scale_n8_sweep.txt) -- except at the
    end of backward: the slices of the last ``tail_blocks`` blocks and of the embedding are never merged into their
    neighbours, because nothing is left to hide a late all-reduce behind: block 1 and block 0 each fire alone and the
    bucket that fires at "embed" is only the embedding's own parameters (0.4 M elements for ViT-B/16; a forward merge
    made it block 0 + embedding = 7.5 M elements = 30 MB of fp32, fully exposed)."""
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
    if min_bucket_elems >= flat.numel:            # one bucket: everything at the end
        return [("embed", 0, flat.numel)]
    alone = {"embed"} | {i for i in range(tail_blocks)}
    merged = []
    for tag, a, b in ranges:     # merge small buckets forward (a merged bucket fires at its LAST tag)
        if merged and merged[-1][0] not in alone and tag not in alone and merged[-1][2] - merged[-1][1] < min_bucket_elems:
            merged[-1][0], merged[-1][2] = tag, b
        else:
            merged.append([tag, a, b])
    return [tuple(r) for r in merged]


def default_wire_dtype():
    """Element type of the gradient exchange: bf16 (half the NVLink bytes and half the time the NCCL
    kernels share the SMs with backward) unless MEMB_DP_WIRE=fp32 asks for the reference's fp32 all-reduce (DDP
    reduces the fp32 ``.grad`` buffers under autocast)."""
    import os
    return torch.float32 if os.environ.get("MEMB_DP_WIRE", "bf16").lower() in ("fp32", "float32", "f32") else torch.bfloat16


class GradReducer:
    """Bucketed, overlapped all-reduce(SUM) of ``flat.grad`` across the default process group.

    ``wire_dtype`` bf16: a bucket is rounded to bf16 into a staging buffer, summed over the ranks in bf16 and written
    back to the fp32 gradient, all on the side stream (the rounding, 2^-9 relative per element, is of the size of the
    bf16 GEMM operand rounding that produced the gradient).  fp32: the slice of the gradient buffer is reduced in place."""

    def __init__(self, flat_grad, ranges, group=None, wire_dtype=None):
        self.grad, self.ranges, self.group = flat_grad, list(ranges), group
        self.by_tag = {r[0]: r for r in self.ranges}
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.cuda = flat_grad.is_cuda
        self.stream = torch.cuda.Stream(device=flat_grad.device) if self.cuda else None
        self.wire_dtype = wire_dtype if wire_dtype is not None else default_wire_dtype()
        self.staging = None
        if self.wire_dtype != flat_grad.dtype and self.world > 1:
            self.staging = torch.empty(flat_grad.numel(), dtype=self.wire_dtype, device=flat_grad.device)
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

    def _reduce(self, a, b):
        view = self.grad[a:b]
        if self.staging is None:
            return dist.all_reduce(view, op=dist.ReduceOp.SUM, group=self.group, async_op=True), None
        wire = self.staging[a:b]
        wire.copy_(view)                               # fp32 -> bf16, round to nearest even
        return dist.all_reduce(wire, op=dist.ReduceOp.SUM, group=self.group, async_op=True), (view, wire)

    def hook(self, tag):
        r = self.by_tag.get(tag)
        if r is None or self.world == 1:
            return
        if self.cuda:
            self.stream.wait_stream(torch.cuda.current_stream(self.grad.device))
            with torch.cuda.stream(self.stream):
                work, back = self._reduce(r[1], r[2])
                if back is not None:                   # stream-ordered after the collective, still on the side stream
                    work.wait()
                    back[0].copy_(back[1])
                    work = None
        else:
            work, back = self._reduce(r[1], r[2])
        if work is not None or back is not None:
            self.pending.append((work, back if not self.cuda else None))
        self.launched += 1

    def finish(self):
        """Make the compute stream wait for every launched bucket (no host sync on CUDA)."""
        for work, back in self.pending:
            if work is not None:
                work.wait()
            if back is not None:
                back[0].copy_(back[1])
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
