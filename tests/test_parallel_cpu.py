"""CPU (gloo, world_size 2): host logic of the data-parallel path -- bucket partition of the flat
gradient buffer, overlapped all-reduce launch/finish protocol, DistributedSampler-compatible sharding,
and the fused meter synchronisation."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mem_b200 import registry, utils
from mem_b200 import modeling_pretrain  # noqa: F401
from mem_b200.parallel import GradReducer, bucket_ranges, shard_indices
from mem_b200.vit_engine import engine_of
from oracle import vit_ref


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_bucket_ranges_cover_flat_buffer_in_backward_order():
    model = registry.create_model("pt_vit", **dict(vit_ref.TINY, depth=4))
    flat = engine_of(model).flat()
    r = bucket_ranges(flat, 4, min_bucket_elems=1)
    assert [t for t, _, _ in r] == ["head", 3, 2, 1, 0, "embed"]
    assert r[0][1] == 0 and r[-1][2] == flat.numel
    for (_, a, b), (_, c, d) in zip(r, r[1:]):
        assert b == c and a < b
    # every parameter of block i lies inside the bucket tagged i; the shared rel-pos table is in "embed"
    by_tag = {t: (a, b) for t, a, b in r}
    for n in flat.names:
        o = flat.offsets[n]
        if n.startswith("blocks."):
            a, b = by_tag[int(n.split(".")[1])]
        elif n.startswith(("lm_head", "norm")):
            a, b = by_tag["head"]
        else:
            a, b = by_tag["embed"]
        assert a <= o < b, n
    merged = bucket_ranges(flat, 4, min_bucket_elems=10 ** 9)
    assert len(merged) == 1 and merged[0] == ("embed", 0, flat.numel)
    # merging stops before the end of backward: the last two blocks fire alone and "embed" carries only the embedding
    blk = by_tag[3][1] - by_tag[3][0]
    tail = bucket_ranges(flat, 4, min_bucket_elems=3 * blk)
    assert [t for t, _, _ in tail] == [2, 1, 0, "embed"]
    assert tail[0][1] == 0 and tail[-1][1:] == by_tag["embed"] and tail[1][1:] == by_tag[1] and tail[2][1:] == by_tag[0]


def test_shard_indices_matches_distributed_sampler():
    from torch.utils.data import DistributedSampler
    data = list(range(103))
    for epoch in (0, 3):
        for world in (2, 8):
            for rank in range(world):
                s = DistributedSampler(data, num_replicas=world, rank=rank, shuffle=True, seed=5)
                s.set_epoch(epoch)
                assert list(s) == shard_indices(len(data), rank, world, epoch=epoch, seed=5)
    s = DistributedSampler(data, num_replicas=4, rank=1, shuffle=False, drop_last=True)
    assert list(s) == shard_indices(len(data), 1, 4, shuffle=False, drop_last=True)


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(0)
        model = registry.create_model("pt_vit", **vit_ref.TINY)
        flat = engine_of(model).flat()
        g = torch.Generator().manual_seed(100 + rank)
        flat.grad.copy_(torch.randn(flat.numel, generator=g))
        mine = flat.grad.clone()
        red = GradReducer(flat.grad, bucket_ranges(flat, 2, min_bucket_elems=1), wire_dtype=torch.float32)
        for tag in ("head", 1, 0, "embed"):       # the order VitEngine.backward_pretrain fires them
            red.hook(tag)
        red.finish()
        other = torch.randn(flat.numel, generator=torch.Generator().manual_seed(100 + (1 - rank)))
        ok = torch.allclose(flat.grad, mine + other, atol=1e-6) and red.launched == 4
        # bf16 wire (the default): each rank's slice is rounded to bf16, summed in bf16, written back as fp32
        flat.grad.copy_(mine)
        red16 = GradReducer(flat.grad, bucket_ranges(flat, 2, min_bucket_elems=1))
        assert red16.wire_dtype == torch.bfloat16
        for tag in ("head", 1, 0, "embed"):
            red16.hook(tag)
        red16.finish()
        want16 = (mine.bfloat16() + other.bfloat16()).float()
        ok = ok and torch.equal(flat.grad, want16) and red16.launched == 4
        # fused meter all-reduce
        ml = utils.MetricLogger()
        ml.update(loss=1.0 + rank, mlm_acc=0.5 * rank)
        ml.update(loss=3.0 + rank)
        ml.synchronize_between_processes()
        ok = ok and abs(ml.loss.global_avg - (1 + 3 + 2 + 4) / 4) < 1e-12 and ml.loss.count == 4
        ok = ok and abs(ml.mlm_acc.global_avg - 0.25) < 1e-12
        out[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


def test_grad_reducer_and_meters_world2_gloo():
    port = _free_port()
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
        assert dict(out) == {0: True, 1: True}


def _sync_worker(rank, world, port, out):
    """Replicas built from rank-dependent seeds (the reference seeds with args.seed + rank and lets the
    DistributedDataParallel constructor broadcast rank 0's weights, run_mem_pretraining.py:255, :365-367)."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from mem_b200 import engine_for_finetuning, engine_for_pretraining
        torch.manual_seed(1234 + rank)
        model = registry.create_model("pt_vit", **vit_ref.TINY)
        before = engine_of(model).flat().data.clone()
        red = engine_for_pretraining._reducer_for(model)          # builds the reducer -> broadcasts rank 0's weights
        after = engine_of(model).flat().data
        gathered = [torch.empty_like(after) for _ in range(world)]
        dist.all_gather(gathered, after)
        same = all(torch.equal(gathered[0], g) for g in gathered)
        moved = (rank == 0 and torch.equal(before, after)) or (rank != 0 and not torch.equal(before, after))
        # parameters seen through the module are the broadcast ones (views of the flat buffer)
        views = torch.equal(dict(model.named_parameters())["lm_head.weight"].detach().flatten(),
                            after[engine_of(model).flat().offsets["lm_head.weight"]:][:model.lm_head.weight.numel()])
        # the finetuning loop's one-time sync
        torch.manual_seed(99 + rank)
        ft = registry.create_model("ft_vit", **vit_ref.TINY_FT)
        engine_for_finetuning._sync_replicas(ft)
        f = engine_of(ft).flat().data
        g2 = [torch.empty_like(f) for _ in range(world)]
        dist.all_gather(g2, f)
        out[rank] = bool(same and moved and views and red is not None and torch.equal(g2[0], g2[1]))
    finally:
        dist.destroy_process_group()


def test_replicas_are_broadcast_from_rank0_world2_gloo():
    port = _free_port()
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_sync_worker, args=(2, port, out), nprocs=2, join=True)
        assert dict(out) == {0: True, 1: True}
