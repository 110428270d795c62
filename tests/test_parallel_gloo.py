"""Multi-rank sharding logic on CPU (gloo, world_size 2 and 3): sharded-by-window grounding of one long clip with the single
packed all-gather must reproduce the single-process result exactly, in temporal order."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from grove_b200 import parallel
from oracle import box_eval


def _fake_window_fn(P):
    # a deterministic stand-in for the per-window hot path: the record of frame t, phrase p is a function of (t, p)
    def fn(frame_ids, out=None):
        t = torch.tensor(frame_ids, dtype=torch.float32)[:, None, None]
        p = torch.arange(P, dtype=torch.float32)[None, :, None]
        k = torch.arange(5, dtype=torch.float32)[None, None, :]
        return torch.sin(0.37 * t + 1.3 * p + 0.11 * k)
    return fn


def _worker(rank, world, port, num_frames, P, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        out = parallel.ground_sharded_clip(_fake_window_fn(P), num_frames, P)
        ret[rank] = out
    finally:
        dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("world,num_frames", [(2, 128), (2, 40), (3, 64), (2, 50), (3, 61)])
def test_sharded_clip_allgather_matches_single_process(world, num_frames):
    P = 16 if num_frames == 128 else 3
    ref = parallel.ground_sharded_clip(_fake_window_fn(P), num_frames, P)       # world size 1 path
    full = _fake_window_fn(P)(list(range(num_frames)))
    assert torch.equal(ref, full)                                               # every frame produced once, in temporal order
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), num_frames, P, ret), nprocs=world, join=True)
    for r in range(world):
        assert torch.equal(ret[r], ref), f"rank {r} differs"


def test_window_schedule_matches_reference_port():
    for n in (8, 48, 50, 61, 64, 128):
        assert parallel.sliding_segment_with_mask(n, 8) == box_eval.sliding_segment_with_mask(n, 8)
    assert parallel.units_of_rank(16, 8, 3) == [3, 11]
    assert sorted(sum((parallel.units_of_rank(5, 2, r) for r in range(2)), [])) == list(range(5))


def _grad_worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(100 + rank)
        grads = [torch.randn(s, generator=g) for s in ((300, 7), (5,), (1,), (64, 64), (1000,))]
        parallel.allreduce_gradients(grads, bucket_elems=2200)          # forces three buckets, one of them a single tensor
        ret[rank] = grads
    finally:
        dist.destroy_process_group()


def _reducer_worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(200 + rank)
        groups = [[torch.randn(40, 3, generator=g), torch.randn(7, generator=g)], [torch.randn(500, generator=g)], [torch.randn(1, generator=g)]]
        red = parallel.GradientReducer(bucket_elems=100)
        for grp in groups:                 # handed over group by group, as the backward pass produces them
            red.ready(grp)
        red.finish()
        ret[rank] = (groups, red.calls, red.reduced_elems)
    finally:
        dist.destroy_process_group()


def test_gradient_reducer_groups():
    """config 4: the overlapped reducer (here on CPU tensors: synchronous) averages every group it is handed"""
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_reducer_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
    per_rank = []
    for r in range(world):
        g = torch.Generator().manual_seed(200 + r)
        per_rank.append([[torch.randn(40, 3, generator=g), torch.randn(7, generator=g)], [torch.randn(500, generator=g)], [torch.randn(1, generator=g)]])
    for r in range(world):
        groups, calls, elems = ret[r]
        assert calls == 3 and elems == 120 + 7 + 500 + 1
        for gi, grp in enumerate(groups):
            for ti, t in enumerate(grp):
                assert torch.allclose(t, (per_rank[0][gi][ti] + per_rank[1][gi][ti]) / 2, atol=1e-6)
    single = parallel.GradientReducer()    # no process group: a no-op
    x = torch.ones(3)
    single.ready([x]); single.finish()
    assert torch.equal(x, torch.ones(3)) and single.calls == 0


def test_gradient_allreduce_buckets():
    """config 4's data-parallel gradient averaging: bucketed all-reduce equals the mean of the per-rank gradients"""
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_grad_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
    per_rank = []
    for r in range(world):
        g = torch.Generator().manual_seed(100 + r)
        per_rank.append([torch.randn(s, generator=g) for s in ((300, 7), (5,), (1,), (64, 64), (1000,))])
    for i in range(5):
        mean = (per_rank[0][i] + per_rank[1][i]) / 2
        for r in range(world):
            assert torch.allclose(ret[r][i], mean, atol=1e-6), (r, i)
