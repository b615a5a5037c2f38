"""CPU, world_size 2 over gloo: the N>1 path of the multi-GPU run (sequence ownership + the single final gather)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from pytracking_b200 import shard


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, num_seq, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        mine = shard.assign_sequences(num_seq, world, rank)
        results = {}
        for q in mine:
            t = 5 + q                                     # ragged lengths
            boxes = torch.arange(t * 4, dtype=torch.float32).reshape(t, 4) + 100 * q
            times = torch.full((t,), 0.001 * (q + 1))
            results[q] = (boxes, times)
        merged = shard.gather_results(results, num_seq, max_frames=16)
        ok = sorted(merged) == list(range(num_seq))
        for q, (boxes, times) in merged.items():
            t = 5 + q
            ok = ok and boxes.shape == (t, 4) and torch.equal(boxes, torch.arange(t * 4, dtype=torch.float32).reshape(t, 4) + 100 * q)
            ok = ok and torch.allclose(times, torch.full((t,), 0.001 * (q + 1)))
        fps = shard.aggregate_fps(merged)
        ret[rank] = (ok, mine, fps)
    finally:
        dist.destroy_process_group()


def test_assign_sequences():
    assert shard.assign_sequences(8, 8, 3) == [3]
    assert shard.assign_sequences(5, 2, 0) == [0, 2, 4] and shard.assign_sequences(5, 2, 1) == [1, 3]
    assert shard.assign_sequences(1, 4, 2) == []
    owners = sorted(q for r in range(3) for q in shard.assign_sequences(10, 3, r))
    assert owners == list(range(10))
    with pytest.raises(ValueError):
        shard.assign_sequences(4, 2, 2)


def test_single_process_gather():
    res = {0: (torch.ones(3, 4), torch.ones(3)), 1: (torch.zeros(2, 4), torch.ones(2))}
    merged = shard.gather_results(res, 2, max_frames=4)
    assert merged[0][0].shape == (3, 4) and merged[1][0].shape == (2, 4)
    assert abs(shard.aggregate_fps(merged) - 5 / 3.0) < 1e-6
    with pytest.raises(RuntimeError):
        shard.gather_results({0: res[0]}, 2, max_frames=4)          # sequence 1 was never tracked
    with pytest.raises(ValueError):
        shard.pack_results({0: (torch.ones(9, 4), torch.ones(9))}, max_frames=4)


@pytest.mark.timeout(120)
def test_two_rank_gather_gloo():
    world, num_seq = 2, 5
    mgr = mp.get_context("spawn").Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), num_seq, ret), nprocs=world, join=True)
    assert len(ret) == world
    for rank in range(world):
        ok, mine, fps = ret[rank]
        assert ok and mine == shard.assign_sequences(num_seq, world, rank)
    assert abs(ret[0][2] - ret[1][2]) < 1e-9
