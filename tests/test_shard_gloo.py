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


def test_iou_overlap_matches_reference_calc_iou_overlap():
    """shard.iou_overlap against the reference's calc_iou_overlap (pytracking/analysis/extract_results.py:29-39) on random boxes,
    and the bench's synthetic ground truth against the sequence generator."""
    import numpy as np
    import torch
    from pytracking_b200 import shard, synth
    g = torch.Generator().manual_seed(0)
    a = torch.cat([100 * torch.rand(64, 2, generator=g), 5 + 60 * torch.rand(64, 2, generator=g)], 1)
    b = torch.cat([100 * torch.rand(64, 2, generator=g), 5 + 60 * torch.rand(64, 2, generator=g)], 1)
    b[:8] = a[:8]
    mine = shard.iou_overlap(a, b)
    assert torch.allclose(mine[:8], torch.ones(8))
    from baseline import ref_env
    if ref_env.reference_available():
        ref_env.install()
        from pytracking.analysis.extract_results import calc_iou_overlap
        assert torch.equal(mine, calc_iou_overlap(a, b))
    frames, bb = synth.make_sequence(3, num_frames=5)
    gt = synth.sequence_ground_truth(3, 5)
    assert gt[0] == bb and len(gt) == len(frames)
    x, y, w, h = [int(v) for v in gt[4]]
    assert frames[4][y:y + h, x:x + w].mean() > 100 and frames[4][:y - 2].mean() < 40
