"""Sequence sharding for the multi-GPU run.

The reference runs independent (sequence, tracker) pairs over a process pool with no communication
(pytracking/evaluation/running.py:189-219).  Here: one process per GPU, rank r owns the sequences
{q : q mod world == r}, no collective inside the frame loop, and ONE all_gather of fixed-size padded per-sequence
result tensors at the very end (boxes [T,4], per-frame time [T], valid length) -- NCCL on the GPU box, gloo in the CPU tests.
"""
import torch
import torch.distributed as dist


def assign_sequences(num_sequences, world_size, rank):
    """Round-robin ownership: rank r tracks sequences r, r + world, r + 2*world, ..."""
    if not (0 <= rank < world_size):
        raise ValueError("rank %d outside world of %d" % (rank, world_size))
    return list(range(rank, num_sequences, world_size))


def pack_results(results, max_frames):
    """results: {seq_id: (boxes [T,4] float tensor, times [T] float tensor)} of this rank -> one padded tensor
    [n_local, max_frames, 6] = (x, y, w, h, time, valid) plus the sequence ids [n_local]."""
    ids = sorted(results)
    out = torch.zeros(len(ids), max_frames, 6, dtype=torch.float32)
    for i, q in enumerate(ids):
        boxes, times = results[q]
        t = boxes.shape[0]
        if t > max_frames:
            raise ValueError("sequence %d has %d frames > max_frames=%d" % (q, t, max_frames))
        out[i, :t, :4] = boxes
        out[i, :t, 4] = times
        out[i, :t, 5] = 1.0
    return torch.tensor(ids, dtype=torch.int64), out


def gather_results(results, num_sequences, max_frames, device=None, group=None):
    """The single collective of the run. Every rank returns {seq_id: (boxes [T,4], times [T])} for ALL sequences."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    per_rank = (num_sequences + world - 1) // world
    ids, packed = pack_results(results, max_frames)
    ids_pad = torch.full((per_rank,), -1, dtype=torch.int64)
    ids_pad[:ids.numel()] = ids
    buf = torch.zeros(per_rank, max_frames, 6, dtype=torch.float32)
    buf[:packed.shape[0]] = packed
    if device is not None:
        ids_pad, buf = ids_pad.to(device), buf.to(device)
    if world > 1:
        all_ids = [torch.empty_like(ids_pad) for _ in range(world)]
        all_buf = [torch.empty_like(buf) for _ in range(world)]
        dist.all_gather(all_ids, ids_pad, group=group)
        dist.all_gather(all_buf, buf, group=group)
    else:
        all_ids, all_buf = [ids_pad], [buf]
    merged = {}
    for r in range(world):
        for i, q in enumerate(all_ids[r].cpu().tolist()):
            if q < 0:
                continue
            if q in merged:
                raise RuntimeError("sequence %d reported by two ranks" % q)
            row = all_buf[r][i].cpu()
            t = int(row[:, 5].sum().item())
            merged[q] = (row[:t, :4].clone(), row[:t, 4].clone())
    missing = [q for q in range(num_sequences) if q not in merged]
    if missing:
        raise RuntimeError("sequences %s were not tracked by any rank (rank %d of %d)" % (missing, rank, world))
    return merged


def aggregate_fps(merged):
    """Whole-job throughput as the reference computes it per sequence (running.py:173-180), summed over sequences
    tracked concurrently: total frames / max over sequences of the summed per-frame time."""
    total = sum(v[0].shape[0] for v in merged.values())
    longest = max(float(v[1].sum()) for v in merged.values())
    return total / max(longest, 1e-12)


def iou_overlap(pred_bb, anno_bb):
    """Per-frame IoU of (x, y, w, h) boxes with the inclusive-pixel convention of `calc_iou_overlap`
    (pytracking/analysis/extract_results.py:29-39): a box covers the pixels x .. x + w - 1."""
    lo = torch.maximum(pred_bb[:, :2], anno_bb[:, :2])
    hi = torch.minimum(pred_bb[:, :2] + pred_bb[:, 2:], anno_bb[:, :2] + anno_bb[:, 2:]) - 1.0
    inter = (hi - lo + 1.0).clamp(min=0).prod(dim=1)
    return inter / (pred_bb[:, 2:].prod(dim=1) + anno_bb[:, 2:].prod(dim=1) - inter)
