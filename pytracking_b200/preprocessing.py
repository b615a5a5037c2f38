"""Per-sequence host constants of the hot path with bit-exact mirrors of the reference's constructors: the score-map output windows
(pytracking/libs/dcf.py:8-37) and the ToMP sine position encoding (ltr/models/transformer/position_encoding.py:33-58).  (Crop
sampling is not here: the engine samples crops on the GPU, csrc/dimp_tracker.cu; the torch-CPU restatement used by the tests lives
in oracle/preprocessing_ref.py.)
"""
import math

import torch


# ---- output windows (host side, built once per sequence; pytracking/libs/dcf.py:8-37) ------------------------------------------
def hann1d(sz, centered=True):
    """1-D cosine window of `sz` taps: centred (peak in the middle, zero just outside both ends) or with its peak at tap 0 and the
    taps wrapped around (the layout of an un-shifted score map).  Bit-exact mirror of dcf.hann1d (float32 arithmetic)."""
    sz = int(sz)
    if centered:
        i = torch.arange(sz, dtype=torch.float32)
        return 0.5 * (1 - torch.cos((2 * math.pi / (sz + 1)) * (i + 1)))
    half = 0.5 * (1 + torch.cos((2 * math.pi / (sz + 2)) * torch.arange(0, sz // 2 + 1).float()))
    idx = torch.tensor([i if i <= sz // 2 else sz - i for i in range(sz)], dtype=torch.long)
    return half[idx]


def hann2d(sz, centered=True):
    """[1,1,H,W] separable cosine window (dcf.hann2d); sz = (H, W)."""
    return hann1d(int(sz[0]), centered).reshape(1, 1, -1, 1) * hann1d(int(sz[1]), centered).reshape(1, 1, 1, -1)


def hann2d_clipped(sz, effective_sz, centered=True):
    """dcf.hann2d_clipped: a centred cosine window of `effective_sz` (made to differ from `sz` by an even amount), extended to `sz`
    by replicating its border, optionally rotated so that its peak sits at index (0, 0)."""
    sz = [int(sz[0]), int(sz[1])]
    eff = [int(effective_sz[0]), int(effective_sz[1])]
    eff = [e + (e - s) % 2 for e, s in zip(eff, sz)]
    win = hann2d(eff, True)
    rows = (torch.arange(sz[0]) - (sz[0] - eff[0]) // 2).clamp(0, eff[0] - 1)
    cols = (torch.arange(sz[1]) - (sz[1] - eff[1]) // 2).clamp(0, eff[1] - 1)
    win = win[:, :, rows][:, :, :, cols]
    if centered:
        return win
    return torch.roll(win, shifts=(-int(sz[0] / 2), -int(sz[1] / 2)), dims=(2, 3))


# ---- ToMP position encoding (host side, constant per feature size; ltr/models/transformer/position_encoding.py:6-58) -----------
def tomp_position_encoding(h, w, d_model=256, max_spatial_resolution=18):
    """[d_model, h, w] float32: PositionEmbeddingSine(num_pos_feats=d_model//2, sine_type='lin_sine', avoid_aliazing=True,
    max_spatial_resolution=...) evaluated on an all-valid mask, as FilterPredictor.get_positional_encoding does
    (ltr/models/transformer/filter_predictor.py:34-35,41-47).  Pixel centres are normalised to (0,1) per axis, then every
    coordinate pair (x, y) is expanded into sin / cos of `depth = d_model/4` linearly spaced frequencies k * (res/depth) * pi:
    channel layout = [sin k=1 (x,y), ..., sin k=depth (x,y), cos k=1 (x,y), ..., cos k=depth (x,y)].  Bit-exact mirror."""
    depth = (d_model // 2) // 2
    factor = float(max_spatial_resolution) / depth
    ys = torch.arange(1, h + 1, dtype=torch.float32).reshape(h, 1).expand(h, w)
    xs = torch.arange(1, w + 1, dtype=torch.float32).reshape(1, w).expand(h, w)
    ys = (ys - 0.5) / (ys[-1:, :] + 1e-6)
    xs = (xs - 0.5) / (xs[:, -1:] + 1e-6)
    pos = torch.stack([xs, ys], dim=-1)                                           # [h, w, 2]
    waves = [torch.sin((k + 1) * factor * math.pi * pos) for k in range(depth)] + \
            [torch.cos((k + 1) * factor * math.pi * pos) for k in range(depth)]
    return torch.cat(waves, dim=-1).permute(2, 0, 1).contiguous()
