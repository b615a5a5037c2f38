"""Host-side mirror of the reference crop sampling (pytracking/features/preprocessing.py:6-7, 33-148) for the
border mode the trackers of this path use ('replicate').  It runs the same torch-CPU operations in the same order as the
reference, so the crops are bit-identical (asserted against the reference itself in oracle/gen_track_golden.py); it exists
so that bench / tests on a machine without the reference tree can feed the engine exactly what the tracker would.
"""
import math

import torch
import torch.nn.functional as F


def numpy_to_torch(a):
    """preprocessing.py:6-7: HxWx3 uint8/float ndarray -> [1,3,H,W] float32."""
    return torch.from_numpy(a).float().permute(2, 0, 1).unsqueeze(0)


def sample_patch(im, pos, sample_sz, output_sz=None, mode="replicate"):
    """preprocessing.py:55-148 (mode 'replicate'). Returns (patch [1,C,h,w], patch_coord [1,4] = (tl_y, tl_x, br_y, br_x))."""
    if mode != "replicate":
        raise NotImplementedError("sample_patch mirror: only border_mode 'replicate'")
    posl = pos.long().clone()
    if output_sz is not None:
        resize_factor = torch.min(sample_sz.float() / output_sz.float()).item()
        df = int(max(int(resize_factor - 0.1), 1))
    else:
        df = int(1)
    sz = sample_sz.float() / df
    if df > 1:
        os_ = posl % df
        posl = (posl - os_) / df
        im2 = im[..., os_[0].item()::df, os_[1].item()::df]
    else:
        im2 = im
    szl = torch.max(sz.round(), torch.Tensor([2])).long()
    tl = posl - (szl - 1) / 2
    br = posl + szl / 2 + 1
    pad = (-tl[1].int().item(), br[1].int().item() - im2.shape[3], -tl[0].int().item(), br[0].int().item() - im2.shape[2])
    im_patch = F.pad(im2, pad, "replicate")
    patch_coord = df * torch.cat((tl, br)).view(1, 4)
    if output_sz is None or (im_patch.shape[-2] == output_sz[0] and im_patch.shape[-1] == output_sz[1]):
        return im_patch.clone(), patch_coord
    im_patch = F.interpolate(im_patch, output_sz.long().tolist(), mode="bilinear")
    return im_patch, patch_coord


def sample_patch_multiscale(im, pos, scales, image_sz, mode="replicate", max_scale_change=None):
    """preprocessing.py:33-52."""
    if isinstance(scales, (int, float)):
        scales = [scales]
    patch_iter, coord_iter = zip(*(sample_patch(im, pos, s * image_sz, image_sz, mode=mode) for s in scales))
    return torch.cat(list(patch_iter)), torch.cat(list(coord_iter))


def sample_init_patch(im, pos, scale, img_sample_sz, aug_expansion_factor=None):
    """The un-augmented first-frame sample of DiMP.generate_init_samples (pytracking/tracker/dimp/dimp.py:353-389):
    the patch is sampled at the augmentation expansion size and the Identity transform crops its centre back to the sample
    size (pytracking/features/augmentation.py:20-40: F.pad with negative 'replicate' padding)."""
    import math
    aug_sz = img_sample_sz.clone()
    out_sz = None
    if aug_expansion_factor is not None and aug_expansion_factor != 1:
        aug_sz = (img_sample_sz * aug_expansion_factor).long()
        aug_sz += (aug_sz - img_sample_sz.long()) % 2
        aug_sz = aug_sz.float()
        out_sz = img_sample_sz.long().tolist()
    patch, _ = sample_patch(im, pos, scale * aug_sz, aug_sz)
    if out_sz is None:
        return patch
    pad_h = (out_sz[0] - patch.shape[2]) / 2
    pad_w = (out_sz[1] - patch.shape[3]) / 2
    return F.pad(patch, (math.floor(pad_w), math.ceil(pad_w), math.floor(pad_h), math.ceil(pad_h)), "replicate")


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
