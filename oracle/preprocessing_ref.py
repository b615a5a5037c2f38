"""TEST INFRASTRUCTURE ONLY (never imported by pytracking_b200/): torch-CPU restatement of the reference's crop sampling
(pytracking/features/preprocessing.py:55-148, border mode 'replicate') split the way the engine splits it -- an integer / float32
*geometry* step (`sample_patch_geometry`, what csrc/dimp_tracker.cu `plan_patch` computes on the host) and a *resampling* step
(clamped gather + torch's own CPU bilinear `F.interpolate`, what `sample_patch_kernel` computes on the device).
Pinned: oracle/gen_track_golden.py and gen_atom_track_golden.py assert bit-equality with the reference's `sample_patch` /
`sample_patch_transformed` on every recorded frame.  The GPU tests use it to check the crop kernel at arbitrary geometries.
"""
import math

import torch
import torch.nn.functional as F


def numpy_to_torch(a):
    """HxWx3 ndarray -> [1,3,H,W] float32 (preprocessing.py:6-7)."""
    return torch.from_numpy(a).float().permute(2, 0, 1).unsqueeze(0)


def sample_patch_geometry(im, pos, sample_sz, output_sz):
    """-> (df, os_r, os_c, tl_r, tl_c, in_h, in_w, patch_coord[1,4] float32).  float32 tensor arithmetic exactly as the reference
    performs it (preprocessing.py:75-140): truncating `.long()`, true division of integer tensors, round-half-even."""
    centre = pos.long()
    df = 1
    if output_sz is not None:
        df = max(int(torch.min(sample_sz.float() / output_sz.float()).item() - 0.1), 1)
    offs = centre % df if df > 1 else torch.zeros(2, dtype=torch.long)
    centre = (centre - offs) / df if df > 1 else centre
    extent = torch.max((sample_sz.float() / df).round(), torch.Tensor([2])).long()
    lo = centre - (extent - 1) / 2
    hi = centre + extent / 2 + 1
    lo_i, hi_i = lo.int(), hi.int()
    coord = df * torch.cat((lo, hi)).view(1, 4)
    return (df, int(offs[0]), int(offs[1]), int(lo_i[0]), int(lo_i[1]), int(hi_i[0] - lo_i[0]), int(hi_i[1] - lo_i[1]), coord)


def sample_patch(im, pos, sample_sz, output_sz=None, mode="replicate"):
    """-> (patch [1,C,h,w], patch_coord [1,4]); replicate padding is a clamped gather of the decimated image."""
    if mode != "replicate":
        raise NotImplementedError("only border_mode 'replicate'")
    df, os_r, os_c, tl_r, tl_c, in_h, in_w, coord = sample_patch_geometry(im, pos, sample_sz, output_sz)
    dec = im[..., os_r::df, os_c::df]
    rows = (torch.arange(in_h) + tl_r).clamp(0, dec.shape[-2] - 1)
    cols = (torch.arange(in_w) + tl_c).clamp(0, dec.shape[-1] - 1)
    patch = dec[..., rows, :][..., cols]
    if output_sz is None or (patch.shape[-2] == output_sz[0] and patch.shape[-1] == output_sz[1]):
        return patch.clone(), coord
    return F.interpolate(patch, output_sz.long().tolist(), mode="bilinear"), coord


def sample_patch_multiscale(im, pos, scales, image_sz, mode="replicate", max_scale_change=None):
    """preprocessing.py:33-52."""
    scales = [scales] if isinstance(scales, (int, float)) else scales
    out = [sample_patch(im, pos, s * image_sz, image_sz, mode=mode) for s in scales]
    return torch.cat([o[0] for o in out]), torch.cat([o[1] for o in out])


def sample_init_patch(im, pos, scale, img_sample_sz, aug_expansion_factor=None):
    """The un-augmented first-frame sample of DiMP.generate_init_samples (pytracking/tracker/dimp/dimp.py:353-389): sampled at the
    augmentation expansion size, then the Identity transform keeps the centre window (augmentation.py:20-40)."""
    big = img_sample_sz.clone()
    if aug_expansion_factor is not None and aug_expansion_factor != 1:
        big = (img_sample_sz * aug_expansion_factor).long()
        big += (big - img_sample_sz.long()) % 2
        big = big.float()
    patch, _ = sample_patch(im, pos, scale * big, big)
    h, w = [int(v) for v in img_sample_sz.long()]
    top, left = -math.floor((h - patch.shape[2]) / 2), -math.floor((w - patch.shape[3]) / 2)
    return patch[..., top:top + h, left:left + w].clone()
