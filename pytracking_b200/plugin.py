"""Host-side mirror of the reference's plug-in seams for the hot path -- same names, argument meaning,
return contracts and shape conventions as the reference functions / modules they stand in for, computing through the
C-ABI library.  `install()` rebinds the seams inside an importable, unmodified pytracking checkout.

Shapes the CUDA path does not claim (several sequences per call, filter sizes other than 4x4, dilation, groups,
training mode) raise `NotImplementedError` here; `install()` keeps the reference implementation for exactly those cases.
"""
import math

import torch

from . import ops


# ---------------------------------------------------------------------------------------------------------------
# ltr/models/layers/filter.py
# ---------------------------------------------------------------------------------------------------------------
def apply_filter(feat, filter, dilation_factors=None):
    """filter.py:5-57. feat (images_in_sequence, [sequences], feat_dim, H, W); filter (sequences, feat_dim, fH, fW)
    -> scores (images_in_sequence, [sequences], yH, yW)."""
    multiple_filters = (filter.dim() == 5)
    if dilation_factors is not None or multiple_filters:
        raise NotImplementedError("b200trk apply_filter: dilation / multiple filters are not on the CUDA path")
    num_images = feat.shape[0]
    num_sequences = feat.shape[1] if feat.dim() == 5 else 1
    if num_sequences != 1 or filter.shape[0] != 1:
        raise NotImplementedError("b200trk apply_filter: one sequence per call")
    f = feat.reshape(num_images, *feat.shape[-3:])
    if filter.shape[-1] == 1 and filter.shape[-2] == 1:
        # _apply_filter_ksz1 (filter.py:60-88, the ToMP classifier): a matmul over the channels == a 1x1 convolution
        scores = ops.conv1x1(f, filter.reshape(1, filter.shape[-3], 1, 1))
        return scores.reshape(num_images, num_sequences, scores.shape[-2], scores.shape[-1])
    if filter.shape[-1] != 4 or filter.shape[-2] != 4:
        raise NotImplementedError("b200trk apply_filter: 4x4 or 1x1 filters")
    scores = ops.apply_filter(f, filter.reshape(1, *filter.shape[-3:]))
    return scores.reshape(num_images, num_sequences, scores.shape[-2], scores.shape[-1])


def apply_feat_transpose(feat, input, filter_ksz, training=True, groups=1):
    """filter.py:91-107 (v2 and v3 are the same adjoint). input (images, [sequences], yH, yW) -> (sequences, feat_dim, fH, fW)."""
    if groups != 1:
        raise NotImplementedError("b200trk apply_feat_transpose: groups != 1")
    if isinstance(filter_ksz, int):
        filter_ksz = (filter_ksz, filter_ksz)
    num_images = feat.shape[0]
    num_sequences = feat.shape[1] if feat.dim() == 5 else 1
    if num_sequences != 1 or tuple(filter_ksz) != (4, 4):
        raise NotImplementedError("b200trk apply_feat_transpose: one sequence and a 4x4 filter per call")
    f = feat.reshape(num_images, *feat.shape[-3:])
    r = input.reshape(num_images, 1, input.shape[-2], input.shape[-1])
    return ops.apply_feat_transpose(f, r, 4)


# ---------------------------------------------------------------------------------------------------------------
# pytracking/libs/dcf.py, pytracking/libs/operation.py, pytracking/libs/fourier.py (ATOM)
# ---------------------------------------------------------------------------------------------------------------
def softmax_reg(x, dim, reg=None):
    """ltr/models/layers/activation.py:7-16 (last dimension only)."""
    if dim % x.dim() != x.dim() - 1:
        raise NotImplementedError("b200trk softmax_reg: only the last dimension")
    return ops.softmax_reg(x, reg)


def max2d(a):
    """dcf.py:156-164 -> (max_val, argmax [.., 2] as (row, col))."""
    return ops.max2d(a)


def conv2d(input, weight, bias=None, stride=1, padding=0, dilation=1, groups=1, mode=None):
    """operation.py:5-32 for ATOM.apply_filter: one 4x4 filter, mode='same'."""
    if weight is None:
        return input
    if mode != "same" or bias is not None or stride != 1 or padding != 0 or dilation != 1 or groups != 1 or weight.shape[0] != 1:
        raise NotImplementedError("b200trk conv2d: only mode='same' with a single 4x4 filter is on the CUDA path")
    return ops.conv2d_same(input, weight)


def conv1x1(input, weight):
    """operation.py:35-42."""
    if weight is None:
        return input
    return ops.conv1x1(input, weight)


def localize_scores_fs(scores_raw, kernel_size, output_sz):
    """ATOM.localize_target's Fourier chain (atom.py:304-316) for one feature type: cfft2 / (H*W) -> shift_fs -> sum_fs -> sample_fs."""
    return ops.fourier_interp(scores_raw, kernel_size, output_sz)


# ---------------------------------------------------------------------------------------------------------------
# ltr/models/target_classifier/optimizer.py  (forward(weights, feat, bb, sample_weight, num_iter, compute_losses)
#                                            -> (weights, weight_iterates, losses))
# ---------------------------------------------------------------------------------------------------------------
def _one_sequence(feat, bb, sample_weight):
    num_images = feat.shape[0]
    num_sequences = feat.shape[1] if feat.dim() == 5 else 1
    if num_sequences != 1:
        raise NotImplementedError("b200trk optimiser modules: one sequence per call (the tracker's inference configuration)")
    f = feat.reshape(num_images, *feat.shape[-3:])
    b = None if bb is None else bb.reshape(num_images, 4).float()
    sw = None
    if isinstance(sample_weight, torch.Tensor):
        sw = sample_weight.reshape(num_images).float()
    elif sample_weight is not None:
        raise NotImplementedError("b200trk optimiser modules: sample_weight must be a tensor or None")
    return f, b, sw


def _iterates(its, losses, compute_losses):
    return [w.unsqueeze(0) for w in its], ([l for l in losses] if compute_losses else [])


class DiMPSteepestDescentGN:
    """optimizer.py:11-170. Built from the reference module (or its state_dict entries)."""

    def __init__(self, label_map_weight, target_mask_weight, spatial_weight_weight, log_step_length, filter_reg, num_iter=1,
                 feat_stride=16, min_filter_reg=1e-3, alpha_eps=0.0, bin_displacement=0.1):
        self.luts = [t.detach().float().reshape(-1).contiguous() for t in (label_map_weight, target_mask_weight, spatial_weight_weight)]
        self.step_length = float(torch.exp(log_step_length.detach().float()).item())
        self.reg_weight = max(float(filter_reg.detach().float().item()) ** 2, min_filter_reg ** 2)
        self.num_iter, self.feat_stride, self.alpha_eps, self.bin_displacement = num_iter, feat_stride, alpha_eps, bin_displacement

    @classmethod
    def from_module(cls, m):
        return cls(m.label_map_predictor.weight, m.target_mask_predictor[0].weight, m.spatial_weight_predictor.weight,
                   m.log_step_length, m.filter_reg, m.num_iter, m.feat_stride, m.min_filter_reg, m.alpha_eps,
                   m.distance_map.bin_displacement)

    def forward(self, weights, feat, bb, sample_weight=None, num_iter=None, compute_losses=True):
        num_iter = self.num_iter if num_iter is None else num_iter
        f, b, sw = _one_sequence(feat, bb, sample_weight)
        luts = [t.to(f.device) for t in self.luts]
        w, its, losses = ops.dimp_sd_gn(weights, f, b, sw, luts[0], luts[1], luts[2], num_iter, self.step_length, self.reg_weight,
                                        self.alpha_eps, self.bin_displacement, self.feat_stride, return_iterates=True,
                                        compute_losses=compute_losses)
        return (w,) + _iterates(its, losses, compute_losses)

    __call__ = forward


class PrDiMPSteepestDescentNewton:
    """optimizer.py:294-439."""

    def __init__(self, log_step_length, filter_reg, num_iter=1, feat_stride=16, gauss_sigma=1.0, min_filter_reg=1e-3, alpha_eps=0.0,
                 softmax_reg=None, label_shrink=0.0, label_threshold=0.0, normalize_label=False, uni_weight=0.0):
        self.step_length = float(torch.exp(log_step_length.detach().float()).item())
        self.reg_weight = max(float(filter_reg.detach().float().item()) ** 2, min_filter_reg ** 2)
        self.num_iter, self.feat_stride, self.gauss_sigma, self.alpha_eps = num_iter, feat_stride, gauss_sigma, alpha_eps
        self.softmax_reg, self.label_shrink, self.label_threshold = softmax_reg, label_shrink, label_threshold
        self.normalize_label, self.uni_weight = normalize_label, uni_weight

    @classmethod
    def from_module(cls, m):
        return cls(m.log_step_length, m.filter_reg, m.num_iter, m.feat_stride, m.gauss_sigma, m.min_filter_reg, m.alpha_eps,
                   m.softmax_reg, m.label_shrink, m.label_threshold, getattr(m, "normalize_label", False), m.uni_weight)

    def forward(self, weights, feat, bb, sample_weight=None, num_iter=None, compute_losses=True):
        num_iter = self.num_iter if num_iter is None else num_iter
        f, b, sw = _one_sequence(feat, bb, sample_weight)
        w, its, losses = ops.prdimp_sd_newton(weights, f, b, sw, num_iter, self.gauss_sigma, self.step_length, self.reg_weight,
                                              self.alpha_eps, self.softmax_reg, self.label_threshold, self.normalize_label,
                                              self.label_shrink, self.uni_weight, self.feat_stride, return_iterates=True,
                                              compute_losses=compute_losses)
        return (w,) + _iterates(its, losses, compute_losses)

    __call__ = forward


class DiMPL2SteepestDescentGN:
    """optimizer.py:172-291."""

    def __init__(self, log_step_length, filter_reg, num_iter=1, feat_stride=16, gauss_sigma=1.0, hinge_threshold=-999,
                 min_filter_reg=1e-3, alpha_eps=0.0):
        self.step_length = float(torch.exp(log_step_length.detach().float()).item())
        self.reg_weight = max(float(filter_reg.detach().float().item()) ** 2, min_filter_reg ** 2)
        self.num_iter, self.feat_stride, self.gauss_sigma = num_iter, feat_stride, gauss_sigma
        self.hinge_threshold, self.alpha_eps = hinge_threshold, alpha_eps

    def forward(self, weights, feat, bb, sample_weight=None, num_iter=None, compute_losses=True):
        num_iter = self.num_iter if num_iter is None else num_iter
        f, b, sw = _one_sequence(feat, bb, sample_weight)
        w, its, losses = ops.dimp_l2_sd_gn(weights, f, b, sw, num_iter, self.gauss_sigma, self.hinge_threshold, self.step_length,
                                           self.reg_weight, self.alpha_eps, self.feat_stride, return_iterates=True,
                                           compute_losses=compute_losses)
        return (w,) + _iterates(its, losses, compute_losses)

    __call__ = forward


# ---------------------------------------------------------------------------------------------------------------
# pytracking/libs/optimization.py on the ATOM problems (run(...) updates the variable in place)
# ---------------------------------------------------------------------------------------------------------------
class ConjugateGradient:
    """optimization.py:199-289 specialised to ConvProblem (atom/optim.py:71-99): `run(num_cg_iter)` updates `filter` in place."""

    def __init__(self, training_samples, y, filter_reg, sample_weights, filter, response_activation=("mlu", 0.05),
                 fletcher_reeves=True, direction_forget_factor=0):
        if direction_forget_factor != 0:
            raise NotImplementedError("b200trk ConjugateGradient: direction_forget_factor must be 0 (the ATOM default)")
        self.x, self.y, self.reg, self.sw, self.filter = training_samples, y, float(filter_reg), sample_weights, filter
        self.act, self.act_param = response_activation if isinstance(response_activation, tuple) else (response_activation, 0.0)
        self.fletcher_reeves = fletcher_reeves

    def run(self, num_cg_iter):
        if num_cg_iter == 0:
            return
        ops.atom_cg_filter(self.filter, self.x, self.y, self.sw, self.reg, num_cg_iter, self.act, self.act_param, self.fletcher_reeves,
                           out=self.filter)


class GaussNewtonCG:
    """optimization.py:293-435 specialised to FactorizedConvProblem (atom/optim.py:6-68): `run(num_cg_iter, num_gn_iter)`
    updates the filter and the projection matrix in place."""

    def __init__(self, training_samples, y, filter_reg, projection_reg, sample_weights, filter, projection_matrix,
                 response_activation=("mlu", 0.05), fletcher_reeves=True):
        self.x, self.y, self.sw, self.filter, self.P = training_samples, y, sample_weights, filter, projection_matrix
        self.reg, self.preg = float(filter_reg), float(projection_reg)
        self.act, self.act_param = response_activation if isinstance(response_activation, tuple) else (response_activation, 0.0)
        self.fletcher_reeves = fletcher_reeves

    def run(self, num_cg_iter, num_gn_iter=None):
        if isinstance(num_cg_iter, (list, tuple)):
            if len(set(num_cg_iter)) > 1:
                raise NotImplementedError("b200trk GaussNewtonCG: a per-GN-iteration CG schedule must be constant")
            num_cg_iter, num_gn_iter = (num_cg_iter[0] if num_cg_iter else 0), len(num_cg_iter)
        elif num_gn_iter is None:
            raise ValueError("Must specify number of GN iter if CG iter is constant")
        ops.atom_gn_joint_(self.filter, self.P, self.x, self.y, self.sw, self.reg, self.preg, num_cg_iter, num_gn_iter, self.act,
                           self.act_param, self.fletcher_reeves)


# ---------------------------------------------------------------------------------------------------------------
def install(net=None):
    """Rebind the functional seams inside an importable pytracking / ltr checkout (no reference file is edited).
    Unsupported shapes fall through to the reference implementation. Returns the list of rebound attributes."""
    import importlib
    done = []
    fl = importlib.import_module("ltr.models.layers.filter")
    ref_apply, ref_tr = fl.apply_filter, fl.apply_feat_transpose

    def _apply(feat, filter, dilation_factors=None):
        if feat.is_cuda and not torch.is_grad_enabled():
            try:
                return apply_filter(feat, filter, dilation_factors)
            except NotImplementedError:
                pass
        return ref_apply(feat, filter, dilation_factors)

    def _tr(feat, input, filter_ksz, training=True, groups=1):
        if feat.is_cuda and not torch.is_grad_enabled():
            try:
                return apply_feat_transpose(feat, input, filter_ksz, training, groups)
            except NotImplementedError:
                pass
        return ref_tr(feat, input, filter_ksz, training, groups)
    fl.apply_filter, fl.apply_feat_transpose = _apply, _tr
    done += ["ltr.models.layers.filter.apply_filter", "ltr.models.layers.filter.apply_feat_transpose"]
    dcf = importlib.import_module("pytracking.libs.dcf")
    ref_max2d = dcf.max2d
    dcf.max2d = lambda a: max2d(a) if (a.is_cuda and a.dtype == torch.float32) else ref_max2d(a)
    done.append("pytracking.libs.dcf.max2d")
    return done
