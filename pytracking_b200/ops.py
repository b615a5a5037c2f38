"""Torch-tensor wrappers over the C ABI (device pointers + current stream; torch is plumbing only).

Every function requires CUDA tensors and raises if they are not (no CPU / eager fallback).
"""
import ctypes as C

import torch

from . import _lib


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _dev(t, name, contiguous=True):
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise RuntimeError("b200trk: '%s' must be a CUDA tensor (the engine has no CPU path)" % name)
    if t.dtype != torch.float32:
        raise RuntimeError("b200trk: '%s' must be float32, got %s" % (name, t.dtype))
    if t.device.index != torch.cuda.current_device():
        # the library's stream, scratch buffers and SM count belong to the CURRENT device (cudaGetDevice)
        raise RuntimeError("b200trk: '%s' lives on %s but the current device is cuda:%d; wrap the call in torch.cuda.device(%r)"
                           % (name, t.device, torch.cuda.current_device(), str(t.device)))
    return t.contiguous() if contiguous else t


def _p(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def apply_filter(feat, filt, return_max=False):
    """feat [n,C,H,W], filt [1,C,k,k] -> scores [n,1,Ho,Wo] (and max2d values / indices)."""
    feat, filt = _dev(feat, "feat"), _dev(filt, "filter")
    n, c, h, w = feat.shape
    k = filt.shape[-1]
    if filt.shape != (1, c, k, k):
        raise RuntimeError("b200trk.apply_filter: filter shape %s does not match features %s" % (tuple(filt.shape), tuple(feat.shape)))
    ho, wo = h + (k + 1) % 2, w + (k + 1) % 2
    scores = torch.empty(n, 1, ho, wo, device=feat.device, dtype=torch.float32)
    mv = mi = None
    if return_max:
        mv = torch.empty(n, device=feat.device, dtype=torch.float32)
        mi = torch.empty(n, 2, device=feat.device, dtype=torch.int64)
    _lib.check(_lib.lib().b200trk_apply_filter(_p(feat), _p(filt), _p(scores), n, c, h, w, k, _p(mv), _p(mi), _stream()),
               "apply_filter")
    return (scores, mv, mi) if return_max else scores


def apply_feat_transpose(feat, resid, k):
    """feat [n,C,H,W], resid [n,1,Ho,Wo] -> [1,C,k,k]."""
    feat, resid = _dev(feat, "feat"), _dev(resid, "input")
    n, c, h, w = feat.shape
    ho, wo = h + (k + 1) % 2, w + (k + 1) % 2
    if resid.numel() != n * ho * wo:
        raise RuntimeError("b200trk.apply_feat_transpose: input has %d elements, expected %d" % (resid.numel(), n * ho * wo))
    grad = torch.empty(1, c, k, k, device=feat.device, dtype=torch.float32)
    _lib.check(_lib.lib().b200trk_apply_feat_transpose(_p(feat), _p(resid), _p(grad), n, c, h, w, k, _stream()),
               "apply_feat_transpose")
    return grad


def softmax_reg(x, reg=None):
    """activation.softmax_reg(x, dim=-1, reg): softmax over the last dimension with an optional extra constant logit."""
    x = _dev(x, "x")
    L = x.shape[-1]
    n = x.numel() // L
    out = torch.empty_like(x)
    _lib.check(_lib.lib().b200trk_softmax_reg(_p(x), _p(out), n, L, 0 if reg is None else 1, 0.0 if reg is None else float(reg), _stream()),
               "softmax_reg")
    return out


def max2d(a):
    """a [..., H, W] -> (max values [...], indices [..., 2]) with the reference's tie order."""
    a = _dev(a, "a")
    lead = a.shape[:-2]
    h, w = a.shape[-2:]
    n = 1
    for d in lead:
        n *= d
    mv = torch.empty(n, device=a.device, dtype=torch.float32)
    mi = torch.empty(n, 2, device=a.device, dtype=torch.int64)
    _lib.check(_lib.lib().b200trk_max2d(_p(a), n, h, w, _p(mv), _p(mi), _stream()), "max2d")
    return mv.reshape(lead), mi.reshape(*lead, 2)


def dimp_sd_gn(weights, feat, bb, sample_weight, label_lut, mask_lut, spatial_lut, num_iter, step_length, reg_weight,
               alpha_eps=0.0, bin_displacement=0.1, feat_stride=16.0, return_iterates=False, compute_losses=False,
               out=None):
    weights, feat, bb = _dev(weights, "weights"), _dev(feat, "feat"), _dev(bb, "bb")
    n, c, h, w = feat.shape
    k = weights.shape[-1]
    if sample_weight is not None:
        sample_weight = _dev(sample_weight, "sample_weight")
    luts = [_dev(t, "lut").reshape(-1) for t in (label_lut, mask_lut, spatial_lut)]
    nb = luts[0].numel()
    wout = out if out is not None else torch.empty_like(weights)
    its = torch.empty(num_iter + 1, c, k, k, device=feat.device, dtype=torch.float32) if return_iterates else None
    losses = torch.empty(num_iter + 1, device=feat.device, dtype=torch.float32) if compute_losses else None
    _lib.check(_lib.lib().b200trk_dimp_sd_gn(
        _p(weights), _p(wout), _p(feat), _p(bb), _p(sample_weight), n, c, h, w, k, int(num_iter),
        _p(luts[0]), _p(luts[1]), _p(luts[2]), nb, float(bin_displacement), float(feat_stride),
        float(step_length), float(reg_weight), float(alpha_eps), _p(its), _p(losses), _stream()), "dimp_sd_gn")
    return wout, its, losses


def dimp_l2_sd_gn(weights, feat, bb, sample_weight, num_iter, gauss_sigma, hinge_threshold, step_length, reg_weight,
                  alpha_eps=0.0, feat_stride=16.0, return_iterates=False, compute_losses=False, out=None):
    """DiMPL2SteepestDescentGN.forward (one sequence)."""
    weights, feat, bb = _dev(weights, "weights"), _dev(feat, "feat"), _dev(bb, "bb")
    n, c, h, w = feat.shape
    k = weights.shape[-1]
    if sample_weight is not None:
        sample_weight = _dev(sample_weight, "sample_weight")
    wout = out if out is not None else torch.empty_like(weights)
    its = torch.empty(num_iter + 1, c, k, k, device=feat.device, dtype=torch.float32) if return_iterates else None
    losses = torch.empty(num_iter + 1, device=feat.device, dtype=torch.float32) if compute_losses else None
    _lib.check(_lib.lib().b200trk_dimp_l2_sd_gn(
        _p(weights), _p(wout), _p(feat), _p(bb), _p(sample_weight), n, c, h, w, k, int(num_iter), float(gauss_sigma),
        float(hinge_threshold), float(feat_stride), float(step_length), float(reg_weight), float(alpha_eps), _p(its), _p(losses),
        _stream()), "dimp_l2_sd_gn")
    return wout, its, losses


def gn_sd_hinge(weights, feat, train_label, sample_weight, num_iter, filter_reg, hinge_threshold=-999.0, activation_leak=0.0,
                score_act="relu", act_param=1.0, steplength_reg=0.0, return_iterates=False, compute_losses=False, out=None):
    """GNSteepestDescent.forward with the LinearFilterHinge residual module (one sequence)."""
    weights, feat, train_label = _dev(weights, "weights"), _dev(feat, "feat"), _dev(train_label, "train_label")
    n, c, h, w = feat.shape
    k = weights.shape[-1]
    if train_label.numel() != n * (h + 1) * (w + 1):
        raise RuntimeError("b200trk.gn_sd_hinge: train_label must be [n,1,H+1,W+1]")
    if sample_weight is not None:
        sample_weight = _dev(sample_weight, "sample_weight")
    wout = out if out is not None else torch.empty_like(weights)
    its = torch.empty(num_iter + 1, c, k, k, device=feat.device, dtype=torch.float32) if return_iterates else None
    losses = torch.empty(num_iter + 1, device=feat.device, dtype=torch.float32) if compute_losses else None
    _lib.check(_lib.lib().b200trk_gn_sd_hinge(
        _p(weights), _p(wout), _p(feat), _p(train_label), _p(sample_weight), n, c, h, w, k, int(num_iter), float(filter_reg),
        float(hinge_threshold), float(activation_leak), {"relu": 0, "bentpar": 1}[score_act], float(act_param),
        float(steplength_reg), _p(its), _p(losses), _stream()), "gn_sd_hinge")
    return wout, its, losses


def prdimp_sd_newton(weights, feat, bb, sample_weight, num_iter, gauss_sigma, step_length, reg_weight, alpha_eps=0.0,
                     softmax_reg=None, label_threshold=0.0, normalize_label=False, label_shrink=0.0, uni_weight=0.0,
                     feat_stride=16.0, return_iterates=False, compute_losses=False, out=None):
    weights, feat, bb = _dev(weights, "weights"), _dev(feat, "feat"), _dev(bb, "bb")
    n, c, h, w = feat.shape
    k = weights.shape[-1]
    if sample_weight is not None:
        sample_weight = _dev(sample_weight, "sample_weight")
    wout = out if out is not None else torch.empty_like(weights)
    its = torch.empty(num_iter + 1, c, k, k, device=feat.device, dtype=torch.float32) if return_iterates else None
    losses = torch.empty(num_iter + 1, device=feat.device, dtype=torch.float32) if compute_losses else None
    _lib.check(_lib.lib().b200trk_prdimp_sd_newton(
        _p(weights), _p(wout), _p(feat), _p(bb), _p(sample_weight), n, c, h, w, k, int(num_iter),
        float(gauss_sigma), float(feat_stride), float(step_length), float(reg_weight), float(alpha_eps),
        0 if softmax_reg is None else 1, 0.0 if softmax_reg is None else float(softmax_reg), float(label_threshold),
        1 if normalize_label else 0, float(label_shrink), float(uni_weight), _p(its), _p(losses), _stream()),
        "prdimp_sd_newton")
    return wout, its, losses


def conv2d_same(feat, filt):
    """operation.conv2d(feat, filt, mode='same') for one 4x4 filter: [n,C,H,W] x [1,C,4,4] -> [n,1,H,W]."""
    feat, filt = _dev(feat, "input"), _dev(filt, "weight")
    n, c, h, w = feat.shape
    k = filt.shape[-1]
    out = torch.empty(n, 1, h, w, device=feat.device, dtype=torch.float32)
    _lib.check(_lib.lib().b200trk_conv2d_same(_p(feat), _p(filt), _p(out), n, c, h, w, k, _stream()), "conv2d_same")
    return out


def conv1x1(x, weight):
    """operation.conv1x1: [S,Cin,H,W] x [Cout,Cin,1,1] -> [S,Cout,H,W]."""
    x, weight = _dev(x, "input"), _dev(weight, "weight")
    s, cin, h, w = x.shape
    cout = weight.shape[0]
    if weight.numel() != cout * cin:
        raise RuntimeError("b200trk.conv1x1: weight %s does not match input channels %d" % (tuple(weight.shape), cin))
    out = torch.empty(s, cout, h, w, device=x.device, dtype=torch.float32)
    _lib.check(_lib.lib().b200trk_conv1x1(_p(x), _p(weight), _p(out), s, cin, cout, h, w, _stream()), "conv1x1")
    return out


def feature_normalize_(feat, normalize_power=2.0):
    """In-place p-norm normalisation of MultiFeatureBase.get_feature."""
    if not feat.is_cuda or feat.dtype != torch.float32 or not feat.is_contiguous():
        raise RuntimeError("b200trk.feature_normalize_: needs a contiguous CUDA float32 tensor")
    s, c, h, w = feat.shape
    _lib.check(_lib.lib().b200trk_feature_normalize(_p(feat), s, c, h, w, float(normalize_power), _stream()), "feature_normalize")
    return feat


def fourier_interp(scores, kernel_size, output_sz):
    """ATOM.localize_target: Fourier-series upsampling of [S,1,H,W] score maps to [S,1,oh,ow]."""
    scores = _dev(scores, "scores")
    s, _, h, w = scores.shape
    oh, ow = int(output_sz[0]), int(output_sz[1])
    out = torch.empty(s, 1, oh, ow, device=scores.device, dtype=torch.float32)
    _lib.check(_lib.lib().b200trk_fourier_interp(_p(scores), _p(out), s, h, w, int(kernel_size[0]), int(kernel_size[1]), oh, ow,
                                                 _stream()), "fourier_interp")
    return out


ATOM_ACTIVATIONS = {"none": 0, "relu": 1, "elu": 2, "mlu": 3}


def atom_cg_filter(filt, feat, y, sample_weight, filter_reg, num_iter, activation="mlu", act_param=0.05,
                   fletcher_reeves=False, out=None):
    """ConjugateGradient.run(num_iter) on ConvProblem (ATOM online filter update). Returns the updated filter."""
    filt, feat, y, sample_weight = _dev(filt, "filter"), _dev(feat, "feat"), _dev(y, "y"), _dev(sample_weight, "sample_weight")
    n, c, h, w = feat.shape
    k = filt.shape[-1]
    if y.numel() != n * h * w or sample_weight.numel() != n:
        raise RuntimeError("b200trk.atom_cg_filter: label / weight shapes do not match the sample memory")
    wout = out if out is not None else torch.empty_like(filt)
    _lib.check(_lib.lib().b200trk_atom_cg_filter(_p(filt), _p(wout), _p(feat), _p(y), _p(sample_weight), n, c, h, w, k,
                                                 int(num_iter), float(filter_reg), 1 if fletcher_reeves else 0,
                                                 ATOM_ACTIVATIONS[activation], float(act_param), _stream()), "atom_cg_filter")
    return wout


def eco_filter_cg_(filt, samples, yf, sample_weights, reg_filter, sample_energy, num_iter, new_xf=None, state=None,
                   fletcher_reeves=False, standard_alpha=True, direction_forget_factor=0.0, precond_learning_rate=0.0075,
                   precond_data_param=0.3, precond_reg_param=0.15):
    """FilterOptim.run(num_iter, new_xf) for one ECO feature block (optim.py:140-208).  Updates IN PLACE: `filt` [1,C,H,Wh,2],
    `sample_energy` [1,C,H,Wh] (None: created from `new_xf`), and the CG state `state` = dict(p, r_prev, rho) of device tensors
    (None / empty: no previous direction).  Returns (sample_energy, state)."""
    for name, t in (("filter", filt), ("sample_energy", sample_energy)) + tuple((state or {}).items()):
        if t is not None and (not isinstance(t, torch.Tensor) or not t.is_cuda or t.dtype != torch.float32 or not t.is_contiguous()):
            raise RuntimeError("b200trk.eco_filter_cg_: '%s' must be a contiguous CUDA float32 tensor (updated in place)" % name)
    _dev(filt, "filter")
    samples, yf, sample_weights, reg_filter = _dev(samples, "training_samples"), _dev(yf, "yf"), _dev(sample_weights, "sample_weights"), \
        _dev(reg_filter, "reg_filter")
    if filt.dim() != 5 or filt.shape[0] != 1 or filt.shape[-1] != 2 or samples.dim() != 5:
        raise RuntimeError("b200trk.eco_filter_cg_: filter [1,C,H,Wh,2] and training_samples [H,Wh,N,C,2] expected")
    _, c, h, wh, _ = filt.shape
    n = samples.shape[2]
    if tuple(samples.shape) != (h, wh, n, c, 2) or yf.numel() != h * wh or sample_weights.numel() != n or samples.data_ptr() % 16:
        raise RuntimeError("b200trk.eco_filter_cg_: training_samples %s / yf / sample_weights do not match the filter %s"
                           % (tuple(samples.shape), tuple(filt.shape)))
    if int(num_iter) == 0:
        return sample_energy, state
    if new_xf is not None:
        new_xf = _dev(new_xf, "new_xf")
        if new_xf.numel() != filt.numel():
            raise RuntimeError("b200trk.eco_filter_cg_: new_xf %s does not match the filter" % (tuple(new_xf.shape),))
    has_energy = sample_energy is not None
    if not has_energy:
        if new_xf is None:
            raise RuntimeError("b200trk.eco_filter_cg_: no sample energy and no new sample")
        sample_energy = torch.empty(1, c, h, wh, device=filt.device, dtype=torch.float32)
    elif sample_energy.numel() != c * h * wh:
        raise RuntimeError("b200trk.eco_filter_cg_: sample_energy %s does not match the filter" % (tuple(sample_energy.shape),))
    has_state = bool(state) and state.get("p") is not None
    if not has_state:
        state = {"p": torch.empty_like(filt), "r_prev": None if fletcher_reeves else torch.empty_like(filt),
                 "rho": torch.ones(1, device=filt.device, dtype=torch.float32)}
    elif not fletcher_reeves and state.get("r_prev") is None:
        raise RuntimeError("b200trk.eco_filter_cg_: the Polak-Ribiere formula needs state['r_prev']")
    if state["p"].numel() != filt.numel() or state["rho"].numel() != 1:
        raise RuntimeError("b200trk.eco_filter_cg_: CG state does not match the filter")
    _lib.check(_lib.lib().b200trk_eco_filter_cg(
        _p(filt), _p(samples), _p(yf), _p(sample_weights), _p(reg_filter), int(reg_filter.shape[-2]), int(reg_filter.shape[-1]),
        _p(sample_energy), 1 if has_energy else 0, _p(new_xf), _p(state["p"]), _p(state.get("r_prev")), _p(state["rho"]),
        1 if has_state else 0, h, wh, n, c, int(num_iter), 1 if fletcher_reeves else 0, 1 if standard_alpha else 0,
        float(direction_forget_factor), float(precond_learning_rate), float(precond_data_param), float(precond_reg_param), _stream()),
        "eco_filter_cg")
    return sample_energy, state


def eco_joint_gn_(filt, proj, samples, yf, sample_weights_sqrt, reg_filter, diag_M_filter, diag_M_proj, projection_reg, num_cg_iter,
                  num_gn_iter):
    """GaussNewtonCG.run(num_cg_iter, num_gn_iter) on ECO's FactorizedConvProblem for one feature block (eco/optim.py:8-117); updates
    `filt` [1,C,H,Wh,2] and `proj` [Cin,C] in place."""
    for name, t in (("filter", filt), ("projection_matrix", proj)):
        if not isinstance(t, torch.Tensor) or not t.is_cuda or t.dtype != torch.float32 or not t.is_contiguous():
            raise RuntimeError("b200trk.eco_joint_gn_: '%s' must be a contiguous CUDA float32 tensor (updated in place)" % name)
    _dev(filt, "filter")
    samples, yf, sample_weights_sqrt, reg_filter, diag_M_filter = _dev(samples, "training_samples"), _dev(yf, "yf"), \
        _dev(sample_weights_sqrt, "sample_weights_sqrt"), _dev(reg_filter, "reg_filter"), _dev(diag_M_filter, "diag_M")
    if filt.dim() != 5 or filt.shape[0] != 1 or filt.shape[-1] != 2 or samples.dim() != 5 or proj.dim() != 2:
        raise RuntimeError("b200trk.eco_joint_gn_: filter [1,C,H,Wh,2], projection matrix [Cin,C], training_samples [H,Wh,N,Cin,2] expected")
    _, c, h, wh, _ = filt.shape
    n, cin = samples.shape[2], samples.shape[3]
    if tuple(samples.shape) != (h, wh, n, cin, 2) or tuple(proj.shape) != (cin, c) or yf.numel() != h * wh or \
            sample_weights_sqrt.numel() != n or diag_M_filter.numel() != c * h * wh:
        raise RuntimeError("b200trk.eco_joint_gn_: shapes of training_samples %s / projection matrix %s / yf / weights / diag_M do not "
                           "match the filter %s" % (tuple(samples.shape), tuple(proj.shape), tuple(filt.shape)))
    _lib.check(_lib.lib().b200trk_eco_joint_gn(
        _p(filt), _p(proj), _p(samples), _p(yf), _p(sample_weights_sqrt), _p(reg_filter), int(reg_filter.shape[-2]), int(reg_filter.shape[-1]),
        _p(diag_M_filter), float(diag_M_proj), float(projection_reg), h, wh, n, cin, c, int(num_cg_iter), int(num_gn_iter), _stream()),
        "eco_joint_gn")
    return filt, proj


def eco_apply_filter(filt, sample_xf):
    """ECO.apply_filter for one feature block (eco.py:244-245): filter [1,C,H,Wh,2] x sample_xf [S,C,H,Wh,2] -> [S,1,H,Wh,2]."""
    filt, sample_xf = _dev(filt, "filter"), _dev(sample_xf, "sample_xf")
    if filt.dim() != 5 or sample_xf.dim() != 5 or filt.shape[0] != 1 or filt.shape[-1] != 2 or tuple(sample_xf.shape[1:]) != tuple(filt.shape[1:]):
        raise RuntimeError("b200trk.eco_apply_filter: filter [1,C,H,Wh,2] and sample_xf [S,C,H,Wh,2] expected, got %s and %s"
                           % (tuple(filt.shape), tuple(sample_xf.shape)))
    s, c, h, wh, _ = sample_xf.shape
    sf = torch.empty(s, 1, h, wh, 2, device=filt.device, dtype=torch.float32)
    _lib.check(_lib.lib().b200trk_eco_apply_filter(_p(filt), _p(sample_xf), _p(sf), s, c, h, wh, _stream()), "eco_apply_filter")
    return sf


def eco_sample_fs(sf_blocks, output_sz, weights=None):
    """fourier.sample_fs(fourier.sum_fs(weight * sf), output_sz) of ECO.localize_target (eco.py:247-252; fourier.py:35-61, 95-114): the
    centred half spectra `sf_blocks` (a tensor or a list of tensors [S,1,H_b,Wh_b,2]) summed with `weights` and evaluated on the
    output_sz grid -> scores [S,1,oh,ow]."""
    import ctypes as C
    blocks = [sf_blocks] if isinstance(sf_blocks, torch.Tensor) else list(sf_blocks)
    blocks = [_dev(b, "sf") for b in blocks]
    nb = len(blocks)
    if nb == 0 or any(b.dim() != 5 or b.shape[1] != 1 or b.shape[-1] != 2 or b.shape[0] != blocks[0].shape[0] for b in blocks):
        raise RuntimeError("b200trk.eco_sample_fs: blocks [S,1,H,Wh,2] with a common S expected, got %s" % [tuple(b.shape) for b in blocks])
    if weights is not None and len(weights) != nb:
        raise RuntimeError("b200trk.eco_sample_fs: %d weights for %d blocks" % (len(weights), nb))
    s = blocks[0].shape[0]
    oh, ow = int(output_sz[0]), int(output_sz[1])
    out = torch.empty(s, 1, oh, ow, device=blocks[0].device, dtype=torch.float32)
    ptrs = (C.c_void_p * nb)(*[b.data_ptr() for b in blocks])
    hs, ws = (C.c_int * nb)(*[b.shape[2] for b in blocks]), (C.c_int * nb)(*[b.shape[3] for b in blocks])
    wts = (C.c_float * nb)(*[float(w) for w in weights]) if weights is not None else None
    _lib.check(_lib.lib().b200trk_eco_sample_fs(ptrs, hs, ws, wts, nb, s, oh, ow, _p(out), _stream()), "eco_sample_fs")
    return out


def eco_preprocess_sample_(x, window, interp_y, interp_x):
    """ECO.preprocess_sample for one feature block (eco.py:297-300): windows `x` [S,C,H,W] (any strides) IN PLACE (as the reference does) and returns
    interpolate_dft(cfft2(x), (interp_y, interp_x)) [S,C,H',Wh',2]."""
    x = _dev(x, "x", contiguous=False)                              # windowed in place: the caller's tensor or view, through its strides
    if x.dim() != 4 or min(x.stride()) < 0:
        raise RuntimeError("b200trk.eco_preprocess_sample_: x must be a [S,C,H,W] tensor or view with non-negative strides")
    window, interp_y, interp_x = _dev(window, "window"), _dev(interp_y, "interp_y"), _dev(interp_x, "interp_x")
    s, c, h, w = x.shape
    hp, whp = h + (h + 1) % 2, w // 2 + 1
    if window.numel() != h * w or interp_y.numel() != 2 * hp or interp_x.numel() != 2 * whp:
        raise RuntimeError("b200trk.eco_preprocess_sample_: window %s / interp_y %s / interp_x %s do not match the %dx%d feature map"
                           % (tuple(window.shape), tuple(interp_y.shape), tuple(interp_x.shape), h, w))
    xf = torch.empty(s, c, hp, whp, 2, device=x.device, dtype=torch.float32)
    st = x.stride()
    _lib.check(_lib.lib().b200trk_eco_preprocess_sample(_p(x), st[0], st[1], st[2], st[3], _p(window), _p(interp_y), _p(interp_x), _p(xf), s, c, h, w,
                                                        _stream()),
               "eco_preprocess_sample")
    return xf


def eco_shift_fs(a, shift_y, shift_x):
    """fourier.shift_fs (fourier.py:78-92): the centred half spectrum a [S,C,H,Wh,2] shifted by (shift_y, shift_x) radians per frequency."""
    a = _dev(a, "a")
    if a.dim() != 5 or a.shape[-1] != 2:
        raise RuntimeError("b200trk.eco_shift_fs: a must be the Fourier coefficients [S,C,H,Wh,2], got %s" % (tuple(a.shape),))
    s, c, h, wh, _ = a.shape
    out = torch.empty_like(a)
    _lib.check(_lib.lib().b200trk_eco_shift_fs(_p(a), _p(out), s, c, h, wh, float(shift_y), float(shift_x), _stream()), "eco_shift_fs")
    return out


def atom_gn_joint_(filt, proj, samples, y, sample_weight, filter_reg, projection_reg, num_cg_iter, num_gn_iter,
                   activation="mlu", act_param=0.05, fletcher_reeves=True):
    """GaussNewtonCG.run(num_cg_iter, num_gn_iter) on FactorizedConvProblem; updates `filt` and `proj` in place."""
    for name, t in (("filter", filt), ("projection_matrix", proj)):
        if not t.is_cuda or t.dtype != torch.float32 or not t.is_contiguous():
            raise RuntimeError("b200trk.atom_gn_joint_: '%s' must be a contiguous CUDA float32 tensor (updated in place)" % name)
    samples, y, sample_weight = _dev(samples, "samples"), _dev(y, "y"), _dev(sample_weight, "sample_weight")
    n, cin, h, w = samples.shape
    cc, k = filt.shape[1], filt.shape[-1]
    if proj.numel() != cc * cin:
        raise RuntimeError("b200trk.atom_gn_joint_: projection matrix %s does not match (%d, %d)" % (tuple(proj.shape), cc, cin))
    _lib.check(_lib.lib().b200trk_atom_gn_joint(_p(filt), _p(proj), _p(samples), _p(y), _p(sample_weight), n, cin, cc, h, w, k,
                                                int(num_cg_iter), int(num_gn_iter), float(filter_reg), float(projection_reg),
                                                1 if fletcher_reeves else 0, ATOM_ACTIVATIONS[activation], float(act_param), _stream()),
               "atom_gn_joint")
    return filt, proj


def prroi_pool_forward(features, rois, ph, pw, scale):
    features, rois = _dev(features, "features"), _dev(rois, "rois")
    b, c, h, w = features.shape
    r = rois.shape[0]
    out = torch.zeros(r, c, ph, pw, device=features.device, dtype=torch.float32)
    _lib.check(_lib.lib().b200trk_prroi_pool_forward(_p(features), _p(rois), _p(out), b, c, h, w, r, int(ph), int(pw),
                                                     float(scale), _stream()), "prroi_pool_forward")
    return out


def prroi_pool_backward(features, rois, output, output_grad, ph, pw, scale):
    features, rois, output, output_grad = (_dev(features, "features"), _dev(rois, "rois"), _dev(output, "output"),
                                           _dev(output_grad, "output_diff"))
    b, c, h, w = features.shape
    r = rois.shape[0]
    fg = torch.empty_like(features)
    _lib.check(_lib.lib().b200trk_prroi_pool_backward(_p(features), _p(rois), _p(output), _p(output_grad), _p(fg), b, c, h, w,
                                                      r, int(ph), int(pw), float(scale), _stream()), "prroi_pool_backward")
    return fg


def prroi_pool_coor_backward(features, rois, output, output_grad, ph, pw, scale):
    features, rois, output, output_grad = (_dev(features, "features"), _dev(rois, "rois"), _dev(output, "output"),
                                           _dev(output_grad, "output_diff"))
    b, c, h, w = features.shape
    r = rois.shape[0]
    rg = torch.zeros(r, 5, device=features.device, dtype=torch.float32)
    _lib.check(_lib.lib().b200trk_prroi_pool_coor_backward(_p(features), _p(rois), _p(output), _p(output_grad), _p(rg), b, c,
                                                           h, w, r, int(ph), int(pw), float(scale), _stream()),
               "prroi_pool_coor_backward")
    return rg
