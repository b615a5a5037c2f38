"""TEST INFRASTRUCTURE ONLY -- generates tests/golden/*.npz by running the UNMODIFIED reference.

Run in the build container (needs /root/reference):   python -m oracle.gen_golden [--only NAME]

Inputs are regenerated from seeds by `pytracking_b200.synth` in the tests, so the fixtures only hold
the reference's OUTPUTS (plus tiny inputs where convenient).  The reference code paths exercised:
  corr      ltr/models/layers/filter.py: apply_filter, apply_feat_transpose (eval -> _v2), pytracking/libs/dcf.py: max2d
  labels    ltr/models/layers/distance.py DistanceMap + the three 1x1 predictors (optimizer.py:111-119)
  dimp_sd   ltr/models/target_classifier/optimizer.py DiMPSteepestDescentGN.forward
  prdimp_sd ltr/models/target_classifier/optimizer.py PrDiMPSteepestDescentNewton.forward
  backbone  ltr/models/backbone/resnet.py ResNet.forward + classifier.feature_extractor (dimpnet50 / dimpnet18)
"""
import argparse
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_shims  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")


def _np(t):
    return t.detach().cpu().numpy()


def gen_corr():
    import ltr.models.layers.filter as filter_layer
    from pytracking import dcf
    from pytracking_b200 import synth
    out = {}
    for tag, (n, c, h, k) in {"a": (3, 32, 18, 4), "b": (2, 64, 22, 4), "c": (2, 32, 18, 1), "d": (2, 32, 13, 3)}.items():
        feat = synth.make_clf_features(100 + ord(tag), n, c, h, h, filter_size=max(k, 1))
        g = torch.Generator().manual_seed(200 + ord(tag))
        w = torch.randn(1, c, k, k, generator=g) * 0.5
        s = filter_layer.apply_filter(feat.unsqueeze(1), w)            # (images, sequences=1, Ho, Wo)
        r = torch.randn(s.shape, generator=g)
        gt = filter_layer.apply_feat_transpose(feat.unsqueeze(1), r, (k, k), training=False)
        gt3 = filter_layer.apply_feat_transpose(feat.unsqueeze(1), r, (k, k), training=True)
        assert torch.allclose(gt, gt3, rtol=1e-4, atol=1e-5), "v2 vs v3 adjoint differ"
        mv, mi = dcf.max2d(s.squeeze(1))
        out.update({tag + "_w": _np(w), tag + "_scores": _np(s), tag + "_r": _np(r), tag + "_grad": _np(gt),
                    tag + "_maxval": _np(mv), tag + "_maxidx": _np(mi)})
    np.savez_compressed(os.path.join(GOLDEN, "corr.npz"), **out)


def _ref_dimp_optimizer(lut_seed):
    from ltr.models.target_classifier.optimizer import DiMPSteepestDescentGN
    from pytracking_b200 import synth
    opt = DiMPSteepestDescentGN(num_iter=5, feat_stride=16, init_step_length=0.9, init_filter_reg=0.1,
                                init_gauss_sigma=0.9, num_dist_bins=100, bin_displacement=0.1,
                                mask_init_factor=3.0, score_act='relu', mask_act='sigmoid')
    p = synth.make_dimp_optimizer_params(seed=lut_seed)
    missing = opt.load_state_dict(p, strict=True)
    opt.eval()
    return opt, p


def gen_labels():
    from pytracking_b200 import synth
    opt, p = _ref_dimp_optimizer(lut_seed=5)
    bb = synth.make_boxes(11, 6)
    bb[4] = torch.tensor([-40.0, 300.0, 30.0, 20.0])   # centre far outside the map -> last-bin clamp
    bb[5] = torch.tensor([136.0, 136.0, 16.0, 16.0])   # centre exactly on a cell
    with torch.no_grad():
        center = ((bb[..., :2] + bb[..., 2:] / 2) / 16).reshape(-1, 2).flip((1,))
        dm = opt.distance_map(center, (19, 19))
        y = opt.label_map_predictor(dm)[:, 0]
        m = opt.target_mask_predictor(dm)[:, 0]
        v = opt.spatial_weight_predictor(dm)[:, 0]
    np.savez_compressed(os.path.join(GOLDEN, "labels.npz"), bb=_np(bb), y=_np(y), m=_np(m), v=_np(v))


def gen_dimp_sd():
    from pytracking_b200 import synth
    out = {}
    cases = {"n15_it10": (15, 512, 18, 10, True, 21), "n50_it2": (50, 512, 18, 2, True, 22),
             "n4_c64": (4, 64, 18, 3, False, 23), "n7_22": (7, 128, 22, 4, True, 24)}
    for tag, (n, c, h, it, use_sw, seed) in cases.items():
        opt, p = _ref_dimp_optimizer(lut_seed=seed)
        feat = synth.make_clf_features(seed, n, c, h, h)
        bb = synth.make_boxes(seed + 1, n, center=(h * 16) / 2 - 25)
        g = torch.Generator().manual_seed(seed + 2)
        w0 = torch.randn(1, c, 4, 4, generator=g) * 0.02 if tag != "n15_it10" else torch.zeros(1, c, 4, 4)
        sw = None
        if use_sw:
            sw = torch.rand(n, generator=g) + 0.1
            sw = sw / sw.sum()
        with torch.no_grad():
            wf, its, losses = opt(w0, feat=feat.unsqueeze(1), bb=bb.unsqueeze(1),
                                  sample_weight=None if sw is None else sw.reshape(n, 1), num_iter=it,
                                  compute_losses=True)
        out[tag + "_w0"] = _np(w0)
        out[tag + "_sw"] = _np(sw) if sw is not None else np.zeros(0, np.float32)
        out[tag + "_wfinal"] = _np(wf)
        out[tag + "_w1"] = _np(its[1])
        out[tag + "_losses"] = np.array([float(l) for l in losses], dtype=np.float64)
        print(tag, "losses", [round(float(l), 5) for l in losses])
    np.savez_compressed(os.path.join(GOLDEN, "dimp_sd.npz"), **out)


def gen_prdimp_sd():
    from ltr.models.target_classifier.optimizer import PrDiMPSteepestDescentNewton
    from pytracking_b200 import synth
    out = {}
    # hyper-parameters: ltr/train_settings/dimp/prdimp50.py:95-98 (optim_init_reg=0.05, gauss_sigma=output_sigma*feature_sz,
    # alpha_eps=0.05, normalize_label=True) + pytracking/parameter/dimp/prdimp50.py (softmax_reg etc. overwritten at run time)
    cases = {"n15_22": (15, 512, 22, 10, 31, None, 0.0), "n6_18": (6, 64, 18, 3, 32, -2.0, 0.05)}
    for tag, (n, c, h, it, seed, sreg, lthr) in cases.items():
        sigma = (1 / 4 / 6.0) * h   # output_sigma_factor/search_area_factor * feature_sz (train_settings/dimp/prdimp50.py:20-25)
        opt = PrDiMPSteepestDescentNewton(num_iter=5, feat_stride=16, init_step_length=1.0, init_filter_reg=0.05,
                                          gauss_sigma=sigma, min_filter_reg=0.05, alpha_eps=0.05,
                                          normalize_label=True, softmax_reg=sreg, label_threshold=lthr,
                                          label_shrink=0.0 if sreg is None else 0.1)
        opt.eval()
        feat = synth.make_clf_features(seed, n, c, h, h)
        bb = synth.make_boxes(seed + 1, n, center=(h * 16) / 2 - 25)
        g = torch.Generator().manual_seed(seed + 2)
        w0 = torch.randn(1, c, 4, 4, generator=g) * 0.02
        sw = torch.rand(n, generator=g) + 0.1
        sw = sw / sw.sum()
        with torch.no_grad():
            wf, its, losses = opt(w0, feat=feat.unsqueeze(1), bb=bb.unsqueeze(1), sample_weight=sw.reshape(n, 1),
                                  num_iter=it, compute_losses=True)
        out[tag + "_w0"] = _np(w0)
        out[tag + "_sw"] = _np(sw)
        out[tag + "_wfinal"] = _np(wf)
        out[tag + "_w1"] = _np(its[1])
        out[tag + "_losses"] = np.array([float(l) for l in losses], dtype=np.float64)
        out[tag + "_sigma"] = np.array(sigma)
        print(tag, "losses", [round(float(l), 5) for l in losses])
    np.savez_compressed(os.path.join(GOLDEN, "prdimp_sd.npz"), **out)


def gen_backbone():
    from ltr.models.tracking import dimpnet
    from oracle import dimp_oracle
    from pytracking_b200 import synth
    out = {}
    common = dict(filter_size=4, backbone_pretrained=False, optim_iter=5, clf_feat_norm=True, final_conv=True,
                  optim_init_step=0.9, optim_init_reg=0.1, init_gauss_sigma=0.9, num_dist_bins=100,
                  bin_displacement=0.1, mask_init_factor=3.0, target_mask_act='sigmoid', score_act='relu')
    for arch, ctor, kw in (("resnet50", dimpnet.dimpnet50, dict(clf_feat_blocks=0, out_feature_dim=512)),
                           ("resnet18", dimpnet.dimpnet18, dict(clf_feat_blocks=1, out_feature_dim=256))):
        net = ctor(**common, **kw)
        sd = synth.make_dimp_state_dict(arch, seed=0, lut_seed=3)
        res = net.load_state_dict(sd, strict=False)
        bad = [k for k in res.missing_keys if not (k.startswith("bb_regressor") or "layer4" in k or ".fc." in k)]
        assert not bad and not res.unexpected_keys, (bad, res.unexpected_keys)
        net.eval()
        for size, seed in ((288, 41), (96, 42)):
            im = dimp_oracle.preprocess_image(synth.make_crop(seed, 1, size))
            with torch.no_grad():
                bf = net.extract_backbone_features(im)           # OrderedDict layer2, layer3
                clf = net.extract_classification_feat(bf)
            tag = "%s_%d_" % (arch, size)
            l2 = bf["layer2"]
            out[tag + "layer2"] = _np(l2 if size == 96 else l2[:, ::8])   # full map is 2.6 MB; keep every 8th channel
            out[tag + "layer3"] = _np(bf["layer3"])
            out[tag + "clf"] = _np(clf)
            print(tag, "layer3 absmax %.3f mean %.4f  clf absmax %.5f" % (bf["layer3"].abs().max(), bf["layer3"].mean(), clf.abs().max()))
    np.savez_compressed(os.path.join(GOLDEN, "backbone.npz"), **out)


GENS = {"corr": gen_corr, "labels": gen_labels, "dimp_sd": gen_dimp_sd, "prdimp_sd": gen_prdimp_sd,
        "backbone": gen_backbone}


def gen_atom_cg():
    """ConjugateGradient.run on ConvProblem exactly as ATOM wires it (pytracking/tracker/atom/atom.py:189-217)."""
    from pytracking import TensorList
    from pytracking.libs.optimization import ConjugateGradient
    from pytracking.tracker.atom.optim import ConvProblem
    from pytracking_b200 import synth
    import torch.nn.functional as F
    out = {}
    cfgs = {"n12_c16_pr_mlu": (12, 16, 12, 5, False, "mlu", 51), "n40_c64_pr_mlu": (40, 64, 25, 5, False, "mlu", 52),
            "n9_c32_fr_none": (9, 32, 9, 4, True, "none", 53), "n20_c64_pr_relu": (20, 64, 20, 3, False, "relu", 54)}
    for tag, (n, c, nf, it, fr, act, seed) in cfgs.items():
        x, y, sw = synth.make_atom_memory(seed, n, c, 18, 18, n_filled=nf)
        g = torch.Generator().manual_seed(seed + 100)
        w0 = torch.randn(1, c, 4, 4, generator=g) * 0.02
        if act == "mlu":
            fn = lambda t: F.elu(F.leaky_relu(t, 1 / 0.05), 0.05)
        elif act == "relu":
            fn = torch.nn.ReLU(inplace=False)
        else:
            fn = lambda t: t
        filt = TensorList([w0.clone()])
        prob = ConvProblem(TensorList([x]), TensorList([y]), TensorList([0.1]), TensorList([sw]), fn)
        opt = ConjugateGradient(prob, filt, fletcher_reeves=fr, direction_forget_factor=0, debug=False)
        opt.run(it)
        out.update({tag + "_w0": _np(w0), tag + "_w": _np(filt[0]).copy(), tag + "_iters": np.array(it)})
        # a second run from the updated filter (state is reset every run when the forget factor is 0)
        opt.run(it)
        out[tag + "_w2"] = _np(filt[0])
    np.savez_compressed(os.path.join(GOLDEN, "atom_cg.npz"), **out)


def gen_fourier():
    """ATOM.localize_target's Fourier chain with the reference's own fourier/complex modules (atom.py:304-316)."""
    import math
    from pytracking import fourier, TensorList
    out = {}
    for tag, (S, H, ksz, osz, seed) in {"s18_k4": (3, 18, 4, 288, 61), "s18_k4_o72": (2, 18, 4, 72, 62), "s17_k5": (2, 17, 5, 64, 63),
                                        "s22_k4": (1, 22, 4, 352, 64)}.items():
        g = torch.Generator().manual_seed(seed)
        sc = torch.randn(S, 1, H, H, generator=g)
        sf = fourier.cfft2(TensorList([sc])) / (H * H)
        sf[0] = fourier.shift_fs(sf[0], math.pi * (1 - torch.Tensor([ksz % 2, ksz % 2]) / torch.Tensor([H, H])))
        fs = fourier.sum_fs(sf)
        up = fourier.sample_fs(fs, torch.Tensor([osz, osz]))
        out.update({tag + "_scores": _np(sc), tag + "_up": _np(up)})
    np.savez_compressed(os.path.join(GOLDEN, "fourier.npz"), **out)


def gen_dimp_l2_sd():
    from ltr.models.target_classifier.optimizer import DiMPL2SteepestDescentGN
    from pytracking_b200 import synth
    out = {}
    for tag, (n, c, h, it, use_sw, thr, seed) in {"n8_c64": (8, 64, 18, 4, True, 0.05, 71), "n5_c32_22": (5, 32, 22, 3, False, -999.0, 72)}.items():
        opt = DiMPL2SteepestDescentGN(num_iter=it, feat_stride=16, init_step_length=0.9, gauss_sigma=1.3, hinge_threshold=thr,
                                      init_filter_reg=0.1, min_filter_reg=1e-3, alpha_eps=0.01)
        opt.eval()
        feat = synth.make_clf_features(seed, n, c, h, h)
        bb = synth.make_boxes(seed + 1, n, center=(h * 16) / 2 - 25)
        g = torch.Generator().manual_seed(seed + 2)
        w0 = torch.randn(1, c, 4, 4, generator=g) * 0.05 if tag == "n8_c64" else torch.zeros(1, c, 4, 4)
        sw = (torch.rand(n, generator=g) + 0.5) if use_sw else None
        if sw is not None:
            sw = sw / sw.sum()
        with torch.no_grad():
            wf, its, losses = opt(w0, feat.unsqueeze(1), bb.unsqueeze(1), sample_weight=None if sw is None else sw.unsqueeze(1),
                                  num_iter=it, compute_losses=True)
        out.update({tag + "_w0": _np(w0), tag + "_wfinal": _np(wf), tag + "_w1": _np(its[1]),
                    tag + "_losses": np.array([float(l) for l in losses], dtype=np.float32)})
        if sw is not None:
            out[tag + "_sw"] = _np(sw)
    np.savez_compressed(os.path.join(GOLDEN, "dimp_l2_sd.npz"), **out)


def gen_gn_sd_hinge():
    from ltr.models.meta.steepestdescent import GNSteepestDescent
    from ltr.models.target_classifier.residual_modules import LinearFilterHinge
    from pytracking import TensorList
    from pytracking_b200 import synth
    out = {}
    for tag, (n, c, h, it, use_sw, thr, leak, act, seed) in {"relu_n6_c64": (6, 64, 18, 4, True, 0.05, 0.0, "relu", 81),
                                                            "bent_n4_c32_22": (4, 32, 22, 3, False, 0.1, 0.1, "bentpar", 82)}.items():
        res = LinearFilterHinge(feat_stride=16, init_filter_reg=0.1, hinge_threshold=thr, activation_leak=leak, score_act=act,
                                act_param=0.7 if act == "bentpar" else None)
        opt = GNSteepestDescent(residual_module=res, num_iter=it, residual_batch_dim=1, compute_losses=True, steplength_reg=0.02)
        opt.eval()
        feat = synth.make_clf_features(seed, n, c, h, h)
        g = torch.Generator().manual_seed(seed + 2)
        ctr = torch.rand(n, 2, generator=g) * 4 + (h / 2 - 2)
        k0 = torch.arange(h + 1, dtype=torch.float32).view(1, -1, 1)
        k1 = torch.arange(h + 1, dtype=torch.float32).view(1, 1, -1)
        label = torch.exp(-0.5 / 1.5 ** 2 * (k0 - ctr[:, 0].view(-1, 1, 1)) ** 2) * torch.exp(-0.5 / 1.5 ** 2 * (k1 - ctr[:, 1].view(-1, 1, 1)) ** 2)
        w0 = torch.randn(1, c, 4, 4, generator=g) * 0.05
        sw = (torch.rand(n, generator=g) + 0.5) if use_sw else None
        if sw is not None:
            sw = sw / sw.sum()
        wf, its, losses = opt(TensorList([w0.clone()]), feat=feat.unsqueeze(1), train_label=label.unsqueeze(1),
                              sample_weight=None if sw is None else sw.reshape(n, 1, 1, 1), num_iter=it)
        out.update({tag + "_w0": _np(w0), tag + "_label": _np(label), tag + "_wfinal": _np(wf[0]), tag + "_w1": _np(its[1][0]),
                    tag + "_losses": np.array([float(l) for l in losses], dtype=np.float32)})
        if sw is not None:
            out[tag + "_sw"] = _np(sw)
    np.savez_compressed(os.path.join(GOLDEN, "gn_sd_hinge.npz"), **out)


TRANSFORMER_CASES = {"small": (64, 2, 128, 2, 2, 40, 2, 91, True), "tomp_l72": (256, 8, 2048, 6, 6, 72, 2, 92, True),
                     "tomp_l48_nomask": (256, 8, 2048, 6, 6, 48, 1, 93, False)}


def gen_transformer():
    """The reference's Transformer module (ltr/models/transformer/transformer.py) with seeded weights."""
    from ltr.models.transformer.transformer import Transformer
    from pytracking_b200 import synth
    out = {}
    for tag, (d, nh, ff, ne, nd, L, B, seed, use_mask) in TRANSFORMER_CASES.items():
        net = Transformer(d_model=d, nhead=nh, num_encoder_layers=ne, num_decoder_layers=nd, dim_feedforward=ff, dropout=0.1,
                          activation="relu", normalize_before=False)
        sd = synth.make_transformer_state_dict(seed, d, nh, ff, ne, nd)
        missing = net.load_state_dict(sd, strict=True)
        net.eval()
        g = torch.Generator().manual_seed(seed + 1)
        src = torch.randn(L, B, d, generator=g)
        pos = torch.randn(L, 1, d, generator=g) * 0.5
        qe = torch.randn(1, d, generator=g)
        mask = None
        if use_mask:
            mask = torch.zeros(B, L, dtype=torch.bool)
            mask[B - 1, L // 3: L // 2] = True
        with torch.no_grad():
            hs, mem = net(src, mask, qe, pos)
        out.update({tag + "_hs": _np(hs), tag + "_memory": _np(mem)})
    np.savez_compressed(os.path.join(GOLDEN, "transformer.npz"), **out)


ATOM_GN_CASES = {"n6_c32_16": (6, 32, 16, 3, 2, True, "mlu", 101), "n10_c64_32_pr": (10, 64, 32, 4, 3, False, "relu", 102)}


def gen_atom_gn():
    """GaussNewtonCG on FactorizedConvProblem as ATOM.init_optimization wires it (atom.py:157-178)."""
    from pytracking import TensorList
    from pytracking.libs.optimization import GaussNewtonCG
    from pytracking.tracker.atom.optim import FactorizedConvProblem
    from pytracking_b200 import synth
    import torch.nn.functional as F
    out = {}
    for tag, (n, cin, cc, ncg, ngn, fr, act, seed) in ATOM_GN_CASES.items():
        x, y, sw = synth.make_atom_memory(seed, n, cin, 18, 18)
        g = torch.Generator().manual_seed(seed + 7)
        w0 = torch.zeros(1, cc, 4, 4) if tag.startswith("n6") else torch.randn(1, cc, 4, 4, generator=g) * 0.02
        P0 = torch.randn(cc, cin, 1, 1, generator=g) * (1.0 / cin ** 0.5)
        fn = (lambda t: F.elu(F.leaky_relu(t, 1 / 0.05), 0.05)) if act == "mlu" else torch.nn.ReLU(inplace=False)
        filt, proj = TensorList([w0.clone()]), TensorList([P0.clone()])
        prob = FactorizedConvProblem(TensorList([x]), TensorList([y]), TensorList([0.1]), TensorList([1e-2]), None, TensorList([sw]),
                                     lambda t: t, fn)
        var = filt.concat(proj)
        opt = GaussNewtonCG(prob, var, fletcher_reeves=fr, debug=False)
        opt.run(ncg, ngn)
        out.update({tag + "_w0": _np(w0), tag + "_P0": _np(P0), tag + "_w": _np(var[0]).copy(), tag + "_P": _np(var[1]).copy()})
    np.savez_compressed(os.path.join(GOLDEN, "atom_gn.npz"), **out)


def gen_softmax_reg():
    """ltr/models/layers/activation.py:7-16 softmax_reg over the flattened score map, with and without the extra logit
    (PrDiMP score pre-processing, pytracking/tracker/dimp/dimp.py:206-210)."""
    from ltr.models.layers.activation import softmax_reg
    out = {}
    g = torch.Generator().manual_seed(4242)
    for tag, (n, h, reg) in {"a": (5, 23, None), "b": (3, 19, -1.5), "c": (2, 19, 4.0)}.items():
        x = torch.randn(n, 1, h, h, generator=g) * 3.0
        y = softmax_reg(x.view(n, 1, -1), dim=2, reg=reg).view(x.shape)
        out[tag + "_x"] = _np(x).copy()
        out[tag + "_y"] = _np(y).copy()
        out[tag + "_reg"] = np.array([np.nan if reg is None else reg], dtype=np.float32)
    np.savez_compressed(os.path.join(GOLDEN, "softmax_reg.npz"), **out)


def gen_hann():
    """pytracking/libs/dcf.py: hann1d / hann2d / hann2d_clipped (the optional output window of localize_target, built at init)."""
    from pytracking import dcf
    out = {}
    for tag, (h, w, c) in {"a": (19, 19, True), "b": (19, 19, False), "c": (18, 18, False), "d": (23, 21, True), "e": (22, 17, False)}.items():
        out["h2_" + tag] = _np(dcf.hann2d(torch.tensor([h, w]), centered=c)).copy()
        out["h2_" + tag + "_arg"] = np.array([h, w, int(c)])
    for tag, (h, w, eh, ew, c) in {"a": (36, 36, 25, 26, True), "b": (36, 36, 25, 26, False), "c": (19, 19, 13, 13, False), "d": (288, 288, 200, 201, False)}.items():
        out["hc_" + tag] = _np(dcf.hann2d_clipped(torch.tensor([h, w]), torch.tensor([eh, ew]), centered=c)).copy()
        out["hc_" + tag + "_arg"] = np.array([h, w, eh, ew, int(c)])
    np.savez_compressed(os.path.join(GOLDEN, "hann.npz"), **out)


def gen_tomp_pos():
    """ltr/models/transformer/position_encoding.py PositionEmbeddingSine as FilterPredictor builds and calls it."""
    import contextlib, io
    from ltr.models.transformer.position_encoding import PositionEmbeddingSine
    out = {}
    for tag, (h, w, d, res) in {"a": (18, 18, 256, 18), "b": (22, 22, 256, 22), "c": (18, 16, 128, 18)}.items():
        with contextlib.redirect_stdout(io.StringIO()):
            pe = PositionEmbeddingSine(num_pos_feats=d // 2, sine_type='lin_sine', avoid_aliazing=True, max_spatial_resolution=res)
        out[tag] = _np(pe(torch.zeros((1, h, w), dtype=torch.bool))[0]).copy()
        out[tag + "_arg"] = np.array([h, w, d, res])
    np.savez_compressed(os.path.join(GOLDEN, "tomp_pos.npz"), **out)


GENS["tomp_pos"] = gen_tomp_pos
GENS["hann"] = gen_hann
GENS["softmax_reg"] = gen_softmax_reg
GENS["atom_gn"] = gen_atom_gn
GENS["transformer"] = gen_transformer
GENS["gn_sd_hinge"] = gen_gn_sd_hinge
GENS["dimp_l2_sd"] = gen_dimp_l2_sd
GENS["atom_cg"] = gen_atom_cg
GENS["fourier"] = gen_fourier


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default=None)
    args = ap.parse_args()
    ref_shims.install()
    os.makedirs(GOLDEN, exist_ok=True)
    torch.set_num_threads(8)
    for name, fn in GENS.items():
        if args.only and name != args.only:
            continue
        print("== golden:", name)
        fn()


if __name__ == "__main__":
    main()
