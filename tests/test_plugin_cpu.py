"""CPU checks of `pytracking_b200.plugin.install()`: every seam of SURVEY.md 8(b) is rebound in the unmodified reference checkout, CPU /
autograd / unsupported-shape calls fall through to the reference implementation (reference configs with 14x14 or 16x16
feature maps must not crash after install()), and `uninstall()` restores the checkout.  No CUDA call is made here."""
import numpy as np
import pytest
import torch

from baseline import ref_env

pytestmark = pytest.mark.skipif(not ref_env.reference_available(), reason="reference tree not staged (baseline/_ref)")

SEAMS = ["filter.apply_filter", "filter.apply_feat_transpose", "dcf.max2d", "DiMPSteepestDescentGN.forward",
         "PrDiMPSteepestDescentNewton.forward", "DiMPL2SteepestDescentGN.forward", "GNSteepestDescent.forward", "NetWithBackbone.extract_backbone",
         "DiMPnet.extract_classification_feat", "functional._prroi_pooling", "operation.conv2d", "operation.conv1x1",
         "ConjugateGradient.run", "GaussNewtonCG.run", "FilterOptim.run", "Transformer.forward", "AtomIoUNet.get_iou_feat", "AtomIoUNet.predict_iou", "Head.extract_head_feat",
         "DenseBoxRegressor.forward", "FilterPredictor.predict_cls_bbreg_filters_parallel"]


@pytest.fixture()
def installed():
    from oracle import ref_shims
    ref_shims.install(prroi_cpu=False)
    from pytracking_b200 import plugin
    names = plugin.install()
    yield plugin, names
    plugin.uninstall()


def test_install_rebinds_every_seam_and_uninstall_restores(installed):
    plugin, names = installed
    for s in SEAMS:
        assert any(n.endswith(s) for n in names), (s, names)
    import ltr.models.layers.filter as fl
    import ltr.models.target_classifier.optimizer as opt
    patched = fl.apply_filter, opt.DiMPSteepestDescentGN.forward
    plugin.uninstall()
    assert fl.apply_filter is not patched[0] and opt.DiMPSteepestDescentGN.forward is not patched[1]
    assert fl.apply_filter.__module__ == "ltr.models.layers.filter"
    plugin.install()


def test_cpu_and_unsupported_shapes_fall_through(installed):
    plugin, _ = installed
    import ltr.models.layers.filter as fl
    from oracle import dimp_oracle as O
    g = torch.Generator().manual_seed(0)
    for hw, k in ((14, 4), (16, 4), (18, 3), (18, 4)):           # dimp50_vot18 (14x14), dimp50_vot19 (16x16), odd filter
        feat = torch.randn(3, 1, 32, hw, hw, generator=g)
        filt = torch.randn(1, 32, k, k, generator=g)
        s = fl.apply_filter(feat, filt)                           # CPU tensors: the reference path
        assert s.shape[-1] == hw + (k + 1) % 2
        if k == 4:
            assert torch.allclose(s.reshape(3, 1, *s.shape[-2:]), O.apply_filter(feat[:, 0], filt), atol=1e-4)
        r = torch.randn(3, 1, s.shape[-2], s.shape[-1], generator=g)
        assert fl.apply_feat_transpose(feat, r, (k, k), training=False).shape == (1, 32, k, k)
    assert not plugin.stats.get("apply_filter")
    # the mirror functions themselves reject what the library would reject -- with NotImplementedError, not RuntimeError
    for hw, c, k in ((14, 32, 4), (16, 32, 4), (18, 24, 4), (18, 32, 3)):
        with pytest.raises(NotImplementedError):
            plugin.apply_filter(torch.zeros(1, 1, c, hw, hw), torch.zeros(1, c, k, k))
        with pytest.raises(NotImplementedError):
            plugin.apply_feat_transpose(torch.zeros(1, 1, c, hw, hw), torch.zeros(1, 1, hw + 1, hw + 1), (k, k))
    with pytest.raises(NotImplementedError):
        plugin.apply_filter(torch.zeros(1, 1, 32, 18, 18, dtype=torch.float64), torch.zeros(1, 32, 4, 4, dtype=torch.float64))


def test_unsupported_optimizer_modules_are_rejected(installed):
    plugin, _ = installed
    import ltr.models.target_classifier.optimizer as opt
    m = opt.DiMPSteepestDescentGN(num_dist_bins=10, score_act="bentpar", act_param=0.1)
    with pytest.raises(NotImplementedError):
        plugin.DiMPSteepestDescentGN.from_module(m)
    m = opt.DiMPSteepestDescentGN(num_dist_bins=10, mask_act="linear")
    with pytest.raises(NotImplementedError):
        plugin.DiMPSteepestDescentGN.from_module(m)
    plugin.DiMPSteepestDescentGN.from_module(opt.DiMPSteepestDescentGN(num_dist_bins=10))


def test_response_activation_probe():
    from pytracking_b200 import plugin
    import torch.nn.functional as F
    assert plugin._probe_activation(lambda x: x) == ("none", 0.0)
    assert plugin._probe_activation(torch.nn.ReLU(inplace=True)) == ("relu", 0.0)
    assert plugin._probe_activation(torch.nn.ELU(inplace=True)) == ("elu", 1.0)
    name, par = plugin._probe_activation(lambda x: F.elu(F.leaky_relu(x, 1 / 0.05), 0.05))       # atom.py:468
    assert name == "mlu" and abs(par - 0.05) < 1e-6
    assert plugin._probe_activation(torch.tanh) is None


def test_reference_tracker_on_cpu_is_unchanged_by_install():
    """The whole reference DiMP tracker (CPU) gives bit-identical boxes with and without the plug-in installed."""
    from oracle import ref_shims
    ref_shims.install()
    from baseline import ref_tracker
    from pytracking_b200 import plugin, synth
    torch.set_num_threads(8)
    frames, bb = synth.make_sequence(0, num_frames=4)
    ov = dict(net_opt_iter=2, net_opt_update_iter=1)
    a = ref_tracker.run_sequence(ref_tracker.build_dimp("cpu", overrides=ov, use_augmentation=False), frames, bb)
    plugin.install()
    try:
        b = ref_tracker.run_sequence(ref_tracker.build_dimp("cpu", overrides=ov, use_augmentation=False), frames, bb)
    finally:
        plugin.uninstall()
    assert np.array_equal(a["target_bbox"], b["target_bbox"])


def test_eco_filter_optim_falls_through_on_cpu(installed):
    """FilterOptim.run (eco/optim.py:140) is rebound; CPU tensors keep the reference implementation and reproduce the golden run."""
    plugin, _ = installed
    import os
    from pytracking import TensorList
    from pytracking.tracker.eco.optim import FilterOptim
    from pytracking.utils import TrackerParams
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "eco_cg.npz"))
    case = "pr_forget"
    fr, sa, dff, pdp, prp = g[case + "/params"]
    params = TrackerParams()
    params.fletcher_reeves, params.standard_alpha, params.direction_forget_factor, params.debug = bool(fr), bool(sa), float(dff), 0
    params.precond_data_param, params.precond_reg_param = float(pdp), float(prp)
    params.precond_learning_rate = TensorList([float(g["%s/b%d/lr" % (case, b)]) for b in range(2)])
    T = lambda k: torch.from_numpy(g[k].copy())
    k0 = [case + "/run0/b%d/" % b for b in range(2)]
    filt = TensorList([T(k + "hf_in") for k in k0])
    reg = TensorList([T("%s/b%d/reg_filter" % (case, b)) for b in range(2)])
    opt = FilterOptim(params, reg.view(-1) @ reg.view(-1))
    opt.register(filt, TensorList([T(k + "samples") for k in k0]), TensorList([T("%s/b%d/yf" % (case, b)) for b in range(2)]),
                 TensorList([T(k + "sw") for k in k0]), reg)
    served = plugin.stats.get("FilterOptim.run", 0)
    opt.run(int(g[k0[0] + "num_iter"]), TensorList([T(k + "new_xf") for k in k0]))
    for b, k in enumerate(k0):
        assert torch.allclose(filt[b], T(k + "hf_out"), rtol=0, atol=1e-6 * float(T(k + "hf_out").abs().max()))
    assert plugin.stats.get("FilterOptim.run", 0) == served


def test_gn_steepest_descent_seam_glue(installed, monkeypatch):
    """GNSteepestDescent.forward over LinearFilterHinge (the SuperDiMPSimple / KeepTrack classifier, dimp_simple.py:685-689): CPU calls
    keep the reference (autograd) implementation; the seam's argument mapping and return structure are checked by putting the oracle
    (explicit g = J^T r, h = J g) behind `ops.gn_sd_hinge` and comparing with the reference module itself."""
    plugin, _ = installed
    from oracle import dimp_oracle as O
    from pytracking import TensorList
    import ltr.models.meta.steepestdescent as sdm
    import ltr.models.target_classifier.residual_modules as rm
    from pytracking_b200 import ops, synth
    g = torch.Generator().manual_seed(4)
    n, c, hw = 5, 32, 18
    feat = synth.make_clf_features(21, n, c, hw, hw)
    label = torch.rand(n, 1, hw + 1, hw + 1, generator=g) * 0.6
    sw = (torch.rand(n, generator=g) + 0.2).view(-1, 1, 1, 1)
    w0 = torch.randn(1, c, 4, 4, generator=g) * 0.05
    for act, thr, leak in (("relu", 0.05, 0.0), ("bentpar", 0.1, 0.1)):
        mod = sdm.GNSteepestDescent(rm.LinearFilterHinge(init_filter_reg=0.1, hinge_threshold=thr, activation_leak=leak, score_act=act,
                                                         act_param=0.7), num_iter=3, residual_batch_dim=1, compute_losses=True,
                                    steplength_reg=0.02).eval()
        served = plugin.stats.get("GNSteepestDescent.forward", 0)
        with torch.no_grad():
            ref_w, ref_its, ref_losses = mod(TensorList([w0.clone()]), num_iter=3, feat=feat, bb=None, train_label=label, sample_weight=sw)
        assert plugin.stats.get("GNSteepestDescent.forward", 0) == served            # CPU tensors: the reference ran

        def fake(weights, f, train_label, sample_weight, num_iter, filter_reg, hinge_threshold, activation_leak, score_act, act_param,
                 steplength_reg, return_iterates=False, compute_losses=False, out=None):
            assert train_label.shape == (n, 1, hw + 1, hw + 1) and sample_weight.shape == (n,) and f.shape == (n, c, hw, hw)
            w, its, losses = O.gn_sd_hinge(weights, f, train_label[:, 0], sample_weight, filter_reg, num_iter, hinge_threshold, activation_leak,
                                           score_act, act_param, steplength_reg, compute_losses)
            return w, torch.cat(its, 0), torch.stack(losses) if compute_losses else None

        with monkeypatch.context() as mp:
            mp.setattr(ops, "gn_sd_hinge", fake)
            mp.setattr(plugin, "_inference", lambda *ts: all(isinstance(t, torch.Tensor) and t.dtype == torch.float32 for t in ts))
            mp.setattr(torch.Tensor, "is_cuda", property(lambda self: True))
            with torch.no_grad():
                w, its, losses = mod(TensorList([w0.clone()]), num_iter=3, feat=feat, bb=None, train_label=label, sample_weight=sw)
                w_t, its_t, _ = mod(w0.clone(), num_iter=2, feat=feat, train_label=label, sample_weight=sw)       # a bare tensor in, a tensor out
        assert plugin.stats.get("GNSteepestDescent.forward", 0) == served + 2
        assert isinstance(w, TensorList) and len(its) == 4 and isinstance(its[0], TensorList) and its[0][0].shape == w0.shape and len(losses) == 4
        assert isinstance(w_t, torch.Tensor) and len(its_t) == 3 and isinstance(its_t[0], torch.Tensor)
        assert torch.allclose(w[0], ref_w[0], rtol=0, atol=2e-5 * float(ref_w[0].abs().max()))
        for a, b in zip(its, ref_its):
            assert torch.allclose(a[0], b[0], rtol=0, atol=2e-5 * float(ref_w[0].abs().max()))
        assert torch.allclose(torch.stack(list(losses)), torch.stack([l.detach() for l in ref_losses]), rtol=2e-4)
