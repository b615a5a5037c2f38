"""-m gpu test of the `GNSteepestDescent.forward` seam (SuperDiMPSimple / KeepTrack classifiers; SURVEY 8 row S3.4 + the "also" of f4):
the UNMODIFIED reference module `GNSteepestDescent(LinearFilterHinge)` (ltr/models/meta/steepestdescent.py:32-105, autograd) on stock
PyTorch-CUDA against the same module with `plugin.install()` (the kernel behind `b200trk_gn_sd_hinge`, itself pinned to reference goldens in
tests/test_gpu_parity.py::test_gn_sd_hinge_golden)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max() / (b.abs().max() + 1e-30))


@pytest.mark.parametrize("act,thr,leak,n,c,hw", [("relu", 0.05, 0.0, 50, 512, 18), ("bentpar", 0.1, 0.1, 15, 256, 22)])
def test_reference_gn_steepest_descent_above_the_engine(act, thr, leak, n, c, hw):
    from baseline import ref_env
    if not ref_env.reference_available():
        pytest.skip("reference tree not staged (baseline/_ref)")
    from oracle import ref_shims
    ref_shims.install()
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    import ltr.models.meta.steepestdescent as sdm
    import ltr.models.target_classifier.residual_modules as rm
    from pytracking import TensorList
    from pytracking_b200 import plugin, synth
    g = torch.Generator().manual_seed(4)
    feat = synth.make_clf_features(21, n, c, hw, hw).cuda()
    label = (torch.rand(n, 1, hw + 1, hw + 1, generator=g) * 0.6).cuda()
    sw = (torch.rand(n, generator=g) + 0.2).view(-1, 1, 1, 1).cuda()
    w0 = (torch.randn(1, c, 4, 4, generator=g) * 0.05).cuda()
    mod = sdm.GNSteepestDescent(rm.LinearFilterHinge(init_filter_reg=0.1, hinge_threshold=thr, activation_leak=leak, score_act=act, act_param=0.7),
                                num_iter=5, residual_batch_dim=1, compute_losses=True, steplength_reg=0.02).cuda().eval()
    with torch.no_grad():
        ref_w, ref_its, ref_losses = mod(TensorList([w0.clone()]), num_iter=5, feat=feat, bb=None, train_label=label, sample_weight=sw)
    plugin.install()
    try:
        before = plugin.stats.get("GNSteepestDescent.forward", 0)
        with torch.no_grad():
            w, its, losses = mod(TensorList([w0.clone()]), num_iter=5, feat=feat, bb=None, train_label=label, sample_weight=sw)
        assert plugin.stats.get("GNSteepestDescent.forward", 0) == before + 1
    finally:
        plugin.uninstall()
    assert isinstance(w, TensorList) and len(its) == 6 and len(losses) == 6
    assert _rel(w[0], ref_w[0]) < 2e-4, _rel(w[0], ref_w[0])
    for a, b in zip(its, ref_its):
        assert _rel(a[0], b[0]) < 2e-4
    assert torch.allclose(torch.stack(list(losses)).cpu(), torch.stack([l.detach() for l in ref_losses]).cpu(), rtol=1e-3)


NOT_GNSD = ("apply_filter", "apply_feat_transpose", "max2d", "extract_backbone", "extract_classification_feat", "get_iou_feat", "predict_iou",
            "_prroi_pooling", "_import_prroi_pooling", "conv2d", "conv1x1", "extract_head_feat", "predict_cls_bbreg_filters_parallel", "run",
            "softmax_reg", "DiMPSteepestDescentGN.forward", "PrDiMPSteepestDescentNewton.forward", "DiMPL2SteepestDescentGN.forward",
            "Transformer.forward", "DenseBoxRegressor.forward", "fourier", "preprocess_sample")


def test_reference_dimp_simple_tracker_above_the_engine():
    """The unmodified reference DiMPSimple tracker (SuperDiMPSimple parameters, random-init dimpnet50_simple, no IoUNet) on stock
    PyTorch-CUDA vs the same tracker with its classifier's GNSteepestDescent(LinearFilterHinge) module on `b200trk_gn_sd_hinge` (only
    that seam bound, so the comparison isolates it); CPU counterpart with the oracle behind the op: tests/test_eco_tracker_cpu.py."""
    from baseline import ref_env
    if not ref_env.reference_available():
        pytest.skip("reference tree not staged (baseline/_ref)")
    from oracle import ref_shims
    ref_shims.install()
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    from baseline import ref_tracker
    from pytracking_b200 import plugin, synth
    n_frames = 8
    frames, bb = synth.make_sequence(0, num_frames=n_frames)
    ov = dict(train_skipping=2, use_iou_net=False)

    def drive():
        trk = ref_tracker.build_dimp_simple(device="cuda", overrides=ov)
        torch.manual_seed(0)
        trk.initialize(frames[0], {"init_bbox": list(bb)})
        w_init = trk.target_filter.clone()
        return [trk.track(frames[i])["target_bbox"] for i in range(1, n_frames + 1)], w_init

    ref_boxes, ref_w = drive()
    plugin.install(skip=NOT_GNSD)
    try:
        before = plugin.stats.get("GNSteepestDescent.forward", 0)
        boxes, w = drive()
        assert plugin.stats.get("GNSteepestDescent.forward", 0) >= before + 1 + n_frames // 2 - 1
    finally:
        plugin.uninstall()
    # the first-frame filter (net_opt_iter = 10 iterations from the initialiser's output) is compared directly, the closed loop through
    # the boxes of the first frames only: a non-smooth update (LeakyReluPar) amplifies rounding differences over time
    assert _rel(w, ref_w) < 2e-4, _rel(w, ref_w)
    for a, b in zip(ref_boxes[:2], boxes[:2]):
        assert max(abs(x - y) for x, y in zip(a, b)) < 1e-2, (a, b)
