"""-m gpu parity tests of the ToMP row beyond `Transformer.forward` (tests/test_gpu_parity.py): the head feature
extractor on the network plan, the `DenseBoxRegressor` tower on the engine, and the UNMODIFIED reference ToMP-101 tracker
(pytracking/tracker/tomp/tomp.py, baseline/_ref) above the engine against the same tracker on stock PyTorch-CUDA (TF32 off)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _ref():
    from baseline import ref_env
    if not ref_env.reference_available():
        pytest.skip("reference tree not staged (baseline/_ref)")
    from oracle import ref_shims
    ref_shims.install()
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False


def _rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max() / (b.abs().max() + 1e-30))


def test_box_regression_tower_matches_reference_module():
    _ref()
    import ltr.models.transformer.heads as heads
    from pytracking_b200.transformer_engine import BoxTower
    torch.manual_seed(3)
    m = heads.DenseBoxRegressor(num_channels=256).eval()
    g = torch.Generator().manual_seed(4)
    with torch.no_grad():
        for p in m.parameters():                       # non-trivial GroupNorm affines / biases
            if p.dim() == 1:
                p.copy_(0.5 + torch.rand(p.shape, generator=g) if p.numel() == 256 and p.mean() > 0.5 else 0.1 * torch.randn(p.shape, generator=g))
        m.bbreg_layer.weight.mul_(0.3)
    feat = torch.randn(1, 1, 256, 18, 18, generator=g) * 0.3
    filt = torch.randn(1, 256, 1, 1, generator=g) * 0.2
    with torch.no_grad():
        ref_cpu = m(feat, filt)                                               # [1,1,4,18,18]
        mc = m.cuda()
        ref_cuda = mc(feat.cuda(), filt.cuda())
        fp = mc.linear(filt.cuda().reshape(-1, 256)).reshape(1, 256, 1, 1)
        att = (feat.cuda()[0] * fp).sum(1)                                     # the 1x1 correlation
    tw = BoxTower(m.state_dict(), 18, 18, max_batch=2)
    out = tw.forward(feat.cuda()[0], att)
    assert out.shape == (1, 4, 18, 18)
    assert _rel(out, ref_cpu[0]) < 1e-4 and _rel(out, ref_cuda[0]) < 1e-4, (_rel(out, ref_cpu[0]), _rel(out, ref_cuda[0]))
    assert abs(tw.flops / 2 / 1e6 - (4 * 191.1 + 191.1 * 4 / 256)) < 2.0      # 9 * 256 * 256 * 324 = 191.1 MMAC per 256 -> 256 layer at 18x18
    two = tw.forward(torch.cat([feat.cuda()[0], feat.cuda()[0]]), torch.cat([att, att]))
    # the two samples of one launch are bitwise identical; a batch-1 launch may pick another split-K factor (different summation order)
    assert torch.equal(two[0], two[1]) and _rel(two[0], out[0]) < 1e-5
    tw.close()


def test_reference_tomp_tracker_above_engine():
    """ToMP-101 (parameter/tomp/tomp101.py, random-init tompnet101): every frame's backbone + head feature, transformer, classifier
    correlation and box-regression tower are served by the library; boxes (floating-point ltrb regression outputs) agree with the stock
    PyTorch-CUDA run to 1e-2 px while the trajectories coincide."""
    _ref()
    from baseline import ref_tracker
    from pytracking_b200 import plugin, synth
    n = 12
    frames, bb = synth.make_sequence(0, num_frames=n)
    plugin.install()
    plugin.stats.clear()
    try:
        eng = ref_tracker.run_sequence(ref_tracker.build_tomp("cuda"), frames, bb)
        stats = dict(plugin.stats)
    finally:
        plugin.uninstall()
    stock = ref_tracker.run_sequence(ref_tracker.build_tomp("cuda"), frames, bb)
    for seam in ("extract_backbone", "Head.extract_head_feat", "FilterPredictor.predict_cls_bbreg_filters_parallel", "Transformer.forward",
                 "apply_filter", "DenseBoxRegressor.forward", "max2d"):
        assert stats.get(seam, 0) >= n, (seam, stats)
    d = np.abs(eng["target_bbox"] - stock["target_bbox"]).max(axis=1)
    print("ToMP-101 above the engine vs stock PyTorch-CUDA, max abs box difference per frame [px]:", np.round(d, 5).tolist())
    assert d[:3].max() < 1e-2, d


def test_token_assembly_matches_reference_filter_predictor():
    """`b200trk_tomp_tokens` against the reference's own token construction (filter_predictor.py:92-135), intercepted at the
    transformer call, for 2 training frames + 1 test frame; with and without the test-frame embedding."""
    _ref()
    import ltr.models.transformer.filter_predictor as fp
    from pytracking_b200.transformer_engine import TokenBuilder
    for use_test in (False, True):
        captured = {}

        class FakeTransformer(torch.nn.Module):
            d_model = 256

            def forward(self, feat, mask=None, query_embed=None, pos_embed=None):
                captured["feat"], captured["mask"], captured["pos"] = feat.clone(), mask, pos_embed
                return torch.zeros(1, 1, feat.shape[1], 256), torch.zeros_like(feat)
        torch.manual_seed(7)
        m = fp.FilterPredictor(FakeTransformer(), feature_sz=18, use_test_frame_encoding=use_test).eval()
        g = torch.Generator().manual_seed(8)
        with torch.no_grad():
            for mod in m.box_encoding:
                if isinstance(mod, torch.nn.BatchNorm1d):
                    mod.running_mean.copy_(0.1 * torch.randn(mod.running_mean.shape, generator=g))
                    mod.running_var.copy_(0.6 + 0.8 * torch.rand(mod.running_var.shape, generator=g))
                    mod.weight.copy_(0.8 + 0.4 * torch.rand(mod.weight.shape, generator=g))
                    mod.bias.copy_(0.1 * torch.randn(mod.bias.shape, generator=g))
        train_feat = torch.randn(2, 1, 256, 18, 18, generator=g)
        test_feat = torch.randn(1, 1, 256, 18, 18, generator=g)
        label = torch.rand(2, 1, 18, 18, generator=g)
        ltrb = torch.rand(2, 1, 4, 18, 18, generator=g)
        with torch.no_grad():
            m.predict_cls_bbreg_filters_parallel(train_feat, test_feat, label, 1, ltrb)
        ref = captured["feat"]                                                      # [972, 2, 256]
        tb = TokenBuilder(m.state_dict())
        out = tb.build(train_feat[:, 0].cuda(), test_feat[:, 0].cuda(), label[:, 0].cuda(), ltrb[:, 0].cuda(), B=2, use_test_token=use_test)
        assert out.shape == ref.shape
        assert _rel(out, ref) < 1e-5, _rel(out, ref)
        assert captured["mask"].shape == (2, 972) and int(captured["mask"][1].sum()) == 324 and int(captured["mask"][0].sum()) == 0
