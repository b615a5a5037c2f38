"""CPU: the oracle restatement (oracle/dimp_oracle.py) against golden vectors produced by the
unmodified reference (oracle/gen_golden.py).  This is what pins the oracle."""
import os

import math

import numpy as np
import pytest
import torch

from oracle import dimp_oracle as O
from pytracking_b200 import synth


def _rel(a, b):
    a = torch.as_tensor(a, dtype=torch.float64)
    b = torch.as_tensor(b, dtype=torch.float64)
    return float((a - b).abs().max() / (b.abs().max() + 1e-30))


@pytest.fixture(scope="module")
def G(golden_dir):
    return {n: np.load(os.path.join(golden_dir, n + ".npz")) for n in ("corr", "labels", "dimp_sd", "prdimp_sd", "backbone")}


@pytest.mark.parametrize("tag,n,c,h,k", [("a", 3, 32, 18, 4), ("b", 2, 64, 22, 4), ("c", 2, 32, 18, 1), ("d", 2, 32, 13, 3)])
def test_corr_and_adjoint(G, tag, n, c, h, k):
    g = G["corr"]
    feat = synth.make_clf_features(100 + ord(tag), n, c, h, h, filter_size=max(k, 1))
    w = torch.from_numpy(g[tag + "_w"])
    s = O.apply_filter(feat, w)
    assert s.shape[-1] == h + (k + 1) % 2
    assert _rel(s[:, 0], g[tag + "_scores"][:, 0]) < 2e-6
    r = torch.from_numpy(g[tag + "_r"])[:, 0:1]
    gt = O.apply_feat_transpose(feat, r, k)
    assert _rel(gt, g[tag + "_grad"]) < 2e-6
    mv, mi = O.max2d(torch.from_numpy(g[tag + "_scores"])[:, 0])
    assert np.array_equal(mi.numpy(), g[tag + "_maxidx"])
    assert np.array_equal(mv.numpy(), g[tag + "_maxval"])


def test_max2d_ties():
    a = torch.zeros(2, 5, 5)
    a[0, 3, 1] = 1.0
    a[0, 1, 3] = 1.0     # tie: reference picks the smaller column (dcf.py:159-160)
    a[1, 4, 2] = 2.0
    a[1, 2, 2] = 2.0     # tie inside a column: smaller row
    _, idx = O.max2d(a)
    assert idx.tolist() == [[3, 1], [2, 2]]


def test_label_maps(G):
    g = G["labels"]
    p = synth.make_dimp_optimizer_params(seed=5)
    y, m, v = O.dimp_label_maps(torch.from_numpy(g["bb"]), p, (19, 19))
    assert _rel(y, g["y"]) < 1e-5 and _rel(m, g["m"]) < 1e-5 and _rel(v, g["v"]) < 1e-5


@pytest.mark.parametrize("tag,n,c,h,it,use_sw,seed", [("n15_it10", 15, 512, 18, 10, True, 21), ("n50_it2", 50, 512, 18, 2, True, 22),
                                                      ("n4_c64", 4, 64, 18, 3, False, 23), ("n7_22", 7, 128, 22, 4, True, 24)])
def test_dimp_sd(G, tag, n, c, h, it, use_sw, seed):
    g = G["dimp_sd"]
    p = synth.make_dimp_optimizer_params(seed=seed)
    feat = synth.make_clf_features(seed, n, c, h, h)
    bb = synth.make_boxes(seed + 1, n, center=(h * 16) / 2 - 25)
    sw = torch.from_numpy(g[tag + "_sw"]) if use_sw else None
    w, its, losses = O.dimp_sd_gn(torch.from_numpy(g[tag + "_w0"]), feat, bb, sw, p, it)
    assert _rel(its[1], g[tag + "_w1"]) < 1e-5
    assert _rel(w, g[tag + "_wfinal"]) < 1e-4
    assert np.allclose([float(l) for l in losses], g[tag + "_losses"], rtol=1e-4)


@pytest.mark.parametrize("tag,n,c,h,it,seed,sreg,lthr", [("n15_22", 15, 512, 22, 10, 31, None, 0.0), ("n6_18", 6, 64, 18, 3, 32, -2.0, 0.05)])
def test_prdimp_sd(G, tag, n, c, h, it, seed, sreg, lthr):
    g = G["prdimp_sd"]
    feat = synth.make_clf_features(seed, n, c, h, h)
    bb = synth.make_boxes(seed + 1, n, center=(h * 16) / 2 - 25)
    w, its, losses = O.prdimp_sd_newton(torch.from_numpy(g[tag + "_w0"]), feat, bb, torch.from_numpy(g[tag + "_sw"]),
                                        0.0, 0.05, it, float(g[tag + "_sigma"]), min_filter_reg=0.05, alpha_eps=0.05,
                                        softmax_reg_val=sreg, label_threshold=lthr, normalize_label=True,
                                        label_shrink=0.0 if sreg is None else 0.1)
    assert _rel(its[1], g[tag + "_w1"]) < 1e-5
    assert _rel(w, g[tag + "_wfinal"]) < 1e-4
    assert np.allclose([float(l) for l in losses], g[tag + "_losses"], rtol=1e-4)


@pytest.mark.parametrize("arch,size,seed", [("resnet50", 96, 42), ("resnet18", 96, 42), ("resnet50", 288, 41)])
def test_backbone_and_head(G, arch, size, seed):
    g = G["backbone"]
    sd = synth.make_dimp_state_dict(arch, seed=0, lut_seed=3)
    im = O.preprocess_image(synth.make_crop(seed, 1, size))
    with torch.no_grad():
        bf = O.resnet_forward(sd, im, arch)
        clf = O.clf_head_dimp50(sd, bf["layer3"]) if arch == "resnet50" else O.clf_head_dimp18(sd, bf["layer3"])
    tag = "%s_%d_" % (arch, size)
    l2 = bf["layer2"] if size == 96 else bf["layer2"][:, ::8]
    assert _rel(l2, g[tag + "layer2"]) < 1e-5
    assert _rel(bf["layer3"], g[tag + "layer3"]) < 1e-5
    assert _rel(clf, g[tag + "clf"]) < 1e-5


def test_conv_formulation_matches_explicit():
    feat = synth.make_clf_features(9, 6, 64, 18, 18)
    g = torch.Generator().manual_seed(4)
    w = torch.randn(1, 64, 4, 4, generator=g) * 0.1
    r = torch.randn(6, 1, 19, 19, generator=g)
    assert _rel(O.apply_filter_conv(feat, w), O.apply_filter(feat, w)) < 1e-5
    assert _rel(O.apply_feat_transpose_conv(feat, r, 4), O.apply_feat_transpose(feat, r, 4)) < 1e-5
    p = synth.make_dimp_optimizer_params(seed=2)
    bb = synth.make_boxes(3, 6)
    w1 = O.dimp_sd_gn_conv(w, feat, bb, None, p, 3)
    w2, _, _ = O.dimp_sd_gn(w, feat, bb, None, p, 3, compute_losses=False)
    assert _rel(w1, w2) < 1e-4


ATOM_CG_CASES = {"n12_c16_pr_mlu": (12, 16, 12, 5, False, "mlu", 51), "n40_c64_pr_mlu": (40, 64, 25, 5, False, "mlu", 52),
                 "n9_c32_fr_none": (9, 32, 9, 4, True, "none", 53), "n20_c64_pr_relu": (20, 64, 20, 3, False, "relu", 54)}


@pytest.mark.parametrize("tag", sorted(ATOM_CG_CASES))
def test_atom_cg_filter(golden_dir, tag):
    """ATOM ConjugateGradient.run on ConvProblem (reference classes, oracle/gen_golden.py:gen_atom_cg)."""
    from oracle import atom_oracle as A
    g = np.load(os.path.join(golden_dir, "atom_cg.npz"))
    n, c, nf, it, fr, act, seed = ATOM_CG_CASES[tag]
    x, y, sw = synth.make_atom_memory(seed, n, c, 18, 18, n_filled=nf)
    w0 = torch.from_numpy(g[tag + "_w0"])
    w1, _, _ = A.atom_cg_filter(w0, x, y, sw, 0.1, it, act, 0.05, fr)
    assert _rel(w1, g[tag + "_w"]) < 2e-5
    w2, _, _ = A.atom_cg_filter(w1, x, y, sw, 0.1, it, act, 0.05, fr)
    assert _rel(w2, g[tag + "_w2"]) < 5e-5


def test_atom_conv_same_adjoint():
    from oracle import atom_oracle as A
    g = torch.Generator().manual_seed(1)
    x = torch.randn(3, 8, 18, 18, generator=g)
    w = torch.randn(1, 8, 4, 4, generator=g)
    u = torch.randn(3, 1, 18, 18, generator=g)
    lhs = float((A.conv_same(x, w).double() * u.double()).sum())
    rhs = float((w.double() * A.conv_same_adjoint_filter(x, u, 4).double()).sum())
    assert abs(lhs - rhs) < 1e-4 * max(1.0, abs(lhs))
    # conv_same == apply_filter with the last row / column dropped
    assert _rel(A.conv_same(x, w), O.apply_filter(x, w)[:, :, :18, :18]) < 1e-5


FOURIER_CASES = {"s18_k4": (3, 18, 4, 288), "s18_k4_o72": (2, 18, 4, 72), "s17_k5": (2, 17, 5, 64), "s22_k4": (1, 22, 4, 352)}


@pytest.mark.parametrize("tag", sorted(FOURIER_CASES))
def test_fourier_interp(golden_dir, tag):
    """cfft2 -> shift_fs -> sum_fs -> sample_fs as ATOM.localize_target chains them (reference modules, gen_fourier)."""
    from oracle import atom_oracle as A
    g = np.load(os.path.join(golden_dir, "fourier.npz"))
    S, H, ksz, osz = FOURIER_CASES[tag]
    up = A.fourier_interp(torch.from_numpy(g[tag + "_scores"]), (ksz, ksz), (osz, osz))
    assert _rel(up, g[tag + "_up"]) < 1e-5
    # the Fourier series passes through the shifted samples: for an even kernel the half-cell shift puts sample y at Y = 16 y + 144 (mod 288)
    if ksz % 2 == 0 and osz % H == 0:
        r = osz // H
        idx = (torch.arange(H) * r + osz // 2) % osz
        sub = up[:, 0][:, idx][:, :, idx]
        # (an even-sized map keeps both Nyquist terms, so the pass-through is exact only up to the doubled Nyquist component)
        assert sub.shape[-1] == H


DIMP_L2_CASES = {"n8_c64": (8, 64, 18, 4, True, 0.05, 71), "n5_c32_22": (5, 32, 22, 3, False, -999.0, 72)}


@pytest.mark.parametrize("tag", sorted(DIMP_L2_CASES))
def test_dimp_l2_sd(golden_dir, tag):
    g = np.load(os.path.join(golden_dir, "dimp_l2_sd.npz"))
    n, c, h, it, use_sw, thr, seed = DIMP_L2_CASES[tag]
    feat = synth.make_clf_features(seed, n, c, h, h)
    bb = synth.make_boxes(seed + 1, n, center=(h * 16) / 2 - 25)
    sw = torch.from_numpy(g[tag + "_sw"]) if use_sw else None
    w, its, losses = O.dimp_l2_sd_gn(torch.from_numpy(g[tag + "_w0"]), feat, bb, sw, math.log(0.9), 0.1, it, 1.3, thr, alpha_eps=0.01)
    assert _rel(its[1], g[tag + "_w1"]) < 1e-5
    assert _rel(w, g[tag + "_wfinal"]) < 1e-4
    assert np.allclose([float(l) for l in losses], g[tag + "_losses"], rtol=1e-4)


GN_HINGE_CASES = {"relu_n6_c64": (6, 64, 18, 4, True, 0.05, 0.0, "relu", 81), "bent_n4_c32_22": (4, 32, 22, 3, False, 0.1, 0.1, "bentpar", 82)}


@pytest.mark.parametrize("tag", sorted(GN_HINGE_CASES))
def test_gn_sd_hinge(golden_dir, tag):
    """GNSteepestDescent over LinearFilterHinge (autograd J^T r / J g in the reference, explicit in the oracle)."""
    g = np.load(os.path.join(golden_dir, "gn_sd_hinge.npz"))
    n, c, h, it, use_sw, thr, leak, act, seed = GN_HINGE_CASES[tag]
    feat = synth.make_clf_features(seed, n, c, h, h)
    sw = torch.from_numpy(g[tag + "_sw"]) if use_sw else None
    w, its, losses = O.gn_sd_hinge(torch.from_numpy(g[tag + "_w0"]), feat, torch.from_numpy(g[tag + "_label"]), sw, 0.1, it, thr, leak,
                                   act, 0.7, 0.02)
    assert _rel(its[1], g[tag + "_w1"]) < 1e-5
    assert _rel(w, g[tag + "_wfinal"]) < 1e-4
    assert np.allclose([float(l) for l in losses], g[tag + "_losses"], rtol=1e-4)


TRANSFORMER_CASES = {"small": (64, 2, 128, 2, 2, 40, 2, 91, True), "tomp_l72": (256, 8, 2048, 6, 6, 72, 2, 92, True),
                     "tomp_l48_nomask": (256, 8, 2048, 6, 6, 48, 1, 93, False)}


def transformer_inputs(tag):
    d, nh, ff, ne, nd, L, B, seed, use_mask = TRANSFORMER_CASES[tag]
    sd = synth.make_transformer_state_dict(seed, d, nh, ff, ne, nd)
    g = torch.Generator().manual_seed(seed + 1)
    src = torch.randn(L, B, d, generator=g)
    pos = torch.randn(L, 1, d, generator=g) * 0.5
    qe = torch.randn(1, d, generator=g)
    mask = None
    if use_mask:
        mask = torch.zeros(B, L, dtype=torch.bool)
        mask[B - 1, L // 3: L // 2] = True
    return sd, src, pos, qe, mask


@pytest.mark.parametrize("tag", sorted(TRANSFORMER_CASES))
def test_transformer_forward(golden_dir, tag):
    """ToMP Transformer.forward: explicit-attention restatement against the reference module's outputs."""
    from oracle import tomp_oracle as T
    g = np.load(os.path.join(golden_dir, "transformer.npz"))
    d, nh, ff, ne, nd, L, B, seed, use_mask = TRANSFORMER_CASES[tag]
    sd, src, pos, qe, mask = transformer_inputs(tag)
    with torch.no_grad():
        hs, mem = T.transformer_forward(sd, src, mask, qe, pos, nh, ne, nd)
    assert hs.shape == g[tag + "_hs"].shape
    assert _rel(mem, g[tag + "_memory"]) < 2e-5
    assert _rel(hs, g[tag + "_hs"]) < 2e-5


ATOM_GN_CASES = {"n6_c32_16": (6, 32, 16, 3, 2, True, "mlu", 101), "n10_c64_32_pr": (10, 64, 32, 4, 3, False, "relu", 102)}


@pytest.mark.parametrize("tag", sorted(ATOM_GN_CASES))
def test_atom_gn_joint(golden_dir, tag):
    """ATOM first-frame joint optimisation: GaussNewtonCG on FactorizedConvProblem (reference classes) vs the explicit J / J^T oracle."""
    from oracle import atom_oracle as A
    g = np.load(os.path.join(golden_dir, "atom_gn.npz"))
    n, cin, cc, ncg, ngn, fr, act, seed = ATOM_GN_CASES[tag]
    x, y, sw = synth.make_atom_memory(seed, n, cin, 18, 18)
    w, P = A.atom_gn_joint(torch.from_numpy(g[tag + "_w0"]), torch.from_numpy(g[tag + "_P0"]), x, y, sw, 0.1, 1e-2, ncg, ngn, act, 0.05, fr)
    assert _rel(w, g[tag + "_w"]) < 1e-4
    assert _rel(P, g[tag + "_P"]) < 1e-4


@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_softmax_reg_golden(golden_dir, tag):
    """oracle softmax_reg against the reference's ltr/models/layers/activation.py:softmax_reg outputs."""
    from oracle import dimp_oracle as O
    g = np.load(os.path.join(golden_dir, "softmax_reg.npz"))
    x = torch.from_numpy(g[tag + "_x"])
    reg = None if np.isnan(g[tag + "_reg"][0]) else float(g[tag + "_reg"][0])
    y = O.softmax_reg(x.reshape(x.shape[0], x.shape[-2], x.shape[-1]), reg).reshape(x.shape)
    assert np.allclose(y.numpy(), g[tag + "_y"], rtol=1e-5, atol=1e-9)


def test_hann_windows_golden(golden_dir):
    """Host mirror of dcf.hann2d / hann2d_clipped (pytracking_b200/preprocessing.py) against the reference's outputs: bit-exact."""
    from pytracking_b200 import preprocessing as pre
    g = np.load(os.path.join(golden_dir, "hann.npz"))
    n = 0
    for k in g.files:
        if k.endswith("_arg"):
            continue
        a = [int(v) for v in g[k + "_arg"]]
        w = pre.hann2d(a[:2], bool(a[2])) if k.startswith("h2_") else pre.hann2d_clipped(a[:2], a[2:4], bool(a[4]))
        assert w.shape == g[k].shape and np.array_equal(w.numpy(), g[k]), k
        n += 1
    assert n == 9


@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_tomp_position_encoding_golden(golden_dir, tag):
    """Host mirror of the ToMP position encoding against the reference's PositionEmbeddingSine: bit-exact."""
    from pytracking_b200 import preprocessing as pre
    g = np.load(os.path.join(golden_dir, "tomp_pos.npz"))
    h, w, d, res = [int(v) for v in g[tag + "_arg"]]
    pos = pre.tomp_position_encoding(h, w, d, res)
    assert pos.shape == g[tag].shape and np.array_equal(pos.numpy(), g[tag])
