"""The CUDA-core kernels of the ToMP head executed ON THE CPU: token assembly (csrc/tomp_tokens_kernels.cuh:
`FilterPredictor.predict_cls_bbreg_filters_parallel` up to the transformer call, ltr/models/transformer/filter_predictor.py:92-135) and
the non-convolution kernels of the `DenseBoxRegressor` tower (csrc/tower_kernels.cuh: ltr/models/transformer/heads.py:101-141) -- the same
source files the CUDA build compiles (`cuobjdump -sass` identical before and after they moved into headers), built as host code under
tests/cpu_emul/cuda_shim.h, against plain PyTorch (float64) restatements of the reference modules' arithmetic."""
import ctypes as C
import os
import shutil
import subprocess

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def emul(tmp_path_factory):
    if shutil.which("g++") is None:
        pytest.skip("no g++")
    out = os.path.join(str(tmp_path_factory.mktemp("tomp_emul")), "libtomp_emul.so")
    subprocess.run(["g++", "-std=c++17", "-O1", "-g", "-pthread", "-shared", "-fPIC", "-ffp-contract=off", "-fno-gnu-unique", "-Wno-unknown-pragmas",
                    os.path.join(ROOT, "tests", "cpu_emul", "tomp_emul.cpp"), "-o", out], check=True, capture_output=True)
    return C.CDLL(out)


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def _rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def _np(t):
    return np.ascontiguousarray(t.numpy())


@pytest.mark.parametrize("D,D1,n_train,n_test,hw,B,use_test_token", [(256, 64, 2, 1, 6, 2, False), (64, 16, 1, 2, 5, 1, True)])
def test_token_assembly_kernel_source(emul, D, D1, n_train, n_test, hw, B, use_test_token):
    """train token = (feat + fg_token * label) + MLP(ltrb) with the two BatchNorms already folded into the 1x1 convolutions (the host's job),
    test token = feat (+ test token); every token written once per batch entry (filter_predictor.py:104-135)."""
    g = torch.Generator().manual_seed(D)
    HW = hw * hw
    trf, tef = torch.randn(n_train, D, hw, hw, generator=g), torch.randn(n_test, D, hw, hw, generator=g)
    label, ltrb = torch.rand(n_train, hw, hw, generator=g), torch.rand(n_train, 4, hw, hw, generator=g)
    fg, tt = torch.randn(D, generator=g), torch.randn(D, generator=g)
    w1, b1 = torch.randn(D1, 4, generator=g), torch.randn(D1, generator=g)
    w2, b2 = torch.randn(D, D1, generator=g) / D1 ** 0.5, torch.randn(D, generator=g)
    w3, b3 = torch.randn(D, D, generator=g) / D ** 0.5, torch.randn(D, generator=g)
    out = np.full(((n_train + n_test) * HW, B, D), np.nan, np.float32)
    rc = emul.tomp_emul_tokens(_p(_np(trf)), _p(_np(tef)), _p(_np(label)), _p(_np(ltrb)), _p(_np(fg)), _p(_np(tt)) if use_test_token else None, _p(_np(w1)),
                               _p(_np(b1)), _p(_np(w2.t())), _p(_np(b2)), _p(_np(w3.t())), _p(_np(b3)), _p(out), n_train, n_test, hw, hw, D, D1, B)
    assert rc == 0
    x = ltrb.double().reshape(n_train, 4, HW)
    h1 = torch.relu(torch.einsum("jk,fkp->fjp", w1.double(), x) + b1.double().view(1, -1, 1))
    h2 = torch.relu(torch.einsum("cj,fjp->fcp", w2.double(), h1) + b2.double().view(1, -1, 1))
    enc = torch.einsum("ck,fkp->fcp", w3.double(), h2) + b3.double().view(1, -1, 1)
    tr = (trf.double().reshape(n_train, D, HW) + fg.double().view(1, -1, 1) * label.double().reshape(n_train, 1, HW)) + enc
    te = tef.double().reshape(n_test, D, HW) + (tt.double().view(1, -1, 1) if use_test_token else 0.0)
    ref = torch.cat([tr.permute(0, 2, 1).reshape(-1, D), te.permute(0, 2, 1).reshape(-1, D)], 0)          # tokens: frame-major, then cell
    for b in range(B):
        assert _rel(out[:, b], ref.numpy()) < 2e-6


def test_box_tower_kernel_sources(emul):
    g = torch.Generator().manual_seed(2)
    S, Cc, hw = 2, 40, 7
    HW = hw * hw
    feat, att = torch.randn(S, Cc, hw, hw, generator=g), torch.rand(S, hw, hw, generator=g)
    out = np.full((S, HW, Cc), np.nan, np.float32)
    assert emul.tomp_emul_import_scaled(_p(_np(feat)), _p(_np(att)), _p(out), S, HW, Cc) == 0          # feats_att = attention * feat, NCHW -> NHWC
    assert np.array_equal(out, _np((feat * att.unsqueeze(1)).permute(0, 2, 3, 1).reshape(S, HW, Cc)))
    assert emul.tomp_emul_import_scaled(_p(_np(feat)), None, _p(out), S, HW, Cc) == 0
    assert np.array_equal(out, _np(feat.permute(0, 2, 3, 1).reshape(S, HW, Cc)))
    x = torch.randn(S, Cc, hw, hw, generator=g) * 2 + 0.5
    gam, bet = torch.rand(Cc, generator=g) + 0.5, torch.randn(Cc, generator=g)
    xn = _np(x.permute(0, 2, 3, 1).reshape(S, HW, Cc)).copy()
    assert emul.tomp_emul_groupnorm1_relu(_p(xn), _p(_np(gam)), _p(_np(bet)), S, HW, Cc) == 0            # heads.py:8-15: GroupNorm(1, C) + ReLU
    ref = torch.relu(torch.nn.functional.group_norm(x.double(), 1, gam.double(), bet.double(), 1e-5))
    assert _rel(xn, _np(ref.permute(0, 2, 3, 1).reshape(S, HW, Cc))) < 2e-6
    y = torch.randn(S, HW, 4, generator=g)
    e = np.full((S, 4, HW), np.nan, np.float32)
    assert emul.tomp_emul_export_exp(_p(_np(y)), _p(e), S, HW, 4) == 0                                   # heads.py:139: ltrb = exp(tower output)
    assert _rel(e, torch.exp(y.double()).permute(0, 2, 1).numpy()) < 1e-6
