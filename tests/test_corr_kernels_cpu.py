"""The stage-2 kernels of csrc/corr_kernels.cuh + corr.cuh (SURVEY 8 rows S2.1 `apply_filter` / `apply_feat_transpose`, S2.2
`operation.conv2d(mode='same')`, S2.4 `dcf.max2d`) executed ON THE CPU: the same source files the CUDA build compiles (`cuobjdump -sass`
identical before and after the kernels moved into the header), built as host code under tests/cpu_emul/cuda_shim.h with the launch
arithmetic of csrc/corr_api.cu -- channel chunks, sample groups, the last-CTA reduction over the self-resetting arrival counters.
CPU-tier counterparts of tests/test_gpu_parity.py::test_apply_filter_golden / _oracle / test_max2d_ties: goldens from the unmodified
reference (`ltr/models/layers/filter.py`, `dcf.max2d`)."""
import ctypes as C
import os
import shutil
import subprocess

import numpy as np
import pytest
import torch

from pytracking_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="module")
def emul(tmp_path_factory):
    if shutil.which("g++") is None:
        pytest.skip("no g++")
    out = os.path.join(str(tmp_path_factory.mktemp("corr_emul")), "libcorr_emul.so")
    subprocess.run(["g++", "-std=c++17", "-O1", "-g", "-pthread", "-shared", "-fPIC", "-ffp-contract=off", "-fno-gnu-unique", "-Wno-unknown-pragmas",
                    os.path.join(ROOT, "tests", "cpu_emul", "corr_emul.cpp"), "-o", out], check=True, capture_output=True)
    return C.CDLL(out)


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def _rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def _apply(emul, feat, w, crop=0, want_max=True):
    n, c, h, _ = feat.shape
    o = h if crop else h + 1
    s = np.full((n, 1, o, o), np.nan, np.float32)
    mv, mi = np.full(n, np.nan, np.float32), np.full((n, 2), -1, np.int64)
    rc = emul.corr_emul_apply_filter(_p(np.ascontiguousarray(feat.numpy())), _p(np.ascontiguousarray(w.numpy())), _p(s), n, c, h, h,
                                     _p(mv) if want_max else None, _p(mi) if want_max else None, crop)
    assert rc == 0, rc
    return s, mv, mi


def _transpose(emul, feat, r, sms=148):
    n, c, h, _ = feat.shape
    g = np.full((1, c, 4, 4), np.nan, np.float32)
    assert emul.corr_emul_feat_transpose(_p(np.ascontiguousarray(feat.numpy())), _p(np.ascontiguousarray(r.numpy())), _p(g), n, c, h, h, sms) == 0
    return g


@pytest.mark.parametrize("tag,n,c,h", [("a", 3, 32, 18), ("b", 2, 64, 22)])
def test_apply_filter_kernel_sources_vs_reference_golden(emul, tag, n, c, h):
    g = np.load(os.path.join(GOLDEN, "corr.npz"))
    feat = synth.make_clf_features(100 + ord(tag), n, c, h, h)
    s, mv, mi = _apply(emul, feat, torch.from_numpy(g[tag + "_w"]))
    assert _rel(s, g[tag + "_scores"].reshape(s.shape)) < 1e-5
    assert np.array_equal(mi, g[tag + "_maxidx"]) and _rel(mv, g[tag + "_maxval"]) < 1e-5
    gt = _transpose(emul, feat, torch.from_numpy(g[tag + "_r"])[:, 0:1].contiguous())
    assert _rel(gt, g[tag + "_grad"]) < 1e-5


@pytest.mark.parametrize("n,c,h", [(1, 512, 18), (5, 128, 18), (17, 64, 18), (3, 128, 22), (20, 32, 22)])
def test_apply_filter_kernel_sources_vs_oracle(emul, n, c, h):
    """Several channel chunks and sample groups (the tracker's classify call is n = 1, C = 512: 32 CTAs reduced by the last one), the 'same'
    crop of operation.conv2d, the arg-max of the kernel's own map by the reference rule, adjointness <A w, r> = <w, A^T r>."""
    from oracle import atom_oracle as A
    from oracle import dimp_oracle as O
    feat = synth.make_clf_features(7 * n + c, n, c, h, h)
    w = torch.randn(1, c, 4, 4, generator=torch.Generator().manual_seed(n)) * 0.3
    s, mv, mi = _apply(emul, feat, w)
    assert _rel(s, O.apply_filter(feat, w).numpy()) < 1e-5
    mv_ref, mi_ref = O.max2d(torch.from_numpy(s)[:, 0])
    assert np.array_equal(mi, mi_ref.numpy()) and np.array_equal(mv, mv_ref.numpy())
    sc, _, _ = _apply(emul, feat, w, crop=1, want_max=False)
    assert _rel(sc, A.conv_same(feat, w).numpy()) < 1e-5
    r = torch.randn(n, 1, h + 1, h + 1, generator=torch.Generator().manual_seed(n + 1))
    for sms in (148, 8):                                           # few "SMs": several samples per group
        g = _transpose(emul, feat, r, sms)
        assert _rel(g, O.apply_feat_transpose(feat, r, 4).numpy()) < 1e-5
    lhs = float((s.astype(np.float64) * r.double().numpy()).sum())
    rhs = float((w.double().numpy() * g.astype(np.float64)).sum())
    assert abs(lhs - rhs) <= 1e-4 * max(1.0, abs(lhs))


def test_max2d_kernel_source_ties(emul):
    a = torch.zeros(3, 19, 19)
    a[0, 3, 1] = 1.0; a[0, 1, 3] = 1.0
    a[1, 4, 2] = 2.0; a[1, 2, 2] = 2.0
    a[2] = -1.0; a[2, 18, 18] = -0.5
    mv, mi = np.full(3, np.nan, np.float32), np.full((3, 2), -1, np.int64)
    assert emul.corr_emul_max2d(_p(np.ascontiguousarray(a.numpy())), 3, 19, 19, _p(mv), _p(mi)) == 0
    assert mi.tolist() == [[3, 1], [2, 2], [18, 18]] and mv.tolist() == [1.0, 2.0, -0.5]
