"""The persistent cooperative optimiser kernels on the cp.async sweeps of csrc/corr2.cuh executed ON THE CPU (the same source files the CUDA
build compiles; `cuobjdump -sass` identical before and after they moved into headers and the inline cp.async / globaltimer statements got a
host form): `atom_cg_kernel` (ATOM's per-frame ConjugateGradient.run on ConvProblem).  ONE CTA per launch: the kernels keep
small arrays in static __shared__ storage, which the shim can give to a single live block only, so the decomposition is the one the launcher
picks on a 1-SM device (all channels, <= 8 samples in one CTA); the cross-CTA exchange stays with the `-m gpu` tests."""
import ctypes as C
import os
import shutil
import subprocess

import numpy as np
import pytest
import torch

from pytracking_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ACT = {"none": 0, "relu": 1, "elu": 2, "mlu": 3}


@pytest.fixture(scope="module")
def emul(tmp_path_factory):
    if shutil.which("g++") is None:
        pytest.skip("no g++")
    out = os.path.join(str(tmp_path_factory.mktemp("cg_emul")), "libcg_emul.so")
    subprocess.run(["g++", "-std=c++17", "-O1", "-g", "-pthread", "-shared", "-fPIC", "-ffp-contract=off", "-Wno-unknown-pragmas",
                    os.path.join(ROOT, "tests", "cpu_emul", "cg_emul.cpp"), "-o", out], check=True, capture_output=True)
    return C.CDLL(out)


def _rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


@pytest.mark.parametrize("n,c,h,it,fr,act", [(6, 32, 18, 4, False, "mlu"), (8, 64, 18, 3, True, "relu"), (5, 16, 22, 5, False, "none")])
def test_atom_cg_kernel_source_vs_oracle(emul, n, c, h, it, fr, act):
    from oracle import atom_oracle as A
    x, y, sw = synth.make_atom_memory(31 + n, n, c, h, h)
    w = torch.randn(1, c, 4, 4, generator=torch.Generator().manual_seed(n)) * 0.02
    out = np.full((1, c, 4, 4), np.nan, np.float32)
    xs, ys, sws, ws = [np.ascontiguousarray(t.numpy()) for t in (x, y, sw, w)]
    rc = emul.cg_emul_atom_cg_filter(ws.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p), xs.ctypes.data_as(C.c_void_p),
                                     ys.ctypes.data_as(C.c_void_p), sws.ctypes.data_as(C.c_void_p), n, c, h, h, it, C.c_float(0.1), int(fr), ACT[act], C.c_float(0.05))
    assert rc == 0
    ref64 = A.atom_cg_filter(w.double(), x.double(), y.double(), sw.double(), 0.1, it, act, 0.05, fr)[0]
    ref32 = A.atom_cg_filter(w, x, y, sw, 0.1, it, act, 0.05, fr)[0]
    tol = max(20 * _rel(ref32.numpy(), ref64.numpy()), 2e-5)           # float32 CG against the float64 solution, calibrated by the oracle's own spread
    assert _rel(out, ref64.numpy()) < tol, (_rel(out, ref64.numpy()), tol)


# ---- sd_kernel: the DiMP-family online optimisers against the UNMODIFIED reference modules' golden outputs ------------
@pytest.fixture(scope="module")
def sd(tmp_path_factory):
    if shutil.which("g++") is None:
        pytest.skip("no g++")
    out = os.path.join(str(tmp_path_factory.mktemp("sd_emul")), "libsd_emul.so")
    subprocess.run(["g++", "-std=c++17", "-O1", "-g", "-pthread", "-shared", "-fPIC", "-ffp-contract=off", "-Wno-unknown-pragmas",
                    os.path.join(ROOT, "tests", "cpu_emul", "sd_emul.cpp"), "-o", out], check=True, capture_output=True)
    return C.CDLL(out)


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def _arr(t):
    return np.ascontiguousarray(t.numpy() if torch.is_tensor(t) else t, dtype=np.float32)


def _check(g, tag, w, its, losses):
    assert _rel(its[1], g[tag + "_w1"][0]) < 1e-4 and _rel(w, g[tag + "_wfinal"]) < 1e-4 and _rel(its[-1], g[tag + "_wfinal"][0]) < 1e-4
    assert np.allclose(losses, g[tag + "_losses"], rtol=1e-4)


def _golden(name):
    return np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))


# (the reference golden "n7_22" -- 7 samples of 128 channels at 22x22 -- does not fit one CTA's shared memory: -m gpu only)
@pytest.mark.parametrize("tag,n,c,h,it,use_sw,seed", [("n4_c64", 4, 64, 18, 3, False, 23)])
def test_dimp_sd_kernel_source_vs_reference_golden(sd, tag, n, c, h, it, use_sw, seed):
    """DiMPSteepestDescentGN.forward (ltr/models/target_classifier/optimizer.py:85-170): goldens of the reference module (n <= 8 fits one CTA)."""
    g = _golden("dimp_sd")
    p = synth.make_dimp_optimizer_params(seed=seed)
    feat, bb = _arr(synth.make_clf_features(seed, n, c, h, h)), _arr(synth.make_boxes(seed + 1, n, center=(h * 16) / 2 - 25))
    sw = _arr(g[tag + "_sw"]) if use_sw else None
    luts = [_arr(p[k].reshape(-1)) for k in ("label_map_predictor.weight", "target_mask_predictor.0.weight", "spatial_weight_predictor.weight")]
    w0, w = _arr(g[tag + "_w0"]), np.full((1, c, 4, 4), np.nan, np.float32)
    its, losses = np.full((it + 1, c, 4, 4), np.nan, np.float32), np.full(it + 1, np.nan, np.float32)
    rc = sd.sd_emul_dimp_sd_gn(_ptr(w0), _ptr(w), _ptr(feat), _ptr(bb), _ptr(sw), n, c, h, h, it, _ptr(luts[0]), _ptr(luts[1]), _ptr(luts[2]), luts[0].size,
                               C.c_float(0.1), C.c_float(16.0), C.c_float(float(torch.exp(p["log_step_length"]))),
                               C.c_float(max(float(p["filter_reg"]) ** 2, 1e-3 ** 2)), C.c_float(0.0), _ptr(its), _ptr(losses))
    assert rc == 0
    _check(g, tag, w, its, losses)


def test_prdimp_sd_kernel_source_vs_reference_golden(sd):
    """PrDiMPSteepestDescentNewton.forward (optimizer.py:355-439) with softmax_reg, label threshold, normalisation and shrink."""
    g, tag, n, c, h, it, seed = _golden("prdimp_sd"), "n6_18", 6, 64, 18, 3, 32
    feat, bb = _arr(synth.make_clf_features(seed, n, c, h, h)), _arr(synth.make_boxes(seed + 1, n, center=(h * 16) / 2 - 25))
    w0, w, sw = _arr(g[tag + "_w0"]), np.full((1, c, 4, 4), np.nan, np.float32), _arr(g[tag + "_sw"])
    its, losses = np.full((it + 1, c, 4, 4), np.nan, np.float32), np.full(it + 1, np.nan, np.float32)
    rc = sd.sd_emul_prdimp_sd_newton(_ptr(w0), _ptr(w), _ptr(feat), _ptr(bb), _ptr(sw), n, c, h, h, it, C.c_float(float(g[tag + "_sigma"])), C.c_float(16.0),
                                     C.c_float(1.0), C.c_float(0.05 ** 2), C.c_float(0.05), 1, C.c_float(-2.0), C.c_float(0.05), 1, C.c_float(0.1),
                                     C.c_float(0.0), _ptr(its), _ptr(losses))
    assert rc == 0
    _check(g, tag, w, its, losses)


@pytest.mark.parametrize("tag,n,c,h,it,use_sw,thr,seed", [("n8_c64", 8, 64, 18, 4, True, 0.05, 71), ("n5_c32_22", 5, 32, 22, 3, False, -999.0, 72)])
def test_dimp_l2_sd_kernel_source_vs_reference_golden(sd, tag, n, c, h, it, use_sw, thr, seed):
    """DiMPL2SteepestDescentGN.forward (optimizer.py:211-291)."""
    g = _golden("dimp_l2_sd")
    feat, bb = _arr(synth.make_clf_features(seed, n, c, h, h)), _arr(synth.make_boxes(seed + 1, n, center=(h * 16) / 2 - 25))
    sw = _arr(g[tag + "_sw"]) if use_sw else None
    w0, w = _arr(g[tag + "_w0"]), np.full((1, c, 4, 4), np.nan, np.float32)
    its, losses = np.full((it + 1, c, 4, 4), np.nan, np.float32), np.full(it + 1, np.nan, np.float32)
    rc = sd.sd_emul_dimp_l2_sd_gn(_ptr(w0), _ptr(w), _ptr(feat), _ptr(bb), _ptr(sw), n, c, h, h, it, C.c_float(1.3), C.c_float(thr), C.c_float(16.0),
                                  C.c_float(0.9), C.c_float(max(0.1 ** 2, 1e-6)), C.c_float(0.01), _ptr(its), _ptr(losses))
    assert rc == 0
    _check(g, tag, w, its, losses)


@pytest.mark.parametrize("tag,n,c,h,it,use_sw,thr,leak,act,seed", [("relu_n6_c64", 6, 64, 18, 4, True, 0.05, 0.0, 0, 81),
                                                                  ("bent_n4_c32_22", 4, 32, 22, 3, False, 0.1, 0.1, 1, 82)])
def test_gn_sd_hinge_kernel_source_vs_reference_golden(sd, tag, n, c, h, it, use_sw, thr, leak, act, seed):
    """GNSteepestDescent.forward over LinearFilterHinge (ltr/models/meta/steepestdescent.py:32-105, residual_modules.py:89-135)."""
    g = _golden("gn_sd_hinge")
    feat = _arr(synth.make_clf_features(seed, n, c, h, h))
    sw = _arr(g[tag + "_sw"]) if use_sw else None
    label = _arr(g[tag + "_label"])
    w0, w = _arr(g[tag + "_w0"]), np.full((1, c, 4, 4), np.nan, np.float32)
    its, losses = np.full((it + 1, c, 4, 4), np.nan, np.float32), np.full(it + 1, np.nan, np.float32)
    rc = sd.sd_emul_gn_sd_hinge(_ptr(w0), _ptr(w), _ptr(feat), _ptr(label), _ptr(sw), n, c, h, h, it, C.c_float(0.1), C.c_float(thr), C.c_float(leak), act,
                                C.c_float(0.7), C.c_float(0.02), _ptr(its), _ptr(losses))
    assert rc == 0
    _check(g, tag, w, its, losses)
