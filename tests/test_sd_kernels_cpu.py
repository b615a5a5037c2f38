"""The persistent cooperative optimiser kernels on the cp.async sweeps of csrc/corr2.cuh executed ON THE CPU (the same source files the CUDA
build compiles; `cuobjdump -sass` identical before and after they moved into headers and the inline cp.async / globaltimer statements got a
host form): `sd_kernel` (DiMP SD-GN, PrDiMP SD-Newton, DiMP-L2, GNSteepestDescent + hinge) against the golden outputs
of the UNMODIFIED reference modules, `atom_cg_kernel` (row S3.5: ATOM's per-frame ConjugateGradient.run on ConvProblem) against the reference
classes' goldens and the float64 oracle.  The kernels keep small arrays in static __shared__ storage, so they run in two shim modes: OS
thread per CUDA thread (one CTA: the decomposition the launcher picks on a 1-SM device; this is the mode tests/test_kernels_tsan_cpu.py puts
under ThreadSanitizer) and block = OS thread with fibers (`launch_coop`, static __shared__ = thread_local): the whole cooperative grid for a
given SM count, cross-CTA exchange through the grid barriers included."""
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
ATOM_CG_FILLED = {"n12_c16_pr_mlu": 12, "n40_c64_pr_mlu": 25, "n9_c32_fr_none": 9, "n20_c64_pr_relu": 20}      # filled memory slots of the golden cases


def _build(tmp_path_factory, name, coop):
    if shutil.which("g++") is None:
        pytest.skip("no g++")
    out = os.path.join(str(tmp_path_factory.mktemp("%s_emul%d" % (name, coop))), "lib%s_emul.so" % name)
    # -fno-gnu-unique: the two builds of a harness define the same inline / function-local symbols with DIFFERENT storage (static vs
    # thread_local __shared__ arrays); as GNU unique symbols they would be merged process-wide when both are loaded
    subprocess.run(["g++", "-std=c++17", "-O1", "-g", "-pthread", "-shared", "-fPIC", "-ffp-contract=off", "-fno-gnu-unique", "-Wno-unknown-pragmas"] +
                   (["-DB200_EMUL_COOP_FIBERS"] if coop else []) + [os.path.join(ROOT, "tests", "cpu_emul", name + "_emul.cpp"), "-o", out],
                   check=True, capture_output=True)
    return C.CDLL(out)


@pytest.fixture(scope="module")
def emul(tmp_path_factory):
    return _build(tmp_path_factory, "cg", coop=False)


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
                                     ys.ctypes.data_as(C.c_void_p), sws.ctypes.data_as(C.c_void_p), n, c, h, h, it, C.c_float(0.1), int(fr), ACT[act], C.c_float(0.05), 1)
    assert rc == 0
    ref64 = A.atom_cg_filter(w.double(), x.double(), y.double(), sw.double(), 0.1, it, act, 0.05, fr)[0]
    ref32 = A.atom_cg_filter(w, x, y, sw, 0.1, it, act, 0.05, fr)[0]
    tol = max(20 * _rel(ref32.numpy(), ref64.numpy()), 2e-5)           # float32 CG against the float64 solution, calibrated by the oracle's own spread
    assert _rel(out, ref64.numpy()) < tol, (_rel(out, ref64.numpy()), tol)


# ---- sd_kernel: the DiMP-family online optimisers against the UNMODIFIED reference modules' golden outputs ------------
@pytest.fixture(scope="module")
def sd(tmp_path_factory):
    return _build(tmp_path_factory, "sd", coop=False)


@pytest.fixture(scope="module")
def sd_coop(tmp_path_factory):
    """block = one OS thread, its threads = fibers, static __shared__ = thread_local: whole cooperative grids (launch_coop)"""
    return _build(tmp_path_factory, "sd", coop=True)


@pytest.fixture(scope="module")
def cg_coop(tmp_path_factory):
    return _build(tmp_path_factory, "cg", coop=True)


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
                               C.c_float(max(float(p["filter_reg"]) ** 2, 1e-3 ** 2)), C.c_float(0.0), _ptr(its), _ptr(losses), 1)
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
                                     C.c_float(0.0), _ptr(its), _ptr(losses), 1)
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
                                  C.c_float(0.9), C.c_float(max(0.1 ** 2, 1e-6)), C.c_float(0.01), _ptr(its), _ptr(losses), 1)
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
                                C.c_float(0.7), C.c_float(0.02), _ptr(its), _ptr(losses), 1)
    assert rc == 0
    _check(g, tag, w, its, losses)


# ---- whole cooperative grids: channel chunks x sample groups, the cross-CTA exchange (gpart / qpart / scalars) through the grid barriers ------------
@pytest.mark.parametrize("tag,n,c,h,it,use_sw,seed,sms", [("n15_it10", 15, 512, 18, 10, True, 21, 12), ("n7_22", 7, 128, 22, 4, True, 24, 6),
                                                          ("n4_c64", 4, 64, 18, 3, False, 23, 5)])
def test_dimp_sd_kernel_source_multi_cta_vs_reference_golden(sd_coop, tag, n, c, h, it, use_sw, seed, sms):
    """The decomposition launch_sd picks for `sms` SMs (e.g. n = 15, C = 512 on 12 SMs: 4 channel chunks x 3 sample groups), against the
    reference module's goldens -- the tracker's own first-frame configuration (15 augmented samples, 10 iterations) included."""
    g = _golden("dimp_sd")
    p = synth.make_dimp_optimizer_params(seed=seed)
    feat, bb = _arr(synth.make_clf_features(seed, n, c, h, h)), _arr(synth.make_boxes(seed + 1, n, center=(h * 16) / 2 - 25))
    sw = _arr(g[tag + "_sw"]) if use_sw else None
    luts = [_arr(p[k].reshape(-1)) for k in ("label_map_predictor.weight", "target_mask_predictor.0.weight", "spatial_weight_predictor.weight")]
    w0, w = _arr(g[tag + "_w0"]), np.full((1, c, 4, 4), np.nan, np.float32)
    its, losses = np.full((it + 1, c, 4, 4), np.nan, np.float32), np.full(it + 1, np.nan, np.float32)

    def run(out_w):
        return sd_coop.sd_emul_dimp_sd_gn(_ptr(w0), _ptr(out_w), _ptr(feat), _ptr(bb), _ptr(sw), n, c, h, h, it, _ptr(luts[0]), _ptr(luts[1]), _ptr(luts[2]),
                                          luts[0].size, C.c_float(0.1), C.c_float(16.0), C.c_float(float(torch.exp(p["log_step_length"]))),
                                          C.c_float(max(float(p["filter_reg"]) ** 2, 1e-3 ** 2)), C.c_float(0.0), _ptr(its), _ptr(losses), sms)
    assert run(w) == 0
    _check(g, tag, w, its, losses)
    w2 = np.full_like(w, np.nan)
    assert run(w2) == 0 and np.array_equal(w, w2)                  # fixed summation orders across the CTAs: bitwise run to run


def test_prdimp_sd_kernel_source_multi_cta_vs_reference_golden(sd_coop):
    g, tag, n, c, h, it, seed = _golden("prdimp_sd"), "n15_22", 15, 512, 22, 10, 31
    feat, bb = _arr(synth.make_clf_features(seed, n, c, h, h)), _arr(synth.make_boxes(seed + 1, n, center=(h * 16) / 2 - 25))
    w0, w, sw = _arr(g[tag + "_w0"]), np.full((1, c, 4, 4), np.nan, np.float32), _arr(g[tag + "_sw"])
    its, losses = np.full((it + 1, c, 4, 4), np.nan, np.float32), np.full(it + 1, np.nan, np.float32)
    rc = sd_coop.sd_emul_prdimp_sd_newton(_ptr(w0), _ptr(w), _ptr(feat), _ptr(bb), _ptr(sw), n, c, h, h, it, C.c_float(float(g[tag + "_sigma"])), C.c_float(16.0),
                                          C.c_float(1.0), C.c_float(0.05 ** 2), C.c_float(0.05), 0, C.c_float(0.0), C.c_float(0.0), 1, C.c_float(0.0),
                                          C.c_float(0.0), _ptr(its), _ptr(losses), 12)
    assert rc == 0
    _check(g, tag, w, its, losses)


@pytest.mark.parametrize("tag,n,c,it,fr,act,seed,sms", [("n12_c16_pr_mlu", 12, 16, 5, False, "mlu", 51, 4), ("n40_c64_pr_mlu", 40, 64, 5, False, "mlu", 52, 10),
                                                        ("n9_c32_fr_none", 9, 32, 4, True, "none", 53, 3), ("n20_c64_pr_relu", 20, 64, 3, False, "relu", 54, 6)])
def test_atom_cg_kernel_source_multi_cta_vs_reference_golden(cg_coop, tag, n, c, it, fr, act, seed, sms):
    """ATOM ConjugateGradient.run on ConvProblem against the outputs of the reference classes (tests/golden/atom_cg.npz, two consecutive
    runs), on the grid launch_cg picks for `sms` SMs."""
    g = _golden("atom_cg")
    x, y, sw = synth.make_atom_memory(seed, n, c, 18, 18, n_filled=ATOM_CG_FILLED.get(tag))
    w = _arr(g[tag + "_w0"])
    xs, ys, sws = _arr(x), _arr(y), _arr(sw)
    for run in (1, 2):
        out = np.full_like(w, np.nan)
        rc = cg_coop.cg_emul_atom_cg_filter(_ptr(w), _ptr(out), _ptr(xs), _ptr(ys), _ptr(sws), n, c, 18, 18, it, C.c_float(0.1), int(fr), ACT[act], C.c_float(0.05), sms)
        assert rc == 0
        ref = g[tag + ("_w" if run == 1 else "_w2")]
        assert _rel(out, ref) < 1e-4, (tag, run, _rel(out, ref))
        w = out
