"""ATOM's first-frame joint optimisation (GaussNewtonCG on FactorizedConvProblem, pytracking/libs/optimization.py:328-421,
pytracking/tracker/atom/optim.py:6-68) executed ON THE CPU: the eight `gn_*` kernels of csrc/atom_gn_kernels.cuh together with the stage-2
kernels they are built from (conv1x1, conv2d 'same', apply_feat_transpose) -- the same source files the CUDA build compiles (`cuobjdump
-sass` identical before and after they moved into headers) -- run launch for launch as `b200trk_atom_gn_joint` issues them, against the
golden outputs of the unmodified reference classes (autograd Jacobians; tests/golden/atom_gn.npz) and the explicit-Jacobian oracle.
CPU-tier counterpart of tests/test_gpu_parity.py::test_atom_gn_joint_golden."""
import ctypes as C
import os
import shutil
import subprocess

import numpy as np
import pytest
import torch

from pytracking_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CASES = {"n6_c32_16": (6, 32, 16, 3, 2, True, "mlu", 101), "n10_c64_32_pr": (10, 64, 32, 4, 3, False, "relu", 102)}
ACT = {"none": 0, "relu": 1, "elu": 2, "mlu": 3}


@pytest.fixture(scope="module")
def emul(tmp_path_factory):
    if shutil.which("g++") is None:
        pytest.skip("no g++")
    out = os.path.join(str(tmp_path_factory.mktemp("atom_gn_emul")), "libatom_gn_emul.so")
    subprocess.run(["g++", "-std=c++17", "-O1", "-g", "-pthread", "-shared", "-fPIC", "-ffp-contract=off", "-fno-gnu-unique", "-Wno-unknown-pragmas",
                    os.path.join(ROOT, "tests", "cpu_emul", "atom_gn_emul.cpp"), "-o", out], check=True, capture_output=True)
    return C.CDLL(out)


def _rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


@pytest.mark.parametrize("tag", sorted(CASES))
def test_atom_gn_joint_kernel_sources_vs_reference_golden(emul, tag):
    from oracle import atom_oracle as A
    g = np.load(os.path.join(ROOT, "tests", "golden", "atom_gn.npz"))
    n, cin, cc, ncg, ngn, fr, act, seed = CASES[tag]
    x, y, sw = synth.make_atom_memory(seed, n, cin, 18, 18)
    w, P = g[tag + "_w0"].copy(), np.ascontiguousarray(g[tag + "_P0"]).copy()
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    rc = emul.atom_gn_emul_joint(p(w), p(P), p(np.ascontiguousarray(x.numpy())), p(np.ascontiguousarray(y.numpy())), p(np.ascontiguousarray(sw.numpy())), n, cin,
                                 cc, 18, 18, ncg, ngn, C.c_float(0.1), C.c_float(1e-2), int(fr), ACT[act], C.c_float(0.05), 148)
    assert rc == 0
    assert _rel(w, g[tag + "_w"]) < 1e-4 and _rel(P, g[tag + "_P"]) < 1e-4, (_rel(w, g[tag + "_w"]), _rel(P, g[tag + "_P"]))
    w64, P64 = A.atom_gn_joint(torch.from_numpy(g[tag + "_w0"]).double(), torch.from_numpy(g[tag + "_P0"]).double(), x.double(), y.double(), sw.double(), 0.1,
                               1e-2, ncg, ngn, act, 0.05, fr)
    assert _rel(w, w64.numpy()) < 1e-4 and _rel(P, P64.numpy()) < 1e-4
