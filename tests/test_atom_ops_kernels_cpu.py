"""The stage-2 kernels of csrc/atom_ops_kernels.cuh (SURVEY 8 rows S1.3 feature normalisation, S2.2 conv1x1, S2.3 Fourier score
interpolation, S2.4 softmax_reg) executed ON THE CPU: the same source file the CUDA build compiles (`cuobjdump -sass` identical before
and after the kernels moved into the header), built as host code under tests/cpu_emul/cuda_shim.h with the launch shapes of
csrc/atom_ops.cu.  CPU-tier counterparts of the `-m gpu` tests in tests/test_gpu_parity.py: the golden vectors come from the unmodified
reference (`activation.softmax_reg`, the `fourier` chain of ATOM.localize_target), the other two are checked against the oracle."""
import ctypes as C
import os
import shutil
import subprocess

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="module")
def emul(tmp_path_factory):
    if shutil.which("g++") is None:
        pytest.skip("no g++")
    out = os.path.join(str(tmp_path_factory.mktemp("atom_emul")), "libatom_emul.so")
    subprocess.run(["g++", "-std=c++17", "-O1", "-g", "-pthread", "-shared", "-fPIC", "-ffp-contract=off", "-fno-gnu-unique", "-Wno-unknown-pragmas",
                    os.path.join(ROOT, "tests", "cpu_emul", "atom_emul.cpp"), "-o", out], check=True, capture_output=True)
    return C.CDLL(out)


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_softmax_reg_kernel_source_vs_reference_golden(emul, tag):
    g = np.load(os.path.join(GOLDEN, "softmax_reg.npz"))
    x = np.ascontiguousarray(g[tag + "_x"].reshape(g[tag + "_x"].shape[0], -1))
    reg = g[tag + "_reg"][0]
    y = np.empty_like(x)
    assert emul.atom_emul_softmax_reg(_p(x), _p(y), x.shape[0], x.shape[1], 0 if np.isnan(reg) else 1, C.c_float(0.0 if np.isnan(reg) else float(reg))) == 0
    assert _rel(y, g[tag + "_y"].reshape(x.shape[0], -1)) < 1e-5


@pytest.mark.parametrize("tag,S,H,ksz,osz", [("s18_k4", 3, 18, 4, 288), ("s18_k4_o72", 2, 18, 4, 72), ("s17_k5", 2, 17, 5, 64), ("s22_k4", 1, 22, 4, 352)])
def test_fourier_interp_kernel_source_vs_reference_golden(emul, tag, S, H, ksz, osz):
    """cfft2 -> shift_fs -> sum_fs -> sample_fs of the reference (pytracking/libs/fourier.py) against the direct evaluation; the arg-max
    on the upsampled grid (what ATOM.localize_target reads) must be the reference's."""
    g = np.load(os.path.join(GOLDEN, "fourier.npz"))
    sc = np.ascontiguousarray(g[tag + "_scores"])
    up = np.empty((S, 1, osz, osz), np.float32)
    assert emul.atom_emul_fourier_interp(_p(sc), _p(up), S, H, H, ksz, ksz, osz, osz) == 0
    assert _rel(up, g[tag + "_up"]) < 1e-4
    assert np.array_equal(up.reshape(S, -1).argmax(1), g[tag + "_up"].reshape(S, -1).argmax(1))


def test_feature_normalize_and_conv1x1_kernel_sources_vs_oracle(emul):
    from oracle import atom_oracle as A
    g = torch.Generator().manual_seed(5)
    for p in (2.0, 1.5):
        x = torch.randn(3, 24, 18, 18, generator=g)
        y = x.clone().numpy()
        assert emul.atom_emul_feature_normalize(_p(y), 3, 24, 18, 18, C.c_float(p)) == 0
        assert _rel(y, A.feature_normalize(x, p).numpy()) < 1e-5
    for (S, cin, cout, hw) in ((2, 40, 24, 18), (1, 256, 64, 18), (3, 16, 70, 11)):        # partial tiles in every dimension
        x = torch.randn(S, cin, hw, hw, generator=g)
        P = torch.randn(cout, cin, 1, 1, generator=g) / cin ** 0.5
        out = np.empty((S, cout, hw, hw), np.float32)
        assert emul.atom_emul_conv1x1(_p(np.ascontiguousarray(x.numpy())), _p(np.ascontiguousarray(P.numpy())), _p(out), S, cin, cout, hw, hw) == 0
        assert _rel(out, A.conv1x1(x, P).numpy()) < 1e-5
