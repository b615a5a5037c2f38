"""The Precise RoI Pooling kernels of csrc/prroi_kernels.cuh (the reference's one native op, forward / backward /
coordinate backward) executed ON THE CPU: the same source file the CUDA build compiles (`cuobjdump -sass` identical before and after
the kernels moved into the header), built as host code under tests/cpu_emul/cuda_shim.h with the launch shapes of csrc/prroi.cu.
CPU-tier counterparts of tests/test_gpu_parity.py::test_prroi_known_answer / test_prroi_all_three_kernels."""
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
    out = os.path.join(str(tmp_path_factory.mktemp("prroi_emul")), "libprroi_emul.so")
    subprocess.run(["g++", "-std=c++17", "-O1", "-g", "-pthread", "-shared", "-fPIC", "-ffp-contract=off", "-fno-gnu-unique", "-Wno-unknown-pragmas",
                    os.path.join(ROOT, "tests", "cpu_emul", "prroi_emul.cpp"), "-o", out], check=True, capture_output=True)
    return C.CDLL(out)


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def _forward(emul, feat, rois, ph, pw, scale):
    B, Cc, H, W = feat.shape
    out = np.full((rois.shape[0], Cc, ph, pw), np.nan, np.float32)
    assert emul.prroi_emul_forward(_p(feat), _p(rois), _p(out), B, Cc, H, W, rois.shape[0], ph, pw, C.c_float(scale)) == 0
    return out


def test_prroi_kernel_source_known_answer(emul):
    """The reference's only known-answer test (PreciseRoIPooling/pytorch/tests/test_prroi_pooling2d.py:21-35): on integer-aligned RoIs with
    spatial_scale 0.5, PrRoIPool 7x7 equals avg_pool2d(k=2, s=1) slices."""
    import torch.nn.functional as F
    feat = torch.rand(4, 16, 24, 32, generator=torch.Generator().manual_seed(0))
    rois = np.array([[0, 0, 0, 14, 14], [1, 14, 14, 28, 28]], dtype=np.float32)
    out = _forward(emul, np.ascontiguousarray(feat.numpy()), rois, 7, 7, 0.5)
    ref = F.avg_pool2d(feat, kernel_size=2, stride=1).numpy()
    assert np.allclose(out[0], ref[0, :, :7, :7], atol=1e-6) and np.allclose(out[1], ref[1, :, 7:14, 7:14], atol=1e-6)


@pytest.mark.parametrize("B,Cc,H,W,R,ph,pw,scale,seed", [(2, 16, 18, 18, 6, 4, 4, 1.0 / 16, 1), (1, 32, 36, 36, 10, 5, 5, 1.0 / 8, 2),
                                                        (3, 8, 9, 11, 4, 3, 1, 0.9, 3)])
def test_prroi_kernel_sources_all_three(emul, B, Cc, H, W, R, ph, pw, scale, seed):
    from oracle import prroi_oracle as P
    g = torch.Generator().manual_seed(seed)
    feat = torch.randn(B, Cc, H, W, generator=g)
    ext_w, ext_h = W / scale, H / scale
    x1 = torch.rand(R, generator=g) * ext_w * 0.5
    y1 = torch.rand(R, generator=g) * ext_h * 0.5
    bw = (0.15 + 0.45 * torch.rand(R, generator=g)) * ext_w
    bh = (0.15 + 0.45 * torch.rand(R, generator=g)) * ext_h
    rois = torch.stack([torch.randint(0, B, (R,), generator=g).float(), x1, y1, x1 + bw, y1 + bh], 1).contiguous()
    rois[0, 3] = ext_w + 5.0                       # a box leaving the feature map: zero padding outside
    og = torch.randn(R, Cc, ph, pw, generator=g)
    f, r, o = np.ascontiguousarray(feat.numpy()), np.ascontiguousarray(rois.numpy()), np.ascontiguousarray(og.numpy())
    out = _forward(emul, f, r, ph, pw, scale)
    out_ref = P.forward(f, r, ph, pw, scale)
    assert _rel(out, out_ref) < 1e-5
    fg = np.full_like(f, np.nan)
    assert emul.prroi_emul_backward(_p(r), _p(o), _p(fg), B, Cc, H, W, R, ph, pw, C.c_float(scale)) == 0
    assert _rel(fg, P.backward(f, r, o, ph, pw, scale)) < 1e-5
    rg = np.full((R, 5), np.nan, np.float32)
    assert emul.prroi_emul_coor_backward(_p(f), _p(r), _p(out), _p(o), _p(rg), B, Cc, H, W, R, ph, pw, C.c_float(scale)) == 0
    assert _rel(rg, P.coor_backward(f, r, out_ref, o, ph, pw, scale)) < 1e-4
    z = np.array([[0, 5.0, 5.0, 5.0, 9.0]], np.float32)      # degenerate RoI (zero area): zero output, as the reference (win_size == 0)
    assert float(np.abs(_forward(emul, f, z, ph, pw, scale)).max()) == 0.0
