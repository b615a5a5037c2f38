"""ECO's score computation (eco.py:244-252): ECO.apply_filter and fourier.sample_fs(fourier.sum_fs(weight * sf), output_sz).
CPU tier: (1) the restatement in oracle/eco_oracle.py against outputs of the UNMODIFIED reference functions (tests/golden/eco_loc.npz,
oracle/gen_eco_golden.py loc); (2) the kernel source of pytracking_b200/csrc/eco_loc_kernels.cuh, compiled as host code under
tests/cpu_emul/cuda_shim.h and launched as the C ABI launches it, against the same goldens incl. the arg-max of dcf.max2d; (3) ECO's
real sizes (63x32 + 15x8 coefficients, five scales, a 250x250 grid) against the float64 oracle; (4) what the entry point rejects."""
import ctypes as C
import os
import shutil
import subprocess

import numpy as np
import pytest
import torch

from oracle import eco_oracle as E

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden", "eco_loc.npz")
CASES = ("two_blocks_even", "two_blocks_odd", "one_block", "tight")


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLD)


def _blocks(g, name):
    nb = len({k.split("/")[1] for k in g.files if k.startswith(name + "/b")})
    return [(g["%s/b%d/filter" % (name, b)], g["%s/b%d/xf" % (name, b)], g["%s/b%d/sf" % (name, b)]) for b in range(nb)]


def _rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / np.abs(b).max())


@pytest.mark.parametrize("name", CASES)
@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
def test_oracle_matches_reference_functions(gold, name, dtype):
    sfs = []
    for filt, xf, sf_ref in _blocks(gold, name):
        sf = E.apply_filter(torch.from_numpy(filt).to(dtype), torch.from_numpy(xf).to(dtype))
        assert _rel(sf.numpy(), sf_ref) < 1e-6
        sfs.append(sf)
    scores = E.sample_fs(E.sum_fs(sfs, gold[name + "/weights"].tolist()), gold[name + "/out"])
    assert scores.shape == gold[name + "/scores"].shape and _rel(scores.numpy(), gold[name + "/scores"]) < 2e-6
    mv, mi = E.max2d(scores)
    assert np.array_equal(mi.numpy().reshape(-1, 2), gold[name + "/max_disp"].reshape(-1, 2))
    assert np.allclose(mv.numpy().reshape(-1), gold[name + "/max_score"].reshape(-1), rtol=1e-5)


# ---- the kernel source on the CPU ------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def emul(tmp_path_factory):
    if shutil.which("g++") is None:
        pytest.skip("no g++")
    out = os.path.join(str(tmp_path_factory.mktemp("eco_loc")), "libeco_loc_emul.so")
    cmd = ["g++", "-std=c++17", "-O2", "-g", "-pthread", "-shared", "-fPIC", "-fno-gnu-unique", "-ffp-contract=off", "-Wno-unknown-pragmas",
           "-fsanitize=alignment", os.path.join(ROOT, "tests", "cpu_emul", "eco_loc_emul.cpp"), "-o", out]
    if subprocess.run(cmd, capture_output=True).returncode != 0:
        cmd.remove("-fsanitize=alignment")
        subprocess.run(cmd, check=True, capture_output=True)
    return C.CDLL(out)


P = lambda a: a.ctypes.data_as(C.c_void_p)


def _apply(emul, filt, xf):
    s, c, h, wh, _ = xf.shape
    sf = np.full((s, 1, h, wh, 2), np.nan, np.float32)
    assert emul.eco_loc_emul_apply_filter(P(filt), P(xf), P(sf), s, c, h, wh) == 0
    return sf


def _sample(emul, sfs, weights, out_sz):
    nb, s = len(sfs), sfs[0].shape[0]
    scores = np.full((s, 1, int(out_sz[0]), int(out_sz[1])), np.nan, np.float32)
    ptrs = (C.c_void_p * nb)(*[a.ctypes.data for a in sfs])
    hs, ws = (C.c_int * nb)(*[a.shape[2] for a in sfs]), (C.c_int * nb)(*[a.shape[3] for a in sfs])
    wts = (C.c_float * nb)(*[float(w) for w in weights]) if weights is not None else None
    rc = emul.eco_loc_emul_sample_fs(ptrs, hs, ws, wts, nb, s, int(out_sz[0]), int(out_sz[1]), P(scores))
    return rc, scores


@pytest.mark.parametrize("name", CASES)
def test_emulated_kernels_match_reference_functions(emul, gold, capfd, name):
    sfs = []
    for filt, xf, sf_ref in _blocks(gold, name):
        sf = _apply(emul, filt, xf)
        assert _rel(sf, sf_ref) < 2e-6
        sfs.append(sf)
    rc, scores = _sample(emul, sfs, gold[name + "/weights"], gold[name + "/out"])
    assert rc == 0 and "runtime error" not in capfd.readouterr().err
    assert _rel(scores, gold[name + "/scores"]) < 5e-6, _rel(scores, gold[name + "/scores"])
    mv, mi = E.max2d(torch.from_numpy(scores))                       # dcf.max2d on the kernel's map: the reference's arg-max cell
    assert np.array_equal(mi.numpy().reshape(-1, 2), gold[name + "/max_disp"].reshape(-1, 2))


def test_emulated_kernels_at_eco_default_sizes(emul, capfd):
    """parameter/eco/default.py: shallow block 16 channels on 63x32 coefficients, deep block 64 on 15x8, five scales, output_sz 250x250."""
    g = torch.Generator().manual_seed(3)
    blocks = [(63, 32, 16), (15, 8, 64)]
    filt = [0.1 * torch.randn(1, c, h, wh, 2, generator=g) for (h, wh, c) in blocks]
    xf = [torch.randn(5, c, h, wh, 2, generator=g) for (h, wh, c) in blocks]
    ref = E.sample_fs(E.sum_fs([E.apply_filter(f.double(), x.double()) for f, x in zip(filt, xf)], [1.0, 0.6]), (250, 250))
    sfs = [_apply(emul, f.numpy(), x.numpy()) for f, x in zip(filt, xf)]
    rc, scores = _sample(emul, sfs[::-1], [0.6, 1.0], (250, 250))      # the caller's order does not matter: sum_fs sorts by rows
    assert rc == 0 and "runtime error" not in capfd.readouterr().err
    assert _rel(scores, ref.numpy()) < 5e-6, _rel(scores, ref.numpy())
    assert torch.equal(E.max2d(torch.from_numpy(scores))[1], E.max2d(ref)[1])


def test_entry_point_rejects_what_the_kernel_does_not_claim(emul):
    sf = np.zeros((1, 1, 9, 5, 2), np.float32)
    assert _sample(emul, [sf], None, (8, 20))[0] == 2                  # grid smaller than the series (fourier.py:47-48 raises)
    assert _sample(emul, [sf], None, (9, 9))[0] == 2                   # equal size: the reference's other branch
    assert _sample(emul, [np.zeros((1, 1, 8, 5, 2), np.float32)], None, (20, 20))[0] == 2       # even number of rows
    assert _sample(emul, [sf, np.zeros((1, 1, 5, 7, 2), np.float32)], None, (20, 20))[0] == 2   # more columns than the largest block
    assert _sample(emul, [sf], None, (10, 9))[0] == 0


def test_score_seams_claim_only_what_the_library_serves(gold):
    """`plugin.install()` on the ECO module: `ECO.apply_filter` and the module-scoped `fourier.sample_fs` go to the library for CUDA float32
    series on a larger grid and to the reference for everything else (CPU tensors here, TensorLists, rescale=False, a grid equal to the
    series); every other name of `fourier` is the reference module's; ATOM's module keeps the reference `fourier`; uninstall restores."""
    import unittest.mock as um
    from baseline import ref_env
    if not ref_env.reference_available():
        pytest.skip("reference tree not staged (baseline/_ref)")
    from oracle import ref_shims
    ref_shims.install()
    import pytracking.libs.fourier as ref_fourier
    import pytracking.tracker.atom.atom as atom_mod
    import pytracking.tracker.eco.eco as eco_mod
    from pytracking import TensorList
    from pytracking_b200 import ops, plugin
    name = "two_blocks_even"
    sf = [torch.from_numpy(b[2]) for b in _blocks(gold, name)]
    a = ref_fourier.sum_fs(TensorList(sf))
    out = torch.Tensor(gold[name + "/out"].tolist())
    ref = ref_fourier.sample_fs(a, out)
    calls = []
    plugin.install()
    try:
        assert eco_mod.fourier is not ref_fourier and atom_mod.fourier is ref_fourier
        assert eco_mod.fourier.cfft2 is ref_fourier.cfft2 and eco_mod.fourier.sum_fs is ref_fourier.sum_fs
        assert torch.equal(eco_mod.fourier.sample_fs(a, out), ref)                                   # CPU tensor: the reference
        with um.patch.object(ops, "eco_sample_fs", lambda s, o, weights=None: calls.append(tuple(o)) or E.sample_fs(s, o)), \
                um.patch.object(plugin, "_inference", lambda *ts: all(isinstance(t, torch.Tensor) and t.dtype == torch.float32 for t in ts)):
            got = eco_mod.fourier.sample_fs(a, out)
            assert calls == [tuple(int(v) for v in gold[name + "/out"])] and _rel(got.numpy(), ref.numpy()) < 2e-6
            eco_mod.fourier.sample_fs(a, out, False)                                                 # rescale=False
            eco_mod.fourier.sample_fs(TensorList([a]), out)                                          # a TensorList
            eco_mod.fourier.sample_fs(a, torch.Tensor([a.shape[2], 2 * a.shape[3] - 1]))             # the series' own size
            eco_mod.fourier.sample_fs(a)                                                             # no grid
            assert len(calls) == 1
    finally:
        plugin.uninstall()
    assert eco_mod.fourier is ref_fourier and "apply_filter" in eco_mod.ECO.__dict__


# ---- ECO.preprocess_sample (eco.py:297-300) -----------------------------------------------------------------------------------------
PREP_GOLD = os.path.join(ROOT, "tests", "golden", "eco_prep.npz")


@pytest.mark.parametrize("name", ["even", "odd", "rect"])
@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
def test_preprocess_oracle_matches_reference_method(name, dtype):
    g = np.load(PREP_GOLD)
    xf, v = E.preprocess_sample(torch.from_numpy(g[name + "/x"]).to(dtype), torch.from_numpy(g[name + "/window"]),
                                torch.from_numpy(g[name + "/interp_y"]), torch.from_numpy(g[name + "/interp_x"]))
    assert xf.shape == g[name + "/xf"].shape and _rel(xf.numpy(), g[name + "/xf"]) < 1e-6
    assert _rel(v.numpy(), g[name + "/x_after"]) < 1e-6


def _preprocess(emul, x, window, iy, ix, permuted=False):
    """permuted: x as the tracker hands it over -- a [S,C,H,W] view of the projection's contiguous [H,W,S,C] result (eco.py:304-309)"""
    s, c, h, w = x.shape
    x = np.ascontiguousarray(x.transpose(2, 3, 0, 1)).transpose(2, 3, 0, 1) if permuted else np.ascontiguousarray(x).copy()
    st = [C.c_longlong(v // 4) for v in x.strides]
    xf = np.full((s, c, h + (h + 1) % 2, w // 2 + 1, 2), np.nan, np.float32)
    base = x.ctypes.data if not permuted else x.transpose(2, 3, 0, 1).ctypes.data
    assert emul.eco_loc_emul_preprocess(C.c_void_p(base), st[0], st[1], st[2], st[3], P(np.ascontiguousarray(window)), P(np.ascontiguousarray(iy)),
                                        P(np.ascontiguousarray(ix)), P(xf), s, c, h, w) == 0
    return xf, x


@pytest.mark.parametrize("permuted", [False, True])
@pytest.mark.parametrize("name", ["even", "odd", "rect"])
def test_emulated_preprocess_kernel_matches_reference_method(emul, capfd, name, permuted):
    g = np.load(PREP_GOLD)
    xf, x_after = _preprocess(emul, g[name + "/x"], g[name + "/window"], g[name + "/interp_y"], g[name + "/interp_x"], permuted)
    assert "runtime error" not in capfd.readouterr().err
    assert _rel(xf, g[name + "/xf"]) < 2e-6, _rel(xf, g[name + "/xf"])
    assert np.array_equal(x_after, g[name + "/x_after"])             # the argument windowed in place, bit for bit


@pytest.mark.parametrize("s,c,hw", [(5, 16, 62), (5, 64, 15), (2, 3, 61)])
def test_emulated_preprocess_kernel_at_eco_default_sizes(emul, capfd, s, c, hw):
    """parameter/eco/default.py feature maps (62x62 shallow -> 63x32 coefficients, 15x15 deep -> 15x8) against the float64 oracle."""
    g = torch.Generator().manual_seed(hw)
    x = torch.randn(s, c, hw, hw, generator=g)
    hann = 0.5 * (1 - torch.cos(2 * torch.pi * torch.arange(1, hw + 1).float() / (hw + 1)))
    window = (hann.view(-1, 1) * hann.view(1, -1)).view(1, 1, hw, hw)
    hp, whp = hw + (hw + 1) % 2, hw // 2 + 1
    iy, ix = torch.randn(1, 1, hp, 1, 2, generator=g) / hw, torch.randn(1, 1, 1, whp, 2, generator=g) / hw
    ref, _ = E.preprocess_sample(x.double(), window.double(), iy.double(), ix.double())
    xf, _ = _preprocess(emul, x.numpy(), window.numpy(), iy.numpy(), ix.numpy())
    assert "runtime error" not in capfd.readouterr().err
    assert _rel(xf, ref.numpy()) < 5e-6, _rel(xf, ref.numpy())


# ---- fourier.shift_fs (fourier.py:78-92) ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["small", "large", "one_axis"])
def test_shift_fs_oracle_and_emulated_kernel_match_reference_function(emul, name):
    g = np.load(os.path.join(ROOT, "tests", "golden", "eco_shift.npz"))
    a, shift, ref = g[name + "/a"], g[name + "/shift"], g[name + "/out"]
    assert np.array_equal(E.shift_fs(torch.from_numpy(a), shift).numpy(), ref)               # the float32 restatement: bit for bit
    assert _rel(E.shift_fs(torch.from_numpy(a).double(), shift).numpy(), ref) < 2e-6
    out = np.full_like(a, np.nan)
    s, c, h, wh, _ = a.shape
    assert emul.eco_loc_emul_shift_fs(P(a), P(out), s, c, h, wh, C.c_float(shift[0]), C.c_float(shift[1])) == 0
    assert _rel(out, ref) < 1e-6, _rel(out, ref)
