"""-m gpu parity tests of SURVEY 8 row f4: ECO's per-frame Fourier-domain filter optimiser (`b200trk_eco_filter_cg`, csrc/eco_cg.cu)
through the C ABI against (1) the golden vectors of the UNMODIFIED reference FilterOptim (tests/golden/eco_cg.npz), (2) the oracle
(oracle/eco_oracle.py, pinned to the same vectors on the CPU) at ECO's real block sizes, with the tolerance calibrated by the
float32-vs-float64 spread of the oracle itself, and (3) the unmodified reference FilterOptim object on stock PyTorch-CUDA with the
plug-in seam installed.  (The same kernel source also runs on the CPU under tests/cpu_emul -- tests/test_eco_cpu.py.)"""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "eco_cg.npz")
CASES = [(c, r, b) for c in ("pr_forget", "fr_reset") for r in range(3) for b in range(2)]


def _rel(a, b):
    a, b = torch.as_tensor(a).detach().double().cpu(), torch.as_tensor(b).detach().double().cpu()
    return float((a - b).abs().max() / (b.abs().max() + 1e-30))


def _state_from_golden(g, k, fr):
    if not int(g[k + "has_state"]):
        return None
    return {"p": torch.from_numpy(g[k + "p_in"].copy()).cuda(),
            "r_prev": None if fr else torch.from_numpy(g[k + "r_prev_in"].copy()).cuda(),
            "rho": torch.tensor([float(g[k + "rho_in"])], dtype=torch.float32).cuda()}


@pytest.mark.parametrize("case,run,bi", CASES)
def test_eco_filter_cg_matches_reference_golden(case, run, bi):
    from pytracking_b200 import ops
    g = np.load(GOLD)
    fr, sa, dff, pdp, prp = g[case + "/params"]
    k = "%s/run%d/b%d/" % (case, run, bi)
    D = lambda key: torch.from_numpy(g[key].copy()).cuda()
    hf = D(k + "hf_in")
    en = D(k + "energy_in") if int(g[k + "has_energy"]) else None
    en, st = ops.eco_filter_cg_(hf, D(k + "samples"), D("%s/b%d/yf" % (case, bi)), D(k + "sw"), D("%s/b%d/reg_filter" % (case, bi)), en,
                                int(g[k + "num_iter"]), D(k + "new_xf"), _state_from_golden(g, k, bool(fr)), bool(fr), bool(sa), float(dff),
                                float(g["%s/b%d/lr" % (case, bi)]), float(pdp), float(prp))
    torch.cuda.synchronize()
    assert _rel(hf, g[k + "hf_out"]) < 5e-5
    assert _rel(en, g[k + "energy_out"]) < 1e-6
    assert _rel(st["p"], g[k + "p_out"]) < 5e-5
    assert abs(float(st["rho"]) - float(g[k + "rho_out"])) < 5e-5 * abs(float(g[k + "rho_out"]))
    if (k + "r_prev_out") in g:
        assert _rel(st["r_prev"], g[k + "r_prev_out"]) < 5e-5


def _problem(h, wh, n, c, stored, seed):
    g = torch.Generator().manual_seed(seed)
    samples = torch.zeros(h, wh, n, c, 2)
    samples[:, :, :stored] = torch.randn(h, wh, stored, c, 2, generator=g)
    sw = torch.zeros(n)
    sw[:stored] = torch.rand(stored, generator=g) + 0.1
    sw /= sw.sum()
    ky = torch.arange(-(h - 1) // 2, h // 2 + 1, dtype=torch.float32).view(-1, 1)
    kx = torch.arange(0, wh, dtype=torch.float32).view(1, -1)
    yf = torch.exp(-0.05 * (ky ** 2 + kx ** 2)).view(1, 1, h, wh)                       # a label function's spectrum: real, decaying
    reg = torch.tensor([[0.0, 0.02, 0.05, 0.02, 0.0], [0.02, 0.1, 0.2, 0.1, 0.02], [0.05, 0.2, 0.9, 0.2, 0.05],
                        [0.02, 0.1, 0.2, 0.1, 0.02], [0.0, 0.02, 0.05, 0.02, 0.0]]).view(1, 1, 5, 5)
    hf = 0.01 * torch.randn(1, c, h, wh, 2, generator=g)
    new_xf = [torch.randn(1, c, h, wh, 2, generator=g) for _ in range(2)]
    return samples, sw, yf, reg, hf, new_xf


# ECO default blocks (parameter/eco/default.py): memory 200; deep 64 channels on 15x8 coefficients (one resident slab per CTA, eight
# warps per slab), shallow 16 channels on 63x32 (14 coefficients per CTA, 8 slabs resident + 6 streamed); a partly filled memory
@pytest.mark.parametrize("h,wh,n,c,stored", [(15, 8, 200, 64, 200), (63, 32, 200, 16, 200), (17, 9, 200, 32, 37), (13, 7, 50, 128, 50)])
def test_eco_filter_cg_full_size_vs_oracle(h, wh, n, c, stored):
    from oracle import eco_oracle as E
    from pytracking_b200 import ops
    samples, sw, yf, reg, hf0, new_xf = _problem(h, wh, n, c, stored, seed=h * 100 + c)
    kw = dict(precond_learning_rate=0.0075, precond_data_param=0.3, precond_reg_param=0.15, fletcher_reeves=False, standard_alpha=True,
              direction_forget_factor=(1 - 0.0075) ** 75)
    res = {}
    for dt in (torch.float32, torch.float64):                       # two consecutive runs: the second continues from the CG state
        x, en, st = hf0.to(dt), None, {}
        for r in range(2):
            x, en, st = E.filter_optim_run(x, samples.to(dt), yf.to(dt), sw.to(dt), reg.to(dt), en, st, 5, new_xf[r].to(dt), **kw)
        res[dt] = (x, en, st["p"])
    hf, en, st = hf0.clone().cuda(), None, None
    for r in range(2):
        en, st = ops.eco_filter_cg_(hf, samples.cuda(), yf.cuda(), sw.cuda(), reg.cuda(), en, 5, new_xf[r].cuda(), st, **kw)
    torch.cuda.synchronize()
    spread = max(_rel(res[torch.float32][0], res[torch.float64][0]), _rel(res[torch.float32][2], res[torch.float64][2]))
    tol = max(20 * spread, 2e-5)                                    # float32 CG against the float64 solution, calibrated by the oracle's own spread
    assert _rel(hf, res[torch.float64][0]) < tol, (_rel(hf, res[torch.float64][0]), spread)
    assert _rel(st["p"], res[torch.float64][2]) < tol, (_rel(st["p"], res[torch.float64][2]), spread)
    assert _rel(en, res[torch.float64][1]) < 1e-5
    # bitwise determinism (fixed summation orders)
    hf2, en2, st2 = hf0.clone().cuda(), None, None
    for r in range(2):
        en2, st2 = ops.eco_filter_cg_(hf2, samples.cuda(), yf.cuda(), sw.cuda(), reg.cuda(), en2, 5, new_xf[r].cuda(), st2, **kw)
    assert torch.equal(hf, hf2) and torch.equal(st["p"], st2["p"]) and torch.equal(st["rho"], st2["rho"])


def test_eco_filter_cg_rejects_what_the_kernel_does_not_claim():
    from pytracking_b200 import ops
    samples, sw, yf, reg, hf, new_xf = _problem(9, 5, 8, 16, 8, seed=1)
    with pytest.raises(RuntimeError):                               # compressed_dim 24
        ops.eco_filter_cg_(torch.zeros(1, 24, 9, 5, 2).cuda(), torch.zeros(9, 5, 8, 24, 2).cuda(), yf.cuda(), sw.cuda(), reg.cuda(), None, 2,
                           torch.zeros(1, 24, 9, 5, 2).cuda())
    with pytest.raises(RuntimeError):                               # regularisation filter wider than the half spectrum
        ops.eco_filter_cg_(hf.cuda(), samples.cuda(), yf.cuda(), sw.cuda(), torch.ones(1, 1, 3, 7).cuda(), None, 2, new_xf[0].cuda())
    with pytest.raises(RuntimeError):                               # no energy and nothing to initialise it from
        ops.eco_filter_cg_(hf.cuda(), samples.cuda(), yf.cuda(), sw.cuda(), reg.cuda(), None, 2, None)
    h0 = hf.clone().cuda()
    ops.eco_filter_cg_(h0, samples.cuda(), yf.cuda(), sw.cuda(), reg.cuda(), None, 0, new_xf[0].cuda())     # num_iter = 0: untouched
    assert torch.equal(h0.cpu(), hf)


def test_reference_filter_optim_above_the_engine():
    """The unmodified reference FilterOptim object (two-block TensorLists, ECO default CG settings) on CUDA tensors: stock PyTorch vs
    the plug-in seam, three consecutive runs with memory updates; the CG state stays in the reference's own attributes."""
    from baseline import ref_env
    if not ref_env.reference_available():
        pytest.skip("reference tree not staged (baseline/_ref)")
    from oracle import ref_shims
    ref_shims.install()
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    from pytracking import TensorList
    from pytracking.tracker.eco.optim import FilterOptim
    from pytracking.utils import TrackerParams
    from pytracking_b200 import plugin

    blocks = [(15, 8, 16, 0.025), (9, 5, 64, 0.0075)]
    n = 30

    def make():
        g = torch.Generator().manual_seed(5)
        params = TrackerParams()
        params.fletcher_reeves, params.standard_alpha, params.debug = False, True, 0
        params.direction_forget_factor = (1 - 0.025) ** 75
        params.precond_data_param, params.precond_reg_param = 0.3, 0.15
        params.precond_learning_rate = TensorList([b[3] for b in blocks])
        probs = [_problem(h, wh, n, c, 12, seed=7 + i) for i, (h, wh, c, _) in enumerate(blocks)]
        filt = TensorList([p[4].clone().cuda() for p in probs])
        samples = TensorList([p[0].clone().cuda() for p in probs])
        sw = TensorList([p[1].clone().cuda() for p in probs])
        reg = TensorList([p[3].clone().cuda() for p in probs])
        opt = FilterOptim(params, reg.view(-1) @ reg.view(-1))
        opt.register(filt, samples, TensorList([p[2].clone().cuda() for p in probs]), sw, reg)
        news = [TensorList([torch.randn(1, c, h, wh, 2, generator=g).cuda() for (h, wh, c, _) in blocks]) for _ in range(3)]
        return opt, filt, samples, sw, news

    def drive(opt, filt, samples, sw, news):
        outs = []
        for r, nx in enumerate(news):
            if r:
                for s, w_, xf in zip(samples, sw, nx):
                    s[:, :, 12 + r:13 + r] = xf.permute(2, 3, 0, 1, 4)
                    w_ *= 0.9
                    w_[12 + r] = 0.1
                    w_ /= w_.sum()
            opt.run(5, nx)
            for hf in filt:                                           # ECO.symmetrize_filter (eco.py:381-383)
                hf[:, :, :, 0, :] /= 2
            outs.append([hf.clone() for hf in filt])
        return outs

    ref = drive(*make())
    plugin.install()
    try:
        before = plugin.stats.get("FilterOptim.run", 0)
        opt, filt, samples, sw, news = make()
        got = drive(opt, filt, samples, sw, news)
        assert plugin.stats.get("FilterOptim.run", 0) == before + 3
        assert isinstance(opt.p, TensorList) and opt.p[0].shape == filt[0].shape and opt.rho[0].dim() == 0
    finally:
        plugin.uninstall()
    for r in range(3):
        for b in range(2):
            assert _rel(got[r][b], ref[r][b]) < 2e-4, (r, b, _rel(got[r][b], ref[r][b]))


# ---- score computation (b200trk_eco_apply_filter, b200trk_eco_sample_fs; eco.py:244-252) ----------------------------------------------
LOC_GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "eco_loc.npz")


@pytest.mark.parametrize("name", ["two_blocks_even", "two_blocks_odd", "one_block", "tight"])
def test_eco_score_kernels_match_reference_golden(name):
    """Outputs of the unmodified reference functions (complex.mult(..).sum(1), fourier.sum_fs, fourier.sample_fs, dcf.max2d;
    oracle/gen_eco_golden.py loc) against the two entry points and ops.max2d."""
    from pytracking_b200 import ops
    g = np.load(LOC_GOLD)
    nb = len({k.split("/")[1] for k in g.files if k.startswith(name + "/b")})
    sfs = []
    for b in range(nb):
        sf = ops.eco_apply_filter(torch.from_numpy(g["%s/b%d/filter" % (name, b)]).cuda(), torch.from_numpy(g["%s/b%d/xf" % (name, b)]).cuda())
        assert _rel(sf, g["%s/b%d/sf" % (name, b)]) < 2e-6
        sfs.append(sf)
    scores = ops.eco_sample_fs(sfs, g[name + "/out"].tolist(), g[name + "/weights"].tolist())
    assert tuple(scores.shape) == g[name + "/scores"].shape and _rel(scores, g[name + "/scores"]) < 5e-6, _rel(scores, g[name + "/scores"])
    mv, mi = ops.max2d(scores)
    assert np.array_equal(mi.cpu().numpy().reshape(-1, 2), g[name + "/max_disp"].reshape(-1, 2))
    assert np.allclose(mv.cpu().numpy().reshape(-1), g[name + "/max_score"].reshape(-1), rtol=1e-5)


def test_eco_score_kernels_at_eco_default_sizes_vs_oracle():
    """parameter/eco/default.py: shallow block 16 channels on 63x32 coefficients, deep block 64 on 15x8, five scales, a 250x250 grid."""
    from oracle import eco_oracle as E
    from pytracking_b200 import ops
    g = torch.Generator().manual_seed(3)
    blocks = [(63, 32, 16), (15, 8, 64)]
    filt = [0.1 * torch.randn(1, c, h, wh, 2, generator=g) for (h, wh, c) in blocks]
    xf = [torch.randn(5, c, h, wh, 2, generator=g) for (h, wh, c) in blocks]
    ref = E.sample_fs(E.sum_fs([E.apply_filter(f.double(), x.double()) for f, x in zip(filt, xf)], [1.0, 0.6]), (250, 250))
    sfs = [ops.eco_apply_filter(f.cuda(), x.cuda()) for f, x in zip(filt, xf)]
    scores = ops.eco_sample_fs(sfs[::-1], (250, 250), [0.6, 1.0])        # the caller's order does not matter: sum_fs sorts by rows
    assert _rel(scores, ref) < 5e-6, _rel(scores, ref)
    assert torch.equal(ops.max2d(scores)[1].cpu().view(-1, 2), E.max2d(ref)[1].view(-1, 2))
    assert torch.equal(scores, ops.eco_sample_fs(sfs[::-1], (250, 250), [0.6, 1.0]))
    sf = torch.zeros(1, 1, 9, 5, 2).cuda()
    for bad in ((8, 20), (9, 9)):                                        # smaller than the series / equal to it (fourier.py:43-48)
        with pytest.raises(RuntimeError):
            ops.eco_sample_fs(sf, bad)
    with pytest.raises(RuntimeError):
        ops.eco_sample_fs(torch.zeros(1, 1, 8, 5, 2).cuda(), (20, 20))   # a centred half spectrum has an odd number of rows


@pytest.mark.parametrize("name", ["small", "large", "one_axis"])
def test_eco_shift_fs_matches_reference_golden(name):
    """Outputs of the unmodified fourier.shift_fs (oracle/gen_eco_golden.py shift)."""
    from pytracking_b200 import ops
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "eco_shift.npz"))
    out = ops.eco_shift_fs(torch.from_numpy(g[name + "/a"]).cuda(), float(g[name + "/shift"][0]), float(g[name + "/shift"][1]))
    assert _rel(out, g[name + "/out"]) < 1e-6, _rel(out, g[name + "/out"])


@pytest.mark.parametrize("permuted", [False, True])
@pytest.mark.parametrize("name", ["even", "odd", "rect"])
def test_eco_preprocess_sample_matches_reference_golden(name, permuted):
    """Outputs of the unmodified `ECO.preprocess_sample` (oracle/gen_eco_golden.py prep); `permuted`: x as the tracker hands it over, a
    [S,C,H,W] view of the projection's contiguous [H,W,S,C] result (eco.py:304-309)."""
    from pytracking_b200 import ops
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "eco_prep.npz"))
    x = torch.from_numpy(g[name + "/x"]).cuda()
    if permuted:
        x = x.permute(2, 3, 0, 1).contiguous().permute(2, 3, 0, 1)
    xf = ops.eco_preprocess_sample_(x, torch.from_numpy(g[name + "/window"]).cuda(), torch.from_numpy(g[name + "/interp_y"]).cuda(),
                                    torch.from_numpy(g[name + "/interp_x"]).cuda())
    assert tuple(xf.shape) == g[name + "/xf"].shape and _rel(xf, g[name + "/xf"]) < 2e-6, _rel(xf, g[name + "/xf"])
    assert torch.equal(x.cpu(), torch.from_numpy(g[name + "/x_after"]))          # windowed in place, bit for bit


@pytest.mark.parametrize("s,c,hw", [(5, 16, 62), (5, 64, 15), (30, 96, 62)])
def test_eco_preprocess_sample_at_eco_default_sizes_vs_oracle(s, c, hw):
    """parameter/eco/default.py feature maps: 62x62 (-> 63x32 coefficients) and 15x15 (-> 15x8); the last case is the first frame's call on
    the 30 augmented, not yet projected samples."""
    from oracle import eco_oracle as E
    from pytracking_b200 import ops
    g = torch.Generator().manual_seed(hw)
    x = torch.randn(s, c, hw, hw, generator=g)
    hann = 0.5 * (1 - torch.cos(2 * torch.pi * torch.arange(1, hw + 1).float() / (hw + 1)))
    window = (hann.view(-1, 1) * hann.view(1, -1)).view(1, 1, hw, hw)
    hp, whp = hw + (hw + 1) % 2, hw // 2 + 1
    iy, ix = torch.randn(1, 1, hp, 1, 2, generator=g) / hw, torch.randn(1, 1, 1, whp, 2, generator=g) / hw
    n = min(s, 3)                                                   # the float64 oracle on the first samples only
    ref, _ = E.preprocess_sample(x[:n].double(), window.double(), iy.double(), ix.double())
    xd = x.clone().cuda()
    xf = ops.eco_preprocess_sample_(xd, window.cuda(), iy.cuda(), ix.cuda())
    assert _rel(xf[:n], ref) < 5e-6, _rel(xf[:n], ref)
    assert torch.equal(xd.cpu(), x * window)


# ---- first-frame joint optimisation (b200trk_eco_joint_gn) -----------------------------------------------------------------------------
JOINT_GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "eco_joint.npz")


@pytest.mark.parametrize("bi", [0, 1])
def test_eco_joint_gn_matches_reference_golden(bi):
    """Against the reference's GaussNewtonCG on FactorizedConvProblem (J, J^T by autograd): 3 GN x 4 CG iterations."""
    from pytracking_b200 import ops
    g, k = np.load(JOINT_GOLD), "b%d/" % bi
    D = lambda key: torch.from_numpy(np.ascontiguousarray(g[key])).cuda()
    num_cg, num_gn, lam = int(g["params"][0]), int(g["params"][1]), float(g["params"][2])
    hf, P = D(k + "hf_in"), D(k + "P_in")
    ops.eco_joint_gn_(hf, P, D(k + "samples"), D(k + "yf"), D(k + "sample_weights").sqrt(), D(k + "reg_filter"),
                      D(k + "diag_M_filter").reshape(1, hf.shape[1], hf.shape[2], hf.shape[3]).contiguous(), float(g[k + "diag_M_proj"]), lam,
                      num_cg, num_gn)
    torch.cuda.synchronize()
    assert _rel(hf, g[k + "hf_out"]) < 2e-5 and _rel(P, g[k + "P_out"]) < 2e-5, (_rel(hf, g[k + "hf_out"]), _rel(P, g[k + "P_out"]))


# ECO's first frame (parameter/eco/default.py): 30 augmented samples; deep block 256 -> 64 channels on 15x8 coefficients (one resident
# slab per CTA), shallow block 96 -> 16 on 63x32 (14 coefficients per CTA, 9 slabs resident); 2 GN x 5 CG iterations here
@pytest.mark.parametrize("h,wh,n,cin,c", [(15, 8, 30, 256, 64), (63, 32, 30, 96, 16), (11, 6, 9, 40, 32)])
def test_eco_joint_gn_full_size_vs_oracle(h, wh, n, cin, c):
    from oracle import eco_oracle as E
    from pytracking_b200 import ops
    g = torch.Generator().manual_seed(h * 7 + c)
    samples = torch.randn(h, wh, n, cin, 2, generator=g)
    P0 = torch.linalg.qr(torch.randn(cin, cin, generator=g))[0][:, :c].contiguous()
    _, _, yf, reg, _, _ = _problem(h, wh, 2, 16, 2, seed=3)
    sw = torch.full((n,), 1.0 / n)
    hf0 = torch.zeros(1, c, h, wh, 2)
    res = {}
    for dt in (torch.float32, torch.float64):
        res[dt] = E.joint_gn_run(hf0.to(dt), P0.to(dt), samples.to(dt), yf.to(dt), sw.to(dt), reg.to(dt), 5, 2)
    dMh, dMP, _ = E.joint_precond(samples, P0, yf, reg, 0.3, 0.15, 35.0, 5e-8)
    hf, P = hf0.clone().cuda(), P0.clone().cuda()
    ops.eco_joint_gn_(hf, P, samples.cuda(), yf.cuda(), sw.sqrt().cuda(), reg.cuda(), dMh.reshape(1, c, h, wh).contiguous().cuda(), float(dMP),
                      5e-8, 5, 2)
    torch.cuda.synchronize()
    spread = max(_rel(res[torch.float32][0], res[torch.float64][0]), _rel(res[torch.float32][1], res[torch.float64][1]))
    tol = max(20 * spread, 2e-5)
    assert _rel(hf, res[torch.float64][0]) < tol and _rel(P, res[torch.float64][1]) < tol, \
        (_rel(hf, res[torch.float64][0]), _rel(P, res[torch.float64][1]), spread)
    hf2, P2 = hf0.clone().cuda(), P0.clone().cuda()
    ops.eco_joint_gn_(hf2, P2, samples.cuda(), yf.cuda(), sw.sqrt().cuda(), reg.cuda(), dMh.reshape(1, c, h, wh).contiguous().cuda(), float(dMP),
                      5e-8, 5, 2)
    assert torch.equal(hf, hf2) and torch.equal(P, P2)                  # bitwise determinism


def test_reference_joint_optimizer_above_the_engine():
    """The unmodified reference FactorizedConvProblem + GaussNewtonCG objects on CUDA tensors (two blocks, permuted sample view):
    stock PyTorch autograd vs the plug-in seam."""
    from baseline import ref_env
    if not ref_env.reference_available():
        pytest.skip("reference tree not staged (baseline/_ref)")
    from oracle import ref_shims
    ref_shims.install()
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    from pytracking import TensorList
    from pytracking.libs.optimization import GaussNewtonCG
    from pytracking.tracker.eco.optim import FactorizedConvProblem
    from pytracking.utils import TrackerParams
    from pytracking_b200 import plugin

    blocks = [(15, 8, 48, 16), (9, 5, 96, 64)]
    n = 12

    def make():
        g = torch.Generator().manual_seed(9)
        params = TrackerParams()
        params.precond_data_param, params.precond_reg_param, params.precond_proj_param, params.projection_reg = 0.3, 0.15, 35, 5e-8
        X = TensorList([torch.randn(n, cin, h, wh, 2, generator=g).cuda().permute(2, 3, 0, 1, 4) for (h, wh, cin, c) in blocks])   # eco.py:133
        P = TensorList([torch.linalg.qr(torch.randn(cin, cin, generator=g))[0][:, :c].contiguous().cuda() for (h, wh, cin, c) in blocks])
        hf = TensorList([torch.zeros(1, c, h, wh, 2).cuda() for (h, wh, cin, c) in blocks])
        probs = [_problem(h, wh, 2, 16, 2, seed=3) for (h, wh, cin, c) in blocks]
        prob = FactorizedConvProblem(X, TensorList([p[2].cuda() for p in probs]), TensorList([p[3].cuda() for p in probs]), P, params,
                                     TensorList([torch.ones(1).cuda() / n for _ in blocks]))
        var = hf.concat(P)
        return GaussNewtonCG(prob, var, debug=False), var

    opt, ref = make()
    opt.run(5, 3)
    plugin.install()
    try:
        before = plugin.stats.get("GaussNewtonCG.run[eco]", 0)
        opt, got = make()
        opt.run(5, 3)
        assert plugin.stats.get("GaussNewtonCG.run[eco]", 0) == before + 1
    finally:
        plugin.uninstall()
    for i in range(4):
        assert _rel(got[i], ref[i]) < 1e-4, (i, _rel(got[i], ref[i]))


@pytest.mark.parametrize("score_seams", [False, True])
def test_reference_eco_tracker_above_the_engine(score_seams):
    """The UNMODIFIED reference ECO tracker (parameter/eco/default.py, seeded random-init ResNet18m1 features) on the synthetic sequence:
    stock PyTorch-CUDA vs `plugin.install()` (first-frame GaussNewtonCG.run and every FilterOptim.run on the library; with `score_seams`
    also ECO.preprocess_sample, ECO.apply_filter and the sample_fs of ECO.localize_target).  The CPU counterpart with the oracle behind the entry points is
    tests/test_eco_tracker_cpu.py."""
    from baseline import ref_env
    if not ref_env.reference_available():
        pytest.skip("reference tree not staged (baseline/_ref)")
    from baseline import ref_tracker
    from pytracking_b200 import plugin, synth
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    n_frames, ov = 12, dict(init_CG_iter=40, init_GN_iter=4, train_skipping=3)
    frames, bb = synth.make_sequence(0, num_frames=n_frames)

    def drive():
        trk = ref_tracker.build_eco(device="cuda", overrides=ov)
        torch.manual_seed(0)
        trk.initialize(frames[0], {"init_bbox": list(bb)})
        return [trk.track(frames[i])["target_bbox"] for i in range(1, n_frames + 1)], trk

    try:
        ref_boxes, ref_trk = drive()
    except Exception as e:                                          # not the engine's doing: the stock reference on this torch build
        pytest.skip("the reference ECO tracker does not run on stock PyTorch-CUDA here: %r" % (e,))
    plugin.install(skip=() if score_seams else ("ECO.apply_filter", "fourier", "preprocess_sample"))
    try:
        before = dict(plugin.stats)
        boxes, trk = drive()
        runs = sum(1 for f in range(2, n_frames + 2) if f % ov["train_skipping"] == 1)
        assert plugin.stats.get("GaussNewtonCG.run[eco]", 0) == before.get("GaussNewtonCG.run[eco]", 0) + 1
        assert plugin.stats.get("FilterOptim.run", 0) == before.get("FilterOptim.run", 0) + runs
        assert plugin.stats.get("ECO.apply_filter", 0) == before.get("ECO.apply_filter", 0) + (n_frames if score_seams else 0)
        assert plugin.stats.get("fourier.sample_fs[eco]", 0) == before.get("fourier.sample_fs[eco]", 0) + (n_frames if score_seams else 0)
        assert plugin.stats.get("ECO.preprocess_sample", 0) == before.get("ECO.preprocess_sample", 0) + (n_frames + 1 if score_seams else 0)
    finally:
        plugin.uninstall()
    for b in range(2):
        assert _rel(trk.filter[b], ref_trk.filter[b]) < 1e-3, (b, _rel(trk.filter[b], ref_trk.filter[b]))
    for a, b in zip(ref_boxes, boxes):
        assert max(abs(x - y) for x, y in zip(a, b)) < 1e-2, (a, b)
