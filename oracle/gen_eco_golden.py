"""TEST INFRASTRUCTURE ONLY -- golden vectors of ECO's online filter optimiser from the UNMODIFIED reference.

    python -m oracle.gen_eco_golden          (needs the reference: baseline/_ref or /root/reference)

Drives `pytracking.tracker.eco.optim.FilterOptim` (optim.py:120-208) exactly as `ECO.initialize` / `ECO.track` do
(eco.py:166-170, 236-246): a two-block TensorList (a "shallow" and a "deep" block with different sizes, channel counts,
regularisation filters from `dcf.get_reg_filter` and learning rates), three consecutive `run` calls with a memory
update in between, so that the CG state (p, rho, r_prev) carried between runs with direction_forget_factor != 0 and the
running sample energy are covered.  Two parameter sets: the ECO default (Polak-Ribiere, forgetting) and Fletcher-Reeves
with a state reset (direction_forget_factor = 0) and the non-standard alpha.
Writes tests/golden/eco_cg.npz: per case / run / block the inputs and outputs of the run; `joint()` writes tests/golden/eco_joint.npz
(the first-frame joint optimisation of filter and projection matrix).
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")

BLOCKS = [dict(H=9, Wh=5, C=16, lr=0.025, reg=dict(reg_window_min=1e-4, reg_window_edge=10e-3, reg_sparsity_threshold=0.05)),
          dict(H=7, Wh=4, C=64, lr=0.0075, reg=dict(reg_window_min=10e-4, reg_window_edge=50e-3, reg_sparsity_threshold=0.1))]
N = 6
CASES = {"pr_forget": dict(fletcher_reeves=False, standard_alpha=True, direction_forget_factor=(1 - 0.025) ** 75, iters=(4, 3, 3)),
         "fr_reset": dict(fletcher_reeves=True, standard_alpha=False, direction_forget_factor=0, iters=(3, 2, 2))}


def main():
    from oracle import ref_shims
    ref_shims.install()
    from pytracking import TensorList, dcf
    from pytracking.tracker.eco.optim import FilterOptim
    from pytracking.utils import TrackerParams

    out = {}
    for cname, case in CASES.items():
        g = torch.Generator().manual_seed(11)
        params = TrackerParams()
        params.fletcher_reeves = case["fletcher_reeves"]
        params.standard_alpha = case["standard_alpha"]
        params.direction_forget_factor = case["direction_forget_factor"]
        params.debug = 0
        params.precond_data_param = 0.3
        params.precond_reg_param = 0.15
        params.precond_learning_rate = TensorList([b["lr"] for b in BLOCKS])
        reg_filter, yf, samples, sw, filt = TensorList(), TensorList(), TensorList(), TensorList(), TensorList()
        for b in BLOCKS:
            fp = TrackerParams()
            fp.use_reg_window = True
            fp.reg_window_power = 2
            for k, v in b["reg"].items():
                setattr(fp, k, v)
            reg_filter.append(dcf.get_reg_filter(torch.Tensor([240., 240.]), torch.Tensor([50., 64.]), fp))
            yf.append(dcf.label_function(torch.Tensor([b["H"], 2 * b["Wh"] - 1]), torch.Tensor([1.0, 1.3])))
            samples.append(torch.zeros(b["H"], b["Wh"], N, b["C"], 2))
            sw.append(torch.zeros(N))
            filt.append(0.05 * torch.randn(1, b["C"], b["H"], b["Wh"], 2, generator=g))
        # three stored samples to start with (eco.py:141-158 leaves the unused slots zero with zero weight)
        for s, w_, b in zip(samples, sw, BLOCKS):
            s[:, :, :3] = torch.randn(b["H"], b["Wh"], 3, b["C"], 2, generator=g)
            w_[:3] = torch.tensor([0.5, 0.3, 0.2])
        reg_energy = reg_filter.view(-1) @ reg_filter.view(-1)
        opt = FilterOptim(params, reg_energy)
        opt.register(filt, samples, yf, sw, reg_filter)
        for bi, (b, rf) in enumerate(zip(BLOCKS, reg_filter)):
            out["%s/b%d/reg_filter" % (cname, bi)] = rf.numpy().copy()
            out["%s/b%d/yf" % (cname, bi)] = yf[bi].numpy().copy()
            out["%s/b%d/lr" % (cname, bi)] = np.float32(b["lr"])
        for run, it in enumerate(case["iters"]):
            new_xf = TensorList([torch.randn(1, b["C"], b["H"], b["Wh"], 2, generator=g) for b in BLOCKS])
            if run > 0:                                        # memory update of ECO.update_memory (eco.py:337-341): one slot per run
                for s, w_, xf in zip(samples, sw, new_xf):
                    slot = 2 + run
                    s[:, :, slot:slot + 1] = xf.permute(2, 3, 0, 1, 4)
                    w_ *= 0.9
                    w_[slot] = 0.1
                    w_ /= w_.sum()
            for bi in range(len(BLOCKS)):
                k = "%s/run%d/b%d/" % (cname, run, bi)
                out[k + "hf_in"] = filt[bi].numpy().copy()
                out[k + "samples"] = samples[bi].numpy().copy()
                out[k + "sw"] = sw[bi].numpy().copy()
                out[k + "new_xf"] = new_xf[bi].numpy().copy()
                out[k + "has_energy"] = np.int32(opt.sample_energy is not None)
                if opt.sample_energy is not None:
                    out[k + "energy_in"] = opt.sample_energy[bi].numpy().copy()
                out[k + "has_state"] = np.int32(opt.p is not None)
                if opt.p is not None:
                    out[k + "p_in"] = opt.p[bi].numpy().copy()
                    out[k + "rho_in"] = np.float32(float(opt.rho[bi]))
                    if opt.r_prev is not None:
                        out[k + "r_prev_in"] = opt.r_prev[bi].numpy().copy()
            opt.run(it, new_xf)
            for bi in range(len(BLOCKS)):
                k = "%s/run%d/b%d/" % (cname, run, bi)
                out[k + "num_iter"] = np.int32(it)
                out[k + "hf_out"] = filt[bi].numpy().copy()
                out[k + "energy_out"] = opt.sample_energy[bi].numpy().copy()
                out[k + "p_out"] = opt.p[bi].numpy().copy()
                out[k + "rho_out"] = np.float32(float(opt.rho[bi]))
                if opt.r_prev is not None:
                    out[k + "r_prev_out"] = opt.r_prev[bi].numpy().copy()
            # ECO.symmetrize_filter (eco.py:381-383) runs after every optimiser call in the tracker
            for hf in filt:
                hf[:, :, :, 0, :] /= 2
        out[cname + "/params"] = np.array([float(case["fletcher_reeves"]), float(case["standard_alpha"]),
                                           float(case["direction_forget_factor"]), 0.3, 0.15], dtype=np.float64)
    os.makedirs(GOLDEN, exist_ok=True)
    np.savez_compressed(os.path.join(GOLDEN, "eco_cg.npz"), **out)
    print("wrote eco_cg.npz with %d arrays" % len(out))


JOINT_BLOCKS = [dict(H=9, Wh=5, Cin=24, C=16, reg=dict(reg_window_min=1e-4, reg_window_edge=10e-3, reg_sparsity_threshold=0.05)),
                dict(H=7, Wh=4, Cin=80, C=64, reg=dict(reg_window_min=10e-4, reg_window_edge=50e-3, reg_sparsity_threshold=0.1))]
JOINT_N = 6


def joint():
    """First-frame joint optimisation: the reference's FactorizedConvProblem (eco/optim.py:8-117) under GaussNewtonCG with ECO's settings
    (eco.py:155-162: defaults of the optimiser = Fletcher-Reeves, state reset; init_CG_iter // init_GN_iter CG iterations per GN
    iteration), J and J^T by autograd as the reference does.  Writes tests/golden/eco_joint.npz."""
    from oracle import ref_shims
    ref_shims.install()
    from pytracking import TensorList, dcf
    from pytracking.libs.optimization import GaussNewtonCG
    from pytracking.tracker.eco.optim import FactorizedConvProblem
    from pytracking.utils import TrackerParams

    g = torch.Generator().manual_seed(23)
    params = TrackerParams()
    params.precond_data_param, params.precond_reg_param, params.precond_proj_param, params.projection_reg = 0.3, 0.15, 35, 5e-8
    X, yf, rf, P, hf, sw = TensorList(), TensorList(), TensorList(), TensorList(), TensorList(), TensorList()
    for b in JOINT_BLOCKS:
        fp = TrackerParams()
        fp.use_reg_window, fp.reg_window_power = True, 2
        for k, v in b["reg"].items():
            setattr(fp, k, v)
        rf.append(dcf.get_reg_filter(torch.Tensor([240., 240.]), torch.Tensor([50., 64.]), fp))
        yf.append(dcf.label_function(torch.Tensor([b["H"], 2 * b["Wh"] - 1]), torch.Tensor([1.0, 1.3])))
        # the tracker hands over a permuted view (eco.py:133); the values are what matters here
        X.append(torch.randn(b["H"], b["Wh"], JOINT_N, b["Cin"], 2, generator=g))
        P.append(torch.linalg.qr(torch.randn(b["Cin"], b["Cin"], generator=g))[0][:, :b["C"]].clone())    # eco.py:117-120: orthonormal columns
        hf.append(torch.zeros(1, b["C"], b["H"], b["Wh"], 2))                                             # eco.py:150-151
        sw.append(torch.ones(1) / JOINT_N)                                                                 # eco.py:132
    out = {}
    for bi in range(len(JOINT_BLOCKS)):
        k = "b%d/" % bi
        out[k + "samples"], out[k + "yf"], out[k + "reg_filter"] = X[bi].numpy().copy(), yf[bi].numpy().copy(), rf[bi].numpy().copy()
        out[k + "P_in"], out[k + "hf_in"] = P[bi].numpy().copy(), hf[bi].numpy().copy()
        out[k + "sample_weights"] = np.full(JOINT_N, 1.0 / JOINT_N, np.float32)
    prob = FactorizedConvProblem(X, yf, rf, P, params, sw)
    var = hf.concat(P)
    opt = GaussNewtonCG(prob, var, debug=False)
    nb = len(JOINT_BLOCKS)
    for bi in range(nb):
        k = "b%d/" % bi
        out[k + "diag_M_filter"] = prob.diag_M[bi].numpy().copy()                 # [1,C,H,Wh,1]
        out[k + "diag_M_proj"] = np.float32(float(prob.diag_M[nb + bi]))
        out[k + "sample_energy"] = prob.sample_energy[bi].numpy().copy()
    opt.run(4, 3)                                                                 # 3 GN iterations x 4 CG iterations
    for bi in range(nb):
        k = "b%d/" % bi
        out[k + "hf_out"], out[k + "P_out"] = var[bi].detach().numpy().copy(), var[nb + bi].detach().numpy().copy()
    out["params"] = np.array([4, 3, 5e-8, 0.3, 0.15, 35.0], dtype=np.float64)   # num_cg, num_gn, projection_reg, precond data / reg / proj
    np.savez_compressed(os.path.join(GOLDEN, "eco_joint.npz"), **out)
    print("wrote eco_joint.npz with %d arrays" % len(out))


# (block sizes [(H, Wh, C)], scales, output size): two blocks as ECO's default (a large shallow and a small deep one), an even and an odd
# output grid, one single-block case, and a grid exactly one step larger than the series
LOC_CASES = {"two_blocks_even": dict(blocks=[(15, 8, 8), (9, 5, 12)], S=3, out=(24, 24), weights=(1.0, 0.7)),
             "two_blocks_odd": dict(blocks=[(7, 4, 12), (13, 7, 6)], S=5, out=(25, 23), weights=(0.4, 1.0)),
             "one_block": dict(blocks=[(11, 6, 16)], S=2, out=(32, 20), weights=(1.0,)),
             "tight": dict(blocks=[(9, 5, 4), (5, 3, 8)], S=1, out=(10, 10), weights=(1.0, 1.0))}


def loc():
    """Score computation and localisation of `ECO.track` (eco.py:194-196, 244-274): `apply_filter` = complex.mult(filter, sample_xf).sum(1)
    per block, then score_fusion_strategy 'weightedsum': fourier.sample_fs(fourier.sum_fs(weight * sf), output_sz) and dcf.max2d -- the
    reference's own functions (irfft of the zero-padded series).  Writes tests/golden/eco_loc.npz."""
    from oracle import ref_shims
    ref_shims.install()
    from pytracking import TensorList, complex, dcf, fourier
    out = {}
    for name, c in LOC_CASES.items():
        g = torch.Generator().manual_seed(len(name))
        filt = TensorList([0.1 * torch.randn(1, C, H, Wh, 2, generator=g) for (H, Wh, C) in c["blocks"]])
        xf = TensorList([torch.randn(c["S"], C, H, Wh, 2, generator=g) for (H, Wh, C) in c["blocks"]])
        sf = complex.mult(filt, xf).sum(1, keepdim=True)                                     # ECO.apply_filter (eco.py:244-245)
        weight = TensorList(list(c["weights"]))
        scores = fourier.sample_fs(fourier.sum_fs(weight * sf), torch.Tensor(list(c["out"])))   # eco.py:250-252
        max_score, max_disp = dcf.max2d(scores)                                              # eco.py:271
        out[name + "/out"] = np.array(c["out"], np.int64)
        out[name + "/weights"] = np.array(c["weights"], np.float32)
        for b in range(len(filt)):
            out["%s/b%d/filter" % (name, b)] = filt[b].numpy()
            out["%s/b%d/xf" % (name, b)] = xf[b].numpy()
            out["%s/b%d/sf" % (name, b)] = sf[b].numpy()
        out[name + "/scores"] = scores.numpy()
        out[name + "/max_score"] = max_score.numpy()
        out[name + "/max_disp"] = max_disp.numpy()
    np.savez_compressed(os.path.join(GOLDEN, "eco_loc.npz"), **out)
    print("wrote eco_loc.npz: %d arrays, %.1f KB" % (len(out), os.path.getsize(os.path.join(GOLDEN, "eco_loc.npz")) / 1e3))

# feature map sizes (H, W), channels, samples: even (ECO's shallow 62x62 -> 63x32 in small), odd (deep 15x15 -> 15x8 in small), non-square
PREP_CASES = {"even": dict(sz=(10, 10), C=3, S=2), "odd": dict(sz=(7, 7), C=4, S=3), "rect": dict(sz=(6, 9), C=2, S=1)}


def prep():
    """`ECO.preprocess_sample` (eco.py:297-300) called unbound on a stand-in `self` that carries what ECO.initialize builds (eco.py:65-78):
    the Hann window of the feature size, the interpolation kernel's Fourier coefficients of the filter size (bicubic, a = -0.75, centred:
    parameter/eco/default.py:72-75).  x *= window (in place), fourier.cfft2, dcf.interpolate_dft.  Writes tests/golden/eco_prep.npz."""
    from oracle import ref_shims
    ref_shims.install()
    import types
    from pytracking import TensorList, dcf
    from pytracking.tracker.eco.eco import ECO
    out = {}
    names = list(PREP_CASES)
    g = torch.Generator().manual_seed(21)
    feature_sz = TensorList([torch.Tensor(list(PREP_CASES[n]["sz"])) for n in names])
    filter_sz = feature_sz + (feature_sz + 1) % 2
    me = types.SimpleNamespace(window=TensorList([dcf.hann2d(sz) for sz in feature_sz]),
                               interp_fs=TensorList([dcf.get_interp_fourier(sz, "bicubic", -0.75, True, False, "cpu") for sz in filter_sz]))
    x = TensorList([torch.randn(PREP_CASES[n]["S"], PREP_CASES[n]["C"], *PREP_CASES[n]["sz"], generator=g) for n in names])
    x_in = [e.clone() for e in x]
    xf = ECO.preprocess_sample(me, x)
    for i, n in enumerate(names):
        out[n + "/x"] = x_in[i].numpy()
        out[n + "/x_after"] = x[i].numpy()                         # the reference windows its argument in place
        out[n + "/window"] = me.window[i].numpy()
        out[n + "/interp_y"] = me.interp_fs[i][0].numpy()
        out[n + "/interp_x"] = me.interp_fs[i][1].numpy()
        out[n + "/xf"] = xf[i].numpy()
    np.savez_compressed(os.path.join(GOLDEN, "eco_prep.npz"), **out)
    print("wrote eco_prep.npz: %d arrays, %.1f KB; shapes %s" % (len(out), os.path.getsize(os.path.join(GOLDEN, "eco_prep.npz")) / 1e3,
                                                                   [tuple(e.shape) for e in xf]))


SHIFT_CASES = {"small": dict(shape=(2, 3, 7, 4), shift=(0.3, -1.1)), "large": dict(shape=(1, 5, 11, 6), shift=(-2.9, 3.0)),
               "one_axis": dict(shape=(3, 2, 5, 3), shift=(0.0, 0.7))}


def shift():
    """fourier.shift_fs (fourier.py:78-92) as ECO.track / ECO.initialize call it (eco.py:119-127, 226-227).  Writes tests/golden/eco_shift.npz."""
    from oracle import ref_shims
    ref_shims.install()
    from pytracking import fourier
    out = {}
    g = torch.Generator().manual_seed(5)
    for name, c in SHIFT_CASES.items():
        a = torch.randn(*c["shape"], 2, generator=g)
        out[name + "/a"] = a.numpy()
        out[name + "/shift"] = np.array(c["shift"], np.float64)
        out[name + "/out"] = fourier.shift_fs(a, shift=torch.Tensor(list(c["shift"]))).numpy()
    np.savez_compressed(os.path.join(GOLDEN, "eco_shift.npz"), **out)
    print("wrote eco_shift.npz: %d arrays" % len(out))


if __name__ == "__main__":
    if "shift" in sys.argv[1:]:
        shift()
    elif "loc" in sys.argv[1:] or "prep" in sys.argv[1:]:
        if "loc" in sys.argv[1:]:
            loc()
        if "prep" in sys.argv[1:]:
            prep()
    else:
        main()
        joint()
        loc()
        prep()
        shift()
