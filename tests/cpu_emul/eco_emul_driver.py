"""TEST INFRASTRUCTURE ONLY -- numpy / ctypes driver of the CPU-emulated ECO kernel (tests/cpu_emul/eco_emul.cpp), shared by
tests/test_eco_cpu.py and the ThreadSanitizer subprocess it spawns (which must not import torch)."""
import ctypes as C

import numpy as np


def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def run_emulated(lib, g, case, run, bi, max_ctas, block, force_res):
    fr, sa, dff, pdp, prp = g[case + "/params"]
    k = "%s/run%d/b%d/" % (case, run, bi)
    P = lambda a: a.ctypes.data_as(C.c_void_p) if a is not None else None
    hf = g[k + "hf_in"].copy()
    samples, sw, yf, rf, nx = g[k + "samples"], g[k + "sw"], g["%s/b%d/yf" % (case, bi)], g["%s/b%d/reg_filter" % (case, bi)], g[k + "new_xf"]
    _, cc, h, wh, _ = hf.shape
    has_e, has_s = int(g[k + "has_energy"]), int(g[k + "has_state"])
    en = g[k + "energy_in"].copy() if has_e else np.full((1, cc, h, wh), np.nan, np.float32)
    p = g[k + "p_in"].copy() if has_s else np.full_like(hf, np.nan)                 # unused buffers are poisoned
    rp = g[k + "r_prev_in"].copy() if (k + "r_prev_in") in g else np.full_like(hf, np.nan)
    rho = np.array([g[k + "rho_in"] if has_s else np.nan], np.float32)
    plan = (C.c_int * 6)()
    rc = lib.eco_emul_filter_cg(P(hf), P(samples), P(yf), P(sw), P(rf), rf.shape[2], rf.shape[3], P(en), has_e, P(nx), P(p), P(rp), P(rho),
                                has_s, h, wh, samples.shape[2], cc, int(g[k + "num_iter"]), int(fr), int(sa), C.c_float(dff),
                                C.c_float(float(g["%s/b%d/lr" % (case, bi)])), C.c_float(pdp), C.c_float(prp), max_ctas, block, force_res, plan)
    assert rc == 0
    errs = [rel(hf, g[k + "hf_out"]), rel(en, g[k + "energy_out"]), rel(p, g[k + "p_out"]),
            abs(float(rho[0]) - float(g[k + "rho_out"])) / abs(float(g[k + "rho_out"]))]
    if (k + "r_prev_out") in g:
        errs.append(rel(rp, g[k + "r_prev_out"]))
    return max(errs), list(plan), (hf, p, rho)




def run_emulated_joint(lib, g, bi, max_ctas, block, force_res):
    """First-frame joint optimisation (csrc/eco_joint_kernel.cuh) of block `bi` of tests/golden/eco_joint.npz on the emulated grid."""
    k = "b%d/" % bi
    P = lambda a: a.ctypes.data_as(C.c_void_p)
    num_cg, num_gn, lam = int(g["params"][0]), int(g["params"][1]), float(g["params"][2])
    hf, proj = g[k + "hf_in"].copy(), np.ascontiguousarray(g[k + "P_in"]).copy()
    samples, yf, rf = np.ascontiguousarray(g[k + "samples"]), g[k + "yf"], g[k + "reg_filter"]
    sw_sqrt = np.sqrt(g[k + "sample_weights"]).astype(np.float32)
    dmh = np.ascontiguousarray(g[k + "diag_M_filter"].reshape(g[k + "diag_M_filter"].shape[:4]))
    _, cc, h, wh, _ = hf.shape
    n, cin = samples.shape[2], samples.shape[3]
    plan = (C.c_int * 6)()
    rc = lib.eco_emul_joint_gn(P(hf), P(proj), P(samples), P(yf), P(sw_sqrt), P(rf), rf.shape[2], rf.shape[3], P(dmh),
                               C.c_float(float(g[k + "diag_M_proj"])), C.c_float(lam), h, wh, n, cin, cc, num_cg, num_gn, max_ctas, block,
                               force_res, plan)
    assert rc == 0
    return max(rel(hf, g[k + "hf_out"]), rel(proj, g[k + "P_out"])), list(plan), (hf, proj)
