"""The UNMODIFIED reference ECO tracker (pytracking/tracker/eco/eco.py, parameter/eco/default.py, seeded random-init ResNet18m1 features)
driven on the CPU over the synthetic sequence, once as it is and once with `plugin.install()` binding its two optimiser seams
(`GaussNewtonCG.run` on the first frame, `FilterOptim.run` every `train_skipping` frames) and its score-path seams (`ECO.preprocess_sample`,
`ECO.apply_filter`, `fourier.sample_fs` as `ECO.localize_target` calls it, `fourier.shift_fs`).  There is no GPU here, so the two library entry
points are replaced by the oracle (oracle/eco_oracle.py -- itself pinned to the reference's golden vectors, as is the CUDA kernel source
under the CPU shim, tests/test_eco_cpu.py): what this test pins is everything ABOVE the C ABI inside a real tracker run -- the tensors the
tracker actually hands over (permuted sample views, the column-major projection matrix out of torch.svd, one-element sample weights, the
aliased sample energy), the CG state carried from run to run in the reference's own attributes, `symmetrize_filter` on the tensors updated
in place -- by requiring the same boxes and filters as the stock run."""
import unittest.mock as um

import pytest
import torch

from baseline import ref_env

pytestmark = pytest.mark.skipif(not ref_env.reference_available(), reason="reference tree not staged (baseline/_ref)")

OTHER_SEAMS = ("ltr.models.layers.filter.apply_filter", "apply_feat_transpose", "max2d", "forward", "extract_backbone", "extract_classification_feat", "get_iou_feat",
               "predict_iou", "_prroi_pooling", "_import_prroi_pooling", "conv2d", "conv1x1", "extract_head_feat",
               "predict_cls_bbreg_filters_parallel", "ConjugateGradient.run", "softmax_reg")
OVERRIDES = dict(init_CG_iter=12, init_GN_iter=2, train_skipping=2, sample_memory_size=36)
FRAMES = 6


def _drive(frames, bb):
    from baseline import ref_tracker
    trk = ref_tracker.build_eco(overrides=OVERRIDES)
    torch.manual_seed(0)                                            # the dropout augmentation of the first frame (eco.py:322-326)
    trk.initialize(frames[0], {"init_bbox": list(bb)})
    return [trk.track(frames[i])["target_bbox"] for i in range(1, FRAMES + 1)], trk


def test_reference_eco_tracker_with_both_optimiser_seams_bound():
    from oracle import eco_oracle as E
    from pytracking_b200 import ops, plugin, synth
    frames, bb = synth.make_sequence(0, num_frames=FRAMES)
    ref_boxes, ref_trk = _drive(frames, bb)
    calls = {"cg": 0, "gn": 0, "apply": 0, "sample": 0, "prep": 0, "shift": 0}

    def oracle_cg(filt, samples, yf, sw, reg, energy, num_iter, new_xf=None, state=None, fletcher_reeves=False, standard_alpha=True,
                  direction_forget_factor=0.0, precond_learning_rate=0.0075, precond_data_param=0.3, precond_reg_param=0.15):
        calls["cg"] += 1
        assert filt.is_contiguous() and samples.is_contiguous() and tuple(samples.shape[:2]) == tuple(filt.shape[2:4])
        st = {} if not state else {"p": state["p"], "rho": float(state["rho"]), "r_prev": state.get("r_prev")}
        x, en, st2 = E.filter_optim_run(filt, samples, yf, sw, reg, energy, st, num_iter, new_xf, precond_learning_rate, precond_data_param,
                                        precond_reg_param, fletcher_reeves, standard_alpha, direction_forget_factor)
        filt.copy_(x)                                               # the library updates filter and energy in place
        if energy is not None:
            energy.copy_(en)
            en = energy
        return en, {"p": st2["p"].contiguous(), "r_prev": st2["r_prev"], "rho": torch.as_tensor(st2["rho"]).reshape(1).float()}

    def oracle_gn(filt, proj, samples, yf, sw_sqrt, reg, dMh, dMP, lam, ncg, ngn):
        calls["gn"] += 1
        assert proj.is_contiguous() and samples.is_contiguous() and sw_sqrt.numel() == samples.shape[2]
        # the library takes the problem's preconditioner as an argument: it must be the one the oracle derives from the same inputs
        dh, dp, _ = E.joint_precond(samples, proj, yf, reg, 0.3, 0.15, 35.0, lam)
        assert torch.allclose(dh.reshape(-1), dMh.reshape(-1), rtol=1e-4) and abs(float(dp) - dMP) < 1e-4 * dMP
        h, P, _ = E.joint_gn_run(filt, proj, samples, yf, sw_sqrt * sw_sqrt, reg, ncg, ngn, lam)
        filt.copy_(h)
        proj.copy_(P)
        return filt, proj

    def oracle_apply(filt, xf):
        calls["apply"] += 1
        assert filt.is_contiguous() and xf.is_contiguous()
        return E.apply_filter(filt, xf)

    def oracle_prep(x, window, iy, ix):
        calls["prep"] += 1
        xf, v = E.preprocess_sample(x, window, iy, ix)
        x.copy_(v)                                                  # the library windows its argument in place, as the reference does
        return xf

    def oracle_shift(a, sy, sx):
        calls["shift"] += 1
        assert a.is_contiguous()
        return E.shift_fs(a, (sy, sx))

    def oracle_sample(sf, out_sz, weights=None):
        calls["sample"] += 1
        return E.sample_fs(sf, out_sz)

    with um.patch.object(ops, "eco_filter_cg_", oracle_cg), um.patch.object(ops, "eco_joint_gn_", oracle_gn), \
            um.patch.object(ops, "eco_apply_filter", oracle_apply), um.patch.object(ops, "eco_sample_fs", oracle_sample), \
            um.patch.object(ops, "eco_preprocess_sample_", oracle_prep), \
            um.patch.object(ops, "eco_shift_fs", oracle_shift), \
            um.patch.object(plugin, "_inference", lambda *ts: all(isinstance(t, torch.Tensor) and t.dtype == torch.float32 for t in ts)), \
            um.patch.object(torch.Tensor, "is_cuda", property(lambda self: True)):
        plugin.install(skip=OTHER_SEAMS)
        served = dict(plugin.stats)
        try:
            boxes, trk = _drive(frames, bb)
        finally:
            plugin.uninstall()
    runs = sum(1 for f in range(2, FRAMES + 2) if f % OVERRIDES["train_skipping"] == 1)        # eco.py:244: frame_num % train_skipping == 1
    # two feature blocks per optimiser call and per ECO.apply_filter; one summed series per frame through sample_fs
    # preprocess_sample: two blocks on the first frame (the augmented samples) and on every tracked frame
    shifts = calls.pop("shift")                                     # eco.py:119-127 (augmentation shifts, sub-pixel position) and :226-227
    assert shifts >= 2 * FRAMES, shifts
    assert calls == {"gn": 2, "cg": 2 * runs, "apply": 2 * FRAMES, "sample": FRAMES, "prep": 2 * (FRAMES + 1)}, calls
    assert plugin.stats.get("GaussNewtonCG.run[eco]", 0) == served.get("GaussNewtonCG.run[eco]", 0) + 1
    assert plugin.stats.get("FilterOptim.run", 0) == served.get("FilterOptim.run", 0) + runs
    for a, b in zip(ref_boxes, boxes):
        assert max(abs(x - y) for x, y in zip(a, b)) < 1e-3, (a, b)
    for b in range(2):
        d = float((ref_trk.filter[b] - trk.filter[b]).abs().max() / ref_trk.filter[b].abs().max())
        assert d < 2e-5, (b, d)
        assert float((ref_trk.projection_matrix[b] - trk.projection_matrix[b]).abs().max()) < 2e-5


# every seam except GNSteepestDescent.forward stays on the reference in the DiMPSimple run below
NOT_GNSD = ("apply_filter", "apply_feat_transpose", "max2d", "extract_backbone", "extract_classification_feat", "get_iou_feat", "predict_iou",
            "_prroi_pooling", "_import_prroi_pooling", "conv2d", "conv1x1", "extract_head_feat", "predict_cls_bbreg_filters_parallel", "run",
            "softmax_reg", "DiMPSteepestDescentGN.forward", "PrDiMPSteepestDescentNewton.forward", "DiMPL2SteepestDescentGN.forward",
            "Transformer.forward", "DenseBoxRegressor.forward", "fourier", "preprocess_sample")


def test_reference_dimp_simple_tracker_with_the_gn_steepest_descent_seam_bound():
    """The UNMODIFIED reference DiMPSimple tracker (SuperDiMPSimple parameters, random-init dimpnet50_simple): its classifier's
    GNSteepestDescent(LinearFilterHinge) module is called by `LinearFilter.get_filter` with a bare weight tensor and 5-D features on the
    first frame and by `update_classifier` with a TensorList every `train_skipping` frames (dimp_simple.py:619,685).  Stock run vs
    `plugin.install()` with the oracle behind `ops.gn_sd_hinge` (the CUDA kernel is pinned to the same oracle on the GPU)."""
    from oracle import dimp_oracle as O
    from oracle import ref_shims
    ref_shims.install()
    from baseline import ref_tracker
    from pytracking_b200 import ops, plugin, synth
    frames, bb = synth.make_sequence(0, num_frames=4)
    ov = dict(train_skipping=2, use_iou_net=False)

    def drive():
        trk = ref_tracker.build_dimp_simple(overrides=ov)
        torch.manual_seed(0)
        trk.initialize(frames[0], {"init_bbox": list(bb)})
        return [trk.track(frames[i])["target_bbox"] for i in range(1, 5)], trk

    ref_boxes, ref_trk = drive()
    calls = []

    def oracle_op(weights, f, train_label, sample_weight, num_iter, filter_reg, hinge_threshold, activation_leak, score_act, act_param,
                  steplength_reg, return_iterates=False, compute_losses=False, out=None):
        calls.append((tuple(weights.shape), tuple(f.shape), num_iter))
        w, its, losses = O.gn_sd_hinge(weights, f, train_label[:, 0], sample_weight, filter_reg, num_iter, hinge_threshold, activation_leak,
                                       score_act, act_param, steplength_reg, compute_losses)
        return w, torch.cat(its, 0), torch.stack(losses) if compute_losses else None

    with um.patch.object(ops, "gn_sd_hinge", oracle_op), \
            um.patch.object(plugin, "_inference", lambda *ts: all(isinstance(t, torch.Tensor) and t.dtype == torch.float32 for t in ts)), \
            um.patch.object(torch.Tensor, "is_cuda", property(lambda self: True)):
        assert [n.split(".")[-2:] for n in plugin.install(skip=NOT_GNSD)] == [["GNSteepestDescent", "forward"]]
        try:
            boxes, trk = drive()
        finally:
            plugin.uninstall()
    assert len(calls) >= 3 and calls[0][2] == 10 and all(c[0] == (1, 512, 4, 4) and c[1][1:] == (512, 22, 22) for c in calls), calls
    for a, b in zip(ref_boxes, boxes):
        assert max(abs(x - y) for x, y in zip(a, b)) < 1e-3, (a, b)
    assert float((ref_trk.target_filter - trk.target_filter).abs().max() / ref_trk.target_filter.abs().max()) < 2e-5
