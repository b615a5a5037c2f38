"""Host-side logic above the C ABI against the recorded trajectory of the unmodified reference tracker (no GPU needed):
the sample-weight bookkeeping of DiMP.update_sample_weights (pytracking/tracker/dimp/dimp.py:445-484) mirrored by
pytracking_b200.frame_engine.SampleWeights, and the oracle's full per-frame path (backbone + head + classify + update)
replayed open loop from the reference's own filters."""
import os

import numpy as np
import pytest
import torch

from oracle import dimp_oracle as O
from oracle import preprocessing_ref as pre
from pytracking_b200 import synth


def test_sample_weights_follow_the_reference_trajectory(golden_dir):
    from pytracking_b200.frame_engine import SampleWeights
    g = np.load(os.path.join(golden_dir, "dimp_track.npz"))
    swm = SampleWeights(50, 1, learning_rate=0.01, init_samples_minimum_weight=0.25)
    updates = 0
    for t in range(1, 13):
        k = "f%02d_" % t
        if not int(g[k + "updated"]):
            continue
        lr = float(g[k + "lr"])
        r = swm.step(None if lr < 0 else lr)
        n = int(g[k + "n_stored"])
        assert r == int(g[k + "replace_ind"]) and n == swm.num_stored
        assert np.allclose(swm.w[:n], g[k + "sample_weights"], rtol=1e-6, atol=1e-9)
        assert abs(float(swm.w[:n].sum()) - 1.0) < 1e-5
        updates += 1
    assert updates >= 8


def test_oracle_replays_the_reference_tracker_frames(golden_dir):
    """Open-loop replay of the first tracked frames through the CPU oracle: reference crop requests -> bit-exact crops ->
    oracle backbone / head / classify with the reference's filter of the previous frame -> the reference's score map and arg-max."""
    g = np.load(os.path.join(golden_dir, "dimp_track.npz"))
    frames, init_bbox = synth.make_sequence(0, num_frames=3)
    assert np.allclose(init_bbox, g["init_bbox"])
    sd = synth.make_dimp_state_dict("resnet50", seed=0, lut_seed=3)
    sz = torch.from_numpy(g["img_sample_sz"])
    prev_filter = torch.from_numpy(g["init_filter"])
    for t in (1, 2):
        k = "f%02d_" % t
        im = pre.numpy_to_torch(frames[t])
        crop, _ = pre.sample_patch_multiscale(im, torch.from_numpy(g[k + "crop_pos"]), [float(s) for s in g[k + "crop_scale"]], sz)
        with torch.no_grad():
            bf = O.resnet_forward(sd, O.preprocess_image(crop), "resnet50")
            clf = O.clf_head_dimp50(sd, bf["layer3"])
            scores = O.apply_filter(clf, prev_filter)
        ref = torch.from_numpy(g[k + "scores"]).reshape(scores.shape)
        assert float((scores - ref).abs().max() / ref.abs().max()) < 1e-4
        _, mi = O.max2d(scores.reshape(1, 19, 19))
        _, mi_ref = O.max2d(ref.reshape(1, 19, 19))
        assert mi.tolist() == mi_ref.tolist()
        if int(g[k + "updated"]):
            prev_filter = torch.from_numpy(g[k + "filter"])
