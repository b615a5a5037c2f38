"""The host half of the native whole-frame tracker (csrc/dimp_tracker.cu: init_state / plan_crop / commit -- pure host float32
arithmetic behind the C ABI, no GPU) replayed against trajectories of the UNMODIFIED reference DiMP tracker recorded by
oracle/gen_host_logic_golden.py: fed with the reference's own localisation results it must reproduce, bit for bit, every crop
request (sample_patch's patch_coord), every output box, every stored target box, the replaced memory slots and the optimiser
schedule; the sample weights to float32 rounding (torch's vectorised sum order is not restated)."""
import os

import numpy as np
import pytest

from pytracking_b200 import _lib
from pytracking_b200.tracker import HostLogic, make_params

from tracker_cases import DIMP50, GOLDEN, OVERRIDES, _loc


@pytest.mark.parametrize("name", ["cfg2", "stress", "noaug"])
def test_host_logic_replays_reference_trajectory(name):
    d = np.load(os.path.join(GOLDEN, "dimp_host_%s.npz" % name))
    hl = HostLogic(make_params(**dict(DIMP50, **OVERRIDES[name])))
    H, W = [int(v) for v in d["image_hw"]]
    if name == "noaug":
        # native initialisation: scalar state, first-frame crop request and init target box from init_bbox alone
        g, box = hl.init_state(H, W, d["init_bbox"])
        assert np.array_equal(hl.state(), d["init_state"])
        assert np.array_equal(box, d["init_target_box"])
        assert (g.out_h, g.win_r, g.win_c) == (576, 144, 144)
        assert np.array_equal(np.array(g.sample_pos), d["init_sample_pos"]) and np.float32(g.sample_scale) == d["init_sample_scale"]
    else:
        hl.adopt(H, W, d["init_state"], d["init_sw"], d["init_counts"][0], d["init_counts"][1])
    T = len(d["flag"])
    max_sw_err = 0.0
    for t in range(T):
        g = hl.plan_crop()
        assert np.array_equal(np.array(g.coord, dtype=np.float32), d["coord"][t]), (name, t, list(g.coord), d["coord"][t])
        info, sw = hl.commit(g, _loc(d, t))
        assert np.array_equal(np.array(info.bbox, dtype=np.float32), d["bbox"][t]), (name, t)
        assert np.array_equal(hl.state()[:5], d["state"][t]), (name, t)
        assert info.updated == d["updated"][t], (name, t)
        if info.updated:
            assert info.replace_ind == d["replace_ind"][t], (name, t)
            assert np.array_equal(np.array(info.target_box, dtype=np.float32), d["target_box"][t]), (name, t)
            assert np.float32(info.learning_rate) == np.float32(d["lr"][t])
            max_sw_err = max(max_sw_err, float(np.abs(sw - d["sw"][t]).max() / d["sw"][t].max()))
        assert info.num_iter == d["num_iter"][t], (name, t, info.num_iter, d["num_iter"][t])
        if info.num_iter:
            assert info.n_stored == d["n_stored"][t]
    assert max_sw_err < 2e-6, max_sw_err
    hl.close()


def test_crop_geometry_edge_cases():
    """Integer pre-decimation (df > 1), crops hanging over every image border, tiny targets (size clamp 2)."""
    import torch
    from oracle import preprocessing_ref as R
    for (H, W, bb) in [(480, 640, [300, 200, 80, 60]), (720, 1280, [2, 3, 400, 300]), (1080, 1920, [1500, 900, 410, 170]),
                       (240, 320, [310, 230, 9, 9]), (480, 640, [100, 100, 3, 2]), (2160, 3840, [1000, 1000, 1500, 900])]:
        hl = HostLogic(make_params(**dict(DIMP50, augmentation_expansion_factor=None)))
        hl.init_state(H, W, bb)
        g = hl.plan_crop()
        st = hl.state()
        im = torch.zeros(1, 3, H, W)
        pos, scale = torch.tensor([st[0], st[1]]), torch.tensor(st[4])
        ref_geom = R.sample_patch_geometry(im, pos, scale * torch.tensor([288.0, 288.0]), torch.tensor([288.0, 288.0]))
        assert (g.df, g.os_r, g.os_c, g.tl_r, g.tl_c, g.in_h, g.in_w) == ref_geom[:7], ((H, W, bb), ref_geom)
        assert np.array_equal(np.array(g.coord, dtype=np.float32), ref_geom[7].numpy().reshape(4))
        hl.close()
