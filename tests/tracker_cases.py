"""Tracker parameter sets and golden-trajectory helpers shared by the CPU host-logic test and the GPU tracker tests."""
import os

from pytracking_b200 import _lib

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
DIMP50 = dict(image_sample_size=288, search_area_scale=5, sample_memory_size=50, learning_rate=0.01, init_samples_minimum_weight=0.25,
              train_skipping=20, update_classifier=True, net_opt_iter=10, net_opt_update_iter=2, net_opt_hn_iter=1,
              advanced_localization=True, target_not_found_threshold=0.25, distractor_threshold=0.8, hard_negative_threshold=0.5,
              target_neighborhood_scale=2.2, dispalcement_scale=0.8, hard_negative_learning_rate=0.02, augmentation_expansion_factor=2,
              use_iou_net=False)
OVERRIDES = {
    "cfg2": dict(target_not_found_threshold=-1e9, train_skipping=1, net_opt_update_iter=10),
    "stress": dict(target_not_found_threshold=0.052, uncertain_threshold=0.0555, hard_sample_threshold=0.058, distractor_threshold=0.3,
                   hard_negative_threshold=0.07, train_skipping=3, sample_memory_size=24, net_opt_update_iter=2, dispalcement_scale=0.25),
    "noaug": dict(target_not_found_threshold=-1e9, train_skipping=1, net_opt_update_iter=2),
}


def _loc(d, t):
    loc = _lib.LocResult()
    loc.flag, loc.scale_ind = int(d["flag"][t]), 0
    loc.score1, loc.r1, loc.c1 = float(d["m1"][t, 0]), int(d["m1"][t, 1]), int(d["m1"][t, 2])
    loc.score2, loc.r2, loc.c2 = float(d["m2"][t, 0]), int(d["m2"][t, 1]), int(d["m2"][t, 2])
    loc.use_second = int(d["use2"][t])
    loc.max_score = loc.score1
    return loc


