"""TEST INFRASTRUCTURE ONLY -- records a short trajectory of the UNMODIFIED reference DiMP tracker (CPU, through
oracle/ref_shims.py) so that the GPU tests can replay the hot path inside a real tracker run.

Run in the build container (needs /root/reference):   python -m oracle.gen_track_golden

Configuration: `pytracking/parameter/dimp/dimp50.py` with plain-attribute overrides only -- CPU, random-init DiMP-50 built by
the reference constructor and loaded with `synth.make_dimp_state_dict('resnet50', seed=0, lut_seed=3)`, use_iou_net=False,
no init augmentation, filter_init_zero=True, update every frame with 2 SD iterations, target_not_found_threshold=-1e9
(random weights give small scores).  Recorded per frame: the crop request (centre, scale), the raw score map, the
localisation outcome, the memory update the tracker performed (slot, box, sample weights, iterations) and the filter
after it; nothing else is needed to replay the device-side work.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_shims  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")
NUM_FRAMES = 12
SEQ = 0


def main():
    ref_shims.install()
    torch.set_num_threads(8)
    from pytracking_b200 import synth
    from oracle import preprocessing_ref as mirror_pre
    import ltr.models.tracking.dimpnet as dimpnet
    from pytracking.parameter.dimp import dimp50 as dimp50_params
    from pytracking.tracker.dimp.dimp import DiMP
    from pytracking.features.net_wrappers import NetWithBackbone
    import pytracking.features.preprocessing as ref_pre

    # ---- network: reference constructor, weights = the seeded synthetic state dict the engine is built from ----
    torch.manual_seed(0)
    net = dimpnet.dimpnet50(filter_size=4, backbone_pretrained=False, optim_iter=5, clf_feat_norm=True, final_conv=True,
                            optim_init_step=0.9, optim_init_reg=0.1, init_gauss_sigma=0.9, num_dist_bins=100,
                            bin_displacement=0.1, mask_init_factor=3.0)
    sd = synth.make_dimp_state_dict("resnet50", seed=0, lut_seed=3)
    missing, unexpected = net.load_state_dict(sd, strict=False)
    assert not unexpected, unexpected
    assert all(k.startswith(("bb_regressor.", "feature_extractor.layer4", "feature_extractor.fc")) or "num_batches_tracked" in k
               for k in missing), [k for k in missing][:10]
    net.eval()

    params = dimp50_params.parameters()
    params.use_gpu = False
    params.device = "cpu"
    wrapper = NetWithBackbone(net_path="unused", use_gpu=False)
    wrapper.net = net                                     # NetWrapper.load_network is bypassed: the net is already built
    wrapper.load_network = lambda: None
    params.net = wrapper
    params.use_iou_net = False
    params.augmentation = {}
    params.use_augmentation = False
    params.filter_init_zero = True
    params.train_skipping = 1
    params.net_opt_update_iter = 2
    params.target_not_found_threshold = -1e9
    tracker = DiMP(params)

    frames, init_bbox = synth.make_sequence(SEQ, num_frames=NUM_FRAMES)
    rec = {"init_bbox": np.array(init_bbox, dtype=np.float32)}
    log = []

    # ---- recording hooks (attribute rebinding only) ----
    orig_extract = tracker.extract_backbone_features
    orig_classify = tracker.classify_target
    orig_update = tracker.update_classifier
    orig_memory = tracker.update_memory
    cur = {}

    def extract_hook(im, pos, scales, sz):
        cur.clear()
        cur["crop_pos"] = pos.clone().numpy()
        cur["crop_scale"] = np.array([float(s) for s in scales], dtype=np.float32)
        out = orig_extract(im, pos, scales, sz)
        # the host-side mirror of sample_patch must reproduce the reference crop bit for bit
        mine, coords = mirror_pre.sample_patch_multiscale(im, pos, scales, sz)
        assert torch.equal(mine, out[2]) and torch.equal(coords, out[1]), "preprocessing mirror differs from the reference"
        return out

    def classify_hook(x):
        s = orig_classify(x)
        cur["scores"] = s.detach().clone().numpy()
        return s

    def memory_hook(sample_x, target_box, learning_rate=None):
        orig_memory(sample_x, target_box, learning_rate)
        cur["replace_ind"] = int(tracker.previous_replace_ind[0])
        cur["target_box"] = target_box.clone().numpy()

    def update_hook(train_x, target_box, learning_rate=None, scores=None):
        orig_update(train_x, target_box, learning_rate, scores)
        n = int(tracker.num_stored_samples[0])
        cur["n_stored"] = n
        cur["sample_weights"] = tracker.sample_weights[0][:n].clone().numpy()
        cur["lr"] = -1.0 if learning_rate is None else float(learning_rate)
        cur["filter"] = tracker.target_filter.detach().clone().numpy()

    tracker.extract_backbone_features = extract_hook
    tracker.classify_target = classify_hook
    tracker.update_memory = memory_hook
    tracker.update_classifier = update_hook

    tracker.initialize(frames[0], {"init_bbox": init_bbox})
    rec["init_pos"] = tracker.init_sample_pos.clone().numpy()
    rec["init_scale"] = np.array(float(tracker.init_sample_scale), dtype=np.float32)
    rec["init_target_box"] = tracker.target_boxes[0].clone().numpy()
    rec["init_filter"] = tracker.target_filter.detach().clone().numpy()
    rec["init_num_iter"] = np.array(params.net_opt_iter)
    rec["aug_expansion_factor"] = np.array(float(params.augmentation_expansion_factor))
    # the mirror of the first-frame sample (expanded patch + Identity centre crop) must equal the reference's
    im0 = ref_pre.numpy_to_torch(frames[0])
    aug_sz = (tracker.img_sample_sz * params.augmentation_expansion_factor).long()
    aug_sz += (aug_sz - tracker.img_sample_sz.long()) % 2
    ref_patch = ref_pre.sample_patch_transformed(im0, tracker.init_sample_pos, tracker.init_sample_scale, aug_sz.float(), tracker.transforms[:1])
    mine = mirror_pre.sample_init_patch(im0, tracker.init_sample_pos, tracker.init_sample_scale, tracker.img_sample_sz,
                                       params.augmentation_expansion_factor)
    assert torch.equal(ref_patch, mine), "first-frame sample mirror differs from the reference"
    rec["img_sample_sz"] = tracker.img_sample_sz.clone().numpy()
    boxes, flags = [], []
    for t in range(1, NUM_FRAMES + 1):
        out = tracker.track(frames[t])
        boxes.append(out["target_bbox"])
        flags.append(tracker.debug_info["flag"])
        for k in ("crop_pos", "crop_scale", "scores"):
            rec["f%02d_%s" % (t, k)] = cur[k]
        updated = "filter" in cur
        rec["f%02d_updated" % t] = np.array(1 if updated else 0)
        if updated:
            for k in ("replace_ind", "target_box", "n_stored", "sample_weights", "lr"):
                rec["f%02d_%s" % (t, k)] = np.array(cur[k])
            rec["f%02d_filter" % t] = cur["filter"]
        log.append((t, flags[-1], [round(v, 2) for v in boxes[-1]], float(cur["scores"].max())))
    rec["boxes"] = np.array(boxes, dtype=np.float32)
    rec["flags"] = np.array(flags)
    os.makedirs(GOLDEN, exist_ok=True)
    np.savez_compressed(os.path.join(GOLDEN, "dimp_track.npz"), **rec)
    for l in log:
        print(l)


if __name__ == "__main__":
    main()
