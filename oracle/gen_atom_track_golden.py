"""TEST INFRASTRUCTURE ONLY -- records a short trajectory of the UNMODIFIED reference ATOM tracker (BASELINE configs[0]:
ATOM ResNet-18, one synthetic sequence, reference PyTorch CPU path) for the GPU replay test.

Run in the build container (needs /root/reference):   python -m oracle.gen_atom_track_golden

Configuration: `pytracking/parameter/atom/multiscale_no_iounet.py` (5 scales, no IoUNet) with plain-attribute overrides only --
CPU, `atom_resnet18` built by the reference constructor with the seeded synthetic backbone weights, no init augmentation,
filter update every frame (train_skipping = 1), target_not_found_threshold = -1e9.
Recorded: the first-frame joint optimisation (initial filter / projection matrix, label, result), and per frame the crop
request, raw score maps, arg-max of the Fourier-upsampled maps per scale, the memory update and the filter after the CG run.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_shims  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")
NUM_FRAMES = 8
SEQ = 1
BACKBONE_SEED = 5


def main():
    ref_shims.install()
    torch.set_num_threads(8)
    torch.manual_seed(1234)
    from pytracking_b200 import synth
    from oracle import preprocessing_ref as mirror_pre
    import ltr.models.bbreg.atom as atom_models
    from pytracking.parameter.atom import multiscale_no_iounet as atom_params
    import pytracking.tracker.atom.atom as atom_mod
    import pytracking.features.deep as deep
    import pytracking.features.preprocessing as ref_pre
    from pytracking.libs import fourier, dcf

    net = atom_models.atom_resnet18(backbone_pretrained=False)
    sd = synth.make_backbone_state_dict("resnet18", seed=BACKBONE_SEED)
    missing, unexpected = net.load_state_dict(sd, strict=False)
    assert not unexpected, unexpected
    assert all(k.startswith(("bb_regressor.", "feature_extractor.layer4", "feature_extractor.fc")) or "num_batches_tracked" in k for k in missing)
    net.eval()
    deep.load_network = lambda path: net

    params = atom_params.parameters()
    params.use_gpu = False
    params.device = "cpu"
    params.features.features[0].use_gpu = False
    params.augmentation = {}
    params.train_skipping = 1
    params.target_not_found_threshold = -1e9
    tracker = atom_mod.ATOM(params)

    frames, init_bbox = synth.make_sequence(SEQ, num_frames=NUM_FRAMES)
    rec = {"init_bbox": np.array(init_bbox, dtype=np.float32)}
    cur = {}

    # ---- hooks ----
    class RecGN(atom_mod.GaussNewtonCG):
        def run(self, num_cg_iter, num_gn_iter=None):
            rec["init_w0"] = self.x[0].detach().clone().numpy()
            rec["init_P0"] = self.x[1].detach().clone().numpy()
            rec["init_num_cg"] = np.array(num_cg_iter)
            rec["init_num_gn"] = np.array(num_gn_iter)
            out = super().run(num_cg_iter, num_gn_iter)
            rec["init_w"] = self.x[0].detach().clone().numpy()
            rec["init_P"] = self.x[1].detach().clone().numpy()
            return out
    atom_mod.GaussNewtonCG = RecGN

    orig_extract = params.features.extract

    def extract_hook(im, pos, scales, image_sz, return_patches=False):
        cur.clear()
        cur["crop_pos"] = pos.clone().numpy()
        cur["crop_scales"] = np.array([float(s) for s in scales], dtype=np.float32)
        mine, _ = mirror_pre.sample_patch_multiscale(im, pos, scales, image_sz)
        ref_p = torch.cat([ref_pre.sample_patch(im, pos, s * image_sz, image_sz)[0] for s in scales])
        assert torch.equal(mine, ref_p), "preprocessing mirror differs from the reference"
        return orig_extract(im, pos, scales, image_sz, return_patches)
    params.features.extract = extract_hook

    orig_apply = tracker.apply_filter

    def apply_hook(sample_x):
        s = orig_apply(sample_x)
        cur["scores_raw"] = s[0].detach().clone().numpy()
        return s
    tracker.apply_filter = apply_hook

    orig_sample_fs = fourier.sample_fs

    def sample_fs_hook(a, grid_sz=None, rescale=True):
        out = orig_sample_fs(a, grid_sz, rescale)
        mv, mi = dcf.max2d(out)
        cur["up_maxval"] = mv.reshape(-1).clone().numpy()
        cur["up_maxidx"] = mi.reshape(-1, 2).clone().numpy()
        return out
    fourier.sample_fs = sample_fs_hook

    orig_memory = tracker.update_memory

    def memory_hook(sample_x, sample_y, learning_rate=None):
        orig_memory(sample_x, sample_y, learning_rate)
        cur["replace_ind"] = int(tracker.previous_replace_ind[0])
        cur["train_y"] = sample_y[0].clone().numpy()
        cur["lr"] = -1.0 if learning_rate is None else float(learning_rate)
        cur["sample_weights"] = tracker.sample_weights[0].clone().numpy()
    tracker.update_memory = memory_hook

    tracker.initialize(frames[0], {"init_bbox": init_bbox})
    rec["init_pos"] = tracker.pos.clone().numpy()
    rec["init_scale"] = np.array(float(tracker.target_scale), dtype=np.float32)
    rec["img_sample_sz"] = tracker.img_sample_sz.clone().numpy()
    rec["aug_expansion_factor"] = np.array(float(params.augmentation_expansion_factor))
    rec["init_y"] = tracker.y[0][:1].clone().numpy()
    rec["filter_reg"] = np.array(float(tracker.filter_reg[0]))
    rec["projection_reg"] = np.array(float(tracker.fparams.attribute("projection_reg")[0]))
    rec["output_sz"] = tracker.output_sz.clone().numpy()
    assert tuple(rec["init_w"].shape) == (1, 64, 4, 4)
    # the mirror of the first-frame sample must equal the reference's
    im0 = ref_pre.numpy_to_torch(frames[0])
    aug_sz = (tracker.img_sample_sz * params.augmentation_expansion_factor).long()
    aug_sz += (aug_sz - tracker.img_sample_sz.long()) % 2
    ref_patch, _ = ref_pre.sample_patch(im0, torch.from_numpy(rec["init_pos"]), tracker.target_scale * aug_sz.float(), aug_sz.float())
    ref_patch = tracker.transforms[0](ref_patch)
    mine = mirror_pre.sample_init_patch(im0, torch.from_numpy(rec["init_pos"]), tracker.target_scale, tracker.img_sample_sz,
                                       params.augmentation_expansion_factor)
    assert torch.equal(ref_patch, mine), "first-frame sample mirror differs from the reference"

    orig_run = tracker.filter_optimizer.run
    boxes = []
    for t in range(1, NUM_FRAMES + 1):
        ran = {}

        def run_hook(n, _ran=ran):
            _ran["iters"] = int(n)
            return orig_run(n)
        tracker.filter_optimizer.run = run_hook
        out = tracker.track(frames[t])
        boxes.append(out["target_bbox"])
        k = "f%02d_" % t
        for name in ("crop_pos", "crop_scales", "scores_raw", "up_maxval", "up_maxidx"):
            rec[k + name] = cur[name]
        rec[k + "flag"] = np.array(tracker.debug_info["flag"])
        rec[k + "updated"] = np.array(1 if "replace_ind" in cur else 0)
        if "replace_ind" in cur:
            for name in ("replace_ind", "train_y", "lr", "sample_weights"):
                rec[k + name] = np.array(cur[name])
        rec[k + "cg_iters"] = np.array(ran.get("iters", 0))
        rec[k + "filter"] = tracker.filter[0].detach().clone().numpy()
        print(t, tracker.debug_info["flag"], [round(v, 2) for v in boxes[-1]], float(cur["up_maxval"].max()), ran.get("iters", 0))
    rec["boxes"] = np.array(boxes, dtype=np.float32)
    np.savez_compressed(os.path.join(GOLDEN, "atom_track.npz"), **rec)


if __name__ == "__main__":
    main()
