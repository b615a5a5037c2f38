"""Builds and drives the UNMODIFIED reference trackers (baseline/_ref, through baseline/ref_env.py) on seeded random-init networks
and the synthetic sequence of SURVEY.md 8(d).  Used by `bench.py --impl reference` (CPU arm), by bench.py's PyTorch-CUDA reference
timing, by the plug-in tests (reference tracker above the engine) and by the golden generators.

The frame loop restates `Tracker._track_sequence` (pytracking/evaluation/tracker.py:176-233): `initialize(image, info)`, then for
every frame `time.time()` around `tracker.track(image, info)` -- the reference's own clock.  (`Tracker.run_sequence` itself reads
frames from disk through cv2, which this image does not have, so the loop is restated around the unmodified tracker class.)
"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

# BASELINE configs[1] ("DiMP-50 single-GPU, synthetic 288x288 crops, 10 SD iters/frame"): parameter/dimp/dimp50.py plus the
# SURVEY.md 8(d) overrides (random-init scores are ~0.1, so the not-found test is disabled; update every frame with 10 iterations)
CONFIG_DIMP50 = dict(target_not_found_threshold=-1e9, train_skipping=1, net_opt_update_iter=10)


def build_dimp_net(arch="resnet50", seed=0):
    """Reference constructor (ltr/train_settings/dimp/dimp50.py:91-95 hyper-parameters) + the seeded synthetic weights the engine
    tests use for the backbone / clf head / optimiser; the IoUNet keeps the constructor's (seeded) random init."""
    from baseline import ref_env
    ref_env.install()
    from pytracking_b200 import synth
    import ltr.models.tracking.dimpnet as dimpnet
    torch.manual_seed(seed)
    ctor = dimpnet.dimpnet50 if arch == "resnet50" else dimpnet.dimpnet18
    net = ctor(filter_size=4, backbone_pretrained=False, optim_iter=5, clf_feat_norm=True, final_conv=True,
               optim_init_step=0.9, optim_init_reg=0.1, init_gauss_sigma=0.9, num_dist_bins=100,
               bin_displacement=0.1, mask_init_factor=3.0)
    sd = synth.make_dimp_state_dict(arch, seed=seed, lut_seed=3)
    missing, unexpected = net.load_state_dict(sd, strict=False)
    assert not unexpected, unexpected
    # the IoUNet of a random-init net predicts O(1e2) IoUs with O(1e3) box gradients; scale its last layer so that the refinement
    # steps are a few pixels, as with a trained net (the weights stay a deterministic function of `seed`)
    with torch.no_grad():
        net.bb_regressor.iou_predictor.weight.mul_(0.02)
    net.eval()
    return net


def build_dimp(device="cpu", arch="resnet50", use_iou_net=False, overrides=None, seed=0, use_augmentation=True, dropout=True):
    """-> reference `DiMP` tracker object (pytracking/tracker/dimp/dimp.py), parameters = parameter/dimp/dimp50.py + overrides.
    dropout=False removes the 'dropout' entry of params.augmentation: F.dropout2d draws from the DEVICE generator, so a CPU run and a
    CUDA run of the stock reference see different masks (and different filters) with it."""
    from baseline import ref_env
    ref_env.install()
    from pytracking.parameter.dimp import dimp50 as dimp50_params
    from pytracking.tracker.dimp.dimp import DiMP
    from pytracking.features.net_wrappers import NetWithBackbone
    net = build_dimp_net(arch, seed)
    use_gpu = (device != "cpu")
    if use_gpu:
        net = net.cuda()
    params = dimp50_params.parameters()
    params.use_gpu = use_gpu
    params.device = "cuda" if use_gpu else "cpu"
    wrapper = NetWithBackbone(net_path="unused", use_gpu=use_gpu)
    wrapper.net = net                    # NetWrapper.load_network reads a checkpoint file; the net is already built
    wrapper.load_network = lambda: None
    params.net = wrapper
    params.use_iou_net = use_iou_net
    params.use_augmentation = use_augmentation
    if not dropout:
        params.augmentation = {k: v for k, v in params.augmentation.items() if k != "dropout"}
    for k, v in dict(CONFIG_DIMP50, **(overrides or {})).items():
        setattr(params, k, v)
    return DiMP(params)


def run_sequence(tracker, frames, init_bbox, sync=None, seed=0, on_frame=None):
    """pytracking/evaluation/tracker.py:176-233 around an already constructed tracker.  `sync` (e.g. torch.cuda.synchronize) is
    called inside the timed bracket of every frame when given.  -> dict(target_bbox [T,4], time [T], init_time)"""
    torch.manual_seed(seed)              # DiMP.initialize draws random augmentation shifts / IoUNet proposal jitter
    np.random.seed(seed)
    t0 = time.time()
    tracker.initialize(frames[0], {"init_bbox": list(init_bbox)})
    if sync:
        sync()
    out = {"init_time": time.time() - t0, "target_bbox": [], "time": []}
    for t in range(1, len(frames)):
        start = time.time()
        o = tracker.track(frames[t], {})
        if sync:
            sync()
        out["time"].append(time.time() - start)
        out["target_bbox"].append(o["target_bbox"])
        if on_frame:
            on_frame(t, tracker, o)
    out["target_bbox"] = np.array(out["target_bbox"], dtype=np.float64)
    out["time"] = np.array(out["time"])
    return out


# ---------------------------------------------------------------------------------------------------------------------------
# the other BASELINE configurations (random-init networks from the reference constructors, seeded)
# ---------------------------------------------------------------------------------------------------------------------------
def _wrap(net, use_gpu):
    from pytracking.features.net_wrappers import NetWithBackbone
    if use_gpu:
        net = net.cuda()
    wrapper = NetWithBackbone(net_path="unused", use_gpu=use_gpu)
    wrapper.net = net
    wrapper.load_network = lambda: None
    return wrapper


def build_prdimp(device="cpu", use_iou_net=False, overrides=None, seed=0, use_augmentation=True, dropout=True):
    """BASELINE configs[2]: PrDiMP-50 (parameter/dimp/prdimp50.py; klcedimpnet50 with the ltr/train_settings/dimp/prdimp50.py
    hyper-parameters): 352^2 crops, 22x22 features, PrDiMPSteepestDescentNewton, softmax score pre-processing."""
    from baseline import ref_env
    ref_env.install()
    from pytracking_b200 import synth
    import ltr.models.tracking.dimpnet as dimpnet
    from pytracking.parameter.dimp import prdimp50 as P
    from pytracking.tracker.dimp.dimp import DiMP
    torch.manual_seed(seed)
    output_sigma = 1 / 4 / 6.0                                   # output_sigma_factor / search_area_factor (train settings)
    net = dimpnet.klcedimpnet50(filter_size=4, backbone_pretrained=False, optim_iter=5, clf_feat_norm=True, final_conv=True,
                                optim_init_step=1.0, optim_init_reg=0.05, optim_min_reg=0.05, gauss_sigma=output_sigma * 22,
                                alpha_eps=0.05, normalize_label=True, init_initializer="zero")
    sd = synth.make_dimp_state_dict("resnet50", seed=seed, lut_seed=3)
    sd = {k: v for k, v in sd.items() if not k.startswith(("classifier.filter_optimizer.", "classifier.filter_initializer."))}
    missing, unexpected = net.load_state_dict(sd, strict=False)
    assert not unexpected, unexpected
    with torch.no_grad():
        net.bb_regressor.iou_predictor.weight.mul_(0.02)
    net.eval()
    use_gpu = device != "cpu"
    params = P.parameters()
    params.use_gpu, params.device = use_gpu, ("cuda" if use_gpu else "cpu")
    params.net = _wrap(net, use_gpu)
    params.use_iou_net, params.use_augmentation = use_iou_net, use_augmentation
    if not dropout:
        params.augmentation = {k: v for k, v in params.augmentation.items() if k != "dropout"}
    # (filter_init_zero: FilterInitializerLinear pools with PrRoIPool even with init_weights='zero'; a stock checkout cannot run it)
    for k, v in dict(dict(target_not_found_threshold=-1e9, train_skipping=1, net_opt_update_iter=10, filter_init_zero=True),
                     **(overrides or {})).items():
        setattr(params, k, v)
    return DiMP(params)


def build_atom(device="cpu", overrides=None, seed=5):
    """BASELINE configs[0]: ATOM ResNet-18 (parameter/atom/multiscale_no_iounet.py: 5 scales, no IoUNet), first-frame GaussNewtonCG
    on FactorizedConvProblem, per-frame ConjugateGradient on ConvProblem."""
    from baseline import ref_env
    ref_env.install()
    from pytracking_b200 import synth
    import ltr.models.bbreg.atom as atom_models
    from pytracking.parameter.atom import multiscale_no_iounet as P
    import pytracking.tracker.atom.atom as atom_mod
    import pytracking.features.deep as deep
    torch.manual_seed(1234)
    net = atom_models.atom_resnet18(backbone_pretrained=False)
    missing, unexpected = net.load_state_dict(synth.make_backbone_state_dict("resnet18", seed=seed), strict=False)
    assert not unexpected, unexpected
    net.eval()
    deep.load_network = lambda path: net
    use_gpu = device != "cpu"
    params = P.parameters()
    params.use_gpu, params.device = use_gpu, ("cuda" if use_gpu else "cpu")
    params.features.features[0].use_gpu = use_gpu
    for k, v in dict(dict(train_skipping=1, target_not_found_threshold=-1e9), **(overrides or {})).items():
        setattr(params, k, v)
    return atom_mod.ATOM(params)


def build_tomp(device="cpu", overrides=None, seed=0):
    """BASELINE configs[3]: ToMP-101 (parameter/tomp/tomp101.py; tompnet101 with the ltr/train_settings/tomp/tomp101.py arguments)."""
    from baseline import ref_env
    ref_env.install()
    from pytracking_b200 import synth
    import ltr.models.tracking.tompnet as tompnet
    from pytracking.parameter.tomp import tomp101 as P
    from pytracking.tracker.tomp.tomp import ToMP
    torch.manual_seed(seed)
    net = tompnet.tompnet101(filter_size=1, backbone_pretrained=False, head_feat_blocks=0, head_feat_norm=True, final_conv=True,
                             out_feature_dim=256, feature_sz=18, frozen_backbone_layers=(), num_encoder_layers=6,
                             num_decoder_layers=6, use_test_frame_encoding=False)
    missing, unexpected = net.load_state_dict(synth.make_backbone_state_dict("resnet101", seed=seed), strict=False)
    assert not unexpected, unexpected
    net.eval()
    use_gpu = device != "cpu"
    params = P.parameters()
    params.use_gpu, params.device = use_gpu, ("cuda" if use_gpu else "cpu")
    params.net = _wrap(net, use_gpu)
    for k, v in dict(dict(target_not_found_threshold=-1e9), **(overrides or {})).items():
        setattr(params, k, v)
    return ToMP(params)


def build_eco(device="cpu", overrides=None, seed=7):
    """The reference ECO tracker (pytracking/tracker/eco/eco.py with parameter/eco/default.py: ResNet18m1 features `vggconv1` + `layer3`,
    compressed to 16 + 64 channels, memory 200, first-frame GaussNewtonCG on FactorizedConvProblem, FilterOptim.run every
    `train_skipping` frames) on a seeded random-init `resnet18_vggmconv1` (the checkpoint file cannot be fetched here)."""
    from baseline import ref_env
    ref_env.install()
    import ltr.models.backbone.resnet18_vggm as rv
    import pytracking.features.deep as deep
    import pytracking.tracker.eco.eco as eco_mod
    from pytracking.parameter.eco import default as P

    def random_init(output_layers=None, path=None, **kwargs):        # deep.ResNet18m1.initialize calls this with a checkpoint path
        torch.manual_seed(seed)
        return rv.resnet18_vggmconv1(output_layers, path=None, **kwargs)
    deep.resnet18_vggmconv1 = random_init
    use_gpu = device != "cpu"
    params = P.parameters()
    params.use_gpu, params.device = use_gpu, ("cuda" if use_gpu else "cpu")
    params.features.features[0].use_gpu = use_gpu
    for k, v in (overrides or {}).items():
        setattr(params, k, v)
    return eco_mod.ECO(params)


def build_dimp_simple(device="cpu", overrides=None, seed=0):
    """The reference DiMPSimple tracker (pytracking/tracker/dimp_simple/dimp_simple.py with parameter/dimp_simple/super_dimp_simple.py) on a
    seeded random-init `dimpnet50_simple` built with the arguments of ltr/train_settings/dimp/super_dimp_simple.py:104-107: the classifier
    whose online optimiser is GNSteepestDescent over LinearFilterHinge (also KeepTrack's base tracker)."""
    from baseline import ref_env
    ref_env.install()
    from pytracking_b200 import synth
    import ltr.models.tracking.dimpnet as dimpnet
    from pytracking.parameter.dimp_simple import super_dimp_simple as P
    from pytracking.tracker.dimp_simple.dimp_simple import DiMPSimple
    torch.manual_seed(seed)
    net = dimpnet.dimpnet50_simple(filter_size=4, backbone_pretrained=False, optim_iter=5, clf_feat_norm=True, clf_feat_blocks=0,
                                   final_conv=True, out_feature_dim=512, optim_init_reg=0.1, score_act='relu', hinge_threshold=0.05,
                                   activation_leak=0.1, frozen_backbone_layers=['conv1', 'bn1', 'layer1', 'layer2'])
    missing, unexpected = net.load_state_dict(synth.make_backbone_state_dict("resnet50", seed=seed), strict=False)
    assert not unexpected, unexpected
    net.eval()
    use_gpu = device != "cpu"
    params = P.parameters()
    params.use_gpu, params.device = use_gpu, ("cuda" if use_gpu else "cpu")
    params.net = _wrap(net, use_gpu)
    for k, v in dict(dict(target_not_found_threshold=-1e9), **(overrides or {})).items():
        setattr(params, k, v)
    return DiMPSimple(params)
