"""Seeded synthetic networks, crops and sequences (no datasets / checkpoints are reachable here).

Everything is generated on the CPU torch generator so that the reference (in the build
container), the oracle and the engine (on the GPU box) see bit-identical inputs.

The state-dict key names are the reference's own (`DiMPnet.state_dict()`):
  feature_extractor.*                        ltr/models/backbone/resnet.py:96-110
  classifier.feature_extractor.0.weight      ltr/models/target_classifier/features.py:66
  classifier.filter_initializer.filter_conv  ltr/models/target_classifier/initializer.py:133
  classifier.filter_optimizer.*              ltr/models/target_classifier/optimizer.py:40-70
so `reference_net.load_state_dict(sd, strict=False)` takes them unchanged.
"""
import math
from collections import OrderedDict

import numpy as np
import torch

# (block type, blocks per layer) -- ltr/models/backbone/resnet.py:242-291
RESNET_ARCH = {
    "resnet18": ("basic", [2, 2, 2, 2]),
    "resnet50": ("bottleneck", [3, 4, 6, 3]),
    "resnet101": ("bottleneck", [3, 4, 23, 3]),
}


def _gen(seed):
    g = torch.Generator()
    g.manual_seed(int(seed))
    return g


def _conv_w(g, cout, cin, k, gain=1.0):
    # fan-out He init as the reference does (resnet.py:131-134), values drawn from our own generator
    std = gain * math.sqrt(2.0 / (k * k * cout))
    return torch.randn(cout, cin, k, k, generator=g) * std


def _bn(g, sd, prefix, c, gamma_scale=1.0):
    # non-trivial eval-mode statistics so that BN folding is actually exercised
    sd[prefix + ".weight"] = (0.8 + 0.4 * torch.rand(c, generator=g)) * gamma_scale
    sd[prefix + ".bias"] = 0.1 * torch.randn(c, generator=g)
    sd[prefix + ".running_mean"] = 0.1 * torch.randn(c, generator=g)
    sd[prefix + ".running_var"] = 0.6 + 0.8 * torch.rand(c, generator=g)
    sd[prefix + ".num_batches_tracked"] = torch.tensor(0, dtype=torch.long)


def make_backbone_state_dict(arch="resnet50", seed=0, prefix="feature_extractor.", last_layer=3):
    """Random-init ResNet weights for conv1..layer{last_layer} in torchvision/reference naming."""
    block, layers = RESNET_ARCH[arch]
    g = _gen(seed)
    sd = OrderedDict()
    sd[prefix + "conv1.weight"] = _conv_w(g, 64, 3, 7)
    _bn(g, sd, prefix + "bn1", 64)
    inplanes = 64
    for li in range(last_layer):
        planes = 64 * (2 ** li)
        for bi in range(layers[li]):
            p = "%slayer%d.%d." % (prefix, li + 1, bi)
            stride = 2 if (li > 0 and bi == 0) else 1
            if block == "bottleneck":
                sd[p + "conv1.weight"] = _conv_w(g, planes, inplanes, 1)
                _bn(g, sd, p + "bn1", planes)
                sd[p + "conv2.weight"] = _conv_w(g, planes, planes, 3)
                _bn(g, sd, p + "bn2", planes)
                sd[p + "conv3.weight"] = _conv_w(g, planes * 4, planes, 1)
                _bn(g, sd, p + "bn3", planes * 4, gamma_scale=0.5)
                outp = planes * 4
            else:
                sd[p + "conv1.weight"] = _conv_w(g, planes, inplanes, 3)
                _bn(g, sd, p + "bn1", planes)
                sd[p + "conv2.weight"] = _conv_w(g, planes, planes, 3)
                _bn(g, sd, p + "bn2", planes, gamma_scale=0.5)
                outp = planes
            if bi == 0 and (stride != 1 or inplanes != outp):
                sd[p + "downsample.0.weight"] = _conv_w(g, outp, inplanes, 1)
                _bn(g, sd, p + "downsample.1", outp)
            inplanes = outp
    return sd


def make_dimp_optimizer_params(num_dist_bins=100, bin_displacement=0.1, init_gauss_sigma=0.9,
                               mask_init_factor=3.0, init_step=0.9, init_reg=0.1, seed=None):
    """Learned scalars + the three 100-bin radial LUTs of DiMPSteepestDescentGN, initialised as the
    reference constructor does (optimizer.py:40-70; hyper-parameters ltr/train_settings/dimp/dimp50.py:91-95).
    With `seed` the LUTs are additionally perturbed (a trained net has non-trivial LUTs)."""
    d = torch.arange(num_dist_bins, dtype=torch.float32) * bin_displacement
    init_gauss = torch.exp(-0.5 * (d / init_gauss_sigma) ** 2)
    label = init_gauss - init_gauss.min()
    mask = mask_init_factor * torch.tanh(2.0 - d)
    spatial = torch.ones(num_dist_bins)
    if seed is not None:
        g = _gen(seed)
        label = label + 0.02 * torch.randn(num_dist_bins, generator=g)
        mask = mask + 0.1 * torch.randn(num_dist_bins, generator=g)
        spatial = spatial + 0.1 * torch.rand(num_dist_bins, generator=g)
    sd = OrderedDict()
    sd["log_step_length"] = torch.tensor([math.log(init_step)], dtype=torch.float32)
    sd["filter_reg"] = torch.tensor([init_reg], dtype=torch.float32)
    sd["label_map_predictor.weight"] = label.reshape(1, -1, 1, 1).clone()
    sd["target_mask_predictor.0.weight"] = mask.reshape(1, -1, 1, 1).clone()
    sd["spatial_weight_predictor.weight"] = spatial.reshape(1, -1, 1, 1).clone()
    return sd


def make_dimp_state_dict(arch="resnet50", seed=0, filter_size=4, lut_seed=None):
    """State dict for DiMP-50 / DiMP-18 (backbone to layer3 + clf head + initializer + optimizer)."""
    sd = make_backbone_state_dict(arch, seed)
    g = _gen(seed + 7919)
    if arch == "resnet18":
        # residual_basic_block(num_blocks=1, final_conv=True): BasicBlock(256,256) + conv3x3 256->256
        p = "classifier.feature_extractor.0."
        sd[p + "conv1.weight"] = _conv_w(g, 256, 256, 3)
        _bn(g, sd, p + "bn1", 256)
        sd[p + "conv2.weight"] = _conv_w(g, 256, 256, 3)
        _bn(g, sd, p + "bn2", 256, gamma_scale=0.5)
        sd["classifier.feature_extractor.1.weight"] = _conv_w(g, 256, 256, 3)
        cdim = 256
    else:
        # residual_bottleneck(num_blocks=0, final_conv=True): one conv3x3 1024->512 (features.py:66)
        sd["classifier.feature_extractor.0.weight"] = _conv_w(g, 512, 1024, 3)
        cdim = 512
    sd["classifier.filter_initializer.filter_conv.weight"] = _conv_w(g, cdim, cdim, 3)
    sd["classifier.filter_initializer.filter_conv.bias"] = 0.01 * torch.randn(cdim, generator=g)
    for k, v in make_dimp_optimizer_params(seed=lut_seed).items():
        sd["classifier.filter_optimizer." + k] = v
    return sd


def make_crop(seed, n=1, size=288):
    """Synthetic search crops in the tracker's pixel range [0,255], NCHW float32 (what
    NetWithBackbone.extract_backbone receives, pytracking/features/net_wrappers.py:71-75)."""
    g = _gen(seed)
    base = torch.rand(n, 3, size // 8, size // 8, generator=g)
    im = torch.nn.functional.interpolate(base, size=(size, size), mode="bilinear", align_corners=False)
    im = im + 0.15 * torch.rand(n, 3, size, size, generator=g)
    return (im.clamp(0, 1) * 255.0).contiguous()


def make_clf_features(seed, n, c=512, h=18, w=18, filter_size=4):
    """L2-normalised classification features as the clf head produces them (InstanceL2Norm with
    scale sqrt(1/(C*k*k)), ltr/models/tracking/dimpnet.py:159) -- PrDiMP overflows on raw randn."""
    g = _gen(seed)
    x = torch.randn(n, c, h, w, generator=g)
    # low-pass a little so neighbouring cells correlate like real features
    x = torch.nn.functional.avg_pool2d(x, 3, stride=1, padding=1) + 0.3 * x
    scale = math.sqrt(1.0 / (c * filter_size * filter_size))
    nrm = torch.sqrt((c * h * w) / ((x * x).reshape(n, -1).sum(1) + 1e-5)).reshape(n, 1, 1, 1)
    return (x * (scale * nrm)).contiguous()


def make_boxes(seed, n, center=120.0, size=50.0, jitter=20.0):
    """Target boxes (x, y, w, h) in crop pixels: [120,120,50,50] +- U(-20,20)."""
    g = _gen(seed)
    bb = torch.tensor([center, center, size, size]).repeat(n, 1)
    bb = bb + (torch.rand(n, 4, generator=g) * 2 - 1) * jitter
    return bb.contiguous()


def make_sequence(q, num_frames=200, height=480, width=640):
    """The survey's trackable synthetic sequence: uint8 RGB frames with an 80x60
    textured rectangle moving +3/+2 px per frame from (300,200). Returns (frames list, init_bbox)."""
    rng = np.random.RandomState(1000 + q)
    tex = (rng.rand(60, 80, 3) * 30).astype(np.float32)
    base_col = np.array([200, 120, 60], dtype=np.float32) + 10 * q
    frames = []
    for t in range(num_frames + 1):
        im = (rng.rand(height, width, 3) * 40).astype(np.float32)
        x0, y0 = 300 + 3 * t, 200 + 2 * t
        x0 = min(x0, width - 81)
        y0 = min(y0, height - 61)
        im[y0:y0 + 60, x0:x0 + 80, :] = np.clip(base_col + tex, 0, 255)
        frames.append(im.astype(np.uint8))
    return frames, [300.0, 200.0, 80.0, 60.0]


def sequence_ground_truth(q, num_frames=200, height=480, width=640):
    """Ground-truth boxes (x, y, w, h) of `make_sequence(q, num_frames)`, one per frame including frame 0."""
    out = []
    for t in range(num_frames + 1):
        out.append([float(min(300 + 3 * t, width - 81)), float(min(200 + 2 * t, height - 61)), 80.0, 60.0])
    return out


def make_atom_memory(seed, n, c=64, h=18, w=18, n_filled=None, sigma=1.5):
    """ATOM sample memory as `ATOM.init_memory`/`update_memory` keep it (pytracking/tracker/atom/atom.py:561-600):
    features [n,c,h,w] (p-norm normalised, featurebase.py:105-108 -> unit mean square), Gaussian labels
    y [n,1,h,w] (`dcf.label_function_spatial`, pytracking/libs/dcf.py:56-59, centre jittered per sample) and
    sample weights [n] (zero for unfilled slots, summing to one)."""
    g = _gen(seed)
    x = torch.randn(n, c, h, w, generator=g)
    x = torch.nn.functional.avg_pool2d(x, 3, stride=1, padding=1) + 0.3 * x
    x = x / torch.sqrt((x * x).reshape(n, -1).mean(1) + 1e-10).reshape(n, 1, 1, 1)
    n_filled = n if n_filled is None else n_filled
    ctr = (torch.rand(n, 2, generator=g) * 2 - 1) * 3.0
    k0 = torch.arange(h, dtype=torch.float32).view(1, -1, 1) - (h - 1) / 2
    k1 = torch.arange(w, dtype=torch.float32).view(1, 1, -1) - (w - 1) / 2
    y = torch.exp(-0.5 / sigma ** 2 * (k0 - ctr[:, 0].view(-1, 1, 1)) ** 2) * torch.exp(-0.5 / sigma ** 2 * (k1 - ctr[:, 1].view(-1, 1, 1)) ** 2)
    sw = torch.rand(n, generator=g) + 0.2
    sw[n_filled:] = 0
    sw = sw / sw.sum()
    x[n_filled:] = 0
    y[n_filled:] = 0
    return x.contiguous(), y.unsqueeze(1).contiguous(), sw.contiguous()


def make_transformer_state_dict(seed, d_model=256, nhead=8, dim_ff=2048, n_enc=6, n_dec=6):
    """Random-init parameters in the key layout of ltr/models/transformer/transformer.py `Transformer.state_dict()`
    (nn.MultiheadAttention: in_proj_weight/in_proj_bias/out_proj.weight/out_proj.bias; linear1/2; norm1..3; decoder.norm).
    Xavier-like scales so that activations stay O(1) through 12 layers."""
    g = _gen(seed)
    sd = OrderedDict()

    def lin(prefix, n_out, n_in, wname=".weight", bname=".bias"):
        sd[prefix + wname] = torch.randn(n_out, n_in, generator=g) * math.sqrt(2.0 / (n_in + n_out))
        sd[prefix + bname] = torch.randn(n_out, generator=g) * 0.02

    def mha(prefix):
        sd[prefix + ".in_proj_weight"] = torch.randn(3 * d_model, d_model, generator=g) * math.sqrt(2.0 / (2 * d_model))
        sd[prefix + ".in_proj_bias"] = torch.randn(3 * d_model, generator=g) * 0.02
        lin(prefix + ".out_proj", d_model, d_model)

    def norm(prefix):
        sd[prefix + ".weight"] = 1.0 + 0.1 * torch.randn(d_model, generator=g)
        sd[prefix + ".bias"] = 0.05 * torch.randn(d_model, generator=g)

    for i in range(n_enc):
        p = "encoder.layers.%d" % i
        mha(p + ".self_attn"); lin(p + ".linear1", dim_ff, d_model); lin(p + ".linear2", d_model, dim_ff)
        norm(p + ".norm1"); norm(p + ".norm2")
    for i in range(n_dec):
        p = "decoder.layers.%d" % i
        mha(p + ".self_attn"); mha(p + ".multihead_attn"); lin(p + ".linear1", dim_ff, d_model); lin(p + ".linear2", d_model, dim_ff)
        norm(p + ".norm1"); norm(p + ".norm2"); norm(p + ".norm3")
    norm("decoder.norm")
    return sd
