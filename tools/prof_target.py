"""Small workload for ncu captures: SD optimiser (n=50, 10 it), classification (n=1), one backbone pass."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pytracking_b200 import ops, synth
from pytracking_b200.engine import BackboneEngine

what = sys.argv[1] if len(sys.argv) > 1 else "all"
p = synth.make_dimp_optimizer_params(seed=3)
luts = [p[k].cuda() for k in ("label_map_predictor.weight", "target_mask_predictor.0.weight", "spatial_weight_predictor.weight")]
if what in ("all", "sd"):
    n = 50
    feat = synth.make_clf_features(3, n, 512, 18, 18).cuda()
    bb = synth.make_boxes(4, n).cuda()
    sw = torch.full((n,), 1.0 / n).cuda()
    w0 = torch.zeros(1, 512, 4, 4).cuda()
    for _ in range(3):
        ops.dimp_sd_gn(w0, feat, bb, sw, *luts, 10, 0.9, 0.01)
    f1 = synth.make_clf_features(5, 1, 512, 18, 18).cuda()
    w1 = torch.randn(1, 512, 4, 4).cuda()
    for _ in range(3):
        ops.apply_filter(f1, w1, return_max=True)
if what in ("all", "net"):
    sd = synth.make_dimp_state_dict("resnet50", seed=0, lut_seed=3)
    eng = BackboneEngine(sd, arch="resnet50", max_batch=1, crop_size=288, precision=int(os.environ.get("PREC", "0")))
    im = synth.make_crop(1, 1, 288).cuda()
    for _ in range(3):
        eng.forward(im, want=("classification",))
torch.cuda.synchronize()
