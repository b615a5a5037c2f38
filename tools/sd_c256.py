"""SD-GN timing at C=256 (DiMP-18-like width) for the decomposition choice (B200TRK_SD_PASSES forces one)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pytracking_b200 import ops, synth
from stage_bench import timeit
C = int(sys.argv[1]) if len(sys.argv) > 1 else 256
p = synth.make_dimp_optimizer_params(seed=3)
luts = [p[k].cuda() for k in ("label_map_predictor.weight", "target_mask_predictor.0.weight", "spatial_weight_predictor.weight")]
for n in (15, 50):
    feat = synth.make_clf_features(3, n, C, 18, 18).cuda()
    bb = synth.make_boxes(4, n).cuda()
    sw = torch.full((n,), 1.0 / n).cuda()
    w0 = torch.zeros(1, C, 4, 4).cuda()
    out = torch.empty_like(w0)
    print("C=%d n=%d it=10: median %.1f us min %.1f us (B200TRK_SD_PASSES=%s)" % ((C, n) + timeit(lambda: ops.dimp_sd_gn(w0, feat, bb, sw, *luts, 10, 0.9, 0.01, out=out)) + (os.environ.get("B200TRK_SD_PASSES", "auto"),)))
