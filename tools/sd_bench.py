import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pytracking_b200 import ops, synth
from tools.stage_bench import timeit
p = synth.make_dimp_optimizer_params(seed=3)
luts = [p[k].cuda() for k in ("label_map_predictor.weight", "target_mask_predictor.0.weight", "spatial_weight_predictor.weight")]
for n in (15, 50):
    feat = synth.make_clf_features(3, n, 512, 18, 18).cuda()
    bb = synth.make_boxes(4, n).cuda()
    sw = torch.full((n,), 1.0 / n).cuda()
    w0 = torch.zeros(1, 512, 4, 4).cuda()
    out = torch.empty_like(w0)
    r = []
    for it in (0, 1, 2, 10):
        r.append(timeit(lambda: ops.dimp_sd_gn(w0, feat, bb, sw, *luts, it, 0.9, 0.01, out=out))[0])
    print("n=%d  it0 %.1f it1 %.1f it2 %.1f it10 %.1f  -> per-iter %.1f us" % (n, r[0], r[1], r[2], r[3], (r[3] - r[2]) / 8))
