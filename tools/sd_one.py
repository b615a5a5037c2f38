"""One SD-GN call per path (for ncu): B200TRK_SD_TC selects the tcgen05 kernel."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pytracking_b200 import ops, synth
p = synth.make_dimp_optimizer_params(seed=3)
luts = [p[k].cuda() for k in ("label_map_predictor.weight", "target_mask_predictor.0.weight", "spatial_weight_predictor.weight")]
n = int(sys.argv[1]) if len(sys.argv) > 1 else 50
feat = synth.make_clf_features(3, n, 512, 18, 18).cuda(); bb = synth.make_boxes(4, n).cuda(); sw = torch.full((n,), 1.0 / n).cuda()
w0 = torch.zeros(1, 512, 4, 4).cuda()
for _ in range(3):
    ops.dimp_sd_gn(w0, feat, bb, sw, *luts, 10, 0.9, 0.01)
torch.cuda.synchronize()
