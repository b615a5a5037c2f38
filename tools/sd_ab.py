"""A/B timing of the tcgen05 SD kernel under B200TRK_SD_DBG settings that keep the result valid (0 = product, 6 = no cross-sweep
prefetch, 7 = blocking barrier waits instead of the look-ahead tests): CUDA events, n = 50, C = 512, 18x18, 10 iterations."""
import os, sys
os.environ["B200TRK_SD_TC"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pytracking_b200 import ops, synth
p = synth.make_dimp_optimizer_params(seed=3)
luts = [p[k].cuda() for k in ("label_map_predictor.weight", "target_mask_predictor.0.weight", "spatial_weight_predictor.weight")]
feat = synth.make_clf_features(3, 50, 512, 18, 18).cuda(); bb = synth.make_boxes(4, 50).cuda(); sw = torch.full((50,), 1.0 / 50).cuda()
w0 = torch.zeros(1, 512, 4, 4).cuda()
ref = None
for dbg in (0, 7, 6, 0, 7):
    os.environ["B200TRK_SD_DBG"] = str(dbg)
    for _ in range(5):
        out = ops.dimp_sd_gn(w0, feat, bb, sw, *luts, 10, 0.9, 0.01)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50):
        out = ops.dimp_sd_gn(w0, feat, bb, sw, *luts, 10, 0.9, 0.01)
    e1.record(); torch.cuda.synchronize()
    w = out[0] if isinstance(out, (tuple, list)) else out
    if ref is None: ref = w.clone()
    print("B200TRK_SD_DBG=%d: %.1f us per call, max |w - w(dbg 0)| = %.1e" % (dbg, e0.elapsed_time(e1) * 1e3 / 50, float((w - ref).abs().max())))
