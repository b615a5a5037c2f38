import sys, os, ctypes as C
os.environ["B200TRK_SD_TRACE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
from pytracking_b200 import ops, synth, _lib
p = synth.make_dimp_optimizer_params(seed=3)
luts = [p[k].cuda() for k in ("label_map_predictor.weight", "target_mask_predictor.0.weight", "spatial_weight_predictor.weight")]
n = int(sys.argv[1]) if len(sys.argv) > 1 else 50
feat = synth.make_clf_features(3, n, 512, 18, 18).cuda()
bb = synth.make_boxes(4, n).cuda()
sw = torch.full((n,), 1.0 / n).cuda()
w0 = torch.zeros(1, 512, 4, 4).cuda()
for _ in range(4):
    ops.dimp_sd_gn(w0, feat, bb, sw, *luts, 4, 0.9, 0.01)
torch.cuda.synchronize()
buf = (C.c_uint64 * 64)()
_lib.check(_lib.lib().b200trk_debug_sd_trace(buf))
t = np.array(list(buf), dtype=np.float64)
t0 = t[0]
print("prologue %.2f | s0 sweep %.2f | barrier %.2f | s-sum %.2f" % ((t[1]-t[0])/1e3, (t[2]-t[1])/1e3, (t[3]-t[2])/1e3, (t[4]-t[3])/1e3))
for it in range(4):
    b = 8 + it * 10
    d = [(t[b+k+1]-t[b+k])/1e3 for k in range(8)]
    print("it %d: resid %.2f | sweepT %.2f | barrier1 %.2f | gsum %.2f | sweepA %.2f | barrier2 %.2f | qsum+h %.2f | barrier3 %.2f | total %.2f" % (
        it, *d, (t[b+10]-t[b])/1e3 if t[b+10] else (t[b+8]-t[b])/1e3))
