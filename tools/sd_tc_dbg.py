"""Phase trace of the tcgen05 SD kernel under the timing experiments (B200TRK_SD_DBG)."""
import sys, os, ctypes as C
os.environ["B200TRK_SD_TC"] = "1"; os.environ["B200TRK_SD_TRACE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
from pytracking_b200 import ops, synth, _lib
p = synth.make_dimp_optimizer_params(seed=3)
luts = [p[k].cuda() for k in ("label_map_predictor.weight", "target_mask_predictor.0.weight", "spatial_weight_predictor.weight")]
feat = synth.make_clf_features(3, 50, 512, 18, 18).cuda(); bb = synth.make_boxes(4, 50).cuda(); sw = torch.full((50,), 1.0 / 50).cuda()
w0 = torch.zeros(1, 512, 4, 4).cuda()
for dbg in (0, 1, 2, 3, 4, 5):
    os.environ["B200TRK_SD_DBG"] = str(dbg)
    for _ in range(3):
        ops.dimp_sd_gn(w0, feat, bb, sw, *luts, 3, 0.9, 0.01)
    torch.cuda.synchronize()
    buf = (C.c_uint64 * 64)()
    _lib.check(_lib.lib().b200trk_debug_sd_trace(buf))
    t = np.array(list(buf), dtype=np.float64)
    b = 8 + 10
    d = [(t[b+k+1]-t[b+k])/1e3 for k in range(8)]
    print("dbg %d: prologue %.2f |" % (dbg, (t[1]-t[0])/1e3), end=" ")
    print("dbg %d: s0 sweepA %.2f | it1: resid %.2f | sweepT %.2f | barrier1 %.2f | gsum+b1b+FT %.2f | sweepA %.2f | barrier2 %.2f | qsum+h %.2f | barrier3 %.2f" % (dbg, (t[2]-t[1])/1e3, *d), flush=True)
names = "slot_free tma_issued | conv_at_unit landed tmem_free regs st_retired published | mma_sees mma_committed"
order = [0, 1, 8, 2, 3, 9, 4, 5, 6, 7]
_lib.lib().b200trk_debug_sd_units((C.c_uint64 * 256)())
for dbgu in (6, 0, 4, 1, 3):
  os.environ["B200TRK_SD_DBG"] = str(dbgu)
  for _ in range(3):
    ops.dimp_sd_gn(w0, feat, bb, sw, *luts, 3, 0.9, 0.01)
  torch.cuda.synchronize()
  ub = (C.c_uint64 * 256)()
  _lib.check(_lib.lib().b200trk_debug_sd_units(ub))
  u = np.array(list(ub), dtype=np.float64).reshape(16, 16)
  t0 = u[u > 0].min()
  print("dbg %d  unit: %s   (SM clocks since first; -1 = not stamped)" % (dbgu, names))
  for i in range(16):
    print("%2d: " % i + " ".join("%7d" % ((u[i, k] - t0) if u[i, k] > 0 else -1) for k in order))
