"""Workload for ncu launch lists: a few tracked frames of the bench.py step (device-resident crops, eager launches)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("B200TRK_GRAPH", "0")      # one launch per kernel, so that ncu lists them individually
import torch
import bench
eng, sw, synth = bench.setup_engine(0, 0)
crops = synth.make_crop(2000, 4, bench.CROP).cuda()
boxes = synth.make_boxes(3000, 4).numpy()
for i in range(int(sys.argv[1]) if len(sys.argv) > 1 else 3):
    eng.localize_device(crops[i % 4:i % 4 + 1])
    r = sw.step()
    eng.update(0, r, boxes[i % 4], sw.w, bench.MEMORY, bench.SD_ITERS)
torch.cuda.synchronize()
