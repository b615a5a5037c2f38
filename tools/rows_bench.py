"""CUDA-event timings of the non-DiMP rows of SURVEY 8(a) at BASELINE sizes (development aid / profiles)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pytracking_b200 import ops, synth
from pytracking_b200.transformer_engine import TransformerEngine
from tools.stage_bench import timeit

res = {}
x, y, sw = synth.make_atom_memory(77, 250, 64, 18, 18, n_filled=250)
x, y, sw = x.cuda(), y.cuda(), sw.cuda()
w = (torch.randn(1, 64, 4, 4) * 0.02).cuda()
out = torch.empty_like(w)
res["atom_cg_filter n=250 C=64, 5 PR-CG it (us)"] = timeit(lambda: ops.atom_cg_filter(w, x, y, sw, 0.1, 5, "mlu", 0.05, False, out=out))
res["atom_cg_filter n=250 C=64, 1 it (us)"] = timeit(lambda: ops.atom_cg_filter(w, x, y, sw, 0.1, 1, "mlu", 0.05, False, out=out))
xi, yi, swi = synth.make_atom_memory(111, 30, 256, 18, 18)
xi, yi, swi = xi.cuda(), yi.cuda(), swi.cuda()
P0 = (torch.randn(64, 256, 1, 1) / 16).cuda()
def gn():
    wz, Pz = torch.zeros(1, 64, 4, 4, device="cuda"), P0.clone()
    ops.atom_gn_joint_(wz, Pz, xi, yi, swi, 0.1, 1e-4, 10, 6, "mlu", 0.05, True)
res["atom_gn_joint n=30 256->64, 6 GN x 10 CG (us)"] = timeit(gn, iters=5, warm=2)
s5 = torch.randn(5, 256, 18, 18).cuda()
Pm = (torch.randn(64, 256, 1, 1) * 0.05).cuda()
def atom_localize():
    xn = ops.feature_normalize_(s5.clone(), 2.0)
    pr = ops.conv1x1(xn, Pm)
    sc = ops.conv2d_same(pr, w)
    up = ops.fourier_interp(sc, (4, 4), (288, 288))
    return ops.max2d(up[:, 0])
res["ATOM stage 2, 5 scales: normalise+project+conv_same+Fourier 288^2+max2d (us)"] = timeit(atom_localize)
d, nh, ff, L, B = 256, 8, 2048, 972, 2
sd = synth.make_transformer_state_dict(95, d, nh, ff, 6, 6)
eng = TransformerEngine(sd, L, B, d, nh, ff, 6, 6)
src, pos, qe = torch.randn(L, B, d).cuda(), (torch.randn(L, 1, d) * 0.5).cuda(), torch.randn(1, d).cuda()
mask = torch.zeros(B, L, dtype=torch.bool); mask[1, 324:648] = True
mask = mask.cuda()
res["ToMP Transformer.forward 972x2 tokens, 6+6 layers (us)"] = timeit(lambda: eng.forward(src, mask, qe, pos), iters=10, warm=3)
for k, v in res.items():
    print("%-86s median %9.1f  min %9.1f" % (k, v[0], v[1]))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/rows_bench.json", "w"), indent=1)
try:
    from baseline import ref_env, ref_tracker
    if ref_env.reference_available():
        from pytracking_b200.iou import IoUPredictor
        from pytracking_b200.transformer_engine import BoxTower, TokenBuilder
        net = ref_tracker.build_dimp_net()
        pred = IoUPredictor(net.state_dict())
        f3, f4 = torch.relu(torch.randn(1, 256, 36, 36)).cuda(), torch.relu(torch.randn(1, 256, 18, 18)).cuda()
        mod = [torch.rand(256).cuda(), torch.rand(256).cuda()]
        boxes = torch.tensor([[100.0, 110.0, 70.0, 55.0]] * 10).cuda()
        extra = {"IoUNet refine R=10, 5 ascent steps (us)": timeit(lambda: pred.refine(mod, [f3, f4], boxes, 5, 1.0)),
                 "IoUNet predict_iou + box gradient R=10 (us)": timeit(lambda: pred.predict_iou(mod, [f3, f4], boxes, return_grad=True))}
        for k, v in extra.items():
            print("%-86s median %9.1f  min %9.1f" % (k, v[0], v[1]))
        res.update(extra)
        json.dump(res, open("gpurun_out/rows_bench.json", "w"), indent=1)
except Exception as e:
    print("IoU rows skipped:", repr(e))
