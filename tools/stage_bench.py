"""Per-stage CUDA-event timings of the hot path (development aid; bench.py is the contract)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pytracking_b200 import ops, synth
from pytracking_b200.engine import BackboneEngine


def timeit(fn, iters=20, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(iters + 1)]
    ev[0].record()
    for i in range(iters):
        fn()
        ev[i + 1].record()
    torch.cuda.synchronize()
    ts = sorted(ev[i].elapsed_time(ev[i + 1]) * 1e3 for i in range(iters))
    return ts[len(ts) // 2], ts[0]


def main():
    prec = int(os.environ.get("PREC", "0"))
    res = {}
    sd = synth.make_dimp_state_dict("resnet50", seed=0, lut_seed=3)
    eng = BackboneEngine(sd, arch="resnet50", max_batch=13, crop_size=288, precision=prec)
    im = synth.make_crop(1, 1, 288).cuda()
    res["backbone+head S=1 (us)"] = timeit(lambda: eng.forward(im, want=("classification",)))
    im13 = synth.make_crop(2, 13, 288).cuda()
    res["backbone+head S=13 (us)"] = timeit(lambda: eng.forward(im13, want=("classification",)), iters=5, warm=2)
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        eng.forward(im, want=("classification",))
        torch.cuda.synchronize()
        with torch.cuda.graph(g, stream=s):
            o = eng.forward(im, want=("classification",))
    res["backbone+head S=1 graph (us)"] = timeit(lambda: g.replay())
    p = synth.make_dimp_optimizer_params(seed=3)
    luts = [p[k].cuda() for k in ("label_map_predictor.weight", "target_mask_predictor.0.weight", "spatial_weight_predictor.weight")]
    for n in (15, 50):
        feat = synth.make_clf_features(3, n, 512, 18, 18).cuda()
        bb = synth.make_boxes(4, n).cuda()
        sw = torch.full((n,), 1.0 / n).cuda()
        w0 = torch.zeros(1, 512, 4, 4).cuda()
        out = torch.empty_like(w0)
        for it in (0, 1, 2, 10):
            res["dimp_sd n=%d it=%d (us)" % (n, it)] = timeit(lambda: ops.dimp_sd_gn(w0, feat, bb, sw, *luts, it, 0.9, 0.01, out=out))
        res["apply_filter n=%d (us)" % n] = timeit(lambda: ops.apply_filter(feat, w0, return_max=True))
        r = torch.randn(n, 1, 19, 19).cuda()
        res["feat_transpose n=%d (us)" % n] = timeit(lambda: ops.apply_feat_transpose(feat, r, 4))
    f1 = synth.make_clf_features(5, 1, 512, 18, 18).cuda()
    w1 = torch.randn(1, 512, 4, 4).cuda()
    res["apply_filter+max2d n=1 (us)"] = timeit(lambda: ops.apply_filter(f1, w1, return_max=True))
    for k, v in res.items():
        print("%-36s median %9.1f  min %9.1f" % (k, v[0], v[1]))
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(res, open("gpurun_out/stage_bench_prec%d.json" % prec, "w"), indent=1)


if __name__ == "__main__":
    main()
