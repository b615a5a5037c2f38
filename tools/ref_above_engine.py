"""Runs the UNMODIFIED reference DiMP tracker (baseline/_ref) on the synthetic sequence three ways -- stock PyTorch on the CPU, stock
PyTorch on CUDA (TF32 off), and on CUDA above the engine (`pytracking_b200.plugin.install()`) -- and reports where the boxes / score
maps first differ plus the reference's own per-frame clock.  Writes gpurun_out/ref_above_engine_<tag>.json.

    python tools/ref_above_engine.py --frames 200 [--iou] [--tag r02a]
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def run(kind, frames, bb, use_iou, overrides, aug):
    from baseline import ref_tracker
    from pytracking_b200 import plugin
    dev = "cpu" if kind == "cpu" else "cuda"
    if kind == "engine":
        plugin.install(precision=int(os.environ.get("RAE_PRECISION", "0")))
        plugin.stats.clear()
    try:
        trk = ref_tracker.build_dimp(dev, use_iou_net=use_iou, overrides=overrides, use_augmentation=aug, dropout=False)
        scores, flags = [], []
        orig = trk.classify_target

        def hook(x):
            s = orig(x)
            scores.append(s.detach().float().cpu().numpy().copy())
            return s
        trk.classify_target = hook
        r = ref_tracker.run_sequence(trk, frames, bb, sync=(torch.cuda.synchronize if dev == "cuda" else None),
                                     on_frame=lambda t, tr, o: flags.append(tr.debug_info["flag"]))
        r["scores"] = np.stack(scores)
        r["flags"] = flags
        r["filter"] = trk.target_filter.detach().float().cpu().numpy()
        r["stats"] = dict(plugin.stats) if kind == "engine" else {}
    finally:
        if kind == "engine":
            plugin.uninstall()
    return r


def compare(a, b):
    nb = min(len(a["target_bbox"]), len(b["target_bbox"]))
    same = np.all(a["target_bbox"][:nb] == b["target_bbox"][:nb], axis=1)
    first_diff = int(np.argmin(same)) if not same.all() else -1
    sd = [float(np.abs(a["scores"][i] - b["scores"][i]).max() / (np.abs(a["scores"][i]).max() + 1e-30)) for i in range(nb)]
    fl = [i for i in range(nb) if a["flags"][i] != b["flags"][i]]
    return {"frames": nb, "boxes_identical": bool(same.all()), "first_box_diff_frame": first_diff, "first_flag_diff_frame": fl[0] if fl else -1,
            "score_rel_diff_by_frame": [round(v, 7) for v in sd[:12]] + ["..."] + [round(v, 7) for v in sd[-3:]],
            "max_box_abs_diff": float(np.abs(a["target_bbox"][:nb] - b["target_bbox"][:nb]).max()),
            "score_rel_diff_first": sd[0], "score_rel_diff_max": float(max(sd)),
            "score_rel_diff_until_first_box_diff": float(max(sd[:first_diff]) if first_diff > 0 else max(sd)),
            "flags_identical": a["flags"][:nb] == b["flags"][:nb],
            "filter_rel_diff": float(np.abs(a["filter"] - b["filter"]).max() / np.abs(a["filter"]).max())}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=200)
    ap.add_argument("--cpu-frames", type=int, default=40)
    ap.add_argument("--iou", action="store_true")
    ap.add_argument("--no-aug", action="store_true")
    ap.add_argument("--tag", default="r02")
    ap.add_argument("--threads", type=int, default=16)
    ap.add_argument("--default-schedule", action="store_true", help="stock dimp50 update schedule (train_skipping 20, 2 iterations)")
    args = ap.parse_args()
    from oracle import ref_shims
    ref_shims.install()
    from pytracking_b200 import synth
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.set_num_threads(args.threads)
    frames, bb = synth.make_sequence(0, num_frames=args.frames)
    out = {"frames": args.frames, "use_iou_net": args.iou, "augmentation": not args.no_aug}
    res = {}
    kinds = ["engine", "cpu"] if args.iou else ["cuda", "engine", "cpu"]       # stock CUDA has no PrRoIPool build
    for kind in kinds:
        n = args.cpu_frames if kind == "cpu" else args.frames
        # stock CUDA has no PrRoIPool: the learned filter initialiser cannot run there -> zero initialiser in every arm of that comparison
        ov = {} if args.iou else dict(filter_init_zero=True)
        if args.default_schedule:
            ov.update(train_skipping=20, net_opt_update_iter=2)
        res[kind] = run(kind, frames[:n + 1], bb, args.iou, ov, not args.no_aug)
        t = res[kind]["time"]
        out[kind] = {"ms_per_frame_median": float(np.median(t) * 1e3), "ms_per_frame_mean": float(t.mean() * 1e3),
                     "fps_reference_clock": float(len(t) / t.sum()), "init_s": res[kind]["init_time"], "stats": res[kind]["stats"],
                     "last_box": res[kind]["target_bbox"][-1].tolist()}
        print(kind, json.dumps(out[kind]), flush=True)
    ks = list(res)
    for i in range(len(ks)):
        for j in range(i + 1, len(ks)):
            out["%s_vs_%s" % (ks[i], ks[j])] = compare(res[ks[i]], res[ks[j]])
            print(ks[i], "vs", ks[j], json.dumps(out["%s_vs_%s" % (ks[i], ks[j])]), flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "ref_above_engine_%s%s%s.json" % (args.tag, "_iou" if args.iou else "", "_sched" if args.default_schedule else "")), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
