"""Markdown table of the round's bench lines (profiles/r02_bench.json, r02_bench_reference_arm.json, r02_bench_other_configs.jsonl,
r02_scale_*.json) for DESIGN.md section 5."""
import json
import os
import sys

P = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles")


def load(name):
    f = os.path.join(P, name)
    return [json.loads(l) for l in open(f) if l.startswith("{")] if os.path.exists(f) else []


out = []
b = load("r02_bench.json")
if b:
    d = b[-1]
    tc = d.get("torch_cuda_baseline", {})
    out.append("| DiMP-50 (BASELINE configs[1]), 1 x B200 | frames/s |")
    out.append("|---|---|")
    out.append("| native tracker, frames resident in HBM (`value`) | **%.0f** (%.3f ms/frame) |" % (d["value"], d["ms_per_step"]))
    out.append("| native tracker, pinned host frames in, box out (`e2e`) | **%.0f** (median %.3f ms/frame) |" % (d["e2e"]["value"], d["e2e"].get("ms_per_frame_median", float("nan"))))
    if "reference_above_engine" in tc:
        out.append("| unmodified reference tracker above the engine (`plugin.install()`) | %.0f |" % tc["reference_above_engine"]["value"])
    if tc:
        out.append("| unmodified reference tracker, stock PyTorch-CUDA (cuDNN / cuBLAS, TF32 off), same GPU | %.1f |" % tc["value"])
    if "cpu_baseline" in d:
        out.append("| unmodified reference tracker, PyTorch-CPU, %d host cores (`cpu_baseline`) | %.1f |" % (d["cpu_baseline"]["cores"], d["cpu_baseline"]["value"]))
    r = load("r02_bench_reference_arm.json")
    if r:
        out.append("| `bench.py --impl reference` (the same, %d steps) | %.1f |" % (r[-1]["steps"], r[-1]["value"]))
    out.append("")
    out.append("| kernel | µs per launch | roofline | fraction |")
    out.append("|---|---|---|---|")
    for k in d["rooflines"]:
        out.append("| %s | %.1f | %s: %.1f of %.1f %s | %.3f |" % (k["kernel"], k["us_per_launch"], k["bound"], k["achieved"], k["peak"], k["unit"], k["frac"]))
    out.append("")
    out.append("mean IoU against the synthetic ground truth %.3f; clocks %s / %s MHz, throttle reasons %s; %d launches in %d timed frames." % (
        d["tracking"]["mean_iou_vs_synthetic_ground_truth"], d["clocks"]["sm_mhz"], d["clocks"]["sm_max_mhz"], d["clocks"]["reasons"], d["gpu_launches"], d["steps"]))
    out.append("")
o = load("r02_bench_other_configs.jsonl")
if o:
    out.append("| other BASELINE configurations (unmodified reference tracker objects, frames/s) | above the engine | stock PyTorch-CUDA | PyTorch-CPU |")
    out.append("|---|---|---|---|")
    for d in o:
        out.append("| %s | **%.1f** | %.1f | %.1f |" % (d["metric"].split(" tracked")[0], d["value"], d["torch_cuda_baseline"]["value"], d["cpu_baseline"]["value"]))
    out.append("")
sc = []
for n in (1, 2, 4, 8):
    for tag in ("", "_reference"):
        x = load("r02_scale_n%d%s.json" % (n, tag))
        if x:
            sc.append((n, tag, x[-1]))
if sc:
    out.append("| GPUs | arm | frames/s (whole job) | mean IoU |")
    out.append("|---|---|---|---|")
    for n, tag, d in sc:
        out.append("| %d | %s | %.0f | %s |" % (n, "reference (CPU)" if tag else "b200", d["value"], ("%.3f" % d["tracking"]["mean_iou_vs_synthetic_ground_truth"]) if "tracking" in d else "-"))
    out.append("")
print("\n".join(out))
