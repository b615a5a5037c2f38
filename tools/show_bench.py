"""Print the headline fields of bench.py JSON lines found in the given log files."""
import json
import sys

for f in sys.argv[1:]:
    try:
        d = json.loads([l for l in open(f) if l.startswith("{")][-1])
    except Exception as e:
        print(f, "no JSON line:", repr(e)[:100])
        continue
    roof = [(r["kernel"][:22], round(r["us_per_launch"], 1), round(r["frac"], 4)) for r in d.get("rooflines", [])]
    extra = {k: (round(d[k]["value"], 1) if isinstance(d.get(k), dict) and "value" in d[k] else None) for k in ("cpu_baseline", "torch_cuda_baseline")}
    print(f, "value %.1f e2e %.1f n_gpus %s" % (d["value"], d["e2e"]["value"], d.get("n_gpus")), roof, extra,
          d.get("tracking", {}).get("mean_iou_vs_synthetic_ground_truth"))
