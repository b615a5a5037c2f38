"""tools/eco_bench.py (the script bench.py runs in a subprocess for the ECO rows of `rooflines[]`) has never executed on a GPU box; a
shape or argument slip in it would silently turn those rows into "unavailable".  Dry run on the CPU: the script itself, with `.cuda()` a
no-op, the CUDA-event timer replaced by a single call, and the C library replaced by a recorder -- so that every argument validation of the
real `pytracking_b200.ops` wrappers runs on the tensors the script builds, and the C entry points are called with the argument counts the
ctypes table declares."""
import json
import os
import runpy
import sys
import unittest.mock as um

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class _Recorder:
    def __init__(self, table):
        self.table, self.calls = table, []

    def __getattr__(self, name):
        if name not in self.table:
            raise AttributeError(name)
        nargs = len(self.table[name][1])

        def fn(*args):
            assert len(args) == nargs, (name, len(args), nargs)
            self.calls.append(name)
            return 0
        return fn


def test_eco_bench_script_builds_valid_calls(tmp_path):
    from pytracking_b200 import _lib, ops
    import tools.stage_bench as sb
    table = next(v for v in vars(_lib).values() if isinstance(v, dict) and "b200trk_eco_filter_cg" in v)
    rec = _Recorder(table)
    out = str(tmp_path / "eco_bench.json")

    def dev(t, name, contiguous=True):
        assert isinstance(t, torch.Tensor) and t.dtype == torch.float32, name
        return t.contiguous() if contiguous else t

    def timeit(fn, iters=20, warm=5):
        fn()
        fn()                                                        # a second call: the carried CG state / the energy of the first
        return 10.0, 9.0

    with um.patch.object(torch.Tensor, "cuda", lambda self, *a, **k: self), um.patch.object(torch.Tensor, "is_cuda", property(lambda self: True)), \
            um.patch.object(torch.cuda, "current_device", lambda: 0), um.patch.object(ops, "_dev", dev), \
            um.patch.object(ops, "_stream", lambda: None), um.patch.object(_lib, "lib", lambda: rec), um.patch.object(_lib, "check", lambda rc, what: None), \
            um.patch.object(sb, "timeit", timeit), um.patch.object(sys, "argv", ["eco_bench.py", "--json", out]):
        runpy.run_path(os.path.join(ROOT, "tools", "eco_bench.py"), run_name="__main__")
    data = json.load(open(out))
    assert len([k for k in data if not k.startswith(("joint ", "scores "))]) == 2 and len([k for k in data if k.startswith("joint ")]) == 2
    assert len([k for k in data if k.startswith("scores ")]) == 1, list(data)
    for name, n in (("b200trk_eco_filter_cg", 4), ("b200trk_eco_joint_gn", 4), ("b200trk_eco_preprocess_sample", 4), ("b200trk_eco_apply_filter", 4),
                    ("b200trk_eco_sample_fs", 2), ("b200trk_max2d", 2)):
        assert rec.calls.count(name) == n, (name, rec.calls.count(name))
    import bench
    rows = bench.eco_rows_from(data, {"hbm_gbs": 6000.0})
    assert len(rows) == 5 and all(r["us_per_launch"] == 10.0 for r in rows)
