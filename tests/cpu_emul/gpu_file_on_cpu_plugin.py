"""TEST INFRASTRUCTURE ONLY -- pytest plugin (`-p gpu_file_on_cpu_plugin`, tests/cpu_emul on sys.path) that lets the `-m gpu` test file of
the ECO entry points run WITHOUT a GPU: `.cuda()` is the identity, and `pytracking_b200._lib.lib()` returns a handle to the host build of
the launchers + kernel sources (tests/cpu_emul/eco_abi_emul.cpp, path in $B200_ECO_ABI_EMUL) instead of libb200trk.so.  Everything between
the test and the C ABI is the product's own: pytracking_b200/ops.py (argument validation, ctypes calls with the signatures of
_lib.SIGNATURES), pytracking_b200/plugin.py, the launchers' host code, the kernels.  Only used by tests/test_eco_gpu_file_on_cpu.py, in a
subprocess; nothing in the product knows about it."""
import ctypes as C
import os

import torch

ECO = ("b200trk_eco_filter_cg", "b200trk_eco_joint_gn", "b200trk_eco_apply_filter", "b200trk_eco_sample_fs", "b200trk_eco_preprocess_sample",
       "b200trk_eco_shift_fs")


class _Handle:
    def __init__(self, path, signatures):
        h = C.CDLL(path)
        for name in ECO:
            fn = getattr(h, name)
            fn.restype, fn.argtypes = signatures[name]
            setattr(self, name, fn)
        h.b200trk_last_error.restype = C.c_char_p
        self.b200trk_last_error = h.b200trk_last_error

    def __getattr__(self, name):
        raise AssertionError("the CPU run of the ECO GPU tests reached %s, which is not an ECO entry point" % name)


def pytest_configure(config):
    from oracle import eco_oracle as E
    from pytracking_b200 import _lib, ops, plugin
    handle = _Handle(os.environ["B200_ECO_ABI_EMUL"], _lib.SIGNATURES)
    _lib.lib = lambda: handle

    def dev(t, name, contiguous=True):
        if not isinstance(t, torch.Tensor):
            raise RuntimeError("b200trk: '%s' must be a CUDA tensor (the engine has no CPU path)" % name)
        if t.dtype != torch.float32:
            raise RuntimeError("b200trk: '%s' must be float32, got %s" % (name, t.dtype))
        return t.contiguous() if contiguous else t
    ops._dev = dev
    ops._stream = lambda: C.c_void_p(0)

    def max2d(a):                                                   # b200trk_max2d (GPU-validated) is not part of this run: dcf.max2d restated
        mv, mi = E.max2d(a.reshape(-1, 1, a.shape[-2], a.shape[-1]))
        return mv.reshape(a.shape[:-2]), mi.reshape(*a.shape[:-2], 2)
    ops.max2d = max2d
    plugin.max2d = max2d
    plugin._inference = lambda *ts: all(isinstance(t, torch.Tensor) and t.dtype == torch.float32 and not (torch.is_grad_enabled() and t.requires_grad)
                                        for t in ts)
    try:                                                            # the tracker-level tests build the reference tracker "on cuda"
        from baseline import ref_tracker
        build_eco = ref_tracker.build_eco
        ref_tracker.build_eco = lambda device="cpu", overrides=None, **k: build_eco(device="cpu", overrides=overrides, **k)
    except Exception:
        pass
    torch.Tensor.cuda = lambda self, *a, **k: self
    torch.Tensor.is_cuda = property(lambda self: True)
    torch.cuda.synchronize = lambda *a, **k: None
    torch.cuda.current_device = lambda: 0
