"""Out-of-tree compatibility shims that let the UNMODIFIED reference (visionml/pytracking @ 7eb9e74) import and run on this
image's torch 2.11 -- on the CPU or on CUDA.  The reference tree is taken from baseline/_ref (staged byte for byte by
baseline/stage_reference.py; this is what travels to the GPU box) or, failing that, from /root/reference.  It is never modified.

Used by the reference arm of bench.py (`--impl reference`, and the PyTorch-CUDA reference timing), by the plug-in tests that run
the reference trackers above the engine, and (through oracle/ref_shims.py, which adds the CPU PrRoIPool restatement) by the golden
generators.  Nothing under pytracking_b200/ imports this module.

Each shim addresses one incompatibility listed in SURVEY.md section 8(c):
  1. missing optional deps imported at module scope (matplotlib, visdom, jpeg4py, timm, cv2...)
  2. torchvision.models.resnet.model_urls removed     (ltr/models/backbone/resnet.py:5)
  3. torch.rfft / torch.irfft removed                  (pytracking/libs/fourier.py:24,31)
  4. TensorList.__getattr__ answers __torch_function__ (pytracking/libs/tensorlist.py:173-180)
  5. torch.load(weights_only=True) default             (ltr/admin/loading.py:125)
  6. local.py environment modules                      (pytracking/evaluation/environment.py:57-67)
(7, PrRoIPool: the CUDA replacement is installed by pytracking_b200.plugin.install(); the CPU restatement by oracle/ref_shims.py.)
"""
import importlib
import os
import sys
import types

_HERE = os.path.dirname(os.path.abspath(__file__))


def reference_root():
    env = os.environ.get("PYTRACKING_REFERENCE")
    for cand in ([env] if env else []) + [os.path.join(_HERE, "_ref"), "/root/reference"]:
        if cand and os.path.isdir(os.path.join(cand, "pytracking")):
            return cand
    return None


def reference_available() -> bool:
    return reference_root() is not None


def _stub_module(name, **attrs):
    mod = types.ModuleType(name)
    mod.__dict__.update(attrs)
    mod.__path__ = []  # behave as a package so that submodule imports resolve
    sys.modules[name] = mod
    return mod


class _Anything:
    """Object that swallows any attribute access / call (for plotting & UI stubs)."""

    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return _Anything()

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Anything()


_installed = False


def install(scratch_dir=None):
    """Install all shims and put the reference on sys.path. Idempotent."""
    global _installed
    if _installed:
        return
    REFERENCE_ROOT = reference_root()
    if REFERENCE_ROOT is None:
        raise RuntimeError("reference tree not found (baseline/_ref is staged by baseline/stage_reference.py)")
    import torch
    import torchvision.models.resnet as tv_resnet

    # -- 1. optional deps --------------------------------------------------------------------
    for name in ("matplotlib", "visdom", "jpeg4py", "timm", "cv2", "pandas_stub"):
        try:
            importlib.import_module(name)
        except Exception:
            if name == "matplotlib":
                m = _stub_module("matplotlib", use=lambda *a, **k: None)
                _stub_module("matplotlib.pyplot", **{k: _Anything() for k in (
                    "figure", "plot", "draw", "pause", "imshow", "cla", "axis", "title", "subplots")})
                _stub_module("matplotlib.patches", Rectangle=_Anything)
                m.pyplot = sys.modules["matplotlib.pyplot"]
                m.patches = sys.modules["matplotlib.patches"]
            elif name == "visdom":
                _stub_module("visdom", Visdom=_Anything)
                _stub_module("visdom.server", download_scripts=lambda *a, **k: None)
            elif name == "jpeg4py":
                _stub_module("jpeg4py", JPEG=_Anything)
            elif name == "timm":
                _stub_module("timm")
                _stub_module("timm.models")
                _stub_module("timm.models.layers", DropPath=torch.nn.Identity,
                             to_2tuple=lambda x: (x, x) if not isinstance(x, tuple) else x,
                             trunc_normal_=lambda t, **k: t)
            elif name == "cv2":
                _stub_module("cv2", **{k: _Anything() for k in (
                    "imread", "cvtColor", "COLOR_BGR2RGB", "resize", "warpAffine", "getRotationMatrix2D",
                    "BORDER_REPLICATE", "INTER_LINEAR", "setNumThreads")})
    for name in ("tikzplotlib", "pycocotools", "lvis", "tensorboardX", "skimage"):
        try:
            importlib.import_module(name)
        except Exception:
            _stub_module(name)

    # -- 2. torchvision model_urls ---------------------------------------------------------------
    if not hasattr(tv_resnet, "model_urls"):
        tv_resnet.model_urls = {}

    # -- 3. legacy FFT API -------------------------------------------------------------------------
    if not hasattr(torch, "rfft") or not callable(getattr(torch, "rfft", None)):
        def _rfft(a, signal_ndim, normalized=False, onesided=True):
            dims = tuple(range(-signal_ndim, 0))
            f = torch.fft.rfftn(a, dim=dims) if onesided else torch.fft.fftn(a, dim=dims)
            return torch.view_as_real(f)

        def _irfft(a, signal_ndim, normalized=False, onesided=True, signal_sizes=None):
            dims = tuple(range(-signal_ndim, 0))
            c = torch.view_as_complex(a.contiguous())
            return torch.fft.irfftn(c, s=tuple(signal_sizes) if signal_sizes is not None else None, dim=dims)

        torch.rfft = _rfft
        torch.irfft = _irfft

    # -- 5. torch.load ---------------------------------------------------------------------------------
    _orig_load = torch.load
    if not getattr(_orig_load, "_b200_shim", False):
        def _load(*a, **k):
            k.setdefault("weights_only", False)
            return _orig_load(*a, **k)
        _load._b200_shim = True
        torch.load = _load

    # -- sys.path -----------------------------------------------------------------------------------------
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)

    # -- 6. local env modules ------------------------------------------------------------------------------
    scratch = scratch_dir or os.environ.get("B200TRK_SCRATCH", "/tmp/b200trk_scratch")
    os.makedirs(os.path.join(scratch, "networks"), exist_ok=True)
    os.makedirs(os.path.join(scratch, "results"), exist_ok=True)

    import pytracking.evaluation.environment as pt_env  # light module (importlib, os only)

    def _pt_local_env_settings():
        s = pt_env.EnvSettings()
        s.network_path = os.path.join(scratch, "networks") + "/"
        s.results_path = os.path.join(scratch, "results") + "/"
        return s
    _stub_module("pytracking.evaluation.local", local_env_settings=_pt_local_env_settings)

    class _LtrEnv:
        def __init__(self):
            self.workspace_dir = scratch
            self.tensorboard_dir = scratch
            self.pretrained_networks = os.path.join(scratch, "networks") + "/"
    _stub_module("ltr.admin.local", EnvironmentSettings=_LtrEnv)

    # -- 4. TensorList dunder spoofing ---------------------------------------------------------------------
    from pytracking.libs import tensorlist as tl_mod
    _orig_getattr = tl_mod.TensorList.__getattr__

    def _safe_getattr(self, name):
        if name.startswith("__") and name.endswith("__"):
            raise AttributeError(name)
        return _orig_getattr(self, name)
    tl_mod.TensorList.__getattr__ = _safe_getattr

    # -- 8. ATOM indexes CPU tensors with the CUDA tensor `scale_ind` (pytracking/tracker/atom/atom.py:322-330: `disp[scale_ind, ...]`),
    #       which current torch rejects ("indices should be either on cpu or on the same device"); its own `dcf.max2d` results are
    #       handed to it on the CPU (the next line of the reference moves them there anyway, atom.py:323). Scoped to the ATOM module.
    try:
        import pytracking.tracker.atom.atom as atom_mod
        import pytracking.libs.dcf as dcf_mod

        class _DcfForAtom:
            def __getattr__(self, name):
                return getattr(dcf_mod, name)

            @staticmethod
            def max2d(a):
                v, i = dcf_mod.max2d(a)          # resolved at call time: the plug-in's binding when installed
                return v.cpu(), i.cpu()
        atom_mod.dcf = _DcfForAtom()
    except Exception as e:  # pragma: no cover
        sys.stderr.write("[ref_env] ATOM index shim not installed: %r\n" % (e,))

    _installed = True
