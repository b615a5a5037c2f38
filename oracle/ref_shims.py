"""TEST INFRASTRUCTURE ONLY -- never imported by the product path.

Out-of-tree compatibility shims that let the UNMODIFIED reference
(visionml/pytracking @ 7eb9e74, mounted read-only at /root/reference) import and
run on this container's torch 2.11 / CPU.  Used only by `oracle/gen_golden.py`
(golden-vector generation) and by the CPU tests that cross-check the oracle
restatement against the real reference when /root/reference is present.

Each shim addresses one incompatibility listed in SURVEY.md section 8(c):
  1. missing optional deps imported at module scope (matplotlib, visdom, jpeg4py, timm, cv2...)
  2. torchvision.models.resnet.model_urls removed     (ltr/models/backbone/resnet.py:5)
  3. torch.rfft / torch.irfft removed                  (pytracking/libs/fourier.py:24,31)
  4. TensorList.__getattr__ answers __torch_function__ (pytracking/libs/tensorlist.py:173-180)
  5. torch.load(weights_only=True) default             (ltr/admin/loading.py:125)
  6. local.py environment modules                      (pytracking/evaluation/environment.py:57-67)
  7. PrRoIPool has no CPU path / does not build        (ltr/external/PreciseRoIPooling)
The reference tree is never modified.
"""
import importlib
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("PYTRACKING_REFERENCE", "/root/reference")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "pytracking"))


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
    if not reference_available():
        raise RuntimeError("reference tree not found at %s" % REFERENCE_ROOT)
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

    # -- 7. PrRoIPool CPU path --------------------------------------------------------------------------------
    try:
        from oracle import prroi_oracle
        import ltr.external.PreciseRoIPooling.pytorch.prroi_pool.functional as prf
        prf._prroi_pooling = prroi_oracle.RefModuleCPU()
        prf._import_prroi_pooling = lambda: prf._prroi_pooling
        prroi_oracle.patch_reference_function(prf)
    except Exception as e:  # pragma: no cover - prroi oracle optional at early bring-up
        sys.stderr.write("[ref_shims] PrRoIPool CPU shim not installed: %r\n" % (e,))

    _installed = True
