"""TEST INFRASTRUCTURE ONLY -- never imported by the product path.

`install()` = baseline/ref_env.py (the out-of-tree compatibility shims 1-6 of SURVEY.md 8(c) that let the UNMODIFIED reference
import on this image) + shim 7: the CPU restatement of PrRoIPool (oracle/prroi_oracle.py) bound at the reference's own
`_prroi_pooling` seam (ltr/external/PreciseRoIPooling/pytorch/prroi_pool/functional.py:18-38), because the reference's native
module neither builds on current torch nor has a CPU path.  Used by the golden generators and the CPU cross-checks.
"""
import sys

from baseline import ref_env
from baseline.ref_env import reference_available, reference_root  # noqa: F401

_installed = False


def install_prroi_cpu():
    from oracle import prroi_oracle
    import ltr.external.PreciseRoIPooling.pytorch.prroi_pool.functional as prf
    prf._prroi_pooling = prroi_oracle.RefModuleCPU()
    prf._import_prroi_pooling = lambda: prf._prroi_pooling
    prroi_oracle.patch_reference_function(prf)


def install(scratch_dir=None, prroi_cpu=True):
    """Install all shims and put the reference on sys.path. Idempotent."""
    global _installed
    ref_env.install(scratch_dir)
    if _installed or not prroi_cpu:
        return
    try:
        install_prroi_cpu()
    except Exception as e:  # pragma: no cover
        sys.stderr.write("[ref_shims] PrRoIPool CPU shim not installed: %r\n" % (e,))
    _installed = True
