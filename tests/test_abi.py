"""The C-ABI library loads without a GPU and exports every symbol include/amx.h declares."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _built_lib():
    from automix_b200 import build

    return build.build()


def test_library_exports_every_declared_symbol():
    path = _built_lib()
    L = ctypes.CDLL(path)
    hdr = open(os.path.join(ROOT, "include", "amx.h")).read() + open(os.path.join(ROOT, "include", "amx_layout.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(amx_[a-z0-9_]+)\s*\(", hdr))
    declared -= {"amx_scalar_fn", "amx_batched_fn"}
    assert len(declared) > 25
    missing = [s for s in sorted(declared) if not hasattr(L, s)]
    assert not missing, f"declared in include/*.h but not exported: {missing}"
    from automix_b200 import _lib

    assert set(_lib.EXPORTS) <= declared | {"amx_fam_plan", "amx_fam_pack"}


def test_no_cpu_fallback_without_device():
    """Without a CUDA device every compute entry point must fail loudly (AMX_ENODEV)."""
    from automix_b200 import _lib

    L = _lib.lib()
    if L.amx_device_count() > 0:
        return  # on the GPU box this property is not observable
    import numpy as np
    import pytest

    with pytest.raises(_lib.AmxError, match="no CUDA device"):
        _lib.mix_logpdf([1.0], [0.0], [1.0], np.zeros((4, 1)))
    with pytest.raises(_lib.AmxError):
        _lib.Target(dict(kind="quad", dims=[1], center=[0.0], scale=[1.0]))
