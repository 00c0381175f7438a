"""The C-ABI library loads without a GPU and exports every symbol include/prgpu.h declares."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "prgpu.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(prgpu_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported(P):
    P.build()
    L = ctypes.CDLL(P.lib_path())
    names = _declared()
    assert len(names) >= 20
    for n in names:
        assert hasattr(L, n), "missing export: " + n
    # the binding knows every declared symbol too
    from pr_ompl_b200 import capi
    assert set(names) == set(capi.SIGNATURES)


def test_no_cpu_fallback(P):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    assert P.lib().prgpu_device_count() == 0
    with pytest.raises(P.PrgpuError):
        P.KnnIndex(7)
    with pytest.raises(P.PrgpuError):
        P.World(*[x for x in (P.synth.default_arm() + P.synth.obstacles(3))])
