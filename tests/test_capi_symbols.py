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


def test_kd_leaf_capacity_rule(P, monkeypatch):
    """The k-d index takes the leaf capacity (64 or 96 slots) that leaves fewer slots of its balanced tree's 2^L
    leaves empty (host arithmetic behind prgpu_diag_kd_leaf_capacity: no GPU needed)."""
    import ctypes as C
    monkeypatch.delenv("PRGPU_KD_PPL", raising=False)

    def cap(n):
        s, l = C.c_int(), C.c_int()
        assert P.lib().prgpu_diag_kd_leaf_capacity(n, C.byref(s), C.byref(l)) == 0
        return s.value, l.value

    assert cap(10_000_000) == (96, 17)            # 76 nodes per leaf: 79 % of 96 slots (64 slots: 60 % at depth 18)
    assert cap(1 << 17) == (64, 11)               # exactly full 64-slot leaves
    assert cap(1_250_000) == (96, 14)             # 76 per leaf again
    assert cap(128) == (64, 1) and cap(191) == (64, 2) and cap(192) == (96, 1)
    for n in (129, 1000, 4097, 300_017, 999_983, 5_000_000, 123_456_789):
        slots, lv = cap(n)
        per_leaf = -(-n // (1 << lv))             # ceil
        assert per_leaf <= slots and slots in (64, 96)
        fill = n / ((1 << lv) * slots)
        assert fill > 0.5                          # a fixed 64-slot leaf only guarantees > 0.5; the choice does better
        other = 96 if slots == 64 else 64
        lo = 0
        while -(-n // (1 << lo)) > other:
            lo += 1
        if n >= 2 * other:
            assert fill >= n / ((1 << lo) * other) - 1e-9
    monkeypatch.setenv("PRGPU_KD_PPL", "2")
    assert cap(10_000_000) == (64, 18)
    s, l = C.c_int(), C.c_int()
    assert P.lib().prgpu_diag_kd_leaf_capacity(5, C.byref(s), C.byref(l)) != 0
