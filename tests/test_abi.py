"""The C-ABI library loads without a GPU and exports every symbol include/*.h declares."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "cusumtools_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ct_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_entry_points():
    syms = declared_symbols()
    for must in ("ct_filtfilt_u16", "ct_filtfilt_f32", "ct_detect_f32", "ct_block_stats_f32", "ct_last_error"):
        assert must in syms


def test_library_exports_every_declared_symbol():
    from cusumtools_b200 import _lib
    _lib.build()
    L = ctypes.CDLL(_lib.LIB_PATH)
    missing = [s for s in declared_symbols() if not hasattr(L, s)]
    assert not missing, missing


def test_binding_covers_header():
    from cusumtools_b200 import _lib
    assert sorted(_lib.SIGNATURES) == declared_symbols()


def test_version_and_struct_layout():
    from cusumtools_b200 import _lib
    L = _lib.lib()
    assert L.ct_version() == 2
    assert L.ct_filter_seq_tile() == 64
    # struct: 2 int32 + (3*5 + 11 + 1) floats
    assert ctypes.sizeof(_lib.CtFilterCoef) == 8 + 4 * (15 + 11 + 1)


def test_argument_errors_are_reported_without_gpu():
    from cusumtools_b200 import _lib
    L = _lib.lib()
    coef = _lib.CtFilterCoef()
    coef.nsec = 4
    rc = L.ct_filtfilt_u16(None, 10, 1000, 0.0, 0xFFFC, 1.0, 0.0, ctypes.byref(coef), 238, 0, None, None, 0, None, None)
    assert rc == -1 and b"bad argument" in L.ct_last_error()
    assert L.ct_filtfilt_workspace_bytes(10_000_000, 1000, 238) >= 4 * 10_001_000
    with pytest.raises(ValueError):
        _lib.check(rc, "ct_filtfilt_u16")


def test_product_refuses_cpu_tensors():
    import torch
    from cusumtools_b200 import filters, synth
    raw = torch.zeros(1024, dtype=torch.uint16)
    with pytest.raises(RuntimeError, match="no CPU path"):
        filters.dequant_filtfilt(raw, synth.CHIMERA_SETTINGS, 1e5, 8, median_codes=(0, 0))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "cusumtools_b200")
    for dp, _, fns in os.walk(pkg):
        for fn in fns:
            if fn.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dp, fn)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), fn
                assert "scipy.signal" not in txt or fn == "design.py" or "import scipy" not in txt, fn
