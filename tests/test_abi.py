"""The C-ABI library loads and exports every symbol include/moonb200.h declares (no GPU needed)."""
import ctypes

import pytest

from moonrtx_b200 import _lib


def test_library_loads_and_exports_every_declared_symbol():
    lib = _lib.load()
    declared = _lib.header_symbols()
    assert len(declared) >= 40
    for name in declared:
        assert hasattr(lib, name), f"{name} is declared in include/moonb200.h but not exported"
    # and the binding table covers the header exactly
    assert sorted(_lib._SIGNATURES) == declared


def test_abi_version_and_error_string():
    lib = _lib.load()
    assert lib.mrtx_abi_version() == 1
    assert isinstance(lib.mrtx_last_error(), bytes)


def test_bad_arguments_are_rejected_without_a_gpu():
    lib = _lib.load()
    # null context -> MRTX_ERR_INVALID -> ValueError on the Python side
    rc = lib.mrtx_resize(None, 16, 16)
    assert rc == -1
    with pytest.raises(ValueError):
        _lib.check(rc)
    rc = lib.mrtx_set_float(None, b"scene_epsilon", 1e-4)
    assert rc == -1


def test_no_cpu_fallback_in_product_package():
    """The product package must not import the oracle (parity claims depend on it)."""
    import os
    import re
    root = os.path.dirname(os.path.abspath(_lib.__file__))
    for dirpath, _, files in os.walk(root):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f), encoding="utf-8").read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f
