"""CPU: the C-ABI library builds for sm_100a, loads, and exports every symbol include/b200chan.h declares."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "b200chan.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(rcb_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_all_exported(built_lib):
    names = _declared()
    assert len(names) >= 30
    for n in names:
        assert hasattr(built_lib, n), "missing export %s" % n
    from radiocapture_rf_b200 import _lib
    assert sorted(_lib._PROTOTYPES) == names, "ctypes prototypes out of sync with the header"


def test_error_strings_and_no_gpu_behaviour(built_lib):
    from radiocapture_rf_b200 import _lib
    assert built_lib.rcb_version() >= 100
    for code in range(-7, 1):
        assert built_lib.rcb_strerror(code)
    assert built_lib.rcb_open(0, None) == _lib.RCB_EINVAL
    n = C.c_int(-1)
    st = built_lib.rcb_device_count(C.byref(n))
    if st != 0 or n.value == 0:
        # CPU box: opening must fail cleanly (never a silent CPU fallback)
        h = C.c_void_p()
        assert built_lib.rcb_open(0, C.byref(h)) == _lib.RCB_ENODEV
        from radiocapture_rf_b200.engine import Engine
        with pytest.raises(_lib.B200ChanError):
            Engine(0)


def test_product_never_imports_oracle():
    """The product path must not route through the oracle (or any CPU fallback)."""
    pkg = os.path.join(ROOT, "radiocapture_rf_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), f
                assert "gr_cpu" not in txt, f


def test_sass_is_sm100a(built_lib):
    import subprocess
    from radiocapture_rf_b200 import _lib
    out = subprocess.run(["cuobjdump", "-lelf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out
