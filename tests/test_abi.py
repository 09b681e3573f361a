"""The C-ABI library on the CPU (-m "not gpu"): it loads, exports every symbol the header declares,
and refuses to run without a B200 (no compute calls here)."""
import ctypes
import os
import re

import pytest

from tlxcv_b200 import runtime

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_functions():
    src = open(os.path.join(ROOT, "include", "tlxcv_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(tlxcv_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = runtime.load_library()
    names = _header_functions()
    assert len(names) >= 15
    for n in names:
        assert hasattr(lib, n), f"{n} is declared in include/tlxcv_b200.h but not exported"
    assert sorted(runtime.EXPORTS) == names, "ctypes table and header disagree"
    assert lib.tlxcv_abi_version() == runtime.ABI_VERSION


def test_struct_layouts_match_header():
    assert ctypes.sizeof(runtime.TensorDesc) == 24
    assert ctypes.sizeof(runtime.OpDesc) == 14 * 4 + 6 * 8 + 8
    assert ctypes.sizeof(runtime.OpInfo) == 48 + 8 + 16 + 16


def test_no_device_is_an_error_not_a_fallback():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    lib = runtime.load_library()
    h = ctypes.c_void_p()
    rc = lib.tlxcv_create(0, ctypes.byref(h))
    assert rc == -4 and not h.value
    assert b"no CPU fallback" in lib.tlxcv_last_error(None)
    assert lib.tlxcv_plan_run(None, None, None, None, 0) == -1


def test_missing_library_fails_loudly(monkeypatch):
    monkeypatch.setattr(runtime, "_lib", None)
    monkeypatch.setattr(runtime, "_LIB_PATH", "/nonexistent/libtlxcv_b200.so")
    with pytest.raises(runtime.B200RuntimeError, match="no CPU or eager fallback"):
        runtime.load_library()


def test_library_is_sm100a_native():
    """SASS carries the Blackwell mnemonics (tcgen05.mma -> UTCHMMA, TMA -> UTMALDG, tcgen05.ld -> LDTM)."""
    import shutil
    import subprocess

    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not on PATH")
    out = subprocess.run(["cuobjdump", "-sass", runtime._LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out
    for mnem in ("UTCHMMA", "UTMALDG.4D.IM2COL", "UTMALDG.2D", "LDTM"):
        assert mnem in out, mnem
    assert "HMMA.16816" not in out, "legacy mma.sync tensor path present"
