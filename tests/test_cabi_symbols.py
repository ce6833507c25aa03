"""The C-ABI library loads (no GPU needed) and exports every function include/tclight.h declares."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    src = open(os.path.join(ROOT, "include", "tclight.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = re.findall(r"\b(tcl_[a-z0-9_]+)\s*\(", src)
    return sorted(set(names))


def test_library_exports_every_declared_symbol():
    from tclight_b200 import _lib

    names = declared_functions()
    assert len(names) >= 25
    missing = [n for n in names if not hasattr(_lib.lib, n)]
    assert not missing, f"libtclight.so lacks: {missing}"
    assert _lib.lib.tcl_version() >= 100
    assert _lib.lib.tcl_last_error() is not None


def test_argument_errors_are_reported_without_a_gpu():
    import ctypes as C
    from tclight_b200 import _lib

    rc = _lib.lib.tcl_igemm(None, None)
    assert rc != 0 and b"null descriptor" in _lib.lib.tcl_last_error()
    assert _lib.lib.tcl_postopt_workspace_bytes(720, 1280, 16) > 16 * 3 * 720 * 1280 * 4
    assert _lib.lib.tcl_postopt_pyramid_elems(720, 1280) == 360 * 640 + 180 * 320 + 90 * 160 + 45 * 80


def test_no_cpu_fallback():
    """Product ops refuse CPU tensors instead of silently computing elsewhere."""
    import pytest
    import torch
    from tclight_b200 import ops
    from tclight_b200._lib import TclError

    with pytest.raises(TclError):
        ops.layernorm(torch.zeros(4, 64, dtype=torch.float16), torch.ones(64), torch.zeros(64))
    with pytest.raises(TclError):
        ops.linear(torch.zeros(4, 64, dtype=torch.float16), torch.zeros(64, 64, dtype=torch.float16))


def test_product_never_imports_oracle():
    """Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline/reference legs may touch oracle/."""
    pkg = os.path.join(ROOT, "tclight_b200")
    offenders = []
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                txt = open(os.path.join(dp, f)).read()
                for m in re.finditer(r"^\s*(from|import)\s+oracle\b.*$", txt, flags=re.M):
                    # postopt.smoke_check is the smoke() leg
                    start = txt.rfind("\ndef ", 0, m.start())
                    fn = txt[start:start + 60]
                    if "smoke_check" not in fn:
                        offenders.append((f, m.group(0).strip()))
    assert not offenders, offenders
