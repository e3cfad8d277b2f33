"""CPU: the C-ABI library loads and exports exactly what include/deltakd.h declares (no compute calls)."""
import ctypes
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "deltakd.h")


def declared():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    out = {}
    for m in re.finditer(r"DKD_API\s+[\w\s\*]+?\b(dkd_\w+)\s*\(([^;]*?)\)\s*;", src, flags=re.S):
        args = m.group(2).strip()
        out[m.group(1)] = 0 if args in ("", "void") else len(args.split(","))
    return out


def test_header_declares_functions():
    d = declared()
    assert "dkd_logit_kd_fwdbwd" in d and "dkd_last_error" in d and len(d) >= 7


def test_library_loads_and_exports_every_symbol():
    from deltakd_b200 import _lib
    assert os.path.exists(_lib.LIB_PATH)
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared():
        assert hasattr(lib, name), f"{name} declared in deltakd.h but not exported"
    nm = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True).stdout
    exported = {ln.split()[-1] for ln in nm.splitlines() if " T " in ln and ln.split()[-1].startswith("dkd_")}
    assert exported == set(declared()), exported ^ set(declared())


def test_python_binding_table_matches_header():
    from deltakd_b200 import _lib
    d = declared()
    assert set(_lib.SIGNATURES) == set(d)
    for name, (_res, args) in _lib.SIGNATURES.items():
        assert len(args) == d[name], f"{name}: binding has {len(args)} args, header {d[name]}"


def test_version_and_error_string():
    from deltakd_b200 import _lib
    assert _lib.lib.dkd_version() >= 100
    assert isinstance(_lib.last_error(), str)


def test_no_cpu_fallback():
    """Without a CUDA device the product refuses to compute (and never touches oracle/)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from deltakd_b200 import _lib
    from deltakd_b200 import functional as Fn
    assert _lib.lib.dkd_check_device() == -4
    with pytest.raises(RuntimeError):
        Fn.logit_kd_loss(torch.zeros(2, 4), None, None, torch.zeros(2, 4), kd_kind="none")
    with pytest.raises(RuntimeError):
        Fn.mask_rank(torch.zeros(2, 4), 2)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "deltakd_b200")
    for dirpath, _dirs, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), f
