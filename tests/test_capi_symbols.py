"""The C-ABI library loads on a CPU-only box and exports every symbol include/dxb.h declares (no compute calls)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_functions():
    txt = open(os.path.join(ROOT, "include", "dxb.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    names = set(re.findall(r"\b(dxb_[a-z0-9_]+)\s*\(", txt))
    return sorted(names)


def test_header_and_binding_table_agree():
    from opendxmc_b200 import _capi
    declared = _declared_functions()
    assert len(declared) > 50
    missing = [n for n in declared if n not in _capi.SIGNATURES]
    extra = [n for n in _capi.SIGNATURES if n not in declared]
    assert not missing, f"declared in dxb.h but not bound: {missing}"
    assert not extra, f"bound but not declared in dxb.h: {extra}"


def test_library_exports_every_declared_symbol():
    from opendxmc_b200 import _capi
    lib = ctypes.CDLL(_capi.LIB_PATH)
    for name in _declared_functions():
        assert hasattr(lib, name), f"libdxmc_b200.so does not export {name}"
    assert _capi.load().dxb_abi_version() == 2


def test_no_cpu_fallback_without_device():
    """dxb_create must fail with DXB_ECUDA when no CUDA device is present (the product has no CPU path)."""
    from opendxmc_b200 import _capi
    lib = _capi.load()
    if lib.dxb_device_count() > 0:
        return
    h = ctypes.c_void_p()
    assert lib.dxb_create(ctypes.byref(h), None, 0) == _capi.DXB_ECUDA
    assert not h.value


def test_product_does_not_link_or_reference_the_oracle():
    """only tests/, __graft_entry__.smoke() and bench.py may touch oracle/."""
    for dirpath, _, files in os.walk(os.path.join(ROOT, "opendxmc_b200")):
        for f in files:
            if f.endswith((".py", ".cpp", ".cu", ".hpp", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f), errors="replace").read()
                assert "oracle_py" not in txt and "liboracle" not in txt and "orc_" not in txt, f
    for f in os.listdir(os.path.join(ROOT, "include")):
        p = os.path.join(ROOT, "include", f)
        if os.path.isfile(p):
            assert "orc_" not in open(p).read()
