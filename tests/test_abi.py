"""The C-ABI library loads here (no GPU) and exports every symbol that
include/semb200.h declares; error paths that need no device work."""

import ctypes as C
import os
import re

import pytest

from sempy_b200 import _capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    hdr = open(os.path.join(ROOT, "include", "semb200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(semb_[a-z0-9_]+)\s*\(", hdr)) - {"semb_apply_fn"})


def test_library_exports_every_declared_symbol():
    lib = C.CDLL(_capi.lib_path())
    names = _declared()
    assert len(names) >= 25
    for name in names:
        assert hasattr(lib, name), f"{name} declared in semb200.h but not exported"
    assert set(names) == set(_capi.SIGNATURES), "ctypes table and header disagree"


def test_version_and_argument_errors():
    lib = _capi.lib()
    assert lib.semb_version() >= 100
    h = C.c_void_p()
    assert lib.semb_create(C.byref(h), 0, 0, 8) != 0
    assert b"num_elems" in lib.semb_last_error()
    assert lib.semb_create(C.byref(h), 0, 4, 40) != 0
    assert b"Nq" in lib.semb_last_error()
    with pytest.raises(RuntimeError, match="semb200 error"):
        _capi.check(lib.semb_set_stream(None, None))


def test_no_cpu_fallback_in_product():
    """The package must not import the oracle or route around the CUDA library."""
    pkg = os.path.join(ROOT, "sempy_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert "oracle" not in src.replace("no CPU fallback", ""), fn


def test_ctypes_signatures_match_header_arity():
    """Every prototype in semb200.h has as many parameters as its ctypes entry (a mismatch would corrupt
    the call stack only at run time, on the GPU box)."""
    hdr = open(os.path.join(ROOT, "include", "semb200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    protos = re.findall(r"\b(semb_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", hdr, flags=re.S)
    seen = 0
    for name, params in protos:
        if name == "semb_apply_fn":
            continue
        params = params.strip()
        n = 0 if params in ("", "void") else len([p for p in params.split(",") if p.strip()])
        assert name in _capi.SIGNATURES, name
        assert len(_capi.SIGNATURES[name][1]) == n, f"{name}: header has {n} parameters, ctypes table {len(_capi.SIGNATURES[name][1])}"
        seen += 1
    assert seen == len(_capi.SIGNATURES)


def test_build_targets_sm_100a():
    sh = open(os.path.join(ROOT, "sempy_b200", "csrc", "build.sh")).read()
    assert "arch=compute_100a,code=sm_100a" in sh and "-lineinfo" in sh
    entry = open(os.path.join(ROOT, "__graft_entry__.py")).read()
    assert "build.sh" in entry
