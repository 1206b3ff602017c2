"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol that
include/nicp_b200.h declares, struct layouts match, and -- with no GPU -- it fails loudly instead of
falling back to a CPU path.  No compute calls are made here."""
import ctypes as C
import os
import re

import pytest

from conftest import ROOT
from g2o_frontend_b200 import capi


def header_symbols():
    text = open(os.path.join(ROOT, "include", "nicp_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(nicp_[a-z0-9_]+)\s*\(", text)))


@pytest.mark.parametrize("verify", [False, True])
def test_library_exports_every_declared_symbol(verify):
    L = capi.load(verify)
    names = header_symbols()
    assert len(names) >= 25
    for n in names:
        assert hasattr(L, n), n
    assert sorted(capi.SYMBOLS) == names
    assert bool(L.nicp_is_verification_build()) == verify


def test_struct_layouts():
    assert C.sizeof(capi.AlignResult) == 256
    assert C.sizeof(capi.Projector) == 9 * 4 + 4 * 4
    assert C.sizeof(capi.StatsParams) == 6 * 4 + 9 * 4
    assert C.sizeof(capi.AlignParams) == 8 * 4
    assert C.sizeof(capi.Prior) == 4 + (16 + 16 + 36) * 4


def test_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    L = capi.load()
    h = C.c_void_p()
    rc = L.nicp_create(0, C.byref(h))
    assert rc != 0 and not h
    assert b"no CPU fallback" in L.nicp_last_error()
    with pytest.raises(capi.NicpError):
        capi.Context(0)


def test_missing_library_raises(monkeypatch, tmp_path):
    monkeypatch.setattr(capi, "LIB_DIR", str(tmp_path))
    monkeypatch.setattr(capi, "_LIBS", {})
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        capi.load()


def test_product_does_not_import_oracle():
    """only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline legs may touch oracle/"""
    pkg = os.path.join(ROOT, "g2o_frontend_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                text = open(os.path.join(dirpath, f)).read()
                assert "pwn_oracle" not in text and "liboracle" not in text, os.path.join(dirpath, f)
    for sub in ("include", "tools", "integration"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, sub)):
            for f in files:
                text = open(os.path.join(dirpath, f)).read()
                assert "pwn_oracle" not in text and "liboracle" not in text and "voxel_oracle" not in text, os.path.join(dirpath, f)
