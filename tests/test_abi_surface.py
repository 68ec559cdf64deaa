"""CPU-side checks of the drop-in boundary: the shared library loads without a GPU, exports
every symbol include/scf_b200.h declares, the ctypes table covers the same set, and the
product refuses to run without a device (no CPU fallback)."""
import os
import re

import pytest

import scf_b200
from scf_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "scf_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(scfb_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported_and_bound():
    lib = capi.lib()
    names = _declared()
    assert len(names) >= 15
    for name in names:
        assert hasattr(lib, name), f"{name} declared in scf_b200.h but not exported"
    assert sorted(capi._SIGNATURES) == names, "ctypes table and header disagree"


def test_version():
    assert capi.lib().scfb_version() >= 100


def test_product_does_not_import_the_oracle():
    """Only tests/, __graft_entry__.smoke() and bench.py's CPU baseline may touch oracle/."""
    for top in ("selfconsistentfield.jl_b200", "examples", "benchmarks", "include"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, top)):
            for f in files:
                if f.endswith((".py", ".cu", ".h", ".jl")):
                    src = open(os.path.join(dirpath, f)).read()
                    assert not re.search(r"^\s*(from|import)\s+oracle", src, flags=re.M), f
                    assert "scf_oracle" not in src and "c_oracle" not in src, f
    shim = open(os.path.join(ROOT, "scf_b200.py")).read()
    assert "oracle" not in shim


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(scf_b200.ScfbError) as ei:
        scf_b200.JKBuilderCUDA(8)
    assert ei.value.code == capi.ERR_CUDA


def test_argument_errors_need_no_device():
    import ctypes as C
    lib = capi.lib()
    h = capi._h()
    assert lib.scfb_create(C.byref(h), 0, 0, 0, 0, 1) == capi.ERR_ARG
    assert b"n_bas" in lib.scfb_last_error(None)
    assert lib.scfb_create(C.byref(h), 8, 7, 0, 0, 1) == capi.ERR_ARG
    assert lib.scfb_create(C.byref(h), 8, 0, 0, 3, 2) == capi.ERR_ARG
    assert lib.scfb_destroy(None) == capi.OK
    assert lib.scfb_sync(None) == capi.ERR_ARG


def test_abstract_builder_fallback_raises():
    # src/types/ScfProblem.jl:4-9
    with pytest.raises(NotImplementedError):
        scf_b200.TwoElectronBuilder().add_2e_term(None, None, None)
