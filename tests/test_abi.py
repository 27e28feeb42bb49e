"""The C-ABI library loads and exports every symbol include/heatflow_b200.h declares (no compute)."""
import ctypes
import os
import re

import pytest

from heatflow_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    text = open(os.path.join(ROOT, "include", "heatflow_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(hf_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_expected_surface():
    names = header_functions()
    for must in ("hf_create", "hf_destroy", "hf_set_mesh", "hf_set_materials", "hf_set_bcs", "hf_build_operator",
                 "hf_get_csr", "hf_set_state", "hf_step", "hf_run", "hf_get_state", "hf_sample", "hf_project_gradient",
                 "hf_ens_create", "hf_ens_run", "hf_last_error"):
        assert must in names


def test_library_exports_every_declared_symbol():
    if not os.path.isfile(_lib.LIB_PATH):
        pytest.fail(f"{_lib.LIB_PATH} missing - run `python __graft_entry__.py` (the driver's build() does)")
    lib = ctypes.CDLL(_lib.LIB_PATH)
    missing = [n for n in header_functions() if not hasattr(lib, n)]
    assert not missing, missing


def test_ctypes_signatures_cover_the_header():
    assert sorted(_lib.SIGNATURES) == header_functions()
    lib = _lib.load()
    assert lib.hf_version() >= 100
    assert isinstance(lib.hf_last_error(), bytes)


def test_no_cpu_fallback_without_a_device():
    lib = _lib.load()
    if lib.hf_device_count() > 0:
        pytest.skip("a CUDA device is present")
    from heatflow_b200.solver import HeatSolver
    with pytest.raises(_lib.HeatflowError, match="no CUDA device"):
        HeatSolver(0)


def test_product_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "heatflow_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("the oracle", ""), f"{f} references the oracle"
