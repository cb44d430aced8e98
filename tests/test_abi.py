"""The C-ABI library loads and exports every symbol include/drs_b200.h declares
(no compute calls: there is no GPU in the CPU test tier)."""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "drs_b200.h")


def _declared():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(drs_[a-z0-9_]+)\s*\(", text)))


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as g
    g.build()                                   # nvcc cross-compiles without a GPU
    from deeprecsys_b200 import _lib
    return _lib


def test_header_declares_the_expected_entry_points():
    names = _declared()
    for must in ("drs_create", "drs_destroy", "drs_load_table", "drs_load_fc", "drs_stage_batch",
                 "drs_infer", "drs_infer_staged", "drs_sparse_lengths_sum", "drs_fetch_stage",
                 "drs_last_error", "drs_shard_export", "drs_shard_connect"):
        assert must in names


def test_library_exports_every_declared_symbol(lib):
    out = subprocess.run(["nm", "-D", "--defined-only", lib.lib_path()], stdout=subprocess.PIPE,
                         text=True, check=True).stdout
    exported = set(re.findall(r" T (drs_[a-z0-9_]+)", out))
    declared = set(_declared())
    assert declared <= exported, sorted(declared - exported)
    # and nothing is exported that the header does not declare
    assert exported <= declared, sorted(exported - declared)


def test_ctypes_binding_covers_the_header(lib):
    assert set(lib.exported_symbols()) == set(_declared())
    handle = lib.load()
    assert handle.drs_abi_version() == lib.ABI_VERSION == 7
    assert isinstance(handle.drs_last_error(), bytes)


def test_sm100a_cubin_is_embedded(lib):
    out = subprocess.run(["cuobjdump", "-lelf", lib.lib_path()], stdout=subprocess.PIPE,
                         stderr=subprocess.STDOUT, text=True).stdout
    assert "sm_100a" in out, out


def test_no_cpu_fallback_when_library_missing(monkeypatch, tmp_path):
    from deeprecsys_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib._build, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(ImportError):
        _lib.load()


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "deeprecsys_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f
