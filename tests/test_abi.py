"""The C-ABI library loads and exports every symbol include/pz.h declares
(no compute calls: runs without a GPU)."""
import ctypes
import os
import re

from conftest import ROOT


def header_symbols():
    text = open(os.path.join(ROOT, "include", "pz.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(pz_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_entry_points():
    syms = header_symbols()
    for must in ("pz_create", "pz_set_graph", "pz_run_rows", "pz_run_fused", "pz_set_ps",
                 "pz_convolve", "pz_micro_finalize", "pz_canon_export", "pz_make_perms"):
        assert must in syms


def test_library_exports_every_declared_symbol():
    from pypercolate_b200 import _native
    lib = ctypes.CDLL(_native.LIB_PATH)
    missing = [s for s in header_symbols() if not hasattr(lib, s)]
    assert not missing, missing


def test_python_binding_lists_the_same_symbols():
    from pypercolate_b200 import _native
    assert sorted(_native.SYMBOLS) == header_symbols()
    _native.load()


def test_no_torch_types_in_signatures():
    text = open(os.path.join(ROOT, "include", "pz.h")).read()
    assert "torch" not in text.lower().replace("no torch", "")
    assert "at::" not in text and "Tensor" not in text


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "pypercolate_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, f
                assert "pz_oracle" not in src.replace("oracle/pz_oracle.c restates", ""), f
