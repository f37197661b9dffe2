"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol include/mtfjsp.h
declares (no compute calls without a GPU), and argument errors come back as codes, not crashes."""
import ctypes as C
import importlib
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def libmod():
    mod = importlib.import_module("e2e-mappo-for-mt-fjsp_b200._lib")
    mod.build()
    return mod


def test_header_symbols_are_exported(libmod):
    hdr = open(os.path.join(ROOT, "include", "mtfjsp.h")).read()
    declared = set(re.findall(r"\b(mtfjsp_[a-z0-9_]+)\s*\(", hdr))
    declared.discard("mtfjsp_env")
    assert len(declared) >= 20
    lib = C.CDLL(libmod.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(lib, name), name
    assert declared == set(libmod.SIGNATURES), declared ^ set(libmod.SIGNATURES)


def test_bad_arguments_return_codes(libmod):
    lib = libmod.lib()
    assert lib.mtfjsp_version().startswith(b"mtfjsp-b200")
    h = C.c_void_p()
    assert lib.mtfjsp_create(C.byref(h), 4, 6, 1, 2, 1, 0) == -1      # M < 2
    assert lib.mtfjsp_create(C.byref(h), 4, 600, 20, 2, 1, 0) == -1   # N > 4096
    assert lib.mtfjsp_create(None, 4, 6, 6, 2, 1, 0) == -1
    assert b"size out of range" in lib.mtfjsp_last_error() or b"null" in lib.mtfjsp_last_error()
    assert lib.mtfjsp_step(None, None, None, None, None, None, None, None) == -1
    assert lib.mtfjsp_destroy(None) == 0
    assert lib.mtfjsp_launch_count(None) == 0


def test_product_package_never_imports_the_oracle():
    pkg_dir = os.path.join(ROOT, "e2e-mappo-for-mt-fjsp_b200")
    for dirpath, _, files in os.walk(pkg_dir):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "mtfjsp_oracle" not in src and "from oracle" not in src and "import oracle" not in src, f
