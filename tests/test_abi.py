"""The C-ABI library: it loads, exports every function include/gbdr.h declares (and nothing the
header does not declare is relied on by the Python host), and fails loudly without a GPU.  No compute
calls are made here."""
import os
import re
import subprocess

import pytest

from gbnns_dim_red_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_functions():
    src = open(os.path.join(ROOT, "include", "gbdr.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(gbdr_[a-z0-9_]+)\s*\(", src)))


def test_header_and_python_symbol_lists_agree():
    assert _header_functions() == sorted(capi.SYMBOLS)


def test_library_exports_every_declared_symbol():
    lib = capi.lib()
    for name in _header_functions():
        assert hasattr(lib, name), f"libgbdr.so does not export {name}"
    out = subprocess.run(["nm", "-D", "--defined-only", capi.LIB_PATH], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (gbdr_[a-z0-9_]+)", out))
    assert exported == set(_header_functions())


def test_library_is_sm100a_only():
    out = subprocess.run(["cuobjdump", "-lelf", capi.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_header_compiles_as_plain_c(tmp_path):
    c = tmp_path / "t.c"
    c.write_text('#include "gbdr.h"\nint main(void){return gbdr_version() == GBDR_VERSION ? 0 : 1;}\n')
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), "-c", str(c), "-o",
                    str(tmp_path / "t.o")], check=True)


def test_version():
    assert capi.lib().gbdr_version() == 100


def test_no_device_means_loud_failure_not_fallback():
    if capi.device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(capi.GbdrError) as e:
        capi.Index(0)
    assert e.value.code == -2  # GBDR_E_NO_DEVICE
    assert "no CPU fallback" in str(e.value)
    import numpy as np

    with pytest.raises(capi.GbdrError):
        capi.knn(np.zeros((4, 4), np.float32), np.zeros((4, 4), np.float32), 1)


def test_product_never_references_the_oracle():
    """Nothing under gbnns_dim_red_b200/ may import, link or open anything under oracle/."""
    pkg = os.path.join(ROOT, "gbnns_dim_red_b200")
    for dirpath, _, files in os.walk(pkg):
        if "build" in dirpath.split(os.sep)[-1:]:
            continue
        for f in files:
            if not f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".hpp")):
                continue
            text = open(os.path.join(dirpath, f), errors="replace").read()
            if f == "build.py":  # builds the checker (allowed), never loads it
                assert "CDLL" not in text
                continue
            assert "liboracle" not in text and "_oracle" not in text and "oracle/" not in text, f
    out = subprocess.run(["ldd", capi.LIB_PATH], capture_output=True, text=True).stdout
    assert "oracle" not in out
