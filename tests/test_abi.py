"""The C-ABI library loads and exports every symbol include/*.h declares; on a box without a GPU
compute entry points fail loudly (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

from cg_mrslam_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared(header):
    text = open(os.path.join(ROOT, "include", header)).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b((?:cgm|pgo)_[a-z0-9_]+)\s*\(", text)))


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as g
    if not os.path.exists(_lib.LIB_PATH):
        g.build()
    return _lib.load()


@pytest.mark.parametrize("header", sorted(h for h in os.listdir(os.path.join(ROOT, "include"))
                                          if h.endswith(".h")))
def test_exports_every_declared_symbol(lib, header):
    names = _declared(header)
    assert names, header
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing


def test_no_cpu_fallback(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by the gpu suite")
    from cg_mrslam_b200 import matcher
    with pytest.raises(matcher.MatcherError) as e:
        matcher.Matcher((-35.0, -35.0), (35.0, 35.0), 0.1, 0.5)
    assert e.value.code == -2  # CGM_ERR_CUDA


def test_product_does_not_import_oracle():
    """Nothing under cg_mrslam_b200/ may reference oracle/ (it is test infrastructure)."""
    bad = []
    for dirpath, _, files in os.walk(os.path.join(ROOT, "cg_mrslam_b200")):
        for f in files:
            if f.endswith((".py", ".cpp", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                if re.search(r"^\s*(from|import)\s+oracle|#include\s+\"[^\"]*oracle", text, re.M):
                    bad.append(f)
    assert not bad, bad
