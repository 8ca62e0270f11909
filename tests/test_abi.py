"""CPU tests of the boundary: the shared library loads, exports every symbol include/*.h declares, and refuses to run
without a GPU (no CPU fallback)."""
import ctypes
import glob
import os
import re

import numpy as np
import pytest

import gxbeam_b200
from gxbeam_b200 import api

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    names = set()
    for h in glob.glob(os.path.join(ROOT, "include", "*.h")):
        src = open(h).read()
        src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
        names |= set(re.findall(r"\b(gxb_[a-z0-9_]+)\s*\(", src))
    return names


def test_library_exports_every_declared_symbol():
    path = api.library_path()
    assert os.path.exists(path), "build the CUDA library first: python -c 'import __graft_entry__ as g; g.build()'"
    lib = ctypes.CDLL(path)
    declared = _declared_symbols()
    assert len(declared) >= 15
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/ but not exported"
    assert declared == set(api.EXPORTS), "api.EXPORTS must list exactly the declared entry points"


def test_struct_layouts_match_header():
    lib = api.load_library()
    o = api.Opts()
    lib.gxb_default_opts(ctypes.byref(o))
    assert (o.linear, o.two_dimensional, o.ftol, o.iterations, o.maxstep, o.c1, o.rho_hi, o.rho_lo, o.profile) == \
        (0, 0, 1e-9, 1000, 1e6, 1e-4, 0.5, 0.1, 0)


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(api.GXBError, match="no usable CUDA device|CUDA"):
        api.Context(0)


def test_product_package_does_not_import_oracle():
    pkg = os.path.join(ROOT, "gxbeam_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".sh")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "pyoracle" not in txt and "libgxb_oracle" not in txt and "oracle/" not in txt.replace("no CPU fallback", ""), f
