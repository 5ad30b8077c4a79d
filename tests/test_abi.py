"""CPU tests of the drop-in boundary: the C-ABI library loads without a GPU, exports every symbol
include/icp_b200.h declares, the ctypes mirrors match the C structs, and the product path fails
loudly (no CPU fallback) when no CUDA device is present."""
import ctypes
import os
import re

import numpy as np
import pytest

import object_alignment_b200 as oab
from object_alignment_b200 import _native

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_functions():
    text = open(os.path.join(ROOT, "include", "icp_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(icp_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    names = _declared_functions()
    assert len(names) >= 25
    lib = _native.lib()
    for n in names:
        assert n in _native.SIGNATURES, "ctypes binding missing for %s" % n
        assert getattr(lib, n) is not None
    assert set(_native.SIGNATURES) == set(names)
    assert lib.icp_version() == 100


def test_struct_layouts_match_header():
    # icp_params: 2 x int32, 3 x double, 8 x int32 = 64 bytes; icp_iter_stats: int64 + 20 doubles
    assert ctypes.sizeof(_native.IcpParams) == 64
    assert _native.IcpParams.thresh.offset == 8
    assert _native.IcpParams.stride.offset == 32
    assert ctypes.sizeof(_native.IcpIterStats) == 8 + 20 * 8


def test_no_cpu_fallback_without_gpu():
    lib = _native.lib()
    if lib.icp_device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(oab.IcpError):
        oab.IcpEngine()
    with pytest.raises(oab.IcpError):
        oab.affine_matrix_from_points(np.zeros((3, 5)), np.ones((3, 5)), shear=False)
    with pytest.raises(TypeError):
        oab.make_pairs(None, None, object(), [0], 0.5)      # a CPU BVHTree is not accepted


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "object_alignment_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
                assert "liboracle" not in src, f
