"""CPU: libslamb200.so builds for sm_100a, loads without a GPU and exports every symbol that
include/slamb200.h declares.  No compute call is made here."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib_path():
    from slam_indoor_code_b200 import build
    return build.build()


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "slamb200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(slamb200_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported(lib_path):
    lib = ctypes.CDLL(lib_path)
    names = _declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in slamb200.h but not exported"


def test_binding_table_covers_header(lib_path):
    from slam_indoor_code_b200 import _capi
    assert sorted(_capi.SYMBOLS) == _declared_symbols()
    _capi.load()


def test_dmatch_is_cv_dmatch_layout():
    from slam_indoor_code_b200 import _capi
    assert _capi.DMATCH.itemsize == 16
    assert [_capi.DMATCH.fields[k][1] for k in ("queryIdx", "trainIdx", "imgIdx", "distance")] == \
        [0, 4, 8, 12]


def test_sass_is_sm100a(lib_path):
    import subprocess
    out = subprocess.run(["cuobjdump", "-lelf", lib_path], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_no_gpu_fails_loudly(lib_path):
    """Without a device the library reports an error instead of computing on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from slam_indoor_code_b200 import _capi
    from slam_indoor_code_b200.feature_matching import Context
    with pytest.raises(_capi.Slamb200Error) as e:
        Context(0)
    assert e.value.code == _capi.ERR_CUDA


def test_product_does_not_import_oracle():
    """The product package must not reference oracle/ in any way."""
    pkg = os.path.join(ROOT, "slam_indoor_code_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                s = open(os.path.join(dp, f), errors="replace").read()
                assert "import oracle" not in s and "from oracle" not in s and "liboracle" not in s, f
