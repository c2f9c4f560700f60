"""CPU: the host-side narrowing used by slamb200_upload_desc_packed (csrc/host_pack.cpp) accepts
exactly the Mats whose rows are integers in [0,255] with squared norm below 2^20 -- the exact-mode
condition the device applies (sift_prep.cu) -- and reproduces their bytes."""
import ctypes

import numpy as np
import pytest

from oracle import synth
from slam_indoor_code_b200 import _capi


@pytest.fixture(scope="module")
def pack():
    lib = _capi.load()
    fn = lib.slamb200_host_pack_u8
    fn.restype = ctypes.c_int
    fn.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_void_p]

    def run(a, stride_floats=None):
        a = np.asarray(a, np.float32)
        out = np.full((a.shape[0], 128), 0xAB, np.uint8)
        ok = fn(a.ctypes.data, stride_floats or a.strides[0] // 4, a.shape[0], out.ctypes.data)
        return ok, out
    return run


def test_integer_rows_are_packed_exactly(pack):
    q = synth.sift_like(1003, 5)
    ok, out = pack(q)
    assert ok == 1 and np.array_equal(out, q.astype(np.uint8))
    # full byte range, one value per position
    a = np.zeros((256, 128), np.float32)
    a[:, 0] = np.arange(256)
    a[np.arange(128), np.arange(128)] = 255
    ok, out = pack(a)
    assert ok == 1 and np.array_equal(out, a.astype(np.uint8))


def test_row_pitch(pack):
    wide = np.zeros((300, 160), np.float32)
    wide[:, :128] = synth.sift_like(300, 6)
    wide[:, 128:] = 0.5                                  # the padding must not be looked at
    ok, out = pack(wide[:, :128], 160)
    assert ok == 1 and np.array_equal(out, wide[:, :128].astype(np.uint8))


@pytest.mark.parametrize("value", [0.5, 254.99998, -1.0, -0.5, 256.0, 1e9, -1e9, 3e38, np.inf, -np.inf, np.nan,
                                   1e-30, 255.00002])
@pytest.mark.parametrize("pos", [(0, 0), (17, 31), (999, 127), (500, 64)])
def test_anything_else_is_refused(pack, value, pos):
    q = synth.sift_like(1000, 7)
    q[pos] = value
    ok, _ = pack(q)
    assert ok == 0


def test_negative_zero_and_norm_limit(pack):
    q = synth.sift_like(64, 8)
    q[3, 3] = -0.0                                       # an integer in range: accepted, byte 0
    ok, out = pack(q)
    assert ok == 1 and out[3, 3] == 0
    big = np.zeros((2, 128), np.float32)
    big[0, :16] = 255                                    # 16 * 255^2 = 1 040 400 < 2^20
    ok, _ = pack(big)
    assert ok == 1
    big[1, :17] = 255                                    # 17 * 255^2 = 1 105 425 >= 2^20
    ok, _ = pack(big)
    assert ok == 0
