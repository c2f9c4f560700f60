"""Host logic of the batch search (cycleProcessing/batch.cpp:101-226): the rule
that picks the next good frame from the per-element match counts, in the Python mirror and in the
C++ drop-in unit, against a literal transcription of the reference's loop.  No GPU needed."""
import ctypes
import os

import numpy as np
import pytest

from slam_indoor_code_b200 import batch_search as bs
from slam_indoor_code_b200 import build


def reference_loop(sizes, required, first_fit, skip_head):
    """batch.cpp:120-148 with BatchElement::matches.size() given: size_t compared with int."""
    def as_size_t(v):
        return v + (1 << 64) if v < 0 else v
    good_index, good_size = -1, 0
    batch_index = len(sizes) - 1
    while batch_index >= skip_head:
        if sizes[batch_index] >= as_size_t(required) and sizes[batch_index] >= good_size:
            good_index, good_size = batch_index, sizes[batch_index]
            if first_fit:
                break
        batch_index -= 1
    return good_index


@pytest.fixture(scope="module")
def host():
    build.build()
    lib = ctypes.CDLL(os.path.join(build.LIBDIR, "libslamb200_hostshim.so"))
    lib.hostshim_select_good_frame.restype = ctypes.c_int
    lib.hostshim_select_good_frame.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int]
    return lib


def test_selection_rule_examples():
    # best fit: the LOWEST index among the elements with the most matches (ties replace, walking down)
    assert bs.selectGoodFrameFromMatchCounts([5, 9, 9, 3], 4, False) == 1
    # first fit: the first element from the end that reaches the requirement
    assert bs.selectGoodFrameFromMatchCounts([5, 9, 9, 3], 4, True) == 2
    assert bs.selectGoodFrameFromMatchCounts([5, 9, 9, 3], 3, True) == 3
    # nothing reaches it / empty batch / the head of the batch is not looked at
    assert bs.selectGoodFrameFromMatchCounts([5, 9, 9, 3], 10, False) == bs.FRAME_NOT_FOUND
    assert bs.selectGoodFrameFromMatchCounts([], 0, False) == bs.FRAME_NOT_FOUND
    assert bs.selectGoodFrameFromMatchCounts([50, 9, 9, 3], 4, False, skipFramesFromBatchHead=1) == 1
    # a requirement of zero accepts empty match lists too
    assert bs.selectGoodFrameFromMatchCounts([0, 0], 0, False) == 0
    # negative requirement: size_t >= int converts it to a huge unsigned number
    assert bs.selectGoodFrameFromMatchCounts([5, 9], -1, False) == bs.FRAME_NOT_FOUND


def test_selection_rule_equals_reference_loop(host):
    rng = np.random.default_rng(2025)
    for _ in range(3000):
        n = int(rng.integers(0, 12))
        sizes = [int(v) for v in rng.integers(0, 8, n)]
        required = int(rng.integers(-1, 9))
        first_fit = bool(rng.integers(0, 2))
        skip = int(rng.integers(0, 4))
        want = reference_loop(sizes, required, first_fit, skip)
        assert bs.selectGoodFrameFromMatchCounts(sizes, required, first_fit, skip) == want
        arr = np.asarray(sizes, np.int64)
        got = host.hostshim_select_good_frame(arr.ctypes.data_as(ctypes.c_void_p), n, required, int(first_fit), skip)
        assert got == want
