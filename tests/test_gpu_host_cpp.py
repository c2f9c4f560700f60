"""GPU: the C++ drop-in translation unit (host/featureMatchingB200.cpp) driven like the reference
drives featureMatchingCPU.cpp -- cv::Mat descriptors in, std::vector<cv::DMatch> out -- through
the cv_shim build.  Checker: the CPU oracle."""
import ctypes
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import c_oracle, synth
from slam_indoor_code_b200 import _capi, build


@pytest.fixture(scope="module")
def host():
    build.build()
    lib = ctypes.CDLL(os.path.join(build.LIBDIR, "libslamb200_hostshim.so"))
    lib.hostshim_set_ratio.argtypes = [ctypes.c_double]
    lib.hostshim_match_features.restype = ctypes.c_int
    lib.hostshim_match_features.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_size_t, ctypes.c_void_p,
                                            ctypes.c_int, ctypes.c_size_t, ctypes.c_int, ctypes.c_void_p,
                                            ctypes.c_int]
    lib.hostshim_match_batch.restype = ctypes.c_int
    lib.hostshim_match_batch.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p,
                                         ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_int,
                                         ctypes.c_void_p]
    return lib


def _match(host, q, t, mtype, ratio=0.7):
    host.hostshim_set_ratio(ratio)
    out = np.zeros(max(len(q), 1), _capi.DMATCH)
    n = host.hostshim_match_features(_capi.ptr(q), len(q), q.strides[0] if len(q) else 0, _capi.ptr(t),
                                     len(t), t.strides[0] if len(t) else 0, mtype, _capi.ptr(out), len(out))
    return n, out[: max(n, 0)]


@pytest.mark.parametrize("mtype", [0, 1, 2])
def test_match_features_cpp(host, mtype):
    q, t = synth.orb_pair(700, 900, 31) if mtype == 2 else synth.sift_pair(700, 900, 31)
    n, got = _match(host, q, t, mtype)
    assert n >= 0 and np.array_equal(got, c_oracle.match_features(mtype, q, t, 0.7))
    n, got = _match(host, q, t, mtype, ratio=0.9)
    assert np.array_equal(got, c_oracle.match_features(mtype, q, t, 0.9))


def test_cv_mat_step_and_empty_mats(host):
    q, t = synth.sift_pair(300, 400, 32)
    wide = np.zeros((400, 144), np.float32)
    wide[:, :128] = t
    n, got = _match(host, q, wide[:, :128], 0)
    assert np.array_equal(got, c_oracle.match_features(0, q, t, 0.7))
    e = np.zeros((0, 128), np.float32)
    assert _match(host, e, t, 0)[0] == 0          # empty query Mat -> matches cleared
    assert _match(host, q, e, 0)[0] == 0          # empty train Mat


def test_bad_matcher_type_throws(host):
    q, t = synth.sift_pair(10, 10, 33)
    assert _match(host, q, t, 5)[0] == -1         # featureMatchingCPU.cpp:37: throw std::exception()


def test_batch_fast_path_cpp(host):
    q = synth.sift_like(600, 34)
    trains = [synth.sift_train_from_query(q, n, 35 + i) for i, n in enumerate((500, 800, 64))]
    P = len(trains)
    host.hostshim_set_ratio(0.7)
    ptrs = (ctypes.c_void_p * P)(*[t.ctypes.data for t in trains])
    nt = np.array([len(t) for t in trains], np.int32)
    out = np.zeros((P, len(q)), _capi.DMATCH)
    n_out = np.zeros(P, np.int32)
    rc = host.hostshim_match_batch(_capi.ptr(q), len(q), ptrs, _capi.ptr(nt), P, 0, _capi.ptr(out), len(q),
                                   _capi.ptr(n_out))
    assert rc == 0
    for p in range(P):
        assert np.array_equal(out[p, : n_out[p]], c_oracle.match_features(0, q, trains[p], 0.7))


def test_extract_descriptor_orb_cpp(host):
    """extractDescriptor(frame, features, ORB_BF, desc) of the drop-in unit: descriptors on the B200,
    `features` pruned exactly like cv::ORB::compute prunes it."""
    host.hostshim_extract_orb.restype = ctypes.c_int
    host.hostshim_extract_orb.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_size_t,
                                          ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
    frame = synth.textured_frame(300, 400, 41, 3)
    rng = np.random.default_rng(42)
    kps = np.stack([rng.integers(0, 400, 1500), rng.integers(0, 300, 1500), np.full(1500, -1.0)], 1).astype(np.float32)
    kept_xy = np.zeros((1500, 2), np.float32)
    desc = np.zeros((1500, 32), np.uint8)
    n = host.hostshim_extract_orb(_capi.ptr(frame), 300, 400, 3, frame.strides[0], _capi.ptr(kps), 1500,
                                  _capi.ptr(kept_xy), _capi.ptr(desc), 1500)
    keep, want = c_oracle.orb_compute(frame, kps)
    assert n == len(want) > 500
    assert np.array_equal(kept_xy[:n], kps[keep.astype(bool), :2]) and np.array_equal(desc[:n], want)
    # nothing survives the border filter -> empty Mat, empty vector
    edge = np.array([[3, 3, -1], [399, 299, -1]], np.float32)
    assert host.hostshim_extract_orb(_capi.ptr(frame), 300, 400, 3, frame.strides[0], _capi.ptr(edge), 2,
                                     _capi.ptr(kept_xy), _capi.ptr(desc), 1500) == 0


def test_fast_extractor_cpp(host):
    """fastExtractor(frame, points, threshold) of host/fastExtractorB200.cpp: the keypoint vector the
    reference's callers get (batch.cpp:245), from the B200."""
    host.hostshim_fast_extractor.restype = ctypes.c_int
    host.hostshim_fast_extractor.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_size_t,
                                             ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_int]
    frame = synth.textured_frame(360, 480, 51, 3)
    want = c_oracle.fast_detect(frame, 10, True)
    out = np.zeros((len(want) + 8, 5), np.float32)
    n = host.hostshim_fast_extractor(_capi.ptr(frame), 360, 480, 3, frame.strides[0], 10, 1, _capi.ptr(out), len(out))
    assert n == len(want) > 1000
    assert np.array_equal(out[:n, :3], want)
    assert np.all(out[:n, 3] == 7.0) and np.all(out[:n, 4] == -1.0)
    # a frame too small to hold a corner, and a threshold nothing passes
    tiny = synth.textured_frame(6, 6, 52, 3)
    assert host.hostshim_fast_extractor(_capi.ptr(tiny), 6, 6, 3, tiny.strides[0], 10, 1, _capi.ptr(out), len(out)) == 0
    assert host.hostshim_fast_extractor(_capi.ptr(frame), 360, 480, 3, frame.strides[0], 255, 1, _capi.ptr(out), len(out)) == 0


def test_find_good_frame_from_batch_cpp_and_python(host, ctx):
    """The batch search of batch.cpp:101-226 in one matcher call: per-element match counts equal to
    the oracle's, the good index equal to the reference's rule, in the C++ unit and the Python mirror."""
    from slam_indoor_code_b200 import batch_search as bs
    from slam_indoor_code_b200.feature_matching import MatcherType
    host.hostshim_find_good_frame.restype = ctypes.c_int
    host.hostshim_find_good_frame.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p,
                                              ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                              ctypes.c_void_p]
    q = synth.sift_like(700, 61)
    # elements with different numbers of true correspondences (and one unrelated frame)
    trains = [synth.sift_train_from_query(q, 600 + 40 * i, 62 + i, planted=p)
              for i, p in enumerate((0.1, 0.4, 0.0, 0.4, 0.25))]
    counts = [len(c_oracle.match_features(0, q, t, 0.7)) for t in trains]
    host.hostshim_set_ratio(0.7)
    P = len(trains)
    ptrs = (ctypes.c_void_p * P)(*[t.ctypes.data for t in trains])
    nts = np.array([len(t) for t in trains], np.int32)
    Q = ctx.upload(q)
    Ts = [ctx.upload(t) for t in trains]
    for required, first_fit, skip in [(50, False, 0), (50, True, 0), (max(counts) + 1, False, 0),
                                      (1, False, 2), (counts[4], True, 0)]:
        want = bs.selectGoodFrameFromMatchCounts(counts, required, first_fit, skip)
        n_out = np.zeros(P, np.int32)
        good = host.hostshim_find_good_frame(_capi.ptr(q), len(q), ptrs, _capi.ptr(nts), P, 0, required,
                                             int(first_fit), skip, _capi.ptr(n_out))
        assert good == want and list(n_out) == counts
        g2, all_matches = bs.findGoodFrameFromBatch(ctx, Q, Ts, MatcherType.SIFT_BF, required, first_fit, skip)
        assert g2 == want and [len(m) for m in all_matches] == counts
