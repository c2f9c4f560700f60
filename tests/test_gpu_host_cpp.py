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


# ---- the other units of the drop-in (cameraTranslationB200.cpp, poseEstimationB200.cpp,
# ---- triangulateB200.cpp), compiled against cv_shim.h with cv2's CPU solvers injected ---------------
K4 = synth.SAMSUNG_HV_4K
KMAT = np.array([[K4[0], 0, K4[2]], [0, K4[1], K4[3]], [0, 0, 1.0]])


@pytest.fixture(scope="module")
def solvers(host):
    """cv2's 5-point solver, solvePnP and Rodrigues behind the shim's injection points: the C++
    control flow (RANSAC loop, update rule, refit) runs around OpenCV's own minimal solvers."""
    cv2 = pytest.importorskip("cv2")
    dp, fp = ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_float)
    F5 = ctypes.CFUNCTYPE(ctypes.c_int, fp, fp, dp, dp)
    PNP = ctypes.CFUNCTYPE(ctypes.c_int, dp, dp, ctypes.c_int, dp, dp, ctypes.c_int, dp, dp, ctypes.c_int, ctypes.c_int)
    ROD = ctypes.CFUNCTYPE(None, dp, dp)

    def five(a, b, K9, out):
        pa = np.ctypeslib.as_array(a, (5, 2)).copy()
        pb = np.ctypeslib.as_array(b, (5, 2)).copy()
        K = np.ctypeslib.as_array(K9, (3, 3)).copy()
        E = cv2.findEssentialMat(pa, pb, K, cv2.RANSAC, 0.999, 1.0)[0]
        if E is None:
            return 0
        E = np.asarray(E, np.float64).reshape(-1, 9)
        np.ctypeslib.as_array(out, (90,))[: E.size] = E.reshape(-1)
        return E.shape[0]

    def pnp(obj, img, n, K9, dist, nd, rvec, tvec, guess, method):
        o = np.ctypeslib.as_array(obj, (n, 3)).copy()
        m = np.ctypeslib.as_array(img, (n, 2)).copy()
        K = np.ctypeslib.as_array(K9, (3, 3)).copy()
        d = np.ctypeslib.as_array(dist, (nd,)).copy() if nd else None
        r, t = np.ctypeslib.as_array(rvec, (3,)), np.ctypeslib.as_array(tvec, (3,))
        flags = {0: cv2.SOLVEPNP_ITERATIVE, 1: cv2.SOLVEPNP_EPNP, 2: cv2.SOLVEPNP_P3P}[method]
        if guess:
            ok, rr, tt = cv2.solvePnP(o, m, K, d, r.copy().reshape(3, 1), t.copy().reshape(3, 1), True, flags)
        else:
            # the minimal solvers see the float points of the caller (cv::Point3f / Point2f)
            ok, rr, tt = cv2.solvePnP(o.astype(np.float32), m.astype(np.float32), K, d, flags=flags)
        if not ok:
            return 0
        r[:] = rr.reshape(3)
        t[:] = tt.reshape(3)
        return 1

    def rod(r3, R9):
        R = cv2.Rodrigues(np.ctypeslib.as_array(r3, (3,)).copy())[0]
        np.ctypeslib.as_array(R9, (9,))[:] = R.reshape(9)

    keep = (F5(five), PNP(pnp), ROD(rod))
    host.hostshim_set_solvers.argtypes = [F5, PNP, ROD]
    host.hostshim_set_solvers(*keep)
    yield keep


@pytest.mark.parametrize("m,noise,outl,seed", [(600, 0.7, 0.3, 7100), (4000, 0.7, 0.3, 7101), (2500, 3.0, 0.6, 7102)])
def test_find_essential_mat_cpp_unit(host, solvers, m, noise, outl, seed):
    """host/cameraTranslationB200.cpp::findEssentialMatB200 == cv2.findEssentialMat, E and mask bit
    for bit (the seam of cameraTranslation.cpp:41-46)."""
    import cv2
    p1, p2, _, _ = synth.two_view(m, seed, noise_px=noise, outliers=outl)
    Ecv, mcv = cv2.findEssentialMat(p1, p2, KMAT, cv2.RANSAC, 0.999, 5.0)
    E = np.zeros(9)
    mask = np.zeros(m, np.uint8)
    host.hostshim_find_essential_mat.restype = ctypes.c_int
    host.hostshim_find_essential_mat.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p,
                                                 ctypes.c_double, ctypes.c_double, ctypes.c_void_p, ctypes.c_void_p]
    rc = host.hostshim_find_essential_mat(_capi.ptr(p1), _capi.ptr(p2), m, _capi.ptr(np.ascontiguousarray(KMAT)),
                                          0.999, 5.0, _capi.ptr(E), _capi.ptr(mask))
    assert rc == 1
    assert np.array_equal(E.reshape(3, 3), np.asarray(Ecv, np.float64).reshape(3, 3))
    assert np.array_equal(mask, mcv.reshape(-1))


@pytest.mark.parametrize("m,seed", [(800, 7200), (5000, 7201)])
def test_solve_pnp_ransac_cpp_unit(host, solvers, m, seed):
    """host/poseEstimationB200.cpp::solvePnPRansacB200 == cv2.solvePnPRansac (mainCycle.cpp:155-159)."""
    import cv2
    obj, img, _, _ = synth.pnp_scene(m, seed)
    dist = np.array(synth.REF_DIST5, np.float64)
    ok, rcv, tcv, icv = cv2.solvePnPRansac(obj, img, KMAT, dist)
    r, t = np.zeros(3), np.zeros(3)
    inl = np.zeros(m, np.int32)
    n = ctypes.c_int(0)
    host.hostshim_solve_pnp_ransac.restype = ctypes.c_int
    host.hostshim_solve_pnp_ransac.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p,
                                               ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p,
                                               ctypes.c_void_p, ctypes.c_void_p]
    rc = host.hostshim_solve_pnp_ransac(_capi.ptr(obj), _capi.ptr(img), m, _capi.ptr(np.ascontiguousarray(KMAT)),
                                        _capi.ptr(dist), 5, _capi.ptr(r), _capi.ptr(t), _capi.ptr(inl), ctypes.byref(n))
    assert rc == 1 and ok
    assert np.array_equal(inl[: n.value], icv.reshape(-1))
    assert np.array_equal(r, rcv.reshape(3)) and np.array_equal(t, tcv.reshape(3))


def test_triangulation_wrapper_cpp_unit(host, ctx):
    """host/triangulateB200.cpp::triangulationWrapper (triangulate.cpp:57-72): N x 2 CV_64F in, 4 x N out,
    equal to the Python twin's device result and within 1e-12 of the oracle."""
    from slam_indoor_code_b200 import triangulation as tri
    p1, p2, R, t = synth.two_view(3000, 7300, outliers=0.0)
    P1 = np.ascontiguousarray(KMAT @ np.hstack([np.eye(3), np.zeros((3, 1))]))
    P2 = np.ascontiguousarray(KMAT @ np.hstack([R, t.reshape(3, 1)]))
    a, b = np.ascontiguousarray(p1, np.float64), np.ascontiguousarray(p2, np.float64)
    out = np.zeros((4, 3000))
    host.hostshim_triangulate.restype = ctypes.c_int
    host.hostshim_triangulate.argtypes = [ctypes.c_void_p] * 2 + [ctypes.c_int] + [ctypes.c_void_p] * 3
    assert host.hostshim_triangulate(_capi.ptr(a), _capi.ptr(b), 3000, _capi.ptr(P1), _capi.ptr(P2), _capi.ptr(out)) == 0
    assert np.array_equal(out, tri.triangulationWrapper(ctx, p1, p2, P1, P2))
    O4, _ = c_oracle.triangulate(P1, P2, p1, p2)
    sgn = np.sign(np.sum(out * O4, axis=0))
    assert np.max(np.abs(out - O4 * sgn)) < 1e-12


def test_descriptor_cache_of_the_drop_in_unit(host):
    """Caller-owned descriptor Mats are uploaded ONCE (the reference re-uploads per pair,
    featureMatchingCUDA.cpp:98-99): the previous frame's descriptor across the pairs of a search,
    a batch element across searches; a Mat rewritten in place is recognised and uploaded again."""
    host.hostshim_mat_create.restype = ctypes.c_void_p
    host.hostshim_mat_create.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int]
    host.hostshim_mat_free.argtypes = [ctypes.c_void_p]
    host.hostshim_mat_write.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
    host.hostshim_match_mats.restype = ctypes.c_int
    host.hostshim_match_mats.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int]
    host.hostshim_cache_stats.argtypes = [ctypes.c_void_p, ctypes.c_void_p]

    def stats():
        h, m = ctypes.c_longlong(0), ctypes.c_longlong(0)
        host.hostshim_cache_stats(ctypes.byref(h), ctypes.byref(m))
        return h.value, m.value

    host.hostshim_set_ratio(0.7)
    host.hostshim_cache_clear()
    q = synth.sift_like(900, 71)
    trains = [synth.sift_train_from_query(q, 1000 + 10 * i, 72 + i) for i in range(6)]
    CV_32F = 5
    Q = host.hostshim_mat_create(_capi.ptr(q), len(q), 128, CV_32F)
    Ts = [host.hostshim_mat_create(_capi.ptr(t), len(t), 128, CV_32F) for t in trains]
    out = np.zeros(len(q), _capi.DMATCH)
    h0, m0 = stats()
    for rounds in range(2):           # two "searches" over the same window
        for T, t in zip(Ts, trains):
            n = host.hostshim_match_mats(Q, T, 0, _capi.ptr(out), len(out))
            assert np.array_equal(out[:n], c_oracle.match_features(0, q, t, 0.7))
    h1, m1 = stats()
    assert m1 - m0 == 1 + len(trains)                       # every Mat uploaded once
    assert h1 - h0 == 2 * 2 * len(trains) - (1 + len(trains))
    # in-place rewrite of a train Mat (same pointer, same shape): must not be answered from the cache
    t_new = synth.sift_train_from_query(q, len(trains[2]), 99)
    host.hostshim_mat_write(Ts[2], _capi.ptr(t_new))
    n = host.hostshim_match_mats(Q, Ts[2], 0, _capi.ptr(out), len(out))
    assert np.array_equal(out[:n], c_oracle.match_features(0, q, t_new, 0.7))
    assert stats()[1] - m1 == 1
    # many host threads share the query Mat (batch.cpp:181-201): one upload of it
    import threading
    host.hostshim_cache_clear()
    _, m2 = stats()
    res = [None] * len(Ts)

    def work(i):
        o = np.zeros(len(q), _capi.DMATCH)
        k = host.hostshim_match_mats(Q, Ts[i], 0, _capi.ptr(o), len(o))
        res[i] = o[:k]
    th = [threading.Thread(target=work, args=(i,)) for i in range(len(Ts))]
    [x.start() for x in th]
    [x.join() for x in th]
    want = trains[:2] + [t_new] + trains[3:]
    for r, t in zip(res, want):
        assert np.array_equal(r, c_oracle.match_features(0, q, t, 0.7))
    assert stats()[1] - m2 == 1 + len(Ts)
    for m in Ts + [Q]:
        host.hostshim_mat_free(m)
    host.hostshim_cache_clear()


def test_match_frames_pair_features_entry_point_and_timing_lines(host):
    """matchFramesPairFeatures(firstFrameDescriptor, secondFrame, secondFeatures, ORB_BF, matches)
    (featureMatching.h:47-53) end to end in the C++ unit -- describe the second frame on the device,
    match, prune `secondFeatures` -- and the per-call timing lines of the reference's CUDA unit
    (featureMatchingCUDA.cpp:101,107) on logStreams.timeStream."""
    host.hostshim_match_frames_pair_orb.restype = ctypes.c_int
    host.hostshim_match_frames_pair_orb.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int,
                                                    ctypes.c_int, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_int,
                                                    ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]
    host.hostshim_time_log.restype = ctypes.c_int
    host.hostshim_time_log.argtypes = [ctypes.c_char_p, ctypes.c_int]
    host.hostshim_mat_create.restype = ctypes.c_void_p
    host.hostshim_mat_create.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int]
    host.hostshim_mat_free.argtypes = [ctypes.c_void_p]
    host.hostshim_set_ratio(0.8)
    frame = synth.textured_frame(480, 640, 81, 3)
    kps = c_oracle.fast_detect(frame, 20, True)[:, :2]
    kps = np.concatenate([kps, np.full((len(kps), 1), -1.0, np.float32)], 1).astype(np.float32)
    keep, desc2 = c_oracle.orb_compute(frame, kps)
    # the "previous frame": the same frame shifted by a few pixels, described by the oracle
    prev = np.roll(frame, (2, 3), (0, 1))
    _, desc1 = c_oracle.orb_compute(prev, kps)
    Q = host.hostshim_mat_create(_capi.ptr(desc1), len(desc1), 32, 0)
    buf = ctypes.create_string_buffer(4096)
    host.hostshim_time_log(buf, 4096)
    out = np.zeros(len(desc1), _capi.DMATCH)
    left = ctypes.c_int(0)
    n = host.hostshim_match_frames_pair_orb(Q, _capi.ptr(frame), 480, 640, 3, frame.strides[0], _capi.ptr(kps),
                                            len(kps), _capi.ptr(out), len(out), ctypes.byref(left))
    assert n >= 0 and left.value == len(desc2)
    assert np.array_equal(out[:n], c_oracle.match_features(2, desc1, desc2, 0.8))
    host.hostshim_time_log(buf, 4096)
    lines = buf.value.decode().splitlines()
    assert len(lines) == 2 and lines[0].startswith("Descriptors extracting: ") and lines[1].startswith("Matching: ")
    assert all(l.split(": ")[1].isdigit() for l in lines)
    host.hostshim_mat_free(Q)
