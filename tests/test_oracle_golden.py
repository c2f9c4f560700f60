"""CPU: the oracle (C and NumPy restatements) against the committed cv2-generated fixtures.

tests/golden/*.npz hold the outputs of the reference's own arithmetic owner (OpenCV, through the
cv2 wheel; oracle/gen_golden.py) for the calls the reference makes at
featureMatchingCPU.cpp:26-40 and cameraTranslation.cpp:41-46.  Bit-exact everywhere: indices,
float distances (bit patterns), masks.
"""
import os

import numpy as np
import pytest

from oracle import c_oracle, np_oracle

L2_CASES = ["sift_int", "sift_float", "sift_ties", "sift_t1", "sift_t2", "sift_q1"]


def _load(golden_dir, name):
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    return g, g["q"].astype(np.float32) if g["q"].dtype != np.float32 else g["q"], \
        g["t"].astype(np.float32) if g["t"].dtype != np.float32 else g["t"]


def _same_knn(idx, dist, g):
    assert np.array_equal(idx, g["idx"])
    valid = g["idx"] >= 0
    assert np.array_equal(dist[valid].view(np.int32), g["dist"][valid].view(np.int32))


@pytest.mark.parametrize("name", L2_CASES)
def test_c_oracle_l2_matches_cv2(golden_dir, name):
    g, q, t = _load(golden_dir, name)
    idx, dist = c_oracle.l2_knn2(q, t)
    _same_knn(idx, dist, g)


@pytest.mark.parametrize("name", L2_CASES)
def test_np_oracle_l2_matches_cv2(golden_dir, name):
    g, q, t = _load(golden_dir, name)
    idx, dist = np_oracle.knn2_from_matrix(np_oracle.l2_dist_matrix(q, t))
    _same_knn(idx, dist, g)


@pytest.mark.parametrize("name", ["orb", "orb_t1"])
def test_oracles_hamming_match_cv2(golden_dir, name):
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    idx, dist = c_oracle.hamming_knn2(g["q"], g["t"])
    _same_knn(idx, dist, g)
    idx, dist = np_oracle.knn2_from_matrix(np_oracle.hamming_dist_matrix(g["q"], g["t"]))
    _same_knn(idx, dist, g)


def test_ties_keep_lowest_train_index(golden_dir):
    g, q, t = _load(golden_dir, "sift_ties")
    idx, dist = c_oracle.l2_knn2(q, t)
    # query 5 has two exact copies in the train set at rows 7 and 20: both distance 0, 7 first
    assert list(idx[5]) == [7, 20] and dist[5, 0] == 0 and dist[5, 1] == 0
    # query 9 has one exact copy at row 60
    assert idx[9, 0] == 60 and dist[9, 0] == 0
    # identical queries get identical rows
    assert np.array_equal(idx[11], idx[12])


def test_short_train_sets(golden_dir):
    g, q, t = _load(golden_dir, "sift_t1")
    assert list(g["lens"]) == [1] * q.shape[0]          # cv2: one-element lists when T == 1
    idx, _ = c_oracle.l2_knn2(q, t)
    assert np.all(idx[:, 0] == 0) and np.all(idx[:, 1] == -1)
    # getGoodMatches would read [1] out of bounds there (reference UB): defined as "reject"
    assert len(c_oracle.match_features(0, q, t, 0.7)) == 0
    # empty sets
    e = np.zeros((0, 128), np.float32)
    assert c_oracle.l2_knn2(e, t)[0].shape == (0, 2)
    idx, _ = c_oracle.l2_knn2(q, e)
    assert np.all(idx == -1)
    assert len(c_oracle.match_features(0, q, e, 0.7)) == 0


def test_ratio_test_is_strict_and_in_double():
    idx = np.array([[3, 4], [5, 6], [7, 8], [1, -1], [-1, -1]], np.int32)
    d1 = np.float32(10.0)
    on = np.float32(np.float64(0.7) * np.float64(d1))        # (float) of the double product
    dist = np.array([[on, d1], [np.nextafter(on, np.float32(0)), d1], [0, 0], [1, 0], [0, 0]],
                    np.float32)
    got = c_oracle.ratio_test(idx, dist, 0.7)
    ref = np_oracle.ratio_test(idx, dist, 0.7)
    assert np.array_equal(got, ref)
    exp = [q for q in range(3) if float(dist[q, 0]) < 0.7 * float(dist[q, 1])]
    assert list(got["queryIdx"]) == exp
    assert 2 not in got["queryIdx"]                           # 0 < 0.7*0 is false
    assert np.all(got["imgIdx"] == 0)


def test_c_and_np_oracles_agree_on_seeded_inputs():
    from oracle import synth
    q, t = synth.sift_pair(150, 170, 77)
    a = c_oracle.match_features(0, q, t, 0.7)
    idx, dist = np_oracle.knn2_from_matrix(np_oracle.l2_dist_matrix(q, t))
    b = np_oracle.ratio_test(idx, dist, 0.7)
    assert np.array_equal(a, b) and len(a) > 10
    q, t = synth.orb_pair(120, 140, 78)
    a = c_oracle.match_features(2, q, t, 0.7)
    idx, dist = np_oracle.knn2_from_matrix(np_oracle.hamming_dist_matrix(q, t))
    assert np.array_equal(a, np_oracle.ratio_test(idx, dist, 0.7)) and len(a) > 10
    with pytest.raises(ValueError):
        c_oracle.match_features(4, q, t, 0.7)


def test_ransac_masks_match_cv2(golden_dir):
    g = np.load(os.path.join(golden_dir, "ransac.npz"))
    for c in range(4):
        p1, p2 = g[f"p1_{c}"], g[f"p2_{c}"]
        for orc in (c_oracle, np_oracle):
            counts, best, mask, _ = orc.score_essential(p1, p2, g["K4"], g[f"E_cv_{c}"],
                                                        float(g["threshold_px"]))[:4]
            assert best == 0
            assert np.array_equal(mask, g[f"mask_cv_{c}"])
            assert counts[0] == int(g[f"mask_cv_{c}"].sum())


def test_ransac_c_vs_np_on_hypothesis_lists(golden_dir):
    g = np.load(os.path.join(golden_dir, "ransac.npz"))
    for c in range(4):
        p1, p2, hyp = g[f"p1_{c}"], g[f"p2_{c}"], g[f"hyp_{c}"]
        ca, ba, ma, alla = c_oracle.score_essential(p1, p2, g["K4"], hyp, 5.0, want_all_masks=True)
        cb, bb, mb, allb = np_oracle.score_essential(p1, p2, g["K4"], hyp, 5.0)
        assert np.array_equal(ca, cb) and ba == bb and np.array_equal(ma, mb)
        assert np.array_equal(alla, allb)
        # first-best-wins, and only above 4 inliers
        if ba >= 0:
            assert ca[ba] == ca.max() and ca[ba] > 4 and np.argmax(ca) == ba


def test_ransac_first_best_and_min_count():
    from oracle import synth
    p1, p2, R, t = synth.two_view(200, 42)
    E = synth.pose_hypotheses(8, R, t, 43)
    E2 = np.concatenate([E, E])                               # duplicates: the first must win
    c, b, m, _ = c_oracle.score_essential(p1, p2, synth.SAMSUNG_HV_4K, E2, 5.0)
    assert b == int(np.argmax(c)) and b < 8
    # three matches can never exceed the "> 4" floor
    c, b, m, _ = c_oracle.score_essential(p1[:3], p2[:3], synth.SAMSUNG_HV_4K, E, 5.0)
    assert b == -1 and not m.any()


# ---- solvePnPRansac scoring ------------------------------------------------------
def test_pnp_masks_match_cv2_project_points(golden_dir):
    """The reprojection-inlier restatement against masks built from cv2.projectPoints, for the
    reference's five-coefficient model, no distortion, and the twelve-coefficient model."""
    g = np.load(os.path.join(golden_dir, "pnp.npz"))
    for c in range(5):
        obj, img, dist, poses = g[f"obj_{c}"], g[f"img_{c}"], g[f"dist_{c}"], g[f"poses_{c}"]
        want = np.unpackbits(g[f"masks_{c}"], axis=1)[:, :len(obj)]
        counts, best, mask, allm = c_oracle.score_pnp(obj, img, g["K4"], dist, poses, float(g["reproj"]),
                                                      want_all_masks=True)
        assert np.array_equal(allm, want)
        assert np.array_equal(counts, want.sum(1))
        if best >= 0:
            assert best == int(np.argmax(counts)) and counts[best] > 4
            assert np.array_equal(mask, want[best])


def test_pnp_cv2_ransac_inliers_are_the_mask_of_some_minimal_model(golden_dir):
    """cv2.solvePnPRansac's inlier list has the size the update rule promises: no scored
    hypothesis of the fixture beats it by the strict rule unless it was never drawn (sanity)."""
    g = np.load(os.path.join(golden_dir, "pnp.npz"))
    for c in range(5):
        assert bool(g[f"cv_ok_{c}"])
        inl = g[f"cv_inliers_{c}"]
        assert len(inl) > 4 and np.all(np.diff(inl) > 0)


# ---- NORM_L1 mode ---------------------------------------------------------------
@pytest.mark.parametrize("name", ["sift_int", "sift_float", "sift_ties", "sift_t1", "sift_t2", "sift_q1"])
def test_c_oracle_l1_matches_cv2(golden_dir, name):
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    l1 = np.load(os.path.join(golden_dir, "sift_l1.npz"))
    idx, dist = c_oracle.l1_knn2(g["q"].astype(np.float32), g["t"].astype(np.float32))
    assert np.array_equal(idx, l1[name + "_idx"])
    valid = idx >= 0
    assert np.array_equal(dist[valid].view(np.int32), l1[name + "_dist"][valid].view(np.int32))


# ---- ORB descriptors of given keypoints -----------------------------------------
def test_orb_descriptor_oracle_matches_cv2(golden_dir):
    g = np.load(os.path.join(golden_dir, "orb_desc.npz"))
    keep, desc = c_oracle.orb_compute(g["frame"], g["kps"])
    assert np.array_equal(g["kps"][keep.astype(bool), :2], g["kept_xy"])
    assert np.array_equal(desc, g["desc"]) and len(desc) > 150


# ---- FAST keypoints (fastExtractor.cpp:7-13) ---------------------------------------
def test_fast_oracle_matches_cv2_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "fast.npz"))
    for thr, nms in ((10, True), (10, False), (25, True)):
        ref = g[f"kp_t{thr}_n{int(nms)}"]
        got = c_oracle.fast_detect(g["frame"], thr, nms)
        assert np.array_equal(got, ref) and (len(ref) > 20 or thr > 10)


def test_fast_np_statement_matches_cv2_golden_and_c_oracle(golden_dir):
    """FAST restated twice: whole-image array operations (np_oracle) against the cv2 fixture and
    against the C loops on fresh frames (gray and BGR, with and without suppression)."""
    from oracle import synth
    g = np.load(os.path.join(golden_dir, "fast.npz"))
    for thr, nms in ((10, True), (10, False), (25, True)):
        assert np.array_equal(np_oracle.fast_detect(g["frame"], thr, nms), g[f"kp_t{thr}_n{int(nms)}"])
    for h, w, ch, seed, thr, nms in [(200, 300, 3, 611, 10, True), (133, 157, 1, 612, 0, True),
                                     (128, 128, 3, 613, 40, False), (7, 9, 1, 614, 10, True)]:
        f = synth.textured_frame(h, w, seed, ch)
        assert np.array_equal(np_oracle.fast_detect(f, thr, nms), c_oracle.fast_detect(f, thr, nms))


# ---- the next rows restated twice: the C oracle against the independent NumPy statement -----------
def test_c_and_np_oracles_agree_on_next_rows(golden_dir):
    import re
    from oracle import synth
    q, t = synth.float_pair(90, 130, 601)
    idx, dist = np_oracle.knn2_from_matrix(np_oracle.l1_dist_matrix(q, t))
    ci, cd = c_oracle.l1_knn2(q, t)
    assert np.array_equal(idx, ci) and np.array_equal(dist.view(np.int32), cd.view(np.int32))
    obj, img, R, tt = synth.pnp_scene(700, 602)
    poses = synth.pnp_hypotheses(24, R, tt, 603)
    for dist_c in (synth.REF_DIST5, None, (0.1, -0.2, 1e-3, -1e-3, 0.05, 0.01, -0.02, 0.003, 1e-3, 0, 2e-4, 0)):
        a = c_oracle.score_pnp(obj, img, synth.SAMSUNG_HV_4K, dist_c, poses, 8.0, want_all_masks=True)
        b = np_oracle.score_pnp(obj, img, synth.SAMSUNG_HV_4K, dist_c, poses, 8.0)
        assert np.array_equal(a[0], b[0]) and a[1] == b[1] and np.array_equal(a[2], b[2])
        assert np.array_equal(a[3], b[3])
    here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pat = np.array([[int(v) for v in re.findall(r"-?\d+", l)]
                    for l in open(os.path.join(here, "slam_indoor_code_b200", "csrc", "orb_pattern.h"))
                    if l.strip().startswith("{")])
    g = np.load(os.path.join(golden_dir, "orb_desc.npz"))
    keep, desc = np_oracle.orb_compute(g["frame"], g["kps"], pat)
    ck, cdsc = c_oracle.orb_compute(g["frame"], g["kps"])
    assert np.array_equal(keep, ck) and np.array_equal(desc, cdsc) and np.array_equal(desc, g["desc"])


def test_sift_descriptors_golden(golden_dir):
    """extractDescriptor's SIFT branch (featureMatchingCPU.cpp:45-66 -> cv::SIFT::compute on octave-0
    keypoints): the restated working image equals cv2.GaussianBlur's bit for bit (frame width not a
    multiple of 8: the filter's scalar tails included); the descriptors are held to the stated
    tolerance against cv2.SIFT.compute -- every element within 1, >= 99.9 % equal."""
    g = np.load(os.path.join(golden_dir, "sift_desc.npz"))
    base = c_oracle.sift_base(g["frame"])
    assert np.array_equal(base[[0, 1, 74, 148, 149]], g["base_rows"])
    assert np.array_equal(base[:, [0, 1, 199, 200, 203, 204]], g["base_cols"])
    got = c_oracle.sift_compute(g["frame"], g["kps"])
    want = g["desc"].astype(np.float32)
    assert got.shape == want.shape == (len(g["kps"]), 128)
    assert np.abs(got - want).max() <= 1.0
    assert np.mean(got == want) >= 0.999
    assert np.all(got == np.rint(got)) and got.min() >= 0 and got.max() <= 255
