"""CPU: the host-side RANSAC control (RNG, subsets, iteration budget, update rule) reproduces
cv2.findEssentialMat bit for bit when the scorer is the CPU oracle.  The same control drives the
GPU scorer in tests/test_gpu_ransac.py."""
import numpy as np
import pytest

cv2 = pytest.importorskip("cv2")

from oracle import c_oracle, synth
from slam_indoor_code_b200 import ransac_host
from slam_indoor_code_b200.camera_translation import _cv2_five_point

K4 = np.array(synth.SAMSUNG_HV_4K)
KMAT = np.array([[K4[0], 0, K4[2]], [0, K4[1], K4[3]], [0, 0, 1.0]])


def test_cv_rng_sequence():
    rng = ransac_host.CvRNG()
    a = [rng.next() for _ in range(4)]
    assert a[0] == (0xFFFFFFFF * 4164903690 + 0xFFFFFFFF) & 0xFFFFFFFF
    r2 = ransac_host.CvRNG()
    s = ransac_host.get_subset(r2, 1000)
    assert len(set(s)) == 5 and all(0 <= v < 1000 for v in s)


def test_update_num_iters():
    assert ransac_host.update_num_iters(0.999, 0.3, 5, 1000) == 38
    assert ransac_host.update_num_iters(0.999, 0.0, 5, 1000) == 0
    assert ransac_host.update_num_iters(0.999, 1.0, 5, 1000) == 1000
    assert ransac_host.update_num_iters(0.999, 0.9, 5, 1000) == 1000


@pytest.mark.parametrize("m,noise,outl,seed", [(600, 0.7, 0.3, 7000), (1500, 1.5, 0.5, 7001),
                                               (200, 0.7, 0.2, 7002), (80, 2.0, 0.6, 7004)])
@pytest.mark.parametrize("chunk", [1, 32])
def test_control_loop_equals_cv2(m, noise, outl, seed, chunk):
    p1, p2, _, _ = synth.two_view(m, seed, noise_px=noise, outliers=outl)
    Ecv, mcv = cv2.findEssentialMat(p1, p2, KMAT, cv2.RANSAC, 0.999, 5.0)
    score = lambda models: c_oracle.score_essential(p1, p2, K4, models, 5.0)[0]
    E, _, iters, scored = ransac_host.ransac_essential(p1, p2, K4, 0.999, 5.0, _cv2_five_point(K4), score,
                                                       chunk=chunk)
    assert np.array_equal(np.asarray(Ecv, np.float64).reshape(-1)[:9], E)
    mask = c_oracle.score_essential(p1, p2, K4, E.reshape(1, 9), 5.0)[2]
    assert np.array_equal(mask, mcv.reshape(-1))
    assert 0 < iters <= 1000 and scored > 0
