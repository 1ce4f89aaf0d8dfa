"""The cv2-based CPU baseline port (oracle/cv_ref.py) must behave like the C oracle, and like the real
reference classes when the checkout is present (build container only)."""
import numpy as np
import pytest

cv2 = pytest.importorskip("cv2")
from oracle import cv_ref, oracle as orc, ref_shim  # noqa: E402


def _seq(rng, n=3):
    for k in range(n):
        img = rng.integers(0, 256, (70, 90, 3), dtype=np.uint8)
        th = rng.uniform(-0.6, 0.6)
        yield img, np.array([[np.cos(th), -np.sin(th), rng.uniform(0, 80)], [np.sin(th), np.cos(th), rng.uniform(0, 60)],
                             [rng.uniform(-2e-4, 2e-4), 0.0, 1.0]])


@pytest.mark.parametrize("flags,interp", [(cv2.INTER_LINEAR, 1), (cv2.INTER_CUBIC, 2)])
def test_ref_mapper_equals_oracle(flags, interp):
    rng = np.random.default_rng(1)
    mp = cv_ref.RefMapper(200, 150, alpha_blend=0.3, flags=flags)
    canvas = np.zeros((150, 200, 3), np.uint8); cmask = np.zeros((150, 200), np.uint8)
    for img, H in _seq(rng):
        mp.update(img, H)
        orc.mapper_update(canvas, cmask, img, H, 0.3, interp)
        assert np.array_equal(mp.output, canvas) and np.array_equal(mp._canvas_mask, cmask)


def test_ref_matcher_equals_oracle():
    rng = np.random.default_rng(2)
    q = rng.integers(0, 256, (90, 32), dtype=np.uint8); t = rng.integers(0, 256, (120, 32), dtype=np.uint8)
    for norm, name in ((cv2.NORM_HAMMING, "hamming"), (cv2.NORM_L2, "l2")):
        m = cv_ref.RefMatcher(norm).knn_match(q, t)
        idx, dist = orc.knn2(q, t, name)
        assert [[x.trainIdx for x in r] for r in m] == idx.tolist()
        keep = orc.ratio_test(idx, dist, 0.7)
        assert [g.queryIdx for g in cv_ref.lowe_ratio(m, 0.7)] == np.flatnonzero(keep).tolist()


@pytest.mark.skipif(not ref_shim.available(), reason="reference checkout only exists in the build container")
def test_port_equals_real_reference():
    reg, mos = ref_shim.import_reference()
    rng = np.random.default_rng(3)
    real, port = mos.Mapper(200, 150, alpha_blend=0.5), cv_ref.RefMapper(200, 150, alpha_blend=0.5)
    for img, H in _seq(rng):
        kps = [cv2.KeyPoint(float(x), float(y), 1) for x, y in rng.uniform(5, 60, (20, 2))]
        real.update(img, H, None, kps); port.update(img, H, None, kps)
        assert np.array_equal(real.output, port.output) and np.array_equal(real._canvas_mask, port._canvas_mask)
    q = rng.integers(0, 256, (50, 32), dtype=np.uint8); t = rng.integers(0, 256, (60, 32), dtype=np.uint8)
    a, b = reg.Matcher().knn_match(q, t), cv_ref.RefMatcher().knn_match(q, t)
    assert [[(x.trainIdx, x.distance) for x in r] for r in a] == [[(x.trainIdx, x.distance) for x in r] for r in b]
