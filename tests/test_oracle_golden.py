"""Pins the CPU oracle (oracle/mosaic_oracle.c) against the committed golden fixtures that
oracle/make_golden.py produced from the real reference classes + cv2, and,
when cv2 is importable, against live cv2 calls on fresh seeded inputs."""
import numpy as np
import pytest

from oracle import oracle as orc

KNN_TAGS_HAM = ["orb", "ties", "nt1", "nt2", "b64", "realorb"]
KNN_TAGS_L2 = KNN_TAGS_HAM + ["siftint", "f32", "f32d61", "realsift"]


@pytest.mark.parametrize("tag", KNN_TAGS_HAM)
def test_knn_hamming_golden(golden, tag):
    g = golden.knn
    idx, dist = orc.knn2(g[f"{tag}_q"], g[f"{tag}_t"], "hamming")
    assert np.array_equal(idx, g[f"{tag}_ham_idx"])
    assert np.array_equal(dist, g[f"{tag}_ham_dist"])
    assert np.array_equal(orc.ratio_test(idx, dist, 0.7), g[f"{tag}_ham_keep"])


@pytest.mark.parametrize("tag", KNN_TAGS_L2)
def test_knn_l2_golden(golden, tag):
    g = golden.knn
    idx, dist = orc.knn2(g[f"{tag}_q"], g[f"{tag}_t"], "l2")
    assert np.array_equal(idx, g[f"{tag}_l2_idx"])
    assert np.array_equal(dist, g[f"{tag}_l2_dist"])          # bit-exact, incl. generic fp32
    assert np.array_equal(orc.ratio_test(idx, dist, 0.7), g[f"{tag}_l2_keep"])


def test_warp_golden(golden):
    g = golden.warp
    for n in range(int(g["n"])):
        for tag, interp in (("lin", orc.INTER_LINEAR), ("cub", orc.INTER_CUBIC)):
            got = orc.warp_perspective(g[f"img{n}"], g[f"H{n}"], tuple(g[f"dsize{n}"]), interp)
            assert np.array_equal(got, g[f"{tag}{n}"]), (n, tag)


def test_mapper_golden(golden):
    g = golden.mapper
    for ci, (interp, alpha, use_kp, nupd, Hc, Wc) in enumerate(g["cases"]):
        canvas = np.zeros((int(Hc), int(Wc), 3), np.uint8)
        cmask = np.zeros((int(Hc), int(Wc)), np.uint8)
        for k in range(int(nupd)):
            roi = g[f"c{ci}_roi{k}"] if use_kp else None
            orc.mapper_update(canvas, cmask, g[f"c{ci}_img{k}"], g[f"c{ci}_H{k}"], float(alpha), int(interp), roi)
            assert np.array_equal(canvas, g[f"c{ci}_canvas{k}"]), (ci, k)
            assert np.array_equal(cmask, g[f"c{ci}_mask{k}"]), (ci, k)


def test_ratio_integer_form_equals_float64_form():
    # mosaic.py:430 on Hamming ints:  d1 < 0.7*d2 (float64)  <=>  10*d1 < 7*d2  for 0 <= d <= 512
    d1, d2 = np.meshgrid(np.arange(513), np.arange(513), indexing="ij")
    assert np.array_equal(d1.astype(np.float64) < 0.7 * d2.astype(np.float64), 10 * d1 < 7 * d2)


def test_oracle_vs_live_cv2():
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(5)
    img = rng.integers(0, 256, (90, 130, 3), dtype=np.uint8)
    for k in range(6):
        th = rng.uniform(-1, 1)
        H = np.array([[np.cos(th), -np.sin(th), rng.uniform(0, 80)], [np.sin(th), np.cos(th), rng.uniform(0, 60)],
                      [rng.uniform(-1e-3, 1e-3), rng.uniform(-1e-3, 1e-3), 1.0]])
        for fl, interp in ((cv2.INTER_LINEAR, 1), (cv2.INTER_CUBIC, 2)):
            ref = cv2.warpPerspective(img, H, (200, 160), None, fl, borderMode=cv2.BORDER_CONSTANT, borderValue=0)
            assert np.array_equal(ref, orc.warp_perspective(img, H, (200, 160), interp))
    q = rng.integers(0, 256, (100, 32), dtype=np.uint8)
    t = rng.integers(0, 256, (140, 32), dtype=np.uint8)
    for norm, name in ((cv2.NORM_HAMMING, "hamming"), (cv2.NORM_L2, "l2")):
        m = cv2.BFMatcher(norm).knnMatch(q, t, k=2)
        idx, dist = orc.knn2(q, t, name)
        assert [[mm.trainIdx for mm in r] for r in m] == idx.tolist()
        assert np.array_equal(np.array([[mm.distance for mm in r] for r in m], np.float32), dist)


def test_reference_pipeline_golden(golden):
    """The reference's own integration test run for real (SequentialMosaic on its test video, ORB + keypoint ROI, alpha = 1,
    INTER_CUBIC as shipped): the recorded Mapper.update calls replayed through the oracle give the reference's canvas."""
    cv2 = pytest.importorskip("cv2")
    g = golden.pipeline
    Wt, Ht = (int(v) for v in g["tile"])
    y0, y1, x0, x1 = (int(v) for v in g["crop"])
    canvas = np.zeros((Ht, Wt, 3), np.uint8)
    cmask = np.zeros((Ht, Wt), np.uint8)
    for k in range(int(g["n"])):
        img, kp = g[f"img{k}"], g[f"kp{k}"]
        roi = None
        if len(kp):                                            # mosaic.py:140-157
            roi = np.zeros(img.shape[:2], np.uint8)
            roi = cv2.fillConvexPoly(roi, cv2.convexHull(kp.astype(np.int32)), 255)
        orc.mapper_update(canvas, cmask, img, g[f"H{k}"], float(g["alpha"]), orc.INTER_CUBIC, roi)
    assert np.array_equal(canvas[y0:y1, x0:x1], g["canvas_crop"])
    assert np.array_equal(cmask[y0:y1, x0:x1], g["mask_crop"])
    assert int(canvas.sum()) == int(g["canvas_crop"].sum())    # nothing outside the recorded crop


def test_match_features_golden(golden):
    """The good matches SequentialMosaic.match_features kept on the reference's test video (mosaic.py:426-430, metric NORM_L2
    on the ORB bytes as registration.py:231 creates it): oracle kNN + ratio test reproduce queryIdx / trainIdx / distance."""
    g = golden.matchfeatures
    for k in range(int(g["npairs"])):
        idx, dist = orc.knn2(g[f"q{k}"], g[f"t{k}"], "l2")
        keep = orc.ratio_test(idx, dist, 0.7)
        got = np.stack([np.nonzero(keep)[0].astype(np.float64), idx[keep, 0].astype(np.float64), dist[keep, 0].astype(np.float64)], 1)
        assert np.array_equal(got, g[f"good{k}"]), k


def test_window_oracle_equals_full_canvas_crop():
    rng = np.random.default_rng(8)
    Hc, Wc = 420, 560
    canvas = np.zeros((Hc, Wc, 3), np.uint8); cm = np.zeros((Hc, Wc), np.uint8)
    for (wx, wy, ww, wh) in ((0, 0, 300, 200), (250, 180, 310, 240), (100, 60, 200, 150)):
        win = canvas[wy:wy + wh, wx:wx + ww].copy(); wm = cm[wy:wy + wh, wx:wx + ww].copy()
        full, fm = canvas.copy(), cm.copy()
        img = rng.integers(0, 256, (150, 220, 3), dtype=np.uint8)
        img[rng.random((150, 220)) < 0.03] = 0
        th = rng.uniform(-0.6, 0.6)
        H = np.array([[np.cos(th), -np.sin(th), wx + 30.0], [np.sin(th), np.cos(th), wy + 40.0], [2e-5, 1e-5, 1.0]])
        for interp in (orc.INTER_LINEAR, orc.INTER_CUBIC):
            f2, m2, w2, wm2 = full.copy(), fm.copy(), win.copy(), wm.copy()
            orc.mapper_update(f2, m2, img, H, 0.4, interp)
            orc.mapper_update_window(w2, wm2, (Hc, Wc), (wx, wy), img, H, 0.4, interp)
            assert np.array_equal(f2[wy:wy + wh, wx:wx + ww], w2) and np.array_equal(m2[wy:wy + wh, wx:wx + ww], wm2)
        orc.mapper_update(canvas, cm, img, H, 0.4, orc.INTER_LINEAR)


def test_preproc_golden(golden):
    """preprocessing.make_gray / make_bgr and DistortionMapper.apply of the real reference (preprocessing.py:68-98,314-346):
    the oracle's cvtColor / remap restatements equal them bit for bit (uint8)."""
    g = golden.preproc
    for tag in ("a", "b"):
        assert np.array_equal(orc.cvt_color(g[f"{tag}_bgr"], orc.CVT_BGR2GRAY), g[f"{tag}_gray_of_bgr"])
        assert np.array_equal(orc.cvt_color(g[f"{tag}_bgra"], orc.CVT_BGRA2GRAY), g[f"{tag}_gray_of_bgra"])
        assert np.array_equal(orc.cvt_color(g[f"{tag}_gray"], orc.CVT_GRAY2BGR), g[f"{tag}_bgr_of_gray"])
        assert np.array_equal(orc.cvt_color(g[f"{tag}_bgra"], orc.CVT_BGRA2BGR), g[f"{tag}_bgr_of_bgra"])
    for t in ("fwd", "inv"):
        xm, ym = g[f"und_{t}_xmap"], g[f"und_{t}_ymap"]
        assert np.array_equal(orc.remap(g["und_img"], xm, ym, orc.INTER_CUBIC), g[f"und_{t}_bgr"])
        assert np.array_equal(orc.remap(g["und_gray"], xm, ym, orc.INTER_CUBIC), g[f"und_{t}_gray"])
        assert np.array_equal(orc.remap(g["und_img"], xm, ym, orc.INTER_LINEAR), g[f"und_{t}_linear"])


def test_cvt_gray_all_triples_live_cv2():
    """The 15-bit gray coefficients against cv2 itself on every BGR triple (2^24 pixels)."""
    cv2 = pytest.importorskip("cv2")
    v = np.arange(256, dtype=np.uint8)
    B, G, R = np.meshgrid(v, v, v, indexing="ij")
    img = np.stack([B.ravel(), G.ravel(), R.ravel()], 1).reshape(4096, 4096, 3)
    assert np.array_equal(orc.cvt_color(img, orc.CVT_BGR2GRAY), cv2.cvtColor(img, cv2.COLOR_BGR2GRAY))
