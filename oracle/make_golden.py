"""Generate tests/golden/*.npz from the REAL reference (mosaic-library) + cv2.

TEST INFRASTRUCTURE, build container only: needs /root/reference (read-only checkout) and
cv2.  Run `python oracle/make_golden.py` from the repo root; the fixtures it writes are
committed so that the oracle and the CUDA path can be pinned where the reference is absent
(the GPU box).  Inputs are stored next to the outputs, so nothing depends on RNG streams.

What produces each expected value:
  knn_*      mosaicking.registration.Matcher / CompositeMatcher.knn_match (registration.py:233-290,
             NORM_L2) and cv2.BFMatcher(NORM_HAMMING).knnMatch(k=2) (the north-star metric);
             ratio flags from the literal expression of mosaic.py:430.
  warp       cv2.warpPerspective(img, H, dsize, None, INTER_LINEAR|INTER_CUBIC, BORDER_CONSTANT, 0)
             -- the call of mosaic.py:114.
  mapper_*   mosaicking.mosaic.Mapper.update (mosaic.py:159-172 -> _update_cpu :111-138), as shipped
             (INTER_CUBIC) and with the flag patched to INTER_LINEAR (the north-star kernel), with and
             without the keypoint-hull ROI (mosaic.py:140-157); canvas + mask after every update.
  pipeline   the reference's own integration test (tests/test_mosaic.py:28-73) run for real on its test video: the
             first Mapper.update calls of SequentialMosaic.generate (mosaic.py:665-691) and the canvas they produce.
  matchfeatures  SequentialMosaic.match_features / .registration (mosaic.py:402-470) of the same run: descriptor dicts as handed to
             CompositeMatcher.knn_match, the good matches kept by the ratio comprehension (mosaic.py:430), RANSAC inputs / (H, inliers).
  preproc    mosaicking.preprocessing.make_gray / make_bgr (preprocessing.py:314-346) and DistortionMapper.apply, forward and
             inverse (preprocessing.py:68-98: cv2.remap INTER_CUBIC with the maps of cv2.initUndistortRectifyMap), on odd-sized
             random images; the float32 maps are stored next to the outputs.
  routing    mosaicking.mosaic.get_corners / get_mosaic_dimensions (mosaic.py:708-768),
             registration.bbox_overlap on int-truncated corners (mosaic.py:614), and
             splitter.calculate_tiles / assign_image_to_tiles (splitter.py:11-58).
"""
from __future__ import annotations

import sys
from pathlib import Path

import cv2
import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from oracle import ref_shim  # noqa: E402

OUT = ROOT / "tests" / "golden"
FLT_MAX = np.finfo(np.float32).max


def matches_to_arrays(matches, nq):
    idx = np.full((nq, 2), -1, np.int32)
    dist = np.full((nq, 2), FLT_MAX, np.float32)
    for i, row in enumerate(matches):
        for k, m in enumerate(row):
            assert m.queryIdx == i and m.imgIdx == 0
            idx[i, k] = m.trainIdx
            dist[i, k] = m.distance
    return idx, dist


def ratio_flags(matches):
    # the literal comprehension condition of mosaic.py:430
    return np.array([len(m) == 2 and m[0].distance < 0.7 * m[1].distance for m in matches], bool)


def smooth_img(rng, h, w, black=0.02):
    base = rng.integers(0, 256, (h // 8 + 2, w // 8 + 2, 3), dtype=np.uint8)
    img = cv2.resize(base, (w, h), interpolation=cv2.INTER_CUBIC)
    img = np.clip(img, 1, 255).astype(np.uint8)
    img[rng.random((h, w)) < black] = 0
    return img


def gen_knn(reg):
    rng = np.random.default_rng(11)
    out = {}
    matcher = reg.Matcher()
    comp = reg.CompositeMatcher()

    def add(tag, q, t, hamming):
        out[f"{tag}_q"], out[f"{tag}_t"] = q, t
        m = matcher.knn_match(q, t)
        out[f"{tag}_l2_idx"], out[f"{tag}_l2_dist"] = matches_to_arrays(m, len(q))
        out[f"{tag}_l2_keep"] = ratio_flags(m)
        mc = comp.knn_match({"x": q, "none": None}, {"x": t, "none": None})
        assert mc["none"] == tuple()
        ci, cd = matches_to_arrays(mc["x"], len(q))
        assert np.array_equal(ci, out[f"{tag}_l2_idx"]) and np.array_equal(cd, out[f"{tag}_l2_dist"])
        if hamming:
            mh = cv2.BFMatcher(cv2.NORM_HAMMING, crossCheck=False).knnMatch(q, t, k=2)
            out[f"{tag}_ham_idx"], out[f"{tag}_ham_dist"] = matches_to_arrays(mh, len(q))
            out[f"{tag}_ham_keep"] = ratio_flags(mh)

    # ORB-like random bytes with planted duplicates / near-duplicates (forces distance ties)
    q = rng.integers(0, 256, (257, 32), dtype=np.uint8)
    t = rng.integers(0, 256, (391, 32), dtype=np.uint8)
    t[50] = q[3]; t[100] = q[3]; t[300] = q[3]
    t[10] = q[7]; t[11] = q[7]; t[11, 0] ^= 1
    for i in range(20, 60):           # genuine matches: a few flipped bits
        t[i + 100] = q[i]
        flips = rng.integers(0, 256, 3)
        for f in flips:
            t[i + 100, f // 8] ^= np.uint8(1 << (f % 8))
    add("orb", q, t, True)
    # low-entropy bytes: many exact ties in both metrics
    q2 = (rng.integers(0, 3, (130, 32)) * 85).astype(np.uint8)
    t2 = (rng.integers(0, 3, (190, 32)) * 85).astype(np.uint8)
    add("ties", q2, t2, True)
    # degenerate train sizes (cv2 returns fewer than k matches)
    add("nt1", q[:9].copy(), t[:1].copy(), True)
    add("nt2", q[:9].copy(), t[:2].copy(), True)
    # 64-byte binary descriptors (e.g. BRISK / AKAZE-MLDB width)
    add("b64", rng.integers(0, 256, (70, 64), dtype=np.uint8), rng.integers(0, 256, (90, 64), dtype=np.uint8), True)
    # SIFT-like: integer-valued float32, D = 128
    qs = rng.integers(0, 180, (150, 128)).astype(np.float32)
    ts = rng.integers(0, 180, (210, 128)).astype(np.float32)
    ts[5] = qs[0]; ts[9] = qs[0]
    for i in range(10, 50):
        ts[i + 60] = np.clip(qs[i] + rng.integers(-6, 7, 128), 0, 255)
    add("siftint", qs, ts, False)
    # generic float32 descriptors (non-integer)
    qg = rng.standard_normal((120, 128)).astype(np.float32)
    tg = rng.standard_normal((170, 128)).astype(np.float32)
    qg /= np.linalg.norm(qg, axis=1, keepdims=True); tg /= np.linalg.norm(tg, axis=1, keepdims=True)
    for i in range(10, 40):
        v = qg[i] + 0.05 * rng.standard_normal(128).astype(np.float32)
        tg[i + 50] = v / np.linalg.norm(v)
    add("f32", qg.astype(np.float32), tg.astype(np.float32), False)
    add("f32d61", rng.standard_normal((40, 61)).astype(np.float32), rng.standard_normal((55, 61)).astype(np.float32), False)

    # real descriptors: ORB on two frames of the reference's own test video, SIFT on its scene images
    vid = ref_shim.REFERENCE_SRC.parent / "tests" / "data" / "sparus_snippet.mp4"
    cap = cv2.VideoCapture(str(vid))
    frames = [cap.read()[1] for _ in range(3)]
    cap.release()
    orb = reg.OrbDetector() if hasattr(reg, "OrbDetector") else None
    det = cv2.ORB.create(nfeatures=300)
    d0 = det.detectAndCompute(cv2.cvtColor(frames[0], cv2.COLOR_BGR2GRAY), None)[1]
    d1 = det.detectAndCompute(cv2.cvtColor(frames[2], cv2.COLOR_BGR2GRAY), None)[1]
    add("realorb", d1, d0, True)      # query = current frame, train = previous (mosaic.py:426)
    s1 = cv2.imread(str(ref_shim.REFERENCE_SRC.parent / "data" / "registration" / "scene1.png"), cv2.IMREAD_GRAYSCALE)
    s2 = cv2.imread(str(ref_shim.REFERENCE_SRC.parent / "data" / "registration" / "scene2.png"), cv2.IMREAD_GRAYSCALE)
    sift = cv2.SIFT.create(nfeatures=250)
    ds1 = sift.detectAndCompute(cv2.resize(s1, (640, 360)), None)[1]
    ds2 = sift.detectAndCompute(cv2.resize(s2, (640, 360)), None)[1]
    add("realsift", ds2, ds1, False)
    np.savez_compressed(OUT / "knn.npz", **out)
    print("knn.npz", {k: v.shape for k, v in out.items() if k.endswith("_q")})


def rand_H(rng, persp, tx, ty):
    th = rng.uniform(-0.6, 0.6)
    s = rng.uniform(0.7, 1.4)
    H = np.array([[s * np.cos(th), -s * np.sin(th), rng.uniform(-tx, tx)],
                  [s * np.sin(th), s * np.cos(th), rng.uniform(-ty, ty)],
                  [0, 0, 1.0]])
    H[2, 0] = rng.uniform(-persp, persp)
    H[2, 1] = rng.uniform(-persp, persp)
    return H


def gen_warp():
    rng = np.random.default_rng(12)
    out = {}
    n = 0
    for trial in range(10):
        h, w = int(rng.integers(30, 110)), int(rng.integers(40, 150))
        cn = [3, 3, 1, 4, 3][trial % 5]
        img = rng.integers(0, 256, (h, w, cn), dtype=np.uint8) if trial % 2 else smooth_img(rng, h, w)[:, :, :cn] if cn <= 3 else rng.integers(0, 256, (h, w, cn), dtype=np.uint8)
        if cn == 1:
            img = np.ascontiguousarray(img[:, :, 0])
        H = rand_H(rng, [0, 1e-4, 2e-3][trial % 3], 60, 40)
        if trial == 3:
            H = np.array([[1, 0, 12.0], [0, 1, -7.0], [0, 0, 1.0]])            # integer translation
        if trial == 4:
            H = np.array([[1, 0, 17.515625], [0, 1, 3.5], [0, 0, 1.0]])        # exact half-way roundings
        if trial == 5:
            H = np.eye(3)
        if trial == 6:
            H = np.array([[0.31, 0.02, 5.0], [-0.01, 0.29, 3.0], [0, 0, 1.0]])  # strong down-scale
        Wd, Hd = int(rng.integers(70, 230)), int(rng.integers(50, 170))
        out[f"img{n}"], out[f"H{n}"], out[f"dsize{n}"] = img, H, np.array([Wd, Hd])
        for fl, tag in ((cv2.INTER_LINEAR, "lin"), (cv2.INTER_CUBIC, "cub")):
            out[f"{tag}{n}"] = cv2.warpPerspective(img, H, (Wd, Hd), None, fl, borderMode=cv2.BORDER_CONSTANT, borderValue=0)
        n += 1
    out["n"] = np.array(n)
    np.savez_compressed(OUT / "warp.npz", **out)
    print("warp.npz cases", n)


def gen_mapper(mos):
    rng = np.random.default_rng(13)
    out = {}
    cases = []
    real_cubic = cv2.INTER_CUBIC
    for interp in (1, 2):
        for alpha, use_kp in ((0.5, False), (1.0, False), (0.3, True), (0.0, False)):
            tag = f"c{len(cases)}"
            Hc, Wc = 150, 210
            if interp == 1:
                cv2.INTER_CUBIC = cv2.INTER_LINEAR   # the LINEAR-patched reference (north-star kernel)
            try:
                mp = mos.Mapper(Wc, Hc, alpha_blend=alpha)
                for k in range(3):
                    h, w = int(rng.integers(50, 90)), int(rng.integers(60, 120))
                    img = smooth_img(rng, h, w)
                    th = rng.uniform(-0.5, 0.5)
                    s = rng.uniform(0.8, 1.3)
                    H = np.array([[s * np.cos(th), -s * np.sin(th), rng.uniform(-10, 110)],
                                  [s * np.sin(th), s * np.cos(th), rng.uniform(-10, 70)],
                                  [rng.uniform(-2e-4, 2e-4), rng.uniform(-2e-4, 2e-4), 1.0]])
                    if k == 1:
                        H[2, :2] = 0.0   # affine member of every sequence
                    kps = None
                    if use_kp:
                        pts = np.stack([rng.uniform(3, w - 3, 30), rng.uniform(3, h - 3, 30)], 1).astype(np.float32)
                        kps = [cv2.KeyPoint(float(x), float(y), 1) for x, y in pts]
                        out[f"{tag}_kp{k}"] = pts
                        out[f"{tag}_roi{k}"] = mp._create_keypoint_mask(cv2.KeyPoint.convert(kps), (w, h))
                    mp.update(img, H, None, kps)
                    out[f"{tag}_img{k}"], out[f"{tag}_H{k}"] = img, H
                    out[f"{tag}_canvas{k}"] = mp.output.copy()
                    out[f"{tag}_mask{k}"] = mp._canvas_mask.copy()
            finally:
                cv2.INTER_CUBIC = real_cubic
            cases.append((interp, alpha, int(use_kp), 3, Hc, Wc))
    out["cases"] = np.array(cases, np.float64)
    np.savez_compressed(OUT / "mapper.npz", **out)
    print("mapper.npz cases", len(cases))


def gen_routing(reg, mos):
    rng = np.random.default_rng(14)
    sys.path.insert(0, str(ref_shim.REFERENCE_SRC))
    from mosaicking import splitter
    Hs = np.stack([rand_H(rng, 1e-5, 3000, 2000) for _ in range(12)])
    Hs[:, 0, 2] += 3000; Hs[:, 1, 2] += 2000
    w, h = 640, 360
    out = {"H": Hs, "wh": np.array([w, h])}
    out["corners"] = mos.get_corners(Hs, w, h)
    out["corners_single"] = mos.get_corners(Hs[0], w, h)
    out["dims"] = np.array(mos.get_mosaic_dimensions(Hs, w, h), np.float64)
    out["nice"] = mos.find_nice_homographies(Hs)
    tile = (1024, 768)
    txs = list(range(0, 7000, tile[0])); tys = list(range(0, 5000, tile[1]))
    ov = np.zeros((len(Hs), len(txs), len(tys)), bool)
    for k, H in enumerate(Hs):
        crn = mos.get_corners(H, w, h)
        for i, tx in enumerate(txs):
            for j, ty in enumerate(tys):
                tc = np.array([[tx, ty], [tx + tile[0], ty], [tx + tile[0], ty + tile[1]], [tx, ty + tile[1]]], dtype=int)
                ov[k, i, j] = reg.bbox_overlap(tc, crn.squeeze().astype(int))     # mosaic.py:612-614
    out["tile"] = np.array(tile); out["overlap"] = ov
    tiles = splitter.calculate_tiles((0, 0, 6500, 4700), tile)
    out["split_tiles"] = np.array(tiles)
    asg = np.zeros((len(Hs), len(tiles)), bool)
    for k, H in enumerate(Hs):
        crn = mos.get_corners(H, w, h).reshape(-1, 2)
        for a in splitter.assign_image_to_tiles(crn, tiles):
            asg[k, a] = True
    out["split_assign"] = asg
    np.savez_compressed(OUT / "routing.npz", **out)
    print("routing.npz")


def gen_pipeline(reg, mos):
    """The reference's own integration test (tests/test_mosaic.py:28-73: ORB, keypoint_roi=True, alpha=1.0, CLAHE, tile
    4096x4096) run for real on tests/data/sparus_snippet.mp4, with Mapper.update instrumented: the first NUPD calls
    (frame after preprocessing, tile-local H, inlier keypoints) and the canvas / mask they produce are recorded."""
    import tempfile
    import yaml
    NUPD = 6
    data = ref_shim.REFERENCE_SRC.parent / "tests" / "data"
    with open(data / "sparus_time_offset.yaml") as f:
        offset = yaml.safe_load(f)["video"]
    K = np.array([405.6384738851233, 0, 189.9054317917407, 0, 405.588335378204, 139.9149578253755, 0, 0, 1]).reshape((3, 3))
    D = np.array([-0.3670656233416921, 0.203001968694465, 0.003336917744124004, -0.000487426354679637, 0])
    rec = {"calls": [], "canvas": None, "mask": None}
    real_update = mos.Mapper.update

    def spy(self, image, H, stream=None, keypoints=None):
        real_update(self, image, H, stream, keypoints)
        if len(rec["calls"]) < NUPD:
            kp = None if not keypoints else np.array(cv2.KeyPoint.convert(keypoints), np.float32)
            rec["calls"].append((np.ascontiguousarray(image).copy(), np.array(H, np.float64), kp))
            if len(rec["calls"]) == NUPD:
                rec["canvas"], rec["mask"] = self._canvas.copy(), self._canvas_mask.copy()

    mos.Mapper.update = spy
    try:
        with tempfile.TemporaryDirectory() as tmp:
            m = mos.SequentialMosaic(data_path=data / "sparus_snippet.mp4", project_path=Path(tmp), feature_types=("ORB",),
                                     intrinsics={"K": K, "D": D}, orientation_path=data / "sparus_snippet.csv",
                                     orientation_time_offset=offset, force_cpu=True, keypoint_roi=True, bovw_clusters=5, alpha=1.0)
            m.extract_features(); m.match_features(); m.registration(); m.global_registration()
            m.generate((4096, 4096))
    finally:
        mos.Mapper.update = real_update
    assert len(rec["calls"]) == NUPD, "the reference pipeline produced fewer updates than expected"
    ys, xs = np.nonzero(rec["mask"])
    y0, y1, x0, x1 = max(0, ys.min() - 8), ys.max() + 9, max(0, xs.min() - 8), xs.max() + 9
    out = {"n": np.array(NUPD), "alpha": np.array(1.0), "tile": np.array([4096, 4096]), "crop": np.array([y0, y1, x0, x1]),
           "canvas_crop": rec["canvas"][y0:y1, x0:x1], "mask_crop": rec["mask"][y0:y1, x0:x1],
           "outside_is_zero": np.array(int(rec["canvas"].sum() == rec["canvas"][y0:y1, x0:x1].sum()))}
    for k, (img, H, kp) in enumerate(rec["calls"]):
        out[f"img{k}"], out[f"H{k}"] = img, H
        out[f"kp{k}"] = kp if kp is not None else np.zeros((0, 2), np.float32)
    np.savez_compressed(OUT / "pipeline.npz", **out)
    print("pipeline.npz", NUPD, "updates, crop", (y0, y1, x0, x1), "frame", rec["calls"][0][0].shape)


def gen_matchfeatures(reg, mos):
    """SequentialMosaic.match_features (mosaic.py:402-438) and .registration (mosaic.py:440-470) of the reference run for real on
    its test video, instrumented: for the first NP consecutive-frame pairs the descriptor dicts handed to
    CompositeMatcher.knn_match (query = later frame, train = earlier frame), the good matches the reference keeps after the
    literal ratio comprehension of mosaic.py:430, and the keypoint arrays + (H, inliers) of the RANSAC call of mosaic.py:459."""
    import tempfile
    import yaml
    NP = 4
    data = ref_shim.REFERENCE_SRC.parent / "tests" / "data"
    with open(data / "sparus_time_offset.yaml") as f:
        offset = yaml.safe_load(f)["video"]
    K = np.array([405.6384738851233, 0, 189.9054317917407, 0, 405.588335378204, 139.9149578253755, 0, 0, 1]).reshape((3, 3))
    D = np.array([-0.3670656233416921, 0.203001968694465, 0.003336917744124004, -0.000487426354679637, 0])
    rec = {"knn": [], "good": [], "ransac": []}
    real_knn = reg.CompositeMatcher.knn_match
    real_est = reg.HomographyEstimation.__call__

    def spy_knn(self, query, train, mask=None):
        out = real_knn(self, query, train, mask)
        if len(rec["knn"]) < NP:
            rec["knn"].append(({k: np.ascontiguousarray(v).copy() for k, v in query.items()}, {k: np.ascontiguousarray(v).copy() for k, v in train.items()}))
            good = {ft: tuple(m[0] for m in ms if len(m) == 2 and m[0].distance < 0.7 * m[1].distance) for ft, ms in out.items()}   # mosaic.py:430
            rec["good"].append({ft: np.array([(m.queryIdx, m.trainIdx, m.distance) for m in g], np.float64).reshape(-1, 3) for ft, g in good.items()})
        return out

    def spy_est(self, src_points, dst_points, *args, **kwargs):
        H, inl = real_est(self, src_points, dst_points, *args, **kwargs)
        if len(rec["ransac"]) < NP:
            rec["ransac"].append((np.array(src_points, np.float32), np.array(dst_points, np.float32), np.array(H, np.float64), np.array(inl, np.uint8).ravel(), str(self.method)))
        return H, inl

    reg.CompositeMatcher.knn_match = spy_knn
    reg.HomographyEstimation.__call__ = spy_est
    try:
        with tempfile.TemporaryDirectory() as tmp:
            m = mos.SequentialMosaic(data_path=data / "sparus_snippet.mp4", project_path=Path(tmp), feature_types=("ORB",),
                                     intrinsics={"K": K, "D": D}, orientation_path=data / "sparus_snippet.csv",
                                     orientation_time_offset=offset, force_cpu=True, keypoint_roi=True, bovw_clusters=5, alpha=1.0)
            m.extract_features(); m.match_features(); m.registration()
    finally:
        reg.CompositeMatcher.knn_match = real_knn
        reg.HomographyEstimation.__call__ = real_est
    assert len(rec["knn"]) == NP and len(rec["ransac"]) >= 1
    out = {"npairs": np.array(NP), "nransac": np.array(len(rec["ransac"]))}
    for k, ((q, t), good) in enumerate(zip(rec["knn"], rec["good"])):
        assert list(q) == ["ORB"], list(q)
        out[f"q{k}"], out[f"t{k}"], out[f"good{k}"] = q["ORB"], t["ORB"], good["ORB"]
    for k, (src, dst, H, inl, method) in enumerate(rec["ransac"]):
        out[f"rsrc{k}"], out[f"rdst{k}"], out[f"rH{k}"], out[f"rinl{k}"] = src, dst, H, inl
        out["rmethod"] = np.array(method)
    np.savez_compressed(OUT / "matchfeatures.npz", **out)
    print("matchfeatures.npz", NP, "pairs", [out[f"q{k}"].shape for k in range(NP)], [len(out[f"good{k}"]) for k in range(NP)], "ransac", len(rec["ransac"]), out["rmethod"],
          [int(out[f"rinl{k}"].sum()) for k in range(len(rec["ransac"]))])


def gen_preproc():
    import mosaicking.preprocessing as pre      # the real reference module (imported after ref_shim.import_reference())
    rng = np.random.default_rng(77)
    out = {}
    for tag, (h, w) in {"a": (131, 203), "b": (96, 160)}.items():      # odd width: scalar row tails; 160: the 16-pixel vector path
        bgr = smooth_img(rng, h, w); bgra = np.dstack([bgr, rng.integers(0, 256, (h, w), dtype=np.uint8)]); gray = rng.integers(0, 256, (h, w), dtype=np.uint8)
        out[f"{tag}_bgr"] = bgr; out[f"{tag}_bgra"] = bgra; out[f"{tag}_gray"] = gray
        out[f"{tag}_gray_of_bgr"] = pre.make_gray(bgr); out[f"{tag}_gray_of_bgra"] = pre.make_gray(bgra)
        out[f"{tag}_bgr_of_gray"] = pre.make_bgr(gray); out[f"{tag}_bgr_of_bgra"] = pre.make_bgr(bgra)
        assert pre.make_gray(gray) is gray and pre.make_bgr(bgr) is bgr
    h, w = 192, 256
    K = np.array([[210.0, 0, 126.3], [0, 205.0, 97.7], [0, 0, 1]]); D = np.array([-0.28, 0.09, 0.0012, -0.0007, -0.01])
    img = smooth_img(rng, h, w); g1 = rng.integers(0, 256, (h, w), dtype=np.uint8)
    out["K"] = K; out["D"] = D; out["und_img"] = img; out["und_gray"] = g1
    for inv in (False, True):
        dm = pre.DistortionMapper(K, D, inverse=inv)
        t = "inv" if inv else "fwd"
        out[f"und_{t}_bgr"] = dm.apply(img); out[f"und_{t}_gray"] = dm.apply(g1)
        if not inv:
            xmap, ymap = cv2.initUndistortRectifyMap(K, D, None, K, (w, h), cv2.CV_32FC1)
        else:
            xmap, ymap = cv2.initInverseRectificationMap(K, D, np.eye(3), K, (w, h), cv2.CV_32FC1)
        out[f"und_{t}_xmap"] = xmap; out[f"und_{t}_ymap"] = ymap
        out[f"und_{t}_linear"] = cv2.remap(img, xmap, ymap, cv2.INTER_LINEAR)
    np.savez_compressed(OUT / "preproc.npz", **out)
    print("preproc.npz", {k: v.shape for k, v in out.items() if k.startswith("und_")})


def main():
    OUT.mkdir(parents=True, exist_ok=True)
    reg, mos = ref_shim.import_reference()
    cv2.setRNGSeed(0)
    if len(sys.argv) > 1 and sys.argv[1] == "preproc":
        gen_preproc()
        return
    if len(sys.argv) > 1 and sys.argv[1] == "matchfeatures":      # add this fixture without regenerating the others
        gen_matchfeatures(reg, mos)
        return
    gen_knn(reg)
    gen_warp()
    gen_mapper(mos)
    gen_routing(reg, mos)
    gen_pipeline(reg, mos)
    gen_matchfeatures(reg, mos)
    gen_preproc()
    (OUT / "PROVENANCE.txt").write_text(
        f"generated by oracle/make_golden.py\ncv2 {cv2.__version__}, numpy {np.__version__}\n"
        f"reference: DTUAqua-ObsTek/mosaic-library checkout at {ref_shim.REFERENCE_SRC.parent}\n")


if __name__ == "__main__":
    main()
