"""Diagnostics for the device RANSAC against cv2 and the ground truth (numbers behind the tolerances in tests/test_gpu_parity.py)."""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np, cv2, torch
import mosaicking_b200 as mb

def corr(rng, n, H, noise, of, w=1920, h=1080):
    src = np.stack([rng.uniform(0, w, n), rng.uniform(0, h, n)], 1)
    p = np.c_[src, np.ones(n)] @ H.T
    dst = p[:, :2] / p[:, 2:3] + rng.normal(0, noise, (n, 2))
    out = rng.random(n) < of
    dst[out] = np.stack([rng.uniform(0, w, int(out.sum())), rng.uniform(0, h, int(out.sum()))], 1)
    return src.astype(np.float32), dst.astype(np.float32)

corners = np.array([[0, 0, 1], [1920, 0, 1], [1920, 1080, 1], [0, 1080, 1.0]])
pr = lambda M: (corners @ M.T)[:, :2] / (corners @ M.T)[:, 2:3]
for method in ("affine", "perspective"):
    rng = np.random.default_rng(5 if method == "affine" else 6)
    for trial in range(8):
        th = rng.uniform(-0.3, 0.3)
        H = np.array([[np.cos(th) * 1.02, -np.sin(th), rng.uniform(-60, 60)], [np.sin(th), np.cos(th) * 0.98, rng.uniform(-40, 40)], [0, 0, 1.0]])
        if method == "perspective":
            H[2, :2] = rng.uniform(-2e-5, 2e-5, 2)
        n = [900, 2500, 120, 5000][trial % 4]
        src, dst = corr(rng, n, H, 0.25, [0.3, 0.5, 0.2, 0.4][trial % 4])
        est = mb.HomographyEstimation(method)
        Hg, inl = est(src, dst, method=cv2.RANSAC, ransacReprojThreshold=1.0)
        if method == "affine":
            A, ci = cv2.estimateAffine2D(src, dst, method=cv2.RANSAC, ransacReprojThreshold=1.0); Hc = np.concatenate((A, [[0., 0., 1.]]), 0)
        else:
            Hc, ci = cv2.findHomography(src, dst, method=cv2.RANSAC, ransacReprojThreshold=1.0)
        p = np.c_[src.astype(np.float64), np.ones(n)] @ H.T
        te = np.linalg.norm(p[:, :2] / p[:, 2:3] - dst, axis=1)
        g, c = inl.ravel().astype(bool), ci.ravel().astype(bool)
        print(f"{method} n={n}: recall(te<.5) gpu {g[te < .5].mean():.4f} cv2 {c[te < .5].mean():.4f}; max te of inliers gpu {te[g].max():.2f} cv2 {te[c].max():.2f}; "
              f"mask diff {np.mean(g != c):.4f}; inliers {g.sum()} / {c.sum()}; corners vs gt gpu {np.abs(pr(Hg) - pr(H)).max():.3f} cv2 {np.abs(pr(Hc) - pr(H)).max():.3f} gpu-cv2 {np.abs(pr(Hg) - pr(Hc)).max():.3f}")
# timing: device-resident call
rng = np.random.default_rng(1)
H = np.array([[1.01, -0.05, 20], [0.05, 0.99, -10], [1e-5, 0, 1.0]])
src, dst = corr(rng, 2000, H, 0.25, 0.4)
for method in ("affine", "perspective"):
    est = mb.HomographyEstimation(method)
    s, d = torch.from_numpy(src).cuda(), torch.from_numpy(dst).cuda()
    for _ in range(3): est.estimate_device(s, d, 1.0)
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): est.estimate_device(s, d, 1.0)
    e1.record(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        (cv2.estimateAffine2D(src, dst, method=cv2.RANSAC, ransacReprojThreshold=1.0) if method == "affine" else cv2.findHomography(src, dst, method=cv2.RANSAC, ransacReprojThreshold=1.0))
    print(f"{method} 2000 matches: device {e0.elapsed_time(e1) / 20 * 1e3:.1f} us per call (2048 hypotheses + refit), cv2 {(time.perf_counter() - t0) / 5 * 1e3:.2f} ms")
