"""K4 timing on the bench trajectory: isolated launches (L2 flushed, CUDA events around one launch) and back-to-back groups
over distinct frames (working set > L2); MK_STRIP=0 in the environment selects the tile-per-CTA kernel for A/B runs.
Also checks a few launches against the oracle on a small canvas (bit-exact)."""
import os, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np, torch
import mosaicking_b200 as mb
from bench import trajectory

dev = torch.device("cuda:0")
flush = torch.empty(1024 << 20, dtype=torch.uint8, device=dev)
rng = np.random.default_rng(0)
tag = "strip" if os.environ.get("MK_STRIP", "1") != "0" else "tile"
INTERP = os.environ.get("K4_INTERP", "linear")        # K4_INTERP=cubic times the reference Mapper's default interpolation


def parity():
    from oracle import oracle as orc
    bad = 0
    for interp, oi in (("linear", orc.INTER_LINEAR), ("cubic", orc.INTER_CUBIC)):
        for alpha in (0.5, 0.3, 1.0, 0.0):
            Wc, Hc = 1500, 1100
            mp = mb.Mapper(Wc, Hc, alpha_blend=alpha, interp=interp)
            canvas = np.zeros((Hc, Wc, 3), np.uint8); cmask = np.zeros((Hc, Wc), np.uint8)
            for k in range(4):
                img = rng.integers(0, 256, (400, 640, 3), dtype=np.uint8)
                img[rng.random((400, 640)) < 0.02] = 0
                th = 0.15 * k - 0.1
                H = np.array([[np.cos(th), -np.sin(th), 300.0 + 170 * k], [np.sin(th), np.cos(th), 200.0 + 90 * k], [2e-5 * (k % 2), 1e-5 * (k % 2), 1.0]])
                if k == 3:
                    H[0, 2] = 1100.0; H[1, 2] = -150.0      # crosses the canvas corner
                mp.update(img, H)
                orc.mapper_update(canvas, cmask, img, H, alpha, oi)
            ok = np.array_equal(mp.output, canvas) and np.array_equal(mp.mask, cmask)
            if not ok:
                d = np.argwhere((mp.output != canvas).any(axis=2)); dm = np.argwhere(mp.mask != cmask)
                print(f"[{tag}] MISMATCH {interp} alpha={alpha}: {len(d)} px, {len(dm)} mask; first {d[:3].tolist()} {dm[:3].tolist()}")
                bad += 1
            mp.release()
    print(f"[{tag}] parity vs oracle: {'OK' if not bad else f'{bad} FAILED'}")


def timing(h, w, n=24):
    C = 16384
    frames = [torch.from_numpy(rng.integers(1, 256, (h, w, 3), dtype=np.uint8)).to(dev) for _ in range(n)]
    Hs = [trajectory(k + 5, w, h, C) for k in range(n)]
    mp = mb.Mapper(C, C, alpha_blend=0.5, interp=INTERP)
    for k in range(3):
        mp.update_device(frames[k], Hs[k])
    iso = []
    for k in range(n):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); mp.update_device(frames[k], Hs[k]); b.record(); torch.cuda.synchronize()
        iso.append(a.elapsed_time(b) * 1e3)
    grp = []
    for rep in range(5):
        for _ in range(6):
            flush.zero_()          # ~1 ms of queued GPU work: the host gets ahead of the GPU, the span below is GPU time only
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for k in range(n):
            mp.update_device(frames[k], Hs[k])
        b.record(); torch.cuda.synchronize()
        grp.append(a.elapsed_time(b) * 1e3 / n)
    nbytes = np.mean([mp.footprint_bytes(h, w, H) for H in Hs])
    iso = np.array(iso)
    print(f"[{tag}] {INTERP} {w}x{h}: isolated min/med/max {iso.min():.1f}/{np.median(iso):.1f}/{iso.max():.1f} us -> {nbytes / np.median(iso) / 1e3:.0f} GB/s; "
          f"back-to-back {np.median(grp):.1f} us/launch -> {nbytes / np.median(grp) / 1e3:.0f} GB/s ({nbytes / 1e6:.1f} MB algorithmic)")
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    mp.release()


if __name__ == "__main__":
    parity()
    timing(1080, 1920)
    timing(2160, 3840, n=8)
