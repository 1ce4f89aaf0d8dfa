"""K3s (plain warpPerspective) timing: isolated calls (L2 flushed, CUDA events, device frame in place, preallocated out) and
back-to-back trains over distinct frames; bit-exact check against the oracle first."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np, torch
import mosaicking_b200 as mb
from oracle import oracle as orc

dev = torch.device("cuda:0")
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
rng = np.random.default_rng(3)
bad = 0
for interp, code in (("linear", 1), ("cubic", 2)):
    for (h, w, Wd, Hd, th, px) in ((300, 400, 640, 480, 0.3, 0.0), (1080, 1920, 1920, 1080, 0.1, 1e-6), (700, 900, 1500, 333, -2.0, 0.0), (64, 64, 4000, 40, 0.0, 0.0)):
        img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        H = np.array([[np.cos(th), -np.sin(th), 50.5], [np.sin(th), np.cos(th), 20.25], [px, -px, 1.0]])
        got = mb.warp_perspective(img, H, (Wd, Hd), interp).cpu().numpy()
        if not np.array_equal(got, orc.warp_perspective(img, H, (Wd, Hd), code)):
            bad += 1; print("MISMATCH", interp, h, w, Wd, Hd)
print("parity vs oracle:", "OK" if not bad else f"{bad} FAILED")
for name, (h, w) in (("1080p", (1080, 1920)), ("4k", (2160, 3840))):
    imgs = [torch.from_numpy(rng.integers(1, 256, (h, w, 3), dtype=np.uint8)).to(dev) for _ in range(12)]
    outs = [torch.empty((h, w, 3), dtype=torch.uint8, device=dev) for _ in range(12)]
    th = 0.1
    c, s = np.cos(th), np.sin(th)
    H = np.array([[c, -s, w / 2 - c * w / 2 + s * h / 2], [s, c, h / 2 - s * w / 2 - c * h / 2], [0, 0, 1.0]])
    for interp in ("linear", "cubic"):
        for i in range(3): mb.warp_perspective(imgs[i], H, (w, h), interp, out=outs[i])
        iso = []
        for i in range(12):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); mb.warp_perspective(imgs[i], H, (w, h), interp, out=outs[i]); b.record(); torch.cuda.synchronize()
            iso.append(a.elapsed_time(b) * 1e3)
        grp = []
        for rep in range(5):
            for _ in range(6): flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for i in range(12): mb.warp_perspective(imgs[i], H, (w, h), interp, out=outs[i])
            b.record(); torch.cuda.synchronize()
            grp.append(a.elapsed_time(b) * 1e3 / 12)
        nb = 2.0 * w * h * 3
        print(f"warp {name} {interp}: isolated min/med {min(iso):.1f}/{np.median(iso):.1f} us -> {nb / np.median(iso) / 1e3:.0f} GB/s; back-to-back {np.median(grp):.1f} us -> {nb / np.median(grp) / 1e3:.0f} GB/s")
