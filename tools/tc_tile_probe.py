"""Steady-state cost per 128x256 tile of the tcgen05 kNN kernel: one full wave of CTAs (148 x 128 queries, one split),
train rows swept; a linear fit of time vs tiles separates the per-tile cost from the per-call constant."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np, torch
import mosaicking_b200 as mb

dev = torch.device("cuda:0")
def timeit(fn, n=10, warm=2):
    for _ in range(warm): fn()
    ts = []
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    return sorted(ts)[len(ts) // 2]

rng = np.random.default_rng(0)
nq = 148 * 128
for kind in ("hamming_tc", "l2_u8", "l2_f32"):
    xs, ys = [], []
    for tiles in (8, 16, 32, 64, 128):
        nt = tiles * 256
        if kind == "l2_f32":
            q = torch.from_numpy(rng.integers(0, 200, (nq, 128)).astype(np.float32)).to(dev)
            t = torch.from_numpy(rng.integers(0, 200, (nt, 128)).astype(np.float32)).to(dev)
            m = mb.Matcher()
        else:
            q = torch.from_numpy(rng.integers(0, 256, (nq, 32), dtype=np.uint8)).to(dev)
            t = torch.from_numpy(rng.integers(0, 256, (nt, 32), dtype=np.uint8)).to(dev)
            m = mb.Matcher(norm="hamming", engine="tensor") if kind == "hamming_tc" else mb.Matcher(norm="l2")
        us = timeit(lambda: m.knn_match_device(q, t, 0.7))
        xs.append(tiles); ys.append(us)
        print(f"{kind} nq={nq} nt={nt} ({tiles} tiles/CTA): {us:.1f} us", flush=True)
    a, b = np.polyfit(xs, ys, 1)
    print(f"  -> {a:.3f} us per tile ({a*1.965e3:.0f} cycles at 1965 MHz), constant {b:.1f} us")
