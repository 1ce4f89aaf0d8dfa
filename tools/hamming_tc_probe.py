"""Hamming kNN: XOR+POPC kernel vs the optional tcgen05 fp8 engine (CUDA events, L2 flushed)."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np, torch
import mosaicking_b200 as mb

dev = torch.device("cuda:0")
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)

def timeit(fn, n=20, warm=3):
    for _ in range(warm): fn()
    ts = []
    for _ in range(n):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    ts.sort()
    return ts[len(ts) // 2], ts[0]

rng = np.random.default_rng(0)
for n in (2000, 5000, 10000, 20000):
    q = torch.from_numpy(rng.integers(0, 256, (n, 32), dtype=np.uint8)).to(dev)
    t = torch.from_numpy(rng.integers(0, 256, (n, 32), dtype=np.uint8)).to(dev)
    res = {}
    for eng in ("popc", "tensor"):
        m = mb.Matcher(norm="hamming", engine=eng)
        med, best = timeit(lambda: m.knn_match_device(q, t, 0.7))
        res[eng] = [x.cpu() for x in m.knn_match_device(q, t, 0.7)]
        print(f"hamming {eng:6s} {n}x{n}x32B: median {med:.1f} us best {best:.1f} us -> {n*n/med*1e6:.3e} pairs/s", flush=True)
    same = all(torch.equal(a, b) for a, b in zip(res["popc"], res["tensor"]))
    print("  identical results:", same)
