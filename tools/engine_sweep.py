"""Crossover between the XOR+POPC kernel and the tensor-core engine for Hamming (and SIMT vs TC for L2-u8)."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np, torch
import mosaicking_b200 as mb
dev = torch.device("cuda:0")
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
def timeit(fn, n=15, warm=3):
    for _ in range(warm): fn()
    ts = []
    for _ in range(n):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    return sorted(ts)[len(ts) // 2]
rng = np.random.default_rng(0)
L = mb._lib.lib()
for n in (256, 512, 768, 1000, 1500, 2000, 3000):
    q = torch.from_numpy(rng.integers(0, 256, (n, 32), dtype=np.uint8)).to(dev)
    t = torch.from_numpy(rng.integers(0, 256, (n, 32), dtype=np.uint8)).to(dev)
    a = timeit(lambda: mb.Matcher(norm="hamming", engine="popc").knn_match_device(q, t, 0.7))
    m2 = mb.Matcher(norm="hamming", engine="tensor")
    b = timeit(lambda: m2.knn_match_device(q, t, 0.7))
    print(f"n={n}: popc {a:.1f} us, auto {b:.1f} us (tensor cores used: {L.mk_knn2_uses_tensor_cores(3, n, n, 32)})", flush=True)
