import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np, torch
import mosaicking_b200 as mb
from collections import defaultdict
rng = np.random.default_rng(0)
mp = mb.Mapper(16384, 16384, alpha_blend=0.5, interp="linear")
frames = [torch.from_numpy(rng.integers(1, 256, (1080, 1920, 3), dtype=np.uint8)).pin_memory() for _ in range(4)]
desc = [torch.from_numpy(rng.integers(0, 256, (5000, 32), dtype=np.uint8)).pin_memory() for _ in range(4)]
H = np.array([[1, 0, 4000.3], [0, 1, 3000.7], [0, 0, 1.0]])
m = mb.Matcher(norm="hamming")
L = mb._lib.lib()
acc = defaultdict(float); cnt = defaultdict(int)
class Timed:
    def __init__(self, name, fn): self.name, self.fn = name, fn
    def __call__(self, *a):
        t0 = time.perf_counter(); r = self.fn(*a); acc[self.name] += time.perf_counter() - t0; cnt[self.name] += 1; return r
class Proxy:
    def __init__(self, L): self._L = L; self._c = {}
    def __getattr__(self, k):
        if k not in self._c: self._c[k] = Timed(k, getattr(self._L, k))
        return self._c[k]
P = Proxy(L)
m._matcher["lib"] = P
mp._lib = P
def piped(n):
    p = None
    for k in range(n):
        nxt = m.knn_match_arrays_async(desc[(k + 1) % 4].numpy(), desc[k % 4].numpy(), 0.7)
        mp.update(frames[k % 4].numpy(), H)
        if p is not None: p.result()
        p = nxt
    p.result()
piped(20); torch.cuda.synchronize(); acc.clear(); cnt.clear()
t0 = time.perf_counter(); piped(300); torch.cuda.synchronize(); tot = (time.perf_counter() - t0) / 300 * 1e6
print("loop %.0f us/frame" % tot)
for k, v in sorted(acc.items(), key=lambda kv: -kv[1]):
    print(f"  {k:34s} {v / 300 * 1e6:7.1f} us/frame  ({cnt[k] / 300:.0f} calls/frame, {v / cnt[k] * 1e6:.1f} us each)")
