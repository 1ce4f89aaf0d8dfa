import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np, torch
import mosaicking_b200 as mb
rng = np.random.default_rng(0)
mp = mb.Mapper(16384, 16384, alpha_blend=0.5, interp="linear")
frames = [torch.from_numpy(rng.integers(1, 256, (1080, 1920, 3), dtype=np.uint8)).pin_memory() for _ in range(4)]
desc = [torch.from_numpy(rng.integers(0, 256, (5000, 32), dtype=np.uint8)).pin_memory() for _ in range(4)]
H = np.array([[1, 0, 4000.3], [0, 1, 3000.7], [0, 0, 1.0]])
m = mb.Matcher(norm="hamming")
def loop(fn, n=200):
    for _ in range(5): fn(0)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for k in range(n): fn(k)
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e6
print("mapper.update host pinned  : %.0f us" % loop(lambda k: mp.update(frames[k % 4].numpy(), H)))
print("matcher host pinned        : %.0f us" % loop(lambda k: m.knn_match_arrays(desc[(k + 1) % 4].numpy(), desc[k % 4].numpy(), 0.7)))
print("both                       : %.0f us" % loop(lambda k: (mp.update(frames[k % 4].numpy(), H), m.knn_match_arrays(desc[(k + 1) % 4].numpy(), desc[k % 4].numpy(), 0.7))))
dev = frames[0].cuda()
print("mapper.update_device       : %.0f us" % loop(lambda k: mp.update_device(dev, H)))
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
for k in range(200): mp.update(frames[k % 4].numpy(), H); m.knn_match_arrays(desc[(k + 1) % 4].numpy(), desc[k % 4].numpy(), 0.7)
pr.disable(); pstats.Stats(pr).sort_stats("cumulative").print_stats(18)

# ---- pipelined loop: where does the host time go? (pure CPU cost of each call, GPU work is asynchronous) ----
def cpu_cost(fn, n=300):
    for _ in range(5): fn(0)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for k in range(n): fn(k)
    dt = (time.perf_counter() - t0) / n * 1e6
    torch.cuda.synchronize()
    return dt
pend = []
print("CPU cost update()               : %.0f us" % cpu_cost(lambda k: mp.update(frames[k % 4].numpy(), H), 60))
print("CPU cost knn_match_arrays_async : %.0f us" % cpu_cost(lambda k: pend.append(m.knn_match_arrays_async(desc[(k + 1) % 4].numpy(), desc[k % 4].numpy(), 0.7)), 60))
t0 = time.perf_counter(); [p.result() for p in pend[-40:]]; print("CPU cost result() (already done): %.0f us" % ((time.perf_counter() - t0) / 40 * 1e6))
def piped(n):
    p = None
    for k in range(n):
        nxt = m.knn_match_arrays_async(desc[(k + 1) % 4].numpy(), desc[k % 4].numpy(), 0.7)
        mp.update(frames[k % 4].numpy(), H)
        if p is not None: p.result()
        p = nxt
    p.result()
piped(10); torch.cuda.synchronize(); t0 = time.perf_counter(); piped(300); torch.cuda.synchronize()
print("pipelined loop                  : %.0f us/frame" % ((time.perf_counter() - t0) / 300 * 1e6))
pr = cProfile.Profile(); pr.enable(); piped(300); pr.disable(); pstats.Stats(pr).sort_stats("tottime").print_stats(22)
