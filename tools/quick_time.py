"""Quick device-side timings of each kernel (CUDA events, L2 flushed between launches)."""
import sys, time
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
for norm in ("hamming", "l2"):
    for n in (2000, 5000):
        q = torch.from_numpy(rng.integers(0, 256, (n, 32), dtype=np.uint8)).to(dev)
        t = torch.from_numpy(rng.integers(0, 256, (n, 32), dtype=np.uint8)).to(dev)
        m = mb.Matcher(norm=norm)
        med, best = timeit(lambda: m.knn_match_device(q, t, 0.7))
        print(f"knn {norm} u8 {n}x{n}: median {med:.1f} us best {best:.1f} us -> {n*n/med*1e6:.3e} pairs/s")
for n in (4000, 8000):
    q = torch.from_numpy(rng.integers(0, 200, (n, 128)).astype(np.float32)).to(dev)
    t = torch.from_numpy(rng.integers(0, 200, (n, 128)).astype(np.float32)).to(dev)
    m = mb.Matcher()
    med, best = timeit(lambda: m.knn_match_device(q, t, 0.7), n=5, warm=1)
    print(f"knn l2 f32 (tcgen05 when integer-valued) {n}x{n}x128: median {med:.1f} us -> {n*n/med*1e6:.3e} pairs/s")

for (h, w, C) in ((1080, 1920, 16384), (2160, 3840, 16384)):
    for interp in ("linear", "cubic"):
        for blend in ("alpha", "feather"):
            mp = mb.Mapper(C, C, alpha_blend=0.5, interp=interp, blend=blend)
            frame = torch.from_numpy(rng.integers(1, 256, (h, w, 3), dtype=np.uint8)).to(dev)
            for th in (0.0, 0.1, 0.5):
                H = np.array([[np.cos(th), -np.sin(th), 4000.3], [np.sin(th), np.cos(th), 3000.7], [0, 0, 1.0]])
                nbytes = mp.footprint_bytes(h, w, H)
                med, best = timeit(lambda: mp.update_device(frame, H))
                print(f"mapper {w}x{h} {interp} {blend} theta={th}: median {med:.1f} us best {best:.1f} -> {nbytes/med/1e3:.0f} GB/s algorithmic ({nbytes/1e6:.1f} MB), {1e6/med:.0f} frames/s")
            Hp = np.array([[np.cos(0.1), -np.sin(0.1), 4000.3], [np.sin(0.1), np.cos(0.1), 3000.7], [1e-6, 2e-6, 1.0]])
            med, best = timeit(lambda: mp.update_device(frame, Hp))
            print(f"mapper {w}x{h} {interp} {blend} perspective: median {med:.1f} us -> {mp.footprint_bytes(h, w, Hp)/med/1e3:.0f} GB/s")
            mp.release()
# e2e host path
mp = mb.Mapper(16384, 16384, alpha_blend=0.5, interp="linear")
host = torch.from_numpy(rng.integers(1, 256, (1080, 1920, 3), dtype=np.uint8)).pin_memory()
H = np.array([[1, 0, 4000.3], [0, 1, 3000.7], [0, 0, 1.0]])
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(50): mp.update(host.numpy(), H)
torch.cuda.synchronize(); print(f"e2e host->update 1080p: {(time.perf_counter()-t0)/50*1e6:.0f} us/frame")
