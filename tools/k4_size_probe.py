import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np, torch
import mosaicking_b200 as mb
dev = torch.device("cuda:0")
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
rng = np.random.default_rng(0)
mp = mb.Mapper(16384, 16384, alpha_blend=0.5, interp="linear")
H = np.array([[np.cos(0.05), -np.sin(0.05), 4000.3], [np.sin(0.05), np.cos(0.05), 3000.7], [0, 0, 1.0]])
def timeit(fn, n=20, warm=3, do_flush=True):
    for _ in range(warm): fn()
    ts = []
    for _ in range(n):
        if do_flush: flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    return sorted(ts)[len(ts) // 2]
empty = torch.empty(1, device=dev)
print("event-bracketed tiny torch kernel: %.1f us" % timeit(lambda: empty.zero_()))
for (h, w) in ((32, 64), (128, 128), (256, 256), (512, 512), (720, 1280), (1080, 1920), (2160, 3840)):
    fr = torch.from_numpy(rng.integers(1, 256, (h, w, 3), dtype=np.uint8)).to(dev)
    a = timeit(lambda: mp.update_device(fr, H)); b = timeit(lambda: mp.update_device(fr, H), do_flush=False)
    print(f"{w}x{h}: {a:.1f} us cold L2, {b:.1f} us warm   ({h*w/1e6:.2f} Mpx)")

# back-to-back trains (device time only: ~1 ms of flushes queued in front so that the host is ahead of the GPU)
big = torch.empty(1024 << 20, dtype=torch.uint8, device=dev)
for (h, w) in ((32, 64), (256, 256), (540, 960), (720, 1280), (1080, 1920), (1530, 2720), (2160, 3840)):
    frs = [torch.from_numpy(rng.integers(1, 256, (h, w, 3), dtype=np.uint8)).to(dev) for _ in range(8)]
    for n in (8, 24, 48):
        for _ in range(6): big.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for k in range(n): mp.update_device(frs[k % 8], H)
        b.record(); torch.cuda.synchronize()
        print(f"{w}x{h}: train of {n}: {a.elapsed_time(b) * 1e3 / n:.2f} us/launch")
