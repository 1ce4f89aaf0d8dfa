"""K4s with an axis-aligned trajectory (small source boxes: lets 4 CTAs per SM fit for the occupancy experiment)."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np, torch
import mosaicking_b200 as mb
dev = torch.device("cuda:0")
flush = torch.empty(1024 << 20, dtype=torch.uint8, device=dev)
rng = np.random.default_rng(0)
for (h, w, n) in ((1080, 1920, 24), (2160, 3840, 8)):
    C = 16384
    frames = [torch.from_numpy(rng.integers(1, 256, (h, w, 3), dtype=np.uint8)).to(dev) for _ in range(n)]
    Hs = [np.array([[1.0, 0, 2000.3 + 96 * k], [0, 1.0, 3000.7 + 11 * k], [0, 0, 1.0]]) for k in range(n)]
    mp = mb.Mapper(C, C, alpha_blend=0.5, interp="linear")
    for k in range(3):
        mp.update_device(frames[k], Hs[k])
    grp = []
    for rep in range(5):
        for _ in range(6):
            flush.zero_()          # ~1 ms of queued GPU work: the host gets ahead, the span below is GPU time only
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for k in range(n):
            mp.update_device(frames[k], Hs[k])
        b.record(); torch.cuda.synchronize()
        grp.append(a.elapsed_time(b) * 1e3 / n)
    print(f"{w}x{h} axis-aligned: back-to-back {np.median(grp):.1f} us/launch")
    mp.release()
