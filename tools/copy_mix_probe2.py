"""Copy-engine cost of the e2e transfer pattern (frame 6.2 MB + descriptor sets 160 KB up, 85 KB result down) by stream layout.
Uses the same huge-page page-locked buffers as the bench (pinned_empty)."""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np, torch
import mosaicking_b200 as mb
dev = torch.device("cuda:0")
fh, fw = 1080, 1920
frames = [torch.from_numpy(mb.pinned_empty((fh, fw, 3))) for _ in range(4)]
descs = [torch.from_numpy(mb.pinned_empty((5000, 32))) for _ in range(4)]
res = torch.from_numpy(mb.pinned_empty((85000,)))
dF = [torch.empty((fh, fw, 3), dtype=torch.uint8, device=dev) for _ in range(3)]
dq = torch.empty((5000, 32), dtype=torch.uint8, device=dev); dt = torch.empty_like(dq)
dres = torch.empty(85000, dtype=torch.uint8, device=dev)
sA, sB, sC = torch.cuda.Stream(), torch.cuda.Stream(), torch.cuda.Stream()
def run(mode, n=300):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for k in range(n):
        if mode == "frame | 2 small + d2h":
            with torch.cuda.stream(sB):
                dq.copy_(descs[k % 4], non_blocking=True); dt.copy_(descs[(k + 1) % 4], non_blocking=True)
                res.copy_(dres, non_blocking=True)
        elif mode == "frame | 1 small + d2h":
            with torch.cuda.stream(sB):
                dq.copy_(descs[k % 4], non_blocking=True); res.copy_(dres, non_blocking=True)
        elif mode == "frame | 2 small | d2h":
            with torch.cuda.stream(sB):
                dq.copy_(descs[k % 4], non_blocking=True); dt.copy_(descs[(k + 1) % 4], non_blocking=True)
            with torch.cuda.stream(sC):
                res.copy_(dres, non_blocking=True)
        elif mode == "frame + 2 small | d2h":
            with torch.cuda.stream(sA):
                dq.copy_(descs[k % 4], non_blocking=True); dt.copy_(descs[(k + 1) % 4], non_blocking=True)
            with torch.cuda.stream(sC):
                res.copy_(dres, non_blocking=True)
        elif mode == "frame + 1 small | d2h":
            with torch.cuda.stream(sA):
                dq.copy_(descs[k % 4], non_blocking=True)
            with torch.cuda.stream(sC):
                res.copy_(dres, non_blocking=True)
        elif mode == "frame | d2h":
            with torch.cuda.stream(sC):
                res.copy_(dres, non_blocking=True)
        with torch.cuda.stream(sA):
            dF[k % 3].copy_(frames[k % 4], non_blocking=True)
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e6
t_begin = time.perf_counter()
while time.perf_counter() - t_begin < 3.0:
    run("frame only", 200)
for _ in range(2):
    for mode in ("frame only", "frame | d2h", "frame | 2 small + d2h", "frame | 1 small + d2h", "frame | 2 small | d2h", "frame + 2 small | d2h", "frame + 1 small | d2h"):
        run(mode, 100)
        print(f"{mode:28s}: {run(mode):.1f} us / frame", flush=True)
