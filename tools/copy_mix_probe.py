import time
import numpy as np, torch
dev = torch.device("cuda:0")
fh, fw = 1080, 1920
frames = [torch.empty((fh, fw, 3), dtype=torch.uint8).pin_memory() for _ in range(4)]
descs = [torch.empty((5000, 32), dtype=torch.uint8).pin_memory() for _ in range(4)]
both = [torch.empty((10000, 32), dtype=torch.uint8).pin_memory() for _ in range(4)]
res = torch.empty(85000, dtype=torch.uint8).pin_memory()
dF = [torch.empty((fh, fw, 3), dtype=torch.uint8, device=dev) for _ in range(3)]
dq = torch.empty((5000, 32), dtype=torch.uint8, device=dev); dt = torch.empty_like(dq); dqt = torch.empty((10000, 32), dtype=torch.uint8, device=dev)
dres = torch.empty(85000, dtype=torch.uint8, device=dev)
sA, sB = torch.cuda.Stream(), torch.cuda.Stream()
def run(mode, n=300):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for k in range(n):
        if mode in ("frame+2small", "frame+2small+d2h"):
            with torch.cuda.stream(sB):
                dq.copy_(descs[k % 4], non_blocking=True); dt.copy_(descs[(k + 1) % 4], non_blocking=True)
                if mode.endswith("d2h"): res.copy_(dres, non_blocking=True)
        if mode == "frame+1small":
            with torch.cuda.stream(sB):
                dqt.copy_(both[k % 4], non_blocking=True)
        if mode == "frame,2small same stream":
            with torch.cuda.stream(sA):
                dq.copy_(descs[k % 4], non_blocking=True); dt.copy_(descs[(k + 1) % 4], non_blocking=True)
        with torch.cuda.stream(sA):
            dF[k % 3].copy_(frames[k % 4], non_blocking=True)
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e6
for _ in range(2):
    for mode in ("frame only", "frame+2small", "frame+2small+d2h", "frame+1small", "frame,2small same stream"):
        run(mode, 100)
        print(f"{mode:28s}: {run(mode):.0f} us / frame", flush=True)
