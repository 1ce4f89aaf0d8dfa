import sys, time
import numpy as np, torch
dev = torch.device("cuda:0")
fh, fw = 1080, 1920
bufs = [torch.empty((fh, fw, 3), dtype=torch.uint8).pin_memory() for _ in range(4)]
dd = torch.empty((fh, fw, 3), dtype=torch.uint8, device=dev)
def burst(n, label):
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(n + 1)]
    evs[0].record()
    for k in range(n):
        dd.copy_(bufs[k % 4], non_blocking=True); evs[k + 1].record()
    torch.cuda.synchronize()
    ts = [evs[k].elapsed_time(evs[k + 1]) * 1e3 for k in range(n)]
    print(label, "copies 0-9: %.0f us, 10-49: %.0f, 50-99: %.0f, 100-199: %.0f, 200-399: %.0f, 400+: %.0f" % (
        np.mean(ts[:10]), np.mean(ts[10:50]), np.mean(ts[50:100]), np.mean(ts[100:200]), np.mean(ts[200:400]), np.mean(ts[400:])), flush=True)
burst(600, "cold start      ")
burst(600, "right after     ")
time.sleep(0.5); burst(600, "after 0.5 s idle")
time.sleep(0.05); burst(600, "after 50 ms idle")
x = torch.empty(1 << 30, dtype=torch.uint8, device=dev)
for _ in range(50): x.zero_()
torch.cuda.synchronize(); burst(600, "after GPU-busy   ")
time.sleep(0.5)
big = [torch.empty(16384 * 16384, dtype=torch.float32, device=dev) for _ in range(3)]
for b in big: b.zero_()
torch.cuda.synchronize(); del big; torch.cuda.empty_cache()
b2 = [torch.empty((fh, fw, 3), dtype=torch.uint8).pin_memory() for _ in range(4)]
bufs = b2
burst(600, "after empty_cache + new pinned")
burst(600, "again                        ")
