import sys, time
import numpy as np, torch
dev = torch.device("cuda:0")
fh, fw = 1080, 1920
rng = np.random.default_rng(0)
dd = torch.empty((fh, fw, 3), dtype=torch.uint8, device=dev)
def rate(bufs, n=40):
    for k in range(4): dd.copy_(bufs[k % len(bufs)], non_blocking=True)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for k in range(n): dd.copy_(bufs[k % len(bufs)], non_blocking=True)
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e6
def mk_a(n): return [torch.from_numpy(rng.integers(1, 256, (fh, fw, 3), dtype=np.uint8)).pin_memory() for _ in range(n)]
def mk_b(n, src):
    out = [torch.empty((fh, fw, 3), dtype=torch.uint8).pin_memory() for _ in range(n)]
    for i, pf in enumerate(out): pf.copy_(src[i].cpu())
    return out
def mk_c(n): return [torch.empty((fh, fw, 3), dtype=torch.uint8).pin_memory() for _ in range(n)]
a = mk_a(4)
print("A from_numpy().pin_memory(), 4 bufs cycled: %.0f us" % rate(a), " single: %.0f us" % rate(a[:1]))
devf = [x.to(dev) for x in a]
b = mk_b(4, devf)
print("B empty().pin_memory()+copy_(dev.cpu()), 4 bufs: %.0f us" % rate(b), " single: %.0f us" % rate(b[:1]))
c = mk_c(4)
print("C empty().pin_memory() untouched, 4 bufs: %.0f us" % rate(c), " single: %.0f us" % rate(c[:1]))
big = torch.empty(8 << 30, dtype=torch.uint8, device=dev); big.zero_(); torch.cuda.synchronize()
print("after touching 8 GiB of HBM:  A %.0f  B %.0f  C %.0f" % (rate(a), rate(b), rate(c)))
b2 = mk_b(4, devf)
print("B allocated now: %.0f us" % rate(b2))
for i in range(3):
    print("repeat A %.0f  B %.0f  C %.0f  B2 %.0f" % (rate(a), rate(b), rate(c), rate(b2)))
