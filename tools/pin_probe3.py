"""H2D rate vs number of page-locked buffers cycled, torch pin_memory() vs one madvise(HUGEPAGE) region registered with CUDA."""
import mmap, time, ctypes
import numpy as np, torch
dev = torch.device("cuda:0")
fh, fw = 1080, 1920
FB = fh * fw * 3
dd = torch.empty(FB, dtype=torch.uint8, device=dev)
def rate(bufs, n=200):
    for k in range(40): dd.copy_(bufs[k % len(bufs)], non_blocking=True)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for k in range(n): dd.copy_(bufs[k % len(bufs)], non_blocking=True)
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e6
junk = [np.random.default_rng(i).integers(0, 255, 50_000_000, dtype=np.uint8) for i in range(20)]   # 1 GB of host churn first
del junk
pins = [torch.empty(FB, dtype=torch.uint8).pin_memory() for _ in range(8)]
for b in pins: b.fill_(7)
print("pin_memory():     " + "  ".join(f"{n} bufs {rate(pins[:n]):.0f} us" for n in (1, 2, 4, 8)), flush=True)
# one 2 MB-aligned anonymous mapping with MADV_HUGEPAGE, registered as page-locked
SZ = ((8 * FB + (2 << 20) - 1) >> 21) << 21
mm = mmap.mmap(-1, SZ + (2 << 20), flags=mmap.MAP_PRIVATE | mmap.MAP_ANONYMOUS)
try:
    mm.madvise(mmap.MADV_HUGEPAGE)
except Exception as ex:
    print("madvise failed:", ex)
arr = np.frombuffer(mm, dtype=np.uint8)
off = (-arr.ctypes.data) % (2 << 20)
arr = arr[off:off + SZ]
arr[:] = 7                                            # first touch
rc = torch.cuda.cudart().cudaHostRegister(arr.ctypes.data, SZ, 0)
print("cudaHostRegister rc:", rc)
big = torch.from_numpy(arr)
regs = [big[i * FB:(i + 1) * FB] for i in range(8)]
print("hugepage region:  " + "  ".join(f"{n} bufs {rate(regs[:n]):.0f} us" for n in (1, 2, 4, 8)), flush=True)
print("pin_memory() again: " + "  ".join(f"{n} bufs {rate(pins[:n]):.0f} us" for n in (1, 4, 8)), flush=True)
try:
    print("THP:", open("/sys/kernel/mm/transparent_hugepage/enabled").read().strip(), "| AnonHugePages:", [l for l in open("/proc/meminfo") if "AnonHuge" in l][0].strip())
except Exception as ex:
    print(ex)
