"""Device-resident pipeline WITHOUT the bench's L2 flush: frames and descriptors already in HBM, back-to-back steps
(matching on its own stream).  Shows what a real frame loop gets and whether the host or the GPU paces it."""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np, torch
import mosaicking_b200 as mb
from bench import trajectory
dev = torch.device("cuda:0")
rng = np.random.default_rng(0)
N = 200
frames = [torch.from_numpy(rng.integers(1, 256, (1080, 1920, 3), dtype=np.uint8)).to(dev) for _ in range(32)]
descs = [torch.from_numpy(rng.integers(0, 256, (5000, 32), dtype=np.uint8)).to(dev) for _ in range(33)]
Hs = [trajectory(k, 1920, 1080, 16384) for k in range(N)]
mp = mb.Mapper(16384, 16384, alpha_blend=0.5, interp="linear")
m = mb.Matcher(norm="hamming")
s_match = torch.cuda.Stream()
main = torch.cuda.current_stream()
def step(k):
    s_match.wait_stream(main)
    with torch.cuda.stream(s_match):
        m.knn_match_device(descs[k % 32 + 1], descs[k % 32], 0.7)
    mp.update_device(frames[k % 32], Hs[k])
    main.wait_stream(s_match)
def step_async(k):
    h = m.knn_match_device_async(descs[k % 32 + 1], descs[k % 32], 0.7)
    mp.update_device(frames[k % 32], Hs[k])
    return h.result()
for k in range(20): step(k)
torch.cuda.synchronize()
t0 = time.perf_counter()
for k in range(20, N): step(k)
t_cpu = time.perf_counter() - t0
torch.cuda.synchronize()
t_all = time.perf_counter() - t0
print(f"back-to-back steps, no flush: {t_all / (N - 20) * 1e6:.1f} us/step wall ({(N - 20) / t_all:.0f} frames/s); host issue time {t_cpu / (N - 20) * 1e6:.1f} us/step")
t0 = time.perf_counter()
for k in range(20, N): mp.update_device(frames[k % 32], Hs[k])
t_cpu = time.perf_counter() - t0; torch.cuda.synchronize(); t_all = time.perf_counter() - t0
print(f"compositing only: {t_all / (N - 20) * 1e6:.1f} us/step wall, host {t_cpu / (N - 20) * 1e6:.1f}")
t0 = time.perf_counter()
for k in range(20, N): m.knn_match_device(descs[k % 32 + 1], descs[k % 32], 0.7)
t_cpu = time.perf_counter() - t0; torch.cuda.synchronize(); t_all = time.perf_counter() - t0
print(f"matching only: {t_all / (N - 20) * 1e6:.1f} us/step wall, host {t_cpu / (N - 20) * 1e6:.1f}")

for k in range(20): step_async(k)
torch.cuda.synchronize()
t0 = time.perf_counter()
for k in range(20, N): step_async(k)
t_cpu = time.perf_counter() - t0; torch.cuda.synchronize(); t_all = time.perf_counter() - t0
print(f"knn_match_device_async + update_device: {t_all / (N - 20) * 1e6:.1f} us/step wall ({(N - 20) / t_all:.0f} frames/s), host {t_cpu / (N - 20) * 1e6:.1f}")
