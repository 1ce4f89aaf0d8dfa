"""How should matching and compositing of one frame be ordered on two streams?  Variants of the bench step."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np, torch
import mosaicking_b200 as mb

dev = torch.device("cuda:0")
C, fh, fw, N = 16384, 1080, 1920, 5000
rng = np.random.default_rng(0)
frames = [torch.from_numpy(rng.integers(1, 256, (fh, fw, 3), dtype=np.uint8)).to(dev) for _ in range(8)]
descs = [torch.from_numpy(rng.integers(0, 256, (N, 32), dtype=np.uint8)).to(dev) for _ in range(9)]
Hs = [np.array([[np.cos(0.002 * k), -np.sin(0.002 * k), 3000.3 + 25 * k], [np.sin(0.002 * k), np.cos(0.002 * k), 3000.7 + 8 * k], [0, 0, 1.0]]) for k in range(64)]
flush = torch.empty(1 << 30, dtype=torch.uint8, device=dev)
mapper = mb.Mapper(C, C, alpha_blend=0.5, interp="linear", device=dev)
main = torch.cuda.current_stream(dev)

def run(name, s_match, delay_cycles, k4_first=False, steps=40):
    matcher = mb.Matcher(norm="hamming", device=dev)
    ts = []
    for k in range(steps + 5):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(main)
        s_match.wait_stream(main)
        if k4_first:
            mapper.update_device(frames[k % 8], Hs[k % 64])
        with torch.cuda.stream(s_match):
            matcher.knn_match_device(descs[k % 8 + 1], descs[k % 8], 0.7)
        if not k4_first:
            if delay_cycles:
                torch.cuda._sleep(delay_cycles)
            mapper.update_device(frames[k % 8], Hs[k % 64])
        main.wait_stream(s_match)
        e1.record(main)
        torch.cuda.synchronize()
        if k >= 5:
            ts.append(e0.elapsed_time(e1) * 1e3)
    print(f"{name:50s} {np.mean(ts):6.1f} us/step (min {np.min(ts):.1f})", flush=True)

lo, hi = torch.cuda.Stream.priority_range() if hasattr(torch.cuda.Stream, "priority_range") else (0, -1)
s_norm = torch.cuda.Stream(device=dev)
s_high = torch.cuda.Stream(device=dev, priority=-1)
run("match first, same priority (bench today)", s_norm, 0)
run("K4 first, then match", s_norm, 0, k4_first=True)
run("match first, match stream high priority", s_high, 0)
for us in (2, 4, 6, 8, 12):
    run(f"match first, K4 delayed {us} us", s_norm, int(us * 1965))
    run(f"match first (high prio), K4 delayed {us} us", s_high, int(us * 1965))
