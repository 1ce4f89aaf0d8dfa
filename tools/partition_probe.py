"""Does giving matching and compositing their own SM groups (green contexts) shorten the frame step?"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np, torch
import mosaicking_b200 as mb
from bench import trajectory
dev = torch.device("cuda:0")
rng = np.random.default_rng(0)
frames = [torch.from_numpy(rng.integers(1, 256, (1080, 1920, 3), dtype=np.uint8)).to(dev) for _ in range(8)]
descs = [torch.from_numpy(rng.integers(0, 256, (5000, 32), dtype=np.uint8)).to(dev) for _ in range(9)]
Hs = [trajectory(k, 1920, 1080, 16384) for k in range(64)]
flush = torch.empty(1 << 30, dtype=torch.uint8, device=dev)
mapper = mb.Mapper(16384, 16384, alpha_blend=0.5, interp="linear", device=dev)
main = torch.cuda.current_stream(dev)

def run(name, s_match, s_comp, steps=40):
    matcher = mb.Matcher(norm="hamming", device=dev)
    ts, tm, tc = [], [], []
    for k in range(steps + 5):
        flush.zero_()
        e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        e[0].record(main)
        s_match.wait_stream(main); s_comp.wait_stream(main)
        with torch.cuda.stream(s_match):
            matcher.knn_match_device(descs[k % 8 + 1], descs[k % 8], 0.7); e[1].record(s_match)
        with torch.cuda.stream(s_comp):
            mapper.update_device(frames[k % 8], Hs[k % 64]); e[2].record(s_comp)
        main.wait_stream(s_match); main.wait_stream(s_comp)
        e[3].record(main)
        torch.cuda.synchronize()
        if k >= 5:
            ts.append(e[0].elapsed_time(e[3]) * 1e3); tm.append(e[0].elapsed_time(e[1]) * 1e3); tc.append(e[0].elapsed_time(e[2]) * 1e3)
    print(f"{name:44s} step {np.mean(ts):6.1f} us   match done at {np.mean(tm):5.1f}   compositing done at {np.mean(tc):5.1f}", flush=True)

run("two ordinary streams", torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev))
for n in (24, 32, 40, 48, 56):
    try:
        part = mb.SmPartition(n, dev)
    except Exception as ex:
        print("partition failed:", ex); break
    run(f"green contexts {part.sms_first} + {part.sms_rest} SMs", part.first, part.rest)
    part.close()
