"""How should one frame step (kNN match + fused update) be put on the GPU?  Same step, same events and L2 flush as bench.py,
different stream arrangements: concurrent (bench default), matcher stream with high priority, serial in either order."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np, torch
import mosaicking_b200 as mb
from bench import trajectory, make_descriptors, frames_along, CANVAS, RATIO, ALPHA

dev = torch.device("cuda:0")
K, W = 40, 5
fw, fh = 1920, 1080
Hs = [trajectory(k, fw, fh, CANVAS) for k in range(K + W)]
frames = frames_along(torch, mb, dev, Hs, fw, fh, 7, 0.0)
desc_h = make_descriptors(K + W + 1, 11)
desc_d = [torch.from_numpy(d).to(dev) for d in desc_h]
flush = torch.empty(1024 << 20, dtype=torch.uint8, device=dev)
matcher = mb.Matcher(norm="hamming", device=dev)
mapper = mb.Mapper(CANVAS, CANVAS, alpha_blend=ALPHA, interp="linear", device=dev)
main = torch.cuda.current_stream(dev)
lo, hi = torch.cuda.Stream.priority_range() if hasattr(torch.cuda.Stream, "priority_range") else (0, -1)


def run(name, step):
    for k in range(W): step(k)
    torch.cuda.synchronize()
    ts = []
    for k in range(K):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(main); step(W + k); b.record(main); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    print(f"{name:46s}: median {np.median(ts):.1f} us  min {min(ts):.1f}  max {max(ts):.1f}")


def concurrent(s_match):
    def step(k):
        s_match.wait_stream(main)
        with torch.cuda.stream(s_match):
            matcher.knn_match_device(desc_d[k + 1], desc_d[k], RATIO)
        mapper.update_device(frames[k], Hs[k])
        main.wait_stream(s_match)
    return step


def concurrent_k4_first(s_match):
    def step(k):
        s_match.wait_stream(main)
        mapper.update_device(frames[k], Hs[k])
        with torch.cuda.stream(s_match):
            matcher.knn_match_device(desc_d[k + 1], desc_d[k], RATIO)
        main.wait_stream(s_match)
    return step


def serial_match_first(k):
    matcher.knn_match_device(desc_d[k + 1], desc_d[k], RATIO)
    mapper.update_device(frames[k], Hs[k])


def serial_k4_first(k):
    mapper.update_device(frames[k], Hs[k])
    matcher.knn_match_device(desc_d[k + 1], desc_d[k], RATIO)


def k4_on_side(s_comp):
    def step(k):
        s_comp.wait_stream(main)
        matcher.knn_match_device(desc_d[k + 1], desc_d[k], RATIO)
        with torch.cuda.stream(s_comp):
            mapper.update_device(frames[k], Hs[k])
        main.wait_stream(s_comp)
    return step


for rep in range(2):
    run("concurrent, match issued first (bench)", concurrent(torch.cuda.Stream(device=dev)))
    run("concurrent, matcher stream high priority", concurrent(torch.cuda.Stream(device=dev, priority=-1)))
    run("concurrent, K4 issued first", concurrent_k4_first(torch.cuda.Stream(device=dev)))
    run("concurrent, K4 first, matcher high priority", concurrent_k4_first(torch.cuda.Stream(device=dev, priority=-1)))
    run("concurrent, K4 on a high-priority side stream", k4_on_side(torch.cuda.Stream(device=dev, priority=-1)))
    run("serial, match then K4", serial_match_first)
    run("serial, K4 then match", serial_k4_first)
