"""Where do the ~30 us per frame between the e2e loop (bench.e2e_single) and the copies-only ceiling go?
Same objects as the bench (Mapper(async_upload=True), pinned_empty buffers, knn_match_arrays_async); prints per-call host
times, the time the host spends blocked in result(), and the loop with parts removed."""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np, torch
import mosaicking_b200 as mb

dev = torch.device("cuda:0")
rng = np.random.default_rng(0)
fh, fw, CANVAS = 1080, 1920, 16384
N = int(sys.argv[1]) if len(sys.argv) > 1 else 400
OPTS = sys.argv[2:]
SEQ = "seq" in OPTS
SWAP = "swap" in OPTS                       # frame first, then the matching call
SPIN = next((float(o[5:]) for o in OPTS if o.startswith("spin=")), 0.0) * 1e-6     # busy-wait between the two calls (a slower host)
def spin(dt):
    t = time.perf_counter() + dt
    while time.perf_counter() < t:
        pass
pin_frames = [mb.pinned_empty((fh, fw, 3)) for _ in range(4)]
for f in pin_frames:
    f[...] = rng.integers(1, 256, (fh, fw, 3), dtype=np.uint8)
pin_desc = [mb.pinned_empty((5000, 32)) for _ in range(8)]
for d in pin_desc:
    d[...] = rng.integers(0, 256, (5000, 32), dtype=np.uint8)
Hs = []
for k in range(64):
    th = 0.002 * k
    c, s = np.cos(th), np.sin(th)
    Hs.append(np.array([[c, -s, 6000.3 + 48 * k], [s, c, 6000.7 + 16 * k], [0, 0, 1.0]]))
matcher = mb.Matcher(norm="hamming", device=dev)
amapper = mb.Mapper(CANVAS, CANVAS, alpha_blend=0.5, interp="linear", device=dev, async_upload=True)


def inputs(k):
    return pin_frames[k % 4], pin_desc[(k + 1) % 8], pin_desc[k % 8]


def warm_link():
    wd = torch.empty((fh, fw, 3), dtype=torch.uint8, device=dev)
    tf = [torch.from_numpy(f) for f in pin_frames]
    t_begin = time.perf_counter()
    while time.perf_counter() - t_begin < 3.0:
        for k in range(200):
            wd.copy_(tf[k % 4], non_blocking=True)
        torch.cuda.synchronize()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for k in range(200):
        wd.copy_(tf[k % 4], non_blocking=True)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 200
    print(f"plain frame copies: {dt * 1e6:.1f} us / frame = {fh * fw * 3 / dt / 1e9:.1f} GB/s", flush=True)


def run(mode, n=N, detail=False):
    """mode: full | noresult (results collected at the end) | mapper (frames only) | matcher (matching only)"""
    acc = {"submit": 0.0, "update": 0.0, "result": 0.0}
    pend_all = []
    pending = None
    torch.cuda.synchronize()
    t_start = time.perf_counter()
    for k in range(n):
        img, q, t = inputs(k)
        t0 = time.perf_counter()
        if SWAP and mode != "matcher":
            amapper.update(img, Hs[k % len(Hs)])
        nxt = (matcher.match_next_async(q, 0.7) if SEQ else matcher.knn_match_arrays_async(q, t, 0.7)) if mode != "mapper" else None
        t1 = time.perf_counter()
        if SPIN:
            spin(SPIN)
        if not SWAP and mode != "matcher":
            amapper.update(img, Hs[k % len(Hs)])
        t2 = time.perf_counter()
        if mode in ("full", "matcher") and pending is not None:
            int(pending.result()[2].sum())
        t3 = time.perf_counter()
        pending = nxt
        if mode == "noresult":
            pend_all.append(nxt)
            if len(pend_all) >= 3:
                pend_all.pop(0).result()
        acc["submit"] += t1 - t0; acc["update"] += t2 - t1; acc["result"] += t3 - t2
    if pending is not None and mode in ("full", "matcher"):
        pending.result()
    for p in pend_all:
        p.result()
    torch.cuda.synchronize()
    total = (time.perf_counter() - t_start) / n * 1e6
    print(f"{mode:10s}: {total:7.1f} us/frame = {1e6 / total:7.0f} frames/s   host: submit {acc['submit'] / n * 1e6:5.1f}  update {acc['update'] / n * 1e6:5.1f}"
          f"  result {acc['result'] / n * 1e6:5.1f} us", flush=True)
    return total


warm_link()
for _ in range(6):
    run("full", 64)
print("---")
for rep in range(3):
    for mode in ("full", "noresult"):
        run(mode)
if "prof" not in OPTS:
    sys.exit(0)
# host cost alone: the same calls with nothing to wait for (tiny queue)
import cProfile, pstats
pr = cProfile.Profile(); pr.enable(); run("full", 300); pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(16)
