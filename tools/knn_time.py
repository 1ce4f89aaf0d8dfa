"""kNN engine timings (isolated calls, L2 flushed, CUDA events) with a bit-exact check of the tensor-core engines against the
SIMT kernels on the same inputs: Hamming 5000x5000x32 B, L2-on-bytes 5000x5000x32, SIFT-like f32 8000x8000x128, ragged sizes."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np, torch
import mosaicking_b200 as mb

dev = torch.device("cuda:0")
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
rng = np.random.default_rng(0)


def timeit(fn, n=20, warm=3):
    for _ in range(warm): fn()
    ts = []
    for _ in range(n):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    return np.median(ts), min(ts)


def train(fn, n=20, reps=5):
    """us per call in a train of n calls (no flush inside; a few memsets first so the host gets ahead of the GPU)"""
    out = []
    for _ in range(reps):
        for _ in range(4): flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(n): fn()
        b.record(); torch.cuda.synchronize()
        out.append(a.elapsed_time(b) * 1e3 / n)
    return np.median(out)


def same(a, b):
    return all(torch.equal(x, y) for x, y in zip(a, b))


for nq, nt in ((5000, 5000), (4999, 5003), (777, 9001)):
    q = torch.from_numpy(rng.integers(0, 256, (nq, 32), dtype=np.uint8)).to(dev)
    t = torch.from_numpy(rng.integers(0, 256, (nt, 32), dtype=np.uint8)).to(dev)
    t[5] = q[7]; t[100] = q[7]
    ref = mb.Matcher(norm="hamming", engine="popc").knn_match_device(q, t, 0.7)
    m = mb.Matcher(norm="hamming", engine="tensor")
    ok = same(ref, m.knn_match_device(q, t, 0.7))
    med, best = timeit(lambda: m.knn_match_device(q, t, 0.7))
    print(f"hamming tcgen05 {nq}x{nt}: median {med:.1f} best {best:.1f} us, train {train(lambda: m.knn_match_device(q, t, 0.7)):.2f} us/call  identical to XOR+POPC: {ok}")
    m2 = mb.Matcher(norm="l2")
    med, best = timeit(lambda: m2.knn_match_device(q, t, 0.7))
    print(f"l2-u8 tcgen05 {nq}x{nt}: median {med:.1f} best {best:.1f} us, train {train(lambda: m2.knn_match_device(q, t, 0.7)):.2f} us/call")
for n, nt in ((8000, 8000), (4000, 4000), (8000, 7777)):
    q = torch.from_numpy(rng.integers(0, 200, (n, 128)).astype(np.float32)).to(dev)
    t = torch.from_numpy(rng.integers(0, 200, (nt, 128)).astype(np.float32)).to(dev)
    m = mb.Matcher()
    out = m.knn_match_device(q, t, 0.7)
    # SIMT reference: a non-integer value in row 0 sends the call to the bit-exact fallback; compare all other rows
    q2 = q.clone(); q2[0, 0] += 0.5
    ref = m.knn_match_device(q2, t, 0.7)
    ok = all(torch.equal(x[1:], y[1:]) for x, y in zip(out, ref))
    med, best = timeit(lambda: m.knn_match_device(q, t, 0.7), n=10, warm=2)
    tr = train(lambda: m.knn_match_device(q, t, 0.7), n=10)
    print(f"l2 f32 tcgen05 {n}x{nt}x128: median {med:.1f} best {best:.1f} us, train {tr:.2f} us/call -> {2.0*n*nt*128/tr/1e6:.0f} TFLOP/s  identical to SIMT: {ok}")
