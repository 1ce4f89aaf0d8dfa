"""Tiny driver for ncu: Hamming / L2 kNN-2 on ORB-like 5000 x 5000 x 32 B descriptors."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np, torch
import mosaicking_b200 as mb

norm = sys.argv[1] if len(sys.argv) > 1 else "hamming"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 5000
rng = np.random.default_rng(0)
if norm == "f32":
    q = torch.from_numpy(rng.integers(0, 200, (n, 128)).astype(np.float32)).cuda()
    t = torch.from_numpy(rng.integers(0, 200, (n, 128)).astype(np.float32)).cuda()
    m = mb.Matcher()
else:
    q = torch.from_numpy(rng.integers(0, 256, (n, 32), dtype=np.uint8)).cuda()
    t = torch.from_numpy(rng.integers(0, 256, (n, 32), dtype=np.uint8)).cuda()
    m = mb.Matcher(norm="hamming", engine="tensor") if norm == "hamming_tc" else mb.Matcher(norm=norm)
for _ in range(5):
    m.knn_match_device(q, t, 0.7)
torch.cuda.synchronize()
print("done", mb.launch_count())
