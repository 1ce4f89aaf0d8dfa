"""ncu driver: Hamming tensor engine with explicit nq, nt."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np, torch
import mosaicking_b200 as mb
nq, nt = int(sys.argv[1]), int(sys.argv[2])
rng = np.random.default_rng(0)
q = torch.from_numpy(rng.integers(0, 256, (nq, 32), dtype=np.uint8)).cuda()
t = torch.from_numpy(rng.integers(0, 256, (nt, 32), dtype=np.uint8)).cuda()
m = mb.Matcher(norm="hamming", engine="tensor")
for _ in range(4):
    m.knn_match_device(q, t, 0.7)
torch.cuda.synchronize()
