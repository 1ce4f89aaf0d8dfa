"""A few K4 launches on the bench trajectory for ncu (python tools/k4_ncu.py [h w])."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import os
import numpy as np, torch
import mosaicking_b200 as mb
from bench import trajectory
h, w = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (1080, 1920)
dev = torch.device("cuda:0")
rng = np.random.default_rng(0)
frames = [torch.from_numpy(rng.integers(1, 256, (h, w, 3), dtype=np.uint8)).to(dev) for _ in range(4)]
mp = mb.Mapper(16384, 16384, alpha_blend=0.5, interp=os.environ.get("K4_INTERP", "linear"))
for k in range(6):
    mp.update_device(frames[k % 4], trajectory(k + 20, w, h, 16384))
torch.cuda.synchronize()
