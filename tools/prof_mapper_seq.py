"""Driver for ncu launch lists: a natural sequence of fused updates (no artificial L2 flush)."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np, torch
import mosaicking_b200 as mb
from bench import trajectory
h, w = (2160, 3840) if (len(sys.argv) > 1 and sys.argv[1] == "4k") else (1080, 1920)
rng = np.random.default_rng(0)
mp = mb.Mapper(16384, 16384, alpha_blend=0.5, interp="linear")
N = 40
frames = [torch.from_numpy(rng.integers(1, 256, (h, w, 3), dtype=np.uint8)).cuda() for _ in range(N)]
for k in range(N):
    mp.update_device(frames[k], trajectory(k, w, h, 16384))
torch.cuda.synchronize()
print("done", mb.launch_count())
