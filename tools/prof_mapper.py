"""Tiny driver for ncu: a few fused warp+blend updates (1080p frame, 16k x 16k canvas)."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np, torch
import mosaicking_b200 as mb

blend = sys.argv[1] if len(sys.argv) > 1 else "alpha"
h, w = (2160, 3840) if (len(sys.argv) > 2 and sys.argv[2] == "4k") else (1080, 1920)
rng = np.random.default_rng(0)
mp = mb.Mapper(16384, 16384, alpha_blend=0.5, interp="linear", blend=blend)
frame = torch.from_numpy(rng.integers(1, 256, (h, w, 3), dtype=np.uint8)).cuda()
flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
for k in range(6):
    th = 0.05 * k
    H = np.array([[np.cos(th), -np.sin(th), 4000.3 + 100 * k], [np.sin(th), np.cos(th), 3000.7 + 40 * k], [0, 0, 1.0]])
    flush.zero_()
    mp.update_device(frame, H)
torch.cuda.synchronize()
print("done", mb.launch_count())
