import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np, torch
import mosaicking_b200 as mb
dev = torch.device("cuda:0")
rng = np.random.default_rng(0)
mp = mb.Mapper(16384, 16384, alpha_blend=0.5, interp="linear")
H = np.array([[np.cos(0.05), -np.sin(0.05), 4000.3], [np.sin(0.05), np.cos(0.05), 3000.7], [0, 0, 1.0]])
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
for (h, w) in ((128, 256), (1080, 1920)):
    fr = torch.from_numpy(rng.integers(1, 256, (h, w, 3), dtype=np.uint8)).to(dev)
    for _ in range(3):
        flush.zero_(); mp.update_device(fr, H); torch.cuda.synchronize()
    print("---", w, h, flush=True)
