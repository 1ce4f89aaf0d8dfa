import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np, torch
import mosaicking_b200 as mb
from oracle import oracle as orc
rng = np.random.default_rng(0)
Wc, Hc = 640, 480
mp = mb.Mapper(Wc, Hc, alpha_blend=0.5, interp="linear")
canvas = np.zeros((Hc, Wc, 3), np.uint8); cmask = np.zeros((Hc, Wc), np.uint8)
img = rng.integers(1, 256, (180, 240, 3), dtype=np.uint8)
H = np.array([[1.0, 0, 150.0], [0, 1.0, 90.0], [0, 0.0, 1.0]])
mp.update(img, H)
torch.cuda.synchronize()
orc.mapper_update(canvas, cmask, img, H, 0.5, orc.INTER_LINEAR)
out, msk = mp.output, mp.mask
print("equal", np.array_equal(out, canvas), np.array_equal(msk, cmask))
d = np.argwhere((out != canvas).any(axis=2)); print(len(d), d[:5].tolist(), d[-3:].tolist() if len(d) else None)
