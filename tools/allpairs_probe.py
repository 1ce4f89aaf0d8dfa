"""BASELINE configs[4] sample: all-pairs matching over a keyframe set (SIFT 4000 kp each), one GPU, batched tcgen05 path."""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np, torch
import mosaicking_b200 as mb
nimg = int(sys.argv[1]) if len(sys.argv) > 1 else 64
nkp, dim = 4000, 128
rng = np.random.default_rng(0)
dev = torch.device("cuda:0")
desc = torch.from_numpy(rng.integers(0, 120, (nimg * nkp, dim)).astype(np.float32)).to(dev)
set_off = torch.arange(0, (nimg + 1) * nkp, nkp, dtype=torch.int32, device=dev)
pr = np.array([(a, b) for a in range(nimg) for b in range(a)], np.int32)
pairs = torch.from_numpy(pr).to(dev)
out_off = torch.arange(0, (len(pr) + 1) * nkp, nkp, dtype=torch.int32, device=dev)
rows = len(pr) * nkp
idx = torch.empty((rows, 2), dtype=torch.int32, device=dev); dist = torch.empty((rows, 2), dtype=torch.float32, device=dev)
keep = torch.empty((rows,), dtype=torch.uint8, device=dev); nk = torch.zeros((len(pr),), dtype=torch.int32, device=dev)
L = mb._lib.lib()
ws = torch.empty(L.mk_knn2_batched_workspace_bytes(2, len(pr), nkp, nkp, dim, nimg * nkp), dtype=torch.uint8, device=dev)
def run():
    mb._lib.check(L.mk_knn2_batched(2, desc.data_ptr(), nimg * nkp, dim, dim * 4, set_off.data_ptr(), pairs.data_ptr(), len(pr),
                                    out_off.data_ptr(), nkp, nkp, idx.data_ptr(), dist.data_ptr(), 0.7, keep.data_ptr(), nk.data_ptr(),
                                    ws.data_ptr(), ws.numel(), torch.cuda.current_stream().cuda_stream))
run(); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); run(); e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
npairs_desc = len(pr) * nkp * nkp
print(f"{nimg} images x {nkp} kp: {len(pr)} image pairs in {ms:.2f} ms -> {len(pr)/ms*1e3:.0f} image pairs/s, {npairs_desc/ms*1e3:.3e} descriptor pairs/s, "
      f"{2*npairs_desc*dim/ms*1e3/1e12:.0f} TFLOP/s; 2000-image all-pairs (1 999 000 pairs) would take {1999000/(len(pr)/ms*1e3):.1f} s on one GPU")
print("kept per pair (first 5):", nk[:5].tolist())
