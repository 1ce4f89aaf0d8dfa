"""Isolated timings of the pre-processing kernels at 3840x2160 (row N4): python tools/preproc_time.py"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np, torch
import mosaicking_b200 as mb
from bench import Iso, preproc_record, stats_us

dev = torch.device("cuda:0")
flush = torch.empty(1024 << 20, dtype=torch.uint8, device=dev)
res = preproc_record(torch, mb, dev, Iso(torch, flush))
for k, v in res.items():
    print(f"{k:18s} {v['launch_us']:8.1f} us  {v['achieved']:7.0f} GB/s  {100 * v['frac']:5.1f} % of the HBM copy peak")
