"""BASELINE configs[4] sample over N GPUs: ShardedMatcher.match_all_pairs on a keyframe set (SIFT 4000 kp each).
torchrun --nproc-per-node N tools/allpairs_sharded.py [n_images]"""
import os, sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np, torch
import torch.distributed as dist
import mosaicking_b200 as mb
from mosaicking_b200.distributed import ShardedMatcher, all_pairs_schedule

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
nimg = int(sys.argv[1]) if len(sys.argv) > 1 else 64
nkp, dim = 4000, 128
rng = np.random.default_rng(0)                      # same sets on every rank (replicated descriptors)
sets = [rng.integers(0, 120, (nkp, dim)).astype(np.float32) for _ in range(nimg)]
sets[9][:300] = sets[3][:300]
sm = ShardedMatcher(mb.Matcher(device=dev), dev, rank=rank, world=world)
sm.match_all_pairs(sets[:8])                        # warm-up (module load, NCCL init)
torch.cuda.synchronize()
if world > 1: dist.barrier()
t0 = time.perf_counter()
local_res, counts = sm.match_all_pairs(sets)          # counts only: 4 bytes per pair come back
torch.cuda.synchronize()
if world > 1: dist.barrier()
dt = time.perf_counter() - t0
handle = sm.matcher.upload_sets(sets)                 # kernel-side rate with the sets resident
mine = all_pairs_schedule(nimg, world, rank)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); sm.matcher.knn_match_sets(handle, mine, 0.7, output="device"); e1.record(); torch.cuda.synchronize()
kt = torch.tensor([e0.elapsed_time(e1)], device=dev)
if world > 1: dist.all_reduce(kt, op=dist.ReduceOp.MAX)
npairs = nimg * (nimg - 1) // 2
if rank == 0:
    print(f"N={world}: {nimg} images x {nkp} SIFT rows, {npairs} image pairs in {dt*1e3:.1f} ms wall (upload + batched tcgen05 + download + all_reduce) "
          f"-> {npairs/dt:.0f} image pairs/s, {npairs*nkp*nkp/dt:.3e} descriptor pairs/s; this rank owned {len(local_res)} pairs; "
          f"survivors (9,3): {counts[9,3]}, total edges with >=50 survivors: {int((counts >= 50).sum())}")
    print(f"      resident sets, device outputs: slowest rank {kt.item():.2f} ms -> {npairs*nkp*nkp/(kt.item()*1e-3):.3e} descriptor pairs/s over {world} GPU(s); "
          f"2000-image all-pairs (1 999 000 pairs) -> {1999000/(npairs/(kt.item()*1e-3)):.1f} s")
if world > 1: dist.destroy_process_group()
