"""N-GPU numerical check of ShardedMosaic: planned steps, real Mapper kernels, gather -> compare bit for bit with rendering every
tile sequentially on one GPU (reference-tile semantics).  Two transports: "nccl" (frames sent with ncclSend/Recv, the exchange of
step k+1 posted ahead), "peer" (no copy: the compositing kernel reads the other GPU's frame over NVLink, distributed.PeerFrames)
and "peer_copy" (the frame is pulled whole out of the other GPU's memory by the copy engine, one step ahead).
torchrun --nproc-per-node N tools/sharded_check.py"""
import os, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np, torch
import torch.distributed as dist
import mosaicking_b200 as mb
from mosaicking_b200.distributed import PeerFrames, ShardedMosaic, tiles_of_footprint

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
TILE, GRID, FRAME, STEPS = (1024, 768), (4, 2), (360, 480), 12
rng = np.random.default_rng(4)                                   # same scenario on every rank
frames = rng.integers(1, 256, (STEPS, world, FRAME[0], FRAME[1], 3), dtype=np.uint8)
Hs = np.zeros((STEPS, world, 3, 3))
for k in range(STEPS):
    for r in range(world):
        th = 0.05 * k + 0.3 * r
        Hs[k, r] = [[np.cos(th), -np.sin(th), 300 + 230 * k + 170 * r], [np.sin(th), np.cos(th), 150 + 70 * k + 90 * r], [1e-6, 0, 1]]
factory = lambda w, h: mb.Mapper(w, h, alpha_blend=0.5, interp="linear", device=dev)
fr = [torch.from_numpy(frames[k, rank]).to(dev) for k in range(STEPS)]
want = None
for mode in ("nccl", "peer", "peer_copy"):
    sm = ShardedMosaic(GRID, TILE, factory, FRAME, dev)
    sm.connect(); sm.plan(Hs[:, rank])
    peers = PeerFrames(fr, dev) if mode != "nccl" else None
    for k in range(STEPS):
        if peers is None:
            sm.step_planned(k, fr[k], fr[k + 1] if k + 1 < STEPS else None)
        else:
            sm.step_planned(k, None, peers=peers, peer_copy=(mode == "peer_copy"))
    torch.cuda.synchronize()
    if peers is not None:
        peers.close()
    got = sm.gather(0)
    if rank == 0:
        if want is None:
            tiles = {}
            for k in range(STEPS):
                for r in range(world):
                    for (xi, yi) in tiles_of_footprint(Hs[k, r], FRAME[1], FRAME[0], TILE, GRID):
                        m = tiles.setdefault((xi, yi), factory(*TILE))
                        m.update_device(torch.from_numpy(frames[k, r]).to(dev), mb.homogeneous_translation(-xi * TILE[0], -yi * TILE[1]) @ Hs[k, r])
            out = {t: m.output for t, m in tiles.items()}
            out.setdefault((GRID[0] - 1, GRID[1] - 1), np.zeros((TILE[1], TILE[0], 3), np.uint8))
            want = mb.combine_tiles(out, TILE)
        same = np.array_equal(got, want)
        print(f"N={world} {mode}: {sm.bytes_sent} frame bytes routed by rank 0, gathered mosaic == sequential tiles: {same}; nonzero px {int((want.sum(2) > 0).sum())}")
        assert same
    dist.barrier()
dist.destroy_process_group()
