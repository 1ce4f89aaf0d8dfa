"""world_size-2 gloo test of the tile-sharded mosaic (host logic + routing + P2P frame delivery).
The per-tile compositor is an oracle-backed CPU stand-in, so this runs without a GPU; the result
must equal rendering every tile sequentially the way SequentialMosaic.generate does
(src/mosaicking/mosaic.py:665-691: one Mapper per tile, H_tile = T(-tx,-ty) @ H)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import mosaicking_b200 as mb
from mosaicking_b200.distributed import ShardedMatcher, ShardedMosaic, all_pairs_schedule, tiles_of_footprint
from oracle import oracle as orc

TILE = (128, 96)
GRID = (3, 2)
FRAME = (60, 80)   # h, w
STEPS = 5


class OracleMapper:
    """CPU stand-in with Mapper's device-side interface (update_device / output)."""

    def __init__(self, w, h):
        self.canvas = np.zeros((h, w, 3), np.uint8)
        self.cmask = np.zeros((h, w), np.uint8)

    def update_device(self, frame, H):
        orc.mapper_update(self.canvas, self.cmask, frame.numpy(), H, 0.5, orc.INTER_LINEAR)

    @property
    def output(self):
        return self.canvas


def _scenario(world):
    rng = np.random.default_rng(4)
    frames = rng.integers(1, 256, (STEPS, world, FRAME[0], FRAME[1], 3), dtype=np.uint8)
    Hs = np.zeros((STEPS, world, 3, 3))
    for k in range(STEPS):
        for r in range(world):
            th = 0.1 * k + 0.3 * r
            Hs[k, r] = [[np.cos(th), -np.sin(th), 70 + 55 * k + 20 * r], [np.sin(th), np.cos(th), 40 + 18 * k + 60 * r], [0, 0, 1]]
    return frames, Hs


def _expected(world):
    frames, Hs = _scenario(world)
    tiles = {}
    for k in range(STEPS):
        for r in range(world):
            for (xi, yi) in tiles_of_footprint(Hs[k, r], FRAME[1], FRAME[0], TILE, GRID):
                m = tiles.setdefault((xi, yi), OracleMapper(*TILE))
                m.update_device(torch.from_numpy(frames[k, r]), mb.homogeneous_translation(-xi * TILE[0], -yi * TILE[1]) @ Hs[k, r])
    out = {t: m.output for t, m in tiles.items()}
    out.setdefault((GRID[0] - 1, GRID[1] - 1), np.zeros((TILE[1], TILE[0], 3), np.uint8))
    return mb.combine_tiles(out, TILE), set(tiles)


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        frames, Hs = _scenario(world)
        sm = ShardedMosaic(GRID, TILE, OracleMapper, FRAME, torch.device("cpu"))
        sm.connect()
        sm.plan(Hs[:, rank])                 # one all_gather of every pose (the bench path) ...
        fr = [torch.from_numpy(frames[k, rank]) for k in range(STEPS)]
        for k in range(STEPS):
            if k in (0, 3):                  # planned step that also posts the exchange of step k+1 ahead of time
                sm.step_planned(k, fr[k], next_frame=fr[k + 1])
            elif k in (1, 4):                # ... consumed here
                sm.step_planned(k, fr[k])
            else:                            # ... and the per-step all_gather path
                sm.step(fr[k], Hs[k, rank])
        assert not sm._posted
        assert all(sm.owner(t) == rank for t in sm.mappers)
        res = sm.gather(0)
        if rank == 0:
            q.put((res, sm.bytes_sent))
    finally:
        dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.timeout(180)
def test_sharded_mosaic_two_ranks_matches_sequential_tiles():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, _free_port_cached(), q)) for r in range(world)]
    for p in procs:
        p.start()
    got, sent = q.get(timeout=150)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    want, touched = _expected(world)
    assert len(touched) >= 4            # the scenario really crosses tile seams
    assert sent > 0                     # ... and frames really travelled between ranks
    assert np.array_equal(got, want)


_PORT = None


def _free_port_cached():
    global _PORT
    if _PORT is None:
        _PORT = _free_port()
    return _PORT


def test_single_rank_is_plain_tiling():
    frames, Hs = _scenario(1)
    sm = ShardedMosaic(GRID, TILE, OracleMapper, FRAME, torch.device("cpu"), rank=0, world=1)
    for k in range(STEPS):
        sm.step(torch.from_numpy(frames[k, 0]), Hs[k, 0])
    want, _ = _expected(1)
    assert np.array_equal(sm.gather(0), want)


def test_tiles_of_footprint_equals_full_scan():
    rng = np.random.default_rng(8)
    for _ in range(40):
        th = rng.uniform(-3, 3)
        H = np.array([[np.cos(th), -np.sin(th), rng.uniform(-50, 400)], [np.sin(th), np.cos(th), rng.uniform(-50, 250)], [0, 0, 1.0]])
        crn = mb.get_corners(H, FRAME[1], FRAME[0]).squeeze().astype(int)
        full = []
        for xi in range(GRID[0]):
            for yi in range(GRID[1]):
                tx, ty = xi * TILE[0], yi * TILE[1]
                tc = np.array([[tx, ty], [tx + TILE[0], ty], [tx + TILE[0], ty + TILE[1]], [tx, ty + TILE[1]]])
                if mb.bbox_overlap(tc, crn):     # mosaic.py:612-614 over every tile
                    full.append((xi, yi))
        assert sorted(tiles_of_footprint(H, FRAME[1], FRAME[0], TILE, GRID)) == sorted(full)


# ------------------------------------------------------------------------------------------------
# all-pairs matching sharded over ranks (every (query image, train image) pair is independent)
# ------------------------------------------------------------------------------------------------
class OracleSetMatcher:
    """CPU stand-in with Matcher.knn_match_sets' interface."""

    def knn_match_sets(self, sets, pairs, ratio, output="host"):
        out = []
        for a, b in pairs:
            idx, dist_ = orc.knn2(sets[a], sets[b], "hamming")
            out.append((idx, dist_, orc.ratio_test(idx, dist_, ratio)))
        if output == "counts":
            return np.array([np.count_nonzero(k) for _, _, k in out], np.int32)
        return out


def _desc_sets():
    rng = np.random.default_rng(21)
    sets = [rng.integers(0, 256, (40 + 7 * k, 32), dtype=np.uint8) for k in range(9)]
    sets[5][:20] = sets[2][:20]            # two images that really match
    sets[7][3:30] = sets[2][10:37]
    return sets


def _match_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        sets = _desc_sets()
        sm = ShardedMatcher(OracleSetMatcher(), torch.device("cpu"))
        local, counts = sm.match_all_pairs(sets, 0.7, chunk=5, keep_results=True)
        owned, counts2 = sm.match_all_pairs(sets, 0.7, chunk=7)          # counts-only form: same matrix, same ownership
        assert sorted(owned) == sorted(local) and np.array_equal(counts, counts2)
        q.put((rank, sorted(local), counts))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_sharded_all_pairs_matching_two_ranks():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_match_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = [q.get(timeout=150) for _ in range(world)]
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    sets = _desc_sets()
    n = len(sets)
    want = np.zeros((n, n), np.int64)
    for i in range(n):
        for j in range(i + 1, n):
            idx, dist_ = orc.knn2(sets[j], sets[i], "hamming")
            want[j, i] = np.count_nonzero(orc.ratio_test(idx, dist_, 0.7))
    owned = [set(map(tuple, pairs)) for _, pairs, _ in sorted(got)]
    assert not (owned[0] & owned[1]) and (owned[0] | owned[1]) == {(j, i) for i in range(n) for j in range(i + 1, n)}
    assert min(len(o) for o in owned) > 0
    for _, _, counts in got:                 # every rank ends up with the whole survivor-count matrix
        assert np.array_equal(counts, want)
    assert want[5, 2] >= 10 and want[7, 2] >= 10


def test_all_pairs_schedule_is_a_balanced_partition():
    for n, world in ((37, 8), (300, 8), (5, 2), (1, 2), (64, 3)):
        seen, loads = set(), []
        for r in range(world):
            p = all_pairs_schedule(n, world, r)
            assert not (seen & set(p))
            seen |= set(p); loads.append(len(p))
        assert seen == {(j, i) for i in range(n) for j in range(i + 1, n)}
        if n >= 4 * world:
            assert max(loads) - min(loads) <= 64
