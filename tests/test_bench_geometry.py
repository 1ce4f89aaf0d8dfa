"""Host-side check of the N > 1 bench workload (bench.lane_geometry / lane_pose): frames are routed at every world size, and the
routing is balanced -- every rank receives exactly one neighbouring lane's straddling frames (reference routing rule:
src/mosaicking/mosaic.py:595-622 through distributed.tiles_of_footprint)."""
import sys
from pathlib import Path

import pytest

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import bench  # noqa: E402
from mosaicking_b200.distributed import tiles_of_footprint  # noqa: E402


@pytest.mark.parametrize("world", [2, 4, 8])
@pytest.mark.parametrize("frame", [(1920, 1080), (3840, 2160)])
def test_lanes_route_one_neighbour_to_every_rank(world, frame):
    fw, fh = frame
    gx = bench.BIG_GRID[0]
    owner = lambda t: (t[1] * gx + t[0]) % world  # noqa: E731  (ShardedMosaic.owner)
    received = {r: set() for r in range(world)}
    for lane in range(world):
        tx, ty, side = bench.lane_geometry(lane)
        assert owner((tx, ty)) == lane and owner((tx + side, ty)) != lane
        for k in range(25):
            tiles = tiles_of_footprint(bench.lane_pose(k, lane, fw, fh), fw, fh, (bench.CANVAS, bench.CANVAS), bench.BIG_GRID)
            owners = sorted(owner(t) for t in tiles)
            if k % 2 == 0:      # centred on the seam: its own tile and the neighbour's
                assert sorted(tiles) == sorted([(tx, ty), (tx + side, ty)])
                received[owner((tx + side, ty))].add(lane)
            else:               # inside its own tile: nothing is routed
                assert owners == [lane]
    assert all(len(v) == 1 for v in received.values()), received
