"""CPU-side checks of the boundary: the C-ABI library loads and exports every symbol that
include/mosaic_b200.h declares (no compute calls), argument validation happens before any CUDA
work, and the host-side geometry mirrors the reference (golden routing fixture)."""
import ctypes
import re
from pathlib import Path

import numpy as np
import pytest

import mosaicking_b200 as mb

ROOT = Path(__file__).resolve().parent.parent


def test_header_symbols_are_exported():
    header = (ROOT / "include" / "mosaic_b200.h").read_text()
    declared = set(re.findall(r"\b(mk_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations found"
    L = ctypes.CDLL(str(mb._lib.LIB_PATH))
    for name in declared:
        assert hasattr(L, name), f"{name} declared in mosaic_b200.h but not exported"
    assert declared == set(mb._lib.EXPORTS)


def test_argument_validation_without_gpu():
    L = mb._lib.lib()
    assert L.mk_abi_version() == 1
    H = np.eye(3)
    # null canvas -> MK_ERR_ARG before touching CUDA
    rc = L.mk_mapper_update(None, 0, None, 0, None, 0, 10, 10, None, 4, 4, 12, None, 0, H.ctypes.data, 0.5, 1, 0, 0.0, None, None)
    assert rc == -1 and b"null pointer" in L.mk_last_error()
    rc = L.mk_knn2(0, None, 1, None, 1, 32, 24, None, None, 0.7, None, None, None, 0, None)
    assert rc == -2  # stride not a multiple of 16
    rc = L.mk_knn2(7, None, 1, None, 1, 32, 32, None, None, 0.7, None, None, None, 0, None)
    assert rc == -1
    assert L.mk_knn2_workspace_bytes(0, 5000, 5000, 32) > 0


def test_footprint_is_conservative():
    L = mb._lib.lib()
    rng = np.random.default_rng(0)
    reg = (ctypes.c_int32 * 4)()
    for _ in range(50):
        th = rng.uniform(-3, 3)
        H = np.array([[np.cos(th), -np.sin(th), rng.uniform(-200, 900)], [np.sin(th), np.cos(th), rng.uniform(-200, 700)],
                      [rng.uniform(-1e-4, 1e-4), rng.uniform(-1e-4, 1e-4), 1.0]])
        assert L.mk_mapper_footprint(600, 800, 120, 160, H.ctypes.data, 1, reg) == 0
        x0, y0, x1, y1 = reg
        c = mb.get_corners(H, 160, 120).reshape(-1, 2)
        for cx, cy in c:
            if 0 <= cx < 800 and 0 <= cy < 600:
                assert x0 <= cx < x1 and y0 <= cy < y1


def test_no_gpu_means_loud_failure():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(mb.MosaicB200Error):
        mb.Mapper(64, 64)
    with pytest.raises(mb.MosaicB200Error):
        mb.Matcher().knn_match(np.zeros((3, 32), np.uint8), np.zeros((3, 32), np.uint8))


def test_routing_matches_reference_golden(golden):
    g = golden.routing
    Hs, (w, h) = g["H"], g["wh"]
    assert np.allclose(mb.get_corners(Hs, int(w), int(h)), g["corners"], rtol=0, atol=0)
    assert np.array_equal(mb.get_corners(Hs[0], int(w), int(h)), g["corners_single"])
    assert np.array_equal(np.array(mb.get_mosaic_dimensions(Hs, int(w), int(h)), np.float64), g["dims"])
    assert np.array_equal(mb.find_nice_homographies(Hs), g["nice"])
    tile = tuple(int(v) for v in g["tile"])
    ov = g["overlap"]
    for k, H in enumerate(Hs):
        crn = mb.get_corners(H, int(w), int(h)).squeeze().astype(int)
        for i in range(ov.shape[1]):
            for j in range(ov.shape[2]):
                tx, ty = i * tile[0], j * tile[1]
                tc = np.array([[tx, ty], [tx + tile[0], ty], [tx + tile[0], ty + tile[1]], [tx, ty + tile[1]]])
                assert mb.bbox_overlap(tc, crn) == bool(ov[k, i, j]), (k, i, j)
    tiles = mb.mosaic.calculate_tiles((0, 0, 6500, 4700), tile)
    assert np.array_equal(np.array(tiles), g["split_tiles"])
    for k, H in enumerate(Hs):
        crn = mb.get_corners(H, int(w), int(h)).reshape(-1, 2)
        got = np.zeros(len(tiles), bool)
        got[list(mb.mosaic.assign_image_to_tiles(crn, tiles))] = True
        assert np.array_equal(got, g["split_assign"][k])


def test_matcher_pickle_drops_handles():
    import pickle
    m = mb.CompositeMatcher(norm="hamming")
    m2 = pickle.loads(pickle.dumps(m))
    assert m2._norm == "hamming" and m2._matcher["workspace"] is None
