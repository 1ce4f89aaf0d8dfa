"""BASELINE configs[0], end to end and per phase: the reference's own offline pipeline (SequentialMosaic: extract_features ->
match_features -> registration -> global_registration -> generate, src/mosaicking/mosaic.py:304-691, examples/offline.py:123-409)
on a synthetic 1280x720 BGR video with known homography drift, ORB 2000 keypoints, run TWICE:

  reference : the UNMODIFIED reference from baseline/_ref on the host CPU (cv2.BFMatcher + Mapper._update_cpu)
  drop-in   : the same SequentialMosaic with the two edits of INTEGRATION.md section 3 applied as patches
              (registration.CompositeMatcher -> mosaicking_b200.CompositeMatcher, Mapper -> mosaicking_b200.Mapper)

and checks that the tile PNGs the two runs write are identical.  Prints one JSON object with the per-phase wall times.

    python tools/config0_pipeline.py [--frames 200] [--tile 4096]

Needs baseline/_ref (tools/install_reference.py) and, for the second run, a CUDA device."""
import argparse
import json
import os
import sys
import tempfile
import time
from pathlib import Path

import cv2
import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def make_video(path: Path, nframes: int, w=1280, h=720, seed=3):
    """World texture (multi-octave noise, no pure black) seen through a camera that drifts by (25, 8) px and 0.002 rad per frame."""
    rng = np.random.default_rng(seed)
    W, Hh = w + 25 * nframes + 400, h + 8 * nframes + 400
    world = np.zeros((Hh, W, 3), np.float32)
    for cell, amp in ((64, 1.0), (16, 0.6), (4, 0.4), (2, 0.25)):
        g = rng.random((Hh // cell + 2, W // cell + 2, 3)).astype(np.float32)
        world += amp * cv2.resize(g, (W, Hh), interpolation=cv2.INTER_CUBIC)
    world = (world - world.min()) / (world.max() - world.min())
    world = (1 + 254 * world).astype(np.uint8)
    vw = cv2.VideoWriter(str(path), cv2.VideoWriter_fourcc(*"mp4v"), 25.0, (w, h))
    assert vw.isOpened(), "cv2.VideoWriter could not open an mp4v stream"
    for k in range(nframes):
        th = 0.002 * k
        c, s = np.cos(th), np.sin(th)
        Hk = np.array([[c, -s, 200 + 25.0 * k], [s, c, 200 + 8.0 * k], [0, 0, 1.0]])      # frame -> world
        vw.write(cv2.warpPerspective(world, np.linalg.inv(Hk), (w, h), flags=cv2.INTER_LINEAR))
    vw.release()


def run(mos, video: Path, out: Path, tile: int) -> dict:
    t = {}
    m = mos.SequentialMosaic(data_path=video, project_path=out, feature_types=("ORB",), extractor_kwargs={"nFeatures": 2000},
                             bovw_clusters=5, force_cpu=True, alpha=0.5)
    for name, fn in (("extract_features", m.extract_features), ("match_features", m.match_features), ("registration", m.registration),
                     ("global_registration", m.global_registration), ("generate", lambda: m.generate((tile, tile)))):
        t0 = time.perf_counter()
        fn()
        t[name] = time.perf_counter() - t0
    t["total"] = sum(t.values())
    return t


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=200)
    ap.add_argument("--tile", type=int, default=4096)
    ap.add_argument("--skip-b200", action="store_true")
    args = ap.parse_args()
    from oracle import ref_shim
    reg, mos = ref_shim.import_reference(ROOT / "baseline" / "_ref")
    res = {"workload": f"BASELINE configs[0]: {args.frames}-frame synthetic 1280x720 video (drift 25,8 px + 0.002 rad per frame), ORB 2000, "
                       f"SequentialMosaic offline pipeline, tiles {args.tile}^2, alpha 0.5, INTER_CUBIC as shipped",
           "host_cpus": os.cpu_count(), "cv2_threads": cv2.getNumThreads(), "cv2": cv2.__version__}
    with tempfile.TemporaryDirectory() as tmp:
        tmp = Path(tmp)
        video = tmp / "synthetic.mp4"
        make_video(video, args.frames)
        (tmp / "ref").mkdir(); (tmp / "b200").mkdir()
        cv2.setRNGSeed(0)
        res["reference_cpu_s"] = run(mos, video, tmp / "ref", args.tile)
        ref_pngs = sorted((tmp / "ref").glob("*.png"))
        res["tiles"] = [p.name for p in ref_pngs]
        if not args.skip_b200:
            import mosaicking_b200 as mb
            real_matcher, real_mapper = reg.CompositeMatcher, mos.Mapper
            reg.CompositeMatcher, mos.Mapper = mb.CompositeMatcher, mb.Mapper           # INTEGRATION.md section 3, edits 2 and 3
            try:
                cv2.setRNGSeed(0)
                n0 = mb.launch_count()
                res["dropin_b200_s"] = run(mos, video, tmp / "b200", args.tile)
                res["dropin_gpu_launches"] = mb.launch_count() - n0
            finally:
                reg.CompositeMatcher, mos.Mapper = real_matcher, real_mapper
            same = []
            for p in ref_pngs:
                a, b = cv2.imread(str(p), cv2.IMREAD_UNCHANGED), cv2.imread(str(tmp / "b200" / p.name), cv2.IMREAD_UNCHANGED)
                same.append(bool(b is not None and a.shape == b.shape and np.array_equal(a, b)))
            res["tile_pngs_identical"] = same
            res["speedup_total"] = res["reference_cpu_s"]["total"] / res["dropin_b200_s"]["total"]
            res["speedup_hot_path"] = ((res["reference_cpu_s"]["match_features"] + res["reference_cpu_s"]["generate"]) /
                                       (res["dropin_b200_s"]["match_features"] + res["dropin_b200_s"]["generate"]))
    print(json.dumps(res))


if __name__ == "__main__":
    main()
