"""Import shim for the real reference (mosaic-library) -- TEST INFRASTRUCTURE, build container only.

/root/reference exists only in the build container, never on the GPU box.  This module is
used by oracle/make_golden.py (fixture generation) and by tests that pin the oracle against
the reference's own classes when the checkout is present; nothing on the GPU run path
imports it.

The unmodified reference does not import on stock opencv-python: it needs `shapely`
(registration.py:15) and evaluates cv2.cuda.SURF_CUDA / cv2.xfeatures2d in def-time
annotations (registration.py:56,121).  The shim stubs exactly those three names.
"""
from __future__ import annotations

import os
import sys
import types
from pathlib import Path

REFERENCE_SRC = Path(os.environ.get("MOSAIC_REFERENCE", "/root/reference")) / "src"


def available() -> bool:
    return (REFERENCE_SRC / "mosaicking" / "mosaic.py").exists()


def _install_stubs() -> None:
    import cv2
    import numpy as np

    if "shapely" not in sys.modules:
        try:
            import shapely  # noqa: F401
        except ImportError:
            shapely = types.ModuleType("shapely")
            geometry = types.ModuleType("shapely.geometry")

            class Polygon:  # only use: registration.py:375-380 (convex quads)
                def __init__(self, pts):
                    self.pts = np.asarray(pts, np.float32).reshape(-1, 2)

                def intersects(self, other) -> bool:
                    a = cv2.convexHull(self.pts)
                    b = cv2.convexHull(other.pts)
                    area, _ = cv2.intersectConvexConvex(a, b)
                    if area > 0:
                        return True
                    for p in a.reshape(-1, 2):
                        if cv2.pointPolygonTest(b, (float(p[0]), float(p[1])), False) >= 0:
                            return True
                    for p in b.reshape(-1, 2):
                        if cv2.pointPolygonTest(a, (float(p[0]), float(p[1])), False) >= 0:
                            return True
                    return False

            geometry.Polygon = Polygon
            shapely.geometry = geometry
            sys.modules["shapely"] = shapely
            sys.modules["shapely.geometry"] = geometry
    if not hasattr(cv2.cuda, "SURF_CUDA"):
        cv2.cuda.SURF_CUDA = type("SURF_CUDA", (), {})
    if not hasattr(cv2, "xfeatures2d"):
        cv2.xfeatures2d = types.SimpleNamespace(SURF=type("SURF", (), {}))


def import_reference(src: Path | None = None):
    """-> (mosaicking.registration, mosaicking.mosaic) of the real reference, CPU mode.  `src` = directory that holds the
    `mosaicking` package: the checkout's src/ (default, build container) or baseline/_ref (the pip-installed, unmodified
    copy that travels to the GPU box for bench.py --impl reference)."""
    src = REFERENCE_SRC if src is None else Path(src)
    if not (src / "mosaicking" / "mosaic.py").exists():
        raise ImportError(f"reference not found under {src}")
    _install_stubs()
    if str(src) not in sys.path:
        sys.path.insert(0, str(src))
    import mosaicking
    import mosaicking.mosaic
    import mosaicking.registration

    assert not mosaicking.HAS_CUDA
    return mosaicking.registration, mosaicking.mosaic
