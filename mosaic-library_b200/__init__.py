"""mosaic-library_b200: B200-native registration + compositing hot path of mosaic-library.

Drop-in for the three hot call sites of the reference:
    registration.Matcher / CompositeMatcher   (reference: src/mosaicking/registration.py:226-298)
    mosaic.Mapper                             (reference: src/mosaicking/mosaic.py:36-183)
    frame -> tile routing helpers             (reference: src/mosaicking/mosaic.py:571-622,708-768)
    preprocessing.make_gray / make_bgr / DistortionMapper on the device (reference: src/mosaicking/preprocessing.py:68-98,314-346)

Everything numeric runs in libmosaic_b200.so (hand-written sm_100a CUDA behind the C ABI of
include/mosaic_b200.h).  HAS_B200 sits next to the reference's HAS_CUDA / HAS_CODEC flags
(src/mosaicking/__init__.py:18-56): True when the library is built and a CUDA device is visible.
There is no CPU fallback; constructing Matcher / Mapper without a GPU raises.
"""
from . import _lib
from ._lib import MosaicB200Error, build, launch_count


def _has_b200() -> bool:
    try:
        import torch
        return _lib.LIB_PATH.exists() and torch.cuda.is_available()
    except Exception:
        return False


HAS_B200 = _has_b200()

from . import mosaic, preprocessing, registration  # noqa: E402
from .mosaic import (Mapper, assign_frames_to_tiles, bbox_overlap, combine_tiles, find_nice_homographies,  # noqa: E402
                     get_corners, get_mosaic_dimensions, homogeneous_translation, warp_perspective)
from .registration import (CompositeMatcher, DMatch, HomographyEstimation, HomographyEstimationType, Matcher, gather_matched_keypoints,  # noqa: E402
                           get_matches, propagate_homographies)
from .partition import SmPartition  # noqa: E402
from .hostmem import pinned_empty  # noqa: E402

__all__ = ["HAS_B200", "MosaicB200Error", "build", "launch_count", "Mapper", "Matcher", "CompositeMatcher", "DMatch",
           "get_matches", "get_corners", "get_mosaic_dimensions", "find_nice_homographies", "bbox_overlap",
           "assign_frames_to_tiles", "combine_tiles", "homogeneous_translation", "warp_perspective", "mosaic",
           "registration", "preprocessing", "SmPartition", "pinned_empty", "HomographyEstimation", "HomographyEstimationType", "gather_matched_keypoints",
           "propagate_homographies"]
