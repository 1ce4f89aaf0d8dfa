"""The reference's CPU path restated over its own L0 dependency (cv2 + numpy) -- TEST INFRASTRUCTURE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  It is the timed CPU baseline ("kind": "port"): the reference is pure Python
glue over OpenCV, /root/reference does not exist on the GPU box, so the glue is restated here
statement by statement while the arithmetic stays in the very same cv2 / numpy calls, with all
the host threads cv2 uses by default.

  RefMatcher.knn_match      src/mosaicking/registration.py:230-242   cv2.BFMatcher(NORM_L2).knnMatch(k=2)
  lowe_ratio                src/mosaicking/mosaic.py:430
  RefMapper.update          src/mosaicking/mosaic.py:111-138 (Mapper._update_cpu), :159-172
  RefMapper._keypoint_mask  src/mosaicking/mosaic.py:140-157

Differences from the reference, all opt-in: `norm` (the reference is always NORM_L2; NORM_HAMMING is
the north-star ORB metric) and `flags` (the reference hard-codes INTER_CUBIC; INTER_LINEAR is the
north-star kernel).  tests/test_cv_ref.py checks this port against the real reference classes when
/root/reference is present, and against the C oracle everywhere.
"""
from __future__ import annotations

import cv2
import numpy as np


class RefMatcher:
    def __init__(self, norm: int = cv2.NORM_L2):
        self._matcher = cv2.BFMatcher.create(norm, crossCheck=False)   # registration.py:231

    def knn_match(self, query: np.ndarray, train: np.ndarray, mask=None):
        return self._matcher.knnMatch(query, train, k=2, mask=mask)   # registration.py:242


def lowe_ratio(matches, ratio: float = 0.7):
    """mosaic.py:430"""
    return tuple(m[0] for m in matches if len(m) == 2 and m[0].distance < ratio * m[1].distance)


def make_bgr(image: np.ndarray) -> np.ndarray:
    """preprocessing.py:332-346 (CPU branch)"""
    if image.ndim == 2:
        return cv2.cvtColor(image, cv2.COLOR_GRAY2BGR)
    if image.shape[-1] == 4:
        return cv2.cvtColor(image, cv2.COLOR_BGRA2BGR)
    return image


class RefMapper:
    def __init__(self, output_width: int, output_height: int, alpha_blend: float = 0.5, flags: int = cv2.INTER_CUBIC):
        assert 0.0 <= alpha_blend <= 1.0, "Alpha blend must in interval [0, 1]."
        self._alpha = alpha_blend
        self._flags = flags
        self._canvas = np.zeros((output_height, output_width, 3), dtype=np.uint8)     # mosaic.py:60-63
        self._canvas_mask = np.zeros((output_height, output_width), dtype=np.uint8)

    def _update_cpu(self, image, H, mask=None):
        image = make_bgr(image)
        dsize = self._canvas.shape[1::-1]
        warped = cv2.warpPerspective(image, H, dsize, None, self._flags, borderMode=cv2.BORDER_CONSTANT, borderValue=(0, 0, 0))
        if mask is None:
            warped_mask = np.where(warped.any(axis=2), 255, 0).astype(np.uint8)
        else:
            warped_mask = cv2.warpPerspective(mask, H, dsize, None, self._flags, borderMode=cv2.BORDER_CONSTANT, borderValue=0)
        warped_mask = cv2.morphologyEx(warped_mask, cv2.MORPH_ERODE, cv2.getStructuringElement(cv2.MORPH_RECT, (5, 5)))
        warped_mask_bin = warped_mask.copy()
        output_mask_bin = cv2.threshold(self._canvas_mask, 1, 255, cv2.THRESH_BINARY)[1]
        mask_intersect = cv2.bitwise_and(output_mask_bin, warped_mask_bin)
        warped_mask_only = cv2.bitwise_and(warped_mask_bin, cv2.bitwise_not(output_mask_bin))
        self._canvas = np.where(warped_mask_only[..., None], warped, self._canvas)
        blended = cv2.addWeighted(self._canvas, self._alpha, warped, 1 - self._alpha, 0.0)
        self._canvas = np.where(mask_intersect[..., None], blended, self._canvas)
        self._canvas_mask = cv2.bitwise_or(self._canvas_mask, warped_mask)

    @staticmethod
    def _keypoint_mask(keypoints, image_dims):
        mask = np.zeros(image_dims[1::-1], dtype=np.uint8)
        hull = cv2.convexHull(np.array(keypoints).astype(np.int32))
        return cv2.fillConvexPoly(mask, hull, 255)

    def update(self, image, H, stream=None, keypoints=None):
        mask = None
        if keypoints and isinstance(keypoints[0], cv2.KeyPoint):
            keypoints = cv2.KeyPoint.convert(keypoints)
            mask = self._keypoint_mask(keypoints, image.shape[1::-1])
        self._update_cpu(image, H, mask)

    @property
    def output(self) -> np.ndarray:
        return self._canvas
