"""B200 drop-in for mosaicking.mosaic.Mapper and the frame -> tile routing helpers.

Mirrors the reference (paths relative to the mosaic-library checkout):
  Mapper                       src/mosaicking/mosaic.py:36-183   (update -> _update_cpu :111-138)
  get_corners                  src/mosaicking/mosaic.py:727-768
  get_mosaic_dimensions        src/mosaicking/mosaic.py:708-725
  find_nice_homographies       src/mosaicking/mosaic.py:693-706
  bbox_overlap                 src/mosaicking/registration.py:375-380 (shapely Polygon.intersects)
  homogeneous_translation      src/mosaicking/transformations.py:533-536
  assign_frames_to_tiles       the tile loop of SequentialMosaic._create_tile_graph, mosaic.py:595-619
  calculate_tiles / assign_image_to_tiles / calculate_translation_homography   splitter.py:11-77
  combine_tiles                examples/combine_tiles.py:46-62

The per-frame arithmetic (warp, validity mask, 5x5 erosion, blend, mask accumulation) is ONE fused
CUDA kernel (csrc/mk_mapper.cu) reached through the C ABI of include/mosaic_b200.h; the canvas and
its mask stay resident in HBM for the Mapper's lifetime.  There is no CPU fallback.
"""
from __future__ import annotations

import ctypes
import math
from typing import Sequence

import numpy as np
import torch

from . import _lib
from .registration import _cuda_device, _current_device, _raw_stream

try:
    import cv2 as _cv2
except Exception:  # pragma: no cover
    _cv2 = None

INTERP = {"linear": _lib.INTER_LINEAR, "cubic": _lib.INTER_CUBIC, 1: _lib.INTER_LINEAR, 2: _lib.INTER_CUBIC}
BLEND = {"alpha": _lib.BLEND_ALPHA, "feather": _lib.BLEND_FEATHER}


def _round_up(x: int, m: int) -> int:
    return (x + m - 1) // m * m


def make_bgr_host(image: np.ndarray) -> np.ndarray:
    """preprocessing.make_bgr (preprocessing.py:332-346) for host arrays: gray -> BGR, BGRA -> BGR."""
    if image.ndim == 2:
        return np.repeat(image[:, :, None], 3, axis=2)
    if image.shape[-1] == 4:
        return np.ascontiguousarray(image[:, :, :3])
    if image.shape[-1] == 1:
        return np.repeat(image, 3, axis=2)
    return image


class Mapper:
    """Accumulates warped frames into a canvas resident on the GPU (reference: mosaic.py:36-183).

    Mapper(output_width, output_height, alpha_blend=0.5) and update(image, H, stream, keypoints)
    keep the reference's signature.  Extra keyword-only knobs:
      interp   "cubic" (what the reference hard-codes, mosaic.py:114) or "linear" (north-star kernel)
      blend    "alpha" (reference semantics) or "feather" (SURVEY.md A.4b extension)
      async_upload  False (default): update() returns once a HOST image has been consumed (its host-to-device copy is
               complete), like the reference's synchronous update() -- the caller may overwrite the array at once.
               True: the copy of a page-locked array stays in flight after update() returns (it overlaps the previous
               frame's kernel).  The array is kept alive by the Mapper, `last_upload` is an event that completes when the
               copy has read it, and the host never runs more than len(ring) = 3 uploads ahead: a page-locked buffer may
               be refilled after `last_upload.synchronize()` or once update() has been called three more times.
    """

    def __init__(self, output_width: int, output_height: int, alpha_blend: float = 0.5, *, interp="cubic",
                 blend: str = "alpha", feather_ramp: float | None = None, device=None, async_upload: bool = False):
        assert 0.0 <= alpha_blend <= 1.0, "Alpha blend must in interval [0, 1]."
        self._alpha = float(alpha_blend)
        self._interp = INTERP[interp]
        self._blend = BLEND[blend]
        self._ramp = feather_ramp
        self._device = _cuda_device(device)
        self._lib = _lib.lib()
        self._width, self._height = int(output_width), int(output_height)
        self._canvas, self._canvas_mask, self._wacc = self._create_canvas(self._width, self._height)
        # host frames are uploaded on a side stream into a small ring of staging buffers so that the
        # H2D copy of frame k+1 overlaps the kernel of frame k
        self._copy_stream = torch.cuda.Stream(device=self._device)
        self._copy_sptr = self._copy_stream.cuda_stream
        self._cargs = None          # cached canvas / mask / weight pointers and pitches (they only change on release)
        self._slots = [{"buf": None, "roi": None, "free": None, "ready": None, "host": None} for _ in range(3)]
        self._slot = 0
        self._async_upload = bool(async_upload)
        self.last_upload = None     # event: the most recent host image has been read by its host-to-device copy
        self._region_buf = (ctypes.c_int32 * 4)()
        self._flag = True
        self.last_region = (0, 0, 0, 0)

    # canvas rows are padded so that every 64-px tile row is made of whole 16-byte vectors
    def _create_canvas(self, output_width: int, output_height: int):
        cp = _round_up(_round_up(output_width, 64) * 3, 256)
        mp = _round_up(_round_up(output_width, 64), 256)
        canvas = torch.zeros((output_height, cp), dtype=torch.uint8, device=self._device)
        mask = torch.zeros((output_height, mp), dtype=torch.uint8, device=self._device)
        wacc = None
        if self._blend == _lib.BLEND_FEATHER:
            wacc = torch.zeros((output_height, _round_up(output_width, 64)), dtype=torch.float32, device=self._device)
        return canvas, mask, wacc

    def _create_keypoint_mask(self, keypoints, image_dims: tuple[int, int]) -> np.ndarray:
        """mosaic.py:140-157: filled convex hull of the keypoints (host side, as in the reference)."""
        if _cv2 is None:
            raise _lib.MosaicB200Error("keypoint ROI masks need cv2 for convexHull / fillConvexPoly")
        mask = np.zeros(image_dims[1::-1], dtype=np.uint8)
        hull = _cv2.convexHull(np.array(keypoints).astype(np.int32))
        return _cv2.fillConvexPoly(mask, hull, 255)

    @staticmethod
    def _as_rows(image, channels: int) -> tuple[torch.Tensor, int, int]:
        t = image if isinstance(image, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(image))
        if t.dtype != torch.uint8:
            raise TypeError("Mapper.update needs uint8 images")
        h, w = int(t.shape[0]), int(t.shape[1])
        return t.reshape(h, w * channels), h, w

    @staticmethod
    def _usable_in_place(t2: torch.Tensor) -> bool:
        return t2.is_cuda and t2.stride(1) == 1 and t2.stride(0) % 4 == 0 and t2.data_ptr() % 4 == 0

    def _stage_upload(self, slot: dict, key: str, t2: torch.Tensor) -> torch.Tensor:
        """Copy rows into the slot's pitched device buffer on the copy stream."""
        h, row = int(t2.shape[0]), int(t2.shape[1])
        pitch = _round_up(row, 128)
        buf = slot[key]
        if buf is None or buf.shape[0] < h or buf.shape[1] != pitch:
            buf = torch.empty((h, pitch), dtype=torch.uint8, device=self._device)
            slot[key] = buf
        buf[:h, :row].copy_(t2, non_blocking=True)
        return buf

    def update_device(self, frame: torch.Tensor, H: np.ndarray, roi: torch.Tensor | None = None) -> None:
        """One fused update with a frame that already lives in HBM: uint8 [h, w, 3].  Rows are used in
        place when their stride is a multiple of 4 bytes (16 for the bulk-copy fast path).  `roi` is an
        optional uint8 [h, w] keypoint-hull mask on the device."""
        # lean path: a dense uint8 [h, w, 3] device tensor whose rows can be used in place
        if (roi is None and type(frame) is torch.Tensor and frame.is_cuda and frame.dtype == torch.uint8 and frame.dim() == 3
                and frame.shape[2] == 3 and frame.is_contiguous() and (frame.shape[1] * 3) % 4 == 0 and frame.data_ptr() % 4 == 0):
            self._launch_raw(frame.data_ptr(), int(frame.shape[0]), int(frame.shape[1]), int(frame.shape[1]) * 3, H, frame.device)
            return
        self._update_any(frame, H, roi)

    def update_pointer(self, ptr: int, h: int, w: int, pitch: int, H: np.ndarray) -> None:
        """update_device for a raw device pointer to a BGR uint8 frame [h][pitch] -- e.g. a frame that lives in ANOTHER GPU's
        memory and was mapped with distributed.PeerFrames (cudaIpc): the kernel reads the footprint it needs over NVLink.
        The pointer must be 4-byte aligned with pitch % 4 == 0 (16 / 16 for the TMA-staged fast path)."""
        self._launch_raw(int(ptr), int(h), int(w), int(pitch), H, self._device)

    def _launch_raw(self, ptr: int, h: int, w: int, pitch: int, H, device) -> None:
        Hm = H if (isinstance(H, np.ndarray) and H.dtype == np.float64 and H.shape == (3, 3) and H.flags.c_contiguous) else \
            np.ascontiguousarray(np.asarray(H, dtype=np.float64).reshape(3, 3))
        ca = self._cargs
        if ca is None:
            ca = self._cargs = (self._canvas.data_ptr(), self._canvas.stride(0), self._canvas_mask.data_ptr(), self._canvas_mask.stride(0),
                                self._wacc.data_ptr() if self._wacc is not None else None,
                                self._wacc.stride(0) * 4 if self._wacc is not None else 0, self._height, self._width)
        region = self._region_buf
        ramp = self._ramp if self._ramp is not None else 0.1 * min(h, w)
        with _current_device(self._device):
            rc = self._lib.mk_mapper_update(*ca, ptr, h, w, pitch, None, 0, Hm.__array_interface__["data"][0], self._alpha, self._interp,
                                            self._blend, float(ramp), region, _raw_stream(device))
        if rc:
            _lib.check(rc)
        self.last_region = (region[0], region[1], region[2], region[3])

    def _slot_buffer(self, slot: dict, key: str, h: int, row: int) -> torch.Tensor:
        pitch = _round_up(row, 128)
        buf = slot[key]
        if buf is None or buf.shape[0] < h or buf.shape[1] != pitch:
            buf = torch.empty((h, pitch), dtype=torch.uint8, device=self._device)
            slot[key] = buf
        return buf

    def _update_any(self, image, H, roi) -> None:
        cur = torch.cuda.current_stream(self._device)
        host_np = isinstance(image, np.ndarray) and (roi is None or isinstance(roi, np.ndarray))
        if not host_np:
            t2, h, w = self._as_rows(image, 3)
            r2 = None
            if roi is not None:
                r2, rh, rw = self._as_rows(roi, 1)
                assert (rh, rw) == (h, w)
            if self._usable_in_place(t2) and (r2 is None or self._usable_in_place(r2)):
                self._launch(t2, h, w, r2, H, cur)
                return
        else:
            if image.dtype != np.uint8:
                raise TypeError("Mapper.update needs uint8 images")
            if not image.flags.c_contiguous:
                image = np.ascontiguousarray(image)
            h, w = image.shape[:2]
        slot = self._slots[self._slot]
        self._slot = (self._slot + 1) % len(self._slots)
        cs = self._copy_stream
        if slot["free"] is None:
            slot["free"], slot["ready"] = torch.cuda.Event(), torch.cuda.Event()
        else:
            if host_np and self._async_upload:
                slot["ready"].synchronize()      # bounds the host's run-ahead to the ring depth (see async_upload)
            cs.wait_event(slot["free"])          # the kernel that last read this slot has finished
        rbuf = None
        if host_np:
            # straight cudaMemcpy2DAsync on the copy stream: asynchronous when the array is page-locked
            buf = self._slot_buffer(slot, "buf", h, w * 3)
            _lib.check(self._lib.mk_upload_2d(buf.data_ptr(), buf.stride(0), image.__array_interface__["data"][0], w * 3, w * 3, h,
                                              self._copy_sptr))
            if roi is not None:
                roi = np.ascontiguousarray(roi)
                rbuf = self._slot_buffer(slot, "roi", h, w)
                _lib.check(self._lib.mk_upload_2d(rbuf.data_ptr(), rbuf.stride(0), roi.ctypes.data, w, w, h, cs.cuda_stream))
        else:
            if t2.is_cuda:
                cs.wait_stream(cur)
            with torch.cuda.stream(cs):
                buf = self._stage_upload(slot, "buf", t2)
                rbuf = self._stage_upload(slot, "roi", r2) if r2 is not None else None
        slot["ready"].record(cs)
        cur.wait_event(slot["ready"])
        self._launch(buf, h, w, rbuf, H, cur)
        slot["free"].record(cur)
        if host_np:
            if self._async_upload:
                slot["host"] = (image, roi)      # keep the source arrays alive while the DMA may still read them
                self.last_upload = slot["ready"]
            else:
                slot["ready"].synchronize()      # the reference consumes the image before update() returns

    def _launch(self, buf, h, w, rbuf, H, cur=None) -> None:
        Hm = H if (isinstance(H, np.ndarray) and H.dtype == np.float64 and H.shape == (3, 3) and H.flags.c_contiguous) else \
            np.ascontiguousarray(np.asarray(H, dtype=np.float64).reshape(3, 3))
        region = self._region_buf
        ramp = self._ramp if self._ramp is not None else 0.1 * min(h, w)
        ca = self._cargs
        if ca is None:
            ca = self._cargs = (self._canvas.data_ptr(), self._canvas.stride(0), self._canvas_mask.data_ptr(), self._canvas_mask.stride(0),
                                self._wacc.data_ptr() if self._wacc is not None else None,
                                self._wacc.stride(0) * 4 if self._wacc is not None else 0, self._height, self._width)
        if cur is None:
            cur = torch.cuda.current_stream(self._device)
        with _current_device(self._device):
            rc = self._lib.mk_mapper_update(
                *ca, buf.data_ptr(), h, w, buf.stride(0),
                rbuf.data_ptr() if rbuf is not None else None, rbuf.stride(0) if rbuf is not None else 0,
                Hm.__array_interface__["data"][0], self._alpha, self._interp, self._blend, float(ramp), region, cur.cuda_stream)
        if rc:
            _lib.check(rc)
        self.last_region = (region[0], region[1], region[2], region[3])

    def update_batch(self, frames: Sequence, Hs: Sequence[np.ndarray]) -> None:
        """Composite many frames in sequence order with ONE call: a back-to-back train of strip-kernel launches (alpha
        blending), or one gather launch (feather: every canvas tile is read once, updated by all the frames that reach it,
        written once).  Same result as calling update() frame by frame -- the loop SequentialMosaic.generate runs per tile
        (mosaic.py:678-686).  Frames: uint8 [h, w, 3] arrays / tensors;
        host frames are uploaded first (they all have to be resident during the launch)."""
        n = len(frames)
        if n != len(Hs):
            raise ValueError("frames and Hs must have the same length")
        if n == 0:
            return
        keep, ptrs, hs, ws, pitches = [], (ctypes.c_void_p * n)(), (ctypes.c_int * n)(), (ctypes.c_int * n)(), (ctypes.c_size_t * n)()
        for k, fr in enumerate(frames):
            if isinstance(fr, np.ndarray):
                fr = make_bgr_host(fr)
            t2, h, w = self._as_rows(fr, 3)
            if not self._usable_in_place(t2):
                pitch = _round_up(w * 3, 128)
                buf = torch.empty((h, pitch), dtype=torch.uint8, device=self._device)
                buf[:, : w * 3].copy_(t2, non_blocking=True)
                t2 = buf
            keep.append(t2)
            ptrs[k], hs[k], ws[k], pitches[k] = t2.data_ptr(), h, w, t2.stride(0)
        Hm = np.ascontiguousarray(np.stack([np.asarray(H, np.float64).reshape(3, 3) for H in Hs]))
        ramp = self._ramp if self._ramp is not None else 0.1 * min(int(hs[0]), int(ws[0]))
        with _current_device(self._device):
            _lib.check(self._lib.mk_mapper_update_batch(
                self._canvas.data_ptr(), self._canvas.stride(0), self._canvas_mask.data_ptr(), self._canvas_mask.stride(0),
                self._wacc.data_ptr() if self._wacc is not None else None, self._wacc.stride(0) * 4 if self._wacc is not None else 0,
                self._height, self._width, n, ptrs, hs, ws, pitches, Hm.ctypes.data, self._alpha, self._interp, self._blend, float(ramp),
                torch.cuda.current_stream(self._device).cuda_stream))
        self._batch_keepalive = keep      # the launch is asynchronous: keep the staged frames alive until the next call

    def update(self, image, H: np.ndarray, stream=None, keypoints=None):
        """mosaic.py:159-172.  `stream` is accepted and ignored, as in the reference."""
        mask = None
        if keypoints is not None and len(keypoints) and _cv2 is not None and isinstance(keypoints[0], _cv2.KeyPoint):
            keypoints = _cv2.KeyPoint.convert(keypoints)
            height, width = image.shape[:2]
            mask = self._create_keypoint_mask(keypoints, (width, height))
        if isinstance(image, np.ndarray):
            image = make_bgr_host(image)
        self._update_any(image, H, mask)

    def release(self):
        self._canvas = self._canvas_mask = self._wacc = None
        self._cargs = None
        self._slots = [{"buf": None, "roi": None, "free": None, "ready": None, "host": None} for _ in range(3)]
        self.last_upload = None

    @property
    def output(self) -> np.ndarray:
        """Host copy of the canvas, uint8 [H, W, 3] (mosaic.py:179-183)."""
        return self.output_device.cpu().numpy()

    @property
    def output_device(self) -> torch.Tensor:
        return self._canvas[:, : self._width * 3].reshape(self._height, self._width, 3)

    @property
    def mask(self) -> np.ndarray:
        """Host copy of Mapper._canvas_mask, uint8 [H, W]."""
        return self._canvas_mask[:, : self._width].cpu().numpy()

    @property
    def weights(self) -> np.ndarray | None:
        return None if self._wacc is None else self._wacc[:, : self._width].cpu().numpy()

    def footprint_bytes(self, h: int, w: int, H: np.ndarray) -> int:
        """Algorithmic bytes of one update: src + 2*3*A + 2*A (+ 8*A in feather mode)."""
        Hm = np.ascontiguousarray(np.asarray(H, dtype=np.float64).reshape(3, 3))
        region = (ctypes.c_int32 * 4)()
        _lib.check(self._lib.mk_mapper_footprint(self._height, self._width, h, w, Hm.ctypes.data, self._interp, region))
        area = max(0, region[2] - region[0]) * max(0, region[3] - region[1])
        return h * w * 3 + area * (8 + (8 if self._blend == _lib.BLEND_FEATHER else 0))


def warp_perspective(image, H: np.ndarray, dsize: tuple[int, int], interp="linear", device=None, out: torch.Tensor | None = None) -> torch.Tensor:
    """cv2.warpPerspective(image, H, dsize, None, interp, BORDER_CONSTANT, 0) on the GPU
    (mosaic.py:114 / transformations.py:116).  uint8, 1 or 3 channels; returns a device tensor.

    A contiguous device tensor whose rows are a multiple of 16 bytes (1920 or 3840 px x 3) is read in place; anything else is
    copied into a pitched staging buffer first.  `out` (device, uint8, contiguous, Hd x Wd[ x cn], rows a multiple of 16 bytes)
    receives the result without an allocation -- like cv2's `dst` argument."""
    device = _cuda_device(device)
    t = image if isinstance(image, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(image))
    cn = 1 if t.ndim == 2 else int(t.shape[2])
    h, w = int(t.shape[0]), int(t.shape[1])
    if t.is_cuda and t.device == device and t.dtype == torch.uint8 and t.is_contiguous() and (w * cn) % 16 == 0 and t.data_ptr() % 16 == 0:
        src, pitch = t.reshape(h, w * cn), w * cn
    else:
        pitch = _round_up(w * cn, 128)
        src = torch.empty((h, pitch), dtype=torch.uint8, device=device)
        src[:, : w * cn].copy_(t.reshape(h, w * cn), non_blocking=True)
    Wd, Hd = int(dsize[0]), int(dsize[1])
    if out is not None:
        want = (Hd, Wd) if cn == 1 else (Hd, Wd, cn)
        if (not out.is_cuda or out.device != device or out.dtype != torch.uint8 or tuple(out.shape) != want or not out.is_contiguous()
                or (Wd * cn) % 16 or out.data_ptr() % 16):
            raise ValueError(f"warp_perspective: out must be a contiguous uint8 tensor {want} on {device} with rows of a multiple of 16 bytes")
        dst, dpitch = out.reshape(Hd, Wd * cn), Wd * cn
    else:
        dpitch = _round_up(Wd * cn, 128)
        dst = torch.empty((Hd, dpitch), dtype=torch.uint8, device=device)
    Hm = np.ascontiguousarray(np.asarray(H, dtype=np.float64).reshape(3, 3))
    with _current_device(device):
        _lib.check(_lib.lib().mk_warp_perspective_u8(src.data_ptr(), h, w, cn, pitch, Hm.ctypes.data, dst.data_ptr(), Hd, Wd,
                                                     dpitch, INTERP[interp], torch.cuda.current_stream(device).cuda_stream))
    if out is not None:
        return out
    res = dst[:, : Wd * cn]
    return res.reshape(Hd, Wd) if cn == 1 else res.reshape(Hd, Wd, cn)


# ------------------------------------------------------------------------------------------------
# Host-side geometry (3x3 float64 math; defines the multi-GPU partition)
# ------------------------------------------------------------------------------------------------
def homogeneous_translation(x: float, y: float) -> np.ndarray:
    out = np.eye(3)
    out[[0, 1], 2] = [x, y]
    return out


def find_nice_homographies(H: np.ndarray, eps: float = 1e-3) -> np.ndarray:
    det = np.linalg.det(H[:, :2, :2]) if H.ndim > 2 else np.linalg.det(H)
    return det > eps


def get_corners(H: np.ndarray, width, height) -> np.ndarray:
    """Warped positions of the image corners (taken at pixel centres, +0.5), shape M x N x 1 x 2."""
    if isinstance(width, (int, np.integer)):
        src_crn = np.array([[[0, 0]], [[width - 1, 0]], [[width - 1, height - 1]], [[0, height - 1]]], np.float32) + 0.5
    elif len(width):
        crns = [[[0, 0]]]
        for w, h in zip(width, height):
            crns.extend([[[w - 1, 0]], [[w - 1, h - 1]], [[0, h - 1]]])
        src_crn = np.array(crns, np.float32) + 0.5
    else:
        raise ValueError("width and height must either be ints or sequences of ints with same length.")
    H = np.asarray(H)
    if H.ndim > 2:
        src_crn = np.stack([src_crn] * len(H), 0)
    else:
        src_crn = src_crn[None, ...]
        H = H[None, ...]
    M, N = src_crn.shape[:2]
    src_h = np.concatenate((src_crn, np.ones((M, N, 1, 1))), axis=-1)
    dst_h = np.swapaxes(H @ np.swapaxes(src_h.squeeze(axis=2), 1, 2), 1, 2)
    dst = dst_h[:, :, :2] / dst_h[:, :, -1:]
    return dst[:, :, None, :]


def get_mosaic_dimensions(H: np.ndarray, width, height):
    dst = get_corners(H, width, height).reshape(-1, 2)
    min_x, min_y = dst.min(axis=0)
    max_x, max_y = dst.max(axis=0)
    return min_x, min_y, max_x - min_x, max_y - min_y


def _convex_intersects(a: np.ndarray, b: np.ndarray) -> bool:
    """Closed convex polygons share a point <=> no separating axis with a strict gap."""
    for poly in (a, b):
        n = len(poly)
        for i in range(n):
            ex, ey = poly[(i + 1) % n] - poly[i]
            if ex == 0 and ey == 0:
                continue
            nx, ny = -ey, ex
            pa = a[:, 0] * nx + a[:, 1] * ny
            pb = b[:, 0] * nx + b[:, 1] * ny
            if pa.max() < pb.min() or pb.max() < pa.min():
                return False
    return True


def bbox_overlap(shape_1: np.ndarray, shape_2: np.ndarray) -> bool:
    """registration.py:375-380: Polygon(shape_1).intersects(Polygon(shape_2)) for convex quads."""
    a = np.asarray(shape_1, np.float64).reshape(-1, 2)
    b = np.asarray(shape_2, np.float64).reshape(-1, 2)
    return _convex_intersects(a, b)


def assign_frames_to_tiles(H_all: np.ndarray, dims: Sequence[tuple[int, int]], tile_size: tuple[int, int]):
    """The tile loop of SequentialMosaic._create_tile_graph (mosaic.py:589-619).

    H_all: N x 3 x 3 absolute homographies (top-left referenced); dims: N x (width, height).
    Returns {(tile_x_idx, tile_y_idx): {frame_index: H_tile}} with H_tile = T(-tx,-ty) @ H (:617)."""
    H_all = np.asarray(H_all, np.float64)
    widths = [int(d[0]) for d in dims]
    heights = [int(d[1]) for d in dims]
    mosaic_dims = get_mosaic_dimensions(H_all, widths, heights)
    tile_x = range(0, math.ceil(mosaic_dims[2]), tile_size[0])
    tile_y = range(0, math.ceil(mosaic_dims[3]), tile_size[1])
    corners = [get_corners(H_all[k], widths[k], heights[k]).squeeze().astype(int) for k in range(len(H_all))]
    tiles = {}
    for xi, tx in enumerate(tile_x):
        for yi, ty in enumerate(tile_y):
            tile_crns = np.array([[tx, ty], [tx + tile_size[0], ty], [tx + tile_size[0], ty + tile_size[1]],
                                  [tx, ty + tile_size[1]]], dtype=int)
            frames = {}
            for k in range(len(H_all)):
                if bbox_overlap(tile_crns, corners[k]):
                    frames[k] = homogeneous_translation(-tx, -ty) @ H_all[k]
            if frames:
                tiles[(xi, yi)] = frames
    return tiles


def calculate_tiles(bounding_box, tile_size):
    """splitter.py:11-39."""
    min_x, min_y, width, height = bounding_box
    max_x, max_y = min_x + width, min_y + height
    tw, th = tile_size
    tiles = []
    for i in range(int(np.ceil(width / tw))):
        for j in range(int(np.ceil(height / th))):
            x0, y0 = min_x + i * tw, min_y + j * th
            tiles.append((x0, y0, min(x0 + tw, max_x), min(y0 + th, max_y)))
    return tiles


def assign_image_to_tiles(warped_corners, tiles):
    """splitter.py:42-58: a tile is selected when one of the warped corners falls inside it."""
    out = []
    for idx, (min_x, min_y, max_x, max_y) in enumerate(tiles):
        if np.any((warped_corners[:, 0] >= min_x) & (warped_corners[:, 0] < max_x) &
                  (warped_corners[:, 1] >= min_y) & (warped_corners[:, 1] < max_y)):
            out.append(idx)
    return tuple(out)


def calculate_translation_homography(homography, tile_origin):
    """splitter.py:61-77."""
    return np.array([[1, 0, -tile_origin[0]], [0, 1, -tile_origin[1]], [0, 0, 1]]) @ homography


def combine_tiles(tiles: dict, tile_size: tuple[int, int]) -> np.ndarray:
    """examples/combine_tiles.py:46-62: paste tile (x, y) at (x*tile_w, y*tile_h) of one canvas."""
    tw, th = tile_size
    max_x = max(x for x, _ in tiles) * tw + tw
    max_y = max(y for _, y in tiles) * th + th
    canvas = np.zeros((max_y, max_x, 3), dtype=np.uint8)
    for (x, y), img in tiles.items():
        canvas[y * th:y * th + th, x * tw:x * tw + tw] = img
    return canvas
