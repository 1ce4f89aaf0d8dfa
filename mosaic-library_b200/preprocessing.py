"""B200 drop-in for the per-frame part of mosaicking.preprocessing (row N4: pre-processing on the device).

Mirrors the reference (paths relative to the mosaic-library checkout):
  make_gray / make_bgr        src/mosaicking/preprocessing.py:314-346   (cv2.cvtColor BGR[A]2GRAY, GRAY2BGR, BGRA2BGR)
  DistortionMapper            src/mosaicking/preprocessing.py:68-98     (cv2.remap with the maps of cv2.initUndistortRectifyMap)
  Crop                        src/mosaicking/preprocessing.py:49-66     (a view, no kernel)
  Pipeline                    src/mosaicking/preprocessing.py:121-152   (function chain)

A frame handed in as a CUDA tensor stays on the device: it is converted / undistorted by mk_cvt_color_u8 / mk_remap_u8
(csrc/mk_preproc.cu) and can go straight into Matcher / Mapper.update_device.  numpy input is uploaded, processed on the device
and returned as a numpy array, like the reference's ndarray-in / ndarray-out contract.  Results equal cv2's bit for bit.
The maps themselves are 3x3 / 5-coefficient host math done once per camera by cv2 (as in the reference) and kept on the device;
the reference rebuilds them for every frame.  There is no CPU fallback for the pixel work.
"""
from __future__ import annotations

from functools import reduce
from typing import Any, Callable, Sequence

import numpy as np
import torch

from . import _lib
from .mosaic import INTERP, _round_up
from .registration import _cuda_device, _current_device

try:
    import cv2 as _cv2
except Exception:  # pragma: no cover
    _cv2 = None

CVT_BGR2GRAY, CVT_BGRA2GRAY, CVT_GRAY2BGR, CVT_BGRA2BGR = 0, 1, 2, 3


def _to_device_rows(image, device) -> tuple[torch.Tensor, int, int, int, bool]:
    """-> (uint8 [h][pitch] device tensor, h, w, channels, input was a host array).  Device tensors whose rows are already
    16-byte aligned are used in place."""
    host = not isinstance(image, torch.Tensor)
    t = torch.from_numpy(np.ascontiguousarray(image)) if host else image
    if t.dtype != torch.uint8:
        raise TypeError("uint8 images only")
    cn = 1 if t.ndim == 2 else int(t.shape[2])
    h, w = int(t.shape[0]), int(t.shape[1])
    if (not host) and t.is_cuda and t.is_contiguous() and (w * cn) % 16 == 0 and t.data_ptr() % 16 == 0:
        return t.reshape(h, w * cn), h, w, cn, host
    pitch = _round_up(w * cn, 128)
    buf = torch.empty((h, pitch), dtype=torch.uint8, device=device)
    buf[:, : w * cn].copy_(t.reshape(h, w * cn), non_blocking=True)
    return buf, h, w, cn, host


def _finish(dst: torch.Tensor, h: int, w: int, cn: int, host: bool):
    out = dst[:, : w * cn]
    out = out.reshape(h, w) if cn == 1 else out.reshape(h, w, cn)
    return out.cpu().numpy() if host else out


def _cvt(image, code: int, dst_cn: int, device=None):
    device = _cuda_device(device if device is not None else (image.device if isinstance(image, torch.Tensor) and image.is_cuda else None))
    src, h, w, cn, host = _to_device_rows(image, device)
    dpitch = _round_up(w * dst_cn, 128)
    dst = torch.empty((h, dpitch), dtype=torch.uint8, device=device)
    with _current_device(device):
        _lib.check(_lib.lib().mk_cvt_color_u8(src.data_ptr(), h, w, src.stride(0), code, dst.data_ptr(), dpitch,
                                              torch.cuda.current_stream(device).cuda_stream))
    return _finish(dst, h, w, dst_cn, host)


def make_gray(image, device=None):
    """preprocessing.make_gray (preprocessing.py:314-329): BGR / BGRA -> gray, anything else is returned unchanged."""
    if image.ndim > 2:
        if image.shape[-1] == 3:
            return _cvt(image, CVT_BGR2GRAY, 1, device)
        if image.shape[-1] == 4:
            return _cvt(image, CVT_BGRA2GRAY, 1, device)
    return image


def make_bgr(image, device=None):
    """preprocessing.make_bgr (preprocessing.py:332-346): gray / BGRA -> BGR, anything else is returned unchanged."""
    if image.ndim == 2:
        return _cvt(image, CVT_GRAY2BGR, 3, device)
    if image.shape[-1] == 4:
        return _cvt(image, CVT_BGRA2BGR, 3, device)
    return image


def remap(image, xmap, ymap, interp="cubic", device=None):
    """cv2.remap(image, xmap, ymap, interp) with float32 maps and BORDER_CONSTANT 0; maps may be numpy arrays or device tensors."""
    device = _cuda_device(device if device is not None else (image.device if isinstance(image, torch.Tensor) and image.is_cuda else None))
    src, h, w, cn, host = _to_device_rows(image, device)
    mx = xmap if isinstance(xmap, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(xmap, dtype=np.float32))
    my = ymap if isinstance(ymap, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(ymap, dtype=np.float32))
    mx = mx.to(device=device, dtype=torch.float32).contiguous(); my = my.to(device=device, dtype=torch.float32).contiguous()
    if mx.shape != my.shape or mx.ndim != 2:
        raise ValueError("xmap / ymap must be two float32 planes of the same shape")
    Hd, Wd = int(mx.shape[0]), int(mx.shape[1])
    dpitch = _round_up(Wd * cn, 128)
    dst = torch.empty((Hd, dpitch), dtype=torch.uint8, device=device)
    with _current_device(device):
        _lib.check(_lib.lib().mk_remap_u8(src.data_ptr(), h, w, cn, src.stride(0), mx.data_ptr(), my.data_ptr(), Wd * 4, dst.data_ptr(),
                                          Hd, Wd, dpitch, INTERP[interp], torch.cuda.current_stream(device).cuda_stream))
    return _finish(dst, Hd, Wd, cn, host)


class Preprocessor:
    """preprocessing.Preprocessor (preprocessing.py:20-29)."""

    def apply(self, img, stream=None):
        raise NotImplementedError


class Crop(Preprocessor):
    """preprocessing.Crop (preprocessing.py:49-66): roi = (x, y, width, height); a view for arrays and tensors alike."""

    def __init__(self, roi: tuple[int, int, int, int]):
        self._roi = roi

    def apply(self, img, stream=None):
        x, y, width, height = self._roi
        return img[y:y + height, x:x + width]


class DistortionMapper(Preprocessor):
    """preprocessing.DistortionMapper (preprocessing.py:68-98): undistortion (or its inverse) of every frame of a camera
    with intrinsics K and distortion D, INTER_CUBIC like the reference.  The maps are built once per frame size and kept on
    the device (the reference calls cv2.initUndistortRectifyMap for every frame)."""

    def __init__(self, K: np.ndarray, D: np.ndarray, inverse: bool = False, device=None):
        self._K = np.asarray(K, dtype=np.float64)
        self._D = np.asarray(D, dtype=np.float64)
        self._inverse = inverse
        self._device = _cuda_device(device)
        self._maps: dict[tuple[int, int], tuple[torch.Tensor, torch.Tensor]] = {}

    def maps(self, width: int, height: int) -> tuple[torch.Tensor, torch.Tensor]:
        key = (width, height)
        if key not in self._maps:
            if _cv2 is None:
                raise _lib.MosaicB200Error("DistortionMapper needs cv2 for initUndistortRectifyMap (host-side camera model)")
            if not self._inverse:
                xmap, ymap = _cv2.initUndistortRectifyMap(self._K, self._D, None, self._K, (width, height), _cv2.CV_32FC1)
            else:
                xmap, ymap = _cv2.initInverseRectificationMap(self._K, self._D, np.eye(3), self._K, (width, height), _cv2.CV_32FC1)
            self._maps[key] = (torch.from_numpy(xmap).to(self._device), torch.from_numpy(ymap).to(self._device))
        return self._maps[key]

    def apply(self, image, stream=None):
        height, width = int(image.shape[0]), int(image.shape[1])
        xmap, ymap = self.maps(width, height)
        return remap(image, xmap, ymap, "cubic", self._device)


class Pipeline(Preprocessor):
    """preprocessing.Pipeline (preprocessing.py:121-152): applies the preprocessors in order.  A numpy frame is uploaded once,
    runs through the whole chain on the device and comes back as a numpy array (the reference does the same with a GpuMat)."""

    def __init__(self, preprocessors: Sequence[Preprocessor], args: Sequence[dict[str, Any]] = None, device=None):
        if args is None:
            args = [{}] * len(preprocessors)
        if len(preprocessors) != len(args):
            raise ValueError("The number of preprocessors must match the number of argument dictionaries.")
        self._device = device
        self._pipeline: Callable = reduce(self.chain, zip(preprocessors, args), self.identity)

    @staticmethod
    def identity(img, **kwargs):
        return img

    @staticmethod
    def chain(f: Callable, g_arg: tuple[Preprocessor, dict[str, Any]]) -> Callable:
        g, arg = g_arg

        def chained(img, **kwargs):
            return g.apply(f(img, **kwargs), **arg)
        return chained

    def apply(self, img, stream=None):
        if isinstance(img, np.ndarray):
            tmp = torch.from_numpy(np.ascontiguousarray(img)).to(_cuda_device(self._device))
            output = self._pipeline(tmp, stream=stream)
            return output.cpu().numpy() if isinstance(output, torch.Tensor) else output
        return self._pipeline(img, stream=stream)
