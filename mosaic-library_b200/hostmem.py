"""Page-locked host arrays that DMA at link speed from the first copy.

torch's pin_memory() / cudaHostAlloc hand out 4 KB pages; cycling a few frame-sized buffers of those through the copy engine
runs at 50-90 % of the link rate on the test hosts until the I/O translation caches have warmed up (sometimes for the whole
life of a short job: tools/pin_probe3.py).  A 2 MB-aligned anonymous mapping with MADV_HUGEPAGE, registered with CUDA, runs at
the full rate immediately.  pinned_empty() carves numpy arrays out of such regions:

    frame_ring = [pinned_empty((1080, 1920, 3)) for _ in range(4)]      # fill these from the decoder, pass them to Mapper.update
"""
from __future__ import annotations

import mmap

import numpy as np
import torch

_HUGE = 2 << 20
_CHUNK = 64 << 20


class _Arena:
    def __init__(self, nbytes: int):
        size = ((nbytes + _HUGE - 1) // _HUGE) * _HUGE
        self._mm = mmap.mmap(-1, size + _HUGE, flags=mmap.MAP_PRIVATE | mmap.MAP_ANONYMOUS)
        try:
            self._mm.madvise(mmap.MADV_HUGEPAGE)
        except (AttributeError, OSError):          # no THP on this host: still correct, just ordinary pages
            pass
        raw = np.frombuffer(self._mm, dtype=np.uint8)
        off = (-raw.ctypes.data) % _HUGE
        self.buf = raw[off:off + size]
        self.buf[::4096] = 0                       # first touch: this is when the huge pages are handed out
        torch.cuda.init()
        rc = torch.cuda.cudart().cudaHostRegister(self.buf.ctypes.data, size, 0)
        if int(rc) != 0:
            raise RuntimeError(f"cudaHostRegister failed: {rc}")
        self.size, self.used = size, 0

    def take(self, nbytes: int):
        start = (self.used + 255) & ~255
        if start + nbytes > self.size:
            return None
        self.used = start + nbytes
        return self.buf[start:start + nbytes]


_arenas: list[_Arena] = []


def pinned_empty(shape, dtype=np.uint8) -> np.ndarray:
    """Uninitialised page-locked (CUDA-registered, huge-page backed) host array.  The memory lives as long as the process."""
    dtype = np.dtype(dtype)
    nbytes = int(np.prod(shape)) * dtype.itemsize
    for a in _arenas:
        blk = a.take(nbytes)
        if blk is not None:
            break
    else:
        _arenas.append(_Arena(max(_CHUNK, nbytes)))
        blk = _arenas[-1].take(nbytes)
    return blk.view(dtype).reshape(shape)
