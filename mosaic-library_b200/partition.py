"""Spatial partition of the GPU between the two halves of a frame step (green contexts, mk_sm_partition_create).

Matching wants one large CTA per SM, compositing five small ones; they cannot share an SM, so on ordinary streams the
kernel that comes second mostly waits.  SmPartition gives each its own group of SMs:

    part = SmPartition(40)                       # >= 40 SMs for matching, the rest for compositing
    with torch.cuda.stream(part.first):  matcher.knn_match_device(q, t)
    with torch.cuda.stream(part.rest):   mapper.update_device(frame, H)

The streams are torch.cuda.ExternalStream objects; order them against other streams with wait_stream / events as usual."""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib


class SmPartition:
    def __init__(self, min_sms_first: int, device=None):
        device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        L = _lib.lib()
        part, s0, s1 = C.c_void_p(), C.c_void_p(), C.c_void_p()
        n0, n1 = C.c_int(), C.c_int()
        with torch.cuda.device(device):
            torch.cuda.current_stream()          # primary context up
            _lib.check(L.mk_sm_partition_create(int(min_sms_first), C.byref(part), C.byref(s0), C.byref(s1), C.byref(n0), C.byref(n1)))
        self._lib, self._handle, self.device = L, part, device
        self.sms_first, self.sms_rest = n0.value, n1.value
        self.first = torch.cuda.ExternalStream(s0.value, device=device)
        self.rest = torch.cuda.ExternalStream(s1.value, device=device)

    def close(self) -> None:
        if self._handle is not None:
            torch.cuda.synchronize(self.device)
            self._lib.mk_sm_partition_destroy(self._handle)
            self._handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
