"""B200 drop-in for mosaicking.registration.Matcher / CompositeMatcher.

Mirrors src/mosaicking/registration.py:226-298 of the reference (same class names, method
names, argument meaning, return shapes and pickling contract); the arithmetic runs in
libmosaic_b200 (csrc/mk_knn.cu) instead of cv2.BFMatcher.  PyTorch only owns device memory.

Two things the reference does not have, both optional:
  * `norm`: the reference Matcher is always NORM_L2, even for ORB bytes (registration.py:231);
    "l2" (default) reproduces that bit for bit, "hamming" is the north-star ORB metric
    (cv2.BFMatcher(NORM_HAMMING) semantics).
  * `knn_match_arrays`: idx / dist / keep arrays with the Lowe ratio test (mosaic.py:430) fused
    on the device, so callers need not materialise 2*Nq cv2.DMatch objects.
"""
from __future__ import annotations

import warnings
from itertools import repeat
from typing import Sequence

import numpy as np
import torch

from . import _lib
from .hostmem import pinned_empty

try:  # cv2 is only used for the DMatch value type the reference API returns
    import cv2 as _cv2
    DMatch = _cv2.DMatch
except Exception:  # pragma: no cover - cv2 is part of the image
    _cv2 = None

    class DMatch:  # minimal stand-in with cv2.DMatch's attributes
        __slots__ = ("queryIdx", "trainIdx", "imgIdx", "distance")

        def __init__(self, queryIdx=-1, trainIdx=-1, imgIdx=-1, distance=float("inf")):
            self.queryIdx, self.trainIdx, self.imgIdx, self.distance = queryIdx, trainIdx, imgIdx, distance


def _cuda_device(device=None) -> torch.device:
    if not torch.cuda.is_available():
        raise _lib.MosaicB200Error("no CUDA device: the B200 path has no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)


def _raw_stream(device: torch.device) -> int:
    """cudaStream_t of torch's current stream on `device` (the private fast accessor when this torch has it)."""
    try:
        return torch._C._cuda_getCurrentRawStream(device.index if device.index is not None else torch.cuda.current_device())
    except AttributeError:                        # pragma: no cover - older / newer torch without the accessor
        return torch.cuda.current_stream(device).cuda_stream


def _bad_input(msg: str):
    if _cv2 is not None:
        return _cv2.error(msg)
    return ValueError(msg)


def _norm_code(norm: str, dtype, engine: str = "auto") -> int:
    if dtype == np.uint8:
        if norm == "hamming":
            return _lib.NORM_HAMMING if engine == "popc" else _lib.NORM_HAMMING_TC
        if norm == "l2":
            return _lib.NORM_L2_U8
    elif dtype == np.float32:
        if norm == "l2":
            return _lib.NORM_L2_F32
        if norm == "hamming":
            raise _bad_input("hamming norm needs uint8 descriptors")
    raise _bad_input(f"unsupported descriptor dtype {dtype} / norm {norm} (cv2 accepts CV_8U and CV_32F)")


def upload_descriptors(desc, device=None) -> tuple[torch.Tensor, int]:
    """Host ndarray / tensor [n, dim] -> device tensor whose rows are zero-padded to a multiple of
    16 bytes (zero padding changes neither Hamming nor L2 distances).  Returns (tensor, dim)."""
    device = _cuda_device(device)
    if isinstance(desc, torch.Tensor):
        t = desc
    else:
        a = np.asarray(desc)
        if a.ndim != 2:
            raise _bad_input("descriptors must be 2-D")
        t = torch.from_numpy(np.ascontiguousarray(a))
    if t.dtype not in (torch.uint8, torch.float32):
        raise _bad_input(f"unsupported descriptor dtype {t.dtype}")
    n, dim = t.shape
    row_bytes = dim * t.element_size()
    pad_bytes = (-row_bytes) % 16
    if pad_bytes == 0 and t.is_cuda and t.is_contiguous():
        return t, dim
    pad = pad_bytes // t.element_size()
    out = torch.zeros((n, dim + pad), dtype=t.dtype, device=device)
    if n:
        out[:, :dim].copy_(t, non_blocking=True)
    return out, dim


class DescriptorSets:
    """Descriptor sets resident on the device (Matcher.upload_sets): one row matrix + the offsets of the sets in it."""

    def __init__(self, desc, set_off, sizes, dtype, dim, total):
        self.desc, self.set_off, self.sizes, self.dtype, self.dim, self.total = desc, set_off, sizes, dtype, dim, total

    def __len__(self):
        return len(self.sizes)


class _EventPool:
    """The library events of Matcher.knn_match_device_async.  Pending handles hold a reference, so the events are destroyed
    only after the matcher AND every unjoined handle are gone."""

    def __init__(self, lib, n: int):
        self.lib, self.ev = lib, [lib.mk_event_create() for _ in range(n)]

    def __del__(self):
        try:
            for e in self.ev:
                self.lib.mk_event_destroy(e)
        except Exception:
            pass


class _current_device:
    """`with _current_device(dev):` -- make `dev` the current CUDA device for the C-ABI calls inside (the library allocates
    its tables and sets kernel attributes on the current device); free when it already is."""

    __slots__ = ("_idx", "_prev")

    def __init__(self, device):
        self._idx = device.index if isinstance(device, torch.device) else int(device)

    def __enter__(self):
        self._prev = None
        if self._idx is not None:
            cur = torch.cuda.current_device()
            if cur != self._idx:
                self._prev = cur
                torch.cuda.set_device(self._idx)

    def __exit__(self, *exc):
        if self._prev is not None:
            torch.cuda.set_device(self._prev)
        return False


class PendingDeviceMatch:
    """Handle of one Matcher.knn_match_device_async call: the kernels run on the matcher's own stream."""

    __slots__ = ("_outs", "_lib", "_event", "_own", "_device", "_keepalive", "_pool")

    def __init__(self, outs, lib, event, own_stream, device, keepalive, pool=None):
        self._outs, self._lib, self._event, self._own, self._device, self._keepalive = outs, lib, event, own_stream, device, keepalive
        self._pool = pool            # keeps the library events alive while this handle may still join on one

    def result(self):
        """-> (idx, dist, keep, n_keep) device tensors; the CURRENT stream is made to wait for the matching kernels (no host
        synchronisation), so anything queued afterwards may read them."""
        if self._event is not None:
            self._lib.mk_stream_follow(self._event, _lib.STREAM_NONE, _raw_stream(self._device))   # join: current stream waits
            self._event = self._keepalive = self._pool = None
        return self._outs

    def __del__(self):
        if self._event is not None and self._outs is not None:   # dropped without joining: keep the allocator from recycling early
            try:
                for t in self._outs:
                    t.record_stream(self._own)
            except Exception:
                pass


class PendingMatch:
    """Handle of one queued Matcher.knn_match_arrays_async call."""

    DEPTH = 4

    def __init__(self, slot, nq, off_keep, keepalive):
        self._slot, self._nq, self._off_keep, self._keepalive, self._res = slot, nq, off_keep, keepalive, None

    @classmethod
    def done(cls, idx, dist, keep):
        p = cls(None, 0, 0, None)
        p._res = (idx, dist, keep)
        return p

    def _resolve(self):
        if self._res is None:
            self._slot["event"].synchronize()
            hb, nq = self._slot["host"].numpy(), self._nq
            self._res = (hb[: nq * 8].view(np.int32).reshape(nq, 2).copy(), hb[nq * 8: nq * 16].view(np.float32).reshape(nq, 2).copy(),
                         hb[self._off_keep: self._off_keep + nq].astype(bool))
            self._slot["owner"] = None
            self._slot = self._keepalive = None
        return self._res

    def ready(self) -> bool:
        return self._res is not None or self._slot["event"].query()

    def result(self):
        """-> (idx [nq,2] int32, dist [nq,2] float32, keep [nq] bool), the values knn_match_arrays returns."""
        return self._resolve()


class Matcher:
    """Brute-force k=2 matcher (reference: registration.py:226-264)."""

    def __init__(self, norm: str = "l2", device=None, engine: str = "auto"):
        """engine only matters for norm="hamming": "popc" always runs the XOR+POPC kernel; "auto" (default, alias
        "tensor") lets the library send large problems with descriptors <= 32 bytes through the tcgen05 fp8 GEMM on
        bit-expanded rows (mk_knn2_uses_tensor_cores tells which) -- the results are identical bit for bit."""
        if norm not in ("l2", "hamming"):
            raise ValueError("norm must be 'l2' (reference behaviour) or 'hamming'")
        if engine not in ("auto", "popc", "tensor"):
            raise ValueError("engine must be 'auto', 'popc' or 'tensor'")
        self._norm = norm
        self._engine = engine
        self._device = device
        self._matcher = self._create_matcher()

    # -- reference-shaped plumbing ------------------------------------------------------------
    def _create_matcher(self):
        # the reference builds a cv2.BFMatcher here (registration.py:230-231); ours is a handle on
        # the C-ABI library plus a growable device workspace
        return {"lib": _lib.lib(), "workspace": None}

    def __getstate__(self):
        state = self.__dict__.copy()
        del state["_matcher"]  # no CUDA handles in the pickle (registration.py:255-259)
        return state

    def __setstate__(self, state):
        self.__dict__.update(state)
        self._matcher = self._create_matcher()

    def _workspace(self, nbytes: int, device, key: str = "workspace", stream=None) -> torch.Tensor:
        """Growable scratch.  `stream`: the torch stream the scratch is USED on when that is not the current one -- the block
        is then allocated from that stream's pool, so that replacing it cannot hand memory a kernel on the private stream
        still uses to a later allocation of the caller's stream."""
        ws = self._matcher.get(key)
        if ws is None or ws.numel() < nbytes or ws.device != device:
            if stream is not None:
                with torch.cuda.stream(stream):
                    ws = torch.empty(max(nbytes, 1 << 16), dtype=torch.uint8, device=device)
            else:
                ws = torch.empty(max(nbytes, 1 << 16), dtype=torch.uint8, device=device)
            self._matcher[key] = ws
        return ws

    # -- device-side entry points --------------------------------------------------------------
    def knn_match_device(self, query, train, ratio: float = 0.7, _stream: int | None = None, _tstream=None):
        """-> (idx int32 [nq,2], dist float32 [nq,2], keep uint8 [nq], n_keep int32 [1]) on the device."""
        # lean path for what a frame loop passes: two well-formed device tensors (about a third of the host time of the
        # general path below, which matters once the kernels take ~20 us)
        if (type(query) is torch.Tensor and type(train) is torch.Tensor and query.is_cuda and train.is_cuda
                and query.dtype == train.dtype and query.dim() == 2 and train.dim() == 2 and query.is_contiguous()
                and train.is_contiguous() and query.shape[1] == train.shape[1] and query.shape[0] > 0
                and (query.shape[1] * query.element_size()) % 16 == 0 and query.dtype in (torch.uint8, torch.float32)
                and query.device == train.device):
            device = query.device
            nq, dim = query.shape
            nt = train.shape[0]
            fp = self._matcher.setdefault("fast", {})
            key = (query.dtype, dim, nq, nt, device)
            ent = fp.get(key)
            L = self._matcher["lib"]
            if ent is None:
                code = _norm_code(self._norm, np.uint8 if query.dtype == torch.uint8 else np.float32, self._engine)
                ent = fp[key] = (code, int(L.mk_knn2_workspace_bytes(code, nq, nt, dim)), dim * query.element_size())
                if len(fp) > 256:
                    fp.clear(); fp[key] = ent
            code, ws_bytes, stride = ent
            ws = self._workspace(ws_bytes, device, "workspace" if _stream is None else "workspace_own_stream", _tstream)
            idx = torch.empty((nq, 2), dtype=torch.int32, device=device)
            dist = torch.empty((nq, 2), dtype=torch.float32, device=device)
            keep = torch.empty((nq,), dtype=torch.uint8, device=device)
            n_keep = torch.empty((1,), dtype=torch.int32, device=device)
            with _current_device(device):
                rc = L.mk_knn2(code, query.data_ptr(), nq, train.data_ptr() if nt else None, nt, dim, stride, idx.data_ptr(), dist.data_ptr(),
                               float(ratio), keep.data_ptr(), n_keep.data_ptr(), ws.data_ptr(), ws.numel(),
                               _raw_stream(device) if _stream is None else _stream)
            if rc:
                _lib.check(rc)
            return idx, dist, keep, n_keep
        device = _cuda_device(self._device)
        q, dim = upload_descriptors(query, device)
        t, dim_t = upload_descriptors(train, device)
        if q.dtype != t.dtype:
            raise _bad_input("query and train descriptors must have the same type")
        if dim != dim_t:
            raise _bad_input("query and train descriptors must have the same width")
        code = _norm_code(self._norm, np.uint8 if q.dtype == torch.uint8 else np.float32, self._engine)
        nq, nt = q.shape[0], t.shape[0]
        idx = torch.empty((nq, 2), dtype=torch.int32, device=device)
        dist = torch.empty((nq, 2), dtype=torch.float32, device=device)
        keep = torch.empty((nq,), dtype=torch.uint8, device=device)
        if nq == 0:
            return idx, dist, keep, torch.zeros((1,), dtype=torch.int32, device=device)
        n_keep = torch.empty((1,), dtype=torch.int32, device=device)   # the library zeroes it on the stream
        L = self._matcher["lib"]
        ws = self._workspace(L.mk_knn2_workspace_bytes(code, nq, nt, dim), device, "workspace" if _stream is None else "workspace_own_stream", _tstream)
        stride = q.stride(0) * q.element_size()
        assert stride == t.stride(0) * t.element_size()
        with _current_device(device):
            _lib.check(L.mk_knn2(code, q.data_ptr(), nq, t.data_ptr() if nt else None, nt, dim, stride,
                                 idx.data_ptr(), dist.data_ptr(), float(ratio), keep.data_ptr(), n_keep.data_ptr(),
                                 ws.data_ptr(), ws.numel(), torch.cuda.current_stream(device).cuda_stream if _stream is None else _stream))
        return idx, dist, keep, n_keep

    def knn_match_device_async(self, query: torch.Tensor, train: torch.Tensor, ratio: float = 0.7) -> PendingDeviceMatch:
        """knn_match_device on the matcher's OWN stream: the call orders that stream after the caller's current stream (so
        the descriptors are ready), launches, and returns at once; handle.result() orders the caller's current stream after
        the kernels and hands back the device tensors.  A frame loop writes
            h = matcher.knn_match_device_async(desc_k1, desc_k); mapper.update_device(frame, H); idx, dist, keep, n = h.result()
        and gets matching and compositing side by side without touching streams or events itself."""
        device = query.device
        L = self._matcher["lib"]
        st = self._matcher.setdefault("dasync", {})
        if st.get("device") != device:
            st.clear()
            with _current_device(device):
                st.update(device=device, stream=torch.cuda.Stream(device=device), pool=_EventPool(L, 16), pos=0)
            st["sptr"] = st["stream"].cuda_stream
        k = st["pos"] = (st["pos"] + 2) % 16
        pool = st["pool"]
        ev_in, ev_out = pool.ev[k], pool.ev[k + 1]
        L.mk_stream_follow(ev_in, _raw_stream(device), st["sptr"])      # own stream after whatever produced the descriptors
        outs = self.knn_match_device(query, train, ratio, _stream=st["sptr"], _tstream=st["stream"])
        L.mk_stream_follow(ev_out, st["sptr"], _lib.STREAM_NONE)        # completion of THIS call, joined in result()
        return PendingDeviceMatch(outs, L, ev_out, st["stream"], device, (query, train), pool)

    def knn_match_arrays(self, query, train, ratio: float = 0.7):
        """-> (idx [nq,2] int32, dist [nq,2] float32, keep [nq] bool) as numpy arrays.
        idx = -1 / dist = FLT_MAX where the train set has fewer than 2 rows.
        Host arrays take the lean path: persistent device buffers, plain cudaMemcpyAsync uploads, ONE packed
        device->host copy of idx | dist | keep into page-locked memory and a single stream synchronisation."""
        if isinstance(query, np.ndarray) and isinstance(train, np.ndarray):
            return self._knn_host(query, train, ratio)
        idx, dist, keep, _ = self.knn_match_device(query, train, ratio)
        return idx.cpu().numpy(), dist.cpu().numpy(), keep.cpu().numpy().astype(bool)

    def _knn_host(self, q: np.ndarray, t: np.ndarray, ratio: float):
        return self._knn_host_submit(q, t, ratio).result()

    def knn_match_arrays_async(self, query: np.ndarray, train: np.ndarray, ratio: float = 0.7) -> "PendingMatch":
        """knn_match_arrays without the wait: uploads, kernels and the packed device->host copy are queued on the matcher's
        private stream and the call returns a handle; handle.result() -> (idx, dist, keep) waits for that call only.
        Lets a frame loop queue frame k's matching, hand frame k to Mapper.update and collect frame k-1's matches, so the
        PCIe uploads never wait for a host round trip.  The input arrays must stay alive and unmodified until result();
        at most PendingMatch.DEPTH calls are kept in flight (older handles are resolved when their slot is reused)."""
        if not (isinstance(query, np.ndarray) and isinstance(train, np.ndarray)):
            raise TypeError("knn_match_arrays_async takes host ndarrays")
        return self._knn_host_submit(query, train, ratio)

    def match_next_async(self, desc: np.ndarray, ratio: float = 0.7) -> "PendingMatch":
        """Sequential matching the way SequentialMosaic.match_features walks a video (mosaic.py:402-438:
        knn_match(descriptors, descriptors_prev) for consecutive frames): match `desc` (query) against the set passed to the
        PREVIOUS match_next call (train), which is still in device memory -- every descriptor set crosses the host link once
        instead of twice.  Same results as knn_match_arrays_async(desc, previous_desc, ratio).  The first call of a sequence has
        nothing to match against: its handle resolves to idx = -1, dist = FLT_MAX, keep = False (what an empty train set gives).
        `desc` must stay alive and unmodified until the handle's result() (or the next call's) has returned.
        reset_sequence() starts a new sequence; any knn_match_arrays* call in between does so too."""
        if not isinstance(desc, np.ndarray):
            raise TypeError("match_next_async takes a host ndarray")
        return self._knn_host_submit(desc, None, ratio, seq=True)

    def match_next(self, desc: np.ndarray, ratio: float = 0.7):
        """Blocking form of match_next_async -> (idx, dist, keep)."""
        return self.match_next_async(desc, ratio).result()

    def reset_sequence(self) -> None:
        """Forget the descriptor set kept on the device by match_next: the next match_next call starts a new sequence."""
        hp = self._matcher.get("host_plan")
        if hp is not None:
            hp["seq"] = None

    def _host_plan(self, device, dtype, dim: int, nq: int, nt: int) -> dict:
        """Everything the host path reuses call after call: private stream, device row buffers (rows padded to 16 bytes,
        padding zero), packed output buffer, the ring of page-locked result buffers.  Rebuilt only when the descriptor
        type / width changes or a set outgrows its buffer, so the per-call path is a handful of C-ABI calls."""
        hp = self._matcher.get("host_plan")
        if (hp is not None and hp["key"] == (device, dtype.char, dim) and nq <= hp["cap"] and nt <= hp["cap"]):
            return hp
        stream = self._matcher.get("stream")
        if stream is None or stream.device != device:
            stream = torch.cuda.Stream(device=device)
            self._matcher["stream"] = stream
        if hp is not None:
            for slot in hp["ring"]:
                if slot["owner"] is not None:
                    slot["owner"]._resolve()     # results of queued calls live in the buffers we are about to replace
            stream.synchronize()
        row = dim * dtype.itemsize
        stride = (row + 15) & ~15
        cap = max(nq, nt, 1)
        cap = 1 << (cap - 1).bit_length()
        total_cap = (cap * 17 + 3 & ~3) + 4
        dstream = self._matcher.get("dstream")
        if dstream is None or dstream.device != device:
            dstream = torch.cuda.Stream(device=device)
            self._matcher["dstream"] = dstream
        with torch.cuda.stream(stream):          # allocations and fills happen on the private stream
            bufs = [torch.zeros(cap * stride, dtype=torch.uint8, device=device) for _ in range(2)]
            outs = [torch.empty(total_cap, dtype=torch.uint8, device=device) for _ in range(PendingMatch.DEPTH)]
        # every slot: page-locked result buffer + its own device output buffer (the packed result leaves on the DOWNLOAD stream,
        # so the next call's uploads and kernels never queue behind a device->host copy: 140 -> 130 us per frame of copy-engine
        # time in the frame loop, tools/copy_mix_probe2.py)
        ring = [{"host": torch.from_numpy(pinned_empty((total_cap,))), "event": torch.cuda.Event(), "kdone": torch.cuda.Event(),
                 "out": outs[i], "out_ptr": outs[i].data_ptr(), "owner": None} for i in range(PendingMatch.DEPTH)]
        for slot in ring:
            slot["host_ptr"] = slot["host"].data_ptr()
        seq = None
        if hp is not None and hp["seq"] is not None and hp["key"] == (device, dtype.char, dim):
            # a set outgrew the buffers in the middle of a match_next sequence: the resident set moves into the new buffers
            seq = hp["seq"]
            with torch.cuda.stream(stream):
                nb = seq[1] * stride
                bufs[seq[0]][:nb].copy_(hp["bufs"][seq[0]][:nb])
        hp = {"key": (device, dtype.char, dim), "cap": cap, "stream": stream, "sptr": stream.cuda_stream, "dstream": dstream,
              "dsptr": dstream.cuda_stream, "row": row, "stride": stride, "bufs": bufs, "buf_ptr": [b.data_ptr() for b in bufs],
              "seq": seq, "ring": ring, "pos": 0, "code": _norm_code(self._norm, dtype, self._engine), "ws": {}, "device": device}
        self._matcher["host_plan"] = hp
        return hp

    def _knn_host_submit(self, q: np.ndarray, t, ratio: float, seq: bool = False) -> "PendingMatch":
        """t: the train descriptors (host), or with seq=True nothing: the train set is then the query set of the previous
        seq call, which is still in device memory (match_next_async)."""
        if q.ndim != 2 or (not seq and (t.ndim != 2 or q.shape[1] != t.shape[1])):
            raise _bad_input("query and train descriptors must have the same width")
        if not seq and q.dtype != t.dtype:
            raise _bad_input("query and train descriptors must have the same type")
        if q.dtype not in (np.uint8, np.float32):
            raise _bad_input(f"unsupported descriptor dtype {q.dtype} (cv2 accepts CV_8U and CV_32F)")
        if not q.flags.c_contiguous:
            q = np.ascontiguousarray(q)
        if not seq and not t.flags.c_contiguous:
            t = np.ascontiguousarray(t)          # both stay referenced by the handle until the copies are done
        nq, dim = q.shape
        prev = None
        if seq:
            hp0 = self._matcher.get("host_plan")
            prev = hp0["seq"] if hp0 is not None else None
            if prev is not None and hp0["key"] != (_cuda_device(self._device), q.dtype.char, dim):
                raise _bad_input("match_next: descriptor type or width changed inside a sequence; call reset_sequence() first")
            nt = prev[1] if prev is not None else 0
        else:
            nt = t.shape[0]
        if nq == 0 and not seq:
            return PendingMatch.done(np.empty((0, 2), np.int32), np.empty((0, 2), np.float32), np.empty((0,), bool))
        # host in, host out: nothing depends on the caller's stream, so the call runs on a private stream and
        # overlaps whatever the caller has in flight (e.g. Mapper's frame upload + fused kernel)
        hp = self._host_plan(_cuda_device(self._device), q.dtype, dim, max(nq, 1), nt)
        L, sptr, row, stride, code = self._matcher["lib"], hp["sptr"], hp["row"], hp["stride"], hp["code"]
        ws = hp["ws"].get((nq, nt))
        if ws is None:
            with torch.cuda.stream(hp["stream"]):
                wt = self._workspace(L.mk_knn2_workspace_bytes(code, nq, nt, dim), hp["device"], "workspace_host_path")
            ws = hp["ws"][(nq, nt)] = (wt.data_ptr(), wt.numel())
            if len(hp["ws"]) > 64:
                hp["ws"] = {(nq, nt): ws}
        elif self._matcher["workspace_host_path"].data_ptr() != ws[0]:      # the workspace grew since: refresh
            wt = self._matcher["workspace_host_path"]
            ws = hp["ws"][(nq, nt)] = (wt.data_ptr(), wt.numel())
        # packed outputs: idx (nq*8) | dist (nq*8) | keep (nq) | pad | n_keep (4)
        off_keep, off_n = nq * 16, (nq * 17 + 3) & ~3
        total = off_n + 4
        # device buffers are reused call after call (stream order protects them); the page-locked result buffer is the
        # only thing a later call could overwrite under the reader, so it comes from a small ring
        k = hp["pos"] = (hp["pos"] + 1) % PendingMatch.DEPTH
        slot = hp["ring"][k]
        if slot["owner"] is not None:
            slot["owner"]._resolve()             # an uncollected older call still points at this slot
        base = slot["out_ptr"]
        if seq:      # the previous query set stays where it is and becomes the train set; the new set goes into the other buffer
            qi = 1 - prev[0] if prev is not None else 0
            q_ptr, t_ptr = hp["buf_ptr"][qi], hp["buf_ptr"][1 - qi]
            hp["seq"] = (qi, nq)
        else:
            q_ptr, t_ptr = hp["buf_ptr"]
            hp["seq"] = None
        with _current_device(hp["device"]):
            if nq:
                _lib.check(L.mk_upload_2d(q_ptr, stride, q.__array_interface__["data"][0], row, row, nq, sptr))
            if seq and (prev is None or nq == 0):
                # first set of a sequence (nothing to match against yet) or an empty set: the upload is all there is to do
                if nq:
                    slot["event"].record(hp["stream"])
                    slot["event"].synchronize()      # once per sequence: the caller may reuse `q` as soon as this returns
                pending = PendingMatch.done(np.full((nq, 2), -1, np.int32), np.full((nq, 2), np.finfo(np.float32).max, np.float32),
                                            np.zeros((nq,), bool))
                return pending
            if nt and not seq:
                _lib.check(L.mk_upload_2d(t_ptr, stride, t.__array_interface__["data"][0], row, row, nt, sptr))
            _lib.check(L.mk_knn2(code, q_ptr, nq, t_ptr if nt else None, nt, dim, stride, base, base + nq * 8,
                                 float(ratio), base + off_keep, base + off_n, ws[0], ws[1], sptr))
            slot["kdone"].record(hp["stream"])
            hp["dstream"].wait_event(slot["kdone"])
            _lib.check(L.mk_download_2d(slot["host_ptr"], total, base, total, total, 1, hp["dsptr"]))
        slot["event"].record(hp["dstream"])
        pending = PendingMatch(slot, nq, off_keep, (q, t))
        slot["owner"] = pending
        return pending

    def upload_sets(self, sets: Sequence[np.ndarray]) -> "DescriptorSets":
        """Concatenate descriptor arrays of equal width / dtype into ONE device matrix (rows padded to 16 bytes) plus the
        set-offset table mk_knn2_batched indexes it with.  Pass the handle to knn_match_sets as often as needed: all-pairs
        matching uploads the keyframe set once and then only sends pair lists."""
        device = _cuda_device(self._device)
        arrs = [np.ascontiguousarray(a) for a in sets]
        if not arrs:
            raise _bad_input("no descriptor sets")
        dtype, dim = arrs[0].dtype, arrs[0].shape[1]
        if any(a.dtype != dtype or a.ndim != 2 or a.shape[1] != dim for a in arrs):
            raise _bad_input("all descriptor sets must share dtype and width")
        sizes = np.array([a.shape[0] for a in arrs], np.int64)
        total = int(sizes.sum())
        desc, _ = upload_descriptors(np.concatenate(arrs) if total else np.zeros((0, dim), dtype), device)
        set_off = torch.from_numpy(np.concatenate([[0], np.cumsum(sizes)]).astype(np.int32)).to(device)
        return DescriptorSets(desc, set_off, sizes, dtype, dim, total)

    def knn_match_sets(self, sets, pairs: Sequence[tuple[int, int]], ratio: float = 0.7, output: str = "host"):
        """Many (query set, train set) pairs in ONE launch: the consecutive-frame loop of
        SequentialMosaic.match_features (mosaic.py:402-438: query = frame k+1, train = frame k) or all-pairs loop closure.
        sets: descriptor arrays of equal width / dtype, or the handle upload_sets returned; pairs: (query index, train
        index) into `sets`.
        output="host"   -> list of (idx [nq,2] int32, dist [nq,2] float32, keep [nq] bool) per pair (= knn_match_arrays);
        output="device" -> (idx, dist, keep, out_off): device tensors over all pairs' query rows, rows of pair p at
                           out_off[p]:out_off[p+1] (host int32 array);
        output="counts" -> int32 array, ratio-test survivors per pair (only 4 bytes per pair leave the device)."""
        if output not in ("host", "device", "counts"):
            raise ValueError("output must be 'host', 'device' or 'counts'")
        device = _cuda_device(self._device)
        if not len(pairs):
            return [] if output == "host" else (np.zeros((0,), np.int32) if output == "counts" else None)
        ds = sets if isinstance(sets, DescriptorSets) else self.upload_sets(sets)
        desc, set_off, sizes, dtype, dim, total = ds.desc, ds.set_off, ds.sizes, ds.dtype, ds.dim, ds.total
        code = _norm_code(self._norm, dtype, self._engine)
        pr = np.asarray(pairs, np.int32).reshape(-1, 2)
        nq_per = sizes[pr[:, 0]]
        out_off_h = np.concatenate([[0], np.cumsum(nq_per)]).astype(np.int32)
        rows = int(out_off_h[-1])
        max_nq, max_nt = int(nq_per.max()), int(sizes[pr[:, 1]].max())
        pairs_d, out_off = torch.from_numpy(pr).to(device), torch.from_numpy(out_off_h).to(device)
        idx = torch.empty((rows, 2), dtype=torch.int32, device=device)
        dist = torch.empty((rows, 2), dtype=torch.float32, device=device)
        keep = torch.empty((rows,), dtype=torch.uint8, device=device)
        n_keep = torch.empty((len(pr),), dtype=torch.int32, device=device) if output == "counts" else None
        L = self._matcher["lib"]
        CH = 65535                                   # grid.z limit per launch
        for lo in range(0, len(pr), CH):
            hi = min(len(pr), lo + CH)
            ws = self._workspace(L.mk_knn2_batched_workspace_bytes(code, hi - lo, max_nq, max_nt, dim, total), device)
            stride = desc.stride(0) * desc.element_size()
            with _current_device(device):
                _lib.check(L.mk_knn2_batched(code, desc.data_ptr(), total, dim, stride, set_off.data_ptr(),
                                             pairs_d.data_ptr() + lo * 8, hi - lo, out_off.data_ptr() + lo * 4, max_nq, max_nt,
                                             idx.data_ptr(), dist.data_ptr(), float(ratio), keep.data_ptr(),
                                             n_keep.data_ptr() + lo * 4 if n_keep is not None else None,
                                             ws.data_ptr(), ws.numel(), torch.cuda.current_stream(device).cuda_stream))
        if output == "counts":
            return n_keep.cpu().numpy()
        if output == "device":
            return idx, dist, keep, out_off_h
        idx_h, dist_h, keep_h = idx.cpu().numpy(), dist.cpu().numpy(), keep.cpu().numpy().astype(bool)
        out = []
        for p in range(len(pr)):
            sl = slice(out_off_h[p], out_off_h[p + 1])
            out.append((idx_h[sl], dist_h[sl], keep_h[sl]))
        return out

    # -- reference API ---------------------------------------------------------------------------
    def knn_match(self, query: np.ndarray, train: np.ndarray, mask: np.ndarray = None) -> Sequence[Sequence[DMatch]]:
        """registration.py:233-242: cv2.BFMatcher.knnMatch(query, train, k=2, mask=mask)."""
        if mask is not None:
            raise NotImplementedError("match masks are not on the accelerated path (the reference passes mask=None)")
        if len(query) == 0:
            return ()
        idx, dist, _ = self.knn_match_arrays(query, train, ratio=0.0)
        if idx.shape[1] == 2 and idx.min() >= 0:
            # every row has its two neighbours (Nt >= 2, the hot path): build the 2 * Nq objects with C-level loops
            # (map over columns, zip into pairs) -- a third cheaper than nested generator expressions; what is left is
            # cv2.DMatch's own constructor
            q = range(idx.shape[0])
            m0 = map(DMatch, q, idx[:, 0].tolist(), repeat(0), dist[:, 0].tolist())
            m1 = map(DMatch, q, idx[:, 1].tolist(), repeat(0), dist[:, 1].tolist())
            return tuple(zip(m0, m1))
        il, dl = idx.tolist(), dist.tolist()      # plain Python numbers: far cheaper than numpy scalar indexing
        return tuple(tuple(DMatch(i, j, 0, d) for j, d in zip(ir, dr) if j >= 0) for i, (ir, dr) in enumerate(zip(il, dl)))

    def match(self, query: np.ndarray, train: np.ndarray, mask: np.ndarray = None) -> Sequence[DMatch]:
        """registration.py:244-253: best match per query row (rows without any match are dropped)."""
        if mask is not None:
            raise NotImplementedError("match masks are not on the accelerated path")
        if len(query) == 0:
            return ()
        idx, dist, _ = self.knn_match_arrays(query, train, ratio=0.0)
        return tuple(DMatch(i, int(idx[i, 0]), 0, float(dist[i, 0])) for i in range(idx.shape[0]) if idx[i, 0] >= 0)

    def good_matches(self, query, train, ratio: float = 0.7) -> Sequence[DMatch]:
        """knn_match + the Lowe ratio test of mosaic.py:430, materialising only the survivors."""
        if len(query) == 0:
            return ()
        idx, dist, keep = self.knn_match_arrays(query, train, ratio)
        return tuple(DMatch(int(i), int(idx[i, 0]), 0, float(dist[i, 0])) for i in np.flatnonzero(keep))


class CompositeMatcher(Matcher):
    """Per-feature-type matching (reference: registration.py:266-298)."""

    def knn_match(self, query: dict, train: dict, mask: np.ndarray = None) -> dict:
        if mask is not None:
            warnings.warn("mask variable unused.")
        matches = dict()
        for feature_type in query.keys():
            if query[feature_type] is None or train[feature_type] is None:
                matches.update({feature_type: tuple()})
                continue
            matches.update({feature_type: super().knn_match(query[feature_type], train[feature_type])})
        return matches

    def match(self, query: dict, train: dict, mask: np.ndarray = None) -> dict:
        if mask is not None:
            warnings.warn("mask variable unused.")
        matches = dict()
        for feature_type in query:
            matches.update({feature_type: super().match(query[feature_type], train[feature_type])})
        return matches

    def good_matches(self, query: dict, train: dict, ratio: float = 0.7) -> dict:
        out = dict()
        for feature_type in query.keys():
            if query[feature_type] is None or train[feature_type] is None:
                out[feature_type] = tuple()
            else:
                out[feature_type] = super().good_matches(query[feature_type], train[feature_type], ratio)
        return out


def get_matches(descriptors1, descriptors2, matcher: Matcher, minmatches: int):
    """registration.py:300-319 with its evident intent (k=2 + ratio 0.7); the reference helper calls
    `matcher.match` but unpacks pairs, so it only ever worked with a knn matcher."""
    if not isinstance(descriptors1, (tuple, list)):
        descriptors1 = [descriptors1]
    if not isinstance(descriptors2, (tuple, list)):
        descriptors2 = [descriptors2]
    mlength = nlength = 0
    good = []
    for d1, d2 in zip(descriptors1, descriptors2):
        for m in Matcher.good_matches(matcher, d1, d2, 0.7):
            good.append(DMatch(m.queryIdx + mlength, m.trainIdx + nlength, 0, m.distance))
        mlength += d1.shape[0]
        nlength += d2.shape[0]
    return minmatches <= len(good), good


# ------------------------------------------------------------------------------------------------
# registration glue on the device
# ------------------------------------------------------------------------------------------------
from enum import Enum  # noqa: E402


class HomographyEstimationType(Enum):
    """registration.py:383-398."""
    PERSPECTIVE = "perspective"
    AFFINE = "affine"
    PARTIAL = "partial"

    @staticmethod
    def from_string(s: str) -> "HomographyEstimationType":
        try:
            return HomographyEstimationType(s.lower())
        except ValueError:
            raise ValueError(f"Unknown homography type: {s.lower()}")

    def __str__(self) -> str:
        return self.value


class HomographyEstimation:
    """Drop-in for registration.HomographyEstimation (registration.py:399-419): RANSAC model fit on the GPU
    (csrc/mk_ransac.cu) instead of cv2.findHomography / cv2.estimateAffine2D.

    Called like the reference does (mosaic.py:459): est(kp, kp_prev, method=cv2.RANSAC, ransacReprojThreshold=1.0)
    -> (H 3x3 float64, inliers uint8 [n, 1]).  PERSPECTIVE fits 8 degrees of freedom; AFFINE and PARTIAL both call
    cv2.estimateAffine2D in the reference (6 degrees of freedom; PARTIAL drops the caller's arguments, i.e. threshold 3.0),
    which is reproduced.  The estimator is randomised, like cv2's: same inlier set up to points within rounding of the
    threshold, same model up to fit noise; deterministic for a given `seed`."""

    def __init__(self, method, device=None, iters: int = 2048, seed: int = 0x5EED):
        if isinstance(method, HomographyEstimationType):
            self.method = method
        elif isinstance(method, str):
            self.method = HomographyEstimationType.from_string(method)
        else:
            raise ValueError("method must be either a HomographyEstimationType or a string")
        self._device = _cuda_device(device)
        self._lib = _lib.lib()
        self.iters, self.seed = int(iters), int(seed)
        self._ws = None

    def _threshold(self, kwargs) -> float:
        if self.method == HomographyEstimationType.PARTIAL:
            return 3.0                                   # registration.py:415 calls estimateAffine2D(src, dst) with cv2's defaults
        return float(kwargs.get("ransacReprojThreshold", 3.0))

    def estimate_device(self, src: torch.Tensor, dst: torch.Tensor, threshold: float):
        """src / dst: float32 [n, 2] on the device -> (H float64 [3, 3], inliers uint8 [n], n_inliers int32 [1]), all on the
        device, asynchronous on the current stream."""
        dev = self._device
        src = src.to(dev, torch.float32).contiguous(); dst = dst.to(dev, torch.float32).contiguous()
        n = int(src.shape[0])
        if self._ws is None:
            self._ws = torch.empty(int(self._lib.mk_ransac_workspace_bytes()), dtype=torch.uint8, device=dev)
        H = torch.empty((3, 3), dtype=torch.float64, device=dev)
        inl = torch.empty((max(n, 1),), dtype=torch.uint8, device=dev)
        cnt = torch.empty((1,), dtype=torch.int32, device=dev)
        model = _lib.MODEL_PERSPECTIVE if self.method == HomographyEstimationType.PERSPECTIVE else _lib.MODEL_AFFINE
        with _current_device(dev):
            _lib.check(self._lib.mk_ransac_transform(model, src.data_ptr(), dst.data_ptr(), n, None, float(threshold), self.iters, self.seed,
                                                     H.data_ptr(), inl.data_ptr(), cnt.data_ptr(), self._ws.data_ptr(), _raw_stream(dev)))
        return H, inl[:n], cnt

    def __call__(self, src_points, dst_points, *args, **kwargs):
        src = torch.from_numpy(np.ascontiguousarray(np.asarray(src_points, np.float32).reshape(-1, 2)))
        dst = torch.from_numpy(np.ascontiguousarray(np.asarray(dst_points, np.float32).reshape(-1, 2)))
        H, inl, cnt = self.estimate_device(src, dst, self._threshold(kwargs))
        if int(cnt.item()) == 0:
            return None, None                                        # what cv2 returns when no model is found
        return H.cpu().numpy(), inl.cpu().numpy().reshape(-1, 1)


def gather_matched_keypoints(kp_query: torch.Tensor, kp_train: torch.Tensor, idx: torch.Tensor, keep: torch.Tensor):
    """mosaic.py:449-457 on the device: keypoints of the ratio-test survivors in match order.  kp_*: float32 [n, 2], idx int32
    [nq, 2] and keep uint8 [nq] as Matcher.knn_match_device returns them -> (src [nq, 2], dst [nq, 2], count int32 [1]); only the
    first `count` rows are meaningful.  No host round trip."""
    dev = kp_query.device
    nq = int(idx.shape[0])
    src = torch.empty((max(nq, 1), 2), dtype=torch.float32, device=dev)
    dst = torch.empty((max(nq, 1), 2), dtype=torch.float32, device=dev)
    cnt = torch.empty((1,), dtype=torch.int32, device=dev)
    with _current_device(dev):
        _lib.check(_lib.lib().mk_gather_matches(kp_query.contiguous().data_ptr(), kp_train.contiguous().data_ptr(), idx.contiguous().data_ptr(),
                                                keep.contiguous().data_ptr(), nq, src.data_ptr(), dst.data_ptr(), None, cnt.data_ptr(), _raw_stream(dev)))
    return src, dst, cnt


def propagate_homographies(H_seq, device=None) -> np.ndarray:
    """core/interface.py:97-113 (_propagate_homographies) for an ordered chain: [H0, H1, ...] -> accumulate(matmul), float64,
    computed on the device (one launch, left to right like itertools.accumulate)."""
    dev = _cuda_device(device)
    H = torch.from_numpy(np.ascontiguousarray(np.asarray(H_seq, np.float64).reshape(-1, 3, 3))).to(dev)
    out = torch.empty_like(H)
    with _current_device(dev):
        _lib.check(_lib.lib().mk_chain_homographies(H.data_ptr(), int(H.shape[0]), out.data_ptr(), _raw_stream(dev)))
    return out.cpu().numpy()
