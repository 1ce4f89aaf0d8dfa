"""Tile-sharded mosaic canvas over the GPUs of one box (one process per GPU, torch.distributed).

Mirrors how the reference decomposes a large mosaic: SequentialMosaic._create_tile_graph cuts the
canvas into a regular grid of tiles (src/mosaicking/mosaic.py:595-604), a frame is assigned to every
tile its warped corner quad overlaps (get_corners + registration.bbox_overlap on int-truncated
corners, mosaic.py:612-614), each tile is rendered by its own Mapper with the tile-local homography
H_tile = T(-tx, -ty) @ H (mosaic.py:617), and examples/combine_tiles.py:46-62 pastes the tiles back
together.  Tiles never read each other, so the path shards with NO data-path collective between
tiles; every tile is eroded independently exactly like the reference's per-tile Mapper, hence no
seam (halo) exchange is needed for reference-tile semantics.

What does cross GPUs:
  * the per-frame homographies: all_gather of 9 doubles per rank and step, so that every rank can
    derive the same routing table;
  * the frames themselves, only towards the ranks that own a tile the footprint overlaps
    (batched isend / irecv = ncclSend / ncclRecv over NVLink);
  * the final tile gather (= combine_tiles).
Each rank ingests its own frame stream ("camera"), so per-GPU work is constant as ranks are added.
"""
from __future__ import annotations

import os

from typing import Callable, Sequence

import numpy as np
import torch
import torch.distributed as dist

from .mosaic import bbox_overlap, get_corners, homogeneous_translation


def tiles_of_footprint(H: np.ndarray, width: int, height: int, tile_size: tuple[int, int], grid: tuple[int, int]):
    """Tiles (xi, yi) of a grid[0] x grid[1] tile grid that the warped frame overlaps -- the test of
    mosaic.py:612-614 restricted to the tiles the corner bounding box can reach."""
    crn = get_corners(H, width, height).squeeze().astype(int)
    tw, th = tile_size
    # one extra tile on each side: closed polygons that merely touch a tile edge still intersect it
    x0 = max(0, int(np.floor(crn[:, 0].min() / tw)) - 1); x1 = min(grid[0] - 1, int(np.floor(crn[:, 0].max() / tw)) + 1)
    y0 = max(0, int(np.floor(crn[:, 1].min() / th)) - 1); y1 = min(grid[1] - 1, int(np.floor(crn[:, 1].max() / th)) + 1)
    out = []
    for xi in range(x0, x1 + 1):
        for yi in range(y0, y1 + 1):
            tx, ty = xi * tw, yi * th
            tile_crns = np.array([[tx, ty], [tx + tw, ty], [tx + tw, ty + th], [tx, ty + th]], dtype=int)
            if bbox_overlap(tile_crns, crn):
                out.append((xi, yi))
    return out


class PeerFrames:
    """The frames of EVERY rank, readable from this rank without a copy: local tensors stay where they are, the other ranks'
    frames are mapped into this process through cudaIpc (mk_ipc_export / mk_ipc_open, include/mosaic_b200.h) and read by the
    compositing kernel over NVLink / NVSwitch -- it pulls only the source rectangles its tiles need.

    frames: this rank's frames, uint8 [h, w, 3] contiguous CUDA tensors (all ranks pass the same number).  Collective: every rank
    must construct it at the same time.  The frames must stay alive and unchanged until every rank has called close()."""

    def __init__(self, frames: Sequence[torch.Tensor], device, rank: int | None = None, world: int | None = None):
        import ctypes as C

        from . import _lib
        self._lib = _lib.lib()
        self.rank = dist.get_rank() if rank is None else rank
        self.world = dist.get_world_size() if world is None else world
        self.device = device
        self.frames = list(frames)
        for f in self.frames:
            assert f.is_cuda and f.dtype == torch.uint8 and f.dim() == 3 and f.shape[2] == 3 and f.is_contiguous()
        if device.type == "cuda":
            torch.cuda.synchronize(device)                     # the frames are complete before anybody else may read them
        mine = []
        for f in self.frames:
            h = (C.c_uint8 * 64)()
            off = C.c_uint64(0)
            _lib.check(self._lib.mk_ipc_export(f.data_ptr(), h, C.byref(off)))
            mine.append((bytes(h), int(off.value), int(f.shape[0]), int(f.shape[1])))
        if self.world == 1:
            everyone = [mine]
        else:
            everyone = [None] * self.world
            dist.all_gather_object(everyone, mine)            # 64-byte handles + offsets, once
        self._bases: dict[bytes, int] = {}
        self.ptr: list[list[tuple[int, int, int]]] = []        # [source rank][k] -> (device pointer, h, w)
        with torch.cuda.device(device):
            for r, lst in enumerate(everyone):
                row = []
                for k, (hb, off, fh, fw) in enumerate(lst):
                    if r == self.rank:
                        row.append((self.frames[k].data_ptr(), fh, fw))
                        continue
                    base = self._bases.get(hb)
                    if base is None:
                        out = C.c_void_p(0)
                        buf = (C.c_uint8 * 64).from_buffer_copy(hb)
                        _lib.check(self._lib.mk_ipc_open(buf, C.byref(out)))
                        base = self._bases[hb] = int(out.value)
                    row.append((base + off, fh, fw))
                self.ptr.append(row)
        if self.world > 1:
            dist.barrier()                                     # everyone has mapped everyone before anyone may finish and free

    def close(self) -> None:
        """Collective: unmap the peers' allocations (after every rank is done reading)."""
        if self.world > 1:
            if self.device.type == "cuda":
                torch.cuda.synchronize(self.device)
            dist.barrier()
        for base in self._bases.values():
            self._lib.mk_ipc_close(base)
        self._bases.clear()
        if self.world > 1:
            dist.barrier()          # nobody frees an exported frame while another rank still has it mapped


class ShardedMosaic:
    """Canvas of grid[0] x grid[1] tiles, tile (xi, yi) owned by rank (yi * grid[0] + xi) % world.

    mapper_factory(tile_w, tile_h) builds the per-tile compositor: mosaicking_b200.Mapper on the GPU
    path; the CPU tests inject an oracle-backed stand-in and run this class over gloo."""

    def __init__(self, grid: tuple[int, int], tile_size: tuple[int, int], mapper_factory: Callable, frame_shape: tuple[int, int],
                 device, rank: int | None = None, world: int | None = None):
        self.grid, self.tile_size = grid, tile_size
        self.rank = dist.get_rank() if rank is None else rank
        self.world = dist.get_world_size() if world is None else world
        self.device = device
        self.h, self.w = frame_shape
        self._factory = mapper_factory
        self.mappers: dict[tuple[int, int], object] = {}
        self._recv = {}          # (source rank, step parity) -> receive buffer
        self._posted = {}        # step -> exchange posted ahead of time (step_planned with next_frame)
        self._pc = None          # state of the copy-engine pulls (step_planned with peers and peer_copy=True)
        self.frames_applied = 0
        self.bytes_sent = 0
        self._host_group = None   # gloo group for the host-side pose exchange of the unplanned step()

    def owner(self, tile: tuple[int, int]) -> int:
        return (tile[1] * self.grid[0] + tile[0]) % self.world

    def owned_tiles(self) -> list[tuple[int, int]]:
        return [(xi, yi) for yi in range(self.grid[1]) for xi in range(self.grid[0]) if self.owner((xi, yi)) == self.rank]

    def _mapper(self, tile):
        if tile not in self.mappers:
            self.mappers[tile] = self._factory(*self.tile_size)
        return self.mappers[tile]

    def routing(self, H_all: Sequence[np.ndarray]):
        """-> per source rank: {owner rank: [tiles]} for the frame of that source."""
        table = []
        for s, H in enumerate(H_all):
            dests: dict[int, list] = {}
            if H is not None:
                for tile in tiles_of_footprint(H, self.w, self.h, self.tile_size, self.grid):
                    dests.setdefault(self.owner(tile), []).append(tile)
            table.append(dests)
        return table

    def exchange_homographies(self, H: np.ndarray | None) -> list:
        """all_gather of the 9 doubles of every rank's current frame (NaN = no frame this step)."""
        mine = torch.full((9,), float("nan"), dtype=torch.float64)
        if H is not None:
            mine = torch.from_numpy(np.ascontiguousarray(np.asarray(H, np.float64).reshape(9)))
        if self.world == 1:
            return [None if H is None else np.asarray(H, np.float64).reshape(3, 3)]
        # the poses are needed on the HOST (routing is host logic): exchange them over a host-side (gloo) group, so the
        # per-frame path neither copies them to the device and back nor synchronises the GPU
        if self._host_group is None and self.device.type == "cuda":
            try:
                self._host_group = dist.new_group(backend="gloo")
            except Exception:            # no gloo in this build: fall back to the device group (one small D2H per frame)
                self._host_group = False
        if self._host_group:
            out = [torch.empty_like(mine) for _ in range(self.world)]
            dist.all_gather(out, mine, group=self._host_group)
            allH = torch.stack(out).numpy()
        else:
            mine = mine.to(self.device)
            out = [torch.empty_like(mine) for _ in range(self.world)]
            dist.all_gather(out, mine)
            allH = torch.stack(out).cpu().numpy()
        return [None if np.isnan(a[0]) else a.reshape(3, 3) for a in allH]

    def connect(self) -> None:
        """Establish every rank-to-rank NCCL send/recv channel up front (they are otherwise created lazily by the
        first routed frame, which costs hundreds of milliseconds once)."""
        if self.world == 1:
            return
        token = torch.zeros(8, dtype=torch.uint8, device=self.device)
        sink = [torch.empty_like(token) for _ in range(self.world)]
        ops = []
        for peer in range(self.world):
            if peer != self.rank:
                ops.append(dist.P2POp(dist.isend, token, peer))
                ops.append(dist.P2POp(dist.irecv, sink[peer], peer))
        for req in dist.batch_isend_irecv(ops):
            req.wait()
        if self.device.type == "cuda":
            torch.cuda.synchronize(self.device)

    def plan(self, H_local: np.ndarray) -> None:
        """Exchange the homographies of ALL upcoming frames once (the reference also knows every pose
        before it renders: global_registration precedes generate, and _create_tile_graph routes every
        frame before the first Mapper.update, mosaic.py:571-622).  H_local: [K, 3, 3] for this rank."""
        H_local = np.ascontiguousarray(np.asarray(H_local, np.float64).reshape(-1, 9))
        if self.world == 1:
            allH = H_local[None]
        else:
            mine = torch.from_numpy(H_local).to(self.device)
            out = [torch.empty_like(mine) for _ in range(self.world)]
            dist.all_gather(out, mine)
            allH = torch.stack(out).cpu().numpy()
        self._plan_H = allH.reshape(self.world, -1, 3, 3)
        self._plan_table = [self.routing([self._plan_H[s, k] for s in range(self.world)]) for k in range(self._plan_H.shape[1])]

    def step_planned(self, k: int, frame: torch.Tensor | None, next_frame: torch.Tensor | None = None, peers: "PeerFrames | None" = None,
                     peer_copy: bool = False) -> int:
        """Step k of a planned sequence: no collective, only the frame send/recv the routing asks for.
        With `next_frame` (this rank's frame of step k+1, already on the device) the exchange of step k+1 is posted BEFORE
        step k is composited, so the NVLink transfer runs under the kernels of step k (receive buffers alternate by step
        parity).  Every rank must then pass next_frame for the same steps, and the next call must be step k+1.
        With `peers` (PeerFrames over every rank's frames of the planned sequence) nothing is sent at all: a frame that another
        rank ingested is composited straight out of that GPU's memory (the kernel's TMA loads cross NVLink), frame k of rank s
        being peers.ptr[s][k].
        With `peers` and peer_copy=True the frames other ranks ingested are PULLED whole by the copy engine instead
        (mk_copy_peer: cudaMemcpyAsync out of the mapped peer memory on a side stream, the pulls of steps k+1 and k+2 queued before
        step k is composited): no SM is taken from the persistent compositing kernel -- an NCCL send/recv kernel running next to it
        costs the step ~10 us -- and the kernel reads local memory.  The next call must then be step k+1."""
        H_all = [self._plan_H[s, k] for s in range(self.world)]
        if peers is not None and peer_copy:
            return self._composite_peer_copy(k, H_all, peers)
        if peers is not None:
            return self._composite_peer(k, H_all, self._plan_table[k], peers)
        posted = self._posted.pop(k, None)
        if posted is None:
            posted = self._post(frame, self._plan_table[k], k & 1)
        if next_frame is not None and k + 1 < len(self._plan_table):
            self._posted[k + 1] = self._post(next_frame, self._plan_table[k + 1], (k + 1) & 1)
        return self._composite(frame, H_all, self._plan_table[k], posted)

    def step(self, frame: torch.Tensor | None, H: np.ndarray | None, H_all: Sequence | None = None, table=None) -> int:
        """One step: this rank contributes `frame` [h, w, 3] uint8 (already on self.device) with absolute
        homography H (frame -> global canvas).  Frames are applied tile by tile in source-rank order so
        that every tile sees the same sequence on every run.  Returns the number of tile updates done."""
        if H_all is None:
            H_all = self.exchange_homographies(H)
        if table is None:
            table = self.routing(H_all)
        return self._composite(frame, H_all, table, self._post(frame, table, 0))

    def _post(self, frame, table, parity: int):
        """Start the sends / receives one step's routing table asks of this rank -> (requests, {source rank: buffer})."""
        ops, incoming = [], {}
        if self.world > 1:
            for s, dests in enumerate(table):
                if s == self.rank:
                    for d in dests:
                        if d != self.rank:
                            ops.append(dist.P2POp(dist.isend, frame, d))
                            self.bytes_sent += frame.numel()
                elif self.rank in dests:
                    buf = self._recv.get((s, parity))
                    if buf is None:
                        buf = torch.empty((self.h, self.w, 3), dtype=torch.uint8, device=self.device)
                        self._recv[(s, parity)] = buf
                    incoming[s] = buf
                    ops.append(dist.P2POp(dist.irecv, buf, s))
        return (dist.batch_isend_irecv(ops) if ops else []), incoming

    def _composite(self, frame, H_all, table, posted) -> int:
        reqs, incoming = posted
        waited = False
        done = 0
        for s, dests in enumerate(table):
            if self.rank not in dests:
                continue
            if s != self.rank and not waited:     # local tiles that come first in source order do not wait for the exchange
                for req in reqs:
                    req.wait()
                waited = True
            src = frame if s == self.rank else incoming[s]
            for (xi, yi) in dests[self.rank]:
                H_tile = homogeneous_translation(-xi * self.tile_size[0], -yi * self.tile_size[1]) @ H_all[s]   # mosaic.py:617
                self._mapper((xi, yi)).update_device(src, H_tile)
                done += 1
        if not waited:
            for req in reqs:
                req.wait()                        # sends must be complete before the caller reuses the frame
        self.frames_applied += done
        return done

    def _tile_tensor(self, tile) -> torch.Tensor:
        """Contiguous [tile_h, tile_w, 3] uint8 tensor of an owned tile on self.device."""
        m = self.mappers[tile]
        t = m.output_device if hasattr(m, "output_device") else torch.from_numpy(np.ascontiguousarray(m.output))
        return t.to(self.device).contiguous()

    def _composite_peer(self, k: int, H_all, table, peers: "PeerFrames") -> int:
        done = 0
        for s, dests in enumerate(table):
            if self.rank not in dests:
                continue
            ptr, fh, fw = peers.ptr[s][k]
            for (xi, yi) in dests[self.rank]:
                H_tile = homogeneous_translation(-xi * self.tile_size[0], -yi * self.tile_size[1]) @ H_all[s]   # mosaic.py:617
                self._mapper((xi, yi)).update_pointer(ptr, fh, fw, fw * 3, H_tile)
                done += 1
                if s != self.rank:
                    self.bytes_sent += fh * fw * 3        # upper bound of what crosses NVLink: only the footprint's source rectangles do
        self.frames_applied += done
        return done

    PULL_AHEAD = int(os.environ.get("MK_PULL_AHEAD", "2"))   # peer_copy: the pulls of step k + 2 are queued before step k is composited (a 6 MB pull over NVSwitch
                             # takes longer than one ~35 us step when all eight GPUs pull at once)

    def _peer_pull(self, k: int, peers: "PeerFrames") -> dict:
        """Queue the copy-engine pulls of the frames other ranks ingested and this rank composites at step k
        -> {source rank: (local buffer, event 'landed')}.  PULL_AHEAD + 1 buffers per source rank, by step modulo."""
        pc = self._pc
        if pc is None:
            from . import _lib
            st = torch.cuda.Stream(device=self.device)
            pc = self._pc = {"lib": _lib.lib(), "check": _lib.check, "stream": st, "sptr": st.cuda_stream, "bufs": {}, "landed": {}, "free": {},
                             "posted": {}}
        got = {}
        for s, dests in enumerate(self._plan_table[k]):
            if s == self.rank or self.rank not in dests:
                continue
            slot = k % (self.PULL_AHEAD + 1)
            key = (s, slot)
            buf = pc["bufs"].get(key)
            if buf is None:
                buf = pc["bufs"][key] = torch.empty((self.h, self.w, 3), dtype=torch.uint8, device=self.device)
                pc["landed"][key] = torch.cuda.Event()
            free = pc["free"].get(slot)
            if free is not None:
                pc["stream"].wait_event(free)          # the kernels of step k - (PULL_AHEAD + 1) have read this buffer
            ptr, fh, fw = peers.ptr[s][k]
            with torch.cuda.device(self.device):
                pc["check"](pc["lib"].mk_copy_peer(buf.data_ptr(), ptr, fh * fw * 3, pc["sptr"]))
            pc["landed"][key].record(pc["stream"])
            got[s] = (buf, pc["landed"][key])
            self.bytes_sent += fh * fw * 3
        return got

    def _composite_peer_copy(self, k: int, H_all, peers: "PeerFrames") -> int:
        if self._pc is None or k not in self._pc["posted"]:
            first = self._peer_pull(k, peers)
            self._pc["posted"][k] = first
        for j in range(k + 1, min(k + self.PULL_AHEAD, len(self._plan_table) - 1) + 1):
            if j not in self._pc["posted"]:
                self._pc["posted"][j] = self._peer_pull(j, peers)          # runs under the kernels of steps k .. j-1
        pulled = self._pc["posted"].pop(k)
        main = torch.cuda.current_stream(self.device)
        done = 0
        for s, dests in enumerate(self._plan_table[k]):
            if self.rank not in dests:
                continue
            if s != self.rank:
                main.wait_event(pulled[s][1])
            for (xi, yi) in dests[self.rank]:
                H_tile = homogeneous_translation(-xi * self.tile_size[0], -yi * self.tile_size[1]) @ H_all[s]   # mosaic.py:617
                if s == self.rank:
                    ptr, fh, fw = peers.ptr[s][k]
                    self._mapper((xi, yi)).update_pointer(ptr, fh, fw, fw * 3, H_tile)
                else:
                    self._mapper((xi, yi)).update_device(pulled[s][0], H_tile)
                done += 1
        slot = k % (self.PULL_AHEAD + 1)
        ev = self._pc["free"].get(slot)
        if ev is None:
            ev = self._pc["free"][slot] = torch.cuda.Event()
        ev.record(main)
        self.frames_applied += done
        return done

    def gather(self, dst: int = 0, out: np.ndarray | None = None) -> np.ndarray | None:
        """combine_tiles (examples/combine_tiles.py:46-62): every touched tile travels to rank `dst`, which pastes tile (x, y) at
        (x*tile_w, y*tile_h) of one host canvas.  Tiles nobody touched stay black, like missing PNGs.  The tiles move as
        device tensors (ncclSend / ncclRecv, one tile-sized staging buffer on `dst`), never through pickle; `out` may be a
        caller-provided (e.g. page-locked) [grid_h*tile_h, grid_w*tile_w, 3] uint8 array."""
        tw, th = self.tile_size
        mine = sorted(self.mappers.keys())
        if self.world == 1:
            owned = [mine]
        else:
            owned = [None] * self.world
            dist.all_gather_object(owned, mine)            # tile coordinates only (a few bytes)
        if self.rank != dst:
            for t in mine:
                dist.send(self._tile_tensor(t), dst)
            return None
        if out is None:
            out = np.zeros((self.grid[1] * th, self.grid[0] * tw, 3), np.uint8)
        else:
            assert out.shape == (self.grid[1] * th, self.grid[0] * tw, 3) and out.dtype == np.uint8
        stage = None
        for r, tiles in enumerate(owned):
            for (xi, yi) in tiles:
                if r == self.rank:
                    src = self._tile_tensor((xi, yi))
                else:
                    if stage is None:
                        stage = torch.empty((th, tw, 3), dtype=torch.uint8, device=self.device)
                    dist.recv(stage, r)
                    src = stage
                out[yi * th:(yi + 1) * th, xi * tw:(xi + 1) * tw] = src.cpu().numpy()
        return out


# ------------------------------------------------------------------------------------------------
# matching: every (query image, train image) pair is an independent unit (BASELINE configs[4])
# ------------------------------------------------------------------------------------------------
def all_pairs_blocks(n_images: int, block: int = 8) -> list[list[tuple[int, int]]]:
    """Upper-triangular pair list (query j > train i, the direction SequentialMosaic.match_features uses for consecutive
    frames, mosaic.py:402-438: query = later frame) cut into block x block squares of the pair matrix, so that the pairs
    of one unit of work share their descriptor sets (2 * block sets instead of block^2)."""
    blocks = []
    for bi in range(0, n_images, block):
        for bj in range(bi, n_images, block):
            pairs = [(j, i) for i in range(bi, min(bi + block, n_images)) for j in range(max(bj, i + 1), min(bj + block, n_images))]
            if pairs:
                blocks.append(pairs)
    return blocks


def all_pairs_schedule(n_images: int, world: int, rank: int, block: int = 8) -> list[tuple[int, int]]:
    """Pairs owned by `rank`: blocks are dealt out longest-first to the least-loaded rank (diagonal blocks hold fewer
    pairs), which keeps the per-rank pair counts within one block of each other.  Deterministic, identical on all ranks."""
    block = max(1, min(block, n_images // (2 * world)))          # few images: smaller squares, or one rank would own them all
    blocks = sorted(all_pairs_blocks(n_images, block), key=lambda b: (-len(b), b[0]))
    load = [0] * world
    mine: list[tuple[int, int]] = []
    for b in blocks:
        r = min(range(world), key=lambda x: (load[x], x))
        load[r] += len(b)
        if r == rank:
            mine.extend(b)
    return mine


class ShardedMatcher:
    """All-pairs descriptor matching over the GPUs of one box: descriptor sets replicated on every rank (2000 x 4000 SIFT
    rows as bf16 = 2 GB), the pair list partitioned by all_pairs_schedule, each rank matching its pairs with ONE batched
    launch per chunk (Matcher.knn_match_sets -> mk_knn2_batched), and one all_reduce of the [n, n] survivor-count matrix --
    the loop-closure graph needs "how many ratio-test survivors does pair (j, i) have", the match lists stay on the rank
    that computed them.  No data-path collective: the pairs are independent.

    matcher: anything with knn_match_sets(sets, pairs, ratio) (mosaicking_b200.Matcher on the GPU; the CPU tests inject
    an oracle-backed stand-in and run this over gloo)."""

    def __init__(self, matcher, device, rank: int | None = None, world: int | None = None, block: int = 8):
        self.matcher, self.device, self.block = matcher, device, block
        self.rank = dist.get_rank() if rank is None else rank
        self.world = dist.get_world_size() if world is None else world

    def match_all_pairs(self, sets: Sequence[np.ndarray], ratio: float = 0.7, chunk: int = 4096, keep_results: bool = False, handle=None):
        """-> (local, counts): counts [n, n] int64 on the host with counts[query, train] = number of ratio-test survivors,
        summed over ranks; local = {(query, train): (idx, dist, keep)} for the pairs this rank owns when keep_results is
        set (host copies of the match lists), else just the list of owned pairs.  The descriptor sets are uploaded once
        per call; without keep_results only 4 bytes per pair leave the device."""
        n = len(sets)
        mine = all_pairs_schedule(n, self.world, self.rank, self.block)
        local = {} if keep_results else list(mine)
        counts = torch.zeros((n, n), dtype=torch.int64)
        if handle is None:       # `handle` = Matcher.upload_sets(sets) from an earlier call: the sets are already resident
            handle = self.matcher.upload_sets(sets) if mine and hasattr(self.matcher, "upload_sets") else sets
        for lo in range(0, len(mine), chunk):
            part = mine[lo:lo + chunk]
            if keep_results:
                for (qj, ti), (idx, dist_, keep) in zip(part, self.matcher.knn_match_sets(handle, part, ratio)):
                    local[(qj, ti)] = (idx, dist_, keep)
                    counts[qj, ti] = int(np.count_nonzero(keep))
            else:
                pr = np.asarray(part, np.int64)
                counts[pr[:, 0], pr[:, 1]] = torch.from_numpy(self.matcher.knn_match_sets(handle, part, ratio, output="counts").astype(np.int64))
        if self.world > 1:
            c = counts.to(self.device)
            dist.all_reduce(c)
            counts = c.cpu()
        return local, counts.numpy()
