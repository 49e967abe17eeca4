"""Multi-GPU host logic: one process per GPU, ``torch.distributed`` for the plumbing only.

Two ways the path shards:

* frame sharding — frame pairs are independent; pair ``i`` goes to rank ``i % world``; no collective on
  the data path, results (KB-sized MotionEntry lists / verdicts) are gathered at the end;
* spatial tiling — one large frame is cut into horizontal strips aligned to block rows.  Blocks never
  straddle strips, so the current frame needs no halo; the previous frame needs ``search`` rows above and
  below each strip.  That is the one real exchange step: grouped send/recv of halo rows with the two
  neighbours (NCCL over NVLink on GPUs, gloo in the CPU tests), overlapped with the interior block rows,
  which do not depend on it.

The detector / densifier consume the gathered entry list in strip order (= raster order), which keeps
the reference's input order and therefore its f32 sums bit-exact; partial sums are never all-reduced.
The Almeida estimator does not shard (30 dependent iterations of 12 scalars): replicas only.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np


def shard_round_robin(n_items: int, rank: int, world: int) -> list[int]:
    """Indices of the items (frame pairs) owned by ``rank``."""
    return list(range(rank, n_items, world))


@dataclass(frozen=True)
class Strip:
    rank: int
    by0: int          # first block row
    nby: int          # block rows in the strip
    y0: int           # first pixel row  (= by0 * block)
    rows: int         # pixel rows of cur (= nby * block)
    own_rows: int     # prev rows stored by this rank (last rank also keeps the frame's remainder rows)
    halo_top: int     # prev rows needed from the rank above
    halo_bottom: int  # prev rows needed from the rank below (beyond own_rows)


def strip_plan(h: int, block: int, search: int, world: int) -> list[Strip]:
    """Cut ``h`` rows into ``world`` strips of whole block rows (first strips take the extra rows)."""
    nby = h // block
    if world < 1 or nby < world:
        raise ValueError(f"cannot cut {nby} block rows into {world} strips")
    base, extra = divmod(nby, world)
    strips, by0 = [], 0
    for r in range(world):
        n = base + (1 if r < extra else 0)
        y0, rows = by0 * block, n * block
        last = r == world - 1
        own = (h - y0) if last else rows
        halo_top = min(search, y0)
        halo_bottom = 0 if last else min(search, h - (y0 + rows))
        strips.append(Strip(r, by0, n, y0, rows, own, halo_top, halo_bottom))
        by0 += n
    for s in strips:
        above = strips[s.rank - 1].own_rows if s.rank > 0 else 0
        below = strips[s.rank + 1].own_rows if s.rank + 1 < world else 0
        if s.halo_top > above or s.halo_bottom > below:
            raise ValueError("search range exceeds a neighbouring strip: use fewer ranks or a smaller range")
    return strips


def exchange_halos(buf, plan: list[Strip], rank: int, stride: int, group=None):
    """Fill the halo rows of ``buf`` from the neighbours.

    ``buf`` is a 1-D uint8 torch tensor laid out as ``[halo_top | own_rows | halo_bottom]`` rows of ``stride``
    bytes; the own rows are already in place.  Returns the list of outstanding requests (wait on them before
    launching the edge block rows)."""
    import torch.distributed as dist
    s = plan[rank]
    ops = []
    own0 = s.halo_top * stride
    if rank > 0:
        up = plan[rank - 1]
        # my first rows are the upper neighbour's bottom halo
        if up.halo_bottom:
            ops.append(dist.P2POp(dist.isend, buf[own0:own0 + up.halo_bottom * stride], rank - 1, group))
        if s.halo_top:
            ops.append(dist.P2POp(dist.irecv, buf[0:own0], rank - 1, group))
    if rank + 1 < len(plan):
        dn = plan[rank + 1]
        end_own = own0 + s.own_rows * stride
        # my last rows are the lower neighbour's top halo
        if dn.halo_top:
            ops.append(dist.P2POp(dist.isend, buf[end_own - dn.halo_top * stride:end_own], rank + 1, group))
        if s.halo_bottom:
            ops.append(dist.P2POp(dist.irecv, buf[end_own:end_own + s.halo_bottom * stride], rank + 1, group))
    if not ops:
        return []
    return dist.batch_isend_irecv(ops)


class TiledBlockMatcher:
    """Block matching of one large frame pair tiled over ``world`` GPUs (one instance per rank).

    ``load(prev_frame, cur_frame)`` takes this rank's rows from host frames (numpy) — in production each
    rank would receive only its strip; ``run()`` exchanges halo rows of the previous frame with the
    neighbours over NCCL and launches the strip kernel: interior block rows first (they overlap the
    transfer), the edge block rows once the halos have landed."""

    def __init__(self, ctx, w: int, h: int, block: int, search: int, metric: int, rank: int, world: int,
                 group=None, overlap: bool = True):
        import torch
        self.torch = torch
        self.ctx, self.w, self.h, self.block, self.search, self.metric = ctx, w, h, block, search, metric
        self.rank, self.world, self.group, self.overlap = rank, world, group, overlap
        self.plan = strip_plan(h, block, search, world)
        s = self.strip = self.plan[rank]
        dev = torch.device("cuda", ctx.device)
        self.nbx = w // block
        self.prev = torch.zeros((s.halo_top + s.own_rows + s.halo_bottom) * w, dtype=torch.uint8, device=dev)
        self.cur = torch.zeros(s.rows * w, dtype=torch.uint8, device=dev)
        self.entries = torch.zeros((s.nby * self.nbx, 4), dtype=torch.float32, device=dev)
        self.mv = torch.zeros((s.nby * self.nbx, 2), dtype=torch.int16, device=dev)
        self.cost = torch.zeros(s.nby * self.nbx, dtype=torch.int32, device=dev)
        self.comm_stream = torch.cuda.Stream(device=dev)
        self.kernel_stream = torch.cuda.ExternalStream(ctx.get_stream(), device=dev)

    def load(self, prev_frame: np.ndarray, cur_frame: np.ndarray, fill_halos: bool = False):
        """Copy this rank's rows to the device.  ``fill_halos`` also copies the halo rows straight from
        the full frame (single-process emulation of the exchange, used by the 1-GPU tests)."""
        torch, s, w = self.torch, self.strip, self.w
        if fill_halos:
            r0, r1 = s.y0 - s.halo_top, s.y0 + s.own_rows + s.halo_bottom
            self.prev.copy_(torch.from_numpy(np.ascontiguousarray(prev_frame[r0:r1]).reshape(-1)))
        else:
            own = torch.from_numpy(np.ascontiguousarray(prev_frame[s.y0:s.y0 + s.own_rows]).reshape(-1))
            self.prev[s.halo_top * w:(s.halo_top + s.own_rows) * w].copy_(own)
        self.cur.copy_(torch.from_numpy(np.ascontiguousarray(cur_frame[s.y0:s.y0 + s.rows]).reshape(-1)))
        torch.cuda.synchronize()

    def _launch(self, br0: int, br1: int):
        """Block rows [br0, br1) of this strip."""
        if br1 <= br0:
            return
        s, w, B, R = self.strip, self.w, self.block, self.search
        y_first = br0 * B                                  # strip-relative pixel row
        rows = (br1 - br0) * B
        above = s.halo_top + y_first                       # prev rows available above this sub-strip
        below = s.own_rows + s.halo_bottom - (y_first + rows)
        prev_ptr = self.prev.data_ptr() + (s.halo_top + y_first) * w
        cur_ptr = self.cur.data_ptr() + y_first * w
        off = br0 * self.nbx
        self.ctx.block_match_strip_dev(prev_ptr, cur_ptr, w, rows, w, min(R, above), min(R, below), s.y0 + y_first,
                                       self.h, B, R, self.metric, self.mv.data_ptr() + off * 4,
                                       self.cost.data_ptr() + off * 4, self.entries.data_ptr() + off * 16)

    def run(self, exchange: bool = True):
        torch, s, B, R = self.torch, self.strip, self.block, self.search
        top_rows = min(s.nby, -(-R // B)) if s.halo_top else 0          # block rows touching the top halo
        bot_rows = min(s.nby - top_rows, -(-R // B)) if s.halo_bottom else 0
        if self.world == 1:
            self._launch(0, s.nby)
            return
        if not exchange:                                                # halos already in place (emulation)
            self._launch(top_rows, s.nby - bot_rows)
            self._launch(0, top_rows)
            self._launch(s.nby - bot_rows, s.nby)
            return
        self.comm_stream.wait_stream(self.kernel_stream)
        with torch.cuda.stream(self.comm_stream):
            reqs = exchange_halos(self.prev, self.plan, self.rank, self.w, self.group)
            for r in reqs:
                r.wait()
        if self.overlap:
            self._launch(top_rows, s.nby - bot_rows)                    # interior: independent of the halos
            self.kernel_stream.wait_stream(self.comm_stream)
            self._launch(0, top_rows)
            self._launch(s.nby - bot_rows, s.nby)
        else:
            self.kernel_stream.wait_stream(self.comm_stream)
            self._launch(0, s.nby)

    def gather_entries(self):
        """All ranks' MotionEntry lists in strip (= raster) order, on every rank, as a numpy array."""
        import torch.distributed as dist
        torch = self.torch
        self.ctx.sync()
        if self.world == 1:
            return self.entries.cpu().numpy()
        max_n = max(p.nby for p in self.plan) * self.nbx
        pad = torch.zeros((max_n, 4), dtype=torch.float32, device=self.entries.device)
        pad[:self.entries.shape[0]] = self.entries
        out = [torch.empty_like(pad) for _ in range(self.world)]
        dist.all_gather(out, pad, group=self.group)
        return np.concatenate([o[:p.nby * self.nbx].cpu().numpy() for o, p in zip(out, self.plan)])


class TiledStreamMatcher:
    """Spatial tiling of a STREAM of large frames: rank r keeps its strip (plus halo rows) of all n_frames
    frames; the halo rows of every frame are exchanged with the two neighbours in ONE grouped NCCL
    send/recv per step, then a single batched strip launch matches all n_frames-1 pairs.  The exchange
    latency that dominates a single tiled pair is amortised over the batch.

    Buffer layout per frame: [halo_top | own_rows | halo_bottom] rows of w bytes; pair i = frames (i, i+1),
    so cur = prev + one frame buffer (the "stream" layout of ofpsb_block_match_strip_batch_dev)."""

    def __init__(self, ctx, w: int, h: int, block: int, search: int, metric: int, n_frames: int, rank: int, world: int,
                 group=None):
        import torch
        self.torch = torch
        self.ctx, self.w, self.h, self.block, self.search, self.metric = ctx, w, h, block, search, metric
        self.rank, self.world, self.group, self.n_frames = rank, world, group, n_frames
        self.plan = strip_plan(h, block, search, world)
        s = self.strip = self.plan[rank]
        dev = torch.device("cuda", ctx.device)
        self.nbx = w // block
        self.rows = s.halo_top + s.own_rows + s.halo_bottom
        self.frames = torch.zeros((n_frames, self.rows, w), dtype=torch.uint8, device=dev)
        nb = s.nby * self.nbx
        self.entries = torch.zeros((n_frames - 1, nb, 4), dtype=torch.float32, device=dev)
        self.comm_stream = torch.cuda.Stream(device=dev)
        self.kernel_stream = torch.cuda.ExternalStream(ctx.get_stream(), device=dev)

    def load(self, frames: np.ndarray, fill_halos: bool = False):
        """frames: [n_frames, h, w] uint8 on the host (every rank takes its rows)."""
        torch, s = self.torch, self.strip
        if fill_halos:
            r0, r1 = s.y0 - s.halo_top, s.y0 + s.own_rows + s.halo_bottom
            self.frames.copy_(torch.from_numpy(np.ascontiguousarray(frames[:, r0:r1])))
        else:
            own = torch.from_numpy(np.ascontiguousarray(frames[:, s.y0:s.y0 + s.own_rows]))
            self.frames[:, s.halo_top:s.halo_top + s.own_rows].copy_(own)
        torch.cuda.synchronize()

    def exchange(self):
        """One grouped send/recv with each neighbour for the halo rows of ALL frames."""
        import torch.distributed as dist
        s, plan, rank = self.strip, self.plan, self.rank
        ops, recvs = [], []
        own0, own1 = s.halo_top, s.halo_top + s.own_rows
        if rank > 0:
            up = plan[rank - 1]
            if up.halo_bottom:
                ops.append(dist.P2POp(dist.isend, self.frames[:, own0:own0 + up.halo_bottom].contiguous(), rank - 1, self.group))
            if s.halo_top:
                buf = self.torch.empty((self.n_frames, s.halo_top, self.w), dtype=self.torch.uint8, device=self.frames.device)
                ops.append(dist.P2POp(dist.irecv, buf, rank - 1, self.group))
                recvs.append((buf, 0, s.halo_top))
        if rank + 1 < len(plan):
            dn = plan[rank + 1]
            if dn.halo_top:
                ops.append(dist.P2POp(dist.isend, self.frames[:, own1 - dn.halo_top:own1].contiguous(), rank + 1, self.group))
            if s.halo_bottom:
                buf = self.torch.empty((self.n_frames, s.halo_bottom, self.w), dtype=self.torch.uint8, device=self.frames.device)
                ops.append(dist.P2POp(dist.irecv, buf, rank + 1, self.group))
                recvs.append((buf, own1, own1 + s.halo_bottom))
        if ops:
            for r in dist.batch_isend_irecv(ops):
                r.wait()
        for buf, a, b in recvs:
            self.frames[:, a:b].copy_(buf)

    def match(self):
        s, w = self.strip, self.w
        fb = self.rows * w
        prev = self.frames.data_ptr() + s.halo_top * w
        # rows below the last block row that this rank stores (the frame's remainder on the last rank) are valid
        # search rows, exactly as in TiledBlockMatcher._launch and in the whole-frame rule
        below = min(self.search, s.halo_bottom + s.own_rows - s.rows)
        self.ctx.block_match_strip_batch_dev(prev, prev + fb, w, s.rows, w, fb, self.n_frames - 1, s.halo_top, below,
                                             s.y0, self.h, self.block, self.search, self.metric, None, None,
                                             self.entries.data_ptr())

    def run(self, exchange: bool = True):
        torch = self.torch
        if exchange and self.world > 1:
            self.comm_stream.wait_stream(self.kernel_stream)
            with torch.cuda.stream(self.comm_stream):
                self.exchange()
            self.kernel_stream.wait_stream(self.comm_stream)
        self.match()


class PeerTiledMatcher:
    """Spatial tiling with NO exchange step (ofpsb_tiled_*, csrc/tiled.cu): every rank maps its neighbours' frame
    buffers (CUDA IPC) and the matching kernel reads the halo rows of the previous frame straight from their HBM.
    ``torch.distributed`` only carries the 128-byte handles once, at set-up; per pair there is no collective."""

    def __init__(self, ctx, w: int, h: int, block: int, search: int, rank: int, world: int, n_slots: int = 2, group=None):
        import torch
        import torch.distributed as dist
        from . import capi
        self.torch, self.ctx, self.rank, self.world, self.group = torch, ctx, rank, world, group
        self.t = capi.Tiled(ctx, rank, world, w, h, block, search, n_slots)
        self.w, self.h = w, h
        if world > 1:
            blobs = [None] * world
            dist.all_gather_object(blobs, self.t.export(), group=group)
            self.t.connect(blobs[rank - 1] if rank > 0 else None, blobs[rank + 1] if rank + 1 < world else None)
            dist.barrier(group=group)
        dev = torch.device("cuda", ctx.device)
        self.entries = torch.zeros((self.t.n_blocks, 4), dtype=torch.float32, device=dev)

    def load(self, slot: int, frame: np.ndarray):
        """This rank's rows of a whole host frame -> slot, then tell the neighbours."""
        t = self.t
        t.upload(slot, frame[t.y0:t.y0 + t.own_rows])
        t.publish(slot)

    def match(self, prev_slot: int = 0, cur_slot: int = 1, wait: bool = True):
        self.t.match(prev_slot, cur_slot, self.entries.data_ptr(), wait=wait)

    def gather_entries(self):
        """All ranks' MotionEntry lists in strip (= raster) order, on every rank, as a numpy array."""
        import torch.distributed as dist
        torch = self.torch
        self.ctx.sync()
        if self.world == 1:
            return self.entries.cpu().numpy()
        counts = [None] * self.world
        dist.all_gather_object(counts, self.t.n_blocks, group=self.group)
        pad = torch.zeros((max(counts), 4), dtype=torch.float32, device=self.entries.device)
        pad[:self.t.n_blocks] = self.entries
        out = [torch.empty_like(pad) for _ in range(self.world)]
        dist.all_gather(out, pad, group=self.group)
        return np.concatenate([o[:n].cpu().numpy() for o, n in zip(out, counts)])

    def close(self):
        self.t.close()
