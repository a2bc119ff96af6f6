"""The one exchange step of the multi-GPU path (one process per GPU): block rows -> primary.

The reference merges inside one process: every worker calls `primary.MergeOutput(self, blockReq)` and the
`aggregateAccumulator` kernel reads the peer device's buffer through a shared OpenCL context
(tracer/opencl/tracer.go:279-286, resources.go:108-124, renderer/default.go:188-191).  With one process
per GPU the same data movement is a gather of each rank's block rows (16 B per pixel, rows
[BlockY, BlockY+BlockH) only) to rank 0, which adds them into its frame accumulator with
`pc_merge_rows`.  It is a grouped send/recv, not a reduce: outside its own rows a rank's accumulator is
zero, so an all-reduce would move world_size times more bytes.

Two transports for that one step:
  * `IpcRowExchange` (default on the B200 box): every rank exports two frame-sized buffers through CUDA IPC
    (`pc_ipc_export`), rank 0 maps them once (`pc_ipc_open`); per pass a rank publishes its block rows into slot
    pass % 2 and rank 0's `k_merge` LOADS them from the peer GPU over NVLink while adding -- the transfer is the add,
    there is no staging copy and no bulk collective.  This is the reference's shared-context merge
    (device/context.go:11-28) across process boundaries.
  * `RowGather`: grouped send/recv of the rows (NCCL on GPUs, gloo in the CPU tests), then `pc_merge_rows`.
Hand-shakes ride on the scheduler's stats all-gather (`StatsExchange`), the only collective of the data path.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


class RowGather:
    """One exchange in flight.  Split phase so that the next pass can be traced while the rows travel: `start` snapshots
    this rank's block rows (the tracer clears its accumulator at the start of the next `Trace`, tracer.go:215) and posts
    the sends / receives, `finish` waits and hands rank 0 the per-rank row tensors."""

    def __init__(self, rows, frame_w: int, rank: int, world: int):
        self.rows = [int(r) for r in rows]
        self.frame_w, self.rank, self.world = frame_w, rank, world
        self.works, self.blocks, self._keep = [], None, None

    def start(self, mine: torch.Tensor, recv: torch.Tensor | None = None, snapshot: bool = True):
        rows, w, rank, world = self.rows, self.frame_w, self.rank, self.world
        assert mine.numel() == rows[rank] * w * 4, (mine.numel(), rows[rank], w)
        mine = mine.reshape(-1)
        if snapshot and (world > 1 and rank != 0):
            mine = mine.clone()  # 16 B/px of the block, device to device
        self._keep = mine
        if world == 1:
            self.blocks = [mine.reshape(rows[0], w, 4)]
            return self
        if rank != 0:
            self.works = dist.batch_isend_irecv([dist.P2POp(dist.isend, mine.contiguous(), 0)])
            return self
        need = sum(rows[1:]) * w * 4
        if recv is None or recv.numel() < need:
            recv = torch.empty(need, dtype=torch.float32, device=mine.device)
        self._keep = (mine, recv)
        out, off, ops = [mine.reshape(rows[0], w, 4)], 0, []
        for r in range(1, world):
            n = rows[r] * w * 4
            view = recv[off:off + n]
            ops.append(dist.P2POp(dist.irecv, view, r))
            out.append(view.reshape(rows[r], w, 4))
            off += n
        self.works = dist.batch_isend_irecv(ops) if ops else []  # one grouped launch instead of world-1 serialised ones
        self.blocks = out
        return self

    def finish(self):
        for wk in self.works:
            wk.wait()
        self.works = []
        return self.blocks if self.rank == 0 else None


class IpcRowExchange:
    """Block rows -> primary through CUDA IPC mappings of the workers' export buffers (see the module docstring).

    Ordering contract (the caller provides it with `StatsExchange`, see bench.py): rank r may `publish` pass i only after
    rank 0 finished `merge` of pass i - 2 (same slot), and rank 0 may `merge` pass i only after every rank returned from
    `publish` of pass i."""

    def __init__(self, tracer, rank: int, world: int, frame_w: int):
        self.tr, self.rank, self.world, self.w = tracer, rank, world, frame_w
        handles = [tracer.ipc_export(0), tracer.ipc_export(1)] if rank != 0 else None
        everyone = [None] * world
        dist.all_gather_object(everyone, handles)
        self.peers = {}
        if rank == 0:
            for r in range(1, world):
                self.peers[r] = [tracer.ipc_open(h) for h in everyone[r]]

    def publish(self, block_req, pass_index: int):
        if self.rank != 0:
            self.tr.ipc_publish_rows(block_req, pass_index % 2)

    def merge(self, rows, pass_index: int, make_req):
        """rank 0: add every peer's rows of pass `pass_index`; make_req(r) -> the block request of rank r's block."""
        if self.rank != 0:
            return
        y = int(rows[0])
        for r in range(1, self.world):
            if rows[r]:
                self.tr.merge_rows(self.peers[r][pass_index % 2] + 16 * self.w * y, True, make_req(r, y))
            y += int(rows[r])

    def close(self):
        for ptrs in self.peers.values():
            for p in ptrs:
                self.tr.ipc_close(p)
        self.peers = {}


def gather_rows_to_primary(mine: torch.Tensor, rows, frame_w: int, rank: int, world: int, recv: torch.Tensor | None = None):
    """`mine`: this rank's block rows, float32, rows[rank]*frame_w*4 elements (any shape).
    Returns on rank 0 the list of per-rank row tensors shaped (rows[r], frame_w, 4) (entry 0 is `mine`),
    on every other rank None.  `recv` is an optional reusable flat receive buffer on rank 0."""
    return RowGather(rows, frame_w, rank, world).start(mine, recv, snapshot=False).finish()


class StatsExchange:
    """`exchange_stats`, split phase: post the all-gather, read it a pass later (the scheduler then works from timings that
    are one pass old, which changes nothing once the assignment has converged)."""

    def __init__(self, values, world: int, device="cpu"):
        self.world = world
        self.t = torch.tensor([float(x) for x in values], dtype=torch.float64, device=device)
        self.all = [torch.zeros_like(self.t) for _ in range(world)]
        self.work = dist.all_gather(self.all, self.t, async_op=True) if world > 1 else None
        if world == 1:
            self.all = [self.t]

    def result(self):
        if self.work is not None:
            self.work.wait()
            self.work = None
        return [a.tolist() for a in self.all]


def exchange_stats(block_h: int, render_time_s: float, rank: int, world: int, device="cpu", extra=()):
    """All-gather (BlockH, RenderTime, *extra) of every tracer: what `Tracer.Stats()` hands the perfect
    scheduler (tracer/scheduler.go:58-72) when the tracers live in other processes.  Returns a list of
    tuples, one per rank, identical on every rank."""
    vals = [float(block_h), float(render_time_s), *[float(x) for x in extra]]
    if world == 1:
        return [(int(vals[0]), vals[1], *vals[2:])]
    t = torch.tensor(vals, dtype=torch.float64, device=device)
    allt = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(allt, t)
    return [(int(a[0].item()), float(a[1].item()), *[float(x) for x in a[2:].tolist()]) for a in allt]
