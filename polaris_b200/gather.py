"""The one exchange step of the multi-GPU path (one process per GPU): block rows -> primary.

The reference merges inside one process: every worker calls `primary.MergeOutput(self, blockReq)` and the
`aggregateAccumulator` kernel reads the peer device's buffer through a shared OpenCL context
(tracer/opencl/tracer.go:279-286, resources.go:108-124, renderer/default.go:188-191).  With one process
per GPU the same data movement is a gather of each rank's block rows (16 B per pixel, rows
[BlockY, BlockY+BlockH) only) to rank 0, which adds them into its frame accumulator with
`pc_merge_rows`.  It is a grouped send/recv, not a reduce: outside its own rows a rank's accumulator is
zero, so an all-reduce would move world_size times more bytes.

Backend agnostic: CUDA tensors over NCCL/NVLink on the B200 box, CPU tensors over gloo in the tests.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def gather_rows_to_primary(mine: torch.Tensor, rows, frame_w: int, rank: int, world: int, recv: torch.Tensor | None = None):
    """`mine`: this rank's block rows, float32, rows[rank]*frame_w*4 elements (any shape).
    Returns on rank 0 the list of per-rank row tensors shaped (rows[r], frame_w, 4) (entry 0 is `mine`),
    on every other rank None.  `recv` is an optional reusable flat receive buffer on rank 0."""
    rows = [int(r) for r in rows]
    assert mine.numel() == rows[rank] * frame_w * 4, (mine.numel(), rows[rank], frame_w)
    if world == 1:
        return [mine.reshape(rows[0], frame_w, 4)]
    if rank != 0:
        dist.send(mine.reshape(-1).contiguous(), dst=0)
        return None
    need = sum(rows[1:]) * frame_w * 4
    if recv is None or recv.numel() < need:
        recv = torch.empty(need, dtype=torch.float32, device=mine.device)
    out, works, off = [mine.reshape(rows[0], frame_w, 4)], [], 0
    for r in range(1, world):
        n = rows[r] * frame_w * 4
        view = recv[off:off + n]
        works.append(dist.irecv(view, src=r))
        out.append(view.reshape(rows[r], frame_w, 4))
        off += n
    for wk in works:
        wk.wait()
    return out


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
