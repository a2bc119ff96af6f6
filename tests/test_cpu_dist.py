"""CPU tier, world_size 2 over gloo: the host logic of the one exchange step -- every rank
schedules the same rows from all-gathered stats, traces its row block, and the block rows are gathered to
rank 0 and added into the frame accumulator (aggregateAccumulator with offset FrameW*BlockY,
resources.go:108-124).  The oracle stands in for the tracer so the test runs without a GPU; the merged
frame must equal ONE tracer executing the same block requests sequentially (appendix E 'Merge')."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_path, split_phase=False):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from polaris_b200 import _lib
    from polaris_b200 import tracer as T
    from polaris_b200.gather import RowGather, StatsExchange, exchange_stats, gather_rows_to_primary
    from polaris_b200.scheduler import PerfectScheduler, StaticSpeed
    from tests import common as C

    w, h, spp = 64, 48, 1
    sc = C.small_scene("c2", w, h)
    tr = C.oracle_for(sc, w, h)
    sched = PerfectScheduler()
    speeds = [StaticSpeed(10) for _ in range(world)]
    acc_samples = 0
    history = []
    for p in range(2):  # two passes: naive split, then perfect rebalancing from the exchanged timings
        rows = [int(r) for r in sched.schedule(speeds, h)]
        by = sum(rows[:rank])
        req = T.make_block_request(w, h, block_y=by, block_h=rows[rank], spp=spp, accumulated_samples=acc_samples)
        seeds = T.splitmix_seeds(100 + 10 * p + rank, spp * 6)
        tr.trace(req, seeds)
        mine = tr.read_buffer(_lib.BUF_TRACE_ACCUMULATOR, w * h * 4, np.float32).reshape(h, w, 4)[by:by + rows[rank]].copy()
        if split_phase:  # what bench.py does: post the exchange, trace on, collect later (here: at once, after clobbering `mine`)
            rg = RowGather(rows, w, rank, world).start(torch.from_numpy(mine))
            se = StatsExchange([rows[rank], 0.001 * (1 + 3 * rank)], world)
            if rank != 0:
                mine[:] = -1.0  # the tracer clears its accumulator at the next Trace: the snapshot must already be taken
            blocks = rg.finish()
            stats = [(int(a[0]), a[1]) for a in se.result()]
        else:
            blocks = gather_rows_to_primary(torch.from_numpy(mine), rows, w, rank, world)
            # fake, deterministic render times so both ranks compute the same next assignment
            stats = exchange_stats(rows[rank], 0.001 * (1 + 3 * rank), rank, world)
        for r in range(world):
            speeds[r].set_stats(*stats[r])
        history.append(rows)
        if rank == 0:
            for r, blk in enumerate(blocks):
                np.save(f"{out_path}.p{p}.r{r}.npy", blk.numpy())
        acc_samples += spp
    if rank == 0:
        np.save(out_path + ".rows.npy", np.array(history))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("split_phase", [False, True])
def test_row_gather_world2(tmp_path, split_phase):
    sys.path.insert(0, ROOT)
    from polaris_b200 import _lib
    from polaris_b200 import tracer as T
    from tests import common as C

    out = str(tmp_path / "gather")
    port = _free_port()
    mp.start_processes(_worker, args=(2, port, out, split_phase), nprocs=2, join=True, start_method="spawn")
    rows = np.load(out + ".rows.npy")
    w, h, spp = 64, 48, 1
    assert rows.shape == (2, 2) and rows[0].tolist() == [24, 24] and rows.sum(axis=1).tolist() == [h, h]
    assert rows[1][0] > rows[1][1]  # rank 1 reported 4x the render time -> fewer rows next pass (scheduler.go:50-80)
    sc = C.small_scene("c2", w, h)
    single = C.oracle_for(sc, w, h)
    frame = np.zeros((h, w, 3), np.float32)
    for p in range(2):
        for r in range(2):
            by = int(rows[p][:r].sum())
            req = T.make_block_request(w, h, block_y=by, block_h=int(rows[p][r]), spp=spp, accumulated_samples=p * spp)
            single.trace(req, T.splitmix_seeds(100 + 10 * p + r, spp * 6))
            want = single.read_buffer(_lib.BUF_TRACE_ACCUMULATOR, w * h * 4, np.float32).reshape(h, w, 4)[by:by + int(rows[p][r])]
            got = np.load(f"{out}.p{p}.r{r}.npy")
            assert got.shape == want.shape and got.tobytes() == want.tobytes(), (p, r)
            frame[by:by + int(rows[p][r])] += got[..., :3]
    assert frame.sum() > 0
