#!/usr/bin/env python
"""Measure how the per-ray cost of each kernel class depends on the block size (rays per launch)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from polaris_b200 import _lib, scenes, tracer as T

w, h = 3840, 2160
sc, _, _, _ = scenes.build("c5_cornell_4k", w, h)
tr = T.CudaTracer("cuda:0", 0); tr.init()
tr.update_state(T.SYNCHRONOUS, T.FRAME_DIMENSIONS, (w, h))
tr.update_state(T.SYNCHRONOUS, T.SCENE_DATA, sc)
tr.update_state(T.SYNCHRONOUS, T.CAMERA_DATA, sc.camera)
spp = 8
seeds = T.splitmix_seeds(5, spp * 6)
for timers in (0, 1):
    tr.set_option(_lib.OPT_KERNEL_TIMERS, timers)
    for bh in (1, 4, 16, 64, 135, 270, 540, 1080, 2160)[:int(os.environ.get("NB", 9))]:
        by = (h - bh) // 2
        for rep in range(2):
            req = T.make_block_request(w, h, block_y=by, block_h=bh, spp=spp)
            tr.trace(req, seeds)
        d = tr.stats().device
        rays = d["query_rays"] + d["occlusion_rays"]
        line = f"timers={timers} block_h={bh:5d} paths/launch={w*bh/1e6:6.2f}M  device {d['device_time_ns']/1e6:8.2f} ms  {rays/d['device_time_ns']*1e3:8.1f} Mrays/s"
        if timers:
            line += "  per-class us/launch: " + " ".join(f"{n}={d['kernel_time_ns'][i]/max(1,d['kernel_count'][i])/1e3:.0f}" for i, n in enumerate(_lib.KERNEL_CLASS_NAMES) if d['kernel_count'][i])
        print(line, flush=True)
