#!/usr/bin/env python
"""Config 2's scene and camera at growing square resolutions: Mrays/s of the production path (4 chains, graph) vs frame size.
How much of the small-frame inefficiency a larger launch (more samples per launch) could recover."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from polaris_b200 import _lib, scenes, tracer as T

spp = 64
for side in (256, 512, 724, 1024, 1448, 2048, 2896, 4000):
    sc, _, _, _ = scenes.build("c2_cornell", side, side)
    tr = T.CudaTracer("cuda:0", 0); tr.init()
    tr.update_state(T.SYNCHRONOUS, T.FRAME_DIMENSIONS, (side, side))
    tr.update_state(T.SYNCHRONOUS, T.SCENE_DATA, sc)
    tr.update_state(T.SYNCHRONOUS, T.CAMERA_DATA, sc.camera)
    seeds = T.splitmix_seeds(2, spp * 6)
    best = 0
    for rep in range(4):
        req = T.make_block_request(side, side, spp=spp)
        tr.trace(req, seeds)
        d = tr.stats().device
        rays = d["query_rays"] + d["occlusion_rays"]
        best = max(best, rays / d["device_time_ns"] * 1e3)
    print(f"{side}x{side}: {side*side/1e6:6.2f} Mpx, rays/path {rays/(side*side*spp):.2f}, best of 4: {best:8.1f} Mrays/s", flush=True)
    tr.close()
