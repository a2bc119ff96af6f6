#!/usr/bin/env python
"""Experiment: how much do two independent sample chains running concurrently on ONE GPU gain?
Two tracer handles (own stream + buffers) trace half of the samples each from two host threads."""
import sys, os, threading, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from polaris_b200 import scenes, tracer as T

name, w, h = (sys.argv[1], int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else ("c2_cornell", 1024, 1024)
spp = 64
sc, _, _, _ = scenes.build(name, w, h)
def mk(i):
    tr = T.CudaTracer(f"cuda:0/{i}", 0); tr.init()
    tr.update_state(T.SYNCHRONOUS, T.FRAME_DIMENSIONS, (w, h))
    tr.update_state(T.SYNCHRONOUS, T.SCENE_DATA, sc)
    tr.update_state(T.SYNCHRONOUS, T.CAMERA_DATA, sc.camera)
    return tr
for nh in (1, 2, 3, 4):
    trs = [mk(i) for i in range(nh)]
    rays = [0] * nh
    def work(i):
        req = T.make_block_request(w, h, spp=spp // nh)
        trs[i].trace(req, T.splitmix_seeds(2 + i, (spp // nh) * 6))
        d = trs[i].stats().device
        rays[i] = d["query_rays"] + d["occlusion_rays"]
    for rep in range(3):
        th = [threading.Thread(target=work, args=(i,)) for i in range(nh)]
        t0 = time.perf_counter()
        for t in th: t.start()
        for t in th: t.join()
        dt = time.perf_counter() - t0
    print(f"{name} {w}x{h}: {nh} concurrent handle(s), {spp} spp total: {sum(rays)/dt/1e6:.1f} Mrays/s ({dt*1e3:.1f} ms)", flush=True)
    for t in trs: t.close()
