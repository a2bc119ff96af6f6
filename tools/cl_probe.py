#!/usr/bin/env python
"""GPU-box probe: can the reference's OpenCL program be built and run on this box's OpenCL driver?
   python tools/cl_probe.py   (writes gpurun_out/cl_probe.txt)"""
import os, subprocess, sys, time, traceback
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

out = open(os.path.join("gpurun_out", "cl_probe.txt"), "w") if os.path.isdir("gpurun_out") else sys.stdout


def say(*a):
    print(*a, flush=True)
    if out is not sys.stdout:
        print(*a, file=out, flush=True)


for lib in ("/usr/lib/libnvidia-opencl.so.1", "/usr/local/nvidia/lib/libnvidia-opencl.so.1"):
    if os.path.exists(lib):
        r = subprocess.run(["nm", "-D", "--defined-only", lib], capture_output=True, text=True)
        say(lib, "exports:", " ".join(l.split()[-1] for l in r.stdout.splitlines())[:600])
        break
try:
    from oracle import cl_device as CLD
    from oracle import ref_binding
    from polaris_b200 import _lib, tracer as T
    from tests import common as C

    say("available:", CLD.available())
    a = CLD.api()
    say("loader:", a.how)
    tr = CLD.ClDeviceTracer()
    tr.init()
    say("device:", tr.dev.describe())
    say("build log:", tr.dev.build_log[:1500] or "<empty>")
    w = h = 128
    sc = C.small_scene("c2", w, h)
    C.setup(tr, sc, w, h)
    rays = C.fixed_rays(sc, w, h)
    ref = C.ref_for(sc, w, h) if hasattr(C, "ref_for") else None
    cu = C.cuda_for(sc, w, h)
    for mode, nm in ((0, "query"), (1, "test"), (2, "packet query")):
        f_cl, h_cl = tr.debug_intersect(rays, mode)
        f_cu, h_cu = cu.debug_intersect(rays, 1 if mode == 1 else 0)
        same_flag = int((f_cl == f_cu).sum())
        msg = f"{nm}: {len(rays)} rays, flags equal {same_flag}"
        if mode != 1:
            hit = (f_cl != 0) & (f_cu != 0)
            ids = (h_cl["mesh_instance"] == h_cu["mesh_instance"]) & (h_cl["tri_index"] == h_cu["tri_index"])
            bits = (h_cl["wuvt"].view(np.uint32) == h_cu["wuvt"].view(np.uint32)).all(axis=1)
            dt = np.abs(h_cl["wuvt"][:, 3] - h_cu["wuvt"][:, 3])[hit] / np.maximum(np.abs(h_cu["wuvt"][:, 3][hit]), 1e-6)
            msg += f", both hit {int(hit.sum())}, ids equal {int((ids & hit).sum())}, wuvt bit-equal {int((bits & hit).sum())}, max rel dt {dt.max() if len(dt) else 0:.3g}"
        say(msg)
    seeds = T.splitmix_seeds(1, 2)
    for name, t in (("opencl", tr), ("cuda", cu)):
        t.set_option(_lib.OPT_PRIMARY_PACKETS, 0)
    accs = {}
    for name, t in (("opencl", tr), ("cuda", cu)):
        req = T.make_block_request(w, h, spp=1, num_bounces=1)
        t0 = time.perf_counter()
        t.trace(req, seeds)
        accs[name] = C.acc_of(t, _lib.BUF_TRACE_ACCUMULATOR, w, h).copy()
        say(name, "bounce-0 trace:", f"{time.perf_counter() - t0:.4f}s", t.stats().device if name == "opencl" else "")
    err = C.rel_err(accs["opencl"], accs["cuda"])
    say(f"bounce-0 radiance opencl vs cuda: max rel {err.max():.3g}, >1e-4: {(err > 1e-4).sum()} / {len(err)}, >1e-3: {(err > 1e-3).sum()}, mean cl {accs['opencl'].mean():.6f} cuda {accs['cuda'].mean():.6f}")
    # throughput of the reference discipline on this device: c2 at 512x512, 4 spp, 5 bounces
    w = h = 512
    sc = C.scene("c2_cornell", w, h)
    C.setup(tr, sc, w, h)
    tr.set_option(_lib.OPT_PRIMARY_PACKETS, 1)
    seeds = T.splitmix_seeds(2, 4 * 6)
    for rep in range(2):
        req = T.make_block_request(w, h, spp=4)
        dt = tr.trace(req, seeds)
        d = tr.stats().device
        say(f"opencl c2 {w}x{h} 4spp: {dt:.3f}s, {(d['query_rays'] + d['occlusion_rays']) / dt / 1e6:.1f} Mrays/s, {d}")
    tr.merge_output(tr, req)
    tr.sync_framebuffer(T.make_block_request(w, h, spp=4))
    say("frame mean RGBA:", tr.frame_buffer.reshape(-1, 4).mean(axis=0))
    tr.close()
except Exception:
    say(traceback.format_exc())
