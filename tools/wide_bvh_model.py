#!/usr/bin/env python
"""Design tool: would a 4-ary collapse of the reference's BVH pay?  (tests/emul pe_simt_wide vs pe_simt)

Runs the real bounce rays of a config through (a) the binary while-while walk the kernels use and (b) the same walk over
wide nodes holding each inner node's grandchildren, 32 lanes in lockstep, and reports warp-level iteration counts (the
length of a warp's dependent chain), lane efficiency -- and checks that the wide walk returns the SAME hit records bit for
bit (flags, instance, triangle, w/u/v/t), which is what makes it admissible at all.

    python tools/wide_bvh_model.py [scene] [w] [h]
"""
import ctypes, sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests import common as C
from polaris_b200 import _lib, tracer as T

name = sys.argv[1] if len(sys.argv) > 1 else "c2_cornell"
w = int(sys.argv[2]) if len(sys.argv) > 2 else 384
h = int(sys.argv[3]) if len(sys.argv) > 3 else 384
sc = C.scene(name, w, h)
orc = C.oracle_for(sc, w, h)
orc.trace(T.make_block_request(w, h, spp=1, num_bounces=2), T.splitmix_seeds(2, 3))
cnt = orc.read_buffer(_lib.BUF_RAY_COUNTERS, 3, np.int32)
ind = orc.read_buffer(_lib.BUF_RAYS0, w * h, _lib.RAY_DTYPE)[:cnt[0]].copy()
occ = orc.read_buffer(_lib.BUF_RAYS2, w * h, _lib.RAY_DTYPE)[:cnt[2]].copy()
emu = C.Emul(sc, w, h)
vp = ctypes.c_void_p
emu.lib.pe_simt.argtypes = [vp, vp, ctypes.c_uint32] + [ctypes.c_int] * 5 + [vp]
emu.lib.pe_simt_wide.argtypes = [vp, vp, ctypes.c_uint32, ctypes.c_int, vp, vp, vp]
C_NODE, C_WIDE, C_TRI, C_OTHER, C_ROUND = 50, 105, 45, 20, 12  # instruction-ish weights of one warp iteration

print(f"{name} {w}x{h}: layout {emu.layout_info()}")
for nm, rays, ah in (("indirect (closest hit)", ind, 0), ("occlusion (any hit)", occ, 1)):
    n = len(rays)
    o = np.zeros(8)
    emu.lib.pe_simt(emu.h, rays.ctypes.data, n, ah, 0, 0, 0, 0, o.ctypes.data)
    wi, li, wt, lt, wo, lo, rounds, _ = o
    ow = np.zeros(8)
    flags = np.zeros(n, np.uint32)
    hits = np.zeros(n, _lib.INTERSECTION_DTYPE)
    emu.lib.pe_simt_wide(emu.h, rays.ctypes.data, n, ah, ow.ctypes.data, flags.ctypes.data, hits.ctypes.data)
    ww, lw, wt2, lt2, wo2, lo2, rounds2, _ = ow
    f0, h0, _ = emu.intersect(rays, 1 if ah else 0)
    same_flags = bool((f0 == flags).all())
    hit = f0 == 1
    same_hits = ah == 1 or (hits["wuvt"][hit].tobytes() == h0["wuvt"][hit].tobytes() and (hits["mesh_instance"][hit] == h0["mesh_instance"][hit]).all()
                            and (hits["tri_index"][hit] == h0["tri_index"][hit]).all())
    units = n / 32
    chain_b = (wi * C_NODE + wt * C_TRI + wo * C_OTHER + rounds * C_ROUND) / units
    chain_w = (ww * C_WIDE + wt2 * C_TRI + wo2 * C_OTHER + rounds2 * C_ROUND) / units
    print(f"{nm}: {n} rays; results identical: flags {same_flags}, hit records {same_hits}")
    print(f"  binary: {wi / units:6.1f} node + {wt / units:6.1f} triangle + {wo / units:5.1f} other warp iterations per unit in {rounds / units:5.1f} rounds; "
          f"lanes per node step {li / max(1, wi):4.1f}, box tests per ray {2 * li / n:5.1f}, triangles per ray {lt / n:5.1f}; chain {chain_b:7.0f}")
    print(f"  4-ary : {ww / units:6.1f} node + {wt2 / units:6.1f} triangle + {wo2 / units:5.1f} other warp iterations per unit in {rounds2 / units:5.1f} rounds; "
          f"lanes per node step {lw / max(1, ww):4.1f}, box tests per ray ~{3.5 * lw / n:5.1f}, triangles per ray {lt2 / n:5.1f}; chain {chain_w:7.0f} "
          f"({100 * (chain_w / chain_b - 1):+.0f} %)")
