#!/usr/bin/env python
"""Design tool: SIMT schedule model of the traversal loop (tests/emul pe_simt) on real bounce rays of a config.
   python tools/simt_model.py [scene] [w] [h]"""
import ctypes, sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests import common as C
from polaris_b200 import _lib, tracer as T

name = sys.argv[1] if len(sys.argv) > 1 else "c2_cornell"
w = int(sys.argv[2]) if len(sys.argv) > 2 else 512
h = int(sys.argv[3]) if len(sys.argv) > 3 else 512
sc = C.scene(name, w, h)
orc = C.oracle_for(sc, w, h)
orc.trace(T.make_block_request(w, h, spp=1, num_bounces=2), T.splitmix_seeds(2, 3))
cnt = orc.read_buffer(_lib.BUF_RAY_COUNTERS, 3, np.int32)
ind = orc.read_buffer(_lib.BUF_RAYS0, w * h, _lib.RAY_DTYPE)[:cnt[0]].copy()
occ = orc.read_buffer(_lib.BUF_RAYS2, w * h, _lib.RAY_DTYPE)[:cnt[2]].copy()
emu = C.Emul(sc, w, h)
emu.lib.pe_simt.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint32] + [ctypes.c_int] * 5 + [ctypes.c_void_p]
Cn, Ct, Co, Cr = 50, 45, 20, 12


def sim(rays, any_hit, variant, tricap=0, refill=0, inner_min=0):
    out = np.zeros(8)
    emu.lib.pe_simt(emu.h, rays.ctypes.data, len(rays), any_hit, variant, tricap, refill, inner_min, out.ctypes.data)
    wi, li, wt, lt, wo, lo, rounds, n = out
    cost = wi * Cn + wt * Ct + wo * Co + rounds * Cr
    useful = (li * Cn + lt * Ct + lo * Co) / 32
    return dict(inner_eff=li / max(1, wi) / 32, tri_eff=lt / max(1, wt) / 32, nodes=li / n, tris=lt / n, other=lo / n,
                rounds_per_warp=rounds / (n / 32), cost_per_ray=cost / n, simd_eff=useful / cost)


if __name__ == "__main__":
    for nm, rays, ah in (("indirect", ind, 0), ("occlusion", occ, 1)):
        print(nm, len(rays), "rays")
        for v, refill, im in ((0, 0, 0), (1, 0, 0), (1, 20, 0), (1, 24, 0), (1, 28, 0), (1, 20, 4), (1, 20, 8), (1, 20, 12), (1, 20, 16), (1, 24, 8), (1, 28, 8), (1, 28, 12), (1, 0, 8)):
            r = sim(rays, ah, v, 0, refill, im)
            print(f"  variant {v} refill {refill:2d} innerMin {im:2d}: " + " ".join(f"{k}={x:.3f}" for k, x in r.items()))
