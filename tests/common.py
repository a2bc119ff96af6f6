"""Shared helpers for the parity tests: tracer setup, the test-only emulation binding, comparisons."""
from __future__ import annotations

import ctypes
import functools
import os

import numpy as np

from oracle.binding import OracleTracer
from polaris_b200 import _lib, scenes
from polaris_b200 import tracer as T

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
vp = ctypes.c_void_p


@functools.lru_cache(maxsize=8)
def scene(name, w, h, **sizes):
    sc, _, _, _ = scenes.build(name, w, h, **sizes)
    return sc


SMALL = {  # reduced-size variants of the BASELINE configs that the oracle finishes in seconds
    "c1": ("c1_sphere", {}),
    "c2": ("c2_cornell", {"sphere_seg": (32, 16)}),
    "c3": ("c3_instancing", {"grid": 48, "lattice": 4}),
    "c4": ("c4_terrain", {"n": 129, "tex": 128, "sky": (256, 128)}),
}


def small_scene(key, w, h):
    name, sizes = SMALL[key]
    return scene(name, w, h, **{k: v for k, v in sizes.items()})


def setup(tr, sc, w, h):
    tr.init()
    tr.update_state(T.SYNCHRONOUS, T.FRAME_DIMENSIONS, (w, h))
    tr.update_state(T.SYNCHRONOUS, T.SCENE_DATA, sc)
    tr.update_state(T.SYNCHRONOUS, T.CAMERA_DATA, sc.camera)
    return tr


def oracle_for(sc, w, h):
    return setup(OracleTracer(), sc, w, h)


def cuda_for(sc, w, h, ordinal=0, **opts):
    tr = setup(T.CudaTracer(f"cuda:{ordinal}", ordinal), sc, w, h)
    for k, v in opts.items():
        tr.set_option(getattr(_lib, "OPT_" + k.upper()), v)
    return tr


def acc_of(tr, which, w, h):
    return tr.read_buffer(which, w * h * 4, np.float32).reshape(h * w, 4)[:, :3]


def rel_err(a, b, floor=1e-6):
    """per-pixel max-channel relative error with an absolute floor"""
    return (np.abs(a - b) / np.maximum(np.maximum(np.abs(a), np.abs(b)), floor)).max(axis=1)


def fixed_rays(sc, w, h, n=4096, seed=7):
    """The fixed ray set of SURVEY §8(c): primary rays of a seeded sample plus cosine-ish bounce
    rays leaving the oracle's primary hit points, half of them with finite max distance."""
    rng = np.random.default_rng(seed)
    orc = oracle_for(sc, w, h)
    req = T.make_block_request(w, h, spp=1, num_bounces=1)
    orc.trace(req, T.splitmix_seeds(9, 2))
    cnt = orc.read_buffer(_lib.BUF_RAY_COUNTERS, 3, np.int32)
    prim = orc.read_buffer(_lib.BUF_RAYS0, w * h, _lib.RAY_DTYPE)
    ind = orc.read_buffer(_lib.BUF_RAYS1, w * h, _lib.RAY_DTYPE)[: cnt[1]]
    occ = orc.read_buffer(_lib.BUF_RAYS2, w * h, _lib.RAY_DTYPE)[: cnt[2]]
    orc.close()
    parts = [prim[rng.choice(len(prim), min(n // 2, len(prim)), replace=False)]]
    if len(ind):
        parts.append(ind[rng.choice(len(ind), min(n // 4, len(ind)), replace=False)])
    if len(occ):
        parts.append(occ[rng.choice(len(occ), min(n // 4, len(occ)), replace=False)])
    rays = np.concatenate(parts).copy()
    return rays


class Emul:
    """tests/emul/libpc_emul.so: the CUDA device functions compiled for the host (TEST ONLY)."""

    def __init__(self, sc, w, h, variant=""):
        lib = ctypes.CDLL(os.path.join(ROOT, "tests", "emul", f"libpc_emul{variant}.so"))
        lib.pe_create.restype = vp
        lib.pe_create.argtypes = [ctypes.POINTER(_lib.SceneView)]
        lib.pe_destroy.argtypes = [vp]
        lib.pe_error.restype = ctypes.c_char_p
        lib.pe_error.argtypes = [vp]
        lib.pe_intersect.argtypes = [vp, vp, ctypes.c_uint32, ctypes.c_int, vp, vp, vp]
        lib.pe_resize.argtypes = [vp, ctypes.c_uint32, ctypes.c_uint32]
        lib.pe_set_camera.argtypes = [vp, vp, vp]
        lib.pe_trace.argtypes = [vp, ctypes.POINTER(_lib.BlockRequest), vp, ctypes.c_size_t]
        lib.pe_read_buffer.argtypes = [vp, ctypes.c_int, vp, ctypes.c_uint64]
        lib.pe_layout_info.argtypes = [vp, vp, vp, vp, vp]
        lib.pe_bxdf.argtypes = [vp, vp, ctypes.c_uint32, vp]
        lib.pe_tonemap.argtypes = [vp, ctypes.c_uint32, ctypes.c_float, ctypes.c_float, vp]
        self.lib = lib
        self.view, self.keep = _lib.scene_view(sc)
        self.h = lib.pe_create(ctypes.byref(self.view))
        err = lib.pe_error(self.h)
        if err:
            raise RuntimeError(err.decode())
        lib.pe_resize(self.h, w, h)
        eye = np.ascontiguousarray(sc.camera.position, dtype=np.float32)
        fr = np.ascontiguousarray(sc.camera.frustrum, dtype=np.float32)
        lib.pe_set_camera(self.h, eye.ctypes.data, fr.ctypes.data)

    def layout_info(self):
        a, b, c = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
        d = ctypes.c_uint32()
        self.lib.pe_layout_info(self.h, ctypes.byref(a), ctypes.byref(b), ctypes.byref(c), ctypes.byref(d))
        return {"top_depth": a.value, "mesh_depth": b.value, "stack_need": c.value, "inner_nodes": d.value}

    def intersect(self, rays, mode):
        rays = np.ascontiguousarray(rays, dtype=_lib.RAY_DTYPE)
        n = rays.shape[0]
        flags = np.zeros(n, np.uint32)
        hits = np.zeros(n, _lib.INTERSECTION_DTYPE)
        cnt = np.zeros(3, np.uint64)
        rc = self.lib.pe_intersect(self.h, rays.ctypes.data, n, mode, flags.ctypes.data, hits.ctypes.data, cnt.ctypes.data)
        assert rc == 0, rc
        return flags, hits, cnt

    def trace(self, req, seeds):
        seeds = np.ascontiguousarray(seeds, dtype=np.uint32)
        rc = self.lib.pe_trace(self.h, ctypes.byref(req), seeds.ctypes.data, seeds.size)
        assert rc == 0, rc

    def read_buffer(self, which, count, dtype):
        out = np.empty(count, dtype=dtype)
        self.lib.pe_read_buffer(self.h, which, out.ctypes.data, out.nbytes)
        return out

    def bxdf(self, records):
        records = np.ascontiguousarray(records, dtype=_lib.BXDF_IN_DTYPE)
        out = np.zeros(records.shape[0], dtype=_lib.BXDF_OUT_DTYPE)
        self.lib.pe_bxdf(self.h, records.ctypes.data, records.shape[0], out.ctypes.data)
        return out

    def close(self):
        if self.h:
            self.lib.pe_destroy(self.h)
            self.h = None


def bxdf_records(sc, n_theta=8, n_rand=8, seed=3):
    """Per-BxDF sample/pdf/eval table inputs: every leaf material node x (theta_i, r) grid."""
    from polaris_b200.material import OP_MIX

    rng = np.random.default_rng(seed)
    leaves = [i for i, m in enumerate(sc.material_nodes) if 2 < int(m["union1"][0]) < OP_MIX]
    recs = []
    for leaf in leaves:
        for ti in range(n_theta):
            for sign in (1.0, -1.0):  # hit from outside and from inside
                ct = (ti + 0.5) / n_theta
                st = np.sqrt(max(0.0, 1 - ct * ct))
                for _ in range(n_rand):
                    phi = rng.uniform(0, 2 * np.pi)
                    r = np.zeros((), dtype=_lib.BXDF_IN_DTYPE)
                    r["normal"] = (0, 1, 0)
                    r["mat_node"] = leaf
                    r["in_dir"] = (st * np.cos(phi), sign * ct, st * np.sin(phi))
                    o = rng.normal(size=3)
                    o /= np.linalg.norm(o)
                    r["out_dir"] = o
                    r["rnd"] = rng.uniform(0, 1, size=2)
                    r["uv"] = rng.uniform(0, 1, size=2)
                    recs.append(r)
    return np.array(recs, dtype=_lib.BXDF_IN_DTYPE)
