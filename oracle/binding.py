"""ctypes binding of oracle/libpolaris_oracle.so with the same Tracer surface as CudaTracer.

TEST INFRASTRUCTURE: the checker, never the thing measured or shipped.
"""
from __future__ import annotations

import ctypes
import os

import numpy as np

from polaris_b200 import _lib
from polaris_b200._lib import BlockRequest, SceneView, Stats
from polaris_b200.tracer import CPU_DEVICE, LOCAL, _HandleTracer

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libpolaris_oracle.so")
_P = ctypes.POINTER
vp, u32, u64, f32 = ctypes.c_void_p, ctypes.c_uint32, ctypes.c_uint64, ctypes.c_float
_SYMS = [
    ("po_create", vp, []),
    ("po_destroy", None, [vp]),
    ("po_set_option", ctypes.c_int, [vp, ctypes.c_int, ctypes.c_int]),
    ("po_resize", ctypes.c_int, [vp, u32, u32]),
    ("po_upload_scene", ctypes.c_int, [vp, _P(SceneView)]),
    ("po_set_camera", ctypes.c_int, [vp, _P(f32), _P(f32)]),
    ("po_trace", ctypes.c_int, [vp, _P(BlockRequest), vp, ctypes.c_size_t, _P(Stats)]),
    ("po_trace_debug", ctypes.c_int, [vp, _P(BlockRequest), vp, ctypes.c_size_t, u32, vp, u64, vp, u32, _P(u32), _P(Stats)]),
    ("po_merge_output", ctypes.c_int, [vp, vp, _P(BlockRequest)]),
    ("po_sync_framebuffer", ctypes.c_int, [vp, _P(BlockRequest), vp]),
    ("po_read_buffer", ctypes.c_int, [vp, ctypes.c_int, vp, u64]),
    ("po_debug_intersect", ctypes.c_int, [vp, vp, u32, ctypes.c_int, vp, vp, vp]),
    ("po_debug_bxdf", ctypes.c_int, [vp, vp, u32, vp]),
    ("po_debug_rng", ctypes.c_int, [vp, u32, u32, vp]),
    ("po_debug_tonemap", ctypes.c_int, [vp, u32, f32, f32, vp]),
    ("po_build_bvh", u32, [vp, vp, vp, u32, ctypes.c_int, vp, vp]),
]
_lib_handle = None


def load():
    global _lib_handle
    if _lib_handle is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} not built; run `make -C oracle`")
        lib = ctypes.CDLL(LIB_PATH)
        for name, res, args in _SYMS:
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = res, args
        _lib_handle = lib
    return _lib_handle


class OracleTracer(_HandleTracer):
    _prefix = "po_"

    def __init__(self, tracer_id="oracle"):
        super().__init__(tracer_id)
        self._lib = None

    def _fn(self, name):
        return getattr(self._lib, "po_" + name)

    def _has(self, name):
        return name != "last_error"

    def init(self):
        if self._h is None:
            self._lib = load()
            self._h = ctypes.c_void_p(self._lib.po_create())

    def close(self):
        if self._h is not None:
            self._lib.po_destroy(self._h)
            self._h = None
        self._has_scene = False

    def flags(self):
        return LOCAL | CPU_DEVICE

    def speed(self):
        return os.cpu_count() or 1

    def debug_intersect(self, rays, mode, want_counters=False):
        rays = np.ascontiguousarray(rays, dtype=_lib.RAY_DTYPE)
        n = rays.shape[0]
        flags = np.zeros(n, dtype=np.uint32)
        hits = np.zeros(n, dtype=_lib.INTERSECTION_DTYPE)
        cnt = np.zeros(3, dtype=np.uint64)
        self._check(self._lib.po_debug_intersect(self._h, rays.ctypes.data, n, mode, flags.ctypes.data, hits.ctypes.data, cnt.ctypes.data))
        return (flags, hits, cnt) if want_counters else (flags, hits)

    def debug_bxdf(self, records):
        records = np.ascontiguousarray(records, dtype=_lib.BXDF_IN_DTYPE)
        out = np.zeros(records.shape[0], dtype=_lib.BXDF_OUT_DTYPE)
        self._check(self._lib.po_debug_bxdf(self._h, records.ctypes.data, records.shape[0], out.ctypes.data))
        return out

    def debug_rng(self, states, draws):
        states = np.ascontiguousarray(states, dtype=np.uint32).copy()
        n = states.shape[0]
        out = np.zeros((n, draws, 2), dtype=np.float32)
        load().po_debug_rng(states.ctypes.data, n, draws, out.ctypes.data)
        return out, states

    def debug_tonemap(self, acc, sample_weight, exposure):
        acc = np.ascontiguousarray(acc, dtype=np.float32).reshape(-1, 4)
        out = np.zeros((acc.shape[0], 4), dtype=np.uint8)
        load().po_debug_tonemap(acc.ctypes.data, acc.shape[0], sample_weight, exposure, out.ctypes.data)
        return out


def build_bvh_literal(bmin, bmax, center, min_leaf_items):
    """bvh.Build restated literally (O(planes x items)); returns (nodes, leaf-ordered items)."""
    from polaris_b200.scene import BVH_NODE_DTYPE

    bmin = np.ascontiguousarray(bmin, dtype=np.float32)
    bmax = np.ascontiguousarray(bmax, dtype=np.float32)
    center = np.ascontiguousarray(center, dtype=np.float32)
    n = bmin.shape[0]
    nodes = np.zeros(2 * n + 1, dtype=BVH_NODE_DTYPE)
    order = np.zeros(n, dtype=np.uint32)
    cnt = load().po_build_bvh(bmin.ctypes.data, bmax.ctypes.data, center.ctypes.data, n, min_leaf_items,
                              nodes.ctypes.data, order.ctypes.data)
    return nodes[:cnt].copy(), order
