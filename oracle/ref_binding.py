"""ctypes binding of oracle/_ref/libpolaris_clref.so: the reference's own OpenCL C kernels compiled
for the CPU by oracle/build_ref.py, behind the same Tracer surface as CudaTracer / OracleTracer.

TEST INFRASTRUCTURE: the checker (tests/) and the CPU baseline of bench.py, never the product path.
The library is built in the container that holds /root/reference and travels to the GPU box as a
prebuilt, git-ignored file; `available()` says whether it is there.
"""
from __future__ import annotations

import ctypes
import os

import numpy as np

from polaris_b200 import _lib
from polaris_b200._lib import BlockRequest, SceneView, Stats
from polaris_b200.tracer import CPU_DEVICE, LOCAL, _HandleTracer

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libpolaris_clref.so")
_P = ctypes.POINTER
vp, u32, u64, f32 = ctypes.c_void_p, ctypes.c_uint32, ctypes.c_uint64, ctypes.c_float
_SYMS = [
    ("pr_create", vp, []),
    ("pr_destroy", None, [vp]),
    ("pr_set_option", ctypes.c_int, [vp, ctypes.c_int, ctypes.c_int]),
    ("pr_resize", ctypes.c_int, [vp, u32, u32]),
    ("pr_upload_scene", ctypes.c_int, [vp, _P(SceneView)]),
    ("pr_set_camera", ctypes.c_int, [vp, _P(f32), _P(f32)]),
    ("pr_trace", ctypes.c_int, [vp, _P(BlockRequest), vp, ctypes.c_size_t, _P(Stats)]),
    ("pr_trace_debug", ctypes.c_int, [vp, _P(BlockRequest), vp, ctypes.c_size_t, u32, vp, u64, vp, u32, _P(u32), _P(Stats)]),
    ("pr_merge_output", ctypes.c_int, [vp, vp, _P(BlockRequest)]),
    ("pr_sync_framebuffer", ctypes.c_int, [vp, _P(BlockRequest), vp]),
    ("pr_read_buffer", ctypes.c_int, [vp, ctypes.c_int, vp, u64]),
    ("pr_debug_intersect", ctypes.c_int, [vp, vp, u32, ctypes.c_int, vp, vp, vp]),
    ("pr_debug_bxdf", ctypes.c_int, [vp, vp, u32, vp]),
    ("pr_debug_rng", ctypes.c_int, [vp, u32, u32, vp]),
    ("pr_debug_tonemap", ctypes.c_int, [vp, u32, f32, f32, vp]),
    ("pr_stack_need", ctypes.c_int, [vp]),
    ("pr_num_threads", ctypes.c_int, []),
    ("pr_set_num_threads", None, [ctypes.c_int]),
]
_lib_handle = None


def available() -> bool:
    return os.path.exists(LIB_PATH)


def load():
    global _lib_handle
    if _lib_handle is None:
        if not available():
            raise RuntimeError(f"{LIB_PATH} not built; run `python oracle/build_ref.py` where /root/reference exists")
        lib = ctypes.CDLL(LIB_PATH)
        for name, res, args in _SYMS:
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = res, args
        _lib_handle = lib
    return _lib_handle


class RefTracer(_HandleTracer):
    """The reference's kernels on the host CPU (what `tracer/opencl` runs on a CPU device,
    pipeline.go:107-111: per-ray traversal for primary rays too)."""

    _prefix = "pr_"

    def __init__(self, tracer_id="clref"):
        super().__init__(tracer_id)
        self._lib = None

    def _fn(self, name):
        return getattr(self._lib, "pr_" + name)

    def _has(self, name):
        return name != "last_error"

    def init(self):
        if self._h is None:
            self._lib = load()
            self._h = ctypes.c_void_p(self._lib.pr_create())

    def close(self):
        if self._h is not None:
            self._lib.pr_destroy(self._h)
            self._h = None
        self._has_scene = False

    def flags(self):
        return LOCAL | CPU_DEVICE

    def speed(self):
        return os.cpu_count() or 1

    def threads(self):
        return int(load().pr_num_threads())

    @staticmethod
    def set_threads(n: int):
        """Use n OpenMP threads from now on (a launcher may have exported OMP_NUM_THREADS=1)."""
        load().pr_set_num_threads(int(n))

    def stack_need(self):
        return int(self._lib.pr_stack_need(self._h))

    def debug_intersect(self, rays, mode):
        rays = np.ascontiguousarray(rays, dtype=_lib.RAY_DTYPE)
        n = rays.shape[0]
        flags = np.zeros(n, dtype=np.uint32)
        hits = np.zeros(n, dtype=_lib.INTERSECTION_DTYPE)
        self._check(self._lib.pr_debug_intersect(self._h, rays.ctypes.data, n, mode, flags.ctypes.data, hits.ctypes.data, None))
        return flags, hits

    def debug_bxdf(self, records):
        records = np.ascontiguousarray(records, dtype=_lib.BXDF_IN_DTYPE)
        out = np.zeros(records.shape[0], dtype=_lib.BXDF_OUT_DTYPE)
        self._check(self._lib.pr_debug_bxdf(self._h, records.ctypes.data, records.shape[0], out.ctypes.data))
        return out

    def debug_rng(self, states, draws):
        states = np.ascontiguousarray(states, dtype=np.uint32).copy()
        n = states.shape[0]
        out = np.zeros((n, draws, 2), dtype=np.float32)
        load().pr_debug_rng(states.ctypes.data, n, draws, out.ctypes.data)
        return out, states

    def debug_tonemap(self, acc, sample_weight, exposure):
        acc = np.ascontiguousarray(acc, dtype=np.float32).reshape(-1, 4)
        out = np.zeros((acc.shape[0], 4), dtype=np.uint8)
        load().pr_debug_tonemap(acc.ctypes.data, acc.shape[0], sample_weight, exposure, out.ctypes.data)
        return out
