"""Python mirror of the reference's tracer seam, bound to libpolaris_cuda.so.

`Tracer` restates the Go interface `tracer.Tracer` (reference tracer/tracer.go:80-111) with the
same method names (snake-cased), argument meaning and error behaviour; `CudaTracer` is the
`tracer/cuda` implementation a Go maintainer would write over the C ABI (INTEGRATION.md,
go/tracer/cuda/tracer.go), here over ctypes because the image has no Go toolchain.
"""
from __future__ import annotations

import abc
import ctypes
import random
import time
from dataclasses import dataclass

import numpy as np

from . import _lib
from ._lib import BlockRequest, Stats

# tracer.Flag (tracer/tracer.go:49-61)
LOCAL, REMOTE, CPU_DEVICE = 1, 2, 4
# tracer.UpdateMode (tracer/tracer.go:63-69)
SYNCHRONOUS, ASYNCHRONOUS = 0, 1
# tracer.ChangeType (tracer/tracer.go:71-78)
FRAME_DIMENSIONS, SCENE_DATA, CAMERA_DATA = 0, 1, 2


class TracerError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(msg)
        self.code = code


class ErrNoSceneData(TracerError):
    """tracer/opencl/errors.go:21"""


class ErrUnsupportedChangeType(TracerError):
    """tracer/opencl/errors.go:18 / tracer.go:186"""


class ErrUnsupportedTracer(TracerError):
    """tracer/opencl/tracer.go:282 'merge failed: unsupported tracer instance'"""


@dataclass
class TracerStats:
    """tracer.Stats (tracer/tracer.go:37-47); times in seconds."""

    block_w: int = 0
    block_h: int = 0
    update_time: float = 0.0
    render_time: float = 0.0
    device: dict | None = None  # raw pc_stats of the last trace


def make_block_request(frame_w, frame_h, block_y=0, block_h=None, spp=1, num_bounces=5,
                       min_bounces_for_rr=3, exposure=1.2, seed=0, accumulated_samples=0) -> BlockRequest:
    r = BlockRequest()
    r.frame_w, r.frame_h = frame_w, frame_h
    r.block_x, r.block_y = 0, block_y
    r.block_w, r.block_h = frame_w, frame_h if block_h is None else block_h
    r.samples_per_pixel, r.num_bounces, r.min_bounces_for_rr = spp, num_bounces, min_bounces_for_rr
    r.exposure, r.seed, r.accumulated_samples = exposure, seed, accumulated_samples
    return r


def splitmix_seeds(config_number: int, count: int) -> np.ndarray:
    """Host seed list of SURVEY §8(d): successive splitmix64 outputs from state
    0x501A2150 + config_number, low 32 bits."""
    out = np.empty(count, dtype=np.uint32)
    s = (0x501A2150 + config_number) & 0xFFFFFFFFFFFFFFFF
    m = 0xFFFFFFFFFFFFFFFF
    for i in range(count):
        s = (s + 0x9E3779B97F4A7C15) & m
        z = s
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & m
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & m
        z ^= z >> 31
        out[i] = z & 0xFFFFFFFF
    return out


class Tracer(abc.ABC):
    """tracer.Tracer (tracer/tracer.go:80-111)."""

    @abc.abstractmethod
    def id(self) -> str: ...
    @abc.abstractmethod
    def flags(self) -> int: ...
    @abc.abstractmethod
    def speed(self) -> int: ...
    @abc.abstractmethod
    def init(self) -> None: ...
    @abc.abstractmethod
    def close(self) -> None: ...
    @abc.abstractmethod
    def stats(self) -> TracerStats: ...
    @abc.abstractmethod
    def update_state(self, mode: int, change_type: int, data) -> float: ...
    @abc.abstractmethod
    def trace(self, block_req: BlockRequest, seeds=None) -> float: ...
    @abc.abstractmethod
    def merge_output(self, other: "Tracer", block_req: BlockRequest) -> float: ...
    @abc.abstractmethod
    def sync_framebuffer(self, block_req: BlockRequest) -> float: ...


class _HandleTracer(Tracer):
    """Shared host logic: change buffering (tracer.go:150-191) over a pc_*-shaped C interface."""

    _prefix = "pc_"

    def __init__(self, tracer_id: str):
        self._id = tracer_id
        self._h = None
        self._stats = TracerStats()
        self._change_buffer = {}
        self._keep = None  # keeps scene arrays alive for backends that borrow them
        self._has_scene = False
        self.frame_buffer = None  # RGBA8 (H, W, 4) after sync_framebuffer

    # -- backend hooks
    def _fn(self, name):
        raise NotImplementedError

    def _check(self, rc):
        if rc == 0:
            return
        msg = self._fn("last_error")(self._h) if self._has("last_error") else None
        msg = msg.decode() if msg else f"error code {rc}"
        cls = {_lib.ERR_NO_SCENE_DATA: ErrNoSceneData, _lib.ERR_UNSUPPORTED_TRACER: ErrUnsupportedTracer}.get(rc, TracerError)
        raise cls(rc, f"{self._prefix[:-1]} tracer: {msg}")

    def _has(self, name):
        return True

    # -- tracer.Tracer
    def id(self):
        return self._id

    def stats(self):
        return self._stats

    def update_state(self, mode, change_type, data):
        self._change_buffer[change_type] = data
        if mode == SYNCHRONOUS:
            return self._commit_changes()
        return 0.0

    def _commit_changes(self):
        if not self._change_buffer:
            return 0.0
        t0 = time.perf_counter()
        for change_type, data in list(self._change_buffer.items()):
            if change_type == FRAME_DIMENSIONS:
                w, h = data
                self._check(self._fn("resize")(self._h, int(w), int(h)))
            elif change_type == SCENE_DATA:
                view, keep = _lib.scene_view(data)
                self._check(self._fn("upload_scene")(self._h, ctypes.byref(view)))
                self._keep = (data, keep)
                self._has_scene = True
            elif change_type == CAMERA_DATA:
                eye = np.ascontiguousarray(data.position, dtype=np.float32)
                fr = np.ascontiguousarray(data.frustrum, dtype=np.float32).reshape(16)
                self._check(self._fn("set_camera")(self._h, eye.ctypes.data_as(ctypes.POINTER(ctypes.c_float)),
                                                   fr.ctypes.data_as(ctypes.POINTER(ctypes.c_float))))
            else:
                raise ErrUnsupportedChangeType(_lib.ERR_INVALID_ARGUMENT, f"unsupported change type {change_type}")
        self._change_buffer = {}
        dt = time.perf_counter() - t0
        self._stats.update_time = dt
        return dt

    def trace(self, block_req, seeds=None):
        t0 = time.perf_counter()
        self._commit_changes()
        if not self._has_scene:
            raise ErrNoSceneData(_lib.ERR_NO_SCENE_DATA, "no scene data uploaded")
        per_sample = 1 + block_req.num_bounces
        need = per_sample * block_req.samples_per_pixel
        if seeds is None:
            # the reference draws from Go's global math/rand (tracer.go:222, pipeline.go:146)
            seeds = np.array([random.getrandbits(32) for _ in range(need)], dtype=np.uint32)
        seeds = np.ascontiguousarray(seeds, dtype=np.uint32)
        if seeds.size < need:
            raise TracerError(_lib.ERR_INVALID_ARGUMENT, f"need {need} seeds, got {seeds.size}")
        st = Stats()
        self._check(self._fn("trace")(self._h, ctypes.byref(block_req), seeds.ctypes.data, seeds.size, ctypes.byref(st)))
        dt = time.perf_counter() - t0
        self._stats.block_w, self._stats.block_h = block_req.block_w, block_req.block_h
        self._stats.render_time = dt
        self._stats.device = st.as_dict()
        return dt

    def trace_debug(self, block_req, seeds, debug_flags):
        """Trace with opencl.DebugFlag stages (pipeline.go:17-30,113-200).  Returns [(flag, bounce, rgba (H, W, 4) uint8)]
        for the last sample, in the order the reference writes its debug-*.png files."""
        self._commit_changes()
        if not self._has_scene:
            raise ErrNoSceneData(_lib.ERR_NO_SCENE_DATA, "no scene data uploaded")
        seeds = np.ascontiguousarray(seeds, dtype=np.uint32)
        w, h = block_req.frame_w, block_req.frame_h
        n = _lib.debug_frame_count(debug_flags, block_req.num_bounces)
        frames = np.zeros((max(n, 1), h, w, 4), dtype=np.uint8)
        infos = np.zeros(max(n, 1), dtype=_lib.DEBUG_FRAME_DTYPE)
        got = ctypes.c_uint32(0)
        st = Stats()
        self._check(self._fn("trace_debug")(self._h, ctypes.byref(block_req), seeds.ctypes.data, seeds.size, int(debug_flags),
                                            frames.ctypes.data, frames.nbytes, infos.ctypes.data, len(infos), ctypes.byref(got),
                                            ctypes.byref(st)))
        self._stats.device = st.as_dict()
        return [(int(infos[i]["flag"]), int(infos[i]["bounce"]), frames[i]) for i in range(got.value)]

    def merge_output(self, other, block_req):
        if type(other) is not type(self):
            raise ErrUnsupportedTracer(_lib.ERR_UNSUPPORTED_TRACER, "merge failed: unsupported tracer instance")
        t0 = time.perf_counter()
        self._check(self._fn("merge_output")(self._h, other._h, ctypes.byref(block_req)))
        return time.perf_counter() - t0

    def sync_framebuffer(self, block_req, want_pixels=True):
        t0 = time.perf_counter()
        if not self._has_scene:
            raise ErrNoSceneData(_lib.ERR_NO_SCENE_DATA, "no scene data uploaded")
        out = self._frame_out(block_req.frame_h, block_req.frame_w) if want_pixels else None
        self._check(self._fn("sync_framebuffer")(self._h, ctypes.byref(block_req), out.ctypes.data if want_pixels else None))
        self.frame_buffer = out
        return time.perf_counter() - t0

    def _frame_out(self, h, w):
        return np.empty((h, w, 4), dtype=np.uint8)

    # -- test hooks
    def set_option(self, option, value):
        self._check(self._fn("set_option")(self._h, option, int(value)))

    def read_buffer(self, which, count, dtype):
        out = np.empty(count, dtype=dtype)
        self._check(self._fn("read_buffer")(self._h, which, out.ctypes.data, out.nbytes))
        return out


class CudaTracer(_HandleTracer):
    """The `cuda` backend: one handle == one B200."""

    _prefix = "pc_"

    def __init__(self, tracer_id: str = "cuda:0", ordinal: int = 0):
        super().__init__(tracer_id)
        self.ordinal = ordinal
        self._lib = None

    def _fn(self, name):
        return getattr(self._lib, "pc_" + name)

    def init(self):
        if self._h is not None:
            return
        self._lib = _lib.load()
        h = ctypes.c_void_p()
        rc = self._lib.pc_create(self.ordinal, self._id.encode(), ctypes.byref(h))
        if rc != 0:
            msg = self._lib.pc_last_error(None)
            raise TracerError(rc, f"cuda tracer: {msg.decode() if msg else rc}")
        self._h = h

    def close(self):
        if self._h is not None:
            if getattr(self, "_pinned_frame", None) is not None:
                self._lib.pc_host_unregister(self._pinned_frame.ctypes.data)
                self._pinned_frame = None
            self._lib.pc_destroy(self._h)
            self._h = None
        self._has_scene = False
        self._keep = None

    def flags(self):
        return int(self._lib.pc_flags(self._h))

    def _frame_out(self, h, w):
        # one page-locked RGBA8 frame per tracer, reused by every SyncFramebuffer (the caller copies it if it keeps it)
        buf = getattr(self, "_pinned_frame", None)
        if buf is None or buf.shape[:2] != (h, w):
            if buf is not None:
                self._lib.pc_host_unregister(buf.ctypes.data)
            buf = np.empty((h, w, 4), dtype=np.uint8)
            self._lib.pc_host_register(buf.ctypes.data, buf.nbytes)
            self._pinned_frame = buf
        return buf

    def speed(self):
        return int(self._lib.pc_speed(self._h))

    def merge_rows(self, rows, is_device, block_req):
        ptr = rows if isinstance(rows, int) else np.ascontiguousarray(rows, dtype=np.float32).ctypes.data
        self._check(self._lib.pc_merge_rows(self._h, ptr, int(is_device), ctypes.byref(block_req)))

    def trace_rows(self, block_req):
        p, n = ctypes.c_void_p(), ctypes.c_uint64()
        self._check(self._lib.pc_trace_rows(self._h, ctypes.byref(block_req), ctypes.byref(p), ctypes.byref(n)))
        return p.value, n.value

    def wait_for_kernels(self):
        self._check(self._lib.pc_wait_for_kernels(self._h))

    def kernel_timings(self):
        """(classes uint32[n], microseconds float32[n]) of every launch of the last trace (PC_OPT_KERNEL_TIMERS)."""
        n = ctypes.c_uint32(0)
        self._check(self._lib.pc_get_kernel_timings(self._h, None, None, 0, ctypes.byref(n)))
        cls, us = np.zeros(n.value, np.uint32), np.zeros(n.value, np.float32)
        if n.value:
            self._check(self._lib.pc_get_kernel_timings(self._h, cls.ctypes.data, us.ctypes.data, n.value, ctypes.byref(n)))
        return cls, us

    # -- one process per GPU: the shared-context reach of device/context.go:11-28 across processes (pc_ipc_*)
    def ipc_export(self, slot: int) -> bytes:
        buf = ctypes.create_string_buffer(64)
        self._check(self._lib.pc_ipc_export(self._h, int(slot), buf))
        return buf.raw

    def ipc_publish_rows(self, block_req, slot: int):
        self._check(self._lib.pc_ipc_publish_rows(self._h, ctypes.byref(block_req), int(slot)))

    def ipc_open(self, handle: bytes) -> int:
        p = ctypes.c_void_p()
        self._check(self._lib.pc_ipc_open(self._h, ctypes.create_string_buffer(handle, 64), ctypes.byref(p)))
        return p.value

    def ipc_close(self, ptr: int):
        self._check(self._lib.pc_ipc_close(self._h, ctypes.c_void_p(ptr)))

    def debug_intersect(self, rays, mode):
        rays = np.ascontiguousarray(rays, dtype=_lib.RAY_DTYPE)
        n = rays.shape[0]
        flags = np.zeros(n, dtype=np.uint32)
        hits = np.zeros(n, dtype=_lib.INTERSECTION_DTYPE)
        self._check(self._lib.pc_debug_intersect(self._h, rays.ctypes.data, n, mode, flags.ctypes.data, hits.ctypes.data))
        return flags, hits

    def debug_bxdf(self, records):
        records = np.ascontiguousarray(records, dtype=_lib.BXDF_IN_DTYPE)
        out = np.zeros(records.shape[0], dtype=_lib.BXDF_OUT_DTYPE)
        self._check(self._lib.pc_debug_bxdf(self._h, records.ctypes.data, records.shape[0], out.ctypes.data))
        return out

    def debug_rng(self, states, draws):
        states = np.ascontiguousarray(states, dtype=np.uint32).copy()
        n = states.shape[0]
        out = np.zeros((n, draws, 2), dtype=np.float32)
        self._check(self._lib.pc_debug_rng(self._h, states.ctypes.data, n, draws, out.ctypes.data))
        return out, states

    def debug_tonemap(self, acc, sample_weight, exposure):
        acc = np.ascontiguousarray(acc, dtype=np.float32).reshape(-1, 4)
        out = np.zeros((acc.shape[0], 4), dtype=np.uint8)
        self._check(self._lib.pc_debug_tonemap(self._h, acc.ctypes.data, acc.shape[0], sample_weight, exposure, out.ctypes.data))
        return out


def device_count() -> int:
    return int(_lib.load().pc_device_count())


def device_info(ordinal: int) -> dict:
    lib = _lib.load()
    name = ctypes.create_string_buffer(256)
    sm, mhz, speed = ctypes.c_uint32(), ctypes.c_uint32(), ctypes.c_uint32()
    rc = lib.pc_device_info(ordinal, name, 256, ctypes.byref(sm), ctypes.byref(mhz), ctypes.byref(speed))
    if rc != 0:
        raise TracerError(rc, f"pc_device_info({ordinal}) failed")
    return {"name": name.value.decode(), "sm_count": sm.value, "clock_mhz": mhz.value, "speed": speed.value}
