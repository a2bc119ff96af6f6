"""The reference's OWN OpenCL program on an OpenCL device of this box (on the GPU box: the B200,
through NVIDIA's OpenCL driver) behind the same Tracer surface as CudaTracer / OracleTracer / RefTracer.

TEST INFRASTRUCTURE: a second, independent checker and the "reference kernels on the same GPU"
baseline of bench.py.  Nothing under polaris_b200/ imports this file.

What runs: the include-expanded text of the reference's tracer/opencl/CL/main.cl, embedded verbatim in
oracle/_ref/libpolaris_clref.so by oracle/build_ref.py, compiled at run time by the device's OpenCL
compiler with the reference's build options (none besides `-I`, device/device.go:133-140), and driven
with the reference's launch discipline restated from
    tracer/opencl/tracer.go:194-286     Trace / MergeOutput / SyncFramebuffer
    tracer/opencl/pipeline.go:94-213    MonteCarloIntegrator (packet query for primary rays on GPUs)
    tracer/opencl/resources.go:81-360   argument order and NDRange extents of every launch
    tracer/opencl/device/kernel.go:89-215   one clFinish after every launch
    tracer/opencl/buffers.go:127-175    frame-sized buffers
The Go host cannot be built here (no Go toolchain, un-vendored cgo OpenCL binding), so the host side
is this file: ctypes over the OpenCL C API.  The image has neither OpenCL headers nor an ICD loader,
only the vendor driver (libnvidia-opencl.so.1); `_Api` therefore is its own minimal ICD loader: it
takes the platform from clIcdGetPlatformIDsKHR and calls through the platform's dispatch table, whose
slot order is fixed by the cl_khr_icd extension (the OpenCL 1.0 block is all this file needs).

Differences from a run of the Go binary that matter when reading results:
  * arithmetic is the DEVICE compiler's: NVIDIA contracts a*b+c into FMA and maps native_recip /
    native_sqrt / native_sin / native_cos to approximate SFU instructions, so values differ from the
    IEEE CPU oracle in the last bits and a few rays take another branch at a discontinuity;
  * bounce >= 1 ray order is atomic arrival order (pt_integrator.cl:162,176,188-197): the
    per-pixel result of one sample is reproducible only through bounce 0;
  * the ray counters are read back once per bounce (12 bytes) to report Mrays/s -- the reference never
    reads them.
"""
from __future__ import annotations

import ctypes
import os
import re
import time

import numpy as np

from polaris_b200 import _lib
from polaris_b200.tracer import (CAMERA_DATA, FRAME_DIMENSIONS, LOCAL, SCENE_DATA, SYNCHRONOUS, ErrNoSceneData,
                                 ErrUnsupportedTracer, Tracer, TracerError, TracerStats)

from . import ref_binding

vp, u32, u64, i32, sz = ctypes.c_void_p, ctypes.c_uint32, ctypes.c_uint64, ctypes.c_int32, ctypes.c_size_t
P = ctypes.POINTER

# cl.h constants used below
CL_DEVICE_TYPE_CPU, CL_DEVICE_TYPE_GPU, CL_DEVICE_TYPE_ALL = 2, 4, 0xFFFFFFFF
CL_MEM_READ_WRITE, CL_MEM_COPY_HOST_PTR = 1, 32
CL_PLATFORM_VERSION, CL_PLATFORM_NAME = 0x0901, 0x0902
CL_DEVICE_TYPE, CL_DEVICE_MAX_COMPUTE_UNITS, CL_DEVICE_MAX_CLOCK_FREQUENCY = 0x1000, 0x1002, 0x100C
CL_DEVICE_NAME, CL_DRIVER_VERSION, CL_DEVICE_VERSION = 0x102B, 0x102D, 0x102F
CL_PROGRAM_BUILD_LOG = 0x1183
CL_INVALID_ARG_SIZE = -51

# slot -> (name, restype, argtypes) in the cl_khr_icd dispatch table (OpenCL 1.0 block)
_DISPATCH = {
    0: ("clGetPlatformIDs", i32, [u32, P(vp), P(u32)]),
    1: ("clGetPlatformInfo", i32, [vp, u32, sz, vp, P(sz)]),
    2: ("clGetDeviceIDs", i32, [vp, u64, u32, P(vp), P(u32)]),
    3: ("clGetDeviceInfo", i32, [vp, u32, sz, vp, P(sz)]),
    4: ("clCreateContext", vp, [vp, u32, P(vp), vp, vp, P(i32)]),
    7: ("clReleaseContext", i32, [vp]),
    9: ("clCreateCommandQueue", vp, [vp, vp, u64, P(i32)]),
    11: ("clReleaseCommandQueue", i32, [vp]),
    14: ("clCreateBuffer", vp, [vp, u64, sz, vp, P(i32)]),
    18: ("clReleaseMemObject", i32, [vp]),
    26: ("clCreateProgramWithSource", vp, [vp, u32, P(ctypes.c_char_p), P(sz), P(i32)]),
    29: ("clReleaseProgram", i32, [vp]),
    30: ("clBuildProgram", i32, [vp, u32, P(vp), ctypes.c_char_p, vp, vp]),
    33: ("clGetProgramBuildInfo", i32, [vp, vp, u32, sz, vp, P(sz)]),
    34: ("clCreateKernel", vp, [vp, ctypes.c_char_p, P(i32)]),
    37: ("clReleaseKernel", i32, [vp]),
    38: ("clSetKernelArg", i32, [vp, u32, sz, vp]),
    47: ("clFinish", i32, [vp]),
    48: ("clEnqueueReadBuffer", i32, [vp, vp, u32, sz, sz, vp, u32, vp, vp]),
    49: ("clEnqueueWriteBuffer", i32, [vp, vp, u32, sz, sz, vp, u32, vp, vp]),
    59: ("clEnqueueNDRangeKernel", i32, [vp, vp, u32, P(sz), P(sz), P(sz), u32, vp, vp]),
}
_LIB_CANDIDATES = ("libOpenCL.so.1", "libOpenCL.so", "libnvidia-opencl.so.1", "/usr/lib/libnvidia-opencl.so.1",
                   "/usr/local/nvidia/lib/libnvidia-opencl.so.1", "/usr/lib/x86_64-linux-gnu/libnvidia-opencl.so.1")


class ClError(RuntimeError):
    pass


class _Api:
    """OpenCL entry points, either exported by a real ICD loader (libOpenCL) or taken from the vendor
    driver's dispatch table."""

    def __init__(self):
        self.lib = None
        self.how = None
        self.platform = None
        errors = []
        for cand in _LIB_CANDIDATES + tuple(filter(None, [os.environ.get("POLARIS_OPENCL_LIB")])):
            try:
                lib = ctypes.CDLL(cand)
            except OSError as e:
                errors.append(f"{cand}: {e}")
                continue
            try:
                if self._bind(lib, cand):
                    return
            except Exception as e:  # keep looking
                errors.append(f"{cand}: {e}")
        raise ClError("no usable OpenCL driver: " + "; ".join(errors[-4:]))

    def _bind(self, lib, path):
        n = u32(0)
        plats = (vp * 8)()
        icd = getattr(lib, "clIcdGetPlatformIDsKHR", None)
        if icd is None and hasattr(lib, "clGetExtensionFunctionAddress"):
            lib.clGetExtensionFunctionAddress.restype = vp
            lib.clGetExtensionFunctionAddress.argtypes = [ctypes.c_char_p]
            addr = lib.clGetExtensionFunctionAddress(b"clIcdGetPlatformIDsKHR")
            if addr:
                icd = ctypes.CFUNCTYPE(None)(addr)
        if icd is not None:
            icd.restype, icd.argtypes = i32, [u32, P(vp), P(u32)]
            rc = icd(8, plats, ctypes.byref(n))
            if rc == 0 and n.value > 0:
                self.platform = vp(plats[0])
                table = ctypes.cast(ctypes.cast(self.platform, P(vp))[0], P(vp))  # platform->dispatch
                for slot, (name, res, args) in _DISPATCH.items():
                    fn = ctypes.CFUNCTYPE(res, *args)(table[slot])
                    setattr(self, name, fn)
                self.lib, self.how = lib, f"{path} via clIcdGetPlatformIDsKHR + dispatch table"
                return True
        if hasattr(lib, "clGetPlatformIDs"):  # a real ICD loader
            for _, (name, res, args) in _DISPATCH.items():
                fn = getattr(lib, name)
                fn.restype, fn.argtypes = res, args
                setattr(self, name, fn)
            rc = self.clGetPlatformIDs(8, plats, ctypes.byref(n))
            if rc == 0 and n.value > 0:
                self.platform = vp(plats[0])
                self.lib, self.how = lib, f"{path} (exported entry points)"
                return True
        return False

    def info_str(self, fn, obj, what):
        buf = ctypes.create_string_buffer(1024)
        got = sz(0)
        rc = fn(obj, what, 1024, buf, ctypes.byref(got))
        return buf.value.decode(errors="replace") if rc == 0 else f"<error {rc}>"

    def info_u32(self, fn, obj, what):
        v = u32(0)
        rc = fn(obj, what, 4, ctypes.byref(v), None)
        return v.value if rc == 0 else 0


_api = None


def api() -> _Api:
    global _api
    if _api is None:
        _api = _Api()
    return _api


def available() -> bool:
    """True when the embedded program text and an OpenCL device are both present."""
    if not ref_binding.available():
        return False
    try:
        lib = ref_binding.load()
        if not hasattr(lib, "pr_cl_program_source"):
            return False
        a = api()
        n = u32(0)
        return a.clGetDeviceIDs(a.platform, CL_DEVICE_TYPE_ALL, 0, None, ctypes.byref(n)) == 0 and n.value > 0
    except Exception:
        return False


# C++-style functional casts -- `uint(0)`, `int(x)`, `uchar(x)` -- appear at 14 places of the reference's program
# (texture_sampler.cl:27-30,118-121,201-204, emissive_sampler.cl:236, debug.cl:46).  They are not OpenCL C: the
# compiler the reference was developed against (Apple's) accepts them, NVIDIA's rejects them ("unexpected type
# name 'uint': expected expression", profiles/cl_probe_r01.txt).  `T(x)` -> `(T)(x)` is the same conversion.
_FUNCTIONAL_CAST = re.compile(rb"(?<![\w)])\b(uint|int|uchar)\(")


def program_source(portable=True) -> bytes:
    lib = ref_binding.load()
    lib.pr_cl_program_source.restype = ctypes.c_void_p
    lib.pr_cl_program_source.argtypes = [P(u64)]
    n = u64(0)
    p = lib.pr_cl_program_source(ctypes.byref(n))
    src = ctypes.string_at(p, n.value)
    if portable:
        src = _FUNCTIONAL_CAST.sub(lambda m: b"(" + m.group(1) + b")(", src)
    return src


_KERNELS = ("clearAccumulator", "aggregateAccumulator", "generatePrimaryRays", "rayIntersectionTest", "rayIntersectionQuery",
            "rayPacketIntersectionQuery", "shadeHits", "shadePrimaryRayMisses", "shadeIndirectRayMisses",
            "accumulateEmissiveSamples", "tonemapSimpleReinhard", "debugClearBuffer", "debugRayIntersectionDepth",
            "debugRayIntersectionNormals", "debugEmissiveSamples", "debugThroughput", "debugAccumulator")


class _Buf:
    def __init__(self, dev, nbytes, init=None):
        a = dev.api
        self.nbytes = max(16, int(nbytes))
        err = i32(0)
        if init is not None and len(init):
            host = np.ascontiguousarray(init).view(np.uint8).reshape(-1)
            if host.nbytes < self.nbytes:
                host = np.concatenate([host, np.zeros(self.nbytes - host.nbytes, np.uint8)])
            self.h = a.clCreateBuffer(dev.ctx, CL_MEM_READ_WRITE | CL_MEM_COPY_HOST_PTR, self.nbytes, host.ctypes.data, ctypes.byref(err))
        else:
            self.h = a.clCreateBuffer(dev.ctx, CL_MEM_READ_WRITE, self.nbytes, None, ctypes.byref(err))
        if err.value != 0 or not self.h:
            raise ClError(f"clCreateBuffer({self.nbytes}) failed: {err.value}")
        self.dev = dev
        if init is None:
            self.write(np.zeros(self.nbytes, np.uint8))

    def write(self, arr, offset=0):
        arr = np.ascontiguousarray(arr)
        rc = self.dev.api.clEnqueueWriteBuffer(self.dev.queue, self.h, 1, offset, arr.nbytes, arr.ctypes.data, 0, None, None)
        if rc != 0:
            raise ClError(f"clEnqueueWriteBuffer failed: {rc}")

    def read(self, nbytes, dtype=np.uint8, offset=0):
        out = np.empty(nbytes // np.dtype(dtype).itemsize, dtype=dtype)
        rc = self.dev.api.clEnqueueReadBuffer(self.dev.queue, self.h, 1, offset, out.nbytes, out.ctypes.data, 0, None, None)
        if rc != 0:
            raise ClError(f"clEnqueueReadBuffer failed: {rc}")
        return out

    def release(self):
        if self.h:
            self.dev.api.clReleaseMemObject(self.h)
            self.h = None


class ClDevice:
    """device.Device (tracer/opencl/device/device.go): context, in-order queue, the built program."""

    def __init__(self, prefer_gpu=True, build_options=b""):
        a = self.api = api()
        n = u32(0)
        devs = (vp * 16)()
        kind = CL_DEVICE_TYPE_GPU if prefer_gpu else CL_DEVICE_TYPE_ALL
        rc = a.clGetDeviceIDs(a.platform, kind, 16, devs, ctypes.byref(n))
        if rc != 0 or n.value == 0:
            rc = a.clGetDeviceIDs(a.platform, CL_DEVICE_TYPE_ALL, 16, devs, ctypes.byref(n))
        if rc != 0 or n.value == 0:
            raise ClError(f"clGetDeviceIDs: {rc}, {n.value} devices")
        self.device = vp(devs[0])
        self.name = a.info_str(a.clGetDeviceInfo, self.device, CL_DEVICE_NAME)
        self.platform_name = a.info_str(a.clGetPlatformInfo, a.platform, CL_PLATFORM_NAME)
        self.platform_version = a.info_str(a.clGetPlatformInfo, a.platform, CL_PLATFORM_VERSION)
        self.driver_version = a.info_str(a.clGetDeviceInfo, self.device, CL_DRIVER_VERSION)
        self.compute_units = a.info_u32(a.clGetDeviceInfo, self.device, CL_DEVICE_MAX_COMPUTE_UNITS)
        self.clock_mhz = a.info_u32(a.clGetDeviceInfo, self.device, CL_DEVICE_MAX_CLOCK_FREQUENCY)
        dt = u64(0)
        a.clGetDeviceInfo(self.device, CL_DEVICE_TYPE, 8, ctypes.byref(dt), None)
        self.is_gpu = bool(dt.value & CL_DEVICE_TYPE_GPU)
        err = i32(0)
        dev_arr = (vp * 1)(self.device)
        self.ctx = a.clCreateContext(None, 1, dev_arr, None, None, ctypes.byref(err))
        if err.value != 0:
            raise ClError(f"clCreateContext: {err.value}")
        self.queue = a.clCreateCommandQueue(self.ctx, self.device, 0, ctypes.byref(err))  # device.go:93: in-order, no profiling
        if err.value != 0:
            raise ClError(f"clCreateCommandQueue: {err.value}")
        src = program_source()
        srcs = (ctypes.c_char_p * 1)(src)
        lens = (sz * 1)(len(src))
        t0 = time.perf_counter()
        self.program = a.clCreateProgramWithSource(self.ctx, 1, srcs, lens, ctypes.byref(err))
        if err.value != 0:
            raise ClError(f"clCreateProgramWithSource: {err.value}")
        rc = a.clBuildProgram(self.program, 1, dev_arr, build_options, None, None)
        log = ctypes.create_string_buffer(1 << 16)
        a.clGetProgramBuildInfo(self.program, self.device, CL_PROGRAM_BUILD_LOG, 1 << 16, log, None)
        self.build_log = log.value.decode(errors="replace")
        self.build_seconds = time.perf_counter() - t0
        if rc != 0:
            raise ClError(f"clBuildProgram: {rc}\n{self.build_log[:6000]}")
        self.kernels = {}
        for k in _KERNELS:
            h = a.clCreateKernel(self.program, k.encode(), ctypes.byref(err))
            if err.value != 0:
                raise ClError(f"clCreateKernel({k}): {err.value}")
            self.kernels[k] = h
        self.launches = 0

    def describe(self):
        return {"platform": self.platform_name, "platform_version": self.platform_version, "device": self.name,
                "driver": self.driver_version, "compute_units": self.compute_units, "clock_mhz": self.clock_mhz,
                "loader": self.api.how, "build_seconds": round(self.build_seconds, 2)}

    # Kernel.SetArgs (device/kernel.go:34-83): buffers by handle, scalars by value, Vec3 as 12 bytes
    def set_args(self, kernel, args):
        a, k = self.api, self.kernels[kernel]
        for idx, arg in enumerate(args):
            if isinstance(arg, _Buf):
                h = vp(arg.h)
                rc = a.clSetKernelArg(k, idx, 8, ctypes.byref(h))
            else:
                v = np.ascontiguousarray(arg)
                rc = a.clSetKernelArg(k, idx, v.nbytes, v.ctypes.data)
                if rc == CL_INVALID_ARG_SIZE and v.nbytes == 12:
                    # the reference passes types.Vec3 as 12 bytes (kernel.go:56-58); a conformant runtime wants
                    # sizeof(cl_float3) == 16 for a float3 argument
                    v = np.concatenate([v.view(np.uint8).reshape(-1), np.zeros(4, np.uint8)])
                    rc = a.clSetKernelArg(k, idx, 16, v.ctypes.data)
            if rc != 0:
                raise ClError(f"clSetKernelArg({kernel}, {idx}): {rc}")

    # Kernel.Exec1D / Exec2D (device/kernel.go:89-215): enqueue + clFinish
    def exec(self, kernel, global_size, local_size=None, offset=None, wait=True):
        a = self.api
        dim = len(global_size)
        g = (sz * dim)(*global_size)
        loc = (sz * dim)(*local_size) if local_size else None
        off = (sz * dim)(*offset) if offset and any(offset) else None
        rc = a.clEnqueueNDRangeKernel(self.queue, self.kernels[kernel], dim, off, g, loc, 0, None, None)
        if rc != 0:
            raise ClError(f"clEnqueueNDRangeKernel({kernel}, {tuple(global_size)}): {rc}")
        self.launches += 1
        if wait:
            self.finish(kernel)

    def finish(self, what=""):
        rc = self.api.clFinish(self.queue)
        if rc != 0:
            raise ClError(f"clFinish after {what}: {rc}")

    def close(self):
        a = self.api
        for h in self.kernels.values():
            a.clReleaseKernel(h)
        self.kernels = {}
        if self.program:
            a.clReleaseProgram(self.program)
            self.program = None
        if self.queue:
            a.clReleaseCommandQueue(self.queue)
            self.queue = None
        if self.ctx:
            a.clReleaseContext(self.ctx)
            self.ctx = None


_f32, _u32 = np.float32, np.uint32


class ClDeviceTracer(Tracer):
    """opencl.Tracer (tracer/opencl/tracer.go) restated over ClDevice."""

    def __init__(self, tracer_id="opencl:0", primary_packets=None, build_options=b""):
        self._id = tracer_id
        self.dev = None
        self._opts = build_options
        self._packets = primary_packets  # None: like the reference, packets iff the device is a GPU (pipeline.go:107-111)
        self._stats = TracerStats()
        self._change_buffer = {}
        self.b = {}
        self.scene = {}
        self.w = self.h = 0
        self._has_scene = False
        self._scene = None
        self._cam = None
        self.frame_buffer = None

    # -- tracer.Tracer
    def id(self):
        return self._id

    def flags(self):
        return LOCAL

    def speed(self):
        return self.dev.compute_units * self.dev.clock_mhz // 1000  # device.go:209-222

    def init(self):
        if self.dev is None:
            self.dev = ClDevice(build_options=self._opts)
            if self._packets is None:
                self._packets = self.dev.is_gpu

    def close(self):
        for d in (self.b, self.scene):
            for buf in d.values():
                buf.release()
            d.clear()
        if self.dev is not None:
            self.dev.close()
            self.dev = None
        self._has_scene = False

    def stats(self):
        return self._stats

    def update_state(self, mode, change_type, data):
        self._change_buffer[change_type] = data
        if mode == SYNCHRONOUS:
            return self._commit_changes()
        return 0.0

    def _commit_changes(self):
        t0 = time.perf_counter()
        for change_type, data in list(self._change_buffer.items()):
            if change_type == FRAME_DIMENSIONS:
                self._resize(*data)
            elif change_type == SCENE_DATA:
                self._upload_scene(data)
            elif change_type == CAMERA_DATA:
                self._cam = (np.ascontiguousarray(data.position, dtype=_f32).copy(),
                             np.ascontiguousarray(data.frustrum, dtype=_f32).reshape(4, 4).copy())
            else:
                raise TracerError(_lib.ERR_INVALID_ARGUMENT, f"unsupported change type {change_type}")
        self._change_buffer = {}
        return time.perf_counter() - t0

    def _resize(self, w, h):  # bufferSet.Resize (buffers.go:127-175)
        for buf in self.b.values():
            buf.release()
        px = int(w) * int(h)
        self.w, self.h = int(w), int(h)
        d = self.dev
        self.b = {"rays0": _Buf(d, px * 32), "rays1": _Buf(d, px * 32), "rays2": _Buf(d, px * 32), "paths": _Buf(d, px * 32),
                  "hitFlags": _Buf(d, px * 4), "intersections": _Buf(d, px * 32), "emissiveSamples": _Buf(d, px * 16),
                  "traceAcc": _Buf(d, px * 16), "frameAcc": _Buf(d, px * 16), "frameBuffer": _Buf(d, px * 4), "debugOutput": _Buf(d, px * 4),
                  "cnt0": _Buf(d, 4), "cnt1": _Buf(d, 4), "cnt2": _Buf(d, 4)}

    def _upload_scene(self, sc):  # bufferSet.UploadSceneData (buffers.go:178-201)
        for buf in self.scene.values():
            buf.release()
        view, keep = _lib.scene_view(sc)
        d = self.dev

        def up(ptr, nbytes):
            host = np.frombuffer(ctypes.string_at(ptr, nbytes), dtype=np.uint8) if nbytes else np.zeros(0, np.uint8)
            return _Buf(d, nbytes, host)

        self.scene = {
            "bvhNodes": up(view.bvh_nodes, view.bvh_nodes_bytes), "meshInstances": up(view.mesh_instances, view.mesh_instances_bytes),
            "materialNodes": up(view.material_nodes, view.material_nodes_bytes), "textures": up(view.texture_data, view.texture_data_bytes),
            "textureMetadata": up(view.texture_metadata, view.texture_metadata_bytes), "vertices": up(view.vertices, view.vertices_bytes),
            "normals": up(view.normals, view.normals_bytes), "uv": up(view.uvs, view.uvs_bytes),
            "materialIndices": up(view.material_indices, view.material_indices_bytes), "emissives": up(view.emissives, view.emissives_bytes)}
        self.num_emissives = int(view.emissives_bytes // 80)
        self.scene_diffuse = int(view.scene_diffuse_mat_index)
        self._has_scene = True
        del keep

    # -- launches (resources.go)
    def _counters(self):
        return [int(self.b[f"cnt{i}"].read(4, np.int32)[0]) for i in range(3)]

    def _query(self, buf, num_pixels, packets=False):
        d, b, s = self.dev, self.b, self.scene
        name = "rayPacketIntersectionQuery" if packets else "rayIntersectionQuery"
        d.set_args(name, [b[f"rays{buf}"], b[f"cnt{buf}"], s["bvhNodes"], s["meshInstances"], s["vertices"], b["hitFlags"], b["intersections"]])
        d.exec(name, [num_pixels], [32] if packets else None)

    def _test(self, buf, num_pixels):
        d, b, s = self.dev, self.b, self.scene
        d.set_args("rayIntersectionTest", [b[f"rays{buf}"], b[f"cnt{buf}"], s["bvhNodes"], s["meshInstances"], s["vertices"], b["hitFlags"]])
        d.exec("rayIntersectionTest", [num_pixels])

    # -- debug stages (resources.go:362-520, pipeline.go:113-200)
    def _debug_stage(self, req, flag, bounce, a, frames):
        d, b, s = self.dev, self.b, self.scene
        px, n = self.w * self.h, int(req.frame_w) * int(req.block_h)
        d.set_args("debugClearBuffer", [b["debugOutput"]])
        d.exec("debugClearBuffer", [px])
        if flag == _lib.DEBUG_PRIMARY_DEPTH:
            t = b["intersections"].read(px * 32, _lib.INTERSECTION_DTYPE)["wuvt"][:, 3]  # ReadDataIntoSlice of the whole buffer
            finite = t[t != np.finfo(np.float32).max]
            max_depth = _f32(max(1.0, float(finite.max()))) if finite.size else _f32(1.0)
            d.set_args("debugRayIntersectionDepth", [b[f"cnt{a}"], b["paths"], b["hitFlags"], b["intersections"], max_depth, b["debugOutput"]])
            d.exec("debugRayIntersectionDepth", [n])
        elif flag == _lib.DEBUG_PRIMARY_NORMALS:
            d.set_args("debugRayIntersectionNormals", [b[f"rays{a}"], b[f"cnt{a}"], b["paths"], b["hitFlags"], b["intersections"], s["vertices"],
                                                       s["normals"], s["uv"], s["materialIndices"], s["materialNodes"], s["textureMetadata"],
                                                       s["textures"], b["debugOutput"]])
            d.exec("debugRayIntersectionNormals", [n])
        elif flag in (_lib.DEBUG_ALL_EMISSIVE, _lib.DEBUG_VISIBLE_EMISSIVE, _lib.DEBUG_OCCLUDED_EMISSIVE):
            d.set_args("debugEmissiveSamples", [b["rays2"], b["cnt2"], b["paths"], b["hitFlags"], b["emissiveSamples"],
                                                _u32(flag == _lib.DEBUG_VISIBLE_EMISSIVE), _u32(flag == _lib.DEBUG_OCCLUDED_EMISSIVE), b["debugOutput"]])
            d.exec("debugEmissiveSamples", [n])
        elif flag == _lib.DEBUG_THROUGHPUT:
            d.set_args("debugThroughput", [b["paths"], b["debugOutput"]])
            d.exec("debugThroughput", [n])
        elif flag == _lib.DEBUG_ACCUMULATOR:
            weight = _f32(1.0 / _f32(int(req.accumulated_samples) + int(req.samples_per_pixel)))
            d.set_args("debugAccumulator", [weight, b["paths"], b["traceAcc"], b["debugOutput"]])
            d.exec("debugAccumulator", [n])
        if frames is not None:
            frames.append((flag, bounce, b["debugOutput"].read(px * 4).reshape(self.h, self.w, 4)))

    def trace_debug(self, req, seeds, debug_flags):
        frames = []
        self.trace(req, seeds, debug_flags=debug_flags, frames=frames)
        return frames

    def trace(self, req, seeds=None, debug_flags=0, frames=None):
        t0 = time.perf_counter()
        self._commit_changes()
        if not self._has_scene:
            raise ErrNoSceneData(_lib.ERR_NO_SCENE_DATA, "no scene data uploaded")  # tracer.go:203-205
        d, b, s = self.dev, self.b, self.scene
        nb, spp = int(req.num_bounces), int(req.samples_per_pixel)
        per_sample = 1 + nb
        if seeds is None:
            seeds = np.random.default_rng().integers(0, 1 << 32, per_sample * spp, dtype=np.uint64).astype(_u32)
        seeds = np.ascontiguousarray(seeds, dtype=_u32)
        px = int(req.frame_w) * int(req.frame_h)
        launches0 = d.launches
        if req.accumulated_samples == 0:  # tracer.go:208-213
            d.set_args("clearAccumulator", [b["frameAcc"]])
            d.exec("clearAccumulator", [px])
        d.set_args("clearAccumulator", [b["traceAcc"]])  # tracer.go:215
        d.exec("clearAccumulator", [px])
        num_pixels = int(req.frame_w) * int(req.block_h)  # pipeline.go:96
        eye, fr = self._cam
        texel = np.array([1.0 / _f32(req.frame_w), 1.0 / _f32(req.frame_h)], dtype=_f32)
        zero = np.zeros(1, np.int32)
        q_rays = o_rays = occ_emitted = ind_emitted = 0
        for sample in range(spp):  # tracer.go:221
            ss = seeds[per_sample * sample: per_sample * (sample + 1)]
            d.set_args("generatePrimaryRays", [b["rays0"], b["cnt0"], b["paths"], fr[0], fr[1], fr[2], fr[3], eye, texel,
                                               _u32(req.block_y), _u32(req.block_h), _u32(req.frame_w), _u32(req.frame_h), _u32(ss[0])])
            d.exec("generatePrimaryRays", [int(req.frame_w), int(req.block_h)])
            a = 0
            self._query(a, num_pixels, self._packets and num_pixels % 32 == 0)
            q_rays += num_pixels
            keep = frames if sample + 1 == spp else None  # the reference overwrites its PNGs every sample
            for f in (_lib.DEBUG_PRIMARY_DEPTH, _lib.DEBUG_PRIMARY_NORMALS):
                if debug_flags & f:
                    self._debug_stage(req, f, 0, a, keep)
            for bounce in range(nb):  # pipeline.go:132
                if self.scene_diffuse != -1:
                    name = "shadePrimaryRayMisses" if bounce == 0 else "shadeIndirectRayMisses"
                    d.set_args(name, [b[f"rays{a}"], b[f"cnt{a}"], b["paths"], b["hitFlags"], s["materialNodes"], _u32(self.scene_diffuse),
                                      s["textureMetadata"], s["textures"], b["traceAcc"]])
                    d.exec(name, [num_pixels])
                b["cnt2"].write(zero)  # resources.go:230-238: two blocking counter writes
                b[f"cnt{1 - a}"].write(zero)
                d.set_args("shadeHits", [b[f"rays{a}"], b[f"cnt{a}"], b["paths"], b["hitFlags"], b["intersections"], s["vertices"], s["normals"],
                                         s["uv"], s["materialIndices"], s["materialNodes"], s["emissives"], _u32(self.num_emissives),
                                         s["textureMetadata"], s["textures"], _u32(bounce), _u32(req.min_bounces_for_rr), _u32(ss[1 + bounce]),
                                         b["rays2"], b["cnt2"], b["emissiveSamples"], b[f"rays{1 - a}"], b[f"cnt{1 - a}"], b["traceAcc"]])
                d.exec("shadeHits", [num_pixels])
                if debug_flags & _lib.DEBUG_THROUGHPUT:
                    self._debug_stage(req, _lib.DEBUG_THROUGHPUT, bounce, a, keep)
                cnt = self._counters()  # not in the reference: only to report rays
                occ_emitted += cnt[2]
                ind_emitted += cnt[1 - a]
                self._test(2, num_pixels)  # pipeline.go:160
                o_rays += cnt[2]
                d.set_args("accumulateEmissiveSamples", [b["rays2"], b["cnt2"], b["paths"], b["hitFlags"], b["emissiveSamples"], b["traceAcc"]])
                d.exec("accumulateEmissiveSamples", [num_pixels])  # :165
                for f in (_lib.DEBUG_ALL_EMISSIVE, _lib.DEBUG_VISIBLE_EMISSIVE, _lib.DEBUG_OCCLUDED_EMISSIVE, _lib.DEBUG_ACCUMULATOR):
                    if debug_flags & f:
                        self._debug_stage(req, f, bounce, a, keep)
                if bounce + 1 < nb:  # :203-209
                    a = 1 - a
                    self._query(a, num_pixels)
                    q_rays += cnt[a]
            req.accumulated_samples += 1  # tracer.go:240
        dt = time.perf_counter() - t0
        self._stats.block_w, self._stats.block_h, self._stats.render_time = req.block_w, req.block_h, dt
        self._stats.device = {"query_rays": q_rays, "occlusion_rays": o_rays, "occlusion_emitted": occ_emitted,
                              "indirect_emitted": ind_emitted, "kernel_launches": d.launches - launches0}
        return dt

    def merge_output(self, other, req):  # tracer.go:278-286, resources.go:108-124 (Exec1DNoWait)
        if type(other) is not type(self):
            raise ErrUnsupportedTracer(_lib.ERR_UNSUPPORTED_TRACER, "merge failed: unsupported tracer instance")
        t0 = time.perf_counter()
        self.dev.set_args("aggregateAccumulator", [other.b["traceAcc"], self.b["frameAcc"]])
        self.dev.exec("aggregateAccumulator", [int(req.block_w) * int(req.block_h)], offset=[int(req.frame_w) * int(req.block_y)], wait=False)
        return time.perf_counter() - t0

    def sync_framebuffer(self, req, want_pixels=True):  # tracer.go:250-276, resources.go:344-360
        t0 = time.perf_counter()
        if not self._has_scene:
            raise ErrNoSceneData(_lib.ERR_NO_SCENE_DATA, "no scene data uploaded")
        self.dev.finish("merge")  # WaitForKernels (tracer.go:259)
        weight = _f32(1.0 / _f32(int(req.accumulated_samples) + int(req.samples_per_pixel)))
        self.dev.set_args("tonemapSimpleReinhard", [self.b["frameAcc"], self.b["paths"], self.b["frameBuffer"], weight, _f32(req.exposure)])
        self.dev.exec("tonemapSimpleReinhard", [int(req.frame_w) * int(req.block_h)])
        self.frame_buffer = self.b["frameBuffer"].read(self.w * self.h * 4).reshape(self.h, self.w, 4) if want_pixels else None
        return time.perf_counter() - t0

    # -- test hooks (same shape as the other tracers')
    def set_option(self, option, value):
        if option == _lib.OPT_PRIMARY_PACKETS:
            self._packets = bool(value)

    def read_buffer(self, which, count, dtype):
        names = {_lib.BUF_RAYS0: "rays0", _lib.BUF_RAYS1: "rays1", _lib.BUF_RAYS2: "rays2", _lib.BUF_PATHS: "paths",
                 _lib.BUF_HIT_FLAGS: "hitFlags", _lib.BUF_INTERSECTIONS: "intersections", _lib.BUF_EMISSIVE_SAMPLES: "emissiveSamples",
                 _lib.BUF_TRACE_ACCUMULATOR: "traceAcc", _lib.BUF_FRAME_ACCUMULATOR: "frameAcc", _lib.BUF_FRAME_BUFFER: "frameBuffer"}
        if which == _lib.BUF_RAY_COUNTERS:
            return np.array(self._counters(), dtype=np.int32)[:count].astype(dtype)
        nbytes = count * np.dtype(dtype).itemsize
        return self.b[names[which]].read(nbytes, dtype)

    def debug_intersect(self, rays, mode):
        """mode 0: rayIntersectionQuery, 1: rayIntersectionTest, 2: rayPacketIntersectionQuery over caller-supplied rays."""
        rays = np.ascontiguousarray(rays, dtype=_lib.RAY_DTYPE)
        n = rays.shape[0]
        assert n <= self.w * self.h, "ray set larger than the frame-sized ray buffer"
        self.b["rays0"].write(rays)
        self.b["cnt0"].write(np.array([n], np.int32))
        self.b["hitFlags"].write(np.zeros(n, _u32))
        self.b["intersections"].write(np.zeros(n * 32, np.uint8))
        if mode == 1:
            self._test(0, n)
        else:
            g = (n + 31) // 32 * 32 if mode == 2 else n
            self._query(0, g, packets=(mode == 2))
        return self.b["hitFlags"].read(n * 4, _u32), self.b["intersections"].read(n * 32, _lib.INTERSECTION_DTYPE)
