"""ctypes view of include/polaris_cuda.h and loader of libpolaris_cuda.so.

There is no fallback: if the CUDA library is missing or fails to load, `load()` raises.
"""
from __future__ import annotations

import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("POLARIS_CUDA_LIB") or os.path.join(_HERE, "libpolaris_cuda.so")  # env override: A/B builds only

u32, u64, i32, f32 = ctypes.c_uint32, ctypes.c_uint64, ctypes.c_int32, ctypes.c_float
vp = ctypes.c_void_p


class BlockRequest(ctypes.Structure):
    """pc_block_request == tracer.BlockRequest (reference tracer/tracer.go:6-34)."""

    _fields_ = [("frame_w", u32), ("frame_h", u32), ("block_x", u32), ("block_y", u32),
                ("block_w", u32), ("block_h", u32), ("samples_per_pixel", u32), ("num_bounces", u32),
                ("min_bounces_for_rr", u32), ("exposure", f32), ("seed", u32), ("accumulated_samples", u32)]

    def copy(self):
        c = BlockRequest()
        ctypes.memmove(ctypes.byref(c), ctypes.byref(self), ctypes.sizeof(self))
        return c


class Stats(ctypes.Structure):
    """pc_stats: tracer.Stats (tracer/tracer.go:37-47) + device counters."""

    _fields_ = [("block_w", u32), ("block_h", u32), ("update_time_ns", u64), ("render_time_ns", u64),
                ("device_time_ns", u64), ("query_rays", u64), ("occlusion_rays", u64), ("kernel_launches", u64),
                ("nodes_tested", u64), ("tris_tested", u64), ("instances_entered", u64), ("shaded_hits", u64),
                ("occlusion_emitted", u64), ("indirect_emitted", u64), ("unoccluded", u64), ("missed_query_rays", u64),
                ("kernel_time_ns", u64 * 8), ("kernel_count", u64 * 8)]

    def as_dict(self):
        d = {n: getattr(self, n) for n, _ in self._fields_}
        d["kernel_time_ns"] = list(d["kernel_time_ns"])
        d["kernel_count"] = list(d["kernel_count"])
        return d


class SceneView(ctypes.Structure):
    _fields_ = [("bvh_nodes", vp), ("bvh_nodes_bytes", u64),
                ("mesh_instances", vp), ("mesh_instances_bytes", u64),
                ("material_nodes", vp), ("material_nodes_bytes", u64),
                ("texture_data", vp), ("texture_data_bytes", u64),
                ("texture_metadata", vp), ("texture_metadata_bytes", u64),
                ("vertices", vp), ("vertices_bytes", u64),
                ("normals", vp), ("normals_bytes", u64),
                ("uvs", vp), ("uvs_bytes", u64),
                ("material_indices", vp), ("material_indices_bytes", u64),
                ("emissives", vp), ("emissives_bytes", u64),
                ("scene_diffuse_mat_index", i32), ("scene_emissive_mat_index", i32)]


assert ctypes.sizeof(BlockRequest) == 48

# pc_buffer
BUF_RAYS0, BUF_RAYS1, BUF_RAYS2, BUF_PATHS, BUF_HIT_FLAGS, BUF_INTERSECTIONS = 0, 1, 2, 3, 4, 5
BUF_EMISSIVE_SAMPLES, BUF_TRACE_ACCUMULATOR, BUF_FRAME_ACCUMULATOR, BUF_FRAME_BUFFER, BUF_RAY_COUNTERS = 6, 7, 8, 9, 10
# pc_option
OPT_COUNTERS, OPT_PRIMARY_PACKETS, OPT_REFERENCE_ORDER, OPT_USE_GRAPH, OPT_FIX_Q4, OPT_KERNEL_TIMERS, OPT_SAMPLE_CHAINS, OPT_FUSE_TRACE, OPT_SORT_RAYS, OPT_DEFER_OCCLUSION, OPT_TRACE_REFILL, OPT_SAMPLE_SLOTS = 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11
K_BEGIN_SAMPLE, K_PRIMARY, K_SHADE, K_OCCLUSION, K_QUERY, K_TRACE = 0, 1, 2, 3, 4, 5
KERNEL_CLASS_NAMES = ["k_begin_sample", "k_primary", "k_shade", "k_occlusion", "k_query", "k_trace"]
# pc_debug_flag == opencl.DebugFlag (tracer/opencl/pipeline.go:17-30)
DEBUG_PRIMARY_DEPTH, DEBUG_PRIMARY_NORMALS, DEBUG_ALL_EMISSIVE, DEBUG_VISIBLE_EMISSIVE = 2, 4, 8, 16
DEBUG_OCCLUDED_EMISSIVE, DEBUG_THROUGHPUT, DEBUG_ACCUMULATOR, DEBUG_FRAMEBUFFER = 32, 64, 128, 256
DEBUG_ALL_STAGES = 2 | 4 | 8 | 16 | 32 | 64 | 128
# file names the reference gives the dumps (pipeline.go:115-196)
DEBUG_FILE_NAMES = {2: "debug-primary-intersection-depth.png", 4: "debug-primary-intersection-normals.png", 8: "debug-emissive-all-%03d.png",
                    16: "debug-emissive-vis-%03d.png", 32: "debug-emissive-occ-%03d.png", 64: "debug-throughput-%03d.png",
                    128: "debug-accumulator-%03d.png"}
DEBUG_FRAME_DTYPE = np.dtype([("flag", np.uint32), ("bounce", np.uint32)])


def debug_frame_count(flags: int, num_bounces: int) -> int:
    return bin(flags & 6).count("1") + num_bounces * bin(flags & (8 | 16 | 32 | 64 | 128)).count("1")


# pc_status
OK = 0
ERR_INVALID_ARGUMENT, ERR_NO_DEVICE, ERR_ALLOC, ERR_COPY_TO_DEVICE, ERR_COPY_TO_HOST, ERR_KERNEL = 1, 2, 3, 4, 5, 6
ERR_NO_SCENE_DATA, ERR_NO_FRAME, ERR_UNSUPPORTED_TRACER, ERR_STACK_DEPTH, ERR_BAD_SCENE, ERR_PEER_ACCESS = 7, 8, 9, 10, 11, 12

RAY_DTYPE = np.dtype([("origin", np.float32, 4), ("dir", np.float32, 4)])
PATH_DTYPE = np.dtype([("throughput", np.float32, 4), ("pixel_index", np.uint32), ("flags", np.uint32), ("pad", np.uint32, 2)])
INTERSECTION_DTYPE = np.dtype([("wuvt", np.float32, 4), ("mesh_instance", np.uint32), ("tri_index", np.uint32), ("pad", np.uint32, 2)])
BXDF_IN_DTYPE = np.dtype([("normal", np.float32, 3), ("mat_node", np.uint32), ("in_dir", np.float32, 3), ("p0", np.float32),
                          ("out_dir", np.float32, 3), ("p1", np.float32), ("rnd", np.float32, 2), ("uv", np.float32, 2)])
BXDF_OUT_DTYPE = np.dtype([("sample", np.float32, 3), ("sample_pdf", np.float32), ("dir", np.float32, 3), ("pdf", np.float32),
                           ("eval", np.float32, 3), ("p", np.float32)])
assert RAY_DTYPE.itemsize == 32 and PATH_DTYPE.itemsize == 32 and INTERSECTION_DTYPE.itemsize == 32
assert BXDF_IN_DTYPE.itemsize == 64 and BXDF_OUT_DTYPE.itemsize == 48

# every symbol include/polaris_cuda.h declares: (name, restype, argtypes)
_P = ctypes.POINTER
SYMBOLS = [
    ("pc_abi_version", ctypes.c_int, []),
    ("pc_device_count", ctypes.c_int, []),
    ("pc_device_info", ctypes.c_int, [ctypes.c_int, ctypes.c_char_p, ctypes.c_size_t, _P(u32), _P(u32), _P(u32)]),
    ("pc_create", ctypes.c_int, [ctypes.c_int, ctypes.c_char_p, _P(vp)]),
    ("pc_destroy", None, [vp]),
    ("pc_id", ctypes.c_char_p, [vp]),
    ("pc_flags", u32, [vp]),
    ("pc_speed", u32, [vp]),
    ("pc_last_error", ctypes.c_char_p, [vp]),
    ("pc_resize", ctypes.c_int, [vp, u32, u32]),
    ("pc_upload_scene", ctypes.c_int, [vp, _P(SceneView)]),
    ("pc_set_camera", ctypes.c_int, [vp, _P(f32), _P(f32)]),
    ("pc_set_option", ctypes.c_int, [vp, ctypes.c_int, ctypes.c_int]),
    ("pc_trace", ctypes.c_int, [vp, _P(BlockRequest), vp, ctypes.c_size_t, _P(Stats)]),
    ("pc_get_stats", ctypes.c_int, [vp, _P(Stats)]),
    ("pc_get_kernel_timings", ctypes.c_int, [vp, vp, vp, u32, _P(u32)]),
    ("pc_merge_output", ctypes.c_int, [vp, vp, _P(BlockRequest)]),
    ("pc_merge_rows", ctypes.c_int, [vp, vp, ctypes.c_int, _P(BlockRequest)]),
    ("pc_trace_rows", ctypes.c_int, [vp, _P(BlockRequest), _P(vp), _P(u64)]),
    ("pc_ipc_export", ctypes.c_int, [vp, ctypes.c_int, vp]),
    ("pc_ipc_publish_rows", ctypes.c_int, [vp, _P(BlockRequest), ctypes.c_int]),
    ("pc_ipc_open", ctypes.c_int, [vp, vp, _P(vp)]),
    ("pc_ipc_close", ctypes.c_int, [vp, vp]),
    ("pc_wait_for_kernels", ctypes.c_int, [vp]),
    ("pc_sync_framebuffer", ctypes.c_int, [vp, _P(BlockRequest), vp]),
    ("pc_host_register", ctypes.c_int, [vp, u64]),
    ("pc_host_unregister", ctypes.c_int, [vp]),
    ("pc_read_buffer", ctypes.c_int, [vp, ctypes.c_int, vp, u64]),
    ("pc_debug_frame_count", u32, [u32, u32]),
    ("pc_trace_debug", ctypes.c_int, [vp, _P(BlockRequest), vp, ctypes.c_size_t, u32, vp, u64, vp, u32, _P(u32), _P(Stats)]),
    ("pc_compile_geometry", vp, [ctypes.c_int, vp, u32, vp, u32, vp, vp, u32, ctypes.c_int32]),
    ("pc_build_bvh", vp, [ctypes.c_int, vp, vp, vp, u32, ctypes.c_int, vp]),
    ("pc_compiled_error", ctypes.c_char_p, [vp]),
    ("pc_compiled_get", ctypes.c_int, [vp, ctypes.c_int, _P(vp), _P(u64)]),
    ("pc_compiled_depths", None, [vp, _P(ctypes.c_int), _P(ctypes.c_int)]),
    ("pc_compiled_timing", None, [vp, vp]),
    ("pc_compiled_free", None, [vp]),
    ("pc_debug_intersect", ctypes.c_int, [vp, vp, u32, ctypes.c_int, vp, vp]),
    ("pc_debug_bxdf", ctypes.c_int, [vp, vp, u32, vp]),
    ("pc_debug_rng", ctypes.c_int, [vp, vp, u32, u32, vp]),
    ("pc_debug_tonemap", ctypes.c_int, [vp, vp, u32, f32, f32, vp]),
]

_lib = None


def load():
    """Load libpolaris_cuda.so and bind every declared symbol; raises if it is missing."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} not built; run `python -c 'import __graft_entry__ as g; g.build()'`. "
                               "There is no CPU fallback.")
        lib = ctypes.CDLL(LIB_PATH)
        for name, res, args in SYMBOLS:
            fn = getattr(lib, name)  # AttributeError if the library does not export it
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def pin_scene(scene) -> list:
    """Page-lock the scene's ten flat arrays in place (pc_host_register); returns the arrays to hand to unpin()."""
    lib = load()
    pinned = []
    for field in scene._SECTIONS:
        a = getattr(scene, field)
        if not a.flags["C_CONTIGUOUS"]:
            a = np.ascontiguousarray(a)
            setattr(scene, field, a)
        if a.nbytes and lib.pc_host_register(a.ctypes.data, a.nbytes) == 0:
            pinned.append(a)
    return pinned


def unpin(arrays):
    lib = load()
    for a in arrays:
        lib.pc_host_unregister(a.ctypes.data)


def scene_view(scene) -> tuple[SceneView, list]:
    """Build a pc_scene_view over a polaris_b200.scene.Scene; returns (view, keepalive)."""
    keep = []

    def pb(a):
        a = np.ascontiguousarray(a)
        keep.append(a)
        return (a.ctypes.data if a.nbytes else None), a.nbytes

    v = SceneView()
    for field, arr in (("bvh_nodes", scene.bvh_nodes), ("mesh_instances", scene.mesh_instances),
                       ("material_nodes", scene.material_nodes), ("texture_data", scene.texture_data),
                       ("texture_metadata", scene.texture_metadata), ("vertices", scene.vertices),
                       ("normals", scene.normals), ("uvs", scene.uvs), ("material_indices", scene.material_index),
                       ("emissives", scene.emissives)):
        p, n = pb(arr)
        setattr(v, field, p)
        setattr(v, field + "_bytes", n)
    v.scene_diffuse_mat_index = scene.scene_diffuse_mat_index
    v.scene_emissive_mat_index = scene.scene_emissive_mat_index
    return v, keep
