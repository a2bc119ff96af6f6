"""Compiled scenes: the ten flat buffers of `scene.Scene` plus camera.

`RawScene` plays the role of the reference's `input.Scene` (asset/compiler/input/raw_scene.go):
triangle soups, mesh instances, named material expressions, camera.  `compile_scene` is the
reference's `compiler.Compile` (asset/compiler/compiler.go:45-76): materials are flattened
here (polaris_b200/material.py), geometry is partitioned by the C++ builder in
csrc/scene_compiler.cpp, the camera follows asset/scene/camera.go.

`Scene.save` / `Scene.load` implement the raw little-endian interchange dump SURVEY §8(f)
asks for (the reference's own archive is Go gob inside a zip, asset/scene/writer/zip.go:47-52,
which only Go can read): magic, version, 10 (length, bytes) sections in upload order, the
two scene material indices and the camera.
"""
from __future__ import annotations

import ctypes
import os
import struct
from dataclasses import dataclass, field

import numpy as np

from . import gotypes as gt
from .material import MATERIAL_NODE_DTYPE, MaterialCompiler

F = np.float32

BVH_NODE_DTYPE = np.dtype([("min", F, 3), ("ldata", np.int32), ("max", F, 3), ("rdata", np.int32)])
MESH_INSTANCE_DTYPE = np.dtype([("mesh_index", np.uint32), ("bvh_root", np.uint32), ("pad", np.uint32, 2), ("transform", F, 16)])
EMISSIVE_DTYPE = np.dtype([("transform", F, 16), ("area", F), ("prim_index", np.uint32), ("mat_node_index", np.uint32), ("type", np.uint32)])
TEXTURE_META_DTYPE = np.dtype([("format", np.uint32), ("width", np.uint32), ("height", np.uint32), ("data_offset", np.uint32)])
assert BVH_NODE_DTYPE.itemsize == 32 and MESH_INSTANCE_DTYPE.itemsize == 80
assert EMISSIVE_DTYPE.itemsize == 80 and TEXTURE_META_DTYPE.itemsize == 16

SCENE_DIFFUSE_MATERIAL = "scene_diffuse_material"    # compiler.go:20
SCENE_EMISSIVE_MATERIAL = "scene_emissive_material"  # compiler.go:21

_HERE = os.path.dirname(os.path.abspath(__file__))


# --------------------------------------------------------------------------- camera
@dataclass
class Camera:
    """asset/scene/camera.go:38-141. FOV is used as radians by Perspective4 (matrix.go:156-161)."""

    position: np.ndarray
    look_at: np.ndarray
    up: np.ndarray
    fov: float = 45.0
    pitch: float = 0.0
    yaw: float = 0.0
    invert_y: bool = False
    view_mat: np.ndarray = field(default_factory=gt.ident4)
    proj_mat: np.ndarray = field(default_factory=gt.ident4)
    frustrum: np.ndarray = field(default_factory=lambda: np.zeros((4, 4), dtype=F))
    look_at_given: np.ndarray | None = None  # LookAt as constructed: Update() replaces look_at by position + unit direction

    def __post_init__(self):
        if self.look_at_given is None:
            self.look_at_given = np.array(self.look_at, dtype=F)

    def setup_projection(self, aspect):
        self.proj_mat = gt.perspective4(self.fov, aspect, 1, 1000)
        self.update()

    def update(self):
        pos = np.asarray(self.position, dtype=F)
        # a LookAt (or position) somebody set since the last update is the new "given" one; our own write-back below is not
        own = getattr(self, "_own", None)
        if own is not None and not (np.array_equal(self.look_at, own[0]) and np.array_equal(pos, own[1])):
            self.look_at_given = np.array(self.look_at, dtype=F)
        d = gt.v_normalize(np.asarray(self.look_at, dtype=F) - pos)
        pitch_axis = gt.v_cross(d, self.up)
        pq = gt.quat_from_axis_angle(pitch_axis, self.pitch)
        yq = gt.quat_from_axis_angle(np.asarray(self.up, dtype=F), self.yaw)
        oq = gt.quat_normalize(gt.quat_mul(pq, yq))
        d = gt.quat_rotate(oq, d)
        self.look_at = (pos + np.array([F(c * F(1.0)) for c in d], dtype=F)).astype(F)
        if self.pitch != 0.0 or self.yaw != 0.0:  # rotated away from the given LookAt: only the written-back one describes the view
            self.look_at_given = self.look_at.copy()
        self._own = (self.look_at.copy(), pos.copy())
        self.view_mat = gt.look_at_v(pos, self.look_at, self.up)
        inv = gt.inv4(gt.mul4(self.proj_mat, self.view_mat))
        y_up = F(-1.0) if self.invert_y else F(1.0)
        corners = [(-1, y_up, -1, 1), (1, y_up, -1, 1), (-1, -y_up, -1, 1), (1, -y_up, -1, 1)]
        fr = np.zeros((4, 4), dtype=F)
        for k, c in enumerate(corners):
            v = gt.mul4x1(inv, np.array(c, dtype=F))
            s = F(F(1.0) / v[3])
            fr[k, :3] = np.array([F(v[i] * s) for i in range(3)], dtype=F) - pos
        self.frustrum = fr


# --------------------------------------------------------------------------- raw scene
@dataclass
class RawMesh:
    name: str
    vertices: np.ndarray   # (T, 3, 3) float32
    normals: np.ndarray    # (T, 3, 3)
    uvs: np.ndarray        # (T, 3, 2)
    material: np.ndarray   # (T,) int32 index into RawScene.materials order

    def bbox(self):
        v = np.ascontiguousarray(self.vertices, dtype=F).reshape(-1, 3)
        # min / max over axis 0 of an (N, 3) array crawls (inner dimension 3): fold long rows first
        k = 4096
        n = (v.shape[0] // k) * k
        parts_lo, parts_hi = [], []
        if n:
            w = v[:n].reshape(-1, 3 * k)
            parts_lo.append(w.min(axis=0).reshape(k, 3))
            parts_hi.append(w.max(axis=0).reshape(k, 3))
        if n < v.shape[0]:
            parts_lo.append(v[n:])
            parts_hi.append(v[n:])
        return np.concatenate(parts_lo).min(axis=0).astype(F), np.concatenate(parts_hi).max(axis=0).astype(F)


@dataclass
class RawInstance:
    mesh_index: int
    translation: tuple = (0.0, 0.0, 0.0)  # translation-only, see SURVEY §8(d)


@dataclass
class RawScene:
    meshes: list
    instances: list
    materials: dict               # ordered: name -> expression
    textures: dict = field(default_factory=dict)  # name -> (format, w, h, bytes)
    camera_eye: tuple = (0, 0, 0)
    camera_look: tuple = (0, 0, -1)
    camera_up: tuple = (0, 1, 0)
    camera_fov: float = 45.0


def flat_normals(vertices: np.ndarray) -> np.ndarray:
    """wavefront.go:603-611: face normal e01 x e02, normalised with Vec3.Normalize."""
    e01 = (vertices[:, 1] - vertices[:, 0]).astype(F)
    e02 = (vertices[:, 2] - vertices[:, 0]).astype(F)
    c = np.empty_like(e01)
    c[:, 0] = (e01[:, 1] * e02[:, 2]).astype(F) - (e01[:, 2] * e02[:, 1]).astype(F)
    c[:, 1] = (e01[:, 2] * e02[:, 0]).astype(F) - (e01[:, 0] * e02[:, 2]).astype(F)
    c[:, 2] = (e01[:, 0] * e02[:, 1]).astype(F) - (e01[:, 1] * e02[:, 0]).astype(F)
    l2 = ((c[:, 0] * c[:, 0]).astype(F) + (c[:, 1] * c[:, 1]).astype(F)).astype(F)
    l2 = (l2 + (c[:, 2] * c[:, 2]).astype(F)).astype(F)
    ln = np.sqrt(l2.astype(np.float64)).astype(F)
    with np.errstate(divide="ignore"):
        inv = (F(1.0) / ln).astype(F)
    n = (c * inv[:, None]).astype(F)
    n[inv < 1e-10] = 0
    return np.repeat(n[:, None, :], 3, axis=1)


# --------------------------------------------------------------------------- compiled scene
@dataclass
class Scene:
    bvh_nodes: np.ndarray
    mesh_instances: np.ndarray
    material_nodes: np.ndarray
    emissives: np.ndarray
    texture_data: np.ndarray       # uint8
    texture_metadata: np.ndarray
    vertices: np.ndarray           # (V, 4) float32
    normals: np.ndarray            # (V, 4)
    uvs: np.ndarray                # (V, 2)
    material_index: np.ndarray     # (T,) uint32
    scene_diffuse_mat_index: int = -1
    scene_emissive_mat_index: int = -1
    camera: Camera | None = None
    top_depth: int = 0
    mesh_depth: int = 0
    compile_timing: dict | None = None  # seconds per phase of the native geometry compiler (compile_scene)

    _SECTIONS = ("bvh_nodes", "mesh_instances", "material_nodes", "texture_data", "texture_metadata",
                 "vertices", "normals", "uvs", "material_index", "emissives")
    _DTYPES = (BVH_NODE_DTYPE, MESH_INSTANCE_DTYPE, MATERIAL_NODE_DTYPE, np.uint8, TEXTURE_META_DTYPE,
               F, F, F, np.uint32, EMISSIVE_DTYPE)
    _SHAPES = (None, None, None, None, None, (-1, 4), (-1, 4), (-1, 2), None, None)
    MAGIC = b"PLRSCN2\0"

    @property
    def num_triangles(self):
        return int(self.material_index.shape[0])

    def nbytes(self):
        return sum(getattr(self, s).nbytes for s in self._SECTIONS)

    def save(self, path):
        with open(path, "wb") as f:
            f.write(self.MAGIC)
            f.write(struct.pack("<I", 1))
            for s in self._SECTIONS:
                b = np.ascontiguousarray(getattr(self, s)).tobytes()
                f.write(struct.pack("<Q", len(b)))
                f.write(b)
            f.write(struct.pack("<ii", self.scene_diffuse_mat_index, self.scene_emissive_mat_index))
            cam = self.camera
            f.write(np.asarray(cam.position, dtype=F).tobytes())
            # LookAt as it was BEFORE Update() (camera.go: Update stores Position + normalised direction back into LookAt, and
            # normalising that again differs in the last ulp): a loaded scene then derives the very same frustum
            f.write(np.asarray(cam.look_at_given, dtype=F).tobytes())
            f.write(np.asarray(cam.up, dtype=F).tobytes())
            f.write(struct.pack("<f", cam.fov))

    @classmethod
    def load(cls, path):
        with open(path, "rb") as f:
            if f.read(8) != cls.MAGIC:
                raise ValueError("not a polaris raw scene dump")
            (ver,) = struct.unpack("<I", f.read(4))
            if ver != 1:
                raise ValueError(f"unsupported scene dump version {ver}")
            arrays = {}
            for s, dt, shp in zip(cls._SECTIONS, cls._DTYPES, cls._SHAPES):
                (n,) = struct.unpack("<Q", f.read(8))
                a = np.frombuffer(f.read(n), dtype=dt).copy()
                arrays[s] = a.reshape(shp) if shp else a
            d, e = struct.unpack("<ii", f.read(8))
            pos = np.frombuffer(f.read(12), dtype=F).copy()
            look = np.frombuffer(f.read(12), dtype=F).copy()
            up = np.frombuffer(f.read(12), dtype=F).copy()
            (fov,) = struct.unpack("<f", f.read(4))
        return cls(scene_diffuse_mat_index=d, scene_emissive_mat_index=e,
                   camera=Camera(pos, look, up, fov), **arrays)


# --------------------------------------------------------------------------- C++ compiler binding
class _PsMesh(ctypes.Structure):
    _fields_ = [("vertices", ctypes.c_void_p), ("normals", ctypes.c_void_p), ("uvs", ctypes.c_void_p),
                ("material", ctypes.c_void_p), ("ntris", ctypes.c_uint32)]


class _PsInstance(ctypes.Structure):
    _fields_ = [("mesh_index", ctypes.c_uint32), ("inv_transform", ctypes.c_float * 16),
                ("bbox_min", ctypes.c_float * 3), ("bbox_max", ctypes.c_float * 3), ("center", ctypes.c_float * 3)]


_scene_lib = None


def scene_lib():
    global _scene_lib
    if _scene_lib is None:
        path = os.path.join(_HERE, "libpolaris_scene.so")
        if not os.path.exists(path):
            raise RuntimeError(f"{path} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                               "or `make -C polaris_b200/csrc`")
        lib = ctypes.CDLL(path)
        lib.ps_compile.restype = ctypes.c_void_p
        lib.ps_compile.argtypes = [ctypes.c_void_p, ctypes.c_uint32, ctypes.c_void_p, ctypes.c_uint32,
                                   ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint32, ctypes.c_int32]
        lib.ps_build_bvh.restype = ctypes.c_void_p
        lib.ps_build_bvh.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint32,
                                     ctypes.c_int, ctypes.c_void_p]
        lib.ps_error.restype = ctypes.c_char_p
        lib.ps_error.argtypes = [ctypes.c_void_p]
        lib.ps_get.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(ctypes.c_uint64)]
        lib.ps_depths.argtypes = [ctypes.c_void_p, ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_int)]
        lib.ps_free.argtypes = [ctypes.c_void_p]
        lib.ps_timing.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
        _scene_lib = lib
    return _scene_lib


class _GeometryCompiler:
    """The two producers of the geometry buffers behind one face: "host" = libpolaris_scene.so (ps_*, OpenMP), "cuda" =
    libpolaris_cuda.so (pc_compile_geometry / pc_build_bvh: the same compiler with the SAH build on the device,
    csrc/pc_bvh_build.cu).  Both emit byte-identical buffers."""

    def __init__(self, builder="host", ordinal=0):
        self.builder, self.ordinal = builder, ordinal
        if builder == "host":
            lib = scene_lib()
            self._compile, self._build = lib.ps_compile, lib.ps_build_bvh
            self.ps_error, self.ps_get, self.ps_depths, self.ps_free = lib.ps_error, lib.ps_get, lib.ps_depths, lib.ps_free
            self.ps_timing = lib.ps_timing
        elif builder == "cuda":
            from . import _lib as cabi
            lib = cabi.load()
            self._compile = lambda *a: lib.pc_compile_geometry(ordinal, *a)
            self._build = lambda *a: lib.pc_build_bvh(ordinal, *a)
            self.ps_error, self.ps_get, self.ps_depths, self.ps_free = (lib.pc_compiled_error, lib.pc_compiled_get,
                                                                        lib.pc_compiled_depths, lib.pc_compiled_free)
            self.ps_timing = lib.pc_compiled_timing
        else:
            raise ValueError(f"unknown scene builder {builder!r}")

    def ps_compile(self, *a):
        return self._compile(*a)

    def ps_build_bvh(self, *a):
        return self._build(*a)


def _fetch(lib, h, which, dtype, shape=None):
    ptr, n = ctypes.c_void_p(), ctypes.c_uint64()
    lib.ps_get(h, which, ctypes.byref(ptr), ctypes.byref(n))
    if n.value == 0:
        a = np.zeros(0, dtype=dtype)
    else:
        a = np.ctypeslib.as_array((ctypes.c_ubyte * n.value).from_address(ptr.value)).view(dtype).copy()  # one copy
    return a.reshape(shape) if shape else a


def build_bvh(bmin, bmax, center, min_leaf_items, builder="host", ordinal=0):
    """bvh.Build on bare volumes; returns (nodes, leaf-ordered item indices)."""
    lib = _GeometryCompiler(builder, ordinal)
    bmin = np.ascontiguousarray(bmin, dtype=F)
    bmax = np.ascontiguousarray(bmax, dtype=F)
    center = np.ascontiguousarray(center, dtype=F)
    n = bmin.shape[0]
    order = np.zeros(n, dtype=np.uint32)
    h = lib.ps_build_bvh(bmin.ctypes.data, bmax.ctypes.data, center.ctypes.data, n, min_leaf_items, order.ctypes.data)
    try:
        err = lib.ps_error(h)
        if err:
            raise RuntimeError(err.decode())
        return _fetch(lib, h, 0, BVH_NODE_DTYPE), order
    finally:
        lib.ps_free(h)


def compile_scene(raw: RawScene, aspect: float | None = None, builder: str = "host", ordinal: int = 0) -> Scene:
    lib = _GeometryCompiler(builder, ordinal)
    # --- createLayeredMaterialTrees (compiler.go:271-310); every listed material is "Used"
    mc = MaterialCompiler(raw.materials, raw.textures)
    names = list(raw.materials.keys())
    mat_root = np.full(len(names), -1, dtype=np.int32)
    mat_emissive = np.full(len(names), -1, dtype=np.int32)
    diffuse_idx = emissive_idx = -1
    for i, name in enumerate(names):
        mat_root[i] = mc.generate(name)
        mat_emissive[i] = mc.find_emissive(int(mat_root[i]))
        if name == SCENE_DIFFUSE_MATERIAL:
            diffuse_idx = int(mat_root[i])
        elif name == SCENE_EMISSIVE_MATERIAL:
            emissive_idx = int(mat_root[i])
    env_node = -1
    if emissive_idx != -1:
        # compiler.go:214 indexes emissiveIndexCache with the *node* index; for the usual
        # single-leaf emissive material root == leaf, so look the leaf up directly.
        env_node = mc.find_emissive(emissive_idx)

    # --- geometry
    keep = []
    meshes = (_PsMesh * len(raw.meshes))()
    for i, m in enumerate(raw.meshes):
        v = np.ascontiguousarray(m.vertices, dtype=F)
        n = np.ascontiguousarray(m.normals, dtype=F)
        u = np.ascontiguousarray(m.uvs, dtype=F)
        mat = np.ascontiguousarray(m.material, dtype=np.int32)
        keep += [v, n, u, mat]
        meshes[i] = _PsMesh(v.ctypes.data, n.ctypes.data, u.ctypes.data, mat.ctypes.data, v.shape[0])
    insts = (_PsInstance * len(raw.instances))()
    mesh_boxes = {}  # one min/max pass per MESH, not per instance (1 000 instances share two meshes in config 3)
    for i, inst in enumerate(raw.instances):
        # wavefront.go:505-523: M = S*(R*T); instance AABB = translated mesh AABB corners
        trans = gt.translate4(inst.translation)
        m = gt.mul4(gt.scale4((0, 0, 0)), gt.mul4(gt.ident4(), trans))
        inv = gt.inv4(m)
        if inst.mesh_index not in mesh_boxes:
            mesh_boxes[inst.mesh_index] = raw.meshes[inst.mesh_index].bbox()
        lo, hi = mesh_boxes[inst.mesh_index]
        a = gt.mul4x1(trans, np.array([*lo, 1], dtype=F))[:3]
        b = gt.mul4x1(trans, np.array([*hi, 1], dtype=F))[:3]
        bmin, bmax = np.minimum(a, b), np.maximum(a, b)
        cen = ((bmin + bmax).astype(F) * F(0.5)).astype(F)
        insts[i].mesh_index = inst.mesh_index
        insts[i].inv_transform[:] = [float(x) for x in inv]
        insts[i].bbox_min[:] = [float(x) for x in bmin]
        insts[i].bbox_max[:] = [float(x) for x in bmax]
        insts[i].center[:] = [float(x) for x in cen]
    h = lib.ps_compile(ctypes.addressof(meshes), len(raw.meshes), ctypes.addressof(insts), len(raw.instances),
                       mat_root.ctypes.data, mat_emissive.ctypes.data, len(names), env_node)
    try:
        err = lib.ps_error(h)
        if err:
            raise RuntimeError(err.decode())
        td, md = ctypes.c_int(), ctypes.c_int()
        lib.ps_depths(h, ctypes.byref(td), ctypes.byref(md))
        tim = (ctypes.c_double * 8)()
        lib.ps_timing(h, tim)
        tex_meta = np.array(mc.tex_meta, dtype=np.uint32).reshape(-1, 4).view(TEXTURE_META_DTYPE).reshape(-1) \
            if mc.tex_meta else np.zeros(0, dtype=TEXTURE_META_DTYPE)
        sc = Scene(
            bvh_nodes=_fetch(lib, h, 0, BVH_NODE_DTYPE),
            mesh_instances=_fetch(lib, h, 1, MESH_INSTANCE_DTYPE),
            material_nodes=mc.node_array(),
            emissives=_fetch(lib, h, 2, EMISSIVE_DTYPE),
            texture_data=np.frombuffer(bytes(mc.tex_data), dtype=np.uint8).copy(),
            texture_metadata=tex_meta,
            vertices=_fetch(lib, h, 3, F, (-1, 4)),
            normals=_fetch(lib, h, 4, F, (-1, 4)),
            uvs=_fetch(lib, h, 5, F, (-1, 2)),
            material_index=_fetch(lib, h, 6, np.uint32),
            scene_diffuse_mat_index=diffuse_idx,
            scene_emissive_mat_index=emissive_idx,
            top_depth=td.value,
            mesh_depth=md.value,
        )
        sc.compile_timing = {"builder": builder, "bounds_s": tim[0], "bvh_build_s": tim[1], "flatten_s": tim[2], "gather_s": tim[3],
                             "native_total_s": tim[4], "bvh_build_device_s": tim[5]}
    finally:
        lib.ps_free(h)
    # --- setupCamera (compiler.go:234-241) + cmd/render.go:58
    cam = Camera(np.array(raw.camera_eye, dtype=F), np.array(raw.camera_look, dtype=F),
                 np.array(raw.camera_up, dtype=F), raw.camera_fov)
    if aspect is not None:
        cam.setup_projection(aspect)
    sc.camera = cam
    return sc
