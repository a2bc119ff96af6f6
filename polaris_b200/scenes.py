"""Procedural scenes for the five BASELINE.json configs (SURVEY §8(d) "synthetic inputs").

The reference ships no scenes (they live in another repository) so every config is
generated here as a `RawScene` -- the equivalent of what the wavefront reader would hand
to the compiler -- and compiled with `compile_scene`.  Rules taken from the survey:

  * every instance is translation-only, because the reader derives instance boxes from the
    translation alone (asset/scene/reader/wavefront.go:511-517);
  * geometry that is shaded through instances at identity, because `surfaceInit` uses the
    mesh-space hit point as the world point (CL/util/surface.cl:12-33);
  * emissive meshes stay at identity;
  * `scene_emissive_material` comes first in the material list so its root node is node 0
    and compiler.go:214's index quirk cannot bite;
  * coordinates stay within +-100 units;
  * camera FOV 45.0 is the reader's default (raw_scene.go:154) and is consumed as radians.

`sizes` lets tests shrink a config without changing its character.
"""
from __future__ import annotations

import os

import numpy as np

from .material import TEX_LUMINANCE32F, TEX_RGBA8, TEX_RGBA32F
from .scene import RawInstance, RawMesh, RawScene, compile_scene, flat_normals

F = np.float32

# name, frame_w, frame_h, spp as quoted in BASELINE.json["configs"]
CONFIGS = {
    "c1_sphere": (512, 512, 128),
    "c2_cornell": (1024, 1024, 256),
    "c3_instancing": (1920, 1080, 64),
    "c4_terrain": (3840, 2160, 64),
    "c5_cornell_4k": (3840, 2160, 1024),
}
NUM_BOUNCES = 5          # main.go:92-103 defaults
MIN_BOUNCES_FOR_RR = 3
EXPOSURE = 1.2


# ------------------------------------------------------------------------------ helpers
def _quad(p0, p1, p2, p3, normal=None, uv=((0, 0), (1, 0), (1, 1), (0, 1))):
    """Two triangles (0,1,2) (0,2,3), the reader's quad split (wavefront.go:615-618)."""
    p = [np.asarray(x, dtype=F) for x in (p0, p1, p2, p3)]
    v = np.array([[p[0], p[1], p[2]], [p[0], p[2], p[3]]], dtype=F)
    t = np.array([[uv[0], uv[1], uv[2]], [uv[0], uv[2], uv[3]]], dtype=F)
    if normal is None:
        n = flat_normals(v)
    else:
        n = np.broadcast_to(np.asarray(normal, dtype=F), (2, 3, 3)).copy()
    return v, n, t


def _concat(parts):
    v = np.concatenate([p[0] for p in parts]).astype(F)
    n = np.concatenate([p[1] for p in parts]).astype(F)
    t = np.concatenate([p[2] for p in parts]).astype(F)
    return v, n, t


def _grid_to_tris(P, N, UV, wrap_u=False, wrap_v=False):
    """(nu, nv) vertex grids -> triangles; each cell -> (a,b,c) (a,c,d)."""
    nu, nv = P.shape[:2]
    iu = np.arange(nu if wrap_u else nu - 1)
    iv = np.arange(nv if wrap_v else nv - 1)
    I, J = np.meshgrid(iu, iv, indexing="ij")
    I, J = I.ravel(), J.ravel()
    I1, J1 = (I + 1) % nu, (J + 1) % nv

    def gather(A):
        a, b, c, d = A[I, J], A[I1, J], A[I1, J1], A[I, J1]
        t1 = np.stack([a, b, c], axis=1)
        t2 = np.stack([a, c, d], axis=1)
        return np.stack([t1, t2], axis=1).reshape(-1, 3, A.shape[-1])

    return gather(P).astype(F), gather(N).astype(F), gather(UV).astype(F)


def uv_sphere(center, radius, seg_u, seg_v):
    """UV sphere with smooth normals; poles produce degenerate-free triangle fans."""
    u = np.linspace(0.0, 2.0 * np.pi, seg_u + 1)
    v = np.linspace(0.0, np.pi, seg_v + 1)
    U, V = np.meshgrid(u, v, indexing="ij")
    n = np.stack([np.sin(V) * np.cos(U), np.cos(V), np.sin(V) * np.sin(U)], axis=-1)
    P = np.asarray(center, dtype=np.float64) + radius * n
    UV = np.stack([U / (2 * np.pi), V / np.pi], axis=-1)
    pv, pn, pt = _grid_to_tris(P, n, UV)
    # drop the zero-area triangles at the poles
    e1 = pv[:, 1] - pv[:, 0]
    e2 = pv[:, 2] - pv[:, 0]
    area = np.linalg.norm(np.cross(e1, e2), axis=1)
    keep = area > 1e-9
    return pv[keep], pn[keep], pt[keep]


def box(center, size, yaw_deg=0.0, bottom=True):
    cx, cy, cz = center
    hx, hy, hz = size[0] / 2, size[1] / 2, size[2] / 2
    a = np.deg2rad(yaw_deg)
    ca, sa = np.cos(a), np.sin(a)

    def P(x, y, z):
        return (cx + ca * x + sa * z, cy + y, cz - sa * x + ca * z)

    c = {k: P(x, y, z) for k, (x, y, z) in {
        "lbf": (-hx, -hy, hz), "rbf": (hx, -hy, hz), "rtf": (hx, hy, hz), "ltf": (-hx, hy, hz),
        "lbb": (-hx, -hy, -hz), "rbb": (hx, -hy, -hz), "rtb": (hx, hy, -hz), "ltb": (-hx, hy, -hz)}.items()}
    faces = [
        _quad(c["lbf"], c["rbf"], c["rtf"], c["ltf"]),   # front  (+z)
        _quad(c["rbb"], c["lbb"], c["ltb"], c["rtb"]),   # back   (-z)
        _quad(c["rbf"], c["rbb"], c["rtb"], c["rtf"]),   # right  (+x)
        _quad(c["lbb"], c["lbf"], c["ltf"], c["ltb"]),   # left   (-x)
        _quad(c["ltf"], c["rtf"], c["rtb"], c["ltb"]),   # top    (+y)
    ]
    if bottom:
        faces.append(_quad(c["lbb"], c["rbb"], c["rbf"], c["lbf"]))  # bottom (-y)
    return _concat(faces)


def _hash01(ix, iy, seed):
    """Integer lattice hash -> [0,1), deterministic across numpy versions."""
    h = (ix.astype(np.uint64) * np.uint64(0x9E3779B1) + iy.astype(np.uint64) * np.uint64(0x85EBCA77)
         + np.uint64(seed) * np.uint64(0xC2B2AE3D)) & np.uint64(0xFFFFFFFF)
    h ^= h >> np.uint64(15)
    h = (h * np.uint64(0x2C1B3C6D)) & np.uint64(0xFFFFFFFF)
    h ^= h >> np.uint64(12)
    h = (h * np.uint64(0x297A2D39)) & np.uint64(0xFFFFFFFF)
    h ^= h >> np.uint64(15)
    return h.astype(np.float64) / 4294967296.0


def value_noise(x, y, seed, period=None):
    """Smooth value noise; `period` makes it tile on an integer lattice."""
    x0, y0 = np.floor(x), np.floor(y)
    fx, fy = x - x0, y - y0
    sx, sy = fx * fx * (3 - 2 * fx), fy * fy * (3 - 2 * fy)
    ix, iy = x0.astype(np.int64), y0.astype(np.int64)

    def h(a, b):
        if period is not None:
            a, b = a % period, b % period
        return _hash01(a & 0xFFFFFFFF, b & 0xFFFFFFFF, seed)

    v00, v10, v01, v11 = h(ix, iy), h(ix + 1, iy), h(ix, iy + 1), h(ix + 1, iy + 1)
    return (v00 * (1 - sx) + v10 * sx) * (1 - sy) + (v01 * (1 - sx) + v11 * sx) * sy


def fbm(x, y, octaves, seed, period=None):
    out, amp, freq, norm = 0.0, 1.0, 1.0, 0.0
    for o in range(octaves):
        out = out + amp * value_noise(x * freq, y * freq, seed + o, None if period is None else int(period * freq))
        norm += amp
        amp *= 0.5
        freq *= 2.0
    return out / norm


# ------------------------------------------------------------------------------ C1
def c1_sphere(seg_u=40, seg_v=19):
    """C1: diffuse UV sphere on a ground quad under a 1x1 emissive quad."""
    sphere = uv_sphere((0, 1, 0), 1.0, seg_u, seg_v)
    ground = _quad((-8, 0, 8), (8, 0, 8), (8, 0, -8), (-8, 0, -8), normal=(0, 1, 0))
    light = _quad((-0.5, 3, -0.5), (0.5, 3, -0.5), (0.5, 3, 0.5), (-0.5, 3, 0.5), normal=(0, -1, 0))
    materials = {
        "sphere": "diffuse(reflectance: {0.7, 0.7, 0.7})",
        "ground": "diffuse(reflectance: {0.5, 0.5, 0.5})",
        "light": "emissive(radiance: {1, 1, 1}, scale: 15)",
    }
    meshes = [
        RawMesh("sphere", *sphere, np.zeros(len(sphere[0]), np.int32)),
        RawMesh("ground", *ground, np.full(2, 1, np.int32)),
        RawMesh("light", *light, np.full(2, 2, np.int32)),
    ]
    return RawScene(meshes, [RawInstance(i) for i in range(3)], materials,
                    camera_eye=(0, 1, 4), camera_look=(0, 0.5, 0))


def write_obj(raw: RawScene, obj_path, mtl_path):
    """The config names a '.obj/.mtl'; emit them as text the reference's reader accepts
    (`mat_expr`, `instance` extensions, docs/scene.md) -- informational, nothing reads it back."""
    names = list(raw.materials)
    with open(mtl_path, "w") as f:
        for n in names:
            f.write(f"newmtl {n}\nmat_expr {raw.materials[n]}\n\n")
    with open(obj_path, "w") as f:
        f.write(f"mtllib {mtl_path.split('/')[-1]}\n")
        f.write("camera_eye %g %g %g\ncamera_look %g %g %g\n" % (*raw.camera_eye, *raw.camera_look))
        base = 1
        for m in raw.meshes:
            f.write(f"o {m.name}\n")
            for tri in range(m.vertices.shape[0]):
                for k in range(3):
                    f.write("v %.7g %.7g %.7g\n" % tuple(m.vertices[tri, k]))
                    f.write("vn %.7g %.7g %.7g\n" % tuple(m.normals[tri, k]))
                    f.write("vt %.7g %.7g\n" % tuple(m.uvs[tri, k]))
            mat = -1
            for tri in range(m.vertices.shape[0]):
                if m.material[tri] != mat:
                    mat = int(m.material[tri])
                    f.write(f"usemtl {names[mat]}\n")
                i = base + 3 * tri
                f.write(f"f {i}/{i}/{i} {i+1}/{i+1}/{i+1} {i+2}/{i+2}/{i+2}\n")
            base += 3 * m.vertices.shape[0]
        for inst in raw.instances:
            f.write("instance %s %g %g %g 0 0 0 1 1 1\n" % (raw.meshes[inst.mesh_index].name, *inst.translation))


# ------------------------------------------------------------------------------ C2 / C5
def c2_cornell(sphere_seg=(64, 32)):
    """C2: Cornell box with layered diffuse / conductor / dielectric materials (~8k triangles)."""
    W = _concat([
        _quad((-1, 0, 1), (1, 0, 1), (1, 0, -1), (-1, 0, -1), normal=(0, 1, 0)),      # floor
        _quad((-1, 2, -1), (1, 2, -1), (1, 2, 1), (-1, 2, 1), normal=(0, -1, 0)),     # ceiling
        _quad((-1, 0, -1), (1, 0, -1), (1, 2, -1), (-1, 2, -1), normal=(0, 0, 1)),    # back
    ])
    left = _quad((-1, 0, 1), (-1, 0, -1), (-1, 2, -1), (-1, 2, 1), normal=(1, 0, 0))
    right = _quad((1, 0, -1), (1, 0, 1), (1, 2, 1), (1, 2, -1), normal=(-1, 0, 0))
    room = _concat([W, left, right])
    room_mat = np.array([0] * 6 + [1] * 2 + [2] * 2, dtype=np.int32)
    light = _quad((-0.3, 1.998, -0.3), (0.3, 1.998, -0.3), (0.3, 1.998, 0.3), (-0.3, 1.998, 0.3), normal=(0, -1, 0))
    tall = box((-0.38, 0.6, -0.35), (0.6, 1.2, 0.6), yaw_deg=17.0, bottom=False)
    short = box((0.42, 0.277, 0.2), (0.55, 0.55, 0.55), yaw_deg=-17.0, bottom=True)
    s1 = uv_sphere((-0.42, 0.252, 0.52), 0.25, *sphere_seg)
    s2 = uv_sphere((0.42, 0.554 + 0.252, 0.2), 0.25, *sphere_seg)
    materials = {
        "white": "diffuse(reflectance: {0.73, 0.73, 0.73})",
        "red": "diffuse(reflectance: {0.65, 0.05, 0.05})",
        "green": "diffuse(reflectance: {0.12, 0.45, 0.15})",
        "light": "emissive(radiance: {1, 1, 1}, scale: 17)",
        "gold": 'roughConductor(intIOR: "gold", specularity: {1, 0.766, 0.336}, roughness: 0.25)',
        "glass": 'dielectric(intIOR: "glass")',
        "silvered": 'mix(diffuse(reflectance: {0.8, 0.8, 0.8}), conductor(intIOR: "silver", specularity: {0.97, 0.96, 0.92}), 0.6)',
        "frosted": "roughDielectric(roughness: 0.2)",
    }
    meshes = [
        RawMesh("room", *room, room_mat),
        RawMesh("light", *light, np.full(2, 3, np.int32)),
        RawMesh("tall_box", *tall, np.full(len(tall[0]), 4, np.int32)),
        RawMesh("short_box", *short, np.full(len(short[0]), 5, np.int32)),
        RawMesh("sphere_mix", *s1, np.full(len(s1[0]), 6, np.int32)),
        RawMesh("sphere_frosted", *s2, np.full(len(s2[0]), 7, np.int32)),
    ]
    return RawScene(meshes, [RawInstance(i) for i in range(len(meshes))], materials,
                    camera_eye=(0, 1, 3.2), camera_look=(0, 1, 0))


# ------------------------------------------------------------------------------ C3
def _torus(nu, nv, R=1.0, r=0.4, amp=0.05, seed=1):
    u = np.arange(nu) * (2 * np.pi / nu)
    v = np.arange(nv) * (2 * np.pi / nv)
    U, V = np.meshgrid(u, v, indexing="ij")
    disp = (fbm(U / (2 * np.pi) * 16, V / (2 * np.pi) * 16, 4, seed, period=16) - 0.5) * 2 * amp * r
    rr = r + disp
    n = np.stack([np.cos(V) * np.cos(U), np.sin(V), np.cos(V) * np.sin(U)], axis=-1)
    c = np.stack([R * np.cos(U), np.zeros_like(U), R * np.sin(U)], axis=-1)
    P = c + rr[..., None] * n
    UV = np.stack([U / (2 * np.pi), V / (2 * np.pi)], axis=-1)
    return _grid_to_tris(P, n, UV, wrap_u=True, wrap_v=True)


def c3_instancing(grid=224, lattice=10):
    """C3: lattice^3 instances of a displaced torus (grid^2 * 2 triangles each), two source
    meshes (diffuse / roughConductor), emissive quad above, sky background."""
    t = _torus(grid, grid)
    nt = len(t[0])
    materials = {
        "matte": "diffuse(reflectance: {0.7, 0.55, 0.4})",
        "metal": 'roughConductor(intIOR: "copper", specularity: {0.95, 0.64, 0.54}, roughness: 0.3)',
        "light": "emissive(radiance: {1, 0.95, 0.9}, scale: 12)",
        "scene_diffuse_material": "diffuse(reflectance: {0.35, 0.45, 0.6})",
    }
    pitch = 3.0
    span = pitch * (lattice - 1)
    ext = span / 2 + 4
    top = span / 2 + 3
    light = _quad((-ext, top, -ext), (ext, top, -ext), (ext, top, ext), (-ext, top, ext), normal=(0, -1, 0))
    meshes = [
        RawMesh("torus_matte", *t, np.zeros(nt, np.int32)),
        RawMesh("torus_metal", *t, np.ones(nt, np.int32)),
        RawMesh("light", *light, np.full(2, 2, np.int32)),
    ]
    insts = []
    for i in range(lattice):
        for j in range(lattice):
            for k in range(lattice):
                pos = (i * pitch - span / 2, j * pitch - span / 2, k * pitch - span / 2)
                insts.append(RawInstance((i + j + k) & 1, pos))
    insts.append(RawInstance(2))
    d = span / 2 + 9
    return RawScene(meshes, insts, materials, camera_eye=(d * 0.55, d * 0.35, d), camera_look=(0, 0, 0))


# ------------------------------------------------------------------------------ C4
def c4_terrain(n=2237, tex=2048, sky=(4096, 2048), seed=1):
    """C4: n x n height field ((n-1)^2*2 triangles; 9,999,392 at n=2237) of dispersive rough
    glass with bilinear L32F / RGBA8 / RGBA32F textures and a lat-long HDR sky."""
    size = 100.0
    xs = np.linspace(-size / 2, size / 2, n)
    X, Z = np.meshgrid(xs, xs, indexing="ij")
    H = (fbm(X / size * 8 + 3.1, Z / size * 8 + 1.7, 6, seed) - 0.5) * 12.0
    P = np.stack([X, H, Z], axis=-1)
    gx, gz = np.gradient(H, xs, xs)
    nrm = np.stack([-gx, np.ones_like(H), -gz], axis=-1)
    nrm /= np.linalg.norm(nrm, axis=-1, keepdims=True)
    UV = np.stack([(X / size + 0.5) * 8, (Z / size + 0.5) * 8], axis=-1)
    # triangle winding chosen so the geometric normal points up
    pv, pn, pt = _grid_to_tris(P[:, ::-1], nrm[:, ::-1], UV[:, ::-1])
    ty, tx = np.meshgrid(np.arange(tex), np.arange(tex), indexing="ij")
    rough = (0.1 + 0.5 * fbm(tx / tex * 8, ty / tex * 8, 4, seed + 11, period=8)).astype(F)
    tr = np.stack([
        200 + 55 * fbm(tx / tex * 4, ty / tex * 4, 3, seed + 21, period=4),
        215 + 40 * fbm(tx / tex * 4, ty / tex * 4, 3, seed + 22, period=4),
        225 + 30 * fbm(tx / tex * 4, ty / tex * 4, 3, seed + 23, period=4),
        np.full((tex, tex), 255.0)], axis=-1).astype(np.uint8)
    sw, sh = sky
    sy, sx = np.meshgrid(np.arange(sh), np.arange(sw), indexing="ij")
    lum = np.exp(np.log(0.05) + (np.log(50.0) - np.log(0.05)) * fbm(sx / sw * 8, sy / sh * 4, 5, seed + 31, period=8) ** 2.5)
    lum = lum * np.clip(1.3 - sy / sh * 1.2, 0.05, None)  # darker towards the nadir
    skyimg = np.stack([lum * 0.9, lum * 1.0, lum * 1.15, np.ones_like(lum)], axis=-1).astype(F)
    textures = {
        "sky.exr": (TEX_RGBA32F, sw, sh, skyimg.tobytes()),
        "r.tex": (TEX_LUMINANCE32F, tex, tex, rough.tobytes()),
        "t.tex": (TEX_RGBA8, tex, tex, tr.tobytes()),
    }
    materials = {
        "scene_emissive_material": 'emissive(radiance: "sky.exr", scale: 1)',
        "scene_diffuse_material": 'diffuse(reflectance: "sky.exr")',
        "glass_terrain": 'disperse(roughDielectric(intIOR: 1.5, roughness: "r.tex", transmittance: "t.tex"), '
                         "intIOR: {1.50, 1.52, 1.54}, extIOR: {0, 0, 0})",
    }
    meshes = [RawMesh("terrain", pv, pn, pt, np.full(len(pv), 2, np.int32))]
    return RawScene(meshes, [RawInstance(0)], materials, textures,
                    camera_eye=(0, 14, 46), camera_look=(0, 0, 0))


# ------------------------------------------------------------------------------ front door
_GENERATORS = {"c1_sphere": "c1_sphere", "c2_cornell": "c2_cornell", "c3_instancing": "c3_instancing",
               "c4_terrain": "c4_terrain", "c5_cornell_4k": "c2_cornell"}
_raw_cache = {}


def raw_scene(name: str, **sizes):
    """The procedural RawScene of a config (what the wavefront reader would hand the compiler), memoised: the 10 M-triangle
    terrain takes most of a minute to generate and several callers (host build, device build, tests) want the same one."""
    key = (name, tuple(sorted((k, tuple(v) if isinstance(v, (list, tuple)) else v) for k, v in sizes.items())))
    if key not in _raw_cache:
        if len(_raw_cache) >= 2:
            _raw_cache.pop(next(iter(_raw_cache)))
        _raw_cache[key] = globals()[_GENERATORS[name]](**sizes)
    return _raw_cache[key]


def build(name: str, frame_w=None, frame_h=None, builder: str = "host", **sizes):
    """Compile one config; returns (Scene, frame_w, frame_h, spp).  builder: "host" (libpolaris_scene.so) or "cuda" (the
    device BVH build of libpolaris_cuda.so) -- same bytes either way."""
    w, h, spp = CONFIGS[name]
    w, h = frame_w or w, frame_h or h
    # POLARIS_SCENE_CACHE=<dir>: keep the compiled scene as a PLRSCN2 dump (the 10 M-triangle terrain takes a minute of
    # procedural generation + compilation; several bench / profiler runs inside one session share it)
    cache = os.environ.get("POLARIS_SCENE_CACHE")
    path = os.path.join(cache, f"{name}_{w}x{h}.plrscn") if cache and not sizes else None
    if path and os.path.exists(path):
        from .scene import Scene
        sc = Scene.load(path)
        sc.camera.setup_projection(F(w) / F(h))
        return sc, w, h, spp
    sc = compile_scene(raw_scene(name, **sizes), aspect=F(w) / F(h), builder=builder)
    if path:
        os.makedirs(cache, exist_ok=True)
        sc.save(path + ".tmp")
        os.replace(path + ".tmp", path)
    return sc, w, h, spp
