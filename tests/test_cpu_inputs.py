"""CPU tier: the input producers and host logic, checked against every known answer the reference's own
tests hold for this path -- BVH builder, scene compiler, block schedulers, material type
codes, the fixture cube -- plus the properties the compiled buffers must have for the tracer ABI.
"""
import os
import numpy as np
import pytest

from oracle import binding as OB
from polaris_b200 import material as M
from polaris_b200 import scene as S
from polaris_b200 import scenes
from polaris_b200.scheduler import NaiveScheduler, PerfectScheduler, StaticSpeed

from . import common as C

F = np.float32


# ------------------------------------------------------------------------------------------------
# reference asset/compiler/bvh/bvh_builder_test.go:10-68
FOUR_BOXES = [((-2, 0, -2), (-1, 1, -1)), ((1, 0, -2), (2, 1, -1)), ((-2, 0, 1), (-1, 1, 2)), ((1, 0, 1), (2, 1, 2))]


def _boxes(spec):
    bmin = np.array([b[0] for b in spec], F)
    bmax = np.array([b[1] for b in spec], F)
    return bmin, bmax, ((bmin + bmax) * F(0.5)).astype(F)


@pytest.mark.parametrize("builder", ["binned", "literal"])
def test_bvh_builder_known_answers(builder):
    build = S.build_bvh if builder == "binned" else OB.build_bvh_literal
    bmin, bmax, cen = _boxes(FOUR_BOXES)
    nodes, order = build(bmin, bmax, cen, 1)  # one item per leaf -> 4 leaves, 7 nodes (:45-56)
    assert len(nodes) == 7
    leaves = nodes[nodes["ldata"] <= 0]
    assert len(leaves) == 4
    assert sorted(order.tolist()) == [0, 1, 2, 3]
    nodes, order = build(bmin, bmax, cen, 2)  # two items per leaf -> 2 leaves, 3 nodes (:58-68)
    assert len(nodes) == 3
    assert (nodes["ldata"][1:] <= 0).all() and nodes["ldata"][0] == 1 and nodes["rdata"][0] == 2
    # the root box is the union of the item boxes (bvh_builder.go:134-139)
    assert np.array_equal(nodes["min"][0], bmin.min(axis=0)) and np.array_equal(nodes["max"][0], bmax.max(axis=0))


def test_binned_builder_equals_literal_builder():
    """The production builder evaluates the SAH sweep by binning; it must emit exactly the
    node array of the literal O(planes x items) restatement of bvh_builder.go:124-224."""
    rng = np.random.default_rng(11)
    for n, leaf in ((1, 1), (2, 1), (37, 1), (400, 10), (1500, 10)):
        lo = rng.uniform(-20, 20, size=(n, 3)).astype(F)
        hi = (lo + rng.uniform(0.01, 3, size=(n, 3)).astype(F)).astype(F)
        cen = ((lo + hi) * F(0.5)).astype(F)
        a, ao = S.build_bvh(lo, hi, cen, leaf)
        b, bo = OB.build_bvh_literal(lo, hi, cen, leaf)
        assert a.tobytes() == b.tobytes(), f"n={n} leaf={leaf}: node arrays differ"
        assert np.array_equal(ao, bo)


def _single_triangle_two_instances():
    v = np.array([[[0, 0, 0], [1, 0, 0], [0, 1, 0]]], F)
    mesh = S.RawMesh("tri", v, S.flat_normals(v), np.zeros((1, 3, 2), F), np.zeros(1, np.int32))
    raw = S.RawScene([mesh], [S.RawInstance(0, (0, 0, 0)), S.RawInstance(0, (5, 0, 0))],
                     {"m": "diffuse(reflectance: {0.5, 0.5, 0.5})"})
    return S.compile_scene(raw, aspect=F(1.0))


def test_compiler_known_answers():
    """reference asset/compiler/compiler_test.go:154-193: 1 triangle x 2 instances."""
    sc = _single_triangle_two_instances()
    assert len(sc.vertices) == 3 and len(sc.normals) == 3 and len(sc.uvs) == 3
    assert len(sc.material_index) == 1
    assert len(sc.bvh_nodes) == 4
    assert len(sc.mesh_instances) == 2
    leaf0, leaf1 = sc.bvh_nodes[1], sc.bvh_nodes[2]
    assert leaf0["rdata"] == 0 and -leaf0["ldata"] == 0  # top leaf 0 -> instance 0
    assert leaf1["rdata"] == 0 and -leaf1["ldata"] == 1  # top leaf 1 -> instance 1
    assert sc.mesh_instances["bvh_root"].tolist() == [3, 3]
    mesh_leaf = sc.bvh_nodes[3]
    assert -mesh_leaf["ldata"] == 0 and mesh_leaf["rdata"] == 1  # first triangle 0, 1 triangle
    # instance stores the INVERSE transform (compiler.go:191): translation (5,0,0) -> column 3 = (-5,0,0,1)
    t = sc.mesh_instances["transform"][1].reshape(4, 4)  # column-major: row k of the reshape == column k
    assert np.allclose(t[3], [-5, 0, 0, 1]) and np.allclose(t[:3, :3], np.eye(3))


def test_fixture_cube_scene():
    """tracer/opencl/fixtures/cube.{obj,mtl}: 12 triangles, two instances one unit apart, Kd 0.588."""
    cube = scenes.box((0, 0, 0), (1, 1, 1))
    mesh = S.RawMesh("cube", *cube, np.zeros(len(cube[0]), np.int32))
    raw = S.RawScene([mesh], [S.RawInstance(0, (0, 0, 0)), S.RawInstance(0, (-1.0, 0, 0))],
                     {"cube": "diffuse(reflectance: {0.588, 0.588, 0.588})"}, camera_eye=(0, 0, 4), camera_look=(0, 0, 0))
    sc = S.compile_scene(raw, aspect=F(1.0))
    assert sc.num_triangles == 12 and len(sc.mesh_instances) == 2
    assert len(sc.emissives) == 0
    # leaf order re-orders triangles but keeps the set (compiler.go:132-169)
    got = np.sort(sc.vertices[:, :3].reshape(12, 9), axis=0)
    want = np.sort(cube[0].reshape(12, 9), axis=0)
    assert np.array_equal(got, want)
    assert (sc.vertices[:, 3] == 0).all()
    m = sc.material_nodes[sc.material_index[0]]
    assert m["union1"][0] == M.BXDF_DIFFUSE and np.allclose(m["union2"][:3], 0.588)


def test_align4():
    """compiler_test.go:12-18 TestAlign16 (align4 pads texture blobs to 4 bytes, compiler.go:528-537,556-563)."""
    from polaris_b200.material import align4

    for i in range(1, 17):
        assert align4(i) % 4 == 0 and 0 <= align4(i) - i < 4
    assert [align4(x) for x in (0, 1, 3, 4, 5, 8, 9)] == [0, 4, 4, 4, 8, 8, 12]


# ------------------------------------------------------------------------------------------------
# reference tracer/scheduler_test.go:14-26,46-61
@pytest.mark.parametrize("s1,s2,h,e1,e2", [(1, 2, 10, 4, 6), (2, 1, 10, 7, 3), (1, 1000, 10, 1, 9)])
def test_naive_scheduler_table(s1, s2, h, e1, e2):
    rows = NaiveScheduler().schedule([StaticSpeed(s1), StaticSpeed(s2)], h)
    assert list(rows) == [e1, e2]


def test_perfect_scheduler_table():
    t1, t2 = StaticSpeed(1), StaticSpeed(1)
    sch = PerfectScheduler()
    for h, r1, r2, e1, e2 in [(10, 1, 5, 5, 5), (10, 1, 5, 9, 1), (10, 5, 1, 7, 3)]:
        # the Go test leaves Stats.BlockH at whatever the previous Schedule call assigned
        prev = list(sch.block_assignment) if sch.block_assignment else [0, 0]
        t1.set_stats(prev[0], r1 * 1e-9)
        t2.set_stats(prev[1], r2 * 1e-9)
        rows = sch.schedule([t1, t2], h)
        assert list(rows) == [e1, e2]


def test_scheduler_rows_cover_frame():
    rng = np.random.default_rng(3)
    for _ in range(50):
        n = int(rng.integers(1, 9))
        h = int(rng.integers(n, 2161))
        trs = [StaticSpeed(int(rng.integers(1, 400))) for _ in range(n)]
        sch = PerfectScheduler()
        rows = sch.schedule(trs, h)
        assert sum(rows) >= h and min(rows) >= 1
        for t, r in zip(trs, rows):
            t.set_stats(r, float(rng.uniform(0.01, 2.0)))
        rows = sch.schedule(trs, h)
        assert min(rows) >= 1 and sum(rows) >= h - n  # floor() may leave < n rows unassigned before the top-up


# ------------------------------------------------------------------------------------------------
def test_material_codes_and_defaults():
    """Type codes are the contract with the device code (bxdf.go:6-17 == CL/bxdf/bxdf.cl:13-18,
    op.go:7-17 == CL/samplers/material_sampler.cl:4-8); defaults from defaults.go:6-13."""
    assert (M.BXDF_EMISSIVE, M.BXDF_DIFFUSE, M.BXDF_CONDUCTOR, M.BXDF_ROUGH_CONDUCTOR, M.BXDF_DIELECTRIC,
            M.BXDF_ROUGH_DIELECTRIC) == (2, 4, 8, 16, 32, 64)
    assert (M.OP_MIX, M.OP_MIX_MAP, M.OP_BUMP_MAP, M.OP_NORMAL_MAP, M.OP_DISPERSE) == (10001, 10002, 10003, 10004, 10005)
    mc = M.MaterialCompiler({"a": "diffuse()", "b": 'mix(diffuse(reflectance: {0.8,0.8,0.8}), conductor(intIOR: "silver"), 0.6)',
                             "c": "disperse(roughDielectric(roughness: 0.2), intIOR: {1.50, 1.52, 1.54}, extIOR: {0, 0, 0})"}, {})
    ra, rb, rc = mc.generate("a"), mc.generate("b"), mc.generate("c")
    nodes = mc.node_array()
    a = nodes[ra]
    assert a["union1"].tolist() == [M.BXDF_DIFFUSE, 0, -1, -1] or a["union1"][0] == M.BXDF_DIFFUSE
    assert np.allclose(a["union2"][:3], 0.2)                       # default reflectance
    assert np.isclose(a["union4"][0], 1.51714) and np.isclose(a["union4"][1], 1.0002926)  # Glass / Air on every leaf
    assert a["union5"][0] == -1
    b = nodes[rb]  # post-order: children first, the op node last (compiler.go:398-407,458-459)
    assert b["union1"][0] == M.OP_MIX and b["union1"][1] < rb and b["union1"][2] < rb
    assert np.isclose(b["union2"][0], 0.6)
    left, right = nodes[b["union1"][1]], nodes[b["union1"][2]]
    assert left["union1"][0] == M.BXDF_DIFFUSE and right["union1"][0] == M.BXDF_CONDUCTOR
    assert np.isclose(right["union4"][0], 0.18)                    # ior.go: Silver
    c = nodes[rc]
    assert c["union1"][0] == M.OP_DISPERSE and np.allclose(c["union2"][:3], [1.50, 1.52, 1.54]) and np.allclose(c["union3"][:3], 0)
    assert nodes[c["union1"][1]]["union1"][0] == M.BXDF_ROUGH_DIELECTRIC
    with pytest.raises(M.MaterialError):
        M.MaterialCompiler({"bad": "diffuse(reflectance: {1.5, 0, 0})"}, {}).generate("bad")  # node.go:138-163


# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("key", ["c1", "c2", "c3", "c4"])
def test_compiled_scene_invariants(key):
    sc = C.small_scene(key, 64, 64)
    n = sc.bvh_nodes
    nt = sc.num_triangles
    assert sc.vertices.shape == (3 * nt, 4) and sc.normals.shape == (3 * nt, 4) and sc.uvs.shape == (3 * nt, 2)
    inner = n["ldata"] > 0
    assert (n["rdata"][inner] > 0).all()
    # every node is referenced exactly once (node 0 = scene root, mesh roots via the instances)
    refs = np.concatenate([n["ldata"][inner], n["rdata"][inner], sc.mesh_instances["bvh_root"].astype(np.int64)])
    uniq = np.unique(refs)
    assert set(uniq.tolist()) | {0} == set(range(len(n)))
    top_leaf = (~inner) & (n["rdata"] == 0)
    mesh_leaf = (~inner) & (n["rdata"] > 0)
    assert sorted((-n["ldata"][top_leaf]).tolist()) == list(range(len(sc.mesh_instances)))
    # mesh leaves tile the triangle array without gaps or overlap, in leaf order (compiler.go:128-170)
    first, cnt = -n["ldata"][mesh_leaf], n["rdata"][mesh_leaf]
    o = np.argsort(first)
    assert first[o][0] == 0 and (first[o][1:] == (first[o] + cnt[o])[:-1]).all() and (first[o] + cnt[o])[-1] == nt
    # child boxes lie inside their parent's
    for side in ("ldata", "rdata"):
        ch = n[side][inner]
        assert (n["min"][ch] >= n["min"][inner] - 1e-6).all() and (n["max"][ch] <= n["max"][inner] + 1e-6).all()
    assert (sc.material_index < len(sc.material_nodes)).all()
    assert len(sc.emissives) >= 1
    e = sc.emissives
    area = e["type"] == 0
    assert (e["area"][area] > 0).all() and (e["prim_index"][area] < nt).all()
    assert np.abs(sc.vertices[:, :3]).max() <= 100.0


@pytest.mark.parametrize("key", ["c1", "c2", "c3", "c4"])
def test_scene_dump_roundtrip(tmp_path, key):
    sc = C.small_scene(key, 64, 64)
    p = tmp_path / "scene.bin"
    sc.save(p)
    back = S.Scene.load(p)
    for name in S.Scene._SECTIONS:
        assert np.ascontiguousarray(getattr(sc, name)).tobytes() == np.ascontiguousarray(getattr(back, name)).tobytes(), name
    assert (back.scene_diffuse_mat_index, back.scene_emissive_mat_index) == (sc.scene_diffuse_mat_index, sc.scene_emissive_mat_index)
    # the camera too, bit for bit: Update() writes position + unit direction back into LookAt (camera.go), so the dump keeps
    # the LookAt the camera was constructed with (a c1 camera re-normalised from the updated LookAt is 1 ulp off)
    back.camera.setup_projection(F(1.0))
    assert back.camera.frustrum.tobytes() == sc.camera.frustrum.tobytes()


def test_scene_dump_follows_a_moved_camera(tmp_path):
    """The dump keeps the LookAt the camera was GIVEN -- also after somebody moved the camera and updated it again."""
    import copy
    sc = copy.deepcopy(C.small_scene("c1", 64, 64))
    cam = sc.camera
    cam.position = (cam.position + np.array([0.25, -0.5, 0.75], F)).astype(F)
    cam.look_at = np.array([0.1, 0.2, -0.3], F)
    cam.update()
    cam.update()  # a second update re-normalises the written-back LookAt, like the interactive renderer does every frame
    p = tmp_path / "moved.bin"
    sc.save(p)
    back = S.Scene.load(p)
    back.camera.setup_projection(F(1.0))
    fresh = S.Camera(cam.position.copy(), np.array([0.1, 0.2, -0.3], F), cam.up.copy(), cam.fov)
    fresh.setup_projection(F(1.0))
    assert back.camera.frustrum.tobytes() == fresh.frustrum.tobytes()
    assert np.allclose(back.camera.frustrum, cam.frustrum, atol=1e-6)


def test_scene_cache_returns_the_same_scene(tmp_path, monkeypatch):
    """POLARIS_SCENE_CACHE (scenes.build): the second build loads the dump and must be indistinguishable from the first."""
    from polaris_b200 import scenes
    from tests.golden.make_golden import scene_digest
    monkeypatch.setenv("POLARIS_SCENE_CACHE", str(tmp_path))
    a = scenes.build("c1_sphere", 96, 64)[0]
    b = scenes.build("c1_sphere", 96, 64)[0]
    assert os.path.exists(tmp_path / "c1_sphere_96x64.plrscn")
    assert scene_digest(a) == scene_digest(b)


def test_camera_frustum():
    """camera.go:121-141: corners of inv(Proj*View) minus eye; symmetric for a centred camera, FOV
    consumed as radians (matrix.go:156-161)."""
    cam = S.Camera(np.array([0, 0, 5], F), np.array([0, 0, 0], F), np.array([0, 1, 0], F), 45.0)
    cam.setup_projection(F(1.0))
    tl, tr, bl, br = cam.frustrum[:, :3]
    assert np.allclose(tl * [-1, 1, 1], tr, atol=1e-5) and np.allclose(bl * [-1, 1, 1], br, atol=1e-5)
    assert np.allclose(tl * [1, -1, 1], bl, atol=1e-5)
    assert (cam.frustrum[:, 2] < 0).all() and (cam.frustrum[:, 3] == 0).all()
    f = 1.0 / np.tan(45.0 / 2.0)  # radians!
    assert np.isclose(abs(tl[0] / tl[2]), abs(1.0 / f), rtol=1e-4)


def test_instance_matrix_known_answers():
    """wavefront_test.go TestMeshInstancing: `instance testObj 1 0 1 0 0 0 1 1 1` maps (0,0,0) -> (1,0,1)
    (translation-only instances are the ones the synthetic scenes use)."""
    from polaris_b200 import gotypes as gt

    m = gt.mul4(gt.scale4((1, 1, 1)), gt.mul4(gt.ident4(), gt.translate4((1, 0, 1))))
    assert np.allclose(gt.mul4x1(m, np.array([0, 0, 0, 1], F))[:3], [1, 0, 1])
    assert np.allclose(gt.scale4((0, 0, 0)), gt.ident4())  # Scale4 maps a zero scale to 1 (matrix.go:42-53)
    inv = gt.inv4(m)
    assert np.allclose(gt.mul4x1(inv, np.array([1, 0, 1, 1], F))[:3], [0, 0, 0])
    assert np.array_equal(gt.inv4(np.zeros(16, F)), np.zeros(16, F))  # |det| < 1e-10 -> zero matrix (matrix.go:108-138)
