"""The device BVH build / scene compiler (csrc/pc_bvh_build.cu) against the host build: the reference's
tree (asset/compiler/bvh/bvh_builder.go:124-224, asset/compiler/compiler.go:81-231) must come out byte for byte -- nodes,
leaf order, leaf-ordered triangle arrays, instances, emissives -- or hit ids are no longer comparable index by index."""
import time

import numpy as np
import pytest

from polaris_b200 import scenes
from polaris_b200.scene import build_bvh, compile_scene

from . import common as C

pytestmark = pytest.mark.gpu

F = np.float32


def _boxes(centres, half):
    centres = np.asarray(centres, F)
    half = np.broadcast_to(np.asarray(half, F), centres.shape)
    return (centres - half).astype(F), (centres + half).astype(F), centres


def test_known_answers_of_the_reference_tests():
    """bvh_builder_test.go:45-68: 4 separated boxes, leaf size 1 -> 7 nodes / 4 leaves; leaf size 2 -> 3 nodes / 2 leaves."""
    lo, hi, c = _boxes([[0, 0, 0], [10, 0, 0], [0, 10, 0], [10, 10, 0]], 1.0)
    for leaf, want_nodes, want_leaves in ((1, 7, 4), (2, 3, 2)):
        nodes, order = build_bvh(lo, hi, c, leaf, builder="cuda")
        ref_nodes, ref_order = build_bvh(lo, hi, c, leaf, builder="host")
        assert len(nodes) == want_nodes and int((nodes["ldata"] <= 0).sum()) == want_leaves
        assert nodes.tobytes() == ref_nodes.tobytes() and order.tobytes() == ref_order.tobytes()


@pytest.mark.parametrize("n,leaf,kind", [(1, 1, "uniform"), (2, 1, "uniform"), (37, 1, "uniform"), (1000, 1, "lattice"), (5000, 10, "uniform"),
                                          (5000, 10, "clustered"), (4097, 10, "flat"), (3000, 10, "same_centre"), (20000, 10, "ties"),
                                          (300000, 10, "uniform")])
def test_random_volumes_identical_to_host_build(n, leaf, kind):
    rng = np.random.default_rng(n + leaf)
    if kind == "uniform":
        c = rng.uniform(-50, 50, (n, 3))
    elif kind == "lattice":  # the instancing config's top level: many equal scores -> tie-breaking in (axis, plane) order
        g = np.stack(np.meshgrid(*[np.arange(10)] * 3, indexing="ij"), -1).reshape(-1, 3)[:n]
        c = g * 3.0
    elif kind == "clustered":
        c = rng.normal(0, 1, (n, 3)) * rng.choice([0.01, 1.0, 30.0], (n, 1))
    elif kind == "flat":  # one axis thinner than minSideLength: that axis is skipped (bvh_builder.go:157)
        c = rng.uniform(-20, 20, (n, 3))
        c[:, 1] = 1e-5 * rng.uniform(0, 1, n)
    elif kind == "same_centre":  # no plane separates anything: one leaf with ALL items (:193-195)
        c = np.zeros((n, 3)) + 1.5
    else:  # "ties": coordinates on a coarse grid, many items share a centre exactly
        c = np.round(rng.uniform(-8, 8, (n, 3)) * 2) / 2
    half = rng.uniform(0.01, 0.5, (n, 3)) if kind != "flat" else np.abs(rng.uniform(1e-6, 1e-5, (n, 3)))
    lo, hi, c = _boxes(c, half)
    t0 = time.time()
    nodes, order = build_bvh(lo, hi, c, leaf, builder="cuda")
    t1 = time.time()
    ref_nodes, ref_order = build_bvh(lo, hi, c, leaf, builder="host")
    t2 = time.time()
    print(f"{kind} n={n} leaf={leaf}: {len(nodes)} nodes; device {t1 - t0:.3f}s host {t2 - t1:.3f}s")
    assert len(nodes) == len(ref_nodes)
    assert order.tobytes() == ref_order.tobytes(), "leaf item order differs"
    assert nodes.tobytes() == ref_nodes.tobytes(), "node array differs"
    assert sorted(order.tolist()) == list(range(n))


def _same_scene(a, b):
    for sec in a._SECTIONS:
        x, y = getattr(a, sec), getattr(b, sec)
        assert x.shape == y.shape and x.tobytes() == y.tobytes(), f"section {sec} differs between the device and the host build"
    assert (a.scene_diffuse_mat_index, a.scene_emissive_mat_index, a.top_depth, a.mesh_depth) == \
           (b.scene_diffuse_mat_index, b.scene_emissive_mat_index, b.top_depth, b.mesh_depth)


@pytest.mark.parametrize("key", ["c1", "c2", "c3", "c4"])
def test_small_configs_compile_identically(key):
    name, sizes = C.SMALL[key]
    raw = scenes.raw_scene(name, **sizes)
    _same_scene(compile_scene(raw, aspect=1.5, builder="cuda"), compile_scene(raw, aspect=1.5, builder="host"))


@pytest.mark.parametrize("name", ["c1_sphere", "c2_cornell", "c3_instancing"] + ([] if __import__("os").environ.get("POLARIS_SKIP_FULL_C4") else ["c4_terrain"]))
def test_baseline_configs_compile_identically(name):
    """The five BASELINE configs at full size (c5 is c2's scene): 1 000 instances of a 100 k-triangle mesh, the
    10 M-triangle terrain.  Prints the two build times (the device one includes every host<->device copy)."""
    raw = scenes.raw_scene(name)
    t0 = time.time()
    dev = compile_scene(raw, aspect=16 / 9, builder="cuda")
    t1 = time.time()
    host = compile_scene(raw, aspect=16 / 9, builder="host")
    t2 = time.time()
    print(f"{name}: {dev.num_triangles} triangles, {len(dev.bvh_nodes)} nodes: device build {t1 - t0:.2f}s, host build {t2 - t1:.2f}s (whole compile_scene)")
    _same_scene(dev, host)
