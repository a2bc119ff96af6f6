"""CPU tier: golden vectors generated from the reference's own kernels (tests/golden/make_golden.py)
against (a) the oracle port and (b) the CUDA tracer's device functions compiled for the host
(tests/emul, TEST ONLY -- the same pc_device.cuh / pc_layout.hpp the GPU runs: derived node64/tri48
layout, nearest-first traversal with closest-hit culling and the reference-order tie-break, shading).

Bars: bit-exact for RNG, hit flags, instance / triangle ids, barycentrics, distances, ray counters and
tonemapped bytes; radiance bit-exact for the oracle, and for the emulated CUDA code too because both are
built with -ffp-contract=off against the same libm (on the GPU itself the transcendental functions differ in
the last ulp, so tests/test_gpu_parity.py uses the tolerances BASELINE.json states).
"""
import os

import numpy as np
import pytest

from oracle.binding import OracleTracer
from polaris_b200 import _lib
from polaris_b200 import tracer as T

from . import common as C
from .golden.make_golden import CONFIGS, scene_digest

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def scene_for(key, g):
    w, h = CONFIGS[key]
    sc = C.small_scene(key, w, h)
    assert scene_digest(sc) == str(g["scene_sha256"]), \
        f"the procedural scene {key} no longer matches the one the golden vectors were made from: regenerate tests/golden"
    return sc, w, h


def test_golden_rng_and_tonemap():
    g = load("scalars")
    out, final = OracleTracer().debug_rng(g["rng_states"], 8)
    assert out.tobytes() == g["rng_out"].tobytes() and final.tobytes() == g["rng_final"].tobytes()
    # first draw of state (0,0): x = 0 -> (871483, 234234) * 2^-32 (random_sampler.cl:10-15, by hand)
    assert out[0, 0, 0] == np.float32(871483) * np.float32(1.0 / 4294967296.0)
    assert out[0, 0, 1] == np.float32(234234) * np.float32(1.0 / 4294967296.0)
    rgba = OracleTracer().debug_tonemap(g["tonemap_acc"], float(g["tonemap_weight"]), float(g["tonemap_exposure"]))
    assert np.array_equal(rgba, g["tonemap_rgba"])
    assert rgba[0].tolist() == [0, 0, 0, 255]  # black stays black, alpha is 255 (hdr.cl:22-27)
    sc = C.small_scene("c1", 32, 32)
    e = C.Emul(sc, 32, 32)
    out = np.zeros((len(g["tonemap_acc"]), 4), np.uint8)
    acc = np.ascontiguousarray(g["tonemap_acc"], np.float32)
    e.lib.pe_tonemap(acc.ctypes.data, len(acc), float(g["tonemap_weight"]), float(g["tonemap_exposure"]), out.ctypes.data)
    assert np.array_equal(out, g["tonemap_rgba"])
    e.close()


@pytest.mark.parametrize("key", ["c1", "c2", "c3", "c4"])
def test_golden_hit_records(key):
    g = load(key)
    sc, w, h = scene_for(key, g)
    hit = g["flags"] == 1
    orc = C.oracle_for(sc, w, h)
    emu = C.Emul(sc, w, h)
    cases = {"oracle": orc.debug_intersect(g["rays"], 0), "cuda per-ray (host build)": emu.intersect(g["rays"], 0)[:2],
             "cuda reference-order (host build)": emu.intersect(g["rays"], 2)[:2]}
    for label, (flags, hits) in cases.items():
        assert np.array_equal(flags, g["flags"]), label
        names = ("mesh_instance", "tri_index") if "mesh_instance" in hits.dtype.names else ("inst", "tri")
        for mine, gold in zip(names, ("mesh_instance", "tri_index")):
            assert np.array_equal(hits[mine][hit], g["hits"][gold][hit]), f"{label}: {gold}"
        assert hits["wuvt"][hit].tobytes() == g["hits"]["wuvt"][hit].tobytes(), f"{label}: wuvt"
    assert np.array_equal(orc.debug_intersect(g["occ_rays"], 1)[0], g["occ_flags"])
    assert np.array_equal(emu.intersect(g["occ_rays"], 1)[0], g["occ_flags"])
    assert np.array_equal(emu.intersect(g["occ_rays"], 3)[0], g["occ_flags"])
    assert 0 < g["occ_flags"].sum() < len(g["occ_flags"])
    emu.close()


@pytest.mark.parametrize("key", ["c1", "c2", "c3", "c4"])
def test_golden_frames(key):
    g = load(key)
    sc, w, h = scene_for(key, g)
    spp, seeds = int(g["spp"]), g["seeds"]
    orc = C.oracle_for(sc, w, h)
    r = T.make_block_request(w, h, spp=spp)
    orc.trace(r, seeds)
    assert C.acc_of(orc, _lib.BUF_TRACE_ACCUMULATOR, w, h).tobytes() == g["full_acc"].tobytes()
    assert np.array_equal(orc.read_buffer(_lib.BUF_RAY_COUNTERS, 3, np.int32), g["counters"])
    st = orc.stats().device
    assert (st["query_rays"], st["occlusion_rays"]) == (int(g["query_rays"]), int(g["occlusion_rays"]))
    orc.merge_output(orc, r)
    orc.sync_framebuffer(T.make_block_request(w, h, spp=spp))
    assert np.array_equal(orc.frame_buffer, g["rgba"])
    r0 = T.make_block_request(w, h, spp=1, num_bounces=1)
    orc.trace(r0, seeds[:2])
    assert orc.read_buffer(_lib.BUF_RAYS0, w * h, _lib.RAY_DTYPE).tobytes() == g["primary_rays"].tobytes()
    assert C.acc_of(orc, _lib.BUF_TRACE_ACCUMULATOR, w, h).tobytes() == g["bounce0_acc"].tobytes()
    # the CUDA device code, host build: same frame
    emu = C.Emul(sc, w, h)
    emu.trace(T.make_block_request(w, h, spp=spp), seeds)
    acc = emu.read_buffer(_lib.BUF_TRACE_ACCUMULATOR, w * h * 4, np.float32).reshape(-1, 4)[:, :3]
    assert acc.tobytes() == g["full_acc"].tobytes()
    assert np.array_equal(emu.read_buffer(_lib.BUF_RAY_COUNTERS, 3, np.int32), g["counters"])
    emu.close()


@pytest.mark.parametrize("key", ["c2", "c4"])
def test_golden_bxdf_tables(key):
    g = load("bxdf_" + key)
    w, h = CONFIGS[key]
    sc = C.small_scene(key, w, h)
    assert scene_digest(sc) == str(g["scene_sha256"])
    o = C.oracle_for(sc, w, h).debug_bxdf(g["records"])
    emu = C.Emul(sc, w, h)
    e = emu.bxdf(g["records"])
    for f in ("sample", "sample_pdf", "dir", "pdf", "eval"):
        assert o[f].tobytes() == g["out"][f].tobytes(), f"oracle {f}"
        assert e[f].tobytes() == g["out"][f].tobytes(), f"cuda host build {f}"
    emu.close()


def test_derived_layout_properties():
    """pc_layout.hpp: the 16-byte aligned traversal layout derived at upload from the reference buffers."""
    for key in ("c1", "c2", "c3", "c4"):
        w, h = CONFIGS[key]
        sc = C.small_scene(key, w, h)
        emu = C.Emul(sc, w, h)
        info = emu.layout_info()
        n = sc.bvh_nodes
        assert info["inner_nodes"] == int((n["ldata"] > 0).sum())
        assert info["top_depth"] >= 0 and info["mesh_depth"] >= 0 and info["stack_need"] >= 1
        assert info["stack_need"] <= 64  # PC_STACK_SIZE; deeper scenes are refused with PC_ERR_STACK_DEPTH
        emu.close()


def test_triangle_pretest_is_conservative():
    """(PC_TRI_PRETEST, measured slower and off by default, kept as a documented experiment.)
    pc_device.cuh triTest rejects without the IEEE reciprocal only when the reference's own `u < 0 || u > 1`
    (intersect.cl:266-269) is certain to reject: check the implication on 4M (a, det) pairs spanning the float32
    range, including denormals, huge values and the neighbourhood of the decision boundaries."""
    rng = np.random.default_rng(11)
    n = 1 << 20
    parts = []
    for lo, hi in ((-150, 128), (-30, 30), (-10, 10)):
        a = (rng.choice([-1.0, 1.0], n) * np.exp2(rng.uniform(lo, hi, n))).astype(np.float32)
        det = (rng.choice([-1.0, 1.0], n) * np.exp2(rng.uniform(max(lo, -17), hi, n))).astype(np.float32)  # |det| >= 1e-5
        parts.append((a, det))
    det = (rng.choice([-1.0, 1.0], n) * np.exp2(rng.uniform(-16, 20, n))).astype(np.float32)
    a = (det.astype(np.float64) * (1.0 + rng.uniform(-3e-5, 3e-5, n))).astype(np.float32)  # u close to 1
    parts.append((a, det))
    parts.append(((rng.choice([-1.0, 1.0], n) * np.exp2(rng.uniform(-24, -16, n))).astype(np.float32), det))  # |a| around 2^-20
    for k, (a, det) in enumerate(parts):
        keep = np.abs(det) >= np.float32(1e-5)
        a, det = a[keep], det[keep]
        with np.errstate(all="ignore"):
            inv = (np.float32(1.0) / det).astype(np.float32)
            u = (a * inv).astype(np.float32)
            ref_reject = (u < 0) | (u > 1)
            neg = (a < 0) != (det < 0)
            aa = np.abs(a)
            pre = np.where(neg, aa >= np.float32(9.5367431640625e-07), aa > (np.abs(det) * np.float32(1.00001)).astype(np.float32))
        assert not (pre & ~ref_reject).any()
        if k < 3:
            assert pre.mean() > 0.3  # and the filter does fire on generic inputs


@pytest.mark.parametrize("key", ["c1", "c2", "c3", "c4"])
def test_wide_collapse_preserves_hit_records(key):
    """Design-tool check (tools/wide_bvh_model.py): walking 4-ary nodes that hold every inner node's grandchildren visits
    the same leaves as the binary walk -- the skipped intermediate boxes contain the tested ones and the float slab test is
    monotone in the bounds -- so flags, instance / triangle ids and w,u,v,t are identical bit for bit."""
    import ctypes

    w, h = CONFIGS[key]
    sc = C.small_scene(key, w, h)
    rays = C.fixed_rays(sc, w, h, n=2048)
    emu = C.Emul(sc, w, h)
    vp = ctypes.c_void_p
    emu.lib.pe_simt_wide.argtypes = [vp, vp, ctypes.c_uint32, ctypes.c_int, vp, vp, vp]
    for any_hit in (0, 1):
        f0, h0, _ = emu.intersect(rays, any_hit)
        out = np.zeros(8)
        flags = np.zeros(len(rays), np.uint32)
        hits = np.zeros(len(rays), _lib.INTERSECTION_DTYPE)
        emu.lib.pe_simt_wide(emu.h, rays.ctypes.data, len(rays), any_hit, out.ctypes.data, flags.ctypes.data, hits.ctypes.data)
        assert np.array_equal(flags, f0)
        if not any_hit:
            hit = f0 == 1
            assert hits["wuvt"][hit].tobytes() == h0["wuvt"][hit].tobytes()
            assert np.array_equal(hits["mesh_instance"][hit], h0["mesh_instance"][hit]) and np.array_equal(hits["tri_index"][hit], h0["tri_index"][hit])
        assert out[0] > 0 and out[7] == len(rays)


@pytest.mark.parametrize("key", ["c1", "c2", "c3", "c4"])
def test_wide_traversal_variant_reproduces_the_golden_vectors(key):
    """PC_WIDE_BVH (the 4-ary traversal experiment, compiled out of the product build): the SAME device code the kernels
    would run -- pc_layout.hpp build_wide + pc_device.cuh travInnerWide -- built for the host, against the golden hit records
    and the golden full-depth frame made from the reference's kernels: bit for bit, like the binary walk."""
    g = load(key)
    sc, w, h = scene_for(key, g)
    emu = C.Emul(sc, w, h, variant="_wide")
    assert emu.layout_info()["stack_need"] <= 64
    flags, hits, _ = emu.intersect(g["rays"], 0)
    assert np.array_equal(flags, g["flags"])
    hit = g["flags"] == 1
    assert hits["wuvt"][hit].tobytes() == g["hits"]["wuvt"][hit].tobytes()
    assert np.array_equal(hits["mesh_instance"][hit], g["hits"]["mesh_instance"][hit]) and np.array_equal(hits["tri_index"][hit], g["hits"]["tri_index"][hit])
    occ_flags, _, _ = emu.intersect(g["occ_rays"], 1)
    assert np.array_equal(occ_flags, g["occ_flags"])
    spp, seeds = int(g["spp"]), g["seeds"]
    emu.trace(T.make_block_request(w, h, spp=spp), seeds)
    assert C.acc_of(emu, _lib.BUF_TRACE_ACCUMULATOR, w, h).tobytes() == g["full_acc"].tobytes()


@pytest.mark.parametrize("key", ["c1", "c2", "c3", "c4"])
def test_pop_culling_variant_reproduces_the_golden_vectors(key):
    """PC_POP_CULL (stacked far children carry their entry distance and are tested again when popped; measured slower on the
    GPU and compiled out of the product build): the same device code built for the host returns the golden
    hit records and the golden full-depth frame bit for bit -- a subtree culled at pop time cannot hold a hit with
    t <= best -- while visiting no more nodes or triangles than the plain walk."""
    g = load(key)
    sc, w, h = scene_for(key, g)
    emu, plain = C.Emul(sc, w, h, variant="_popcull"), C.Emul(sc, w, h)
    flags, hits, cnt = emu.intersect(g["rays"], 0)
    _, _, cnt0 = plain.intersect(g["rays"], 0)
    assert np.array_equal(flags, g["flags"])
    hit = g["flags"] == 1
    assert hits["wuvt"][hit].tobytes() == g["hits"]["wuvt"][hit].tobytes()
    assert np.array_equal(hits["mesh_instance"][hit], g["hits"]["mesh_instance"][hit]) and np.array_equal(hits["tri_index"][hit], g["hits"]["tri_index"][hit])
    assert cnt[0] <= cnt0[0] and cnt[1] <= cnt0[1]
    occ_flags, _, _ = emu.intersect(g["occ_rays"], 1)
    assert np.array_equal(occ_flags, g["occ_flags"])
    spp, seeds = int(g["spp"]), g["seeds"]
    emu.trace(T.make_block_request(w, h, spp=spp), seeds)
    assert C.acc_of(emu, _lib.BUF_TRACE_ACCUMULATOR, w, h).tobytes() == g["full_acc"].tobytes()
