"""Pins the oracle port (oracle/polaris_oracle.cpp) against the reference's OWN kernels.

oracle/_ref/libpolaris_clref.so is tracer/opencl/CL/*.cl of the reference compiled for the CPU by
oracle/build_ref.py (verbatim sources + an OpenCL-C compatibility header).  Both sides use IEEE float32
without contraction and the same libm, so the bar is BIT-EXACT everywhere: RNG, primary rays, hit
records, BxDF tables, ray counters, per-pixel radiance at full depth and over several samples, merge
and tonemap.  A restatement slip in the oracle shows up here as a single differing bit.

The library is prebuilt where /root/reference exists and travels with the snapshot; if it is absent
these tests skip and tests/test_cpu_golden.py (vectors generated FROM this library) still pins the oracle.
"""
import numpy as np
import pytest

from oracle import ref_binding
from oracle.binding import OracleTracer
from polaris_b200 import _lib
from polaris_b200 import tracer as T

from . import common as C

pytestmark = pytest.mark.skipif(not ref_binding.available(), reason="oracle/_ref not built (needs /root/reference)")


def ref_for(sc, w, h):
    return C.setup(ref_binding.RefTracer(), sc, w, h)


def test_rng_bit_exact():
    states = np.array([[0, 0], [1, 2], [0xFFFFFFFF, 7], [0x501A2150, 123456], [17, 0xDEADBEEF]], dtype=np.uint32)
    r, rs = ref_binding.RefTracer().debug_rng(states, 16)
    o, os_ = OracleTracer().debug_rng(states, 16)
    assert r.tobytes() == o.tobytes() and rs.tobytes() == os_.tobytes()
    assert r.min() >= 0.0 and r.max() <= 1.0  # can be exactly 1.0


def test_tonemap_bit_exact():
    rng = np.random.default_rng(5)
    acc = np.zeros((4096 + 16, 4), np.float32)
    acc[:16, :3] = np.array([[0, 0, 0], [1e-6, 1, 100], [0.5, 0.25, 0.125], [1e4, 3, 0.01]] * 4, np.float32)
    acc[16:, :3] = np.exp(rng.uniform(-8, 8, size=(4096, 3))).astype(np.float32)
    r = ref_binding.RefTracer().debug_tonemap(acc, 1.0 / 16, 1.2)
    o = OracleTracer().debug_tonemap(acc, 1.0 / 16, 1.2)
    assert np.array_equal(r, o) and (r[:, 3] == 255).all()


@pytest.mark.parametrize("key,w,h", [("c1", 128, 128), ("c2", 128, 128), ("c3", 160, 96), ("c4", 128, 96)])
def test_hit_records_bit_exact(key, w, h):
    sc = C.small_scene(key, w, h)
    rays = C.fixed_rays(sc, w, h)
    ref, orc = ref_for(sc, w, h), C.oracle_for(sc, w, h)
    rf, rh = ref.debug_intersect(rays, 0)
    of, oh = orc.debug_intersect(rays, 0)
    assert np.array_equal(rf, of)
    hit = of == 1
    assert 0 < hit.sum()
    for f in ("wuvt", "mesh_instance", "tri_index"):  # a missed ray's record is undefined in the reference
        assert rh[f][hit].tobytes() == oh[f][hit].tobytes(), f
    assert rh["wuvt"][~hit, 3].tobytes() == oh["wuvt"][~hit, 3].tobytes()  # ... except wuvt.w == tmax
    occ = rays.copy()
    t = oh["wuvt"][:, 3]
    with np.errstate(over="ignore"):
        occ["origin"][:, 3] = np.where(of == 1, t * np.where(np.arange(len(t)) % 3 == 0, np.float32(0.5), np.float32(1.5)), np.float32(3.0))
    assert np.array_equal(ref.debug_intersect(occ, 1)[0], orc.debug_intersect(occ, 1)[0])


@pytest.mark.parametrize("key", ["c2", "c4"])
def test_bxdf_tables_bit_exact(key):
    sc = C.small_scene(key, 64, 64)
    recs = C.bxdf_records(sc)
    r, o = ref_for(sc, 64, 64).debug_bxdf(recs), C.oracle_for(sc, 64, 64).debug_bxdf(recs)
    for f in ("sample", "sample_pdf", "dir", "pdf", "eval"):
        assert r[f].tobytes() == o[f].tobytes(), f


@pytest.mark.parametrize("key,w,h", [("c1", 96, 96), ("c2", 128, 128), ("c3", 160, 96), ("c4", 128, 96)])
def test_full_depth_frames_bit_exact(key, w, h):
    """3 samples x 5 bounces with Russian roulette: every buffer of the bufferSet ends up identical."""
    sc = C.small_scene(key, w, h)
    spp = 3
    seeds = T.splitmix_seeds(7, spp * 6)
    ref, orc = ref_for(sc, w, h), C.oracle_for(sc, w, h)
    rr, ro = T.make_block_request(w, h, spp=spp), T.make_block_request(w, h, spp=spp)
    ref.trace(rr, seeds)
    orc.trace(ro, seeds)
    assert (rr.seed, rr.accumulated_samples) == (ro.seed, ro.accumulated_samples)
    cr, co = ref.read_buffer(_lib.BUF_RAY_COUNTERS, 3, np.int32), orc.read_buffer(_lib.BUF_RAY_COUNTERS, 3, np.int32)
    assert np.array_equal(cr, co)
    n = w * h
    assert ref.read_buffer(_lib.BUF_RAYS0, n, _lib.RAY_DTYPE)[: cr[0]].tobytes() == orc.read_buffer(_lib.BUF_RAYS0, n, _lib.RAY_DTYPE)[: co[0]].tobytes()
    assert ref.read_buffer(_lib.BUF_RAYS2, n, _lib.RAY_DTYPE)[: cr[2]].tobytes() == orc.read_buffer(_lib.BUF_RAYS2, n, _lib.RAY_DTYPE)[: co[2]].tobytes()
    pr, po = ref.read_buffer(_lib.BUF_PATHS, n, _lib.PATH_DTYPE), orc.read_buffer(_lib.BUF_PATHS, n, _lib.PATH_DTYPE)
    for f in ("pixel_index", "flags"):
        assert np.array_equal(pr[f], po[f]), f
    assert pr["throughput"][:, :3].tobytes() == po["throughput"][:, :3].tobytes()
    a, b = C.acc_of(ref, _lib.BUF_TRACE_ACCUMULATOR, w, h), C.acc_of(orc, _lib.BUF_TRACE_ACCUMULATOR, w, h)
    assert a.tobytes() == b.tobytes()
    assert a.sum() > 0
    sr, so = ref.stats().device, orc.stats().device
    for k in ("query_rays", "occlusion_rays", "indirect_emitted", "occlusion_emitted"):
        assert sr[k] == so[k], k
    for tr, r in ((ref, rr), (orc, ro)):
        tr.merge_output(tr, r)
        tr.sync_framebuffer(T.make_block_request(w, h, spp=spp))
    assert C.acc_of(ref, _lib.BUF_FRAME_ACCUMULATOR, w, h).tobytes() == C.acc_of(orc, _lib.BUF_FRAME_ACCUMULATOR, w, h).tobytes()
    assert np.array_equal(ref.frame_buffer, orc.frame_buffer)


def test_row_block_and_q4():
    """BlockY > 0: the reference adds an emissive hit at accumulator[rayPathIndex] (pt_integrator.cl:106,
    SURVEY Q4).  The oracle reproduces that with fix_q4 = 0 and differs from it -- only in rows that see
    the light -- with the fix on."""
    w, h = 96, 96
    sc = C.small_scene("c2", w, h)
    seeds = T.splitmix_seeds(8, 6)
    ref, orc, fixed = ref_for(sc, w, h), C.oracle_for(sc, w, h), C.oracle_for(sc, w, h)
    orc.set_option(_lib.OPT_FIX_Q4, 0)
    accs = []
    for tr in (ref, orc, fixed):
        tr.trace(T.make_block_request(w, h, block_y=0, block_h=40, spp=1), seeds)  # the block that looks at the ceiling light
        accs.append(C.acc_of(tr, _lib.BUF_TRACE_ACCUMULATOR, w, h).copy())
        tr.trace(T.make_block_request(w, h, block_y=56, block_h=40, spp=1), seeds)
        accs.append(C.acc_of(tr, _lib.BUF_TRACE_ACCUMULATOR, w, h).copy())
    assert accs[0].tobytes() == accs[2].tobytes() == accs[4].tobytes()  # BlockY == 0: all agree
    assert accs[1].tobytes() == accs[3].tobytes()                       # BlockY > 0: literal behaviour
    assert accs[5].reshape(h, w, 3)[:56].sum() == 0                     # the fix keeps radiance inside the block's rows


def test_reference_stack_limit_is_enforced():
    """The kernels reserve 32 unchecked stack entries (intersect.cl:4); the driver refuses a
    scene that could overflow them instead of running into undefined behaviour."""
    sc = C.small_scene("c3", 64, 64)
    ref = ref_for(sc, 64, 64)
    assert 0 < ref.stack_need() <= 32
