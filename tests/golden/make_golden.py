#!/usr/bin/env python
"""Generate the golden vectors of tests/golden/*.npz FROM THE REFERENCE'S OWN KERNELS.

Run in the container that holds /root/reference, after `python oracle/build_ref.py`:

    python tests/golden/make_golden.py

Every output below is produced by oracle/_ref/libpolaris_clref.so (tracer/opencl/CL/*.cl compiled for
the CPU), not by the oracle port and not by the CUDA path: the fixtures are what pins both of those
when the library itself is absent (tests/test_cpu_golden.py, tests/test_gpu_parity.py::test_golden_*).
Inputs are the procedural small scenes of tests/common.py; a SHA-256 of the compiled scene buffers is
stored with each fixture so generator drift is reported as such, not as a parity failure.
"""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import ref_binding  # noqa: E402
from polaris_b200 import _lib  # noqa: E402
from polaris_b200 import tracer as T  # noqa: E402
from tests import common as C  # noqa: E402

CONFIGS = {"c1": (64, 64), "c2": (64, 64), "c3": (80, 48), "c4": (64, 48)}
N_RAYS = 1024
SPP = 2


def scene_digest(sc):
    h = hashlib.sha256()
    for name in sc._SECTIONS:
        h.update(np.ascontiguousarray(getattr(sc, name)).tobytes())
    h.update(np.ascontiguousarray(sc.camera.frustrum).tobytes())
    return h.hexdigest()


def main():
    assert ref_binding.available(), "build oracle/_ref first: python oracle/build_ref.py"
    ref0 = ref_binding.RefTracer()
    # --- RNG (random_sampler.cl:7-16) and tonemap (hdr.cl:5-28)
    states = np.array([[0, 0], [1, 2], [0xFFFFFFFF, 7], [0x501A2150, 123456]], dtype=np.uint32)
    out, final = ref0.debug_rng(states, 8)
    rng = np.random.default_rng(5)
    acc = np.zeros((64, 4), np.float32)
    acc[:16, :3] = np.array([[0, 0, 0], [1e-6, 1, 100], [0.5, 0.25, 0.125], [1e4, 3, 0.01]] * 4, np.float32)
    acc[16:, :3] = np.exp(rng.uniform(-8, 8, size=(48, 3))).astype(np.float32)
    rgba = ref0.debug_tonemap(acc, 1.0 / 16, 1.2)
    np.savez_compressed(os.path.join(HERE, "scalars.npz"), rng_states=states, rng_out=out, rng_final=final,
                        tonemap_acc=acc, tonemap_rgba=rgba, tonemap_weight=np.float32(1.0 / 16), tonemap_exposure=np.float32(1.2))
    # --- per config: fixed ray set -> hit records; full-depth frame; bounce-0 frame
    for key, (w, h) in CONFIGS.items():
        sc = C.small_scene(key, w, h)
        ref = C.setup(ref_binding.RefTracer(), sc, w, h)
        rays = C.fixed_rays(sc, w, h, n=N_RAYS)
        flags, hits = ref.debug_intersect(rays, 0)
        occ = rays.copy()
        t = hits["wuvt"][:, 3]
        with np.errstate(over="ignore"):
            occ["origin"][:, 3] = np.where(flags == 1, t * np.where(np.arange(len(t)) % 3 == 0, np.float32(0.5), np.float32(1.5)), np.float32(3.0))
        occ_flags, _ = ref.debug_intersect(occ, 1)
        hit = flags == 1
        hits_clean = hits.copy()
        hits_clean["pad"] = 0
        for f in ("mesh_instance", "tri_index"):  # undefined on a miss in the reference: zero them in the fixture
            hits_clean[f][~hit] = 0
        hits_clean["wuvt"][~hit, :3] = 0
        seeds = T.splitmix_seeds(40 + int(key[1]), SPP * 6)
        r = T.make_block_request(w, h, spp=SPP)
        ref.trace(r, seeds)
        full = C.acc_of(ref, _lib.BUF_TRACE_ACCUMULATOR, w, h).copy()
        counters = ref.read_buffer(_lib.BUF_RAY_COUNTERS, 3, np.int32).copy()
        st = ref.stats().device
        ref.merge_output(ref, r)
        ref.sync_framebuffer(T.make_block_request(w, h, spp=SPP))
        rgba = ref.frame_buffer.copy()
        r0 = T.make_block_request(w, h, spp=1, num_bounces=1)
        ref.trace(r0, seeds[:2])
        b0 = C.acc_of(ref, _lib.BUF_TRACE_ACCUMULATOR, w, h).copy()
        prim = ref.read_buffer(_lib.BUF_RAYS0, w * h, _lib.RAY_DTYPE).copy()
        np.savez_compressed(
            os.path.join(HERE, f"{key}.npz"), scene_sha256=np.array(scene_digest(sc)), frame=np.array([w, h]), rays=rays,
            flags=flags, hits=hits_clean, occ_rays=occ, occ_flags=occ_flags, seeds=seeds, spp=np.array(SPP), full_acc=full,
            counters=counters, query_rays=np.array(st["query_rays"]), occlusion_rays=np.array(st["occlusion_rays"]),
            rgba=rgba, bounce0_acc=b0, primary_rays=prim)
        print(f"{key}: {w}x{h}, {int(hit.sum())}/{len(rays)} fixed rays hit, {st['query_rays']} query + {st['occlusion_rays']} occlusion rays, "
              f"mean radiance {full.mean():.5f}")
        ref.close()
    # --- BxDF tables (bxdf/*.cl) on the layered Cornell and the dispersive terrain materials
    for key in ("c2", "c4"):
        sc = C.small_scene(key, *CONFIGS[key])
        recs = C.bxdf_records(sc, n_theta=6, n_rand=4)
        ref = C.setup(ref_binding.RefTracer(), sc, *CONFIGS[key])
        np.savez_compressed(os.path.join(HERE, f"bxdf_{key}.npz"), records=recs, out=ref.debug_bxdf(recs), scene_sha256=np.array(scene_digest(sc)))
        ref.close()
    make_debug()
    print("golden vectors written to", HERE)


DEBUG_BOUNCES = 3


def make_debug():
    """debug pipeline stages (kernels/debug.cl + pipeline.go:113-200) from the reference's kernels: 2 + 5 x bounces frames"""
    for key in ("c2", "c4"):
        w, h = CONFIGS[key]
        sc = C.small_scene(key, w, h)
        ref = C.setup(ref_binding.RefTracer(), sc, w, h)
        seeds = T.splitmix_seeds(60 + int(key[1]), SPP * (1 + DEBUG_BOUNCES))
        frames = ref.trace_debug(T.make_block_request(w, h, spp=SPP, num_bounces=DEBUG_BOUNCES), seeds, _lib.DEBUG_ALL_STAGES)
        np.savez_compressed(os.path.join(HERE, f"debug_{key}.npz"), scene_sha256=np.array(scene_digest(sc)), seeds=seeds,
                            flags=np.array([f for f, _, _ in frames], np.uint32), bounces=np.array([b for _, b, _ in frames], np.uint32),
                            frames=np.stack([a for _, _, a in frames]),
                            path_flags=ref.read_buffer(_lib.BUF_PATHS, w * h, _lib.PATH_DTYPE)["flags"].copy())
        print(f"debug_{key}: {len(frames)} frames of {w}x{h}")
        ref.close()


if __name__ == "__main__":
    if "--only-debug" in sys.argv:
        make_debug()
    else:
        main()
