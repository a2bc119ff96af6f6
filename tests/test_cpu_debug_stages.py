"""CPU tier: the debug pipeline stages (reference kernels/debug.cl:16-156 driven by pipeline.go:113-200).

  * oracle port (po_trace_debug) == the reference's own debug kernels compiled for the CPU (pr_trace_debug), byte for byte,
    including the side effect of the normals stage on the paths' dispersion bits and the unchanged radiance;
  * both == the committed golden frames (tests/golden/debug_*.npz, made from the reference's kernels);
  * frame order and count == the order in which the reference writes its debug-*.png files.
"""
import os

import numpy as np
import pytest

from oracle import ref_binding
from polaris_b200 import _lib
from polaris_b200 import tracer as T

from . import common as C
from .golden.make_golden import CONFIGS, DEBUG_BOUNCES, SPP, scene_digest

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
ORDER_PER_BOUNCE = [_lib.DEBUG_THROUGHPUT, _lib.DEBUG_ALL_EMISSIVE, _lib.DEBUG_VISIBLE_EMISSIVE, _lib.DEBUG_OCCLUDED_EMISSIVE, _lib.DEBUG_ACCUMULATOR]


def expected_order(flags, nb):
    out = [(f, 0) for f in (_lib.DEBUG_PRIMARY_DEPTH, _lib.DEBUG_PRIMARY_NORMALS) if flags & f]
    for b in range(nb):
        out += [(f, b) for f in ORDER_PER_BOUNCE if flags & f]
    return out


def test_debug_flag_values_and_file_names():
    # DebugFlag = 1 << iota with iota == 1 on the first flag's line (pipeline.go:19-30)
    assert [_lib.DEBUG_PRIMARY_DEPTH, _lib.DEBUG_PRIMARY_NORMALS, _lib.DEBUG_ALL_EMISSIVE, _lib.DEBUG_VISIBLE_EMISSIVE,
            _lib.DEBUG_OCCLUDED_EMISSIVE, _lib.DEBUG_THROUGHPUT, _lib.DEBUG_ACCUMULATOR, _lib.DEBUG_FRAMEBUFFER] == [2 << i for i in range(8)]
    assert _lib.debug_frame_count(_lib.DEBUG_ALL_STAGES, 5) == 2 + 5 * 5
    assert _lib.debug_frame_count(_lib.DEBUG_PRIMARY_DEPTH | _lib.DEBUG_ACCUMULATOR, 3) == 1 + 3
    assert _lib.DEBUG_FILE_NAMES[_lib.DEBUG_THROUGHPUT] % 2 == "debug-throughput-002.png"


@pytest.mark.parametrize("key", ["c2", "c4"])
def test_debug_stages_golden(key):
    g = np.load(os.path.join(GOLDEN, f"debug_{key}.npz"))
    w, h = CONFIGS[key]
    sc = C.small_scene(key, w, h)
    assert scene_digest(sc) == str(g["scene_sha256"]), "regenerate tests/golden (python tests/golden/make_golden.py --only-debug)"
    orc = C.oracle_for(sc, w, h)
    frames = orc.trace_debug(T.make_block_request(w, h, spp=SPP, num_bounces=DEBUG_BOUNCES), g["seeds"], _lib.DEBUG_ALL_STAGES)
    assert [(f, b) for f, b, _ in frames] == expected_order(_lib.DEBUG_ALL_STAGES, DEBUG_BOUNCES)
    assert [(f, b) for f, b, _ in frames] == list(zip(g["flags"].tolist(), g["bounces"].tolist()))
    for i, (f, b, a) in enumerate(frames):
        assert a.tobytes() == g["frames"][i].tobytes(), f"stage {f} bounce {b} differs from the reference's debug kernel"
        assert (a[..., 3] == 255).all()
    assert np.array_equal(orc.read_buffer(_lib.BUF_PATHS, w * h, _lib.PATH_DTYPE)["flags"], g["path_flags"])
    # the stages only read the tracer state (apart from the normals stage's matSelectNode): radiance is what pc_trace gives
    acc_dbg = C.acc_of(orc, _lib.BUF_TRACE_ACCUMULATOR, w, h).copy()
    orc.trace(T.make_block_request(w, h, spp=SPP, num_bounces=DEBUG_BOUNCES), g["seeds"])
    if key == "c2":  # no dispersion in the Cornell materials: identical
        assert acc_dbg.tobytes() == C.acc_of(orc, _lib.BUF_TRACE_ACCUMULATOR, w, h).tobytes()
    # not every frame is blank
    assert all(fr[..., :3].any() for fr in g["frames"][:3])
    orc.close()


@pytest.mark.skipif(not ref_binding.available(), reason="oracle/_ref not built")
@pytest.mark.parametrize("key,flags,spp,nb", [("c1", _lib.DEBUG_ALL_STAGES, 1, 2), ("c3", _lib.DEBUG_PRIMARY_DEPTH | _lib.DEBUG_VISIBLE_EMISSIVE, 2, 2),
                                              ("c4", _lib.DEBUG_PRIMARY_NORMALS | _lib.DEBUG_THROUGHPUT | _lib.DEBUG_ACCUMULATOR, 3, 4)])
def test_debug_stages_port_vs_reference_kernels(key, flags, spp, nb):
    w, h = CONFIGS[key]
    sc = C.small_scene(key, w, h)
    seeds = T.splitmix_seeds(77, spp * (1 + nb))
    orc, ref = C.oracle_for(sc, w, h), C.setup(ref_binding.RefTracer(), sc, w, h)
    orc.set_option(_lib.OPT_FIX_Q4, 0)  # literal pt_integrator.cl:106 so that BlockY > 0 is comparable
    by = 8  # a row block: pixelIndex != work-item index, the accumulator stage indexes by work-item (debug.cl:153)
    fo = orc.trace_debug(T.make_block_request(w, h, block_y=by, block_h=h - 2 * by, spp=spp, num_bounces=nb), seeds, flags)
    fr = ref.trace_debug(T.make_block_request(w, h, block_y=by, block_h=h - 2 * by, spp=spp, num_bounces=nb), seeds, flags)
    assert [(f, b) for f, b, _ in fo] == [(f, b) for f, b, _ in fr] == expected_order(flags, nb)
    for (f, b, a), (_, _, r) in zip(fo, fr):
        assert a.tobytes() == r.tobytes(), f"{key}: stage {f} bounce {b}"
    assert C.acc_of(orc, _lib.BUF_TRACE_ACCUMULATOR, w, h).tobytes() == C.acc_of(ref, _lib.BUF_TRACE_ACCUMULATOR, w, h).tobytes()
    orc.close()
    ref.close()


def test_debug_frames_buffer_too_small_is_an_error():
    import ctypes

    from polaris_b200._lib import Stats

    w, h = 32, 32
    sc = C.small_scene("c1", w, h)
    orc = C.oracle_for(sc, w, h)
    req = T.make_block_request(w, h, spp=1, num_bounces=2)
    seeds = T.splitmix_seeds(1, 3)
    frames = np.zeros((1, h, w, 4), np.uint8)
    infos = np.zeros(1, _lib.DEBUG_FRAME_DTYPE)
    got = ctypes.c_uint32(0)
    rc = orc._fn("trace_debug")(orc._h, ctypes.byref(req), seeds.ctypes.data, seeds.size, _lib.DEBUG_ALL_STAGES, frames.ctypes.data,
                                frames.nbytes, infos.ctypes.data, 1, ctypes.byref(got), ctypes.byref(Stats()))
    assert rc == _lib.ERR_INVALID_ARGUMENT and got.value == 1
    orc.close()
