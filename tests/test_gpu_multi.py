"""GPU tests of the multi-tracer side of the interface: MergeOutput across handles, devices and processes, the renderer's
worker-thread contract driven from plain C, and the literal converged-image bar of BASELINE.json's north_star.

  * renderer/default.go:106-196 -- one worker thread per tracer, every worker calls primary.MergeOutput(self) concurrently
    (tests/cabi/render_multi.c, pthreads over the C ABI).  Checked by REPLAYING the run on one handle: same block list,
    same seeds -> the primary's frame accumulator must be bit-identical.
  * tracer/opencl/tracer.go:279-286 + resources.go:108-124 -- the merge reads the peer DEVICE's accumulator: tracers on
    CUDA ordinals 0 and 1 (skipped on a one-GPU box; `gpurun --gpus 2` runs it, log under profiles/).
  * device/context.go:11-28 -- the shared context, across processes: pc_ipc_export / pc_ipc_open.
  * "converged images at high spp must agree with reference RMSE <= 1e-3 of mean luminance": CUDA against oracle/_ref (the
    reference's own kernels), shared seeds, stable compaction order on both sides.
"""
import os
import subprocess
import sys

import numpy as np
import pytest

from polaris_b200 import _lib
from polaris_b200 import tracer as T

from . import common as C

pytestmark = pytest.mark.gpu


def _cam_args(sc):
    return [repr(float(x)) for x in np.asarray(sc.camera.frustrum, np.float32).reshape(16)] + \
           [repr(float(x)) for x in np.asarray(sc.camera.position, np.float32)]


def _replay(sc, w, h, spp, rows_per_frame, ordinal=0):
    """The same block requests, frame by frame, tracer by tracer, SEQUENTIALLY: one handle traces every block, a second one
    only receives the merges (a tracer that traced two first-pass blocks of one frame would reset its own frame
    accumulator in between, like the reference's Trace does: tracer.go:208-213)."""
    cu, primary = C.cuda_for(sc, w, h, ordinal=ordinal), C.cuda_for(sc, w, h, ordinal=ordinal)
    for fr, rows in enumerate(rows_per_frame):
        y = 0
        for i, bh in enumerate(rows):
            req = T.make_block_request(w, h, block_y=y, block_h=int(bh), spp=spp, accumulated_samples=fr * spp)
            cu.trace(req, T.splitmix_seeds(7 + 100 * i + fr, spp * 6))
            primary.merge_output(cu, req)
            y += int(bh)
        primary.sync_framebuffer(T.make_block_request(w, h, spp=spp, accumulated_samples=fr * spp), want_pixels=fr + 1 == len(rows_per_frame))
    acc = primary.read_buffer(_lib.BUF_FRAME_ACCUMULATOR, w * h * 4, np.float32)
    rgba = primary.frame_buffer.copy()
    cu.close()
    primary.close()
    return acc, rgba


@pytest.mark.parametrize("tracers", [2, 5])
def test_c_host_worker_threads(tmp_path, tracers):
    """tests/cabi/render_multi.c: the renderer's goroutine structure as pthreads over the C ABI, concurrent MergeOutput into
    the primary, perfect scheduler between progressive frames.  Tracer i runs on device i % device_count."""
    from .test_cpu_abi import _build_c_client

    w, h, spp, frames = 128, 96, 4, 3
    sc = C.small_scene("c2", w, h)
    dump, out = str(tmp_path / "c2.plrscn"), str(tmp_path / "multi.bin")
    sc.save(dump)
    exe = _build_c_client(tmp_path, "render_multi")
    r = subprocess.run([exe, dump, str(w), str(h), str(spp), str(tracers), str(frames), out] + _cam_args(sc), capture_output=True, text=True)
    print(r.stdout)
    assert r.returncode == 0, r.stderr
    raw = open(out, "rb").read()
    assert raw[:8] == b"PCMULTI1"
    n, nf = np.frombuffer(raw, np.uint32, 2, 8)
    assert (n, nf) == (tracers, frames)
    rows = np.frombuffer(raw, np.uint32, n * nf, 16).reshape(nf, n)
    assert (rows.sum(axis=1) == h).all() and (rows >= 1).all()
    off = 16 + 4 * n * nf
    acc = np.frombuffer(raw, np.float32, w * h * 4, off)
    rgba = np.frombuffer(raw, np.uint8, w * h * 4, off + w * h * 16).reshape(h, w, 4)
    want_acc, want_rgba = _replay(sc, w, h, spp, rows)
    assert acc.tobytes() == want_acc.tobytes(), "concurrently merged frame differs from the sequential replay"
    assert np.array_equal(rgba, want_rgba)
    assert acc.reshape(h, w, 4)[..., :3].sum() > 0


def test_merge_across_devices():
    """pc_merge_output with the source tracer on ANOTHER GPU: k_merge on the primary loads the peer's rows over NVLink.
    Bit-identical to the same requests with both handles on one device."""
    if T.device_count() < 2:
        pytest.skip("needs two CUDA devices (run with gpurun --gpus 2)")
    w, h, spp = 160, 120, 3
    sc = C.small_scene("c2", w, h)
    blocks = [(0, 50), (50, 70)]
    frames = {}
    for name, ordinals in (("two devices", (0, 1)), ("one device", (0, 0))):
        trs = [C.cuda_for(sc, w, h, ordinal=o) for o in ordinals]
        for fr in range(2):  # two progressive frames: the second merge must not reset
            for i in (1, 0):  # the worker first: its merge must survive the primary's own first-pass Trace
                by, bh = blocks[i]
                r = T.make_block_request(w, h, block_y=by, block_h=bh, spp=spp, accumulated_samples=fr * spp)
                trs[i].trace(r, T.splitmix_seeds(30 + 10 * fr + i, spp * 6))
                trs[0].merge_output(trs[i], r)
            trs[0].sync_framebuffer(T.make_block_request(w, h, spp=spp, accumulated_samples=fr * spp))
        frames[name] = (trs[0].read_buffer(_lib.BUF_FRAME_ACCUMULATOR, w * h * 4, np.float32).tobytes(), trs[0].frame_buffer.tobytes())
        for t in trs:
            t.close()
    assert frames["two devices"] == frames["one device"]
    fa = np.frombuffer(frames["two devices"][0], np.float32).reshape(h, w, 4)
    assert fa[:50, :, :3].sum() > 0 and fa[50:, :, :3].sum() > 0


_IPC_CHILD = r"""
import sys
sys.path.insert(0, {root!r})
import numpy as np
from polaris_b200 import tracer as T
from tests import common as C
w, h, spp, by, bh, ordinal = {w}, {h}, {spp}, {by}, {bh}, {ordinal}
sc = C.small_scene("c2", w, h)
tr = C.cuda_for(sc, w, h, ordinal=ordinal)
handles = [tr.ipc_export(s) for s in (0, 1)]
for slot, cfg in ((0, 41), (1, 42)):
    req = T.make_block_request(w, h, block_y=by, block_h=bh, spp=spp)
    tr.trace(req, T.splitmix_seeds(cfg, spp * 6))
    tr.ipc_publish_rows(req, slot)
print("HANDLES", handles[0].hex(), handles[1].hex(), flush=True)
sys.stdin.readline()  # keep the allocation alive until the parent has merged
tr.close()
"""


def test_ipc_rows_between_processes():
    """One process per GPU: a worker PROCESS publishes its block rows into an exported buffer, the primary maps it with
    pc_ipc_open and merges with peer loads (pc_merge_rows on the mapped pointer) -- bit-identical to merging the same
    rows traced in-process."""
    w, h, spp, by, bh = 128, 96, 3, 32, 48
    ordinal = 1 if T.device_count() >= 2 else 0
    code = _IPC_CHILD.format(root=C.ROOT, w=w, h=h, spp=spp, by=by, bh=bh, ordinal=ordinal)
    child = subprocess.Popen([sys.executable, "-c", code], stdin=subprocess.PIPE, stdout=subprocess.PIPE, text=True, cwd=C.ROOT)
    try:
        line = ""
        for line in child.stdout:
            if line.startswith("HANDLES"):
                break
        assert line.startswith("HANDLES"), "worker process did not publish"
        handles = [bytes.fromhex(x) for x in line.split()[1:3]]
        sc = C.small_scene("c2", w, h)
        cu = C.cuda_for(sc, w, h)
        got = []
        for slot, cfg in ((0, 41), (1, 42)):
            ptr = cu.ipc_open(handles[slot])
            req = T.make_block_request(w, h, block_y=by, block_h=bh, spp=spp, accumulated_samples=spp)
            cu.merge_rows(ptr + 16 * w * by, True, req)
            cu.sync_framebuffer(T.make_block_request(w, h, spp=spp))
            got.append(cu.read_buffer(_lib.BUF_FRAME_ACCUMULATOR, w * h * 4, np.float32).copy())
            cu.ipc_close(ptr)
            # the same rows traced and merged in this process
            r2 = T.make_block_request(w, h, block_y=by, block_h=bh, spp=spp)
            cu.trace(r2, T.splitmix_seeds(cfg, spp * 6))
            cu.merge_output(cu, r2)
            cu.sync_framebuffer(T.make_block_request(w, h, spp=spp))
            want = cu.read_buffer(_lib.BUF_FRAME_ACCUMULATOR, w * h * 4, np.float32)
            assert got[-1].tobytes() == want.tobytes(), f"slot {slot}: rows merged through the IPC mapping differ"
            assert want.reshape(h, w, 4)[by:by + bh, :, :3].sum() > 0
        assert got[0].tobytes() != got[1].tobytes()  # two different passes went through the two slots
        cu.close()
    finally:
        try:
            child.stdin.write("\n")
            child.stdin.flush()
        except Exception:
            pass
        child.wait(timeout=60)


def test_converged_image_vs_reference_kernels():
    """north_star, literally: "converged images at high spp must agree with reference RMSE <= 1e-3 of mean luminance".
    Config 2's scene, shared seed list, CUDA against oracle/_ref = the reference's own kernels compiled for the CPU (the
    port when _ref is absent).  Both sides compact in stable order, so every sample follows the same paths and what is left
    is libm's last ulp in sin / cos / atan / acos / pow flipping a branch in a few samples."""
    try:
        from oracle import ref_binding
        ref = ref_binding.RefTracer() if ref_binding.available() else None
    except Exception:
        ref = None
    w = h = 128
    spp = 1024
    sc = C.small_scene("c2", w, h)
    ref = C.setup(ref, sc, w, h) if ref is not None else C.oracle_for(sc, w, h)
    cu = C.cuda_for(sc, w, h)
    seeds = T.splitmix_seeds(2, spp * 6)
    imgs = []
    for tr in (ref, cu):
        tr.trace(T.make_block_request(w, h, spp=spp), seeds)
        imgs.append(C.acc_of(tr, _lib.BUF_TRACE_ACCUMULATOR, w, h).astype(np.float64) / spp)
    lum = lambda x: 0.2126 * x[:, 0] + 0.7152 * x[:, 1] + 0.0722 * x[:, 2]  # noqa: E731
    lr, lg = lum(imgs[0]), lum(imgs[1])
    rmse = float(np.sqrt(np.mean((lg - lr) ** 2)))
    mean = float(lr.mean())
    worst = int(np.argmax(np.abs(lg - lr)))
    print(f"c2 {w}x{h} @ {spp} spp, shared seeds: mean luminance reference {mean:.6f} cuda {lg.mean():.6f}; "
          f"RMSE(lum) / mean = {rmse / mean:.3e}; pixels that differ at all: {int((lg != lr).sum())} / {lr.size}; "
          f"worst pixel {worst}: reference {lr[worst]:.6f} cuda {lg[worst]:.6f}")
    assert rmse / mean <= 1e-3
    cu.close()
