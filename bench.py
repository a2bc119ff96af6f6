#!/usr/bin/env python
"""bench.py -- throughput of the tracer hot path on B200 (driver contract).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
  torchrun --nproc-per-node N ... bench.py --gpus N ...          (one rank per GPU, N > 1)

N = 1  workload = BASELINE.json configs[1]: layered-material Cornell box, 1024x1024, 256 spp,
       5 bounces, RR after 3.  One step = one frame: Trace + MergeOutput + SyncFramebuffer.
N > 1  workload = configs[4]: the same scene at 3840x2160, rows split across the ranks by the
       restated `perfect` block scheduler; one step = one 64-spp pass: every rank traces its row
       block, the blocks are gathered to rank 0 over NCCL/NVLink and added, rank 0 tonemaps.
Metric = Mrays/s over all bounces (closest-hit + occlusion rays, counted on the device).

`value`   whole-job Mrays/s, scene resident in HBM before the timed region.
`e2e`     same metric through the public Tracer API with HOST scene buffers: every step uploads the
          scene + camera from host memory, traces, and reads the RGBA8 frame back to the host.
`roofline`  dominant kernel's algorithmic bytes per launch / its CUDA-event time per launch (a second
          pass of the same workload in PC_OPT_KERNEL_TIMERS mode, device counters on), against the
          measured HBM copy bandwidth of MEASURED_PEAKS.json.
`cpu_baseline`  the CPU implementation (oracle/_ref = the reference's own kernels compiled for the
          CPU when present, else the oracle port) on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

NUM_BOUNCES, MIN_RR, EXPOSURE = 5, 3, 1.2
HBM_FALLBACK_GBS = 6650.0  # /opt/skills/guides/B200_PROFILING.md fallback


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.rows = []
        self.proc = None
        self.gpu_index = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", str(self.gpu_index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) < 9:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return HBM_FALLBACK_GBS, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------------
# algorithmic bytes per launch of each kernel class: SURVEY §8(d) per-unit figures x the units the
# launch processed (device counters).  DESIGN.md §Roofline restates this table.
def algorithmic_bytes(st, frame_px, nb):
    q, o = st["query_rays"], st["occlusion_rays"]
    trav = 64 * st["nodes_tested"] + 80 * st["instances_entered"] + 48 * st["tris_tested"]
    # split traversal payload between the kernels by their ray share (counters are per trace call)
    share_q = q / max(1, q + o)
    prim_rays = frame_px * st["_spp"]
    ind_rays = q - prim_rays
    by = {
        "k_primary": 64 * prim_rays + 36 * prim_rays + trav * share_q * (prim_rays / max(1, q)),
        "k_query": (32 + 4 + 32) * ind_rays + trav * share_q * (ind_rays / max(1, q)),
        "k_occlusion": 32 * o + trav * (1 - share_q) + (4 + 16 + 32) * st["unoccluded"],
        "k_shade": (36 * q + (64 + 124 + 64 + 80 + 120 + 32) * st["shaded_hits"] + 48 * st["occlusion_emitted"]
                    + 32 * st["indirect_emitted"] + 96 * st["missed_query_rays"] * (1 if st["_background"] else 0)),
    }
    return by


# ------------------------------------------------------------------------------------------------
def build_scene(name, w, h):
    from polaris_b200 import scenes

    t0 = time.time()
    sc, _, _, _ = scenes.build(name, w, h)
    log(f"[bench] scene {name} {w}x{h}: {sc.num_triangles} triangles, {len(sc.bvh_nodes)} BVH nodes, "
        f"{sc.nbytes() / 1e6:.2f} MB, compiled in {time.time() - t0:.1f}s")
    return sc


def cpu_trace_sample(sc, w, h, spp, config_number, rows=None):
    """Bounded CPU run of the same workload: returns (Mrays/s, cores, kind, seconds, rays)."""
    from polaris_b200 import tracer as T

    kind = "port"
    tr = None
    try:
        from oracle import ref_binding

        if ref_binding.available():
            tr = ref_binding.RefTracer()
            kind = "reference"
    except Exception as e:  # pragma: no cover
        log(f"[bench] oracle/_ref unavailable ({e}); using the oracle port")
    if tr is None:
        from oracle.binding import OracleTracer

        tr = OracleTracer()
    tr.init()
    tr.update_state(T.SYNCHRONOUS, T.FRAME_DIMENSIONS, (w, h))
    tr.update_state(T.SYNCHRONOUS, T.SCENE_DATA, sc)
    tr.update_state(T.SYNCHRONOUS, T.CAMERA_DATA, sc.camera)
    bh = min(rows or h, h)
    by = (h - bh) // 2  # a bounded sample takes the MIDDLE rows of the frame: the top rows of the terrain frame are sky only
    req = T.make_block_request(w, h, block_y=by, block_h=bh, spp=spp, num_bounces=NUM_BOUNCES, min_bounces_for_rr=MIN_RR, exposure=EXPOSURE)
    seeds = T.splitmix_seeds(config_number, spp * (1 + NUM_BOUNCES))
    t0 = time.perf_counter()
    tr.trace(req, seeds)
    tr.merge_output(tr, req)
    tr.sync_framebuffer(T.make_block_request(w, h, spp=spp, exposure=EXPOSURE))
    dt = time.perf_counter() - t0
    d = tr.stats().device
    rays = d["query_rays"] + d["occlusion_rays"]
    cores = tr.threads() if hasattr(tr, "threads") else (os.cpu_count() or 1)
    tr.close()
    return rays / dt / 1e6, cores, kind, dt, rays


def opencl_reference_sample(sc, w, h, spp, config_number, rows=None):
    """The reference's own OpenCL program + launch discipline on this box's OpenCL device (on the GPU box: the same
    B200 through NVIDIA's OpenCL driver, oracle/cl_device.py), bounded sample of the same workload.  A reported
    baseline next to `cpu_baseline`; returns a dict for the JSON line."""
    try:
        from oracle import cl_device
        from polaris_b200 import tracer as T

        if not cl_device.available():
            return {"unavailable": "no OpenCL device or no embedded reference program (oracle/_ref) on this box"}
        tr = cl_device.ClDeviceTracer()
        tr.init()
        tr.update_state(T.SYNCHRONOUS, T.FRAME_DIMENSIONS, (w, h))
        tr.update_state(T.SYNCHRONOUS, T.SCENE_DATA, sc)
        tr.update_state(T.SYNCHRONOUS, T.CAMERA_DATA, sc.camera)
        bh = min(rows or h, h)
        by = (h - bh) // 2  # middle rows, like the CPU sample
        seeds = T.splitmix_seeds(config_number, spp * (1 + NUM_BOUNCES))
        best = None
        for _ in range(2):  # first pass warms the driver's lazily built kernels
            req = T.make_block_request(w, h, block_y=by, block_h=bh, spp=spp, num_bounces=NUM_BOUNCES, min_bounces_for_rr=MIN_RR, exposure=EXPOSURE)
            t0 = time.perf_counter()
            tr.trace(req, seeds)
            tr.merge_output(tr, req)
            tr.sync_framebuffer(T.make_block_request(w, h, spp=spp, exposure=EXPOSURE), want_pixels=False)
            dt = time.perf_counter() - t0
            d = tr.stats().device
            rays = d["query_rays"] + d["occlusion_rays"]
            if best is None or rays / dt > best[0]:
                best = (rays / dt, dt, rays, d["kernel_launches"])
        desc = tr.dev.describe()
        tr.close()
        return {"value": best[0] / 1e6, "unit": "Mrays/s", "kind": "reference OpenCL kernels + launch discipline (clFinish per launch), same GPU",
                "device": desc["device"], "platform": desc["platform_version"], "launches": int(best[3]),
                "sample": f"rows {by}..{by + bh} of the {w}x{h} frame, {spp} spp ({best[2]} rays in {best[1]:.3f}s, best of 2)"}
    except Exception as e:  # a baseline must never take the bench down
        return {"unavailable": f"{type(e).__name__}: {str(e)[:200]}"}


# ------------------------------------------------------------------------------------------------
def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path on the host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    n = args.gpus
    if n == 1:
        name, w, h, spp_full, cfgno = "c2_cornell", 1024, 1024, 256, 2
    else:
        name, w, h, spp_full, cfgno = "c5_cornell_4k", 3840, 2160, 1024, 5
    sc = build_scene(name, w, h)
    try:  # torchrun exports OMP_NUM_THREADS=1 to its ranks; the reference arm uses every host core
        from oracle import ref_binding
        if ref_binding.available():
            ref_binding.RefTracer.set_threads(len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1))
    except Exception as e:  # pragma: no cover
        log(f"[bench:reference] could not raise the OpenMP thread count: {e}")
    # bounded sample per step so that warmup+steps finish within a few minutes
    spp = 2 if n == 1 else 1
    rows = h if n == 1 else 540
    vals, times = [], []
    for i in range(args.warmup + args.steps):
        v, cores, kind, dt, rays = cpu_trace_sample(sc, w, h, spp, cfgno, rows)
        log(f"[bench:reference] step {i}: {v:.2f} Mrays/s ({dt:.2f}s, {rays} rays, kind={kind}, cores={cores})")
        if i >= args.warmup:
            vals.append(v)
            times.append(dt)
    value = float(np.mean(vals))
    sample = f"the middle {rows} rows of the {w}x{h} frame, {spp} spp of {spp_full} per step"
    line = {
        "impl": "reference", "metric": "Mrays/s (all bounces)", "value": value, "unit": "Mrays/s", "n_gpus": n,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": float(np.mean(times) * 1e3), "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(n, w, h, spp_full),
        "cpu_baseline": {"value": value, "unit": "Mrays/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    if not args.no_opencl:
        line["reference_on_gpu"] = opencl_reference_sample(sc, w, h, 8 if n == 1 else 2, cfgno, rows)
        log(f"[bench:reference] reference OpenCL path on this box's GPU: {line['reference_on_gpu']}")
    emit(line)
    return 0


WORKLOADS = {
    "c1": "configs[0]: procedurally generated diffuse sphere, 512x512, 128 spp",
    "c2": "configs[1]: synthetic Cornell box, layered diffuse/conductor/dielectric materials, 1024x1024, 256 spp",
    "c3": "configs[2]: mesh-instancing stress, 1,000 instances of a 100k-triangle procedural mesh, 1920x1080, 64 spp",
    "c4": "configs[3]: 10M-triangle displaced terrain, roughDielectric + dispersion, synthetic HDR textures, 3840x2160, 64 spp",
    "c5": "configs[4] on ONE GPU: 3840x2160 layered Cornell box, one 64-spp pass of the 1024-spp job",
}


def workload_config(n, w, h, spp, config="c2", sc=None):
    if n == 1:
        wl = WORKLOADS[config]
        if sc is not None and config != "c2":
            return {"workload": wl, "frame": [w, h], "spp": spp, "num_bounces": NUM_BOUNCES, "min_bounces_for_rr": MIN_RR,
                    "triangles": int(sc.num_triangles), "instances": int(len(sc.mesh_instances)), "scene_bytes": int(sc.nbytes()),
                    "l2": "per-step ray/path state (220 B/px x frame, re-written every bounce) exceeds the 126 MB L2"}
    else:
        wl = f"configs[4]: 3840x2160 layered Cornell box, 1024 spp in 64-spp passes, rows split over {n} GPUs by the perfect scheduler"
    return {"workload": wl, "frame": [w, h], "spp": spp, "num_bounces": NUM_BOUNCES, "min_bounces_for_rr": MIN_RR,
            "l2": "per-step ray/path state (220 B/px x frame, re-written every bounce) exceeds the 126 MB L2; scene data is L2 resident by design"}


def measured_profile(config, kernel):
    """What ncu measured for this kernel class on this config (profiles/traffic.json, written by tools/summarize_ncu.py from
    `ncu --set full` captures kept under profiles/): DRAM bytes per launch, lanes active, IPC, cache hit rates."""
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        t = json.load(open(tp))
    except Exception:
        return None
    ent = t.get(config, {}) if isinstance(t.get(config), dict) else {}
    return ent.get(kernel)


def roofline_pass(tr, sc, w, h, block_y, block_h, prof_spp, seeds, config):
    """A second pass of the same workload in PC_OPT_KERNEL_TIMERS mode (CUDA events around every launch on the tracer's
    stream, one sample chain) with device counters on.  One untimed warm-up trace, then prof_spp >= 64 samples; a kernel
    class's time is the MEDIAN over the samples of (sum of that class's launches within the sample), so that a cold first
    sample or a clock ramp does not move it.  achieved = algorithmic bytes of the dominant class / that time."""
    from polaris_b200 import _lib
    from polaris_b200 import tracer as T

    tr.set_option(_lib.OPT_KERNEL_TIMERS, 1)
    tr.set_option(_lib.OPT_COUNTERS, 1)
    per = 1 + NUM_BOUNCES
    mk = lambda n: T.make_block_request(w, h, block_y=block_y, block_h=block_h, spp=n, num_bounces=NUM_BOUNCES,  # noqa: E731
                                        min_bounces_for_rr=MIN_RR, exposure=EXPOSURE)
    tr.trace(mk(2), seeds[: 2 * per])  # warm-up: clocks, caches, lazily created events
    tr.trace(mk(prof_spp), seeds[: prof_spp * per])
    st = tr.stats().device
    cls, us = tr.kernel_timings()
    tr.set_option(_lib.OPT_KERNEL_TIMERS, 0)
    tr.set_option(_lib.OPT_COUNTERS, 0)
    st["_spp"], st["_background"] = prof_spp, sc.scene_diffuse_mat_index != -1
    by = algorithmic_bytes(st, w * block_h, NUM_BOUNCES)
    names = _lib.KERNEL_CLASS_NAMES
    # the launch sequence repeats per BATCH (a set of launches carries several samples, PC_OPT_SAMPLE_SLOTS): the period is
    # the distance between two k_begin_sample launches
    begins = np.nonzero(cls == _lib.K_BEGIN_SAMPLE)[0]
    lps = int(begins[1] - begins[0]) if len(begins) > 1 else len(cls)
    assert len(cls) % lps == 0, (len(cls), lps)
    n_batches = len(cls) // lps
    cls2, us2 = cls.reshape(n_batches, lps), us.reshape(n_batches, lps).astype(np.float64)
    assert (cls2 == cls2[0]).all()
    med_us, mean_us, counts = {}, {}, {}
    for ci, n in enumerate(names):
        m = cls2[0] == ci
        if m.any():
            per_sample = us2[:, m].sum(axis=1)
            med_us[n], mean_us[n], counts[n] = float(np.median(per_sample)), float(per_sample.mean()), int(m.sum())
    if "k_trace" in med_us:
        # PC_OPT_FUSE_TRACE: the occlusion test and the next bounce's query share a launch, and the device counters are
        # per trace call, not per launch: report the traversal classes as one ("k_trace" = the fused launches + the last
        # bounce's stand-alone occlusion launch)
        by["k_trace"] = by.pop("k_query") + by.pop("k_occlusion")
        for d in (med_us, mean_us, counts):
            d["k_trace"] = d["k_trace"] + d.pop("k_query", 0) + d.pop("k_occlusion", 0)
    classes = [k for k in ("k_primary", "k_shade", "k_occlusion", "k_query", "k_trace") if k in med_us]
    total = sum(med_us.values())
    dom = max(classes, key=lambda k: med_us[k])
    peak, peak_src = measured_hbm_peak()
    kern = {k: {"launches_per_batch": counts[k], "avg_us": med_us[k] / counts[k], "mean_avg_us": mean_us[k] / counts[k],
                "share": med_us[k] / total, "alg_GBps": by[k] / n_batches / (med_us[k] * 1e3)} for k in classes}
    log("[bench] kernel classes (median over %d batches of %d samples): %s" % (n_batches, prof_spp // n_batches, json.dumps(kern)))
    achieved = by[dom] / n_batches / (med_us[dom] * 1e3)  # bytes per ns == GB/s
    avg_launch_us = med_us[dom] / counts[dom]
    rays = st["query_rays"] + st["occlusion_rays"]
    prof = measured_profile(config, dom)
    measured, traffic = None, None
    if prof:
        traffic = prof.get("dram_bytes")
        measured = dict(prof)
        if traffic:
            measured["dram_GBps"] = traffic / (avg_launch_us * 1e3)
            measured["dram_frac"] = measured["dram_GBps"] / peak
        if prof.get("l2_bytes"):
            measured["l2_GBps"] = prof["l2_bytes"] / (avg_launch_us * 1e3)
        if prof.get("ipc"):
            # the roofline these kernels actually sit under: warp instructions issued per SM cycle of the 4 an SM can issue, and
            # how many of a warp instruction's 32 lanes do useful work
            measured["issue_slot_frac"] = prof["ipc"] / 4.0
            if prof.get("lanes_active"):
                measured["simd_frac"] = prof["lanes_active"] / 32.0
        measured["note"] = ("ncu --set full of this kernel class on this config (profiles/); bytes are per launch, rates use the "
                            "launch time measured HERE.  bound = what the counters show limits the kernel")
    return {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
            "traffic": traffic, "peak_source": peak_src, "bytes_per_launch": by[dom] / n_batches / counts[dom],
            "avg_launch_us": avg_launch_us, "samples_per_launch": prof_spp // n_batches,
            "timing": f"median over {n_batches} batches ({prof_spp} samples) after a warm-up trace, CUDA events per launch",
            "kernels": kern, "measured": measured,
            "rays_per_path": rays / (w * block_h * prof_spp), "nodes_per_ray": st["nodes_tested"] / max(1, rays),
            "tris_per_ray": st["tris_tested"] / max(1, rays)}


# ------------------------------------------------------------------------------------------------
def run_cuda_single(args):
    import torch  # device plumbing only (event/synchronize helpers are not needed; kept for parity with N>1)

    from polaris_b200 import _lib
    from polaris_b200 import tracer as T

    assert torch.cuda.is_available(), "bench.py needs a CUDA device; there is no CPU fallback"
    from polaris_b200 import scenes

    cfg_name = {"c1": "c1_sphere", "c2": "c2_cornell", "c3": "c3_instancing", "c4": "c4_terrain", "c5": "c5_cornell_4k"}[args.config]
    w, h, cfg_spp = scenes.CONFIGS[cfg_name]
    spp, cfgno = args.spp or cfg_spp, int(args.config[1])
    if args.config == "c5":
        spp = args.spp or 64  # one 64-spp pass of the 1024-spp job (what a rank does per step at N > 1)
    sc = build_scene(cfg_name, w, h)
    seeds = T.splitmix_seeds(cfgno, spp * (1 + NUM_BOUNCES))
    tr = T.CudaTracer("cuda:0", 0)
    tr.init()
    _lib.pin_scene(sc)  # e2e copies the scene from page-locked host memory (pc_host_register)
    tr.update_state(T.SYNCHRONOUS, T.FRAME_DIMENSIONS, (w, h))
    tr.update_state(T.SYNCHRONOUS, T.SCENE_DATA, sc)
    tr.update_state(T.SYNCHRONOUS, T.CAMERA_DATA, sc.camera)
    if args.chains:
        tr.set_option(_lib.OPT_SAMPLE_CHAINS, args.chains)
    for kv in args.opt:
        k, v = kv.split("=")
        tr.set_option(getattr(_lib, "OPT_" + k.upper()), int(v))

    def frame(e2e=False):
        if e2e:  # host scene buffers -> device, every step
            tr.update_state(T.SYNCHRONOUS, T.SCENE_DATA, sc)
            tr.update_state(T.SYNCHRONOUS, T.CAMERA_DATA, sc.camera)
        req = T.make_block_request(w, h, spp=spp, num_bounces=NUM_BOUNCES, min_bounces_for_rr=MIN_RR, exposure=EXPOSURE)
        tr.trace(req, seeds)
        tr.merge_output(tr, req)
        tr.sync_framebuffer(T.make_block_request(w, h, spp=spp, exposure=EXPOSURE), want_pixels=e2e)
        d = tr.stats().device
        return d["query_rays"] + d["occlusion_rays"], d["kernel_launches"] + 2, d["device_time_ns"]

    for i in range(args.warmup):
        t0 = time.perf_counter()
        rays, _, _ = frame()
        log(f"[bench] warmup {i}: {rays / (time.perf_counter() - t0) / 1e6:.1f} Mrays/s")
    clocks = ClockSampler(0)
    clocks.start()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    tot_rays = tot_launch = tot_dev_ns = 0
    for _ in range(args.steps):
        rays, launches, dev_ns = frame()
        tot_rays += rays
        tot_launch += launches
        tot_dev_ns += dev_ns
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    clk = clocks.stop()
    value = tot_rays / dt / 1e6
    log(f"[bench] timed: {args.steps} steps in {dt:.3f}s -> {value:.1f} Mrays/s ({tot_dev_ns / 1e9:.3f}s device time in pc_trace)")

    # ---- e2e: through the public API with host buffers, H2D scene + D2H frame inside the timed region
    frame(e2e=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    e_rays = 0
    for _ in range(args.steps):
        r, _, _ = frame(e2e=True)
        e_rays += r
    torch.cuda.synchronize()
    e_dt = time.perf_counter() - t0
    h2d = sc.nbytes() + seeds.nbytes + 76
    d2h = w * h * 4

    # ---- roofline pass: same workload, per-kernel CUDA events + device counters
    roofline = roofline_pass(tr, sc, w, h, 0, h, min(spp, 64), seeds, args.config)
    tr.close()

    # ---- cold start: raw triangles -> compiled scene ON THE DEVICE (pc_compile_geometry: the reference's SAH build as CUDA
    # kernels) -> upload -> first frame, on a fresh tracer.  Not part of `value` / `e2e`.
    cold = None
    if args.cold_scene or (cfg_name, ()) in scenes._raw_cache:
        try:
            from polaris_b200.scene import compile_scene
            raw = scenes.raw_scene(cfg_name)
            compile_scene(raw, aspect=np.float32(w) / np.float32(h), builder="cuda")  # CUDA context + first-use allocations
            t0 = time.perf_counter()
            sc2 = compile_scene(raw, aspect=np.float32(w) / np.float32(h), builder="cuda")
            t1 = time.perf_counter()
            tr2 = T.CudaTracer("cuda:0", 0)
            tr2.init()
            tr2.update_state(T.SYNCHRONOUS, T.FRAME_DIMENSIONS, (w, h))
            tr2.update_state(T.SYNCHRONOUS, T.SCENE_DATA, sc2)
            tr2.update_state(T.SYNCHRONOUS, T.CAMERA_DATA, sc2.camera)
            req = T.make_block_request(w, h, spp=spp, num_bounces=NUM_BOUNCES, min_bounces_for_rr=MIN_RR, exposure=EXPOSURE)
            tr2.trace(req, seeds)
            tr2.merge_output(tr2, req)
            tr2.sync_framebuffer(T.make_block_request(w, h, spp=spp, exposure=EXPOSURE), want_pixels=True)
            t2 = time.perf_counter()
            tr2.close()
            same = all(getattr(sc, k).tobytes() == getattr(sc2, k).tobytes() for k in sc._SECTIONS)
            cold = {"compile_scene_s": t1 - t0, "upload_and_first_frame_s": t2 - t1, "total_s": t2 - t0, "builder": "cuda (pc_compile_geometry)",
                    "native": sc2.compile_timing, "identical_to_host_build": bool(same),
                    "note": "second device compile of this process (the first one pays CUDA context creation); first frame = the full-spp frame"}
            log(f"[bench] cold start: {cold}")
        except Exception as e:  # never take the bench down
            cold = {"unavailable": f"{type(e).__name__}: {str(e)[:200]}"}

    # ---- CPU baseline on a bounded sample (rank 0, N=1 only)
    cpu = None
    if not args.no_cpu:
        cpu_rows = h if w * h <= 1024 * 1024 else max(16, (1024 * 1024) // w)  # bound the sample on the big frames
        v, cores, kind, cdt, crays = cpu_trace_sample(sc, w, h, args.cpu_spp, cfgno, cpu_rows)
        cpu = {"value": v, "unit": "Mrays/s", "cores": cores, "kind": kind,
               "sample": f"the middle {cpu_rows} rows of the {w}x{h} frame, {args.cpu_spp} spp of {spp} ({crays} rays in {cdt:.1f}s)"}
        log(f"[bench] cpu baseline ({kind}, {cores} cores): {v:.2f} Mrays/s")
    ocl = None
    if not args.no_cpu and not args.no_opencl:
        ocl_rows = h if w * h <= 1024 * 1024 else max(16, (2048 * 1024) // w)
        ocl = opencl_reference_sample(sc, w, h, 8, cfgno, ocl_rows)
        log(f"[bench] reference OpenCL path on this GPU: {ocl}")

    line = {
        "metric": "Mrays/s (all bounces)", "value": value, "unit": "Mrays/s", "n_gpus": 1, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "strong",
        "device_ms_per_step": tot_dev_ns / args.steps / 1e6,  # CUDA events around pc_trace on the tracer's stream (the rest of a step is merge + tonemap)
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload_config(1, w, h, spp, args.config, sc),
        "spp_mpix_per_s": w * h * spp * args.steps / dt / 1e6, "gpu_launches": int(tot_launch), "clocks": clk,
        "e2e": {"value": e_rays / e_dt / 1e6, "unit": "Mrays/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h)},
        "roofline": roofline, "cpu_baseline": cpu, "reference_on_gpu": ocl, "cold_start": cold,
    }
    if ocl and ocl.get("value"):
        # the same-hardware baseline: the reference's own OpenCL program and launch discipline on this very GPU
        line["vs_reference_on_gpu"] = {"value": value / ocl["value"], "e2e": line["e2e"]["value"] / ocl["value"]}
    emit(line)
    return 0


# ------------------------------------------------------------------------------------------------
class _DevPtr:
    """Expose a raw device pointer to torch through __cuda_array_interface__ (zero copy)."""

    def __init__(self, ptr, nfloats):
        self.__cuda_array_interface__ = {"shape": (nfloats,), "typestr": "<f4", "data": (ptr, False), "version": 3, "strides": None}


def run_cuda_multi(args):
    import torch
    import torch.distributed as dist

    from polaris_b200 import tracer as T
    from polaris_b200.gather import IpcRowExchange, RowGather, StatsExchange
    from polaris_b200.scheduler import PerfectScheduler, StaticSpeed, assign_blocks_based_on_speed

    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    w, h, pass_spp, cfgno = 3840, 2160, args.spp or 64, 5
    frame_passes = max(1, 1024 // pass_spp)  # the 1024-spp job of configs[4]: the frame is tonemapped when it is complete
    sc = build_scene("c5_cornell_4k", w, h) if rank == 0 else None
    objs = [sc]
    dist.broadcast_object_list(objs, src=0)  # every GPU holds the full scene (default.go:70-72)
    sc = objs[0]
    tr = T.CudaTracer(f"cuda:{local}", local)
    tr.init()
    from polaris_b200 import _lib
    _lib.pin_scene(sc)  # e2e copies the scene from page-locked host memory (pc_host_register)
    if args.chains:
        tr.set_option(_lib.OPT_SAMPLE_CHAINS, args.chains)
    tr.update_state(T.SYNCHRONOUS, T.FRAME_DIMENSIONS, (w, h))
    tr.update_state(T.SYNCHRONOUS, T.SCENE_DATA, sc)
    tr.update_state(T.SYNCHRONOUS, T.CAMERA_DATA, sc.camera)
    sched = PerfectScheduler()
    speeds = [StaticSpeed(tr.speed()) for _ in range(world)]
    seed_cache = {}

    def pass_seeds(i):  # every pass of every rank draws its own seeds: 16 passes x 64 spp ARE 1024 different samples
        k = i % 64
        if k not in seed_cache:
            seed_cache[k] = T.splitmix_seeds(cfgno + 100 * rank + 10000 * k, pass_spp * (1 + NUM_BOUNCES))
        return seed_cache[k]

    for k in range(min(64, args.warmup + 2 * args.steps + 4)):
        pass_seeds(k)
    # ---- the exchange step.  Default: CUDA IPC mappings + peer loads in k_merge (polaris_b200/gather.py); --exchange nccl:
    # grouped send/recv of the rows.  Either way it is split phase: pass i's rows are added on rank 0 while pass i+1 is
    # traced, and the only collective of the data path is the scheduler's (BlockH, RenderTime) all-gather, which doubles as
    # the hand-shake: stats(i) is posted by a rank after it published pass i and, on rank 0, after it merged pass i-1.
    use_ipc = args.exchange == "ipc"
    xch = IpcRowExchange(tr, rank, world, w) if use_ipc else None
    recv_bufs = [torch.empty(w * h * 4, dtype=torch.float32, device="cuda") if (rank == 0 and not use_ipc) else None for _ in range(2)]
    state = {"i": 0, "acc": 0}
    pending_stats = {}    # pass index -> StatsExchange
    stats_done = {}       # pass index -> result, kept until the scheduler has used it
    pending_merge = []    # rank 0, at most one: (pass index, rows, acc_samples before that pass, e2e, RowGather | None)
    totals = {"rays": 0.0, "launches": 0.0, "device_ns": 0.0, "fed": -1}

    def wait_stats(i):
        if i in pending_stats:
            res = pending_stats.pop(i).result()
            stats_done[i] = res
            totals["rays"] += sum(x[2] for x in res)
            totals["launches"] += sum(x[3] for x in res)
            totals["device_ns"] += max(x[4] for x in res)  # CUDA-event time of pc_trace, max over ranks
            if rank == 0 and args.verbose:
                log(f"[bench] pass {i}: rows {[int(x[0]) for x in res]} trace ms {[round(x[1] * 1e3, 1) for x in res]}")

    def feed_scheduler(upto):  # Stats() of every tracer -> the perfect scheduler (scheduler.go:58-72), identical on every rank
        for i in sorted(k for k in stats_done if k <= upto):
            res = stats_done.pop(i)
            if i > totals["fed"]:
                for r in range(world):
                    speeds[r].set_stats(int(res[r][0]), float(res[r][1]))
                totals["fed"] = i

    def finish_merge(last=False):
        if not pending_merge:
            return
        i, rows, acc0, e2e, rg = pending_merge.pop(0)
        wait_stats(i)  # every rank has published pass i
        if rank == 0:
            def mk(r, y):
                return T.make_block_request(w, h, block_y=y, block_h=int(rows[r]), spp=pass_spp, accumulated_samples=acc0 + pass_spp)
            if use_ipc:
                xch.merge(rows, i, mk)
            else:
                blocks = rg.finish()
                torch.cuda.current_stream().synchronize()  # the received rows are complete before k_merge reads them on the frame stream
                y = int(rows[0])
                for r in range(1, world):
                    tr.merge_rows(blocks[r].data_ptr(), True, mk(r, y))
                    y += int(rows[r])
            if e2e or last or (i + 1) % frame_passes == 0:  # SyncFramebuffer once per frame (default.go:159-161); e2e reads every pass
                tr.sync_framebuffer(T.make_block_request(w, h, spp=pass_spp, exposure=EXPOSURE, accumulated_samples=acc0), want_pixels=e2e)
            else:
                tr.wait_for_kernels()  # the peers' slots of pass i are free again once this returns (before stats(i+1) is posted)
        elif rg is not None:
            rg.finish()

    def drain():
        finish_merge(last=True)
        for i in sorted(pending_stats):
            wait_stats(i)
        feed_scheduler(state["i"])

    def step(e2e=False):
        i, acc0 = state["i"], state["acc"]
        feed_scheduler(i - 2)  # timings that every rank is known to hold (they were waited for during pass i-1)
        if totals["fed"] < 0 and not totals.get("primed"):  # no Stats() yet (passes 0 and 1): the naive split by Speed(), what the perfect scheduler starts with too
            rows = list(assign_blocks_based_on_speed(speeds, h))
            sched.block_assignment = list(rows)
        else:
            rows = list(sched.schedule(speeds, h))
        by = int(sum(rows[:rank]))
        req = T.make_block_request(w, h, block_y=by, block_h=int(rows[rank]), spp=pass_spp, num_bounces=NUM_BOUNCES,
                                   min_bounces_for_rr=MIN_RR, exposure=EXPOSURE, accumulated_samples=acc0)
        if e2e:
            tr.update_state(T.SYNCHRONOUS, T.SCENE_DATA, sc)
            tr.update_state(T.SYNCHRONOUS, T.CAMERA_DATA, sc.camera)
        t0 = time.perf_counter()
        tr.trace(req, pass_seeds(i))
        t_trace = time.perf_counter() - t0
        d = tr.stats().device
        if rank == 0:
            tr.merge_output(tr, req)  # the primary's own rows go straight into its frame accumulator
        finish_merge()   # pass i-1: arrived / published while this pass was traced (waits for stats(i-1))
        wait_stats(i - 1)  # ranks != 0: rank 0 posted stats(i-1) after it merged pass i-2 -> slot i % 2 is free
        rg = None
        if use_ipc:
            xch.publish(req, i)
        else:
            ptr, nbytes = tr.trace_rows(req)
            mine = torch.as_tensor(_DevPtr(ptr, nbytes // 4), device="cuda")
            rg = RowGather(rows, w, rank, world).start(mine, recv_bufs[i % 2])
            if rank != 0:
                torch.cuda.current_stream().synchronize()  # the snapshot is taken before the next Trace clears the accumulator
        pending_merge.append((i, rows, acc0, e2e, rg))
        # Stats().RenderTime for the scheduler = the device time of the trace (CUDA events): host-side one-offs (state growing
        # with the block, graph re-instantiation) must not look like a slow GPU
        t_sched = d["device_time_ns"] * 1e-9 if d["device_time_ns"] else t_trace
        pending_stats[i] = StatsExchange([rows[rank], t_sched, d["query_rays"] + d["occlusion_rays"], d["kernel_launches"], d["device_time_ns"]], world, "cuda")
        state["i"], state["acc"] = i + 1, acc0 + pass_spp

    # the SAME workload on ONE of these GPUs (rank 0 traces the whole frame, the others wait): the N = 1 default of this
    # script is configs[1], a different frame, so the 1-GPU point of the configs[4] scaling curve is measured here
    one_gpu = None
    if not args.no_single:
        if rank == 0:
            vals = []
            for i in range(4):
                req1 = T.make_block_request(w, h, spp=pass_spp, num_bounces=NUM_BOUNCES, min_bounces_for_rr=MIN_RR, exposure=EXPOSURE)
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                tr.trace(req1, pass_seeds(i))
                tr.merge_output(tr, req1)
                tr.sync_framebuffer(T.make_block_request(w, h, spp=pass_spp, exposure=EXPOSURE), want_pixels=False)
                torch.cuda.synchronize()
                d1 = tr.stats().device
                if i:
                    vals.append((d1["query_rays"] + d1["occlusion_rays"]) / (time.perf_counter() - t0) / 1e6)
            one_gpu = {"value": float(np.mean(vals)), "unit": "Mrays/s", "passes": len(vals),
                       "note": "one 64-spp pass of the whole frame on rank 0's GPU alone, same scene / kernels, measured in this run"}
            log(f"[bench] the same workload on one GPU: {one_gpu['value']:.1f} Mrays/s")
        dist.barrier()
    # The perfect scheduler converges over a few frames (scheduler.go:50-80: rows ~ BlockH / RenderTime of the previous
    # frame, and rows near the top / bottom of this frame are cheaper than the middle).  In the interactive renderer that
    # happens once, at start-up; here a few cheap low-spp passes with a blocking stats exchange do it before the warm-up, so
    # that neither warm-up nor timed passes start from the naive equal split.  They also size every rank's per-block state.
    from polaris_b200.gather import exchange_stats
    prime_spp = max(16, pass_spp // 4)
    for k in range(8):
        rows = list(assign_blocks_based_on_speed(speeds, h)) if k == 0 else list(sched.schedule(speeds, h))
        if k == 0:
            sched.block_assignment = list(rows)
        req = T.make_block_request(w, h, block_y=int(sum(rows[:rank])), block_h=int(rows[rank]), spp=prime_spp, num_bounces=NUM_BOUNCES,
                                   min_bounces_for_rr=MIN_RR, exposure=EXPOSURE, accumulated_samples=pass_spp)
        tr.trace(req, pass_seeds(0)[: prime_spp * (1 + NUM_BOUNCES)])
        d = tr.stats().device
        res = exchange_stats(rows[rank], d["device_time_ns"] * 1e-9, rank, world, "cuda")
        for r in range(world):
            speeds[r].set_stats(int(res[r][0]), float(res[r][1]))
        if rank == 0 and args.verbose:
            log(f"[bench] priming {k}: rows {[int(x[0]) for x in res]} device ms {[round(x[1] * 1e3, 2) for x in res]}")
    totals["primed"] = True
    for i in range(args.warmup):
        step()
    drain()
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    dist.barrier()
    torch.cuda.synchronize()
    totals["rays"] = totals["launches"] = totals["device_ns"] = 0.0
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    drain()  # the last pass's rows are added and the frame is tonemapped inside the timed region
    dist.barrier()
    torch.cuda.synchronize()
    rays, launches, device_ns = totals["rays"], totals["launches"], totals["device_ns"]
    dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
    dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    dt = float(dt.item())
    clk = clocks.stop() if rank == 0 else None
    # e2e: scene + camera re-sent from host memory on every rank and the RGBA8 frame read back on rank 0, every pass
    step(e2e=True)
    drain()
    dist.barrier()
    torch.cuda.synchronize()
    totals["rays"] = 0.0
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step(e2e=True)
    drain()
    dist.barrier()
    torch.cuda.synchronize()
    e_rays = totals["rays"]
    e_dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
    dist.all_reduce(e_dt, op=dist.ReduceOp.MAX)
    e_dt = float(e_dt.item())
    roofline = None
    if rank == 0:  # rank 0's own row block of the last assignment, per-kernel CUDA events (the other ranks are done)
        rows_last = [int(sp.block_h) for sp in speeds]
        roofline = roofline_pass(tr, sc, w, h, 0, max(1, rows_last[0]), min(pass_spp, 64), pass_seeds(0), "c5")
        roofline["note"] = f"rank 0's block ({rows_last[0]} rows of {h}) of the last row assignment"
    if rank == 0:
        seeds_bytes = pass_seeds(0).nbytes
        line = {
            "metric": "Mrays/s (all bounces)", "value": rays / dt / 1e6, "unit": "Mrays/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "strong",
            "device_ms_per_step": device_ns / args.steps / 1e6,  # CUDA-event time of pc_trace, max over ranks, per pass
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload_config(world, w, h, 1024),
            "spp_mpix_per_s": w * h * pass_spp * args.steps / dt / 1e6, "gpu_launches": int(launches) + 2 * args.steps * world,
            "clocks": clk, "rows_last_step": [int(s.block_h) for s in speeds],
            "exchange": ("CUDA IPC mappings of the workers' export buffers, k_merge on rank 0 loads the rows over NVLink while adding"
                         if use_ipc else "NCCL grouped send/recv of the rows + k_merge"),
            "e2e": {"value": e_rays / e_dt / 1e6, "unit": "Mrays/s",
                    "h2d_bytes_per_step": int((sc.nbytes() + seeds_bytes + 76) * world), "d2h_bytes_per_step": w * h * 4},
            "roofline": roofline, "cpu_baseline": None, "one_gpu_same_workload": one_gpu,
        }
        if one_gpu:
            line["speedup_vs_one_gpu_same_workload"] = {"value": line["value"] / one_gpu["value"], "e2e": line["e2e"]["value"] / one_gpu["value"]}
        emit(line)
    if xch is not None:
        tr.wait_for_kernels()
        dist.barrier()
        xch.close()
    tr.close()
    dist.destroy_process_group()
    return 0


_REAL_STDOUT = None


def emit(line: dict):
    """Print the ONE JSON line on the real stdout (see main(): fd 1 is pointed at stderr while we run)."""
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is not None:
        os.write(_REAL_STDOUT, data)
    else:
        sys.stdout.write(data.decode())
        sys.stdout.flush()


def main():
    global _REAL_STDOUT
    # NCCL / torch / CUDA libraries occasionally write to fd 1 (e.g. "NCCL version ..."): keep fd 1 for the JSON
    # line only by routing everything else to stderr at the file-descriptor level.
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--spp", type=int, default=0, help="override samples per step (debugging only; invalidates the config)")
    ap.add_argument("--cpu-spp", type=int, default=8, help="spp of the bounded CPU-baseline sample")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-single", action="store_true", help="N > 1: skip the one-GPU measurement of the same workload")
    ap.add_argument("--no-opencl", action="store_true", help="skip the reference-OpenCL-kernels-on-this-GPU baseline")
    ap.add_argument("--verbose", action="store_true", help="per-step breakdown on stderr (N > 1)")
    ap.add_argument("--chains", type=int, default=0, help="override PC_OPT_SAMPLE_CHAINS (experiments)")
    ap.add_argument("--cold-scene", action="store_true", help="N=1: also time raw triangles -> device scene compile -> first frame "
                    "(default: only when the raw scene is in memory anyway, i.e. not loaded from POLARIS_SCENE_CACHE)")
    ap.add_argument("--exchange", default="ipc", choices=["ipc", "nccl"],
                    help="N > 1: how block rows reach rank 0 (ipc: CUDA IPC mappings + peer loads, nccl: grouped send/recv)")
    ap.add_argument("--opt", action="append", default=[], help="NAME=VALUE tracer option, e.g. PRIMARY_PACKETS=0 (experiments)")
    ap.add_argument("--config", default="c2", choices=["c1", "c2", "c3", "c4", "c5"],
                    help="N=1 only: BASELINE config to run (default c2 = configs[1], the one the metric is quoted on)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        return run_cuda_multi(args)
    if args.gpus > 1:
        log("[bench] --gpus > 1 must be launched with torch.distributed.run (one rank per GPU)")
        return 2
    return run_cuda_single(args)


if __name__ == "__main__":
    sys.exit(main())
