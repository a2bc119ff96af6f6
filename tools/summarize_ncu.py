#!/usr/bin/env python
"""Summarise gpurun_out/prof_<kernel>_<round>.ncu-rep + launches_<round>.csv into profiles/ (tracked):
   profiles/ncu_<round>_summary.md, profiles/launches_<round>.csv, profiles/traffic.json (bench.py reads it).
   python tools/summarize_ncu.py r02 c3      (tag, config; files gpurun_out/prof_<kernel>_<tag>_<config>.ncu-rep)
   traffic.json is keyed by config: {"c3": {"k_trace": {"dram_bytes": per launch, "l2_bytes": ..., "lanes_active": ..,
   "ipc": .., "l1_hit_pct": .., "l2_hit_pct": .., "bound": ".."}}}"""
import csv, json, os, subprocess, sys, collections

R = sys.argv[1]
CFG = sys.argv[2] if len(sys.argv) > 2 else "c2"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "gpurun_out")
P = os.environ.get("PROFILES_OUT") or os.path.join(ROOT, "profiles")
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__inst_executed.avg.per_cycle_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "l1tex__t_sector_hit_rate.pct",
        "lts__t_sector_hit_rate.pct", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct", "sm__sass_branch_targets_threads_divergent.sum",
        "smsp__sass_average_branch_targets_threads_uniform.pct", "launch__grid_size", "launch__block_size"]


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    return rows[0], rows[1], rows[2:]


md = [f"# ncu summary {R} (bench.py --config {CFG} --steps 1 --warmup 1 --spp 8 --chains 1, one B200)\n",
      "`--set full --clock-control none --import-source on`, every launch of the class in one batch of samples (the third of the run).  Durations under ncu are",
      "cold-cache and serialised: compare shares, not absolutes.  dram bytes are per launch.\n"]
traffic = {}
for k in ("k_shade", "k_trace", "k_query", "k_occlusion", "k_primary"):
    rep = os.path.join(G, f"prof_{k}_{R}_{CFG}.ncu-rep")
    if not os.path.exists(rep):
        continue
    hdr, units, rows = raw(rep)
    md.append(f"## {k}\n")
    md.append("| metric | unit | " + " | ".join(f"launch {k + 1}" for k in range(len(rows))) + " |")
    md.append("|---|---|" + "---|" * len(rows))
    vals = []
    for key in KEYS:
        if key in hdr:
            i = hdr.index(key)
            md.append(f"| {key} | {units[i]} | " + " | ".join(r[i] for r in rows) + " |")
    stall = collections.OrderedDict()
    for i, h in enumerate(hdr):
        if h.startswith("smsp__pcsamp_warps_issue_stalled_") and not h.endswith("_not_issued"):
            try:
                stall[h.replace("smsp__pcsamp_warps_issue_stalled_", "")] = float(rows[0][i])
            except ValueError:
                pass
    tot = sum(stall.values()) or 1
    md.append("\nstall reasons (launch 1): " + ", ".join(f"{h} {100*v/tot:.0f}%" for h, v in sorted(stall.items(), key=lambda kv: -kv[1])[:8]) + "\n")
    try:
        ir, iw = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
        scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        tscale = {"ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0}
        tb = [float(r[ir]) * scale.get(units[ir], 1) + float(r[iw]) * scale.get(units[iw], 1) for r in rows]
        def col(name, row=0, default=None):
            try:
                return float(rows[row][hdr.index(name)])
            except Exception:
                return default
        ent = {"dram_bytes": sum(tb) / len(tb)}
        try:
            il0 = hdr.index("lts__t_sectors.sum")
            ent["l2_bytes"] = sum(float(r[il0]) * 32.0 for r in rows) / len(rows)
        except Exception:
            pass
        ent["launches_captured"] = len(rows)
        # instruction-weighted over the captured launches (the first bounce's launch dominates)
        def wmean(name):
            try:
                wi = hdr.index("smsp__inst_executed.sum")
                w = [float(r[wi]) for r in rows]
                v = [float(r[hdr.index(name)]) for r in rows]
                return sum(a * b for a, b in zip(v, w)) / max(1.0, sum(w))
            except Exception:
                return col(name)
        ent["lanes_active"] = wmean("smsp__thread_inst_executed_per_inst_executed.ratio")
        ent["ipc"] = wmean("sm__inst_executed.avg.per_cycle_elapsed")
        ent["l1_hit_pct"] = wmean("l1tex__t_sector_hit_rate.pct")
        ent["l2_hit_pct"] = wmean("lts__t_sector_hit_rate.pct")
        ent["dram_pct_of_peak_under_ncu"] = wmean("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed")
        ent["registers"] = col("launch__registers_per_thread")
        top = sorted(stall.items(), key=lambda kv: -kv[1])[:2]
        ent["top_stalls"] = {h: round(v / tot, 3) for h, v in top}
        dram_pct = ent["dram_pct_of_peak_under_ncu"] or 0.0
        ent["bound"] = "hbm" if dram_pct >= 60.0 else ("latency/simt (divergence + dependent-load latency)" if (ent["lanes_active"] or 32) < 20 else "latency")
        traffic[k] = ent
        # derived: achieved HBM and L2 GB/s under the profiler (cold caches, serialised launches)
        it, il = hdr.index("gpu__time_duration.sum"), hdr.index("lts__t_sectors.sum")
        secs = [float(r[it]) * tscale.get(units[it], 1e-9) for r in rows]
        l2 = [float(r[il]) * 32.0 for r in rows]  # 32-byte sectors
        md.append("derived: HBM " + " / ".join(f"{b / t / 1e9:.0f}" for b, t in zip(tb, secs)) + " GB/s, L2 "
                  + " / ".join(f"{b / t / 1e9:.0f}" for b, t in zip(l2, secs)) + " GB/s (per launch, under the profiler)\n")
    except Exception:
        pass
os.makedirs(P, exist_ok=True)
open(os.path.join(P, f"ncu_{R}_{CFG}_summary.md"), "w").write("\n".join(md) + "\n")
src = os.path.join(G, f"launches_{R}_{CFG}.csv")
if os.path.exists(src):
    rows = [r for r in csv.reader(open(src)) if len(r) > 14 and r[0].isdigit()]
    tot = collections.defaultdict(float); cnt = collections.Counter()
    for r in rows:
        name = r[4].split("(")[0].replace("void ", "")
        tot[name] += float(r[14]); cnt[name] += 1
    s = sum(tot.values())
    with open(os.path.join(P, f"launches_{R}_{CFG}.csv"), "w") as f:
        f.write("kernel,launches,total_ns,share,avg_ns\n")
        for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
            f.write(f"{k},{cnt[k]},{v:.0f},{v/s:.4f},{v/cnt[k]:.0f}\n")
if traffic:
    tp = os.path.join(P, "traffic.json")
    try:
        allt = json.load(open(tp))
    except Exception:
        allt = {}
    allt = {k: v for k, v in allt.items() if isinstance(v, dict)}  # drop the round-1 flat layout
    traffic["_note"] = f"per launch, ncu --set full --clock-control none, {R}, bench.py --config {CFG} --spp 8 --chains 1 (mean over every launch of the class in one batch of samples)"
    allt[CFG] = traffic
    json.dump(allt, open(tp, "w"), indent=1)
print("\n".join(md[:12]))
print(traffic)
