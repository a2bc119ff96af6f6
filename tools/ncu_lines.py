#!/usr/bin/env python
"""Join an ncu source-page CSV (per SASS instruction: stall samples, executed, threads) with nvdisasm -g line
info of the same binary, and aggregate per CUDA source line / per device function.

  ncu -i prof.ncu-rep --page source --csv > prof.csv
  cuobjdump -xelf all lib.so; nvdisasm -g -c x.cubin > all.sass
  python tools/ncu_lines.py prof.csv all.sass <kernel mangled-name substring> [--top 40]
"""
import csv, re, sys, collections

prof, sass, kname = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 40

# ---- nvdisasm: sections in file order -> list of (line, text)
sections = {}
cur = None
line = ("?", 0)
order = []
for l in open(sass):
    m = re.match(r"\s*\.section\s+\.text\.(\S+?),", l)
    if m:
        cur = m.group(1); sections[cur] = []; order.append(cur); line = ("?", 0); continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
    if m:
        line = (m.group(1).split("/")[-1], int(m.group(2))); continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m and cur:
        sections[cur].append((int(m.group(1), 16), line, m.group(2).strip()))
target = [s for s in order if kname in s]
assert target, f"no section matches {kname}"
target = target[0]

# ---- ncu csv: first launch only
rows = list(csv.reader(open(prof)))
hdr = rows[1]
ia, isrc = hdr.index("Address"), hdr.index("Source")
iall, iex, ithr = hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Instructions Executed"), hdr.index("Avg. Threads Executed")
insts = []
for r in rows[2:]:
    try:
        insts.append((int(r[ia], 16), r[isrc].strip(), int(r[iall]), int(r[iex]), float(r[ithr])))
    except Exception:
        pass
# split launches: addresses restart
first = [insts[0]]
for x in insts[1:]:
    if x[0] <= first[-1][0]:
        break
    first.append(x)
base = first[0][0]
sec = sections[target]
print(f"kernel section {target[:60]}: {len(sec)} SASS instructions; ncu rows (first launch): {len(first)}")
by_off = {off: (ln, txt) for off, ln, txt in sec}
tot_s = sum(x[2] for x in first); tot_e = sum(x[3] for x in first)
agg = collections.defaultdict(lambda: [0, 0, 0.0])
unmatched = 0
for addr, src, s_, ex, thr in first:
    off = addr - base
    if off in by_off:
        ln = by_off[off][0]
    else:
        ln = ("(outside kernel section)", 0); unmatched += 1
    a = agg[ln]; a[0] += s_; a[1] += ex; a[2] += ex * thr
print(f"total stall samples {tot_s}, warp instructions {tot_e}, unmatched rows {unmatched}")
print(f"{'file:line':34s} {'samples%':>8s} {'instr%':>7s} {'avg thr':>7s}")
for ln, (s_, ex, wt) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{ln[0]+':'+str(ln[1]):34s} {100*s_/tot_s:8.2f} {100*ex/max(1,tot_e):7.2f} {wt/max(1,ex):7.1f}")
