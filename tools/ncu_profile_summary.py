#!/usr/bin/env python3
"""Turn one round's ncu artefacts into the committed summaries under profiles/.

usage: tools/ncu_profile_summary.py <tag> <launches.csv> <full.ncu-rep> [dominant-kernel-regex] [streams-per-launch]
  <launches.csv>  from `ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file ... python bench.py ...`
  <full.ncu-rep>  from `ncu --set full --clock-control none --import-source on -k regex:... -o ... python bench.py ...`
writes profiles/<tag>_launches.txt, profiles/<tag>_full.txt and updates profiles/ncu_summary.json (read by bench.py for
roofline.traffic).  ncu launch times are cold-cache and serialised: use the kernels' SHARES, not the absolute values."""
import collections
import csv
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag, launches, rep = sys.argv[1], sys.argv[2], sys.argv[3]   # rep: one report or several joined with ","
dom = sys.argv[4] if len(sys.argv) > 4 else "sb_enc_nsq"
streams_per_launch = int(sys.argv[5]) if len(sys.argv) > 5 else None
os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)

# ---- launch list ----
rows = [r for r in csv.reader(l for l in open(launches) if l.startswith('"'))]
hdr = rows[0]
ki, vi, gi, bi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size"), hdr.index("Block Size")
per = collections.OrderedDict()
for r in rows[1:]:
    name = re.sub(r"\(.*", "", r[ki])
    per.setdefault(name, []).append((float(r[vi].replace(",", "")), r[gi], r[bi]))
tot = sum(sum(t for t, _, _ in v) for k, v in per.items() if k.startswith("sb_") and "init" not in k)
with open(os.path.join(ROOT, "profiles", tag + "_launches.txt"), "w") as f:
    f.write("# %s: ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised launches)\n" % tag)
    f.write("# share = kernel's part of the codec kernels' total (init and torch kernels excluded)\n")
    f.write("%-28s %8s %12s %12s %8s  %s\n" % ("kernel", "launches", "avg_us", "total_us", "share", "grid x block"))
    for k, v in per.items():
        t = sum(x for x, _, _ in v)
        share = "%.1f%%" % (100 * t / tot) if k.startswith("sb_") and "init" not in k else "-"
        f.write("%-28s %8d %12.1f %12.1f %8s  %s x %s\n" % (k[:28], len(v), t / len(v) / 1e3, t / 1e3, share, v[0][1], v[0][2]))
print(open(os.path.join(ROOT, "profiles", tag + "_launches.txt")).read())

# ---- full capture ----
h, data, data_units = None, [], []
for one in rep.split(","):
    raw = subprocess.run(["ncu", "-i", one, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rr = list(csv.reader(raw.splitlines()))
    if h is None:
        h = rr[0]
    pos = {n: i for i, n in enumerate(rr[0])}     # align columns (and units: they are per report) by metric name
    for r in rr[2:]:
        data.append([r[pos[n]] if n in pos else "" for n in h])
        data_units.append([rr[1][pos[n]] if n in pos else "" for n in h])
want = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct", "smsp__inst_executed.sum",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "smsp__pcsamp_warps_issue_stalled_long_scoreboard", "smsp__pcsamp_warps_issue_stalled_wait",
    "smsp__pcsamp_warps_issue_stalled_short_scoreboard", "smsp__pcsamp_warps_issue_stalled_math_pipe_throttle",
    "smsp__pcsamp_warps_issue_stalled_not_selected", "smsp__pcsamp_warps_issue_stalled_selected",
    "smsp__pcsamp_warps_issue_stalled_branch_resolving", "smsp__pcsamp_warps_issue_stalled_barrier",
    "smsp__pcsamp_warps_issue_stalled_lg_throttle", "smsp__pcsamp_warps_issue_stalled_mio_throttle",
]
summary = {}


def num(s):
    try:
        return float(s.replace(",", ""))
    except Exception:
        return None


with open(os.path.join(ROOT, "profiles", tag + "_full.txt"), "w") as f:
    f.write("# %s: ncu --set full --clock-control none --import-source on (one launch per kernel, mid-run)\n" % tag)
    seen = set()
    for r, units in zip(data, data_units):
        name = re.sub(r"\(.*", "", r[h.index("Kernel Name")])
        if name in seen:
            continue
        seen.add(name)
        f.write("\n== %s\n" % name)
        vals = {}
        for w in want:
            if w in h:
                i = h.index(w)
                f.write("  %-62s %16s %s\n" % (w, r[i], units[i]))
                vals[w] = (num(r[i]), units[i])
        rd, wr = vals.get("dram__bytes_read.sum"), vals.get("dram__bytes_write.sum")

        def to_bytes(v):
            if not v or v[0] is None:
                return None
            return v[0] * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(v[1], 1)
        if rd and wr:
            tb = to_bytes(rd) + to_bytes(wr)
            f.write("  %-62s %16.0f byte\n" % ("dram traffic per launch (read + write)", tb))
            summary[name] = {"dram_bytes_per_launch": tb, "duration_ns": (vals["gpu__time_duration.sum"][0] or 0) * {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9}.get(vals["gpu__time_duration.sum"][1], 1),
                             "grid": vals.get("launch__grid_size", (None,))[0], "issue_active_pct": vals.get("smsp__issue_active.avg.pct", (None,))[0]}
print(open(os.path.join(ROOT, "profiles", tag + "_full.txt")).read())
js = os.path.join(ROOT, "profiles", "ncu_summary.json")
cur = json.load(open(js)) if os.path.exists(js) else {}
cur[tag] = summary
for k, v in summary.items():
    if re.search(dom, k):
        cur["dominant_kernel"] = k
        cur["encode_kernel_dram_bytes_per_launch"] = v["dram_bytes_per_launch"]
        cur["encode_kernel_grid"] = v["grid"]
        if streams_per_launch:
            cur["encode_kernel_streams_per_launch"] = streams_per_launch
        cur["from"] = tag
json.dump(cur, open(js, "w"), indent=1)
