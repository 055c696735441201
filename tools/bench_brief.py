#!/usr/bin/env python3
"""Print the few numbers of a bench.py JSON line that matter while tuning (stdin: bench output)."""
import json
import sys

for line in sys.stdin:
    line = line.strip()
    if not line.startswith("{"):
        continue
    d = json.loads(line)
    r = d.get("roofline", {})
    print("value %.0f  e2e %.0f  %s  ms/step %.2f  kernels/wave %s  frac %.4f" % (
        d["value"], d["e2e"]["value"], d["unit"], d["ms_per_step"], {k: round(v, 2) for k, v in (r.get("kernel_ms_per_wave") or r.get("kernel_ms_all") or {}).items()}, r.get("frac", 0)))
