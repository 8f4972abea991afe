#!/usr/bin/env python
"""Condense a bench.py JSON line (stdin) to one short line."""
import json, sys
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
r = d.get("roofline") or {}
print("%-50s %8.3f GTEPS %8.3f ms/step  e2e %.2f GTEPS  dominant %s %.3f ms frac %.2f  kernels %s" % (
    d["config"]["workload"][:50], d["value"], d["ms_per_step"], d["e2e"]["value"], r.get("kernel"), r.get("kernel_ms") or 0,
    r.get("frac") or 0, {k: round(v, 3) for k, v in (r.get("kernel_ms_by_name") or {}).items()}))
