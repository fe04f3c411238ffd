"""Prints the interesting keys of a bench.py JSON line read from stdin (used by the tuning scripts)."""
import json
import sys

tag = sys.argv[1] if len(sys.argv) > 1 else ""
try:
    d = json.loads(sys.stdin.read().strip().splitlines()[-1])
    c = d.get("comm") or {}
    print(tag, d["value"], d["ms_per_step"], d["roofline"]["kernel_ms"], d["roofline"]["frac"], d["clocks"]["sm_mhz"], d["clocks"]["reasons"],
          c.get("step_minus_sweep_ms"), c.get("timeline_ms"))
except Exception as e:
    print(tag, "ERR", e)
