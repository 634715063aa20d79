#!/usr/bin/env python3
"""Print the fields of a bench.py JSON line that matter at a glance (reads stdin or a file)."""
import json
import sys

src = open(sys.argv[1]) if len(sys.argv) > 1 else sys.stdin
lines = [json.loads(l) for l in src if l.startswith("{")]
d = lines[-1]
e = d.get("e2e", {})
print("N", d.get("n_gpus"), "ms/step %.4f" % d["ms_per_step"], "value %.0f" % d["value"], "frac %.3f" % d["roofline"]["frac"],
      "| e2e ms %.2f" % e.get("ms_per_step", 0), "value %.0f" % e.get("value", 0),
      "blocking %.2f" % e.get("blocking", {}).get("ms_per_step", 0), "| staging %.3f" % d["staging"]["ms"],
      "|", d.get("multi_gpu"))
print("  between results:", e.get("ms_between_results"))
