"""Prints the headline fields of a bench.py JSON line read from stdin."""
import json, sys
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print("ms_per_step", round(d["ms_per_step"], 3), "value", round(d["value"] / 1e6, 2), "M/s", "launches", d.get("gpu_launches"))
print("breakdown", {k: round(v, 3) for k, v in d.get("breakdown_ms", {}).items()})
print("serial", {k: round(v, 3) for k, v in d.get("linear_kernels_ms_serialised", {}).items()})
print("roofline", round(d["roofline"]["achieved"], 2), d["roofline"]["unit"], "frac", round(d["roofline"]["frac"], 3))
