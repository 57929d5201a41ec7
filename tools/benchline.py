"""Print the short form of a bench.py JSON line read from stdin: workload, ms/step, launches, per-kernel ms."""
import json
import sys

for raw in sys.stdin:
    raw = raw.strip()
    if not raw.startswith("{"):
        continue
    d = json.loads(raw)
    km = (d.get("roofline") or {}).get("kernel_ms") or {}
    print(d["config"]["workload"][:34], "ms/step", round(d["ms_per_step"], 3), "launches", d.get("gpu_launches"),
          {k.split(" ")[0]: round(v, 2) for k, v in km.items()}, "e2e", (d.get("e2e") or {}).get("value"))
