#!/usr/bin/env python
"""Print the headline numbers, the per-kernel times and the extra legs of a bench.py JSON line.  Usage: bench_summary.py bench.json"""
import json
import sys

d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
rb = d["roofline_backbone"]
print(f"value {d['value']:.0f} {d['unit']}  step {d['ms_per_step']:.3f} ms  e2e {d['e2e']['value']:.0f}  backbone {rb['ms']:.3f} ms  "
      f"frac {rb['frac']:.3f}  det heads {rb.get('det_heads_ms', 0):.3f} ms  launches {d['gpu_launches']}  clocks {d.get('clocks')}")
print("dominant:", {k: d["roofline"][k] for k in ("kernel", "frac", "ms_per_launch", "traffic")})
print(" ".join(f"{l['kernel']}={l['ms']:.3f}({l['frac']:.2f})" for l in d["layers"]))
for k in ("cpu_baseline", "sustained", "e2e_trained_weights", "latency"):
    if d.get(k):
        print(k, json.dumps(d[k], default=float))
for k, v in (d.get("configs") or {}).items():
    print(k, json.dumps(v, default=float))
if d.get("train"):
    print("train", json.dumps(d["train"], default=float))
