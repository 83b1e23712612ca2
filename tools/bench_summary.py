#!/usr/bin/env python
"""Print the headline numbers and the per-layer times of a bench.py JSON line.  Usage: bench_summary.py bench.json"""
import json
import sys

d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
rb = d["roofline_backbone"]
print(f"value {d['value']:.0f} {d['unit']}  step {d['ms_per_step']:.3f} ms  e2e {d['e2e']['value']:.0f}  backbone {rb['ms']:.3f} ms  "
      f"frac {rb['frac']:.3f}  det heads {rb.get('det_heads_ms', 0):.3f} ms  launches {d['gpu_launches']}")
print(" ".join(f"{l['kernel']}={l['ms']:.3f}" for l in d["layers"]))
