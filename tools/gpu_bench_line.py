#!/usr/bin/env python
"""one-line summary of a bench.py JSON line: usage gpu_bench_line.py file [label]"""
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
r = d["roofline"]
print(f"{sys.argv[2] if len(sys.argv) > 2 else sys.argv[1]}: {d['value'] / 1e6:.2f} M solves/s  frac {r['frac']:.3f}  launch {r['avg_launch_ms']:.2f} ms  kernel {r['kernel']}  e2e {(d['e2e'] or {}).get('value', 0) / 1e6:.2f}  {d.get('stream_kernel')}")
