#!/usr/bin/env python3
"""Per-step stage times of the resident arm (looks for one-off stalls): python tools/step_probe.py <preset> <scale> [steps]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import bench
from portcullis_b200 import junction_builder as jb
preset = sys.argv[1]; scale = float(sys.argv[2]); steps = int(sys.argv[3]) if len(sys.argv) > 3 else 12
prep, meta = bench.make_workload(preset, scale, 0, 16)
p = jb.PrepDir(prep)
runs = p.decode_segment_lean(1, 0, 0, 1 << 40, threads=16, keep_mate=False, copy=True)
g = jb.JuncGpu(0, "UNKNOWN"); g.set_targets(p.lengths)
for r in runs: g.set_genome(r["tid"], p.genome(r["tid"]))
n_rec = sum(len(r["pos"]) for r in runs)
g.shard_begin(n_rec, sum(len(r["cigar"]) for r in runs), 2 * sum(len(r["seq2"]) for r in runs))
for r in runs: g.submit_lean(r)
for it in range(steps):
    t0 = time.perf_counter(); g.run(); dt = (time.perf_counter() - t0) * 1e3
    ms, nl, st = g.timing()
    print("step %2d wall %.2f dev %.2f | %s" % (it, dt, ms, " ".join("%s=%.2f" % (k, v) for k, v in st)))
g.close()
