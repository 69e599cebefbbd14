#!/usr/bin/env python3
"""Which ingredient of bench.py's resident arm makes timed steps 2-9 stall?  python tools/stall_probe.py preset flags
flags: t = timing() after every step, s = torch.cuda.synchronize() before the loop, w = copy=False/with_whole decode, f = fetch() first,
       r = cudaHostRegister of the columns, g = gc.disable()"""
import gc, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import bench
from portcullis_b200 import junction_builder as jb
preset, flags = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "")
prep, meta = bench.make_workload(preset, 1.0, 0, 16)
if "s" in flags: torch.cuda.set_device(0); torch.cuda.synchronize()
p = jb.PrepDir(prep)
if "w" in flags:
    runs, whole = p.decode_segment_lean(1, 0, 0, 1 << 40, threads=16, keep_mate=False, copy=False, with_whole=True)
else:
    runs, whole = p.decode_segment_lean(1, 0, 0, 1 << 40, threads=16, keep_mate=False, copy=True), {}
if "r" in flags:
    cudart = torch.cuda.cudart()
    for k, v in whole.items():
        if v.nbytes: cudart.cudaHostRegister(v.ctypes.data, v.nbytes, 0)
g = jb.JuncGpu(0, "UNKNOWN"); g.set_targets(p.lengths)
for r in runs: g.set_genome(r["tid"], p.genome(r["tid"]))
g.shard_begin(sum(len(r["pos"]) for r in runs), sum(len(r["cigar"]) for r in runs), 2 * sum(len(r["seq2"]) for r in runs))
for r in runs: g.submit_lean(r)
g.run()
if "f" in flags: rows, st = g.fetch(); del rows
if "g" in flags: gc.disable()
for _ in range(5): g.run()
if "s" in flags: torch.cuda.synchronize()
tr = []
for _ in range(24):
    a = time.perf_counter(); g.run(); d = (time.perf_counter() - a) * 1e3
    if "t" in flags: g.timing()
    tr.append(round(d, 2))
print("%-3s [%-6s] max %.2f  steps: %s" % (preset, flags, max(tr), tr[:12]))
g.close()
