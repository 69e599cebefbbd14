#!/usr/bin/env python3
"""Where does an e2e step (C ABI, lean batches in page-locked host memory) spend its time?  python tools/e2e_probe.py [preset] [scale]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import bench
from portcullis_b200 import junction_builder as jb, _lib as L
from portcullis_b200.columnar import lean_nbytes
preset = sys.argv[1] if len(sys.argv) > 1 else "c2"; scale = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
prep, meta = bench.make_workload(preset, scale, 0, 16)
p = jb.PrepDir(prep)
runs, whole = p.decode_segment_lean(1, 0, 0, 1 << 40, threads=16, keep_mate=False, copy=False, with_whole=True)
cudart = torch.cuda.cudart()
for k, v in whole.items():
    if v.nbytes:
        rc = cudart.cudaHostRegister(v.ctypes.data, v.nbytes, 0)
        print("register", k, v.nbytes, "->", rc, int(rc))
n_rec = sum(len(r["pos"]) for r in runs); ncig = sum(len(r["cigar"]) for r in runs); ns2 = sum(len(r["seq2"]) for r in runs)
print("records", n_rec, "lean MB", sum(lean_nbytes(r) for r in runs) / 1e6)
g = jb.JuncGpu(0, "UNKNOWN"); g.set_targets(p.lengths)
for r in runs: g.set_genome(r["tid"], p.genome(r["tid"]))
def sync(): torch.cuda.synchronize()
for it in range(5):
    sync(); t0 = time.perf_counter()
    g.shard_begin(n_rec, ncig, 2 * ns2); sync(); t1 = time.perf_counter()
    for r in runs: g.submit_lean(r)
    t2 = time.perf_counter(); sync(); t3 = time.perf_counter()
    nj = g.run(); sync(); t4 = time.perf_counter()
    rows, st = g.fetch(); sync(); t5 = time.perf_counter()
    print("step %d: begin %.2f  submit(host) %.2f  submit(drain) %.2f  run %.2f (device %.2f)  fetch %.2f  total %.2f ms" % (
        it, (t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3, (t4 - t3) * 1e3, g.timing()[0], (t5 - t4) * 1e3, (t5 - t0) * 1e3))
# raw link speed from the same registered memory
big = whole["seq2"]; t = torch.empty(big.nbytes, dtype=torch.uint8, device="cuda")
src = torch.from_numpy(big)
for _ in range(3):
    sync(); t0 = time.perf_counter(); t.copy_(src, non_blocking=True); sync(); dt = time.perf_counter() - t0
    print("torch copy of seq2 (%.0f MB): %.1f GB/s (is_pinned=%s)" % (big.nbytes / 1e6, big.nbytes / dt / 1e9, src.is_pinned()))
