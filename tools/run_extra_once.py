#!/usr/bin/env python3
"""One junc + `--extra` pass over a pjsynth preset through the C ABI (for ncu captures of the k_x_* kernels).

    python tools/run_extra_once.py --preset c2 --scale 1.0 [--repeat 2]
"""
import argparse
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--preset", default="c2")
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--repeat", type=int, default=2)
    ap.add_argument("--workdir", default="/tmp/pj_extra_once")
    a = ap.parse_args()
    import numpy as np
    from portcullis_b200 import junction_builder as jb
    prep = os.path.join(a.workdir, "prep")
    if not os.path.exists(os.path.join(prep, "synth.json")):
        subprocess.check_call([os.path.join(ROOT, "portcullis_b200", "bin", "pjsynth"), "--preset", a.preset, "--scale", str(a.scale), "--out", prep],
                              stderr=subprocess.DEVNULL)
    p = jb.PrepDir(prep)
    cols = p.decode(-1, os.cpu_count() or 1, names=True)
    g = jb.JuncGpu(0, "UNKNOWN", extra=True)
    g.set_targets(p.lengths)
    for t in range(len(p.names)):
        g.set_genome(t, p.genome(t))
    for it in range(a.repeat):
        g.shard_begin(len(cols["pos"]), len(cols["cigar"]), len(cols["seq4"]))
        g.submit(cols)
        g.run()
        rows, st = g.fetch()
        t0 = time.perf_counter()
        x, over = g.extra(rows, int(st["maxq"].max()))
        print("pass %d: %d junctions, extra wall %.1f ms, device %.3f ms, stages %s, capped targets %s" %
              (it, len(rows), (time.perf_counter() - t0) * 1e3, g.extra_timing()[0], g.extra_timing()[2], sorted(over)))
    print("sum up_aln %d down_aln %d mm_m %d cov %d" % (x["up_aln"].astype(np.int64).sum(), x["down_aln"].astype(np.int64).sum(),
                                                       x["mm_m"].astype(np.int64).sum(), x["cov_sum"].astype(np.int64).sum()))
    g.close()


if __name__ == "__main__":
    main()
