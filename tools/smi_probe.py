#!/usr/bin/env python3
"""How much does clock sampling perturb the timed steps?  python tools/smi_probe.py [preset] [scale]"""
import os, subprocess, sys, threading, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import bench
from portcullis_b200 import junction_builder as jb
preset = sys.argv[1] if len(sys.argv) > 1 else "c2"; scale = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
prep, meta = bench.make_workload(preset, scale, 0, 16)
p = jb.PrepDir(prep)
runs = p.decode_segment_lean(1, 0, 0, 1 << 40, threads=16, keep_mate=False, copy=True)
g = jb.JuncGpu(0, "UNKNOWN"); g.set_targets(p.lengths)
for r in runs: g.set_genome(r["tid"], p.genome(r["tid"]))
g.shard_begin(sum(len(r["pos"]) for r in runs), sum(len(r["cigar"]) for r in runs), 2 * sum(len(r["seq2"]) for r in runs))
for r in runs: g.submit_lean(r)
for _ in range(5): g.run()

def loop(seconds):
    n = 0; worst = 0.0; t0 = time.perf_counter()
    while time.perf_counter() - t0 < seconds:
        a = time.perf_counter(); g.run(); d = time.perf_counter() - a; worst = max(worst, d); n += 1
    return (time.perf_counter() - t0) / n * 1e3, worst * 1e3, n

FULL = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
MIN = "index,clocks.sm,clocks.max.sm,clocks_event_reasons.active"
print("no sampling      : mean %.3f ms  worst %.2f ms  (%d steps)" % loop(3.0))
for name, q, ms in (("smi -lms 250 full", FULL, 250), ("smi -lms 250 min ", MIN, 250), ("smi -lms 1000 full", FULL, 1000)):
    pr = subprocess.Popen(["nvidia-smi", "-i", "0", "--query-gpu=" + q, "--format=csv,noheader,nounits", "-lms", str(ms)], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    time.sleep(1.5)
    print("%s: mean %.3f ms  worst %.2f ms  (%d steps)" % ((name,) + loop(3.0)))
    pr.terminate(); pr.wait()
try:
    import pynvml
    pynvml.nvmlInit(); h = pynvml.nvmlDeviceGetHandleByIndex(0)
    stop = threading.Event(); samples = []
    def poll(fn, iv):
        while not stop.is_set():
            samples.append(fn()); stop.wait(iv)
    for name, fn, iv in (("pynvml clock+reasons 100ms", lambda: (pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM), pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)), 0.1),
                         ("pynvml clock only 100ms   ", lambda: pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM), 0.1),
                         ("pynvml reasons only 100ms ", lambda: pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h), 0.1),
                         ("pynvml power 100ms        ", lambda: pynvml.nvmlDeviceGetPowerUsage(h), 0.1)):
        stop.clear(); samples.clear()
        t = threading.Thread(target=poll, args=(fn, iv), daemon=True); t.start(); time.sleep(0.3)
        print("%s: mean %.3f ms  worst %.2f ms  (%d steps), %d samples, last %s" % ((name,) + loop(3.0) + (len(samples), samples[-1] if samples else None)))
        stop.set(); t.join()
except Exception as e:
    print("pynvml unavailable:", e)
print("no sampling again: mean %.3f ms  worst %.2f ms  (%d steps)" % loop(2.0))
g.close()
