#!/usr/bin/env python3
"""Benchmark of the junc hot path (BASELINE.json metric: spliced alignments / second through junc).

    python bench.py --gpus N --steps K --warmup W            # our CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's own CPU implementation (oracle/_ref)

Workload: config "c2" of BASELINE.json (synthetic 100 Mb genome = 10 targets x 10 Mb, 10 M 2x150 alignments), made on
the box by portcullis_b200/bin/pjsynth.  With N ranks every rank owns its own 10 targets / 10 M alignments (weak
scaling: targets are independent shards, no data-path collective).  One "step" = one pass of the hot path over the
rank's whole shard.

 value : spliced alignments / s, whole job, alignment columns and packed genome already resident in HBM
 e2e   : same metric through the C ABI (pj_shard_begin / pj_batch_submit / pj_shard_run / pj_shard_fetch) with the
         columns in PINNED HOST memory: H2D of every column and D2H of the junction rows inside the timed region
 e2e_bam (extra): the JunctionBuilder front end from the BAM file (BGZF decode, genome load, GPU, writers), once
"""
import argparse
import json
import os
import shutil
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

PJSYNTH = os.path.join(ROOT, "portcullis_b200", "bin", "pjsynth")
REF_BIN = os.path.join(ROOT, "oracle", "_ref", "portcullis_ref")
WORKDIR = os.environ.get("PJ_BENCH_DIR", "/tmp/pj_bench")
METRIC = "spliced_alignments_per_sec"
UNIT = "spliced alignments/s"


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def make_workload(preset, scale, seed, threads):
    d = os.path.join(WORKDIR, "%s_x%g_s%d" % (preset, scale, seed))
    meta = os.path.join(d, "synth.json")
    if not os.path.exists(meta):
        shutil.rmtree(d, ignore_errors=True)
        os.makedirs(d, exist_ok=True)
        subprocess.check_call([PJSYNTH, "--preset", preset, "--scale", str(scale), "--seed", str(seed), "--threads", str(threads),
                               "--out", d], stderr=subprocess.DEVNULL)
    with open(meta) as f:
        return d, json.load(f)


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        super().__init__(daemon=True)
        self.gpu = gpu_index
        self.rows = []
        self.stop = threading.Event()

    def run(self):
        while not self.stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits"],
                                     stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            self.stop.wait(0.2)

    def summary(self):
        sm = sorted(int(float(r[1])) for r in self.rows if len(r) > 2 and r[1].replace(".", "").isdigit())
        mx = [int(float(r[2])) for r in self.rows if len(r) > 2 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 9:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(self.rows)}


def algorithmic_bytes(cols, n_pairs, n_junc):
    """SURVEY §8(d) formula, evaluated on the actual workload (see DESIGN.md 'Algorithmic bytes')."""
    import numpy as np
    n_rec = len(cols["pos"])
    n_cig = len(cols["cigar"])
    seq_bytes = len(cols["seq4"])                    # 4-bit SEQ of spliced records only
    # CIGAR words of spliced records (the only ones the per-pair kernels touch)
    has_seq = np.diff(cols["seq_off"].astype(np.int64)) > 0
    n_cig_spliced = int(np.diff(cols["cigar_off"].astype(np.int64))[has_seq].sum())
    fixed = 32 * n_rec + 4 * n_cig + seq_bytes
    pair = 48 * n_pairs                               # 8-B key + 16-B payload, written once and read once
    genome = seq_bytes // 2                           # SURVEY: 2-bit genome window of the anchor bases
    rows = 326 * n_junc
    total = fixed + pair + genome + rows
    per_stage = {
        # what each kernel must at least move (inputs read once, outputs written once)
        # fused front end: every record column once, every CIGAR word once, one (key, PairA, PairB) per pair written
        "scan_emit": (4 + 4 + 2 + 1 + 1 + 4 + 4 + 4 + 4) * n_rec + 4 * n_cig + (8 + 32) * n_pairs,
        "scan_reads": (4 + 4 + 2 + 4 + 4) * n_rec + 4 * n_cig + 8 * n_rec,
        "pair_offsets": 8 * n_rec,
        "emit_pairs": (4 + 4 + 2 + 1 + 1 + 4 + 4 + 4 + 4 + 4) * n_rec + 4 * n_cig + (8 + 32) * n_pairs,
        "radix_sort": 2 * 12 * n_pairs,              # one read + one write of (key,val): any extra pass is overhead
        "segments": (8 + 4 + 4) * n_pairs,
        "reduce1": (4 + 4 + 32) * n_pairs + 4 * n_pairs,
        "entropy": 12 * n_pairs + 8 * n_junc,
        # k_match: vals + jid + PairA/B + result per pair; CIGAR and SEQ of the read; the same number of genome bases
        # from the 4-bit plane (SEQ and genome windows have equal length)
        "match": (4 + 4 + 32 + 16) * n_pairs + 4 * n_cig_spliced + seq_bytes + seq_bytes,
        "reduce2": (4 + 16) * n_pairs,
        "finalize": (256 + 100 + 84) * n_junc,
    }
    return total, per_stage


def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    from portcullis_b200 import _lib as L
    from portcullis_b200 import junction_builder as jb

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device visible; the junc path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    cores = max(1, host_cores() // world)

    prep, meta = make_workload(args.preset, args.scale, args.seed + rank, cores)
    t0 = time.time()
    p = jb.PrepDir(prep)
    cols = p.decode(-1, cores)
    t_decode = time.time() - t0
    n_rec = len(cols["pos"])
    genomes = [p.genome(t) for t in range(len(p.names))]

    # pinned host copies of every column (the e2e arm copies from these inside the timed region)
    pinned = {}
    keep = []
    for k, v in cols.items():
        t = torch.from_numpy(np.ascontiguousarray(v).view(np.uint8)).pin_memory() if v.size else torch.zeros(0, dtype=torch.uint8)
        keep.append(t)
        pinned[k] = t.numpy().view(v.dtype) if v.size else v
    h2d_bytes = int(sum(v.nbytes for v in cols.values()))

    g = jb.JuncGpu(local, "UNKNOWN")
    g.set_targets(p.lengths)
    t0 = time.time()
    for t, s in enumerate(genomes):
        g.set_genome(t, s)
    g.shard_begin(n_rec, len(cols["cigar"]), len(cols["seq4"]))
    g.submit(pinned)
    nj = g.run()                                     # also finishes the genome upload
    t_setup = time.time() - t0
    rows, st = g.fetch()
    n_spliced = int(st["spliced"].sum())
    n_pairs = int(rows["nb_raw_aln"].astype(np.int64).sum())
    d2h_bytes = int(rows.nbytes + 32 * len(p.lengths))

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # ---------------- resident arm ----------------
    for _ in range(args.warmup):
        g.run()
    sampler = ClockSampler(local)
    sampler.start()
    sync_all()
    t0 = time.perf_counter()
    dev_ms, launches, stage_acc = 0.0, 0, {}
    for _ in range(args.steps):
        g.run()
        ms, nl, stages = g.timing()
        dev_ms += ms
        launches += nl
        for name, v in stages:
            stage_acc[name] = stage_acc.get(name, 0.0) + v
    sync_all()
    wall = time.perf_counter() - t0
    # ---------------- e2e arm (host buffers through the C ABI) ----------------
    # results land in pinned host memory too (what a C++ host would allocate with cudaMallocHost): a pageable destination makes
    # the D2H copy go through the driver's bounce buffers and page-faults 29 MB per step
    rows_pin_t = torch.empty(int(nj + 16) * L.JUNCTION_DTYPE.itemsize, dtype=torch.uint8).pin_memory()
    rows_pin = rows_pin_t.numpy().view(L.JUNCTION_DTYPE)
    for _ in range(max(3, args.warmup)):          # the link and the arena need a few passes of their own on a fresh box
        g.shard_begin(n_rec, len(cols["cigar"]), len(cols["seq4"])); g.submit(pinned); g.run(); g.fetch(rows_pin)
    sync_all()
    t1 = time.perf_counter()
    for _ in range(args.steps):
        g.shard_begin(n_rec, len(cols["cigar"]), len(cols["seq4"]))
        g.submit(pinned)
        g.run()
        rows2, _ = g.fetch(rows_pin)
    sync_all()
    wall_e2e = time.perf_counter() - t1
    sampler.stop.set()
    sampler.join()
    assert len(rows2) == nj

    # max over ranks, sums of units
    tv = torch.tensor([wall, wall_e2e, dev_ms], dtype=torch.float64, device="cuda")
    uv = torch.tensor([n_spliced, n_rec, h2d_bytes, d2h_bytes, launches], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tv, op=dist.ReduceOp.MAX)
        dist.all_reduce(uv, op=dist.ReduceOp.SUM)
    wall_m, wall_e2e_m, dev_ms_m = [float(x) for x in tv.tolist()]
    tot_spliced, tot_rec, tot_h2d, tot_d2h, tot_launches = [float(x) for x in uv.tolist()]

    if rank == 0:
        total_alg, per_stage = algorithmic_bytes(cols, n_pairs, nj)
        stage_ms = {k: v / args.steps for k, v in stage_acc.items()}
        dom = max((k for k in stage_ms if k in per_stage), key=lambda k: stage_ms[k])
        peaks = {}
        try:
            with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
                peaks = json.load(f)
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        achieved = per_stage[dom] / (stage_ms[dom] * 1e-3) / 1e9
        traffic = None
        try:
            with open(os.path.join(ROOT, "profiles", "dominant_kernel_traffic.json")) as f:
                tj = json.load(f)
                traffic = tj.get(dom)
        except Exception:
            pass
        line = {
            "metric": METRIC, "value": tot_spliced * args.steps / wall_m, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": wall_m / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "int32", "data": "synthetic",
            "config": {"workload": "c2: synthetic 100 Mb genome (10 x 10 Mb), 10M 2x150 alignments per GPU (pjsynth preset %s scale %g)" % (args.preset, args.scale),
                       "records_per_gpu": n_rec, "spliced_per_gpu": n_spliced, "read_junction_pairs_per_gpu": n_pairs,
                       "junctions_per_gpu": int(nj), "l2": "inputs (%.2f GB of columns per GPU) are larger than the 126 MB L2" % (h2d_bytes / 1e9),
                       "sharding": "targets -> ranks, independent shards, no collective"},
            "device_ms_per_step": dev_ms_m / args.steps,
            "all_alignments_per_sec": tot_rec * args.steps / wall_m,
            "gpu_launches": int(tot_launches),
            "clocks": sampler.summary(),
            "e2e": {"value": tot_spliced * args.steps / wall_e2e_m, "unit": UNIT, "h2d_bytes_per_step": int(tot_h2d), "d2h_bytes_per_step": int(tot_d2h),
                    "ms_per_step": wall_e2e_m / args.steps * 1e3},
            "roofline": {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6650 GB/s (of fallback)",
                         "algorithmic_bytes_per_launch": per_stage[dom], "avg_launch_ms": stage_ms[dom],
                         "pipeline_algorithmic_bytes": total_alg, "pipeline_achieved_gbs": total_alg / (dev_ms_m / args.steps * 1e-3) / 1e9,
                         "stage_ms": {k: round(v, 4) for k, v in stage_ms.items()}},
            "setup": {"decode_s": round(t_decode, 3), "genome_and_first_run_s": round(t_setup, 3), "host_threads": cores},
        }
        if world == 1 and not args.no_bam:          # before the CPU baseline: the reference's runs leave the host busy with write-back
            line["e2e_bam"] = e2e_bam(prep, cores, n_spliced)
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(args, cores)
        if world == 1 and args.extra:
            g.close()
            line["extra_metrics"] = extra_leg(p, local, genomes, cores, max(2, min(args.steps, 5)))
        print(json.dumps(line))
    g.close()
    if world > 1:
        dist.destroy_process_group()


def extra_leg(p, device, genomes, cores, steps):
    """The `--extra` metrics on the bench workload: device time of pj_extra_run (CUDA events inside the
    library) and of the whole junc + extra step, inputs resident in pinned host memory."""
    import numpy as np
    from portcullis_b200 import junction_builder as jb
    cols = p.decode(-1, cores, names=True)
    n = len(cols["pos"])
    g = jb.JuncGpu(device, "UNKNOWN", extra=True)
    g.set_targets(p.lengths)
    for t, s in enumerate(genomes):
        g.set_genome(t, s)
    acc, stage_acc, junc_ms, cov_s = 0.0, {}, 0.0, 0.0
    for it in range(steps + 1):
        g.shard_begin(n, len(cols["cigar"]), len(cols["seq4"]))
        g.submit(cols)
        g.run()
        rows, st = g.fetch()
        x = g.extra_run(int(st["maxq"].max()))
        t0 = time.perf_counter()
        covered = np.array([g.target_pileup(t)[0] for t in range(len(p.lengths))], dtype=np.uint8)
        src = jb.coverage_source(covered)
        for t in np.unique(rows["tid"]):
            if src[t] >= 0:
                sel = np.nonzero(rows["tid"] == t)[0]
                x["cov_sum"][sel] = g.coverage(src[t], rows["start"][sel], rows["end"][sel])
        dt = time.perf_counter() - t0
        if it == 0:
            continue                                   # warm-up
        ms, nl, stages = g.extra_timing()
        acc += ms
        cov_s += dt
        junc_ms += g.timing()[0]
        for name, v in stages:
            stage_acc[name] = stage_acc.get(name, 0.0) + v
    g.close()
    x = jb.extra_finalize(x)
    return {"pj_extra_run_device_ms": round(acc / steps, 3), "launches": nl, "stage_ms": {k: round(v / steps, 4) for k, v in stage_acc.items()},
            "coverage_queries_wall_ms": round(cov_s / steps * 1e3, 3), "junc_pipeline_device_ms_in_extra_mode": round(junc_ms / steps, 3),
            "junctions": int(len(rows)), "unspliced_flank_total": int(x["up_aln"].astype(np.int64).sum() + x["down_aln"].astype(np.int64).sum()),
            "steps": steps}


def e2e_bam(prep, cores, n_spliced):
    """Whole front end from the BAM file: BGZF decode + genome load + GPU + finalize + writers, each run with a fresh library
    context.  One untimed run first (page cache, allocator pools), then three timed ones; the median is reported."""
    from portcullis_b200 import junction_builder as jb
    out = os.path.join(WORKDIR, "out_e2e", "p")
    runs = []
    for it in range(4):
        b = jb.JunctionBuilder(prep, out)
        b.setThreads(cores)
        t0 = time.perf_counter()
        rep = b.process()
        dt = time.perf_counter() - t0
        if it:
            runs.append((dt, rep))
    runs.sort(key=lambda r: r[0])
    dt, rep = runs[len(runs) // 2]
    return {"value": n_spliced / dt, "unit": UNIT, "seconds": round(dt, 3), "seconds_all": [round(r[0], 3) for r in runs], "host_threads": cores,
            "breakdown_s": {k: round(rep[k], 4) for k in ("t_open_s", "t_genome_s", "t_decode_s", "t_finalize_s", "t_write_s")},
            "gpu_pipeline_ms": round(rep["t_gpu_ms"], 3)}


def run_reference_once(prep, threads):
    """The unmodified reference `junc` (oracle/_ref/portcullis_ref) on the box's host cores; returns (seconds, spliced)."""
    out = os.path.join(WORKDIR, "out_ref", "r")
    t0 = time.perf_counter()
    pr = subprocess.run([REF_BIN, "junc", "-t", str(threads), "-o", out, prep], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    dt = time.perf_counter() - t0
    if pr.returncode != 0:
        raise RuntimeError("reference junc failed: " + pr.stderr[-500:])
    spliced, find_s = None, None
    for ln in pr.stdout.split("\n"):
        if "junctions from" in ln and "spliced alignments" in ln:
            spliced = int(ln.split("from")[1].split("spliced")[0])
        if "Wall time taken" in ln and find_s is None:
            find_s = float(ln.split(":")[1].strip().rstrip("s"))        # first timer = findJunctions
    return dt, spliced, find_s


def cpu_sample(args, cores):
    scale = args.scale * args.cpu_sample_frac
    prep, meta = make_workload(args.preset, scale, args.seed, cores)
    threads = min(cores, meta["n_targets"])                              # the reference caps threads at #targets
    return prep, meta, threads, scale


def cpu_baseline(args, cores):
    if not os.path.exists(REF_BIN):
        return {"value": None, "unit": UNIT, "cores": 0, "kind": "reference", "sample": "oracle/_ref/portcullis_ref missing"}
    prep, meta, threads, scale = cpu_sample(args, cores)
    run_reference_once(prep, threads)                                    # warm the page cache
    best = min((run_reference_once(prep, threads) for _ in range(2)), key=lambda r: r[0])
    dt, spliced, find_s = best
    return {"value": spliced / dt, "unit": UNIT, "cores": threads, "kind": "reference",
            "sample": "pjsynth preset %s scale %g (%d alignments, %d spliced): unmodified reference `junc -t %d`, whole-run wall %.2fs (findJunctions %.1fs), best of 2 warm"
                      % (args.preset, scale, meta["n_records"], spliced, threads, dt, find_s or -1)}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    cores = host_cores()
    if not os.path.exists(REF_BIN):
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/portcullis_ref was not built (needs /root/reference at build time)"}))
        return
    prep, meta, threads, scale = cpu_sample(args, cores)
    for _ in range(max(1, min(args.warmup, 1))):
        run_reference_once(prep, threads)
    t0 = time.perf_counter()
    spliced = 0
    for _ in range(args.steps):
        dt, sp, _ = run_reference_once(prep, threads)
        spliced += sp
    wall = time.perf_counter() - t0
    v = spliced / wall
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": wall / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
        "config": {"workload": "c2: synthetic 100 Mb genome (10 x 10 Mb), 10M 2x150 alignments per GPU (pjsynth preset %s scale %g)" % (args.preset, args.scale),
                   "step": "bounded sample: preset %s scale %g (%d alignments) through the unmodified reference `junc -t %d` from the BAM file"
                           % (args.preset, scale, meta["n_records"], threads)},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads, "kind": "reference",
                         "sample": "pjsynth preset %s scale %g, %d alignments per step" % (args.preset, scale, meta["n_records"])},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--preset", default="c2")
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--cpu-sample-frac", type=float, default=1.0, help="fraction of the workload the CPU reference is timed on (1.0 = the whole c2 workload, about 4.5 s per pass on 10 host threads)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-bam", action="store_true")
    ap.add_argument("--extra", action="store_true", help="also time the `--extra` metrics phase (pj_extra_run + coverage) on the same workload; adds an extra_metrics object")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
