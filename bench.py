#!/usr/bin/env python3
"""Benchmark of the junc hot path (BASELINE.json metric: spliced alignments / second through junc at 1/2/4/8 B200).

    python bench.py --gpus N --steps K --warmup W            # our CUDA path (torchrun launches one rank per GPU for N > 1)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's own CPU implementation (oracle/_ref)
    python bench.py --preset c2|c4|c5 [--scale F]            # the other named shapes (committed under profiles/)

Workload (default): BASELINE config 3, the one the metric is quoted on — synthetic human-scale genome (24 targets with the
GRCh38 lengths, 3.1 Gb) and 200 M 2x150 alignments, made on the box by portcullis_b200/bin/pjsynth; it fits one B200.
The WHOLE job is sharded over the N ranks by the product's own range plan (pjh_plan_*: contiguous record-balanced
ranges of the BAM, cut inside a target only where no spliced read spans) — strong scaling, no data-path collective.
One "step" = one pass of the hot path over every rank's part of the job.

 value   : spliced alignments / s, whole job: every rank's alignment columns and packed genome already resident in HBM,
           pj_shard_run per step (CUDA-event device time reported beside the wall time)
 e2e     : same metric through the C ABI (pj_shard_begin / pj_batch_submit / pj_shard_run / pj_shard_fetch) with the
           decoded columns in pinned HOST memory: H2D of every column and D2H of the junction rows inside the timed region
 e2e_bam : same metric from the BAM FILE to the output files (BGZF decode, genome load, GPU, gather on rank 0, A12/A13,
           writers) — the like-for-like counterpart of the reference arm, which also starts at the BAM file
"""
import argparse
import hashlib
import json
import os
import shutil
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

PJSYNTH = os.path.join(ROOT, "portcullis_b200", "bin", "pjsynth")
REF_BIN = os.path.join(ROOT, "oracle", "_ref", "portcullis_ref")
WORKDIR = os.environ.get("PJ_BENCH_DIR", "/tmp/pj_bench")
METRIC = "spliced_alignments_per_sec"
UNIT = "spliced alignments/s"

WORKLOADS = {
    "c2": "c2: synthetic 100 Mb genome (10 x 10 Mb), 10M 2x150 spliced alignments",
    "c3": "c3: synthetic human-scale 3.1 Gb genome (24 targets, GRCh38 lengths), 200M 2x150 alignments, sharded over the GPUs",
    "c4": "c4: deep skewed coverage, 8 hot loci with 1-4M reads per junction, heavy multi-mapping (100 Mb genome)",
    "c5": "c5: long-read-style alignments (1-10 kb, up to 20 introns per read, indels near splice sites; 100 Mb genome)",
}
# fraction of the workload the CPU reference is timed on (about 10-30 s of CPU work per pass)
CPU_SAMPLE = {"c2": 1.0, "c3": 0.05, "c4": 0.1, "c5": 0.05}


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def bind_to_gpu_numa(gpu_index):
    """One process per GPU: keep this rank's threads and memory (decoded columns, pinned staging) on the NUMA node its GPU hangs off,
    like `numactl --cpunodebind --preferred` would (round 1's 8-GPU e2e was limited by cross-socket host traffic).  Reads the
    CPU / NUMA affinity columns of `nvidia-smi topo -m`; does nothing when the box does not expose them."""
    info = {"node": None, "cpus": None}
    try:
        import ctypes
        import re
        out = subprocess.run(["nvidia-smi", "topo", "-m"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True, timeout=20).stdout
        out = re.sub(r"\x1b\[[0-9;]*m", "", out)
        lines = [l for l in out.split("\n") if l.strip()]
        hdr = next(l for l in lines if "CPU Affinity" in l).split("\t")
        row = next(l for l in lines if l.startswith("GPU%d\t" % gpu_index) or l.startswith("GPU%d " % gpu_index)).split("\t")
        ci, ni = hdr.index("CPU Affinity"), hdr.index("NUMA Affinity")
        cpus = set()
        for part in row[ci].strip().split(","):
            if "-" in part:
                a, b = part.split("-"); cpus.update(range(int(a), int(b) + 1))
            elif part.strip().isdigit():
                cpus.add(int(part))
        allowed = os.sched_getaffinity(0)
        use = cpus & allowed
        if use and use != allowed:
            os.sched_setaffinity(0, use)
            info["cpus"] = "%d of %d allowed" % (len(use), len(allowed))
        node = row[ni].strip().split(",")[0].split("-")[0]
        if node.isdigit():
            mask = ctypes.c_ulong(1 << int(node))
            rc = ctypes.CDLL(None, use_errno=True).syscall(238, 1, ctypes.byref(mask), 64)      # set_mempolicy(MPOL_PREFERRED, {node})
            if rc == 0:
                info["node"] = int(node)
    except Exception:
        pass
    return info


def config_of(args):
    """Identical in both arms (the driver compares them)."""
    return {"workload": "%s (pjsynth preset %s scale %g)" % (WORKLOADS[args.preset], args.preset, args.scale),
            "preset": args.preset, "scale": args.scale,
            "sharding": "whole job over the ranks: contiguous record-balanced ranges of the BAM, cut inside a target only where no spliced read spans; no collective",
            "l2": "inputs (GBs of columns per GPU) are larger than the 126 MB L2"}


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


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe).  ONE long-running
    `nvidia-smi -lms` process writes a row per GPU every 250 ms; nothing is forked from this process while steps are being timed
    (a Python thread that spawned nvidia-smi per sample held the GIL through each spawn and showed up as 30-60 ms stalls of the
    rank that ran it)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index                            # None: every GPU of the box
        self.rows = []
        self.proc = None
        self.path = os.path.join(WORKDIR, "clocks_%d.csv" % os.getpid())

    def start(self):
        try:
            os.makedirs(WORKDIR, exist_ok=True)
            self.out = open(self.path, "w")
            cmd = ["nvidia-smi"] + (["-i", str(self.gpu)] if self.gpu is not None else []) + \
                  ["--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", os.environ.get("PJ_BENCH_SMI_MS", "250")]
            self.proc = subprocess.Popen(cmd, stdout=self.out, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def n_rows(self):
        try:
            return sum(1 for _ in open(self.path))
        except Exception:
            return 0

    def stop_and_collect(self):
        if self.proc is not None:
            self.proc.terminate()                       # the exact PID we started
            try:
                self.proc.wait(timeout=10)
            except Exception:
                self.proc.kill()
            self.out.close()
        try:
            with open(self.path) as f:
                self.rows = [[c.strip() for c in line.split(",")] for line in f if line.strip()]
            os.remove(self.path)
        except Exception:
            pass

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


def algorithmic_bytes(n_rec, n_cig, seq_bytes, n_cig_spliced, n_pairs, n_junc):
    """SURVEY §8(d), literally: per record 32 B of fixed columns + 4 B per CIGAR op + the 4-bit SEQ of spliced records +
    48 B per read-junction pair (a 24-B pair record = 8-B key + 16-B payload, written once and read once) + ceil(a/4) bytes
    of 2-bit genome under the anchor bases (a = read bases of spliced records = 2 * seq_bytes); per junction a 320-B row + 6 B of
    motif reads.  Sort traffic is NOT algorithmic.  Per stage: the bytes of that formula the stage has to move (24-B pair
    records; the sort and the segmentation are charged one read + one write of what they permute / label, which the
    pipeline total does not contain — they are overhead by §8(d))."""
    genome = seq_bytes // 2
    total = 32 * n_rec + 4 * n_cig + seq_bytes + 48 * n_pairs + genome + 326 * n_junc
    per_stage = {
        "scan_emit": 32 * n_rec + 4 * n_cig + 24 * n_pairs,            # every column + CIGAR once, pair records written
        "radix_sort": 2 * 24 * n_pairs,                                 # one permutation of the pair records (overhead stage)
        "segments": (8 + 4) * n_pairs,                                  # key read, junction id written (overhead stage)
        "reduce1": 24 * n_pairs + 64 * n_junc,
        "entropy": 8 * n_pairs + 8 * n_junc,
        "match": (24 + 16) * n_pairs + 4 * n_cig_spliced + seq_bytes + genome,   # pair record, result, CIGAR + SEQ of the read, 2-bit genome window
        "reduce2": 16 * n_pairs + 100 * n_junc,
        "finalize": 326 * n_junc,
    }
    return total, per_stage


def md5_file(path, chunk=1 << 24):
    h = hashlib.md5()
    with open(path, "rb") as f:
        while True:
            b = f.read(chunk)
            if not b:
                break
            h.update(b)
    return h.hexdigest()


def golden_md5(preset, scale):
    try:
        with open(os.path.join(ROOT, "tests", "golden", "fullsize.json")) as f:
            return json.load(f).get("%s@%g" % (preset, scale))
    except Exception:
        return None


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
        # NCCL prints its version banner on stdout when the communicator is created: keep stdout for the ONE JSON line
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_stdout, 1)
            os.close(saved_stdout)
    all_cores = host_cores()
    cores = max(1, all_cores // world)
    numa = bind_to_gpu_numa(local) if world > 1 else {"node": None, "cpus": None}
    if numa.get("cpus"):
        cores = max(1, min(cores, host_cores()))

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # ---------------- workload: made once per box, by rank 0 ----------------
    t0 = time.time()
    if rank == 0:
        make_workload(args.preset, args.scale, args.seed, all_cores)
    barrier()
    prep, meta = make_workload(args.preset, args.scale, args.seed, all_cores)
    t_gen = time.time() - t0

    # ---------------- this rank's part of the job, decoded once into host memory ----------------
    t0 = time.time()
    p = jb.PrepDir(prep)
    T = len(p.names)
    HUGE = 1 << 40                                     # one segment per part: the whole part is resident for the `value` arm
    seg, cuts = p.plan(world, HUGE)
    # the LEAN batch form, exactly what the product's driver ships over PCIe (pj_batch.lean): one batch per target stretch
    from portcullis_b200.columnar import lean_nbytes
    runs, whole = (p.decode_segment_lean(world, rank, 0, HUGE, threads=cores, keep_mate=False, copy=False, with_whole=True) if int(seg[rank]) > 0 else ([], {}))
    t_decode = time.time() - t0
    n_rec = int(sum(len(r["pos"]) for r in runs))
    cudart = torch.cuda.cudart()
    registered = []
    for k, v in whole.items():                         # page-lock the decoded columns in place (no second copy of the shard)
        if v.nbytes:
            rc = cudart.cudaHostRegister(v.ctypes.data, v.nbytes, 0)
            try:
                ok = int(rc) == 0
            except Exception:
                ok = str(rc).lower().endswith("success")
            if ok:
                registered.append(v.ctypes.data)
    h2d_bytes = int(sum(lean_nbytes(r) for r in runs))
    my_targets = [r["tid"] for r in runs]
    n_cig_all = int(sum(len(r["cigar"]) for r in runs))
    seq2_all = int(sum(len(r["seq2"]) for r in runs))

    def submit_all():
        g.shard_begin(n_rec, n_cig_all, 2 * seq2_all)
        for r in runs:
            g.submit_lean(r)

    g = jb.JuncGpu(local, "UNKNOWN")
    g.set_targets(p.lengths)
    t0 = time.time()
    for t in my_targets:
        g.set_genome(t, p.genome(t))
    submit_all()
    nj = g.run()                                       # also finishes the genome upload
    t_setup = time.time() - t0
    rows, st = g.fetch()
    n_spliced = int(st["spliced"].sum())
    n_pairs = int(rows["nb_raw_aln"].astype(np.int64).sum())
    d2h_bytes = int(rows.nbytes + 32 * T)
    del rows

    # ---------------- resident arm ----------------
    # The clock sampler (one `nvidia-smi -lms` process, the recipe's line) starts BEFORE the warm-up, so that its NVML start-up is over
    # when the timed region begins; steady sampling does not move the step time (tools/smi_probe.py: 1.818 ms with and without).
    # The warm-up runs for at least W steps AND at least --warmup-seconds of continuous load: for the first ~100 ms of work after the
    # host-side set-up (seconds of decode with an idle GPU) single stages stall by 2-30 ms whatever else is running (per-step trace in
    # profiles/r2_history.md), which 5 steps of a 2-7 ms pipeline do not cover.  The count actually run is reported as `warmup_steps_run`.
    sampler = ClockSampler(local if world == 1 else None)
    if rank == 0:
        sampler.start()
    warm_run, t_warm = 0, time.perf_counter()
    while warm_run < args.warmup or time.perf_counter() - t_warm < args.warmup_seconds:
        g.run()
        warm_run += 1
    barrier()
    t0 = time.perf_counter()
    dev_ms, launches, stage_acc = 0.0, 0, {}
    trace = []
    for _ in range(args.steps):
        ts = time.perf_counter()
        g.run()
        ms, nl, stages = g.timing()
        trace.append((round((time.perf_counter() - ts) * 1e3, 3), round(ms, 3), max(stages, key=lambda kv: kv[1])[0]))
        dev_ms += ms
        launches += nl
        for name, v in stages:
            stage_acc[name] = stage_acc.get(name, 0.0) + v
    barrier()
    wall = time.perf_counter() - t0
    if os.environ.get("PJ_BENCH_TRACE"):
        print("rank %d steps (wall ms, device ms, longest stage): %s" % (rank, trace), file=sys.stderr)
    # ---------------- e2e arm (host buffers through the C ABI) ----------------
    wall_e2e = 0.0
    e2e_steps = args.steps if not args.resident_only else 0
    if e2e_steps:
        # results land in pinned host memory too (what a C++ host would allocate with cudaMallocHost)
        rows_pin_t = torch.empty(int(nj + 16) * L.JUNCTION_DTYPE.itemsize, dtype=torch.uint8).pin_memory()
        rows_pin = rows_pin_t.numpy().view(L.JUNCTION_DTYPE)
        for _ in range(3):                             # the link and the arena need a few passes of their own on a fresh box
            submit_all(); g.run(); g.fetch(rows_pin)
        barrier()
        t1 = time.perf_counter()
        for _ in range(e2e_steps):
            submit_all()
            g.run()
            rows2, _ = g.fetch(rows_pin)
        barrier()
        wall_e2e = time.perf_counter() - t1
        assert len(rows2) == nj
    if rank == 0:
        sampler.stop_and_collect()

    # per-rank workload numbers for the roofline (before the columns are released)
    n_cig, n_cig_spliced, seq_bytes = n_cig_all, 0, 0                # seq_bytes: the 4-bit SEQ bytes of SURVEY 8(d), (l + 1) / 2 per spliced record
    for r in runs:
        off = np.concatenate([[0], np.cumsum(r["n_cigar"].astype(np.int64))])
        isn = np.concatenate([[0], np.cumsum((r["cigar"] & 15) == 3)])
        spl = (isn[off[1:]] - isn[off[:-1]]) > 0
        n_cig_spliced += int(r["n_cigar"].astype(np.int64)[spl].sum())
        lq = r["l_qseq"].astype(np.int64)[spl]
        seq_bytes += int(((lq[lq > 0] + 1) // 2).sum())
    g.close()
    for ptr in registered:
        cudart.cudaHostUnregister(ptr)
    del runs, whole
    p.close()

    # ---------------- e2e_bam arm: BAM file -> output files, the product's own driver ----------------
    e2e_bam = None
    if not args.no_bam and not args.resident_only:
        e2e_bam = e2e_bam_leg(args, prep, rank, world, local, cores, barrier, np, jb)

    # max over ranks, sums of units
    tv = torch.tensor([wall, wall_e2e, dev_ms], dtype=torch.float64, device="cuda")
    uv = torch.tensor([n_spliced, n_rec, h2d_bytes, d2h_bytes, launches, n_pairs, nj], dtype=torch.float64, device="cuda")
    per_rank = torch.zeros(world, dtype=torch.float64, device="cuda")
    per_rank[rank] = dev_ms / max(args.steps, 1)
    if world > 1:
        dist.all_reduce(tv, op=dist.ReduceOp.MAX)
        dist.all_reduce(uv, op=dist.ReduceOp.SUM)
        dist.all_reduce(per_rank, op=dist.ReduceOp.SUM)
    wall_m, wall_e2e_m, dev_ms_m = [float(x) for x in tv.tolist()]
    tot_spliced, tot_rec, tot_h2d, tot_d2h, tot_launches, tot_pairs, tot_junc = [float(x) for x in uv.tolist()]

    if rank == 0:
        total_alg, per_stage = algorithmic_bytes(n_rec, n_cig, seq_bytes, n_cig_spliced, n_pairs, nj)     # rank 0's part
        stage_ms = {k: v / args.steps for k, v in stage_acc.items()}
        peaks = {}
        try:
            with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
                peaks = json.load(f)
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        dom = max((k for k in stage_ms if k in per_stage), key=lambda k: stage_ms[k])
        achieved = per_stage[dom] / (stage_ms[dom] * 1e-3) / 1e9
        traffic = None
        try:
            with open(os.path.join(ROOT, "profiles", "dominant_kernel_traffic.json")) as f:
                tj = json.load(f)
                traffic = tj.get(args.preset, {}).get(dom) if isinstance(tj.get(args.preset), dict) else None
        except Exception:
            pass
        stage_table = {k: {"ms": round(stage_ms[k], 4), "alg_bytes": int(per_stage[k]), "gbs": round(per_stage[k] / (stage_ms[k] * 1e-3) / 1e9, 1),
                           "frac": round(per_stage[k] / (stage_ms[k] * 1e-3) / 1e9 / peak, 4)} for k in stage_ms if k in per_stage and stage_ms[k] > 0}
        pipe_ms = dev_ms / args.steps
        cfg = config_of(args)
        line = {
            "metric": METRIC, "value": tot_spliced * args.steps / wall_m, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "warmup_steps_run": warm_run, "ms_per_step": wall_m / args.steps * 1e3, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "int32", "data": "synthetic", "config": cfg,
            "workload_stats": {"records": int(tot_rec), "spliced": int(tot_spliced), "read_junction_pairs": int(tot_pairs), "junctions": int(tot_junc),
                               "records_rank0": n_rec, "gap_cuts_in_plan": int(cuts), "host_cores": all_cores, "decode_threads_per_rank": cores, "numa_rank0": numa},
            "device_ms_per_step": dev_ms_m / args.steps,
            "device_ms_per_step_by_rank": [round(float(x), 4) for x in per_rank.tolist()],
            "all_alignments_per_sec": tot_rec * args.steps / wall_m,
            "gpu_launches": int(tot_launches),
            "clocks": sampler.summary(),
            "e2e": ({"value": tot_spliced * e2e_steps / wall_e2e_m, "unit": UNIT, "h2d_bytes_per_step": int(tot_h2d), "d2h_bytes_per_step": int(tot_d2h),
                     "ms_per_step": wall_e2e_m / e2e_steps * 1e3,
                     "what": "C ABI with the decoded columns (lean batch form: 2-bit SEQ, per-record op counts) in pinned host memory: pj_shard_begin / pj_batch_submit / pj_shard_run / pj_shard_fetch"} if e2e_steps else None),
            "roofline": {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6650 GB/s (of fallback)",
                         "algorithmic_bytes_per_launch": per_stage[dom], "avg_launch_ms": stage_ms[dom],
                         "bytes_model": "SURVEY 8(d) literally: 32 B/record + 4 B/CIGAR op + 4-bit SEQ of spliced records + 48 B/pair (24-B record written once, read once) + ceil(a/4) B of 2-bit genome + 326 B/junction",
                         "pipeline_algorithmic_bytes": total_alg, "pipeline_ms": pipe_ms, "pipeline_achieved_gbs": total_alg / (pipe_ms * 1e-3) / 1e9,
                         "pipeline_frac": total_alg / (pipe_ms * 1e-3) / 1e9 / peak, "scope": "rank 0's part of the job",
                         "stages": stage_table},
            "setup": {"generate_s": round(t_gen, 1), "decode_s": round(t_decode, 3), "genome_and_first_run_s": round(t_setup, 3), "host_threads": cores},
        }
        if e2e_bam is not None:
            line["e2e_bam"] = e2e_bam
        if world == 1 and not args.no_cpu_baseline and not args.resident_only:
            line["cpu_baseline"] = cpu_baseline(args, all_cores)
            if e2e_bam is not None and line["cpu_baseline"].get("value"):
                e2e_bam["ratio_vs_cpu_baseline"] = round(e2e_bam["value"] / line["cpu_baseline"]["value"], 2)
                e2e_bam["ratio_note"] = "both arms start at the BAM file and end with the output files written; the CPU arm runs on a bounded sample (see cpu_baseline.sample)"
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def e2e_bam_leg(args, prep, rank, world, local, cores, barrier, np, jb):
    """BAM file -> output files through the product's driver.  N = 1: JunctionBuilder.process() (the call a user makes).
    N > 1: one process per GPU — every rank decodes and runs its part of the range plan (process_part), drops its rows into
    /dev/shm, rank 0 concatenates them in part order and runs finish() (A12/A13 + writers).  The only synchronisation is
    the barrier; no collective touches the data.  One untimed pass first, then three timed ones (median)."""
    out = os.path.join(WORKDIR, "out_e2e_%s" % args.preset, "p")
    shm = "/dev/shm/pj_bench_rows_%d" % os.getuid()
    if rank == 0:
        shutil.rmtree(shm, ignore_errors=True)
        os.makedirs(shm, exist_ok=True)
        os.makedirs(os.path.dirname(out), exist_ok=True)
    barrier()
    runs = []
    reps = []
    n_spliced = 0
    for it in range(args.bam_passes + 1):
        b = jb.JunctionBuilder(prep, out)
        b.setThreads(cores)
        barrier()
        t0 = time.perf_counter()
        if world == 1:
            rep = b.process()
            n_spliced = rep["n_spliced"]
        else:
            rows, stats, rep = b.process_part(rank, world, device=local)
            rows.tofile(os.path.join(shm, "rows_%d.bin" % rank))
            stats.tofile(os.path.join(shm, "stats_%d.bin" % rank))
            barrier()
            if rank == 0:
                from portcullis_b200 import _lib as L
                sizes = [os.path.getsize(os.path.join(shm, "rows_%d.bin" % r)) // L.JUNCTION_DTYPE.itemsize for r in range(world)]
                allrows = np.empty(sum(sizes), dtype=L.JUNCTION_DTYPE)            # parts land in place, in part order: one copy
                at = 0
                for r, k in enumerate(sizes):
                    if k:
                        with open(os.path.join(shm, "rows_%d.bin" % r), "rb") as fh:
                            fh.readinto(allrows[at:at + k].view(np.uint8).reshape(-1))
                    at += k
                allstats = jb.merge_target_stats([np.fromfile(os.path.join(shm, "stats_%d.bin" % r), dtype=jb.TARGET_STATS_DTYPE) for r in range(world)])
                _, frep = b.finish(allrows, allstats)
                n_spliced = frep["n_spliced"]
                rep = dict(rep, t_finalize_s=frep["t_finalize_s"], t_write_s=frep["t_write_s"], n_junctions=frep["n_junctions"])
        barrier()
        dt = time.perf_counter() - t0
        if it:
            runs.append(dt)
            reps.append(rep)
    if rank != 0:
        return None
    order = sorted(range(len(runs)), key=lambda i: runs[i])
    mid = order[len(order) // 2]
    dt, rep = runs[mid], reps[mid]
    res = {"value": n_spliced / dt, "unit": UNIT, "seconds": round(dt, 3), "seconds_all": [round(r, 3) for r in runs], "host_threads_per_rank": cores,
           "breakdown_s_rank0": {k: round(rep[k], 4) for k in ("t_open_s", "t_init_s", "t_genome_s", "t_decode_s", "t_run_s", "t_teardown_s", "t_finalize_s", "t_write_s", "t_total_s")},
           "gpu_pipeline_ms_rank0": round(rep["t_gpu_ms"], 3), "segments_rank0": rep.get("n_segments"), "junctions": rep.get("n_junctions"),
           "what": "BAM file -> junctions.tab/.bed written; BGZF decode on the host cores, %d GPU(s)" % world}
    gold = golden_md5(args.preset, args.scale)
    if gold and args.seed == 0:
        got = md5_file(out + ".junctions.tab")
        res["parity"] = {"junctions_tab_md5": got, "equals_reference_md5": got == gold["md5"]["junctions.tab"],
                         "reference": "unmodified reference junc on the same prep directory (tests/golden/fullsize.json)"}
    shutil.rmtree(shm, ignore_errors=True)
    return res


def run_reference_once(prep, threads, tag="r"):
    """The unmodified reference `junc` (oracle/_ref/portcullis_ref) on the box's host cores; returns (seconds, spliced)."""
    out = os.path.join(WORKDIR, "out_ref", tag)
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
    frac = args.cpu_sample_frac if args.cpu_sample_frac > 0 else CPU_SAMPLE.get(args.preset, 0.1)
    scale = args.scale * frac
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
            "sample": "pjsynth preset %s scale %g (%d alignments, %d spliced): unmodified reference `junc -t %d` from the BAM file, whole-run wall %.2fs (findJunctions %.1fs), best of 2 warm"
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
    sample = ("bounded sample of the workload: pjsynth preset %s scale %g (%d alignments, %d spliced) per step through the unmodified reference `junc -t %d`, BAM file -> output files, on %d host cores"
              % (args.preset, scale, meta["n_records"], meta["n_spliced"], threads, cores))
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": wall / args.steps * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
        "config": config_of(args),
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads, "kind": "reference", "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--warmup-seconds", type=float, default=1.0, help="the warm-up also lasts at least this long (continuous load before the timed region)")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--preset", default="c3", choices=sorted(WORKLOADS))
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--cpu-sample-frac", type=float, default=0.0, help="fraction of the workload the CPU reference is timed on (0 = per-preset default: about 10-30 s of CPU work)")
    ap.add_argument("--bam-passes", type=int, default=3, help="timed passes of the BAM-file arm (after one untimed pass)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-bam", action="store_true")
    ap.add_argument("--resident-only", action="store_true", help="only the resident arm (profiling runs)")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
