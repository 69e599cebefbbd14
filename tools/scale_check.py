#!/usr/bin/env python3
"""Parity + wall-clock check at larger sizes, on the GPU box: generate a preset with pjsynth, run the unmodified
reference `junc` (oracle/_ref) and our `portcullis junc`, compare the four output files, print one JSON line.

    python tools/scale_check.py --preset c3 --scale 0.25 [--gpus 1] [--threads 16] [--orientation FR]
"""
import argparse
import filecmp
import json
import os
import shutil
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--preset", default="c3")
    ap.add_argument("--scale", type=float, default=0.25)
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--threads", type=int, default=os.cpu_count())
    ap.add_argument("--orientation", default=None)
    ap.add_argument("--workdir", default="/tmp/pj_scale")
    ap.add_argument("--skip-reference", action="store_true")
    ap.add_argument("--extra", action="store_true", help="run both sides with --extra (mm_score, coverage, up_aln, down_aln)")
    a = ap.parse_args()
    from compare import assert_exon_gff_equal, assert_tab_equal
    d = os.path.join(a.workdir, "%s_%g" % (a.preset, a.scale))
    shutil.rmtree(d, ignore_errors=True)
    t0 = time.time()
    subprocess.check_call([os.path.join(ROOT, "portcullis_b200", "bin", "pjsynth"), "--preset", a.preset, "--scale", str(a.scale),
                           "--threads", str(a.threads), "--out", d + "/prep"], stderr=subprocess.DEVNULL)
    t_gen = time.time() - t0
    meta = json.load(open(d + "/prep/synth.json"))
    out = {"preset": a.preset, "scale": a.scale, "records": meta["n_records"], "spliced": meta["n_spliced"], "pairs": meta["n_pairs"],
           "generate_s": round(t_gen, 1), "gpus": a.gpus, "host_threads": a.threads}
    extra = ["--orientation", a.orientation] if a.orientation else []
    if a.extra:
        extra.append("--extra")
        out["extra"] = True
    # the reference shells out to `samtools index` for --extra: oracle/samtools_shim answers with htslib-1.3's indexer
    ref_env = dict(os.environ, PATH=os.path.join(ROOT, "oracle", "samtools_shim") + os.pathsep + os.environ.get("PATH", ""))
    cmd = [os.path.join(ROOT, "portcullis_b200", "bin", "portcullis"), "junc", "-t", str(a.threads), "--gpus", str(a.gpus), "--exon_gff", "--intron_gff",
           "-o", d + "/ours/p"] + extra + [d + "/prep"]
    t0 = time.time()
    p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    out["ours_s"] = round(time.time() - t0, 2)
    if p.returncode != 0:
        out["ours_error"] = p.stderr[-400:]
        print(json.dumps(out)); return 1
    if a.extra:
        out["ours_warnings"] = [l for l in p.stderr.split("\n") if l.startswith("Warning")][:3]
    out["ours_runtime_line"] = [l.strip() for l in p.stdout.split("\n") if "Total runtime" in l][-1]
    out["ours_spliced_per_s"] = round(meta["n_spliced"] / out["ours_s"])
    if not a.skip_reference:
        nt = min(a.threads, meta["n_targets"])
        cmd = [os.path.join(ROOT, "oracle", "_ref", "portcullis_ref"), "junc", "-t", str(nt), "--exon_gff", "--intron_gff", "-o", d + "/ref/p"] + extra + [d + "/prep"]
        t0 = time.time()
        p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, env=ref_env)
        out["reference_s"] = round(time.time() - t0, 2)
        out["reference_threads"] = nt
        if p.returncode != 0:
            out["reference_error"] = p.stderr[-400:]
            print(json.dumps(out)); return 1
        out["reference_spliced_per_s"] = round(meta["n_spliced"] / out["reference_s"])
        out["speedup_wall"] = round(out["reference_s"] / out["ours_s"], 1)
        try:
            assert_tab_equal(d + "/ours/p.junctions.tab", d + "/ref/p.junctions.tab")
            assert filecmp.cmp(d + "/ours/p.junctions.bed", d + "/ref/p.junctions.bed", shallow=False), "bed differs"
            assert filecmp.cmp(d + "/ours/p.junctions.intron.gff3", d + "/ref/p.junctions.intron.gff3", shallow=False), "intron gff differs"
            assert_exon_gff_equal(d + "/ours/p.junctions.exon.gff3", d + "/ref/p.junctions.exon.gff3")
            out["parity"] = "ok"
            out["tab_byte_identical"] = filecmp.cmp(d + "/ours/p.junctions.tab", d + "/ref/p.junctions.tab", shallow=False)
        except AssertionError as e:
            out["parity"] = "FAILED: %s" % str(e)[:300]
        out["junctions"] = sum(1 for _ in open(d + "/ref/p.junctions.tab")) - 2
    print(json.dumps(out))
    shutil.rmtree(d, ignore_errors=True)
    return 0


if __name__ == "__main__":
    sys.exit(main())
