#!/usr/bin/env python3
"""Full-size parity anchors for BASELINE configs 3, 4 and 5 (and 2): generate the preset with pjsynth, run the
UNMODIFIED reference `junc` (oracle/_ref/portcullis_ref) on it and record md5 sums of everything it wrote into
tests/golden/fullsize.json.  CPU only; run once in the build container (the reference needs 10-60 min per preset).

    python tests/golden/make_fullsize.py --preset c4 [--scale 1.0] [--threads 8] [--workdir /tmp/pj_full]

Per preset the record holds:
  * synth.json of the generated data and md5s of the BAM / BAI / FASTA (the generator must reproduce them on the GPU box),
  * md5 of the reference's junctions.tab / .bed / .intron.gff3 / .exon.gff3,
  * `tab_nofp_md5`: md5 of junctions.tab with the floating-point columns (compare.FP_COLS) blanked, the strict
    integer/string anchor,
  * `entropy_sum`, `entropy_head`: sum and the first values of the entropy column (tolerance anchor),
  * reference wall time and thread count (a reported CPU baseline on the build container's cores).
"""
import argparse
import hashlib
import json
import os
import shutil
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
OUT = os.path.join(ROOT, "tests", "golden", "fullsize.json")


def md5_file(path, chunk=1 << 24):
    h = hashlib.md5()
    with open(path, "rb") as f:
        while True:
            b = f.read(chunk)
            if not b:
                break
            h.update(b)
    return h.hexdigest()


def tab_anchors(path):
    """md5 of the tab with FP columns blanked + entropy statistics (see compare.FP_COLS)."""
    from compare import FP_COLS
    h = hashlib.md5()
    n = 0
    ent_sum = 0.0
    head = []
    with open(path) as f:
        for ln, line in enumerate(f):
            c = line.rstrip("\n").split("\t")
            if ln >= 1 and len(c) > 40:
                n += 1
                e = float(c[32])
                ent_sum += e
                if len(head) < 8:
                    head.append(c[32])
                for k in FP_COLS:
                    if k < len(c):
                        c[k] = "#"
            h.update(("\t".join(c) + "\n").encode())
    return h.hexdigest(), n, ent_sum, head


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--preset", required=True)
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--threads", type=int, default=os.cpu_count())
    ap.add_argument("--workdir", default="/tmp/pj_full")
    ap.add_argument("--keep", action="store_true")
    a = ap.parse_args()
    d = os.path.join(a.workdir, "%s_%g" % (a.preset, a.scale))
    shutil.rmtree(d, ignore_errors=True)
    os.makedirs(d)
    t0 = time.time()
    subprocess.check_call([os.path.join(ROOT, "portcullis_b200", "bin", "pjsynth"), "--preset", a.preset, "--scale", str(a.scale),
                           "--threads", str(a.threads), "--out", d + "/prep"])
    rec = {"preset": a.preset, "scale": a.scale, "generate_s": round(time.time() - t0, 1),
           "synth": json.load(open(d + "/prep/synth.json"))}
    for f in ("portcullis.sorted.alignments.bam", "portcullis.sorted.alignments.bam.bai", "portcullis.genome.fa", "portcullis.genome.fa.fai"):
        rec.setdefault("input_md5", {})[f] = md5_file(os.path.join(d, "prep", f))
        rec.setdefault("input_bytes", {})[f] = os.path.getsize(os.path.join(d, "prep", f))
    nt = min(a.threads, rec["synth"]["n_targets"])
    cmd = [os.path.join(ROOT, "oracle", "_ref", "portcullis_ref"), "junc", "-t", str(nt), "--exon_gff", "--intron_gff", "-o", d + "/ref/p", d + "/prep"]
    t0 = time.time()
    p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    rec["reference_s"] = round(time.time() - t0, 1)
    rec["reference_threads"] = nt
    rec["reference_host"] = "%d-core build container" % os.cpu_count()
    if p.returncode != 0:
        print("reference failed:", p.stderr[-2000:], file=sys.stderr)
        return 1
    rec["reference_stdout_tail"] = [l for l in p.stdout.split("\n") if l.strip()][-12:]
    rec["md5"] = {}
    for suffix in ("junctions.tab", "junctions.bed", "junctions.intron.gff3", "junctions.exon.gff3"):
        rec["md5"][suffix] = md5_file(d + "/ref/p." + suffix)
    nofp, n, es, head = tab_anchors(d + "/ref/p.junctions.tab")
    rec["tab_nofp_md5"] = nofp
    rec["junctions"] = n
    rec["entropy_sum"] = es
    rec["entropy_head"] = head
    allrec = {}
    if os.path.exists(OUT):
        allrec = json.load(open(OUT))
    allrec["%s@%g" % (a.preset, a.scale)] = rec
    with open(OUT + ".tmp", "w") as f:
        json.dump(allrec, f, indent=1, sort_keys=True)
        f.write("\n")
    os.replace(OUT + ".tmp", OUT)
    print(json.dumps(rec)[:600])
    if not a.keep:
        shutil.rmtree(d, ignore_errors=True)
    return 0


if __name__ == "__main__":
    sys.exit(main())
