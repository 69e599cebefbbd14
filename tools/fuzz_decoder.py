#!/usr/bin/env python3
"""Corrupted-input fuzzing of the host readers (no GPU needed): random byte flips and truncations of a fixture's BAM or BAI, then
open + plan + decode (with names) + `--separate`.  Every case must either decode or be rejected with a PjError — never crash.
    python tools/fuzz_decoder.py bam|bai <first seed> <cases>"""
import os
import random
import shutil
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from portcullis_b200 import junction_builder as jb, _lib as L
src = os.path.join(ROOT, "tests", "golden", "short_pe")
work = "/tmp/fuzz_dec_%d" % os.getpid(); shutil.rmtree(work, ignore_errors=True); os.makedirs(work)
bam = open(src + "/reads.bam", "rb").read(); bai = open(src + "/reads.bam.bai", "rb").read()
mode = sys.argv[1]; seed0 = int(sys.argv[2]); n = int(sys.argv[3])
ok = err = 0
for seed in range(seed0, seed0 + n):
    rng = random.Random(seed)
    b = bytearray(bam); x = bytearray(bai)
    tgt = b if mode == "bam" else x
    k = rng.choice([1, 1, 2, 5, 20])
    for _ in range(k):
        i = rng.randrange(len(tgt)); tgt[i] = rng.randrange(256)
    if rng.random() < 0.2:
        del tgt[rng.randrange(len(tgt)):]
    d = os.path.join(work, "p"); shutil.rmtree(d, ignore_errors=True); os.makedirs(d)
    open(d + "/portcullis.sorted.alignments.bam", "wb").write(bytes(b)); open(d + "/portcullis.sorted.alignments.bam.bai", "wb").write(bytes(x))
    os.symlink(src + "/genome.fa", d + "/portcullis.genome.fa"); os.symlink(src + "/genome.fa.fai", d + "/portcullis.genome.fa.fai")
    try:
        p = jb.PrepDir(d); p.decode(-1, 2, names=True); p.plan_shards(2)
        jb.separate_bams(d, work + "/s/p", threads=2)
        ok += 1
    except L.PjError:
        err += 1
print("mode", mode, "cases", n, "decoded", ok, "rejected", err)
