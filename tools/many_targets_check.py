#!/usr/bin/env python3
"""Many small targets (transcriptome-like headers): `junc --extra` on 1 and 2 contexts against the reference, with timings.
    python tools/many_targets_check.py 3000        (GPU box; needs oracle/_ref)"""
import os
import sys
import time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import synth, refrun
from compare import assert_tab_equal
from portcullis_b200 import junction_builder as jb
n = int(sys.argv[1])
t0 = time.time()
ds = synth.make_dataset(77, n_targets=n, target_len=3000, genes_per_target=1, reads_per_gene=(5, 25), unspliced_frac=0.5, multimap_frac=0.05)
print("generated %d records on %d targets in %.1fs" % (len(ds["records"]), n, time.time() - t0))
prep = refrun.make_prep_dir(ds, "/tmp/mt/w")
t0 = time.time(); refrun.run_reference(prep, "/tmp/mt/ref/r", threads=8, extra=True, exon_gff=False, intron_gff=False); tr = time.time() - t0
for gpus in (1, 2):
    b = jb.JunctionBuilder(prep, "/tmp/mt/o%d/p" % gpus); b.setThreads(8); b.setExtra(True); b.setGpus(gpus); b.gpu_ids = [0] * gpus
    t0 = time.time(); rep = b.process(); to = time.time() - t0
    assert_tab_equal("/tmp/mt/o%d/p.junctions.tab" % gpus, "/tmp/mt/ref/r.junctions.tab")
    print("gpus %d: ours %.2fs (extra %.2fs, decode %.2fs, genome %.2fs, init %.2fs) reference %.2fs, %d junctions, tab equal" % (gpus, to, rep["t_extra_s"], rep["t_decode_s"], rep["t_genome_s"], rep["t_init_s"], tr, rep["n_junctions"]))
