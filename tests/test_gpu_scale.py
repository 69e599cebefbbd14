"""GPU tests at larger sizes.

* The four synthetic presets of BASELINE.json (c2 short paired, c3 human-like with N runs and soft-masking, c4 hot
  junctions / heavy multi-mapping, c5 long reads with indels) at sizes the reference binary finishes in seconds:
  our `junc` front end must reproduce the files of the unmodified reference `junc` run on the same prep directory.
* The full c2 workload (10 M alignments) through size-independent properties.
"""
import filecmp
import json
import os
import subprocess

import numpy as np
import pytest

import oracle_binding as ob
import refrun
from compare import assert_exon_gff_equal, assert_tab_equal
from portcullis_b200 import junction_builder as jb

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PJSYNTH = os.path.join(ROOT, "portcullis_b200", "bin", "pjsynth")


def synth(tmpdir, preset, scale, seed=0):
    d = os.path.join(str(tmpdir), "%s_%g" % (preset, scale))
    subprocess.check_call([PJSYNTH, "--preset", preset, "--scale", str(scale), "--seed", str(seed), "--out", d], stderr=subprocess.DEVNULL)
    with open(os.path.join(d, "synth.json")) as f:
        return d, json.load(f)


@pytest.mark.parametrize("preset,scale,orient,gpus", [("c2", 0.05, None, 1), ("c2", 0.03, "FR", 1), ("c3", 0.004, None, 1),
                                                      ("c4", 0.05, None, 1), ("c5", 0.03, None, 1),
                                                      ("c2", 0.05, None, 2), ("c3", 0.004, "RF", 4)])
def test_presets_reproduce_reference_files(tmp_path, preset, scale, orient, gpus):
    import torch
    if not os.path.exists(ob.REF_BIN):
        pytest.skip("oracle/_ref/portcullis_ref not built")
    if torch.cuda.device_count() < gpus:
        pytest.skip("needs %d GPUs" % gpus)
    prep, meta = synth(tmp_path, preset, scale)
    ref_prefix = os.path.join(str(tmp_path), "ref", "r")
    refrun.run_reference(prep, ref_prefix, threads=min(8, meta["n_targets"]), orientation=orient)
    out = os.path.join(str(tmp_path), "ours", "o")
    b = jb.JunctionBuilder(prep, out)
    b.setThreads(8)
    b.setGpus(gpus)
    b.setOutputExonGFF(True)
    b.setOutputIntronGFF(True)
    if orient:
        b.setOrientation(orient)
    rep = b.process()
    assert rep["n_spliced"] == meta["n_spliced"] and rep["n_spliced"] + rep["n_unspliced"] == meta["n_records"]
    assert_tab_equal(out + ".junctions.tab", ref_prefix + ".junctions.tab")
    assert filecmp.cmp(out + ".junctions.bed", ref_prefix + ".junctions.bed", shallow=False)
    assert filecmp.cmp(out + ".junctions.intron.gff3", ref_prefix + ".junctions.intron.gff3", shallow=False)
    assert_exon_gff_equal(out + ".junctions.exon.gff3", ref_prefix + ".junctions.exon.gff3")


def test_full_c2_properties(tmp_path):
    """10 M alignments: counts add up, order is (tid,start,end), results are deterministic and independent of batching."""
    from test_gpu_parity import gpu_run
    prep, meta = synth(tmp_path, "c2", 1.0)
    p = jb.PrepDir(prep)
    cols = p.decode(-1, 16)
    assert len(cols["pos"]) == meta["n_records"]
    genomes = [p.genome(t) for t in range(len(p.names))]
    rows, st, timing = gpu_run(cols, p.lengths, genomes, n_batches=1)
    rows2, st2, _ = gpu_run(cols, p.lengths, genomes, n_batches=5, pinned=True)
    assert rows.tobytes() == rows2.tobytes(), "result depends on batching / is not deterministic"
    ops = cols["cigar"] & 0xF
    n_pairs = int((ops == 3).sum())
    assert int(rows["nb_raw_aln"].astype(np.int64).sum()) == n_pairs == meta["n_pairs"]
    assert int(st["spliced"].sum()) == meta["n_spliced"]
    assert int(st["spliced"].sum() + st["unspliced"].sum()) == meta["n_records"]
    assert int(st["sumq"].sum()) == int(cols["l_qseq"].astype(np.int64).sum())
    assert np.all(np.diff(rows["tid"]) >= 0)
    order = np.lexsort((rows["end"], rows["start"], rows["tid"]))
    assert np.array_equal(order, np.arange(len(rows))), "rows not sorted by (tid,start,end)"
    assert len(np.unique(np.stack([rows["tid"], rows["start"], rows["end"]], 1), axis=0)) == len(rows)
    # per-junction identities
    tot = rows["nb_r1_pos"].astype(np.int64) + rows["nb_r1_neg"] + rows["nb_r2_pos"] + rows["nb_r2_neg"]
    assert np.array_equal(tot, rows["nb_raw_aln"].astype(np.int64))
    assert np.all(rows["nb_dist_aln"] <= rows["nb_raw_aln"]) and np.all(rows["nb_dist_aln"] >= 1)
    assert np.all(rows["nb_rel_aln"] == rows["nb_um_aln"])                       # orientation UNKNOWN (Q11)
    assert np.all(rows["jad"][:, 0] <= rows["nb_raw_aln"]) and np.all(np.diff(rows["jad"].astype(np.int64), axis=1) <= 0)
    assert np.all(rows["left"] < rows["start"]) and np.all(rows["right"] > rows["end"])
    assert np.all((rows["entropy"] >= 0) & (rows["entropy"] <= np.log2(np.maximum(rows["nb_raw_aln"], 1)) + 1e-9))
    w = rows["nb_raw_aln"].astype(np.float64)
    canon = float((w * (rows["canonical_ss"] == ord("C"))).sum() / w.sum())
    assert canon > 0.9, canon                                                      # planted GT..AG / CT..AC carry most reads
