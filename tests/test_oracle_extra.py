"""The `--extra` metrics: the oracle restatement (oj_extra) against the UNMODIFIED reference.

Golden vectors: tests/golden/<fixture>/ref_extra.junctions.tab, written by tests/golden/make_golden.py from
`portcullis_ref junc --extra` (the reference needs `samtools index`; oracle/samtools_shim answers it with htslib-1.3's own
indexer).  Where oracle/_ref is present the pileup cap of htslib (8000 reads, sam.c:1622/1906) is replayed live as well.
"""
import os
import tempfile

import numpy as np
import pytest

import oracle_binding as ob
import synth
from compare import extra_tab_columns
from conftest import EXTRA_FIXTURES, GOLDEN, make_prep
from portcullis_b200 import junction_builder as jb


def oracle_all(prep_dir):
    p = jb.PrepDir(prep_dir)
    cols = p.decode(-1, 2, names=True)
    genomes = [p.genome(t) for t in range(len(p.names))]
    rows, st = ob.run(cols, p.lengths, genomes)
    tot = int(st["spliced"].sum() + st["unspliced"].sum())
    rows = ob.finalize(rows, st["sumq"].sum() / max(tot, 1))
    x, capped = ob.extra(cols, p.lengths, rows, int(st["maxq"].max()))
    return cols, p, rows, st, x, capped


def tab_extra_columns(path):
    with open(path) as f:
        lines = [l.split("\t") for l in f.read().split("\n")[1:] if l]
    return [l[50:54] for l in lines]


@pytest.mark.parametrize("fixture", EXTRA_FIXTURES)
def test_oracle_extra_matches_reference_golden(fixture, tmp_path):
    _, _, rows, _, x, capped = oracle_all(make_prep(tmp_path, fixture))
    ref = tab_extra_columns(os.path.join(GOLDEN, fixture, "ref_extra.junctions.tab"))
    assert len(ref) == len(rows)
    assert extra_tab_columns(x) == ref
    assert capped == 0


def test_extra_fixture_exercises_every_metric():
    """extra_mm must hold multi-mapping names, flanking reads, coverage on some targets and none on others (Q14)."""
    with tempfile.TemporaryDirectory() as d:
        _, p, rows, _, x, _ = oracle_all(make_prep(d, "extra_mm"))
    assert (x["mm_m"] > x["mm_n"]).sum() > 10 and (x["up_aln"] > 0).any() and (x["down_aln"] > 0).any()
    by_tid = {t: x["coverage"][rows["tid"] == t] for t in range(len(p.names))}
    assert not by_tid[0].any() and not by_tid[1].any()          # first covered target, and the target without unspliced reads
    assert by_tid[2].any() and by_tid[3].any()                  # scored against target 0's depth / its own (last batch)


def test_coverage_source_rule():
    src = jb.coverage_source
    assert list(src([1, 1, 1])) == [-1, 0, 2]
    assert list(src([1, 0, 1, 1])) == [-1, -1, 0, 3]
    assert list(src([0, 1, 0])) == [-1, 1, -1]
    assert list(src([0, 0])) == [-1, -1]
    assert list(src([1, 1, 0])) == [-1, 1, -1]


def test_extra_finalize_arithmetic():
    from portcullis_b200 import _lib as L
    x = np.zeros(3, dtype=L.EXTRA_DTYPE)
    x["mm_n"] = [5, 3, 0]; x["mm_m"] = [8, 3, 0]
    x["cov_sum"] = [[9, 10, 20, 18], [0, 0, 0, 0], [1, 2, 3, 4]]
    y = jb.extra_finalize(x)
    assert y["mm_score"][0] == 5 / 8 and y["mm_score"][1] == 1.0 and np.isnan(y["mm_score"][2])
    m10, m11 = 1.0 / 9, 1.0 / 10
    assert y["coverage"][0] == (m10 * 9 - m11 * 10) + (m11 * 20 - m10 * 18)
    assert y["coverage"][1] == 0.0


@pytest.mark.skipif(not os.path.exists(ob.REF_BIN), reason="oracle/_ref not built (needs /root/reference)")
def test_oracle_pileup_cap_matches_reference_live(tmp_path):
    """Piles of > 8000 identical unspliced reads next to crafted junctions: htslib drops reads at its cap, the
    restatement drops the same ones (coverage is sensitive to a single read here)."""
    import refrun
    ds = synth.deep_dataset()
    prep = refrun.make_prep_dir(ds, str(tmp_path / "deep"))
    refrun.run_reference(prep, str(tmp_path / "ref"), extra=True, exon_gff=False, intron_gff=False)
    _, _, rows, _, x, capped = oracle_all(prep)
    assert capped > 5000
    assert extra_tab_columns(x) == tab_extra_columns(str(tmp_path / "ref.junctions.tab"))
