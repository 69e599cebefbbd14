"""`filt` feature extraction against the UNMODIFIED reference: oracle/_ref/feature_ref drives
portcullis::ml::ModelFeatures (lib/src/model_features.cc) + the Markov models + Junction::calc* on a junctions.tab and dumps
the feature matrix; pj_features_* must reproduce it — integer / copied columns exactly, log-probability columns within 1e-6.

Negative-strand windows over LOWER-CASE bases are undefined behaviour in the reference (REVCOMP_LOOKUP is indexed past its end,
seq_utils.hpp:115; the bytes read differ between translation units of one build), so those cases run on an upper-cased copy of
the genome; soft-masking itself is covered on the junctions that are not on the negative strand."""
import os
import subprocess

import numpy as np
import pytest

import refrun
from conftest import GOLDEN, make_prep
from portcullis_b200 import junction_builder as jb
from portcullis_b200 import model_features as mf

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FEATURE_REF = os.path.join(ROOT, "oracle", "_ref", "feature_ref")
PJSYNTH = os.path.join(ROOT, "portcullis_b200", "bin", "pjsynth")
TOL = 1e-6        # relative, for the log-probability sums (north star / VERDICT r1)


def reference_features(tab, fasta, coding, passm, failm, workdir):
    sub = os.path.join(workdir, "subsets.txt")
    with open(sub, "w") as f:
        for c, p, q in zip(coding, passm, failm):
            f.write("%d %d %d\n" % (c, p, q))
    out = os.path.join(workdir, "ref_features.tsv")
    subprocess.check_call([FEATURE_REF, tab, fasta, sub, out], stderr=subprocess.DEVNULL)
    rows, l95 = [], None
    with open(out) as f:
        for line in f:
            if line.startswith("#L95"):
                l95 = int(line.split("\t")[1])
            else:
                rows.append([float(x) for x in line.rstrip("\n").split("\t")])
    return np.array(rows, dtype=np.float64).reshape(len(rows), mf.NB_FEATURES), l95


def our_features(tab, sequences, lengths, coding, passm, failm):
    rows = mf.load_junction_tab(tab)
    g = jb.JuncGpu(0, "UNKNOWN")
    try:
        g.set_targets(lengths)
        for t, s in enumerate(sequences):
            g.set_genome(t, s)
        m = mf.ModelFeatures(g)
        if np.any(coding):
            m.calcIntronThreshold(rows, coding)
            m.trainCodingPotentialModel(rows, coding)
        if np.any(passm) or np.any(failm):
            m.trainSplicingModels(rows, passm, failm)
        x = m.juncs2FeatureVectors(rows)
        l95 = m.L95
        m.close()
    finally:
        g.close()
    return rows, x, l95


def assert_features_equal(got, exp, sel=None):
    if sel is not None:
        got, exp = got[sel], exp[sel]
    assert got.shape == exp.shape
    for c in range(1, 9):                          # counts and the doubles copied from the table
        assert np.array_equal(got[:, c], exp[:, c]), mf.VAR_NAMES[c]
    assert np.array_equal(got[:, 10], exp[:, 10]), "dna_minhamm"
    for c in [9, 11, 12, 13] + list(range(14, 34)):
        ok = np.isclose(got[:, c], exp[:, c], rtol=TOL, atol=1e-9)
        assert ok.all(), "%s: row %d got %r expected %r (%d of %d differ)" % (mf.VAR_NAMES[c], int(np.flatnonzero(~ok)[0]), got[~ok][0, c], exp[~ok][0, c], int((~ok).sum()), len(ok))


def subsets(n, seed):
    rng = np.random.default_rng(seed)
    coding = rng.random(n) < 0.6
    passm = rng.random(n) < 0.5
    failm = (~passm) & (rng.random(n) < 0.6)
    return coding, passm, failm


def upper_copy(fasta, workdir):
    out = os.path.join(workdir, "genome_upper.fa")
    with open(fasta) as f, open(out, "w") as o:
        for line in f:
            o.write(line if line.startswith(">") else line.upper())
    refrun.write_fai(out)
    return out


def fasta_sequences(path):
    seqs, cur, seen = [], [], False
    with open(path) as f:
        for line in f:
            if line.startswith(">"):
                if seen:
                    seqs.append("".join(cur)); cur = []
                seen = True
            else:
                cur.append(line.strip())
    seqs.append("".join(cur))
    return [s.encode() for s in seqs]


@pytest.mark.parametrize("fixture", ["short_pe", "indel_rich", "long_se", "extra_mm"])
def test_features_match_reference_on_fixtures(tmp_path, fixture):
    assert os.path.exists(FEATURE_REF), "oracle/_ref/feature_ref not built (run __graft_entry__.build() where /root/reference exists)"
    src = os.path.join(GOLDEN, fixture)
    tab = os.path.join(src, "ref.junctions.tab")
    n = sum(1 for _ in open(tab)) - 2
    coding, passm, failm = subsets(n, 5)
    p = jb.PrepDir(make_prep(tmp_path, fixture))
    # (1) upper-cased genome: every junction, both strands
    fa_up = upper_copy(os.path.join(src, "genome.fa"), str(tmp_path))
    exp, l95 = reference_features(tab, fa_up, coding, passm, failm, str(tmp_path))
    rows, got, our_l95 = our_features(tab, fasta_sequences(fa_up), p.lengths, coding, passm, failm)
    assert our_l95 == l95
    assert_features_equal(got, exp)
    # (2) the soft-masked genome as it is: train and compare on the junctions that are not on the negative strand
    pos = rows["consensus_strand"] != 1
    if pos.sum() >= 3:
        exp, l95 = reference_features(tab, os.path.join(src, "genome.fa"), coding & pos, passm & pos, failm & pos, str(tmp_path))
        rows, got, our_l95 = our_features(tab, [p.genome(t) for t in range(len(p.names))], p.lengths, coding & pos, passm & pos, failm & pos)
        assert our_l95 == l95
        assert_features_equal(got, exp, sel=pos)


def test_untrained_models_give_the_reference_constants(tmp_path):
    """No training at all: dna_coding = 0 (isCodingPotentialModelEmpty), every k-mer lookup misses, the positional product is 0."""
    src = os.path.join(GOLDEN, "short_pe")
    tab = os.path.join(src, "ref.junctions.tab")
    n = sum(1 for _ in open(tab)) - 2
    z = np.zeros(n, bool)
    p = jb.PrepDir(make_prep(tmp_path, "short_pe"))
    fa_up = upper_copy(os.path.join(src, "genome.fa"), str(tmp_path))
    exp, l95 = reference_features(tab, fa_up, z, z, z, str(tmp_path))
    rows, got, our_l95 = our_features(tab, fasta_sequences(fa_up), p.lengths, z, z, z)
    assert our_l95 == 0 and l95 == 0
    assert_features_equal(got, exp)
    assert np.all(got[:, 11] == 0.0) and np.all(got[:, 12] == -600.0)


@pytest.mark.parametrize("preset,scale", [("c2", 0.03), ("c5", 0.03)])
def test_features_match_reference_on_synthetic_presets(tmp_path, preset, scale):
    d = str(tmp_path / "prep")
    subprocess.check_call([PJSYNTH, "--preset", preset, "--scale", str(scale), "--out", d], stderr=subprocess.DEVNULL)
    ref_prefix = os.path.join(str(tmp_path), "ref", "r")
    refrun.run_reference(d, ref_prefix, threads=8, exon_gff=False, intron_gff=False)
    tab = ref_prefix + ".junctions.tab"
    n = sum(1 for _ in open(tab)) - 2
    coding, passm, failm = subsets(n, 9)
    fasta = os.path.join(d, "portcullis.genome.fa")                    # pjsynth c2 / c5 genomes are upper case
    exp, l95 = reference_features(tab, fasta, coding, passm, failm, str(tmp_path))
    p = jb.PrepDir(d)
    rows, got, our_l95 = our_features(tab, [p.genome(t) for t in range(len(p.names))], p.lengths, coding, passm, failm)
    assert our_l95 == l95 and n > 1000
    assert_features_equal(got, exp)
