"""GPU parity tests (run with -m gpu on a B200): the CUDA path, called through the C ABI, against the CPU oracle
on the same inputs and against the committed outputs of the reference binary."""
import filecmp
import os

import numpy as np
import pytest

import oracle_binding as ob
import synth
from compare import assert_exon_gff_equal, assert_rows_equal, assert_tab_equal
from conftest import FIXTURES, GOLDEN, ORIENTED, make_prep
from portcullis_b200 import _lib as L
from portcullis_b200 import junction_builder as jb

pytestmark = pytest.mark.gpu


def gpu_run(cols, lengths, genomes, orientation="UNKNOWN", n_batches=1, pinned=False, device=0, match_group=0, legacy_sort=0):
    g = jb.JuncGpu(device, orientation, match_group, legacy_sort)
    try:
        g.set_targets(lengths)
        for t, s in enumerate(genomes):
            if s is not None:
                g.set_genome(t, s)
        n = len(cols["pos"])
        g.shard_begin(n // 2, 0, 0)            # deliberately small hints: exercises arena growth
        edges = np.linspace(0, n, n_batches + 1).astype(int)
        for a, b in zip(edges[:-1], edges[1:]):
            sub = slice_cols(cols, a, b)
            (g.submit_pinned if pinned else g.submit)(sub)
        g.run()
        rows, st = g.fetch()
        timing = g.timing()
    finally:
        g.close()
    return rows, st, timing


def slice_cols(cols, a, b):
    out = {k: cols[k][a:b] for k in ("tid", "pos", "flag", "mapq", "xs", "l_qseq", "mtid", "mpos")}
    co, so = cols["cigar_off"], cols["seq_off"]
    out["cigar_off"] = (co[a:b + 1] - co[a]).astype(np.uint32)
    out["seq_off"] = (so[a:b + 1] - so[a]).astype(np.uint64)
    out["cigar"] = cols["cigar"][int(co[a]):int(co[b])]
    out["seq4"] = cols["seq4"][int(so[a]):int(so[b])]
    if cols.get("name_code") is not None:
        out["name_code"] = cols["name_code"][a:b]
    return out


def check_against_oracle(cols, lengths, genomes, orientation="UNKNOWN", **kw):
    exp_rows, exp_st = ob.run(cols, lengths, genomes, L.ORIENT[orientation])
    rows, st, timing = gpu_run(cols, lengths, genomes, orientation, **kw)
    assert_rows_equal(rows, exp_rows, "gpu vs oracle")
    for f in ("spliced", "unspliced", "sumq", "minq", "maxq"):
        assert np.array_equal(st[f], exp_st[f]), f
    # host finalize on top of GPU rows == oracle finalize on top of oracle rows
    total = int(st["spliced"].sum() + st["unspliced"].sum())
    if total:
        meanq = float(st["sumq"].sum()) / total
        assert_rows_equal(jb.finalize(rows.copy(), meanq), ob.finalize(exp_rows.copy(), meanq), "finalized", finalize_fields=True)
    assert timing[1] > 0, "no kernel launches recorded"
    return rows


@pytest.mark.parametrize("fixture", FIXTURES)
def test_cabi_matches_oracle_on_golden_fixtures(tmp_path, fixture):
    p = jb.PrepDir(make_prep(tmp_path, fixture))
    cols = p.decode(-1, 2)
    genomes = [p.genome(t) for t in range(len(p.names))]
    for orient in ["UNKNOWN"] + ([ORIENTED[fixture]] if fixture in ORIENTED else []):
        check_against_oracle(cols, p.lengths, genomes, orient)


@pytest.mark.parametrize("fixture", FIXTURES)
def test_junction_builder_reproduces_reference_files(tmp_path, fixture):
    """`JunctionBuilder(prep, out).process()` == files written by the unmodified reference `junc`."""
    prep = make_prep(tmp_path, fixture)
    for orient in [None] + ([ORIENTED[fixture]] if fixture in ORIENTED else []):
        out = str(tmp_path / ("out_%s" % orient) / "p")
        b = jb.JunctionBuilder(prep, out)
        b.setThreads(2)
        b.setOutputExonGFF(True)
        b.setOutputIntronGFF(True)
        if orient:
            b.setOrientation(orient)
        rep = b.process()
        assert rep["n_kernel_launches"] > 0
        tag = "ref" if orient is None else "ref_" + orient
        ref = os.path.join(GOLDEN, fixture, tag + ".junctions.")
        assert_tab_equal(out + ".junctions.tab", ref + "tab")
        assert filecmp.cmp(out + ".junctions.bed", ref + "bed", shallow=False)
        assert filecmp.cmp(out + ".junctions.intron.gff3", ref + "intron.gff3", shallow=False)
        assert_exon_gff_equal(out + ".junctions.exon.gff3", ref + "exon.gff3")


@pytest.mark.parametrize("seed,kw,orient,batches,pinned", [
    (101, dict(), "UNKNOWN", 1, False),
    (102, dict(n_targets=3, target_len=15000), "FR", 4, True),
    (103, dict(long_reads=True, read_len=(500, 3000), genes_per_target=6, target_len=60000, paired=False), "UNKNOWN", 3, False),
    (104, dict(n_targets=4, indel_rate=0.5, clip_rate=0.5, retain_rate=0.4, sub_rate=0.04), "RF", 7, True),
    (105, dict(hot=20000, n_targets=1, genes_per_target=4), "FF", 2, False),
    (106, dict(n_targets=6, target_len=8000, genes_per_target=3, reads_per_gene=(1, 4)), "UNKNOWN", 1, True),
])
def test_random_datasets_match_oracle(seed, kw, orient, batches, pinned):
    ds = synth.make_dataset(seed, **kw)
    cols = synth.to_columns(ds)
    check_against_oracle(cols, ds["lengths"], ds["genomes"], orient, n_batches=batches, pinned=pinned)


def test_edge_cases():
    ds = synth.make_dataset(7, n_targets=2, target_len=6000, genes_per_target=2, reads_per_gene=(2, 5))
    cols = synth.to_columns(ds)
    # empty shard
    empty = slice_cols(cols, 0, 0)
    rows, st, _ = gpu_run(empty, ds["lengths"], ds["genomes"])
    assert len(rows) == 0 and int(st["spliced"].sum()) == 0 and list(st["minq"]) == [2**31 - 1] * 2
    # unspliced records only
    keep = [i for i, r in enumerate(ds["records"]) if "N" not in r["cigar"]]
    un = synth.from_records([ds["records"][i] for i in keep])
    rows, st, _ = gpu_run(un, ds["lengths"], ds["genomes"])
    assert len(rows) == 0 and int(st["unspliced"].sum()) == len(keep)
    # a single spliced read
    first = next(i for i, r in enumerate(ds["records"]) if "N" in r["cigar"])
    one = synth.from_records([ds["records"][first]])
    check_against_oracle(one, ds["lengths"], ds["genomes"])
    # SEQ '*' (l_qseq == 0) on a spliced read: junction.cc:168-185
    r = dict(ds["records"][first]); r["seq"] = None; r["l_qseq"] = 0
    check_against_oracle(synth.from_records([r, ds["records"][first]] if r["pos"] <= ds["records"][first]["pos"] else [ds["records"][first], r]),
                         ds["lengths"], ds["genomes"])


def test_rejected_inputs_fail_loudly():
    ds = synth.make_dataset(8, n_targets=1, target_len=6000, genes_per_target=2, reads_per_gene=(2, 5))
    recs = [r for r in ds["records"] if "N" in r["cigar"]][:3]
    cols = synth.from_records(recs)
    # spliced read whose SEQ bytes are missing although l_qseq says otherwise
    bad = dict(cols); bad["seq_off"] = np.zeros_like(cols["seq_off"]); bad["seq4"] = np.zeros(0, np.uint8)
    with pytest.raises(L.PjError) as ei:
        gpu_run(bad, ds["lengths"], ds["genomes"])
    assert ei.value.code == L.PJ_EDATA
    # target without genome sequence but with junctions: the reference throws in processJunctionWindow
    with pytest.raises(L.PjError) as ei:
        gpu_run(cols, ds["lengths"], [None])
    assert ei.value.code == L.PJ_EDATA
    # the oracle rejects the same input
    with pytest.raises(ob.OracleError):
        ob.run(cols, ds["lengths"], [b""])
    # prefix columns that are not prefix offsets (a host bug): rejected before any kernel walks them, not a device fault
    def swap_middle(a):
        b = a.copy(); b[1], b[2] = a[2] + 7, a[1]; return b
    for col, edit in (("cigar_off", lambda a: a[::-1].copy()), ("cigar_off", swap_middle), ("seq_off", lambda a: np.concatenate([a[:1], a[1:][::-1]])), ("seq_off", swap_middle)):
        bad = dict(cols); bad[col] = edit(cols[col]).astype(cols[col].dtype)
        if bad[col][-1] < bad[col][0]:
            bad[col][0], bad[col][-1] = cols[col][0], cols[col][-1]        # keep the ends plausible: the host checks those itself
        assert not np.all(np.diff(bad[col].astype(np.int64)) >= 0)
        with pytest.raises(L.PjError) as ei:
            gpu_run(bad, ds["lengths"], ds["genomes"])
        assert ei.value.code in (L.PJ_EINVAL, L.PJ_EDATA), col


@pytest.mark.parametrize("group", [1, 2, 4, 8, 16, 32])
def test_match_kernel_group_widths(group):
    """Every lanes-per-pair instantiation of k_match gives the same answer (short, indel-rich and long reads)."""
    for seed, kw in ((201, dict(indel_rate=0.5, clip_rate=0.5, retain_rate=0.4, sub_rate=0.04)),
                     (202, dict(long_reads=True, read_len=(300, 2500), genes_per_target=5, target_len=50000, paired=False, sub_rate=0.03))):
        ds = synth.make_dataset(seed, **kw)
        check_against_oracle(synth.to_columns(ds), ds["lengths"], ds["genomes"], match_group=group)


def test_both_sort_implementations_agree():
    """The one-sweep sort (default) and the multi-kernel histogram/scan/scatter sort give identical rows."""
    ds = synth.make_dataset(301, n_targets=3, target_len=30000, genes_per_target=10, hot=30000)
    cols = synth.to_columns(ds)
    a, _, _ = gpu_run(cols, ds["lengths"], ds["genomes"])
    b, _, _ = gpu_run(cols, ds["lengths"], ds["genomes"], legacy_sort=1)
    assert a.tobytes() == b.tobytes()
    check_against_oracle(cols, ds["lengths"], ds["genomes"], legacy_sort=1)


@pytest.mark.skipif(not os.path.exists(ob.REF_BIN), reason="oracle/_ref not built")
@pytest.mark.parametrize("fixture,strandedness", [("kat", None), ("clipped3", "firststrand"), ("short_pe", "secondstrand"), ("long_se", None), ("indel_rich", "unstranded")])
def test_strand_analysis_report_matches_reference(tmp_path, fixture, strandedness):
    """A15: JunctionSystem::determineStrandedness(true) (junction_system.cc:455-560) is a stdout report — totals, the four
    correlation ratios (printed `-nan` when a class is empty), the two 'Determined ...' lines — plus a warning on stderr when
    --strandedness disagrees.  Our CLI prints the same block as the reference binary."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    prep = make_prep(tmp_path, fixture)
    extra = ["--strandedness", strandedness] if strandedness else []

    def block(cmd):
        p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
        assert p.returncode == 0, p.stderr[-400:]
        lines = p.stdout.split("\n")
        a = lines.index("Strand Analysis")
        b = max(i for i, l in enumerate(lines) if l.startswith("Determined RNAseq strandedness"))
        warn = [l for l in p.stderr.split("\n") if l.startswith("Warning!")]
        return lines[a:b + 1], warn

    ours = block([os.path.join(root, "portcullis_b200", "bin", "portcullis"), "junc", "-t", "2", "-o", str(tmp_path / "o" / "p")] + extra + [prep])
    ref = block([ob.REF_BIN, "junc", "-t", "1", "-o", str(tmp_path / "r" / "p")] + extra + [prep])
    assert ours[0] == ref[0]
    assert ours[1] == ref[1]
