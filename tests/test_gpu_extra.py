"""GPU parity of the `--extra` metrics (run with -m gpu on a B200): pj_extra_* through the C ABI
against the oracle restatement (oj_extra) on the same columns, and the CLI-level driver against the committed outputs of
`portcullis_ref junc --extra`."""
import os

import numpy as np
import pytest

import oracle_binding as ob
import synth
from compare import assert_extra_equal, assert_rows_equal, assert_tab_equal, extra_tab_columns
from conftest import EXTRA_FIXTURES, GOLDEN, make_prep
from portcullis_b200 import _lib as L
from portcullis_b200 import junction_builder as jb
from test_gpu_fuzz import make_case
from test_gpu_parity import slice_cols
from test_oracle_extra import tab_extra_columns

pytestmark = pytest.mark.gpu


def oracle_extra(cols, lengths, genomes, orientation="UNKNOWN"):
    rows, st = ob.run(cols, lengths, genomes, L.ORIENT[orientation])
    tot = int(st["spliced"].sum() + st["unspliced"].sum())
    rows = ob.finalize(rows, st["sumq"].sum() / max(tot, 1))
    maxq = int(st["maxq"].max()) if len(st) else 0
    x, capped = ob.extra(cols, lengths, rows, maxq)
    return rows, st, x, capped, maxq


def gpu_extra(cols, lengths, genomes, maxq, n_batches=1, pinned=False, orientation="UNKNOWN"):
    g = jb.JuncGpu(0, orientation, extra=True)
    try:
        g.set_targets(lengths)
        for t, s in enumerate(genomes):
            g.set_genome(t, s)
        n = len(cols["pos"])
        g.shard_begin(n // 2, 0, 0)
        edges = np.linspace(0, n, n_batches + 1).astype(int)
        for a, b in zip(edges[:-1], edges[1:]):
            (g.submit_pinned if pinned else g.submit)(slice_cols(cols, a, b))
        g.run()
        rows, st = g.fetch()
        x, over = g.extra(rows, maxq)
    finally:
        g.close()
    return rows, st, x, over


@pytest.mark.parametrize("fixture", EXTRA_FIXTURES)
def test_extra_matches_oracle_and_reference_on_golden_fixtures(tmp_path, fixture):
    p = jb.PrepDir(make_prep(tmp_path, fixture))
    cols = p.decode(-1, 2, names=True)
    genomes = [p.genome(t) for t in range(len(p.names))]
    erows, _, ex, capped, maxq = oracle_extra(cols, p.lengths, genomes)
    rows, _, x, over = gpu_extra(cols, p.lengths, genomes, maxq, n_batches=3, pinned=True)
    assert_rows_equal(rows, erows, fixture)
    assert_extra_equal(x, ex, erows, fixture)
    assert capped == 0 and not over
    assert extra_tab_columns(x) == tab_extra_columns(os.path.join(GOLDEN, fixture, "ref_extra.junctions.tab"))


@pytest.mark.parametrize("fixture", EXTRA_FIXTURES)
@pytest.mark.parametrize("gpus", [1, 2])
def test_junction_builder_extra_reproduces_reference_tab(tmp_path, fixture, gpus):
    """`JunctionBuilder.setExtra(True)` (the CLI's --extra) == junctions.tab of the unmodified reference run with --extra.
    gpus=2 shards the targets over two contexts (both on device 0 when the box has one GPU): names and depth vectors cross
    contexts exactly as they would cross GPUs."""
    import torch
    out = str(tmp_path / "o" / "p")
    b = jb.JunctionBuilder(make_prep(tmp_path, fixture), out)
    b.setThreads(2)
    b.setExtra(True)
    if gpus == 2:
        b.setGpus(2)
        if torch.cuda.device_count() < 2:
            b.gpu_ids = [0, 0]
    rep = b.process()
    assert rep["n_kernel_launches"] > 0
    assert_tab_equal(out + ".junctions.tab", os.path.join(GOLDEN, fixture, "ref_extra.junctions.tab"))


@pytest.mark.parametrize("seed,kw,batches", [
    (201, dict(multimap_frac=0.3, unspliced_indel=0.5, unspliced_frac=2.0), 1),
    (202, dict(n_targets=5, target_len=8000, multimap_frac=0.1, unspliced_indel=0.3, unspliced_frac=1.0, no_background=(0, 3)), 4),
    (203, dict(long_reads=True, read_len=(500, 3000), genes_per_target=6, target_len=60000, paired=False, multimap_frac=0.2,
               unspliced_indel=0.5, unspliced_frac=3.0), 3),
    (204, dict(n_targets=3, unspliced_frac=0.0, no_background=(0, 1, 2), multimap_frac=0.5), 2),
])
def test_extra_on_synthetic_sets(seed, kw, batches):
    ds = synth.make_dataset(seed, **kw)
    cols = synth.to_columns(ds)
    erows, _, ex, capped, maxq = oracle_extra(cols, ds["lengths"], ds["genomes"])
    rows, _, x, over = gpu_extra(cols, ds["lengths"], ds["genomes"], maxq, n_batches=batches, pinned=bool(seed & 1))
    assert_rows_equal(rows, erows, "synthetic %d" % seed)
    assert_extra_equal(x, ex, erows, "synthetic %d" % seed)
    assert capped == 0 and not over


@pytest.mark.parametrize("seed", list(range(8)))
def test_extra_fuzz(seed):
    """Random CIGARs (clips, indels, =/X, P, multi-N): unspliced spans, pileup columns and region tests on odd alignments."""
    cols, lengths, genomes = make_case(3000 + seed)
    try:
        erows, _, ex, capped, maxq = oracle_extra(cols, lengths, genomes)
    except ob.OracleError as e:
        assert e.code == L.PJ_EDATA
        return
    rows, _, x, over = gpu_extra(cols, lengths, genomes, maxq, n_batches=1 + seed % 3)
    assert_rows_equal(rows, erows, "fuzz %d" % seed)
    assert_extra_equal(x, ex, erows, "fuzz %d" % seed)


def test_two_contexts_exchange_names_and_depth():
    """The multi-GPU recipe by hand: targets split over two contexts; spliced names are exported / imported, each depth
    vector is queried on the context that owns it."""
    ds = synth.make_dataset(205, n_targets=4, target_len=9000, genes_per_target=6, multimap_frac=0.4, unspliced_indel=0.3, unspliced_frac=1.5)
    cols = synth.to_columns(ds)
    erows, est, ex, _, maxq = oracle_extra(cols, ds["lengths"], ds["genomes"])
    owner = [0, 1, 1, 0]
    ctxs = [jb.JuncGpu(0, extra=True) for _ in range(2)]
    try:
        rows, xs = [], []
        for k, g in enumerate(ctxs):
            g.set_targets(ds["lengths"])
            sel = np.nonzero(np.isin(cols["tid"], [t for t in range(4) if owner[t] == k]))[0]
            g.shard_begin(len(sel), 0, 0)
            for t in range(4):
                if owner[t] == k:
                    g.set_genome(t, ds["genomes"][t])
                    idx = np.nonzero(cols["tid"] == t)[0]
                    g.submit(slice_cols(cols, int(idx[0]), int(idx[-1]) + 1))
            g.run()
            rows.append(g.fetch()[0])
        names = [g.export_names() for g in ctxs]
        ctxs[0].import_names(names[1]); ctxs[1].import_names(names[0])
        xs = [g.extra_run(maxq) for g in ctxs]
        covered = np.array([ctxs[owner[t]].target_pileup(t)[0] for t in range(4)], dtype=np.uint8)
        src = jb.coverage_source(covered)
        for k in range(2):
            for t in np.unique(rows[k]["tid"]):
                if src[t] >= 0:
                    sel = np.nonzero(rows[k]["tid"] == t)[0]
                    xs[k]["cov_sum"][sel] = ctxs[owner[src[t]]].coverage(src[t], rows[k]["start"][sel], rows[k]["end"][sel])
    finally:
        for g in ctxs:
            g.close()
    allrows = np.concatenate(rows); allx = np.concatenate(xs)
    order = np.lexsort((allrows["end"], allrows["start"], allrows["tid"]))
    assert_rows_equal(allrows[order], erows, "two contexts")
    assert_extra_equal(jb.extra_finalize(allx[order]), ex, erows, "two contexts")
    assert (ex["mm_m"] > ex["mm_n"]).any()


@pytest.mark.parametrize("general", [False, True])
def test_pileup_cap_is_replayed(general, monkeypatch):
    """> 8000 reads on one column: htslib's pileup stops accepting reads there (sam.c:1906).  The oracle models it (and is
    pinned against the reference on the same data in test_oracle_extra); the library replays it on the GPU (k_x_cap), so
    even the coverage column stays bit-equal."""
    if general:
        monkeypatch.setenv("PJ_CAP_GENERAL", "1")      # the one-warp-per-target kernel with global counters (reads longer than the ring)
    ds = synth.deep_dataset()
    cols = synth.to_columns(ds)
    erows, _, ex, capped, maxq = oracle_extra(cols, ds["lengths"], ds["genomes"])
    rows, _, x, over = gpu_extra(cols, ds["lengths"], ds["genomes"], maxq, n_batches=2)
    assert capped > 5000 and sorted(over) == [0, 1] and over[1] >= 8110
    assert_rows_equal(rows, erows, "deep")
    assert_extra_equal(x, ex, erows, "deep")
    assert ex["cov_sum"].max() > 8000 * 5


def test_extra_call_sequence_and_rejections():
    ds = synth.make_dataset(206, n_targets=1, target_len=6000, genes_per_target=3)
    cols = synth.to_columns(ds)
    g = jb.JuncGpu(0, extra=True)
    try:
        g.set_targets(ds["lengths"]); g.set_genome(0, ds["genomes"][0])
        g.shard_begin(10, 0, 0)
        with pytest.raises(L.PjError) as e:
            g.extra_run(100)                                     # before pj_shard_run
        assert e.value.code == L.PJ_ESTATE
        nameless = {k: v for k, v in cols.items() if k != "name_code"}
        with pytest.raises(L.PjError) as e:
            g.submit(nameless)                                   # the context needs name_code
        assert e.value.code == L.PJ_EINVAL
        g.submit(cols); g.run(); rows, st = g.fetch()
        g.extra_run(int(st["maxq"].max()))
        with pytest.raises(L.PjError) as e:
            g.extra_run(100)                                     # once per shard
        assert e.value.code == L.PJ_ESTATE
        with pytest.raises(L.PjError) as e:
            g.coverage(5, [1], [2])
        assert e.value.code == L.PJ_EINVAL
    finally:
        g.close()
    plain = jb.JuncGpu(0)
    try:
        plain.set_targets(ds["lengths"]); plain.set_genome(0, ds["genomes"][0])
        plain.shard_begin(10, 0, 0); plain.submit(cols); plain.run()
        with pytest.raises(L.PjError) as e:
            plain.extra_run(100)                                 # context created without extra_metrics
        assert e.value.code == L.PJ_ESTATE
    finally:
        plain.close()
    # a mapped, unspliced record without CIGAR: htslib's pileup asserts on it (sam.c:1537) -> rejected
    recs = [dict(name="a", tid=0, pos=100, flag=0, mapq=60, cigar="20M100N20M", seq="A" * 40, xs=0, mtid=-1, mpos=-1),
            dict(name="b", tid=0, pos=150, flag=0, mapq=60, cigar="", seq="ACGT", xs=0, mtid=-1, mpos=-1)]
    from portcullis_b200.columnar import from_records
    bad = from_records(recs)
    with pytest.raises(ob.OracleError) as oe:
        oracle_extra(bad, ds["lengths"], ds["genomes"])
    assert oe.value.code == L.PJ_EDATA
    g = jb.JuncGpu(0, extra=True)
    try:
        g.set_targets(ds["lengths"]); g.set_genome(0, ds["genomes"][0])
        g.shard_begin(10, 0, 0); g.submit(bad); g.run()
        with pytest.raises(L.PjError) as e:
            g.extra_run(40)
        assert e.value.code == L.PJ_EDATA
    finally:
        g.close()


@pytest.mark.parametrize("fixture", ["extra_mm", "clipped3"])
def test_cli_with_every_junc_flag(tmp_path, fixture):
    """`portcullis junc --separate --extra --exon_gff --intron_gff -t 3 --gpus 1` as a process: tab == the reference's --extra tab,
    gff / bed == the plain ones, the three BAMs == the reference's (md5), exit code 0 and the reference's closing line."""
    import hashlib
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "portcullis_b200", "bin", "portcullis")
    out = str(tmp_path / "o" / "p")
    p = subprocess.run([exe, "junc", "--separate", "--extra", "--exon_gff", "--intron_gff", "-t", "3", "--gpus", "1", "-o", out, make_prep(tmp_path, fixture)],
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    assert p.returncode == 0, p.stderr[-500:]
    assert "Portcullis junc completed." in p.stdout and "Splitting BAM:" in p.stdout and "Calculating extra junction metrics:" in p.stdout
    g = os.path.join(GOLDEN, fixture)
    assert_tab_equal(out + ".junctions.tab", os.path.join(g, "ref_extra.junctions.tab"))
    for ext in ("bed", "intron.gff3"):
        assert open(out + ".junctions." + ext, "rb").read() == open(os.path.join(g, "ref.junctions." + ext), "rb").read(), ext
    want = {l.split("\t")[0]: l.split("\t")[1] for l in open(os.path.join(g, "ref_separate.md5"))}
    for kind in ("spliced", "unspliced", "unmapped"):
        assert hashlib.md5(open("%s.%s.bam" % (out, kind), "rb").read()).hexdigest() == want[kind], kind
    assert os.path.exists(out + ".spliced.bam.bai") and os.path.exists(out + ".unspliced.bam.bai")
