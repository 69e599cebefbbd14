"""`prep` without samtools: the prep directory our `portcullis prep` lays out against the one the
UNMODIFIED reference `prep` lays out (it shells out to samtools: oracle/samtools_shim answers `index` with htslib-1.3's own
indexer and `sort` / `merge` with a restatement of samtools 1.3's published ordering, see oracle/bamtool.c).

CPU tests: genome index, the already-sorted path, CLI errors.  GPU tests (the coordinate order is a device radix sort): an
unsorted input and a three-file merge, record for record."""
import gzip
import os
import random
import subprocess

import numpy as np
import pytest

import oracle_binding as ob
import refrun
import synth
from conftest import GOLDEN
from portcullis_b200 import _lib as L
from portcullis_b200 import junction_builder as jb

HAVE_REF = os.path.exists(ob.REF_BIN)
needs_ref = pytest.mark.skipif(not HAVE_REF, reason="oracle/_ref not built (needs /root/reference)")


def ref_prep(genome, bams, out, extra=()):
    env = dict(os.environ, PATH=refrun.SAMTOOLS_SHIM + os.pathsep + os.environ.get("PATH", ""))
    subprocess.check_call([ob.REF_BIN, "prep", "-o", out, *extra, genome, *bams], env=env, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)


def records_of(bam):
    """(header text, list of raw record bytes) of a BAM file."""
    raw = gzip.decompress(open(bam, "rb").read())
    l_text = int.from_bytes(raw[4:8], "little")
    text = raw[8:8 + l_text].decode()
    p = 8 + l_text
    n_ref = int.from_bytes(raw[p:p + 4], "little"); p += 4
    for _ in range(n_ref):
        l_name = int.from_bytes(raw[p:p + 4], "little"); p += 4 + l_name + 4
    recs = []
    while p < len(raw):
        bs = int.from_bytes(raw[p:p + 4], "little")
        recs.append(raw[p:p + 4 + bs]); p += 4 + bs
    return text, recs


def write_inputs(ds, workdir, shuffle_seed=None, parts=1):
    """genome.fa (no .fai) and `parts` BAM files holding the data set's records; shuffled when a seed is given."""
    os.makedirs(workdir, exist_ok=True)
    fa = os.path.join(workdir, "genome.fa")
    with open(fa, "w") as f:
        f.write(synth.to_fasta(ds))
    recs = list(ds["records"])
    if shuffle_seed is not None:
        random.Random(shuffle_seed).shuffle(recs)
    bams = []
    for k in range(parts):
        part = dict(ds, records=recs[k::parts])
        sam = os.path.join(workdir, "in%d.sam" % k)
        text = synth.to_sam(part)
        if shuffle_seed is not None:
            text = text.replace("@HD\tVN:1.0\tSO:coordinate", "@HD\tVN:1.0\tSO:unsorted")
        with open(sam, "w") as f:
            f.write(text)
        bam = os.path.join(workdir, "in%d.bam" % k)
        subprocess.check_call([ob.BAMTOOL, "sam2bam", sam, bam, "noindex"], stderr=subprocess.DEVNULL)      # prep must cope without an index
        bams.append(bam)
    return fa, bams


@needs_ref
def test_prep_sorted_input_matches_reference(tmp_path):
    """Sorted single BAM: both sides symlink it; the genome index is built (byte-identical to htslib's fai_build) and the BAM
    index answers region queries like `samtools index`'s."""
    ds = synth.make_dataset(51, n_targets=3, target_len=30000, genes_per_target=10, reads_per_gene=(20, 100))
    fa, bams = write_inputs(ds, str(tmp_path / "in"))
    ref_prep(fa, bams, str(tmp_path / "ref"))
    rep = jb.Prepare(str(tmp_path / "ours")).prepare(bams, fa)
    assert rep["sorted_in_process"] == 0
    for name in ("portcullis.genome.fa", "portcullis.sorted.alignments.bam"):
        a, b = str(tmp_path / "ours" / name), str(tmp_path / "ref" / name)
        assert os.path.islink(a) and os.path.islink(b) and os.path.realpath(a) == os.path.realpath(b)
    assert open(str(tmp_path / "ours" / "portcullis.genome.fa.fai"), "rb").read() == open(str(tmp_path / "ref" / "portcullis.genome.fa.fai"), "rb").read()
    rng = random.Random(3)
    regions = ["%d:%d-%d" % (t, b, b + rng.choice([1, 50, 2000])) for t, b in [(rng.randrange(3), rng.randrange(30000)) for _ in range(100)]]
    q = [subprocess.run([ob.BAMTOOL, "query", bams[0], str(tmp_path / d / "portcullis.sorted.alignments.bam.bai")] + regions,
                        stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, check=True).stdout for d in ("ours", "ref")]
    assert q[0] == q[1]
    assert os.path.realpath(str(tmp_path / "ours" / "portcullis.unsorted.alignments.bam")) == os.path.realpath(str(tmp_path / "ref" / "portcullis.unsorted.alignments.bam"))
    # -c: CSI (link mode, like the reference)
    p = jb.Prepare(str(tmp_path / "ours_c")); p.setUseCsi(True)
    p.prepare(bams, fa)
    ref_prep(fa, bams, str(tmp_path / "ref_c"), ["-c"])
    assert sorted(os.listdir(str(tmp_path / "ours_c"))) == sorted(os.listdir(str(tmp_path / "ref_c")))
    q = [subprocess.run([ob.BAMTOOL, "query", bams[0], str(tmp_path / d / "portcullis.sorted.alignments.bam.csi")] + regions,
                        stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, check=True).stdout for d in ("ours_c", "ref_c")]
    assert q[0] == q[1]
    jb.PrepDir(str(tmp_path / "ours_c"), use_csi=True)       # a valid prep directory for junc
    # --copy: the reference leaves a dangling link for a sorted input (it deletes the copy its link points to, prepare.cc:321-324)
    # and fails; here the copy becomes the sorted file
    p = jb.Prepare(str(tmp_path / "ours_k")); p.setUseLinks(False)
    p.prepare(bams, fa)
    for name, src in (("portcullis.genome.fa", fa), ("portcullis.sorted.alignments.bam", bams[0])):
        a = str(tmp_path / "ours_k" / name)
        assert not os.path.islink(a) and open(a, "rb").read() == open(src, "rb").read(), name
    assert not os.path.exists(str(tmp_path / "ours_k" / "portcullis.unsorted.alignments.bam"))
    jb.PrepDir(str(tmp_path / "ours_k"))


def test_fai_builder_on_awkward_fasta(tmp_path):
    """fai_build_core restated (faidx.c:82-155): ragged last lines, blank lines between records, descriptions, CRLF-free."""
    fa = tmp_path / "g.fa"
    fa.write_text(">chrA some description\nACGTACGTAC\nACGTACGTAC\nACG\n\n>chrB\nAC\n>chrC\tx\nACGTAC\nACGTAC\n")
    bam_ds = synth.make_dataset(52, n_targets=1, target_len=5000, genes_per_target=2)
    fa2, bams = write_inputs(bam_ds, str(tmp_path / "in")) if os.path.exists(ob.BAMTOOL) else (None, None)
    if bams is None:
        pytest.skip("oracle/_ref not built")
    jb.Prepare(str(tmp_path / "o")).prepare(bams, str(fa))
    got = open(str(tmp_path / "o" / "portcullis.genome.fa.fai")).read()
    assert got == "chrA\t23\t23\t10\t11\nchrB\t2\t56\t2\t3\nchrC\t12\t67\t6\t7\n"      # what htslib-1.3's fai_build writes for this file
    if HAVE_REF:
        ref_prep(str(fa), bams, str(tmp_path / "r"))
        assert got == open(str(tmp_path / "r" / "portcullis.genome.fa.fai")).read()
    bad = tmp_path / "bad.fa"
    bad.write_text(">x\nACGT\nAC\nACGT\n")
    with pytest.raises(L.PjError) as e:
        jb.Prepare(str(tmp_path / "o2")).prepare(bams, str(bad))
    assert "different line length" in str(e.value)


def test_prep_rejects_bad_arguments(tmp_path):
    with pytest.raises(L.PjError):
        jb.Prepare(str(tmp_path / "o")).prepare(["/nonexistent.bam"], "/nonexistent.fa")
    fa = tmp_path / "g.fa"; fa.write_text(">a\nACGT\n")
    with pytest.raises(L.PjError):
        jb.Prepare(str(tmp_path / "o")).prepare([], str(fa))


@pytest.mark.gpu
def test_coordinate_order_equals_stable_argsort():
    rng = np.random.default_rng(5)
    for n in (0, 1, 1000, 300000):
        tid = rng.integers(-1, 40, n).astype(np.int32)
        pos = np.where(tid < 0, -1, rng.integers(0, 5000, n)).astype(np.int32)
        flag = rng.choice([0, 16, 99, 147, 4], n).astype(np.uint16)
        key = (tid.astype(np.int64).astype(np.uint64) << np.uint64(32)) | ((((pos.astype(np.int64) + 1) << 1) | ((flag >> 4) & 1)).astype(np.uint64) & np.uint64(0xffffffff))
        assert np.array_equal(jb.coordinate_order(tid, pos, flag), np.argsort(key, kind="stable").astype(np.uint32))


@pytest.mark.gpu
@needs_ref
@pytest.mark.parametrize("parts,csi", [(1, False), (3, False), (2, True)])
def test_prep_sort_and_merge_match_reference(tmp_path, parts, csi):
    """Shuffled records in 1-3 unsorted BAMs: our in-process sort / merge writes the same records in the same order as the
    reference's `samtools sort` + `merge` calls (answered by the shim), and `junc` on the two directories agrees."""
    ds = synth.make_dataset(53 + parts, n_targets=3, target_len=20000, genes_per_target=8, reads_per_gene=(20, 120), unspliced_frac=0.5)
    for k in range(4):
        ds["records"].append(dict(name="np%d" % k, tid=-1, pos=-1, flag=4, mapq=0, cigar="", seq="ACGTACGT", xs=0, mtid=-1, mpos=-1))
    fa, bams = write_inputs(ds, str(tmp_path / "in"), shuffle_seed=9, parts=parts)
    opts = ["-c"] if csi else []
    ref_prep(fa, bams, str(tmp_path / "ref"), opts)
    p = jb.Prepare(str(tmp_path / "ours")); p.setThreads(3); p.setUseCsi(csi)
    rep = p.prepare(bams, fa)
    assert rep["sorted_in_process"] == 1 and rep["n_records"] == len(ds["records"])
    t_o, r_o = records_of(str(tmp_path / "ours" / "portcullis.sorted.alignments.bam"))
    t_r, r_r = records_of(str(tmp_path / "ref" / "portcullis.sorted.alignments.bam"))
    assert r_o == r_r
    assert "SO:coordinate" in t_o.split("\n")[0] and [l for l in t_o.split("\n") if l.startswith("@SQ")] == [l for l in t_r.split("\n") if l.startswith("@SQ")]
    # both directories through junc
    outs = []
    for d in ("ours", "ref"):
        b = jb.JunctionBuilder(str(tmp_path / d), str(tmp_path / ("j_" + d) / "p"))
        b.setUseCsi(csi)
        b.process()
        outs.append(open(str(tmp_path / ("j_" + d) / "p.junctions.tab"), "rb").read())
    assert outs[0] == outs[1] and len(outs[0]) > 1000


@pytest.mark.gpu
def test_force_resort_does_not_touch_the_input_index(tmp_path):
    """--force re-sorts an already sorted BAM (prepare.cc:211).  The input's .bai describes the INPUT's block offsets: the
    prep directory must get its own index file, and nothing may be written through the symlink into the user's index."""
    if not os.path.exists(ob.BAMTOOL):
        pytest.skip("oracle/_ref not built")
    ds = synth.make_dataset(57, n_targets=2, target_len=20000, genes_per_target=8, reads_per_gene=(20, 100))
    fa, bams = write_inputs(ds, str(tmp_path / "in"))
    subprocess.check_call([ob.BAMTOOL, "index", bams[0]])
    before = open(bams[0] + ".bai", "rb").read()
    p = jb.Prepare(str(tmp_path / "o")); p.setForce(True); p.setThreads(2)
    rep = p.prepare(bams, fa)
    assert rep["sorted_in_process"] == 1
    assert open(bams[0] + ".bai", "rb").read() == before
    idx = str(tmp_path / "o" / "portcullis.sorted.alignments.bam.bai")
    assert os.path.exists(idx) and not os.path.islink(idx)
    a = jb.PrepDir(str(tmp_path / "o")).decode(-1, 2)
    b = synth.to_columns(ds)
    assert len(a["pos"]) == len(b["pos"]) and np.array_equal(np.sort(a["pos"]), np.sort(b["pos"]))


def test_genome_fetch_block_path_equals_faidx_semantics(tmp_path):
    """FastaFile::fetch_all copies whole blocks of regular lines (thousands of lines per block) and falls back to the line loop for
    the tail: 60 / 70 / 61-column records, a CRLF record, a record that ends on a partial line, one shorter than a block."""
    if not os.path.exists(ob.BAMTOOL):
        pytest.skip("oracle/_ref not built")
    ds = synth.make_dataset(53, n_targets=5, target_len=4000, genes_per_target=1)
    _, bams = write_inputs(ds, str(tmp_path / "in"))
    rng = random.Random(5)
    seqs, text = [], []
    for name, (width, n_bases, eol) in zip(ds["names"], ((60, 60 * 6000 + 17, "\n"), (70, 70 * 5000, "\r\n"), (61, 61 * 4100 + 60, "\n"), (60, 60 * 2048, "\n"), (80, 333, "\n"))):
        sq = "".join(rng.choice("ACGTNacgtn") for _ in range(n_bases))
        seqs.append(sq)
        text.append(">" + name + eol + eol.join(sq[k:k + width] for k in range(0, n_bases, width)) + eol)
    fa = tmp_path / "wide.fa"
    fa.write_bytes("".join(text).encode())
    jb.Prepare(str(tmp_path / "o")).prepare(bams, str(fa))
    p = jb.PrepDir(str(tmp_path / "o"))
    for t, sq in enumerate(seqs):
        assert p.genome(t) == sq.encode(), "target %d" % t
    p.close()
