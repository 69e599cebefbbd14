"""`junc --separate` (JunctionBuilder::separateBams): the three BAM files and their indices.
Host-only code, so these run without a GPU.  Golden vectors: tests/golden/<fixture>/ref_separate.md5 (made by the
unmodified reference); where oracle/_ref is present the files are also compared live and the indices are checked
through htslib-1.3's own region iterator (`bamtool query`)."""
import gzip
import hashlib
import os
import random
import subprocess

import numpy as np
import pytest

import oracle_binding as ob
import synth
from conftest import EXTRA_FIXTURES, GOLDEN, make_prep
from portcullis_b200 import junction_builder as jb

KINDS = ("spliced", "unspliced", "unmapped")


def md5s(path):
    raw = open(path, "rb").read()
    return hashlib.md5(raw).hexdigest(), hashlib.md5(gzip.decompress(raw)).hexdigest()


@pytest.mark.parametrize("fixture", EXTRA_FIXTURES)
def test_separate_matches_reference_golden(tmp_path, fixture):
    prep = make_prep(tmp_path, fixture)
    pre = str(tmp_path / "o" / "p")
    counts = jb.separate_bams(prep, pre, threads=3)
    want = {l.split("\t")[0]: l.rstrip("\n").split("\t")[1:] for l in open(os.path.join(GOLDEN, fixture, "ref_separate.md5"))}
    cols = jb.PrepDir(prep).decode(-1, 1)          # per-target fetch: every placed record of these fixtures
    n_spliced = sum(1 for i in range(len(cols["pos"])) if (cols["cigar"][cols["cigar_off"][i]:cols["cigar_off"][i + 1]] & 15 == 3).any())
    assert counts[0] == n_spliced and sum(counts) == len(cols["pos"])
    for kind in KINDS:
        whole, stream = md5s("%s.%s.bam" % (pre, kind))
        assert stream == want[kind][1], "%s: records differ from the reference's %s.bam" % (fixture, kind)
        # same zlib as the image that made the fixture -> even the compressed bytes agree (block layout follows htslib)
        assert whole == want[kind][0], "%s: %s.bam is not byte-identical to the reference's" % (fixture, kind)
    assert os.path.exists(pre + ".spliced.bam.bai") and os.path.exists(pre + ".unspliced.bam.bai") and not os.path.exists(pre + ".unmapped.bam.bai")


def test_separated_bams_are_usable_prep_inputs(tmp_path):
    """Our own reader over <prefix>.spliced.bam + our BAI / CSI sees exactly the spliced records of the original, per target."""
    prep = make_prep(tmp_path, "extra_mm")
    orig = jb.PrepDir(prep)
    for csi in (False, True):
        if csi:      # a CSI of the input is part of a `-c` prep directory
            if not os.path.exists(ob.BAMTOOL):
                pytest.skip("oracle/_ref not built")
            bam = os.path.join(prep, "portcullis.sorted.alignments.bam")
            subprocess.check_call([ob.BAMTOOL, "index_csi", os.path.realpath(bam)])
            os.symlink(os.path.realpath(bam) + ".csi", bam + ".csi")
        pre = str(tmp_path / ("s%d" % csi) / "p")
        jb.separate_bams(prep, pre, use_csi=csi, threads=2)
        d = str(tmp_path / ("prep_spliced%d" % csi))
        os.makedirs(d)
        ext = ".csi" if csi else ".bai"
        for src, dst in ((pre + ".spliced.bam", "portcullis.sorted.alignments.bam"), (pre + ".spliced.bam" + ext, "portcullis.sorted.alignments.bam" + ext),
                         (os.path.join(prep, "portcullis.genome.fa"), "portcullis.genome.fa"), (os.path.join(prep, "portcullis.genome.fa.fai"), "portcullis.genome.fa.fai")):
            os.symlink(os.path.realpath(src), os.path.join(d, dst))
        sp = jb.PrepDir(d, use_csi=csi)
        for t in range(len(orig.names)):
            a, b = orig.decode(t, 1), sp.decode(t, 2)
            keep = np.array([(a["cigar"][a["cigar_off"][i]:a["cigar_off"][i + 1]] & 15 == 3).any() for i in range(len(a["pos"]))], dtype=bool)
            assert np.array_equal(a["pos"][keep], b["pos"]) and np.array_equal(a["flag"][keep], b["flag"])
        if csi:
            os.remove(os.path.join(prep, "portcullis.sorted.alignments.bam.csi"))
            os.remove(os.path.realpath(os.path.join(prep, "portcullis.sorted.alignments.bam")) + ".csi")


@pytest.mark.skipif(not os.path.exists(ob.REF_BIN), reason="oracle/_ref not built (needs /root/reference)")
@pytest.mark.parametrize("csi", [False, True])
def test_separate_live_against_reference(tmp_path, csi):
    """A multi-block data set: files byte-identical to the reference's for any thread count; our BAI / CSI and the one
    `samtools index` (htslib-1.3) builds answer 200 random region queries identically through htslib's iterator."""
    import refrun
    ds = synth.make_dataset(31, n_targets=3, target_len=60000, genes_per_target=30, reads_per_gene=(100, 400), unspliced_frac=0.8,
                            unspliced_indel=0.3, multimap_frac=0.1)
    # unplaced reads sort last (tid -1): they belong in unmapped.bam and count as n_no_coor in the indices
    for k in range(5):
        ds["records"].append(dict(name="np%d" % k, tid=-1, pos=-1, flag=4, mapq=0, cigar="", seq="ACGTACGT", xs=0, mtid=-1, mpos=-1))
    prep = refrun.make_prep_dir(ds, str(tmp_path / "w"))
    if csi:
        bam = os.path.realpath(os.path.join(prep, "portcullis.sorted.alignments.bam"))
        subprocess.check_call([ob.BAMTOOL, "index_csi", bam])
        os.symlink(bam + ".csi", os.path.join(prep, "portcullis.sorted.alignments.bam.csi"))
    ref = str(tmp_path / "ref" / "p")
    cmd = [ob.REF_BIN, "junc", "--separate", "-o", ref] + (["-c"] if csi else []) + [prep]
    env = dict(os.environ, PATH=refrun.SAMTOOLS_SHIM + os.pathsep + os.environ.get("PATH", ""))
    subprocess.check_call(cmd, env=env, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    ext = ".csi" if csi else ".bai"
    rng = random.Random(7)
    regions = ["%d:%d-%d" % (t, b, b + rng.choice([1, 40, 700, 20000])) for t, b in
               [(rng.randrange(3), rng.randrange(60000)) for _ in range(200)]]
    for threads in (1, 4):
        pre = str(tmp_path / ("t%d" % threads) / "p")
        counts = jb.separate_bams(prep, pre, use_csi=csi, threads=threads)
        assert counts[2] >= 5
        for kind in KINDS:
            assert open("%s.%s.bam" % (pre, kind), "rb").read() == open("%s.%s.bam" % (ref, kind), "rb").read(), kind
        for kind in KINDS[:2]:
            bam = "%s.%s.bam" % (pre, kind)
            q = [subprocess.run([ob.BAMTOOL, "query", bam, ix] + regions, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, check=True).stdout
                 for ix in (bam + ext, "%s.%s.bam%s" % (ref, kind, ext))]
            assert q[0] == q[1] and sum(int(l.split(b"\t")[1]) for l in q[0].splitlines()) > 1000


@pytest.mark.skipif(not os.path.exists(ob.BAMTOOL), reason="oracle/_ref not built")
def test_csi_on_a_target_longer_than_512_mbp(tmp_path):
    """A 700 Mbp target needs CSI depth 6 (BAI cannot index it): our CSI against htslib-1.3's, through htslib's iterator,
    with records on both sides of the 2^29 boundary; the separated files decode through our own CSI reader."""
    L700 = 700_000_000
    sam = ["@HD\tVN:1.0\tSO:coordinate", "@SQ\tSN:big\tLN:%d" % L700, "@SQ\tSN:small\tLN:5000"]
    rng = random.Random(4)
    pos = sorted(rng.randrange(1, L700 - 400) for _ in range(3000)) + []
    pos += []
    for k, p in enumerate(sorted(pos + [2 ** 29 - 60, 2 ** 29 - 1, 2 ** 29, 2 ** 29 + 5, 536_900_000])):
        cigar = "50M" if k % 3 else "20M%dN30M" % (100 + k % 200)
        sam.append("r%d\t0\tbig\t%d\t60\t%s\t*\t0\t0\t%s\t*" % (k, p + 1, cigar, "A" * 50))
    for k in range(20):
        sam.append("s%d\t0\tsmall\t%d\t60\t50M\t*\t0\t0\t%s\t*" % (k, 100 + 10 * k, "C" * 50))
    d = tmp_path / "prep"; d.mkdir()
    (tmp_path / "in.sam").write_text("\n".join(sam) + "\n")
    bam = str(d / "portcullis.sorted.alignments.bam")
    subprocess.check_call([ob.BAMTOOL, "sam2bam", str(tmp_path / "in.sam"), bam, "noindex"], stderr=subprocess.DEVNULL)
    subprocess.check_call([ob.BAMTOOL, "index_csi", bam])
    (d / "portcullis.genome.fa").write_text(">big\nA\n>small\nA\n"); (d / "portcullis.genome.fa.fai").write_text("big\t1\t5\t1\t2\nsmall\t1\t14\t1\t2\n")
    pre = str(tmp_path / "o" / "p")
    counts = jb.separate_bams(str(d), pre, use_csi=True, threads=2)
    assert counts[0] > 900 and counts[1] > 1900
    regions = ["0:%d-%d" % (b, b + w) for b, w in [(2 ** 29 - 100, 200), (2 ** 29, 1), (0, L700), (536_899_990, 100), (123_456_789, 5_000_000)]] + ["1:0-5000", "1:150-160"]
    regions += ["0:%d-%d" % (b, b + rng.choice([1, 500, 100000, 50_000_000])) for b in (rng.randrange(L700) for _ in range(80))]
    for kind in ("spliced", "unspliced"):
        f = "%s.%s.bam" % (pre, kind)
        ref_copy = str(tmp_path / (kind + ".bam"))
        os.symlink(f, ref_copy)
        subprocess.check_call([ob.BAMTOOL, "index_csi", ref_copy])
        q = [subprocess.run([ob.BAMTOOL, "query", f, ix] + regions, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, check=True).stdout
             for ix in (f + ".csi", ref_copy + ".csi")]
        assert q[0] == q[1] and sum(int(l.split(b"\t")[1]) for l in q[0].splitlines()) > 100
    # our reader through our CSI: the unspliced file as a prep directory
    d2 = tmp_path / "prep2"; d2.mkdir()
    for src, dst in ((pre + ".unspliced.bam", "portcullis.sorted.alignments.bam"), (pre + ".unspliced.bam.csi", "portcullis.sorted.alignments.bam.csi"),
                     (str(d / "portcullis.genome.fa"), "portcullis.genome.fa"), (str(d / "portcullis.genome.fa.fai"), "portcullis.genome.fa.fai")):
        os.symlink(src, str(d2 / dst))
    cols = jb.PrepDir(str(d2), use_csi=True).decode(-1, 2)
    assert len(cols["pos"]) == counts[1] and int(cols["pos"].max()) > 2 ** 29
