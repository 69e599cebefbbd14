"""`bamfilt`: our `portcullis bamfilt` against the UNMODIFIED reference BamFilter (src/bam_filter.cc,
compiled into oracle/_ref/portcullis_ref; its `samtools index` call is answered by oracle/samtools_shim).  The survivor test
runs on the GPU (pj_jset_filter), so these are GPU tests; the CPU part checks argument handling."""
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

needs_ref = pytest.mark.skipif(not os.path.exists(ob.REF_BIN), reason="oracle/_ref not built (needs /root/reference)")


def ref_bamfilt(jfile, bam, out, mode="HARD", msrs=False):
    env = dict(os.environ, PATH=refrun.SAMTOOLS_SHIM + os.pathsep + os.environ.get("PATH", ""))
    cmd = [ob.REF_BIN, "bamfilt", "--clip_mode", mode, "-o", out] + (["-m"] if msrs else []) + [jfile, bam]
    p = subprocess.run(cmd, env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, check=True)
    line = [l for l in p.stdout.split("\n") if l.startswith("Filtered out")][0]
    nums = [int(x.strip(";()")) for x in line.replace(":", " ").split() if x.strip(";()").isdigit()]
    return dict(filtered=nums[0], n_in=nums[1], n_out=nums[2], n_modified=nums[3])


def subset_tab(src, dst, keep_every=3, seed=1):
    lines = open(src).read().split("\n")
    rows = [l for l in lines[1:] if l.strip()]
    rng = random.Random(seed)
    kept = [r for r in rows if rng.randrange(keep_every) == 0]
    with open(dst, "w") as f:
        f.write(lines[0] + "\n" + "\n".join(kept) + "\n\n")
    return len(kept)


def test_bamfilt_rejects_bad_input(tmp_path):
    with pytest.raises(L.PjError):
        jb.BamFilter("/nonexistent.tab", os.path.join(GOLDEN, "kat", "reads.bam"), str(tmp_path / "o.bam")).filter()
    bad = tmp_path / "bad.tab"; bad.write_text("index\trefid\n0\t0\t1\t2\n")
    with pytest.raises(L.PjError) as e:
        jb.BamFilter(str(bad), os.path.join(GOLDEN, "kat", "reads.bam"), str(tmp_path / "o.bam")).filter()
    assert "incorrect number of columns" in str(e.value)


@pytest.mark.gpu
@needs_ref
@pytest.mark.parametrize("fixture,mode,msrs", [("extra_mm", "HARD", True), ("short_pe", "SOFT", False), ("long_se", "COMPLETE", True), ("indel_rich", "HARD", False)])
def test_bamfilt_matches_reference(tmp_path, fixture, mode, msrs):
    bam = os.path.join(GOLDEN, fixture, "reads.bam")
    jfile = str(tmp_path / "sub.tab")
    n_kept = subset_tab(os.path.join(GOLDEN, fixture, "ref.junctions.tab"), jfile)
    want = ref_bamfilt(jfile, bam, str(tmp_path / "ref.bam"), mode, msrs)
    f = jb.BamFilter(jfile, bam, str(tmp_path / "ours.bam"))
    f.setClipMode(mode); f.setSaveMSRs(msrs); f.setThreads(3)
    rep = f.filter()
    assert rep["n_junctions"] == n_kept
    assert (rep["n_in"], rep["n_out"], rep["n_modified"]) == (want["n_in"], want["n_out"], want["n_modified"])
    assert 0 < rep["n_out"] < rep["n_in"]
    assert open(str(tmp_path / "ours.bam"), "rb").read() == open(str(tmp_path / "ref.bam"), "rb").read()
    for ext in ((".mod.bam", ".unmod.bam") if msrs else ()):
        assert open(str(tmp_path / "ours.bam") + ext, "rb").read() == open(str(tmp_path / "ref.bam") + ext, "rb").read(), ext
    if not msrs:
        assert not os.path.exists(str(tmp_path / "ours.bam.mod.bam"))
    rng = random.Random(2)
    regions = ["%d:%d-%d" % (t, b, b + rng.choice([1, 60, 3000])) for t, b in [(rng.randrange(2), rng.randrange(9000)) for _ in range(60)]]
    q = [subprocess.run([ob.BAMTOOL, "query", str(tmp_path / "ours.bam"), str(tmp_path / (d + ".bam.bai"))] + regions,
                        stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, check=True).stdout for d in ("ours", "ref")]
    assert q[0] == q[1]


@pytest.mark.gpu
def test_junction_set_filter_semantics():
    """The walk of containsJunctionInSystem (bam_filter.cc:75-99), including its quirk: an N op does not advance lEnd, so the
    later junctions of a multiply spliced read are looked up at shifted coordinates."""
    from portcullis_b200.columnar import from_records
    recs = [dict(tid=0, pos=100, cigar="50M", seq="A" * 50),                      # unspliced: kept
            dict(tid=0, pos=100, cigar="20M100N30M", seq="A" * 50),                # intron [120, 219] in the set: kept
            dict(tid=0, pos=100, cigar="20M101N30M", seq="A" * 50),                # [120, 220] not in the set: dropped
            dict(tid=1, pos=100, cigar="20M100N30M", seq="A" * 50),                # same coordinates, other target: dropped
            dict(tid=0, pos=100, cigar="10M50N10M100N30M", seq="A" * 50),          # true second intron [170, 269]; the walk looks up [120, 219]: kept
            dict(tid=0, pos=100, cigar="5S20M5I100N2D30M", seq="A" * 60),          # S / I do not move lEnd: [120, 219]: kept
            dict(tid=0, pos=100, cigar="10M2D8M100N30M", seq="A" * 48)]            # D moves it: [120, 219]: kept
    cols = from_records([dict(r, flag=0, mapq=60, xs=0, mtid=-1, mpos=-1) for r in recs])
    s = jb.JunctionSet([0, 0, 2], [120, 5000, 120], [219, 6000, 219])
    keep, nn = s.filter(cols)
    s.close()
    assert list(keep) == [1, 1, 0, 0, 1, 1, 1] and list(nn) == [0, 1, 1, 1, 2, 1, 1]
    empty = jb.JunctionSet([], [], [])
    keep, _ = empty.filter(cols)
    empty.close()
    assert list(keep) == [1, 0, 0, 0, 0, 0, 0]
