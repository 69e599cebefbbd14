import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Build the product library and the oracle once per session if they are missing."""
    lib = os.path.join(ROOT, "portcullis_b200", "libportcullis_junc.so")
    if not os.path.exists(lib):
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "portcullis_b200", "csrc")])
    if not os.path.exists(os.path.join(ROOT, "oracle", "liboracle_junc.so")):
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "oracle"])


def chr4_genome():
    """The synthetic Chr4 of config 1a, generated once per machine (18.5 Mb: not committed)."""
    import synth
    d = os.path.join("/tmp", "pj_chr4_%d" % os.getuid())
    fa = os.path.join(d, "genome.fa")
    if not (os.path.exists(fa) and os.path.exists(fa + ".fai") and os.path.getsize(fa) == 6 + 18585056 + (18585056 + 59) // 60):
        os.makedirs(d, exist_ok=True)
        tmp = fa + ".%d.tmp" % os.getpid()
        synth.chr4_fasta(tmp)
        os.replace(tmp + ".fai", fa + ".fai")
        os.replace(tmp, fa)
    return fa


def make_prep(tmpdir, fixture):
    """Lay out a prep directory (src/prepare.hpp:114-140) over a committed fixture."""
    src = os.path.join(GOLDEN, fixture)
    if fixture == "clipped3":          # the reference's bundled BAM; its genome is the survey's synthetic Chr4
        fa = chr4_genome()
        prep = os.path.join(str(tmpdir), "prep_" + fixture)
        os.makedirs(prep, exist_ok=True)
        for a, b in ((fa, "portcullis.genome.fa"), (fa + ".fai", "portcullis.genome.fa.fai"),
                     (os.path.join(src, "reads.bam"), "portcullis.sorted.alignments.bam"),
                     (os.path.join(src, "reads.bam.bai"), "portcullis.sorted.alignments.bam.bai")):
            d = os.path.join(prep, b)
            if not os.path.lexists(d):
                os.symlink(a, d)
        return prep
    prep = os.path.join(str(tmpdir), "prep_" + fixture)
    os.makedirs(prep, exist_ok=True)
    for a, b in (("genome.fa", "portcullis.genome.fa"), ("genome.fa.fai", "portcullis.genome.fa.fai"),
                 ("reads.bam", "portcullis.sorted.alignments.bam"), ("reads.bam.bai", "portcullis.sorted.alignments.bam.bai")):
        d = os.path.join(prep, b)
        if not os.path.lexists(d):
            os.symlink(os.path.join(src, a), d)
    return prep


FIXTURES = ["kat", "clipped3", "short_pe", "long_se", "indel_rich"]
EXTRA_FIXTURES = FIXTURES + ["extra_mm"]      # every fixture also holds ref_extra.junctions.tab (`junc --extra`)
ORIENTED = {"kat": "FR", "clipped3": "FR", "short_pe": "FR", "indel_rich": "RF"}
