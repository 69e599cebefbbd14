#!/usr/bin/env python3
"""Regenerates the committed golden fixtures by running the UNMODIFIED reference binary
(oracle/_ref/portcullis_ref, built from /root/reference by `make -C oracle ref`) in this container.

    python tests/golden/make_golden.py

Each fixture directory holds the inputs (genome.fa[.fai], reads.bam[.bai]) and the reference's outputs
(ref.junctions.tab/.bed/.exon.gff3/.intron.gff3, plus ref_FR.* for --orientation FR where listed, and
ref_extra.junctions.tab from `junc --extra`, and ref_separate.md5: file / inflated-stream md5 of the spliced, unspliced
and unmapped BAMs that run wrote).
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import make_kat  # noqa: E402
import oracle_binding as ob  # noqa: E402
import refrun  # noqa: E402
import synth  # noqa: E402

FIXTURES = synth.GOLDEN_SPECS


def finish(work, out, orientations):
    os.makedirs(out, exist_ok=True)
    for f in ("genome.fa", "genome.fa.fai", "reads.bam", "reads.bam.bai"):
        shutil.copy(os.path.join(work, f), os.path.join(out, f))
    prep = os.path.join(work, "prep")
    for o in orientations:
        tag = "ref" if o is None else "ref_" + o
        refrun.run_reference(prep, os.path.join(work, tag), orientation=o)
        for e in ("tab", "bed", "exon.gff3", "intron.gff3"):
            shutil.copy(os.path.join(work, "%s.junctions.%s" % (tag, e)), os.path.join(out, "%s.junctions.%s" % (tag, e)))
    # the hidden --extra metrics (mm_score, coverage, up_aln, down_aln): only the tab differs
    refrun.run_reference(prep, os.path.join(work, "ref_extra"), extra=True, exon_gff=False, intron_gff=False)
    shutil.copy(os.path.join(work, "ref_extra.junctions.tab"), os.path.join(out, "ref_extra.junctions.tab"))
    # --extra implies --separate: keep the md5 of the three BAM files (whole file, and the inflated record stream)
    import gzip
    import hashlib
    with open(os.path.join(out, "ref_separate.md5"), "w") as f:
        for kind in ("spliced", "unspliced", "unmapped"):
            raw = open(os.path.join(work, "ref_extra.%s.bam" % kind), "rb").read()
            f.write("%s\t%s\t%s\n" % (kind, hashlib.md5(raw).hexdigest(), hashlib.md5(gzip.decompress(raw)).hexdigest()))


def main():
    tmp = "/tmp/pj_golden"
    shutil.rmtree(tmp, ignore_errors=True)
    # 13-read crafted known-answer test
    w = os.path.join(tmp, "kat")
    os.makedirs(w)
    fa, sam = make_kat.build()
    open(os.path.join(w, "genome.fa"), "w").write(fa)
    open(os.path.join(w, "reads.sam"), "w").write(sam)
    refrun.write_fai(os.path.join(w, "genome.fa"))
    subprocess.check_call([ob.BAMTOOL, "sam2bam", os.path.join(w, "reads.sam"), os.path.join(w, "reads.bam")])
    refrun.link_prep_dir(w, os.path.join(w, "genome.fa"), os.path.join(w, "reads.bam"))
    finish(w, os.path.join(HERE, "kat"), [None, "FR"])
    for name, (seed, kw, orients) in FIXTURES.items():
        w = os.path.join(tmp, name)
        ds = synth.make_dataset(seed, **kw)
        refrun.make_prep_dir(ds, w)
        finish(w, os.path.join(HERE, name), orients)
        print(name, len(ds["records"]), "records")


if __name__ == "__main__":
    main()
