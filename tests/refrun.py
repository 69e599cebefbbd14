"""Helpers to build a prep directory from a synthetic data set and to run the UNMODIFIED reference binary
(oracle/_ref/portcullis_ref) on it.  Test infrastructure; only usable where oracle/_ref has been built."""
import os
import subprocess

import oracle_binding as ob
import synth


def write_fai(fasta_path):
    """FASTA index (name, len, offset, line_blen, line_len), deps/htslib-1.3/faidx.c:40-44."""
    entries = []
    with open(fasta_path, "rb") as f:
        data = f.read()
    off = 0
    name = None
    for line in data.split(b"\n"):
        ln = len(line) + 1
        if line.startswith(b">"):
            if name is not None:
                entries.append((name, slen, soff, blen, blen + 1))
            name = line[1:].split()[0].decode()
            slen, soff, blen = 0, off + ln, 0
        elif name is not None and line:
            if blen == 0:
                blen = len(line)
            slen += len(line)
        off += ln
    if name is not None:
        entries.append((name, slen, soff, blen, blen + 1))
    with open(fasta_path + ".fai", "w") as f:
        for e in entries:
            f.write("%s\t%d\t%d\t%d\t%d\n" % e)


def make_prep_dir(ds, workdir):
    """Writes genome.fa/.fai, reads.sam -> reads.bam/.bai (reference htslib) and the prep directory layout of
    src/prepare.hpp:114-140.  Returns the prep dir path."""
    os.makedirs(workdir, exist_ok=True)
    fa = os.path.join(workdir, "genome.fa")
    with open(fa, "w") as f:
        f.write(synth.to_fasta(ds))
    write_fai(fa)
    sam = os.path.join(workdir, "reads.sam")
    with open(sam, "w") as f:
        f.write(synth.to_sam(ds))
    bam = os.path.join(workdir, "reads.bam")
    subprocess.check_call([ob.BAMTOOL, "sam2bam", sam, bam], stderr=subprocess.DEVNULL)
    return link_prep_dir(workdir, fa, bam)


def link_prep_dir(workdir, fa, bam):
    prep = os.path.join(workdir, "prep")
    os.makedirs(prep, exist_ok=True)
    for src, dst in ((fa, "portcullis.genome.fa"), (fa + ".fai", "portcullis.genome.fa.fai"),
                     (bam, "portcullis.sorted.alignments.bam"), (bam + ".bai", "portcullis.sorted.alignments.bam.bai")):
        d = os.path.join(prep, dst)
        if os.path.lexists(d):
            os.remove(d)
        os.symlink(os.path.abspath(src), d)
    return prep


SAMTOOLS_SHIM = os.path.join(ob.ORACLE_DIR, "samtools_shim")


def run_reference(prep, out_prefix, threads=1, orientation=None, exon_gff=True, intron_gff=True, extra=False, separate=False):
    cmd = [ob.REF_BIN, "junc", "-t", str(threads), "-o", out_prefix]
    if extra:
        cmd.append("--extra")
    if separate:
        cmd.append("--separate")
    if exon_gff:
        cmd.append("--exon_gff")
    if intron_gff:
        cmd.append("--intron_gff")
    if orientation:
        cmd += ["--orientation", orientation]
    cmd.append(prep)
    env = dict(os.environ, PATH=SAMTOOLS_SHIM + os.pathsep + os.environ.get("PATH", ""))   # `samtools index` for --separate/--extra
    p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, env=env)
    if p.returncode != 0:
        raise RuntimeError("reference junc failed (%d): %s\n%s" % (p.returncode, p.stderr[-2000:], p.stdout[-500:]))
    return p.stdout
