"""Small deterministic synthetic data sets for parity tests (test infrastructure).

Produces, from a seed: target sequences (with N runs, soft-masked stretches and a few IUPAC codes), gene
models with planted splice motifs, and coordinate-sorted spliced / unspliced alignments with substitutions,
indels near splice sites, soft clips, multi-intron reads, secondary copies, mixed MAPQ and XS tags.  The same
records can be rendered as SAM text (for the reference binary) and as columnar arrays (for the oracle and the
CUDA library).  Constraints of SURVEY §8(d) "value distributions" are respected: no H ops, no leading/trailing N,
every anchor >= 1 bp, SEQ present, junctions >= 10 bp from target ends, XS only of type A.
"""
import numpy as np

from portcullis_b200.columnar import from_records

COMP = {"A": "T", "C": "G", "G": "C", "T": "A", "N": "N"}


def make_genome(rng, n_targets, length, n_frac=0.002, lower_frac=0.2, iupac=0):
    targets = []
    for t in range(n_targets):
        g = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, length)].copy()
        targets.append(g)
    return targets


def plant_genes(rng, genome, n_genes, max_exons=8, exon_len=(20, 200), intron_len=(30, 600), margin=50):
    """Return list of gene models: (strand, [(exon_start, exon_end_exclusive), ...]); plants motifs in `genome`."""
    L = len(genome)
    genes = []
    for _ in range(n_genes):
        ne = int(rng.integers(2, max_exons + 1))
        el = rng.integers(exon_len[0], exon_len[1], ne)
        il = rng.integers(intron_len[0], intron_len[1], ne - 1)
        span = int(el.sum() + il.sum())
        if span + 2 * margin >= L:
            continue
        s = int(rng.integers(margin, L - span - margin))
        strand = "+" if rng.random() < 0.5 else "-"
        exons = []
        p = s
        for k in range(ne):
            exons.append((p, p + int(el[k])))
            p += int(el[k])
            if k < ne - 1:
                i0, i1 = p, p + int(il[k])          # intron [i0, i1)
                r = rng.random()
                if r < 0.90:
                    d, a = ("GT", "AG")
                elif r < 0.95:
                    d, a = ("GC", "AG")
                elif r < 0.98:
                    d, a = ("AT", "AC")
                else:
                    d, a = (None, None)
                if d:
                    if strand == "-":
                        # reverse complement of donor..acceptor as seen on the forward strand
                        d, a = ("".join(COMP[c] for c in reversed(a)), "".join(COMP[c] for c in reversed(d)))
                    genome[i0:i0 + 2] = np.frombuffer(d.encode(), dtype=np.uint8)
                    genome[i1 - 2:i1] = np.frombuffer(a.encode(), dtype=np.uint8)
                p = i1
        genes.append((strand, exons))
    return genes


def decorate_genome(rng, genome, n_runs=3, lower_runs=5, iupac=4):
    L = len(genome)
    for _ in range(n_runs):
        s = int(rng.integers(0, L - 40)); n = int(rng.integers(1, 30))
        genome[s:s + n] = ord("N")
    for _ in range(iupac):
        genome[int(rng.integers(0, L))] = ord("RYKMSWX"[int(rng.integers(0, 7))])
    for _ in range(lower_runs):
        s = int(rng.integers(0, L - 200)); n = int(rng.integers(10, 200))
        seg = genome[s:s + n]
        genome[s:s + n] = np.where((seg >= 65) & (seg <= 90), seg + 32, seg)


def _read_from_gene(rng, gseq_upper, gene, read_len, sub_rate, indel_rate, clip_rate, retain_rate):
    """Sample one alignment from a transcript. Returns (pos, cigar_str, seq) or None."""
    strand, exons = gene
    # optionally retain an intron (merge two exons) so that junction-wide anchors span neighbouring introns (Q5)
    ex = list(exons)
    if len(ex) > 2 and rng.random() < retain_rate:
        k = int(rng.integers(0, len(ex) - 1))
        ex[k:k + 2] = [(ex[k][0], ex[k + 1][1])]
    tlen = sum(e - s for s, e in ex)
    if tlen < 30:
        return None
    rl = min(read_len, tlen)
    t0 = int(rng.integers(0, tlen - rl + 1))
    # walk the transcript interval [t0, t0+rl) over the exons
    blocks = []
    acc = 0
    for s, e in ex:
        n = e - s
        lo, hi = max(t0, acc), min(t0 + rl, acc + n)
        if lo < hi:
            blocks.append((s + lo - acc, s + hi - acc))
        acc += n
    pos = blocks[0][0]
    cigar = []
    seq = []
    for bi, (s, e) in enumerate(blocks):
        if bi > 0:
            cigar.append((blocks[bi][0] - blocks[bi - 1][1], "N"))
        n = e - s
        bases = list(gseq_upper[s:e])
        # indel near a splice site (within 5 bp), keeping >= 1 matched base on both sides of the event
        if n >= 12 and rng.random() < indel_rate:
            off = int(rng.integers(1, 5))
            at = off if rng.random() < 0.5 else n - off - 3
            ln = int(rng.integers(1, 4))
            if rng.random() < 0.5:        # insertion of ln bases before `at`
                ins = ["ACGT"[int(x)] for x in rng.integers(0, 4, ln)]
                cigar += [(at, "M"), (ln, "I"), (n - at, "M")]
                bases = bases[:at] + ins + bases[at:]
            else:                         # deletion of ln reference bases at `at`
                cigar += [(at, "M"), (ln, "D"), (n - at - ln, "M")]
                bases = bases[:at] + bases[at + ln:]
        else:
            cigar.append((n, "M"))
        seq += bases
    # substitutions
    seq = [("ACGT"[(("ACGT".find(c) if c in "ACGT" else 0) + 1 + int(rng.integers(0, 3))) % 4] if rng.random() < sub_rate else c) for c in seq]
    # read bases opposite genome N / IUPAC stay as in the genome (N==N matches); sometimes make the read say N
    if rng.random() < 0.02:
        seq[int(rng.integers(0, len(seq)))] = "N"
    # soft clips
    if rng.random() < clip_rate:
        n5 = int(rng.integers(1, 8)); seq = ["ACGT"[int(x)] for x in rng.integers(0, 4, n5)] + seq; cigar.insert(0, (n5, "S"))
    if rng.random() < clip_rate:
        n3 = int(rng.integers(1, 8)); seq = seq + ["ACGT"[int(x)] for x in rng.integers(0, 4, n3)]; cigar.append((n3, "S"))
    # merge adjacent M ops created by the indel logic with zero length pieces
    out = []
    for ln, op in cigar:
        if ln == 0:
            continue
        if out and out[-1][1] == op and op == "M":
            out[-1] = (out[-1][0] + ln, op)
        else:
            out.append((ln, op))
    return pos, "".join("%d%s" % (ln, op) for ln, op in out), "".join(seq)


def make_dataset(seed, n_targets=2, target_len=20000, genes_per_target=12, reads_per_gene=(5, 120), read_len=(60, 150),
                 long_reads=False, sub_rate=0.01, indel_rate=0.15, clip_rate=0.1, retain_rate=0.15, unspliced_frac=0.2,
                 paired=True, secondary_frac=0.05, lowq_frac=0.15, xs_frac=0.8, hot=None,
                 multimap_frac=0.0, unspliced_indel=0.0, deep=(), no_background=(), base_genomes=None):
    """Returns dict(names, lengths, genomes (list[bytes], original case), records (sorted list of dicts)).

    Options for the `--extra` metrics (all off by default, so the seeds of older data sets keep their records):
    multimap_frac: fraction of alignments that get a second alignment with the SAME read name (a multi-mapping read);
    unspliced_indel: fraction of background reads with a deletion / insertion / soft clips; deep: (tid, pos, count, len)
    piles of identical unspliced reads (htslib's 8000-read pileup cap); no_background: targets without any unspliced
    read."""
    rng = np.random.default_rng(seed)
    genomes = make_genome(rng, n_targets, target_len)
    if base_genomes is not None:          # real sequences (e.g. soft-masked, with n runs): genes are planted into a copy
        genomes = [np.frombuffer(bytes(b), dtype=np.uint8).copy() for b in base_genomes]
        n_targets = len(genomes)
    names = ["chr%s" % (t + 1) for t in range(n_targets)]
    records = []
    rid = 0
    for t in range(n_targets):
        g = genomes[t]
        genes = plant_genes(rng, g, genes_per_target, max_exons=(20 if long_reads else 8))
        if base_genomes is None:
            decorate_genome(rng, g)
        target_len = len(g)
        gu = bytes(np.where((g >= 97) & (g <= 122), g - 32, g).astype(np.uint8)).decode()
        for gi, gene in enumerate(genes):
            nr = int(rng.integers(reads_per_gene[0], reads_per_gene[1]))
            if hot and gi == 0 and t == 0:
                nr = hot
            for _ in range(nr):
                rl = int(rng.integers(read_len[0], read_len[1]))
                r = _read_from_gene(rng, gu, gene, rl, sub_rate, indel_rate, clip_rate, retain_rate)
                if r is None:
                    continue
                pos, cigar, seq = r
                spliced = "N" in cigar
                if not spliced and t in no_background:
                    continue
                rev = rng.random() < 0.5
                flag = 0
                mtid, mpos = -1, -1
                if paired:
                    first = rng.random() < 0.5
                    flag |= 0x1 | (0x40 if first else 0x80) | (0x10 if rev else 0x20)
                    if rng.random() < 0.9:
                        flag |= 0x2
                    mtid = t
                    mpos = max(0, pos + (int(rng.integers(50, 400)) if not rev else -int(rng.integers(50, 400))))
                    if rng.random() < 0.03:
                        flag |= 0x8
                else:
                    flag |= 0x10 if rev else 0
                mapq = 60 if rng.random() > lowq_frac else int(rng.choice([0, 1, 3, 29]))
                if rng.random() < secondary_frac:
                    flag |= 0x100
                xs = 0
                if spliced and rng.random() < xs_frac:
                    xs = gene[0] if rng.random() < 0.93 else ("+" if gene[0] == "-" else "-")
                name = "r%06d" % rid
                if multimap_frac and records and rng.random() < multimap_frac:
                    other = records[int(rng.integers(0, len(records)))]
                    if (other["flag"] & 0xC1) == (flag & 0xC1):      # same deriveName() suffix
                        name = other["name"]; flag |= 0x100
                records.append(dict(name=name, tid=t, pos=pos, flag=flag, mapq=mapq, cigar=cigar, seq=seq, xs=xs,
                                    mtid=mtid, mpos=mpos))
                rid += 1
        # unspliced background + a placed unmapped read
        nb = 0 if t in no_background else int(unspliced_frac * sum(1 for r in records if r["tid"] == t))
        for _ in range(nb):
            rl = int(rng.integers(30, 120)); pos = int(rng.integers(0, target_len - rl))
            cigar, seq, fl = "%dM" % rl, gu[pos:pos + rl].replace("X", "N"), 0
            if unspliced_indel and rl >= 40 and rng.random() < unspliced_indel:
                a = int(rng.integers(5, rl - 20)); d = int(rng.integers(1, 12)); kind = int(rng.integers(0, 3))
                if kind == 0 and pos + rl + d < target_len:      # deletion: the read spans rl + d reference bases
                    cigar = "%dM%dD%dM" % (a, d, rl - a); seq = (gu[pos:pos + a] + gu[pos + a + d:pos + rl + d]).replace("X", "N")
                elif kind == 1:                                   # insertion
                    cigar = "%dM%dI%dM" % (a, d, rl - a); seq = (gu[pos:pos + a] + "A" * d + gu[pos + a:pos + rl]).replace("X", "N")
                else:                                             # soft clips on both sides
                    cigar = "%dS%dM%dS" % (d, rl, d); seq = "C" * d + seq + "G" * d
                fl = int(rng.choice([0, 0x10, 0x100, 0x400, 0x200]))
            records.append(dict(name="u%06d" % rid, tid=t, pos=pos, flag=fl, mapq=60, cigar=cigar, seq=seq, xs=0, mtid=-1, mpos=-1))
            rid += 1
        for (dt, dpos, dcount, dlen) in deep:
            if dt != t:
                continue
            for _ in range(dcount):
                records.append(dict(name="d%06d" % rid, tid=t, pos=dpos, flag=0, mapq=60, cigar="%dM" % dlen,
                                    seq=gu[dpos:dpos + dlen].replace("X", "N"), xs=0, mtid=-1, mpos=-1))
                rid += 1
        records.append(dict(name="x%06d" % rid, tid=t, pos=int(rng.integers(0, target_len - 50)), flag=4, mapq=0, cigar="", seq="ACGTACGTAC", xs=0, mtid=-1, mpos=-1))
        rid += 1
    records.sort(key=lambda r: (r["tid"], r["pos"]))
    return dict(names=names, lengths=np.array([len(g) for g in genomes], dtype=np.int32),
                genomes=[bytes(g) for g in genomes], records=records)


def to_columns(ds):
    return from_records(ds["records"])


def to_sam(ds):
    out = ["@HD\tVN:1.0\tSO:coordinate"]
    for n, l in zip(ds["names"], ds["lengths"]):
        out.append("@SQ\tSN:%s\tLN:%d" % (n, l))
    for r in ds["records"]:
        mt = "*" if r["mtid"] < 0 else ("=" if r["mtid"] == r["tid"] else ds["names"][r["mtid"]])
        xs = ("\tXS:A:%s" % r["xs"]) if r["xs"] else ""
        if r.get("tags_before"):          # other aux fields in front of XS (aligner output order): the decoder must step over every type
            xs = "\t" + r["tags_before"] + xs
        if r.get("tags_after"):
            xs = xs + "\t" + r["tags_after"]
        out.append("%s\t%d\t%s\t%d\t%d\t%s\t%s\t%d\t0\t%s\t*%s" % (
            r["name"], r["flag"], ds["names"][r["tid"]], r["pos"] + 1, r["mapq"], r["cigar"] or "*", mt, r["mpos"] + 1 if r["mpos"] >= 0 else 0,
            r["seq"], xs))
    return "\n".join(out) + "\n"


def to_fasta(ds, width=60):
    out = []
    for n, g in zip(ds["names"], ds["genomes"]):
        out.append(">%s\n" % n)
        s = g.decode()
        out.extend(s[i:i + width] + "\n" for i in range(0, len(s), width))
    return "".join(out)


# data sets behind the committed golden fixtures (tests/golden/<name>/, written by tests/golden/make_golden.py)
GOLDEN_SPECS = {
    # name: (seed, kwargs, orientations the reference was run with)
    "short_pe": (11, dict(n_targets=2, target_len=12000, genes_per_target=8, reads_per_gene=(5, 60)), [None, "FR"]),
    "long_se": (12, dict(n_targets=2, target_len=40000, genes_per_target=5, reads_per_gene=(5, 40), long_reads=True,
                         read_len=(400, 2500), paired=False), [None]),
    # `--extra` metrics: multi-mapping read names, background reads with indels / clips / secondary / duplicate flags,
    # and a target (the second of four) without unspliced reads so that DepthParser's batch pairing (Q14) skips it
    "extra_mm": (14, dict(n_targets=4, target_len=9000, genes_per_target=6, reads_per_gene=(5, 60), multimap_frac=0.2,
                          unspliced_indel=0.5, unspliced_frac=1.5, no_background=(1,)), [None]),
    "indel_rich": (13, dict(n_targets=3, target_len=10000, genes_per_target=6, reads_per_gene=(5, 50), indel_rate=0.5,
                            clip_rate=0.5, retain_rate=0.4, sub_rate=0.03), [None, "RF"]),
}


def deep_dataset(seed=21):
    """Piles of > 8000 identical unspliced reads (htslib's pileup read cap, sam.c:1622/1906) next to crafted junctions whose
    coverage windows touch them, plus reads without reference span ("20S") inside a capped column."""
    import re
    ds = make_dataset(seed, n_targets=2, target_len=6000, genes_per_target=4, reads_per_gene=(5, 40), multimap_frac=0.1,
                      unspliced_indel=0.3, deep=((0, 1500, 4000, 80), (0, 1520, 5000, 90), (0, 1520, 300, 30), (0, 1530, 4000, 60),
                                                 (0, 1531, 20, 10), (0, 4000, 9000, 40), (1, 2500, 8100, 50), (1, 2500, 10, 70)))

    def add(t, pos, cigar):
        g = ds["genomes"][t].decode().upper()
        seq, x = "", pos
        for n, op in re.findall(r"(\d+)([MNS])", cigar):
            if op in "MS":
                seq += g[x:x + int(n)]
            if op != "S":
                x += int(n)
        ds["records"].append(dict(name="m%d_%d_%s_%d" % (t, pos, cigar, len(ds["records"])), tid=t, pos=pos, flag=0, mapq=60, cigar=cigar,
                                  seq=seq.replace("X", "N"), xs=0, mtid=-1, mpos=-1))
    add(0, 1300, "50M145N60M"); add(1, 2545, "15M100N50M"); add(1, 2300, "50M149N40M"); add(0, 3900, "60M95N30M")
    for _ in range(3):
        add(0, 1520, "20S"); add(1, 2500, "12S")
    ds["records"].sort(key=lambda r: (r["tid"], r["pos"]))
    return ds


def chr4_fasta(path, length=18585056, seed=12345):
    """Config 1a of SURVEY.md §8(d) / Appendix B: the genome of the reference's bundled clipped3.bam (artha_chr4.fa) is not shipped,
    so the survey pinned a synthetic one — numpy.random.default_rng(12345).integers(0, 4, L) -> ACGT, header >Chr4, 60 columns.
    Writes the FASTA and its .fai (one sequence, so the index is a single line)."""
    rng = np.random.default_rng(seed)
    g = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, length)]
    full = (length // 60) * 60
    body = np.empty((full // 60, 61), dtype=np.uint8)
    body[:, :60] = g[:full].reshape(-1, 60)
    body[:, 60] = 10
    with open(path, "wb") as f:
        f.write(b">Chr4\n")
        f.write(body.tobytes())
        if length > full:
            f.write(g[full:].tobytes() + b"\n")
    with open(path + ".fai", "w") as f:
        f.write("Chr4\t%d\t6\t60\t61\n" % length)
    return path
