"""Randomised CIGAR fuzzing of the CUDA path against the oracle (columnar level, no BAM involved).

Reads are built op by op from the whole CIGAR alphabet (M I D N S H P = X), including the shapes the reference treats
specially: insertions / deletions glued to an N, leading H+S (clipping quirk Q3), '=' / 'X' ops, P ops, several N per
read, reads that start inside another read's intron (junction-wide windows walking through introns, Q5), genome N /
IUPAC / 'X' bytes, read N.  Whatever the oracle computes the GPU must reproduce bit for bit; whatever the oracle rejects
the GPU must reject too.
"""
import numpy as np
import pytest

import oracle_binding as ob
from compare import assert_rows_equal
from portcullis_b200 import _lib as L
from portcullis_b200.columnar import from_records
from test_gpu_parity import gpu_run

pytestmark = pytest.mark.gpu


def random_read(rng, genome_u, tlen, anchors):
    """Returns (pos, cigar, seq). `anchors` is a shared list of splice coordinates so that reads share junctions."""
    pos = int(rng.integers(20, tlen - 1500))
    ops = []
    seq = []
    r = pos
    if rng.random() < 0.10:
        ops.append((int(rng.integers(1, 6)), "H"))
    if rng.random() < 0.20:
        k = int(rng.integers(1, 8)); ops.append((k, "S")); seq += list(rng.choice(list("ACGT"), k))
    n_exons = int(rng.integers(1, 6))
    for e in range(n_exons):
        # exon body: a run of M/=/X with optional I / D / P inside
        n_seg = int(rng.integers(1, 4))
        for sgm in range(n_seg):
            ln = int(rng.integers(1, 40))
            op = "M" if rng.random() < 0.8 else ("=" if rng.random() < 0.5 else "X")
            ops.append((ln, op)); seq += list(genome_u[r:r + ln]); r += ln
            if sgm < n_seg - 1:
                c = rng.random()
                if c < 0.4:
                    k = int(rng.integers(1, 4)); ops.append((k, "I")); seq += list(rng.choice(list("ACGT"), k))
                elif c < 0.8:
                    k = int(rng.integers(1, 4)); ops.append((k, "D")); r += k
                else:
                    ops.append((int(rng.integers(1, 3)), "P"))
        if e < n_exons - 1:
            # indel glued to the splice site on either side
            if rng.random() < 0.15:
                k = int(rng.integers(1, 3)); ops.append((k, "I")); seq += list(rng.choice(list("ACGT"), k))
            elif rng.random() < 0.10:
                k = int(rng.integers(1, 3)); ops.append((k, "D")); r += k
            # snap the donor to a shared coordinate when one is near, so that junctions collect several reads
            near = [a for a in anchors if 0 < a[0] - r <= 30]
            if near and rng.random() < 0.8:
                a = near[int(rng.integers(0, len(near)))]
                ext = a[0] - r
                ops.append((ext, "M")); seq += list(genome_u[r:r + ext]); r += ext
                il = a[1]
            else:
                il = int(rng.integers(20, 300))
                anchors.append((r, il))
            ops.append((il, "N")); r += il
            if rng.random() < 0.10:
                k = int(rng.integers(1, 3)); ops.append((k, "I")); seq += list(rng.choice(list("ACGT"), k))
    if rng.random() < 0.20:
        k = int(rng.integers(1, 8)); ops.append((k, "S")); seq += list(rng.choice(list("ACGT"), k))
    if rng.random() < 0.05:
        ops.append((int(rng.integers(1, 6)), "H"))
    if r >= tlen - 20:
        return None
    # merge equal neighbours (keeps the CIGAR canonical enough; M next to M after the snap)
    out = []
    for ln, op in ops:
        if out and out[-1][1] == op:
            out[-1] = (out[-1][0] + ln, op)
        else:
            out.append((ln, op))
    seq = [("ACGT"[int(rng.integers(0, 4))] if rng.random() < 0.03 else c) for c in seq]
    if rng.random() < 0.05:
        seq[int(rng.integers(0, len(seq)))] = "N"
    return pos, "".join("%d%s" % o for o in out), "".join(seq)


def make_case(seed, n_reads=250, n_targets=2, tlen=6000, want_records=False):
    rng = np.random.default_rng(seed)
    genomes, records = [], []
    for t in range(n_targets):
        g = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, tlen)].copy()
        for _ in range(6):
            s = int(rng.integers(0, tlen - 30)); g[s:s + int(rng.integers(1, 12))] = ord("N")
        for _ in range(10):
            g[int(rng.integers(0, tlen))] = ord("RYKMSWBDHVXU=*acgtn"[int(rng.integers(0, 19))])
        genomes.append(bytes(g))
        gu = bytes(g).decode().upper()
        anchors = []
        for _ in range(n_reads):
            rd = random_read(rng, gu, tlen, anchors)
            if rd is None:
                continue
            pos, cigar, seq = rd
            flag = int(rng.choice([0, 16, 99, 147, 83, 163, 355, 1024 + 99]))
            xs = [0, "+", "-", "?", "."][int(rng.integers(0, 5))] if "N" in cigar else 0
            records.append(dict(tid=t, pos=pos, flag=flag, mapq=int(rng.choice([0, 3, 29, 30, 60])), cigar=cigar, seq=seq, xs=xs,
                                mtid=t if rng.random() < 0.9 else -1, mpos=int(rng.integers(0, tlen))))
    records.sort(key=lambda r: (r["tid"], r["pos"]))
    for i, r in enumerate(records):
        r["name"] = "f%05d" % i
    if want_records:
        return dict(names=["t%d" % t for t in range(n_targets)], lengths=np.array([tlen] * n_targets, dtype=np.int32), genomes=genomes, records=records)
    return from_records(records), np.array([tlen] * n_targets, dtype=np.int32), genomes


@pytest.mark.parametrize("seed", list(range(24)))
def test_fuzz_against_oracle(seed):
    cols, lengths, genomes = make_case(1000 + seed)
    orient = ["UNKNOWN", "FR", "RF", "FF", "SE"][seed % 5]
    try:
        exp_rows, exp_st = ob.run(cols, lengths, genomes, L.ORIENT[orient])
    except ob.OracleError as e:
        assert e.code == L.PJ_EDATA
        with pytest.raises(L.PjError) as ei:
            gpu_run(cols, lengths, genomes, orient)
        assert ei.value.code == L.PJ_EDATA
        return
    rows, st, _ = gpu_run(cols, lengths, genomes, orient, n_batches=1 + seed % 3, match_group=[0, 1, 2, 4, 8][seed % 5])
    assert_rows_equal(rows, exp_rows, "fuzz seed %d" % seed)
    for f in ("spliced", "unspliced", "sumq", "minq", "maxq"):
        assert np.array_equal(st[f], exp_st[f]), f
