"""The lean batch form (pj_batch.lean, include/portcullis_junc.h): what the host decoder ships over PCIe.

CPU part: the lean decode of a segment carries exactly the information of the classic columns (same records, same CIGAR words,
2-bit SEQ + exception list == 4-bit SEQ).  GPU part: lean batches submitted through the C ABI give the same rows, byte for
byte, as the classic batches, and malformed lean batches are rejected."""
import os

import numpy as np
import pytest

from conftest import make_prep
from portcullis_b200 import _lib as L
from portcullis_b200 import junction_builder as jb

NT16 = "=ACMGRSVTWYHKDBN"


def _unpack4(b4, l):
    nib = np.empty(2 * len(b4), np.uint8)
    nib[0::2] = b4 >> 4
    nib[1::2] = b4 & 15
    return nib[:l]


def _check_lean_equals_classic(prep_dir, keep_mate):
    p = jb.PrepDir(prep_dir)
    cols = p.decode(-1, 2)
    runs = p.decode_segment_lean(1, 0, 0, 1 << 40, threads=2, keep_mate=keep_mate, copy=True)
    assert sum(len(r["pos"]) for r in runs) == len(cols["pos"])
    assert np.array_equal(np.concatenate([np.full(len(r["pos"]), r["tid"], np.int32) for r in runs]), cols["tid"])
    for name in ("pos", "mapq", "xs", "l_qseq") + (("mtid", "mpos") if keep_mate else ()):
        assert np.array_equal(np.concatenate([r[name] for r in runs]), cols[name]), name
    assert ("mtid" in runs[0]) == keep_mate
    flag = np.concatenate([r["flag"] for r in runs])
    assert np.array_equal(flag & 0x7fff, cols["flag"])
    assert np.array_equal(np.concatenate([r["n_cigar"] for r in runs]), np.diff(cols["cigar_off"].astype(np.int64)))
    assert np.array_equal(np.concatenate([r["cigar"] for r in runs]), cols["cigar"])
    so = cols["seq_off"].astype(np.int64)
    n_exc = 0
    i = 0
    for r in runs:
        xp = {int(a): int(c) for a, c in zip(r["seqx_pos"], r["seqx_code"])}
        n_exc += len(xp)
        o2 = 0
        for k in range(len(r["pos"])):
            l = int(cols["l_qseq"][i])
            if so[i + 1] > so[i] and l > 0:
                nib = _unpack4(cols["seq4"][so[i]:so[i + 1]], l)
                nb2 = (l + 3) // 4
                b2 = r["seq2"][o2:o2 + nb2]
                code = np.empty(4 * nb2, np.uint8)
                for q in range(4):
                    code[q::4] = (b2 >> (2 * q)) & 3
                chars = ["ACGT"[c] for c in code[:l]]
                has_x = False
                for q in range(l):
                    if o2 * 4 + q in xp:
                        chars[q] = NT16[xp[o2 * 4 + q]]
                        has_x = True
                assert "".join(chars) == "".join(NT16[v] for v in nib), "record %d" % i
                assert bool(r["flag"][k] & 0x8000) == has_x
                o2 += nb2
            else:
                assert not (r["flag"][k] & 0x8000)
            i += 1
        assert o2 == len(r["seq2"])
    return n_exc


@pytest.mark.parametrize("fixture", ["kat", "short_pe", "indel_rich", "long_se", "clipped3"])
def test_lean_decode_carries_the_classic_columns(tmp_path, fixture):
    _check_lean_equals_classic(make_prep(tmp_path, fixture), keep_mate=(fixture != "long_se"))


def _gpu_rows(submit, lengths, genomes, orientation="UNKNOWN"):
    g = jb.JuncGpu(0, orientation)
    try:
        g.set_targets(lengths)
        for t, s in enumerate(genomes):
            g.set_genome(t, s)
        g.shard_begin(1024, 0, 0)
        submit(g)
        g.run()
        rows, st = g.fetch()
    finally:
        g.close()
    return rows, st


@pytest.mark.gpu
@pytest.mark.parametrize("fixture,orient", [("kat", "FR"), ("short_pe", "UNKNOWN"), ("short_pe", "FR"), ("indel_rich", "RF"), ("long_se", "UNKNOWN"), ("clipped3", "FR")])
def test_lean_batches_give_the_rows_of_classic_batches(tmp_path, fixture, orient):
    p = jb.PrepDir(make_prep(tmp_path, fixture))
    genomes = [p.genome(t) for t in range(len(p.names))]
    cols = p.decode(-1, 2)
    rows_c, st_c = _gpu_rows(lambda g: g.submit(cols), p.lengths, genomes, orient)
    runs = p.decode_segment_lean(1, 0, 0, 1 << 40, threads=2, keep_mate=(orient != "UNKNOWN"), copy=True)
    rows_l, st_l = _gpu_rows(lambda g: [g.submit_lean(r) for r in runs], p.lengths, genomes, orient)
    assert rows_l.tobytes() == rows_c.tobytes()
    assert st_l.tobytes() == st_c.tobytes()
    # several small lean batches per target (records 0..k, k..n): the device-side offset scans must chain
    def split(g):
        for r in runs:
            n = len(r["pos"]); k = n // 3
            off = np.concatenate([[0], np.cumsum(r["n_cigar"].astype(np.int64))])
            isn = np.concatenate([[0], np.cumsum((r["cigar"] & 15) == 3)])
            spl = ((isn[off[1:]] - isn[off[:-1]]) > 0) & (r["l_qseq"] > 0)
            sb = np.concatenate([[0], np.cumsum(np.where(spl, (r["l_qseq"].astype(np.int64) + 3) // 4, 0))])
            for a, b in ((0, k), (k, n)):
                if a == b:
                    continue
                d = {"tid": r["tid"]}
                for name in ("pos", "flag", "mapq", "xs", "l_qseq", "n_cigar", "mtid", "mpos"):
                    if name in r:
                        d[name] = r[name][a:b]
                d["cigar"] = r["cigar"][off[a]:off[b]]
                d["seq2"] = r["seq2"][sb[a]:sb[b]]
                m = (r["seqx_pos"] >= sb[a] * 4) & (r["seqx_pos"] < sb[b] * 4)
                d["seqx_pos"] = r["seqx_pos"][m] - np.uint64(sb[a] * 4)
                d["seqx_code"] = r["seqx_code"][m]
                g.submit_lean(d)
    rows_s, _ = _gpu_rows(split, p.lengths, genomes, orient)
    assert rows_s.tobytes() == rows_c.tobytes()


@pytest.mark.gpu
def test_malformed_lean_batches_are_rejected(tmp_path):
    p = jb.PrepDir(make_prep(tmp_path, "short_pe"))
    genomes = [p.genome(t) for t in range(len(p.names))]
    runs = p.decode_segment_lean(1, 0, 0, 1 << 40, threads=2, keep_mate=False, copy=True)
    r = dict(runs[0])
    bad = dict(r); bad["cigar"] = r["cigar"][:-1]                       # op counts do not add up to the CIGAR words
    with pytest.raises(L.PjError):
        _gpu_rows(lambda g: g.submit_lean(bad), p.lengths, genomes)
    bad = dict(r); bad["seq2"] = r["seq2"][:-1]                         # seq2 bytes do not match the spliced records
    with pytest.raises(L.PjError):
        _gpu_rows(lambda g: g.submit_lean(bad), p.lengths, genomes)
    with pytest.raises(L.PjError):                                       # mate columns are required when the orientation uses them
        _gpu_rows(lambda g: g.submit_lean(r), p.lengths, genomes, "FR")
    bad = dict(r); bad["tid"] = 99
    with pytest.raises(L.PjError):
        _gpu_rows(lambda g: g.submit_lean(bad), p.lengths, genomes)


def _synthetic_prep(tmp_path, seed):
    """A random data set with N / IUPAC bases in reads and genome (tests/synth.py), written as a prep directory through the
    reference's htslib (oracle/_ref/bamtool)."""
    import refrun
    import synth
    ds = synth.make_dataset(seed, n_targets=3, target_len=30000, genes_per_target=8, reads_per_gene=(20, 120))
    return refrun.make_prep_dir(ds, str(tmp_path / ("ds%d" % seed)))


def test_lean_decode_lists_read_bases_that_are_not_acgt(tmp_path):
    import oracle_binding as ob
    if not os.path.exists(ob.BAMTOOL):
        pytest.skip("oracle/_ref/bamtool not built")
    n_exc = _check_lean_equals_classic(_synthetic_prep(tmp_path, 11), keep_mate=True)
    assert n_exc > 0, "the synthetic reads carry N bases: the exception list must not be empty"


@pytest.mark.gpu
@pytest.mark.parametrize("seed", [11, 12])
def test_lean_batches_with_exceptions_match_oracle(tmp_path, seed):
    import oracle_binding as ob
    from compare import assert_rows_equal
    prep = _synthetic_prep(tmp_path, seed)
    p = jb.PrepDir(prep)
    genomes = [p.genome(t) for t in range(len(p.names))]
    cols = p.decode(-1, 2)
    exp, exp_st = ob.run(cols, p.lengths, genomes)
    runs = p.decode_segment_lean(1, 0, 0, 1 << 40, threads=2, keep_mate=False, copy=True)
    assert sum(len(r["seqx_pos"]) for r in runs) > 0
    rows, st = _gpu_rows(lambda g: [g.submit_lean(r) for r in runs], p.lengths, genomes)
    assert_rows_equal(rows, exp, "lean vs oracle")
    assert np.array_equal(st["spliced"], exp_st["spliced"])


def test_lean_seq_repack_at_block_boundaries(tmp_path):
    """The 4-bit -> 2-bit SEQ repack works 32 bases at a time (SSSE3) with a scalar tail: read lengths on every side of the 32- and
    4-base boundaries, odd lengths (padding nibble), N / IUPAC bases in the first, last and boundary positions, all-N reads."""
    import random
    import oracle_binding as ob
    import refrun
    if not os.path.exists(ob.BAMTOOL):
        pytest.skip("oracle/_ref/bamtool not built")
    rng = random.Random(3)
    genome = "".join(rng.choice("ACGT") for _ in range(60000)).encode()
    recs, pos = [], 100
    for l in [2, 3, 4, 5, 7, 8, 9, 15, 16, 17, 30, 31, 32, 33, 34, 35, 36, 47, 48, 49, 63, 64, 65, 66, 95, 96, 97, 127, 128, 129, 160, 161, 255, 256, 257, 1001]:
        for variant in range(4):
            seq = [rng.choice("ACGT") for _ in range(l)]
            if variant == 1:
                for q in (0, l - 1, 15, 16, 31, 32, 33, 63, 64):
                    if q < l:
                        seq[q] = rng.choice("NRYKMSWBDHV")
            if variant == 2:
                seq = ["N"] * l
            if variant == 3 and l > 40:
                seq[rng.randrange(32, l)] = "N"                   # only in the scalar tail or a later block
            a = max(1, l // 2)
            recs.append(dict(name="r%d_%d" % (l, variant), flag=0, tid=0, pos=pos, mapq=60, cigar="%dM200N%dM" % (a, l - a) if l - a > 0 else "%dM" % a,
                             mtid=-1, mpos=-1, seq="".join(seq), xs="+"))
            pos += 7
    ds = dict(names=["chr1"], lengths=np.array([len(genome)], dtype=np.int32), genomes=[genome], records=recs)
    n_exc = _check_lean_equals_classic(refrun.make_prep_dir(ds, str(tmp_path / "edges")), keep_mate=False)
    assert n_exc > 100
