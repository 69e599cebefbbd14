"""The range plan of the junc driver (pjh_plan_describe / pjh_plan_decode), on CPU.

The driver cuts the BAM-ordered decode tasks into parts (one per GPU) and segments (one shard each); a cut inside a target is
moved to the next record no spliced read spans.  Checked here, with the CPU oracle standing in for the device (checker role):
  * the segments partition the records of the BAM, in file order, for any number of parts / segment size;
  * no spliced record of an earlier segment reaches the first position of a later segment of the same target;
  * oracle(segment) rows concatenated in plan order == oracle(whole BAM) rows — junctions never straddle a cut — and the
    per-target scalars add up.  This is the property that makes shards independent (lib/include/portcullis/intron.hpp:69-73)."""
import json
import os
import subprocess

import numpy as np
import pytest

import oracle_binding as ob
from conftest import make_prep
from portcullis_b200 import junction_builder as jb

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PJSYNTH = os.path.join(ROOT, "portcullis_b200", "bin", "pjsynth")


def _ref_end(cols):
    """last reference base of every record and whether it is spliced"""
    ops = cols["cigar"] & 0xF
    lens = (cols["cigar"] >> 4).astype(np.int64)
    consumes = np.isin(ops, [0, 2, 3, 7, 8])
    csum = np.concatenate([[0], np.cumsum(np.where(consumes, lens, 0))])
    nsum = np.concatenate([[0], np.cumsum(ops == 3)])
    co = cols["cigar_off"].astype(np.int64)
    rlen = csum[co[1:]] - csum[co[:-1]]
    spliced = (nsum[co[1:]] - nsum[co[:-1]]) > 0
    return cols["pos"].astype(np.int64) + rlen - 1, spliced


def _check_plan(prep_dir, n_parts, seg_records, run_oracle=True):
    p = jb.PrepDir(prep_dir)
    whole = p.decode(-1, 2)
    seg, cuts = p.plan(n_parts, seg_records)
    pieces = []
    for part in range(n_parts):
        for s in range(int(seg[part])):
            pieces.append(p.decode_segment(n_parts, part, s, seg_records, threads=2))
    assert sum(len(c["pos"]) for c in pieces) == len(whole["pos"])
    for name in ("tid", "pos", "flag", "mapq", "xs", "l_qseq", "mtid", "mpos"):
        assert np.array_equal(np.concatenate([c[name] for c in pieces]) if pieces else whole[name][:0], whole[name]), name
    assert np.array_equal(np.concatenate([c["cigar"] for c in pieces]), whole["cigar"])
    assert np.array_equal(np.concatenate([c["seq4"] for c in pieces]), whole["seq4"])
    # cut validity
    prev_tid, prev_max_end = None, -1
    for c in pieces:
        if len(c["pos"]) == 0:
            continue
        end, spliced = _ref_end(c)
        for t in np.unique(c["tid"]):
            m = c["tid"] == t
            first_pos = int(c["pos"][m][0])
            if prev_tid == t:
                assert first_pos > prev_max_end, "a spliced read of the previous segment spans the cut (target %d)" % t
        t_last = int(c["tid"][-1])
        m = (c["tid"] == t_last) & spliced
        mx = int(end[m].max()) if m.any() else -1
        prev_max_end = max(prev_max_end, mx) if prev_tid == t_last else mx
        prev_tid = t_last
    if not run_oracle:
        return seg, cuts
    genomes = [p.genome(t) for t in range(len(p.names))]
    rows_w, st_w = ob.run(whole, p.lengths, genomes)
    parts_rows, st_sum = [], None
    for c in pieces:
        if len(c["pos"]) == 0:
            continue
        r, st = ob.run(c, p.lengths, genomes)
        parts_rows.append(r)
        if st_sum is None:
            st_sum = st.copy()
        else:
            for f in ("spliced", "unspliced", "sumq"):
                st_sum[f] += st[f]
            st_sum["minq"] = np.minimum(st_sum["minq"], st["minq"]); st_sum["maxq"] = np.maximum(st_sum["maxq"], st["maxq"])
    rows_c = np.concatenate(parts_rows)
    assert rows_c.tobytes() == rows_w.tobytes(), "rows of the segments do not concatenate to the rows of the whole BAM"
    assert st_sum.tobytes() == st_w.tobytes()
    return seg, cuts


@pytest.mark.parametrize("fixture", ["short_pe", "indel_rich", "long_se", "clipped3"])
@pytest.mark.parametrize("n_parts,seg_records", [(1, 0), (2, 0), (3, 200), (8, 50)])
def test_fixture_plans_partition_and_preserve_rows(tmp_path, fixture, n_parts, seg_records):
    _check_plan(make_prep(tmp_path, fixture), n_parts, seg_records)


@pytest.mark.parametrize("preset,scale", [("c2", 0.02), ("c4", 0.01), ("c5", 0.02), ("c3", 0.002)])
def test_synthetic_presets_are_cut_inside_targets(tmp_path, preset, scale):
    d = str(tmp_path / "prep")
    subprocess.check_call([PJSYNTH, "--preset", preset, "--scale", str(scale), "--out", d, "--threads", "4"], stderr=subprocess.DEVNULL)
    meta = json.load(open(os.path.join(d, "synth.json")))
    seg, cuts = _check_plan(d, 8, max(2000, meta["n_records"] // 40), run_oracle=(preset in ("c2", "c5")))
    assert seg.sum() >= 8
    if meta["n_targets"] < 16:
        assert cuts > 0, "with fewer targets than segments some cuts must fall inside a target"


def test_single_part_default_budget_is_one_segment(tmp_path):
    p = jb.PrepDir(make_prep(tmp_path, "short_pe"))
    seg, cuts = p.plan(1)
    assert seg.tolist() == [1] and cuts == 0
    seg, cuts = p.plan(4, whole_targets=True)
    assert cuts == 0 and seg.max() <= 1
