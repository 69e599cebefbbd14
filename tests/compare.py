"""Comparison rules of SURVEY §8(d): integer/string columns bit-exact, FP columns equal as text or within 1e-6 relative."""
import numpy as np

from portcullis_b200 import _lib as L

FP_TOL = 1e-6   # relative tolerance for entropy / ratios (north star)


def assert_rows_equal(got, exp, what="rows", finalize_fields=False):
    assert len(got) == len(exp), "%s: %d junctions, expected %d" % (what, len(got), len(exp))
    for f in L.DEVICE_INT_FIELDS + (L.FINALIZE_FIELDS if finalize_fields else []):
        a, b = got[f], exp[f]
        if a.dtype.kind == "f":
            ok = np.isclose(a, b, rtol=FP_TOL, atol=0)
        else:
            ok = a == b
        if ok.ndim > 1:
            ok = ok.all(axis=1)
        if not ok.all():
            i = int(np.flatnonzero(~ok)[0])
            raise AssertionError("%s: field %s differs at junction %d (tid %d, %d-%d): got %r expected %r; %d of %d rows differ"
                                 % (what, f, i, exp["tid"][i], exp["start"][i], exp["end"][i], a[i], b[i], int((~ok).sum()), len(ok)))
    e = np.isclose(got["entropy"], exp["entropy"], rtol=FP_TOL, atol=1e-12)
    if not e.all():
        i = int(np.flatnonzero(~e)[0])
        raise AssertionError("%s: entropy differs at junction %d: got %r expected %r" % (what, i, got["entropy"][i], exp["entropy"][i]))


FP_COLS = {27, 32, 33, 50, 51}    # 0-based: rel2raw, entropy, mean_mismatches; mm_score and coverage (`--extra`)


def assert_tab_equal(path_got, path_exp):
    """junctions.tab: every column byte-equal; FP columns equal as text or within 1e-6 relative."""
    with open(path_got) as f:
        g = f.read().split("\n")
    with open(path_exp) as f:
        e = f.read().split("\n")
    assert len(g) == len(e), "line count %d vs %d" % (len(g), len(e))
    for ln, (a, b) in enumerate(zip(g, e)):
        if a == b:
            continue
        ca, cb = a.split("\t"), b.split("\t")
        assert len(ca) == len(cb), "line %d: column count" % ln
        for k, (x, y) in enumerate(zip(ca, cb)):
            if x == y:
                continue
            assert k in FP_COLS, "line %d column %d: %r != %r" % (ln, k, x, y)
            assert abs(float(x) - float(y)) <= FP_TOL * abs(float(y)), "line %d column %d: %r vs %r" % (ln, k, x, y)


def assert_exon_gff_equal(path_got, path_exp):
    import re
    pat = re.compile(r"(ent:|Entropy=)([-0-9.e+]+)")
    with open(path_got) as f:
        g = f.read().split("\n")
    with open(path_exp) as f:
        e = f.read().split("\n")
    assert len(g) == len(e)
    for ln, (a, b) in enumerate(zip(g, e)):
        if a == b:
            continue
        va, vb = pat.findall(a), pat.findall(b)
        assert pat.sub(r"\1#", a) == pat.sub(r"\1#", b), "line %d differs outside entropy fields" % ln
        for (_, x), (_, y) in zip(va, vb):
            assert abs(float(x) - float(y)) <= FP_TOL * max(abs(float(y)), 1e-300), "line %d: %s vs %s" % (ln, x, y)


EXTRA_INT_FIELDS = ["up_aln", "down_aln", "mm_n", "mm_m", "cov_sum"]


def assert_extra_equal(got, exp, rows, what="extra"):
    """pj_junction_extra arrays: integer fields bit-exact, the two doubles bit-equal too (same operands, same order)."""
    assert len(got) == len(exp)
    for f in EXTRA_INT_FIELDS + ["mm_score", "coverage"]:
        a, b = got[f], exp[f]
        ok = (a == b) | ((a != a) & (b != b)) if a.dtype.kind == "f" else a == b
        if ok.ndim > 1:
            ok = ok.all(axis=1)
        if not ok.all():
            i = int(np.flatnonzero(~ok)[0])
            raise AssertionError("%s: field %s differs at junction %d (tid %d, %d-%d): got %r expected %r; %d of %d rows differ"
                                 % (what, f, i, rows["tid"][i], rows["start"][i], rows["end"][i], a[i], b[i], int((~ok).sum()), len(ok)))


def extra_tab_columns(x):
    """The four junctions.tab columns (mm_score, coverage, up_aln, down_aln) as the writer prints them (ostream << double
    = %.6g, junction.cc:105-108)."""
    return [["%.6g" % r["mm_score"], "%.6g" % r["coverage"], str(int(r["up_aln"])), str(int(r["down_aln"]))] for r in x]
