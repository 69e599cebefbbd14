"""CPU: the oracle against the UNMODIFIED reference binary on randomised CIGARs (whole alphabet M I D N S H P = X, indels
glued to splice sites, leading H+S, IUPAC / 'X' / '=' genome bytes, every orientation).  Runs wherever oracle/_ref has
been built (this container; `make -C oracle ref`)."""
import filecmp
import os

import numpy as np
import pytest

import oracle_binding as ob
import refrun
import synth
from portcullis_b200 import _lib as L
from portcullis_b200 import junction_builder as jb
from test_gpu_fuzz import make_case


@pytest.mark.parametrize("seed", [0, 1, 2, 3, 4, 7, 13])
def test_oracle_equals_reference_on_fuzzed_cigars(tmp_path, seed):
    if not (os.path.exists(ob.REF_BIN) and os.path.exists(ob.BAMTOOL)):
        pytest.skip("oracle/_ref not built")
    orient = ["UNKNOWN", "FR", "RF", "FF", "SE"][seed % 5]
    ds = make_case(1000 + seed, want_records=True)
    wd = str(tmp_path)
    prep = refrun.make_prep_dir(ds, wd)
    refrun.run_reference(prep, os.path.join(wd, "ref", "k"), orientation=None if orient == "UNKNOWN" else orient, extra=True)
    p = jb.PrepDir(prep)
    cols = p.decode(-1, 2, names=True)
    direct = synth.to_columns(ds)
    for k in direct:
        assert np.array_equal(direct[k], cols[k]), k
    rows, st = ob.run(cols, p.lengths, [p.genome(t) for t in range(len(p.names))], L.ORIENT[orient])
    total = float(st["spliced"].sum() + st["unspliced"].sum())
    fin = jb.finalize(rows.copy(), float(st["sumq"].sum()) / total)
    x, _ = ob.extra(cols, p.lengths, fin, int(st["maxq"].max()))     # the reference ran with --extra: all 75 columns are compared
    jb.write_outputs(os.path.join(wd, "mine", "k"), fin, p.names, p.lengths, exon_gff=True, intron_gff=True, extra=x)
    for ext in ("tab", "bed", "exon.gff3", "intron.gff3"):
        assert filecmp.cmp(os.path.join(wd, "mine", "k.junctions." + ext), os.path.join(wd, "ref", "k.junctions." + ext), shallow=False), ext
