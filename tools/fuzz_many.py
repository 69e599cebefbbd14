#!/usr/bin/env python3
"""One-off fuzz campaign on a GPU box: N random CIGAR cases (tests/test_gpu_fuzz.make_case) through the C ABI against the oracle,
all lanes-per-pair widths, the `--extra` metrics on every fifth case.    python tools/fuzz_many.py 400"""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import oracle_binding as ob
from portcullis_b200 import _lib as L
from test_gpu_fuzz import make_case
from test_gpu_parity import gpu_run
from test_gpu_extra import oracle_extra, gpu_extra
from compare import assert_rows_equal, assert_extra_equal
bad = 0; rej = 0
for seed in range(5000, 5000 + int(sys.argv[1])):
    cols, lengths, genomes = make_case(seed, n_reads=150 + seed % 300)
    orient = ["UNKNOWN", "FR", "RF", "FF", "SE"][seed % 5]
    try:
        exp_rows, exp_st = ob.run(cols, lengths, genomes, L.ORIENT[orient])
    except ob.OracleError as e:
        rej += 1
        try:
            gpu_run(cols, lengths, genomes, orient); print("seed", seed, "GPU accepted what the oracle rejects"); bad += 1
        except L.PjError as ge:
            if ge.code != e.code: print("seed", seed, "codes differ", ge.code, e.code); bad += 1
        continue
    try:
        rows, st, _ = gpu_run(cols, lengths, genomes, orient, n_batches=1 + seed % 3, match_group=[0, 1, 2, 4, 8, 16, 32][seed % 7])
        assert_rows_equal(rows, exp_rows, "seed %d" % seed)
        for f in ("spliced", "unspliced", "sumq", "minq", "maxq"):
            assert np.array_equal(st[f], exp_st[f]), f
        if orient == "UNKNOWN":
            erows, _, ex, capped, maxq = oracle_extra(cols, lengths, genomes)
            rows2, _, x, over = gpu_extra(cols, lengths, genomes, maxq, n_batches=1 + seed % 2)
            assert_extra_equal(x, ex, erows, "extra seed %d" % seed)
    except (AssertionError, L.PjError, ob.OracleError) as e:
        print("seed", seed, "FAILED:", str(e)[:300]); bad += 1
print("done: %d cases, %d rejected by both, %d failures" % (int(sys.argv[1]), rej, bad))
