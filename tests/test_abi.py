"""CPU tests of the drop-in boundary: the shared library loads, exports every symbol the headers declare, and
refuses to compute without a GPU (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from portcullis_b200 import _lib as L

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    names = set()
    for h in ("portcullis_junc.h", "portcullis_junc_host.h"):
        src = open(os.path.join(ROOT, "include", h)).read()
        src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
        names |= set(re.findall(r"\b(pjh?_[a-z0-9_]+)\s*\(", src))
    return names


def test_library_exports_every_declared_symbol():
    lib = L.load()
    declared = _declared_symbols()
    assert len(declared) >= 30
    for name in sorted(declared):
        assert hasattr(lib, name), "symbol %s declared in include/ but not exported" % name
    assert declared == set(L.SYMBOLS), "python binding table out of sync with the headers: %s" % (declared ^ set(L.SYMBOLS))
    assert lib.pj_abi_version() == 2
    assert lib.pj_junction_size() == L.JUNCTION_DTYPE.itemsize == 256


def test_product_does_not_reference_the_oracle():
    """Nothing under portcullis_b200/ may import, link or call oracle/."""
    bad = []
    for dp, _, files in os.walk(os.path.join(ROOT, "portcullis_b200")):
        for f in files:
            if f.endswith((".py", ".cpp", ".cu", ".cuh", ".hpp", ".h")) or f == "Makefile":
                s = open(os.path.join(dp, f), errors="ignore").read()
                if re.search(r"oracle|oj_run|liboracle", s):
                    bad.append(f)
    assert not bad, bad


def test_no_gpu_means_loud_failure():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    lib = L.load()
    cfg = L.PjConfig()
    cfg.device, cfg.orientation = 0, L.ORIENT["UNKNOWN"]
    ctx = C.c_void_p()
    rc = lib.pj_create(C.byref(cfg), C.byref(ctx))
    assert rc == L.PJ_ECUDA and not ctx.value
    assert b"no CPU fallback" in lib.pj_global_last_error()
    from portcullis_b200 import JunctionBuilder
    from conftest import make_prep
    import tempfile
    with tempfile.TemporaryDirectory() as d:
        jbd = JunctionBuilder(make_prep(d, "kat"), os.path.join(d, "out", "k"))
        with pytest.raises(L.PjError) as ei:
            jbd.process()
        assert ei.value.code == L.PJ_ECUDA


def test_host_finalize_edge_cases():
    """A13 boundary behaviour (quirk Q8): J=0, J=1 (stats skipped), J=2 same target, cross-target sentinels."""
    from portcullis_b200 import junction_builder as jb
    import oracle_binding as ob

    def rows(spec):
        r = np.zeros(len(spec), dtype=L.JUNCTION_DTYPE)
        for i, (tid, s, e, raw) in enumerate(spec):
            r[i]["tid"], r[i]["start"], r[i]["end"], r[i]["nb_raw_aln"], r[i]["nb_rel_aln"] = tid, s, e, raw, raw
        return r
    for spec in ([], [(0, 10, 20, 3)], [(0, 10, 20, 3), (0, 30, 40, 5)], [(0, 10, 20, 3), (1, 30, 40, 5)],
                 [(0, 10, 20, 3), (0, 10, 40, 5), (0, 35, 40, 5), (1, 5, 9, 1), (2, 5, 9, 1), (2, 50, 90, 2)]):
        a = jb.finalize(rows(spec), 100.5)
        b = ob.finalize(rows(spec), 100.5)
        assert a.tobytes() == b.tobytes()
    two = jb.finalize(rows([(0, 10, 20, 3), (0, 30, 40, 5)]), 100.5)
    assert list(two["dist_2_down_junc"]) == [0xFFFFFFFF, 10] and list(two["dist_2_up_junc"]) == [10, 0]
    one = jb.finalize(rows([(0, 10, 20, 3)]), 100.5)
    assert one["mean_readlen"][0] == 0 and one["uniq_junc"][0] == 0


def test_host_finalize_parallel_sweep_equals_the_pairwise_walk():
    """A12/A13 on the host is one parallel sweep with per-row closed forms of the reference's pair-by-pair walk (junction_system.cc:250-320);
    the oracle restates the walk itself.  Large random tables (several host threads, groups that straddle thread ranges, many
    single-junction targets, target changes next to the ends) must come out byte-identical, sorted or not on entry."""
    from portcullis_b200 import junction_builder as jb
    import oracle_binding as ob
    rng = np.random.default_rng(7)
    for n, n_targets, span in ((200_000, 7, 60_000), (150_001, 40_000, 500), (70_000, 3, 2_000_000), (5, 5, 100), (3, 1, 50)):
        tid = rng.integers(0, n_targets, n)
        start = rng.integers(10, span, n)
        length = rng.integers(1, 40, n) * 10
        key = (tid.astype(np.int64) << 44) | (start.astype(np.int64) << 20) | length
        _, first = np.unique(key, return_index=True)                     # junction keys are unique
        r = np.zeros(len(first), dtype=L.JUNCTION_DTYPE)
        r["tid"], r["start"], r["end"] = tid[first], start[first], (start + length)[first]
        r["nb_raw_aln"] = rng.integers(1, 50, len(r)); r["nb_rel_aln"] = rng.integers(0, 50, len(r)) % (r["nb_raw_aln"] + 1)
        r["nb_mismatches"] = rng.integers(0, 200, len(r)); r["maxmmes"] = rng.integers(0, 60, len(r)); r["suspicious"] = rng.integers(0, 2, len(r))
        order = np.lexsort((r["end"], r["start"], r["tid"]))
        for rows in (r[order], r[rng.permutation(len(r))]):
            a = jb.finalize(rows.copy(), 151.25)
            b = ob.finalize(rows.copy(), 151.25)
            assert a.tobytes() == b.tobytes()


def test_fast_inflate_matches_zlib():
    """The built-in DEFLATE decoder used for BGZF blocks agrees with zlib on 300 synthetic streams (random, DNA-like,
    run-heavy, LZ-heavy, all-zero data; stored / fixed / dynamic blocks; several strategies)."""
    assert L.load().pjh_inflate_selftest(300) == 0


def test_writer_number_formatting_equals_printf():
    # the writers' fast "%g" / "%.3f" / integer paths must print exactly what printf prints (byte-identical tables)
    assert L.load().pjh_format_selftest(20000) == 0


def test_headers_are_plain_c_and_link(tmp_path):
    """include/*.h compile as C99 and a C program can link the library and call it (no C++ or torch types in the ABI)."""
    import subprocess
    src = tmp_path / "abi.c"
    src.write_text('''
#include "portcullis_junc.h"
#include "portcullis_junc_host.h"
#include <stdio.h>
int main(void) {
    pj_config cfg = {0}; pj_ctx* ctx = 0; pjh_options o; pj_junction row;
    pjh_options_default(&o);
    cfg.device = 0; cfg.orientation = PJ_ORIENT_UNKNOWN;
    int rc = pj_create(&cfg, &ctx);          /* no GPU here: must fail loudly, not fall back */
    printf("%d %d %d %s\\n", pj_abi_version(), pj_junction_size(), (int)sizeof row, rc == PJ_OK ? "gpu" : pj_global_last_error());
    if (ctx) pj_destroy(ctx);
    return (pj_junction_size() == (int)sizeof row && o.threads == 1) ? 0 : 1;
}
''')
    exe = tmp_path / "abi"
    libdir = os.path.join(ROOT, "portcullis_b200")
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe),
                           "-L", libdir, "-lportcullis_junc", "-Wl,-rpath," + libdir])
    out = subprocess.run([str(exe)], stdout=subprocess.PIPE, text=True, check=True).stdout.split()
    assert out[0] == "2" and out[1] == "256" and out[2] == "256"


def test_corrupted_bam_and_index_never_crash_the_readers():
    """80 + 80 random corruptions of a fixture's BAM / BAI (tools/fuzz_decoder.py, in a child process): open, plan, decode and
    `--separate` either work or raise PjError; a crash would show as a non-zero exit without the summary line."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for mode in ("bam", "bai"):
        p = subprocess.run([sys.executable, os.path.join(root, "tools", "fuzz_decoder.py"), mode, "7000", "80"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
        assert p.returncode == 0 and ("mode %s cases 80" % mode) in p.stdout, p.stderr[-600:]


def _bgzf_blocks(buf):
    off = 0
    while off + 18 <= len(buf):
        xlen = int.from_bytes(buf[off + 10:off + 12], "little")
        bsize = int.from_bytes(buf[off + 16:off + 18], "little") + 1
        yield off, xlen, bsize
        off += bsize


def test_bgzf_crc_mismatch_is_detected(tmp_path):
    """a block whose payload inflates cleanly to ISIZE bytes but does not match its CRC32 trailer (bit-rot, or a
    decoder bug) must be rejected by both readers, as htslib's bgzf_read does — not flow into the junction counts."""
    import zlib
    from portcullis_b200 import junction_builder as jb, _lib as L
    src = os.path.join(os.path.dirname(__file__), "golden", "short_pe")
    bam = bytearray(open(src + "/reads.bam", "rb").read())
    blocks = [b for b in _bgzf_blocks(bam) if b[2] > 200]
    off, xlen, bsize = blocks[len(blocks) // 2]
    raw = bytearray(zlib.decompress(bytes(bam[off + 12 + xlen:off + bsize - 8]), -15))
    raw[len(raw) // 2] ^= 0x01                                    # one flipped bit in the uncompressed records
    comp = zlib.compressobj(6, zlib.DEFLATED, -15)
    payload = comp.compress(bytes(raw)) + comp.flush()
    new_block = bytearray(bam[off:off + 12 + xlen]) + payload + bam[off + bsize - 8:off + bsize]     # stale CRC, same ISIZE
    new_block[16:18] = (len(new_block) - 1).to_bytes(2, "little")
    out = bam[:off] + new_block + bam[off + bsize:]
    d = str(tmp_path / "p")
    os.makedirs(d)
    open(d + "/portcullis.sorted.alignments.bam", "wb").write(bytes(out))
    for a, b in (("reads.bam.bai", "portcullis.sorted.alignments.bam.bai"), ("genome.fa", "portcullis.genome.fa"), ("genome.fa.fai", "portcullis.genome.fa.fai")):
        os.symlink(os.path.join(src, a), os.path.join(d, b))
    p = jb.PrepDir(d)                                             # the index still points at valid block starts up to the patched block
    with pytest.raises(L.PjError, match="CRC32"):
        p.decode(-1, 2)
    with pytest.raises(L.PjError, match="CRC32"):
        jb.separate_bams(d, str(tmp_path / "s" / "p"), threads=2)
