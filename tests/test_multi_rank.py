"""Host-side logic of the multi-GPU path, on CPU with a world_size-2 gloo group: every rank computes the same shard
plan, processes only its own targets (here with the CPU oracle standing in for the device — this is the checker
role, the product never does this), rank 0 gathers the rows and runs the host finalize.  The result must equal the
single-process run: junctions never span targets, so the only cross-shard coupling is A12/A13."""
import os
import socket
import sys
import tempfile

import numpy as np
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker_ranges(rank, world, port, prep_dir, outfile, seg_records):
    """The one-process-per-GPU flow of bench.py / pjh_junc_run_part: every rank computes the same range plan, handles the
    segments of its own part (oracle standing in for the device), rank 0 concatenates in part order and finalizes."""
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_binding as ob
    from portcullis_b200 import junction_builder as jb
    from portcullis_b200 import _lib as L
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    p = jb.PrepDir(prep_dir)
    seg, cuts = p.plan(world, seg_records)
    genomes = [p.genome(t) for t in range(len(p.names))]
    rows_parts, stats = [], None
    for s in range(int(seg[rank])):
        cols = p.decode_segment(world, rank, s, seg_records)
        if len(cols["pos"]) == 0:
            continue
        r, st = ob.run(cols, p.lengths, genomes)
        rows_parts.append(r)
        st = st.astype(jb.TARGET_STATS_DTYPE)
        stats = st if stats is None else jb.merge_target_stats([stats, st])
    rows = np.concatenate(rows_parts) if rows_parts else np.zeros(0, dtype=L.JUNCTION_DTYPE)
    gathered = [None] * world
    dist.all_gather_object(gathered, (seg.tolist(), rows.tobytes(), None if stats is None else stats.tobytes()))
    if rank == 0:
        assert all(g[0] == gathered[0][0] for g in gathered), "ranks disagree on the plan"
        allrows = np.concatenate([np.frombuffer(g[1], dtype=L.JUNCTION_DTYPE) for g in gathered])      # part order, no sort
        st = jb.merge_target_stats([np.frombuffer(g[2], dtype=jb.TARGET_STATS_DTYPE) for g in gathered if g[2] is not None])
        fin = jb.finalize(allrows.copy(), float(st["sumq"].sum()) / float(st["spliced"].sum() + st["unspliced"].sum()))
        np.save(outfile, fin)
    dist.barrier()
    dist.destroy_process_group()


def _worker(rank, world, port, prep_dir, outfile):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_binding as ob
    from portcullis_b200 import junction_builder as jb
    from portcullis_b200 import _lib as L
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    p = jb.PrepDir(prep_dir)
    owner = p.plan_shards(world)
    mine = [t for t in range(len(p.names)) if owner[t] == rank]
    genomes = [p.genome(t) if t in mine else b"" for t in range(len(p.names))]
    rows_parts, stats = [], None
    for t in mine:                                   # whole targets, BAM order inside the shard
        cols = p.decode(t, 1)
        r, st = ob.run(cols, p.lengths, genomes)
        rows_parts.append(r)
        stats = st if stats is None else np.array([tuple(max(a, b) if n in ("maxq",) else (min(a, b) if n == "minq" else a + b)
                                                         for n, a, b in zip(st.dtype.names, x, y)) for x, y in zip(stats, st)], dtype=st.dtype)
    rows = np.concatenate(rows_parts) if rows_parts else np.zeros(0, dtype=L.JUNCTION_DTYPE)
    gathered = [None] * world
    dist.all_gather_object(gathered, (owner.tolist(), rows.tobytes(), None if stats is None else stats.tolist()))
    if rank == 0:
        assert all(g[0] == gathered[0][0] for g in gathered), "ranks disagree on the shard plan"
        allrows = np.concatenate([np.frombuffer(g[1], dtype=L.JUNCTION_DTYPE) for g in gathered])
        spliced = unspliced = sumq = 0
        for g in gathered:
            for s in (g[2] or []):
                spliced += s[0]; unspliced += s[1]; sumq += s[2]
        fin = jb.finalize(allrows.copy(), sumq / (spliced + unspliced))
        np.save(outfile, fin)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_equals_single_process():
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_binding as ob
    from conftest import make_prep
    from portcullis_b200 import junction_builder as jb
    with tempfile.TemporaryDirectory() as d:
        prep = make_prep(d, "indel_rich")            # three targets -> both ranks own work
        out = os.path.join(d, "rows.npy")
        mp.spawn(_worker, args=(2, _free_port(), prep, out), nprocs=2, join=True)
        got = np.load(out)
        p = jb.PrepDir(prep)
        cols = p.decode(-1, 2)
        rows, st = ob.run(cols, p.lengths, [p.genome(t) for t in range(len(p.names))])
        exp = jb.finalize(rows.copy(), float(st["sumq"].sum()) / float(st["spliced"].sum() + st["unspliced"].sum()))
        assert got.tobytes() == exp.tobytes()
        owner = p.plan_shards(2)
        assert set(owner.tolist()) == {0, 1}


def test_shard_plan_is_balanced_and_deterministic():
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from conftest import make_prep
    from portcullis_b200 import junction_builder as jb
    with tempfile.TemporaryDirectory() as d:
        p = jb.PrepDir(make_prep(d, "indel_rich"))
        a, b = p.plan_shards(2), p.plan_shards(2)
        assert a.tolist() == b.tolist()
        assert p.plan_shards(1).tolist() == [0] * len(p.names)
        w = np.array([p.target_records(t) for t in range(len(p.names))], dtype=np.int64)
        loads = [w[a == g].sum() for g in range(2)]
        assert max(loads) <= w.sum() - min(w)        # LPT never leaves a GPU empty when there are >= n_gpus targets


def test_two_rank_range_plan_equals_single_process():
    """World-size-2 gloo run of the range plan with small segments (cuts inside targets): rows gathered in part order and
    finalized on rank 0 equal the single-process result bit for bit."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_binding as ob
    from conftest import make_prep
    from portcullis_b200 import junction_builder as jb
    with tempfile.TemporaryDirectory() as d:
        import subprocess
        prep = os.path.join(d, "prep")
        subprocess.check_call([os.path.join(ROOT, "portcullis_b200", "bin", "pjsynth"), "--preset", "c2", "--scale", "0.01", "--threads", "2", "--out", prep],
                              stderr=subprocess.DEVNULL)
        out = os.path.join(d, "rows.npy")
        mp.spawn(_worker_ranges, args=(2, _free_port(), prep, out, 6000), nprocs=2, join=True)
        got = np.load(out)
        p = jb.PrepDir(prep)
        seg, cuts = p.plan(2, 6000)
        assert seg.min() >= 2 and cuts > 0
        cols = p.decode(-1, 2)
        rows, st = ob.run(cols, p.lengths, [p.genome(t) for t in range(len(p.names))])
        exp = jb.finalize(rows.copy(), float(st["sumq"].sum()) / float(st["spliced"].sum() + st["unspliced"].sum()))
        assert got.tobytes() == exp.tobytes()
