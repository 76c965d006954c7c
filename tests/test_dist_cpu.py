"""World-size-2 gloo test of the gene-range sharding plumbing (no GPU): shard ranges partition
the selection, both shards see the covered regions of the whole range, and the per-rank result
rows gather on rank 0 in gene-rank order."""
import os
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, tmp, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from lesseq_b200 import api, dist as lsd
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    fp, mp_ = os.path.join(tmp, "ev.interval"), os.path.join(tmp, "ev.map")
    full = api.Annotation(fp, mp_)
    sb, se = lsd.shard_range(full.n_events, rank, world)
    ann = api.Annotation(fp, mp_, shard=(sb, se))
    assert ann.n_events == se - sb
    assert ann.event_names() == full.event_names()[sb:se]
    ncov = int(full.array("cov_off", full.tables.n_chroms + 1)[-1])
    assert np.array_equal(ann.array("cov_start", ncov), full.array("cov_start", ncov))
    assert np.array_equal(ann.array("cov_end", ncov), full.array("cov_end", ncov))
    # fake per-shard results whose values encode the global event / form index
    foff = full.form_offsets()
    ev = np.arange(sb, se, dtype=np.uint64)
    forms = np.arange(int(foff[sb]), int(foff[se]), dtype=np.float64)
    count = {"n_valid": np.stack([ev, ev + 1000]), "bases": np.stack([ev * 2, ev * 3]),
             "iso_count": forms.astype(np.uint64)}
    solve = {"theta": forms / 7.0, "iters": ev.astype(np.uint32)}
    out = lsd.gather_results(count, solve)
    if rank == 0:
        E, F = full.n_events, full.n_forms
        ok = (np.array_equal(out["n_valid"], np.stack([np.arange(E), np.arange(E) + 1000]).astype(np.uint64)) and
              np.array_equal(out["iso_count"], np.arange(F, dtype=np.uint64)) and
              np.array_equal(out["theta"], np.arange(F) / 7.0) and
              np.array_equal(out["iters"], np.arange(E, dtype=np.uint32)) and out["n_valid"].dtype == np.uint64)
        q.put(bool(ok))
    else:
        assert all(v is None for v in out.values())
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_and_gather(tmp_path):
    sys.path.insert(0, ROOT)
    from lesseq_b200 import build, synth
    build.build()
    A = synth.make_annotation(40, 5, n_chroms=3, types=synth.EVENT_TYPES, exons_per_gene=(4, 8), events_per_gene=(1, 3))
    A.write(str(tmp_path / "ev.interval"), str(tmp_path / "ev.map"))
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, str(tmp_path), q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0
    assert q.get(timeout=5) is True


def test_more_ranks_than_events_gather(tmp_path):
    """A rank whose shard is empty (world size above the number of selected genes) must still take part in the
    gather: three ranks, two events."""
    sys.path.insert(0, ROOT)
    from lesseq_b200 import build, synth
    build.build()
    A = synth.make_annotation(2, 6, n_chroms=1, types=("SE",), exons_per_gene=(3, 3), events_per_gene=(1, 1))
    A.write(str(tmp_path / "ev.interval"), str(tmp_path / "ev.map"))
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 3, port, str(tmp_path), q)) for r in range(3)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0
    assert q.get(timeout=5) is True


def test_shard_ranges_partition():
    sys.path.insert(0, ROOT)
    from lesseq_b200.dist import shard_range
    for n in (0, 1, 7, 69930):
        for world in (1, 2, 3, 8):
            r = [shard_range(n, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[i][1] == r[i + 1][0] for i in range(world - 1))
