"""World-size-2 NCCL test of the multi-GPU path on real hardware (needs two GPUs: `gpurun --gpus 2`): every rank
holds half of ONE global read set, reads are bucketed per gene-range shard on the device and exchanged over NCCL
(lsq_shard_reads), each rank counts and solves its own events, rank 0 gathers (lsq_gather).  The gathered tables
must equal the single-GPU run of the whole selection over the whole read set byte for byte."""
import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, tmp, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import ctypes as C
    import torch
    import torch.distributed as dist
    from lesseq_b200 import api, multi, synth
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ok = True
    try:
        for case in sorted(os.listdir(tmp)):
            d = os.path.join(tmp, case)
            fp, mp_, rp = os.path.join(d, "ev.interval"), os.path.join(d, "ev.map"), os.path.join(d, "r.mrf")
            meta = json.load(open(os.path.join(d, "meta.json")))
            R, rtype, trb = meta["R"], meta["type"], [meta["trb"]]
            job = multi.ShardedJob(fp, mp_, rank, rank, world, begin=meta["begin"], end=meta["end"])
            # the whole read set, parsed on every rank; this rank keeps its contiguous slice of the reads
            whole = api.Reads.from_mrf(rp)
            v = whole.view
            n = int(v.n_reads)
            lo, hi = (n * rank) // world, (n * (rank + 1)) // world
            def arr(ptr, ctype, count):
                return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(ctype)), shape=(count,)).copy()
            off = arr(v.blk_off, C.c_uint32, n + 1)
            blk = arr(v.blk, C.c_int32, 2 * int(v.n_blocks))
            cs = arr(v.read_cs, C.c_uint32, n)
            ln = arr(v.line_num, C.c_uint32, n) if v.line_num else np.arange(1, n + 1, dtype=np.uint32)
            bcs = arr(v.blk_cs, C.c_uint32, int(v.n_blocks)) if v.blk_cs else None
            cn = [v.chrom_names[i].decode() for i in range(v.n_chrom_names)]
            sn = [v.strand_names[i].decode() for i in range(v.n_strand_names)]
            sl = api.Reads.from_numpy(off[lo:hi + 1] - off[lo], blk[2 * off[lo]:2 * off[hi]], cs[lo:hi], ln[lo:hi], cn, sn,
                                      blk_cs=None if bcs is None else bcs[off[lo]:off[hi]])
            cnt, sol = job.step(sl, rtype, R, 1, trb)
            st = job.stats
            assert st["reads_local"] == hi - lo
            if rank == 0:
                # single-GPU run of the whole selection over the whole read set
                ctx = api.Context(0)
                ctx.upload_events(job.full_ann)
                ctx.set_reads(0, rtype, R, whole)
                ctx.count()
                ctx.solve(trb)
                c1, s1 = ctx.get_count(), ctx.get_solve()
                ctx.close()
                for k in c1:
                    if not np.array_equal(c1[k], cnt[k]):
                        ok = False
                        print("MISMATCH", case, k, flush=True)
                for k in ("theta", "rpkm", "loglik", "ll_per_read", "iters"):
                    if not np.array_equal(s1[k], sol[k], equal_nan=(s1[k].dtype.kind == "f")):
                        ok = False
                        print("MISMATCH", case, k, flush=True)
            job.close()
    finally:
        if rank == 0:
            q.put(ok)
        dist.barrier()
        dist.destroy_process_group()


def test_two_gpu_bucketing_exchange_gather_equals_single_gpu(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (gpurun --gpus 2)")
    import torch.multiprocessing as mp
    import fuzzgen
    from lesseq_b200 import build, synth
    build.build()
    # a C2-shaped case large enough for many groups per rank, a junction-heavy one, and adversarial fuzz cases
    # (nested / overlapping events, chimeric reads, names that trigger the read-name tie-break)
    for name, cfg, genes, reads, kw in (("a_c2", "C2", 3000, 600000, {}), ("b_c4", "C4", 1500, 300000, {"overhang_frac": 0.02})):
        A, R, _ = synth.make_config(cfg, n_genes=genes)
        d = tmp_path / name
        d.mkdir()
        A.write(str(d / "ev.interval"), str(d / "ev.map"))
        rd = synth.make_reads(A, reads, R, 21, device="cpu", **kw)
        synth.reads_to_mrf(rd, str(d / "r.mrf"))
        (d / "meta.json").write_text(json.dumps({"R": R, "type": "SHORT_READ", "trb": rd["total_read_bases"], "begin": 0,
                                                 "end": 2 ** 62}))
    for seed in range(7400, 7412):
        case = fuzzgen.make_case(seed)
        d = tmp_path / ("f%d" % seed)
        d.mkdir()
        ca, sa = fuzzgen.write_case(case, str(d))
        os.replace(ca[12], str(d / "r.mrf"))  # the first read set of the case
        (d / "meta.json").write_text(json.dumps({"R": int(ca[11]), "type": ca[10], "trb": float(sa[13]), "begin": int(ca[7]),
                                                 "end": int(ca[8])}))
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 32500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, str(tmp_path), q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(420)
        if p.is_alive():
            p.kill()
        assert p.exitcode == 0
    assert q.get(timeout=5) is True
