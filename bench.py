#!/usr/bin/env python
"""bench.py -- headline benchmark of the count + solve hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--workload C2|C1|C4] [--reads R] [--genes G]

A "step" is one full pass of the device pipeline over one batch of synthetic packed reads:
scan (load filter + candidate lookup + signature tally) -> expand -> general -> finalize -> EM solve.
  value : reads/s of that pass, inputs already resident in HBM, CUDA-event timed, max over ranks
  e2e   : the same through the C ABI with HOST (pinned) buffers: lsq_set_reads (H2D) + lsq_count +
          lsq_solve + lsq_get_count/lsq_get_solve (D2H) inside the timed region
  roofline : the dominant kernel (lsq_scan_kernel): SURVEY 8(d) algorithmic bytes N*(6+8*blocks/read) / its
             CUDA-event duration vs the measured HBM copy peak
  cpu_baseline : the reference's own count + solve binaries (oracle/_ref) on a bounded text sample
`--impl reference` times the reference binaries on all host cores (its own scale-out recipe:
one process per disjoint gene range, each reading the whole read file).
Multi-GPU (BASELINE config 3): one process per GPU (torchrun); ONE read set of fixed size (1 B reads) over all ranks,
events sharded by gene rank; every step buckets each rank's slice per owning rank on the device, exchanges the buckets
over NCCL, counts + solves the own events and gathers the result rows on rank 0 -- all inside the timed region;
scaling is strong.
"""
import argparse
import json
import os
import shutil
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "reads_per_sec_counted_and_solved"
UNIT = "reads/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="C2", choices=["C1", "C2", "C4", "C5"],
                    help="C5 = solve stress: 500k events with 2-16 forms, class tables injected (solve only)")
    ap.add_argument("--events", type=int, default=500000, help="events of the C5 workload")
    ap.add_argument("--reads", type=int, default=None, help="reads per GPU (default: the workload's)")
    ap.add_argument("--reads-total", type=int, default=None,
                    help="N > 1: size of the ONE read set shared by all ranks (default 1e9: BASELINE config 3)")
    ap.add_argument("--genes", type=int, default=None)
    ap.add_argument("--sample-reads", type=int, default=300000, help="reads of the CPU reference sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cli-reads", type=int, default=10_000_000, help="reads of the MRF text the drop-in executables are timed on (0: skip)")
    ap.add_argument("--wide-wire", action="store_true", help="e2e: wide arrays (blk, read_cs) instead of the compact wire format")
    ap.add_argument("--no-stream", action="store_true", help="e2e: copy everything before the filter starts (no LSQ_READS_STREAM)")
    return ap.parse_args()


def workload_config(args, world):
    from lesseq_b200 import synth
    kw, R, n = synth.CONFIGS[args.workload]
    n = args.reads or n
    genes = args.genes or kw["n_genes"]
    desc = {"C1": "C1: single chromosome, SE events only, 50bp MRF_SINGLE reads",
            "C2": "C2: 24 chromosomes, all 8 event types (SE/RI/A5SS/A3SS/MXE/AFE/ALE/T3), 75bp reads",
            "C4": "C4: junction-heavy 100bp reads over short exons, 2% junction overhangs"}[args.workload]
    cfg = {"workload": desc, "genes": genes, "reads_per_gpu": n, "read_length": R, "read_type": "SHORT_READ",
           "sharding": "gene-rank ranges x%d, covered regions of the whole range" % world,
           }
    return cfg, R, n, genes


def multi_gpu_config(cfg, args, world, n):
    """BASELINE config 3 for world > 1: ONE read set of fixed size (1 B reads) over all ranks; rank r holds reads
    [r * total / world, (r + 1) * total / world) of it, whatever events they hit.  Returns (config, reads per GPU,
    total reads); both arms of the bench print this config."""
    if world <= 1:
        return cfg, n, n
    total_reads = args.reads_total or (args.reads * world if args.reads else 1_000_000_000)
    n = total_reads // world
    total_reads = n * world
    cfg["workload"] = cfg["workload"].replace("C2:", "C3 (C2 genome, one read set of %d reads over %d GPUs):" % (total_reads, world))
    cfg["reads_total"] = total_reads
    cfg["reads_per_gpu"] = n
    cfg["sharding"] = ("gene-rank ranges x%d (events), covered regions of the whole range; every rank holds 1/%d of the read "
                       "set, buckets it per owning rank on the device and exchanges the buckets over NCCL" % (world, world))
    return cfg, n, total_reads


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows = []
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm = sorted(int(float(r[0])) for r in self.rows if r and r[0].replace(".", "").isdigit())
        mx = [int(float(r[1])) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 7 for i in range(4) if r[3 + i].startswith("Active")})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def measured_traffic(args, world, n, which):
    """dram__bytes_read+write of the filter kernel / the count kernels from the committed ncu capture
    of this very workload (profiles/count_traffic.json); None for any other configuration."""
    p = os.path.join(ROOT, "profiles", "count_traffic.json")
    if not os.path.exists(p):
        return None
    t = json.load(open(p))
    if t.get("workload") == args.workload and t.get("reads_per_gpu") == n and world == 1 and not args.genes:
        return t.get("dram_bytes_per_step_%s" % which)
    return None


def write_sample(args, tmp):
    """Annotation text + a bounded MRF text sample of the same workload for the CPU reference."""
    from lesseq_b200 import synth
    A, R, _ = synth.make_config(args.workload, n_genes=args.genes)
    fp, mp, rp = os.path.join(tmp, "ev.interval"), os.path.join(tmp, "ev.map"), os.path.join(tmp, "sample.mrf")
    A.write(fp, mp)
    reads = synth.make_reads(A, args.sample_reads, R, 1234, device="cpu")
    synth.reads_to_mrf(reads, rp)
    return A, R, fp, mp, rp, reads["total_read_bases"]


def ref_args(fp, mp, b, e, R, rp, trb=None):
    a = ["0", "bench", "o_", "LH_GENE_TXT", fp, "UCSC_GENE2ISOFORM", mp, str(b), str(e), "MRF_SINGLE", "SHORT_READ",
         str(R), rp]
    return a + ([repr(trb)] if trb is not None else [])


def cpu_reference_once(fp, mp, R, rp, trb, n_events, procs):
    """count then solve with `procs` concurrent reference processes over disjoint gene ranges.
    Returns wall seconds and the kind of CPU implementation used."""
    from oracle import runner
    if runner.ref_available() or runner.build_ref():
        def launch(prog, args):
            env = dict(os.environ)
            env["LD_LIBRARY_PATH"] = os.path.join(runner.REF_DIR, "lib")
            return subprocess.Popen([os.path.join(runner.REF_DIR, prog)] + args, env=env,
                                    stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        kind = "reference"
    else:
        runner.build_oracle()

        def launch(prog, args):
            return subprocess.Popen([runner.ORACLE_BIN, prog] + args, stdout=subprocess.DEVNULL,
                                    stderr=subprocess.DEVNULL)
        kind = "port"
    t0 = time.time()
    for prog in ("count", "solve"):
        ps = []
        for p in range(procs):
            b, e = (n_events * p) // procs, (n_events * (p + 1)) // procs
            ps.append(launch(prog, ref_args(fp, mp, b, e, R, rp, trb if prog == "solve" else None)))
        for p in ps:
            if p.wait() != 0:
                raise RuntimeError("CPU reference process failed")
    return time.time() - t0, kind


def cli_e2e(args, tmp, torch):
    """Wall clock of the drop-in executables (lesseq_b200/bin/count + solve: text in, table out) on an MRF_SINGLE text of
    `--cli-reads` reads of the workload, next to the reference's own binaries on the first `--sample-reads` lines of the
    SAME text (the reference needs ~20 s per million reads; its rate is what is compared)."""
    from lesseq_b200 import api, synth
    from oracle import runner
    A, R, _ = synth.make_config(args.workload, n_genes=args.genes)
    fp, mp, rp, sp = (os.path.join(tmp, x) for x in ("cli.interval", "cli.map", "cli.mrf", "cli_sample.mrf"))
    A.write(fp, mp)
    n = args.cli_reads
    reads = synth.make_reads(A, n, R, 4321, device="cuda")
    host = {k: reads[k].cpu().numpy() for k in ("blk_off", "blk", "read_cs", "line_num")}
    trb = reads["total_read_bases"]
    hr = api.Reads.from_numpy(host["blk_off"], host["blk"], host["read_cs"], host["line_num"], reads["chrom_names"], reads["strand_names"])
    t0 = time.time()
    hr.write_mrf(rp)
    t_write = time.time() - t0
    with open(rp) as f, open(sp, "w") as g:  # the bounded sample: header + the first sample-reads lines
        for i, line in enumerate(f):
            if i > args.sample_reads:
                break
            g.write(line)
    size = os.path.getsize(rp)
    head = ["0", "bench", os.path.join(tmp, "o_"), "LH_GENE_TXT", fp, "UCSC_GENE2ISOFORM", mp, "0", str(10 ** 9), "MRF_SINGLE", "SHORT_READ", str(R)]
    walls, rows, marks = {}, {}, {}
    for prog in ("count", "solve"):
        a = head + [rp] + ([repr(trb)] if prog == "solve" else [])
        best = None
        for _ in range(2):  # first run pays the page-cache / driver warm-up; the second is reported
            t0 = time.time()
            rc, out, err = api.run_program(prog, a, env_extra={"LSQ_CLI_TIMING": "1"})
            dt = time.time() - t0
            if rc != 0:
                raise RuntimeError("bin/%s failed: %s" % (prog, err[-300:]))
            best = dt
            marks[prog] = {}
            for ln in err.splitlines():  # "[timing] <milestone> <ms since start>"
                f = ln.split()
                if len(f) >= 3 and f[0] == "[timing]":
                    try:
                        marks[prog]["_".join(f[1:-1])] = float(f[-1])
                    except ValueError:
                        pass
        walls[prog] = best
        rows[prog] = out.count("\n")
    ref = {}
    if runner.ref_available() or runner.build_ref():
        for prog in ("count", "solve"):
            a = head + [sp] + ([repr(trb * args.sample_reads / n)] if prog == "solve" else [])
            t0 = time.time()
            rc, out, err = runner.run_ref(prog, a, prec17=False)
            ref[prog] = time.time() - t0
            if rc != 0:
                raise RuntimeError("reference %s failed" % prog)
    ours_rate = n / (walls["count"] + walls["solve"])
    res = {"reads": n, "text_bytes": size, "wall_s": {k: round(v, 3) for k, v in walls.items()}, "reads_per_s": ours_rate,
           "text_GB_per_s_count": size / walls["count"] / 1e9, "milestones_ms": marks, "host_cores": os.cpu_count(), "table_rows": rows["count"],
           "note": "bin/count then bin/solve, each: parse the text on all host cores (lsqh_mrf_parse) while a helper thread creates the CUDA "
                   "context and uploads the event tables, compact host->device copy, scan + expand + general (+ solve), print the table; "
                   "second of two runs (warm page cache); writing the text took %.1f s (not counted)" % t_write}
    if ref:
        ref_rate = args.sample_reads / (ref["count"] + ref["solve"])
        res["reference"] = {"reads": args.sample_reads, "wall_s": {k: round(v, 3) for k, v in ref.items()}, "reads_per_s": ref_rate,
                            "note": "the reference's own count + solve binaries (one thread each, as shipped) on the first %d lines of the same text" % args.sample_reads}
        res["speedup_reads_per_s"] = ours_rate / ref_rate
    del reads
    torch.cuda.empty_cache()
    return res


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    world = int(os.environ.get("WORLD_SIZE", str(args.gpus)))
    cfg, R, n, genes = workload_config(args, world)
    cfg, n, _ = multi_gpu_config(cfg, args, world, n)  # the same config line as the GPU arm at this N
    tmp = tempfile.mkdtemp(prefix="lsq_ref_")
    try:
        A, R, fp, mp, rp, trb = write_sample(args, tmp)
        procs = os.cpu_count() or 1
        n_events = len(A.ev_gene)
        times = []
        kind = "reference"
        for i in range(args.warmup + args.steps):
            dt, kind = cpu_reference_once(fp, mp, R, rp, trb, n_events, procs)
            if i >= args.warmup:
                times.append(dt)
        ms = 1000.0 * sum(times) / len(times)
        value = args.sample_reads / (ms / 1000.0)
        sample = ("%d-read MRF text sample of the workload + the full %d-event annotation per step; count then solve, "
                  "%d concurrent reference processes over disjoint gene ranges, each reading the whole file"
                  % (args.sample_reads, n_events, procs))
        line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
                "scaling": "strong" if world > 1 else "weak", "vs_baseline": None, "dtype": "i64+f64", "data": "synthetic", "config": cfg,
                "cpu_baseline": {"value": value, "unit": UNIT, "cores": procs, "kind": kind, "sample": sample},
                "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "events_per_sec": n_events / (ms / 1000.0), "gpu_launches": 0}
        print(json.dumps(line), flush=True)
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


def compact_reads(reads, torch, wide):
    """The compact packed form of a synthetic read set (include/lesseq_b200.h): blocks per read u8, chromosome/strand
    id u8, block start i32 + block length u8 (u16 when a block is 256 or longer).  None when a value does not fit."""
    blk_cnt = (reads["blk_off"][1:] - reads["blk_off"][:-1]).to(torch.uint8)
    cs = reads["read_cs"]
    b2 = reads["blk"].view(-1, 2)
    blen = b2[:, 1] - b2[:, 0]
    ok = len(reads["chrom_names"]) <= 64 and len(reads["strand_names"]) <= 2 and \
        int(blen.max()) < 65536 and int(blen.min()) >= 0 and not wide
    if not ok:
        return None, False
    cs8 = ((cs & 0x3F) | (((cs >> 16) & 1) << 6) | (((cs >> 31) & 1) << 7)).to(torch.uint8)
    len8 = int(blen.max()) < 256  # short reads: one byte per block length
    return {"blk_start": b2[:, 0].contiguous(), "blk_len": blen.to(torch.uint8 if len8 else torch.int16).contiguous(),
            "read_cs8": cs8.contiguous(), "blk_cnt": blk_cnt.contiguous()}, len8


def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    from lesseq_b200 import build
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- lesseq_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if rank == 0:
        build.build()
    if world > 1:
        dist.barrier()
    from lesseq_b200 import api, multi, synth

    cfg, R, n, genes = workload_config(args, world)
    cfg, n, total_reads = multi_gpu_config(cfg, args, world, n)
    tmp = tempfile.mkdtemp(prefix="lsq_bench_")
    A, _, _ = synth.make_config(args.workload, n_genes=args.genes)
    fp, mp = os.path.join(tmp, "ev.interval"), os.path.join(tmp, "ev.map")
    if rank == 0:
        A.write(fp, mp)
    if world > 1:
        box = [tmp]
        dist.broadcast_object_list(box, src=0)
        fp, mp = os.path.join(box[0], "ev.interval"), os.path.join(box[0], "ev.map")
        dist.barrier()
    n_events_total = len(A.ev_gene)
    stream = torch.cuda.current_stream()
    reads = synth.make_reads(A, n, R, 100 + rank, device="cuda")
    compact, len8 = compact_reads(reads, torch, args.wide_wire)
    compact_ok = compact is not None
    if compact_ok:
        dev_reads = api.Reads.from_torch(None, None, None, None, reads["chrom_names"], reads["strand_names"],
                                         blk_cnt=compact["blk_cnt"], read_cs8=compact["read_cs8"], blk_start=compact["blk_start"],
                                         **{"blk_len8" if len8 else "blk_len": compact["blk_len"]})
        pinned = {k: v.cpu().pin_memory() for k, v in compact.items()}
        host_reads = api.Reads.from_torch(None, None, None, None, reads["chrom_names"], reads["strand_names"],
                                          blk_cnt=pinned["blk_cnt"], stream=not args.no_stream and world == 1,
                                          read_cs8=pinned["read_cs8"], blk_start=pinned["blk_start"],
                                          **{"blk_len8" if len8 else "blk_len": pinned["blk_len"]})
    else:
        dev_reads = api.Reads.from_torch(reads["blk_off"], reads["blk"], reads["read_cs"], reads["line_num"],
                                         reads["chrom_names"], reads["strand_names"])
        blk_cnt = (reads["blk_off"][1:] - reads["blk_off"][:-1]).to(torch.uint8)
        pinned = {"blk": reads["blk"].cpu().pin_memory(), "read_cs": reads["read_cs"].cpu().pin_memory(),
                  "blk_cnt": blk_cnt.cpu().pin_memory()}
        host_reads = api.Reads.from_torch(None, pinned["blk"], pinned["read_cs"], None, reads["chrom_names"],
                                          reads["strand_names"], blk_cnt=pinned["blk_cnt"], stream=not args.no_stream and world == 1)
    resident_bytes = int(dev_reads.host_bytes())
    n_blocks = int(reads["n_blocks"])
    cfg["l2_policy"] = "inputs (%.2f GB packed reads per GPU) far exceed the 126 MB L2; no flush needed" % (resident_bytes / 1e9)
    trb_local = torch.tensor([reads["total_read_bases"]], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(trb_local)
    trb = [float(trb_local[0])]
    del reads
    torch.cuda.empty_cache()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    phase_keys = ("clear_ms", "scan_ms", "expand_ms", "general_ms", "finalize_ms", "solve_ms")
    if world == 1:
        ann = api.Annotation(fp, mp)
        ctx = api.Context(local, stream=stream.cuda_stream)
        ctx.upload_events(ann)
        ctx.set_reads(0, "SHORT_READ", R, dev_reads)

        def step_resident():
            ctx.count()
            ctx.solve(trb)

        def step_e2e():
            ctx.set_reads(0, "SHORT_READ", R, host_reads)
            ctx.count()
            ctx.solve(trb)
            return ctx.get_count(), ctx.get_solve()

        def results():
            return ctx.get_count(), ctx.get_solve()
        tctx = ctx
        shard_stats = lambda: None  # noqa: E731
    else:
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)  # NCCL prints its version banner on stdout when a communicator is created: keep the ONE JSON line clean
        try:
            job = multi.ShardedJob(fp, mp, local, rank, world, stream=stream.cuda_stream)
        finally:
            os.dup2(saved_stdout, 1)
            os.close(saved_stdout)
        first_line = 1 + rank * n
        job.full.set_reads(0, "SHORT_READ", R, dev_reads)
        last = {}

        def sharded_pass():
            # bucket this rank's slice per owning rank + exchange over NCCL, count and solve the own events, gather
            last["stats"] = api.shard_reads(job.full, 0, job.shard, 0, job.comm, first_line)
            job.shard.count()
            job.shard.solve(trb)
            last["out"] = api.gather(job.full, job.shard, job.comm, 0)

        def step_resident():
            sharded_pass()

        def step_e2e():
            job.full.set_reads(0, "SHORT_READ", R, host_reads)  # host -> device copy of this rank's slice
            sharded_pass()
            return last["out"]

        def results():
            return last["out"]
        tctx = job.shard
        shard_stats = lambda: last["stats"]  # noqa: E731

    # ---- device-resident pass ---------------------------------------------------------------
    for _ in range(args.warmup):
        step_resident()
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    phases = {k: 0.0 for k in phase_keys}
    route = {"route_ms": 0.0, "exchange_ms": 0.0, "reads_received": 0, "reads_sent": 0}
    launches = 0
    ev0.record(stream)
    for _ in range(args.steps):
        step_resident()
        t = tctx.timings()  # waits for the step: per-kernel CUDA-event times of this very step
        for k in phases:
            phases[k] += t[k]
        launches += t["launches"] + (1 if world > 1 else 0)
        st = shard_stats()
        if st:
            for k in route:
                route[k] += st[k]
    ev1.record(stream)
    barrier()
    ms_local = ev0.elapsed_time(ev1) / args.steps
    tsum = tctx.timings()
    cnt, sol = results()

    # ---- end to end through the C ABI with host buffers ---------------------------------------
    for _ in range(min(args.warmup, 2)):
        step_e2e()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        c2, s2 = step_e2e()
    e1.record(stream)
    barrier()
    e2e_ms_local = e0.elapsed_time(e1) / args.steps
    clocks = sampler.stop() if sampler else None
    if rank == 0:
        assert np.array_equal(c2["iso_count"], cnt["iso_count"]) and np.array_equal(s2["theta"], sol["theta"])
        assert sol["iters"].shape[0] == n_events_total

    red = torch.tensor([ms_local, e2e_ms_local], dtype=torch.float64, device="cuda")
    per_rank = torch.zeros(world, 4, dtype=torch.float64, device="cuda")
    per_rank[rank] = torch.tensor([route["route_ms"] / args.steps, route["exchange_ms"] / args.steps,
                                   route["reads_received"] / args.steps, phases["scan_ms"] / args.steps], dtype=torch.float64)
    if world > 1:
        dist.all_reduce(red, op=dist.ReduceOp.MAX)
        dist.all_reduce(per_rank)
    ms, e2e_ms = float(red[0]), float(red[1])
    if rank == 0:
        peak, peak_src = hbm_peak()
        sk_ms = phases["scan_ms"] / args.steps
        # algorithmic bytes of the count path = one pass over the packed reads, N * (6 + 8 * blocks/read)
        # (2-byte chromosome id + 4-byte offset per read, 8 bytes per block); the resident layout is narrower
        # (`resident_bytes`), so the kernel moves fewer bytes than that
        n_scanned = int(tsum["reads_in"])
        alg_bytes = int(6 * n_scanned + 8 * int(tsum["blocks_in"]))
        achieved = alg_bytes / (sk_ms * 1e-3) / 1e9 if sk_ms > 0 else 0.0
        d2h = int(sum(v.nbytes for v in cnt.values()) + sum(v.nbytes for v in sol.values()))
        line = {
            "metric": METRIC, "value": total_reads / (ms / 1000.0), "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "strong" if world > 1 else "weak", "vs_baseline": None, "dtype": "i32+f64", "data": "synthetic", "config": cfg,
            "clocks": clocks,
            "e2e": {"value": total_reads / (e2e_ms / 1000.0), "unit": UNIT,
                    "h2d_bytes_per_step": int(host_reads.host_bytes()) * world, "d2h_bytes_per_step": d2h,
                    "ms_per_step": e2e_ms,
                    "wire": ("compact (blk_cnt u8, read_cs8 u8, blk_start i32 + blk_len u8/u16)" if compact_ok else "blk_cnt u8 + wide blk/read_cs") +
                            (", copy overlapped with the scan (LSQ_READS_STREAM)" if (not args.no_stream and world == 1) else "")},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "kernel": "lsq_scan_kernel (1 launch per step, %.0f%% of the step)" % (100.0 * sk_ms / ms),
                         "achieved": achieved, "peak": peak,
                         "unit": "GB/s", "frac": achieved / peak, "traffic": measured_traffic(args, world, n, "scan_kernel"),
                         "peak_source": peak_src, "bytes_per_launch": alg_bytes, "ms_per_launch": sk_ms,
                         "bytes_note": "SURVEY 8(d): N*(6+8*blocks/read) over the reads the kernel scanned; the resident compact "
                                       "layout of this rank's slice holds %d bytes" % resident_bytes,
                         "limiter": "instruction issue and scattered L1TEX operations (2 table gathers + 1 reduction per read; "
                                    "floor 1.1 ms per 100 M reads, profiles/r02_microbench_gather_red_fp64.log), not HBM"},
            # every kernel of the step (CUDA-event times, rank 0)
            "kernels": [{"name": "lsq_" + k[:-3] + "_kernel", "ms": phases[k] / args.steps} for k in phase_keys],
            "events_per_sec": n_events_total / (ms / 1000.0),
            "events": n_events_total,
            "phases_ms_per_step_rank0": {k: v / args.steps for k, v in phases.items()},
            "reads_rank0": {k: int(tsum[k]) for k in ("reads_in", "blocks_in", "reads_single", "reads_pair", "reads_triple", "reads_general", "hash_full")},
            "valid_pairs_rank0": int(tsum["pairs_valid"]), "em_exact_replays_rank0": int(tsum["exact_events"]),
            "em_exact_replay_ms_rank0": float(tsum["exact_ms"]),  # part of solve_ms, last step
            "em_iters_mean": float(sol["iters"].mean()), "em_iters_max": int(sol["iters"].max()),
            "em_iters_p99": float(np.percentile(sol["iters"], 99)),
        }
        if world > 1:
            pr = per_rank.cpu().numpy()
            line["bucketing"] = {
                "route_ms_per_rank": [round(float(x), 3) for x in pr[:, 0]],
                "exchange_ms_per_rank": [round(float(x), 3) for x in pr[:, 1]],
                "scan_ms_per_rank": [round(float(x), 3) for x in pr[:, 3]],
                "reads_received_per_rank": [int(x) for x in pr[:, 2]],
                "multiplicity": float(pr[:, 2].sum()) / total_reads,
                "note": "a read is sent to every rank owning a candidate event of any of its blocks' pieces; gene-rank order "
                        "is lexicographic in the event id, so the events of one gene live on several ranks"}
        if world == 1 and not args.no_cpu_baseline:
            try:
                A2, R2, fp2, mp2, rp2, trb2 = write_sample(args, tmp)
                dt, kind = cpu_reference_once(fp2, mp2, R2, rp2, trb2, len(A2.ev_gene), 1)
                line["cpu_baseline"] = {
                    "value": args.sample_reads / dt, "unit": UNIT, "cores": 1, "kind": kind,
                    "sample": "%d-read MRF text sample of the workload + the full %d-event annotation: reference "
                              "count then solve, one single-threaded process each, %.1f s wall"
                              % (args.sample_reads, len(A2.ev_gene), dt)}
            except Exception as ex:  # the oracle always exists; report why it could not run
                line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 0, "kind": "port", "sample": "failed: %s" % ex}
            if args.cli_reads and args.workload != "C5":
                try:
                    ctx.close()
                    line["cli_e2e"] = cli_e2e(args, tmp, torch)
                except Exception as ex:
                    line["cli_e2e"] = {"failed": str(ex)[:300]}
        print(json.dumps(line), flush=True)
    if world == 1:
        ctx.close()
    else:
        dist.barrier()
        job.close()
    if rank == 0:
        shutil.rmtree(tmp, ignore_errors=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def run_solve_stress(args):
    """BASELINE config 5: batched EM (+ standard errors) over injected compatibility-class tables.
    value = events/s of lsq_solve, device resident; e2e adds lsq_set_classes (H2D) and lsq_get_solve."""
    import numpy as np
    import torch
    from lesseq_b200 import build
    build.build()
    from lesseq_b200 import api, synth
    torch.cuda.set_device(0)
    E = args.events
    S = synth.make_solve_stress_fast(E, 5)
    nch = S["n_chroms"]
    cov_off = np.arange(nch + 1)
    ann = api.ArrayAnnotation(
        ev_chrom=S["ev_chrom"], ev_strand=np.zeros(E), ev_gs=S["ev_gs"], ev_ge=S["ev_ge"], ev_seg_off=S["seg_off"],
        ev_form_off=S["form_off"], seg_start=S["seg_start"], seg_end=S["seg_end"], form_seg_off=S["form_seg_off"],
        form_seg_idx=S["form_seg_idx"], cov_off=cov_off, cov_start=np.zeros(nch), cov_end=np.full(nch, 2 ** 31 - 2),
        chrom_names=["chr%d" % (i + 1) for i in range(nch)], strand_names=["+"])
    stream = torch.cuda.current_stream()
    ctx = api.Context(0, stream=stream.cuda_stream)
    ctx.upload_events(ann)
    ctx.set_classes(0, "SHORT_READ", 75, S["cls_off"], S["cls_mask"], S["cls_cnt"], S["bases"])
    trb = [float(S["bases"].sum())]
    for _ in range(args.warmup):
        ctx.solve(trb)
    torch.cuda.synchronize()
    sampler = ClockSampler(0)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    solve_ms = 0.0
    for _ in range(args.steps):
        ctx.solve(trb)
        solve_ms += ctx.timings()["solve_ms"]
    e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    sol = ctx.get_solve()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record(stream)
    for _ in range(args.steps):
        ctx.set_classes(0, "SHORT_READ", 75, S["cls_off"], S["cls_mask"], S["cls_cnt"], S["bases"])
        ctx.solve(trb)
        sol2 = ctx.get_solve()
    t1.record(stream)
    torch.cuda.synchronize()
    e2e_ms = t0.elapsed_time(t1) / args.steps
    clocks = sampler.stop()
    assert np.array_equal(sol["theta"], sol2["theta"])
    C = np.diff(S["cls_off"].astype(np.int64))
    K = S["K"]
    it = sol["iters"].astype(np.float64)
    flops = float((it * C * (7 * K + 2)).sum() + (4 * K).sum())       # C(7K+2) per EM iteration
    bytes_alg = float((16 * C + 24 * K + 16).sum())                    # class table + G/len/theta/rpkm per event
    fp64_peak = ctx.fp64_peak_tflops()  # DFMA micro-benchmark on this very device
    kms = solve_ms / args.steps
    gflops = flops / (kms * 1e-3) / 1e9
    line = {
        "metric": "events_per_sec_solved", "value": E / (ms / 1000.0), "unit": "events/s", "n_gpus": 1,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "C5: solve stress, %d events with 2-16 forms, K..3K compatibility classes each, "
                               "batched FP64 EM + observed-information standard errors" % E,
                   "l2_policy": "class tables (%.0f MB) are streamed once per EM iteration from L2; solve is FP64-latency bound" % (bytes_alg / 1e6)},
        "clocks": clocks,
        "e2e": {"value": E / (e2e_ms / 1000.0), "unit": "events/s",
                "h2d_bytes_per_step": int(S["cls_off"].nbytes + S["cls_mask"].nbytes + S["cls_cnt"].nbytes + S["bases"].nbytes),
                "d2h_bytes_per_step": int(sum(v.nbytes for v in sol.values())), "ms_per_step": e2e_ms},
        "gpu_launches": int(ctx.timings()["launches"]) * args.steps,
        "roofline": {"bound": "fp64", "kernel": "lsq_solve_mid_kernel + lsq_solve_kernel<8> (standard errors)",
                     "achieved": gflops / 1e3, "peak": fp64_peak, "unit": "TFLOP/s", "frac": gflops / 1e3 / fp64_peak, "traffic": None,
                     "peak_source": "measured on this device: DFMA micro-benchmark (lsq_measure_fp64_peak)", "ms_per_launch": kms,
                     "flops_note": "SURVEY 8(d): C*(7K+2) per EM iteration, summed over events with their iteration counts",
                     "hbm_gbs_algorithmic": bytes_alg / (kms * 1e-3) / 1e9},
        "em_iters_mean": float(it.mean()), "em_iters_max": int(it.max()),
    }
    print(json.dumps(line), flush=True)
    ctx.close()


def main():
    args = parse_args()
    if args.workload == "C5":
        if args.impl == "reference":
            print(json.dumps({"impl": "reference", "unavailable":
                              "C5 feeds class tables directly; the reference has no such entry point"}))
            return
        return run_solve_stress(args)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
