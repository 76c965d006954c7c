"""GPU parity tests (run on the B200 box: pytest -m gpu).  Everything goes through the C ABI
(liblesseq_b200.so) -- either via the drop-in count/solve executables or via the ctypes
mirror -- and is compared with (a) the golden outputs of the reference's own binaries and
(b) the CPU oracle on the same seeded inputs."""
import os

import numpy as np
import pytest

import fuzzgen
import golden_util
from parity_util import num_close, tables_match

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def api():
    from lesseq_b200 import build
    build.build()
    from lesseq_b200 import api as _api
    return _api


SEED_CASES = golden_util.load("seed.json")
FUZZ_CASES = (golden_util.load("fuzz.json") + golden_util.load("weird.json") + golden_util.load("em_fragile.json") +
              golden_util.load("em_orderfree.json"))


@pytest.mark.parametrize("case", SEED_CASES, ids=[c["name"] for c in SEED_CASES])
def test_cli_matches_reference_golden(case, api, tmp_path):
    """Drop-in executables (text in, table out) against the reference binaries' outputs."""
    ca, sa = golden_util.materialise(case, str(tmp_path))
    exp = case["expect"]
    for prog, args in (("count", ca), ("solve", sa)):
        for prec in (17, 6):
            rc, out, err = api.run_program(prog, args, precision=prec)
            assert rc == exp["%s_rc" % prog], err
            bad = tables_match(exp["%s%d" % (prog, prec)], out, prog == "solve", rel6=(prec == 6))
            assert bad is None, "%s %s prec=%d: %s" % (case["name"], prog, prec, bad)


def format_tables(api, ann, M, cnt, sol, prec):
    """The two stdout tables (count.cpp:486-492, solve.cpp:836-846) from C-ABI results."""
    g = lambda v: "%.*g" % (prec, v)  # noqa: E731
    names, forms, foff = ann.event_names(), ann.form_names(), ann.form_offsets()
    ct, st = [], []
    for e, name in enumerate(names):
        sup = "".join(g(float(cnt["n_valid"][m, e])) + "\t" for m in range(M))
        for f in range(int(foff[e]), int(foff[e + 1])):
            ct.append("%s\t%s%s\t%s\n" % (name, sup, forms[f], g(float(cnt["iso_count"][f]))))
            st.append("%s\t%s%s\t%s\t%s\t%s\n" % (name, sup, forms[f], g(sol["theta"][f]), g(sol["rpkm"][f]),
                                                  g(sol["ll_per_read"][e])))
    return "".join(ct), "".join(st)


def test_abi_matches_reference_golden_fuzz(api, tmp_path):
    """The 80 fuzz + 40 weird-annotation fixtures (reference binary outputs) through the ctypes mirror of the C ABI in
    one process: count tables identical, solve tables within tolerance."""
    ctx = api.Context(0)
    for ci, case in enumerate(FUZZ_CASES):
        d = tmp_path / case["name"]
        d.mkdir()
        ca, sa = golden_util.materialise(case, str(d))
        exp = case["expect"]
        if exp["count_rc"] != 0:
            continue
        ann = api.Annotation(ca[4], ca[6], case["begin"], case["end"])
        if ann.n_events == 0:
            assert exp["count17"] == ""
            continue
        ctx.upload_events(ann)
        M = len(case["sets"])
        for m, s in enumerate(case["sets"]):
            # every other case goes through the compact wire format (lsqh_reads_compact)
            ctx.set_reads(m, s["type"], s["read_length"], api.Reads.from_mrf(d / ("r%d.mrf" % m), compact=ci % 2 == 1))
        ctx.count()
        cnt = ctx.get_count()
        ctx.solve([s["total_read_bases"] for s in case["sets"]])
        sol = ctx.get_solve()
        for prec in (17, 6):
            ct, st = format_tables(api, ann, M, cnt, sol, prec)
            assert tables_match(exp["count%d" % prec], ct, False) is None, (case["name"], prec)
            bad = tables_match(exp["solve%d" % prec], st, True, rel6=(prec == 6))
            assert bad is None, (case["name"], prec, bad)
    ctx.close()


def test_large_fragile_events_match_reference_golden(api, tmp_path):
    """tests/golden/em_fragile_large.json: events of 40-80 k reads with a fragile EM stop.  The replay takes the path
    of large events (device-wide bitonic sort, lsq_exact_classify_kernel, ordered_sum); theta, RPKM and log-likelihood
    within 1e-9 of the reference binaries, which a class-collapsed EM alone misses (tests/test_host_cpu.py)."""
    for case in golden_util.load("em_fragile_large.json"):
        d = tmp_path / case["name"]
        d.mkdir()
        ca, sa = golden_util.materialise(case, str(d))
        ann = api.Annotation(ca[4], ca[6], case["begin"], case["end"])
        ctx = api.Context(0)
        ctx.upload_events(ann)
        M = len(case["sets"])
        for m, s in enumerate(case["sets"]):
            ctx.set_reads(m, s["type"], s["read_length"], api.Reads.from_mrf(d / ("r%d.mrf" % m), compact=True))
        ctx.count()
        cnt = ctx.get_count()
        ctx.solve([s["total_read_bases"] for s in case["sets"]])
        sol = ctx.get_solve()
        assert ctx.timings()["exact_events"] >= 1, case["name"]
        # the warp-cooperative running sums (default) and the one-lane form that the CPU selftest checks give the same bits
        os.environ["LSQ_EXACT_SCALAR_SUMS"] = "1"
        try:
            ctx.solve([s["total_read_bases"] for s in case["sets"]])
            sol_scalar = ctx.get_solve()
        finally:
            del os.environ["LSQ_EXACT_SCALAR_SUMS"]
        for key in ("theta", "loglik", "rpkm"):
            assert np.array_equal(sol[key].view(np.uint64), sol_scalar[key].view(np.uint64)), (case["name"], key)
        assert np.array_equal(sol["iters"], sol_scalar["iters"]), case["name"]
        ct, st = format_tables(api, ann, M, cnt, sol, 17)
        assert tables_match(case["expect"]["count17"], ct, False) is None, case["name"]
        bad = tables_match(case["expect"]["solve17"], st, True)
        assert bad is None, (case["name"], bad)
        ctx.close()


def _oracle_dump(oracle_bin, prog, args, path):
    rc, out, err = oracle_bin.run_oracle(prog, args, dump=path)
    assert rc == 0, err
    classes, bases, iters = {}, {}, {}
    for line in open(path):
        f = line.rstrip("\n").split("\t")
        if f[0] == "C":
            classes.setdefault((f[1], int(f[2])), {})[int(f[3])] = int(f[4])
        elif f[0] == "B":
            bases[(f[1], int(f[2]))] = (int(float(f[3])), int(float(f[4])))
        elif f[0] == "E":
            iters[f[1]] = int(f[3])
    return out, classes, bases, iters


def test_abi_matches_oracle_on_fuzz(api, oracle_bin, tmp_path):
    """Per-event n_valid, bases, class histogram (bit-exact), EM iteration count, theta, rpkm and
    log-likelihood (1e-9) through the ctypes mirror of the C ABI, one context reused."""
    ctx = api.Context(0)
    checked = 0
    for seed in range(9000, 9150):
        case = fuzzgen.make_case(seed)
        d = tmp_path / ("s%d" % seed)
        d.mkdir()
        ca, sa = fuzzgen.write_case(case, str(d))
        rc, _, _ = oracle_bin.run_oracle("count", ca)
        if rc != 0:
            continue
        out, classes, bases, iters = _oracle_dump(oracle_bin, "solve", sa, str(d / "dump.txt"))
        ann = api.Annotation(ca[4], ca[6], case["begin"], case["end"])
        if ann.n_events == 0:
            continue
        ctx.upload_events(ann)
        M = len(case["sets"])
        for m, s in enumerate(case["sets"]):
            ctx.set_reads(m, s["type"], s["read_length"], api.Reads.from_mrf(d / ("r%d.mrf" % m)))
        ctx.count()
        cnt = ctx.get_count()
        ctx.solve([s["total_read_bases"] for s in case["sets"]])
        sol = ctx.get_solve()
        names = ann.event_names()
        foff = ann.form_offsets()
        rows = [r for r in out.splitlines()]
        for m in range(M):
            off, mask, c = ctx.get_classes(m)
            for e, g in enumerate(names):
                got = {int(mask[i]): int(c[i]) for i in range(int(off[e]), int(off[e + 1]))}
                assert got == classes.get((g, m), {}), (seed, g, m)
                assert (int(cnt["n_valid"][m, e]), int(cnt["bases"][m, e])) == bases[(g, m)], (seed, g, m)
        for e, g in enumerate(names):
            assert int(sol["iters"][e]) == iters[g], (seed, g)
            for f in range(int(foff[e]), int(foff[e + 1])):
                cols = rows[f].split("\t")
                assert num_close(float(cols[-3]), float(sol["theta"][f])), (seed, g, cols, sol["theta"][f])
                assert num_close(float(cols[-2]), float(sol["rpkm"][f])), (seed, g)
                assert num_close(float(cols[-1]), float(sol["ll_per_read"][e])), (seed, g)
        checked += 1
    ctx.close()
    assert checked > 100


def _synth_case(api, tmp_path, name, n_genes, n_reads, seed, **kw):
    from lesseq_b200 import synth
    A, R, _ = synth.make_config(name, n_genes=n_genes)
    fp, mp, rp = str(tmp_path / "ev.interval"), str(tmp_path / "ev.map"), str(tmp_path / "r.mrf")
    A.write(fp, mp)
    reads = synth.make_reads(A, n_reads, R, seed, device="cpu", **kw)
    synth.reads_to_mrf(reads, rp)
    head = ["0", "p", str(tmp_path / "o_"), "LH_GENE_TXT", fp, "UCSC_GENE2ISOFORM", mp, "0", "100000000"]
    return A, R, reads, head, rp


@pytest.mark.parametrize("name,n_genes,n_reads,kw", [
    ("C1", 500, 1000000, {}),   # BASELINE config 1 at its stated size
    ("C2", 400, 200000, {}),
    ("C4", 400, 200000, {"overhang_frac": 0.02}),
    ("C2", 3000, 2000000, {"overhang_frac": 0.01}),
])
def test_synthetic_configs_match_oracle(api, oracle_bin, tmp_path, name, n_genes, n_reads, kw):
    """BASELINE-shaped inputs at a size the oracle finishes in seconds, through the drop-in
    executables (text in, table out), 17 digits."""
    A, R, reads, head, rp = _synth_case(api, tmp_path, name, n_genes, n_reads, 11, **kw)
    ca = head + ["MRF_SINGLE", "SHORT_READ", str(R), rp]
    sa = ca + [repr(reads["total_read_bases"])]
    for prog, args in (("count", ca), ("solve", sa)):
        rc1, o1, e1 = oracle_bin.run_oracle(prog, args)
        rc2, o2, e2 = api.run_program(prog, args, precision=17)
        assert rc1 == 0 and rc2 == 0, e2
        assert len(o1.splitlines()) == 2 * len(A.ev_gene)
        bad = tables_match(o1, o2, prog == "solve")
        assert bad is None, "%s %s: %s" % (name, prog, bad)


def test_c2_full_annotation_5m_reads_match_oracle(api, oracle_bin, tmp_path):
    """BASELINE config 2's FULL annotation (20 000 genes, 69 930 events of all 8 types) with a 5 M-read text subsample of
    the benched workload, through the drop-in executables at 17 digits against the pinned oracle."""
    import torch
    from lesseq_b200 import synth
    A, R, _ = synth.make_config("C2")
    fp, mp, rp = str(tmp_path / "ev.interval"), str(tmp_path / "ev.map"), str(tmp_path / "r.mrf")
    A.write(fp, mp)
    reads = synth.make_reads(A, 5_000_000, R, 100, device="cuda")  # the bench's generator and seed, first 5 M reads' worth
    hr = api.Reads.from_numpy(reads["blk_off"].cpu().numpy(), reads["blk"].cpu().numpy(), reads["read_cs"].cpu().numpy(),
                              reads["line_num"].cpu().numpy(), reads["chrom_names"], reads["strand_names"])
    hr.write_mrf(rp)
    trb = reads["total_read_bases"]
    del reads
    torch.cuda.empty_cache()
    ca = ["0", "p", str(tmp_path / "o_"), "LH_GENE_TXT", fp, "UCSC_GENE2ISOFORM", mp, "0", "100000000", "MRF_SINGLE", "SHORT_READ", str(R), rp]
    for prog, args in (("count", ca), ("solve", ca + [repr(trb)])):
        rc1, o1, e1 = oracle_bin.run_oracle(prog, args)
        rc2, o2, e2 = api.run_program(prog, args, precision=17)
        assert rc1 == 0 and rc2 == 0, e2
        assert len(o1.splitlines()) == 2 * len(A.ev_gene) == 139860
        bad = tables_match(o1, o2, prog == "solve")
        assert bad is None, "%s: %s" % (prog, bad)


def test_c5_shaped_events_match_reference_solve_binary(api, oracle_bin, tmp_path):
    """BASELINE config 5 materialised: 2 000 events with 2-16 forms over K..2K segments and reads drawn from
    the forms, through the drop-in `solve` at 17 digits against the REFERENCE'S OWN BINARY (oracle/_ref/solve; the pinned
    oracle where the binaries are not available) -- the multi-form EM, the mid and warp solve kernels and the hashed
    class tables (K > 8) against the real thing rather than a numpy restatement."""
    from lesseq_b200 import synth
    A, reads = synth.make_multiform_case(2000, 120, 50, 55)
    fp, mp, rp = str(tmp_path / "ev.interval"), str(tmp_path / "ev.map"), str(tmp_path / "r.mrf")
    A.write(fp, mp)
    api.Reads.from_numpy(reads["blk_off"], reads["blk"], reads["read_cs"], reads["line_num"], reads["chrom_names"],
                         reads["strand_names"]).write_mrf(rp)
    ca = ["0", "p", str(tmp_path / "o_"), "LH_GENE_TXT", fp, "UCSC_GENE2ISOFORM", mp, "0", "100000000", "MRF_SINGLE", "MEDIUM_READ", "50", rp]
    sa = ca + [repr(reads["total_read_bases"])]
    use_ref = oracle_bin.ref_available() or oracle_bin.build_ref()
    for prog, args in (("count", ca), ("solve", sa)):
        rc1, o1, e1 = oracle_bin.run_ref(prog, args) if use_ref else oracle_bin.run_oracle(prog, args)
        rc2, o2, e2 = api.run_program(prog, args, precision=17)
        assert rc1 == 0 and rc2 == 0, (e1[-300:], e2[-300:])
        assert len(o1.splitlines()) == len(A.forms) > 15000
        bad = tables_match(o1, o2, prog == "solve")
        assert bad is None, "%s: %s" % (prog, bad)


def test_device_resident_reads_and_properties(api, tmp_path):
    """Size-independent properties on a larger device-generated input: classes sum to n_valid,
    per-form tallies follow from the classes, shuffling the reads changes nothing, counting twice
    is idempotent, and a two-shard run over the same covered regions concatenates to the full run."""
    import torch
    from lesseq_b200 import synth
    A, R, _ = synth.make_config("C2", n_genes=2000)
    fp, mp = str(tmp_path / "ev.interval"), str(tmp_path / "ev.map")
    A.write(fp, mp)
    n = 4_000_000
    reads = synth.make_reads(A, n, R, 5, device="cuda")

    def run(ann, rd):
        ctx = api.Context(0)
        ctx.upload_events(ann)
        ctx.set_reads(0, "SHORT_READ", R, api.Reads.from_torch(rd["blk_off"], rd["blk"], rd["read_cs"],
                                                               rd["line_num"], rd["chrom_names"], rd["strand_names"]))
        ctx.count()
        c1 = ctx.get_count()
        ctx.count()
        c2 = ctx.get_count()
        for k in c1:
            assert np.array_equal(c1[k], c2[k]), "count is not idempotent (%s)" % k
        off, mask, cnt = ctx.get_classes(0)
        ctx.solve([rd["total_read_bases"]])
        sol = ctx.get_solve()
        t = ctx.timings()
        ctx.close()
        return c1, (off, mask, cnt), sol, t

    ann = api.Annotation(fp, mp)
    c, (off, mask, cnt), sol, t = run(ann, reads)
    assert t["reads_in"] == n and 0 < t["reads_single"] + t["reads_pair"] + t["reads_general"] <= n
    foff = ann.form_offsets()
    for e in range(ann.n_events):
        sl = slice(int(off[e]), int(off[e + 1]))
        assert int(cnt[sl].sum()) == int(c["n_valid"][0, e])
        for k in range(int(foff[e + 1] - foff[e])):
            assert int(cnt[sl][(mask[sl] >> np.uint64(k)) & np.uint64(1) == 1].sum()) == int(c["iso_count"][foff[e] + k])
    assert int(c["n_valid"].sum()) > n // 2
    th = sol["theta"]
    for e in range(ann.n_events):
        assert abs(th[foff[e]:foff[e + 1]].sum() - 1.0) < 1e-9
    # permutation invariance (host-side shuffle of the packed reads)
    perm = torch.randperm(n, device="cuda")
    nblk = (reads["blk_off"][1:] - reads["blk_off"][:-1]).to(torch.int64)
    off2 = torch.zeros(n + 1, dtype=torch.int64, device="cuda")
    off2[1:] = torch.cumsum(nblk[perm], 0)
    idx = torch.repeat_interleave(reads["blk_off"][:-1].to(torch.int64)[perm] - off2[:-1], nblk[perm]) + \
        torch.arange(int(off2[-1]), device="cuda")
    shuffled = dict(reads)
    shuffled["blk_off"] = off2.to(torch.int32).contiguous()
    shuffled["blk"] = reads["blk"].view(-1, 2)[idx].contiguous().view(-1)
    shuffled["read_cs"] = reads["read_cs"][perm].contiguous()
    shuffled["line_num"] = reads["line_num"][perm].contiguous()
    c3, cl3, _, _ = run(ann, shuffled)
    for k in c:
        assert np.array_equal(c[k], c3[k]), "shuffling reads changed %s" % k
    # shards: covered regions of the whole range, events of each half
    half = ann.n_events // 2
    parts = []
    for sh in ((0, half), (half, ann.n_events)):
        a2 = api.Annotation(fp, mp, shard=sh)
        parts.append(run(a2, reads)[0])
    for k in ("n_valid", "bases"):
        assert np.array_equal(np.concatenate([p[k] for p in parts], axis=1), c[k]), k
    assert np.array_equal(np.concatenate([p["iso_count"] for p in parts]), c["iso_count"])


def test_streamed_host_reads_match_resident(api, tmp_path):
    """LSQ_READS_STREAM: the filter consumes blk[] while the copy engine is still filling it, from the
    wide or the compact wire format.  Results must be identical to the device-resident run, for a set
    large enough to take many chunks, a tiny one, and when the same context streams repeatedly."""
    import torch
    from lesseq_b200 import synth
    A, R, _ = synth.make_config("C2", n_genes=2000)
    fp, mp = str(tmp_path / "ev.interval"), str(tmp_path / "ev.map")
    A.write(fp, mp)
    ann = api.Annotation(fp, mp)
    ctx = api.Context(0)
    ctx.upload_events(ann)
    for n in (12_000_000, 5_000, 12_000_000):
        rd = synth.make_reads(A, n, R, 11 + n % 7, device="cuda")
        ctx.set_reads(0, "SHORT_READ", R, api.Reads.from_torch(rd["blk_off"], rd["blk"], rd["read_cs"],
                                                               rd["line_num"], rd["chrom_names"], rd["strand_names"]))
        ctx.count()
        ref = ctx.get_count()
        cnt = (rd["blk_off"][1:] - rd["blk_off"][:-1]).to(torch.uint8)
        host = {k: rd[k].cpu().pin_memory() for k in ("blk", "read_cs")}
        host["blk_cnt"] = cnt.cpu().pin_memory()
        b2 = rd["blk"].view(-1, 2)
        cs = rd["read_cs"]
        host["blk_start"] = b2[:, 0].contiguous().cpu().pin_memory()
        host["blk_len"] = (b2[:, 1] - b2[:, 0]).to(torch.int16).cpu().pin_memory()
        host["blk_len8"] = (b2[:, 1] - b2[:, 0]).to(torch.uint8).cpu().pin_memory()  # 75-bp reads: every block < 256
        host["read_cs8"] = ((cs & 0x3F) | (((cs >> 16) & 1) << 6) | (((cs >> 31) & 1) << 7)).to(torch.uint8).cpu().pin_memory()
        for stream in (True, False):
            for compact in (False, "blk_len", "blk_len8"):
                if compact:
                    hr = api.Reads.from_torch(None, None, None, None, rd["chrom_names"], rd["strand_names"],
                                              blk_cnt=host["blk_cnt"], stream=stream, read_cs8=host["read_cs8"],
                                              blk_start=host["blk_start"], **{compact: host[compact]})
                else:
                    hr = api.Reads.from_torch(None, host["blk"], host["read_cs"], None, rd["chrom_names"],
                                              rd["strand_names"], blk_cnt=host["blk_cnt"], stream=stream)
                ctx.set_reads(0, "SHORT_READ", R, hr)
                ctx.count()
                got = ctx.get_count()
                for k in ref:
                    assert np.array_equal(ref[k], got[k]), "streamed=%s compact=%s n=%d: %s differs" % (stream, compact, n, k)
    ctx.close()


def test_streamed_two_read_sets_pageable_and_csr(api, tmp_path):
    """LSQ_READS_STREAM with two read sets in flight on the copy stream at once, from pageable (not
    pinned) host memory, and with CSR offsets (blk_off) instead of block counts: same tables as the
    device-resident run of the same two sets."""
    import torch
    from lesseq_b200 import synth
    A, R, _ = synth.make_config("C4", n_genes=1500)
    fp, mp = str(tmp_path / "ev.interval"), str(tmp_path / "ev.map")
    A.write(fp, mp)
    ann = api.Annotation(fp, mp)
    ctx = api.Context(0)
    ctx.upload_events(ann)
    sets = [synth.make_reads(A, n, R, seed, device="cuda", overhang_frac=0.02) for n, seed in ((3_000_000, 21), (1_000_000, 22))]

    def run(make):
        for m, rd in enumerate(sets):
            ctx.set_reads(m, "SHORT_READ" if m == 0 else "MEDIUM_READ", R, make(rd))
        ctx.count()
        c = ctx.get_count()
        ctx.solve([rd["total_read_bases"] for rd in sets])
        return c, ctx.get_solve()

    keep = []

    def resident(rd):
        return api.Reads.from_torch(rd["blk_off"], rd["blk"], rd["read_cs"], rd["line_num"], rd["chrom_names"],
                                    rd["strand_names"])

    def streamed_pageable_csr(rd):
        h = {k: rd[k].cpu() for k in ("blk_off", "blk", "read_cs", "line_num")}  # pageable
        keep.append(h)
        return api.Reads.from_torch(h["blk_off"], h["blk"], h["read_cs"], h["line_num"], rd["chrom_names"],
                                    rd["strand_names"], stream=True)

    c0, s0 = run(resident)
    c1, s1 = run(streamed_pageable_csr)
    assert int(c0["n_valid"].sum()) > 1_000_000
    for k in c0:
        assert np.array_equal(c0[k], c1[k]), k
    for k in ("theta", "rpkm", "loglik"):
        assert np.array_equal(s0[k], s1[k], equal_nan=True), k
    ctx.close()


def test_cli_read_cache_is_transparent(api, tmp_path):
    """LSQ_READ_CACHE=1: `count` writes <reads>.lsqpack, `solve` reads it; the tables are byte-identical
    to the uncached programs, a changed text file is re-parsed, and nothing is written without the switch."""
    case = next(c for c in SEED_CASES if c["expect"]["count_rc"] == 0 and c["expect"]["count17"])
    ca, sa = golden_util.materialise(case, str(tmp_path))
    mrfs = [a for a in ca if str(a).endswith(".mrf")]
    assert mrfs
    plain = {p: api.run_program(p, a, precision=17) for p, a in (("count", ca), ("solve", sa))}
    assert not any(os.path.exists(str(m) + ".lsqpack") for m in mrfs)
    on = {"LSQ_READ_CACHE": "1"}
    rc, out, err = api.run_program("count", ca, precision=17, env_extra=on)
    assert (rc, out) == plain["count"][:2], err
    assert all(os.path.exists(str(m) + ".lsqpack") for m in mrfs)
    stamps = [os.path.getmtime(str(m) + ".lsqpack") for m in mrfs]
    rc, out, err = api.run_program("solve", sa, precision=17, env_extra=on)
    assert (rc, out) == plain["solve"][:2], err
    assert stamps == [os.path.getmtime(str(m) + ".lsqpack") for m in mrfs]  # reused, not rewritten
    # the text changes: the stale pack must not be used
    with open(str(mrfs[0]), "a") as f:
        f.write("chrNowhere:+:1:50\n")
    rc, out2, err = api.run_program("count", ca, precision=17, env_extra=on)
    rc3, out3, _ = api.run_program("count", ca, precision=17)
    assert rc == rc3 == 0 and out2 == out3


def _walk_reads(rng, forms, L, n):
    """Reads of length <= L walking along randomly chosen forms (blocks cut at exon ends)."""
    lines = []
    for _ in range(n):
        ex = forms[rng.randrange(len(forms))]
        tot = sum(e - s for s, e in ex)
        off = rng.randrange(0, max(1, tot - 1))
        need = L
        acc = 0
        parts = []
        q = 1
        for (s, e) in ex:
            if off >= acc + (e - s):
                acc += e - s
                continue
            bs = s + max(0, off - acc)
            take = min(e - bs, need)
            if parts and parts[-1][1] == bs:  # abutting exons give one block
                parts[-1][1] = bs + take
            else:
                parts.append([bs, bs + take])
            need -= take
            acc += e - s
            off = acc
            if need <= 0:
                break
        st = "+" if rng.random() < 0.5 else "-"
        lines.append(",".join("chr1:%s:%d:%d:%d:%d" % (st, a + 1, b, 1, b - a) for a, b in parts))
    return lines


def test_wide_and_many_form_events_match_oracle(api, oracle_bin, tmp_path):
    """Events outside the common shape: > 64 segments (prefix-count compatibility test) and > 8
    forms (hashed class table), plus a plain event sharing the region."""
    import random
    rng = random.Random(17)
    pos, exons = 1000, []
    for _ in range(80):
        ln = rng.randint(20, 90)
        exons.append((pos, pos + ln))
        pos += ln + (0 if rng.random() < 0.15 else rng.randint(30, 200))
    forms, fl, ml = [], [], []

    def add(ev, name, ex):
        fl.append("%s\tchr1\t+\t%d\t%d\t%d\t%s\t%s\n" % (name, ex[0][0], ex[-1][1], len(ex),
                                                     ",".join(str(s) for s, _ in ex), ",".join(str(e) for _, e in ex)))
        ml.append("%s\t%s\n" % (ev, name))
        forms.append(ex)

    add("WIDE", "W|1", exons)
    add("WIDE", "W|2", [x for i, x in enumerate(exons) if i % 7 != 3])
    add("WIDE", "W|3", [x for i, x in enumerate(exons) if i % 5 != 1 or i < 10])
    sub = exons[20:34]
    for k in range(11):
        keep = [x for i, x in enumerate(sub) if i in (0, len(sub) - 1) or rng.random() < 0.7]
        add("MANY", "M|%d" % (k + 1), keep)
    add("PLAIN", "P|1", exons[40:43])
    add("PLAIN", "P|2", [exons[40], exons[42]])
    fp, mp, rp = tmp_path / "ev.interval", tmp_path / "ev.map", tmp_path / "r.mrf"
    fp.write_text("".join(fl))
    mp.write_text("".join(ml))
    lines = _walk_reads(rng, forms, 100, 30000)
    rp.write_text("AlignmentBlocks\n" + "\n".join(lines) + "\n")
    args = ["0", "p", str(tmp_path / "o_"), "LH_GENE_TXT", str(fp), "UCSC_GENE2ISOFORM", str(mp), "0", "100",
            "MRF_SINGLE", "SHORT_READ", "100", str(rp)]
    for prog, a in (("count", args), ("solve", args + ["3000000"])):
        rc1, o1, e1 = oracle_bin.run_oracle(prog, a)
        rc2, o2, e2 = api.run_program(prog, a, precision=17)
        assert rc1 == 0 and rc2 == 0, e2
        assert sum(float(r.split("\t")[1]) for r in o1.splitlines()) > 10000
        bad = tables_match(o1, o2, prog == "solve")
        assert bad is None, "%s: %s" % (prog, bad)


@pytest.mark.parametrize("depth", [70, 700])
def test_deeply_nested_events_match_oracle(api, oracle_bin, tmp_path, depth):
    """`depth` events stacked on the same exons: stab lists longer than a partitioned record can
    name (>= 64: FAST and PAIR reads are demoted to the GENERAL class) and, at 700, longer than the
    count kernel stages in shared memory (the bucket is read through L1/L2 instead)."""
    import random
    rng = random.Random(depth)
    exons = [(1000, 1080), (1300, 1390), (1390, 1450), (1700, 1800), (2100, 2160)]
    fl, ml, forms = [], [], []

    def add(ev, name, ex):
        fl.append("%s\tchr1\t+\t%d\t%d\t%d\t%s\t%s\n" % (name, ex[0][0], ex[-1][1], len(ex),
                                                     ",".join(str(s) for s, _ in ex), ",".join(str(e) for _, e in ex)))
        ml.append("%s\t%s\n" % (ev, name))
        forms.append(ex)

    for d in range(depth):
        lo = rng.randrange(0, 2)
        hi = rng.randrange(4, 6)
        ex = exons[lo:hi]
        add("N%d" % d, "N%d|1" % d, ex)
        add("N%d" % d, "N%d|2" % d, [x for i, x in enumerate(ex) if i != 1 + rng.randrange(len(ex) - 2)])
    add("SIDE", "S|1", [(5000, 5100), (5300, 5400)])
    add("SIDE", "S|2", [(5000, 5100), (5200, 5250), (5300, 5400)])
    fp, mp, rp = tmp_path / "ev.interval", tmp_path / "ev.map", tmp_path / "r.mrf"
    fp.write_text("".join(fl))
    mp.write_text("".join(ml))
    lines = _walk_reads(rng, forms, 60, 6000 if depth < 100 else 3000)
    rp.write_text("AlignmentBlocks\n" + "\n".join(lines) + "\n")
    args = ["0", "p", str(tmp_path / "o_"), "LH_GENE_TXT", str(fp), "UCSC_GENE2ISOFORM", str(mp), "0", "100000",
            "MRF_SINGLE", "SHORT_READ", "60", str(rp)]
    for prog, a in (("count", args), ("solve", args + ["360000"])):
        rc1, o1, e1 = oracle_bin.run_oracle(prog, a)
        rc2, o2, e2 = api.run_program(prog, a, precision=17)
        assert rc1 == 0 and rc2 == 0, e2
        assert sum(float(r.split("\t")[1]) for r in o1.splitlines()) > 10000
        bad = tables_match(o1, o2, prog == "solve")
        assert bad is None, "%s: %s" % (prog, bad)


def _numpy_em(K, G, classes, n_total):
    """EM of common/read.h:592-660 over (mask, count) classes, sequential double arithmetic."""
    theta = np.full(K, 1.0 / K)

    def ll(th):
        s = 0.0
        for mk, c in classes:
            p = 0.0
            for k in range(K):
                if (mk >> k) & 1:
                    p = p + th[k] * G[k]
            s += c * np.log(p) if p > 0 else -np.inf
        return s

    old = ll(theta)
    it = 0
    while True:
        new = np.zeros(K)
        for k in range(K):
            z = 0.0
            for mk, c in classes:
                if not (mk >> k) & 1:
                    continue
                p = 0.0
                for j in range(K):
                    if (mk >> j) & 1:
                        p = p + theta[j] * G[j]
                lp = theta[k] * G[k]
                if p > 0 and lp > 0:
                    z += c * (lp / p)
            new[k] = z / n_total
        theta = new
        cur = ll(theta)
        it += 1
        with np.errstate(invalid="ignore", divide="ignore"):
            if not (abs(1.0 - old / cur) > 1e-6):
                break
        old = cur
    return theta, cur, it


def test_solve_stress_injected_classes(api):
    """BASELINE config 5 shape at test size: events with 2-16 forms, class tables injected through
    lsq_set_classes, warp-per-event EM against a numpy restatement of read.h:592-660."""
    from lesseq_b200 import synth
    E = 3000
    S = synth.make_solve_stress(E, 5)
    ann = api.ArrayAnnotation(
        ev_chrom=np.zeros(E), ev_strand=np.zeros(E), ev_gs=S["ev_gs"], ev_ge=S["ev_ge"], ev_seg_off=S["seg_off"],
        ev_form_off=S["form_off"], seg_start=S["seg_start"], seg_end=S["seg_end"], form_seg_off=S["form_seg_off"],
        form_seg_idx=S["form_seg_idx"], cov_off=[0, 1], cov_start=[0], cov_end=[int(S["ev_ge"][-1]) + 10],
        chrom_names=["chr1"], strand_names=["+"])
    ctx = api.Context(0)
    ctx.upload_events(ann)
    ctx.set_classes(0, "SHORT_READ", 75, S["cls_off"], S["cls_mask"], S["cls_cnt"], S["bases"])
    ctx.solve([1e9])
    sol = ctx.get_solve()
    cnt = ctx.get_count()
    ctx.close()
    foff, coff = S["form_off"], S["cls_off"]
    rng = np.random.default_rng(0)
    for e in rng.choice(E, 150, replace=False):
        K = int(S["K"][e])
        classes = [(int(S["cls_mask"][i]), int(S["cls_cnt"][i])) for i in range(int(coff[e]), int(coff[e + 1]))]
        n_total = float(sum(c for _, c in classes))
        assert int(cnt["n_valid"][0, e]) == int(n_total)
        G = []
        for k in range(K):
            f = int(foff[e]) + k
            idx = S["form_seg_idx"][int(S["form_seg_off"][f]):int(S["form_seg_off"][f + 1])]
            lens = (S["seg_end"] - S["seg_start"])[int(S["seg_off"][e]) + idx]
            ars = api.ars_length("SHORT_READ", 75, lens)
            G.append(1.0 / ars if ars > 0 else 0.0)
        th, ll, it = _numpy_em(K, G, classes, n_total)
        assert int(sol["iters"][e]) == it, (e, K, sol["iters"][e], it)
        for k in range(K):
            assert num_close(float(sol["theta"][int(foff[e]) + k]), float(th[k])), (e, k)
        assert num_close(float(sol["loglik"][e]), float(ll)), e
        # standard errors (extension, no reference oracle): observed Fisher information of the free
        # parameters, numpy restatement of the per-read term of common/fim.h:322-367
        D = K - 1
        info = np.zeros((D, D))
        for mk, c in classes:
            v = np.array([G[k] if (mk >> k) & 1 else 0.0 for k in range(K)])
            sc = float(np.dot(th, v))
            if sc > 0:
                d = (v[:D] - v[D]) / sc
                info += c * np.outer(d, d)
        se = sol["se"][int(foff[e]):int(foff[e]) + K]
        if np.linalg.cond(info) < 1e8:
            cov = np.linalg.inv(info)
            exp = np.sqrt(np.concatenate([np.diag(cov), [cov.sum()]]))
            assert np.allclose(se, exp, rtol=1e-6, atol=0), (e, K, se, exp)
        else:
            assert np.all(np.isnan(se) | (se >= 0))


def test_other_annotation_formats_match_reference(api, tmp_path):
    """`solve` with UCSC_GENE_TXT / UCSC_GFF / WORMBASE_GFF2 / GENELETS_GFF3 forms and WORMBASE_GENE2ISOFORMS
    maps (solve/solve.cpp:158-329) against the reference binary's outputs (tests/golden/formats.json): every
    case through the ctypes mirror in one process, every seventh through the drop-in executable; `count`
    keeps refusing them (count.cpp:141-183)."""
    cases = golden_util.load("formats.json")
    ctx = api.Context(0)
    for ci, case in enumerate(cases):
        d = tmp_path / ("c%d" % ci)
        d.mkdir()
        ca, sa = golden_util.materialise_formats(case, str(d))
        exp = case["expect"]
        ann = api.Annotation(sa[4], sa[6], case["begin"], case["end"], forms_format=case["forms_format"],
                             map_format=case["map_format"])
        assert ann.n_events > 0
        ctx.upload_events(ann)
        M = len(case["sets"])
        for m, st in enumerate(case["sets"]):
            ctx.set_reads(m, st["type"], st["read_length"], api.Reads.from_mrf(d / ("r%d.mrf" % m), compact=ci % 2 == 1))
        ctx.count()
        cnt = ctx.get_count()
        ctx.solve([st["total_read_bases"] for st in case["sets"]])
        sol = ctx.get_solve()
        for prec in (17, 6):
            _, st_tab = format_tables(api, ann, M, cnt, sol, prec)
            bad = tables_match(exp["solve%d" % prec], st_tab, True, rel6=(prec == 6))
            assert bad is None, (case["name"], prec, bad)
        if ci % 7 == 0:
            rc, out, err = api.run_program("solve", sa, precision=17)
            assert rc == 0, err
            assert tables_match(exp["solve17"], out, True) is None, case["name"]
            rc, out, err = api.run_program("count", ca)
            assert rc == 1 and out == "" and "Unknown file format error" in err
    ctx.close()


def test_empty_and_header_only_reads_match_oracle(api, oracle_bin, tmp_path):
    """A reads file with no bytes at all, or with nothing but its first line, counts nothing: every event prints
    zero counts and the no-read branch of solve (uniform theta, zero rpkm; solve.cpp:797-826), as the reference does."""
    case = golden_util.load("seed.json")[0]
    ca, sa = golden_util.materialise(case, str(tmp_path))
    for text in ("", "AlignmentBlocks\n", "AlignmentBlocks\n# only a comment\n"):
        p = tmp_path / "none.mrf"
        p.write_text(text)
        c2, s2 = list(ca), list(sa)
        c2[-1] = str(p)
        s2[-2] = str(p)
        for prog, a in (("count", c2), ("solve", s2)):
            rc1, o1, _ = oracle_bin.run_oracle(prog, a)
            rc2, o2, err = api.run_program(prog, a, precision=17)
            assert rc1 == 0 and rc2 == 0, err
            assert o1 != "" and tables_match(o1, o2, prog == "solve") is None, (repr(text), prog)


def test_merged_then_appended_then_out_of_order_blocks(api, oracle_bin, tmp_path):
    """The live-fuzz regression of tests/test_host_cpu.py on the device: overlap, append, out-of-order blocks."""
    (tmp_path / "ev.interval").write_text(
        "~t|1\tchr2\t+\t749\t1039\t4\t749,792,888,1000\t770,888,916,1039,\n"
        "~t|2\tchr2\t-\t749\t1039\t6\t749,753,792,888,910,987,\t751,770,892,910,916,1039\n"
        "~t|3\tchr2\t-\t760\t1039\t4\t760,792,888,1000\t770,875,916,1039\n")
    (tmp_path / "ev.map").write_text("~t\t~t|1\n~t\t~t|2\n~t\t~t|3\n")
    (tmp_path / "r.mrf").write_text(
        "AlignmentBlocks\n"
        "chr2:-:843:892:1:50,chr2:-:889:910:51:72,chr2:-:913:916:73:76,chr2:-:911:912:77:78,chr2:-:988:1009:79:100\n"
        "chr2:-:843:892:1:50,chr2:-:889:910:51:72,chr2:-:913:916:73:76,chr2:-:911:912:77:78\n"
        "chr2:-:843:892:1:50,chr2:-:889:910:51:72,chr2:-:913:916:73:76\n")
    args = ["0", "p", str(tmp_path / "o_"), "LH_GENE_TXT", str(tmp_path / "ev.interval"), "UCSC_GENE2ISOFORM",
            str(tmp_path / "ev.map"), "0", "100", "MRF_SINGLE", "SHORT_READ", "36", str(tmp_path / "r.mrf")]
    for prog, a in (("count", args), ("solve", args + ["288.0"])):
        rc1, o1, _ = oracle_bin.run_oracle(prog, a)
        rc2, o2, err = api.run_program(prog, a, precision=17)
        assert rc1 == 0 and rc2 == 0, err
        assert tables_match(o1, o2, prog == "solve") is None, prog
