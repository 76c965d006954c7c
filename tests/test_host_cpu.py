"""CPU-side tests of the product: the C-ABI library loads and exports every declared symbol,
fails loudly without a GPU, and the host front end + device functions (replayed on the CPU by
tests/emul) agree with the oracle."""
import ctypes as C
import math
import os
import re
import subprocess

import numpy as np
import pytest

import fuzzgen
import golden_util

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EMUL = os.path.join(ROOT, "tests", "emul", "_build", "emul")


@pytest.fixture(scope="session")
def built_lib():
    from lesseq_b200 import build, _lib
    build.build()
    return _lib.load()


@pytest.fixture(scope="session")
def emul_bin():
    subprocess.run(["make", "-C", os.path.join(ROOT, "tests", "emul"), "-s"], check=True)
    return EMUL


def test_library_exports_every_header_symbol(built_lib):
    hdr = open(os.path.join(ROOT, "include", "lesseq_b200.h")).read()
    declared = set(re.findall(r"\b(lsqh?_[a-z_0-9]+)\s*\(", hdr))
    from lesseq_b200 import _lib
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    for name in declared:
        assert hasattr(built_lib, name), name


def test_no_cpu_fallback(built_lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    ctx = C.c_void_p()
    rc = built_lib.lsq_ctx_create(0, C.byref(ctx))
    assert rc == -2
    assert b"no CPU fallback" in built_lib.lsq_last_error()
    from lesseq_b200 import api
    with pytest.raises(api.LsqError):
        api.Context(0)


def test_ars_closed_form_against_enumeration(built_lib):
    """ARS length (accessible_read_starts.h:48-89, 221-274) against a direct enumeration of the
    reference's per-segment interval arithmetic."""
    from lesseq_b200 import api
    rng = np.random.default_rng(0)

    def brute(short, R, lens):
        T = sum(lens)
        total, c = 0, 0
        for L in lens:
            c += L
            if c + R > T:
                return total + max(L + 1 - (c + R - T), 0)
            if short:
                pts = set(range(0, max(L - R + 1, 0))) | (set(range(max(L - R, 0), L + 1)) if L > 0 else set())
                total += len(pts)
            else:
                total += L
        return total

    for _ in range(500):
        n = int(rng.integers(1, 6))
        lens = [int(x) for x in rng.integers(1, 120, n)]
        R = int(rng.integers(0, 200))
        assert api.ars_length("SHORT_READ", R, lens) == brute(True, R, lens)
        assert api.ars_length("MEDIUM_READ", R, lens) == brute(False, R, lens)
    assert api.ars_length("SHORT_READ", 50, [100, 100, 100]) == 253
    assert api.ars_length("MEDIUM_READ", 50, [100, 100, 100]) == 251


def test_annotation_front_end_a1(built_lib, tmp_path):
    from lesseq_b200 import api
    case = [c for c in golden_util.load("seed.json") if c["name"] == "A1_se_basics"][0]
    ca, _ = golden_util.materialise(case, str(tmp_path))
    ann = api.Annotation(ca[4], ca[6], 0, 100)
    assert ann.n_events == 1 and ann.n_forms == 2
    assert ann.event_names() == ["1"] and ann.form_names() == ["E1|2|1", "E1|2|2"]
    assert list(ann.array("seg_start", 3)) == [100, 300, 600]
    assert list(ann.array("seg_end", 3)) == [200, 400, 700]
    assert list(ann.form_lengths()) == [300, 200]
    assert list(ann.array("cov_start", 3)) == [100, 300, 600]


def test_mrf_parser_errors_and_packing(built_lib, tmp_path):
    from lesseq_b200 import api
    p = tmp_path / "bad.mrf"
    p.write_text("AlignmentBlocks\nchr1:+:10:x20:1:11\n")
    with pytest.raises(api.LsqError) as ei:
        api.Reads.from_mrf(p)
    assert ei.value.code == -6
    p.write_text("AlignmentBlocks\n# c\nchr1:+:10:20:1:11,chr2:-:5:9:12:16\nAlignmentBlocks\nchr1:+:1:5\nchr1:+:7:9")
    r = api.Reads.from_mrf(p)
    assert r.n_reads == 2 and r.n_blocks == 3  # the unterminated last line is dropped
    ln = np.ctypeslib.as_array(C.cast(r.view.line_num, C.POINTER(C.c_uint32)), shape=(2,))
    assert list(ln) == [2, 4]  # comment and header lines advance the read number
    cs = np.ctypeslib.as_array(C.cast(r.view.read_cs, C.POINTER(C.c_uint32)), shape=(2,))
    assert cs[0] & 0x80000000 and not (cs[1] & 0x80000000)


def _close(a, b, solve):
    la, lb = a.splitlines(), b.splitlines()
    if len(la) != len(lb):
        return False
    for x, y in zip(la, lb):
        fx, fy = x.split("\t"), y.split("\t")
        if len(fx) != len(fy):
            return False
        for u, v in zip(fx, fy):
            if u == v:
                continue
            if not solve:
                return False
            du, dv = float(u), float(v)
            if math.isnan(du) and math.isnan(dv):
                continue
            if abs(du - dv) > 1e-9 * max(abs(du), abs(dv)):
                return False
    return True


def _run_emul(emul, prog, args):
    env = dict(os.environ)
    env["LSQ_PRECISION"] = "17"
    p = subprocess.run([emul, prog] + [str(a) for a in args], env=env, capture_output=True, text=True)
    return p.returncode, p.stdout


def test_device_functions_replayed_on_cpu_match_golden(emul_bin, tmp_path):
    """Host front end + table layout + per-read/per-pair device functions vs the reference
    binaries' golden outputs: counts bit-exact, solve columns within 1e-9 relative."""
    n = 0
    for case in golden_util.all_cases():
        if case["name"].startswith("A6"):
            continue
        d = tmp_path / case["name"]
        d.mkdir()
        ca, sa = golden_util.materialise(case, str(d))
        exp = case["expect"]
        for prog, args in (("count", ca), ("solve", sa)):
            if exp["%s_rc" % prog] != 0:
                continue
            rc, out = _run_emul(emul_bin, prog, args)
            assert rc == 0
            assert _close(exp["%s17" % prog], out, prog == "solve"), (case["name"], prog)
            n += 1
    assert n > 150


def test_device_functions_replayed_on_cpu_match_oracle_fuzz(emul_bin, oracle_bin, tmp_path):
    for seed in range(7000, 7120):
        case = fuzzgen.make_case(seed)
        d = tmp_path / ("s%d" % seed)
        d.mkdir()
        ca, sa = fuzzgen.write_case(case, str(d))
        for prog, args in (("count", ca), ("solve", sa)):
            rc1, o1, _ = oracle_bin.run_oracle(prog, args)
            rc2, o2 = _run_emul(emul_bin, prog, args)
            assert rc1 == rc2 and _close(o1, o2, prog == "solve"), (seed, prog)


def test_device_functions_replayed_on_cpu_match_oracle_midscale(emul_bin, oracle_bin, tmp_path):
    """The C2 / C4 shapes at a size where the piece lookup runs with wide buckets (shift >= 7), crowded buckets and
    their sub-buckets, and the MRF parser with several threads: host front end + table layout + device functions
    replayed on the CPU against the oracle (which is pinned to the reference binaries)."""
    from lesseq_b200 import synth
    for cfg, seed in (("C2", 7), ("C4", 8)):
        A, R, _ = synth.make_config(cfg, n_genes=3000)
        d = tmp_path / cfg
        d.mkdir()
        fp, mp, rp = str(d / "ev.interval"), str(d / "ev.map"), str(d / "r.mrf")
        A.write(fp, mp)
        rd = synth.make_reads(A, 60000, R, seed, device="cpu", overhang_frac=0.02 if cfg == "C4" else 0.0)
        synth.reads_to_mrf(rd, rp)
        assert os.path.getsize(rp) > (1 << 20)
        args = ["0", "p", str(d / "o_"), "LH_GENE_TXT", fp, "UCSC_GENE2ISOFORM", mp, "0", "10000000", "MRF_SINGLE",
                "SHORT_READ", str(R), rp]
        for prog, a in (("count", args), ("solve", args + [repr(rd["total_read_bases"])])):
            rc1, o1, _ = oracle_bin.run_oracle(prog, a)
            rc2, o2 = _run_emul(emul_bin, prog, a)
            assert rc1 == 0 and rc2 == 0 and len(o1.splitlines()) > 15000
            assert _close(o1, o2, prog == "solve"), (cfg, prog)


def test_sharded_replay_with_bucketed_reads_equals_unsharded(emul_bin, tmp_path):
    """Multi-GPU data flow on the CPU (LSQ_EMUL_SHARDS): reads are routed with the owner masks of the whole
    selection's pieces (the rule of lsq_route_kernel), each gene-rank shard counts its own bucket with its own
    tables (covered regions of the whole selection), and the concatenated tables equal the unsharded run byte
    for byte -- count.cpp:202-215, 319 with SURVEY 8(e)'s caveat."""
    from lesseq_b200 import synth
    cases = []
    for seed in range(7300, 7340):  # adversarial fuzz annotations: overlapping / nested events, chimeric reads
        case = fuzzgen.make_case(seed)
        d = tmp_path / ("s%d" % seed)
        d.mkdir()
        cases.append(fuzzgen.write_case(case, str(d)))
    A, R, _ = synth.make_config("C2", n_genes=1500)
    d = tmp_path / "c2"
    d.mkdir()
    fp, mp, rp = str(d / "ev.interval"), str(d / "ev.map"), str(d / "r.mrf")
    A.write(fp, mp)
    rd = synth.make_reads(A, 40000, R, 11, device="cpu")
    synth.reads_to_mrf(rd, rp)
    args = ["0", "p", str(d / "o_"), "LH_GENE_TXT", fp, "UCSC_GENE2ISOFORM", mp, "0", "10000000", "MRF_SINGLE",
            "SHORT_READ", str(R), rp]
    cases.append((args, args + [repr(rd["total_read_bases"])]))
    n = 0
    for ca, sa in cases:
        for prog, a in (("count", ca), ("solve", sa)):
            rc0, full = _run_emul(emul_bin, prog, a)
            if rc0 != 0:
                continue
            for shards in (2, 3, 8):
                env = dict(os.environ, LSQ_PRECISION="17", LSQ_EMUL_SHARDS=str(shards))
                p = subprocess.run([emul_bin, prog] + [str(x) for x in a], env=env, capture_output=True, text=True)
                assert p.returncode == 0 and p.stdout == full, (a[4], prog, shards)
                n += 1
    assert n > 150


def test_cli_argument_errors(built_lib, tmp_path):
    """Exit codes of the drop-in programs for malformed command lines (count.cpp:20-38, 95-129)."""
    from lesseq_b200 import api
    rc, out, err = api.run_program("count", ["0", "p", "o"])
    assert rc == 1 and out == "" and "Usage" in err
    case = golden_util.load("seed.json")[0]
    ca, sa = golden_util.materialise(case, str(tmp_path))
    bad = list(ca)
    bad[7] = "x"
    assert api.run_program("count", bad)[0] == 1
    bad = list(ca)
    bad[3] = "UCSC_GFF"
    rc, out, err = api.run_program("count", bad)
    assert rc == 1 and "Unknown file format error: UCSC_GFF" in err
    bad = list(ca)
    bad[4] = str(tmp_path / "missing")
    assert api.run_program("count", bad)[0] == 134
    rc, out, err = api.run_program("solve", sa[:-1])
    assert rc == 1 and "Usage" in err


def test_parallel_mrf_parser_matches_serial(built_lib, tmp_path):
    """The chunk-parallel MRF parser (csrc/host/mrf.cpp) must produce the serial result: read
    numbering across chunks (comment lines count), dictionary ids in order of first appearance,
    chimeric reads flagged."""
    import random
    from lesseq_b200 import api
    rng = random.Random(3)
    lines = ["AlignmentBlocks"]
    for i in range(60000):
        r = rng.random()
        if r < 0.01:
            lines.append("# comment %d" % i)
        elif r < 0.02:
            lines.append("chrB%d:.:5:9:1:5,chr1:+:100:120:6:26" % (i % 7))
        else:
            s = rng.randint(1, 10 ** 6)
            lines.append("chr%d:%s:%d:%d:1:50" % (rng.randint(1, 5), "+-"[rng.randint(0, 1)], s, s + 49))
    p = tmp_path / "big.mrf"
    p.write_text("\n".join(lines) + "\n")
    assert p.stat().st_size > (1 << 20)

    def load(threads):
        os.environ["LSQ_THREADS"] = str(threads)
        try:
            r = api.Reads.from_mrf(p)
        finally:
            del os.environ["LSQ_THREADS"]
        n, nb = r.n_reads, r.n_blocks
        u32 = C.POINTER(C.c_uint32)
        arr = lambda ptr, k: np.ctypeslib.as_array(C.cast(ptr, u32), shape=(k,)).copy()  # noqa: E731
        return (n, nb, arr(r.view.line_num, n), arr(r.view.read_cs, n), arr(r.view.blk_off, n + 1),
                np.ctypeslib.as_array(C.cast(r.view.blk, C.POINTER(C.c_int32)), shape=(2 * nb,)).copy(),
                [r.view.chrom_names[i] for i in range(r.view.n_chrom_names)],
                [r.view.strand_names[i] for i in range(r.view.n_strand_names)])

    a, b = load(1), load(7)
    assert a[0] == b[0] and a[1] == b[1] and a[6] == b[6] and a[7] == b[7]
    for x, y in zip(a[2:6], b[2:6]):
        assert np.array_equal(x, y)


def test_piece_lookup_bucket_index_selftest(emul_bin):
    """Two-level bucket index (lsq_tables_host.hpp) + find_piece / staged block_lookup
    (lsq_device_fn.cuh) against a binary search over the piece starts: clustered starts, bucket
    shifts 0..15, crowded buckets and sub-buckets, positions before/after every chromosome."""
    p = subprocess.run([emul_bin, "bucket-selftest", "120"], capture_output=True, text=True)
    assert p.returncode == 0, p.stderr[-2000:]
    shifts = {int(line.split()[4]) for line in p.stdout.splitlines() if line.startswith("ok")}
    assert len(shifts) >= 10 and max(shifts) >= 12


def test_ordered_sum_matches_sequential_sum_bit_for_bit(emul_bin):
    """ordered_sum (lsq_device_fn.cuh: the running sum of lsq_exact_kernel taken binade by binade with integer
    arithmetic) against the plain read-by-read loop on random class sequences: up to 64 classes, up to 3 M reads,
    short terms (exact half-way cases), full mantissas, log-like terms, zeros, long runs, both signs."""
    p = subprocess.run([emul_bin, "ordered-sum-selftest", "600"], capture_output=True, text=True)
    assert p.returncode == 0, p.stderr[-2000:]


def test_large_fragile_events_replayed_on_cpu_match_reference(emul_bin, tmp_path):
    """tests/golden/em_fragile_large.json (events of 40-80 k reads whose EM stop is decided inside the summation-order
    noise; outputs of the reference binaries): the CPU replay of the device functions matches with the read-by-read
    replay and does NOT match without it -- so the fixtures do exercise the replay."""
    for case in golden_util.load("em_fragile_large.json"):
        d = tmp_path / case["name"]
        d.mkdir()
        ca, sa = golden_util.materialise(case, str(d))
        rc, out = _run_emul(emul_bin, "solve", sa)
        assert rc == 0 and _close(case["expect"]["solve17"], out, True), case["name"]
        env = dict(os.environ, LSQ_PRECISION="17", LSQ_EMUL_NO_EXACT="1")
        p = subprocess.run([emul_bin, "solve"] + [str(a) for a in sa], env=env, capture_output=True, text=True)
        assert p.returncode == 0 and not _close(case["expect"]["solve17"], p.stdout, True), case["name"]


def test_order_free_stops_are_certified_not_replayed(emul_bin, tmp_path):
    """tests/golden/em_orderfree.json (reference binary outputs): forms no longer than a read drive theta to a vertex (or
    leave it uniform) with a log-likelihood of exactly 0 -- flagged by value (the stop test divides by zero), cleared by
    em_result_is_order_free.  Same tables with the certificate, with a replay instead of it, and with neither."""
    for case in golden_util.load("em_orderfree.json"):
        d = tmp_path / case["name"]
        d.mkdir()
        ca, sa = golden_util.materialise(case, str(d))
        flagged = {}
        for tag, extra in (("certified", {}), ("replayed", {"LSQ_EMUL_NO_CERT": "1"}), ("neither", {"LSQ_EMUL_NO_EXACT": "1"})):
            env = dict(os.environ, LSQ_PRECISION="17", LSQ_EMUL_STATS="1", **extra)
            p = subprocess.run([emul_bin, "solve"] + [str(a) for a in sa], env=env, capture_output=True, text=True)
            assert p.returncode == 0 and _close(case["expect"]["solve17"], p.stdout, True), (case["name"], tag)
            flagged[tag] = sum(1 for line in p.stderr.splitlines() if "fragile event" in line)
        assert flagged["certified"] == 0 and flagged["replayed"] >= 3, (case["name"], flagged)


def test_ctypes_mirrors_have_the_layout_of_the_c_header(tmp_path):
    """lesseq_b200/_lib.py mirrors the structs of include/lesseq_b200.h by hand: same size and same offset of the last
    member, checked against what the C compiler lays out (a field added on one side only would shift everything)."""
    from lesseq_b200 import _lib
    pairs = [("lsq_event_tables", _lib.EventTables), ("lsq_packed_reads", _lib.PackedReads), ("lsq_count_out", _lib.CountOut),
             ("lsq_solve_out", _lib.SolveOut), ("lsq_timings", _lib.Timings), ("lsq_shard_stats", _lib.ShardStats),
             ("lsq_device_results", _lib.DeviceResults)]
    src = tmp_path / "layout.c"
    body = "".join('    printf("%s %%zu %%zu\\n", sizeof(%s), offsetof(%s, %s));\n' % (c, c, c, t._fields_[-1][0]) for c, t in pairs)
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "lesseq_b200.h"\nint main(void) {\n' + body + "    return 0;\n}\n")
    exe = tmp_path / "layout"
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    subprocess.run(["gcc", "-I", os.path.join(root, "include"), "-o", str(exe), str(src)], check=True)
    out = dict((ln.split()[0], (int(ln.split()[1]), int(ln.split()[2]))) for ln in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.splitlines())
    for c, t in pairs:
        assert out[c] == (C.sizeof(t), getattr(t, t._fields_[-1][0]).offset), (c, out[c], C.sizeof(t))


def test_annotation_loader_is_thread_count_independent(tmp_path):
    """lsqh_annotation_load parses the forms file in stretches and replays the covered-region adds chromosome by
    chromosome on several threads (annotation.cpp): one thread and eight give identical tables -- the rows keep
    file order and every chromosome sees its adds in event order (interval_list add_interval is order-dependent)."""
    from lesseq_b200 import api, synth
    A, _, _ = synth.make_config("C2", n_genes=4000)
    fp, mp = str(tmp_path / "ev.interval"), str(tmp_path / "ev.map")
    A.write(fp, mp)
    assert os.path.getsize(fp) > (1 << 20)  # large enough for the parallel parse

    def load(threads):
        os.environ["LSQ_THREADS"] = str(threads)
        try:
            a = api.Annotation(fp, mp)
        finally:
            del os.environ["LSQ_THREADS"]
        t = a.tables
        arrays = {"ev_chrom": t.n_events, "ev_strand": t.n_events, "ev_gs": t.n_events, "ev_ge": t.n_events,
                  "ev_seg_off": t.n_events + 1, "ev_form_off": t.n_events + 1, "seg_start": t.n_segments, "seg_end": t.n_segments,
                  "form_seg_off": t.n_forms + 1, "cov_off": t.n_chroms + 1}
        out = {k: a.array(k, n) for k, n in arrays.items()}
        out["form_seg_idx"] = a.array("form_seg_idx", int(out["form_seg_off"][-1]))
        out["cov_start"] = a.array("cov_start", int(out["cov_off"][-1]))
        out["cov_end"] = a.array("cov_end", int(out["cov_off"][-1]))
        return out, a.event_names(), a.form_names(), a.form_lengths()

    one, eight = load(1), load(8)
    assert one[1] == eight[1] and one[2] == eight[2] and np.array_equal(one[3], eight[3])
    assert one[0]["ev_gs"].size > 10000 and one[0]["cov_start"].size > 10000
    for k in one[0]:
        assert np.array_equal(one[0][k], eight[0][k]), k


def test_compact_wire_format_is_lossless(tmp_path):
    """lsqh_reads_compact: blk_cnt / read_cs8 / blk_start + blk_len expand to the wide arrays."""
    import ctypes as C
    import numpy as np
    from lesseq_b200 import api, _lib
    path = tmp_path / "r.mrf"
    rows = ["# comment", "AlignmentBlocks"]
    rng = np.random.default_rng(3)
    for i in range(3000):
        nb = int(rng.integers(1, 5))
        s = int(rng.integers(1, 10 ** 6))
        blocks = []
        for _ in range(nb):
            ln = int(rng.integers(1, 300))
            chrom = "chr%d" % rng.integers(1, 25) if rng.random() < 0.02 else "chr%d" % (1 + i % 24)
            blocks.append("%s:%s:%d:%d:1:%d" % (chrom, "+-"[i % 2], s, s + ln - 1, ln))
            s += ln + int(rng.integers(0, 5000))
        rows.append(",".join(blocks))
    path.write_text("\n".join(rows) + "\n")
    wide = api.Reads.from_mrf(str(path))
    comp = api.Reads.from_mrf(str(path), compact=True)
    v, w = comp.view, wide.view
    n, nb = comp.n_reads, comp.n_blocks
    assert (n, nb) == (wide.n_reads, wide.n_blocks) and v.blk_cnt and v.read_cs8 and v.blk_start and v.blk_len

    def arr(ptr, ctype, count):
        return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(ctype)), shape=(count,)).copy()
    off = arr(w.blk_off, C.c_uint32, n + 1)
    blk = arr(w.blk, C.c_int32, 2 * nb).reshape(-1, 2)
    cs = arr(w.read_cs, C.c_uint32, n)
    assert np.array_equal(arr(v.blk_cnt, C.c_uint8, n), np.diff(off))
    assert np.array_equal(arr(v.blk_start, C.c_int32, nb), blk[:, 0])
    assert np.array_equal(arr(v.blk_len, C.c_uint16, nb).astype(np.int64), blk[:, 1].astype(np.int64) - blk[:, 0])
    assert not v.blk_len8  # blocks of up to 299 bases do not fit a byte ...
    short = tmp_path / "short.mrf"
    short.write_text("AlignmentBlocks\n" + "\n".join("chr1:+:%d:%d:1:50" % (s, s + 49) for s in range(1, 400, 7)) + "\n")
    sv = api.Reads.from_mrf(str(short), compact=True)  # ... 50-base blocks do
    assert sv.view.blk_len8 and (arr(sv.view.blk_len8, C.c_uint8, sv.n_blocks) == 50).all()
    assert sv.host_bytes() == sv.n_reads * (1 + 1 + 4) + sv.n_blocks * 5
    cs8 = arr(v.read_cs8, C.c_uint8, n).astype(np.uint32)
    assert np.array_equal((cs8 & 0x3F) | ((cs8 & 0x40) << 10) | ((cs8 & 0x80) << 24), cs)
    assert (cs & _lib.LSQ_READ_MIXED).any()
    assert comp.host_bytes() < 0.75 * wide.host_bytes()


def test_read_pack_round_trip_and_staleness(tmp_path):
    """lsqh_reads_save / lsqh_reads_load: the pack reproduces every array of the parsed file
    (chimeric reads, comment lines, empty read lines included); a changed text file, a truncated
    pack and a flipped byte are all refused."""
    import ctypes as C
    import numpy as np
    from lesseq_b200 import api, _lib
    rng = np.random.default_rng(11)
    rows = ["# header comment", "AlignmentBlocks"]
    for i in range(5000):
        if i % 997 == 0:
            rows.append("# comment in the middle")
        nb = int(rng.integers(1, 4))
        s = int(rng.integers(1, 10 ** 7))
        blocks = []
        for _ in range(nb):
            ln = int(rng.integers(1, 120))
            chrom = "chrX" if rng.random() < 0.01 else "chr%d" % (1 + i % 5)
            blocks.append("%s:%s:%d:%d:1:%d" % (chrom, "+-"[i % 2], s, s + ln - 1, ln))
            s += ln + int(rng.integers(0, 3000))
        rows.append(",".join(blocks))
    for mixed in (True, False):
        d = tmp_path / ("m%d" % mixed)
        d.mkdir()
        text = "\n".join(r for r in rows if mixed or "chrX" not in r) + "\n"
        src, pack = d / "r.mrf", d / "r.mrf.lsqpack"
        src.write_text(text)
        a = api.Reads.from_mrf(str(src))
        a.save_pack(str(pack), str(src))
        b = api.Reads.from_pack(str(pack), str(src))

        def arrays(r):
            v, n, nb = r.view, r.n_reads, r.n_blocks

            def arr(ptr, ctype, count):
                return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(ctype)), shape=(count,)).copy() if ptr else None
            names = ([v.chrom_names[i] for i in range(v.n_chrom_names)], [v.strand_names[i] for i in range(v.n_strand_names)])
            return (n, nb, names, arr(v.blk_off, C.c_uint32, n + 1), arr(v.blk, C.c_int32, 2 * nb), arr(v.read_cs, C.c_uint32, n),
                    arr(v.blk_cs, C.c_uint32, nb), arr(v.line_num, C.c_uint32, n))
        xa, xb = arrays(a), arrays(b)
        assert xa[:3] == xb[:3] and xa[0] > 4000
        assert (xa[6] is not None) == mixed
        for u, w in zip(xa[3:], xb[3:]):
            assert (u is None and w is None) or np.array_equal(u, w)
        # compact form of a loaded pack
        c = api.Reads.from_pack(str(pack), None, compact=True)
        assert c.view.blk_cnt and c.view.blk_len8
        # truncated / corrupted packs
        raw = pack.read_bytes()
        for bad in (raw[:len(raw) // 2], raw[:-1], raw[:200] + bytes([raw[200] ^ 1]) + raw[201:], b"LSQPACK0" + raw[8:]):
            (d / "bad.lsqpack").write_bytes(bad)
            with pytest.raises(_lib.LsqError):
                api.Reads.from_pack(str(d / "bad.lsqpack"), None)
        # stale: the text changed after the pack was written
        src.write_text(text + "chr1:+:5:20:1:16\n")
        with pytest.raises(_lib.LsqError):
            api.Reads.from_pack(str(pack), str(src))
        api.Reads.from_pack(str(pack), None)  # without a source stamp check it still loads


def test_other_annotation_formats_replayed_on_cpu_match_reference(emul_bin, tmp_path):
    """UCSC_GENE_TXT / UCSC_GFF / WORMBASE_GFF2 / GENELETS_GFF3 forms and WORMBASE_GENE2ISOFORMS maps
    (solve/solve.cpp:158-329) through the host front end + the CPU replay of the device functions, against
    what the reference's solve binary printed for the same files (tests/golden/formats.json)."""
    cases = golden_util.load("formats.json")
    seen = set()
    for case in cases:
        d = tmp_path / case["name"].replace("|", "_")
        d.mkdir()
        _, sa = golden_util.materialise_formats(case, str(d))
        rc, out = _run_emul(emul_bin, "solve", sa)
        assert rc == 0, case["name"]
        assert _close(case["expect"]["solve17"], out, True), case["name"]
        seen.add((case["forms_format"], case["map_format"]))
    assert len(cases) > 200 and len(seen) == 8


def test_cli_accepts_solve_only_formats_like_the_reference(tmp_path):
    """count.cpp:141-183 knows LH_GENE_TXT / UCSC_GENE2ISOFORM only and answers anything else with exit 1 and
    "Unknown file format error"; solve.cpp:158-329 reads five forms formats and two map formats.  Without a GPU the
    drop-in `solve` gets as far as creating the device context (exit 70: there is no CPU fallback)."""
    from lesseq_b200 import api, build
    build.build()
    cases = golden_util.load("formats.json")
    picked = {}
    for c in cases:
        picked.setdefault((c["forms_format"], c["map_format"]), c)
    assert len(picked) == 8
    for (ff, mf), case in picked.items():
        d = tmp_path / (ff + mf)
        d.mkdir()
        ca, sa = golden_util.materialise_formats(case, str(d))
        ann = api.Annotation(sa[4], sa[6], case["begin"], case["end"], forms_format=ff, map_format=mf)
        assert ann.n_events > 0 and ann.n_forms >= ann.n_events
        rc, out, err = api.run_program("count", ca)
        assert rc == 1 and out == "" and "Unknown file format error" in err
        rc, out, err = api.run_program("solve", sa)
        assert rc in (0, 70), err
        bad = list(sa)
        bad[3] = "UCSC_BED"
        rc, out, err = api.run_program("solve", bad)
        assert rc == 1 and "Unknown file format error: UCSC_BED" in err
    with pytest.raises(api.LsqError):
        api.Annotation(sa[4], sa[6], forms_format="NOT_A_FORMAT")


def test_mrf_parser_reads_pipes_and_reports_lines(tmp_path):
    """The memory-mapped parser falls back to reading when the path is a pipe (`<(zcat reads.mrf.gz)`); parse
    errors name the file's line whatever the number of parser threads; results do not depend on the thread count."""
    import ctypes as C
    import numpy as np
    from lesseq_b200 import api
    lines = ["AlignmentBlocks"] + ["chr%d:%s:%d:%d:1:30" % (1 + i % 7, "+-"[i % 2], 100 + 13 * i, 129 + 13 * i) for i in range(60000)]
    text = "\n".join(lines) + "\n"
    assert len(text) > (1 << 20)  # large enough for the multi-threaded path
    p = tmp_path / "r.mrf"
    p.write_text(text)

    def arrays(r):
        v, n, nb = r.view, r.n_reads, r.n_blocks
        f = lambda ptr, ct, k: np.ctypeslib.as_array(C.cast(ptr, C.POINTER(ct)), shape=(k,)).copy()  # noqa: E731
        return f(v.blk_off, C.c_uint32, n + 1), f(v.blk, C.c_int32, 2 * nb), f(v.read_cs, C.c_uint32, n), f(v.line_num, C.c_uint32, n)
    os.environ["LSQ_THREADS"] = "1"
    one = arrays(api.Reads.from_mrf(str(p)))
    os.environ["LSQ_THREADS"] = "5"
    try:
        many = arrays(api.Reads.from_mrf(str(p)))
        for a, b in zip(one, many):
            assert np.array_equal(a, b)
        assert np.array_equal(one[3], np.arange(1, 60001, dtype=np.uint32))  # the skipped first line is not counted
        fifo = tmp_path / "fifo"
        os.mkfifo(str(fifo))
        w = subprocess.Popen(["bash", "-c", "cat '%s' > '%s'" % (p, fifo)])
        piped = arrays(api.Reads.from_mrf(str(fifo)))
        w.wait()
        for a, b in zip(one, piped):
            assert np.array_equal(a, b)
        bad = tmp_path / "bad.mrf"
        bad.write_text("\n".join(lines[:50001] + ["chr1:+:12x:40:1:30"] + lines[50001:]) + "\n")
        with pytest.raises(api.LsqError) as ei:
            api.Reads.from_mrf(str(bad))
        assert "MRF line 50001)" in str(ei.value)
    finally:
        del os.environ["LSQ_THREADS"]


def test_empty_input_files_behave_like_the_reference(emul_bin, oracle_bin, tmp_path):
    """An EMPTY file passes the reference's open assertion (eof is only set by a read): an empty map selects no
    gene (no output, exit 0), an empty reads file counts nothing, an empty forms file makes a selected gene name a
    missing form (the reference dies on an assertion: exit 134 here).  The oracle was checked against the binaries
    on exactly these inputs."""
    from lesseq_b200 import api
    case = golden_util.load("seed.json")[0]
    ca, sa = golden_util.materialise(case, str(tmp_path))
    empty = tmp_path / "empty"
    empty.write_text("")

    def variant(idx_count, idx_solve):
        c2, s2 = list(ca), list(sa)
        for i in idx_count:
            c2[i] = str(empty)
        for i in idx_solve:
            s2[i] = str(empty)
        return c2, s2
    for name, (c2, s2) in {"map": variant([6], [6]), "forms+map": variant([4, 6], [4, 6]),
                           "reads": variant([len(ca) - 1], [len(sa) - 2])}.items():
        for prog, a in (("count", c2), ("solve", s2)):
            rc1, o1, _ = oracle_bin.run_oracle(prog, a)
            rc2, o2 = _run_emul(emul_bin, prog, a)
            assert rc1 == 0 and rc2 == 0 and _close(o1, o2, prog == "solve"), (name, prog)
            assert (o1 == "") == (name != "reads")
            rc3, o3, err = api.run_program(prog, a, precision=17)
            if name != "reads":
                assert rc3 == 0 and o3 == "", (name, prog, err)  # decided before any device work
            else:
                assert rc3 in (0, 70), err  # 70: this machine has no GPU
                assert rc3 == 70 or _close(o1, o3, prog == "solve")
    c2, s2 = variant([4], [4])  # empty forms, the map still names them
    for prog, a in (("count", c2), ("solve", s2)):
        assert oracle_bin.run_oracle(prog, a)[0] != 0
        rc, out, err = api.run_program(prog, a)
        assert rc == 134 and out == "" and "not in the forms file" in err
    r = api.Reads.from_mrf(str(empty))
    assert r.n_reads == 0 and r.n_blocks == 0


def test_merged_then_appended_then_out_of_order_blocks(emul_bin, oracle_bin, tmp_path):
    """Regression (found by a live fuzz run against the reference binaries): blocks that first overlap (the list
    shrinks to one interval in the ext entry), then append, then arrive out of order --
    [842,892) [888,910) [912,916) [910,912) -- must leave [842,912) [912,916) like interval_list::add_interval."""
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
    rc1, o1, _ = oracle_bin.run_oracle("count", args)
    rc2, o2 = _run_emul(emul_bin, "count", args)
    assert rc1 == 0 and rc2 == 0 and o1 == o2
    assert o1.splitlines() == ["~t\t2\t~t|1\t1", "~t\t2\t~t|2\t2", "~t\t2\t~t|3\t0"]
