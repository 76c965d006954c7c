#!/usr/bin/env python
"""Generate tests/golden/*.json from the REFERENCE'S OWN BINARIES (run in the dev container).

The reference ships no tests or golden vectors, so every expected output
here is produced by /root/reference/bin/{count,solve} (made runnable by oracle/build_ref.sh):
stdout at the program's native 6 significant digits and at 17 digits through the LD_PRELOAD
shim oracle/ref_support/prec17.cpp.  The fixtures hold the exact input files, the argument
list shape and the outputs, so that they can be replayed where /root/reference does not exist.

    python tests/golden/make_golden.py            # rewrites seed.json and fuzz.json

Seed cases A1-A6 follow SURVEY.md Appendix A.
"""
import json
import os
import shutil
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import runner  # noqa: E402
import fuzzgen  # noqa: E402


def T(*f):
    return "\t".join(f)


def mrf(lines, header="AlignmentBlocks"):
    return header + "\n" + "\n".join(lines) + "\n"


def blocks(*bl):
    q = 1
    out = []
    for (c, st, s, e) in bl:
        out.append("%s:%s:%d:%d:%d:%d" % (c, st, s, e, q, q + e - s))
        q += e - s + 1
    return ",".join(out)


def seed_cases():
    cases = []
    # A1 -- SE basics
    cases.append({
        "name": "A1_se_basics",
        "forms": T("E1|2|1", "chr1", "+", "100", "700", "3", "100,300,600", "200,400,700") + "\n" +
                 T("E1|2|2", "chr1", "+", "100", "700", "2", "100,600", "200,700") + "\n",
        "map": "1\tE1|2|1\n1\tE1|2|2\n",
        "sets": [{"type": "SHORT_READ", "read_length": 50, "total_read_bases": 450.0, "mrf": mrf([
            blocks(("chr1", "+", 101, 150)),
            blocks(("chr1", "+", 171, 200), ("chr1", "+", 301, 320)),
            blocks(("chr1", "+", 171, 200), ("chr1", "+", 601, 620)),
            blocks(("chr1", "+", 311, 360)),
            blocks(("chr1", "+", 381, 400), ("chr1", "+", 601, 630)),
            blocks(("chr1", "+", 651, 700)),
            blocks(("chr1", "+", 191, 240)),
            blocks(("chr2", "+", 101, 150)),
            blocks(("chr1", "+", 151, 200))])}],
        "begin": 0, "end": 100})
    # A2 -- mixed quirks (truncation 74/75, opposite strand, G=0, lexicographic order)
    a2_forms = "".join(T(*r) + "\n" for r in [
        ("A|1", "chr1", "+", "1000", "2000", "3", "1000,1300,1900", "1200,1400,2000"),
        ("A|2", "chr1", "+", "1000", "2000", "2", "1000,1900", "1200,2000"),
        ("S|1", "chr1", "+", "8000", "8030", "1", "8000", "8030"),
        ("S|2", "chr1", "+", "8000", "8200", "1", "8000", "8200"),
        ("R|1", "chr1", "-", "3000", "3300", "3", "3000,3100,3200", "3100,3200,3300"),
        ("R|2", "chr1", "-", "3000", "3300", "2", "3000,3200", "3100,3300")])
    a2_map = "2\tA|1\n2\tA|2\n10\tS|1\n10\tS|2\n1\tR|1\n1\tR|2\n"
    a2_reads = mrf([
        blocks(("chr1", "-", 1101, 1175)),
        blocks(("chr1", "+", 1127, 1200), ("chr1", "+", 1301, 1301)),
        blocks(("chr1", "+", 1127, 1200), ("chr1", "+", 1901, 1901)),
        blocks(("chr1", "+", 1001, 1075)),
        blocks(("chr1", "+", 1001, 2000)),
        blocks(("chr1", "+", 8001, 8030)),
        blocks(("chr1", "+", 8002, 8076)),
        blocks(("chr1", "+", 8101, 8175)),
        blocks(("chr1", "+", 3051, 3125)),
        blocks(("chr1", "+", 3051, 3100), ("chr1", "+", 3201, 3225)),
        blocks(("chr1", "+", 3126, 3200)),
        blocks(("chr1", "+", 3126, 3199), ("chr1", "+", 3201, 3201))])
    cases.append({"name": "A2_mixed_quirks", "forms": a2_forms, "map": a2_map,
                  "sets": [{"type": "SHORT_READ", "read_length": 75, "total_read_bases": 900.0,
                            "mrf": a2_reads}], "begin": 0, "end": 100})
    cases.append({"name": "A2_two_sets_range_1_2", "forms": a2_forms, "map": a2_map,
                  "sets": [{"type": "SHORT_READ", "read_length": 75, "total_read_bases": 900.0, "mrf": a2_reads},
                           {"type": "MEDIUM_READ", "read_length": 75, "total_read_bases": 900.0, "mrf": a2_reads}],
                  "begin": 1, "end": 2})
    # A3 -- tie-break on (strand, name)
    a3_forms = "".join(T(*r) + "\n" for r in [
        ("B|1", "chr1", "+", "5000", "5050", "1", "5000", "5050"),
        ("B|2", "chr1", "+", "5000", "5050", "1", "5000", "5050"),
        ("C|1", "chr1", "-", "6000", "6050", "1", "6000", "6050"),
        ("C|2", "chr1", "-", "6000", "6050", "1", "6000", "6050")])
    a3_reads = mrf([
        blocks(("chr1", "+", 5001, 5050)), blocks(("chr1", "-", 5001, 5050)),
        blocks(("chr1", "+", 6001, 6050)), blocks(("chr1", "-", 6001, 6050)),
        blocks(("chr1", "-", 6002, 6050))])
    for nm, mp in (("A3_tiebreak", "z\tB|1\nz\tB|2\n7\tC|1\n7\tC|2\n"),
                   ("A3_tiebreak_no_trailing_newline", "z\tB|1\nz\tB|2\n7\tC|1\n7\tC|2"),
                   ("A3_tiebreak_readname_gene", "read-2\tB|1\nread-2\tB|2\nread-30\tC|1\nread-30\tC|2\n")):
        cases.append({"name": nm, "forms": a3_forms, "map": mp,
                      "sets": [{"type": "SHORT_READ", "read_length": 50, "total_read_bases": 250.0,
                                "mrf": a3_reads}], "begin": 0, "end": 100})
    # A4 -- ARS lengths / non-finite flow
    a4_forms = (T("A|1", "chr1", "+", "1000", "2000", "3", "1000,1300,1900", "1100,1400,2000") + "\n" +
                T("B|1", "chr1", "+", "5000", "5030", "2", "5000,5100", "5020,5110") + "\n")
    a4_reads = mrf([blocks(("chr1", "+", 1011, 1060)), blocks(("chr1", "+", 5002, 5020))])
    for rt in ("SHORT_READ", "MEDIUM_READ"):
        cases.append({"name": "A4_ars_%s" % rt.lower(), "forms": a4_forms, "map": "1\tA|1\n2\tB|1\n",
                      "sets": [{"type": rt, "read_length": 50, "total_read_bases": 1000.0, "mrf": a4_reads}],
                      "begin": 0, "end": 100})
    # A5 -- range dependence of the load filter
    a5_forms = (T("A|1", "chr1", "+", "1000", "1100", "1", "1000", "1100") + "\n" +
                T("B|1", "chr1", "+", "2000", "2100", "1", "2000", "2100") + "\n")
    a5_reads = mrf([blocks(("chr1", "+", 1051, 1100), ("chr1", "+", 2001, 2025))])
    for (b, e) in ((0, 2), (0, 1), (1, 2), (0, 0)):
        cases.append({"name": "A5_range_%d_%d" % (b, e), "forms": a5_forms, "map": "1\tA|1\n2\tB|1\n",
                      "sets": [{"type": "SHORT_READ", "read_length": 75, "total_read_bases": 75.0,
                                "mrf": a5_reads}], "begin": b, "end": e})
    # A6 -- >= 1e6 formatting (MRF stored as a repeat recipe: [line, copies] pairs)
    cases.append({"name": "A6_million_format",
                  "forms": T("Z|1", "chr1", "+", "9000", "9200", "1", "9000", "9200") + "\n" +
                           T("Z|2", "chr1", "+", "9000", "9100", "1", "9000", "9100") + "\n",
                  "map": "4\tZ|1\n4\tZ|2\n",
                  "sets": [{"type": "SHORT_READ", "read_length": 50, "total_read_bases": 61728700.0,
                            "mrf_repeat": [[blocks(("chr1", "+", 9011, 9060)), 1234567],
                                           [blocks(("chr1", "+", 9111, 9160)), 7]]}],
                  "begin": 0, "end": 100})
    # Order-dependent merging of abutting exons in covered regions (interval_list.hpp:462-503):
    # gene "1" contributes [200,300) before gene "2" contributes [100,200) -> two abutting
    # covered intervals, so a read across position 200 is dropped at load; in the other
    # lexicographic order they fuse and the read survives.
    # The two-block read keeps its second block [195,205) only when the intervals fused; it then
    # fails the 0.98 rule (100/105) instead of being counted for Q.
    ab_reads = mrf([blocks(("chr1", "+", 151, 250)), blocks(("chr1", "+", 101, 200), ("chr1", "+", 196, 205)),
                    blocks(("chr1", "+", 201, 300))])
    for nm, g1, g2 in (("abut_unmerged", "1", "2"), ("abut_merged", "2", "1")):
        cases.append({"name": "covered_%s" % nm,
                      "forms": T("P|1", "chr1", "+", "200", "300", "1", "200", "300") + "\n" +
                               T("Q|1", "chr1", "+", "100", "200", "1", "100", "200") + "\n",
                      "map": "%s\tP|1\n%s\tQ|1\n" % (g1, g2),
                      "sets": [{"type": "SHORT_READ", "read_length": 100, "total_read_bases": 300.0,
                                "mrf": ab_reads}], "begin": 0, "end": 100})
    return cases


def materialise(case, d):
    c = dict(case)
    c["sets"] = []
    for s in case["sets"]:
        s2 = dict(s)
        if "mrf_repeat" in s:
            s2["mrf"] = "AlignmentBlocks\n" + "".join((ln + "\n") * n for ln, n in s["mrf_repeat"])
        c["sets"].append(s2)
    return fuzzgen.write_case(c, d)


def run_case(case):
    d = tempfile.mkdtemp(prefix="golden_")
    try:
        ca, sa = materialise(case, d)
        exp = {}
        for prog, args in (("count", ca), ("solve", sa)):
            for tag, p17 in (("17", True), ("6", False)):
                rc, out, _ = runner.run_ref(prog, args, prec17=p17)
                exp["%s%s" % (prog, tag)] = out
                exp["%s_rc" % prog] = rc
        return exp
    finally:
        shutil.rmtree(d, ignore_errors=True)


def main():
    if not runner.ref_available() and not runner.build_ref():
        sys.exit("reference binaries unavailable (need /root/reference)")
    seeds = seed_cases()
    for c in seeds:
        c["expect"] = run_case(c)
        print(c["name"], "rc", c["expect"]["count_rc"], c["expect"]["solve_rc"])
    json.dump(seeds, open(os.path.join(HERE, "seed.json"), "w"), indent=1)
    fuzz = []
    seed = 0
    while len(fuzz) < 80:
        c = fuzzgen.make_case(seed)
        c["name"] = "fuzz_%d" % seed
        seed += 1
        exp = run_case(c)
        if exp["count_rc"] not in (0, 1) or exp["solve_rc"] not in (0, 1):
            continue  # the reference itself crashed (undefined behaviour on nil-only reads)
        c["expect"] = exp
        fuzz.append(c)
    json.dump(fuzz, open(os.path.join(HERE, "fuzz.json"), "w"), indent=0)
    # annotation shapes outside what Events.r emits: exons out of order, zero-length, duplicated,
    # overlapping inside one form (ExonSet::insert / build_isoform_array corner cases)
    weird = []
    seed = 20000
    while len(weird) < 40:
        c = fuzzgen.make_case(seed, weird=True)
        c["name"] = "weird_%d" % seed
        seed += 1
        exp = run_case(c)
        if exp["count_rc"] not in (0, 1) or exp["solve_rc"] not in (0, 1):
            continue
        c["expect"] = exp
        weird.append(c)
    json.dump(weird, open(os.path.join(HERE, "weird.json"), "w"), indent=0)
    print("wrote", len(seeds), "seed,", len(fuzz), "fuzz and", len(weird), "weird-annotation cases")


if __name__ == "__main__":
    main()
