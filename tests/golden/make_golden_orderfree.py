#!/usr/bin/env python
"""tests/golden/em_orderfree.json: events whose EM stop looks fragile by value (the log-likelihood becomes exactly 0 and
the stop test divides by zero) but comes out the same in any summation order -- forms not longer than a read, so that
G = 1 or G = 0 and theta reaches a vertex (or stays uniform) in one step.  The product certifies such runs
(em_result_is_order_free) instead of replaying them read by read.  Outputs are those of the reference's own binaries.

    python tests/golden/make_golden_orderfree.py
"""
import json
import os
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import golden_util  # noqa: E402
from oracle import runner  # noqa: E402


def form(name, chrom, strand, exons):
    return "%s\t%s\t%s\t%d\t%d\t%d\t%s\t%s\n" % (name, chrom, strand, exons[0][0], exons[-1][1], len(exons),
                                                 ",".join(str(a) for a, _ in exons), ",".join(str(b) for _, b in exons))


def read(chrom, strand, blocks):
    out, pos = [], 1
    for a, b in blocks:  # 0-based half-open -> 1-based closed
        out.append("%s:%s:%d:%d:%d:%d" % (chrom, strand, a + 1, b, pos, pos + (b - a) - 1))
        pos += b - a
    return ",".join(out)


def cases():
    R = 75
    out = []
    # (1) one form exactly a read long (G = 1), the other shorter than a read (G = 0): theta -> (1, 0) in one step,
    #     the log-likelihood becomes 0, the next test is 0 / 0
    forms = form("V|T3|1", "chr1", "+", [(1000, 1075)]) + form("V|T3|2", "chr1", "+", [(1000, 1060)])
    forms += form("W|T3|1", "chr2", "-", [(5000, 5075)]) + form("W|T3|2", "chr2", "-", [(5010, 5075)])
    # (2) both forms exactly a read long, reads that fit both: theta stays uniform, every read's probability is 1
    forms += form("U|MXE|1", "chr1", "+", [(3000, 3040), (3200, 3235)]) + form("U|MXE|2", "chr1", "+", [(3000, 3040), (3400, 3435)])
    # (3) an ordinary event next to them
    forms += form("N|SE|1", "chr3", "+", [(100, 300), (500, 650), (900, 1100)]) + form("N|SE|2", "chr3", "+", [(100, 300), (900, 1100)])
    gmap = "".join("%s\t%s\n" % (g, f) for g, f in (("g1", "V|T3|1"), ("g1", "V|T3|2"), ("g2", "W|T3|1"), ("g2", "W|T3|2"),
                                                      ("g3", "U|MXE|1"), ("g3", "U|MXE|2"), ("g4", "N|SE|1"), ("g4", "N|SE|2")))
    for scale in (1, 37, 2000):
        lines = [(read("chr1", "+", [(1000, 1075)]), 9 * scale), (read("chr1", "+", [(1000, 1060)]), 2 * scale),
                 (read("chr2", "-", [(5000, 5075)]), 5 * scale), (read("chr2", "-", [(5010, 5075)]), 1 * scale),
                 (read("chr1", "+", [(3000, 3040)]), 7 * scale), (read("chr1", "+", [(3005, 3040)]), 3 * scale),
                 (read("chr3", "+", [(150, 225)]), 11 * scale), (read("chr3", "+", [(260, 300), (500, 535)]), 6 * scale),
                 (read("chr3", "+", [(270, 300), (900, 945)]), 4 * scale), (read("chr3", "+", [(560, 635)]), 5 * scale)]
        total = float(sum(n for _, n in lines) * R)
        out.append({"name": "orderfree_x%d" % scale, "forms": forms, "map": gmap, "begin": 0, "end": 100,
                    "sets": [{"type": "SHORT_READ", "read_length": R, "total_read_bases": total, "mrf_repeat": [[ln, n] for ln, n in lines]}]})
    return out


def main():
    if not runner.ref_available() and not runner.build_ref():
        sys.exit("reference binaries unavailable (need /root/reference)")
    out = []
    for c in cases():
        d = tempfile.mkdtemp(prefix="golden_of_")
        ca, sa = golden_util.materialise(c, d)
        exp = {}
        for prog, args in (("count", ca), ("solve", sa)):
            for tag, p17 in (("17", True), ("6", False)):
                rc, o, _ = runner.run_ref(prog, args, prec17=p17)
                exp["%s%s" % (prog, tag)] = o
                exp["%s_rc" % prog] = rc
        assert exp["count_rc"] == 0 and exp["solve_rc"] == 0, exp
        c["expect"] = exp
        out.append(c)
        print(c["name"])
        print(exp["solve17"])
    json.dump(out, open(os.path.join(HERE, "em_orderfree.json"), "w"), indent=0)


if __name__ == "__main__":
    main()
