#!/usr/bin/env python
"""tests/golden/em_fragile_large.json: the em_fragile cases with every read line repeated so that the events hold some
10^5 reads.  Scaling all class counts leaves the EM's trajectory (and its distance from the stop threshold) nearly
unchanged while the summation-order noise grows with the read count, so the stop stays fragile and the product has to
replay events that are LARGE: the device-wide sort, the classify kernel and ordered_sum of lsq_exact_kernel.  Outputs
are those of the reference's own binaries; the read text is stored run-length encoded (mrf_repeat).

    python tests/golden/make_golden_em_large.py
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import golden_util  # noqa: E402
from oracle import runner  # noqa: E402

TARGET_READS = 300000


def main():
    if not runner.ref_available() and not runner.build_ref():
        sys.exit("reference binaries unavailable (need /root/reference)")
    import tempfile
    out = []
    for case in golden_util.load("em_fragile.json"):
        c = {k: v for k, v in case.items() if k != "expect"}
        c["name"] = case["name"] + "_large"
        c["sets"] = []
        for s in case["sets"]:
            lines = [ln for ln in s["mrf"].split("\n")[1:] if ln.strip()]
            rep = max(2, TARGET_READS // max(1, len(lines)))
            s2 = {k: v for k, v in s.items() if k != "mrf"}
            s2["mrf_repeat"] = [[ln, rep] for ln in lines]
            s2["total_read_bases"] = float(s["total_read_bases"]) * rep
            c["sets"].append(s2)
        d = tempfile.mkdtemp(prefix="golden_large_")
        ca, sa = golden_util.materialise(c, d)
        exp = {}
        for prog, args in (("count", ca), ("solve", sa)):
            rc, o, _ = runner.run_ref(prog, args, prec17=True)
            exp["%s17" % prog] = o
            exp["%s_rc" % prog] = rc
        assert exp["count_rc"] == 0 and exp["solve_rc"] == 0
        c["expect"] = exp
        out.append(c)
        print(c["name"], [sum(n for _, n in s["mrf_repeat"]) for s in c["sets"]])
    json.dump(out, open(os.path.join(HERE, "em_fragile_large.json"), "w"), indent=0)


if __name__ == "__main__":
    main()
