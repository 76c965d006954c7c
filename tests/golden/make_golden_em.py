#!/usr/bin/env python
"""Golden cases for the EM stop rule (tests/golden/em_fragile.json): fuzz seeds on which a class-collapsed EM stops at
a different iteration than the reference's read-by-read EM (common/read.h:638-660), found by the live differential
soak (tests/soak/live_fuzz.py) before lsq_exact_kernel existed.  Outputs are those of the reference's own binaries.

    python tests/golden/make_golden_em.py
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import fuzzgen  # noqa: E402
from make_golden import run_case  # noqa: E402
from oracle import runner  # noqa: E402

SEEDS = (229576, 246843, 407193)


def main():
    if not runner.ref_available() and not runner.build_ref():
        sys.exit("reference binaries unavailable (need /root/reference)")
    out = []
    for seed in SEEDS:
        c = fuzzgen.make_case(seed)
        c["name"] = "emfragile_%d" % seed
        c["expect"] = run_case(c)
        assert c["expect"]["solve_rc"] == 0
        out.append(c)
    json.dump(out, open(os.path.join(HERE, "em_fragile.json"), "w"), indent=0)
    print("wrote", len(out), "cases")


if __name__ == "__main__":
    main()
