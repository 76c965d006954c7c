"""Helpers to replay tests/golden/*.json fixtures (inputs + reference-binary outputs)."""
import json
import os

import fuzzgen

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    return json.load(open(os.path.join(GOLDEN, name)))


def materialise(case, d):
    """Write the case's input files into directory d; returns (count_args, solve_args)."""
    c = dict(case)
    c["sets"] = []
    for s in case["sets"]:
        s2 = dict(s)
        if "mrf_repeat" in s:
            s2["mrf"] = "AlignmentBlocks\n" + "".join((ln + "\n") * n for ln, n in s["mrf_repeat"])
        c["sets"].append(s2)
    return fuzzgen.write_case(c, d)


def all_cases():
    return load("seed.json") + load("fuzz.json") + load("weird.json") + load("em_fragile.json") + load("em_orderfree.json")


def materialise_formats(case, d):
    """Write a formats.json case (annotation in one of solve's other formats); returns (count_args, solve_args)."""
    fp, mp = os.path.join(d, "ev.forms"), os.path.join(d, "ev.map")
    open(fp, "w").write(case["forms"])
    open(mp, "w").write(case["map"])
    head = ["0", "p", os.path.join(d, "o_"), case["forms_format"], fp, case["map_format"], mp,
            str(case["begin"]), str(case["end"])]
    ca, sa = list(head), list(head)
    for i, s in enumerate(case["sets"]):
        rp = os.path.join(d, "r%d.mrf" % i)
        open(rp, "w").write(s["mrf"])
        ca += ["MRF_SINGLE", s["type"], str(s["read_length"]), rp]
        sa += ["MRF_SINGLE", s["type"], str(s["read_length"]), rp, repr(s["total_read_bases"])]
    return ca, sa
