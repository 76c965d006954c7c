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
    return load("seed.json") + load("fuzz.json") + load("weird.json")
