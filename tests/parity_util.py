"""Comparison helpers for the parity tests.

Tolerances (BASELINE.json north_star): per-form counts, n_valid, bases and compatibility classes
BIT-EXACT; theta, rpkm and log-likelihood within 1e-9 RELATIVE.  A value whose magnitude is below
1e-12 is accepted on an absolute 1e-12 floor: it only occurs for log-likelihoods that are exactly
0 up to the rounding noise of the reference's sequential sum (all G_k == 1), where a relative
criterion is meaningless.
"""
import math

REL = 1e-9
ABS_FLOOR = 1e-12


def num_close(u, v):
    if u == v:
        return True
    if math.isnan(u) and math.isnan(v):
        return True
    if math.isinf(u) or math.isinf(v):
        return False
    return abs(u - v) <= REL * max(abs(u), abs(v)) or abs(u - v) <= ABS_FLOOR


def tables_match(a, b, solve, rel6=False):
    """a, b: stdout tables.  Count tables must be identical strings.  Solve tables: identical
    string columns up to the form name, numeric columns within tolerance (6-digit output is
    compared at 2e-6 relative because the last printed digit may round differently)."""
    la, lb = a.splitlines(), b.splitlines()
    if len(la) != len(lb):
        return "line count %d != %d" % (len(la), len(lb))
    for ln, (x, y) in enumerate(zip(la, lb)):
        if x == y:
            continue
        if not solve:
            return "line %d: %r != %r" % (ln, x, y)
        fx, fy = x.split("\t"), y.split("\t")
        if len(fx) != len(fy) or fx[:-3] != fy[:-3]:
            return "line %d: %r != %r" % (ln, x, y)
        for u, v in zip(fx[-3:], fy[-3:]):
            if u == v:
                continue
            du, dv = float(u), float(v)
            ok = num_close(du, dv) or (rel6 and abs(du - dv) <= 2e-6 * max(abs(du), abs(dv)))
            if not ok:
                return "line %d: %s vs %s" % (ln, u, v)
    return None
