"""Pin the CPU oracle (oracle/lesseq_oracle.cpp) to the reference binaries' golden outputs.

The fixtures were produced by /root/reference/bin/{count,solve} (tests/golden/make_golden.py);
the oracle must reproduce them byte for byte, both at the programs' native 6 significant digits
and at 17 digits (i.e. bit-identical doubles), including exit codes.
"""
import pytest

import golden_util

CASES = golden_util.all_cases()


@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_oracle_matches_reference_binary(case, oracle_bin, tmp_path):
    ca, sa = golden_util.materialise(case, str(tmp_path))
    exp = case["expect"]
    for prog, args in (("count", ca), ("solve", sa)):
        for prec in (17, 6):
            rc, out, err = oracle_bin.run_oracle(prog, args, prec=prec)
            assert rc == exp["%s_rc" % prog], err
            assert out == exp["%s%d" % (prog, prec)], "%s prec=%d differs from reference" % (prog, prec)


def test_live_reference_binary_if_present(oracle_bin, tmp_path):
    """Where oracle/_ref exists (dev container, GPU box), re-check a few fresh seeds live."""
    if not oracle_bin.ref_available():
        pytest.skip("oracle/_ref not built")
    import fuzzgen
    checked = 0
    for seed in range(5000, 5040):
        case = fuzzgen.make_case(seed)
        d = tmp_path / ("s%d" % seed)
        d.mkdir()
        ca, sa = fuzzgen.write_case(case, str(d))
        for prog, args in (("count", ca), ("solve", sa)):
            rc1, o1, _ = oracle_bin.run_ref(prog, args)
            if rc1 not in (0, 1):
                continue  # reference crashed on its own undefined behaviour (nil-only read)
            rc2, o2, _ = oracle_bin.run_oracle(prog, args)
            assert (rc1, o1) == (rc2, o2), "seed %d %s" % (seed, prog)
            checked += 1
    assert checked > 40
