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


def test_oracle_standard_errors_match_a_numpy_restatement(oracle_bin, tmp_path):
    """The oracle's S records restate the reference's dead code for standard errors -- FIM::ofim (common/fim.h:322-367,
    log domain with a sign matrix) summed over an event's valid reads, linalg::invert_matrix (linalg.h:28-71, LU with
    partial pivoting) and the variance decomposition of fim.h:75-93.  No reference program evaluates them, so there is
    no reference output to pin them to; this holds the restatement against an independent one: the same information
    matrix from the dumped class histogram in the linear domain, inverted by numpy, on multi-form events (2-9 forms)."""
    import numpy as np
    from lesseq_b200 import api, synth
    A, reads = synth.make_multiform_case(150, 200, 50, 77, k_range=(2, 9))
    fp, mp, rp = str(tmp_path / "ev.interval"), str(tmp_path / "ev.map"), str(tmp_path / "r.mrf")
    A.write(fp, mp)
    api.Reads.from_numpy(reads["blk_off"], reads["blk"], reads["read_cs"], reads["line_num"], reads["chrom_names"],
                         reads["strand_names"]).write_mrf(rp)
    sa = ["0", "p", str(tmp_path / "o_"), "LH_GENE_TXT", fp, "UCSC_GENE2ISOFORM", mp, "0", "100000000", "MRF_SINGLE",
          "MEDIUM_READ", "50", rp, repr(reads["total_read_bases"])]
    dump = str(tmp_path / "dump.txt")
    rc, out, err = oracle_bin.run_oracle("solve", sa, dump=dump)
    assert rc == 0, err
    theta = {}
    for row in out.splitlines():
        f = row.split("\t")
        theta.setdefault(f[0], []).append(float(f[-3]))
    classes, G, se = {}, {}, {}
    for line in open(dump):
        f = line.rstrip("\n").split("\t")
        if f[0] == "C":
            classes.setdefault(f[1], []).append((int(f[3]), int(f[4])))
        elif f[0] == "G":
            G.setdefault(f[1], {})[int(f[3])] = float(f[4])
        elif f[0] == "S":
            se.setdefault(f[1], {})[int(f[2])] = float(f[3])
    checked = 0
    for g, th in theta.items():
        K = len(th)
        if K < 2 or g not in classes:
            continue
        v_all = np.array([G[g][k] for k in range(K)])
        D = K - 1
        info = np.zeros((D, D))
        ok = True
        for mk, c in classes[g]:
            v = np.where([(mk >> k) & 1 for k in range(K)], v_all, 0.0)
            s = float(np.dot(np.array(th) * (v > 0), v))
            if not s > 0:
                ok = False
                break
            d = (v[:D] - v[D]) / s
            info += c * np.outer(d, d)
        got = np.array([se[g][k] for k in range(K)])
        if not ok or np.linalg.cond(info) > 1e8:
            continue
        cov = np.linalg.inv(info)
        exp = np.sqrt(np.concatenate([np.diag(cov), [cov.sum()]]))
        assert np.allclose(got, exp, rtol=1e-6, atol=0), (g, K, got, exp)
        checked += 1
    assert checked >= 60
