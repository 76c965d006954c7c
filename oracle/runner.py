"""Runners for the two CPU checkers (TEST INFRASTRUCTURE -- never imported by lesseq_b200/).

* ``run_ref``    -- the reference's own prebuilt ``count``/``solve`` binaries from ``oracle/_ref``
                   (assembled by ``oracle/build_ref.sh``); 17-digit output via the LD_PRELOAD shim.
* ``run_oracle`` -- the C++ restatement ``oracle/_build/lesseq_oracle`` (``make -C oracle``).

Both take the reference's positional argument list *after* the program name
(count/count.cpp:95-129, solve/solve.cpp:109-146) and return (returncode, stdout, stderr).
"""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")
ORACLE_BIN = os.path.join(HERE, "_build", "lesseq_oracle")


def ref_available():
    return all(os.path.exists(os.path.join(REF_DIR, p))
               for p in ("count", "solve", "lib/libgsl.so.0", "lib/libgslcblas.so.0", "lib/libprec17.so"))


def oracle_available():
    return os.path.exists(ORACLE_BIN)


def build_oracle():
    subprocess.run(["make", "-C", HERE, "-s"], check=True)


def build_ref():
    return subprocess.run([os.path.join(HERE, "build_ref.sh")], capture_output=True, text=True).returncode == 0


def run_ref(prog, args, prec17=True, cwd=None, timeout=None):
    assert prog in ("count", "solve")
    env = dict(os.environ)
    lib = os.path.join(REF_DIR, "lib")
    env["LD_LIBRARY_PATH"] = lib + (":" + env["LD_LIBRARY_PATH"] if env.get("LD_LIBRARY_PATH") else "")
    if prec17:
        env["LD_PRELOAD"] = os.path.join(lib, "libprec17.so")
    p = subprocess.run([os.path.join(REF_DIR, prog)] + [str(a) for a in args], env=env, cwd=cwd,
                       capture_output=True, text=True, timeout=timeout)
    return p.returncode, p.stdout, p.stderr


def run_oracle(prog, args, prec=17, cwd=None, dump=None, timeout=None):
    assert prog in ("count", "solve")
    env = dict(os.environ)
    env["LSQ_ORACLE_PRECISION"] = str(prec)
    if dump:
        env["LSQ_ORACLE_DUMP"] = dump
    p = subprocess.run([ORACLE_BIN, prog] + [str(a) for a in args], env=env, cwd=cwd,
                       capture_output=True, text=True, timeout=timeout)
    return p.returncode, p.stdout, p.stderr


def parse_table(stdout, n_methods, solve):
    """Parse count/solve stdout into a list of rows.

    count row: (gname, [n_valid_m...], form, count)
    solve row: (gname, [n_valid_m...], form, theta, rpkm, ll_per_read)
    """
    rows = []
    for line in stdout.splitlines():
        if not line:
            continue
        f = line.split("\t")
        g = f[0]
        nv = [float(x) for x in f[1:1 + n_methods]]
        form = f[1 + n_methods]
        vals = [float(x) for x in f[2 + n_methods:]]
        assert len(vals) == (3 if solve else 1), line
        rows.append((g, nv, form) + tuple(vals))
    return rows
