#!/usr/bin/env python
"""Build tuning variants of liblesseq_b200.so and print the one gpurun command that benches them all.

A GPU call costs minutes of a small budget, so kernel parameters are swept in ONE call: every variant is a
separate shared library built here (nvcc cross-compiles without a GPU) from the same sources with different
-DLSQ_* flags, and bench.py picks the library through LESSEQ_B200_LIB.

    python tools/variant_sweep.py FILT_RBATCH=6 "FILT_BLOCK=256 CTR_PACKED=0" FILT_LBATCH=3
    # -> lesseq_b200/lib/var_*.so (git-ignored, they travel with the gpurun snapshot) + the command to run

Hooks that exist today (lsq_kernels.cuh): LSQ_SCAN_BLOCK, LSQ_SCAN_MINB, LSQ_SCAN_FLIGHT;
host side (environment, no rebuild): LSQ_BKT_PER_PIECE (piece-lookup buckets per piece, default 8; 16 gives 128-bp buckets
with four offsets per word at C2), LSQ_SCAN_FLAGS (1: no single-block signatures, 2: no pair signatures -- debugging).  Remove the var_*.so files before committing a round.
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lesseq_b200 import build  # noqa: E402


def main(specs):
    if not specs:
        print(__doc__)
        return 1
    names = []
    for spec in specs:
        flags = ["-DLSQ_" + kv for kv in spec.split()]
        name = "var_" + "_".join(kv.replace("=", "") for kv in spec.split())
        out = os.path.join(build.LIB_DIR, name + ".so")
        cmd = [build.NVCC] + build.ARCH + build.COMMON + flags + ["-shared", "-o", out] + build._sources()
        print("building", name, flush=True)
        p = subprocess.run(cmd + ["-Xptxas", "-v"], capture_output=True, text=True)
        if p.returncode != 0:
            sys.stderr.write(p.stdout + p.stderr)
            return 1
        for line in (p.stdout + p.stderr).splitlines():  # registers / spills of the kernels that matter
            if "Compiling entry function" in line and ("scan" in line or "general" in line):
                print("  ", line.split("'")[1][:60], end=": ")
            elif "Used" in line and "registers" in line:
                print(line.strip()[len("ptxas info    : "):], flush=True)
        names.append(name)
    loop = " ".join(['""'] + names)
    print("\ngpurun --timeout 900 -- 'for v in %s; do LESSEQ_B200_LIB=${v:+$PWD/lesseq_b200/lib/$v.so} python bench.py --steps 10 "
          "--warmup 3 --no-cpu-baseline > gpurun_out/sweep_${v:-base}.json 2> gpurun_out/sweep.err; echo rc=$?; done'" % loop)
    return 0


if __name__ == "__main__":
    sys.exit(main(sys.argv[1:]))
