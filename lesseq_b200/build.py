"""In-tree build of the CUDA library and the drop-in command line programs (sm_100a only).

    python -m lesseq_b200.build          # builds lib/liblesseq_b200.so, bin/count, bin/solve

nvcc cross-compiles without a GPU; the artefacts are git-ignored but travel with the gpurun
snapshot.  Rebuilds only when a source is newer than the artefact.
"""
import glob
import os
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
LIB_DIR = os.path.join(PKG, "lib")
BIN_DIR = os.path.join(PKG, "bin")
LIB = os.path.join(LIB_DIR, "liblesseq_b200.so")

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC,-Wall,-Wno-unused-function,-pthread"]


def _sources():
    return ([os.path.join(CSRC, "lsq_abi.cu")] +
            [os.path.join(CSRC, "host", f) for f in ("annotation.cpp", "mrf.cpp", "pack.cpp", "api.cpp")])


def _deps():
    return (glob.glob(os.path.join(CSRC, "*.cu*")) + glob.glob(os.path.join(CSRC, "*.hpp")) +
            glob.glob(os.path.join(CSRC, "host", "*")) +
            [os.path.join(os.path.dirname(PKG), "include", "lesseq_b200.h"), os.path.abspath(__file__)])


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def _run(cmd, verbose):
    if verbose:
        print(" ".join(cmd), flush=True)
    p = subprocess.run(cmd, capture_output=True, text=True)
    if p.returncode != 0:
        sys.stderr.write(p.stdout + p.stderr)
        raise RuntimeError("build failed: " + " ".join(cmd))
    if verbose and (p.stdout or p.stderr):
        print(p.stdout + p.stderr)


def build(force=False, verbose=False, ptxas_info=False):
    os.makedirs(LIB_DIR, exist_ok=True)
    os.makedirs(BIN_DIR, exist_ok=True)
    deps = _deps()
    if force or _stale(LIB, deps):
        cmd = [NVCC] + ARCH + COMMON + ["-shared", "-o", LIB] + _sources()
        if ptxas_info:
            cmd += ["-Xptxas", "-v"]
        _run(cmd, verbose)
    for prog in ("count", "solve"):
        out = os.path.join(BIN_DIR, prog)
        if force or _stale(out, deps + [LIB]):
            _run([NVCC, "-O2", "-std=c++17", "-o", out, os.path.join(CSRC, "host", prog + "_main.cpp"),
                  "-L" + LIB_DIR, "-llesseq_b200", "-Xlinker", "-rpath=$ORIGIN/../lib"], verbose)
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose=True, ptxas_info="--ptxas" in sys.argv)
    print("built", LIB)
