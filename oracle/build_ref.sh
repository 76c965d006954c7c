#!/bin/bash
# Assemble oracle/_ref: the reference's OWN prebuilt count/solve binaries made runnable here.
#
# The reference cannot be compiled in this image (Boost 1.34.1 headers are not vendored and
# there is no network; libgsl is a missing blob -- SURVEY.md section 8c), but its prebuilt
# x86-64 binaries run once libgsl.so.0 resolves.  They import four gsl_ran_* symbols that are
# never called on the count/solve path, so a stub library satisfies the loader.  Nothing from
# the reference's SOURCES is copied; oracle/_ref/ is git-ignored and travels to the GPU box
# with the gpurun snapshot.
#
#   _ref/count, _ref/solve      copies of /root/reference/bin/{count,solve}
#   _ref/lib/libgslcblas.so.0   copy of /root/reference/gsl/lib/libgslcblas.so.0
#   _ref/lib/libgsl.so.0        stub built from ref_support/gsl_stub.c
#   _ref/lib/libprec17.so       LD_PRELOAD shim (ref_support/prec17.cpp): 17-digit doubles
set -euo pipefail
here="$(cd "$(dirname "$0")" && pwd)"
ref="${LSQ_REFERENCE_ROOT:-/root/reference}"
out="$here/_ref"
mkdir -p "$out/lib"
if [ -d "$ref/bin" ]; then
    cp -f "$ref/bin/count" "$ref/bin/solve" "$out/"
    cp -f "$ref/gsl/lib/libgslcblas.so.0" "$out/lib/"
    chmod u+w,a+x "$out/count" "$out/solve"
elif [ ! -x "$out/count" ]; then
    echo "build_ref.sh: $ref absent and no prebuilt oracle/_ref -- reference arm unavailable" >&2
    exit 3
fi
gcc -shared -fPIC -O2 -o "$out/lib/libgsl.so.0" "$here/ref_support/gsl_stub.c"
g++ -shared -fPIC -O2 -o "$out/lib/libprec17.so" "$here/ref_support/prec17.cpp" -ldl
echo "oracle/_ref ready"
