#!/bin/bash
# Developer helper: build a variant of libgsr_b200.so with extra -D flags (or from another git revision) into
# gs-sr_b200/variants/ for A/B runs with tests/gpu_lib_sweep.py (GSR_B200_LIB).  Not part of the product build.
#   build_variant.sh NAME [-DFOO=1 ...]            current sources + flags
#   GSR_REV=HEAD build_variant.sh NAME [...]       sources of a git revision
set -euo pipefail
here="$(cd "$(dirname "$0")" && pwd)"
name="$1"; shift
tmp="$(mktemp -d)"
trap 'rm -rf "$tmp"' EXIT
mkdir -p "$tmp/gs-sr_b200/csrc" "$tmp/include" "$here/../variants"
if [ -n "${GSR_REV:-}" ]; then
    (cd "$here/../.." && git archive "$GSR_REV" gs-sr_b200/csrc include) | tar -x -C "$tmp"
else
    cp "$here"/*.cu "$here"/*.cuh "$here"/Makefile "$tmp/gs-sr_b200/csrc/"
    cp "$here"/../../include/*.h "$tmp/include/"
fi
make -C "$tmp/gs-sr_b200/csrc" -j8 NVCCFLAGS="-O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC,-O2 -Xptxas -v --expt-relaxed-constexpr -ccbin /usr/bin/g++ $*" >/dev/null
cp "$tmp/gs-sr_b200/libgsr_b200.so" "$here/../variants/libgsr_$name.so"
echo "built variants/libgsr_$name.so ($*)"
