#!/bin/bash
# Developer aid: build a variant of libgsr_b200.so with extra -D flags applied to ONE translation unit.
#   build_variant.sh <name> <file.cu> "<-D flags>" [<file2.cu> "<flags2>"]   ->  ../variants/libgsr_<name>.so
# (A/B timing with GSR_B200_LIB=... ; tests/gpu_lib_sweep.py)
set -e
cd "$(dirname "$0")"
name=$1; shift
mkdir -p ../variants /tmp/gsr_variants/$name
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC,-O2 -Xptxas -v --expt-relaxed-constexpr -ccbin /usr/bin/g++"
objs=""
skip=""
while [ $# -gt 0 ]; do
  f=$1; d=$2; shift 2
  o=/tmp/gsr_variants/$name/${f%.cu}.o
  $NVCC $FLAGS $d -dc -c $f -o $o 2> /tmp/gsr_variants/$name/${f%.cu}.log || (cat /tmp/gsr_variants/$name/${f%.cu}.log; exit 1)
  grep -h "registers\|spill" /tmp/gsr_variants/$name/${f%.cu}.log | sort | uniq -c | sed "s/^/[$name $f] /" | head -8
  objs="$objs $o"; skip="$skip ${f%.cu}.o"
done
for o in *.o; do case " $skip " in *" $o "*) ;; *) objs="$objs $o";; esac; done
$NVCC -gencode arch=compute_100a,code=sm_100a -shared -Xcompiler -fPIC -ccbin /usr/bin/g++ $objs -o ../variants/libgsr_$name.so
echo "built ../variants/libgsr_$name.so"
