#!/bin/sh
# Builds tuning variants of libfsb200.so into foldseek_b200/variants/ (git-ignored *.so) for tools/perf_pf.py (FSB200_LIB=...).
set -e
cd "$(dirname "$0")/.."
mkdir -p foldseek_b200/variants
rm -f foldseek_b200/variants/*.so
build() {
  name=$1; shift
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 --extended-lambda "$@" \
     -Xcompiler -fPIC,-fopenmp,-O3,-mfma -shared -o foldseek_b200/variants/lib_$name.so foldseek_b200/csrc/fsb200.cu -lgomp 2>&1 | grep -i "error" || true
}
for spec in "$@"; do
  name=${spec%%:*}; flags=$(echo "${spec#*:}" | tr ',' ' ')
  build "$name" $flags &
done
wait
ls foldseek_b200/variants/
