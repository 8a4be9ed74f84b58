#!/bin/bash
# Build alternative libraries with different attention compile-time switches and time them (A/B on one box).
# usage (on the GPU box): bash scripts/attn_variants.sh "TB_ATTN_FWD2_POLY=2" "TB_ATTN_FWD2_POLY=3" ...
set -e
cd "$(dirname "$0")/.."
CS=textboost_b200/csrc
for v in "$@"; do
  tag=$(echo "$v" | tr '= ' '__')
  out=gpurun_out/variants/$tag
  mkdir -p $out
  defs=""
  for d in $v; do defs="$defs -D$d"; done
  objs=""
  for f in $CS/*.cu; do
    o=textboost_b200/build/$(basename ${f%.cu}).o
    if [ "$(basename $f)" = "attn.cu" ]; then
      o=$out/attn.o
      nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr $defs -c $f -o $o
    fi
    objs="$objs $o"
  done
  nvcc -shared -gencode arch=compute_100a,code=sm_100a -o $out/lib.so $objs
  echo "== variant $v"
  TB_LIB=$out/lib.so python scripts/probe_attn.py ${PROBE_ARGS:-}
done
