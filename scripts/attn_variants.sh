#!/bin/bash
# A/B timing of attention compile-time switches.
#   here (no GPU):   bash scripts/attn_variants.sh build "TB_ATTN_FWD2_POLY=2" "TB_ATTN_FWD2_POLY=3" ...
#                    -> textboost_b200/build/variants/<tag>.so   (build/ travels with the gpurun snapshot)
#   on the GPU box:  bash scripts/attn_variants.sh run            -> probe_attn.py once per variant
set -e
cd "$(dirname "$0")/.."
CS=textboost_b200/csrc
VD=textboost_b200/build/variants
if [ "$1" = "build" ]; then
  shift
  mkdir -p $VD
  for v in "$@"; do
    tag=$(echo "$v" | tr '= ' '__')
    defs=""
    for d in $v; do defs="$defs -D$d"; done
    objs=""
    for f in $CS/*.cu; do
      o=textboost_b200/build/$(basename ${f%.cu}).o
      if [ "$(basename $f)" = "attn.cu" ]; then
        o=$VD/$tag.attn.o
        nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr $defs -c $f -o $o
      fi
      objs="$objs $o"
    done
    nvcc -shared -gencode arch=compute_100a,code=sm_100a -o $VD/$tag.so $objs
    echo "built $VD/$tag.so"
  done
else
  for so in $VD/*.so; do
    echo "== variant $(basename $so .so)"
    TB_LIB=$so python scripts/probe_attn.py ${PROBE_ARGS:-}
  done
fi
