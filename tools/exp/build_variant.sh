#!/bin/bash
# DEVELOPMENT AID: builds a variant of the library with extra -D flags into tools/exp/_build/lib<tag>.so
# usage: tools/exp/build_variant.sh <tag> [-DUA3_BT_TG=2 ...]
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
PKG="$HERE/../../ua3reo-ddc-transceiver_b200"
TAG="$1"; shift
mkdir -p "$HERE/_build"
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo --fmad=false -std=c++17 \
    -Xcompiler -fPIC,-O2,-Wno-unknown-pragmas -Xptxas -v --shared "$@" -o "$HERE/_build/lib$TAG.so" "$PKG"/csrc/*.cu 2> "$HERE/_build/$TAG.ptxas.log"
grep -A1 "front_bt" "$HERE/_build/$TAG.ptxas.log" | grep -E "registers|spill" || true
echo "built $HERE/_build/lib$TAG.so"
