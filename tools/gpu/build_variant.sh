#!/bin/bash
# builds a variant of the library into ua3reo-ddc-transceiver_b200/lib/variants/<name>.so with extra nvcc flags: build_variant.sh name -DX=1 ...
set -euo pipefail
HERE="$(cd "$(dirname "$0")/../../ua3reo-ddc-transceiver_b200" && pwd)"
name="$1"; shift
mkdir -p "$HERE/lib/variants"
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo --fmad=false -std=c++17 "$@" \
    -Xcompiler -fPIC,-O2,-Wall,-Wno-unknown-pragmas -Xptxas -v --shared -o "$HERE/lib/variants/$name.so" "$HERE"/csrc/*.cu 2> "$HERE/lib/variants/$name.ptxas.log"
grep -A3 "rx_audio_kernel\|rx_filter_kernel" "$HERE/lib/variants/$name.ptxas.log" | grep "Used\|spill" | head -4
