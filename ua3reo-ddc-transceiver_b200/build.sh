#!/bin/bash
# Builds libua3reo_b200.so for sm_100a in-tree (the built .so travels to the GPU box with gpurun).
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
OUT="$HERE/lib"
mkdir -p "$OUT"
SRCS=("$HERE"/csrc/*.cu)
# --fmad=false: the float audio/FFT stages must match the firmware's unfused arithmetic
"$NVCC" -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo --fmad=false -std=c++17 \
    -Xcompiler -fPIC,-O2,-Wall,-Wno-unknown-pragmas -Xptxas -v --shared -o "$OUT/libua3reo_b200.so" "${SRCS[@]}" 2> "$OUT/ptxas.log" \
    || { cat "$OUT/ptxas.log"; exit 1; }
grep -E "error|warning" "$OUT/ptxas.log" || true
# the C host driver (the reference's host language) on top of the C ABI
gcc -O2 -std=gnu11 -Wall -I"$HERE/../include" "$HERE/host/ua3reo_rx_host.c" -o "$OUT/ua3reo_rx_host" \
    -L"$OUT" -lua3reo_b200 -lm -Wl,-rpath,'$ORIGIN'
echo "built $OUT/libua3reo_b200.so and $OUT/ua3reo_rx_host"
