#!/bin/bash
# DEVELOPMENT AID: host emulation build of the CUDA sources (see tools/emu/cuda_runtime.h).
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
CSRC="$HERE/../../ua3reo-ddc-transceiver_b200/csrc"
mkdir -p "$HERE/_build"
SRCS=()
for f in "$CSRC"/*.cu; do
  case "$(basename "$f")" in peak.cu|fanout.cu) continue;; esac   # device micro-benchmarks / CUDA IPC: nothing to emulate
  SRCS+=("$f")
done
g++ -std=c++20 -O1 -g -fPIC -shared -pthread -ffp-contract=off -I"$HERE" -I"$CSRC" -Wno-unknown-pragmas \
    -x c++ "${SRCS[@]}" "$HERE/emu_globals.cpp" -o "$HERE/_build/libua3reo_emu.so"
echo "built $HERE/_build/libua3reo_emu.so"
