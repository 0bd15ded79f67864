#!/usr/bin/env python3
"""DEVELOPMENT AID: per-kernel device times of the DDC for one or more builds of the library.
usage: python tools/exp/bench_front.py lib1.so [lib2.so ...]  (1024 channels, 2^20-sample blocks, 32 blocks)"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import ua3reo_loader  # noqa: E402

pkg = ua3reo_loader.load()
n_ch, block, steps = 1024, 1 << 20, 32
rng = np.random.default_rng(1)
adc = rng.integers(-2048, 2048, block, dtype=np.int16)
fcw = rng.integers(0, 1 << 22, n_ch, dtype=np.uint32)
ref = None
for lib in sys.argv[1:]:
    pkg.LIB_PATH = os.path.abspath(lib)
    rx = pkg.Receiver(n_ch, block)
    rx.set_fcw(fcw)
    for _ in range(3):
        rx.push(adc)
    rx.profile_begin(steps)
    for _ in range(steps):
        rx.push(adc)
    kms, nb = rx.profile_end()
    fr = rx.read_frames()
    if ref is None:
        ref = fr
    same = bool(np.array_equal(fr, ref))
    rx.close()
    print(json.dumps({"lib": os.path.basename(lib), "same_frames_as_first": same,
                      "ms": {k: round(v / nb, 5) for k, v in kms.items()}}))
