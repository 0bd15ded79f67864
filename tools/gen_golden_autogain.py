#!/usr/bin/env python3
"""Generates tests/golden/autogain_cases.npz: TRX_ADC_MAXAMPLITUDE sequences and what the REFERENCE FIRMWARE's own
TRX_DoAutoGain() (trx_manager.c:268-356, host-built unmodified by oracle/ref_harness into oracle/_ref/fw_autogain)
leaves in autogain_stage, autogain_wait_reaction, TRX.Preamp, TRX.ATT, TRX.LPF, TRX.BPF after every call.

Run:  make -C oracle/ref_harness && python tools/gen_golden_autogain.py
"""
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FW = os.path.join(ROOT, "oracle", "_ref", "fw_autogain")


def run(seq):
    out = subprocess.run([FW], input=" ".join(str(int(v)) for v in seq) + "\n", capture_output=True, text=True, check=True).stdout
    return np.array([[int(x) for x in line.split()] for line in out.strip().splitlines()], np.uint8)


def sequences():
    rng = np.random.default_rng(20261018)
    yield "quiet_band_climbs_to_preamp", [60] * 40
    yield "strong_signal_stays_attenuated", [900] * 30
    yield "overload_after_preamp_falls_back", [60] * 28 + [1500] * 3 + [60] * 30
    yield "borderline", [275, 276, 277, 438, 439, 1100, 1101] * 8
    yield "negative_maximum_reads_as_large", [4000, 3000, 2049] * 6 + [50] * 30     # the unsigned 12-bit decode of fpga.c:270
    yield "random_walk", np.clip(np.cumsum(rng.integers(-120, 125, 400)) + 300, 0, 2047)
    yield "random_bursts", np.where(rng.random(400) < 0.04, 1800, rng.integers(20, 400, 400))


def main():
    assert os.path.exists(FW), "build oracle/_ref/fw_autogain first (make -C oracle/ref_harness)"
    out = {}
    for name, seq in sequences():
        seq = np.asarray(seq, np.int16)
        out[name + "/in"] = seq
        out[name + "/out"] = run(seq)
        print("%-36s %4d steps, final stage %d" % (name, seq.size, out[name + "/out"][-1, 0]))
    path = os.path.join(ROOT, "tests", "golden", "autogain_cases.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    sys.exit(main())
