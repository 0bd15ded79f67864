#!/usr/bin/env python3
"""tests/golden/bus_cases.npz: what crosses the MCU <-> FPGA byte bus when the reference sits on both ends - the firmware's
unmodified fpga.c against stm32_interface.v executed from its source text (oracle/_ref/fw_bus_hdl; tools/verilog_eval.py,
oracle/ref_harness/bus_hdl.c).  The GPU tests compare the product's wire-format entry points with these vectors
(ua3reo_duc_push_wire, ua3reo_get_params); tests/test_verilog_pin.py regenerates them and fails when they differ.

  tx_iq      int16 [n, 2]   I, Q the firmware sends (FPGA_Audio_SendBuffer, integer valued)
  tx_wire    uint8 [n, 4]   the four data bytes of command 3 as they crossed the bus
  tx_latched int16 [n, 2]   TX_I, TX_Q of stm32_interface.v after the exchange (what feeds tx_ciccomp)
  gp_adc     int16 [c, m]   ADC samples between two GET PARAMS commands
  gp_flags   int32 [c, 2]   ADC_OTR, DAC_OTR pins during the read
  gp_packet  uint8 [c, 5]   the five bytes of command 2 as the MCU read them
  gp_decoded int32 [c, 2]   TRX_ADC_MINAMPLITUDE, TRX_ADC_MAXAMPLITUDE as FPGA_fpgadata_getparam() decoded them
"""
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
OUT = os.path.join(ROOT, "tests", "golden", "bus_cases.npz")


def generate(tmp):
    import test_verilog_pin as T
    rng = np.random.default_rng(20261019)
    iq = rng.integers(-32768, 32768, (256, 2)).astype(np.int16)
    iq[:6] = [[-32768, 32767], [32767, -32768], [1, -1], [255, 256], [-256, -255], [0, 0]]
    latched, wire = T.run_bus_tx(iq, tmp)
    m = 4096
    ranges = [(-1500, 1199), (-2048, 2047), (-900, -3), (5, 1800), (0, 0), (-1, -1)]
    adcs, flags, packets, decoded = [], [], [], []
    for c, (lo, hi) in enumerate(ranges):
        adc = rng.integers(lo, hi + 1, m).astype(np.int16)
        adc[7], adc[m - 9] = lo, hi
        otr = int(adc.min() == -2048 or adc.max() == 2047)          # the AD9226 raises OTR at either rail
        dac = c & 1
        kv = T.run_bus_params(7100000, 1, 0, 0, adc, otr, dac, tmp)
        adcs.append(adc)
        flags.append([otr, dac])
        packets.append([kv["packet%d" % i] for i in range(5)])
        decoded.append([kv["TRX_ADC_MINAMPLITUDE"], kv["TRX_ADC_MAXAMPLITUDE"]])
    return {"tx_iq": iq, "tx_wire": wire, "tx_latched": latched, "gp_adc": np.array(adcs, np.int16),
            "gp_flags": np.array(flags, np.int32), "gp_packet": np.array(packets, np.uint8), "gp_decoded": np.array(decoded, np.int32)}


if __name__ == "__main__":
    with tempfile.TemporaryDirectory() as tmp:
        data = generate(tmp)
    np.savez_compressed(OUT, **data)
    print("wrote", OUT, {k: v.shape for k, v in data.items()})
