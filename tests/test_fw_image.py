"""Constant tables against the reference's SHIPPED firmware image (STM32/MDK-ARM/UA3REO/UA3REO.hex), read by tools/fw_image.py.

CMSIS-DSP is a third-party dependency the reference does not vendor (prebuilt arm_cortexM4lf_math.lib, CMSIS 5.5.1 / DSP 1.6.0)
and oracle/cmsis_min.c restates it - but its tables are LINKED INTO the reference's own binary, and so are the firmware's
coefficient tables.  Needs /root/reference (this container); the facts checked here do not travel and do not need to: they pin
the oracle and the generated tables, which do."""
import ctypes
import os
import re
import sys

import numpy as np
import pytest

from conftest import ROOT

sys.path.insert(0, os.path.join(ROOT, "tools"))
import fw_image  # noqa: E402

pytestmark = pytest.mark.skipif(not fw_image.available(), reason="needs the reference tree (firmware image)")


@pytest.fixture(scope="module")
def image():
    return fw_image.Image()


def _bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def test_cmsis_tables_are_the_ones_linked_into_the_firmware(image, oracle):
    L = oracle.lib()
    tw_addr, rev_addr, rev_len = image.cfft512_instance()
    assert rev_len == 448                                    # ARMBITREVINDEXTABLE_512_TABLE_LENGTH of CMSIS-DSP 1.6.0
    L.ua3_cmsis_twiddle_512.restype = ctypes.POINTER(ctypes.c_float)
    mine = np.ctypeslib.as_array(L.ua3_cmsis_twiddle_512(), shape=(1024,))
    assert np.array_equal(_bits(mine), _bits(image.floats(tw_addr, 1024))), "twiddleCoef_512"
    L.ua3_cmsis_sin_table.restype = ctypes.POINTER(ctypes.c_float)
    mine = np.ctypeslib.as_array(L.ua3_cmsis_sin_table(), shape=(513,))
    _, theirs = image.sin_table()
    assert np.array_equal(_bits(mine), _bits(theirs)), "sinTable_f32"


def test_bit_reversal_of_the_firmware_table_is_what_cmsis_min_does(image, oracle):
    """arm_bitreversal_32 over the firmware's armBitRevIndexTable512 against the digit reversal oracle/cmsis_min.c restates:
    arm_cfft_f32 with and without bitReverseFlag on distinct values gives the restatement's permutation."""
    L = oracle.lib()
    perm_fw = image.bitrev_permutation()
    assert sorted(perm_fw) == list(range(512))
    assert np.array_equal(perm_fw, [((i & 7) << 6) | (i & 0x38) | (i >> 6) for i in range(512)])     # base-8 digit reversal
    inst = ctypes.c_void_p.in_dll(L, "arm_cfft_sR_f32_len512")
    x = np.random.default_rng(5).normal(0, 1, 1024).astype(np.float32)
    a, b = x.copy(), x.copy()
    L.arm_cfft_f32.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint8, ctypes.c_uint8]
    L.arm_cfft_f32(ctypes.addressof(inst), a.ctypes.data, 0, 0)
    L.arm_cfft_f32(ctypes.addressof(inst), b.ctypes.data, 0, 1)
    a, b = a.view(np.complex64), b.view(np.complex64)
    assert np.array_equal(b, a[perm_fw])
    # and the transform the tables make is the DFT (forward, natural order after the reversal)
    ref = np.fft.fft(x.view(np.complex64).astype(np.complex128))
    assert np.abs(b - ref).max() / np.abs(ref).max() < 1e-6


def test_firmware_coefficient_tables_are_in_the_image(image):
    """oracle/tables/audio_tables.h (generated from audio_filters.c / fft.c by tools/gen_tables.py; the product's
    tables_audio.inc is the same data, tests/test_tables.py) - every filter's coefficient row occurs bit for bit in the image
    the reference ships, i.e. the sources the tables were generated from are the sources the firmware was built from."""
    txt = open(os.path.join(ROOT, "oracle", "tables", "audio_tables.h")).read()

    def table(name):
        m = re.search(r"%s\[\d+\] = \{[^\n]*\n(.*?)\};" % name, txt, re.S)
        body = re.sub(r"/\*.*?\*/", "", m.group(1))
        return np.array([int(v, 0) for v in re.findall(r"0x[0-9a-fA-F]+|\d+", body)], dtype=np.uint32)

    def present(words):
        return bool(image.find(np.asarray(words, dtype="<u4").tobytes()))

    stages = table("UA3_LPF_STAGES")
    pk, pv = table("UA3_LPF_PK").reshape(32, 11), table("UA3_LPF_PV").reshape(32, 12)
    for i in range(32):
        n = int(stages[i])
        assert present(pk[i, :n]) and present(pv[i, :n + 1]), "LPF %d" % i
    hk, hv = table("UA3_HPF_PK").reshape(6, 6), table("UA3_HPF_PV").reshape(6, 7)
    assert all(present(hk[i]) and present(hv[i]) for i in range(6))
    assert present(table("UA3_SQL_HPF_PK")) and present(table("UA3_SQL_HPF_PV"))
    assert present(table("UA3_TX_HILB_I")) and present(table("UA3_TX_HILB_Q"))
    zb, zf = table("UA3_ZOOM_BIQUAD").reshape(4, 20), table("UA3_ZOOM_FIR").reshape(4, 4)
    assert all(present(zb[i]) and present(zf[i]) for i in range(4))
