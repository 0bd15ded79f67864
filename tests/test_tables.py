"""Constant tables: regenerated from the reference where it is present, and the closed forms the
CUDA front kernel relies on (exhaustive over their whole domain)."""
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT, have_reference


def _arr(txt, name):
    m = re.search(name + r"\[\d+\] = \{([^}]*)\}", txt)
    return np.array([int(t) for t in m.group(1).replace("\n", "").split(",") if t.strip()], dtype=np.int64)


@pytest.fixture(scope="module")
def ddc_tables():
    txt = open(os.path.join(ROOT, "oracle", "tables", "ddc_tables.h")).read()
    return {n: _arr(txt, n) for n in ["UA3_NCO_SIN_C", "UA3_NCO_COS_C", "UA3_NCO_SIN_F", "UA3_NCO_COS_F",
                                      "UA3_RXCOMP_H", "UA3_RXHILB_C", "UA3_TXCOMP_C1", "UA3_TXCOMP_C2"]}


@pytest.mark.skipif(not have_reference(), reason="reference tree not present")
def test_tables_match_reference():
    assert subprocess.call([sys.executable, os.path.join(ROOT, "tools", "gen_tables.py"), "--check"]) == 0


def test_product_and_oracle_tables_identical():
    a = open(os.path.join(ROOT, "oracle", "tables", "ddc_tables.h")).read()
    b = open(os.path.join(ROOT, "ua3reo-ddc-transceiver_b200", "csrc", "tables_ddc.inc")).read()
    assert a == b


def test_survey_checksums(ddc_tables):
    h, c = ddc_tables["UA3_RXCOMP_H"], ddc_tables["UA3_RXHILB_C"]
    assert len(h) == 65 and h.sum() == 32700 and np.abs(h).sum() == 84108 and h[32] == 19710
    assert np.array_equal(h, h[::-1]) and np.count_nonzero(h) == 65
    assert len(c) == 256 and np.abs(c).sum() == 125806 and c[127] == -20824 and c[128] == 20824
    assert np.array_equal(c, -c[::-1])
    t1, t2 = ddc_tables["UA3_TXCOMP_C1"], ddc_tables["UA3_TXCOMP_C2"]
    assert t1.sum() == 17190 and t2.sum() == 17190 and np.array_equal(t2, t1[::-1])


def test_nco_rom_closed_forms(ddc_tables):
    k = np.arange(2048)
    assert np.array_equal(ddc_tables["UA3_NCO_SIN_C"], np.round(8191 * np.sin(2 * np.pi * k / 2048)).astype(int))
    assert np.array_equal(ddc_tables["UA3_NCO_COS_C"], np.round(8191 * np.cos(2 * np.pi * k / 2048)).astype(int))
    # the kernels compute the fine-sine ROM arithmetically, in place on the left-aligned phase word
    # (ddc_front.cuh: nco_fine_level, kSinFMul / kSinFBias / kSinFShift)
    assert np.array_equal(ddc_tables["UA3_NCO_SIN_F"], (k * 6433 + (1 << 18)) >> 19)
    assert np.array_equal(ddc_tables["UA3_NCO_SIN_F"], (k * 201 + 8224) >> 14)
    phase = np.arange(1 << 22, dtype=np.uint64)
    P = (phase << np.uint64(10)) & np.uint64(0xFFFFFFFF)
    t = (P & np.uint64(0x1FFC00)) * np.uint64(201) + np.uint64(8224 << 10)
    assert int(t.max()) < (1 << 32)                                      # no 32-bit overflow in the kernel's form
    assert np.array_equal((t >> np.uint64(24)).astype(np.int64), ddc_tables["UA3_NCO_SIN_F"][(phase & np.uint64(0x7FF)).astype(np.int64)])
    # ChunkState bounds of ddc_front.cuh: the two lowest CIC stages of a 512-sample chunk fit 32 bits
    xmax = (2048 * 2047) >> 8
    assert xmax == 16376 and 512 * xmax < (1 << 23) and (512 * 511 // 2) * xmax < (1 << 31)
    assert set(ddc_tables["UA3_NCO_COS_F"].tolist()) == {8191}


def test_nco_full_phase_sweep(ddc_tables, oracle):
    """All 2^22 phases: the fused (x + 2^12) >> 15 form equals s14 rounding followed by nco_shift's [13:2],
    the 14-bit result never overflows, and a sample of phases agrees with the C golden ua3g_nco()."""
    import ctypes
    sc, cc = ddc_tables["UA3_NCO_SIN_C"], ddc_tables["UA3_NCO_COS_C"]
    sf, cf = ddc_tables["UA3_NCO_SIN_F"], ddc_tables["UA3_NCO_COS_F"]
    ph = np.arange(1 << 22, dtype=np.int64)
    kk, jj = ph >> 11, ph & 0x7FF
    s28 = sc[kk] * cf[jj] + sf[jj] * cc[kk]
    c28 = cc[kk] * cf[jj] - sc[kk] * sf[jj]
    s14, c14 = (s28 + 4096) >> 13, (c28 + 4096) >> 13
    assert s14.min() >= -8192 and s14.max() <= 8191 and c14.min() >= -8192 and c14.max() <= 8191
    assert np.array_equal(s14 >> 2, (s28 + 4096) >> 15) and np.array_equal(c14 >> 2, (c28 + 4096) >> 15)
    assert (s14 >> 2).min() == -2048          # so the (-2048)*(-2048) mixer wrap is reachable
    L = oracle.lib()
    rng = np.random.default_rng(0)
    for p in rng.integers(0, 1 << 22, 2000):
        a, b = ctypes.c_int32(), ctypes.c_int32()
        L.ua3g_nco(int(p), ctypes.byref(a), ctypes.byref(b))
        assert a.value == s14[p] and b.value == c14[p]
