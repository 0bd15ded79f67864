"""The C-ABI library loads without a GPU and exports every symbol include/*.h declares."""
import ctypes
import glob
import os

import pytest

from conftest import ROOT


@pytest.fixture(scope="module")
def built_lib(pkg):
    if not os.path.exists(pkg.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    return ctypes.CDLL(pkg.LIB_PATH)


def test_exports_every_declared_symbol(pkg, built_lib):
    names = []
    for h in glob.glob(os.path.join(ROOT, "include", "*.h")):
        names += pkg.declared_symbols(h)
    assert len(names) >= 20
    missing = [n for n in names if not hasattr(built_lib, n)]
    assert not missing, missing


def test_create_fails_loudly_without_gpu(pkg):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(pkg.UA3Error, match="no CUDA device|no CPU fallback"):
        pkg.Receiver(4)


def test_phrase_from_frequency_host_function(pkg, oracle):
    lib = pkg.load_library()
    for f in (7100000, 14150000, 1800000, 24576000, 30000000, 50000000, 145500000, 433000000):
        s = ctypes.c_int()
        assert pkg.phrase_from_frequency(f, lib) == (oracle.lib().ua3g_phrase_from_frequency(f, ctypes.byref(s)), bool(s.value))


def test_missing_library_raises(pkg, tmp_path):
    with pytest.raises(pkg.UA3Error, match="no CPU fallback"):
        pkg.load_library(str(tmp_path / "nope.so"))


def test_smeter_dbm_host_function(pkg):
    """TRX_RX_dBm (stm32f4xx_it.c:398-407): 10*log10(P/1mW) of the RF input voltage derived from the S-meter extremes."""
    import math
    lib = pkg.load_library()
    for mx, mn, g in [(30000.0, -30000.0, 50), (500.0, -480.0, 50), (5.0, -5.0, 20), (0.0, 0.0, 50)]:
        vpp = (mx / g - mn / g) / 16.0
        v = max(vpp / 4095.0 * 0.3535 / 4 * 0.2, 1e-7)
        want = 10 * math.log10(v * v / 0.05)
        got = lib.ua3reo_smeter_dbm(mx, mn, g)
        assert abs(got - want) <= 1.5, (mx, mn, g, got, want)      # log10f_fast is a cubic fit, result truncated to int16
