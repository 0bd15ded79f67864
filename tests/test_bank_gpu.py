"""ua3reo_bank_*: channels sharded over several device contexts from one host process, the ADC block fanned out with peer
copies.  On a one-GPU box the same device is named twice (two contexts, device-to-device fan-out); with two or more GPUs the
second slab lives on device 1 and the fan-out crosses NVLink."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _devices():
    import torch
    return [0, 1] if torch.cuda.device_count() >= 2 else [0, 0]


def test_bank_equals_single_context_and_golden(pkg, oracle):
    n_ch, block = 37, 1 << 17                                # odd: slabs of 19 and 18
    fcw = pkg.random_fcw(n_ch, seed=9)
    mix = [(0, 2700), (1, 2700), (4, 500), (10, 6000), (8, 15000)]
    sets = [dict(mode=mix[c % 5][0], filter_width=mix[c % 5][1], dnr=(c // 5) % 2) for c in range(n_ch)]
    adc = oracle.synth_adc(3 * block, seed=31)
    bank = pkg.Bank(_devices(), n_ch, block)
    assert bank.slabs() == [(0, 19), (19, 18)]
    one = pkg.Receiver(n_ch, block)
    for r in (bank, one):
        r.set_fcw(fcw)
        r.rx_enable(True)
        r.rx_set([one.rx_defaults(**s) for s in sets])
    for b in range(3):
        blk = adc[b * block:(b + 1) * block]
        assert bank.push(blk) == one.push(blk) == block // 1024
        f_b, f_1 = bank.read_frames(), one.read_frames()
        assert np.array_equal(f_b, f_1), "frames of block %d" % b
        assert np.array_equal(bank.read_audio(), one.read_audio()) and np.array_equal(bank.read_spectra(), one.read_spectra())
        if b == 0:
            for c in (0, 18, 19, n_ch - 1):                  # both sides of the slab boundary
                assert np.array_equal(f_b[c], oracle.GoldenDDC(int(fcw[c])).push(blk)), c
    bank.close()
    one.close()


def test_bank_rejects_bad_arguments(pkg):
    with pytest.raises(pkg.UA3Error):
        pkg.Bank([0, 0, 0], 2, 1 << 16)                      # fewer channels than devices
    bank = pkg.Bank([0], 4, 1 << 16)
    with pytest.raises(pkg.UA3Error):
        bank.push(np.zeros((1 << 16) + 1024, np.int16))      # block larger than max_block_samples
    with pytest.raises(pkg.UA3Error):
        bank.set_fcw([1, 2, 3], first=2)                     # range beyond the bank
    bank.close()
