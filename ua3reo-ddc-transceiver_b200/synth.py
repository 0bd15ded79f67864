"""Synthetic 12-bit ADC stream of SURVEY.md 8(d): K tones at random in-band frequencies with random
phases, about -6 dBFS in total, plus white Gaussian noise (sigma 8 LSB), rounded and clipped."""
import numpy as np

SEED = 20261018


def synth_adc(n, seed=SEED, tones=8, noise_lsb=8.0, level_dbfs=-6.0):
    rng = np.random.default_rng(seed)
    t = np.arange(n, dtype=np.float64)
    f = rng.uniform(0.02, 0.48, tones)
    ph = rng.uniform(0, 2 * np.pi, tones)
    amp = 2047.0 * 10 ** (level_dbfs / 20.0) / tones
    x = np.zeros(n)
    for k in range(tones):
        x += amp * np.sin(2 * np.pi * f[k] * t + ph[k])
    x += rng.normal(0.0, noise_lsb, n)
    return np.clip(np.rint(x), -2048, 2047).astype(np.int16)


def random_fcw(n, seed=SEED):
    """Per-channel tuning words ~ U{1 .. 2^21-1}."""
    return np.random.default_rng(seed).integers(1, 1 << 21, n).astype(np.uint32)
