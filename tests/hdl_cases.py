"""Recipes of the HDL-vector cases shared by tools/gen_golden_hdl.py and tests/test_hdl_pin.py.
(name, ADC kind, seed, fcw, t_rx, tau): t_rx = clk_sys ticks between the last common PLL edge and the release of
RX_N, tau = ticks between the 48 kHz edge and the MCU's bus read (oracle/hdl_ref.py)."""
import numpy as np

N_ADC = 1 << 21            # 2048 frames, 4096 CIC outputs per rail

CASES = (
    ("impulse", "impulse", 0, 1234567, 5, 64),          # class (B, 3, 130)
    ("wrap", "minus2048", 0, 1 << 20, 40, 64),          # (-2048) x (-2048): the s23 wrap of rx_mixer_shift.v:9; (B, 2, 129)
    ("square", "square", 0, 987654, 70, 160),
    ("dc", "dc", 0, 0, 100, 64),
    ("random_a", "random", 11, 2345678, 0, 64),         # T_rx mod 32 = 0: alignment A; (A, 3, 130)
    ("random_b", "random", 12, 77777, 517, 64),         # the default class (B, 3, 129)
    ("random_c", "random", 13, (1 << 21) - 1, 127, 64), # (A, 3, 129)
    ("tones", "tones", 14, 605867, 300, 100),           # 7.1 MHz (functions.c:206-226), default class
    ("random_d", "random", 15, 4194303, 63, 64),        # (A, 2, 129)
    ("random_e", "random", 16, 3000001, 45, 64),        # (B, 2, 129)
    ("random_f", "random", 17, 1, 990, 640),            # late bus read: (B, 3, 130)
)


def make_adc(kind, seed, n=N_ADC):
    """12-bit two's-complement ADC stream, int16"""
    rng = np.random.default_rng(1000 + seed)
    if kind == "impulse":
        a = np.zeros(n, np.int16)
        a[5000] = 2047
        a[700000] = -2048
        return a
    if kind == "minus2048":
        return np.full(n, -2048, np.int16)
    if kind == "square":
        t = np.arange(n)
        return np.where((t // 37) & 1, -2048, 2047).astype(np.int16)
    if kind == "dc":
        return np.full(n, 1000, np.int16)
    if kind == "random":
        return rng.integers(-2048, 2048, n).astype(np.int16)
    if kind == "tones":
        t = np.arange(n, dtype=np.float64)
        x = np.zeros(n)
        for _ in range(8):
            f = rng.uniform(0.01, 0.49)
            x += np.cos(2 * np.pi * (f * t + rng.uniform()))
        x = x / 8 * 1024 + rng.normal(0, 8, n)
        return np.clip(np.rint(x), -2048, 2047).astype(np.int16)
    raise ValueError(kind)
