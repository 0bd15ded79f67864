#!/usr/bin/env python3
"""Generates tests/golden/cw_cases.npz: keyed 350 Hz Morse audio and what the REFERENCE FIRMWARE's own
CWDecoder_Process() (cw_decoder.c, host-built unmodified by oracle/ref_harness into oracle/_ref/fw_cw, 4 ms tick per
192-sample block) makes of it - per block the Goertzel magnitude, CW_Decoder_WPM, the element buffer and the
characters appended to the text bar.

Run:  make -C oracle/ref_harness && python tools/gen_golden_cw.py
"""
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FW = os.path.join(ROOT, "oracle", "_ref", "fw_cw")
MORSE = {'A': '.-', 'B': '-...', 'C': '-.-.', 'D': '-..', 'E': '.', 'F': '..-.', 'G': '--.', 'H': '....', 'I': '..', 'J': '.---',
         'K': '-.-', 'L': '.-..', 'M': '--', 'N': '-.', 'O': '---', 'P': '.--.', 'Q': '--.-', 'R': '.-.', 'S': '...', 'T': '-',
         'U': '..-', 'V': '...-', 'W': '.--', 'X': '-..-', 'Y': '-.--', 'Z': '--..', '1': '.----', '2': '..---', '3': '...--',
         '4': '....-', '5': '.....', '6': '-....', '7': '--...', '8': '---..', '9': '----.', '0': '-----', '?': '..--..', '/': '-..-.'}


def keyed_audio(text, wpm, amp=3000.0, noise=30.0, seed=1, lead=0.5, tail=1.0, fs=48000):
    """Keyed 350 Hz tone: PARIS timing (dit = 1.2 / wpm s, dah 3, element gap 1, letter gap 3, word gap 7)."""
    dit = 1.2 / wpm
    seq = [(0, lead)]
    for ch in text:
        if ch == ' ':
            seq.append((0, 4 * dit))
            continue
        for e in MORSE[ch]:
            seq.append((1, dit if e == '.' else 3 * dit))
            seq.append((0, dit))
        seq.append((0, 2 * dit))
    seq.append((0, tail))
    env = np.concatenate([np.full(int(round(d * fs)), on, np.float64) for on, d in seq])
    t = np.arange(env.size) / fs
    rng = np.random.default_rng(seed)
    x = amp * env * np.sin(2 * np.pi * 350.0 * t) + rng.normal(0, noise, env.size)
    return x[:x.size // 192 * 192].astype(np.float32)


def run_reference(audio):
    with tempfile.NamedTemporaryFile(suffix=".f32") as f:
        audio.tofile(f.name)
        out = subprocess.run([FW, f.name], capture_output=True, text=True, check=True).stdout
    mag, wpm, code, text = [], [], [], []
    for line in out.strip().splitlines():
        a, b, c, d = line.split()
        mag.append(float.fromhex(a)); wpm.append(int(b)); code.append("" if c == "=" else c); text.append("" if d == "=" else bytes.fromhex(d).decode("ascii"))
    return np.array(mag, np.float32), np.array(wpm, np.uint16), code, text


CASES = [("cq_20wpm", "CQ CQ DE UA3REO UA3REO K", 20, 3000.0, 30.0), ("fast_32wpm", "TEST 599 5NN TU 73", 32, 2000.0, 20.0),
         ("slow_12wpm_noisy", "SOS SOS", 12, 1500.0, 150.0), ("speed_change", None, 0, 0, 0), ("punctuation", "WHAT? 1/2 OK", 18, 4000.0, 10.0)]


def main():
    assert os.path.exists(FW), "build oracle/_ref/fw_cw first (make -C oracle/ref_harness)"
    out = {}
    for name, text, wpm, amp, noise in CASES:
        if text is None:
            audio = np.concatenate([keyed_audio("PARIS PARIS", 25, seed=2, tail=0.3), keyed_audio("PARIS PARIS", 14, seed=3, lead=0.2)])
        else:
            audio = keyed_audio(text, wpm, amp, noise)
        mag, w, code, txt = run_reference(audio)
        # the audio itself is not stored (10 MB); keyed_audio() regenerates it from the case parameters
        out[name + "/magnitude"] = mag
        out[name + "/wpm"] = w
        out[name + "/code"] = np.array(code)
        out[name + "/text"] = np.array(txt)
        print("%-18s %5d blocks  decoded %r  final wpm %d" % (name, mag.size, "".join(txt), w[-1]))
    path = os.path.join(ROOT, "tests", "golden", "cw_cases.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    sys.exit(main())
