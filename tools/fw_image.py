#!/usr/bin/env python3
"""The reference's shipped firmware image (STM32/MDK-ARM/UA3REO/UA3REO.hex, Intel HEX, linked against the prebuilt
arm_cortexM4lf_math.lib of CMSIS 5.5.1 / DSP 1.6.0) as a source of CONSTANT TABLES.  TEST INFRASTRUCTURE ONLY.

CMSIS-DSP is not vendored in the reference tree and absent from the image, so oracle/cmsis_min.c restates it; its tables,
though, are linked into the reference's own binary, and so are the firmware's filter coefficient tables.  This module reads
them out of the image where it lies under /root/reference (nothing is copied into the repo):

  * twiddleCoef_512 (1024 floats), located through the arm_cfft_sR_f32_len512 instance {fftLen, pTwiddle, pBitRevTable, len};
  * armBitRevIndexTable512 and the permutation arm_bitreversal_32 makes of it;
  * sinTable_f32 (513 floats);
  * any float table of the firmware, found by its bit pattern (find_floats).

  python tools/fw_image.py            prints what it finds
"""
import os
import struct

import numpy as np

HEX = "/root/reference/STM32/MDK-ARM/UA3REO/UA3REO.hex"


def available():
    return os.path.exists(HEX)


class Image:
    def __init__(self, path=HEX):
        mem, base = {}, 0
        for line in open(path):
            line = line.strip()
            if not line.startswith(":"):
                continue
            b = bytes.fromhex(line[1:])
            assert (sum(b) & 0xFF) == 0, "Intel HEX checksum"
            n, addr, typ, data = b[0], (b[1] << 8) | b[2], b[3], b[4:4 + b[0]]
            if typ == 0:
                for i, x in enumerate(data):
                    mem[base + addr + i] = x
            elif typ == 4:
                base = ((data[0] << 8) | data[1]) << 16
            elif typ == 2:
                base = ((data[0] << 8) | data[1]) << 4
        self.lo, hi = min(mem), max(mem)
        img = bytearray(hi - self.lo + 1)
        for a, x in mem.items():
            img[a - self.lo] = x
        self.img = bytes(img)

    def read(self, addr, n):
        off = addr - self.lo
        if off < 0 or off + n > len(self.img):
            raise ValueError("address outside the image")
        return self.img[off:off + n]

    def floats(self, addr, n):
        return np.frombuffer(self.read(addr, 4 * n), dtype="<f4").copy()

    def u16(self, addr, n):
        return np.frombuffer(self.read(addr, 2 * n), dtype="<u2").copy()

    def find(self, pattern, start=0, align=4):
        """every address at which the byte pattern occurs (aligned)"""
        out, off = [], start
        while True:
            off = self.img.find(pattern, off)
            if off < 0:
                return out
            if (self.lo + off) % align == 0:
                out.append(self.lo + off)
            off += 1

    def find_floats(self, values):
        return self.find(np.asarray(values, dtype="<f4").tobytes())

    # ---- CMSIS-DSP ----
    def cfft512_instance(self):
        """(pTwiddle, pBitRevTable, bitRevLength) of arm_cfft_sR_f32_len512: the twiddle table starts {1, 0, cos, sin(2 pi / 512)}"""
        head = np.array([1.0, 0.0], dtype="<f4").tobytes()
        for tw in self.find(head):
            c, s = self.floats(tw + 8, 2)
            if abs(c - np.cos(2 * np.pi / 512)) < 1e-6 and abs(s - np.sin(2 * np.pi / 512)) < 1e-6:
                for inst in self.find(struct.pack("<I", tw)):          # the instance holds a pointer to it
                    fft_len = struct.unpack("<H", self.read(inst - 4, 2))[0]
                    p_rev, rev_len = struct.unpack("<IH", self.read(inst + 4, 6))
                    if fft_len == 512:
                        return tw, p_rev, rev_len
        raise LookupError("arm_cfft_sR_f32_len512 not found in the image")

    def bitrev_permutation(self, n=512):
        """what arm_bitreversal_32(p, bitRevLen, table) does to an array of n complex floats: index i of the result holds
        element perm[i] of the input (the table holds BYTE offsets / 4 of 8-byte complex elements, swapped pairwise)"""
        _, p_rev, rev_len = self.cfft512_instance()
        tab = self.u16(p_rev, rev_len)
        words = np.arange(2 * n, dtype=np.int64)                          # 32-bit words: re, im interleaved
        for i in range(0, rev_len, 2):
            a, b = int(tab[i]) >> 2, int(tab[i + 1]) >> 2
            words[[a, b]] = words[[b, a]]
            words[[a + 1, b + 1]] = words[[b + 1, a + 1]]
        assert np.array_equal(words[1::2], words[0::2] + 1)
        return words[0::2] // 2

    def sin_table(self):
        """sinTable_f32: 513 floats, sin(2 pi k / 512); found by shape (0, ~0.01227, ..., 1 at k = 128, ~0 at 256 and 512)"""
        for a in self.find(struct.pack("<f", 0.0)):
            try:
                t = self.floats(a, 513)
            except ValueError:
                break
            if abs(t[128] - 1.0) < 1e-7 and abs(t[1] - np.sin(2 * np.pi / 512)) < 1e-6 and abs(t[384] + 1.0) < 1e-7 \
                    and np.abs(t - np.sin(2 * np.pi * np.arange(513) / 512)).max() < 1e-6:
                return a, t
        raise LookupError("sinTable_f32 not found in the image")


if __name__ == "__main__":
    im = Image()
    print("image %#x .. %#x" % (im.lo, im.lo + len(im.img)))
    tw, rev, n = im.cfft512_instance()
    print("twiddleCoef_512 at %#x, armBitRevIndexTable at %#x (%d entries)" % (tw, rev, n))
    a, t = im.sin_table()
    print("sinTable_f32 at %#x" % a)
    p = im.bitrev_permutation()
    print("bit reversal permutation: first", p[:10])
