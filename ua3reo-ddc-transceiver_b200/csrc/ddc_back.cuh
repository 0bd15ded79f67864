// ddc_back.cuh - the 96 kHz / 48 kHz half of the FPGA receive chain.
//
//   cic_combine  : rx_cic.vhd:293-404  (five combs + [59:44] output slice), from chunk partial states
//   comp_fir     : rx_ciccomp.vhd:339-616 (65-tap symmetric FIR, decimate by 2, convergent round, wrap)
//   hilb_pack    : rx_hilb.vhd:375-935 (256-tap Hilbert, per-product convergent rounding),
//                  data_delay.v:16-30 (130-sample Q delay), stm32_interface.v:228-271 (8-byte frame)
//
// All three are pure FIR forms over halo-prefixed per-channel arrays, so every output sample is an
// independent thread; the halo (history from the previous ADC block) is rotated by ddc_rotate.
#pragma once
#include "ua3_common.cuh"

namespace ua3 {

// ---- CIC: comb output from five consecutive chunk records --------------------------------------
// With S_m = A512*S_{m-1} + L_m and z_m = (S_m)[stage 5], the five combs give the 5th backward
// difference  c5_m = sum_i (-1)^i C(5,i) z_{m-i}.  Substituting, c5_m = sum_{p=0..4} sum_{k=1..5}
// G[p][k] * L_{m-p}[k]  with  G[p][k] = sum_{i<=p} (-1)^i C(5,i) C(512(p-i), 5-k)  (mod 2^64); the
// p >= 5 terms vanish because a 5th difference annihilates polynomials of degree <= 4.
// G is computed on the host (api.cu: build_cic_weights) and passed in constant memory.
UA3_HD int16_t cic_combine(const uint64_t* __restrict__ Lrec /* record of chunk m, rail offset applied */,
                           const uint64_t* __restrict__ G /* [5][5] */) {
    uint64_t acc = 0;
#pragma unroll
    for (int p = 0; p < 5; ++p) {
        const uint64_t* r = Lrec - (ptrdiff_t)p * kLRec;
#pragma unroll
        for (int k = 0; k < 5; ++k) acc += G[p * 5 + k] * r[k];
    }
    return (int16_t)(uint16_t)(acc >> 44);      // output_typeconvert <= section_out10(59 DOWNTO 44)
}

// ---- compensator: y[k] = sum_j h[j] * u'[2k - j], u' = natural chunk outputs -------------------
// U points at the halo-prefixed array: U[kUHalo + i] = u'[i] of this block, so the window of frame k
// is U[2k .. 2k+64] with U[2k+64-j] multiplying h[j].
UA3_HD int16_t comp_fir(const int16_t* __restrict__ U, const int16_t* __restrict__ h, int k) {
    const int16_t* w = U + 2 * k;
    int64_t acc = 0;
#pragma unroll 5
    for (int j = 0; j < kCompTaps; ++j) acc += (int32_t)h[j] * (int32_t)w[kCompTaps - 1 - j];
    // rx_ciccomp.vhd:616: low 31 bits + 0x3FFF + bit15, wrap at 31 bits, >> 15, keep 16 bits
    const uint32_t a31 = (uint32_t)acc & 0x7FFFFFFFu;
    const uint32_t r31 = (a31 + 0x3FFFu + (((uint32_t)acc >> 15) & 1u)) & 0x7FFFFFFFu;
    const int32_t s = (int32_t)(r31 << 1) >> 16;          // sign-extend from bit 30, then >> 15
    return (int16_t)s;
}

// ---- Hilbert: v[n] = sum_t c[t] * yI[n - t] with per-product rounding -------------------------
// YI points at the halo-prefixed array: YI[kYIHalo + i] = yI[i]; window of frame n is YI[n .. n+255].
UA3_HD int16_t hilb_fir(const int16_t* __restrict__ YI, const int16_t* __restrict__ c, int n) {
    const int16_t* w = YI + n;
    // |sum| <= sum|c| * 2^15 / 2 < 2^31, so the 40-bit accumulator of the HDL never wraps and
    // a 32-bit accumulator is exact.
    int32_t acc = 0;
#pragma unroll 8
    for (int t = 0; t < kHilbTaps; ++t) {
        const int32_t p = (int32_t)c[t] * (int32_t)w[kHilbTaps - 1 - t];
        acc += (p + ((p >> 1) & 1)) >> 1;                  // rx_hilb.vhd:903
    }
    // rx_hilb.vhd:935: low 30 bits + 0x1FFF + bit14, wrap at 30 bits, >> 14, keep 16 bits
    const uint32_t a30 = (uint32_t)acc & 0x3FFFFFFFu;
    const uint32_t r30 = (a30 + 0x1FFFu + (((uint32_t)acc >> 14) & 1u)) & 0x3FFFFFFFu;
    const int32_t s = (int32_t)(r30 << 2) >> 16;          // sign-extend from bit 29, then >> 14
    return (int16_t)s;
}

// stm32_interface.v:228-271 byte order, packed into one little-endian 64-bit store.
UA3_HD uint64_t frame_pack(int16_t spec_q, int16_t spec_i, int16_t voice_q, int16_t voice_i) {
    auto be = [](int16_t v) -> uint64_t { const uint16_t u = (uint16_t)v; return (uint64_t)(uint16_t)((u >> 8) | (u << 8)); };
    return be(spec_q) | (be(spec_i) << 16) | (be(voice_q) << 32) | (be(voice_i) << 48);
}

}  // namespace ua3
