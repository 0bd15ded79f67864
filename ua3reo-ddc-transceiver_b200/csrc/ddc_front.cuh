// ddc_front.cuh - NCO + quadrature mixer + CIC integrators for one (channel, 512-sample chunk).
//
// Replaces, per ADC clock: nco_nco_ii_0.v:299-420 (NCO), nco_shift.v:9, mixer.v:64-71,
// rx_mixer_shift.v:9 and the five integrators of rx_cic.vhd:178-289.
//
// B200 formulation (not the HDL's): the integrator cascade is linear over Z/2^64, so the state
// after a chunk is  S' = A512 * S + L  where L is the state reached from ZERO over the chunk's 512
// samples.  Every (channel, chunk) pair can therefore compute its L independently - no sequential
// carry along time - and ddc_back.cuh combines five consecutive L records with constant weights
// into the comb output.  Inside a chunk the cascade runs in 32-bit registers over 16-sample
// sub-blocks (|x| <= 2^14, so stage 5 stays below 2^14 * C(16,5) < 2^27) and is folded into the
// 64-bit chunk state with the binomial matrix A16 once per sub-block.
#pragma once
#include "ua3_common.cuh"

namespace ua3 {

// NCO fine-sine ROM in closed form: UA3_NCO_SIN_F[j] == (j * 201 + 8224) >> 14 for all j < 2048 (checked
// exhaustively in tests/test_tables.py; (j * 6433 + 2^18) >> 19 is the same function); cos_f ROM is the constant
// 8191.  The small multiplier lets the kernels apply it to the phase field IN PLACE: with the 22-bit phase
// left-aligned in 32 bits, (P & 0x1FFC00) = j << 10 and (j << 10) * 201 + (8224 << 10) stays below 2^32, so
// sf = that >> 24 - one mask, one multiply-add, one shift.
constexpr uint32_t kSinFMul = 201u;
constexpr uint32_t kSinFBias = 8224u;
constexpr int kSinFShift = 14;
constexpr int32_t kCosF = 8191;
UA3_HD uint32_t nco_fine_level(uint32_t P) {
    return ((P & 0x1FFC00u) * kSinFMul + (kSinFBias << 10)) >> (kSinFShift + 10);
}

// Packed coarse ROM word: sin_c in the high half, cos_c in the low half (both s14, sign-extended to 16).
UA3_HD uint32_t nco_pack(int32_t sin_c, int32_t cos_c) {
    return ((uint32_t)sin_c << 16) | ((uint32_t)cos_c & 0xFFFFu);
}

// One ADC sample for one channel: returns the two 15-bit CIC inputs.
//   P   : 22-bit phase, left-aligned in 32 bits (P = phase << 10) so the accumulator wraps for free
//   a9  : ADC sample pre-shifted left by 9, so that (a9 * nco12) >> 17 == s23(adc * nco12) >> 8
UA3_HD void nco_mix(const uint32_t* __restrict__ tab, uint32_t P, int32_t a9, int32_t& xi, int32_t& xq) {
    const uint32_t w = tab[P >> 21];                        // coarse address = phase[21:11]
    const int32_t sc = (int32_t)w >> 16;
    const int32_t cc = (int32_t)(int16_t)(w & 0xFFFFu);
    const int32_t sf = (int32_t)nco_fine_level(P);          // fine-sine ROM value of phase[10:0]
    // 28-bit angle-sum products + round-half-up to 14 bits + nco_shift's [13:2]  ==  (x + 2^12) >> 15
    const int32_t s12 = (sc * kCosF + sf * cc + 4096) >> 15;
    const int32_t c12 = (cc * kCosF - sc * sf + 4096) >> 15;
    xi = (int32_t)((uint32_t)a9 * (uint32_t)s12) >> 17;     // mixer I = ADC * sin (UA3REO.bdf netlist)
    xq = (int32_t)((uint32_t)a9 * (uint32_t)c12) >> 17;     // mixer Q = ADC * cos
}

// Fold a 16-sample partial state (32-bit, from zero) into the chunk state: S = A16*S + l, A16[r][c] = C(16, r-c)
// (the cascade's own response to its state over 16 clocks).  The two lowest stages of the chunk state fit 32 bits:
// over one 512-sample chunk from zero |S0| <= 512 * 16376 < 2^23 and |S1| <= C(512,2) * 16376 = 2 142 242 816 < 2^31
// (|x| <= 2048 * 2047 >> 8 = 16376), so their products with the binomials are single 32x32->64 multiply-adds.
struct ChunkState {
    int32_t s0 = 0, s1 = 0;
    uint64_t s2 = 0, s3 = 0, s4 = 0;
};
UA3_HD void fold16(ChunkState& S, int32_t l1, int32_t l2, int32_t l3, int32_t l4, int32_t l5) {
    const int64_t a0 = S.s0, a1 = S.s1;
    S.s4 += 16u * S.s3 + 120u * S.s2 + (uint64_t)(560 * a1 + 1820 * a0 + (int64_t)l5);
    S.s3 += 16u * S.s2 + (uint64_t)(120 * a1 + 560 * a0 + (int64_t)l4);
    S.s2 += (uint64_t)(16 * a1 + 120 * a0 + (int64_t)l3);
    S.s1 += 16 * S.s0 + l2;
    S.s0 += l1;
}
UA3_HD void chunk_state_out(const ChunkState& S, uint64_t out[5]) {
    out[0] = (uint64_t)(int64_t)S.s0; out[1] = (uint64_t)(int64_t)S.s1; out[2] = S.s2; out[3] = S.s3; out[4] = S.s4;
}

// Whole chunk for one channel.  adc9: 512 pre-shifted samples (shared memory on the device, read
// as 16-byte vectors, every lane of the warp reads the same address -> broadcast).
// P0: left-aligned phase of the chunk's first sample; F: left-aligned tuning word.
// out[0..4] = I-rail partial states, out[5..9] = Q-rail.
UA3_HD void front_chunk(const uint32_t* __restrict__ tab, const I4* __restrict__ adc9, uint32_t P0, uint32_t F,
                        uint64_t out[10]) {
    ChunkState SI, SQ;
    uint32_t P = P0;
#pragma unroll 1
    for (int sb = 0; sb < kCicR / kSub; ++sb) {
        int32_t i1 = 0, i2 = 0, i3 = 0, i4 = 0, i5 = 0;
        int32_t q1 = 0, q2 = 0, q3 = 0, q4 = 0, q5 = 0;
#pragma unroll
        for (int v = 0; v < kSub / 4; ++v) {
            const I4 a = adc9[sb * (kSub / 4) + v];
            const int32_t as[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                int32_t xi, xq;
                nco_mix(tab, P, as[t], xi, xq);
                P += F;
                // registered cascade (rx_cic.vhd:197-289): every stage adds the PREVIOUS value of the stage before it
                i5 += i4; i4 += i3; i3 += i2; i2 += i1; i1 += xi;
                q5 += q4; q4 += q3; q3 += q2; q2 += q1; q1 += xq;
            }
        }
        fold16(SI, i1, i2, i3, i4, i5);
        fold16(SQ, q1, q2, q3, q4, q5);
    }
    chunk_state_out(SI, out);
    chunk_state_out(SQ, out + 5);
}

// ---------------------------------------------------------------------------------------------------
// "Big table" variant: the fine-sine ROM takes only 26 distinct values (0..25), so the whole NCO output
// stage - angle-sum products, round-half-up to 14 bits and nco_shift's [13:2] - is a function of
// (coarse address k, sf) with 2048 x 26 = 53248 entries.  One packed word (sin12 << 16 | cos12 & 0xFFFF)
// per entry is 208 KB: it fits the 227 KB of shared memory of one persistent CTA per SM and removes the
// four angle-sum IMADs, the rounding add and the two shifts from every channel-sample.
// ---------------------------------------------------------------------------------------------------
constexpr int kSfLevels = 26;
constexpr int kBigTabWords = 2048 * kSfLevels;

UA3_HD uint32_t nco_bigtab_index(uint32_t k, uint32_t sf) { return k * kSfLevels + sf; }

UA3_HD uint32_t nco_bigtab_entry(int32_t sc, int32_t cc, int32_t sf) {
    const int32_t s12 = (sc * kCosF + sf * cc + 4096) >> 15;
    const int32_t c12 = (cc * kCosF - sc * sf + 4096) >> 15;
    return ((uint32_t)s12 << 16) | ((uint32_t)c12 & 0xFFFFu);
}

// a + b issued on the ALU pipe (see front_chunk_bt)
UA3_HD int32_t alu_add(int32_t a, int32_t b) {
#if defined(__CUDA_ARCH__)
    return __viaddmax_s32(a, b, (int)0x80000000);
#else
    return a + b;
#endif
}

UA3_HD void nco_mix_bt(const uint32_t* __restrict__ bt, uint32_t P, int32_t a9, int32_t& xi, int32_t& xq) {
    const uint32_t w = bt[nco_bigtab_index(P >> 21, nco_fine_level(P))];
    const int32_t s12 = (int32_t)w >> 16;
    const int32_t c12 = (int32_t)(int16_t)(w & 0xFFFFu);
    xi = (int32_t)((uint32_t)a9 * (uint32_t)s12) >> 17;
    xq = (int32_t)((uint32_t)a9 * (uint32_t)c12) >> 17;
}

UA3_HD void front_chunk_bt(const uint32_t* __restrict__ bt, const I4* __restrict__ adc9, uint32_t P0, uint32_t F,
                           uint64_t out[10]) {
    ChunkState SI, SQ;
    uint32_t P = P0;
#pragma unroll 1
    for (int sb = 0; sb < kCicR / kSub; ++sb) {
        int32_t i1 = 0, i2 = 0, i3 = 0, i4 = 0, i5 = 0;
        int32_t q1 = 0, q2 = 0, q3 = 0, q4 = 0, q5 = 0;
#pragma unroll
        for (int v = 0; v < kSub / 4; ++v) {
            const I4 a = adc9[sb * (kSub / 4) + v];
            const int32_t as[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                int32_t xi, xq;
                nco_mix_bt(bt, P, as[t], xi, xq);
                P += F;
                // ptxas puts every two-operand cascade add on the FMA pipe (IMAD.IADD), which is the binding pipe of
                // this kernel (87 % busy against 73 % for the ALU pipe).  The last stage's add is therefore written as
                // max(a + b, INT_MIN): one VIADDMNMX on the ALU pipe, same value (measured: 0.836 -> 0.814 ms; moving
                // more stages over makes it slower again, 0.866 ms with three, 1.06 ms with all eight).
                i5 = alu_add(i5, i4); i4 += i3; i3 += i2; i2 += i1; i1 += xi;
                q5 = alu_add(q5, q4); q4 += q3; q3 += q2; q2 += q1; q1 += xq;
            }
        }
        fold16(SI, i1, i2, i3, i4, i5);
        fold16(SQ, q1, q2, q3, q4, q5);
    }
    chunk_state_out(SI, out);
    chunk_state_out(SQ, out + 5);
}

}  // namespace ua3
