// duc.cu - sm_100a kernel for the UA3REO transmit DUC, batched over channels:
//   tx_ciccomp.vhd:199-483 (x2 compensating interpolator) -> tx_cic.vhd:153-420 (x512 CIC interpolator with
//   Hogenauer pruning) -> tx_mixer.v:64-71 (I*sin14, Q*cos14) -> tx_summator.v:74-80 -> DAC_corrector.v:15-21.
//
// Unlike the receive CIC, the pruned integrators (an arithmetic >> 8 between stages) are NOT linear, so the
// cascade cannot be split along time: each channel is a strict recurrence over 49.152 MHz clocks.  A channel
// owns a PAIR of lanes (even = I rail, odd = Q rail).  Each lane runs its rail's compensator and CIC and
// multiplies by its own NCO output - the Q lane evaluates the sine path at phase + 2^20, which is bit-identical
// to the cosine path (cos_c[k] == sin_c[k+512], exhaustively checked) - and one shuffle per clock brings the
// two 28-bit products together for the summator / DAC corrector on the I lane, which stores 8 words at a time.
#include "duc_launch.h"
#include "ua3_common.cuh"
#include "ddc_front.cuh"
#include "tables_ddc.inc"

namespace ua3 {

__constant__ int16_t c_tx_c1[24];
__constant__ int16_t c_tx_c2[24];

// NCO sine, 14 bit: (sin_c*cos_f + sin_f*cos_c + 2^12) >> 13, P = phase << 10.  Like the receive front kernel, the DUC
// reads it from a table indexed by (coarse address, fine-sine level) - 2048 x 26 int16 = 104 KB of shared memory per
// CTA - instead of evaluating the angle sum every clock.
UA3_HD int16_t duc_tab_entry(int32_t sc, int32_t cc, int32_t sf) { return (int16_t)((sc * kCosF + sf * cc + 4096) >> 13); }
UA3_D int32_t nco_sin14(const int16_t* __restrict__ tab, uint32_t P) {
    return tab[nco_bigtab_index(P >> 21, nco_fine_level(P))];
}

constexpr int kDucWarps = 4;                                     // 64 channels per CTA: one warp per SM sub-partition
constexpr size_t kDucSmemBytes = (size_t)kBigTabWords * sizeof(int16_t);

UA3_D int16_t tx_round14(int32_t acc) {      // tx_ciccomp.vhd:469 on the low 30 bits, wrap
    const uint32_t a30 = (uint32_t)acc & 0x3FFFFFFFu;
    const uint32_t r30 = (a30 + 0x1FFFu + (((uint32_t)acc >> 14) & 1u)) & 0x3FFFFFFFu;
    return (int16_t)((int32_t)(r30 << 2) >> 16);
}

__global__ void __launch_bounds__(32 * kDucWarps)
duc_kernel(const int16_t* __restrict__ iq_in, uint32_t n_in, uint32_t in_ch_stride, const int16_t* __restrict__ nco_tab,
           const uint32_t* __restrict__ fcw, DucState* __restrict__ state, uint32_t n_ch, uint16_t* __restrict__ dac,
           size_t dac_ch_stride) {
#if defined(UA3_HOST_EMU)
    static int16_t s_tab[kBigTabWords];
#else
    extern __shared__ __align__(16) int16_t s_tab[];
#endif
    const int lane = threadIdx.x & 31, pair = lane >> 1, rail = lane & 1, warp = threadIdx.x >> 5;
    {
        const uint4* src = reinterpret_cast<const uint4*>(nco_tab);
        uint4* dst = reinterpret_cast<uint4*>(s_tab);
        for (int i = threadIdx.x; i < (int)(kDucSmemBytes / 16); i += (int)blockDim.x) dst[i] = src[i];
    }
    __syncthreads();
    const uint32_t cta_warps = blockDim.x >> 5;                 // 1, 2 or 4: chosen by duc_launch so that the warps spread over the SMs
    const uint32_t ch_raw = (blockIdx.x * cta_warps + (uint32_t)warp) * 16u + (uint32_t)pair;
    if ((blockIdx.x * cta_warps + (uint32_t)warp) * 16u >= n_ch) return;     // whole warp beyond the bank (no barriers below)
    const bool live = ch_raw < n_ch;
    const uint32_t ch = live ? ch_raw : (n_ch - 1u);
    DucState& S = state[ch];

    int32_t dp[24];
#pragma unroll
    for (int i = 0; i < 24; ++i) dp[i] = S.dp[rail][i];
    int32_t wreg = S.wreg[rail];
    int64_t D1 = S.d[rail][0], D2 = S.d[rail][1], D3 = S.d[rail][2], D4 = S.d[rail][3], D5 = S.d[rail][4];
    int64_t I6 = S.i[rail][0], I7 = S.i[rail][1], I8 = S.i[rail][2], I9 = S.i[rail][3], I10 = S.i[rail][4];
    const uint32_t F = fcw[ch] << 10;
    // the Q lane reads the sine path a quarter turn ahead == the cosine path
    uint32_t P = (S.phase << 10) + (rail ? (1u << 30) : 0u);
    uint32_t otr = S.otr;
    const int64_t kLow = ~(int64_t)15;
    // The interpolator feeds its first integrator only on the phase_0 clock (zero stuffing), so I6 is constant for the
    // other 511 clocks of a 96 kHz period and the second integrator's increment is a constant that changes once.
    int64_t c6 = (I6 >> 8) & kLow;

    const int16_t* in = iq_in + (size_t)ch * in_ch_stride;
    uint16_t* out = dac + (size_t)ch * dac_ch_stride;

    for (uint32_t s = 0; s < n_in; ++s) {
        // ---- tx_ciccomp: shift the new sample in, two polyphase outputs ----
#pragma unroll
        for (int i = 23; i > 0; --i) dp[i] = dp[i - 1];
        dp[0] = in[2 * s + rail];
        int32_t a1 = 0, a2 = 0;
#pragma unroll
        for (int i = 0; i < 24; ++i) { a1 += (int32_t)c_tx_c1[i] * dp[i]; a2 += (int32_t)c_tx_c2[i] * dp[i]; }
        const int32_t z0 = tx_round14(a1), z1 = tx_round14(a2);
#pragma unroll 1
        for (int h = 0; h < 2; ++h) {
            const int32_t znew = h ? z1 : z0;
            // ---- phase_0: comb chain from the PRE-edge input register, zero-stuffed into the first integrator ----
            const int64_t V1 = (int64_t)((uint64_t)(int64_t)wreg << 47);        // (w << 43), left-aligned by 4
            const int64_t O1 = V1 - D1;
            const int64_t V2 = (O1 >> 1) & kLow;
            const int64_t O2 = V2 - D2;
            const int64_t V3 = (O2 >> 1) & kLow;
            const int64_t O3 = V3 - D3;
            const int64_t V4 = (O3 >> 1) & kLow;
            const int64_t O4 = V4 - D4;
            const int64_t UP = O4 - D5;
            D1 = V1; D2 = V2; D3 = V3; D4 = V4; D5 = O4;
            wreg = znew;
            for (int t8 = 0; t8 < 512; t8 += 8) {
                uint32_t pk[4] = {0, 0, 0, 0};
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const int32_t o14 = (int32_t)(I10 >> 50);                     // output_register <= section_out10(59:46)
                    // integrators, all from PRE-edge values, pruned by 8 bits between stages
                    const int64_t n10 = I10 + ((I9 >> 8) & kLow);
                    const int64_t n9 = I9 + ((I8 >> 8) & kLow);
                    const int64_t n8 = I8 + ((I7 >> 8) & kLow);
                    const int64_t n7 = I7 + c6;
                    if (t8 == 0 && u == 0) { I6 = I6 + UP; c6 = (I6 >> 8) & kLow; }   // upsampling: only on the phase_0 clock
                    I10 = n10; I9 = n9; I8 = n8; I7 = n7;
                    const int32_t m = o14 * nco_sin14(s_tab, P);                  // s14 x s14 -> s28
                    P += F;
                    const int32_t mq = __shfl_xor_sync(0xffffffffu, m, 1);
                    const int32_t full = m + mq;
                    const int32_t sum = (full << 4) >> 4;                         // wrap to s28
                    otr += (sum != full) ? 1u : 0u;
                    const uint32_t word = (uint32_t)((sum >> 14) + 8191) & 0x3FFFu;
                    pk[u >> 1] |= word << (16 * (u & 1));
                }
                if (live && rail == 0) {
                    uint4 v; v.x = pk[0]; v.y = pk[1]; v.z = pk[2]; v.w = pk[3];
                    *reinterpret_cast<uint4*>(out + ((size_t)(2 * s + h) * 512 + t8)) = v;
                }
            }
        }
    }
    if (live) {
#pragma unroll
        for (int i = 0; i < 24; ++i) S.dp[rail][i] = (int16_t)dp[i];
        S.wreg[rail] = (int16_t)wreg;
        S.d[rail][0] = D1; S.d[rail][1] = D2; S.d[rail][2] = D3; S.d[rail][3] = D4; S.d[rail][4] = D5;
        S.i[rail][0] = I6; S.i[rail][1] = I7; S.i[rail][2] = I8; S.i[rail][3] = I9; S.i[rail][4] = I10;
        if (rail == 0) { S.phase = (P >> 10) & 0x3FFFFFu; S.otr = otr; }
    }
}

void build_duc_nco_table(int16_t* tab /* 2048 * 26 */) {
    for (int k = 0; k < 2048; ++k)
        for (int sf = 0; sf < kSfLevels; ++sf)
            tab[nco_bigtab_index((uint32_t)k, (uint32_t)sf)] = duc_tab_entry(UA3_NCO_SIN_C[k], UA3_NCO_COS_C[k], sf);
}

cudaError_t duc_upload_constants() {
    cudaError_t e = cudaSuccess;
#if !defined(UA3_HOST_EMU)
    e = cudaFuncSetAttribute(duc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kDucSmemBytes);
    if (e != cudaSuccess) return e;
#endif
    e = cudaMemcpyToSymbol(c_tx_c1, UA3_TXCOMP_C1, sizeof(int16_t) * 24);
    if (e != cudaSuccess) return e;
    return cudaMemcpyToSymbol(c_tx_c2, UA3_TXCOMP_C2, sizeof(int16_t) * 24);
}

cudaError_t duc_launch(const DucBuffers& b, uint32_t n_in, cudaStream_t st, int* launches) {
    if (!n_in) return cudaSuccess;
    // The chain is a strict recurrence per channel, so a warp (16 channels) is latency bound: spread the warps over as many
    // SMs as there are (two CTAs fit an SM, 104 KB of NCO table each) before stacking them - 1, 2 or 4 warps per CTA.
    int sm_count = 148, dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev);
    const uint32_t n_warps = (b.n_ch + 15u) / 16u;
    uint32_t cta_warps = (uint32_t)kDucWarps;
    while (cta_warps > 1u && (n_warps + cta_warps - 1u) / cta_warps < 2u * (uint32_t)sm_count) cta_warps >>= 1;
    const uint32_t per_cta = 16u * cta_warps;
    UA3_LAUNCH(duc_kernel, (b.n_ch + per_cta - 1u) / per_cta, 32 * cta_warps, kDucSmemBytes, st, b.iq_in, n_in, b.max_in * 2u, b.nco_tab, b.fcw, b.state, b.n_ch,
               b.dac, (size_t)b.max_in * 1024u);
    if (launches) *launches += 1;
    return cudaGetLastError();
}

}  // namespace ua3
