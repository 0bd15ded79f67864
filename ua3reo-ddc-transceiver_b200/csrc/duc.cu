// duc.cu - sm_100a kernel for the UA3REO transmit DUC, batched over channels:
//   tx_ciccomp.vhd:199-483 (x2 compensating interpolator) -> tx_cic.vhd:153-420 (x512 CIC interpolator with
//   Hogenauer pruning) -> tx_mixer.v:64-71 (I*sin14, Q*cos14) -> tx_summator.v:74-80 -> DAC_corrector.v:15-21.
//
// Unlike the receive CIC, the pruned integrators (an arithmetic >> 8 between stages) are NOT linear, so the
// cascade cannot be split along time: each channel's integrators are a strict recurrence over 49.152 MHz clocks.
// A channel owns a PAIR of lanes of the producer warp (even = I rail, odd = Q rail); the NCO / mixer / summator /
// DAC-corrector half has no recurrence and runs on separate warps, parallel over clocks (see duc_kernel).
#include "duc_launch.h"
#include "ua3_common.cuh"
#include "ddc_front.cuh"
#include "tables_ddc.inc"

namespace ua3 {

__constant__ int16_t c_tx_c1[24];
__constant__ int16_t c_tx_c2[24];

// NCO sine, 14 bit: (sin_c*cos_f + sin_f*cos_c + 2^12) >> 13, P = phase << 10
UA3_D int32_t nco_sin14(const uint32_t* __restrict__ tab, uint32_t P) {
    const uint32_t w = tab[P >> 21];
    const int32_t sc = (int32_t)w >> 16;
    const int32_t cc = (int32_t)(int16_t)(w & 0xFFFFu);
    const int32_t j = (int32_t)((P >> 10) & 0x7FFu);
    const int32_t sf = (j * kSinFMul + (1 << 18)) >> 19;
    return (sc * kCosF + sf * cc + 4096) >> 13;
}

UA3_D int16_t tx_round14(int32_t acc) {      // tx_ciccomp.vhd:469 on the low 30 bits, wrap
    const uint32_t a30 = (uint32_t)acc & 0x3FFFFFFFu;
    const uint32_t r30 = (a30 + 0x1FFFu + (((uint32_t)acc >> 14) & 1u)) & 0x3FFFFFFFu;
    return (int16_t)((int32_t)(r30 << 2) >> 16);
}

// Producer / consumer split inside one CTA of 4 warps, 16 channels per CTA:
//   warp 0 (producer) : the part that IS a recurrence - compensator, combs and the four pruned integrators - for one group of
//                       512 clocks (one 96 kHz sample) per step; writes the 14-bit CIC outputs of its 32 lanes to shared memory
//   warps 1-3         : everything that is not a recurrence - NCO (sine and cosine from the packed coarse ROM), the two mixer
//                       products, summator, overflow count, DAC corrector - for the previous group, parallel over clocks, with
//                       coalesced 2-byte stores of consecutive DAC words
// The two halves overlap through a double-buffered shared array; one __syncthreads per group.
constexpr int kDucThreads = 128;
constexpr size_t kDucSmemBytes = 2048 * 4 + 2 * 512 * 32 * 2 + 16 * 4;

__global__ void __launch_bounds__(kDucThreads)
duc_kernel(const int16_t* __restrict__ iq_in, uint32_t n_in, uint32_t in_ch_stride, const uint32_t* __restrict__ nco_tab,
           const uint32_t* __restrict__ fcw, DucState* __restrict__ state, uint32_t n_ch, uint16_t* __restrict__ dac,
           size_t dac_ch_stride) {
#if defined(UA3_HOST_EMU)
    static uint8_t s_dyn[kDucSmemBytes] __attribute__((aligned(16)));
#else
    extern __shared__ __align__(16) uint8_t s_dyn[];
#endif
    uint32_t* s_tab = reinterpret_cast<uint32_t*>(s_dyn);
    int16_t (*s_o)[512][32] = reinterpret_cast<int16_t (*)[512][32]>(s_dyn + 2048 * 4);     // [buffer][clock][lane]
    uint32_t* s_otr = reinterpret_cast<uint32_t*>(s_dyn + 2048 * 4 + 2 * 512 * 32 * 2);     // per channel of the CTA

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, pair = lane >> 1, rail = lane & 1;
    for (int i = tid; i < 2048; i += kDucThreads) s_tab[i] = nco_tab[i];
    if (tid < 16) s_otr[tid] = 0;
    const uint32_t ch_raw = blockIdx.x * 16u + (uint32_t)pair;
    const bool live = ch_raw < n_ch;
    const uint32_t ch = live ? ch_raw : (n_ch - 1u);
    DucState& S = state[ch];

    // producer registers (warp 0 only uses them)
    int32_t dp[24];
    int32_t wreg = 0, z0 = 0, z1 = 0;
    int64_t D1 = 0, D2 = 0, D3 = 0, D4 = 0, D5 = 0, I6 = 0, I7 = 0, I8 = 0, I9 = 0, I10 = 0;
    if (warp == 0) {
#pragma unroll
        for (int i = 0; i < 24; ++i) dp[i] = S.dp[rail][i];
        wreg = S.wreg[rail];
        D1 = S.d[rail][0]; D2 = S.d[rail][1]; D3 = S.d[rail][2]; D4 = S.d[rail][3]; D5 = S.d[rail][4];
        I6 = S.i[rail][0]; I7 = S.i[rail][1]; I8 = S.i[rail][2]; I9 = S.i[rail][3]; I10 = S.i[rail][4];
    }
    const int64_t kLow = ~(int64_t)15;
    const int16_t* in = iq_in + (size_t)ch * in_ch_stride;
    const uint32_t n_groups = 2 * n_in;
    __syncthreads();

    for (uint32_t g = 0; g <= n_groups; ++g) {
        if (warp == 0) {
            if (g < n_groups) {
                const uint32_t s = g >> 1, h = g & 1;
                if (h == 0) {
                    // ---- tx_ciccomp: shift the new sample in, two polyphase outputs ----
#pragma unroll
                    for (int i = 23; i > 0; --i) dp[i] = dp[i - 1];
                    dp[0] = in[2 * s + rail];
                    int32_t a1 = 0, a2 = 0;
#pragma unroll
                    for (int i = 0; i < 24; ++i) { a1 += (int32_t)c_tx_c1[i] * dp[i]; a2 += (int32_t)c_tx_c2[i] * dp[i]; }
                    z0 = tx_round14(a1); z1 = tx_round14(a2);
                }
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
                int16_t (*dst)[32] = s_o[g & 1];
#pragma unroll 4
                for (int t = 0; t < 512; ++t) {
                    dst[t][lane] = (int16_t)(I10 >> 50);                          // output_register <= section_out10(59:46)
                    // integrators, all from PRE-edge values, pruned by 8 bits between stages
                    const int64_t n10 = I10 + ((I9 >> 8) & kLow);
                    const int64_t n9 = I9 + ((I8 >> 8) & kLow);
                    const int64_t n8 = I8 + ((I7 >> 8) & kLow);
                    const int64_t n7 = I7 + ((I6 >> 8) & kLow);
                    if (t == 0) I6 = I6 + UP;                                     // upsampling: only on the phase_0 clock
                    I10 = n10; I9 = n9; I8 = n8; I7 = n7;
                }
            }
        } else if (g > 0) {
            // ---- consumers: group g-1, 16 channels x 512 clocks over 96 threads ----
            const uint32_t gp = g - 1;
            const int16_t (*src)[32] = s_o[gp & 1];
            for (uint32_t item = (uint32_t)(tid - 32); item < 16u * 512u; item += (uint32_t)(kDucThreads - 32)) {
                const uint32_t c = item >> 9, t = item & 511u;
                const uint32_t cch = blockIdx.x * 16u + c;
                if (cch >= n_ch) continue;
                const uint32_t F = fcw[cch] << 10;
                const uint32_t P = (state[cch].phase << 10) + F * (gp * 512u + t);
                const uint32_t w = s_tab[P >> 21];
                const int32_t sc = (int32_t)w >> 16;
                const int32_t cc = (int32_t)(int16_t)(w & 0xFFFFu);
                const int32_t j = (int32_t)((P >> 10) & 0x7FFu);
                const int32_t sf = (j * kSinFMul + (1 << 18)) >> 19;
                const int32_t s14 = (sc * kCosF + sf * cc + 4096) >> 13;
                const int32_t c14 = (cc * kCosF - sc * sf + 4096) >> 13;
                const int32_t full = (int32_t)src[t][2 * c] * s14 + (int32_t)src[t][2 * c + 1] * c14;   // tx_mixer I*sin + Q*cos
                const int32_t sum = (full << 4) >> 4;                                                    // tx_summator: wrap to s28
                if (sum != full) atomicAdd(&s_otr[c], 1u);
                dac[(size_t)cch * dac_ch_stride + (size_t)gp * 512u + t] = (uint16_t)(((sum >> 14) + 8191) & 0x3FFF);
            }
        }
        __syncthreads();
    }
    // ---- state write-back ----
    if (warp == 0 && live) {
#pragma unroll
        for (int i = 0; i < 24; ++i) S.dp[rail][i] = (int16_t)dp[i];
        S.wreg[rail] = (int16_t)wreg;
        S.d[rail][0] = D1; S.d[rail][1] = D2; S.d[rail][2] = D3; S.d[rail][3] = D4; S.d[rail][4] = D5;
        S.i[rail][0] = I6; S.i[rail][1] = I7; S.i[rail][2] = I8; S.i[rail][3] = I9; S.i[rail][4] = I10;
    }
    __syncthreads();          // consumers are done reading state[].phase
    if (warp == 0 && live && rail == 0) {
        S.phase = (S.phase + fcw[ch] * (n_groups * 512u)) & 0x3FFFFFu;
        S.otr += s_otr[pair];
    }
}

cudaError_t duc_upload_constants() {
#if !defined(UA3_HOST_EMU)
    cudaError_t e0 = cudaFuncSetAttribute(duc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kDucSmemBytes);
    if (e0 != cudaSuccess) return e0;
#endif
    cudaError_t e = cudaMemcpyToSymbol(c_tx_c1, UA3_TXCOMP_C1, sizeof(int16_t) * 24);
    if (e != cudaSuccess) return e;
    return cudaMemcpyToSymbol(c_tx_c2, UA3_TXCOMP_C2, sizeof(int16_t) * 24);
}

cudaError_t duc_launch(const DucBuffers& b, uint32_t n_in, cudaStream_t st, int* launches) {
    if (!n_in) return cudaSuccess;
    UA3_LAUNCH(duc_kernel, (b.n_ch + 15u) / 16u, kDucThreads, kDucSmemBytes, st, b.iq_in, n_in, b.max_in * 2u, b.nco_tab, b.fcw, b.state, b.n_ch,
               b.dac, (size_t)b.max_in * 1024u);
    if (launches) *launches += 1;
    return cudaGetLastError();
}

}  // namespace ua3
