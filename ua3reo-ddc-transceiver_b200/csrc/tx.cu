// tx.cu - sm_100a kernel for the STM32 transmit-audio stage, batched over channels:
//   tx_audio_kernel : processTxAudio() (audio_processor.c:61-273) - DC filter, IIR-lattice HPF/LPF, ALC compressor,
//                     SSB (201-tap +/-45 degree Hilbert FIR pair), AM, FM (ModulateFM :590-619) and CW modulators.
// One warp per channel: lane 0 runs the sequential recurrences (DC filter, lattice filters), lane 1 the Q rail's
// DC filter; the two 201-tap FIRs - the bulk of the arithmetic, and free of recurrences - are spread over all 32
// lanes (6 output samples each), every dot product summed in the firmware's tap order.
#include "tx_launch.h"
#include "ua3_common.cuh"
#include "tables_audio.inc"

namespace ua3 {

__constant__ float c_tx_hilb_i[kTxHilbTaps];
__constant__ float c_tx_hilb_q[kTxHilbTaps];
__constant__ float c_sin_table[513];

template <int N>
UA3_D float tx_lattice_step(float x, const float* k, const float* v, float* G) {
    float f = x, acc = 0.0f;
#pragma unroll
    for (int i = 0; i < N; ++i) {
        const float g_old = G[i];
        f = f - (k[i] * g_old);
        const float g_new = (f * k[i]) + g_old;
        acc += g_new * v[i];
        if (i > 0) G[i - 1] = g_new;
    }
    acc += f * v[N];
    G[N - 1] = f;
    return acc;
}

// arm_sin_f32 / arm_cos_f32 (CMSIS-DSP 1.6.0): 512-entry table + linear interpolation
UA3_D float table_sin(float x, float quarter) {
    float in = x * 0.159154943092f + quarter;
    int32_t n = (int32_t)in;
    if ((quarter != 0.0f ? in : x) < 0.0f) n--;
    in = in - (float)n;
    float findex = 512.0f * in;
    uint32_t index = (uint32_t)findex & 0xFFFFu;
    if (index >= 512u) { index = 0; findex -= 512.0f; }
    const float fract = findex - (float)index;
    return (1.0f - fract) * c_sin_table[index] + fract * c_sin_table[index + 1];
}

__global__ void __launch_bounds__(32)
tx_audio_kernel(const int16_t* __restrict__ mic, uint32_t n_blocks, uint32_t mic_ch_stride, const TxParams* __restrict__ params,
                TxState* __restrict__ state, uint32_t n_ch, float* __restrict__ iq_f, int16_t* __restrict__ iq_w,
                int32_t* __restrict__ loop_out, uint32_t out_ch_stride) {
    __shared__ float s_i[kTxHilbTaps - 1 + kAudioBlock];     // [200 history | 192 new] input of the I-buffer FIR
    __shared__ float s_q[kTxHilbTaps - 1 + kAudioBlock];
    __shared__ float s_oi[kAudioBlock], s_oq[kAudioBlock];
    const int lane = threadIdx.x & 31;
    const uint32_t ch = blockIdx.x;
    if (ch >= n_ch) return;
    const TxParams& P = params[ch];
    TxState& S = state[ch];
    const uint8_t mode = P.mode;
    const bool usb = (mode == kModeUSB || mode == kModeDIGIU), lsb = (mode == kModeLSB || mode == kModeDIGIL);
    const bool am = (mode == kModeAM), fm = (mode == kModeNFM || mode == kModeWFM), cw = (mode == kModeCWL || mode == kModeCWU);
    const bool chain = (mode != kModeIQ) && !P.tune;
    const float A1 = (float)(1.0 - 0.00048828125);
    constexpr int H = kTxHilbTaps - 1;

    // which FIR instance filters which buffer: USB/DIGI_U: I buffer <- FIR_Q, Q buffer <- FIR_I; LSB/DIGI_L/AM: I <- FIR_I, Q <- FIR_Q
    // s_i / s_q hold the history of the *instance* FIR_TX_Hilbert_I / _Q respectively.
    for (int t = lane; t < H; t += 32) { s_i[t] = S.fir_hist[0][t]; s_q[t] = S.fir_hist[1][t]; }
    float lpf_g[kLpfMax], hpf_g[kHpfStages];
    float dcx = 0.f, dcy = 0.f, alc = S.alc_gain, fha = S.fm_hpf_a, fhb = S.fm_hpf_b;
    uint32_t facc = S.fm_accum;
    if (lane < 2) { dcx = S.dc_x[lane]; dcy = S.dc_y[lane]; }
    if (lane == 0) {
#pragma unroll
        for (int i = 0; i < kLpfMax; ++i) lpf_g[i] = S.lpf_g[i];
#pragma unroll
        for (int i = 0; i < kHpfStages; ++i) hpf_g[i] = S.hpf_g[i];
    }
    __syncwarp();

    for (uint32_t blk = 0; blk < n_blocks; ++blk) {
        const int16_t* m = mic + (size_t)ch * mic_ch_stride + (size_t)blk * kAudioBlock * 2;
        float amp = P.amplitude;
        // ---- lanes 0/1: sample load, DC filter; lane 0: HPF, LPF (audio_processor.c:81-110) ----
        if (lane < 2) {
            for (int i = 0; i < kAudioBlock; ++i) {
                float x = P.tune ? amp : (float)m[2 * i + lane];
                if (!P.tune) {
                    const float delta_x = x - dcx;
                    const float a1y = A1 * dcy;
                    const float y = delta_x + a1y;
                    dcx = x; dcy = y; x = y;
                }
                if (lane == 0 && chain) {
                    if (P.hpf_set) x = tx_lattice_step<kHpfStages>(x, P.hpf_k, P.hpf_v, hpf_g);
                    if (P.lpf_on) x = tx_lattice_step<kLpfMax>(x, P.lpf_k, P.lpf_v, lpf_g);
                }
                (lane == 0 ? s_oi : s_oq)[i] = x;
            }
        }
        __syncwarp();
        if (chain) {
            // memcpy(Q, I) then the ALC compressor (:111-140)
            float mx = 0.0f;
            for (int i = lane; i < kAudioBlock; i += 32) mx = fmaxf(mx, fabsf(s_oi[i]));
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
            if (mx == 0.0f) mx = 0.001f;
            const float target = amp / mx;
            if (target > alc) alc += (target - alc) / 500.0f;
            else alc -= (alc - target) / 500.0f;
            if (target < alc) alc = target;
            if (alc < 0.0f) alc = 0.0f;
            if ((alc * mx) > (amp * 1.1f)) alc = target;
            if (alc > 500.0f) alc = 500.0f;
            if (mx < 15.0f) alc = 0.0f;
            if ((alc > 1.0f) && (mode == kModeDIGIL || mode == kModeDIGIU || mode == kModeIQ || mode == kModeLoopback)) alc = 1.0f;
            if (P.tune) alc = 1.0f;
            __syncwarp();
            for (int i = lane; i < kAudioBlock; i += 32) {
                const float v = s_oi[i] * alc;        // both buffers hold the same signal after the copy
                s_oi[i] = v; s_oq[i] = v;
            }
            __syncwarp();
            if (cw) {                                  // :143-151
                if (!P.key_down) amp = 0.0f;
                for (int i = lane; i < kAudioBlock; i += 32) { s_oi[i] = amp; s_oq[i] = amp; }
            } else if (usb || lsb || am) {             // :152-193: arm_fir_f32 x2, 201 taps, three 64-sample sub-blocks
                // append the block to both instances' windows (both see the same input signal)
                for (int i = lane; i < kAudioBlock; i += 32) { const float v = s_oi[i]; s_i[H + i] = v; s_q[H + i] = v; }
                __syncwarp();
                for (int n = lane; n < kAudioBlock; n += 32) {
                    float ai = 0.0f, aq = 0.0f;
#pragma unroll 3
                    for (int t = 0; t < kTxHilbTaps; ++t) {
                        ai += s_i[n + t] * c_tx_hilb_i[t];     // FIR_TX_Hilbert_I
                        aq += s_q[n + t] * c_tx_hilb_q[t];     // FIR_TX_Hilbert_Q
                    }
                    // USB: I buffer gets the Q instance's output, Q buffer the I instance's; LSB/AM the other way round
                    s_oi[n] = usb ? aq : ai;
                    s_oq[n] = usb ? ai : aq;
                }
                __syncwarp();
                // slide the windows: dest [0,200) <- src [192,392); the only overlap (192..199) is read and written by the same lane, in program order
                for (int t = lane; t < H; t += 32) { s_i[t] = s_i[kAudioBlock + t]; s_q[t] = s_q[kAudioBlock + t]; }
                __syncwarp();
                if (am) {                              // :183-192
                    for (int i = lane; i < kAudioBlock; i += 32) {
                        const float I = s_oi[i], Q = s_oq[i];
                        s_oi[i] = ((I - Q) + amp) / 2.0f;
                        s_oq[i] = ((Q - I) - amp) / 2.0f;
                    }
                }
            } else if (fm) {                           // ModulateFM :590-619 (sequential: differentiator + phase accumulator)
                if (lane == 0) {
                    for (int i = 0; i < kAudioBlock; ++i) {
                        const float a = s_oi[i];
                        fhb = 0.95f * ((fhb + a) - fha);
                        fha = a;
                        // fm_mod_accum += hpf_prev_b: float sum converted back to uint32 the way the x86-64 host build
                        // of the firmware does it (through a 64-bit integer; a Cortex-M would saturate negatives to 0)
                        facc = (uint32_t)(long long)((float)facc + fhb);
                        facc %= 48000u;
                        const float sin_data = (((float)facc / 48000.0f) * 3.14159265358979f) * P.fm_index;
                        s_oi[i] = amp * table_sin(sin_data, 0.0f);
                        s_oq[i] = amp * table_sin(sin_data, 0.25f);
                    }
                }
            }
            __syncwarp();
        }
        // ---- mute, output (float as left in FPGA_Audio_SendBuffer_I/Q, int16 as sendiq puts it on the wire) ----
        float* of = iq_f + (size_t)ch * out_ch_stride * 2 + (size_t)blk * kAudioBlock * 2;
        int16_t* ow = iq_w + (size_t)ch * out_ch_stride * 2 + (size_t)blk * kAudioBlock * 2;
        int32_t* ol = loop_out + (size_t)ch * out_ch_stride * 2 + (size_t)blk * kAudioBlock * 2;
        const bool loopback = (mode == kModeLoopback) && !P.tune;    // :228: the block goes to the codec, not to the FPGA
        for (int i = lane; i < kAudioBlock; i += 32) {
            float I = s_oi[i], Q = s_oq[i];
            if (P.mute && !P.tune) { I = I * 0.0f; Q = Q * 0.0f; }
            int32_t l = 0;
            if (loopback) { I = I * P.loop_volume; l = (int32_t)I; }  // arm_scale_f32 by Volume / 50, float -> int32, right = left (:231-236)
            of[2 * i] = I; of[2 * i + 1] = Q;
            ow[2 * i] = (int16_t)(int32_t)I; ow[2 * i + 1] = (int16_t)(int32_t)Q;
            ol[2 * i] = l; ol[2 * i + 1] = l;
        }
        __syncwarp();
    }
    if (lane < 2) { S.dc_x[lane] = dcx; S.dc_y[lane] = dcy; }
    if (lane == 0) {
#pragma unroll
        for (int i = 0; i < kLpfMax; ++i) S.lpf_g[i] = lpf_g[i];
#pragma unroll
        for (int i = 0; i < kHpfStages; ++i) S.hpf_g[i] = hpf_g[i];
        S.alc_gain = alc; S.fm_hpf_a = fha; S.fm_hpf_b = fhb; S.fm_accum = facc;
    }
    for (int t = lane; t < H; t += 32) { S.fir_hist[0][t] = s_i[t]; S.fir_hist[1][t] = s_q[t]; }
}

__global__ void tx_init_state_kernel(TxState* __restrict__ state, uint32_t n) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) state[i].alc_gain = 1.0f;             // ALC_need_gain = 1.0f (audio_processor.c:35)
}

__global__ void tx_clear_filters_kernel(TxState* __restrict__ state, const uint8_t* __restrict__ flags, uint32_t first, uint32_t n) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    TxState& S = state[first + i];
    if (flags[i] & 1) for (int k = 0; k < kLpfMax; ++k) S.lpf_g[k] = 0.0f;
    if (flags[i] & 2) for (int k = 0; k < kHpfStages; ++k) S.hpf_g[k] = 0.0f;
}

cudaError_t tx_upload_constants(const float* sin_table) {
    float ci[kTxHilbTaps], cq[kTxHilbTaps];
    for (int i = 0; i < kTxHilbTaps; ++i) {
        uint32_t a = UA3_TX_HILB_I[i], b = UA3_TX_HILB_Q[i];
        memcpy(&ci[i], &a, 4); memcpy(&cq[i], &b, 4);
    }
    cudaError_t e = cudaMemcpyToSymbol(c_tx_hilb_i, ci, sizeof ci);
    if (e != cudaSuccess) return e;
    e = cudaMemcpyToSymbol(c_tx_hilb_q, cq, sizeof cq);
    if (e != cudaSuccess) return e;
    return cudaMemcpyToSymbol(c_sin_table, sin_table, sizeof(float) * 513);
}

cudaError_t tx_launch_init_state(const TxBuffers& b, cudaStream_t st, int* launches) {
    cudaError_t e = cudaMemsetAsync(b.state, 0, sizeof(TxState) * (size_t)b.n_ch, st);
    if (e != cudaSuccess) return e;
    UA3_LAUNCH(tx_init_state_kernel, (b.n_ch + 127u) / 128u, 128, 0, st, b.state, b.n_ch);
    if (launches) *launches += 1;
    return cudaGetLastError();
}

cudaError_t tx_launch_clear(const TxBuffers& b, const uint8_t* flags_dev, uint32_t first, uint32_t n, cudaStream_t st, int* launches) {
    if (!n) return cudaSuccess;
    UA3_LAUNCH(tx_clear_filters_kernel, (n + 127u) / 128u, 128, 0, st, b.state, flags_dev, first, n);
    if (launches) *launches += 1;
    return cudaGetLastError();
}

cudaError_t tx_launch_audio(const TxBuffers& b, uint32_t n_blocks, cudaStream_t st, int* launches) {
    if (!n_blocks) return cudaSuccess;
    UA3_LAUNCH(tx_audio_kernel, b.n_ch, 32, 0, st, b.mic, n_blocks, b.max_blocks * (uint32_t)kAudioBlock * 2u, b.params, b.state,
               b.n_ch, b.iq_f, b.iq_w, b.loop_out, b.max_blocks * (uint32_t)kAudioBlock);
    if (launches) *launches += 1;
    return cudaGetLastError();
}

}  // namespace ua3
