// rx.cu - sm_100a kernels for the STM32 half of the receive path, batched over channels:
//   rx_audio_kernel : processRxAudio()  (audio_processor.c:275-435) - DC filter, RF gain, IIR-lattice HPF/LPF,
//                     SSB/AM/FM demodulation, notch, S-meter, NLMS noise reduction, AGC, volume, int32 pack
//   rx_fft_kernel   : FFT_doFFT()       (fft.c:212-331) - DC filter, notch, Hamming window, 512-point radix-8
//                     complex FFT, magnitude, 2:1 bin compression, auto-range, temporal averaging
//
// Every stage here is a recurrence along time with per-channel state, so parallelism comes from
// channels: the audio kernel gives each channel a PAIR of lanes (even lane = I rail, odd lane = Q rail,
// 16 channels per warp) that run the two long lattice chains side by side and meet through one
// shuffle per sample; the FFT kernel gives each channel a whole warp (2 radix-8 butterflies per lane
// and pass, data in shared memory).  All float arithmetic is IEEE binary32 in the firmware's
// operation order; the library is compiled with --fmad=false so nothing is contracted.
#include "rx_launch.h"
#include "ua3_common.cuh"

namespace ua3 {

__constant__ float c_fft_window[kFftSize];
__constant__ float c_fft_twiddle[2 * kFftSize];
__constant__ uint16_t c_fft_colors[32];
__constant__ float c_zoom_biquad[4][20];      // zoom 2, 4, 8, 16: mag_coeffs (fft.c:72-121)
__constant__ float c_zoom_fir[4][4];          // FirZoomFFTDecimate (fft.c:123-183)

#if defined(UA3_HOST_EMU)
#define UA3_FULL_MASK 0xffffffffu
#else
#define UA3_FULL_MASK 0xffffffffu
#endif

// arm_iir_lattice_f32 for one sample (CMSIS-DSP 1.6.0 order): G[i] holds g_(N-1-i)(n-1).
template <int N>
UA3_D float lattice_step(float x, const float (&k)[N], const float (&v)[N + 1], float (&G)[N]) {
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

UA3_D int16_t frame_word(uint64_t f, int w) {
    const uint32_t hi = (uint32_t)(f >> (16 * w)) & 0xFFu, lo = (uint32_t)(f >> (16 * w + 8)) & 0xFFu;
    return (int16_t)(uint16_t)((hi << 8) | lo);        // fpga.c:295-301: (hi << 8) | lo as int16
}

// ------------------------------------------------------------------------------------------------
// processRxAudio for n_blocks consecutive 192-sample blocks of every channel.
// A warp serves 16 channels and never talks to another warp (no CTA barrier); a CTA is kRxAudioWarps such warps
// whose shared-memory slices fill the SM (6 x 37 KB), so that the kernel occupies ceil(n_ch / 96) WHOLE SMs instead
// of scattering one-warp CTAs over every SM: it is latency bound (a few warps per SM is all it can use), and packed
// this way it runs beside the next block's persistent front kernel, which simply gets that many SMs fewer.
// ------------------------------------------------------------------------------------------------
constexpr int kRxAudioWarps = 6;
constexpr int kRxBufFloats = kAudioBlock * 32;                         // s_buf[sample][lane]: lane's own rail, conflict free
constexpr int kRxWinFloats = (kLmsTaps - 1 + kSubBlock) * 16;          // NLMS input window of the I lanes
constexpr int kRxRefFloats = 2 * kSubBlock * 16;                       // lms2_reference of the I lanes
constexpr size_t kRxAudioSmemPerWarp = (size_t)(kRxBufFloats + kRxWinFloats + kRxRefFloats) * sizeof(float);
constexpr size_t kRxAudioSmemBytes = kRxAudioSmemPerWarp * kRxAudioWarps;

__global__ void __launch_bounds__(32 * kRxAudioWarps, 1)
rx_audio_kernel(const uint64_t* __restrict__ frames, uint32_t ring_mask, uint32_t frame_ch_stride, uint32_t start,
                uint32_t n_blocks, const RxParams* __restrict__ params, RxState* __restrict__ state, uint32_t n_ch,
                int32_t* __restrict__ audio_out, uint32_t out_ch_stride, float* __restrict__ cw_mag, uint32_t cw_ch_stride,
                const uint32_t* __restrict__ order) {
#if defined(UA3_HOST_EMU)
    static float s_dyn[kRxAudioSmemBytes / sizeof(float)];
#else
    extern __shared__ __align__(16) float s_dyn[];
#endif
    const int warp = threadIdx.x >> 5;
    float* s_base = s_dyn + (size_t)warp * (kRxAudioSmemPerWarp / sizeof(float));
    float (*s_buf)[32] = reinterpret_cast<float (*)[32]>(s_base);
    float (*s_win)[16] = reinterpret_cast<float (*)[16]>(s_base + kRxBufFloats);
    float (*s_ref)[16] = reinterpret_cast<float (*)[16]>(s_base + kRxBufFloats + kRxWinFloats);

    const int lane = threadIdx.x & 31, pair = lane >> 1, rail = lane & 1;
    // Channels are visited in the host-built order that groups equal (mode, DNR, notch) settings, so that the
    // 16 channels of a warp take the same branches (mode is per-channel data; without the grouping a warp with
    // mixed modes executes every demodulator for every sample).
    if ((blockIdx.x * (uint32_t)kRxAudioWarps + (uint32_t)warp) * 16u >= n_ch) return;    // whole warp beyond the bank (no barriers below)
    const uint32_t slot = (blockIdx.x * (uint32_t)kRxAudioWarps + (uint32_t)warp) * 16u + (uint32_t)pair;
    const bool live = slot < n_ch;
    const uint32_t ch = order[live ? slot : (n_ch - 1u)];   // dead pairs shadow the last slot's channel and never store
    const RxParams& P = params[ch];
    RxState& S = state[ch];

    const uint8_t mode = P.mode;
    const bool use_spec = (mode == kModeIQ || mode == kModeNFM || mode == kModeWFM || mode == kModeAM);
    const bool is_lsb = (mode == kModeLSB || mode == kModeCWL || mode == kModeDIGIL);
    const bool is_usb = (mode == kModeUSB || mode == kModeCWU || mode == kModeDIGIU);
    const bool is_fm = (mode == kModeNFM || mode == kModeWFM);
    const bool is_am = (mode == kModeAM);
    const bool do_hpf = (mode == kModeLSB || mode == kModeCWL || mode == kModeUSB || mode == kModeCWU) && P.hpf_set;
    const bool do_lpf = (is_lsb || is_usb || is_am || is_fm) && P.lpf_on;
    const bool chain = is_lsb || is_usb || is_am || is_fm;          // NOTCH/DNR/AGC/COPY tail
    const int widx = (use_spec ? 0 : 2) + (((rail == 0) != (P.iq_swap != 0)) ? 1 : 0);

    float lk[kLpfMax], lv[kLpfMax + 1], hk[kHpfStages], hv[kHpfStages + 1];
#pragma unroll
    for (int i = 0; i < kLpfMax; ++i) lk[i] = P.lpf_k[i];
#pragma unroll
    for (int i = 0; i <= kLpfMax; ++i) lv[i] = P.lpf_v[i];
#pragma unroll
    for (int i = 0; i < kHpfStages; ++i) hk[i] = P.hpf_k[i];
#pragma unroll
    for (int i = 0; i <= kHpfStages; ++i) hv[i] = P.hpf_v[i];
    float LG[kLpfMax], HG[kHpfStages];
#pragma unroll
    for (int i = 0; i < kLpfMax; ++i) LG[i] = S.lpf_g[rail][i];
#pragma unroll
    for (int i = 0; i < kHpfStages; ++i) HG[i] = S.hpf_g[rail][i];
    float dc_x = S.dc_x[rail], dc_y = S.dc_y[rail];
    const float A1 = (float)(1.0 - 0.00048828125);      // (1.0 - pow(2.0, -11.0)) (audio_filters.c:360)
    const float nb0 = P.notch[0], nb1 = P.notch[1], nb2 = P.notch[2], na1 = P.notch[3], na2 = P.notch[4];
    float nd1 = S.notch_d[0], nd2 = S.notch_d[1];
    float sm_max = S.smeter_max, sm_min = S.smeter_min;
    float fm_lpf = S.fm_lpf_prev, fm_ha = S.fm_hpf_prev_a, fm_hb = S.fm_hpf_prev_b, fm_ip = S.fm_i_prev, fm_qp = S.fm_q_prev;
    float fm_sql_avg = S.fm_sql_avg;
    uint32_t fm_sql_count = S.fm_sql_count, squelched = S.squelched;
    float agc_gain = S.agc_gain, agc_old = S.agc_gain_old;
    float w[kLmsTaps];
#pragma unroll
    for (int i = 0; i < kLmsTaps; ++i) w[i] = S.lms_w[i];
    float lms_energy = S.lms_energy, lms_x0 = S.lms_x0;
    uint32_t idx_old = S.lms_idx_old, idx_new = S.lms_idx_new;
    if (rail == 0) {
        for (int i = 0; i < kLmsTaps - 1; ++i) s_win[i][pair] = S.lms_hist[i];
        if (P.dnr_on) for (int i = 0; i < 2 * kSubBlock; ++i) s_ref[i][pair] = S.lms_ref[i];
    }

    const uint64_t* fr = frames + (size_t)ch * frame_ch_stride;

    for (uint32_t blk = 0; blk < n_blocks; ++blk) {
        const uint32_t base = start + blk * (uint32_t)kAudioBlock;
        float angle0 = 0.0f;
        // ---------------- pass 0: the lane's 192 input words -> shared memory (independent loads, many in flight) ----
#pragma unroll 8
        for (int i = 0; i < kAudioBlock; ++i)
            s_buf[i][lane] = (float)frame_word(__ldg(fr + ((base + (uint32_t)i) & ring_mask)), widx);   // int16 -> float32 (fpga.c:303-385)
        // ---------------- pass 1: per-sample chain up to NOTCH + SMETER -----------------------------
        for (int i = 0; i < kAudioBlock; ++i) {
            float x = s_buf[i][lane];
            {   // dc_filter (audio_filters.c:358-373)
                const float delta_x = x - dc_x;
                const float a1_y_prev = A1 * dc_y;
                const float y = delta_x + a1_y_prev;
                dc_x = x; dc_y = y; x = y;
            }
            x = x * P.rf_gain;                                       // audio_processor.c:305-306
            if (do_hpf) x = lattice_step<kHpfStages>(x, hk, hv, HG);
            if (do_lpf) x = lattice_step<kLpfMax>(x, lk, lv, LG);
            const float other = __shfl_xor_sync(UA3_FULL_MASK, x, 1); // the pair's other rail
            float out = x;                                           // value this lane keeps for its rail
            if (rail == 0) {
                const float I = x, Q = other;
                if (is_lsb) out = I - Q;                             // :315
                else if (is_usb) out = I + Q;                        // :328
                else if (is_am) {                                    // :338-342
                    const float s = (I * I) + (Q * Q);
                    out = (s >= 0.0f) ? sqrtf(s) : 0.0f;
                } else if (is_fm) {                                  // DemodulateFM :517-546
                    const float yy = (Q * fm_ip) - (I * fm_qp);
                    const float xx = (I * fm_ip) + (Q * fm_qp);
                    const float angle = atan2f(yy, xx);
                    if (i == 0) angle0 = angle;
                    const float a = fm_lpf + (0.05f * (angle - fm_lpf));
                    fm_lpf = a;
                    fm_qp = Q; fm_ip = I;
                    if (!squelched || !P.fm_sql_threshold) {
                        if (mode == kModeWFM) {
                            out = (angle / 3.14159265358979f) * 16384.0f;
                        } else {
                            const float b = 0.96f * ((fm_hb + a) - fm_ha);
                            fm_ha = a; fm_hb = b;
                            out = b * 30000.0f;
                        }
                    } else {
                        out = 0.0f;
                    }
                }
                if (P.notch_on && (is_lsb || is_usb || is_am)) {     // doRX_NOTCH :466-473 (df2T, 1 stage)
                    const float acc1 = nb0 * out + nd1;
                    nd1 = nb1 * out + nd2;
                    nd1 += na1 * acc1;
                    nd2 = nb2 * out;
                    nd2 += na2 * acc1;
                    out = acc1;
                }
            } else {
                if (is_am) out = x * x;                              // Q rail holds Q*Q after arm_mult_f32 (:339)
            }
            // doRX_SMETER (:491-501): running max / min over both rails
            if (out > sm_max) sm_max = out;
            if (out < sm_min) sm_min = out;
            s_buf[i][lane] = out;
        }
        if (is_fm && rail == 0) {                                    // squelch bookkeeping :548-586
            fm_sql_avg = (0.995f * fm_sql_avg) + (0.005f * sqrtf(fabsf(angle0)));
            if (fm_sql_count == 0) {
                if (fm_sql_avg > 0.7f) fm_sql_avg = 0.7f;
                const float b = fm_sql_avg * 10.0f;
                const int thr = P.fm_sql_threshold;
                if (!thr) squelched = 0;
                else if (squelched) {
                    if (b <= (float)((10 - thr) - 0.3f)) squelched = 0;
                } else {
                    if ((10.0f - thr) > 0.3f) { if (b > (float)((10 - thr) + 0.3f)) squelched = 1; }
                    else { if (b > (10.0f - (float)thr)) squelched = 1; }
                }
                fm_sql_count++;                                      // only ever reaches 1: the firmware's
                if (fm_sql_count >= 50) fm_sql_count = 0;            // increment sits inside the == 0 branch
            }
        }
        {   // both lanes of the pair share the S-meter extremes
            const float om = __shfl_xor_sync(UA3_FULL_MASK, sm_max, 1), on = __shfl_xor_sync(UA3_FULL_MASK, sm_min, 1);
            sm_max = fmaxf(sm_max, om); sm_min = fminf(sm_min, on);
        }
        __syncwarp();

        // ---------------- pass 2: doRX_DNR (:475-483) on the I rail, three 64-sample sub-blocks ------
        if (chain && P.dnr_on && rail == 0 && live) {
            for (int sb = 0; sb < kAudioBlock / kSubBlock; ++sb) {
                // arm_copy_f32(bufferIn, &lms2_reference[reference_index_new], 64)
                for (int n = 0; n < kSubBlock; ++n) {
                    const float v = s_buf[sb * kSubBlock + n][lane];
                    s_ref[idx_new + n][pair] = v;
                    s_win[kLmsTaps - 1 + n][pair] = v;
                }
                for (int n = 0; n < kSubBlock; ++n) {                // arm_lms_norm_f32
                    float win[kLmsTaps];
#pragma unroll
                    for (int t = 0; t < kLmsTaps; ++t) win[t] = s_win[n + t][pair];
                    const float in = win[kLmsTaps - 1];
                    lms_energy -= lms_x0 * lms_x0;
                    lms_energy += in * in;
                    float acc = 0.0f;
#pragma unroll
                    for (int t = 0; t < kLmsTaps; ++t) acc += win[t] * w[t];
                    s_buf[sb * kSubBlock + n][lane] = acc;           // output = prediction, in place
                    const float e = s_ref[idx_old + n][pair] - acc;
                    const float wg = (e * 0.000001f) / (lms_energy + 0.000000119209289f);
#pragma unroll
                    for (int t = 0; t < kLmsTaps; ++t) w[t] += wg * win[t];
                    lms_x0 = win[0];
                }
                for (int t = 0; t < kLmsTaps - 1; ++t) s_win[t][pair] = s_win[kSubBlock + t][pair];
                idx_old += kSubBlock;                                // noise_reduction.c:33-36
                if (idx_old >= 2 * kSubBlock) idx_old = 0;
                idx_new = idx_old + kSubBlock;
                if (idx_new >= 2 * kSubBlock) idx_new = 0;
            }
        }
        // ---------------- pass 3: DoAGC (agc.c:21-67) on the I rail ---------------------------------
        if (chain && rail == 0) {
            float amax = s_buf[0][lane];                             // arm_max_f32: signed maximum
            for (int i = 1; i < kAudioBlock; ++i) { const float v = s_buf[i][lane]; if (amax < v) amax = v; }
            if (amax == 0.0f) amax = 0.001f;
            const float target = 7000.0f / amax;
            if (target > agc_gain) {
                float step = (target - agc_gain) / P.agc_step_up;
                if (step > 1.0f) step = 1.0f;
                agc_gain += step;
            } else {
                agc_gain -= (agc_gain - target) / P.agc_step_down;
            }
            if (agc_gain < 0.0f) agc_gain = 0.0f;
            if ((agc_gain * amax) > 10000.0f) agc_gain = target;
            if (!P.agc_on || mode == kModeDIGIL || mode == kModeDIGIU) agc_gain = 1.0f;
            if (agc_old != agc_gain) {
                float gstep = 0.0f;
                if (agc_old > agc_gain) gstep = -(agc_old - agc_gain) / 192.0f;
                if (agc_old < agc_gain) gstep = (agc_gain - agc_old) / 192.0f;
                for (int i = 0; i < kAudioBlock; ++i) {
                    agc_old += gstep;
                    s_buf[i][lane] = s_buf[i][lane] * agc_old;
                }
            } else {
                for (int i = 0; i < kAudioBlock; ++i) s_buf[i][lane] = s_buf[i][lane] * agc_gain;
            }
        }
        // ---------------- doCW_Decode (:437-443): Goertzel front end of CWDecoder_Process (cw_decoder.c:56-66) ------
        if (live && rail == 0) {
            float mag = 0.0f;
            if (P.cw_on) {
                float Q1 = 0.0f, Q2 = 0.0f;
                const float coeff = P.cw_coeff;
                for (int i = 0; i < kAudioBlock; ++i) {
                    const float Q0 = ((coeff * Q1) - Q2) + s_buf[i][lane];
                    Q2 = Q1; Q1 = Q0;
                }
                const float msq = ((Q1 * Q1) + (Q2 * Q2)) - ((Q1 * Q2) * coeff);
                mag = sqrtf(msq);
            }
            cw_mag[(size_t)ch * cw_ch_stride + blk] = mag;
        }
        __syncwarp();
        // ---------------- output: COPYCHANNEL, volume, float -> int32, L/R interleave (:365-394) ----
        if (live) {
            int32_t* dst = audio_out + (size_t)ch * out_ch_stride + (size_t)blk * (2 * kAudioBlock);
            const int src_lane = chain ? (lane & ~1) : lane;         // doRX_COPYCHANNEL: Q <- I
            for (int i = 0; i < kAudioBlock; ++i) {
                float v = s_buf[i][src_lane];
                v = P.mute ? (v * 0.0f) : (v * P.volume);
                dst[2 * i + rail] = (int32_t)v;                      // C cast: truncation toward zero
            }
        }
        __syncwarp();
    }

    // ---------------- state write-back ------------------------------------------------------------
    if (live) {
#pragma unroll
        for (int i = 0; i < kLpfMax; ++i) S.lpf_g[rail][i] = LG[i];
#pragma unroll
        for (int i = 0; i < kHpfStages; ++i) S.hpf_g[rail][i] = HG[i];
        S.dc_x[rail] = dc_x; S.dc_y[rail] = dc_y;
        if (rail == 0) {
            S.notch_d[0] = nd1; S.notch_d[1] = nd2;
            S.smeter_max = sm_max; S.smeter_min = sm_min;
            S.fm_lpf_prev = fm_lpf; S.fm_hpf_prev_a = fm_ha; S.fm_hpf_prev_b = fm_hb; S.fm_i_prev = fm_ip; S.fm_q_prev = fm_qp;
            S.fm_sql_avg = fm_sql_avg; S.fm_sql_count = fm_sql_count; S.squelched = squelched;
            S.agc_gain = agc_gain; S.agc_gain_old = agc_old;
#pragma unroll
            for (int i = 0; i < kLmsTaps; ++i) S.lms_w[i] = w[i];
            for (int i = 0; i < kLmsTaps - 1; ++i) S.lms_hist[i] = s_win[i][pair];
            if (P.dnr_on) for (int i = 0; i < 2 * kSubBlock; ++i) S.lms_ref[i] = s_ref[i][pair];
            S.lms_energy = lms_energy; S.lms_x0 = lms_x0;
            S.lms_idx_old = idx_old; S.lms_idx_new = idx_new;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// FFT_doFFT for n_frames consecutive 512-sample frames of every channel; one warp per channel.
// ------------------------------------------------------------------------------------------------
struct C8 { float re[8], im[8]; };

// One radix-8 decimation-in-frequency butterfly in the operation order of CMSIS-DSP's
// arm_radix8_butterfly_f32; outputs are left un-twiddled in natural butterfly order
// (index m = position i1 + m*n2), the caller applies the twiddles.
UA3_D void radix8(C8& z) {
    const float C81 = 0.70710678118f;
    float r1 = z.re[0] + z.re[4], r5 = z.re[0] - z.re[4];
    float r2 = z.re[1] + z.re[5], r6 = z.re[1] - z.re[5];
    float r3 = z.re[2] + z.re[6], r7 = z.re[2] - z.re[6];
    float r4 = z.re[3] + z.re[7], r8 = z.re[3] - z.re[7];
    float t1 = r1 - r3; r1 = r1 + r3; r3 = r2 - r4; r2 = r2 + r4;
    const float o0re = r1 + r2;
    const float o4re = r1 - r2;
    float s1 = z.im[0] + z.im[4], s5 = z.im[0] - z.im[4];
    float s2 = z.im[1] + z.im[5], s6 = z.im[1] - z.im[5];
    float s3 = z.im[2] + z.im[6], s7 = z.im[2] - z.im[6];
    float s4 = z.im[3] + z.im[7], s8 = z.im[3] - z.im[7];
    float t2 = s1 - s3; s1 = s1 + s3; s3 = s2 - s4; s2 = s2 + s4;
    const float o2re = t1 + s3, o6re = t1 - s3;
    const float o0im = s1 + s2, o4im = s1 - s2;
    const float o2im = t2 - r3, o6im = t2 + r3;
    r1 = (r6 - r8) * C81; r6 = (r6 + r8) * C81;
    s1 = (s6 - s8) * C81; s6 = (s6 + s8) * C81;
    t1 = r5 - r1; r5 = r5 + r1; r8 = r7 - r6; r7 = r7 + r6;
    t2 = s5 - s1; s5 = s5 + s1; s8 = s7 - s6; s7 = s7 + s6;
    z.re[0] = o0re; z.im[0] = o0im;
    z.re[1] = r5 + s7; z.im[1] = s5 - r7;
    z.re[2] = o2re; z.im[2] = o2im;
    z.re[3] = t1 - s8; z.im[3] = t2 + r8;
    z.re[4] = o4re; z.im[4] = o4im;
    z.re[5] = t1 + s8; z.im[5] = t2 - r8;
    z.re[6] = o6re; z.im[6] = o6im;
    z.re[7] = r5 - s7; z.im[7] = s5 + r7;
}

__global__ void __launch_bounds__(32)
rx_fft_kernel(const uint64_t* __restrict__ frames, uint32_t ring_mask, uint32_t frame_ch_stride, uint32_t start,
              uint32_t n_frames, const RxParams* __restrict__ params, RxState* __restrict__ state, uint32_t n_ch,
              float* __restrict__ spectra, uint32_t spec_ch_stride, uint16_t* __restrict__ waterfall,
              uint16_t* __restrict__ wtf_hist, uint32_t* __restrict__ wtf_head, int32_t* __restrict__ wtf_pending_hz) {
    __shared__ float s_re[kFftSize], s_im[kFftSize];
    __shared__ float s_dec[2][kFftSize / 2];                           // decimated samples of the ZoomFFT branch
    const int lane = threadIdx.x & 31;
    const uint32_t ch = blockIdx.x;
    if (ch >= n_ch) return;
    const RxParams& P = params[ch];
    RxState& S = state[ch];
    if (!P.fft_enabled) return;                                       // fft.c:214
    const uint64_t* fr = frames + (size_t)ch * frame_ch_stride;
    const float A1 = (float)(1.0 - 0.00048828125);
    const int swap = P.iq_swap ? 1 : 0;
    const int zoom = P.fft_zoom > 1 ? P.fft_zoom : 1;
    const int zsel = zoom == 2 ? 0 : (zoom == 4 ? 1 : (zoom == 8 ? 2 : 3));

    for (uint32_t fi = 0; fi < n_frames; ++fi) {
        const uint32_t base = start + fi * (uint32_t)kFftSize;
        // all lanes: the frame's SPEC words -> shared memory (FFT_buff filling of fpga.c:305-341, with the I/Q swap)
        for (int i = lane; i < kFftSize; i += 32) {
            const uint64_t f = __ldg(fr + ((base + (uint32_t)i) & ring_mask));
            s_re[i] = (float)frame_word(f, swap ? 0 : 1);
            s_im[i] = (float)frame_word(f, swap ? 1 : 0);
        }
        __syncwarp();
        // lanes 0/1: dc_filter(FFTInput_I, 512, 4) / dc_filter(FFTInput_Q, 512, 5) and the optional notch (:225-233)
        if (lane < 2) {
            const int rail = lane;                                    // 0 = I, 1 = Q
            float dx = S.dc_x[4 + rail], dy = S.dc_y[4 + rail];
            float d1 = S.notch_fft_d[rail][0], d2 = S.notch_fft_d[rail][1];
            const float b0 = P.notch[0], b1 = P.notch[1], b2 = P.notch[2], a1 = P.notch[3], a2 = P.notch[4];
            float* dst = rail == 0 ? s_re : s_im;
            for (int i = 0; i < kFftSize; ++i) {
                float x = dst[i];
                const float delta_x = x - dx;
                const float a1y = A1 * dy;
                const float y = delta_x + a1y;
                dx = x; dy = y; x = y;
                if (P.notch_on) {
                    const float acc1 = b0 * x + d1;
                    d1 = b1 * x + d2;
                    d1 += a1 * acc1;
                    d2 = b2 * x;
                    d2 += a2 * acc1;
                    x = acc1;
                }
                dst[i] = x;
            }
            S.dc_x[4 + rail] = dx; S.dc_y[4 + rail] = dy;
            S.notch_fft_d[rail][0] = d1; S.notch_fft_d[rail][1] = d2;
            if (zoom > 1) {
                // arm_biquad_cascade_df1_f32, 4 stages (fft.c:239-240), then arm_fir_decimate_f32, 4 taps, M = zoom (:242-243)
                float st[4][4];
#pragma unroll
                for (int sg = 0; sg < 4; ++sg)
#pragma unroll
                    for (int q = 0; q < 4; ++q) st[sg][q] = S.zoom_biquad[rail][sg][q];
                float h0 = S.zoom_fir[rail][0], h1 = S.zoom_fir[rail][1], h2 = S.zoom_fir[rail][2];
                const float* bq = c_zoom_biquad[zsel];
                const float f0 = c_zoom_fir[zsel][0], f1 = c_zoom_fir[zsel][1], f2 = c_zoom_fir[zsel][2], f3 = c_zoom_fir[zsel][3];
                for (int i = 0; i < kFftSize; ++i) {
                    float x = dst[i];
#pragma unroll
                    for (int sg = 0; sg < 4; ++sg) {
                        const float acc = ((((bq[sg * 5] * x) + (bq[sg * 5 + 1] * st[sg][0])) + (bq[sg * 5 + 2] * st[sg][1])) +
                                           (bq[sg * 5 + 3] * st[sg][2])) + (bq[sg * 5 + 4] * st[sg][3]);
                        st[sg][1] = st[sg][0]; st[sg][0] = x; st[sg][3] = st[sg][2]; st[sg][2] = acc;
                        x = acc;
                    }
                    // decimator: output o is computed right after the FIRST sample of its group of `zoom` inputs has been
                    // shifted in (arm_fir_decimate_f32 copies M samples, then reads the window that starts M back)
                    if (i % zoom == 0) {
                        float sum0 = 0.0f;
                        sum0 += h0 * f0; sum0 += h1 * f1; sum0 += h2 * f2; sum0 += x * f3;
                        s_dec[rail][i / zoom] = sum0;
                    }
                    h0 = h1; h1 = h2; h2 = x;
                }
#pragma unroll
                for (int sg = 0; sg < 4; ++sg)
#pragma unroll
                    for (int q = 0; q < 4; ++q) S.zoom_biquad[rail][sg][q] = st[sg][q];
                S.zoom_fir[rail][0] = h0; S.zoom_fir[rail][1] = h1; S.zoom_fir[rail][2] = h2;
            }
        }
        __syncwarp();
        if (zoom > 1) {
            // slide FFTInput_ZOOMFFT left by zoomed_width and append the new decimated samples (fft.c:245-260)
            const int zw = kFftSize / zoom;
            for (int i = lane; i < kFftSize; i += 32) {
                if (i < kFftSize - zw) { s_re[i] = S.zoom_buf[2 * (i + zw)]; s_im[i] = S.zoom_buf[2 * (i + zw) + 1]; }
                else { s_re[i] = s_dec[0][i - (kFftSize - zw)]; s_im[i] = s_dec[1][i - (kFftSize - zw)]; }
            }
            __syncwarp();
            for (int i = lane; i < kFftSize; i += 32) { S.zoom_buf[2 * i] = s_re[i]; S.zoom_buf[2 * i + 1] = s_im[i]; }
            __syncwarp();
        }
        // Hamming window (:275-286)
        for (int i = lane; i < kFftSize; i += 32) {
            const float wm = c_fft_window[i];
            s_re[i] = wm * s_re[i];
            s_im[i] = wm * s_im[i];
        }
        __syncwarp();
        // arm_cfft_f32(len512): three radix-8 DIF passes, n2 = 64, 8, 1; twiddle step 1, 8, (none)
#pragma unroll 1
        for (int pass = 0; pass < 3; ++pass) {
            const int n2 = pass == 0 ? 64 : (pass == 1 ? 8 : 1);
            const int n1 = n2 * 8;
            const int tmod = pass == 0 ? 1 : 8;
            for (int b = lane; b < 64; b += 32) {
                const int j = b % n2, g = b / n2;                     // butterfly j of group g
                const int i1 = g * n1 + j;
                C8 z;
#pragma unroll
                for (int m = 0; m < 8; ++m) { z.re[m] = s_re[i1 + m * n2]; z.im[m] = s_im[i1 + m * n2]; }
                radix8(z);
                s_re[i1] = z.re[0]; s_im[i1] = z.im[0];
                if (pass < 2 && j > 0) {
                    // outputs 1..7 are rotated by exp(-j*2*pi*(j*tmod*q)/512); CMSIS pairs output position m with
                    // twiddle index: pos1<-co2(ia1) pos2<-co3(ia2) pos3<-co4(ia3) pos4<-co5(ia4) pos5<-co6(ia5) pos6<-co7(ia6) pos7<-co8(ia7)
                    const int id = j * tmod;
#pragma unroll
                    for (int m = 1; m < 8; ++m) {
                        const float co = c_fft_twiddle[2 * (id * m)], si = c_fft_twiddle[2 * (id * m) + 1];
                        const float p1 = co * z.re[m], p2 = si * z.im[m], p3 = co * z.im[m], p4 = si * z.re[m];
                        s_re[i1 + m * n2] = p1 + p2;
                        s_im[i1 + m * n2] = p3 - p4;
                    }
                } else {
#pragma unroll
                    for (int m = 1; m < 8; ++m) { s_re[i1 + m * n2] = z.re[m]; s_im[i1 + m * n2] = z.im[m]; }
                }
            }
            __syncwarp();
        }
        // bit (digit) reversal, arm_cmplx_mag_f32, 2:1 compression (:288-302); natural bin r lives at digit-reversed slot
        float comp[kFftBins / 32];
        float lmax = 0.0f;
#pragma unroll
        for (int q = 0; q < kFftBins / 32; ++q) {
            const int bin = q * 32 + lane;                            // output bin, natural order
            float acc = 0.0f;
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                const int r = 2 * bin + c;
                const int slot = ((r & 7) << 6) | (r & 0x38) | (r >> 6);
                const float re = s_re[slot], im = s_im[slot];
                const float sq = (re * re) + (im * im);
                acc += (sq >= 0.0f) ? sqrtf(sq) : 0.0f;
            }
            comp[q] = acc / 2.0f;
        }
        {   // FFTInput[0] = FFTInput[1] (:304)
            const float b1v = __shfl_sync(UA3_FULL_MASK, comp[0], 1);
            if (lane == 0) comp[0] = b1v;
        }
        // arm_max_f32 over the 256 bins (:307)
        lmax = comp[0];
#pragma unroll
        for (int q = 1; q < kFftBins / 32; ++q) lmax = fmaxf(lmax, comp[q]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) lmax = fmaxf(lmax, __shfl_xor_sync(UA3_FULL_MASK, lmax, o));
        // auto-range (:308-318), identical scalar sequence on every lane
        float mv = S.fft_max_value;
        const uint32_t errs = S.fft_max_errors;
        const float diff = (lmax - mv) / 10.0f;
        if (errs >= 6 && diff > 0) mv += diff;
        else if (errs <= 1 && diff < 0 && diff < -10.0f) mv += diff;
        else if (errs <= 1 && mv > 10.0f) mv -= 10.0f;
        else if (errs <= 1 && diff < 0 && diff < -1.0f) mv += diff;
        else if (errs <= 1 && mv > 1.0f) mv -= 1.0f;
        if (mv < 100.0f) mv = 100.0f;
        if (P.mode == kModeLoopback) mv = 60000.0f;
        const float inv = 1.0f / mv;                                   // arm_scale_f32(FFTInput, 1.0f / maxValueFFT, ...)
        // temporal averaging into FFTOutput_mean (:324-328) and the display pass's overflow count (:369-373)
        float* out = spectra + (size_t)ch * spec_ch_stride + (size_t)fi * kFftBins;
        uint16_t* wf = waterfall + (size_t)ch * spec_ch_stride + (size_t)fi * kFftBins;
        uint16_t* hist = wtf_hist + (size_t)ch * kWtfRows * kFftBins;
#pragma unroll
        for (int q = 0; q < kFftBins / 32; ++q) {
            const int bin = q * 32 + lane;
            const float x = comp[q] * inv;
            float m = S.fft_mean[bin];
            if (m < x) m += (x - m) / P.fft_averaging;
            else m -= (m - x) / P.fft_averaging;
            S.fft_mean[bin] = m;
            out[bin] = m;                                             // FFTOutput_mean as FFT_doFFT() leaves it
        }
        __syncwarp();
        uint32_t head = wtf_head[ch];
        // ---- FFT_printFFT() part that feeds back into the numbers (fft.c:346-379) ----
        // a retune since the last print shifts the averages and every stored row sideways (FFT_moveWaterfall, :458-504)
        const int32_t hz = wtf_pending_hz[ch];
        if (hz != 0) {
            const int d = (int)(int16_t)(((int)(int16_t)hz / 187) * zoom);    // FFT_HZ_IN_PIXEL = 48000 / 256 in integers (fft.h:23)
            if (d != 0) {
                if (lane == 0) {
                    // FFTOutput_mean is shifted IN PLACE with wrap-around, so wrapped bins read already-moved values
                    if (d > 0) {
                        for (int x = 0; x < kFftBins; ++x) S.fft_mean[x] = S.fft_mean[(x + d) & (kFftBins - 1)];
                    } else {
                        for (int x = kFftBins - 1; x >= 0; --x) S.fft_mean[x] = S.fft_mean[(x + d) & (kFftBins - 1)];
                    }
                }
                for (int y = 0; y < kWtfRows; ++y) {                 // rows move without wrap, zero filled
                    uint16_t* row = hist + (size_t)y * kFftBins;
                    uint16_t v[kFftBins / 32];
#pragma unroll
                    for (int q = 0; q < kFftBins / 32; ++q) {
                        const int nx = q * 32 + lane + d;
                        v[q] = (nx >= 0 && nx < kFftBins) ? row[nx] : (uint16_t)0;
                    }
                    __syncwarp();
#pragma unroll
                    for (int q = 0; q < kFftBins / 32; ++q) row[q * 32 + lane] = v[q];
                }
            }
            __syncwarp();
            if (lane == 0) wtf_pending_hz[ch] = 0;
        }
        __syncwarp();
        // rows move down by one (:353-358), then row 0 is drawn from the averages (:361-379)
        head = (head + kWtfRows - 1) % kWtfRows;
        uint16_t* row0 = hist + (size_t)head * kFftBins;
        uint32_t nerr = 0;
#pragma unroll
        for (int q = 0; q < kFftBins / 32; ++q) {
            const int bin = q * 32 + lane;
            const float m = S.fft_mean[bin];
            // height = (uint16_t)(mean * FFT_MAX_HEIGHT); if (height > FFT_MAX_HEIGHT - 1) maxValueErrors++
            const uint32_t height = (uint32_t)(int32_t)(m * 30.0f) & 0xFFFFu;
            nerr += (height > 29u) ? 1u : 0u;
            // colour of the column, stored fft-shifted
            const uint16_t col = c_fft_colors[height > 29u ? 30u : height];
            wf[bin ^ 128] = col;
            row0[bin ^ 128] = col;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) nerr += __shfl_xor_sync(UA3_FULL_MASK, nerr, o);
        __syncwarp();
        if (lane == 0) wtf_head[ch] = head;
        if (lane == 0) { S.fft_max_value = mv; S.fft_max_errors = nerr; }
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------------------------
__global__ void rx_clear_filters_kernel(RxState* __restrict__ state, const uint8_t* __restrict__ flags, uint32_t first,
                                        uint32_t n) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    RxState& S = state[first + i];
    const uint8_t f = flags[i];
    if (f & 1) for (int r = 0; r < 2; ++r) for (int k = 0; k < kLpfMax; ++k) S.lpf_g[r][k] = 0.0f;
    if (f & 2) for (int r = 0; r < 2; ++r) for (int k = 0; k < kHpfStages; ++k) S.hpf_g[r][k] = 0.0f;
    if (f & 4) {                                   // FFT_Init() on a zoom change: biquad and decimator states cleared (fft.c:189-207)
        for (int r = 0; r < 2; ++r) {
            for (int sg = 0; sg < 4; ++sg) for (int q = 0; q < 4; ++q) S.zoom_biquad[r][sg][q] = 0.0f;
            for (int q = 0; q < 3; ++q) S.zoom_fir[r][q] = 0.0f;
        }
    }
}

// power-on values of the firmware's statics: everything zero except AGC_need_gain_old = 1.0f (agc.c:12)
__global__ void rx_init_state_kernel(RxState* __restrict__ state, uint32_t n) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) state[i].agc_gain_old = 1.0f;
}

cudaError_t rx_launch_init_state(const RxBuffers& b, cudaStream_t st, int* launches) {
    cudaError_t e = cudaMemsetAsync(b.state, 0, sizeof(RxState) * (size_t)b.n_ch, st);
    if (e != cudaSuccess) return e;
    e = cudaMemsetAsync(b.wtf_hist, 0, sizeof(uint16_t) * (size_t)b.n_ch * kWtfRows * kFftBins, st);
    if (e != cudaSuccess) return e;
    e = cudaMemsetAsync(b.wtf_head, 0, sizeof(uint32_t) * (size_t)b.n_ch, st);
    if (e != cudaSuccess) return e;
    e = cudaMemsetAsync(b.wtf_pending_hz, 0, sizeof(int32_t) * (size_t)b.n_ch, st);
    if (e != cudaSuccess) return e;
    UA3_LAUNCH(rx_init_state_kernel, (b.n_ch + 127u) / 128u, 128, 0, st, b.state, b.n_ch);
    if (launches) *launches += 1;
    return cudaGetLastError();
}

cudaError_t rx_upload_constants(const float* window, const float* twiddle, const uint16_t* colors, const float* zoom_biquad,
                                const float* zoom_fir) {
    cudaError_t e = cudaSuccess;
#if !defined(UA3_HOST_EMU)
    e = cudaFuncSetAttribute(rx_audio_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kRxAudioSmemBytes);
    if (e != cudaSuccess) return e;
#endif
    e = cudaMemcpyToSymbol(c_fft_window, window, sizeof(float) * kFftSize);
    if (e != cudaSuccess) return e;
    e = cudaMemcpyToSymbol(c_zoom_biquad, zoom_biquad, sizeof(float) * 80);
    if (e != cudaSuccess) return e;
    e = cudaMemcpyToSymbol(c_zoom_fir, zoom_fir, sizeof(float) * 16);
    if (e != cudaSuccess) return e;
    e = cudaMemcpyToSymbol(c_fft_colors, colors, sizeof(uint16_t) * 32);
    if (e != cudaSuccess) return e;
    return cudaMemcpyToSymbol(c_fft_twiddle, twiddle, sizeof(float) * 2 * kFftSize);
}

// ------------------------------------------------------------------------------------------------
// ADC_MIN / ADC_MAX tracking (stm32_interface.v:384-397) + samples at the rails, one reduction per ADC block
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
adc_stats_kernel(const int16_t* __restrict__ adc, uint32_t n, int32_t* __restrict__ stats /* min, max, n_rail */) {
    int32_t mn = 32767, mx = -32768;
    uint32_t rail = 0;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int32_t v = adc[i];
        mn = min(mn, v); mx = max(mx, v);
        rail += (v <= -2048 || v >= 2047) ? 1u : 0u;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        mn = min(mn, __shfl_xor_sync(0xffffffffu, mn, o));
        mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        rail += __shfl_xor_sync(0xffffffffu, rail, o);
    }
    if ((threadIdx.x & 31) == 0) {
        atomicMin(&stats[0], mn);
        atomicMax(&stats[1], mx);
        atomicAdd(reinterpret_cast<uint32_t*>(&stats[2]), rail);
    }
}

cudaError_t adc_stats_launch(const int16_t* adc, uint32_t n, int32_t* stats, int sm_count, cudaStream_t st, int* launches) {
    if (!n) return cudaSuccess;
    const uint32_t grid = (uint32_t)min((uint64_t)(n + 255u) / 256u, (uint64_t)sm_count * 8);
    UA3_LAUNCH(adc_stats_kernel, grid, 256, 0, st, adc, n, stats);
    if (launches) *launches += 1;
    return cudaGetLastError();
}

// USB-audio packing of processRxAudio (audio_processor.c:415-432): int16 = (int32 sample) * (1.0f / Volume * 100.0f)
__global__ void __launch_bounds__(256)
rx_usb_pack_kernel(const int32_t* __restrict__ audio, uint32_t audio_ch_stride, uint32_t n_words, const float* __restrict__ undo,
                   int16_t* __restrict__ out) {
    const uint32_t ch = blockIdx.y;
    const float k = undo[ch];
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_words; i += gridDim.x * blockDim.x)
        out[(size_t)ch * n_words + i] = (int16_t)(int32_t)((float)audio[(size_t)ch * audio_ch_stride + i] * k);
}

cudaError_t rx_launch_usb_pack(const RxBuffers& b, uint32_t n_blocks, const float* undo_dev, int16_t* out_dev, cudaStream_t st,
                               int* launches) {
    if (!n_blocks) return cudaSuccess;
    const uint32_t n_words = n_blocks * 2u * (uint32_t)kAudioBlock;
    UA3_LAUNCH(rx_usb_pack_kernel, dim3((n_words + 255u) / 256u, b.n_ch), 256, 0, st, b.audio_out, b.audio_ch_stride, n_words, undo_dev,
               out_dev);
    if (launches) *launches += 1;
    return cudaGetLastError();
}

cudaError_t rx_launch_audio(const RxBuffers& b, uint32_t start, uint32_t n_blocks, cudaStream_t st, int* launches) {
    if (!n_blocks) return cudaSuccess;
    const uint32_t per_cta = 16u * (uint32_t)kRxAudioWarps;
    UA3_LAUNCH(rx_audio_kernel, (b.n_ch + per_cta - 1u) / per_cta, 32 * kRxAudioWarps, kRxAudioSmemBytes, st, b.frames, b.ring_mask, b.frame_ch_stride, start, n_blocks,
               b.params, b.state, b.n_ch, b.audio_out, b.audio_ch_stride, b.cw_mag, b.max_audio_blocks, b.order);
    if (launches) *launches += 1;
    return cudaGetLastError();
}

cudaError_t rx_launch_fft(const RxBuffers& b, uint32_t start, uint32_t n_frames, cudaStream_t st, int* launches) {
    if (!n_frames) return cudaSuccess;
    UA3_LAUNCH(rx_fft_kernel, b.n_ch, 32, 0, st, b.frames, b.ring_mask, b.frame_ch_stride, start, n_frames, b.params,
               b.state, b.n_ch, b.spectra, b.spec_ch_stride, b.waterfall, b.wtf_hist, b.wtf_head, b.wtf_pending_hz);
    if (launches) *launches += 1;
    return cudaGetLastError();
}

cudaError_t rx_launch_clear(const RxBuffers& b, const uint8_t* flags_dev, uint32_t first, uint32_t n, cudaStream_t st,
                            int* launches) {
    if (!n) return cudaSuccess;
    UA3_LAUNCH(rx_clear_filters_kernel, (n + 127u) / 128u, 128, 0, st, b.state, flags_dev, first, n);
    if (launches) *launches += 1;
    return cudaGetLastError();
}

}  // namespace ua3
