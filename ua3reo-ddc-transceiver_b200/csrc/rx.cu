// rx.cu - sm_100a kernels for the STM32 half of the receive path, batched over channels:
//   rx_filter_kernel, rx_post_kernel : processRxAudio()  (audio_processor.c:275-435) - DC filter, RF gain, IIR-lattice
//                     HPF/LPF, SSB/AM/FM demodulation, notch, S-meter | NLMS noise reduction, AGC, volume, int32 pack
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
// processRxAudio for n_blocks consecutive 192-sample blocks of every channel, as TWO kernels with different lane
// mappings, everything in registers, joined by an L2-resident scratch in [plane][sample][slot] order (coalesced for
// both mappings):
//   rx_filter_kernel : lane PAIR per channel (even lane = I rail, odd lane = Q rail; 16 channels per warp).
//                      dc_filter, RF gain, lattice HPF + LPF per rail, the demodulator (one shuffle joins the rails),
//                      notch, S-meter.  Writes the block buffers the firmware leaves in FPGA_Audio_Buffer_I/Q_tmp after
//                      doRX_NOTCH: plane 0 = I lane, plane 1 = Q lane.
//   rx_post_kernel   : ONE lane per channel (32 channels per warp) - the stages that only exist on the I buffer:
//                      doRX_DNR (NLMS, window and weights in registers), DoAGC (block maximum, gain law, 192-step
//                      ramp), the CW decoder's Goertzel, volume / mute, float -> int32 and the L/R interleave; the
//                      [channel][sample] output rows leave through a shared-memory transposition, 256 B per store.
// Splitting by mapping instead of running everything on lane pairs halves the instructions of the second half, and no
// stage keeps more than its own state live: 2 x 16 warps fit an SM where the fused kernel (254 registers, 37 KB of
// shared memory per warp) fitted 6, which is what turns a latency-bound chain into issue-bound throughput.
// Channels are visited in the host-built order that groups equal (mode, DNR, notch) settings, so that the lanes of a
// warp take the same branches.  All arithmetic is IEEE binary32 in the firmware's operation order (--fmad=false).
// ------------------------------------------------------------------------------------------------
#ifndef UA3_RX_KPRE
#define UA3_RX_KPRE 4      /* measured: 8 words in flight cost 20 registers and spills in rx_audio_kernel, 0.63 vs 0.54 ms at 1024 channels */
#endif
#ifndef UA3_RX_FUSED_MINB
#define UA3_RX_FUSED_MINB 2
#endif
constexpr int kRxFiltWarps = 4;
constexpr int kRxPostWarps = 4;
constexpr int kRxTile = 32;                                  // samples per output transposition tile
constexpr int kRxTilePitch = kRxTile + 1;                    // int2 words per channel row (+1: conflict-free column writes)

struct RxScratch {
    float* amax;             // [max_audio_blocks][n_slot_pad] signed block maximum of plane 0 (DoAGC's arm_max_f32 when the DNR is off)
    float* base;             // [3 planes][max_samples][n_slot_pad]
    uint32_t max_samples;    // max_audio_blocks * 192
    uint32_t n_slot_pad;     // channels rounded up to 32
    __host__ __device__ size_t at(uint32_t plane, uint32_t sample, uint32_t slot) const {
        return ((size_t)plane * max_samples + sample) * n_slot_pad + slot;
    }
};

// Body of rx_filter_kernel for one warp.  HPF / LPF / FM say whether ANY channel of the warp needs that stage (the host
// groups equal settings, so usually all or none do): a stage the warp needs runs branch-free for every lane and lanes that
// do not want it keep their input (and their state is not written back), a stage nobody needs is not compiled in.  Without
// branches the sample loop is one basic block, which lets the scheduler interleave the two halves of the software
// pipeline - sample n+1 goes through dc_filter / gain / HPF while sample n goes through LPF / demodulator / notch - and
// roughly halves the dependent-issue stalls of this otherwise strictly sequential chain.
// What the two halves of processRxAudio tell each other when they run as warps of ONE CTA (rx_audio_kernel): a producer
// warp publishes the number of blocks whose buffers it has completely written, a consumer warp waits for it.
struct RxBlockSync {
    volatile uint32_t* done;         // shared-memory progress word of this producer warp, or nullptr (separate kernels)
    UA3_D void publish(uint32_t blocks_done, int lane) const {
#if !defined(UA3_HOST_EMU)
        if (!done) return;
        __threadfence_block();       // this lane's scratch stores before the flag ...
        __syncwarp();                // ... for every lane of the warp
        if (lane == 0) *done = blocks_done;
#else
        (void)blocks_done; (void)lane;
#endif
    }
    UA3_D void wait(uint32_t blocks_needed) const {
#if !defined(UA3_HOST_EMU)
        if (!done) return;
        while (*done < blocks_needed) __nanosleep(256);
        __threadfence_block();
#else
        (void)blocks_needed;
#endif
    }
};

template <bool HPF, bool LPF, bool FM>
UA3_D void rx_filter_body(const uint64_t* __restrict__ fr, uint32_t ring_mask, uint32_t start, uint32_t n_blocks,
                          const RxParams& P, RxState& S, int rail, int lane, bool live, float* __restrict__ out_plane, size_t out_step,
                          float* __restrict__ amax_out, uint32_t amax_step, RxBlockSync sync) {
    const uint8_t mode = P.mode;
    const bool use_spec = (mode == kModeIQ || mode == kModeNFM || mode == kModeWFM || mode == kModeAM);
    const bool is_lsb = (mode == kModeLSB || mode == kModeCWL || mode == kModeDIGIL);
    const bool is_usb = (mode == kModeUSB || mode == kModeCWU || mode == kModeDIGIU);
    const bool is_fm = (mode == kModeNFM || mode == kModeWFM);
    const bool is_wfm = (mode == kModeWFM);
    const bool is_am = (mode == kModeAM);
    const bool do_hpf = (mode == kModeLSB || mode == kModeCWL || mode == kModeUSB || mode == kModeCWU) && P.hpf_set;
    const bool do_lpf = (is_lsb || is_usb || is_am || is_fm) && P.lpf_on;
    const bool do_notch = P.notch_on && (is_lsb || is_usb || is_am);
    const bool sql_on = P.fm_sql_threshold != 0;
    const int sql_thr = P.fm_sql_threshold;
    const float rf_gain = P.rf_gain;
    const int widx = (use_spec ? 0 : 2) + (((rail == 0) != (P.iq_swap != 0)) ? 1 : 0);

    float lk[kLpfMax], lv[kLpfMax + 1], hk[kHpfStages], hv[kHpfStages + 1];
    float LG[kLpfMax], HG[kHpfStages];
    if (LPF) {
#pragma unroll
        for (int i = 0; i < kLpfMax; ++i) { lk[i] = P.lpf_k[i]; LG[i] = S.lpf_g[rail][i]; }
#pragma unroll
        for (int i = 0; i <= kLpfMax; ++i) lv[i] = P.lpf_v[i];
    }
    if (HPF) {
#pragma unroll
        for (int i = 0; i < kHpfStages; ++i) { hk[i] = P.hpf_k[i]; HG[i] = S.hpf_g[rail][i]; }
#pragma unroll
        for (int i = 0; i <= kHpfStages; ++i) hv[i] = P.hpf_v[i];
    }
    float dc_x = S.dc_x[rail], dc_y = S.dc_y[rail];
    const float A1 = (float)(1.0 - 0.00048828125);      // (1.0 - pow(2.0, -11.0)) (audio_filters.c:360)
    const float nb0 = P.notch[0], nb1 = P.notch[1], nb2 = P.notch[2], na1 = P.notch[3], na2 = P.notch[4];
    float nd1 = S.notch_d[0], nd2 = S.notch_d[1];
    float sm_max = S.smeter_max, sm_min = S.smeter_min;
    float fm_lpf = S.fm_lpf_prev, fm_ha = S.fm_hpf_prev_a, fm_hb = S.fm_hpf_prev_b, fm_ip = S.fm_i_prev, fm_qp = S.fm_q_prev;
    float fm_sql_avg = S.fm_sql_avg;
    uint32_t fm_sql_count = S.fm_sql_count, squelched = S.squelched;

    // first half of the pipeline: int16 -> float32 (fpga.c:303-385), dc_filter (audio_filters.c:358-373), RF gain
    // (audio_processor.c:305-306), lattice HPF (:445-454)
    auto stage_a = [&](uint64_t word) -> float {
        float x = (float)frame_word(word, widx);
        const float delta_x = x - dc_x;
        const float a1_y_prev = A1 * dc_y;
        const float y = delta_x + a1_y_prev;
        dc_x = x; dc_y = y; x = y;
        x = x * rf_gain;
        if (HPF) { const float h = lattice_step<kHpfStages>(x, hk, hv, HG); x = do_hpf ? h : x; }
        return x;
    };

    constexpr int kPre = UA3_RX_KPRE;                                // frame words in flight ahead of the chain
    const uint32_t total = n_blocks * (uint32_t)kAudioBlock;
    uint64_t cur[kPre], nxt[kPre];
#pragma unroll
    for (int j = 0; j < kPre; ++j) cur[j] = __ldg(fr + ((start + (uint32_t)j) & ring_mask));
    float xa = stage_a(cur[0]);                                      // sample 0 through the first half

    for (uint32_t blk = 0; blk < n_blocks; ++blk) {
        float angle0 = 0.0f, amax = 0.0f;
        for (uint32_t i0 = 0; i0 < (uint32_t)kAudioBlock; i0 += kPre) {
            const uint32_t s0 = blk * (uint32_t)kAudioBlock + i0;    // sample index within the push
#pragma unroll
            for (int j = 0; j < kPre; ++j) {
                const uint32_t sn = s0 + (uint32_t)(kPre + j);
                nxt[j] = __ldg(fr + ((start + (sn < total ? sn : total - 1u)) & ring_mask));
            }
#pragma unroll
            for (int j = 0; j < kPre; ++j) {
                // ---- first half for sample s0 + j + 1 (its result is dropped after the last sample: restore the state) ----
                const float keep_dx = dc_x, keep_dy = dc_y;
                float keep_hg[kHpfStages];
                if (HPF) {
#pragma unroll
                    for (int q = 0; q < kHpfStages; ++q) keep_hg[q] = HG[q];
                }
                const float xa_next = stage_a(j + 1 < kPre ? cur[(j + 1) % kPre] : nxt[0]);
                if (s0 + (uint32_t)j + 1u >= total) {
                    dc_x = keep_dx; dc_y = keep_dy;
                    if (HPF) {
#pragma unroll
                        for (int q = 0; q < kHpfStages; ++q) HG[q] = keep_hg[q];
                    }
                }
                // ---- second half for sample s0 + j: lattice LPF (:456-464), demodulator, notch, S-meter ----
                float x = xa;
                if (LPF) { const float l = lattice_step<kLpfMax>(x, lk, lv, LG); x = do_lpf ? l : x; }
                const float other = __shfl_xor_sync(UA3_FULL_MASK, x, 1);   // the pair's other rail
                const float I = rail == 0 ? x : other, Q = rail == 0 ? other : x;
                float out = x;                                       // what this lane's buffer holds after the demodulator
                {
                    const float dif = I - Q, sum = I + Q;            // :315 / :328
                    const float sq = (I * I) + (Q * Q);              // :338-342
                    const float am = (sq >= 0.0f) ? sqrtf(sq) : 0.0f;
                    float d = is_lsb ? dif : (is_usb ? sum : (is_am ? am : x));
                    if (FM) {                                        // DemodulateFM :517-546
                        const float yy = (Q * fm_ip) - (I * fm_qp);
                        const float xx = (I * fm_ip) + (Q * fm_qp);
                        const float angle = atan2f(yy, xx);
                        const float a = fm_lpf + (0.05f * (angle - fm_lpf));
                        const float b = 0.96f * ((fm_hb + a) - fm_ha);
                        const bool open = !squelched || !sql_on;
                        const float fm_out = open ? (is_wfm ? ((angle / 3.14159265358979f) * 16384.0f) : (b * 30000.0f)) : 0.0f;
                        if (is_fm) {
                            if (i0 == 0 && j == 0) angle0 = angle;
                            fm_lpf = a; fm_qp = Q; fm_ip = I;
                            if (open && !is_wfm) { fm_ha = a; fm_hb = b; }
                            d = fm_out;
                        }
                    }
                    if (rail == 0) out = d;
                    else if (is_am) out = x * x;                     // Q rail holds Q*Q after arm_mult_f32 (:339)
                }
                {   // doRX_NOTCH :466-473 (df2T, 1 stage) on the I buffer
                    const float acc1 = nb0 * out + nd1;
                    float t1 = nb1 * out + nd2;
                    t1 += na1 * acc1;
                    float t2 = nb2 * out;
                    t2 += na2 * acc1;
                    if (do_notch && rail == 0) { nd1 = t1; nd2 = t2; out = acc1; }
                }
                // doRX_SMETER (:491-501): running max / min over both rails
                if (out > sm_max) sm_max = out;
                if (out < sm_min) sm_min = out;
                // arm_max_f32 of the block as DoAGC would take it from this buffer when the DNR is off (agc.c:26)
                amax = (i0 == 0 && j == 0) ? out : ((amax < out) ? out : amax);
                if (live) out_plane[(size_t)(s0 + (uint32_t)j) * out_step] = out;
                xa = xa_next;
            }
#pragma unroll
            for (int j = 0; j < kPre; ++j) cur[j] = nxt[j];
        }
        if (live && rail == 0) amax_out[(size_t)blk * amax_step] = amax;
        sync.publish(blk + 1u, lane);
        if (FM && is_fm && rail == 0) {                              // squelch bookkeeping :548-586
            fm_sql_avg = (0.995f * fm_sql_avg) + (0.005f * sqrtf(fabsf(angle0)));
            if (fm_sql_count == 0) {
                if (fm_sql_avg > 0.7f) fm_sql_avg = 0.7f;
                const float b = fm_sql_avg * 10.0f;
                const int thr = sql_thr;
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
    }
    {   // both lanes of the pair share the S-meter extremes (max / min are order independent: once, at the end)
        const float om = __shfl_xor_sync(UA3_FULL_MASK, sm_max, 1), on = __shfl_xor_sync(UA3_FULL_MASK, sm_min, 1);
        sm_max = fmaxf(sm_max, om); sm_min = fminf(sm_min, on);
    }
    if (live) {
        if (LPF && do_lpf) {
#pragma unroll
            for (int i = 0; i < kLpfMax; ++i) S.lpf_g[rail][i] = LG[i];
        }
        if (HPF && do_hpf) {
#pragma unroll
            for (int i = 0; i < kHpfStages; ++i) S.hpf_g[rail][i] = HG[i];
        }
        S.dc_x[rail] = dc_x; S.dc_y[rail] = dc_y;
        if (rail == 0) {
            S.notch_d[0] = nd1; S.notch_d[1] = nd2;
            S.smeter_max = sm_max; S.smeter_min = sm_min;
            if (FM && is_fm) {
                S.fm_lpf_prev = fm_lpf; S.fm_hpf_prev_a = fm_ha; S.fm_hpf_prev_b = fm_hb; S.fm_i_prev = fm_ip; S.fm_q_prev = fm_qp;
                S.fm_sql_avg = fm_sql_avg; S.fm_sql_count = fm_sql_count; S.squelched = squelched;
            }
        }
    }
}

// one producer warp: 16 channels starting at slot0
UA3_D void rx_filter_warp(const uint64_t* __restrict__ frames, uint32_t ring_mask, uint32_t frame_ch_stride, uint32_t start,
                          uint32_t n_blocks, const RxParams* __restrict__ params, RxState* __restrict__ state, uint32_t n_ch,
                          const uint32_t* __restrict__ order, const RxScratch& scr, uint32_t slot0, int lane, RxBlockSync sync) {
    const int pair = lane >> 1, rail = lane & 1;
    const uint32_t slot = slot0 + (uint32_t)pair;
    const bool live = slot < n_ch;
    const uint32_t ch = order[live ? slot : (n_ch - 1u)];            // dead pairs shadow the last slot's channel and never store
    const RxParams& P = params[ch];
    RxState& S = state[ch];
    const uint8_t mode = P.mode;
    const bool is_ssb = (mode == kModeLSB || mode == kModeCWL || mode == kModeDIGIL || mode == kModeUSB || mode == kModeCWU || mode == kModeDIGIU);
    const bool is_fm = (mode == kModeNFM || mode == kModeWFM);
    const bool do_hpf = (mode == kModeLSB || mode == kModeCWL || mode == kModeUSB || mode == kModeCWU) && P.hpf_set;
    const bool do_lpf = (is_ssb || mode == kModeAM || is_fm) && P.lpf_on;
    const uint32_t any_hpf = __ballot_sync(UA3_FULL_MASK, do_hpf), any_lpf = __ballot_sync(UA3_FULL_MASK, do_lpf);
    const uint32_t any_fm = __ballot_sync(UA3_FULL_MASK, is_fm);
    const uint64_t* fr = frames + (size_t)ch * frame_ch_stride;
    float* out_plane = scr.base + scr.at((uint32_t)rail, 0, slot);
    float* amax_out = scr.amax + slot;
#define UA3_RX_FILTER(H, L, F) rx_filter_body<H, L, F>(fr, ring_mask, start, n_blocks, P, S, rail, lane, live, out_plane, scr.n_slot_pad, amax_out, scr.n_slot_pad, sync)
    if (any_fm) {                                                    // FM never has the HPF, mixed warps take the general body
        if (any_hpf) UA3_RX_FILTER(true, true, true); else UA3_RX_FILTER(false, true, true);
    } else if (any_hpf) {
        UA3_RX_FILTER(true, true, false);
    } else if (any_lpf) {
        UA3_RX_FILTER(false, true, false);
    } else {
        UA3_RX_FILTER(false, false, false);
    }
#undef UA3_RX_FILTER
}

__global__ void __launch_bounds__(32 * kRxFiltWarps, 3)
rx_filter_kernel(const uint64_t* __restrict__ frames, uint32_t ring_mask, uint32_t frame_ch_stride, uint32_t start,
                 uint32_t n_blocks, const RxParams* __restrict__ params, RxState* __restrict__ state, uint32_t n_ch,
                 const uint32_t* __restrict__ order, RxScratch scr) {
    const int warp = threadIdx.x >> 5;
    const uint32_t slot0 = (blockIdx.x * (uint32_t)kRxFiltWarps + (uint32_t)warp) * 16u;
    if (slot0 >= n_ch) return;                                       // whole warp beyond the bank
    rx_filter_warp(frames, ring_mask, frame_ch_stride, start, n_blocks, params, state, n_ch, order, scr, slot0,
                   (int)(threadIdx.x & 31), RxBlockSync{nullptr});
}

// One NLMS step of arm_lms_norm_f32 (numTaps 16, mu 1e-6) on a register window: win[t] = buf[o + t].
template <int O>
UA3_D float nlms_step(const float (&buf)[kLmsTaps - 1 + kLmsTaps], float (&w)[kLmsTaps], float ref, float& energy, float& x0) {
    const float in = buf[O + kLmsTaps - 1];
    energy -= x0 * x0;
    energy += in * in;
    float acc = 0.0f;
#pragma unroll
    for (int t = 0; t < kLmsTaps; ++t) acc += buf[O + t] * w[t];
    const float e = ref - acc;
    const float wg = (e * 0.000001f) / (energy + 0.000000119209289f);
#pragma unroll
    for (int t = 0; t < kLmsTaps; ++t) w[t] += wg * buf[O + t];
    x0 = buf[O];
    return acc;                                                      // output = prediction (noise_reduction.c:32)
}

template <int O>
struct NlmsUnroll {
    UA3_D static void run(const float (&buf)[2 * kLmsTaps - 1], float (&w)[kLmsTaps], const float (&ref)[kLmsTaps],
                          float (&out)[kLmsTaps], float& energy, float& x0) {
        out[O] = nlms_step<O>(buf, w, ref[O], energy, x0);
        NlmsUnroll<O + 1>::run(buf, w, ref, out, energy, x0);
    }
};
template <>
struct NlmsUnroll<kLmsTaps> {
    UA3_D static void run(const float (&)[2 * kLmsTaps - 1], float (&)[kLmsTaps], const float (&)[kLmsTaps], float (&)[kLmsTaps],
                          float&, float&) {}
};

// one consumer warp: 32 channels starting at slot0; s_tile / s_ch are this warp's shared-memory slices
UA3_D void rx_post_warp(uint32_t n_blocks, const RxParams* __restrict__ params, RxState* __restrict__ state, uint32_t n_ch,
                        const uint32_t* __restrict__ order, const RxScratch& scr, int32_t* __restrict__ audio_out, uint32_t out_ch_stride,
                        float* __restrict__ cw_mag, uint32_t cw_ch_stride, uint32_t slot0, int lane, int2* s_tile_w, uint32_t* s_ch_w,
                        RxBlockSync sync_a, RxBlockSync sync_b) {
    const uint32_t slot = slot0 + (uint32_t)lane;
    const bool live = slot < n_ch;
    const uint32_t ch = order[live ? slot : (n_ch - 1u)];
    const uint32_t n_rows = min(32u, n_ch - slot0);                  // live channels of this warp
    s_ch_w[lane] = ch;
    const RxParams& P = params[ch];
    RxState& S = state[ch];
    const uint8_t mode = P.mode;
    const bool is_lsb = (mode == kModeLSB || mode == kModeCWL || mode == kModeDIGIL);
    const bool is_usb = (mode == kModeUSB || mode == kModeCWU || mode == kModeDIGIU);
    const bool chain = is_lsb || is_usb || mode == kModeAM || mode == kModeNFM || mode == kModeWFM;   // NOTCH/DNR/AGC/COPY tail
    const bool dnr = chain && P.dnr_on && live;
    const bool agc_off = !P.agc_on || mode == kModeDIGIL || mode == kModeDIGIU;
    const bool cw_on = P.cw_on != 0;
    const float cw_coeff = P.cw_coeff;
    const float step_up = P.agc_step_up, step_down = P.agc_step_down;
    const bool mute = P.mute != 0;
    const float volume = P.volume;
    float agc_gain = S.agc_gain, agc_old = S.agc_gain_old;
    float w[kLmsTaps];
    float buf[2 * kLmsTaps - 1];                                     // NLMS window: 15 past inputs + 16 new ones
    float lms_energy = 0.0f, lms_x0 = 0.0f;
    uint32_t idx_old = 0, idx_new = 0;
    int half_src[2] = {-1, -1};                                      // which sub-block of this push each half of lms2_reference holds (-1: state)
    if (dnr) {
#pragma unroll
        for (int i = 0; i < kLmsTaps; ++i) w[i] = S.lms_w[i];
#pragma unroll
        for (int i = 0; i < kLmsTaps - 1; ++i) buf[i] = S.lms_hist[i];
        lms_energy = S.lms_energy; lms_x0 = S.lms_x0;
        idx_old = S.lms_idx_old; idx_new = S.lms_idx_new;
    }
    const float* pI = scr.base + scr.at(0, 0, slot);
    const float* pQ = scr.base + scr.at(1, 0, slot);
    float* pD = scr.base + scr.at(2, 0, slot);
    const size_t step = scr.n_slot_pad;
    __syncwarp();

    for (uint32_t blk = 0; blk < n_blocks; ++blk) {
        sync_a.wait(blk + 1u);                                       // both producer warps of these 32 channels have written block blk
        sync_b.wait(blk + 1u);
        const uint32_t b0 = blk * (uint32_t)kAudioBlock;
        const float* src = pI;                                       // what DoAGC sees
        float amax = 0.0f;
        if (chain && live) {
            // ---------------- doRX_DNR (:475-483): three 64-sample sub-blocks ---------------------------------
            if (dnr) {
                src = pD;
                for (int sb = 0; sb < kAudioBlock / kSubBlock; ++sb) {
                    const uint32_t sub = blk * 3u + (uint32_t)sb, n0 = b0 + (uint32_t)sb * kSubBlock;
                    // arm_copy_f32(bufferIn, &lms2_reference[reference_index_new], 64): that half now holds this sub-block
                    half_src[idx_new >> 6] = (int)sub;
                    const int rs = half_src[idx_old >> 6];           // the half the error term reads
                    for (int c = 0; c < kSubBlock / kLmsTaps; ++c) {
                        float ref[kLmsTaps], out[kLmsTaps];
#pragma unroll
                        for (int j = 0; j < kLmsTaps; ++j) buf[kLmsTaps - 1 + j] = pI[(size_t)(n0 + c * kLmsTaps + j) * step];
                        if (rs < 0) {
#pragma unroll
                            for (int j = 0; j < kLmsTaps; ++j) ref[j] = S.lms_ref[idx_old + c * kLmsTaps + j];
                        } else {
#pragma unroll
                            for (int j = 0; j < kLmsTaps; ++j) ref[j] = pI[(size_t)((uint32_t)rs * kSubBlock + c * kLmsTaps + j) * step];
                        }
                        NlmsUnroll<0>::run(buf, w, ref, out, lms_energy, lms_x0);
#pragma unroll
                        for (int j = 0; j < kLmsTaps; ++j) pD[(size_t)(n0 + c * kLmsTaps + j) * step] = out[j];
#pragma unroll
                        for (int j = 0; j < kLmsTaps - 1; ++j) buf[j] = buf[kLmsTaps + j];
                    }
                    idx_old += kSubBlock;                            // noise_reduction.c:33-36
                    if (idx_old >= 2 * kSubBlock) idx_old = 0;
                    idx_new = idx_old + kSubBlock;
                    if (idx_new >= 2 * kSubBlock) idx_new = 0;
                }
            }
            // ---------------- DoAGC (agc.c:21-67) ------------------------------------------------------------
            // arm_max_f32 (signed maximum): of the NLMS output when the DNR ran, else rx_filter_kernel already took it
            if (dnr) {
                amax = src[(size_t)b0 * step];
                for (int i0 = 0; i0 < kAudioBlock; i0 += 16) {       // sixteen loads in flight: the scratch lives in L2
                    float v[16];
#pragma unroll
                    for (int j = 0; j < 16; ++j) v[j] = src[(size_t)(b0 + i0 + j) * step];
#pragma unroll
                    for (int j = 0; j < 16; ++j) if (amax < v[j]) amax = v[j];
                }
            } else {
                amax = scr.amax[(size_t)blk * scr.n_slot_pad + slot];
            }
            if (amax == 0.0f) amax = 0.001f;
            const float target = 7000.0f / amax;
            if (target > agc_gain) {
                float st = (target - agc_gain) / step_up;
                if (st > 1.0f) st = 1.0f;
                agc_gain += st;
            } else {
                agc_gain -= (agc_gain - target) / step_down;
            }
            if (agc_gain < 0.0f) agc_gain = 0.0f;
            if ((agc_gain * amax) > 10000.0f) agc_gain = target;
            if (agc_off) agc_gain = 1.0f;
        }
        const bool ramp = chain && (agc_old != agc_gain);
        float gstep = 0.0f;
        if (ramp) {
            if (agc_old > agc_gain) gstep = -(agc_old - agc_gain) / 192.0f;
            if (agc_old < agc_gain) gstep = (agc_gain - agc_old) / 192.0f;
        }
        float Q1 = 0.0f, Q2 = 0.0f;                                  // Goertzel of CWDecoder_Process (cw_decoder.c:56-66)
        int32_t* dst_blk = audio_out + (size_t)blk * (2 * kAudioBlock);
        for (int t0 = 0; t0 < kAudioBlock; t0 += kRxTile) {
#pragma unroll
            for (int h = 0; h < kRxTile; h += 16) {
                float lv[16], rv[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) {                       // sixteen loads in flight, then the sequential part
                    const size_t o = (size_t)(b0 + t0 + h + j) * step;
                    lv[j] = live ? src[o] : 0.0f;
                    rv[j] = (live && !chain) ? pQ[o] : 0.0f;
                }
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    float l = lv[j], r = rv[j];
                    if (live) {
                        if (chain) {
                            if (ramp) { agc_old += gstep; l = l * agc_old; }
                            else l = l * agc_gain;
                            r = l;                                   // doRX_COPYCHANNEL: Q <- I
                        }
                        if (cw_on) { const float Q0 = ((cw_coeff * Q1) - Q2) + l; Q2 = Q1; Q1 = Q0; }
                        l = mute ? (l * 0.0f) : (l * volume);        // :365-374
                        r = mute ? (r * 0.0f) : (r * volume);
                    }
                    s_tile_w[lane * kRxTilePitch + h + j] = make_int2((int32_t)l, (int32_t)r);   // C cast: truncation toward zero
                }
            }
            __syncwarp();
            for (uint32_t row = 0; row < n_rows; ++row) {            // one channel row per step: 32 samples x (L, R) = 256 B
                int2* d = reinterpret_cast<int2*>(dst_blk + (size_t)s_ch_w[row] * out_ch_stride + 2 * t0);
                d[lane] = s_tile_w[row * kRxTilePitch + lane];
            }
            __syncwarp();
        }
        if (live) {
            float mag = 0.0f;
            if (cw_on) {
                const float msq = ((Q1 * Q1) + (Q2 * Q2)) - ((Q1 * Q2) * cw_coeff);
                mag = sqrtf(msq);
            }
            cw_mag[(size_t)ch * cw_ch_stride + blk] = mag;
        }
    }
    if (live && chain) { S.agc_gain = agc_gain; S.agc_gain_old = agc_old; }
    if (dnr) {
#pragma unroll
        for (int i = 0; i < kLmsTaps; ++i) S.lms_w[i] = w[i];
#pragma unroll
        for (int i = 0; i < kLmsTaps - 1; ++i) S.lms_hist[i] = buf[i];
        S.lms_energy = lms_energy; S.lms_x0 = lms_x0;
        S.lms_idx_old = idx_old; S.lms_idx_new = idx_new;
        for (int h = 0; h < 2; ++h)                                  // the halves of lms2_reference this push rewrote
            if (half_src[h] >= 0)
                for (int j = 0; j < kSubBlock; ++j) S.lms_ref[h * kSubBlock + j] = pI[(size_t)((uint32_t)half_src[h] * kSubBlock + j) * step];
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

__global__ void __launch_bounds__(32 * kRxPostWarps, 3)
rx_post_kernel(uint32_t n_blocks, const RxParams* __restrict__ params, RxState* __restrict__ state, uint32_t n_ch,
               const uint32_t* __restrict__ order, RxScratch scr, int32_t* __restrict__ audio_out, uint32_t out_ch_stride,
               float* __restrict__ cw_mag, uint32_t cw_ch_stride) {
    __shared__ __align__(8) int2 s_tile[kRxPostWarps][32 * kRxTilePitch];
    __shared__ uint32_t s_ch[kRxPostWarps][32];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t slot0 = (blockIdx.x * (uint32_t)kRxPostWarps + (uint32_t)warp) * 32u;
    if (slot0 >= n_ch) return;                                       // whole warp beyond the bank
    rx_post_warp(n_blocks, params, state, n_ch, order, scr, audio_out, out_ch_stride, cw_mag, cw_ch_stride, slot0, lane,
                 s_tile[warp], s_ch[warp], RxBlockSync{nullptr}, RxBlockSync{nullptr});
}

// processRxAudio as ONE warp-specialised kernel: a CTA owns 64 channels - four producer warps (rx_filter_warp, 16 channels
// each) and two consumer warps (rx_post_warp, 32 channels each).  A consumer starts on block b as soon as its two producers
// have published it, while they are already filtering block b+1: the post stage costs one block of latency instead of a
// whole push, which is what decides the step time at small channel counts (the chain is per-channel sequential).
// G such 64-channel groups share a CTA.  Two of them (384 threads x 168 registers) fill an SM's register file with ONE CTA:
// the block scheduler places CTAs breadth first over the idle SMs, so 192-thread CTAs end up one per SM - twice the SMs for
// the same warps - and every SM the stage holds is an SM the next push's front kernel has to wait for.  Twelve warps on
// an SM run each about a quarter slower than six, though, so the packed form is for banks whose step leaves the stage time
// (rx_launch_audio: from 2048 channels, where a block takes about 1.3 ms); measured at 4096 channels 2.865 -> 2.799 ms per step, at 1024 0.774 -> 0.812.
template <int G>
__global__ void __launch_bounds__(32 * 6 * G, G == 1 ? UA3_RX_FUSED_MINB : 1)
rx_audio_kernel(const uint64_t* __restrict__ frames, uint32_t ring_mask, uint32_t frame_ch_stride, uint32_t start,
                uint32_t n_blocks, const RxParams* __restrict__ params, RxState* __restrict__ state, uint32_t n_ch,
                const uint32_t* __restrict__ order, RxScratch scr, int32_t* __restrict__ audio_out, uint32_t out_ch_stride,
                float* __restrict__ cw_mag, uint32_t cw_ch_stride) {
    constexpr int kRxFusedProd = 4 * G, kRxFusedCons = 2 * G;
    __shared__ __align__(8) int2 s_tile[kRxFusedCons][32 * kRxTilePitch];
    __shared__ uint32_t s_ch[kRxFusedCons][32];
    __shared__ uint32_t s_done[kRxFusedProd];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x < kRxFusedProd) s_done[threadIdx.x] = 0;
    __syncthreads();
    const uint32_t cta_slot0 = blockIdx.x * (64u * G);
    if (warp < kRxFusedProd) {
        const uint32_t slot0 = cta_slot0 + (uint32_t)warp * 16u;
        if (slot0 >= n_ch) return;
        rx_filter_warp(frames, ring_mask, frame_ch_stride, start, n_blocks, params, state, n_ch, order, scr, slot0, lane,
                       RxBlockSync{&s_done[warp]});
    } else {
        const int cw = warp - kRxFusedProd;
        const uint32_t slot0 = cta_slot0 + (uint32_t)cw * 32u;
        if (slot0 >= n_ch) return;
        // the second producer of this half may be beyond the bank: then only the first one is waited for
        volatile uint32_t* a = &s_done[2 * cw];
        volatile uint32_t* b = (slot0 + 16u < n_ch) ? &s_done[2 * cw + 1] : &s_done[2 * cw];
        rx_post_warp(n_blocks, params, state, n_ch, order, scr, audio_out, out_ch_stride, cw_mag, cw_ch_stride, slot0, lane,
                     s_tile[cw], s_ch[cw], RxBlockSync{a}, RxBlockSync{b});
    }
}

// ------------------------------------------------------------------------------------------------
// FFT_doFFT, sequential half (fft.c:225-243): dc_filter of FFTInput_I/Q, the optional notch, the ZoomFFT biquad cascade
// and FIR decimator - recurrences along time, so they run with a lane PAIR per channel (even lane I, odd lane Q; 16
// channels per warp) instead of on two lanes of a per-channel warp.  Output: fin[frame][rail][sample][slot] floats
// (512 / zoom samples per frame when zoomed), read back by rx_fft_kernel.
// ------------------------------------------------------------------------------------------------
constexpr int kRxFftPreWarps = 16;      // 512 threads x 117 registers: one CTA fills an SM (128-thread CTAs were placed one per SM, see rx_audio_kernel)

__global__ void __launch_bounds__(32 * kRxFftPreWarps, 1)
rx_fft_pre_kernel(const uint64_t* __restrict__ frames, uint32_t ring_mask, uint32_t frame_ch_stride, uint32_t start,
                  uint32_t n_frames, const RxParams* __restrict__ params, RxState* __restrict__ state, uint32_t n_ch,
                  const uint32_t* __restrict__ order, float* __restrict__ fin, uint32_t n_slot_pad) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, pair = lane >> 1, rail = lane & 1;
    const uint32_t slot = (blockIdx.x * (uint32_t)kRxFftPreWarps + (uint32_t)warp) * 16u + (uint32_t)pair;
    if (slot >= n_ch) return;                                         // no warp-level operation below
    const uint32_t ch = order[slot];
    const RxParams& P = params[ch];
    RxState& S = state[ch];
    if (!P.fft_enabled) return;                                       // fft.c:214
    const uint64_t* fr = frames + (size_t)ch * frame_ch_stride;
    const float A1 = (float)(1.0 - 0.00048828125);
    const int widx = (rail == 0) ? (P.iq_swap ? 0 : 1) : (P.iq_swap ? 1 : 0);   // FFT_buff filling of fpga.c:305-341, with the I/Q swap
    const int zoom = P.fft_zoom > 1 ? P.fft_zoom : 1;
    const int zsel = zoom == 2 ? 0 : (zoom == 4 ? 1 : (zoom == 8 ? 2 : 3));
    const bool notch_on = P.notch_on != 0;
    float dx = S.dc_x[4 + rail], dy = S.dc_y[4 + rail];
    float d1 = S.notch_fft_d[rail][0], d2 = S.notch_fft_d[rail][1];
    const float b0 = P.notch[0], b1 = P.notch[1], b2 = P.notch[2], a1 = P.notch[3], a2 = P.notch[4];
    float st[4][4];
#pragma unroll
    for (int sg = 0; sg < 4; ++sg)
#pragma unroll
        for (int q = 0; q < 4; ++q) st[sg][q] = S.zoom_biquad[rail][sg][q];
    float h0 = S.zoom_fir[rail][0], h1 = S.zoom_fir[rail][1], h2 = S.zoom_fir[rail][2];
    float bq[20];
#pragma unroll
    for (int q = 0; q < 20; ++q) bq[q] = c_zoom_biquad[zsel][q];
    const float f0 = c_zoom_fir[zsel][0], f1 = c_zoom_fir[zsel][1], f2 = c_zoom_fir[zsel][2], f3 = c_zoom_fir[zsel][3];
    constexpr int kPre = 8;
    const uint32_t total = n_frames * (uint32_t)kFftSize;
    uint64_t cur[kPre], nxt[kPre];
#pragma unroll
    for (int j = 0; j < kPre; ++j) cur[j] = __ldg(fr + ((start + (uint32_t)j) & ring_mask));
    for (uint32_t fi = 0; fi < n_frames; ++fi) {
        float* dst = fin + ((size_t)(fi * 2u + (uint32_t)rail) * kFftSize) * n_slot_pad + slot;
        for (uint32_t i0 = 0; i0 < (uint32_t)kFftSize; i0 += kPre) {
            const uint32_t s0 = fi * (uint32_t)kFftSize + i0;
            if (s0 + kPre < total) {
#pragma unroll
                for (int j = 0; j < kPre; ++j) nxt[j] = __ldg(fr + ((start + s0 + (uint32_t)(kPre + j)) & ring_mask));
            }
#pragma unroll
            for (int j = 0; j < kPre; ++j) {
                const uint32_t i = i0 + (uint32_t)j;
                float x = (float)frame_word(cur[j], widx);
                const float delta_x = x - dx;                         // dc_filter(FFTInput_I/Q, 512, 4/5) (:225-226)
                const float a1y = A1 * dy;
                const float y = delta_x + a1y;
                dx = x; dy = y; x = y;
                if (notch_on) {                                       // :228-233
                    const float acc1 = b0 * x + d1;
                    d1 = b1 * x + d2;
                    d1 += a1 * acc1;
                    d2 = b2 * x;
                    d2 += a2 * acc1;
                    x = acc1;
                }
                if (zoom > 1) {
                    // arm_biquad_cascade_df1_f32, 4 stages (fft.c:239-240), then arm_fir_decimate_f32, 4 taps, M = zoom (:242-243)
#pragma unroll
                    for (int sg = 0; sg < 4; ++sg) {
                        const float acc = ((((bq[sg * 5] * x) + (bq[sg * 5 + 1] * st[sg][0])) + (bq[sg * 5 + 2] * st[sg][1])) +
                                           (bq[sg * 5 + 3] * st[sg][2])) + (bq[sg * 5 + 4] * st[sg][3]);
                        st[sg][1] = st[sg][0]; st[sg][0] = x; st[sg][3] = st[sg][2]; st[sg][2] = acc;
                        x = acc;
                    }
                    // decimator: output o is computed right after the FIRST sample of its group of `zoom` inputs has been
                    // shifted in (arm_fir_decimate_f32 copies M samples, then reads the window that starts M back)
                    if (i % (uint32_t)zoom == 0) {
                        float sum0 = 0.0f;
                        sum0 += h0 * f0; sum0 += h1 * f1; sum0 += h2 * f2; sum0 += x * f3;
                        dst[(size_t)(i / (uint32_t)zoom) * n_slot_pad] = sum0;
                    }
                    h0 = h1; h1 = h2; h2 = x;
                } else {
                    dst[(size_t)i * n_slot_pad] = x;
                }
            }
#pragma unroll
            for (int j = 0; j < kPre; ++j) cur[j] = nxt[j];
        }
    }
    S.dc_x[4 + rail] = dx; S.dc_y[4 + rail] = dy;
    S.notch_fft_d[rail][0] = d1; S.notch_fft_d[rail][1] = d2;
    if (zoom > 1) {
#pragma unroll
        for (int sg = 0; sg < 4; ++sg)
#pragma unroll
            for (int q = 0; q < 4; ++q) S.zoom_biquad[rail][sg][q] = st[sg][q];
        S.zoom_fir[rail][0] = h0; S.zoom_fir[rail][1] = h1; S.zoom_fir[rail][2] = h2;
    }
}

// FFT_doFFT, parallel half: one warp per channel, kRxFftWarps channels (consecutive slots) per CTA sharing the window and
// twiddle tables in shared memory (in constant memory their lane-varying indices would be serialised).
#if defined(UA3_HOST_EMU)
constexpr int kRxFftWarps = 1;          // the emulation's __syncwarp is a CTA barrier: one warp per CTA there
#else
constexpr int kRxFftWarps = 8;
#endif

__global__ void __launch_bounds__(32 * kRxFftWarps)
rx_fft_kernel(uint32_t n_frames, const RxParams* __restrict__ params, RxState* __restrict__ state, uint32_t n_ch,
              const uint32_t* __restrict__ order, const float* __restrict__ fin, uint32_t n_slot_pad,
              float* __restrict__ spectra, uint32_t spec_ch_stride, uint16_t* __restrict__ waterfall,
              uint16_t* __restrict__ wtf_hist, uint32_t* __restrict__ wtf_head, int32_t* __restrict__ wtf_pending_hz) {
    __shared__ float s_re_all[kRxFftWarps][kFftSize], s_im_all[kRxFftWarps][kFftSize];
    __shared__ float s_window[kFftSize], s_twiddle[2 * kFftSize];
    for (int i = threadIdx.x; i < kFftSize; i += 32 * kRxFftWarps) s_window[i] = c_fft_window[i];
    for (int i = threadIdx.x; i < 2 * kFftSize; i += 32 * kRxFftWarps) s_twiddle[i] = c_fft_twiddle[i];
    __syncthreads();
    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    float* s_re = s_re_all[warp];
    float* s_im = s_im_all[warp];
    const uint32_t slot = blockIdx.x * (uint32_t)kRxFftWarps + (uint32_t)warp;
    if (slot >= n_ch) return;
    const uint32_t ch = order[slot];
    const RxParams& P = params[ch];
    RxState& S = state[ch];
    if (!P.fft_enabled) return;                                       // fft.c:214
    const int zoom = P.fft_zoom > 1 ? P.fft_zoom : 1;

    for (uint32_t fi = 0; fi < n_frames; ++fi) {
        const float* src_re = fin + ((size_t)(fi * 2u) * kFftSize) * n_slot_pad + slot;
        const float* src_im = fin + ((size_t)(fi * 2u + 1u) * kFftSize) * n_slot_pad + slot;
        if (zoom > 1) {
            // slide FFTInput_ZOOMFFT left by zoomed_width and append the new decimated samples (fft.c:245-260)
            const int zw = kFftSize / zoom;
            for (int i = lane; i < kFftSize; i += 32) {
                if (i < kFftSize - zw) { s_re[i] = S.zoom_buf[2 * (i + zw)]; s_im[i] = S.zoom_buf[2 * (i + zw) + 1]; }
                else {
                    s_re[i] = src_re[(size_t)(i - (kFftSize - zw)) * n_slot_pad];
                    s_im[i] = src_im[(size_t)(i - (kFftSize - zw)) * n_slot_pad];
                }
            }
        } else {
            for (int i = lane; i < kFftSize; i += 32) {
                s_re[i] = src_re[(size_t)i * n_slot_pad];
                s_im[i] = src_im[(size_t)i * n_slot_pad];
            }
        }
        __syncwarp();
        if (zoom > 1)
            for (int i = lane; i < kFftSize; i += 32) { S.zoom_buf[2 * i] = s_re[i]; S.zoom_buf[2 * i + 1] = s_im[i]; }
        __syncwarp();
        // Hamming window (:275-286)
        for (int i = lane; i < kFftSize; i += 32) {
            const float wm = s_window[i];
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
                        const float co = s_twiddle[2 * (id * m)], si = s_twiddle[2 * (id * m) + 1];
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
// Stand-alone stages on a caller buffer, with the channel's state: the firmware's dc_filter(), DoAGC() and
// processNoiseReduction() as entry points of their own (audio_filters.h:51, agc.h:9, noise_reduction.h:16).  One thread:
// these are single-channel calls on 64..512 samples, kept for callers that use the functions individually.
// ------------------------------------------------------------------------------------------------
__global__ void rx_stage_kernel(int stage, float* __restrict__ buf, float* __restrict__ out, uint32_t n, const RxParams* __restrict__ params,
                                RxState* __restrict__ state, uint32_t ch, int arg) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const RxParams& P = params[ch];
    RxState& S = state[ch];
    if (stage == 0) {                                                // dc_filter(buf, n, stateNum = arg) (audio_filters.c:358-373)
        const float A1 = (float)(1.0 - 0.00048828125);
        float dx = S.dc_x[arg], dy = S.dc_y[arg];
        for (uint32_t i = 0; i < n; ++i) {
            const float x = buf[i];
            const float delta_x = x - dx;
            const float a1y = A1 * dy;
            const float y = delta_x + a1y;
            dx = x; dy = y; buf[i] = y;
        }
        S.dc_x[arg] = dx; S.dc_y[arg] = dy;
    } else if (stage == 1) {                                         // DoAGC(buf, n) (agc.c:21-67)
        float agc_gain = S.agc_gain, agc_old = S.agc_gain_old;
        float amax = buf[0];
        for (uint32_t i = 1; i < n; ++i) if (amax < buf[i]) amax = buf[i];
        if (amax == 0.0f) amax = 0.001f;
        const float target = 7000.0f / amax;
        if (target > agc_gain) {
            float st = (target - agc_gain) / P.agc_step_up;
            if (st > 1.0f) st = 1.0f;
            agc_gain += st;
        } else {
            agc_gain -= (agc_gain - target) / P.agc_step_down;
        }
        if (agc_gain < 0.0f) agc_gain = 0.0f;
        if ((agc_gain * amax) > 10000.0f) agc_gain = target;
        if (!P.agc_on || P.mode == kModeDIGIL || P.mode == kModeDIGIU) agc_gain = 1.0f;
        if (agc_old != agc_gain) {
            float gstep = 0.0f;
            if (agc_old > agc_gain) gstep = -(agc_old - agc_gain) / (float)(int)n;
            if (agc_old < agc_gain) gstep = (agc_gain - agc_old) / (float)(int)n;
            for (uint32_t i = 0; i < n; ++i) { agc_old += gstep; buf[i] = buf[i] * agc_old; }
        } else {
            for (uint32_t i = 0; i < n; ++i) buf[i] = buf[i] * agc_gain;
        }
        S.agc_gain = agc_gain; S.agc_gain_old = agc_old;
    } else if (stage == 2) {                                         // processNoiseReduction(buf, out): 64 samples (noise_reduction.c:25-37)
        if (!P.dnr_on) return;
        uint32_t idx_old = S.lms_idx_old, idx_new = S.lms_idx_new;
        float energy = S.lms_energy, x0 = S.lms_x0;
        float win[kLmsTaps - 1 + kSubBlock];
        for (int i = 0; i < kLmsTaps - 1; ++i) win[i] = S.lms_hist[i];
        for (int i = 0; i < kSubBlock; ++i) { win[kLmsTaps - 1 + i] = buf[i]; S.lms_ref[idx_new + i] = buf[i]; }
        for (int i = 0; i < kSubBlock; ++i) {
            const float in = win[i + kLmsTaps - 1];
            energy -= x0 * x0;
            energy += in * in;
            float acc = 0.0f;
            for (int t = 0; t < kLmsTaps; ++t) acc += win[i + t] * S.lms_w[t];
            out[i] = acc;
            const float e = S.lms_ref[idx_old + i] - acc;
            const float wg = (e * 0.000001f) / (energy + 0.000000119209289f);
            for (int t = 0; t < kLmsTaps; ++t) S.lms_w[t] += wg * win[i + t];
            x0 = win[i];
        }
        for (int i = 0; i < kLmsTaps - 1; ++i) S.lms_hist[i] = win[kSubBlock + i];
        idx_old += kSubBlock;
        if (idx_old >= 2 * kSubBlock) idx_old = 0;
        idx_new = idx_old + kSubBlock;
        if (idx_new >= 2 * kSubBlock) idx_new = 0;
        S.lms_energy = energy; S.lms_x0 = x0; S.lms_idx_old = idx_old; S.lms_idx_new = idx_new;
    }
}

cudaError_t rx_launch_stage(const RxBuffers& b, int stage, float* buf_dev, float* out_dev, uint32_t n, uint32_t ch, int arg, cudaStream_t st,
                            int* launches) {
    UA3_LAUNCH(rx_stage_kernel, 1, 32, 0, st, stage, buf_dev, out_dev, n, b.params, b.state, ch, arg);
    if (launches) *launches += 1;
    return cudaGetLastError();
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
    RxScratch scr;
    scr.amax = b.scratch + rx_scratch_floats(b.n_ch, b.max_audio_blocks) - (size_t)b.max_audio_blocks * ((b.n_ch + 31u) & ~31u);
    scr.base = b.scratch; scr.max_samples = b.max_audio_blocks * (uint32_t)kAudioBlock; scr.n_slot_pad = (b.n_ch + 31u) & ~31u;
#if defined(UA3_HOST_EMU)
    UA3_LAUNCH(rx_filter_kernel, (b.n_ch + 16u * kRxFiltWarps - 1u) / (16u * kRxFiltWarps), 32 * kRxFiltWarps, 0, st, b.frames, b.ring_mask,
               b.frame_ch_stride, start, n_blocks, b.params, b.state, b.n_ch, b.order, scr);
    UA3_LAUNCH(rx_post_kernel, (b.n_ch + 32u * kRxPostWarps - 1u) / (32u * kRxPostWarps), 32 * kRxPostWarps, 0, st, n_blocks, b.params, b.state,
               b.n_ch, b.order, scr, b.audio_out, b.audio_ch_stride, b.cw_mag, b.max_audio_blocks);
    if (launches) *launches += 1;
#else
    if (b.split_audio) {       // the two halves as separate kernels (development / profiling: UA3REO_RX_SPLIT=1)
        UA3_LAUNCH(rx_filter_kernel, (b.n_ch + 16u * kRxFiltWarps - 1u) / (16u * kRxFiltWarps), 32 * kRxFiltWarps, 0, st, b.frames, b.ring_mask,
                   b.frame_ch_stride, start, n_blocks, b.params, b.state, b.n_ch, b.order, scr);
        UA3_LAUNCH(rx_post_kernel, (b.n_ch + 32u * kRxPostWarps - 1u) / (32u * kRxPostWarps), 32 * kRxPostWarps, 0, st, n_blocks, b.params, b.state,
                   b.n_ch, b.order, scr, b.audio_out, b.audio_ch_stride, b.cw_mag, b.max_audio_blocks);
        if (launches) *launches += 1;
    } else {
        if (b.n_ch >= 2048u)
            UA3_LAUNCH(rx_audio_kernel<2>, (b.n_ch + 127u) / 128u, 384, 0, st, b.frames, b.ring_mask, b.frame_ch_stride, start, n_blocks,
                       b.params, b.state, b.n_ch, b.order, scr, b.audio_out, b.audio_ch_stride, b.cw_mag, b.max_audio_blocks);
        else
            UA3_LAUNCH(rx_audio_kernel<1>, (b.n_ch + 63u) / 64u, 192, 0, st, b.frames, b.ring_mask, b.frame_ch_stride, start, n_blocks,
                       b.params, b.state, b.n_ch, b.order, scr, b.audio_out, b.audio_ch_stride, b.cw_mag, b.max_audio_blocks);
    }
#endif
    if (launches) *launches += 1;
    return cudaGetLastError();
}

size_t rx_scratch_floats(uint32_t n_ch, uint32_t max_audio_blocks) {
    const size_t pad = (n_ch + 31u) & ~31u;
    return (size_t)3 * max_audio_blocks * kAudioBlock * pad + (size_t)max_audio_blocks * pad;      // three planes + the block maxima
}

int rx_audio_sms(uint32_t n_ch) {
    // whole SMs rx_audio_kernel can fill: 128 channels per SM (register bound; two 64-channel CTAs or one of 128)
    return (int)((n_ch + 127u) / 128u);
}

cudaError_t rx_launch_fft(const RxBuffers& b, uint32_t start, uint32_t n_frames, cudaStream_t st, int* launches) {
    if (!n_frames) return cudaSuccess;
    const uint32_t pad = (b.n_ch + 31u) & ~31u;
    UA3_LAUNCH(rx_fft_pre_kernel, (b.n_ch + 16u * kRxFftPreWarps - 1u) / (16u * kRxFftPreWarps), 32 * kRxFftPreWarps, 0, st, b.frames,
               b.ring_mask, b.frame_ch_stride, start, n_frames, b.params, b.state, b.n_ch, b.order, b.fft_in, pad);
    UA3_LAUNCH(rx_fft_kernel, (b.n_ch + kRxFftWarps - 1u) / kRxFftWarps, 32 * kRxFftWarps, 0, st, n_frames, b.params,
               b.state, b.n_ch, b.order, b.fft_in, pad, b.spectra, b.spec_ch_stride, b.waterfall, b.wtf_hist, b.wtf_head, b.wtf_pending_hz);
    if (launches) *launches += 2;
    return cudaGetLastError();
}

size_t rx_fft_in_floats(uint32_t n_ch, uint32_t max_fft_frames) {
    return (size_t)max_fft_frames * 2 * kFftSize * ((n_ch + 31u) & ~31u);
}

cudaError_t rx_launch_clear(const RxBuffers& b, const uint8_t* flags_dev, uint32_t first, uint32_t n, cudaStream_t st,
                            int* launches) {
    if (!n) return cudaSuccess;
    UA3_LAUNCH(rx_clear_filters_kernel, (n + 127u) / 128u, 128, 0, st, b.state, flags_dev, first, n);
    if (launches) *launches += 1;
    return cudaGetLastError();
}

}  // namespace ua3
