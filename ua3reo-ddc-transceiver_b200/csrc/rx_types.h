// rx_types.h - per-channel parameter and state blocks of the STM32 receive-audio / panorama-FFT stage.
//
// The firmware keeps one copy of this state in file-scope statics (audio_processor.c:29-42,
// audio_filters.c:13-38,348-356, agc.c:9-12, noise_reduction.c:12-16, fft.c:19-31); here every channel
// owns one RxParams (derived on the host from the TRX settings the path reads, settings.h:63-123)
// and one RxState (device resident, carried from block to block).
#pragma once
#include <stdint.h>

namespace ua3 {

// trx_manager.h:11-24
enum : uint8_t {
    kModeLSB = 0, kModeUSB = 1, kModeIQ = 2, kModeCWL = 3, kModeCWU = 4, kModeDIGIL = 5, kModeDIGIU = 6,
    kModeNoTX = 7, kModeNFM = 8, kModeWFM = 9, kModeAM = 10, kModeLoopback = 11
};

constexpr int kAudioBlock = 192;     // FPGA_AUDIO_BUFFER_HALF_SIZE (audio_processor.h:10-11)
constexpr int kSubBlock = 64;        // APROCESSOR_BLOCK_SIZE / NOISE_REDUCTION_BLOCK_SIZE
constexpr int kLpfMax = 11;          // IIR_LPF_STAGES (audio_filters.h:13)
constexpr int kHpfStages = 6;        // IIR_HPF_STAGES
constexpr int kLmsTaps = 16;         // NOISE_REDUCTION_TAPS (noise_reduction.h:11)
constexpr int kFftSize = 512;        // FFT_SIZE (fft.h:10)
constexpr int kFftBins = 256;        // FFT_PRINT_SIZE

struct RxParams {
    // flags
    uint8_t mode, agc_on, dnr_on, notch_on, mute, iq_swap, fm_sql_threshold, fft_enabled;
    uint8_t lpf_on;          // Filter_Width > 0 (audio_processor.c:448)
    uint8_t hpf_set;         // HPF coefficients have been initialised at least once
    uint8_t cw_on;           // TRX.CWDecoder && mode is CW_L / CW_U (audio_processor.c:437-443)
    uint8_t fft_zoom;        // TRX.FFT_Zoom: 1, 2, 4, 8 or 16 (fft.c:185-210)
    float rf_gain;           // (float)TRX.RF_Gain
    float volume;            // (float)TRX.Volume / 100.0f
    float agc_step_up;       // 500.0f / Agc_speed (agc.c:17)
    float agc_step_down;     // step_up / 10.0f   (agc.c:18)
    float fft_averaging;     // (float)TRX.FFT_Averaging
    float cw_coeff;          // Goertzel coefficient 2*cos(omega) of CWDecoder_Init (cw_decoder.c:43-54)
    // lattice coefficients, TOP-PADDED with zero stages to the maximum stage count: a stage with
    // k = v = 0 leaves f and the accumulator bit-identical, so 7-stage filters run in the 11-stage loop.
    float lpf_k[kLpfMax];
    float lpf_v[kLpfMax + 1];
    float hpf_k[kHpfStages];
    float hpf_v[kHpfStages + 1];
    float notch[5];          // b0, b1, b2, a1, a2 as stored by calcBiquad (audio_filters.c:490-494)
};

struct RxState {
    float dc_x[6], dc_y[6];              // dc_filter_state[6] (audio_filters.c:348-356): 0 RX I, 1 RX Q, 4 FFT I, 5 FFT Q
    float lpf_g[2][kLpfMax];             // lattice g states per rail (top-padded like the coefficients)
    float hpf_g[2][kHpfStages];
    float notch_d[2];                    // NOTCH_State
    float notch_fft_d[2][2];             // NOTCH_State_FFT_I/Q
    float smeter_max, smeter_min;        // Processor_RX_Audio_Samples_MAX/MIN_value
    // noise_reduction.c
    float lms_w[kLmsTaps];
    float lms_hist[kLmsTaps - 1];        // previous numTaps-1 inputs (oldest first)
    float lms_energy, lms_x0;
    float lms_ref[2 * kSubBlock];        // lms2_reference
    uint32_t lms_idx_old, lms_idx_new;   // reference_index_old/new (noise_reduction.c:28-36)
    // agc.c
    float agc_gain, agc_gain_old;
    // DemodulateFM statics (audio_processor.c:509-515)
    float fm_lpf_prev, fm_hpf_prev_a, fm_hpf_prev_b, fm_i_prev, fm_q_prev, fm_sql_avg;
    uint32_t fm_sql_count, squelched;
    // fft.c
    float fft_max_value;                 // maxValueFFT
    uint32_t fft_max_errors;             // maxValueErrors (fed back from the display pass, fft.c:372)
    float fft_mean[kFftBins];            // FFTOutput_mean
    // ZoomFFT (fft.c:236-261)
    float zoom_biquad[2][4][4];          // IIR_biquad_Zoom_FFT_I/Q state: per stage x[n-1], x[n-2], y[n-1], y[n-2]
    float zoom_fir[2][3];                // decimZoomFFTI/QState: the 3 newest filtered samples
    float zoom_buf[2 * kFftSize];        // FFTInput_ZOOMFFT, interleaved re, im
};

// user-facing settings block (mirrors the TRX fields the path reads); declared in include/ua3reo_b200.h
struct RxSettings;

}  // namespace ua3
