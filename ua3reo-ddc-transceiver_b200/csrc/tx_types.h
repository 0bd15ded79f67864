// tx_types.h - per-channel parameter / state blocks of the STM32 transmit-audio stage (processTxAudio).
#pragma once
#include <stdint.h>
#include "rx_types.h"

namespace ua3 {

constexpr int kTxHilbTaps = 201;     // IQ_TX_HILBERT_TAPS (audio_filters.h:11)

struct TxParams {
    uint8_t mode, mute, tune, key_down;   // key_down = TRX_key_serial || TRX_ptt_hard || TRX_key_hard (audio_processor.c:146)
    uint8_t lpf_on, hpf_set, pad[2];
    float amplitude;                      // TRX.RF_Power / 100.0f * MAX_TX_AMPLITUDE (audio_processor.c:65)
    float fm_index;                       // ModulateFM's modulation_index for the channel's Filter_Width (:597-605)
    float loop_volume;                    // (float32_t)TRX.Volume / 50.0f: the loopback branch's output volume (audio_processor.c:231)
    float lpf_k[kLpfMax], lpf_v[kLpfMax + 1], hpf_k[kHpfStages], hpf_v[kHpfStages + 1];
};

struct TxState {
    float dc_x[2], dc_y[2];               // dc_filter_state[2], [3]
    float lpf_g[kLpfMax], hpf_g[kHpfStages];
    float alc_gain;                       // ALC_need_gain (starts at 1.0f, audio_processor.c:35)
    float fm_hpf_a, fm_hpf_b;             // ModulateFM statics
    uint32_t fm_accum;
    float fir_hist[2][kTxHilbTaps - 1];   // previous 200 inputs of FIR_TX_Hilbert_I / _Q (oldest first)
};

}  // namespace ua3
