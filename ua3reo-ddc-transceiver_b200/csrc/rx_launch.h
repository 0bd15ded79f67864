// rx_launch.h - host-visible launch interface of rx.cu
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "rx_types.h"

namespace ua3 {

struct RxBuffers {
    uint32_t n_ch = 0;
    const uint64_t* frames = nullptr;   // the DDC's frame ring: [n_ch][ring] 8-byte frames
    uint32_t ring_mask = 0, frame_ch_stride = 0;
    RxParams* params = nullptr;         // [n_ch]
    RxState* state = nullptr;           // [n_ch]
    uint32_t* order = nullptr;          // [n_ch] channel visited by slot i: groups equal settings into the same warp
    int32_t* audio_out = nullptr;       // [n_ch][max_audio_blocks][384]
    uint32_t audio_ch_stride = 0, max_audio_blocks = 0;
    float* cw_mag = nullptr;            // [n_ch][max_audio_blocks] Goertzel magnitude (CW decoder front end)
    bool split_audio = false;           // run processRxAudio as two kernels instead of the warp-specialised one (UA3REO_RX_SPLIT=1)
    float* scratch = nullptr;           // [3][max_audio_blocks * 192][n_ch padded to 32]: block buffers between rx_filter_kernel and rx_post_kernel
    float* fft_in = nullptr;            // [max_fft_frames][2 rails][512][n_ch padded to 32]: FFT input after the sequential filters (rx_fft_pre_kernel)
    float* spectra = nullptr;           // [n_ch][max_fft_frames][256]
    uint16_t* waterfall = nullptr;      // [n_ch][max_fft_frames][256] RGB565 rows, fft-shifted
    uint32_t spec_ch_stride = 0, max_fft_frames = 0;
    // wtf_buffer[FFT_WTF_HEIGHT][FFT_PRINT_SIZE] (fft.c:29) per channel, kept as a ring: firmware row y is ring row
    // (wtf_head + y) % 50, so the "shift every row down by one" of fft.c:353-358 is a head decrement
    uint16_t* wtf_hist = nullptr;       // [n_ch][50][256]
    uint32_t* wtf_head = nullptr;       // [n_ch]
    int32_t* wtf_pending_hz = nullptr;  // [n_ch] CurrentVFO()->Freq - currentFFTFreq not yet applied (fft.c:347-351)
};
constexpr int kWtfRows = 50;            // FFT_WTF_HEIGHT (fft.h:14)

cudaError_t rx_upload_constants(const float* window, const float* twiddle, const uint16_t* colors, const float* zoom_biquad,
                                const float* zoom_fir);
cudaError_t adc_stats_launch(const int16_t* adc, uint32_t n, int32_t* stats, int sm_count, cudaStream_t st, int* launches);
cudaError_t rx_launch_usb_pack(const RxBuffers& b, uint32_t n_blocks, const float* undo_dev, int16_t* out_dev, cudaStream_t st,
                               int* launches);
cudaError_t rx_launch_audio(const RxBuffers& b, uint32_t start, uint32_t n_blocks, cudaStream_t st, int* launches);
size_t rx_scratch_floats(uint32_t n_ch, uint32_t max_audio_blocks);
size_t rx_fft_in_floats(uint32_t n_ch, uint32_t max_fft_frames);
int rx_audio_sms(uint32_t n_ch);        // SMs the STM32 audio kernels can keep busy (16 warps per SM)
cudaError_t rx_launch_fft(const RxBuffers& b, uint32_t start, uint32_t n_frames, cudaStream_t st, int* launches);
cudaError_t rx_launch_stage(const RxBuffers& b, int stage, float* buf_dev, float* out_dev, uint32_t n, uint32_t ch, int arg, cudaStream_t st,
                            int* launches);
cudaError_t rx_launch_clear(const RxBuffers& b, const uint8_t* flags_dev, uint32_t first, uint32_t n, cudaStream_t st,
                            int* launches);
cudaError_t rx_launch_init_state(const RxBuffers& b, cudaStream_t st, int* launches);

}  // namespace ua3
