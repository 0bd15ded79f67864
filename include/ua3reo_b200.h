/*
 * ua3reo_b200.h - C ABI of libua3reo_b200.so: the UA3REO receive path (FPGA DDC + STM32 audio/FFT
 * stage, TX DUC as mirror) batched over many channels that tap one shared 12-bit ADC stream, on
 * one NVIDIA B200 (sm_100a).  Plain C: pointers, sizes and integers only.
 *
 * The reference has no FFI: its path sits behind void functions over global buffers
 * (SURVEY.md 8b).  Each entry point below names the reference interface it stands in for
 * (file:line under the reference tree).  The firmware's own entry points (processRxAudio,
 * processTxAudio, FFT_doFFT, ReinitAudioFilters ...) implemented over this ABI for one receiver are in
 * ua3reo-ddc-transceiver_b200/host/ua3reo_fw_shim.c, which is compiled with the firmware's headers.
 *
 * Conventions
 *   - every function returns 0 on success or a negative UA3_E_* code; ua3reo_last_error() gives
 *     the message of the calling thread's last failure.  (The firmware's functions return void
 *     and report through flags - FPGA_Buffer_underrun, fpga.c:16 - which has no batched analogue.)
 *   - the library owns all device state; the caller owns every host buffer it passes in.
 *   - one host thread per context.  The DDC kernels of a context are ordered on one CUDA stream; the STM32 stage
 *     runs one push behind on a second stream and result copies on a third - every call below states what it
 *     waits for, and ua3reo_sync() waits for everything.
 *   - there is NO CPU fallback: without a usable CUDA device ua3reo_create() fails.
 */
#ifndef UA3REO_B200_H
#define UA3REO_B200_H
#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define UA3_OK 0
#define UA3_E_INVAL (-1)      /* bad argument */
#define UA3_E_CUDA (-2)       /* CUDA runtime error (see ua3reo_last_error) */
#define UA3_E_NODEV (-3)      /* no CUDA device / not an sm_100 device */
#define UA3_E_TOOBIG (-4)     /* push larger than the context's max_block_samples */
#define UA3_E_STATE (-5)      /* call sequence error (e.g. reading frames before any push) */

#define UA3_ADC_PER_FRAME 1024u   /* rx_cic R=512 (rx_cic.vhd:27) x rx_ciccomp /2 (rx_ciccomp.vhd:25) */
#define UA3_FRAME_BYTES 8u        /* stm32_interface.v:228-271 */
#define UA3_AUDIO_BLOCK 192u      /* audio_processor.h:10-12 FPGA_AUDIO_BUFFER_HALF_SIZE */
#define UA3_FFT_SIZE 512u         /* fft.h:10 */
#define UA3_FFT_BINS 256u         /* fft.h:12 FFT_PRINT_SIZE */

typedef struct ua3reo_ctx ua3reo_ctx;

/* Version / capability string, e.g. "ua3reo_b200 0.1 sm_100a". */
const char *ua3reo_version(void);
const char *ua3reo_last_error(void);

/* Replaces the power-on reset of the FPGA fabric + FPGA_Init()/initAudioProcessor()/FFT_Init()
 * (fpga.c:41-55, audio_processor.c:55-59, fft.c:185-210; init order main.c:188-193) for n_channels
 * independent receivers.  max_block_samples: largest ADC block one push may carry (multiple of
 * 1024; 0 selects 2^20).  All tuning words start at the FPGA's power-on value 620407
 * (stm32_interface.v:56). */
int ua3reo_create(int device, uint32_t n_channels, uint32_t max_block_samples, ua3reo_ctx **out);
int ua3reo_destroy(ua3reo_ctx *ctx);

/* reset_n of the RX chain (UA3REO.bdf RX_N net): clears NCO phases and every filter state. */
int ua3reo_reset(ua3reo_ctx *ctx);

/* Clocking class of the I/Q frame.  On the board the four filter modules share reset = RX_N / clk_enable = RX
 * (UA3REO.bdf) and run on three PLL clocks (MAIN_PLL.v:105-116), so the instant the MCU raises RX decides (a) which
 * of the compensator's two delay lines the first CIC output enters (rx_ciccomp.vhd:339-361), (b) how many 48 kHz
 * samples VOICE_I trails the Hilbert sum (serial MAC + output register, rx_hilb.vhd:907-947) and (c) whether the bus
 * read sees data_delay.v's output before or after its clock edge.  Running the reference's VHDL over all 1024 release
 * instants (tools/hdl_clocking_survey.py) gives six classes; the default is the most frequent one:
 *   align_b = 1 (30 of 32 instants; 0 for the other two), d_i = 3 (2 for 1 of 32), d_q = 129 (130 for a late bus read).
 * Applies to all channels; call before the first push or after ua3reo_reset(). */
int ua3reo_ddc_set_clocking(ua3reo_ctx *ctx, int align_b, int d_i, int d_q);
int ua3reo_ddc_get_clocking(const ua3reo_ctx *ctx, int *align_b, int *d_i, int *d_q);

/* Leaves n_sms streaming multiprocessors out of the persistent front kernel's grid for work the CALLER runs beside the
 * pushes - typically the NCCL kernel that broadcasts the next ADC block (a front CTA owns a whole SM, so a collective
 * kernel can only run on an SM the front kernel does not occupy; without this it waits for the gap between two pushes).
 * 0 (default) gives the front kernel every SM.  Independent of the SMs kept for the STM32 stage. */
int ua3reo_reserve_sms(ua3reo_ctx *ctx, int n_sms);

uint32_t ua3reo_n_channels(const ua3reo_ctx *ctx);
uint32_t ua3reo_max_block_samples(const ua3reo_ctx *ctx);

/* 22-bit NCO tuning words for channels [first, first+n).  Replaces the cmd-1 parameter frame
 * (stm32_interface.v:142-171 <- fpga.c:173-220). Takes effect at the next push. */
int ua3reo_set_fcw(ua3reo_ctx *ctx, uint32_t first, uint32_t n, const uint32_t *fcw22);
int ua3reo_get_fcw(ua3reo_ctx *ctx, uint32_t first, uint32_t n, uint32_t *fcw22);

/* uint32_t getPhraseFromFrequency(uint32_t freq) (functions.c:206-226): Nyquist-zone folding and
 * FCW = round(f / 49152000 * 2^22); *iq_swap receives TRX_IQ_swap.  Pure host function. */
uint32_t ua3reo_phrase_from_frequency(uint32_t freq_hz, int *iq_swap);
/* TRX_setFrequency for one channel: FCW from the above, and the channel's IQ-swap flag. */
int ua3reo_set_frequency(ua3reo_ctx *ctx, uint32_t channel, uint32_t freq_hz);

/* The FPGA receive chain for every channel over n new ADC samples (12-bit two's complement,
 * sign-extended to int16; ADC_INPUT[11..0], UA3REO.bdf:26).  Samples are buffered until whole
 * 1024-sample frames are available; *frames_out (optional) receives the number of 48 kHz frames
 * produced per channel by this call.  n + carried remainder must not exceed max_block_samples.
 *   ua3reo_ddc_push        : adc in host memory (copied to the device inside the call)
 *   ua3reo_ddc_push_device : adc already in device memory of the context's device
 * Both are asynchronous with respect to the host; ua3reo_sync() or a read waits.  A host buffer in pinned
 * memory must stay unchanged until then (whole-block host pushes are copied on a separate stream so that
 * the copy of block k+1 overlaps the kernels of block k).
 * DEVICE PUSH CONTRACT: a device push of whole 1024-sample frames from a 16-byte aligned pointer is consumed IN PLACE
 * (zero copy) by kernels on the context's stream (ua3reo_stream, created cudaStreamNonBlocking: it does NOT order
 * itself after the legacy default stream or any other stream).  The caller must therefore
 *   (1) make the context's stream wait for whatever produces adc_dev (cudaStreamWaitEvent on ua3reo_stream) before
 *       the call, and
 *   (2) leave the buffer unmodified and allocated until the push has completed (an event recorded on
 *       ua3reo_stream after the call, or ua3reo_sync).
 * The Python binding's Receiver.push() does both for torch tensors (wait_stream + record_stream + a kept reference),
 * and sharding.AdcBroadcaster does both around its NCCL broadcast. */
int ua3reo_ddc_push(ua3reo_ctx *ctx, const int16_t *adc_host, size_t n, size_t *frames_out);
int ua3reo_ddc_push_device(ua3reo_ctx *ctx, const int16_t *adc_dev, size_t n, size_t *frames_out);

/* The I/Q words of the last push, in the byte order the FPGA puts on the bus on command 4
 * (stm32_interface.v:228-271; consumer FPGA_fpgadata_getiq, fpga.c:286-401):
 *   SPEC_Q hi, lo, SPEC_I hi, lo, VOICE_Q hi, lo, VOICE_I hi, lo.
 * dst is [n_channels][n_frames][8] bytes; n_frames must equal the last push's frames_out. */
int ua3reo_ddc_read_frames(ua3reo_ctx *ctx, uint8_t *dst_host, size_t n_frames);
/* Same copy, enqueued on a second stream that waits for the last push only: returns immediately so that
 * the caller can issue the NEXT push while the frames travel to the host.  dst (pinned memory for a true
 * overlap) is valid after ua3reo_sync().  At most two reads may be outstanding; the push after next waits
 * for the older one because it overwrites those ring slots. */
int ua3reo_ddc_read_frames_async(ua3reo_ctx *ctx, uint8_t *dst_host, size_t n_frames);
/* Device view of the same data.  Each channel owns a ring of *ring_frames 8-byte frames (a power of two),
 * *channel_stride_bytes apart; the last push wrote *n_frames frames starting at ring index *first_frame. */
int ua3reo_ddc_frames_device(ua3reo_ctx *ctx, const uint8_t **base, size_t *first_frame, size_t *n_frames,
                             size_t *ring_frames, size_t *channel_stride_bytes);

int ua3reo_sync(ua3reo_ctx *ctx);
/* The context's CUDA stream (cudaStream_t), for callers that enqueue their own copies. */
int ua3reo_stream(ua3reo_ctx *ctx, void **stream);
/* The stream the pipelined result reads (ua3reo_ddc_read_frames_async, ua3reo_rx_read_*_async) run on, for callers that
 * chain their own work behind them - e.g. an NCCL gather of the spectra a read has just placed in device memory. */
int ua3reo_copy_stream(ua3reo_ctx *ctx, void **stream);
/* Number of kernels this context has launched so far (bench.py's gpu_launches). */
uint64_t ua3reo_launch_count(const ua3reo_ctx *ctx);

/* ---------------------------------------------------------------------------------------------
 * STM32 stage: processRxAudio() and FFT_doFFT() for every channel.
 * ------------------------------------------------------------------------------------------- */

/* The subset of `struct TRX_SETTINGS TRX` / `VFO` (settings.h:57-123) and of the TRX manager globals
 * (trx_manager.h:60-61) that processRxAudio()/FFT_doFFT() read, per channel.  Defaults (ua3reo_rx_defaults)
 * are the firmware's LoadSettings() values (settings.c:33-94). */
typedef struct ua3reo_rx_settings {
    uint8_t mode;              /* TRX_MODE_* (trx_manager.h:11-24): 0 LSB 1 USB 2 IQ 3 CW_L 4 CW_U 5 DIGI_L 6 DIGI_U 8 NFM 9 WFM 10 AM */
    uint8_t agc;               /* TRX.AGC */
    uint8_t agc_speed;         /* TRX.Agc_speed (agc.c:17) */
    uint8_t dnr;               /* TRX.DNR */
    uint8_t notch;             /* TRX.NotchFilter */
    uint8_t mute;              /* TRX.Mute */
    uint8_t volume;            /* TRX.Volume, percent */
    uint8_t rf_gain;           /* TRX.RF_Gain */
    uint8_t fm_sql_threshold;  /* TRX.FM_SQL_threshold */
    uint8_t fft_enabled;       /* TRX.FFT_Enabled */
    uint8_t fft_averaging;     /* TRX.FFT_Averaging */
    uint8_t fft_zoom;          /* TRX.FFT_Zoom: 1, 2, 4, 8 or 16 (ZoomFFT, fft.c:236-261) */
    uint8_t iq_swap;           /* TRX_IQ_swap (functions.c:211-223) */
    uint8_t cw_decoder;        /* TRX.CWDecoder: run the CW decoder's Goertzel front end in CW_L / CW_U (cw_decoder.c:43-66) */
    uint8_t reserved[2];
    uint16_t filter_width;     /* CurrentVFO()->Filter_Width: 0 (LPF off) or one of the 32 table widths (audio_filters.c:59-122) */
    uint16_t ssb_hpf_pass;     /* TRX.SSB_HPF_pass: 60/100/200/300/400/500 */
    uint16_t notch_fc;         /* TRX.NotchFC, Hz */
    uint16_t reserved2;
} ua3reo_rx_settings;

void ua3reo_rx_defaults(ua3reo_rx_settings *s);

/* initAudioProcessor() + FFT_Init() (audio_processor.c:55-59, fft.c:185-210) for every channel, and from
 * then on every ua3reo_ddc_push*() also runs the STM32 stage over the whole 192-sample audio blocks and
 * 512-sample FFT frames that have become available (frames are buffered across pushes).  */
int ua3reo_rx_enable(ua3reo_ctx *ctx, int enable);

/* The STM32 stage alone over I/Q frames that come from elsewhere - what FPGA_fpgadata_getiq() (fpga.c:286-401)
 * receives from a real FPGA, or BASELINE config 1's synthetic 48 kSPS I/Q: frames_host is [n_channels][n][8] bytes in
 * the bus order of stm32_interface.v:228-271; they enter the same ring the DDC writes and are processed like a push. */
int ua3reo_rx_push_frames(ua3reo_ctx *ctx, const uint8_t *frames_host, size_t n);

/* TRX_setMode()/ReinitAudioFilters()/InitNotchFilter()/InitAGC() for channels [first, first+n)
 * (trx_manager.c:193-220, audio_filters.c:141-346, agc.c:14-19).  As in the firmware, selecting filters
 * clears the lattice filter states of the channel.  Fails with UA3_E_INVAL for a filter width, HPF
 * corner or mode the firmware has no table/branch for. */
int ua3reo_rx_set(ua3reo_ctx *ctx, uint32_t first, uint32_t n, const ua3reo_rx_settings *settings);
/* The fields processRxAudio()/FFT_doFFT() read from TRX on EVERY call (mode, AGC/DNR/notch switches, volume, mute,
 * RF gain, squelch threshold, FFT averaging/enable, IQ swap, CW decoder; `Filter_Width > 0`, audio_processor.c:448),
 * applied without what ReinitAudioFilters()/InitNotchFilter()/InitAGC()/FFT_Init() do: filter tables, notch
 * coefficients, AGC step sizes and the ZoomFFT decimator stay as the last full setter left them and no filter state
 * is cleared. */
int ua3reo_rx_set_live(ua3reo_ctx *ctx, uint32_t first, uint32_t n, const ua3reo_rx_settings *settings);
/* InitNotchFilter() (audio_filters.c:341-346) alone: recomputes the notch biquad for notch_fc[i] Hz (one per channel
 * of the range); like the firmware it leaves the biquad states as they are. */
int ua3reo_rx_set_notch(ua3reo_ctx *ctx, uint32_t first, uint32_t n, const uint16_t *notch_fc);
/* InitAGC() (agc.c:14-19) alone: the AGC step sizes for agc_speed[i] (one per channel of the range, > 0).  Like the
 * notch corner, TRX.Agc_speed is not read per call: ua3reo_rx_set_live() leaves the step sizes alone. */
int ua3reo_rx_set_agc_speed(ua3reo_ctx *ctx, uint32_t first, uint32_t n, const uint8_t *agc_speed);
/* FFT_Init() (fft.c:185-210) alone: ZoomFFT factor fft_zoom[i] (1, 2, 4, 8, 16; one per channel of the range).  As in
 * the firmware, a zoom above 1 clears the biquad and FIR-decimator states every time it is called, the accumulated
 * zoom buffer and the averaged spectrum are kept. */
int ua3reo_rx_fft_init(ua3reo_ctx *ctx, uint32_t first, uint32_t n, const uint8_t *fft_zoom);

/* Results of the last push.  Audio: what processRxAudio() leaves in Processor_AudioBuffer_A/B
 * (audio_processor.c:377-394): per channel and block 384 int32, L/R interleaved.  dst is
 * [n_channels][n_blocks][384]; n_blocks must equal *audio_blocks of ua3reo_rx_counts(). */
int ua3reo_rx_counts(ua3reo_ctx *ctx, size_t *audio_blocks, size_t *fft_frames);
int ua3reo_rx_read_audio(ua3reo_ctx *ctx, int32_t *dst_host, size_t n_blocks);
/* Pipelined variants: the copy is enqueued behind the STM32 stage of the last push (which runs on its own stream, one
 * push behind the DDC) and the call returns immediately; dst is pinned host memory or device memory (the copy
 * kind is inferred) and is valid after ua3reo_sync() / for work enqueued on ua3reo_copy_stream().  The next push's DDC kernels overlap both the stage and the copy. */
int ua3reo_rx_read_audio_async(ua3reo_ctx *ctx, int32_t *dst_host, size_t n_blocks);
int ua3reo_rx_read_spectra_async(ua3reo_ctx *ctx, float *dst_host, size_t n_frames);
/* The same audio as processRxAudio() hands to the USB audio class (audio_processor.c:415-432): volume undone,
 * int16, L/R interleaved.  dst is [n_channels][n_blocks][384] int16. */
int ua3reo_rx_read_audio_usb(ua3reo_ctx *ctx, int16_t *dst_host, size_t n_blocks);
/* Spectra: FFTOutput_mean (fft.c:27,324-328) after each FFT_doFFT(): dst is [n_channels][n_frames][256] float. */
int ua3reo_rx_read_spectra(ua3reo_ctx *ctx, float *dst_host, size_t n_frames);
/* Waterfall rows: what FFT_printFFT() writes into wtf_buffer[0] after each FFT (fft.c:361-379): per bin the
 * column height (uint16_t)(mean * 30) mapped through getFFTColor() (fft.c:503-538) to RGB565, red (0xF800) when
 * the column overflows, stored fft-shifted (bin x lands at x +/- 128).  dst is [n_channels][n_frames][256] uint16. */
int ua3reo_rx_read_waterfall(ua3reo_ctx *ctx, uint16_t *dst_host, size_t n_frames);
/* Waterfall history: wtf_buffer[FFT_WTF_HEIGHT = 50][256] (fft.c:29,353-379) of every channel, row 0 the newest;
 * dst is [n_channels][50][256] uint16. */
#define UA3_WTF_ROWS 50u
int ua3reo_rx_read_waterfall_history(ua3reo_ctx *ctx, uint16_t *dst_host);
/* A retune as FFT_printFFT() sees it (fft.c:347-351): freq_diff_hz[i] = new VFO frequency - previous one for channel
 * first+i.  The next FFT frame of the channel applies FFT_moveWaterfall() (fft.c:458-504) between its averaging and
 * its row: every stored row moves (freq_diff / 187) * FFT_Zoom pixels with zero fill and FFTOutput_mean is rotated in
 * place the way the firmware's loop does it.  Differences queued before that frame add up. */
int ua3reo_rx_move_waterfall(ua3reo_ctx *ctx, uint32_t first, uint32_t n, const int32_t *freq_diff_hz);
/* CW decoder front end (cw_decoder.c:43-66): the Goertzel magnitude at 350 Hz of every 192-sample block of the final
 * audio, for channels in CW_L / CW_U with cw_decoder set (0 elsewhere).  dst is [n_channels][n_blocks] float. */
int ua3reo_rx_read_cw(ua3reo_ctx *ctx, float *dst_host, size_t n_blocks);
/* The host half of the CW decoder: what CWDecoder_Process() (cw_decoder.c:69-170) does with the Goertzel magnitude of
 * a 192-sample block - adaptive threshold, 6 ms noise blanker, dit/dah and gap classification against a rolling dit time,
 * Morse table (:172-247).  One call per audio block and channel with that block's ua3reo_rx_read_cw() value and the
 * HAL_GetTick() time in ms (4 ms per block at 48 kHz); decoded characters (a space for a word gap) are written to out
 * (up to out_cap) and their number is returned; `wpm` is CW_Decoder_WPM.  Pure host function. */
typedef struct ua3reo_cw_decoder {
    float level_hi, level_lo;            /* adaptive keyed / idle magnitude levels (magnitudelimit, magnitudelimit_low) */
    uint8_t raw, raw_prev;               /* slicer output this block / last block */
    uint8_t key, key_prev;               /* debounced key state */
    uint8_t flushed;                     /* the pending letter has been written out after a long silence */
    uint8_t reserved[1];
    uint16_t wpm;                        /* CW_Decoder_WPM */
    int64_t t_raw_edge, t_key_down, t_key_up;   /* ms timestamps of the last slicer edge / key-down / key-up */
    int64_t mark_ms, space_ms, dit_ms;   /* last key-down and key-up durations, rolling dit length */
    char code[24];                       /* elements of the letter in progress, '.' and '-' */
} ua3reo_cw_decoder;
void ua3reo_cw_decoder_init(ua3reo_cw_decoder *st);
int ua3reo_cw_decoder_step(ua3reo_cw_decoder *st, float magnitude, uint32_t tick_ms, char *out, int out_cap);

/* ADC_MIN / ADC_MAX tracking of stm32_interface.v:384-397 over the ADC samples pushed since the last reset (the FPGA
 * resets to +2000 / -2000 when the MCU reads them with command 2, stm32_interface.v:172-205 <- fpga.c:222-284), and
 * the number of samples at either rail of the 12-bit range (what the AD9226 flags on its OTR pin). */
int ua3reo_adc_stats(ua3reo_ctx *ctx, int16_t *adc_min, int16_t *adc_max, uint32_t *n_rail, int reset);
/* GET PARAMS, command 2 of the wire protocol: the five bytes stm32_interface.v:172-205 returns (flags, ADC_MIN/ADC_MAX
 * nibbles and low bytes, encoder) built from the tracked ADC extremes, which are reset afterwards as the FPGA does, and
 * the two values FPGA_fpgadata_getparam() (fpga.c:222-284) decodes from them - TRX_ADC_MINAMPLITUDE sign-extended,
 * TRX_ADC_MAXAMPLITUDE not (a firmware quirk, kept).  ADC_OTR (bit 0) = a sample reached a rail since the last read;
 * DAC_OTR (bit 1) = dac_otr as given by the caller (e.g. a change of ua3reo_duc_read_otr()); key/encoder bits are 0.
 * Byte 4 carries bits 7:5 of byte 3, as on the wire: the state machine assigns only DATA_BUS_OUT[4:0] there (k == 204) and
 * the firmware masks them off.  (Bits 7:6 of byte 0 are likewise left over on the board - from the last byte of whatever
 * command came before, so they are not part of this packet's definition and are returned as 0.)
 * packet and either amplitude pointer may be NULL. */
int ua3reo_get_params(ua3reo_ctx *ctx, uint8_t packet[5], int16_t *adc_min_amplitude, int16_t *adc_max_amplitude, int dac_otr);
/* TRX_DoAutoGain() (trx_manager.c:268-356, called every 100 ms from stm32f4xx_it.c:422): the ATT / preamp decision
 * state machine driven by TRX_ADC_MAXAMPLITUDE.  Pure host function; `stage` and `wait` are autogain_stage and
 * autogain_wait_reaction, the four outputs are TRX.Preamp, TRX.ATT, TRX.LPF and TRX.BPF. */
typedef struct ua3reo_autogain {
    uint8_t stage, wait;
    uint8_t preamp, att, lpf, bpf;
} ua3reo_autogain;
void ua3reo_autogain_init(ua3reo_autogain *st);
void ua3reo_autogain_step(ua3reo_autogain *st, int16_t adc_max_amplitude);
/* The stages of processRxAudio() that the firmware also exports as functions of their own, run on a caller buffer with
 * (and updating) the channel's state - for callers that use them individually; processRxAudio itself is ua3reo_ddc_push /
 * ua3reo_rx_push_frames.  buf_host is read and, except for UA3_STAGE_DNR, overwritten in place.
 *   UA3_STAGE_DC_FILTER  void dc_filter(float32_t *buf, int16_t n, uint8_t stateNum)      audio_filters.h:51, audio_filters.c:358-373; arg = stateNum 0..5
 *   UA3_STAGE_AGC        void DoAGC(float32_t *buf, int16_t n)                             agc.h:9, agc.c:21-67
 *   UA3_STAGE_DNR        void processNoiseReduction(float32_t *in, float32_t *out)         noise_reduction.h:16, noise_reduction.c:25-37; n = 64,
 *                        the prediction goes to out_host (which may equal buf_host); nothing happens when the channel's DNR is off */
#define UA3_STAGE_DC_FILTER 0
#define UA3_STAGE_AGC 1
#define UA3_STAGE_DNR 2
int ua3reo_rx_stage(ua3reo_ctx *ctx, uint32_t channel, int stage, float *buf_host, float *out_host, size_t n, int arg);

/* TRX_RX_dBm of the 100 ms housekeeping tick (stm32f4xx_it.c:398-409) from the S-meter extremes: pure host function. */
int16_t ua3reo_smeter_dbm(float sample_max, float sample_min, uint8_t rf_gain);
/* S-meter accumulators Processor_RX_Audio_Samples_MAX/MIN_value (audio_processor.c:491-501): dst [n_channels][2];
 * reset != 0 clears them afterwards, as the 1 s housekeeping tick does (stm32f4xx_it.c:398-409). */
int ua3reo_rx_read_smeter(ua3reo_ctx *ctx, float *dst_host, int reset);

/* ---------------------------------------------------------------------------------------------
 * Transmit DUC (mirror of the DDC): tx_ciccomp (x2) -> tx_cic (x512) -> tx_mixer (I*sin14, Q*cos14) ->
 * tx_summator -> DAC_corrector  (FPGA/tx_ciccomp.vhd, tx_cic.vhd, tx_mixer.v:64-71, tx_summator.v:74-80,
 * DAC_corrector.v:15-21).  Every channel uses its DDC tuning word (the FPGA has one NCO for RX and TX).
 * ------------------------------------------------------------------------------------------- */
/* Allocates the DUC for pushes of up to max_tx_samples 48 kHz I/Q samples (each makes 1024 DAC words). */
int ua3reo_duc_enable(ua3reo_ctx *ctx, uint32_t max_tx_samples);
/* The TX I/Q words the MCU sends on command 3 (stm32_interface.v:206-227 <- FPGA_fpgadata_sendiq, fpga.c:403-436):
 * iq_host is [n_channels][n][2] int16 (I, Q).  Produces n*1024 DAC words per channel. */
int ua3reo_duc_push(ua3reo_ctx *ctx, const int16_t *iq_host, size_t n);
/* The same samples in the byte order of the wire: wire_host is [n_channels][n][4] bytes, for every sample
 * Q hi, Q lo, I hi, I lo - what FPGA_fpgadata_sendiq() writes after command 3 (fpga.c:403-436) and stm32_interface.v:206-227
 * reassembles into Q_HOLD / I_HOLD -> TX_Q / TX_I.  The mirror of the 8-byte RX frame of ua3reo_ddc_read_frames. */
int ua3reo_duc_push_wire(ua3reo_ctx *ctx, const uint8_t *wire_host, size_t n);
/* The 14-bit offset-binary words DAC_corrector.v:15-21 drives onto the DAC pins: dst [n_channels][n*1024] uint16. */
int ua3reo_duc_read_dac(ua3reo_ctx *ctx, uint16_t *dst_host, size_t n);
int ua3reo_duc_dac_device(ua3reo_ctx *ctx, const uint16_t **base, size_t *n_words, size_t *channel_stride_words);
/* tx_summator overflow count per channel since reset (the DAC_OTR flag, stm32_interface.v:172-205): dst [n_channels]. */
int ua3reo_duc_read_otr(ua3reo_ctx *ctx, uint32_t *dst_host);

/* ---------------------------------------------------------------------------------------------
 * STM32 transmit audio: processTxAudio() (audio_processor.c:61-273) for every channel - DC filter, lattice
 * HPF/LPF, ALC compressor, SSB (+/-45 degree Hilbert FIR pair), AM, FM and CW modulators.  Its I/Q output is
 * what FPGA_fpgadata_sendiq() (fpga.c:403-436) puts on the bus and the DUC above consumes.
 * ------------------------------------------------------------------------------------------- */
typedef struct ua3reo_tx_settings {
    uint8_t mode;            /* TRX_MODE_*: LSB USB IQ CW_L CW_U DIGI_L DIGI_U NO_TX NFM WFM AM LOOPBACK (trx_manager.h:11-24) */
    uint8_t mute;            /* TRX.Mute */
    uint8_t tune;            /* TRX_tune (carrier at the selected power) */
    uint8_t key_down;        /* TRX_key_serial || TRX_ptt_hard || TRX_key_hard (CW keying, audio_processor.c:146) */
    uint8_t rf_power;        /* TRX.RF_Power, percent of MAX_TX_AMPLITUDE (settings.h:13) */
    uint8_t volume;          /* TRX.Volume: output level of TRX_MODE_LOOPBACK (audio_processor.c:231) */
    uint8_t reserved[2];
    uint16_t filter_width;   /* CurrentVFO()->Filter_Width (also selects the FM modulation index, :597-605) */
    uint16_t ssb_hpf_pass;   /* TRX.SSB_HPF_pass */
} ua3reo_tx_settings;

void ua3reo_tx_defaults(ua3reo_tx_settings *s);
int ua3reo_tx_enable(ua3reo_ctx *ctx, uint32_t max_blocks);
int ua3reo_tx_set(ua3reo_ctx *ctx, uint32_t first, uint32_t n, const ua3reo_tx_settings *settings);
/* Like ua3reo_rx_set_live(): mode, mute, tune, key, power and the FM index follow the settings, the lattice tables
 * selected by the last ua3reo_tx_set() (ReinitAudioFilters) stay and no filter state is cleared. */
int ua3reo_tx_set_live(ua3reo_ctx *ctx, uint32_t first, uint32_t n, const ua3reo_tx_settings *settings);
/* n_blocks x 192 codec samples per channel: mic_host is [n_channels][n_blocks*192][2] int16 (left, right), the
 * layout of CODEC_Audio_Buffer_TX (wm8731.h:14).  One processTxAudio() per block. */
int ua3reo_tx_process(ua3reo_ctx *ctx, const int16_t *mic_host, size_t n_blocks);
/* I/Q of the last call: iq_words [n_channels][n_blocks*192][2] int16 (I, Q as sent on the wire) and/or iq_float
 * (same shape, FPGA_Audio_SendBuffer_I/Q contents); either may be NULL. */
int ua3reo_tx_read_iq(ua3reo_ctx *ctx, int16_t *iq_words, float *iq_float, size_t n_blocks);
/* TRX_MODE_LOOPBACK (audio_processor.c:228-249): the block does not go to the FPGA but, scaled by Volume / 50, to the codec.
 * dst [n_channels][n_blocks*192][2] int32 (left, right = left) - what the firmware copies into CODEC_Audio_Buffer_RX; zeros for
 * channels in any other mode. */
int ua3reo_tx_read_loopback(ua3reo_ctx *ctx, int32_t *dst_host, size_t n_blocks);
/* Runs the DUC over the I/Q words of the last ua3reo_tx_process() without leaving the device. */
int ua3reo_tx_feed_duc(ua3reo_ctx *ctx);

/* ---------------------------------------------------------------------------------------------
 * Several GPUs from one C host: a bank shards its channels over devices in contiguous slabs (no term of the path couples
 * two channels) and fans every ADC block out to them.  The fan-out is cudaMemcpyPeerAsync from the ingest
 * device - copy engines over NVLink, no SM involved, so it runs under the previous block's kernels (a collective KERNEL
 * cannot: a front CTA owns its whole SM).  Results stay sharded on the devices and are read slab by slab, every device
 * over its own PCIe link.  (bench.py is the multi-PROCESS form of the same thing, one rank per GPU with an NCCL broadcast,
 * as its launch contract prescribes.)
 * ------------------------------------------------------------------------------------------- */
typedef struct ua3reo_bank ua3reo_bank;
/* n_devices contexts on devices[0..n_devices) (a device may be named twice), n_channels in total, slabs differing by at
 * most one channel; devices[0] ingests the ADC stream. */
int ua3reo_bank_create(int n_devices, const int *devices, uint32_t n_channels, uint32_t max_block_samples, ua3reo_bank **out);
int ua3reo_bank_destroy(ua3reo_bank *bank);
int ua3reo_bank_n_devices(const ua3reo_bank *bank);
/* the context of device slot i and its channel slab [*first, *first + *count): every per-channel call of this header
 * (ua3reo_rx_set, ua3reo_set_fcw, the reads ...) can be made on it with slab-relative channel numbers */
int ua3reo_bank_context(ua3reo_bank *bank, int i, ua3reo_ctx **ctx, uint32_t *first, uint32_t *count);
/* bank-wide forms of the calls a host loop needs: tuning words and RX settings by global channel number, push, reads into
 * [n_channels][...] host arrays (each device writes its slab), sync */
int ua3reo_bank_set_fcw(ua3reo_bank *bank, uint32_t first, uint32_t n, const uint32_t *fcw22);
int ua3reo_bank_rx_enable(ua3reo_bank *bank, int enable);
int ua3reo_bank_rx_set(ua3reo_bank *bank, uint32_t first, uint32_t n, const ua3reo_rx_settings *settings);
/* adc_host: n samples in host memory (pinned for an asynchronous copy); every device processes all of them for its slab */
int ua3reo_bank_push(ua3reo_bank *bank, const int16_t *adc_host, size_t n, size_t *frames_out);
int ua3reo_bank_read_frames(ua3reo_bank *bank, uint8_t *dst_host, size_t n_frames);
int ua3reo_bank_rx_counts(ua3reo_bank *bank, size_t *audio_blocks, size_t *fft_frames);
int ua3reo_bank_rx_read_audio(ua3reo_bank *bank, int32_t *dst_host, size_t n_blocks);
int ua3reo_bank_rx_read_spectra(ua3reo_bank *bank, float *dst_host, size_t n_frames);
int ua3reo_bank_sync(ua3reo_bank *bank);

/* ---------------------------------------------------------------------------------------------
 * The same fan-out for one PROCESS per GPU (the torch.distributed launch; a C host with one process per device): no
 * collective kernel and no SM.  Every rank creates its end, the 64-byte handles are exchanged by whatever transport the
 * host has (bench.py: one all_gather at start-up), and from then on a block travels as
 *     ingest rank:  ua3reo_fanout_send(f, block)                    copy engines write the block into every rank's slot over
 *                                                                   NVLink (CUDA IPC mappings) and then store its number there
 *     every rank:   ua3reo_fanout_acquire(f, stream, &blk)           `stream` waits for that number (cuStreamWaitValue32)
 *                   ua3reo_ddc_push_device(ctx, blk, n, &frames)     consumed in place (DEVICE PUSH CONTRACT above)
 *                   ua3reo_fanout_release(f, stream)                 the slot's credit goes back behind the push's kernels
 * Nothing here orders the processes' HOST threads: the waits are executed by the streams.  Sends run `n_buffers` - 1 blocks
 * ahead of the slowest consumer.  One object is driven by one host thread at a time (like a context); acquire and release
 * alternate; every rank makes the same sequence of calls (a rank that stops consuming stalls the ingest rank's copy stream
 * once its credits are used up, nothing is dropped or overwritten).  The reference has one board = one ADC = one channel (fpga.c:286-401 reads it frame by
 * frame); this is the path's only exchange when its channels are spread over devices.
 * ------------------------------------------------------------------------------------------- */
typedef struct ua3reo_fanout ua3reo_fanout;
#define UA3_FANOUT_HANDLE_BYTES 64u
/* rank of world, `src` ingests; slots of block_samples int16 each, 2 <= n_buffers <= 16.  world == 1 needs no connect. */
int ua3reo_fanout_create(int device, int rank, int world, int src, size_t block_samples, int n_buffers, ua3reo_fanout **out);
/* Tear-down in two steps, because an arena must outlive its peers' mappings: every rank disconnects (waits for its own
 * fan-out streams, unmaps its peers) once all ranks' consumer streams are idle, then - after a host barrier - destroys. */
int ua3reo_fanout_disconnect(ua3reo_fanout *f);
int ua3reo_fanout_destroy(ua3reo_fanout *f);
/* this rank's handle (UA3_FANOUT_HANDLE_BYTES) / all ranks' handles in rank order (world x UA3_FANOUT_HANDLE_BYTES) */
int ua3reo_fanout_handle(ua3reo_fanout *f, void *handle64);
int ua3reo_fanout_connect(ua3reo_fanout *f, const void *handles);
/* ingest rank only; block: block_samples int16 in device or pinned host memory, valid until the copy has run; asynchronous */
int ua3reo_fanout_send(ua3reo_fanout *f, const int16_t *block, size_t n);
int ua3reo_fanout_acquire(ua3reo_fanout *f, void *consumer_stream, const int16_t **block_dev);
int ua3reo_fanout_release(ua3reo_fanout *f, void *consumer_stream);
int ua3reo_fanout_sync(ua3reo_fanout *f);      /* waits for this rank's send and credit streams */
/* direct_remote_store: 1 if the driver stores flag words straight onto peer mappings, 0 if they are staged through a copy */
int ua3reo_fanout_info(const ua3reo_fanout *f, int *direct_remote_store, uint64_t *n_sent, uint64_t *n_acquired);

/* The opposite direction, same means: every rank's slab of results (bench.py: the spectra of a push, which north_star wants
 * back on one device) lands in ONE buffer on the root rank - each rank's copy engine writes its slab into its place of the
 * root's slot and stores an arrival number there; the root's consumer stream waits for all of them; credits flow back when the
 * root has consumed the slot.  No collective kernel.
 *     every rank:  ua3reo_gather_send(g, slab_dev, stream)          enqueued ON `stream`, behind whatever produced the slab
 *     root rank:   ua3reo_gather_acquire(g, stream, &all, &stride)   all + r * stride = rank r's slab; `stream` waits for all ranks
 *                  ... use it on `stream` ...
 *                  ua3reo_gather_release(g, stream) */
typedef struct ua3reo_gather ua3reo_gather;
int ua3reo_gather_create(int device, int rank, int world, int root, size_t slab_bytes, int n_buffers, ua3reo_gather **out);
int ua3reo_gather_disconnect(ua3reo_gather *g);   /* two-step tear-down as for the fan-out */
int ua3reo_gather_destroy(ua3reo_gather *g);
int ua3reo_gather_handle(ua3reo_gather *g, void *handle64);
int ua3reo_gather_connect(ua3reo_gather *g, const void *handles);
int ua3reo_gather_send(ua3reo_gather *g, const void *slab_dev, void *producer_stream);
int ua3reo_gather_acquire(ua3reo_gather *g, void *consumer_stream, const void **all_dev, size_t *rank_stride);
int ua3reo_gather_release(ua3reo_gather *g, void *consumer_stream);

/* Per-kernel device timing with CUDA events on the context's stream (bench.py's roofline):
 * after ua3reo_profile_begin(ctx, max_blocks) each processed ADC block records an event before and
 * after every kernel; ua3reo_profile_end() waits for the stream and sums the elapsed times per
 * kernel over the recorded blocks.  kernel_ms[0..4], in launch order = ADC widening (int16 -> int32 << 9), front
 * (NCO+mixer+CIC integrators), cic combs + compensator FIR, Hilbert+delay+frame pack, state rotate; [5..6] = rx_audio,
 * rx_fft when the STM32 stage is enabled. */
#define UA3_DDC_KERNELS 5u
#define UA3_PROFILED_KERNELS 7u
int ua3reo_profile_begin(ua3reo_ctx *ctx, uint32_t max_blocks);
/* As above with only the two events around kernel slot `kernel` (1 = front): the other slots read 0. */
int ua3reo_profile_begin_kernel(ua3reo_ctx *ctx, uint32_t max_blocks, uint32_t kernel);
int ua3reo_profile_end(ua3reo_ctx *ctx, double *kernel_ms, uint32_t n_kernels, uint32_t *blocks);

/* Dependent-free INT32 issue-rate microbenchmark (IADD3 + IMAD interleaved) used as the roofline
 * denominator: returns integer operations per second summed over the whole device. */
int ua3reo_measure_int32_peak(int device, double *ops_per_s);
/* Shared-memory wavefront rate (conflict-free LDS.32 streams, 1024 threads per SM): the roofline denominator of the
 * tensor-core front kernel, whose NCO table look-ups saturate that pipe.  Wavefronts per second over the whole device. */
int ua3reo_measure_lds_peak(int device, double *wavefronts_per_s);

#ifdef __cplusplus
}
#endif
#endif
