/*
 * ua3reo_fw_shim.c - the firmware-side binding: the reference firmware's OWN entry points for the signal path,
 * implemented over libua3reo_b200.so (one channel, channel 0 of a one-channel receiver bank).
 *
 * It is compiled WITH the firmware's headers and linked INSTEAD OF audio_processor.c, audio_filters.c, agc.c,
 * noise_reduction.c, fft.c and cw_decoder.c; the rest of the firmware (fpga.c bus driver, functions.c, trx_manager.c,
 * the ISR cadence) stays as it is and keeps talking to the same globals:
 *
 *   firmware entry point                 (reference)                      GPU call
 *   initAudioProcessor()                 audio_processor.c:55-59          ua3reo_create + ua3reo_rx_enable + ua3reo_tx_enable
 *   ReinitAudioFilters()                 audio_filters.c:141-339          ua3reo_rx_set / ua3reo_tx_set (tables reselected, states cleared)
 *   InitNotchFilter()                    audio_filters.c:341-346          ua3reo_rx_set_notch (coefficients only)
 *   dc_filter() / DoAGC() / processNoiseReduction()   audio_filters.c:358-373, agc.c:21-67, noise_reduction.c:25-37   ua3reo_rx_stage
 *   processRxAudio()                     audio_processor.c:275-434        ua3reo_rx_push_frames(192) + ua3reo_rx_read_audio/_usb/_smeter/_cw
 *   processTxAudio()                     audio_processor.c:61-273         ua3reo_tx_process(1) + ua3reo_tx_read_iq
 *   FFT_Init() / FFT_doFFT()             fft.c:185-328                    second one-channel context: ua3reo_rx_push_frames(512) + read_spectra/_waterfall
 *   FFT_printFFT()/FFT_printWaterfallDMA fft.c:330-500 (display, out of scope) re-arm FFT_need_fft only
 *
 * Host data movement (which ring half to read, which output half to fill, the A/B buffer toggle) is the
 * firmware's and is restated here around the calls; every number comes from the device.  There is no CPU
 * fallback: when the library or a GPU is missing the shim prints the library's error and aborts.
 *
 * Built by oracle/ref_harness/Makefile into oracle/_ref/fw_rx_b200 / fw_tx_b200 (the reference's headers are
 * only present where /root/reference is); tests/test_fw_shim_gpu.py runs them beside the all-CPU firmware build.
 */
#include "stm32f4xx_hal.h"
#include "arm_math.h"
#include "settings.h"
#include "trx_manager.h"
#include "fpga.h"
#include "fft.h"
#include "wm8731.h"
#include "audio_processor.h"
#include "audio_filters.h"
#include "agc.h"
#include "noise_reduction.h"
#include "cw_decoder.h"
#include "functions.h"
#include "usbd_audio_if.h"
#include "ua3reo_b200.h"
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* ---- the globals the replaced translation units define (audio_processor.h, fft.h, audio_filters.h, cw_decoder.h) ---- */
volatile uint32_t AUDIOPROC_samples = 0;
volatile uint32_t AUDIOPROC_TXA_samples = 0;
volatile uint32_t AUDIOPROC_TXB_samples = 0;
int32_t Processor_AudioBuffer_A[FPGA_AUDIO_BUFFER_SIZE] = {0};
int32_t Processor_AudioBuffer_B[FPGA_AUDIO_BUFFER_SIZE] = {0};
volatile uint8_t Processor_AudioBuffer_ReadyBuffer = 0;
volatile bool Processor_NeedRXBuffer = false;
volatile bool Processor_NeedTXBuffer = false;
volatile float32_t Processor_AVG_amplitude = 0.0f;
volatile float32_t Processor_TX_MAX_amplitude = 0.0f;
volatile float32_t ALC_need_gain = 1.0f;
volatile float32_t ALC_need_gain_new = 1.0f;
float32_t FPGA_Audio_Buffer_Q_tmp[FPGA_AUDIO_BUFFER_HALF_SIZE] = {0};
float32_t FPGA_Audio_Buffer_I_tmp[FPGA_AUDIO_BUFFER_HALF_SIZE] = {0};
volatile float32_t fm_sql_avg = 0.0f;
volatile float32_t Processor_RX_Audio_Samples_MAX_value = 0.0f;
volatile float32_t Processor_RX_Audio_Samples_MIN_value = 0.0f;

volatile uint32_t FFT_buff_index = 0;
bool NeedFFTInputBuffer = true;
bool FFT_need_fft = true;
float32_t FFTInput_I[FFT_SIZE] = {0};
float32_t FFTInput_Q[FFT_SIZE] = {0};

volatile bool NeedReinitNotch = false;
volatile uint16_t CW_Decoder_WPM = 0;
char CW_Decoder_Text[CWDECODER_STRLEN] = {0};

/* ---- what the shim publishes beside the firmware's globals (display code and test harnesses read these) ---- */
float ua3reo_shim_fft_mean[FFT_PRINT_SIZE];       /* FFTOutput_mean (fft.c:27) after the last FFT_doFFT() */
uint16_t ua3reo_shim_wtf_row0[FFT_PRINT_SIZE];    /* wtf_buffer[0] (fft.c:29) after the last FFT_printFFT() */
uint16_t ua3reo_shim_wtf_buffer[UA3_WTF_ROWS * FFT_PRINT_SIZE];   /* wtf_buffer (fft.c:29), row 0 newest, after FFT_printFFT() */
float ua3reo_shim_cw_magnitude = 0.0f;            /* Goertzel magnitude of the last CW block (cw_decoder.c:56-66) */

static ua3reo_ctx *rx_ctx;      /* audio: processRxAudio / processTxAudio */
static ua3reo_ctx *fft_ctx;     /* panorama: FFT_doFFT */
static uint16_t latched_notch_fc = 1000;
static uint8_t latched_agc_speed = 3;    /* TRX.Agc_speed as of the last InitAGC() (agc.c:14-19) */
static uint32_t shim_fft_freq = 0;       /* currentFFTFreq (fft.c:31) */
static float smeter_seen_max, smeter_seen_min;
static ua3reo_rx_settings rx_last, fft_last;
static ua3reo_tx_settings tx_last;
static bool rx_last_valid, fft_last_valid, tx_last_valid;

static void die(const char *what, int rc)
{
    fprintf(stderr, "ua3reo_fw_shim: %s failed (%d): %s\n", what, rc, ua3reo_last_error());
    abort();
}
#define CHECK(call) do { int rc_ = (call); if (rc_) die(#call, rc_); } while (0)

static void ensure_contexts(void)
{
    if (rx_ctx) return;
    CHECK(ua3reo_create(0, 1, 1u << 19, &rx_ctx));     /* one block holds 512 frames */
    CHECK(ua3reo_create(0, 1, 1u << 19, &fft_ctx));
    CHECK(ua3reo_rx_enable(rx_ctx, 1));
    CHECK(ua3reo_rx_enable(fft_ctx, 1));
    CHECK(ua3reo_tx_enable(rx_ctx, 1));
}

/* the TRX fields the audio path reads, as they are right now */
static void gather_rx(ua3reo_rx_settings *s, bool for_fft)
{
    ua3reo_rx_defaults(s);
    s->mode = (uint8_t)TRX_getMode();
    s->filter_width = (uint16_t)CurrentVFO()->Filter_Width;
    s->ssb_hpf_pass = TRX.SSB_HPF_pass;
    s->agc = TRX.AGC; s->dnr = TRX.DNR;
    s->agc_speed = latched_agc_speed;
    s->notch = TRX.NotchFilter; s->notch_fc = latched_notch_fc;
    s->volume = TRX.Volume; s->mute = TRX.Mute; s->rf_gain = TRX.RF_Gain;
    s->fm_sql_threshold = TRX.FM_SQL_threshold;
    s->cw_decoder = TRX.CWDecoder;
    s->iq_swap = 0;                       /* FPGA_fpgadata_getiq() has already swapped into the rings (fpga.c:305-385) */
    s->fft_enabled = for_fft ? 1 : 0;     /* the audio context never runs the panorama and vice versa */
    s->fft_averaging = TRX.FFT_Averaging ? TRX.FFT_Averaging : 1;
    s->fft_zoom = TRX.FFT_Zoom ? TRX.FFT_Zoom : 1;
}

static void gather_tx(ua3reo_tx_settings *s)
{
    ua3reo_tx_defaults(s);
    s->mode = (uint8_t)TRX_getMode();
    s->filter_width = (uint16_t)CurrentVFO()->Filter_Width;
    s->ssb_hpf_pass = TRX.SSB_HPF_pass;
    s->mute = TRX.Mute; s->tune = TRX_tune; s->rf_power = TRX.RF_Power; s->volume = TRX.Volume;
    s->key_down = (TRX_key_serial || TRX_ptt_hard || TRX_key_hard) ? 1 : 0;      /* audio_processor.c:146 */
}

static bool tx_mode_ok(uint8_t mode) { (void)mode; return true; }      /* NO_TX and LOOPBACK run the `default:` branch (audio_processor.c:141-213) */

/* full = what TRX_setMode()/ReinitAudioFilters() do; otherwise only the per-call fields, and only when they moved */
static void apply_rx(bool full)
{
    ua3reo_rx_settings s;
    gather_rx(&s, false);
    if (full) CHECK(ua3reo_rx_set(rx_ctx, 0, 1, &s));
    else if (!rx_last_valid || memcmp(&s, &rx_last, sizeof s)) CHECK(ua3reo_rx_set_live(rx_ctx, 0, 1, &s));
    rx_last = s; rx_last_valid = true;
}

static void apply_fft(bool full)
{
    ua3reo_rx_settings s;
    gather_rx(&s, true);
    s.mode = TRX_MODE_IQ; s.filter_width = 0; s.dnr = 0; s.cw_decoder = 0;      /* its audio output is never read */
    if (full) CHECK(ua3reo_rx_set(fft_ctx, 0, 1, &s));
    else if (!fft_last_valid || memcmp(&s, &fft_last, sizeof s)) CHECK(ua3reo_rx_set_live(fft_ctx, 0, 1, &s));
    fft_last = s; fft_last_valid = true;
}

static void apply_tx(bool full)
{
    ua3reo_tx_settings s;
    gather_tx(&s);
    if (!tx_mode_ok(s.mode)) return;
    if (full) CHECK(ua3reo_tx_set(rx_ctx, 0, 1, &s));
    else if (!tx_last_valid || memcmp(&s, &tx_last, sizeof s)) CHECK(ua3reo_tx_set_live(rx_ctx, 0, 1, &s));
    tx_last = s; tx_last_valid = true;
}

/* ---- audio_filters.h / agc.h / noise_reduction.h entry points ---- */
void InitNotchFilter(void)               /* audio_filters.c:341-346: coefficients only, no state cleared */
{
    NeedReinitNotch = false;
    latched_notch_fc = TRX.NotchFC;
    if (!rx_ctx) return;
    CHECK(ua3reo_rx_set_notch(rx_ctx, 0, 1, &latched_notch_fc));
    CHECK(ua3reo_rx_set_notch(fft_ctx, 0, 1, &latched_notch_fc));
    rx_last.notch_fc = fft_last.notch_fc = latched_notch_fc;
}

void ReinitAudioFilters(void)
{
    ensure_contexts();
    apply_rx(true);
    apply_tx(true);
}

void InitAGC(void)                       /* agc.c:14-19: step sizes from TRX.Agc_speed, latched here like the firmware's statics */
{
    latched_agc_speed = TRX.Agc_speed ? TRX.Agc_speed : 1;   /* the menus keep it in 1..4; 0 would divide by zero and the library rejects it */
    if (!rx_ctx) return;
    CHECK(ua3reo_rx_set_agc_speed(rx_ctx, 0, 1, &latched_agc_speed));
    rx_last.agc_speed = latched_agc_speed;
}
void InitNoiseReduction(void) {}         /* noise_reduction.c:18-23: state is created zeroed with the context */

/* The sub-stages the firmware exports as functions of their own (audio_filters.h:51, agc.h:9, noise_reduction.h:16): on the
 * caller's buffer, with channel 0's state on the device - the same state processRxAudio() uses, as in the firmware. */
void dc_filter(float32_t *agcBuffer, int16_t blockSize, uint8_t stateNum)
{
    ensure_contexts();
    CHECK(ua3reo_rx_stage(rx_ctx, 0, UA3_STAGE_DC_FILTER, agcBuffer, NULL, (size_t)blockSize, stateNum));
}

void DoAGC(float32_t *agcBuffer, int16_t blockSize)
{
    ensure_contexts();
    apply_rx(false);                     /* TRX.AGC and the mode are read on every call (agc.c:46) */
    CHECK(ua3reo_rx_stage(rx_ctx, 0, UA3_STAGE_AGC, agcBuffer, NULL, (size_t)blockSize, 0));
}

void processNoiseReduction(float32_t *bufferIn, float32_t *bufferOut)
{
    ensure_contexts();
    apply_rx(false);                     /* `if (!TRX.DNR) return;` (noise_reduction.c:27) */
    CHECK(ua3reo_rx_stage(rx_ctx, 0, UA3_STAGE_DNR, bufferIn, bufferOut, NOISE_REDUCTION_BLOCK_SIZE, 0));
}

void InitAudioFilters(void)              /* audio_filters.c:124-139: lattice/biquad instances, then InitNotchFilter() */
{
    ensure_contexts();                   /* TX Hilbert pair, squelch HPF and DNR state are created zeroed with the context */
    InitNoiseReduction();
    InitNotchFilter();                   /* one coefficient set serves the audio and the FFT notch (audio_filters.c:43-45) */
}


void initAudioProcessor(void)            /* audio_processor.c:55-59 */
{
    InitAudioFilters();
    InitAGC();
}

void CWDecoder_Init(void) {}             /* cw_decoder.c:43-54: the Goertzel coefficient is derived inside the library */

/* ---- processRxAudio (audio_processor.c:275-434) ---- */
static void put16(uint8_t *p, float v)
{
    const int16_t w = (int16_t)v;        /* the rings hold the int16 bus words as floats (fpga.c:300-385) */
    p[0] = (uint8_t)((uint16_t)w >> 8);
    p[1] = (uint8_t)((uint16_t)w & 0xFF);
}

void processRxAudio(void)
{
    if (!Processor_NeedRXBuffer) return;
    ensure_contexts();
    AUDIOPROC_samples++;
    uint16_t idx = FPGA_Audio_Buffer_Index;                       /* :279-283 */
    if (idx == 0) idx = FPGA_AUDIO_BUFFER_SIZE; else idx--;

    static float sq[FPGA_AUDIO_BUFFER_HALF_SIZE], si[FPGA_AUDIO_BUFFER_HALF_SIZE];
    static float vq[FPGA_AUDIO_BUFFER_HALF_SIZE], vi[FPGA_AUDIO_BUFFER_HALF_SIZE];
    readHalfFromCircleBuffer32((uint32_t *)&FPGA_Audio_Buffer_SPEC_Q[0], (uint32_t *)sq, idx, FPGA_AUDIO_BUFFER_SIZE);
    readHalfFromCircleBuffer32((uint32_t *)&FPGA_Audio_Buffer_SPEC_I[0], (uint32_t *)si, idx, FPGA_AUDIO_BUFFER_SIZE);
    readHalfFromCircleBuffer32((uint32_t *)&FPGA_Audio_Buffer_VOICE_Q[0], (uint32_t *)vq, idx, FPGA_AUDIO_BUFFER_SIZE);
    readHalfFromCircleBuffer32((uint32_t *)&FPGA_Audio_Buffer_VOICE_I[0], (uint32_t *)vi, idx, FPGA_AUDIO_BUFFER_SIZE);
    static uint8_t frames[FPGA_AUDIO_BUFFER_HALF_SIZE * UA3_FRAME_BYTES];
    for (int i = 0; i < FPGA_AUDIO_BUFFER_HALF_SIZE; i++) {       /* bus order, stm32_interface.v:228-271 */
        put16(frames + 8 * i + 0, sq[i]); put16(frames + 8 * i + 2, si[i]);
        put16(frames + 8 * i + 4, vq[i]); put16(frames + 8 * i + 6, vi[i]);
    }

    apply_rx(false);
    if (Processor_RX_Audio_Samples_MAX_value != smeter_seen_max || Processor_RX_Audio_Samples_MIN_value != smeter_seen_min) {
        float drop[2];                                            /* the housekeeping tick zeroed the extremes (stm32f4xx_it.c:398-409) */
        CHECK(ua3reo_rx_read_smeter(rx_ctx, drop, 1));
    }
    CHECK(ua3reo_rx_push_frames(rx_ctx, frames, FPGA_AUDIO_BUFFER_HALF_SIZE));

    int32_t *out = (Processor_AudioBuffer_ReadyBuffer == 0) ? Processor_AudioBuffer_B : Processor_AudioBuffer_A;   /* :376-394 */
    CHECK(ua3reo_rx_read_audio(rx_ctx, out, 1));
    Processor_AudioBuffer_ReadyBuffer = (Processor_AudioBuffer_ReadyBuffer == 0) ? 1 : 0;
    if (WM8731_DMA_state) {                                       /* :396-413 codec DMA hand-off */
        memcpy(&CODEC_Audio_Buffer_RX[FPGA_AUDIO_BUFFER_SIZE], out, sizeof(int32_t) * FPGA_AUDIO_BUFFER_SIZE);
        AUDIOPROC_TXA_samples++;
    } else {
        memcpy(&CODEC_Audio_Buffer_RX[0], out, sizeof(int32_t) * FPGA_AUDIO_BUFFER_SIZE);
        AUDIOPROC_TXB_samples++;
    }
    if (USB_AUDIO_need_rx_buffer) {                               /* :415-432 */
        int16_t *usb = USB_AUDIO_current_rx_buffer ? USB_AUDIO_rx_buffer_b : USB_AUDIO_rx_buffer_a;
        CHECK(ua3reo_rx_read_audio_usb(rx_ctx, usb, 1));
        USB_AUDIO_need_rx_buffer = false;
    }
    float sm[2];
    CHECK(ua3reo_rx_read_smeter(rx_ctx, sm, 0));
    CHECK(ua3reo_rx_read_cw(rx_ctx, &ua3reo_shim_cw_magnitude, 1));
    CHECK(ua3reo_sync(rx_ctx));
    Processor_RX_Audio_Samples_MAX_value = smeter_seen_max = sm[0];
    Processor_RX_Audio_Samples_MIN_value = smeter_seen_min = sm[1];
    Processor_NeedRXBuffer = false;
}

/* ---- processTxAudio (audio_processor.c:61-273) ---- */
void processTxAudio(void)
{
    if (!Processor_NeedTXBuffer) return;
    ensure_contexts();
    AUDIOPROC_samples++;
    if (TRX.InputType == 2) {                                     /* USB audio, :67-72 */
        uint16_t buffer_index = USB_AUDIO_GetTXBufferIndex_FS() / 2;
        if ((buffer_index % 2) == 1) buffer_index--;
        readHalfFromCircleUSBBuffer((int16_t *)&USB_AUDIO_tx_buffer[0], (int32_t *)&Processor_AudioBuffer_A[0], buffer_index, USB_AUDIO_TX_BUFFER_SIZE / 2);
    } else {                                                      /* codec, :73-78 */
        uint16_t dma_index = __HAL_DMA_GET_COUNTER(hi2s3.hdmatx) / 2;
        if ((dma_index % 2) == 1) dma_index--;
        readHalfFromCircleBuffer32((uint32_t *)&CODEC_Audio_Buffer_TX[0], (uint32_t *)&Processor_AudioBuffer_A[0], dma_index, CODEC_AUDIO_BUFFER_SIZE);
    }
    static int16_t mic[FPGA_AUDIO_BUFFER_SIZE];
    for (int i = 0; i < FPGA_AUDIO_BUFFER_SIZE; i++) mic[i] = (int16_t)Processor_AudioBuffer_A[i];   /* :88-89 */
    apply_tx(false);
    CHECK(ua3reo_tx_process(rx_ctx, mic, 1));
    static float iq[FPGA_AUDIO_BUFFER_SIZE];
    CHECK(ua3reo_tx_read_iq(rx_ctx, NULL, iq, 1));
    CHECK(ua3reo_sync(rx_ctx));
    for (int i = 0; i < FPGA_AUDIO_BUFFER_HALF_SIZE; i++) {
        FPGA_Audio_Buffer_I_tmp[i] = iq[2 * i];
        FPGA_Audio_Buffer_Q_tmp[i] = iq[2 * i + 1];
    }
    if (TRX_getMode() == TRX_MODE_LOOPBACK && !TRX_tune) {        /* :228-249: to the codec instead of the FPGA */
        CHECK(ua3reo_tx_read_loopback(rx_ctx, Processor_AudioBuffer_A, 1));
        if (WM8731_DMA_state) {
            memcpy(&CODEC_Audio_Buffer_RX[FPGA_AUDIO_BUFFER_SIZE], Processor_AudioBuffer_A, sizeof(int32_t) * FPGA_AUDIO_BUFFER_SIZE);
            AUDIOPROC_TXA_samples++;
        } else {
            memcpy(&CODEC_Audio_Buffer_RX[0], Processor_AudioBuffer_A, sizeof(int32_t) * FPGA_AUDIO_BUFFER_SIZE);
            AUDIOPROC_TXB_samples++;
        }
    } else {
        const int half = FPGA_Audio_Buffer_State ? FPGA_AUDIO_BUFFER_HALF_SIZE : 0;                   /* :253-270 */
        for (int i = 0; i < FPGA_AUDIO_BUFFER_HALF_SIZE; i++) {
            FPGA_Audio_SendBuffer_I[half + i] = iq[2 * i];
            FPGA_Audio_SendBuffer_Q[half + i] = iq[2 * i + 1];
        }
        if (FPGA_Audio_Buffer_State) AUDIOPROC_TXA_samples++; else AUDIOPROC_TXB_samples++;
    }
    Processor_NeedTXBuffer = false;
    Processor_NeedRXBuffer = false;
}

/* ---- panorama: FFT_Init / FFT_doFFT numeric half (fft.c:185-328) ---- */
void FFT_Init(void)                      /* fft.c:185-210 */
{
    ensure_contexts();
    if (!fft_last_valid) apply_fft(true);                        /* first call: the whole settings block */
    const uint8_t zoom = TRX.FFT_Zoom ? TRX.FFT_Zoom : 1;
    CHECK(ua3reo_rx_fft_init(fft_ctx, 0, 1, &zoom));              /* decimator selected, biquad/FIR states cleared when zoom > 1 */
    fft_last.fft_zoom = zoom;
}

void FFT_doFFT(void)
{
    if (!TRX.FFT_Enabled) return;
    if (!FFT_need_fft) return;
    if (NeedFFTInputBuffer) return;
    ensure_contexts();
    apply_fft(false);
    if (CurrentVFO()->Freq != shim_fft_freq) {                    /* fft.c:347-351: applied on the device after this frame's averaging */
        const int32_t diff = (int32_t)(CurrentVFO()->Freq - shim_fft_freq);
        CHECK(ua3reo_rx_move_waterfall(fft_ctx, 0, 1, &diff));
        shim_fft_freq = CurrentVFO()->Freq;
    }
    static uint8_t frames[FFT_SIZE * UA3_FRAME_BYTES];
    memset(frames, 0, sizeof frames);
    for (int i = 0; i < FFT_SIZE; i++) {                          /* FFTInput_I/Q hold the SPEC words (fpga.c:307-339) */
        put16(frames + 8 * i + 0, FFTInput_Q[i]);
        put16(frames + 8 * i + 2, FFTInput_I[i]);
    }
    CHECK(ua3reo_rx_push_frames(fft_ctx, frames, FFT_SIZE));
    CHECK(ua3reo_rx_read_spectra(fft_ctx, ua3reo_shim_fft_mean, 1));
    CHECK(ua3reo_rx_read_waterfall(fft_ctx, ua3reo_shim_wtf_row0, 1));
    CHECK(ua3reo_sync(fft_ctx));
    NeedFFTInputBuffer = true;            /* fft.c:289: the bus driver may refill while the display pass runs */
    FFT_need_fft = false;
}

/* display half (fft.c:330-500) is outside the path: the column heights, colours and the maxValueErrors feedback
 * (fft.c:361-379) were computed on the device with the spectrum; drawing them is the firmware's LCD code. */
void FFT_printFFT(void)
{
    if (!TRX.FFT_Enabled) return;
    if (FFT_need_fft) return;
    CHECK(ua3reo_rx_read_waterfall_history(fft_ctx, ua3reo_shim_wtf_buffer));
    FFT_need_fft = true;
}
void FFT_printWaterfallDMA(void) {}
void FFT_moveWaterfall(int16_t freq_diff) { (void)freq_diff; }
