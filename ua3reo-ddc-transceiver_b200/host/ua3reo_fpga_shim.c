/*
 * ua3reo_fpga_shim.c - the FPGA half of the firmware-side binding: the reference's bus driver entry points
 *     void FPGA_Init(void); void FPGA_fpgadata_iqclock(void); void FPGA_fpgadata_stuffclock(void);        (fpga.h:12-14)
 * implemented over libua3reo_b200.so, linked INSTEAD OF fpga.c.  Where the firmware clocks bytes over an 8-bit bus out
 * of the FPGA fabric, this file calls the CUDA receive chain: ADC samples go in (ua3reo_fpga_shim_adc, the stand-in for
 * the ADC pins), ua3reo_ddc_push runs NCO / mixer / CIC / compensator / Hilbert on the GPU, and every
 * FPGA_fpgadata_iqclock() takes the next 8-byte frame and stores it exactly as FPGA_fpgadata_getiq() does
 * (fpga.c:286-401): int16 words as float32 into the four 384-deep rings, SPEC words into FFTInput_I/Q while
 * NeedFFTInputBuffer is set, I and Q destinations exchanged when TRX_IQ_swap, ring index modulo 384, FFT_buff_index up
 * to 512.  On transmit (fpga.c:403-466) the I/Q words leave through ua3reo_duc_push_wire in the byte order of command 3.
 * FPGA_fpgadata_stuffclock() (fpga.c:120-146): command 1 latches the tuning word (ua3reo_set_fcw) and the RX/TX bit -
 * raising RX releases reset = RX_N of the whole receive chain (UA3REO.bdf), i.e. ua3reo_reset; command 2 decodes the
 * GET PARAMS packet of ua3reo_get_params the way FPGA_fpgadata_getparam() does (fpga.c:222-284).
 * Everything the rest of the firmware reads (fpga.h:16-31) is defined here under the firmware's names.
 * There is no CPU fallback: without the library or a GPU the shim prints the library's error and aborts.
 *
 * Built by oracle/ref_harness/Makefile into oracle/_ref/fw_fpga_b200; tests/test_fw_shim_gpu.py runs it beside
 * oracle/_ref/fw_fpga, the same driver linked with the reference's own fpga.c fed from the golden DDC over the bus stub.
 */
#include "stm32f4xx_hal.h"
#include "main.h"
#include "fpga.h"
#include "functions.h"
#include "trx_manager.h"
#include "audio_processor.h"
#include "settings.h"
#include "ua3reo_b200.h"
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* ---- the globals fpga.c defines (fpga.c:12-27) ---- */
volatile uint32_t FPGA_samples = 0;
volatile bool FPGA_busy = false;
volatile bool FPGA_NeedSendParams = false;
volatile bool FPGA_NeedGetParams = false;
volatile bool FPGA_Buffer_underrun = false;
uint16_t FPGA_Audio_Buffer_Index = 0;
bool FPGA_Audio_Buffer_State = true;
float32_t FPGA_Audio_Buffer_SPEC_Q[FPGA_AUDIO_BUFFER_SIZE] = {0};
float32_t FPGA_Audio_Buffer_SPEC_I[FPGA_AUDIO_BUFFER_SIZE] = {0};
float32_t FPGA_Audio_Buffer_VOICE_Q[FPGA_AUDIO_BUFFER_SIZE] = {0};
float32_t FPGA_Audio_Buffer_VOICE_I[FPGA_AUDIO_BUFFER_SIZE] = {0};
float32_t FPGA_Audio_SendBuffer_Q[FPGA_AUDIO_BUFFER_SIZE] = {0};
float32_t FPGA_Audio_SendBuffer_I[FPGA_AUDIO_BUFFER_SIZE] = {0};

#define SHIM_BLOCK (1u << 16)                  /* ADC samples per push: 64 frames */
#define SHIM_FIFO 4096                         /* frames waiting for iqclock */
#define SHIM_TX_BATCH 64                       /* TX words per DUC push */

static ua3reo_ctx *ctx;
static uint8_t fifo[SHIM_FIFO][UA3_FRAME_BYTES];
static unsigned fifo_head, fifo_count;
static uint8_t tx_wire[SHIM_TX_BATCH][4];
static unsigned tx_count;
static bool fpga_rx = false;                   /* the FPGA's rx register (stm32_interface.v:150) */

static void die(const char *what, int rc)
{
    fprintf(stderr, "ua3reo_fpga_shim: %s failed (%d): %s\n", what, rc, ua3reo_last_error());
    abort();
}
#define CHECK(call) do { int rc_ = (call); if (rc_) die(#call, rc_); } while (0)

static void ensure(void)
{
    if (ctx) return;
    CHECK(ua3reo_create(0, 1, SHIM_BLOCK, &ctx));
    CHECK(ua3reo_duc_enable(ctx, SHIM_TX_BATCH));
}

ua3reo_ctx *ua3reo_fpga_shim_context(void) { ensure(); return ctx; }

/* The ADC pins: n samples (12-bit two's complement in int16) that arrived since the last call.  Frames the chain
 * completes are queued for FPGA_fpgadata_iqclock().  Returns the number of frames now waiting. */
unsigned ua3reo_fpga_shim_adc(const int16_t *adc, size_t n)
{
    static uint8_t frames[(SHIM_BLOCK / UA3_ADC_PER_FRAME + 1) * UA3_FRAME_BYTES];
    ensure();
    while (n) {
        const size_t take = n < SHIM_BLOCK - UA3_ADC_PER_FRAME ? n : SHIM_BLOCK - UA3_ADC_PER_FRAME;
        size_t nf = 0;
        CHECK(ua3reo_ddc_push(ctx, adc, take, &nf));
        CHECK(ua3reo_ddc_read_frames(ctx, frames, nf));
        for (size_t f = 0; f < nf; f++) {
            if (fifo_count == SHIM_FIFO) { fifo_head = (fifo_head + 1) % SHIM_FIFO; fifo_count--; }   /* nobody clocked them out: oldest lost */
            memcpy(fifo[(fifo_head + fifo_count) % SHIM_FIFO], frames + f * UA3_FRAME_BYTES, UA3_FRAME_BYTES);
            fifo_count++;
        }
        adc += take; n -= take;
    }
    return fifo_count;
}

/* DAC words of the TX samples clocked in so far (flushes the batch): dst receives up to max words, returns the count */
size_t ua3reo_fpga_shim_dac(uint16_t *dst, size_t max_words)
{
    ensure();
    if (!tx_count) return 0;
    const size_t n = tx_count;
    if (n * 1024 > max_words) return 0;
    CHECK(ua3reo_duc_push_wire(ctx, &tx_wire[0][0], n));
    CHECK(ua3reo_duc_read_dac(ctx, dst, n));
    tx_count = 0;
    return n * 1024;
}

void FPGA_start_audio_clock(void) {}           /* command 5 (fpga.c:92-104): the 48 kHz clock is the caller's tick here */
void FPGA_stop_audio_clock(void) {}

void FPGA_Init(void)                           /* fpga.c:41-55: GPIO setup, bus test, audio clock on */
{
    ensure();
    FPGA_start_audio_clock();
}

static inline int16_t word(const uint8_t *p) { return (int16_t)(uint16_t)(((uint16_t)p[0] << 8) | p[1]); }

static void getiq(void)                        /* FPGA_fpgadata_getiq, fpga.c:286-401 */
{
    static const uint8_t silence[UA3_FRAME_BYTES] = {0};
    const uint8_t *f = silence;
    FPGA_samples++;
    if (fifo_count) { f = fifo[fifo_head]; fifo_head = (fifo_head + 1) % SHIM_FIFO; fifo_count--; }
    else FPGA_Buffer_underrun = true;          /* the fabric always answers; here nothing has been computed yet */
    const int16_t spec_q = word(f + 0), spec_i = word(f + 2), voice_q = word(f + 4), voice_i = word(f + 6);
    const uint16_t i = FPGA_Audio_Buffer_Index;
    if (TRX_IQ_swap) {
        if (NeedFFTInputBuffer) { FFTInput_I[FFT_buff_index] = spec_q; FFTInput_Q[FFT_buff_index] = spec_i; }
        FPGA_Audio_Buffer_SPEC_I[i] = spec_q; FPGA_Audio_Buffer_SPEC_Q[i] = spec_i;
        FPGA_Audio_Buffer_VOICE_I[i] = voice_q; FPGA_Audio_Buffer_VOICE_Q[i] = voice_i;
    } else {
        if (NeedFFTInputBuffer) { FFTInput_Q[FFT_buff_index] = spec_q; FFTInput_I[FFT_buff_index] = spec_i; }
        FPGA_Audio_Buffer_SPEC_Q[i] = spec_q; FPGA_Audio_Buffer_SPEC_I[i] = spec_i;
        FPGA_Audio_Buffer_VOICE_Q[i] = voice_q; FPGA_Audio_Buffer_VOICE_I[i] = voice_i;
    }
    FPGA_Audio_Buffer_Index++;
    if (FPGA_Audio_Buffer_Index == FPGA_AUDIO_BUFFER_SIZE) FPGA_Audio_Buffer_Index = 0;
    if (NeedFFTInputBuffer) {
        FFT_buff_index++;
        if (FFT_buff_index == FFT_SIZE) { FFT_buff_index = 0; NeedFFTInputBuffer = false; }
    }
}

static void sendiq(void)                       /* FPGA_fpgadata_sendiq, fpga.c:403-466 */
{
    FPGA_samples++;
    const int16_t q = (int16_t)(float32_t)FPGA_Audio_SendBuffer_Q[FPGA_Audio_Buffer_Index];
    const int16_t i = (int16_t)(float32_t)FPGA_Audio_SendBuffer_I[FPGA_Audio_Buffer_Index];
    if (tx_count == SHIM_TX_BATCH) {           /* nobody collected the DAC words: run the batch and drop them */
        CHECK(ua3reo_duc_push_wire(ctx, &tx_wire[0][0], tx_count));
        tx_count = 0;
    }
    tx_wire[tx_count][0] = (uint8_t)((uint16_t)q >> 8); tx_wire[tx_count][1] = (uint8_t)q;       /* Q hi, Q lo, I hi, I lo */
    tx_wire[tx_count][2] = (uint8_t)((uint16_t)i >> 8); tx_wire[tx_count][3] = (uint8_t)i;
    tx_count++;
    FPGA_Audio_Buffer_Index++;
    if (FPGA_Audio_Buffer_Index == FPGA_AUDIO_BUFFER_SIZE) {
        if (Processor_NeedTXBuffer) { FPGA_Buffer_underrun = true; FPGA_Audio_Buffer_Index--; }
        else { FPGA_Audio_Buffer_Index = 0; FPGA_Audio_Buffer_State = true; Processor_NeedTXBuffer = true; }
    } else if (FPGA_Audio_Buffer_Index == FPGA_AUDIO_BUFFER_SIZE / 2) {
        if (Processor_NeedTXBuffer) { FPGA_Buffer_underrun = true; FPGA_Audio_Buffer_Index--; }
        else { FPGA_Audio_Buffer_State = false; Processor_NeedTXBuffer = true; }
    }
}

void FPGA_fpgadata_iqclock(void)               /* fpga.c:148-171 */
{
    ensure();
    FPGA_busy = true;
    if (TRX_on_TX() && TRX_getMode() != TRX_MODE_LOOPBACK) sendiq();
    else getiq();
    FPGA_busy = false;
}

void FPGA_fpgadata_stuffclock(void)            /* fpga.c:120-146 with sendparam :173-220 and getparam :222-284 */
{
    ensure();
    FPGA_busy = true;
    if (FPGA_NeedSendParams) {
        uint32_t phrase = getPhraseFromFrequency(CurrentVFO()->Freq);
        if (!TRX_on_TX()) {
            if (TRX_getMode() == TRX_MODE_CW_L) phrase = getPhraseFromFrequency(CurrentVFO()->Freq + TRX.CW_GENERATOR_SHIFT_HZ);
            else if (TRX_getMode() == TRX_MODE_CW_U) phrase = getPhraseFromFrequency(CurrentVFO()->Freq - TRX.CW_GENERATOR_SHIFT_HZ);
        }
        const bool tx_bit = TRX_on_TX() && TRX_getMode() != TRX_MODE_LOOPBACK;    /* byte 0 bit 3; rx = !tx (stm32_interface.v:148-152) */
        const uint32_t fcw = phrase & 0x3FFFFFu;                                  /* three bytes, 22 bits used (:159-169) */
        CHECK(ua3reo_set_fcw(ctx, 0, 1, &fcw));
        if (!tx_bit && !fpga_rx) { CHECK(ua3reo_reset(ctx)); fifo_count = 0; }    /* RX raised: reset = RX_N released */
        fpga_rx = !tx_bit;
        FPGA_NeedSendParams = false;
    } else if (FPGA_NeedGetParams) {
        uint8_t p[5];
        int16_t mn = 0, mx = 0;
        CHECK(ua3reo_get_params(ctx, p, &mn, &mx, 0));
        TRX_ADC_OTR = p[0] & 1; TRX_DAC_OTR = (p[0] >> 1) & 1;
        TRX_FrontPanel.key_4 = (p[0] >> 2) & 1; TRX_FrontPanel.key_3 = (p[0] >> 3) & 1;
        TRX_FrontPanel.key_2 = (p[0] >> 4) & 1; TRX_FrontPanel.key_1 = (p[0] >> 5) & 1;
        TRX_ADC_MINAMPLITUDE = mn;             /* decoded inside the library as fpga.c:256-271 does, quirks included */
        TRX_ADC_MAXAMPLITUDE = mx;
        int8_t enc = (int8_t)(p[4] & 0xF);
        if (enc > 7) enc |= (int8_t)0xF0;
        TRX_FrontPanel.sec_encoder = enc;
        TRX_FrontPanel.key_enc = (p[4] >> 4) & 1;
        FPGA_NeedGetParams = false;
    }
    FPGA_busy = false;
}
