/*
 * fw_rx.c - drives the reference firmware's own receive-audio and panorama-FFT code (host-built,
 * unmodified sources from /root/reference/STM32/Src) over a stream of 8-byte FPGA I/Q frames.
 * TEST INFRASTRUCTURE ONLY (oracle/_ref/fw_rx).
 *
 *   fw_rx <params.txt> <frames.bin> <audio_out.bin> <fft_out.bin>
 *
 * Cadence conventions (the firmware's are set by ISR timing on real hardware):
 *   - every frame goes through FPGA_fpgadata_iqclock() -> FPGA_fpgadata_getiq() (fpga.c:148-171,286-401)
 *   - processRxAudio() runs once per 192 new frames, called when the ring index is 1 past a half
 *     boundary, which makes readHalfFromCircleBuffer32 (functions.c:21-34) return exactly the 192
 *     newest contiguous samples (audio_processor.c:279-298 reads up to index-1)
 *   - NeedFFTInputBuffer starts true (as TRX_setFrequency leaves it, trx_manager.c:185); each time the
 *     bus driver has filled 512 samples FFT_doFFT() runs, then FFT_printFFT() + the waterfall DMA chain,
 *     which is what feeds maxValueErrors back into the auto-range (fft.c:310-316,372) and re-arms FFT_need_fft
 *   - VFO frequency is held at 0 so that FFT_moveWaterfall() (fft.c:347-351) never shifts the averages
 * params.txt: `key value` lines applied before the init calls, and `at N key value` lines applied when N frames have
 *   been consumed (before that frame's processing) WITHOUT any re-initialisation - what a menu handler writing TRX
 *   does; the pseudo keys `reinit`, `notch_init`, `agc_init`, `fft_init` (and `smeter_reset`, which zeroes the S-meter extremes) call ReinitAudioFilters(), InitNotchFilter(), InitAGC()
 *   and FFT_Init() at that point, as TRX_setMode() / the 1 s tick / the zoom menu do (trx_manager.c:217, stm32f4xx_it.c:395).
 * audio_out: per block 384 int32 (L,R interleaved; audio_processor.c:377-394) + 3 float (S-meter max, min,
 *            CW decoder Goertzel magnitude of the block or 0) + 384 int16 (USB_AUDIO_rx_buffer_a)
 * fft_out  : per FFT frame 256 float (FFTOutput_mean) + 256 uint16 (waterfall row 0) + float maxValueFFT
 * optional 5th argument: file that receives wtf_buffer (50 x 256 uint16, row 0 newest) as it stands at the end
 * `freq` (as key or event) sets CurrentVFO()->Freq only - nothing retunes, FFT_printFFT() just sees the difference and
 *   moves the waterfall and the averages (fft.c:347-351,458-504)
 */
#include "stm32f4xx_hal.h"
#include "arm_math.h"
#include "settings.h"
#include "trx_manager.h"
#include "fpga.h"
#include "fft.h"
#include "audio_processor.h"
#include "audio_filters.h"
#include "agc.h"
#include "functions.h"
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

extern uint8_t ua3_bus_frame[8];
extern int ua3_bus_pos;
void ua3_lcd_stub_init(void);
const float *ua3_fft_output_mean(void);
const uint16_t *ua3_fft_wtf_row0(void);
const uint16_t *ua3_fft_wtf_all(void);
float ua3_fft_max_value(void);
float ua3_cw_magnitude(void);
void ua3_set_tick(uint32_t t);
#include "cw_decoder.h"
#include "usbd_audio_if.h"

static void defaults(void)
{
    memset(&TRX, 0, sizeof TRX);                    /* then the values of settings.c:33-94 that the path reads */
    TRX.clean_flash = 178;
    TRX.VFO_A.Freq = 0; TRX.VFO_A.Mode = TRX_MODE_LSB; TRX.VFO_A.Filter_Width = 2700;
    TRX.VFO_B = TRX.VFO_A; TRX.current_vfo = false;
    TRX.AGC = true; TRX.DNR = false; TRX.Agc_speed = 3; TRX.Volume = 20; TRX.Mute = false;
    TRX.CW_Filter = 500; TRX.SSB_Filter = 2700; TRX.FM_Filter = 15000; TRX.FM_SQL_threshold = 1; TRX.RF_Gain = 50;
    TRX.FFT_Zoom = 1; TRX.NotchFilter = false; TRX.NotchFC = 1000; TRX.CWDecoder = false; TRX.FFT_Enabled = true;
    TRX.FFT_Averaging = 4; TRX.SSB_HPF_pass = 300;
}

static int set_param(const char *k, long v)
{
#define P(name, field) if (!strcmp(k, name)) { field = v; return 1; }
    P("mode", TRX.VFO_A.Mode) P("filter_width", TRX.VFO_A.Filter_Width) P("hpf_pass", TRX.SSB_HPF_pass)
    P("rf_gain", TRX.RF_Gain) P("agc", TRX.AGC) P("agc_speed", TRX.Agc_speed) P("dnr", TRX.DNR)
    P("notch", TRX.NotchFilter) P("notch_fc", TRX.NotchFC) P("volume", TRX.Volume) P("mute", TRX.Mute)
    P("fm_sql", TRX.FM_SQL_threshold) P("fft_enabled", TRX.FFT_Enabled) P("fft_zoom", TRX.FFT_Zoom)
    P("fft_averaging", TRX.FFT_Averaging) P("freq", TRX.VFO_A.Freq) P("iq_swap", TRX_IQ_swap) P("squelched", TRX_squelched) P("cw_decoder", TRX.CWDecoder)
#undef P
    return 0;
}

struct event { unsigned long at; char key[32]; long val; };
static struct event events[64];
static int n_events;

static void fire(const struct event *e)
{
    if (!strcmp(e->key, "reinit")) ReinitAudioFilters();
    else if (!strcmp(e->key, "notch_init")) InitNotchFilter();
    else if (!strcmp(e->key, "fft_init")) FFT_Init();
    else if (!strcmp(e->key, "agc_init")) InitAGC();
    else if (!strcmp(e->key, "smeter_reset")) {       /* the housekeeping tick after it has shown the S-meter (stm32f4xx_it.c:398-409) */
        Processor_RX_Audio_Samples_MAX_value = 0;
        Processor_RX_Audio_Samples_MIN_value = 0;
    }
    else if (!set_param(e->key, e->val)) { fprintf(stderr, "unknown parameter %s\n", e->key); exit(2); }
}

int main(int argc, char **argv)
{
    if (argc < 5) { fprintf(stderr, "usage: fw_rx params.txt frames.bin audio_out.bin fft_out.bin\n"); return 2; }
    defaults();
    FILE *fp = fopen(argv[1], "r");
    if (!fp) { perror(argv[1]); return 2; }
    char key[64]; long val;
    while (fscanf(fp, "%63s", key) == 1) {
        if (!strcmp(key, "at")) {
            struct event *e = &events[n_events];
            if (n_events >= 64 || fscanf(fp, "%lu %31s %ld", &e->at, e->key, &e->val) != 3) { fprintf(stderr, "bad `at` line\n"); return 2; }
            n_events++;
            continue;
        }
        if (fscanf(fp, "%ld", &val) != 1 || !set_param(key, val)) { fprintf(stderr, "unknown parameter %s\n", key); return 2; }
    }
    fclose(fp);
    FILE *fi = fopen(argv[2], "rb"), *fa = fopen(argv[3], "wb"), *ff = fopen(argv[4], "wb");
    if (!fi || !fa || !ff) { perror("open"); return 2; }

    ua3_lcd_stub_init();
    /* init order of main.c:188-193 (the calls that touch the signal path) */
    FFT_Init();
    CWDecoder_Init();                  /* TRX_Init() (trx_manager.c:63-68) */
    initAudioProcessor();              /* InitAudioFilters (incl. InitNoiseReduction, InitNotchFilter) + InitAGC */
    ReinitAudioFilters();              /* as TRX_setMode() does, trx_manager.c:217 */
    NeedFFTInputBuffer = true;         /* trx_manager.c:185 */
    FFT_need_fft = true;

    uint8_t frame[8];
    unsigned long n = 0;
    while (fread(frame, 1, 8, fi) == 8) {
        for (int e = 0; e < n_events; e++) if (events[e].at == n) fire(&events[e]);
        memcpy(ua3_bus_frame, frame, 8);
        ua3_bus_pos = 0;
        FPGA_fpgadata_iqclock();
        n++;
        if (n > 192 && (n % 192) == 1) {
            Processor_NeedRXBuffer = true;
            USB_AUDIO_need_rx_buffer = true;   /* exercise the USB-audio int16 packing (audio_processor.c:415-432) */
            USB_AUDIO_current_rx_buffer = false;
            /* HAL_GetTick() stays 0: the Morse timing logic (cw_decoder.c:88-240, host side, unbounded strcat) never fires;
             * only the Goertzel front end (:56-66) is exercised */
            const uint8_t before = Processor_AudioBuffer_ReadyBuffer;
            processRxAudio();
            const int32_t *out = (before == 0) ? Processor_AudioBuffer_B : Processor_AudioBuffer_A;
            fwrite(out, sizeof(int32_t), FPGA_AUDIO_BUFFER_SIZE, fa);
            float sm[3] = {Processor_RX_Audio_Samples_MAX_value, Processor_RX_Audio_Samples_MIN_value, 0.0f};
            if (TRX.CWDecoder && (TRX_getMode() == TRX_MODE_CW_L || TRX_getMode() == TRX_MODE_CW_U)) sm[2] = ua3_cw_magnitude();
            fwrite(sm, sizeof(float), 3, fa);
            fwrite(USB_AUDIO_rx_buffer_a, sizeof(int16_t), FPGA_AUDIO_BUFFER_SIZE, fa);
        }
        if (!NeedFFTInputBuffer && TRX.FFT_Enabled) {
            FFT_doFFT();
            fwrite(ua3_fft_output_mean(), sizeof(float), FFT_PRINT_SIZE, ff);
            FFT_printFFT();
            for (int guard = 0; !FFT_need_fft && guard < 1000; guard++) FFT_printWaterfallDMA();
            fwrite(ua3_fft_wtf_row0(), sizeof(uint16_t), FFT_PRINT_SIZE, ff);
            float mv = ua3_fft_max_value();
            fwrite(&mv, sizeof(float), 1, ff);
        }
    }
    fclose(fi); fclose(fa); fclose(ff);
    if (argc > 5) {
        FILE *fw = fopen(argv[5], "wb");
        if (!fw) { perror(argv[5]); return 2; }
        fwrite(ua3_fft_wtf_all(), sizeof(uint16_t), FFT_WTF_HEIGHT * FFT_PRINT_SIZE, fw);
        fclose(fw);
    }
    return 0;
}
