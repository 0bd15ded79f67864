/*
 * fw_tx.c - drives the reference firmware's own transmit-audio code, processTxAudio()
 * (audio_processor.c:61-273, host-built unmodified), over a stream of microphone/line samples.
 * TEST INFRASTRUCTURE ONLY (oracle/_ref/fw_tx).
 *
 *   fw_tx <params.txt> <mic.bin> <iq_out.bin>
 *
 * mic.bin: int16 pairs (left, right) at 48 kHz, 192 pairs per block, as the codec DMA leaves them in
 * CODEC_Audio_Buffer_TX (one int32 per channel sample, wm8731.c).  Per block the harness places the 192 pairs
 * where readHalfFromCircleBuffer32() will read them for a DMA counter of 0 (the stub's value), raises
 * Processor_NeedTXBuffer and calls processTxAudio(); FPGA_Audio_Buffer_State alternates as the bus driver
 * does (fpga.c:440-465).  iq_out: per block 192 x (float I, float Q) as left in FPGA_Audio_SendBuffer_I/Q,
 * followed by 192 x (int16 I, int16 Q) as FPGA_fpgadata_sendiq() converts them for the wire (fpga.c:409,424), followed by
 * 384 int32: the half of CODEC_Audio_Buffer_RX that TRX_MODE_LOOPBACK fills instead (audio_processor.c:228-249; zeros in the
 * other modes).  In loopback mode nothing is sent to the FPGA, the float I/Q are then FPGA_Audio_Buffer_I/Q_tmp.
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
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

void ua3_lcd_stub_init(void);

static int set_param(const char *k, long v)
{
#define P(name, field) if (!strcmp(k, name)) { field = v; return 1; }
    P("mode", TRX.VFO_A.Mode) P("filter_width", TRX.VFO_A.Filter_Width) P("hpf_pass", TRX.SSB_HPF_pass)
    P("rf_power", TRX.RF_Power) P("mute", TRX.Mute) P("tune", TRX_tune) P("volume", TRX.Volume)
    P("key", TRX_key_serial) P("input_type", TRX.InputType)
#undef P
    return 0;
}

int main(int argc, char **argv)
{
    if (argc < 4) { fprintf(stderr, "usage: fw_tx params.txt mic.bin iq_out.bin\n"); return 2; }
    memset(&TRX, 0, sizeof TRX);
    TRX.VFO_A.Mode = TRX_MODE_USB; TRX.VFO_A.Filter_Width = 2700; TRX.current_vfo = false;
    TRX.RF_Power = 20; TRX.Volume = 20; TRX.SSB_HPF_pass = 300; TRX.InputType = 0; TRX.FFT_Enabled = false;
    FILE *fp = fopen(argv[1], "r");
    if (!fp) { perror(argv[1]); return 2; }
    char key[64]; long val;
    while (fscanf(fp, "%63s %ld", key, &val) == 2)
        if (!set_param(key, val)) { fprintf(stderr, "unknown parameter %s\n", key); return 2; }
    fclose(fp);
    FILE *fi = fopen(argv[2], "rb"), *fo = fopen(argv[3], "wb");
    if (!fi || !fo) { perror("open"); return 2; }
    ua3_lcd_stub_init();
    initAudioProcessor();
    ReinitAudioFilters();
    NeedFFTInputBuffer = false;

    int16_t mic[192 * 2];
    FPGA_Audio_Buffer_State = true;
    while (fread(mic, sizeof(int16_t), 192 * 2, fi) == 192 * 2) {
        for (int i = 0; i < 192 * 2; i++) CODEC_Audio_Buffer_TX[FPGA_AUDIO_BUFFER_SIZE + i] = mic[i];
        Processor_NeedTXBuffer = true;
        processTxAudio();
        const int half = FPGA_Audio_Buffer_State ? FPGA_AUDIO_BUFFER_HALF_SIZE : 0;
        float f[192 * 2];
        int16_t w[192 * 2];
        const int loopback = (TRX_getMode() == TRX_MODE_LOOPBACK) && !TRX_tune;
        int32_t codec[192 * 2];
        memset(codec, 0, sizeof codec);
        for (int i = 0; i < 192; i++) {
            const float fi_ = loopback ? FPGA_Audio_Buffer_I_tmp[i] : FPGA_Audio_SendBuffer_I[half + i];
            const float fq_ = loopback ? FPGA_Audio_Buffer_Q_tmp[i] : FPGA_Audio_SendBuffer_Q[half + i];
            f[2 * i] = fi_;
            f[2 * i + 1] = fq_;
            w[2 * i] = (int16_t)(float32_t)fi_;      /* fpga.c:424 */
            w[2 * i + 1] = (int16_t)(float32_t)fq_;  /* fpga.c:409 */
        }
        if (loopback) memcpy(codec, &CODEC_Audio_Buffer_RX[WM8731_DMA_state ? FPGA_AUDIO_BUFFER_SIZE : 0], sizeof codec);
        fwrite(f, sizeof(float), 192 * 2, fo);
        fwrite(w, sizeof(int16_t), 192 * 2, fo);
        fwrite(codec, sizeof(int32_t), 192 * 2, fo);
        FPGA_Audio_Buffer_State = !FPGA_Audio_Buffer_State;
    }
    fclose(fi); fclose(fo);
    return 0;
}
