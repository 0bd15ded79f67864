/*
 * fw_fpga.c - drives the firmware's bus-driver entry points (fpga.h:12-14) for n 48 kHz ticks and dumps everything they
 * leave behind.  TEST INFRASTRUCTURE ONLY.  Two builds of this one driver:
 *   oracle/_ref/fw_fpga       linked with the reference's own fpga.c; input = the golden DDC's 8-byte frames, served byte by
 *                             byte through the GPIO bus stub (fw_stubs.c), as stm32_interface.v:228-271 would
 *   oracle/_ref/fw_fpga_b200  linked with ua3reo-ddc-transceiver_b200/host/ua3reo_fpga_shim.c; input = the ADC samples, the
 *                             frames come from the CUDA receive chain
 *   fw_fpga[_b200] <iq_swap> <freq_hz> <input.bin> <out.bin>
 * out: for every tick index, FFT_buff_index, NeedFFTInputBuffer (3 x uint32); at the end the four rings, FFTInput_I/Q.
 * FFT input is re-armed every 700 ticks, as FFT_printFFT() does once the display pass is through.
 */
#include "stm32f4xx_hal.h"
#include "arm_math.h"
#include "settings.h"
#include "trx_manager.h"
#include "fpga.h"
#include "fft.h"
#include "functions.h"
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#ifdef UA3_FPGA_SHIM
unsigned ua3reo_fpga_shim_adc(const int16_t *adc, size_t n);
#else
extern uint8_t ua3_bus_frame[8];
extern int ua3_bus_pos;
#endif

int main(int argc, char **argv)
{
    if (argc < 5) { fprintf(stderr, "usage: fw_fpga iq_swap freq_hz input.bin out.bin\n"); return 2; }
    memset(&TRX, 0, sizeof TRX);
    TRX.VFO_A.Mode = TRX_MODE_USB; TRX.VFO_A.Freq = (uint32_t)atol(argv[2]); TRX.VFO_B = TRX.VFO_A; TRX.current_vfo = false;
    FILE *fi = fopen(argv[3], "rb"), *fo = fopen(argv[4], "wb");
    if (!fi || !fo) { perror("open"); return 2; }
    FPGA_Init();
    FPGA_NeedSendParams = true;                     /* TRX_setFrequency -> FPGA_NeedSendParams (trx_manager.c:183-186) */
    FPGA_fpgadata_stuffclock();
    TRX_IQ_swap = atoi(argv[1]) != 0;               /* getPhraseFromFrequency() sets it; the harness lets the test force it */
    NeedFFTInputBuffer = true;
    unsigned long n = 0;
#ifdef UA3_FPGA_SHIM
    static int16_t adc[1024];
    while (fread(adc, sizeof(int16_t), 1024, fi) == 1024) {
        ua3reo_fpga_shim_adc(adc, 1024);            /* one 48 kHz period of ADC samples, then the tick */
#else
    uint8_t frame[8];
    while (fread(frame, 1, 8, fi) == 8) {
        memcpy(ua3_bus_frame, frame, 8);
        ua3_bus_pos = 0;
#endif
        FPGA_fpgadata_iqclock();
        n++;
        if (n % 700 == 0) NeedFFTInputBuffer = true;
        const uint32_t rec[3] = {FPGA_Audio_Buffer_Index, FFT_buff_index, NeedFFTInputBuffer};
        fwrite(rec, sizeof(uint32_t), 3, fo);
    }
    fwrite(FPGA_Audio_Buffer_SPEC_Q, sizeof(float), FPGA_AUDIO_BUFFER_SIZE, fo);
    fwrite(FPGA_Audio_Buffer_SPEC_I, sizeof(float), FPGA_AUDIO_BUFFER_SIZE, fo);
    fwrite(FPGA_Audio_Buffer_VOICE_Q, sizeof(float), FPGA_AUDIO_BUFFER_SIZE, fo);
    fwrite(FPGA_Audio_Buffer_VOICE_I, sizeof(float), FPGA_AUDIO_BUFFER_SIZE, fo);
    fwrite(FFTInput_I, sizeof(float), FFT_SIZE, fo);
    fwrite(FFTInput_Q, sizeof(float), FFT_SIZE, fo);
    const uint32_t tail[2] = {(uint32_t)FPGA_samples, (uint32_t)FPGA_Buffer_underrun};
    fwrite(tail, sizeof(uint32_t), 2, fo);
    fclose(fi); fclose(fo);
    return 0;
}
