/*
 * fw_stage.c - drives the firmware's stand-alone stage functions dc_filter(), DoAGC() and processNoiseReduction()
 * (audio_filters.h:51, agc.h:9, noise_reduction.h:16) over a float stream.  TEST INFRASTRUCTURE ONLY.  Built twice: with the
 * reference's own audio_filters.c / agc.c / noise_reduction.c (oracle/_ref/fw_stage) and with the GPU shim
 * (oracle/_ref/fw_stage_b200).   fw_stage <agc 0|1> <agc_speed> <dnr 0|1> <mode> <in.f32> <out.f32>
 * Per 192-sample block: dc_filter(block, 192, 0); three processNoiseReduction(sub, sub) calls of 64; DoAGC(block, 192);
 * a second, independent stream goes through dc_filter(.., 5) only (state slots are separate).
 */
#include "stm32f4xx_hal.h"
#include "arm_math.h"
#include "settings.h"
#include "trx_manager.h"
#include "audio_filters.h"
#include "agc.h"
#include "noise_reduction.h"
#include "audio_processor.h"
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

void ua3_lcd_stub_init(void);

int main(int argc, char **argv)
{
    if (argc < 7) { fprintf(stderr, "usage: fw_stage agc agc_speed dnr mode in.f32 out.f32\n"); return 2; }
    memset(&TRX, 0, sizeof TRX);
    TRX.VFO_A.Mode = (uint8_t)atoi(argv[4]); TRX.VFO_A.Filter_Width = 2700; TRX.VFO_B = TRX.VFO_A; TRX.current_vfo = false;
    TRX.AGC = atoi(argv[1]) != 0; TRX.Agc_speed = (uint8_t)atoi(argv[2]); TRX.DNR = atoi(argv[3]) != 0;
    TRX.Volume = 20; TRX.RF_Gain = 50; TRX.SSB_HPF_pass = 300; TRX.NotchFC = 1000; TRX.FFT_Averaging = 4; TRX.FFT_Zoom = 1;
    FILE *fi = fopen(argv[5], "rb"), *fo = fopen(argv[6], "wb");
    if (!fi || !fo) { perror("open"); return 2; }
    ua3_lcd_stub_init();
    initAudioProcessor();
    ReinitAudioFilters();
    float blk[192], side[192];
    while (fread(blk, sizeof(float), 192, fi) == 192) {
        memcpy(side, blk, sizeof side);
        dc_filter(blk, 192, 0);
        for (int sb = 0; sb < 3; sb++) processNoiseReduction(blk + 64 * sb, blk + 64 * sb);
        DoAGC(blk, 192);
        dc_filter(side, 192, 5);
        fwrite(blk, sizeof(float), 192, fo);
        fwrite(side, sizeof(float), 192, fo);
    }
    fclose(fi); fclose(fo);
    return 0;
}
