/*
 * fw_bus_hdl.c - the firmware's bus driver (fpga.c, unmodified) talking to the reference's own stm32_interface.v
 * (translated by tools/verilog_eval.py; glue in bus_hdl.c).  TEST INFRASTRUCTURE ONLY.
 *   fw_bus_hdl rx <iq_swap> <freq_hz> <in.bin> <out.bin>     in: per 48 kHz tick int16 SPEC_I, SPEC_Q, VOICE_I, VOICE_Q as the
 *                                                            filters present them; out: the dump of fw_fpga.c (same layout)
 *   fw_bus_hdl tx <freq_hz> <in.bin> <out.bin>               in: per tick float I, Q for FPGA_Audio_SendBuffer; out: per tick
 *                                                            int16 TX_I, TX_Q as the module holds them after the exchange,
 *                                                            then the 4 data bytes that crossed the bus
 *   fw_bus_hdl params <freq_hz> <mode> <preamp> <ptt> <adc.bin> <adc_otr> <dac_otr>
 *                                                            SEND PARAMS then (after the ADC samples) GET PARAMS; prints
 *                                                            key=value lines: what the FPGA latched and what the MCU decoded
 */
#include "stm32f4xx_hal.h"
#include "arm_math.h"
#include "settings.h"
#include "trx_manager.h"
#include "fpga.h"
#include "fft.h"
#include "functions.h"
#include "audio_processor.h"
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

void ua3_hdl_attach(void);
void ua3_hdl_flush(void);
void ua3_hdl_set_iq(int16_t spec_i, int16_t spec_q, int16_t voice_i, int16_t voice_q);
void ua3_hdl_set_flags(int adc_otr, int dac_otr);
void ua3_hdl_adc_clock(int16_t adc12);
int64_t ua3_hdl_get(const char *name);
extern unsigned long ua3_hdl_clk_edges;
extern uint8_t ua3_hdl_wr_log[64], ua3_hdl_rd_log[64];
extern unsigned ua3_hdl_n_wr, ua3_hdl_n_rd;
void ua3_hdl_log_reset(void);

static void trx_defaults(uint32_t freq, int mode)
{
    memset(&TRX, 0, sizeof TRX);
    TRX.VFO_A.Mode = (uint8_t)mode; TRX.VFO_A.Freq = freq; TRX.VFO_B = TRX.VFO_A; TRX.current_vfo = false;
}

static int run_rx(int argc, char **argv)
{
    if (argc < 6) return 2;
    trx_defaults((uint32_t)atol(argv[3]), TRX_MODE_USB);
    FILE *fi = fopen(argv[4], "rb"), *fo = fopen(argv[5], "wb");
    if (!fi || !fo) { perror("open"); return 2; }
    FPGA_Init();
    FPGA_NeedSendParams = true;
    FPGA_fpgadata_stuffclock();
    TRX_IQ_swap = atoi(argv[2]) != 0;
    NeedFFTInputBuffer = true;
    unsigned long n = 0;
    int16_t w[4];
    while (fread(w, sizeof(int16_t), 4, fi) == 4) {
        ua3_hdl_set_iq(w[0], w[1], w[2], w[3]);
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
    fprintf(stderr, "fw_bus_hdl rx: %lu ticks, %lu FPGA_CLK edges\n", n, ua3_hdl_clk_edges);
    return 0;
}

static int run_tx(int argc, char **argv)
{
    if (argc < 5) return 2;
    trx_defaults((uint32_t)atol(argv[2]), TRX_MODE_USB);
    FILE *fi = fopen(argv[3], "rb"), *fo = fopen(argv[4], "wb");
    if (!fi || !fo) { perror("open"); return 2; }
    FPGA_Init();
    TRX_ptt_hard = true;                            /* TRX_on_TX() (trx_manager.c:57-61) */
    FPGA_NeedSendParams = true;
    FPGA_fpgadata_stuffclock();
    float iq[2];
    unsigned long n = 0;
    while (fread(iq, sizeof(float), 2, fi) == 2) {
        /* processTxAudio() leaves the block in FPGA_Audio_SendBuffer_I/Q; the interrupt sends entry FPGA_Audio_Buffer_Index */
        FPGA_Audio_SendBuffer_I[FPGA_Audio_Buffer_Index] = iq[0];
        FPGA_Audio_SendBuffer_Q[FPGA_Audio_Buffer_Index] = iq[1];
        Processor_NeedTXBuffer = false;
        ua3_hdl_log_reset();
        FPGA_fpgadata_iqclock();
        const int16_t out[2] = {(int16_t)ua3_hdl_get("TX_I"), (int16_t)ua3_hdl_get("TX_Q")};
        fwrite(out, sizeof(int16_t), 2, fo);
        if (ua3_hdl_n_wr != 4) { fprintf(stderr, "fw_bus_hdl tx: %u data bytes on the bus, expected 4\n", ua3_hdl_n_wr); return 1; }
        fwrite(ua3_hdl_wr_log, 1, 4, fo);
        n++;
    }
    fclose(fi); fclose(fo);
    fprintf(stderr, "fw_bus_hdl tx: %lu ticks, tx=%ld rx=%ld\n", n, (long)ua3_hdl_get("tx"), (long)ua3_hdl_get("rx"));
    return 0;
}

static int run_params(int argc, char **argv)
{
    if (argc < 9) return 2;
    trx_defaults((uint32_t)atol(argv[2]), atoi(argv[3]));
    TRX.Preamp = atoi(argv[4]) != 0;
    TRX_ptt_hard = atoi(argv[5]) != 0;
    FPGA_Init();
    FPGA_NeedSendParams = true;
    FPGA_fpgadata_stuffclock();
    printf("freq_out=%ld\npreamp_enable=%ld\nrx=%ld\ntx=%ld\n", (long)ua3_hdl_get("freq_out"), (long)ua3_hdl_get("preamp_enable"),
           (long)ua3_hdl_get("rx"), (long)ua3_hdl_get("tx"));
    /* A first GET PARAMS raises ADC_MINMAX_RESET (stm32_interface.v:196-199); the ADC clock that follows reloads +2000 / -2000
     * and takes its own sample in (:384-397); the next command of any kind - here an RX I/Q tick, as on the board 21 us
     * later - drops the flag on its DATA_SYNC edge (:100).  Then the samples, then the read under test. */
    FPGA_NeedGetParams = true;
    FPGA_fpgadata_stuffclock();
    FILE *fa = fopen(argv[6], "rb");
    if (!fa) { perror("open"); return 2; }
    int16_t a;
    long n = 0;
    if (fread(&a, sizeof a, 1, fa) == 1) { ua3_hdl_adc_clock(a); n++; }
    FPGA_fpgadata_iqclock();
    while (fread(&a, sizeof a, 1, fa) == 1) { ua3_hdl_adc_clock(a); n++; }
    fclose(fa);
    ua3_hdl_set_flags(atoi(argv[7]), atoi(argv[8]));
    FPGA_NeedGetParams = true;
    ua3_hdl_log_reset();
    FPGA_fpgadata_stuffclock();
    printf("packet_bytes=%u\n", ua3_hdl_n_rd);
    for (unsigned i = 0; i < ua3_hdl_n_rd && i < 8; i++) printf("packet%u=%u\n", i, ua3_hdl_rd_log[i]);
    printf("adc_samples=%ld\nhdl_adc_min=%d\nhdl_adc_max=%d\n", n, (int)(int16_t)((int16_t)(ua3_hdl_get("ADC_MIN") << 4) >> 4),
           (int)(int16_t)((int16_t)(ua3_hdl_get("ADC_MAX") << 4) >> 4));
    printf("TRX_ADC_MINAMPLITUDE=%d\nTRX_ADC_MAXAMPLITUDE=%d\nTRX_ADC_OTR=%d\nTRX_DAC_OTR=%d\n", (int)TRX_ADC_MINAMPLITUDE,
           (int)TRX_ADC_MAXAMPLITUDE, (int)TRX_ADC_OTR, (int)TRX_DAC_OTR);
    return 0;
}

int main(int argc, char **argv)
{
    ua3_hdl_attach();
    int rc = 2;
    if (argc >= 2 && !strcmp(argv[1], "rx")) rc = run_rx(argc, argv);
    else if (argc >= 2 && !strcmp(argv[1], "tx")) rc = run_tx(argc, argv);
    else if (argc >= 2 && !strcmp(argv[1], "params")) rc = run_params(argc, argv);
    if (rc == 2) fprintf(stderr, "usage: fw_bus_hdl rx|tx|params ... (see the header of fw_bus_hdl.c)\n");
    ua3_hdl_flush();
    return rc;
}
