/*
 * ua3reo_rx_host.c - C host driver over the C ABI: the equivalent of the firmware's main loop + ISR cadence
 * (main.c:188-215, stm32f4xx_it.c:315-367) for a bank of receivers on one B200.
 *
 *   ua3reo_rx_host [n_channels] [n_blocks] [block_samples] [full_chain 0|1]
 *
 * Generates a synthetic 12-bit ADC stream (tones + noise), tunes every channel to its own frequency,
 * pushes the stream block by block through the FPGA chain (and, with full_chain, the STM32 audio/FFT
 * stage), pulls the results back and prints throughput plus checksums of everything it received.
 */
#include "ua3reo_b200.h"
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

static uint32_t lcg(uint32_t *s) { *s = *s * 1664525u + 1013904223u; return *s; }

static double now(void)
{
    struct timespec t;
    clock_gettime(CLOCK_MONOTONIC, &t);
    return (double)t.tv_sec + 1e-9 * (double)t.tv_nsec;
}

#define CHECK(call) do { if ((call) != UA3_OK) { fprintf(stderr, "%s failed: %s\n", #call, ua3reo_last_error()); return 1; } } while (0)

int main(int argc, char **argv)
{
    const uint32_t n_ch = argc > 1 ? (uint32_t)atoi(argv[1]) : 64;
    const int n_blocks = argc > 2 ? atoi(argv[2]) : 8;
    const uint32_t block = argc > 3 ? (uint32_t)atoi(argv[3]) : (1u << 18);
    const int full = argc > 4 ? atoi(argv[4]) : 1;
    static const uint8_t modes[5] = {0 /*LSB*/, 1 /*USB*/, 4 /*CW_U*/, 10 /*AM*/, 8 /*NFM*/};
    static const uint16_t widths[5] = {2700, 2700, 500, 6000, 15000};

    ua3reo_ctx *bank = NULL;
    CHECK(ua3reo_create(0, n_ch, block, &bank));
    printf("%s: %u channels, %d blocks of %u ADC samples, %s\n", ua3reo_version(), n_ch, n_blocks, block,
           full ? "full RX chain" : "DDC only");

    uint32_t seed = 20261018u;
    for (uint32_t c = 0; c < n_ch; c++) {
        const uint32_t f_hz = 1800000u + lcg(&seed) % 28000000u;         /* 1.8 .. 29.8 MHz */
        CHECK(ua3reo_set_frequency(bank, c, f_hz));
    }
    if (full) {
        CHECK(ua3reo_rx_enable(bank, 1));
        for (uint32_t c = 0; c < n_ch; c++) {
            ua3reo_rx_settings s;
            ua3reo_rx_defaults(&s);
            s.mode = modes[c % 5];
            s.filter_width = widths[c % 5];
            s.dnr = (uint8_t)((c / 5) & 1);
            s.notch = (uint8_t)((c / 5) & 1);
            CHECK(ua3reo_rx_set(bank, c, 1, &s));
        }
    }

    int16_t *adc = (int16_t *)malloc(sizeof(int16_t) * block);
    uint8_t *frames = (uint8_t *)malloc((size_t)n_ch * (block / 1024) * 8);
    int32_t *audio = (int32_t *)malloc((size_t)n_ch * (block / 1024 / 192 + 2) * 384 * sizeof(int32_t));
    float *spectra = (float *)malloc((size_t)n_ch * (block / 1024 / 512 + 2) * 256 * sizeof(float));
    if (!adc || !frames || !audio || !spectra) { fprintf(stderr, "out of memory\n"); return 1; }

    uint64_t sum_frames = 0, sum_audio = 0;
    double sum_spec = 0.0, t_total = 0.0;
    uint64_t t_index = 0;
    for (int b = 0; b < n_blocks; b++) {
        for (uint32_t i = 0; i < block; i++, t_index++) {                 /* 3 tones + uniform noise, ~ -6 dBFS */
            const double t = (double)t_index;
            double x = 400.0 * sin(2.0 * M_PI * 0.1444 * t) + 350.0 * sin(2.0 * M_PI * 0.2871 * t + 1.0) +
                       250.0 * sin(2.0 * M_PI * 0.0391 * t + 2.0);
            x += (double)((int)(lcg(&seed) >> 20) % 17) - 8.0;
            long v = lrint(x);
            adc[i] = (int16_t)(v > 2047 ? 2047 : (v < -2048 ? -2048 : v));
        }
        size_t nf = 0, nb = 0, nfft = 0;
        const double t0 = now();
        CHECK(ua3reo_ddc_push(bank, adc, block, &nf));
        CHECK(ua3reo_ddc_read_frames(bank, frames, nf));
        if (full) {
            CHECK(ua3reo_rx_counts(bank, &nb, &nfft));
            CHECK(ua3reo_rx_read_audio(bank, audio, nb));
            CHECK(ua3reo_rx_read_spectra(bank, spectra, nfft));
        }
        t_total += now() - t0;
        for (size_t i = 0; i < (size_t)n_ch * nf * 8; i++) sum_frames = sum_frames * 1099511628211ull + frames[i];
        for (size_t i = 0; i < (size_t)n_ch * nb * 384; i++) sum_audio = sum_audio * 1099511628211ull + (uint32_t)audio[i];
        for (size_t i = 0; i < (size_t)n_ch * nfft * 256; i++) sum_spec += spectra[i];
    }
    const double units = (double)n_ch * (double)block * n_blocks;
    printf("frames checksum %016llx  audio checksum %016llx  spectra sum %.6g\n", (unsigned long long)sum_frames,
           (unsigned long long)sum_audio, sum_spec);
    printf("end to end: %.3f s, %.4g channel*ADC-samples/s (%.1f real-time channels), %llu kernel launches\n", t_total,
           units / t_total, units / t_total / 49152000.0, (unsigned long long)ua3reo_launch_count(bank));
    ua3reo_destroy(bank);
    free(adc); free(frames); free(audio); free(spectra);
    return 0;
}
