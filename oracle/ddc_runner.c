/*
 * ddc_runner.c - multi-threaded driver of the golden DDC, used ONLY by bench.py's cpu_baseline and
 * --impl reference legs (TEST INFRASTRUCTURE; see ddc_golden.h).  Channels are distributed
 * round-robin over threads, each thread running the scalar register-transfer model.
 */
#include "ddc_golden.h"
#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

typedef struct {
    ua3g_ddc *st;            /* [n_ch] */
    const int16_t *adc;
    size_t n;
    uint8_t *frames;         /* [n_ch][n/1024][8] or NULL */
    int n_ch, tid, n_thr;
} job_t;

static void *worker(void *p)
{
    job_t *j = (job_t *)p;
    const size_t nf = j->n / 1024 + 1;
    uint8_t *scratch = (uint8_t *)malloc(nf * 8);
    for (int c = j->tid; c < j->n_ch; c += j->n_thr) {
        uint8_t *dst = j->frames ? j->frames + (size_t)c * (j->n / 1024) * 8 : scratch;
        ua3g_ddc_push(&j->st[c], j->adc, j->n, dst, j->n / 1024, NULL, NULL, 0, NULL);
    }
    free(scratch);
    return NULL;
}

/* Allocates n_ch channel states with the given tuning words. */
void *ua3g_bank_create(int n_ch, const uint32_t *fcw)
{
    ua3g_ddc *st = (ua3g_ddc *)calloc((size_t)n_ch, sizeof(ua3g_ddc));
    for (int c = 0; c < n_ch; c++) ua3g_ddc_init(&st[c], fcw[c]);
    return st;
}
void ua3g_bank_destroy(void *bank) { free(bank); }

/* Runs one ADC block through all channels on n_thr threads; returns wall seconds. */
double ua3g_bank_push(void *bank, int n_ch, const int16_t *adc, size_t n, uint8_t *frames, int n_thr)
{
    struct timespec t0, t1;
    if (n_thr < 1) n_thr = 1;
    if (n_thr > n_ch) n_thr = n_ch;
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)n_thr);
    job_t *jobs = (job_t *)malloc(sizeof(job_t) * (size_t)n_thr);
    clock_gettime(CLOCK_MONOTONIC, &t0);
    for (int t = 0; t < n_thr; t++) {
        jobs[t] = (job_t){(ua3g_ddc *)bank, adc, n, frames, n_ch, t, n_thr};
        pthread_create(&th[t], NULL, worker, &jobs[t]);
    }
    for (int t = 0; t < n_thr; t++) pthread_join(th[t], NULL);
    clock_gettime(CLOCK_MONOTONIC, &t1);
    free(th); free(jobs);
    return (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
}
