/*
 * duc_golden.c - register-transfer golden model of the UA3REO transmit DUC (mirror of the DDC).
 * TEST INFRASTRUCTURE ONLY (see ddc_golden.h).  Restates
 *   FPGA/tx_ciccomp.vhd:199-483  48-tap CIC-compensating polyphase interpolator x2 (coefficients s16.14)
 *   FPGA/tx_cic.vhd:153-420      5-stage CIC interpolator x512 with Hogenauer pruning shifts
 *   FPGA/tx_mixer.v:64-71        signed 14x14 -> 28 (I * sin14, Q * cos14; un-truncated NCO outputs)
 *   FPGA/tx_summator.v:74-80     signed 28-bit add, overflow flag -> DAC_OTR
 *   FPGA/DAC_corrector.v:15-21   28 -> 14 bit, offset binary
 * [convention] tx_ciccomp (2.208 MHz clock) and tx_cic (49.152 MHz / 512 strobe) run on independent
 * clocks; we define: tx_cic's m-th input latch sees the m-th compensator output z[m].
 * [convention] lpm_mult / lpm_add_sub pipeline registers (1 clock each) are modelled as zero latency.
 */
#include "ddc_golden.h"
#include "tables/ddc_tables.h"
#include <string.h>

static inline int64_t sext64(int64_t v, int bits)
{
    const uint64_t m = 1ull << (bits - 1);
    const uint64_t x = (uint64_t)v & ((1ull << bits) - 1);
    return (int64_t)((x ^ m) - m);
}

/* ---- tx_ciccomp: z[2n] = round14(sum c1[i] dp[i]), z[2n+1] = round14(sum c2[i] dp[i]) ----
 * dp shifts on count 0 (:224-234); counts 1..23 accumulate the phase-1 branch, counts 24..45,0 the phase-2
 * branch; the two folded tap pairs (3,17) and (6,20) (:238-240) are exact because c[17] = -c[3].  Output
 * rounding (:469): convergent on the low 30 bits, wrap. */
void ua3g_tx_ciccomp_reset(ua3g_tx_ciccomp *c) { memset(c, 0, sizeof *c); }

static int16_t tx_round14(int64_t acc)
{
    const int64_t a30 = sext64(acc, 30);
    const int64_t r = sext64(a30 + 0x1FFF + ((acc >> 14) & 1), 30);
    return (int16_t)sext64(r >> 14, 16);
}

void ua3g_tx_ciccomp_push(ua3g_tx_ciccomp *c, int16_t x, int16_t z[2])
{
    memmove(&c->dp[1], &c->dp[0], 23 * sizeof(int16_t));
    c->dp[0] = x;
    int64_t a1 = 0, a2 = 0;
    for (int i = 0; i < 24; i++) {
        a1 += (int64_t)UA3_TXCOMP_C1[i] * c->dp[i];
        a2 += (int64_t)UA3_TXCOMP_C2[i] * c->dp[i];
    }
    z[0] = tx_round14(sext64(a1, 37));
    z[1] = tx_round14(sext64(a2, 37));
}

/* ---- tx_cic: combs at the input rate, zero stuffing, five integrators at 49.152 MHz, 60-bit wrap ---- */
void ua3g_tx_cic_reset(ua3g_tx_cic *c) { memset(c, 0, sizeof *c); }

int16_t ua3g_tx_cic_clock(ua3g_tx_cic *c, int16_t w_if_phase0)
{
    const int phase_0 = (c->cnt == 0);                                   /* :172 */
    int64_t up = 0;
    int64_t v1 = 0, v2 = 0, v3 = 0, v4 = 0, o4 = 0;
    if (phase_0) {
        /* comb chain is combinational from the PRE-edge input_register and diff registers (:186-289) */
        v1 = sext64((int64_t)c->wreg * ((int64_t)1 << 43), 60);          /* section_cast1 :189 */
        const int64_t o1 = sext64(v1 - c->d[0], 60);
        v2 = o1 >> 1;                                                    /* section_in2(59 DOWNTO 1) :211 */
        const int64_t o2 = sext64(v2 - c->d[1], 60);
        v3 = o2 >> 1;                                                    /* :233 */
        const int64_t o3 = sext64(v3 - c->d[2], 60);
        v4 = o3 >> 1;                                                    /* :255 */
        o4 = sext64(v4 - c->d[3], 60);
        up = sext64(o4 - c->d[4], 60);                                   /* section_out5, upsampling :291-293 */
    }
    /* integrators: every stage adds the PRE-edge value of the stage before it, pruned by 8 bits (:296-399) */
    const int64_t n10 = sext64(c->i[4] + (c->i[3] >> 8), 60);
    const int64_t n9 = sext64(c->i[3] + (c->i[2] >> 8), 60);
    const int64_t n8 = sext64(c->i[2] + (c->i[1] >> 8), 60);
    const int64_t n7 = sext64(c->i[1] + (c->i[0] >> 8), 60);
    const int64_t n6 = sext64(c->i[0] + up, 60);
    c->out14 = (int16_t)sext64(c->i[4] >> 46, 14);                       /* output_register <= section_out10(59:46) :403-415 */
    c->i[4] = n10; c->i[3] = n9; c->i[2] = n8; c->i[1] = n7; c->i[0] = n6;
    if (phase_0) {
        c->d[0] = v1; c->d[1] = v2; c->d[2] = v3; c->d[3] = v4; c->d[4] = o4;
        c->wreg = w_if_phase0;                                           /* input_register :176-184 */
    }
    c->cnt = (c->cnt >= 511) ? 0 : c->cnt + 1;                           /* :155-168 */
    return c->out14;
}

/* tx_mixer (I*sin14, Q*cos14), tx_summator (wrap + overflow), DAC_corrector (x[27:14] + 8191, 14 bits) */
uint16_t ua3g_dac_word(int32_t i14, int32_t q14, int32_t sin14, int32_t cos14, int *overflow)
{
    const int32_t mi = i14 * sin14, mq = q14 * cos14;                    /* s28 each */
    const int32_t full = mi + mq;
    const int32_t sum = (int32_t)sext64(full, 28);
    if (overflow) *overflow = (sum != full);
    const int32_t t2 = sum >> 14;                                        /* {sum[27], (2*sum)[27:15]} == sum[27:14] */
    return (uint16_t)((t2 + 8191) & 0x3FFF);
}

void ua3g_duc_init(ua3g_duc *d, uint32_t fcw22)
{
    memset(d, 0, sizeof *d);
    d->fcw = fcw22 & 0x3FFFFF;
}

void ua3g_duc_push(ua3g_duc *d, const int16_t *tx_i, const int16_t *tx_q, size_t n, uint16_t *dac, uint8_t *otr)
{
    for (size_t s = 0; s < n; s++) {
        int16_t zi[2], zq[2];
        ua3g_tx_ciccomp_push(&d->comp_i, tx_i[s], zi);
        ua3g_tx_ciccomp_push(&d->comp_q, tx_q[s], zq);
        for (int h = 0; h < 2; h++) {
            for (int t = 0; t < 512; t++) {
                const int16_t oi = ua3g_tx_cic_clock(&d->cic_i, zi[h]);
                const int16_t oq = ua3g_tx_cic_clock(&d->cic_q, zq[h]);
                int32_t s14, c14;
                ua3g_nco(d->phase, &s14, &c14);
                d->phase = (d->phase + d->fcw) & 0x3FFFFF;
                int ov = 0;
                const uint16_t w = ua3g_dac_word(oi, oq, s14, c14, &ov);
                const size_t k = (s * 2 + (size_t)h) * 512 + (size_t)t;
                dac[k] = w;
                if (otr) otr[k] = (uint8_t)ov;
            }
        }
        d->n_in++;
    }
}
