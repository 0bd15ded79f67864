/*
 * ddc_golden.c - register-transfer golden model of the UA3REO receive DDC (see ddc_golden.h).
 * TEST INFRASTRUCTURE ONLY - never linked into the product.
 *
 * Every function cites the HDL it restates.  "pre-edge" means the register value before the
 * clock edge being simulated; all registers of one module update simultaneously.
 */
#include "ddc_golden.h"
#include "tables/ddc_tables.h"
#include <string.h>
#include <math.h>

static inline int64_t sext(int64_t v, int bits)
{
    const uint64_t m = 1ull << (bits - 1);
    uint64_t x = (uint64_t)v & ((bits == 64) ? ~0ull : ((1ull << bits) - 1));
    return (int64_t)((x ^ m) - m);
}

/* ------------------------------------------------------------------------------------------
 * NCO - nco_nco_ii_0.v:299-420.  apr=22 (line 28), rawc=rawf=11 (:34-37): coarse address =
 * phase[21:11], fine address = phase[10:0] (:322).  ROM words are s14 (mpr=14).  The two
 * multiply-add blocks are wired (:370-392) as
 *     cos = cos_c*cos_f - sin_c*sin_f      (asj_nco_madx_cen)
 *     sin = sin_c*cos_f + sin_f*cos_c      (asj_nco_mady_cen)
 * giving 28-bit results (opr=28) reduced to 14 bits by asj_nco_mob_w (:404-420).
 * [convention] mob_w is encrypted; we define the reduction as round-half-up: (x + 2^12) >> 13.
 * [convention] phase accumulator starts at 0 and has zero latency: the sample multiplied with
 *              ADC sample n uses phase n*FCW mod 2^22.
 * ------------------------------------------------------------------------------------------ */
void ua3g_nco(uint32_t phase22, int32_t *sin14, int32_t *cos14)
{
    const uint32_t k = (phase22 >> 11) & 0x7FF, j = phase22 & 0x7FF;
    const int32_t sc = UA3_NCO_SIN_C[k], cc = UA3_NCO_COS_C[k];
    const int32_t sf = UA3_NCO_SIN_F[j], cf = UA3_NCO_COS_F[j];
    const int32_t sin28 = sc * cf + sf * cc;
    const int32_t cos28 = cc * cf - sc * sf;
    *sin14 = (int32_t)sext((sin28 + 4096) >> 13, 14);
    *cos14 = (int32_t)sext((cos28 + 4096) >> 13, 14);
}

/* nco_shift.v:9 (out = in[13:2]), mixer.v:64-71 (signed 12x12 -> 24), rx_mixer_shift.v:9
 * (out = in[22:0], i.e. the 24-bit product reinterpreted as s23). */
int32_t ua3g_rx_mix(int32_t adc12, int32_t nco14)
{
    const int32_t nco12 = nco14 >> 2;
    const int32_t p24 = adc12 * nco12;
    return (int32_t)sext(p24, 23);
}

/* ------------------------------------------------------------------------------------------
 * rx_cic.vhd - 5 integrators @ clk, 5 combs @ clk/512, 60-bit wrapping registers.
 * 60-bit two's-complement wrap is kept as mod-2^64 arithmetic: only bits [59:44] are ever
 * observed (:391) and those agree mod 2^60 and mod 2^64.
 * ------------------------------------------------------------------------------------------ */
void ua3g_rx_cic_reset(ua3g_rx_cic *c) { memset(c, 0, sizeof *c); }

int ua3g_rx_cic_clock(ua3g_rx_cic *c, int32_t x)
{
    const int phase_1 = (c->cnt == 1);                                   /* :162 */
    const int64_t x15 = (int64_t)(c->inreg >> 8);                        /* section_in1(22 DOWNTO 8), :193 */
    /* comb chain is combinational from pre-edge section_out5 and diff registers (:293-378) */
    const uint64_t c1 = c->s[4] - c->d[0];
    const uint64_t c2 = c1 - c->d[1];
    const uint64_t c3 = c2 - c->d[2];
    const uint64_t c4 = c3 - c->d[3];
    const uint64_t c5 = c4 - c->d[4];
    if (phase_1) {
        c->outreg = (int16_t)((c5 >> 44) & 0xFFFF);                      /* :391,:400-402 */
        c->d[0] = c->s[4]; c->d[1] = c1; c->d[2] = c2; c->d[3] = c3; c->d[4] = c4;  /* :305-307 ... */
    }
    /* integrators: each stage adds the REGISTERED previous stage (:197-289), so update last first */
    c->s[4] += c->s[3];
    c->s[3] += c->s[2];
    c->s[2] += c->s[1];
    c->s[1] += c->s[0];
    c->s[0] += (uint64_t)x15;
    c->inreg = x;                                                        /* :184 */
    c->cnt = (c->cnt >= 511) ? 0 : c->cnt + 1;                           /* :153-157 */
    return phase_1;
}

/* ------------------------------------------------------------------------------------------
 * rx_ciccomp.vhd - 65-tap symmetric FIR, decimate by 2, polyphase (33 + 32 taps), coefficients
 * s16.15 (:82-146).  The serial MAC (:425-484), the shift-only "power of two" taps (:493-578)
 * and the 56-bit accumulator (:589-612) are all exact, so the sum is computed directly in int64.
 * Phase-0 delay line shifts on count 0, phase-1 on count 16 (:339-361): the sample taken into
 * phase-1 is the older one of each pair.
 * Which CIC output falls into which branch depends on when RX is raised relative to the compensator clock
 * (see "Clocking" at ua3g_ddc_init_clocking).  This function implements alignment A when n_in starts at 0:
 * even-indexed CIC outputs u[2k] -> phase-1 pipeline, odd-indexed u[2k+1] -> phase-0 pipeline, one output per
 * pair, computed when the odd sample arrives:  y[k] = sum_j h[j] * u[2k+1-j];  with n_in starting at 1 it is
 * alignment B, y[k] = sum_j h[j] * u[2k-j].  Either way it reproduces rx_ciccomp.vhd sample for sample
 * (tests/test_hdl_pin.py).
 * Output rounding (:616): convergent, on the low 31 bits, wrap not saturate.
 * ------------------------------------------------------------------------------------------ */
void ua3g_rx_ciccomp_reset(ua3g_rx_ciccomp *c) { memset(c, 0, sizeof *c); }

int ua3g_rx_ciccomp_push(ua3g_rx_ciccomp *c, int16_t u, int16_t *y)
{
    const int odd = (int)(c->n_in & 1);
    c->n_in++;
    if (!odd) {                       /* phase_16: shift into input_pipeline_phase1 */
        memmove(&c->p1[1], &c->p1[0], 32 * sizeof(int16_t));
        c->p1[0] = u;
        return 0;
    }
    memmove(&c->p0[1], &c->p0[0], 32 * sizeof(int16_t));   /* phase_0 */
    c->p0[0] = u;
    int64_t acc = 0;
    for (int i = 0; i < 33; i++) acc += (int64_t)UA3_RXCOMP_H[2 * i] * c->p0[i];      /* coeffphase1_(i+1) */
    for (int i = 0; i < 32; i++) acc += (int64_t)UA3_RXCOMP_H[2 * i + 1] * c->p1[i];  /* coeffphase2_(i+1) */
    const int64_t a31 = sext(acc, 31);
    const int64_t r = sext(a31 + 0x3FFF + ((acc >> 15) & 1), 31);
    *y = (int16_t)sext(r >> 15, 16);
    return 1;
}

/* ------------------------------------------------------------------------------------------
 * rx_hilb.vhd - 256-tap Hilbert FIR (I rail), fully serial.  Per tap (:902-903):
 *    mul_temp = x*c (s32, En30); product = (mul_temp + mul_temp[1]) >> 1  (convergent, En29)
 * accumulated in a 40-bit wrapping register (:907-931); output (:935) convergent >>14 on the
 * low 30 bits, wrap.  delay_pipeline(0) is the newest sample and pairs with coeff1 (:375-640).
 * The 2-sample latency of the serial MAC + output register (measured on the VHDL, tests/test_hdl_pin.py) is not
 * in this function; the frame assembly in ua3g_ddc_push applies it (d_i).
 * ------------------------------------------------------------------------------------------ */
void ua3g_rx_hilb_reset(ua3g_rx_hilb *h) { memset(h, 0, sizeof *h); }

int16_t ua3g_rx_hilb_push(ua3g_rx_hilb *h, int16_t yin)
{
    memmove(&h->dl[1], &h->dl[0], (UA3G_HILB_TAPS - 1) * sizeof(int16_t));
    h->dl[0] = yin;
    int64_t acc = 0;
    for (int k = 0; k < UA3G_HILB_TAPS; k++) {
        const int32_t p = (int32_t)h->dl[k] * (int32_t)UA3_RXHILB_C[k];
        const int32_t pr = (int32_t)(((int64_t)p + ((p >> 1) & 1)) >> 1);
        acc = sext(acc + pr, 40);
    }
    const int64_t a30 = sext(acc, 30);
    const int64_t r = sext(a30 + 0x1FFF + ((acc >> 14) & 1), 30);
    return (int16_t)sext(r >> 14, 16);
}

/* data_delay.v:16-30 with N=130 (UA3REO.bdf:1693-1694): out[n] = in[n-130], zeros before. */
void ua3g_delay_reset(ua3g_delay *d) { memset(d, 0, sizeof *d); }

int16_t ua3g_delay_push(ua3g_delay *d, int16_t q)
{
    const int16_t out = d->dl[UA3G_QDELAY - 1];
    memmove(&d->dl[1], &d->dl[0], (UA3G_QDELAY - 1) * sizeof(int16_t));
    d->dl[0] = q;
    return out;
}

/* stm32_interface.v:228-271 (k=400..407 byte order) <-> fpga.c:286-401 */
void ua3g_frame_pack(uint8_t f[8], int16_t spec_q, int16_t spec_i, int16_t voice_q, int16_t voice_i)
{
    f[0] = (uint8_t)((uint16_t)spec_q >> 8);  f[1] = (uint8_t)spec_q;
    f[2] = (uint8_t)((uint16_t)spec_i >> 8);  f[3] = (uint8_t)spec_i;
    f[4] = (uint8_t)((uint16_t)voice_q >> 8); f[5] = (uint8_t)voice_q;
    f[6] = (uint8_t)((uint16_t)voice_i >> 8); f[7] = (uint8_t)voice_i;
}

void ua3g_frame_unpack(const uint8_t f[8], int16_t *spec_q, int16_t *spec_i, int16_t *voice_q, int16_t *voice_i)
{
    *spec_q = (int16_t)((f[0] << 8) | f[1]);
    *spec_i = (int16_t)((f[2] << 8) | f[3]);
    *voice_q = (int16_t)((f[4] << 8) | f[5]);
    *voice_i = (int16_t)((f[6] << 8) | f[7]);
}

/* functions.c:206-226 */
uint32_t ua3g_phrase_from_frequency(uint32_t freq, int *iq_swap)
{
    const uint32_t clk = 49152000u;      /* settings.h:10 ADCDAC_CLOCK */
    int inverted = 0;
    uint32_t f = freq;
    if (f > clk / 2) {
        while (f > clk / 2) { f -= clk / 2; inverted = !inverted; }
        if (inverted) f = clk / 2 - f;
    }
    if (iq_swap) *iq_swap = inverted;
    return (uint32_t)round(((double)f / clk) * 4194304);
}

/* ------------------------------------------------------------------------------------------
 * Whole RX DDC, one channel.  Netlist from UA3REO.bdf (extracted by tools/bdf_netlist.py, checked in
 * tests/test_hdl_pin.py): MIXER_I.datab <- NCO sin, MIXER_Q.datab <- NCO cos; RX_CICCOMP_I/Q.filter_out ->
 * SPEC_I/Q; RX_VOICE_HILBERT_I(comp I) -> VOICE_I; RX_VOICE_DELAY_Q(comp Q) -> VOICE_Q.
 *
 * Clocking.  All four RX filter modules share reset = RX_N and clk_enable = RX, their clocks are MAIN_PLL
 * outputs of clk_sys with zero phase shift (c0 = /32 compensator, c1 = /4 Hilbert, c2 = /1024 Q delay and MCU
 * interrupt), so ONE free parameter decides how the four words of a frame line up: T_rx, the instant the MCU
 * raises RX relative to the PLL clocks (plus tau, how long after the 48 kHz edge the MCU reads the words).
 * tools/hdl_clocking_survey.py sweeps all 1024 values of T_rx through the reference's own VHDL; the frame is
 * always  SPEC = Y[k], VOICE_I = H[k - d_i], VOICE_Q = Y_Q[k - d_q]  with Y the compensator output for one of
 * the two polyphase alignments, H the Hilbert sum of Y_I with zero latency, and
 *     alignment B for 30 of 32 values of T_rx mod 32 (A for the other two),
 *     d_i = 3 for 31 of 32 (2 otherwise: the serial MAC + output register cost two samples, sampling a third),
 *     d_q = 129 when the MCU reads within about 2 us of the edge (tau <= 100 clk_sys ticks: 85 %), else 130.
 * The class is a parameter of the model; the default is the board's most frequent one, (B, 3, 129).  The
 * round-1 model used (A, 0, 130), which no T_rx produces: it skewed VOICE_I against VOICE_Q by 2.5 samples
 * where the board has 1.5 (or 0.5).
 * [convention] ideal PLL: derived clock edges coincide with clk_sys edges and a register clocked at the same
 * instant as its producer sees the producer's previous value.
 * ------------------------------------------------------------------------------------------ */
int ua3g_ddc_init_clocking(ua3g_ddc *d, uint32_t fcw22, int align_b, int d_i, int d_q)
{
    memset(d, 0, sizeof *d);
    d->fcw = fcw22 & 0x3FFFFF;
    if ((align_b != 0 && align_b != 1) || d_i < 0 || d_i > UA3G_MAX_DI || d_q < 1 || d_q > UA3G_QDELAY) return -1;
    d->clk.align_b = align_b;
    d->clk.d_i = d_i;
    d->clk.d_q = d_q;
    /* alignment B: the first CIC output enters input_pipeline_phase0, as if one (zero) output had gone before */
    d->comp_i.n_in = d->comp_q.n_in = (uint32_t)align_b;
    return 0;
}

void ua3g_ddc_init(ua3g_ddc *d, uint32_t fcw22)
{
    ua3g_ddc_init_clocking(d, fcw22, UA3G_CLOCKING_DEFAULT_ALIGN_B, UA3G_CLOCKING_DEFAULT_DI, UA3G_CLOCKING_DEFAULT_DQ);
}

size_t ua3g_ddc_push(ua3g_ddc *d, const int16_t *adc, size_t n, uint8_t *frames, size_t max_frames,
                     int16_t *cic_i, int16_t *cic_q, size_t max_cic, size_t *n_cic)
{
    size_t nf = 0, nc = 0;
    for (size_t t = 0; t < n; t++) {
        int32_t s14, c14;
        ua3g_nco(d->phase, &s14, &c14);
        d->phase = (d->phase + d->fcw) & 0x3FFFFF;
        const int32_t a = adc[t];
        const int32_t xi = ua3g_rx_mix(a, s14);
        const int32_t xq = ua3g_rx_mix(a, c14);
        const int li = ua3g_rx_cic_clock(&d->cic_i, xi);
        const int lq = ua3g_rx_cic_clock(&d->cic_q, xq);
        d->n_adc++;
        if (li && lq) {
            if (cic_i && nc < max_cic) cic_i[nc] = d->cic_i.outreg;
            if (cic_q && nc < max_cic) cic_q[nc] = d->cic_q.outreg;
            nc++;
            int16_t yi = 0, yq = 0;
            const int oi = ua3g_rx_ciccomp_push(&d->comp_i, d->cic_i.outreg, &yi);
            const int oq = ua3g_rx_ciccomp_push(&d->comp_q, d->cic_q.outreg, &yq);
            if (oi && oq) {
                /* VOICE_I trails the zero-latency Hilbert sum by d_i samples, VOICE_Q is SPEC_Q d_q samples ago */
                memmove(&d->vi_fifo[1], &d->vi_fifo[0], UA3G_MAX_DI * sizeof(int16_t));
                d->vi_fifo[0] = ua3g_rx_hilb_push(&d->hilb, yi);
                const int16_t vi = d->vi_fifo[d->clk.d_i];
                const int16_t vq = d->qdelay.dl[d->clk.d_q - 1];
                ua3g_delay_push(&d->qdelay, yq);
                if (frames && nf < max_frames) ua3g_frame_pack(frames + 8 * nf, yq, yi, vq, vi);
                nf++;
                d->n_frames++;
            }
        }
    }
    if (n_cic) *n_cic = nc;
    return nf;
}
