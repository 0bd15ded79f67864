/*
 * ddc_golden.h - CPU golden model of the UA3REO FPGA receive DDC and transmit DUC.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may build, link
 * or execute it, and only as the checker.  The product (libua3reo_b200.so) never links it.
 *
 * The model is a register-transfer restatement, one C statement per HDL register, of
 *   FPGA/nco/synthesis/submodules/nco_nco_ii_0.v:299-420  (NCO II, multiplier architecture)
 *   FPGA/nco_shift.v:9, FPGA/mixer.v:64-71, FPGA/rx_mixer_shift.v:9
 *   FPGA/rx_cic.vhd:147-404, FPGA/rx_ciccomp.vhd:247-630, FPGA/rx_hilb.vhd:353-949
 *   FPGA/data_delay.v:16-30 (N=130, UA3REO.bdf:1693-1694), FPGA/stm32_interface.v:228-271
 *   FPGA/tx_ciccomp.vhd:199-483, FPGA/tx_cic.vhd:153-420, FPGA/tx_mixer.v:64-71,
 *   FPGA/tx_summator.v:74-80, FPGA/DAC_corrector.v:15-21
 *
 * PARITY PINNING: the reference ships no testbench, golden vector or known-answer test for this path, and no HDL
 * simulator exists in the build image.  The model is pinned by EXECUTING THE REFERENCE'S VHDL: tools/vhdl_eval.py
 * parses rx_cic / rx_ciccomp / rx_hilb / tx_cic / tx_ciccomp.vhd where they lie and evaluates them cycle by cycle
 * (a Python interpreter and a C translation built into oracle/_ref/libua3_hdl.so, run against each other);
 * tests/test_hdl_pin.py requires every filter function here to reproduce the HDL's output register edge for edge
 * (rx_cic, tx_cic) or sample for sample (the serial-MAC filters), requires the whole receive chain - four clock
 * domains composed as UA3REO.bdf wires them (tools/bdf_netlist.py) - to reproduce the HDL's frame for every clocking
 * class, and compares both with the committed HDL vectors tests/golden/hdl_cases.npz.  Beyond that: every
 * coefficient/ROM constant is read mechanically from the HDL by tools/gen_tables.py and checksummed, and tests/ hold
 * the analytic identities (DC gain, impulse responses, rate identities).
 * What stays a CONVENTION, marked [convention] in ddc_golden.c: the NCO start phase/latency and its 28 -> 14 bit
 * rounding (encrypted Altera IP, no source to execute), the ideal-PLL edge alignment, and the Verilog one-liners
 * (mixer = signed lpm_mult, nco_shift, rx_mixer_shift, data_delay, tx_summator, DAC_corrector), which are restated.
 */
#ifndef UA3_DDC_GOLDEN_H
#define UA3_DDC_GOLDEN_H
#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define UA3G_CIC_R 512
#define UA3G_COMP_TAPS 65
#define UA3G_HILB_TAPS 256
#define UA3G_QDELAY 130

typedef struct {
    uint32_t cnt;        /* cur_count, 9 bit           rx_cic.vhd:147-160 */
    int32_t inreg;       /* input_register, s23        rx_cic.vhd:178-187 */
    uint64_t s[5];       /* section_out1..5 (mod 2^60 kept mod 2^64)  :197-289 */
    uint64_t d[5];       /* diff1..5                   :300-389 */
    int16_t outreg;      /* output_register            :393-404 */
} ua3g_rx_cic;

typedef struct {
    int16_t p0[33];      /* input_pipeline_phase0      rx_ciccomp.vhd:339-349 */
    int16_t p1[33];      /* input_pipeline_phase1      rx_ciccomp.vhd:351-361 */
    uint32_t n_in;       /* number of inputs consumed (parity selects the branch) */
} ua3g_rx_ciccomp;

typedef struct { int16_t dl[UA3G_HILB_TAPS]; } ua3g_rx_hilb;     /* delay_pipeline rx_hilb.vhd:375-385 */
typedef struct { int16_t dl[UA3G_QDELAY]; } ua3g_delay;          /* data_delay.v:9-30 */

/* Frame clocking class (see "clocking" in ddc_golden.c): which compensator polyphase branch the first CIC output
 * enters, how many 48 kHz samples VOICE_I trails the Hilbert sum, how many samples VOICE_Q trails SPEC_Q. */
typedef struct {
    int align_b;         /* 0: CIC outputs 0,2,4.. -> input_pipeline_phase1 ("A"); 1: -> input_pipeline_phase0 ("B") */
    int d_i;             /* 0..3: VOICE_I[n] = hilbert(SPEC_I)[n - d_i] */
    int d_q;             /* 1..130: VOICE_Q[n] = SPEC_Q[n - d_q] */
} ua3g_clocking;
#define UA3G_CLOCKING_DEFAULT_ALIGN_B 1
#define UA3G_CLOCKING_DEFAULT_DI 3
#define UA3G_CLOCKING_DEFAULT_DQ 129
#define UA3G_MAX_DI 3

typedef struct {
    uint32_t fcw;        /* 22-bit tuning word (stm32_interface.v:159-169) */
    uint32_t phase;      /* 22-bit accumulator */
    ua3g_rx_cic cic_i, cic_q;
    ua3g_rx_ciccomp comp_i, comp_q;
    ua3g_rx_hilb hilb;
    ua3g_delay qdelay;
    uint64_t n_adc;      /* ADC samples consumed */
    uint64_t n_frames;   /* 48 kHz frames produced */
    ua3g_clocking clk;
    int16_t vi_fifo[UA3G_MAX_DI + 1];   /* Hilbert outputs waiting d_i samples */
} ua3g_ddc;

/* NCO: sin14/cos14 for a 22-bit phase (A.1). */
void ua3g_nco(uint32_t phase22, int32_t *sin14, int32_t *cos14);
/* mixer + rx_mixer_shift: returns the s23 value presented to rx_cic.filter_in */
int32_t ua3g_rx_mix(int32_t adc12, int32_t nco14);

void ua3g_rx_cic_reset(ua3g_rx_cic *c);
/* one clk edge with filter_in = x (s23). returns 1 when output_register was (re)loaded on this edge */
int ua3g_rx_cic_clock(ua3g_rx_cic *c, int32_t x);

void ua3g_rx_ciccomp_reset(ua3g_rx_ciccomp *c);
/* one 96 kHz input; returns 1 and writes *y when a 48 kHz output is produced */
int ua3g_rx_ciccomp_push(ua3g_rx_ciccomp *c, int16_t u, int16_t *y);

void ua3g_rx_hilb_reset(ua3g_rx_hilb *h);
int16_t ua3g_rx_hilb_push(ua3g_rx_hilb *h, int16_t y);

void ua3g_delay_reset(ua3g_delay *d);
int16_t ua3g_delay_push(ua3g_delay *d, int16_t q);

/* whole RX DDC for one channel; clocking class = the board's most frequent one (B, 3, 129) */
void ua3g_ddc_init(ua3g_ddc *d, uint32_t fcw22);
/* same with an explicit clocking class; returns -1 when a field is out of range */
int ua3g_ddc_init_clocking(ua3g_ddc *d, uint32_t fcw22, int align_b, int d_i, int d_q);
/* Feeds n ADC samples (12-bit two's complement, sign-extended int16). Writes produced 8-byte
 * frames (stm32_interface order: SPEC_Q hi,lo, SPEC_I hi,lo, VOICE_Q hi,lo, VOICE_I hi,lo) to
 * frames (capacity max_frames) and returns how many were produced.  If cic_i/cic_q are non-NULL
 * they receive every CIC output register load (96 kHz), capacity max_cic, count in *n_cic. */
size_t ua3g_ddc_push(ua3g_ddc *d, const int16_t *adc, size_t n, uint8_t *frames, size_t max_frames,
                     int16_t *cic_i, int16_t *cic_q, size_t max_cic, size_t *n_cic);

/* stm32_interface.v:228-271 <-> fpga.c:286-401 */
void ua3g_frame_pack(uint8_t f[8], int16_t spec_q, int16_t spec_i, int16_t voice_q, int16_t voice_i);
void ua3g_frame_unpack(const uint8_t f[8], int16_t *spec_q, int16_t *spec_i, int16_t *voice_q, int16_t *voice_i);

/* functions.c:206-226 getPhraseFromFrequency: returns FCW and the IQ-swap flag */
uint32_t ua3g_phrase_from_frequency(uint32_t freq_hz, int *iq_swap);

/* ---------------- TX DUC mirror ---------------- */
typedef struct { int16_t dp[24]; } ua3g_tx_ciccomp;              /* tx_ciccomp.vhd:224-240 */
typedef struct {
    uint32_t cnt;
    int16_t wreg;                                                 /* input_register */
    int64_t d[5];                                                 /* comb delays (s60 kept in int64 with explicit wrap) */
    int64_t up;                                                   /* zero-stuffed comb output latched at phase_0 */
    int64_t i[5];                                                 /* integrators section_out6..10 */
    int16_t out14;
} ua3g_tx_cic;
typedef struct {
    uint32_t fcw, phase;
    ua3g_tx_ciccomp comp_i, comp_q;
    ua3g_tx_cic cic_i, cic_q;
    uint64_t n_in;
} ua3g_duc;

void ua3g_tx_ciccomp_reset(ua3g_tx_ciccomp *c);
void ua3g_tx_ciccomp_push(ua3g_tx_ciccomp *c, int16_t x, int16_t z[2]);
void ua3g_tx_cic_reset(ua3g_tx_cic *c);
/* one 49.152 MHz clock; new_in is consumed when the internal counter is at phase 0 */
int16_t ua3g_tx_cic_clock(ua3g_tx_cic *c, int16_t w_if_phase0);
uint16_t ua3g_dac_word(int32_t i14, int32_t q14, int32_t sin14, int32_t cos14, int *overflow);
void ua3g_duc_init(ua3g_duc *d, uint32_t fcw22);
/* n 48 kHz I/Q samples in, 1024*n DAC words (u14 offset binary) out */
void ua3g_duc_push(ua3g_duc *d, const int16_t *tx_i, const int16_t *tx_q, size_t n, uint16_t *dac, uint8_t *otr);

#ifdef __cplusplus
}
#endif
#endif
