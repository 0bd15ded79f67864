/*
 * bus_hdl.c - the MCU <-> FPGA byte bus with the reference on BOTH ends.  TEST INFRASTRUCTURE ONLY.
 * MCU side: the firmware's own bus driver fpga.c, compiled unmodified with -DUA3_BUS_HDL, so that its pin writes
 * (GPIOA = data bus D0..D7, GPIOC pin 0 = FPGA_CLK, pin 1 = FPGA_SYNC; main.h:100-123) arrive here one by one.
 * FPGA side: stm32_interface.v translated from the reference's source text by tools/verilog_eval.py (oracle/_ref/hdl/vlog.c):
 *   - a rising edge of FPGA_CLK is `posedge clk_in` with DATA_SYNC = the SYNC pin and the bus carrying what the MCU drives;
 *   - bus resolution: while the MCU's pins are outputs (GPIOA->MODER, switched by FPGA_setBusOutput / FPGA_setBusInput) the
 *     bus carries the MCU's output latch; otherwise the module's DATA_BUS_OUT while it enables its drivers, else the pull-ups
 *     (0xFF).  [convention] stm32_interface.v keeps DATA_BUS_OE = 1 after a read command until the NEXT command's DATA_SYNC
 *     edge, on which it also samples the command byte - so the MCU's command byte always meets the FPGA's still enabled
 *     drivers.  An event simulator reads X there; the board works because the MCU's push-pull outputs win that fight, which
 *     is the rule modelled here (the only one under which a command can follow a read at all);
 *   - `posedge adcclk_in` (ADC min/max tracking, key debounce) is clocked by the harness (ua3_hdl_adc_clock).
 */
#include "stm32f4xx_hal.h"
#include "../_ref/hdl/vlog.c"
#include <stdio.h>
#include <stdlib.h>

static stm32_interface_t g_if;
static int g_ready = 0;
unsigned long ua3_hdl_clk_edges = 0;
/* what crossed the bus, for the wire-format vectors: bytes the module sampled on data clocks (SYNC low) and bytes the MCU read */
uint8_t ua3_hdl_wr_log[64], ua3_hdl_rd_log[64];
unsigned ua3_hdl_n_wr = 0, ua3_hdl_n_rd = 0;
void ua3_hdl_log_reset(void) { ua3_hdl_n_wr = ua3_hdl_n_rd = 0; }

static int g_trace = 0;                                         /* UA3_HDL_TRACE=1: one line per FPGA_CLK edge on stderr */
static int mcu_drives(void) { return (ua3_gpio_a.MODER & 0xFFFFu) == 0x5555u; }      /* D0..D7 general purpose outputs */
static void ready(void) { if (!g_ready) { stm32_interface_init(&g_if); g_ready = 1; g_trace = getenv("UA3_HDL_TRACE") != NULL; } }

/* the value on the bus pins, as the module's input buffers and the MCU's IDR see it */
static void resolve(void)
{
    const uint64_t mcu = ua3_gpio_a.ODR & 0xFFu;
    g_if.DATA_BUS__ext = mcu_drives() ? mcu : 0xFFu;
    stm32_interface_settle(&g_if);                              /* assign DATA_BUS = DATA_BUS_OE ? DATA_BUS_OUT : 'bZ */
    if (mcu_drives()) g_if.DATA_BUS = mcu;                      /* contention: the MCU wins */
}

static void apply(GPIO_TypeDef *g, int is_clock_port)
{
    const uint32_t v = g->bsrr_slot[0];
    if (!v) return;
    g->bsrr_slot[0] = 0;
    const uint32_t before = g->ODR;
    g->ODR = (before & ~(v >> 16)) | (v & 0xFFFFu);             /* reset bits, then set bits (set wins) */
    if (is_clock_port && !(before & 1u) && (g->ODR & 1u)) {     /* FPGA_CLK rises */
        g_if.DATA_SYNC = (g->ODR >> 1) & 1u;
        resolve();
        if (!g_if.DATA_SYNC && mcu_drives()) ua3_hdl_wr_log[ua3_hdl_n_wr++ & 63] = (uint8_t)g_if.DATA_BUS;
        stm32_interface_posedge_clk_in(&g_if);
        ua3_hdl_clk_edges++;
        if (g_trace) fprintf(stderr, "clk: moder=%04x sync=%d bus_in=%3d -> k=%d oe=%d out=%3d\n", (unsigned)(ua3_gpio_a.MODER & 0xFFFFu), (int)g_if.DATA_SYNC, (int)g_if.DATA_BUS__ext,
                             (int)(int16_t)g_if.k, (int)g_if.DATA_BUS_OE, (int)g_if.DATA_BUS_OUT);
    }
}

int ua3_bsrr_hook(void)
{
    ready();
    apply(&ua3_gpio_a, 0);                                      /* data first: the firmware always sets data before the clock */
    apply(&ua3_gpio_c, 1);
    return 0;
}

uint32_t ua3_hdl_bus_read(void)
{
    ua3_bsrr_hook();
    resolve();
    ua3_hdl_rd_log[ua3_hdl_n_rd++ & 63] = (uint8_t)g_if.DATA_BUS;
    return (uint32_t)g_if.DATA_BUS;
}

void ua3_hdl_attach(void) { ready(); ua3_gpio_a.idr_fn = ua3_hdl_bus_read; }
void ua3_hdl_flush(void) { ua3_bsrr_hook(); }

void ua3_hdl_set_iq(int16_t spec_i, int16_t spec_q, int16_t voice_i, int16_t voice_q)
{
    ready();
    g_if.SPEC_I = (uint16_t)spec_i; g_if.SPEC_Q = (uint16_t)spec_q;
    g_if.VOICE_I = (uint16_t)voice_i; g_if.VOICE_Q = (uint16_t)voice_q;
}

void ua3_hdl_set_flags(int adc_otr, int dac_otr) { ready(); g_if.ADC_OTR = adc_otr & 1; g_if.DAC_OTR = dac_otr & 1; }

void ua3_hdl_adc_clock(int16_t adc12)
{
    ready();
    ua3_bsrr_hook();
    g_if.ADC_IN = (uint64_t)adc12 & 0xFFFu;
    stm32_interface_posedge_adcclk_in(&g_if);
}

/* any signal of the module by name (raw bits); -1 if there is none */
int64_t ua3_hdl_get(const char *name)
{
    ready();
    ua3_bsrr_hook();
    const vl_module *m = vl_find("stm32_interface");
    for (int i = 0; i < m->n_fields; i++)
        if (!strcmp(m->fields[i].name, name)) return (int64_t)*(const uint64_t *)((const char *)&g_if + m->fields[i].offset);
    return -1;
}
