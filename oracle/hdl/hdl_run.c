/*
 * hdl_run.c - drivers for the C translations of the reference's VHDL filters (the *_hdl.c files under oracle/_ref/hdl, generated
 * by tools/vhdl_eval.py from the .vhd files under /root/reference/FPGA where they lie).  TEST INFRASTRUCTURE ONLY.
 *
 * Every module is driven the way UA3REO.bdf wires it (tools/bdf_netlist.py): `reset` = RX_N / TX_N, `clk_enable`
 * = RX / TX, so "reset held, then released with clk_enable = 1" is the only control sequence that exists on the
 * board.  A run is a list of rising clock edges of the module's OWN clock: in[e] is the value on filter_in when
 * edge e arrives, out[e] the value on filter_out after edge e (ce[e] likewise for ce_out, where the module has
 * one).  Clock-domain crossings are composed by the caller by resampling these per-edge traces (oracle/hdl_ref.py).
 */
#include <stdint.h>
#include <stdlib.h>
#include <stddef.h>

#define DECLARE(mod)                                                                        \
    typedef struct mod##_s mod##_t;                                                         \
    size_t mod##_sizeof(void);                                                              \
    void mod##_clock(mod##_t *);                                                            \
    void mod##_settle(mod##_t *);                                                           \
    void mod##_set_clk(mod##_t *, int64_t);                                                 \
    void mod##_set_clk_enable(mod##_t *, int64_t);                                          \
    void mod##_set_reset(mod##_t *, int64_t);                                               \
    void mod##_set_filter_in(mod##_t *, int64_t);                                           \
    int64_t mod##_get_filter_out(const mod##_t *);

#define RUNNER(mod, CE)                                                                     \
    /* state == NULL: reset, run n edges.  Otherwise *state carries the module across calls  \
       (allocated on first use when *state == NULL). */                                     \
    int hdl_##mod##_run(void **state, const int64_t *in, size_t n, int64_t *out, int64_t *ce) \
    {                                                                                       \
        void *local = NULL;                                                                 \
        if (!state) state = &local;                                                         \
        mod##_t *s = (mod##_t *)*state;                                                     \
        if (!s) {                                                                           \
            s = (mod##_t *)calloc(1, mod##_sizeof());                                       \
            if (!s) return -1;                                                              \
            mod##_set_clk(s, 1);                                                            \
            mod##_set_clk_enable(s, 0);                                                     \
            mod##_set_reset(s, 1);                                                          \
            mod##_clock(s);            /* asynchronous reset branch of every process */     \
            mod##_set_reset(s, 0);                                                          \
            mod##_set_clk_enable(s, 1);                                                     \
            mod##_settle(s);                                                                \
            *state = s;                                                                     \
        }                                                                                   \
        for (size_t e = 0; e < n; e++) {                                                    \
            mod##_set_filter_in(s, in[e]);                                                  \
            mod##_clock(s);                                                                 \
            out[e] = mod##_get_filter_out(s);                                               \
            if (ce) ce[e] = CE;                                                             \
        }                                                                                   \
        if (local) free(local);                                                             \
        return 0;                                                                           \
    }

DECLARE(rx_cic)     int64_t rx_cic_get_ce_out(const rx_cic_t *);
DECLARE(rx_ciccomp) int64_t rx_ciccomp_get_ce_out(const rx_ciccomp_t *);
DECLARE(rx_hilb)
DECLARE(tx_cic)     int64_t tx_cic_get_ce_out(const tx_cic_t *);
DECLARE(tx_ciccomp) int64_t tx_ciccomp_get_ce_out(const tx_ciccomp_t *);

RUNNER(rx_cic, rx_cic_get_ce_out(s))
RUNNER(rx_ciccomp, rx_ciccomp_get_ce_out(s))
RUNNER(rx_hilb, 0)
RUNNER(tx_cic, tx_cic_get_ce_out(s))
RUNNER(tx_ciccomp, tx_ciccomp_get_ce_out(s))

void hdl_free(void *state) { free(state); }
