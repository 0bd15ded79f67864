// duc_launch.h - host-visible launch interface of duc.cu (transmit DUC, mirror of the DDC)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace ua3 {

// Per-channel DUC registers between pushes.  60-bit HDL registers are kept LEFT-ALIGNED in 64 bits
// (value * 16) so that two's-complement wrap at 60 bits is the natural wrap of 64-bit arithmetic.
struct DucState {
    int16_t dp[2][24];       // tx_ciccomp delay_pipeline per rail (tx_ciccomp.vhd:224-234)
    int16_t wreg[2];         // tx_cic input_register (tx_cic.vhd:176-184)
    int16_t pad[2];
    int64_t d[2][5];         // comb delays diff1..5, left-aligned
    int64_t i[2][5];         // integrators section_out6..10, left-aligned
    uint32_t phase;          // 22-bit NCO phase
    uint32_t otr;            // summator overflows seen so far (DAC_OTR)
};

struct DucBuffers {
    uint32_t n_ch = 0, max_in = 0;
    const int16_t* nco_tab = nullptr;   // [2048 * 26] 14-bit NCO sine by (coarse address, fine-sine level)
    const uint32_t* fcw = nullptr;      // shared with the DDC: one NCO tuning word per channel
    DucState* state = nullptr;          // [n_ch]
    int16_t* iq_in = nullptr;           // [n_ch][max_in][2]  TX_I, TX_Q (48 kHz, s16)
    uint16_t* dac = nullptr;            // [n_ch][max_in * 1024] 14-bit offset-binary DAC words
};

cudaError_t duc_upload_constants();
void build_duc_nco_table(int16_t* tab);
cudaError_t duc_launch(const DucBuffers& b, uint32_t n_in, cudaStream_t st, int* launches);

}  // namespace ua3
