// tx_launch.h - host-visible launch interface of tx.cu
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>
#include "tx_types.h"

namespace ua3 {

struct TxBuffers {
    uint32_t n_ch = 0, max_blocks = 0;
    TxParams* params = nullptr;     // [n_ch]
    TxState* state = nullptr;       // [n_ch]
    int16_t* mic = nullptr;         // [n_ch][max_blocks * 192][2]  codec samples L, R
    float* iq_f = nullptr;          // [n_ch][max_blocks * 192][2]  I, Q as left in FPGA_Audio_SendBuffer_I/Q
    int16_t* iq_w = nullptr;        // [n_ch][max_blocks * 192][2]  I, Q as FPGA_fpgadata_sendiq() converts them
    int32_t* loop_out = nullptr;    // [n_ch][max_blocks * 192][2]  TRX_MODE_LOOPBACK: what goes to the codec instead (L, R; audio_processor.c:228-249)
};

cudaError_t tx_upload_constants(const float* sin_table);
cudaError_t tx_launch_init_state(const TxBuffers& b, cudaStream_t st, int* launches);
cudaError_t tx_launch_clear(const TxBuffers& b, const uint8_t* flags_dev, uint32_t first, uint32_t n, cudaStream_t st, int* launches);
cudaError_t tx_launch_audio(const TxBuffers& b, uint32_t n_blocks, cudaStream_t st, int* launches);

}  // namespace ua3
