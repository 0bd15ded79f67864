// ddc_launch.h - host-visible launch interface of ddc.cu
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace ua3 {

struct DdcBuffers {
    uint32_t n_ch = 0, n_ch_pad = 0;
    uint32_t max_chunks = 0, max_frames = 0;
    uint32_t* nco_tab = nullptr;   // [2048] packed coarse ROM
    uint32_t* big_tab = nullptr;   // [2048 * 26] packed (sin12, cos12) by (coarse address, fine-sine value)
    uint32_t* tile_counter = nullptr;   // work counter of the persistent front kernel (zeroed by adc_expand_kernel)
    int32_t* adc9 = nullptr;       // [max_block] the current ADC block widened to int32 and pre-shifted << 9 (adc_expand_kernel)
    int front_variant = 0;         // 0 auto, 1 force the 8 KB-table kernel, 2 force the big-table CUDA-core kernel, 3 force the tensor-core kernel
    // tensor-core front kernel (ddc_front_tc.cuh)
    uint32_t* tab_h = nullptr;     // [2048 * 26] the big table as (sin, cos) binary16 pairs
    uint8_t* tc_w = nullptr;       // [8192] byte planes of the integrator weights C(511 - t, k), MMA B-operand layout
    uint64_t tc_fix[5] = {0, 0, 0, 0, 0};   // per-stage offset removed in the recombination (build_tc_weight_planes)
    uint16_t* adc_h = nullptr;     // [max_block] the current ADC block as binary16 (adc_prepare_tc_kernel)
    uint8_t* wrap_flag = nullptr;  // [max_chunks] chunk contains an ADC sample of -2048
    int pdl = 1;                   // programmatic dependent launch between the chain's kernels (UA3REO_PDL=0 turns it off)
    int tc_adc_stage = 1;          // 1: the chunk's samples are staged in shared memory (cp.async, one tile ahead); 0: broadcast global loads
    uint32_t* fcw = nullptr;       // [n_ch_pad] 22-bit tuning words
    uint32_t* phase = nullptr;     // [n_ch_pad] 22-bit phase at the start of the next block
    uint64_t* L = nullptr;         // [n_ch_pad][kLHalo + max_chunks][2][5]
    uint32_t l_ch_stride = 0;
    int16_t* YI = nullptr;         // [n_ch_pad][kYIHalo + max_frames]
    uint32_t yi_stride = 0;
    int16_t* YQ = nullptr;         // [n_ch_pad][kYQHalo + max_frames]
    uint32_t yq_stride = 0;
    uint64_t* frames = nullptr;    // [n_ch][ring] 8-byte frames, a ring per channel (ring is a power of two)
    uint32_t frame_ch_stride = 0;  // ring size in frames
    uint32_t ring_mask = 0;
    // Clocking class of the frame (which of the board's reset-release instants is reproduced; the classes and their
    // frequencies were measured by running the reference's VHDL, oracle/ddc_golden.c "Clocking"):
    int align_b = 1;               // compensator polyphase alignment: 1 = "B" y[k] = sum h[j] u[2k-j], 0 = "A" y[k] = sum h[j] u[2k+1-j]
    int d_i = 3;                   // VOICE_I[n] = hilbert(SPEC_I)[n - d_i], 0..kMaxDI (serial MAC + output register of rx_hilb.vhd)
    int d_q = 129;                 // VOICE_Q[n] = SPEC_Q[n - d_q], 1..130 (data_delay.v, 130 registers, sampled before or after its clock edge)
};

void build_cic_weights(uint64_t G[25]);
void build_nco_table(uint32_t tab[2048]);
void build_nco_big_table(uint32_t* tab);
void build_nco_half_table(uint32_t* tab);
void build_tc_weight_planes(uint8_t* w, uint64_t fix[5]);
constexpr int kTcWeightPlaneBytes = 8192;
cudaError_t ddc_prepare_kernels();
constexpr int kNcoBigTabWords = 2048 * 26;
cudaError_t ddc_upload_constants();
constexpr int kDdcKernels = 5;   // profile slots in launch order: adc_expand, front, cic+comp, hilb, rotate
// ev: optional array of kDdcKernels + 1 events recorded before/after each kernel (profiling mode); ev_mask selects which
// ring_start: ring index that receives the block's first frame
cudaError_t ddc_launch_block(const DdcBuffers& b, const int16_t* adc_dev, uint32_t n_samples, uint32_t ring_start,
                             int sm_count, cudaStream_t st, int* launches, cudaEvent_t* ev, uint32_t ev_mask = 0xFFFFFFFFu);
cudaError_t measure_int32_peak(int sm_count, cudaStream_t st, double* ops_per_s);
cudaError_t measure_lds_peak(int sm_count, cudaStream_t st, double* wavefronts_per_s);

}  // namespace ua3
