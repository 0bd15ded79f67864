// DEVELOPMENT AID (see cuda_runtime.h in this directory): thread-local CUDA built-ins + stubs.
#include "cuda_runtime.h"
thread_local dim3 threadIdx, blockIdx, blockDim, gridDim;
thread_local std::barrier<>* ua3_emu_barrier = nullptr;
namespace ua3 { cudaError_t measure_int32_peak(int, cudaStream_t, double* r) { *r = 0.0; return 0; } }
float ua3_emu_shfl_slots[1024];
uint32_t ua3_emu_pred_slots[1024];
