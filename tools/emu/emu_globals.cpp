// DEVELOPMENT AID (see cuda_runtime.h in this directory): thread-local CUDA built-ins + stubs.
#include "cuda_runtime.h"
thread_local dim3 threadIdx, blockIdx, blockDim, gridDim;
thread_local std::barrier<>* ua3_emu_barrier = nullptr;
namespace ua3 {
cudaError_t measure_int32_peak(int, cudaStream_t, double* r) { *r = 0.0; return 0; }
cudaError_t measure_lds_peak(int, cudaStream_t, double* r) { *r = 0.0; return 0; }
}
// csrc/fanout.cu (CUDA IPC between processes) has nothing to emulate: its entry points exist so that the binding loads, and fail
extern "C" {
#define UA3_EMU_STUB(name) int name(...) { return -5; }
UA3_EMU_STUB(ua3reo_fanout_create) UA3_EMU_STUB(ua3reo_fanout_destroy) UA3_EMU_STUB(ua3reo_fanout_disconnect)
UA3_EMU_STUB(ua3reo_fanout_handle) UA3_EMU_STUB(ua3reo_fanout_connect) UA3_EMU_STUB(ua3reo_fanout_send)
UA3_EMU_STUB(ua3reo_fanout_acquire) UA3_EMU_STUB(ua3reo_fanout_release) UA3_EMU_STUB(ua3reo_fanout_sync)
UA3_EMU_STUB(ua3reo_fanout_info)
UA3_EMU_STUB(ua3reo_gather_create) UA3_EMU_STUB(ua3reo_gather_destroy) UA3_EMU_STUB(ua3reo_gather_disconnect)
UA3_EMU_STUB(ua3reo_gather_handle) UA3_EMU_STUB(ua3reo_gather_connect) UA3_EMU_STUB(ua3reo_gather_send)
UA3_EMU_STUB(ua3reo_gather_acquire) UA3_EMU_STUB(ua3reo_gather_release)
}
float ua3_emu_shfl_slots[1024];
uint32_t ua3_emu_pred_slots[1024];
