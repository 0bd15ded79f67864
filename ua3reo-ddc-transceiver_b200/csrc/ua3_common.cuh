// ua3_common.cuh - shared definitions for the UA3REO B200 receive-path kernels.
//
// The per-thread bodies of the integer kernels are written as UA3_HD functions that take their
// thread/block coordinates as arguments.  The __global__ wrappers in *.cu pass threadIdx/blockIdx;
// tools/emu/ compiles the same bodies with g++ (UA3_HOST_EMU) as a development aid to debug index
// logic without a GPU.  The emulation build is never loaded by the product.
#pragma once
#include <stdint.h>
#include <stddef.h>

#if defined(__CUDACC__) && !defined(UA3_HOST_EMU)
#define UA3_HD __host__ __device__ __forceinline__
#define UA3_D __device__ __forceinline__
#else
#define UA3_HD inline
#define UA3_D inline
#endif

#if !defined(UA3_HOST_EMU)
#define UA3_LAUNCH(kernel, grid, block, smem, stream, ...) kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__)
#endif
#if defined(__CUDACC__) && !defined(UA3_HOST_EMU)
#include <cuda_runtime.h>
#include <utility>
namespace ua3 {
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(bool pdl, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = pdl ? 1u : 0u;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(std::forward<Args>(args))...);
}
}  // namespace ua3
#define UA3_LAUNCH_PDL(pdl, kernel, grid, block, smem, stream, ...) (void)ua3::launch_pdl((pdl), kernel, dim3(grid), dim3(block), (smem), (stream), __VA_ARGS__)
#else
#define UA3_LAUNCH_PDL(pdl, kernel, grid, block, smem, stream, ...) UA3_LAUNCH(kernel, grid, block, smem, stream, __VA_ARGS__)
#endif

// Programmatic dependent launch: a kernel launched with the attribute may be scheduled while its predecessor in the stream
// still runs; it must call pdl_wait() before it touches anything the predecessor (or, transitively, any earlier kernel - every
// kernel of the chain waits) produced.  pdl_trigger() lets the NEXT kernel's CTAs take their places early.  Both are no-ops
// in a kernel launched the ordinary way.
#if defined(__CUDA_ARCH__)
#define UA3_PDL_WAIT() asm volatile("griddepcontrol.wait;" ::: "memory")
#define UA3_PDL_TRIGGER() asm volatile("griddepcontrol.launch_dependents;" ::: "memory")
#else
#define UA3_PDL_WAIT() ((void)0)
#define UA3_PDL_TRIGGER() ((void)0)
#endif

namespace ua3 {

// ---- geometry of the FPGA chain (SURVEY.md 0 / rx_cic.vhd:27-29, rx_ciccomp.vhd:25-47) ----
constexpr int kCicR = 512;          // ADC samples per CIC output (rx_cic.vhd:27)
constexpr int kFrameAdc = 1024;     // ADC samples per 48 kHz frame (CIC 512 x compensator 2)
constexpr int kCompTaps = 65;       // rx_ciccomp.vhd:47
constexpr int kHilbTaps = 256;      // rx_hilb.vhd
constexpr int kQDelay = 130;        // UA3REO.bdf:1693-1694
constexpr int kFrameBytes = 8;      // stm32_interface.v:228-271

// ---- layout of the per-chunk CIC partial-state records in HBM ----
// One record per (channel, 512-sample chunk): 2 rails x 5 integrator partial states (u64).
constexpr int kLHalo = 69;          // chunks of history: 65 CIC outputs for the 65-tap compensator in either polyphase alignment + 4 for the 5-chunk comb window
constexpr int kLRec = 10;           // u64 per record: [rail][stage]
constexpr int kUHalo = 65;          // 96 kHz samples of history for the 65-tap compensator (64 in alignment A, 65 in alignment B)
constexpr int kMaxDI = 3;           // largest VOICE_I latency of a clocking class (ddc_launch.h)
constexpr int kYIHalo = 255 + kMaxDI;   // 48 kHz samples of history for the 256-tap Hilbert FIR evaluated up to kMaxDI samples back
constexpr int kYQHalo = 130;        // data_delay depth (d_q of a clocking class is 129 or 130)

// ---- front kernel tiling ----
constexpr int kFrontWarps = 8;                     // warps per CTA, one 512-sample chunk each
constexpr int kFrontThreads = kFrontWarps * 32;
constexpr int kSub = 16;                           // 32-bit sub-block length (see ddc_front.cuh)
// big-table front kernel: one persistent CTA per SM, 32 warps = 8 channel tiles x 4 chunks per tile
constexpr int kBtCG = 8;                           // channel tiles (of 32 channels) per CTA tile
#ifndef UA3_BT_TG
#define UA3_BT_TG 4
#endif
constexpr int kBtTG = UA3_BT_TG;                   // 512-sample chunks per CTA tile
constexpr int kBtWarps = kBtCG * kBtTG;
constexpr int kBtThreads = kBtWarps * 32;

struct alignas(16) I4 { int32_t x, y, z, w; };

}  // namespace ua3
