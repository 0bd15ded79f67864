// tools/emu/cuda_runtime.h - DEVELOPMENT AID, NOT PRODUCT CODE.
//
// A minimal stand-in for the CUDA runtime that lets csrc/*.cu be compiled by g++ and executed on
// the host so that kernel index logic can be debugged in the build container (which has no GPU).
// Each CTA runs as blockDim.x OS threads with a real barrier for __syncthreads().  It is slow, it
// is never shipped, and the package (ua3reo-ddc-transceiver_b200/__init__.py) never loads it: only
// tools/emu/check_emu.py does, explicitly.
#pragma once
#include <stdint.h>
#include <stddef.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <math.h>
#include <barrier>
#include <thread>
#include <vector>
#include <functional>

#define UA3_HOST_EMU 1
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(...)
#define __align__(n) __attribute__((aligned(n)))
#define __shared__ static
#define __constant__ static

struct dim3 { unsigned x = 1, y = 1, z = 1; dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {} };
struct uint4 { uint32_t x, y, z, w; };
struct ulonglong2 { unsigned long long x, y; };
struct int2 { int x, y; };
static inline int2 make_int2(int a, int b) { return int2{a, b}; }
static inline ulonglong2 make_ulonglong2(unsigned long long a, unsigned long long b) { return ulonglong2{a, b}; }
template <class T> static inline T __ldg(const T* p) { return *p; }
using std::min; using std::max;

extern thread_local dim3 threadIdx, blockIdx, blockDim, gridDim;
extern thread_local std::barrier<>* ua3_emu_barrier;
static inline void __syncthreads() { ua3_emu_barrier->arrive_and_wait(); }

typedef int cudaError_t;
typedef void* cudaStream_t;
typedef void* cudaEvent_t;
enum { cudaSuccess = 0 };
enum cudaMemcpyKind { cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyDefault };
enum { cudaStreamNonBlocking = 1 };
struct cudaDeviceProp { int major = 10, minor = 0, multiProcessorCount = 2; };
static inline const char* cudaGetErrorString(cudaError_t) { return "emu"; }
static inline cudaError_t cudaGetLastError() { return 0; }
static inline cudaError_t cudaGetDeviceCount(int* n) { *n = 1; return 0; }
static inline cudaError_t cudaGetDeviceProperties(cudaDeviceProp* p, int) { *p = cudaDeviceProp(); return 0; }
static inline cudaError_t cudaSetDevice(int) { return 0; }
static inline cudaError_t cudaGetDevice(int* d) { *d = 0; return 0; }
enum { cudaDevAttrMultiProcessorCount = 16 };
static inline cudaError_t cudaDeviceGetAttribute(int* v, int, int) { *v = 2; return 0; }
static inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, int) { *s = nullptr; return 0; }
static inline cudaError_t cudaStreamCreateWithPriority(cudaStream_t* s, int, int) { *s = nullptr; return 0; }
static inline cudaError_t cudaDeviceGetStreamPriorityRange(int* lo, int* hi) { *lo = 0; *hi = 0; return 0; }
static inline cudaError_t cudaStreamDestroy(cudaStream_t) { return 0; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return 0; }
static inline cudaError_t cudaMalloc(void** p, size_t n) { *p = aligned_alloc(256, (n + 255) & ~size_t(255)); return *p ? 0 : 2; }
template <class T> static inline cudaError_t cudaMalloc(T** p, size_t n) { return cudaMalloc((void**)p, n); }
static inline cudaError_t cudaFree(void* p) { free(p); return 0; }
static inline cudaError_t cudaMemsetAsync(void* p, int v, size_t n, cudaStream_t) { memset(p, v, n); return 0; }
static inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t) { memmove(d, s, n); return 0; }
static inline cudaError_t cudaMemcpy2DAsync(void* d, size_t dp, const void* s, size_t sp, size_t w, size_t h, cudaMemcpyKind, cudaStream_t) {
    for (size_t i = 0; i < h; ++i) memcpy((char*)d + i * dp, (const char*)s + i * sp, w);
    return 0;
}
template <class T> static inline cudaError_t cudaMemcpyToSymbol(T& sym, const void* src, size_t n) { memcpy((void*)&sym, src, n); return 0; }
static inline cudaError_t cudaMemcpyPeerAsync(void* d, int, const void* s, int, size_t n, cudaStream_t) { memmove(d, s, n); return 0; }
static inline cudaError_t cudaDeviceCanAccessPeer(int* can, int, int) { *can = 0; return 0; }
static inline cudaError_t cudaDeviceEnablePeerAccess(int, unsigned) { return 0; }
enum { cudaErrorPeerAccessAlreadyEnabled = 704 };
static inline cudaError_t cudaEventCreate(cudaEvent_t*) { return 0; }
enum { cudaEventDisableTiming = 2 };
static inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t*, int) { return 0; }
static inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, int) { return 0; }
static inline cudaError_t cudaEventDestroy(cudaEvent_t) { return 0; }
static inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t) { return 0; }
static inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return 0; }
static inline cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t, cudaEvent_t) { *ms = 1.f; return 0; }

template <class K, class... A>
static void ua3_emu_launch(K kernel, dim3 grid, dim3 block, A... args) {
    for (unsigned bz = 0; bz < grid.z; ++bz)
    for (unsigned by = 0; by < grid.y; ++by)
    for (unsigned bx = 0; bx < grid.x; ++bx) {
        std::barrier<> bar((std::ptrdiff_t)block.x);
        std::vector<std::thread> th;
        for (unsigned t = 0; t < block.x; ++t)
            th.emplace_back([&, t]() {
                threadIdx = dim3(t); blockIdx = dim3(bx, by, bz); blockDim = block; gridDim = grid;
                ua3_emu_barrier = &bar;
                kernel(args...);
                bar.arrive_and_drop();
            });
        for (auto& x : th) x.join();
    }
}
#define UA3_LAUNCH(kernel, grid, block, smem, stream, ...) ua3_emu_launch(kernel, dim3(grid), dim3(block), __VA_ARGS__)

// ---- warp primitives for single-warp CTAs (every lane of the CTA takes part; no divergence around them) ----
extern float ua3_emu_shfl_slots[1024];
static inline float ua3_emu_shfl(float v, int src_lane) {
    ua3_emu_shfl_slots[threadIdx.x] = v;
    ua3_emu_barrier->arrive_and_wait();
    const float r = ua3_emu_shfl_slots[(threadIdx.x & ~31u) | (unsigned)(src_lane & 31)];
    ua3_emu_barrier->arrive_and_wait();
    return r;
}
static inline float __shfl_xor_sync(unsigned, float v, int m) { return ua3_emu_shfl(v, (int)(threadIdx.x & 31) ^ m); }
static inline float __shfl_sync(unsigned, float v, int l) { return ua3_emu_shfl(v, l); }
static inline uint32_t __shfl_xor_sync(unsigned, uint32_t v, int m) {
    float f; memcpy(&f, &v, 4); f = ua3_emu_shfl(f, (int)(threadIdx.x & 31) ^ m); memcpy(&v, &f, 4); return v;
}
static inline int32_t __shfl_xor_sync(unsigned, int32_t v, int m) {
    float f; memcpy(&f, &v, 4); f = ua3_emu_shfl(f, (int)(threadIdx.x & 31) ^ m); memcpy(&v, &f, 4); return v;
}
extern uint32_t ua3_emu_pred_slots[1024];
static inline uint32_t __ballot_sync(unsigned, int pred) {
    ua3_emu_pred_slots[threadIdx.x] = pred ? 1u : 0u;
    ua3_emu_barrier->arrive_and_wait();
    uint32_t w = 0;
    const unsigned base = threadIdx.x & ~31u;
    for (unsigned l = 0; l < 32; ++l) w |= ua3_emu_pred_slots[base + l] << l;
    ua3_emu_barrier->arrive_and_wait();
    return w;
}
static inline int __popc(uint32_t v) { return __builtin_popcount(v); }
static inline uint32_t __funnelshift_r(uint32_t lo, uint32_t hi, uint32_t sh) { return (uint32_t)((((uint64_t)hi << 32) | lo) >> (sh & 31)); }
static inline void __syncwarp() { ua3_emu_barrier->arrive_and_wait(); }
enum { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
template <class F> static inline cudaError_t cudaFuncSetAttribute(F, int, int) { return 0; }
static inline cudaError_t cudaMemset2DAsync(void* p, size_t pitch, int v, size_t w, size_t h, cudaStream_t) {
    for (size_t i = 0; i < h; ++i) memset((char*)p + i * pitch, v, w);
    return 0;
}
#include <mutex>
static std::mutex ua3_emu_atomic_mutex;
static inline int atomicMin(int* p, int v) { std::lock_guard<std::mutex> g(ua3_emu_atomic_mutex); int o = *p; if (v < o) *p = v; return o; }
static inline int atomicMax(int* p, int v) { std::lock_guard<std::mutex> g(ua3_emu_atomic_mutex); int o = *p; if (v > o) *p = v; return o; }
static inline uint32_t atomicAdd(uint32_t* p, uint32_t v) { std::lock_guard<std::mutex> g(ua3_emu_atomic_mutex); uint32_t o = *p; *p = o + v; return o; }
