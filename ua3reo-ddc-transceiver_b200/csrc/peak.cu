// peak.cu - INT32 issue-rate microbenchmark: the roofline denominator for the DDC front kernel.
//
// MEASURED_PEAKS.json carries only HBM and bf16 peaks; the DDC is bound by integer issue
// (SURVEY.md 8d).  Each thread runs kChains independent dependency chains; one step of one chain is
// an IMAD (fma pipe) followed by a LOP3 and an IADD3 (alu pipe), so both integer pipes are busy and
// the result is the best integer op rate a kernel of dependent-free 32-bit work can reach.
#include "ddc_launch.h"

namespace ua3 {

constexpr int kChains = 8;
constexpr int kInner = 64;

__global__ void __launch_bounds__(256) int32_peak_kernel(uint32_t a, uint32_t b, uint32_t k, int iters, uint32_t* sink) {
    uint32_t x[kChains], y[kChains];
#pragma unroll
    for (int i = 0; i < kChains; ++i) { x[i] = threadIdx.x + i; y[i] = blockIdx.x ^ i; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < kInner; ++u) {
#pragma unroll
            for (int i = 0; i < kChains; ++i) {
                x[i] = x[i] * a + b;            // IMAD
                y[i] = (y[i] ^ k) + x[(i + 1) % kChains];   // LOP3 + IADD3 (reads the other chain: not foldable)
            }
        }
    }
    uint32_t r = 0;
#pragma unroll
    for (int i = 0; i < kChains; ++i) r += x[i] ^ y[i];
    if (r == 0x12345678u) sink[0] = r;          // practically never true; keeps the chains live
}

cudaError_t measure_int32_peak(int sm_count, cudaStream_t st, double* ops_per_s) {
    uint32_t* sink = nullptr;
    cudaError_t e = cudaMalloc(&sink, 4);
    if (e != cudaSuccess) return e;
    cudaEvent_t t0, t1;
    cudaEventCreate(&t0); cudaEventCreate(&t1);
    const int grid = sm_count * 8, iters = 256;
    int32_peak_kernel<<<grid, 256, 0, st>>>(3u, 7u, 0x9E3779B9u, 8, sink);      // warm-up
    double best = 0.0;
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(t0, st);
        int32_peak_kernel<<<grid, 256, 0, st>>>(3u, 7u, 0x9E3779B9u, iters, sink);
        cudaEventRecord(t1, st);
        e = cudaEventSynchronize(t1);
        if (e != cudaSuccess) break;
        float ms = 0.f;
        cudaEventElapsedTime(&ms, t0, t1);
        const double ops = 3.0 * kChains * kInner * (double)iters * 256.0 * grid;
        const double rate = ops / (ms * 1e-3);
        if (rate > best) best = rate;
    }
    cudaEventDestroy(t0); cudaEventDestroy(t1);
    cudaFree(sink);
    if (e == cudaSuccess) e = cudaGetLastError();
    *ops_per_s = best;
    return e;
}

// Shared-memory (LSU data pipe) wavefront rate: the roofline denominator of the tensor-core front kernel, whose NCO table
// look-ups saturate that pipe.  One LDS.128 per warp = 32 lanes x 16 contiguous bytes = four conflict-free wavefronts
// (tools/exp/lsu_wavefronts.cu), so the loads and not the issue slots are the limit; 1024 threads per SM, addresses that do
// not depend on the data.
constexpr int kLdsInner = 64;
constexpr int kLdsWavefrontsPerLoad = 4;

__global__ void __launch_bounds__(1024) lds_peak_kernel(int iters, uint32_t* sink) {
    __shared__ uint4 s[2048];                                                // 32 KB
    for (int i = threadIdx.x; i < 2048; i += 1024) s[i] = make_uint4(i, i * 3u, i * 5u, i * 7u);
    __syncthreads();
    const uint32_t lane = threadIdx.x & 31u;
    uint32_t row = (threadIdx.x >> 5) * 32u, acc = 0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < kLdsInner; ++u) {
            const uint4 v = s[((row + 32u * u) & 2047u) | lane];             // compile-time offsets off one register
            acc += v.x ^ v.y ^ v.z ^ v.w;
        }
        row = (row + 32u * 7u) & 2047u;
    }
    if (acc == 0x12345678u) sink[0] = acc;
}

cudaError_t measure_lds_peak(int sm_count, cudaStream_t st, double* wavefronts_per_s) {
    uint32_t* sink = nullptr;
    cudaError_t e = cudaMalloc(&sink, 4);
    if (e != cudaSuccess) return e;
    cudaEvent_t t0, t1;
    cudaEventCreate(&t0); cudaEventCreate(&t1);
    const int grid = sm_count * 2, iters = 512;
    lds_peak_kernel<<<grid, 1024, 0, st>>>(8, sink);
    double best = 0.0;
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(t0, st);
        lds_peak_kernel<<<grid, 1024, 0, st>>>(iters, sink);
        cudaEventRecord(t1, st);
        e = cudaEventSynchronize(t1);
        if (e != cudaSuccess) break;
        float ms = 0.f;
        cudaEventElapsedTime(&ms, t0, t1);
        const double wavefronts = (double)kLdsWavefrontsPerLoad * kLdsInner * iters * (1024.0 / 32.0) * grid;
        const double rate = wavefronts / (ms * 1e-3);
        if (rate > best) best = rate;
    }
    cudaEventDestroy(t0); cudaEventDestroy(t1);
    cudaFree(sink);
    if (e == cudaSuccess) e = cudaGetLastError();
    *wavefronts_per_s = best;
    return e;
}

}  // namespace ua3
