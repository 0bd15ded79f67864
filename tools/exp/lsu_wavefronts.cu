// lsu_wavefronts.cu - how many LSU data-pipe wavefronts does a warp-uniform (broadcast) load cost on sm_100a?
// Run under ncu with --metrics l1tex__data_pipe_lsu_wavefronts.sum,smsp__inst_executed.sum ; each kernel issues
// 1024 loads per warp, 4 warps per CTA, 148 CTAs.  (Experiment behind ddc_front_tc.cuh's choice of ADC delivery.)
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
constexpr int kIter = 1024;
template <int MODE>
__global__ void __launch_bounds__(128) probe(const uint4* __restrict__ g, uint32_t* out) {
    __shared__ uint4 s[kIter];
    for (int i = threadIdx.x; i < kIter; i += 128) s[i] = g[i];
    __syncthreads();
    uint32_t acc = 0;
    const int lane = threadIdx.x & 31;
#pragma unroll 8
    for (int i = 0; i < kIter; ++i) {
        if (MODE == 0) { const uint4 v = __ldg(g + i); acc += v.x ^ v.y ^ v.z ^ v.w; }                      // LDG.128, all lanes one address
        if (MODE == 1) { const uint4 v = s[i]; acc += v.x ^ v.y ^ v.z ^ v.w; }                              // LDS.128 broadcast
        if (MODE == 2) { const uint2 v = reinterpret_cast<const uint2*>(s)[i]; acc += v.x ^ v.y; }          // LDS.64 broadcast
        if (MODE == 3) { acc += reinterpret_cast<const uint32_t*>(s)[i]; }                                  // LDS.32 broadcast
        if (MODE == 4) { uint4 v = make_uint4(0, 0, 0, 0); if (lane == 0) v = __ldg(g + i); acc += v.x ^ v.y ^ v.z ^ v.w; }   // one lane loads
        if (MODE == 5) { acc += __shfl_sync(0xFFFFFFFFu, acc + i, i & 31); }                                // SHFL
        if (MODE == 6) { const uint4 v = __ldg(g + i * 32 + lane); acc += v.x ^ v.y ^ v.z ^ v.w; }          // LDG.128 coalesced (512 B per warp)
        if (MODE == 7) { acc += reinterpret_cast<const uint32_t*>(s)[(i * 33 + lane * 17) & (kIter * 4 - 1)]; }   // LDS.32, distinct banks (17 odd)
        if (MODE == 8) { acc += __ldg(reinterpret_cast<const uint32_t*>(g) + i); }                          // LDG.32 broadcast
    }
    out[blockIdx.x * 128 + threadIdx.x] = acc;
}
int main() {
    uint4* g; uint32_t* out;
    cudaMalloc(&g, sizeof(uint4) * kIter * 32); cudaMemset(g, 1, sizeof(uint4) * kIter * 32); cudaMalloc(&out, 148 * 128 * 4);
    probe<0><<<148, 128>>>(g, out); probe<1><<<148, 128>>>(g, out); probe<2><<<148, 128>>>(g, out); probe<3><<<148, 128>>>(g, out);
    probe<4><<<148, 128>>>(g, out); probe<5><<<148, 128>>>(g, out); probe<6><<<148, 128>>>(g, out); probe<7><<<148, 128>>>(g, out);
    probe<8><<<148, 128>>>(g, out);
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
