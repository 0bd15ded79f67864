// int_rates.cu - DEVELOPMENT micro-benchmark: issue rates of the integer instructions the DDC front kernel uses.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o int_rates int_rates.cu && ./int_rates
#include <cstdio>
#include <cuda_runtime.h>
#include <cstdint>

constexpr int kChains = 8, kInner = 128;

template <int OP>
__global__ void __launch_bounds__(256) k(uint32_t a, uint32_t b, int iters, uint32_t* sink) {
    uint32_t x[kChains];
    unsigned long long w[kChains];
#pragma unroll
    for (int i = 0; i < kChains; ++i) { x[i] = threadIdx.x * 7 + i; w[i] = x[i]; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < kInner; ++u) {
#pragma unroll
            for (int i = 0; i < kChains; ++i) {
                if (OP == 0) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(a), "r"(b));
                if (OP == 1) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"(x[i]), "r"(a));
                if (OP == 2) asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(a), "r"(b));
                if (OP == 3) asm volatile("add.u32 %0, %0, %1;" : "+r"(x[i]) : "r"(a));
                if (OP == 4) asm volatile("shf.r.wrap.b32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(a), "r"(b));
                if (OP == 5) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[i]) : "r"(a), "r"(b));
                if (OP == 6) asm volatile("prmt.b32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(a), "r"(b));
                if (OP == 7) asm volatile("add.u64 %0, %0, %1;" : "+l"(w[i]) : "l"((unsigned long long)a << 20 | b));
                if (OP == 8) asm volatile("shr.s64 %0, %0, 8;" : "+l"(w[i]));
                if (OP == 9) asm volatile("mul.lo.u64 %0, %0, %1;" : "+l"(w[i]) : "l"((unsigned long long)a << 33 | b));
            }
        }
    }
    uint32_t r = 0;
#pragma unroll
    for (int i = 0; i < kChains; ++i) r += x[i] ^ (uint32_t)w[i] ^ (uint32_t)(w[i] >> 32);
    if (r == 0x12345678u) sink[0] = r;
}

template <int OP>
double run(const char* name, int sms) {
    uint32_t* sink; cudaMalloc(&sink, 4);
    cudaEvent_t t0, t1; cudaEventCreate(&t0); cudaEventCreate(&t1);
    const int grid = sms * 8, iters = 64;
    k<OP><<<grid, 256>>>(3, 7, 4, sink);
    double best = 0;
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(t0); k<OP><<<grid, 256>>>(3, 7, iters, sink); cudaEventRecord(t1); cudaEventSynchronize(t1);
        float ms; cudaEventElapsedTime(&ms, t0, t1);
        const double rate = (double)kChains * kInner * iters * 256.0 * grid / (ms * 1e-3);
        if (rate > best) best = rate;
    }
    printf("%-22s %8.2f Tinstr/s  = %6.1f lanes/clk/SM @1.965GHz\n", name, best / 1e12, best / (sms * 1.965e9));
    cudaFree(sink);
    return best;
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    const int sms = p.multiProcessorCount;
    run<0>("mad.lo.u32 (IMAD)", sms);
    run<1>("mad.wide.u32", sms);
    run<2>("mad.hi.u32", sms);
    run<3>("add.u32", sms);
    run<4>("shf.r.wrap", sms);
    run<5>("lop3", sms);
    run<6>("prmt", sms);
    run<7>("add.u64", sms);
    run<8>("shr.s64 8", sms);
    run<9>("mul.lo.u64", sms);
    return 0;
}
