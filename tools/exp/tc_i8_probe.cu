// tc_i8_probe.cu - EXPERIMENT (tools/exp): the smallest tcgen05 kind::i8 MMA with the A operand written to TMEM by the threads
// themselves (tcgen05.st), B in shared memory (K-major, no swizzle), int32 accumulators read back with tcgen05.ld.
// Purpose: pin down operand layouts and descriptor fields before building the integrator contraction on top of it.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tc_i8_probe tc_i8_probe.cu && ./tc_i8_probe
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>

constexpr int M = 128, N = 16, K = 32;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(128) probe(const int8_t* A, const uint8_t* B, int32_t* D, uint32_t lbo16, uint32_t sbo16, int a_signed,
                                             int second /* issue a second accumulating MMA with the same operands */) {
    __shared__ __align__(128) uint8_t s_b[N * K];
    __shared__ __align__(8) uint64_t s_bar;
    __shared__ uint32_t s_tmem;
    const int t = threadIdx.x, warp = t >> 5;
    // B into the canonical K-major no-swizzle layout: core matrix = 8 rows x 16 bytes = 128 contiguous bytes
    for (int i = t; i < N * K; i += 128) {
        const int n = i / K, k = i % K;
        s_b[(k / 16) * (lbo16 * 16) + (n / 8) * (sbo16 * 16) + (n % 8) * 16 + (k % 16)] = B[i];
    }
    if (t == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&s_bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 64;" ::"r"(smem_u32(&s_tmem)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");          // generic-proxy smem writes -> visible to the MMA (async proxy)
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = s_tmem;
    // A: row t, 32 bytes = 8 words, byte k of the row at bits 8*(k%4) of word k/4
    uint32_t a[8];
    for (int w = 0; w < 8; ++w) {
        uint32_t v = 0;
        for (int b = 0; b < 4; ++b) v |= (uint32_t)(uint8_t)A[t * K + 4 * w + b] << (8 * b);
        a[w] = v;
    }
    const uint32_t lane_base = tmem + ((uint32_t)(warp * 32) << 16);  // TMEM address: lane in bits 31:16, column in 15:0
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(lane_base + 0), "r"(a[0]), "r"(a[1]),
                 "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7])
                 : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (t == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        // shared-memory matrix descriptor: start >> 4, LBO >> 4 at bit 16, SBO >> 4 at bit 32, version 1 at bit 46, no swizzle
        const uint64_t bdesc = (uint64_t)((smem_u32(s_b) >> 4) & 0x3FFF) | ((uint64_t)(lbo16 & 0x3FFF) << 16) | ((uint64_t)(sbo16 & 0x3FFF) << 32) |
                               ((uint64_t)1 << 46);
        // instruction descriptor: D = S32 (2 << 4), A format (1 = s8, 0 = u8) << 7, B format u8 (0) << 10, K-major both, N >> 3 at 17, M >> 4 at 24
        const uint32_t idesc = (2u << 4) | ((uint32_t)(a_signed ? 1 : 0) << 7) | (0u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
        const uint32_t d = tmem + 32, a_t = tmem + 0, zero = 0;
        asm volatile(
            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, {%5, %5, %5, %5}, p;\n\t}" ::"r"(d), "r"(a_t), "l"(bdesc), "r"(idesc), "r"(0u), "r"(zero)
            : "memory");
        if (second)
            asm volatile(
                "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                "tcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, {%5, %5, %5, %5}, p;\n\t}" ::"r"(d), "r"(a_t), "l"(bdesc), "r"(idesc), "r"(1u), "r"(zero)
                : "memory");
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&s_bar)) : "memory");
    }
    // everybody waits for the MMA
    {
        uint32_t done = 0;
        while (!done)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smem_u32(&s_bar)), "r"(0u) : "memory");
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]),
          "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(lane_base + 32));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int n = 0; n < N; ++n) D[t * N + n] = (int32_t)r[n];
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 64;" ::"r"(tmem));
}

int main() {
    int8_t hA[M * K];
    uint8_t hB[N * K];
    for (int m = 0; m < M; ++m) for (int k = 0; k < K; ++k) hA[m * K + k] = (int8_t)(((m * 7 + k * 3) % 251) - 100);
    for (int n = 0; n < N; ++n) for (int k = 0; k < K; ++k) hB[n * K + k] = (uint8_t)((n * 37 + k * 11 + 200) % 256);
    int8_t* dA; uint8_t* dB; int32_t* dD;
    cudaMalloc(&dA, sizeof hA); cudaMalloc(&dB, sizeof hB); cudaMalloc(&dD, M * N * 4);
    cudaMemcpy(dA, hA, sizeof hA, cudaMemcpyHostToDevice); cudaMemcpy(dB, hB, sizeof hB, cudaMemcpyHostToDevice);
    static int32_t hD[M * N];
    const uint32_t combos[2][2] = {{16, 8}, {8, 16}};      // (LBO, SBO) in 16-byte units: [k half][n group] or [n group][k half] placement
    for (int a_signed = 1; a_signed >= 0; --a_signed)
        for (int c = 0; c < 2; ++c)
            for (int second = 0; second < 2; ++second) {
                cudaMemset(dD, 0xFF, M * N * 4);
                probe<<<1, 128>>>(dA, dB, dD, combos[c][0], combos[c][1], a_signed, second);
                cudaError_t e = cudaDeviceSynchronize();
                if (e != cudaSuccess) { printf("a_signed %d lbo %u sbo %u: CUDA error %s\n", a_signed, combos[c][0], combos[c][1], cudaGetErrorString(e)); return 1; }
                cudaMemcpy(hD, dD, sizeof hD, cudaMemcpyDeviceToHost);
                int bad = 0, first = -1;
                for (int m = 0; m < M; ++m) for (int n = 0; n < N; ++n) {
                    long ref = 0;
                    for (int k = 0; k < K; ++k) ref += (long)(a_signed ? (int)hA[m * K + k] : (int)(uint8_t)hA[m * K + k]) * (long)hB[n * K + k];
                    ref *= (second ? 2 : 1);
                    if (hD[m * N + n] != (int32_t)ref) { if (first < 0) first = m * N + n; ++bad; }
                }
                printf("a_signed %d lbo %u sbo %u second %d: %d of %d wrong", a_signed, combos[c][0], combos[c][1], second, bad, M * N);
                if (bad) printf(" (first at m %d n %d: got %d)", first / N, first % N, hD[first]);
                printf("\n");
            }
    return 0;
}
