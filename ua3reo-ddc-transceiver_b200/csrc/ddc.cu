// ddc.cu - sm_100a kernels for the batched UA3REO receive DDC (one shared ADC stream, many channels).
//
// Launch sequence per ADC block of n = 512*M samples (M even), all on one stream:
//   adc_expand_kernel  : int16 ADC block -> int32 pre-shifted << 9 (once per block); zeroes the front kernel's tile counter
//   ddc_front_bt_kernel: persistent, one 1024-thread CTA per SM, 208 KB NCO table in shared memory; tiles of 256 channels x
//                        4 chunks handed out from a global counter, pre-shifted ADC tiles double-buffered by TMA, no CTA
//                        barrier in the loop; one warp per chunk, lane = channel; writes chunk partial states L
//                        (ddc_front_kernel: 8 KB-table variant for banks below 256 channels)
//   ddc_ciccomp_kernel : L records by one TMA bulk copy -> run-based comb recurrence -> 96 kHz CIC outputs (shared memory
//                        only) -> 65-tap compensator, two frames per thread -> 48 kHz outputs YI/YQ (int16)
//   ddc_hilb_kernel    : YI (256-tap antisymmetric window, 4 frames per thread), YQ (130 delay) -> 8-byte frames
//   ddc_rotate_kernel  : move the tails of L/YI/YQ into their halos, advance the NCO phases
#include "ddc_launch.h"
#include "ddc_front.cuh"
#include "ddc_back.cuh"
#if !defined(UA3_HOST_EMU)
#include <cuda_fp16.h>
#endif
#include "tables_ddc.inc"

namespace ua3 {

__constant__ int16_t c_comp_h[kCompTaps];
__constant__ int32_t c_hilb_c32[kHilbTaps];       // Hilbert coefficients widened to 32 bit (direct constant-bank IMAD operand)
__constant__ uint32_t c_hilb_b0[8], c_hilb_b1[8];    // bit 0 / bit 1 of the coefficients, reversed (see ddc_hilb_kernel)
__constant__ uint64_t c_cic_g[25];

// ------------------------------------------------------------------------------------------------
// front: grid-stride over tiles (channel tile of 32) x (time tile of kFrontWarps chunks)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kFrontThreads, 3)
ddc_front_kernel(const int16_t* __restrict__ adc, uint32_t n_chunks, const uint32_t* __restrict__ nco_tab,
                 const uint32_t* __restrict__ fcw, const uint32_t* __restrict__ phase, uint32_t n_ch_pad,
                 uint64_t* __restrict__ L, uint32_t l_ch_stride /* u64 per channel */) {
    __shared__ uint32_t s_tab[2048];
    __shared__ I4 s_adc[kFrontWarps * kCicR / 4];

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < 2048; i += kFrontThreads) s_tab[i] = nco_tab[i];

    const uint32_t n_ttiles = (n_chunks + kFrontWarps - 1) / kFrontWarps;
    const uint32_t n_ctiles = n_ch_pad >> 5;
    const uint32_t n_tiles = n_ttiles * n_ctiles;

    for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        // channel tile varies fastest so that concurrently running CTAs share the same ADC tile in L2
        const uint32_t ct = tile % n_ctiles, tt = tile / n_ctiles;
        const uint32_t chunk0 = tt * kFrontWarps;
        __syncthreads();   // previous tile fully consumed (also orders the table fill on first pass)
        // stage: 8 int16 per thread per step (16-byte coalesced loads), pre-shift by 9
        {
            const uint32_t n_valid = min((uint32_t)kFrontWarps, n_chunks - chunk0) * kCicR;  // samples
            const uint4* src = reinterpret_cast<const uint4*>(adc + (size_t)chunk0 * kCicR);
            for (uint32_t v = tid; v < n_valid / 8; v += kFrontThreads) {
                const uint4 q = __ldg(src + v);
                const uint32_t ws[4] = {q.x, q.y, q.z, q.w};
                I4 lo, hi;
                int32_t e[8];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    e[2 * k] = (int32_t)(int16_t)(ws[k] & 0xFFFFu) << 9;
                    e[2 * k + 1] = ((int32_t)ws[k] >> 16) << 9;
                }
                lo.x = e[0]; lo.y = e[1]; lo.z = e[2]; lo.w = e[3];
                hi.x = e[4]; hi.y = e[5]; hi.z = e[6]; hi.w = e[7];
                s_adc[2 * v] = lo;
                s_adc[2 * v + 1] = hi;
            }
        }
        __syncthreads();
        const uint32_t chunk = chunk0 + warp;
        if (chunk < n_chunks) {
            const uint32_t ch = (ct << 5) + lane;
            const uint32_t F = fcw[ch] << 10;
            const uint32_t P0 = (phase[ch] << 10) + F * (chunk * (uint32_t)kCicR);
            uint64_t out[10];
            front_chunk(s_tab, s_adc + warp * (kCicR / 4), P0, F, out);
            uint64_t* dst = L + (size_t)ch * l_ch_stride + (size_t)(kLHalo + chunk) * kLRec;
            ulonglong2* d2 = reinterpret_cast<ulonglong2*>(dst);
#pragma unroll
            for (int k = 0; k < 5; ++k) d2[k] = make_ulonglong2(out[2 * k], out[2 * k + 1]);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// front, big-table variant: one persistent CTA per SM (208 KB NCO table in shared memory, loaded once per
// launch with TMA bulk copies), 32 warps = 8 channel tiles x 4 chunks; the next tile's raw ADC samples are
// fetched by a TMA bulk copy while the current tile computes.
// ------------------------------------------------------------------------------------------------
// big table + two pre-shifted int32 ADC tiles (double buffer) + 3 mbarriers + 2 release counters
constexpr size_t kBtAdcTileBytes = (size_t)kBtTG * kCicR * 4;
constexpr size_t kBtSmemBytes = (size_t)kBigTabWords * 4 + 2 * kBtAdcTileBytes + 64;   // 3 barriers (24 B) + 2 counters + 2 tile slots

#if !defined(UA3_HOST_EMU)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n .reg .pred p;\n WAIT_%=:\n mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n @p bra DONE_%=;\n bra WAIT_%=;\n DONE_%=:\n}"
        ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
#endif

// int16 ADC block -> int32, pre-shifted left by 9 (see nco_mix): done once per block so that the persistent front
// kernel can pull ready-to-use tiles with TMA and never touches the raw samples.
__global__ void __launch_bounds__(256)
adc_expand_kernel(const int16_t* __restrict__ adc, uint32_t n8, int32_t* __restrict__ adc9, uint32_t* __restrict__ tile_counter) {
    const uint32_t v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v == 0) *tile_counter = 0;                 // the front kernel that follows hands out tiles from this counter
    if (v >= n8) return;
    const uint4 q = __ldg(reinterpret_cast<const uint4*>(adc) + v);
    const uint32_t ws[4] = {q.x, q.y, q.z, q.w};
    I4 lo, hi;
    lo.x = (int32_t)(int16_t)(ws[0] & 0xFFFFu) << 9; lo.y = ((int32_t)ws[0] >> 16) << 9;
    lo.z = (int32_t)(int16_t)(ws[1] & 0xFFFFu) << 9; lo.w = ((int32_t)ws[1] >> 16) << 9;
    hi.x = (int32_t)(int16_t)(ws[2] & 0xFFFFu) << 9; hi.y = ((int32_t)ws[2] >> 16) << 9;
    hi.z = (int32_t)(int16_t)(ws[3] & 0xFFFFu) << 9; hi.w = ((int32_t)ws[3] >> 16) << 9;
    I4* dst = reinterpret_cast<I4*>(adc9);
    dst[2 * v] = lo;
    dst[2 * v + 1] = hi;
}

#if !defined(UA3_HOST_EMU)
// int16 ADC block -> binary16 (12-bit integers are exact) for the tensor-core front kernel, plus one flag per 512-sample
// chunk: "contains -2048", the only sample value whose product with the NCO can wrap the 23-bit mixer register.
__global__ void __launch_bounds__(256)
adc_prepare_tc_kernel(const int16_t* __restrict__ adc, uint32_t n8, uint16_t* __restrict__ adc_h, uint8_t* __restrict__ wrap_flag,
                      uint32_t* __restrict__ tile_counter) {
    __shared__ uint32_t s_any[8];
    UA3_PDL_WAIT();                                 // adc_h / wrap_flag / the tile counter are still read by the previous push's front kernel
    UA3_PDL_TRIGGER();                              // the front kernel's CTAs may take their SMs and load the NCO table meanwhile
    const uint32_t v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v == 0) *tile_counter = 0;
    bool hit = false;
    if (v < n8) {
        const uint4 q = __ldg(reinterpret_cast<const uint4*>(adc) + v);
        const uint32_t ws[4] = {q.x, q.y, q.z, q.w};
        uint32_t o[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const short lo = (short)(ws[k] & 0xFFFFu), hi = (short)(ws[k] >> 16);
            hit |= (lo == -2048) | (hi == -2048);
            o[k] = (uint32_t)__half_as_ushort(__short2half_rn(lo)) | ((uint32_t)__half_as_ushort(__short2half_rn(hi)) << 16);
        }
        reinterpret_cast<uint4*>(adc_h)[v] = make_uint4(o[0], o[1], o[2], o[3]);
    }
    const uint32_t any = __ballot_sync(0xFFFFFFFFu, hit);
    if ((threadIdx.x & 31) == 0) s_any[threadIdx.x >> 5] = any;
    __syncthreads();
    if (threadIdx.x < 4) {                                   // 64 threads = 512 samples = one chunk
        const uint32_t chunk = blockIdx.x * 4u + threadIdx.x;
        if (chunk * 64u < n8) wrap_flag[chunk] = (s_any[2 * threadIdx.x] | s_any[2 * threadIdx.x + 1]) ? 1 : 0;
    }
}
#endif

}  // namespace ua3
#include "ddc_front_tc.cuh"
namespace ua3 {

__global__ void __launch_bounds__(kBtThreads, 1)
ddc_front_bt_kernel(const int16_t* __restrict__ adc, const int32_t* __restrict__ adc9, uint32_t n_chunks,
                    const uint32_t* __restrict__ big_tab, const uint32_t* __restrict__ fcw,
                    const uint32_t* __restrict__ phase, uint32_t n_ch_pad, uint64_t* __restrict__ L, uint32_t l_ch_stride,
                    uint32_t* __restrict__ tile_counter) {
#if defined(UA3_HOST_EMU)
    static uint8_t s_dyn[kBtSmemBytes] __attribute__((aligned(128)));
#else
    extern __shared__ __align__(128) uint8_t s_dyn[];
#endif
    uint32_t* s_bt = reinterpret_cast<uint32_t*>(s_dyn);                                    // 53248 words
    I4* s_adc = reinterpret_cast<I4*>(s_dyn + (size_t)kBigTabWords * 4);                    // 2 x kBtTG x 512 int32
    uint64_t* s_bar = reinterpret_cast<uint64_t*>(s_dyn + (size_t)kBigTabWords * 4 + 2 * kBtAdcTileBytes);
    uint32_t* s_done = reinterpret_cast<uint32_t*>(s_bar + 3);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t n_ctiles = n_ch_pad >> 5;
    const uint32_t n_cg = (n_ctiles + kBtCG - 1) / kBtCG, n_tg = (n_chunks + kBtTG - 1) / kBtTG;
    const uint32_t n_tiles = n_cg * n_tg;
    const uint32_t wc = (uint32_t)warp % kBtCG, wt = (uint32_t)warp / kBtCG;

#if !defined(UA3_HOST_EMU)
    // bar[0]: table; bar[1], bar[2]: ADC tile buffers 0 / 1 ("full").  A buffer is released by counting the warps
    // that are done with it; the LAST warp to finish takes the next tile from a global counter and refills the
    // buffer for the step two ahead, so no warp ever waits for the others at a CTA-wide barrier - fast warps run up to
    // two tiles ahead of slow ones - and a CTA that starts late or shares its SM (an NCCL kernel, another stream)
    // simply takes fewer tiles.
    uint32_t* s_tile = s_done + 2;                              // tile index held by each buffer
    constexpr uint32_t kNoTile = 0xFFFFFFFFu;
    if (tid == 0) {
        mbar_init(&s_bar[0], 1);
        mbar_init(&s_bar[1], 1);
        mbar_init(&s_bar[2], 1);
        s_done[0] = s_done[1] = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    auto fill = [&](uint32_t buf) {                            // one thread: claim a tile and fetch its pre-shifted samples
        const uint32_t tile = atomicAdd(tile_counter, 1u);
        if (tile < n_tiles) {
            const uint32_t chunk0 = (tile / n_cg) * kBtTG;
            const uint32_t bytes = min((uint32_t)kBtTG, n_chunks - chunk0) * kCicR * 4u;
            s_tile[buf] = tile;
            mbar_expect_tx(&s_bar[1 + buf], bytes);
            tma_bulk_g2s(reinterpret_cast<uint8_t*>(s_adc) + buf * kBtAdcTileBytes, adc9 + (size_t)chunk0 * kCicR, bytes, &s_bar[1 + buf]);
        } else {
            s_tile[buf] = kNoTile;                              // out of work: complete the phase with a plain arrive
            asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&s_bar[1 + buf])) : "memory");
        }
    };
    if (tid == 0) {
        mbar_expect_tx(&s_bar[0], (uint32_t)kBigTabWords * 4u);
        for (uint32_t off = 0; off < (uint32_t)kBigTabWords * 4u; off += 16384u)
            tma_bulk_g2s(s_dyn + off, reinterpret_cast<const uint8_t*>(big_tab) + off, 16384u, &s_bar[0]);
        fill(0);
        fill(1);
    }
    mbar_wait(&s_bar[0], 0);

    for (uint32_t it = 0;; ++it) {
        const uint32_t buf = it & 1u;
        mbar_wait(&s_bar[1 + buf], (it >> 1) & 1u);
        const uint32_t tile = s_tile[buf];
        if (tile == kNoTile) break;
        const uint32_t cg = tile % n_cg, tg = tile / n_cg;     // channel group fastest: neighbours share the ADC tile in L2
        const uint32_t chunk0 = tg * kBtTG;
        const uint32_t ctile = cg * kBtCG + wc, chunk = chunk0 + wt;
        if (ctile < n_ctiles && chunk < n_chunks) {
            const uint32_t ch = (ctile << 5) + lane;
            const uint32_t F = fcw[ch] << 10;
            const uint32_t P0 = (phase[ch] << 10) + F * (chunk * (uint32_t)kCicR);
            uint64_t out[10];
            front_chunk_bt(s_bt, s_adc + buf * (kBtTG * kCicR / 4) + wt * (kCicR / 4), P0, F, out);
            uint64_t* dst = L + (size_t)ch * l_ch_stride + (size_t)(kLHalo + chunk) * kLRec;
            ulonglong2* d2 = reinterpret_cast<ulonglong2*>(dst);
#pragma unroll
            for (int k = 0; k < 5; ++k) d2[k] = make_ulonglong2(out[2 * k], out[2 * k + 1]);
        }
        __syncwarp();
        if (lane == 0) {
            __threadfence_block();                                         // this warp's reads of the buffer (and of s_tile) are done
            if (atomicAdd(&s_done[buf], 1u) == (uint32_t)kBtWarps - 1u) {  // last warp out refills it
                atomicExch(&s_done[buf], 0u);
                __threadfence_block();
                fill(buf);
            }
        }
    }
#else
    // emulation: static tile order, plain staging of the raw samples behind CTA-wide barriers
    for (int i = tid; i < kBigTabWords; i += kBtThreads) s_bt[i] = big_tab[i];
    __syncthreads();
    for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const uint32_t cg = tile % n_cg, tg = tile / n_cg;
        const uint32_t chunk0 = tg * kBtTG;
        const uint32_t n_valid = min((uint32_t)kBtTG, n_chunks - chunk0) * kCicR;
        __syncthreads();
        int32_t* stage = reinterpret_cast<int32_t*>(s_adc);
        for (uint32_t v = tid; v < n_valid; v += kBtThreads) stage[v] = (int32_t)adc[(size_t)chunk0 * kCicR + v] << 9;
        __syncthreads();
        const uint32_t ctile = cg * kBtCG + wc, chunk = chunk0 + wt;
        if (ctile < n_ctiles && chunk < n_chunks) {
            const uint32_t ch = (ctile << 5) + lane;
            const uint32_t F = fcw[ch] << 10;
            const uint32_t P0 = (phase[ch] << 10) + F * (chunk * (uint32_t)kCicR);
            uint64_t out[10];
            front_chunk_bt(s_bt, s_adc + wt * (kCicR / 4), P0, F, out);
            uint64_t* dst = L + (size_t)ch * l_ch_stride + (size_t)(kLHalo + chunk) * kLRec;
            for (int k = 0; k < 10; ++k) dst[k] = out[k];
        }
    }
    (void)adc9; (void)tile_counter; (void)s_bar; (void)s_done;
#endif
}

// ------------------------------------------------------------------------------------------------
// cic combs + compensator FIR, fused: CTA = (channel, tile of 128 frames).
//   stage  : the tile's chunk records (256 chunks + 69 halo, both rails, 80 B each = 26 KB) arrive in shared memory
//            with ONE TMA bulk copy, in the order the front kernel wrote them (no transposition, no staging code).
//   combs  : run-based.  With S_m = A512 S_{m-1} + L_m and z_m = (S_m)[stage 5] the five combs are the 5th backward
//            difference of z, which annihilates the (degree <= 4 polynomial) response to whatever state preceded a run.
//            A thread therefore starts from S = 0 four records before its run of kCcRun outputs and steps the
//            recurrence - 10 multiply-adds by the 32-bit binomials C(512, 1..4) plus five 64-bit subtractions per
//            output instead of the 25 64x64-bit products of the direct form (ddc_back.cuh: cic_combine, which the
//            tests keep as the cross-check).  The 96 kHz CIC outputs live only in shared memory.
//   FIR    : one thread per (rail, pair of frames): the two 65-tap windows overlap in 63 samples, which are read once
//            as 16-byte vectors of widened samples; 32-bit wrap-around accumulators (only the low 31 bits are kept by
//            rx_ciccomp.vhd:616).
// ------------------------------------------------------------------------------------------------
constexpr int kCcFrames = 128;                         // frames per CTA tile
constexpr int kCcChunks = 2 * kCcFrames;               // chunks per tile
constexpr int kCcRecs = kCcChunks + kLHalo;            // records staged
constexpr int kCcOut = kCcChunks + kUHalo;             // CIC outputs per rail and tile
constexpr int kCcRun = 11;                             // outputs per run; odd, so that the strided 64-bit record reads of a half warp fall into distinct banks
constexpr int kCcWarm = kLHalo - kUHalo;               // 4 records of run-in
constexpr int kCcThreads = 128;
constexpr int kCcUPitch = (kCcOut + 8 + 3) & ~3;       // int32 samples per rail, zero padded for the vector reads; a multiple of 4 keeps rail 1 16-byte aligned
static_assert(kCcUPitch % 4 == 0, "16-byte vector reads of s_u");
static_assert(2 * ((kCcOut + kCcRun - 1) / kCcRun) <= kCcThreads, "one thread per (rail, run)");
static_assert(2 * (kCcFrames / 2) <= kCcThreads, "one thread per (rail, frame pair)");
static_assert((kCcRecs * kLRec * 8) % 16 == 0, "TMA bulk size");

__global__ void __launch_bounds__(kCcThreads)
ddc_ciccomp_kernel(const uint64_t* __restrict__ L, uint32_t l_ch_stride, uint32_t n_frames, int16_t* __restrict__ YI,
                   uint32_t yi_stride, int16_t* __restrict__ YQ, uint32_t yq_stride, uint32_t align_b) {
    __shared__ __align__(128) uint64_t s_rec[kCcRecs * kLRec];       // 324 x 10 x 8 B, record-major as in global memory
    __shared__ __align__(16) int32_t s_u[2][kCcUPitch];
    __shared__ __align__(8) uint64_t s_bar;
    const uint32_t tid = threadIdx.x;
    const uint32_t ch = blockIdx.x;
    const uint32_t k0 = blockIdx.y * kCcFrames;                      // first frame of the tile
    const uint32_t nk = min((uint32_t)kCcFrames, n_frames - k0);
    const uint32_t n_rec = 2 * nk + kLHalo;
    // record index (array, halo included) of the tile's first staged record: chunk 2*k0 - 69 -> array index 2*k0
    const uint64_t* src = L + (size_t)ch * l_ch_stride + (size_t)(2 * k0) * kLRec;
    UA3_PDL_WAIT();                                                  // the front kernel's records
    UA3_PDL_TRIGGER();
#if !defined(UA3_HOST_EMU)
    if (tid == 0) {
        mbar_init(&s_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        mbar_expect_tx(&s_bar, n_rec * kLRec * 8u);
        tma_bulk_g2s(s_rec, src, n_rec * kLRec * 8u, &s_bar);
    }
#else
    for (uint32_t i = tid; i < n_rec * kLRec; i += kCcThreads) s_rec[i] = src[i];
#endif
    for (uint32_t i = tid; i < 2 * kCcUPitch; i += kCcThreads) (&s_u[0][0])[i] = 0;
    __syncthreads();                                                 // barrier initialised / emulation copy done
#if !defined(UA3_HOST_EMU)
    mbar_wait(&s_bar, 0);
#endif
    // ---- combs: CIC outputs u'[c] for chunks c = 2*k0 - 65 .. 2*k0 + 2*nk - 1.  The 65-tap window of frame k ends at
    // u'[2k] in alignment A and at u'[2k - 1] in alignment B (rx_ciccomp.vhd:339-361: which of the two delay lines a
    // CIC output enters); the outputs are stored so that the window is s_u[rail][2k .. 2k + 64] either way: output m
    // goes to s_u[m - 1 + align_b], and alignment A never needs m = 0. ----
    const uint32_t n_out = 2 * nk + kUHalo;
    const uint32_t n_runs = (n_out + kCcRun - 1) / kCcRun;
    if (tid < 2 * n_runs) {
        const uint32_t rail = tid / n_runs, m0 = (tid % n_runs) * kCcRun;      // staged record of output m is m + kCcWarm
        const uint64_t* rec = s_rec + (size_t)m0 * kLRec + rail * 5;
        uint64_t S0 = 0, S1 = 0, S2 = 0, S3 = 0, S4 = 0;
        uint64_t d0 = 0, d1 = 0, d2 = 0, d3 = 0, d4 = 0;                       // comb delay registers (rx_cic.vhd:293-404)
        const uint32_t steps = min((uint32_t)(kCcRun + kCcWarm), n_rec - m0);
#pragma unroll 3
        for (uint32_t j = 0; j < steps; ++j, rec += kLRec) {
            // S <- A512 S + L, A512[r][c] = C(512, r - c): 512, 130816, 22238720, 2829877120
            S4 += 512u * S3 + 130816u * S2 + 22238720u * S1 + 2829877120u * S0 + rec[4];
            S3 += 512u * S2 + 130816u * S1 + 22238720u * S0 + rec[3];
            S2 += 512u * S1 + 130816u * S0 + rec[2];
            S1 += 512u * S0 + rec[1];
            S0 += rec[0];
            uint64_t y = S4, t;
            t = y - d0; d0 = y; y = t;
            t = y - d1; d1 = y; y = t;
            t = y - d2; d2 = y; y = t;
            t = y - d3; d3 = y; y = t;
            t = y - d4; d4 = y; y = t;
            if (j >= (uint32_t)kCcWarm) {                            // output_typeconvert <= section_out10(59 DOWNTO 44)
                const uint32_t pos = m0 + j - kCcWarm + align_b;     // = m + align_b; stored at pos - 1
                if (pos) s_u[rail][pos - 1] = (int32_t)(int16_t)(uint16_t)(y >> 44);
            }
        }
    }
    __syncthreads();
    // ---- compensator: frames k0 + 2g, k0 + 2g + 1 of one rail per thread (rx_ciccomp.vhd:339-616) ----
    const uint32_t n_pairs = (nk + 1) / 2;
    if (tid < 2 * n_pairs) {
        const uint32_t rail = tid / n_pairs, g = tid % n_pairs;
        const I4* w4 = reinterpret_cast<const I4*>(&s_u[rail][4 * g]);       // window of frame k: s_u[2k .. 2k + 64]
        int32_t a0 = 0, a1 = 0;
#pragma unroll
        for (int v = 0; v < 17; ++v) {
            const I4 q = w4[v];
            const int32_t x[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int i = 4 * v + e;                             // window position; tap j multiplies position 64 - j
                if (i <= 64) a0 += (int32_t)c_comp_h[64 - i] * x[e];
                if (i >= 2 && i <= 66) a1 += (int32_t)c_comp_h[66 - i] * x[e];
            }
        }
        const int32_t accs[2] = {a0, a1};
#pragma unroll
        for (int f = 0; f < 2; ++f) {
            const uint32_t k = 2 * g + f;
            if (k < nk) {
                // rx_ciccomp.vhd:616: low 31 bits + 0x3FFF + bit15, wrap at 31 bits, >> 15, keep 16 bits
                const uint32_t a31 = (uint32_t)accs[f] & 0x7FFFFFFFu;
                const uint32_t r31 = (a31 + 0x3FFFu + (((uint32_t)accs[f] >> 15) & 1u)) & 0x7FFFFFFFu;
                const int16_t y = (int16_t)((int32_t)(r31 << 1) >> 16);
                if (rail == 0) YI[(size_t)ch * yi_stride + kYIHalo + k0 + k] = y;
                else           YQ[(size_t)ch * yq_stride + kYQHalo + k0 + k] = y;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// hilb + delay + frame pack: CTA = (channel, tile of 256 frames)
// ------------------------------------------------------------------------------------------------
// rx_hilb.vhd:903 rounds every product: pr = (p + p[1]) >> 1 (convergent).  Because 2*pr = p + adj with
// adj = +1 if p mod 4 == 3, -1 if p mod 4 == 1, 0 otherwise, the 256-tap sum is exactly
//     sum pr = ( sum p + 2 * #{p mod 4 == 3} - #{p odd} ) / 2,
// and p mod 4 depends only on the two low bits of coefficient and sample: the rounding term comes from 256-bit
// popcounts over bit planes of the samples (built with warp ballots) against bit planes of the coefficients.
// The plain sum uses the antisymmetry of the coefficient set, c[t] = -c[255 - t] (checked in tests/test_tables.py):
//     sum p = sum_{t<128} c[t] * (y[k - t] - y[k - 255 + t]),
// one subtraction (either integer pipe) and one IMAD per PAIR of taps, in 32-bit wrap-around arithmetic - only the
// low 31 bits of the doubled sum survive rx_hilb.vhd:935.  A thread produces 4 consecutive frames from a register
// window of widened samples read as 16-byte vectors (2 vector loads per 32 multiply-adds).
constexpr int kHbFrames = 256;                        // frames per CTA tile
constexpr int kHbPer = 4;                             // frames per thread
constexpr int kHbThreads = kHbFrames / kHbPer;        // 64
constexpr int kHbWin = 528;                           // widened window: 511 samples + zero padding for the vector reads

__global__ void __launch_bounds__(kHbThreads)
ddc_hilb_kernel(const int16_t* __restrict__ YI, uint32_t yi_stride, const int16_t* __restrict__ YQ, uint32_t yq_stride,
                uint32_t n_frames, uint64_t* __restrict__ frames, uint32_t frame_ch_stride, uint32_t ring_start,
                uint32_t ring_mask, uint32_t d_i, uint32_t d_q) {
    __shared__ __align__(16) int32_t s_y[kHbWin];  // window: s_y[i] = yI[k0 - d_i - 255 + i], zero padded
    __shared__ uint32_t s_b0[18], s_b1[18];        // bit planes of the window, 32 samples per word
    const uint32_t ch = blockIdx.x;
    const uint32_t k0 = blockIdx.y * kHbFrames;
    const uint32_t nk = min((uint32_t)kHbFrames, n_frames - k0);
    const uint32_t tid = threadIdx.x;
    // VOICE_I of frame k is the Hilbert sum d_i samples back (serial MAC + output register, rx_hilb.vhd:907-947)
    const int16_t* src = YI + (size_t)ch * yi_stride + k0 + (kMaxDI - d_i);
    UA3_PDL_WAIT();                                        // YI / YQ of ddc_ciccomp_kernel
    UA3_PDL_TRIGGER();
    for (uint32_t i = tid; i < 512; i += kHbThreads) {     // every warp: 32 consecutive samples -> one word per plane
        const int32_t v = (i < nk + 255u) ? (int32_t)src[i] : 0;
        s_y[i] = v;
        const uint32_t w0 = __ballot_sync(0xffffffffu, v & 1), w1 = __ballot_sync(0xffffffffu, v & 2);
        if ((tid & 31) == 0) { s_b0[i >> 5] = w0; s_b1[i >> 5] = w1; }
    }
    if (tid < kHbWin - 512) s_y[512 + tid] = 0;
    if (tid < 2) { s_b0[16 + tid] = 0; s_b1[16 + tid] = 0; }
    __syncthreads();
    const uint32_t kt = kHbPer * tid;              // first frame of this thread, tile-relative
    if (kt >= nk) return;
    // ---- plain sums: acc[o] = sum_{t<128} c[t] * (s_y[kt+o+255-t] - s_y[kt+o+t]) ----
    const I4* lv = reinterpret_cast<const I4*>(s_y + kt);            // left stream : s_y[kt + 4v ..]
    const I4* rv = reinterpret_cast<const I4*>(s_y + kt + 128);      // right stream: s_y[kt + 128 + 4u ..]
    uint32_t acc[kHbPer] = {0, 0, 0, 0};
    I4 l0 = lv[0], r1 = rv[32];
#pragma unroll
    for (int v = 0; v < 32; ++v) {
        const I4 l1 = lv[v + 1], r0 = rv[31 - v];
        const int32_t lw[8] = {l0.x, l0.y, l0.z, l0.w, l1.x, l1.y, l1.z, l1.w};     // s_y[kt + 4v + 0..7]
        const int32_t rw[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};     // s_y[kt + 128 + 4(31-v) + 0..7]
#pragma unroll
        for (int e = 0; e < 4; ++e)
#pragma unroll
            for (int o = 0; o < kHbPer; ++o)
                acc[o] += (uint32_t)c_hilb_c32[4 * v + e] * (uint32_t)(rw[3 + o - e] - lw[o + e]);
        l0 = l1; r1 = r0;
    }
    const int16_t* qsrc = YQ + (size_t)ch * yq_stride + k0 + kt;     // qsrc[kYQHalo + o] = yQ[k], qsrc[kYQHalo - d_q + o] = yQ[k - d_q]
    const int16_t* isrc = YI + (size_t)ch * yi_stride + kYIHalo + k0 + kt;   // isrc[o] = yI[k]
    uint64_t* dst = frames + (size_t)ch * frame_ch_stride;
#pragma unroll
    for (int o = 0; o < kHbPer; ++o) {
        if (kt + o >= nk) break;
        // ---- rounding term: u = 255 - t runs over the window from bit kt + o on ----
        const uint32_t wsel = (kt + o) >> 5, sh = (kt + o) & 31;
        uint32_t n_odd = 0, n_three = 0;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const uint32_t x0 = __funnelshift_r(s_b0[wsel + j], s_b0[wsel + j + 1], sh);
            const uint32_t x1 = __funnelshift_r(s_b1[wsel + j], s_b1[wsel + j + 1], sh);
            const uint32_t q0 = x0 & c_hilb_b0[j];
            const uint32_t q1 = (x0 & c_hilb_b1[j]) ^ (x1 & c_hilb_b0[j]);
            n_odd += __popc(q0);
            n_three += __popc(q0 & q1);
        }
        const uint32_t twice = acc[o] + 2u * n_three - n_odd;        // mod 2^32; even, so >> 1 is exact on bits 1..31
        const uint32_t a = twice >> 1;                               // low 31 bits of the HDL accumulator
        // rx_hilb.vhd:935: low 30 bits + 0x1FFF + bit14, wrap at 30 bits, >> 14, keep 16 bits
        const uint32_t a30 = a & 0x3FFFFFFFu;
        const uint32_t r30 = (a30 + 0x1FFFu + ((a >> 14) & 1u)) & 0x3FFFFFFFu;
        const int16_t vi = (int16_t)((int32_t)(r30 << 2) >> 16);
        dst[(ring_start + k0 + kt + o) & ring_mask] = frame_pack(qsrc[kYQHalo + o], isrc[o], qsrc[kYQHalo - d_q + o], vi);
    }
}

// ------------------------------------------------------------------------------------------------
// rotate: CTA per channel; tails -> halos (read everything, barrier, write), NCO phase advance
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
ddc_rotate_kernel(uint64_t* __restrict__ L, uint32_t l_ch_stride,
                  int16_t* __restrict__ YI, uint32_t yi_stride, int16_t* __restrict__ YQ, uint32_t yq_stride,
                  uint32_t n_chunks, uint32_t n_frames, uint32_t* __restrict__ phase, const uint32_t* __restrict__ fcw,
                  uint32_t n_ch) {
    const uint32_t ch = blockIdx.x, t = threadIdx.x;
    uint64_t* l = L + (size_t)ch * l_ch_stride;
    int16_t* yi = YI + (size_t)ch * yi_stride;
    int16_t* yq = YQ + (size_t)ch * yq_stride;
    constexpr int kLPer = (kLHalo * kLRec + 255) / 256;               // halo words per thread
    static_assert(kYIHalo <= 512 && kYQHalo <= 256, "halo rotation: two YI words and one YQ word per thread");
    uint64_t vl[kLPer]; int16_t vyi = 0, vyi2 = 0, vyq = 0;
    // The chunk records and the phases are used by the front kernel and ddc_ciccomp_kernel only, and those have completed when
    // a CTA of this kernel runs at all: launched the ordinary way it starts after ddc_hilb_kernel has finished; launched
    // programmatically it starts after every CTA of ddc_hilb_kernel has passed ITS wait on ddc_ciccomp_kernel (ddc_hilb_kernel
    // waits before it triggers).  So this part runs beside ddc_hilb_kernel; only the YI / YQ halos wait for it.
#pragma unroll
    for (int q = 0; q < kLPer; ++q) { const uint32_t w = t + 256u * q; vl[q] = (w < kLHalo * kLRec) ? l[(size_t)n_chunks * kLRec + w] : 0; }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < kLPer; ++q) { const uint32_t w = t + 256u * q; if (w < kLHalo * kLRec) l[w] = vl[q]; }
    UA3_PDL_WAIT();                                                   // ddc_hilb_kernel still reads the YI / YQ halos
    UA3_PDL_TRIGGER();
    if (t < kYIHalo) vyi = yi[n_frames + t];
    if (t + 256 < kYIHalo) vyi2 = yi[n_frames + t + 256];
    if (t < kYQHalo) vyq = yq[n_frames + t];
    __syncthreads();
    if (t < kYIHalo) yi[t] = vyi;
    if (t + 256 < kYIHalo) yi[t + 256] = vyi2;
    if (t < kYQHalo) yq[t] = vyq;
    if (t == 0 && ch < n_ch) phase[ch] = (phase[ch] + fcw[ch] * (n_chunks * (uint32_t)kCicR)) & 0x3FFFFFu;
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
static uint64_t binom_u64(uint64_t n, int k) {
    if (k < 0) return 0;
    unsigned __int128 r = 1;
    for (int i = 1; i <= k; ++i) r = r * (n - (uint64_t)(k - i)) / (unsigned)i;   // exact at every step
    return (uint64_t)r;
}

void build_cic_weights(uint64_t G[25]) {
    static const int64_t c5[6] = {1, -5, 10, -10, 5, -1};
    for (int p = 0; p < 5; ++p)
        for (int k = 1; k <= 5; ++k) {
            uint64_t g = 0;
            for (int i = 0; i <= p; ++i) g += (uint64_t)c5[i] * binom_u64((uint64_t)kCicR * (uint64_t)(p - i), 5 - k);
            G[p * 5 + (k - 1)] = g;
        }
}

void build_nco_big_table(uint32_t* tab /* kBigTabWords */) {
    for (int k = 0; k < 2048; ++k)
        for (int sf = 0; sf < kSfLevels; ++sf) {
            tab[nco_bigtab_index((uint32_t)k, (uint32_t)sf)] = nco_bigtab_entry(UA3_NCO_SIN_C[k], UA3_NCO_COS_C[k], sf);
        }
}

// binary16 image of the big table for the tensor-core kernel: sin in the high half, cos in the low half
static uint32_t half_bits_of(int32_t v) {              // binary16 encoding of an integer, |v| <= 2048 (exact)
    if (v == 0) return 0;
    const uint32_t sign = v < 0 ? 0x8000u : 0u;
    const uint32_t m = (uint32_t)(v < 0 ? -v : v);
    int e = 0;
    while ((m >> (e + 1)) != 0) ++e;                       // m = 1.f * 2^e, e <= 11
    const uint32_t frac = e <= 10 ? (m << (10 - e)) : (m >> (e - 10));
    return sign | ((uint32_t)(e + 15) << 10) | (frac & 0x3FFu);
}

void build_nco_half_table(uint32_t* tab /* kBigTabWords */) {
    build_nco_big_table(tab);
    for (int i = 0; i < kBigTabWords; ++i) {
        const int32_t s12 = (int32_t)tab[i] >> 16, c12 = (int32_t)(int16_t)(tab[i] & 0xFFFFu);
        tab[i] = (half_bits_of(s12) << 16) | half_bits_of(c12);
    }
}
static_assert(kTcWeightBytes == kTcWeightPlaneBytes, "ddc_launch.h");

// byte planes of C(511 - t, k), k = 0..4, in the canonical K-major no-swizzle layout the MMA's shared-memory descriptor
// walks: one 512-byte block per 32-sample slice, [16-sample half][8-column group][column][sample]
void build_tc_weight_planes(uint8_t* w /* kTcWeightBytes */, uint64_t fix[5]) {
    for (int i = 0; i < kTcWeightBytes; ++i) w[i] = 0;
    for (int k = 0; k < 5; ++k) fix[k] = 0;
    for (int t = 0; t < kCicR; ++t) {
        const int s = t / 32, kk = t % 32;
        int col = 0;
        for (int k = 0; k < 5; ++k) {
            const uint64_t wt = binom_u64((uint64_t)(kCicR - 1 - t), k);
            fix[k] += wt * 16384ull;
            for (int p = 0; p < tc_planes_of(k); ++p, ++col)
                w[(size_t)s * kTcN * 32 + (kk / 16) * 256 + (col / 8) * 128 + (col % 8) * 16 + (kk % 16)] = (uint8_t)(wt >> (8 * p));
        }
    }
}

cudaError_t ddc_prepare_kernels() {
#if !defined(UA3_HOST_EMU)
    cudaError_t e = cudaFuncSetAttribute(ddc_front_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTcSmemBytes);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(ddc_front_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTcSmemBytes);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(ddc_front_bt_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kBtSmemBytes);
#else
    return cudaSuccess;
#endif
}

void build_nco_table(uint32_t tab[2048]) {
    for (int k = 0; k < 2048; ++k) tab[k] = nco_pack(UA3_NCO_SIN_C[k], UA3_NCO_COS_C[k]);
}

cudaError_t ddc_upload_constants() {
    uint64_t G[25];
    build_cic_weights(G);
    cudaError_t e = cudaMemcpyToSymbol(c_cic_g, G, sizeof G);
    if (e != cudaSuccess) return e;
    e = cudaMemcpyToSymbol(c_comp_h, UA3_RXCOMP_H, sizeof(int16_t) * kCompTaps);
    if (e != cudaSuccess) return e;
    int32_t c32[kHilbTaps];
    uint32_t b0[8] = {0}, b1[8] = {0};
    for (int t = 0; t < kHilbTaps; ++t) {
        c32[t] = UA3_RXHILB_C[t];
        const int u = kHilbTaps - 1 - t;                               // window position of tap t
        b0[u >> 5] |= (uint32_t)(UA3_RXHILB_C[t] & 1) << (u & 31);
        b1[u >> 5] |= (uint32_t)((UA3_RXHILB_C[t] >> 1) & 1) << (u & 31);
    }
    e = cudaMemcpyToSymbol(c_hilb_c32, c32, sizeof c32);
    if (e != cudaSuccess) return e;
    e = cudaMemcpyToSymbol(c_hilb_b0, b0, sizeof b0);
    if (e != cudaSuccess) return e;
    return cudaMemcpyToSymbol(c_hilb_b1, b1, sizeof b1);
}

cudaError_t ddc_launch_block(const DdcBuffers& b, const int16_t* adc_dev, uint32_t n_samples, uint32_t ring_start,
                             int sm_count, cudaStream_t st, int* launches, cudaEvent_t* ev, uint32_t ev_mask) {
    const uint32_t n_chunks = n_samples / kCicR, n_frames = n_samples / kFrameAdc;
    if (n_chunks == 0) return cudaSuccess;
    if (ev && ((ev_mask >> 0) & 1u)) cudaEventRecord(ev[0], st);
    // programmatic dependent launch between the kernels of the chain (tensor-core path; every kernel of it calls UA3_PDL_WAIT
    // before it reads or overwrites what an earlier one uses): the launch latencies and the front kernel's 216 KB prologue
    // run under the predecessor's tail.  A launch that directly follows a recorded profiling event is made the ordinary way
    // (the event sits between the two kernels): with bench.py's front-kernel bracket that is the front kernel and the one
    // behind it, the other boundaries keep their overlap inside the timed region.
    auto pdl_at = [&](int point) { return b.pdl != 0 && !(ev && ((ev_mask >> point) & 1u)); };
    // big-table kernel when a CTA tile (256 channels x 4 chunks) can be filled; the 8 KB-table kernel otherwise
    const bool big = b.front_variant != 1 && b.big_tab && (b.front_variant == 2 || ((b.n_ch_pad >> 5) >= (uint32_t)kBtCG && n_chunks >= (uint32_t)kBtTG));
#if !defined(UA3_HOST_EMU)
    // tensor-core kernel: same threshold as the big-table kernel (a bank that fills the SMs' 128-channel warpgroups)
    const bool tcore = b.tab_h && (b.front_variant == 3 || (b.front_variant == 0 && big));
#else
    const bool tcore = false;
#endif
    if (tcore) {
#if !defined(UA3_HOST_EMU)
        const uint32_t n8 = n_chunks * (uint32_t)kCicR / 8u;
        UA3_LAUNCH_PDL(pdl_at(0) && pdl_at(5), adc_prepare_tc_kernel, (n8 + 255u) / 256u, 256, 0, st, adc_dev, n8, b.adc_h, b.wrap_flag, b.tile_counter);
        if (launches) *launches += 1;
        if (ev && ((ev_mask >> 1) & 1u)) cudaEventRecord(ev[1], st);
        const uint32_t n_tiles = ((b.n_ch_pad + 127u) / 128u) * n_chunks;
        const uint32_t grid = (uint32_t)min((uint64_t)(n_tiles + kTcWg - 1) / kTcWg, (uint64_t)sm_count);
        TcFix fix;
        for (int k = 0; k < 5; ++k) fix.c[k] = b.tc_fix[k];
        if (b.tc_adc_stage)
            UA3_LAUNCH_PDL(pdl_at(1), ddc_front_tc_kernel<true>, grid, kTcThreads, kTcSmemBytes, st, b.adc_h, b.wrap_flag, n_chunks, b.tab_h, b.fcw, b.phase,
                       b.n_ch_pad, b.tc_w, b.L, b.l_ch_stride, b.tile_counter, fix);
        else
            UA3_LAUNCH(ddc_front_tc_kernel<false>, grid, kTcThreads, kTcSmemBytes, st, b.adc_h, b.wrap_flag, n_chunks, b.tab_h, b.fcw, b.phase,
                       b.n_ch_pad, b.tc_w, b.L, b.l_ch_stride, b.tile_counter, fix);
#endif
    } else if (big) {
        const uint32_t n_tiles = (((b.n_ch_pad >> 5) + kBtCG - 1) / kBtCG) * ((n_chunks + kBtTG - 1) / kBtTG);
        // Tiles all take the same time, so the kernel lasts ceil(n_tiles / grid) tile times: of the SMs it may use it takes
        // only as many as that number of rounds needs (2048 tiles: 137 CTAs do 15 rounds exactly like 146 would) and
        // leaves the others to whatever runs beside it - the STM32 stage, an NCCL broadcast.
        uint32_t grid = (uint32_t)min((uint64_t)n_tiles, (uint64_t)sm_count);
        const uint32_t rounds = (n_tiles + grid - 1) / grid;
        grid = (n_tiles + rounds - 1) / rounds;
#if !defined(UA3_HOST_EMU)
        const uint32_t n8 = n_chunks * (uint32_t)kCicR / 8u;
        UA3_LAUNCH(adc_expand_kernel, (n8 + 255u) / 256u, 256, 0, st, adc_dev, n8, b.adc9, b.tile_counter);
        if (launches) *launches += 1;
#endif
        if (ev && ((ev_mask >> 1) & 1u)) cudaEventRecord(ev[1], st);
        UA3_LAUNCH(ddc_front_bt_kernel, grid, kBtThreads, kBtSmemBytes, st, adc_dev, b.adc9, n_chunks, b.big_tab, b.fcw, b.phase,
                   b.n_ch_pad, b.L, b.l_ch_stride, b.tile_counter);
    } else {
        const uint32_t n_tiles = ((n_chunks + kFrontWarps - 1) / kFrontWarps) * (b.n_ch_pad >> 5);
        const uint32_t grid = (uint32_t)min((uint64_t)n_tiles, (uint64_t)sm_count * 3);
        if (ev && ((ev_mask >> 1) & 1u)) cudaEventRecord(ev[1], st);
        UA3_LAUNCH(ddc_front_kernel, grid, kFrontThreads, 0, st, adc_dev, n_chunks, b.nco_tab, b.fcw, b.phase, b.n_ch_pad,
                   b.L, b.l_ch_stride);
    }
    if (ev && ((ev_mask >> 2) & 1u)) cudaEventRecord(ev[2], st);
    UA3_LAUNCH_PDL(pdl_at(2), ddc_ciccomp_kernel, dim3(b.n_ch, (n_frames + kCcFrames - 1) / kCcFrames), kCcThreads, 0, st, b.L, b.l_ch_stride, n_frames,
               b.YI, b.yi_stride, b.YQ, b.yq_stride, (uint32_t)b.align_b);
    if (ev && ((ev_mask >> 3) & 1u)) cudaEventRecord(ev[3], st);
    UA3_LAUNCH_PDL(pdl_at(3), ddc_hilb_kernel, dim3(b.n_ch, (n_frames + kHbFrames - 1) / kHbFrames), kHbThreads, 0, st, b.YI, b.yi_stride, b.YQ, b.yq_stride,
               n_frames, b.frames, b.frame_ch_stride, ring_start, b.ring_mask, (uint32_t)b.d_i, (uint32_t)b.d_q);
    if (ev && ((ev_mask >> 4) & 1u)) cudaEventRecord(ev[4], st);
    UA3_LAUNCH_PDL(pdl_at(4), ddc_rotate_kernel, b.n_ch_pad, 256, 0, st, b.L, b.l_ch_stride, b.YI, b.yi_stride,
               b.YQ, b.yq_stride, n_chunks, n_frames, b.phase, b.fcw, b.n_ch_pad);
    if (ev && ((ev_mask >> 5) & 1u)) cudaEventRecord(ev[5], st);
    if (launches) *launches += kDdcKernels - 1;
    return cudaGetLastError();
}

}  // namespace ua3
