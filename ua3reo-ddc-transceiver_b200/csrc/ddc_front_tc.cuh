// ddc_front_tc.cuh - front kernel, tensor-core variant (sm_100a: tcgen05 / TMEM), included by ddc.cu.
//
// NCO + mixer + the five CIC integrators of rx_cic.vhd over one 512-sample chunk are linear in the mixer output
// x (15 bit): integrator stage k at the end of the chunk, started from zero, is sum_t C(511 - t, k - 1) * x_t
// (mod 2^64; ddc_back.cuh folds the chunks together).  That is a GEMM with rows = channels, K = samples and the
// binomial weights as columns, and it is EXACT on the int8 tensor cores once x and the weights are cut into bytes:
//
//   A (TMEM, written by the threads): per slice of 32 samples four byte planes - x low byte and x high byte + 64
//       (both unsigned) of either rail - 8 columns each, lane = channel
//   B (shared memory, 8 KB for the chunk): the byte planes of C(511 - t, k - 1): 1 + 2 + 3 + 4 + 4 = 14 columns (N = 16)
//   D (TMEM, int32): 4 x 16 columns, |sum| <= 512 * 255 * 255 < 2^26; read once per chunk and recombined into the
//       ten 64-bit words of the L record:  sum_p (D_lo[p] + 256 * D_hi[p]) << 8p  -  16384 * sum_t C(511 - t, k - 1).
//
// What is left on the CUDA cores is the NCO and the mixer, and those run on the FP32 pipe: the (coarse address, fine
// level) table holds sin and cos as two binary16 numbers (12-bit integers are exact), the ADC block is binary16 as
// well, and ONE mixed-precision FMA per product, adc * nco + 1.5 * 2^23, leaves the exact 23-bit product in the
// mantissa: byte 1 of the float is the low byte of the CIC input, byte 2 its high byte + 64.  (-2048) * (-2048) = 2^22,
// the one product that wraps in the 23-bit mixer register (UA3REO.bdf mixer -> rx_cic In1[22..8]), carries into the
// exponent's last bit; chunks that contain an ADC sample of -2048 (flagged by adc_prepare_kernel) mask that bit.
//
// One persistent CTA per SM: four warpgroups of compute warps (lane = channel, 128 channels per warpgroup, 128 TMEM
// columns each: D 64, A 2 x 32) and four single-lane MMA issuer warps.  No CTA-wide barrier in the loop:
//   compute thread : 32 samples -> byte planes -> [wait: A buffer free] -> tcgen05.st -> arrive on full[b]
//   issuer lane    : wait full[b] -> 4 MMAs (M128 N16 K32, kind::i8) -> commit -> free[b]   (slice 15: also acc_ready)
//   compute thread : wait acc_ready -> tcgen05.ld -> arrive on acc_free -> recombine -> store the L record
// Tiles (128 channels x 1 chunk) are handed out from a global counter, one fetch per warpgroup and tile, so a CTA that
// starts late or shares its SM simply takes fewer.
//
// Bound (profiles/r02_front_tc.md): the shared-memory pipe - 32 lanes look up 32 unrelated table words per sample,
// 3.6 bank wavefronts per lookup - not the issue slots (13.8 instructions per sample and channel instead of 24.3).
#pragma once
#include "ddc_front.cuh"
#include <type_traits>

namespace ua3 {

constexpr int kTcN = 16;                          // MMA N: 14 weight byte planes, padded
constexpr int kTcSlices = kCicR / 32;             // K = 32 samples per MMA
constexpr int kTcWg = 4;                          // warpgroups per CTA: 4 x 128 TMEM columns
constexpr int kTcThreads = kTcWg * 128 + kTcWg * 32;
constexpr int kTcWeightBytes = kTcSlices * kTcN * 32;
constexpr int kTcAdcStageBytes = kTcWg * 2 * kCicR * 2;                             // per warpgroup two chunks of binary16 samples
constexpr size_t kTcSmemBytes = (size_t)kBigTabWords * 4 + kTcWeightBytes + kTcAdcStageBytes + 256;   // table + weights + ADC chunks + 25 barriers + 10 words
UA3_HD constexpr int tc_planes_of(int k) { return k == 0 ? 1 : (k == 1 ? 2 : (k == 2 ? 3 : 4)); }   // bytes of C(511, k)

struct TcFix { uint64_t c[5]; };                  // 16384 * sum_t C(511 - t, k): the +64 of the high byte planes

#if !defined(UA3_HOST_EMU)
namespace tc {
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void named_bar(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void tmem_st8(uint32_t addr, const uint32_t (&r)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(addr), "r"(r[0]), "r"(r[1]), "r"(r[2]),
                 "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
                 : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t addr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
          "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(addr));
}
__device__ __forceinline__ void mma_i8(uint32_t d, uint32_t a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, {%5, %5, %5, %5}, p;\n\t}" ::"r"(d), "r"(a), "l"(bdesc), "r"(idesc), "r"(accumulate),
        "r"(0u)
        : "memory");
}
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// four samples: a01 / a23 hold the binary16 ADC samples, w[] the (sin, cos) table words; results are raw float bits
__device__ __forceinline__ void mix4(uint32_t a01, uint32_t a23, const uint32_t (&w)[4], uint32_t (&fi)[4], uint32_t (&fq)[4], float magic) {
    asm("{.reg .b16 a0, a1, a2, a3, s, c;\n"
        "mov.b32 {a0, a1}, %8;\n mov.b32 {a2, a3}, %9;\n"
        "mov.b32 {c, s}, %10;\n fma.rn.f32.f16 %0, a0, s, %14;\n fma.rn.f32.f16 %4, a0, c, %14;\n"
        "mov.b32 {c, s}, %11;\n fma.rn.f32.f16 %1, a1, s, %14;\n fma.rn.f32.f16 %5, a1, c, %14;\n"
        "mov.b32 {c, s}, %12;\n fma.rn.f32.f16 %2, a2, s, %14;\n fma.rn.f32.f16 %6, a2, c, %14;\n"
        "mov.b32 {c, s}, %13;\n fma.rn.f32.f16 %3, a3, s, %14;\n fma.rn.f32.f16 %7, a3, c, %14;}\n"
        : "=r"(fi[0]), "=r"(fi[1]), "=r"(fi[2]), "=r"(fi[3]), "=r"(fq[0]), "=r"(fq[1]), "=r"(fq[2]), "=r"(fq[3])
        : "r"(a01), "r"(a23), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]), "f"(magic));
}
// sum_p (lo[p] + 256 * hi[p]) << 8p over a stage's byte planes
__device__ __forceinline__ uint64_t recombine(const uint32_t* lo, const uint32_t* hi, int n_planes) {
    uint64_t acc = 0;
    for (int p = 0; p < n_planes; ++p) acc += ((uint64_t)lo[p] + ((uint64_t)hi[p] << 8)) << (8 * p);
    return acc;
}

// 32 samples of one channel -> the four A byte planes of a slice.  The slice's ADC samples are the same for every lane:
// STAGE = false reads them with broadcast LDG.128 (four load/store-pipe wavefronts per eight samples), STAGE = true from the
// chunk copy in shared memory with broadcast LDS.64 (one wavefront per four samples; tools/exp/lsu_wavefronts.cu).  Passing
// them round by warp shuffles was measured slower (0.630 against 0.600 ms) although shuffles are not data wavefronts.
template <bool WRAP, bool STAGE>
__device__ __forceinline__ void slice_planes(const uint32_t* __restrict__ s_tab, const void* __restrict__ src, uint32_t& P, uint32_t F, float magic,
                                             uint32_t (&ilo)[8], uint32_t (&ihi)[8], uint32_t (&qlo)[8], uint32_t (&qhi)[8]) {
    uint4 av4 = make_uint4(0, 0, 0, 0);
#pragma unroll
    for (int q = 0; q < 8; ++q) {                                    // four samples per step
        uint32_t a01, a23;
        if (STAGE) {
            const uint2 av = reinterpret_cast<const uint2*>(src)[q];
            a01 = av.x; a23 = av.y;
        } else {
            if ((q & 1) == 0) av4 = __ldg(reinterpret_cast<const uint4*>(src) + (q >> 1));
            a01 = (q & 1) ? av4.z : av4.x; a23 = (q & 1) ? av4.w : av4.y;
        }
        uint32_t w[4], fi[4], fq[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) { w[e] = s_tab[nco_bigtab_index(P >> 21, nco_fine_level(P))]; P += F; }
        mix4(a01, a23, w, fi, fq, magic);
        const uint32_t i01 = __byte_perm(fi[0], fi[1], 0x6251), i23 = __byte_perm(fi[2], fi[3], 0x6251);
        const uint32_t q01 = __byte_perm(fq[0], fq[1], 0x6251), q23 = __byte_perm(fq[2], fq[3], 0x6251);
        ilo[q] = __byte_perm(i01, i23, 0x5410); ihi[q] = __byte_perm(i01, i23, 0x7632);
        qlo[q] = __byte_perm(q01, q23, 0x5410); qhi[q] = __byte_perm(q01, q23, 0x7632);
        if (WRAP) { ihi[q] &= 0x7F7F7F7Fu; qhi[q] &= 0x7F7F7F7Fu; }
    }
}
__device__ __forceinline__ void cp_async8(void* dst_smem, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
}  // namespace tc

template <bool STAGE>
__global__ void __launch_bounds__(kTcThreads, 1)
ddc_front_tc_kernel(const uint16_t* __restrict__ adc_h, const uint8_t* __restrict__ wrap_flag, uint32_t n_chunks, const uint32_t* __restrict__ tab_h,
                    const uint32_t* __restrict__ fcw, const uint32_t* __restrict__ phase, uint32_t n_ch_pad, const uint8_t* __restrict__ wplanes,
                    uint64_t* __restrict__ L, uint32_t l_ch_stride, uint32_t* __restrict__ tile_counter, const TcFix fix) {
    extern __shared__ __align__(128) uint8_t s_dyn[];
    uint32_t* s_tab = reinterpret_cast<uint32_t*>(s_dyn);
    uint8_t* s_w = s_dyn + (size_t)kBigTabWords * 4;
    uint8_t* s_adc = s_w + kTcWeightBytes;                                            // [g][2][512] binary16: this tile's and the next tile's chunk
    uint64_t* s_bar = reinterpret_cast<uint64_t*>(s_adc + kTcAdcStageBytes);          // [g][6]: full0 full1 free0 free1 acc_ready acc_free
    uint32_t* s_misc = reinterpret_cast<uint32_t*>(s_bar + kTcWg * 6);                 // [0] TMEM base, [1 + 2g + (it & 1)] tile slots, [9] 1.5 * 2^23
    uint64_t* s_tabbar = reinterpret_cast<uint64_t*>(s_misc + 10);
    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;

    if (t == 0) {
        for (int g = 0; g < kTcWg; ++g) {
            mbar_init(&s_bar[g * 6 + 0], 128); mbar_init(&s_bar[g * 6 + 1], 128);
            mbar_init(&s_bar[g * 6 + 2], 1); mbar_init(&s_bar[g * 6 + 3], 1); mbar_init(&s_bar[g * 6 + 4], 1);
            mbar_init(&s_bar[g * 6 + 5], 128);
        }
        mbar_init(s_tabbar, 1);
        s_misc[9] = 0x4B400000u;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        mbar_expect_tx(s_tabbar, (uint32_t)kBigTabWords * 4u + (uint32_t)kTcWeightBytes);
        for (uint32_t off = 0; off < (uint32_t)kBigTabWords * 4u; off += 16384u)
            tma_bulk_g2s(s_dyn + off, reinterpret_cast<const uint8_t*>(tab_h) + off, 16384u, s_tabbar);
        tma_bulk_g2s(s_w, wplanes, (uint32_t)kTcWeightBytes, s_tabbar);
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&s_misc[0])));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    mbar_wait(s_tabbar, 0);                               // table and weights have landed (async proxy writes, visible to the MMA as well)
    UA3_PDL_WAIT();                                       // everything above ran under adc_prepare_tc_kernel (and its launch); now its outputs
    const uint32_t tmem = s_misc[0];
    const uint32_t n_cg = (n_ch_pad + 127u) / 128u, n_tiles = n_cg * n_chunks;
    const bool issuer = warp >= kTcWg * 4;
    const int g = issuer ? warp - kTcWg * 4 : warp >> 2;                    // warpgroup served
    uint64_t* bar_full = &s_bar[g * 6 + 0];
    uint64_t* bar_free = &s_bar[g * 6 + 2];
    uint64_t* bar_acc_ready = &s_bar[g * 6 + 4];
    uint64_t* bar_acc_free = &s_bar[g * 6 + 5];
    const uint32_t col0 = tmem + 128u * (uint32_t)g;                       // D at col0 .. +63, A buffers at +64 and +96

    if (!issuer) {
        const int tg = t & 127;                                           // thread within the warpgroup = channel within the tile
        const uint32_t lane_off = (uint32_t)((warp & 3) * 32) << 16;
        uint32_t par_free[2] = {0, 0}, par_acc = 0;
        uint32_t n_done = 0;                                              // slices stored so far (first use of each A buffer needs no wait)
        float magic;                                                      // a LOADED value: ptxas re-materialises a constant before every FMA
        asm volatile("ld.shared.f32 %0, [%1];" : "=f"(magic) : "r"(smem_u32(&s_misc[9])));
        // Tiles are fetched one ahead (two slots per warpgroup) so that the NEXT tile's ADC chunk can be copied into shared
        // memory (cp.async, 8 bytes per thread) while the current one is computed.
        uint32_t* s_tile = &s_misc[1 + 2 * g];
        uint8_t* adc_buf = s_adc + (size_t)g * (2 * kCicR * 2);
        if (tg == 0) { s_tile[0] = atomicAdd(tile_counter, 1u); s_tile[1] = atomicAdd(tile_counter, 1u); }
        tc::named_bar(1 + g, 160);
        if (STAGE) {
            const uint32_t first = s_tile[0];
            if (first < n_tiles) tc::cp_async8(adc_buf + tg * 8, adc_h + (size_t)(first / n_cg) * kCicR + tg * 4);
            tc::cp_async_wait_all();
            tc::named_bar(5 + g, 128);
        }
        for (uint32_t it = 0;; ++it) {
            const uint32_t tile = s_tile[it & 1u], next = s_tile[(it + 1u) & 1u];
            if (tile >= n_tiles) break;
            if (STAGE && next < n_tiles) tc::cp_async8(adc_buf + ((it + 1u) & 1u) * (kCicR * 2) + tg * 8, adc_h + (size_t)(next / n_cg) * kCicR + tg * 4);
            const uint32_t cg = tile % n_cg, chunk = tile / n_cg;         // channel group fastest: neighbours share the ADC chunk in L1/L2
            const uint32_t ch = cg * 128u + (uint32_t)tg;
            const bool live = ch < n_ch_pad;
            const uint32_t chl = live ? ch : n_ch_pad - 1u;
            const uint32_t F = fcw[chl] << 10;
            uint32_t P = (phase[chl] << 10) + F * (chunk * (uint32_t)kCicR);
            const uint8_t* a_src = STAGE ? adc_buf + (it & 1u) * (kCicR * 2) : reinterpret_cast<const uint8_t*>(adc_h + (size_t)chunk * kCicR);
            const bool wrap = wrap_flag[chunk] != 0;
            auto slices = [&](auto wrap_tag) {                            // the chunk's 16 slices; two copies of the loop, with and without the mask
                constexpr bool kWrap = decltype(wrap_tag)::value;
#pragma unroll 1
                for (int s = 0; s < kTcSlices; ++s, ++n_done) {
                    const int b = s & 1;
                    uint32_t ilo[8], ihi[8], qlo[8], qhi[8];
                    tc::slice_planes<kWrap, STAGE>(s_tab, a_src + s * 64, P, F, magic, ilo, ihi, qlo, qhi);
                    if (n_done >= 2) { mbar_wait(&bar_free[b], par_free[b]); par_free[b] ^= 1; }
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t a_col = col0 + lane_off + 64 + 32 * b;
                    tc::tmem_st8(a_col + 0, ilo); tc::tmem_st8(a_col + 8, ihi); tc::tmem_st8(a_col + 16, qlo); tc::tmem_st8(a_col + 24, qhi);
                    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
                    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                    tc::mbar_arrive(&bar_full[b]);
                }
            };
            if (wrap) slices(std::true_type{});
            else slices(std::false_type{});
            mbar_wait(bar_acc_ready, par_acc); par_acc ^= 1;
            // every thread of the warpgroup and the issuer have read this tile's slot (their arrivals / MMAs precede acc_ready):
            // it now takes the tile after next
            if (tg == 0) s_tile[it & 1u] = atomicAdd(tile_counter, 1u);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            uint32_t d_ilo[16], d_ihi[16], d_qlo[16], d_qhi[16];
            tc::tmem_ld16(col0 + lane_off + 0, d_ilo); tc::tmem_ld16(col0 + lane_off + 16, d_ihi);
            tc::tmem_ld16(col0 + lane_off + 32, d_qlo); tc::tmem_ld16(col0 + lane_off + 48, d_qhi);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            tc::mbar_arrive(bar_acc_free);                                // the issuer may overwrite the accumulators
            uint64_t out[10];
            int col = 0;
#pragma unroll
            for (int k = 0; k < 5; ++k) {
                out[k] = tc::recombine(d_ilo + col, d_ihi + col, tc_planes_of(k)) - fix.c[k];
                out[5 + k] = tc::recombine(d_qlo + col, d_qhi + col, tc_planes_of(k)) - fix.c[k];
                col += tc_planes_of(k);
            }
            if (live) {
                ulonglong2* d2 = reinterpret_cast<ulonglong2*>(L + (size_t)ch * l_ch_stride + (size_t)(kLHalo + chunk) * kLRec);
#pragma unroll
                for (int k = 0; k < 5; ++k) d2[k] = make_ulonglong2(out[2 * k], out[2 * k + 1]);
            }
            if (STAGE) tc::cp_async_wait_all();
            tc::named_bar(1 + g, 160);                                    // publishes the new slot and the next chunk's samples
        }
    } else {
        // instruction descriptor: D int32, A and B unsigned 8 bit, both K-major, N = 16, M = 128
        const uint32_t idesc = (2u << 4) | ((uint32_t)(kTcN >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        uint32_t par_full[2] = {0, 0}, par_acc_free = 0;
        uint32_t n_tiles_done = 0;
        const uint32_t* s_tile = &s_misc[1 + 2 * g];
        tc::named_bar(1 + g, 160);
        for (uint32_t it = 0;; ++it) {
            const uint32_t tile = s_tile[it & 1u];
            if (tile >= n_tiles) break;
            if (lane == 0) {
                for (int s = 0; s < kTcSlices; ++s) {
                    const int b = s & 1;
                    mbar_wait(&bar_full[b], par_full[b]); par_full[b] ^= 1;
                    if (s == 0 && n_tiles_done) { mbar_wait(bar_acc_free, par_acc_free); par_acc_free ^= 1; }
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    // shared-memory descriptor of the slice's 512-byte weight block: K-major, no swizzle, LBO = 256 B (next 16 K bytes),
                    // SBO = 128 B (next 8 columns), descriptor version 1 (sm_100)
                    const uint64_t bdesc = (uint64_t)(((smem_u32(s_w) + (uint32_t)s * kTcN * 32) >> 4) & 0x3FFF) | ((uint64_t)16 << 16) | ((uint64_t)8 << 32) |
                                           ((uint64_t)1 << 46);
                    const uint32_t acc = s > 0 ? 1u : 0u, a0 = col0 + 64 + 32 * b;
                    tc::mma_i8(col0 + 0, a0 + 0, bdesc, idesc, acc);
                    tc::mma_i8(col0 + 16, a0 + 8, bdesc, idesc, acc);
                    tc::mma_i8(col0 + 32, a0 + 16, bdesc, idesc, acc);
                    tc::mma_i8(col0 + 48, a0 + 24, bdesc, idesc, acc);
                    tc::mma_commit(&bar_free[b]);
                    if (s == kTcSlices - 1) tc::mma_commit(bar_acc_ready);
                }
            }
            ++n_tiles_done;
            __syncwarp();
            tc::named_bar(1 + g, 160);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    UA3_PDL_TRIGGER();                                    // out of tiles: ddc_ciccomp_kernel's CTAs may take this SM (they wait for the whole grid)
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem));
}
#endif  // !UA3_HOST_EMU

}  // namespace ua3
