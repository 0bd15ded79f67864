// api.cu - C ABI of libua3reo_b200.so (declared in include/ua3reo_b200.h).
#include "../../include/ua3reo_b200.h"
#include "ddc_launch.h"
#include "rx_launch.h"
#include "duc_launch.h"
#include "tx_launch.h"
#include "ua3_common.cuh"
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <new>
#include <string>
#include <vector>

using namespace ua3;
#include "rx_host.cpp.inc"

// SMs the front kernel leaves to the STM32 stage (see push_common): none since the stage's CTAs are packed and its streams have
// the higher priority.  Earlier sweeps (tools/gpu/sweep_rx_reserve.sh, unpacked CTAs): 1024 channels 1.022 / 1.037 / 1.071 ms per
// step at 6 / 8 / 12 SMs; 4096 channels 3.707 / 3.739 / 3.768 ms at 4 / 6 / 8.
static constexpr int kRxReserveSmall = 0, kRxReserveLarge = 0;   // below / from 2048 channels (UA3REO_RX_RESERVE_SMS overrides); the front kernel hands back every SM its round count does not need
static constexpr int kProfEvents = kDdcKernels + 3;   // 5 DDC kernels, rx_audio, rx_fft: 8 event points per block
static thread_local std::string g_err;

void ua3_set_last_error(const char* what) { g_err = what; }   // for the other translation units (fanout.cu)

static int fail(int code, const char* what, cudaError_t e = cudaSuccess) {
    g_err = what;
    if (e != cudaSuccess) { g_err += ": "; g_err += cudaGetErrorString(e); }
    return code;
}
#define UA3_CUDA(call)                                                  \
    do {                                                                \
        cudaError_t e__ = (call);                                       \
        if (e__ != cudaSuccess) return fail(UA3_E_CUDA, #call, e__);    \
    } while (0)

struct ua3reo_ctx {
    int device = 0, sm_count = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t copy_stream = nullptr;   // device->host result copies that overlap the next push
    cudaEvent_t ev_push = nullptr;        // end of the last push on `stream`
    cudaEvent_t ev_copy[2] = {nullptr, nullptr};   // end of the async frame read issued after push k (k & 1)
    bool copy_pending[2] = {false, false};
    uint64_t n_push = 0;
    // host pushes of whole blocks go through two staging buffers on their own stream, so that the
    // host->device copy of block k+1 overlaps the kernels of block k
    cudaStream_t h2d_stream = nullptr;
    int16_t* adc_stage2[2] = {nullptr, nullptr};
    cudaEvent_t ev_h2d = nullptr;                  // staging buffer filled
    cudaEvent_t ev_stage_free[2] = {nullptr, nullptr};   // kernels that read staging buffer i have finished
    bool stage_busy[2] = {false, false};
    uint64_t n_host_push = 0;
    uint32_t n_ch = 0, n_ch_pad = 0, max_block = 0;
    int comm_sms = 0;               // SMs left to the caller's concurrent kernels (ua3reo_reserve_sms)
    DdcBuffers b;
    int16_t* adc_stage = nullptr;   // device [max_block + 1024]: carry + new samples
    uint32_t carry = 0;
    size_t last_frames = 0;
    bool pushed = false;
    std::vector<uint32_t> h_fcw;
    std::vector<uint8_t> h_iq_swap;
    uint64_t launches = 0;
    std::vector<void*> allocs;
    std::vector<cudaEvent_t> prof_ev;   // (kDdcKernels + 1) events per profiled block
    uint32_t prof_cap = 0, prof_used = 0;
    uint32_t prof_mask = 0xFFFFFFFFu;   // which of the kProfEvents event points are recorded
    // frame ring bookkeeping (monotonic frame counters; ring index = counter & ring_mask)
    uint64_t w_pos = 0;                 // frames written so far
    uint64_t a_pos = 0, f_pos = 0;      // frames consumed by the audio / FFT stage
    // STM32 stage.  Its kernels run on their own high-priority stream, one push behind the DDC: rx_audio is latency
    // bound and packed into a few whole SMs (rx.cu), so it overlaps the next push's persistent front kernel, which
    // takes its tiles from a counter and simply runs on the SMs that are left.
    cudaStream_t rx_stream = nullptr;
    cudaStream_t rx_stream2 = nullptr;             // FFT_doFFT kernels: independent of the audio kernels, so they run beside them
    cudaEvent_t ev_fft_done = nullptr;
    cudaEvent_t ev_frames = nullptr;               // the frames (and every earlier operation of `stream`) of the push are done
    cudaEvent_t ev_rx_done[2] = {nullptr, nullptr};   // STM32 stage of push k (k & 1) has finished
    bool rx_done_valid[2] = {false, false};
    // results are double buffered (set k & 1 belongs to push k) so that the pipelined reads of push k, which run on
    // copy_stream, overlap the STM32 kernels of push k+1
    int32_t* rx_audio2[2] = {nullptr, nullptr};
    float* rx_cw2[2] = {nullptr, nullptr};
    float* rx_spec2[2] = {nullptr, nullptr};
    uint16_t* rx_wf2[2] = {nullptr, nullptr};
    cudaEvent_t ev_rxcopy[2] = {nullptr, nullptr};    // pipelined reads of result set i have finished
    bool rxcopy_pending[2] = {false, false};
    int rx_last_slot = 0;
    bool rx_on = false, rx_alloc = false;
    RxBuffers rx;
    std::vector<ua3reo_rx_settings> h_set;
    std::vector<RxParams> h_par;
    uint8_t* rx_flags = nullptr;        // device scratch for rx_set's state-clear flags
    float* stage_buf = nullptr;         // device scratch of ua3reo_rx_stage: 2 x 4096 floats
    float* usb_undo = nullptr;          // device scratch of ua3reo_rx_read_audio_usb: per-channel volume undo ...
    int16_t* usb_out = nullptr;         // ... and the packed int16 words of the largest push
    int32_t* adc_stats = nullptr;       // device: min, max, samples at the rails (since the last reset)
    bool adc_stats_on = false;
    size_t last_audio_blocks = 0, last_fft_frames = 0;
    // transmit DUC
    bool duc_alloc = false;
    DucBuffers duc;
    size_t last_tx = 0;
    // transmit audio (processTxAudio)
    bool tx_alloc = false;
    TxBuffers tx;
    std::vector<TxParams> h_txpar;
    uint8_t* tx_flags = nullptr;
    size_t last_tx_blocks = 0;
};

template <class T>
static cudaError_t dev_alloc(ua3reo_ctx* c, T** p, size_t n) {
    cudaError_t e = cudaMalloc((void**)p, n * sizeof(T));
    if (e != cudaSuccess) return e;
    c->allocs.push_back(*p);
    return cudaMemsetAsync(*p, 0, n * sizeof(T), c->stream);
}

// Set-up calls (rx_allocate, ua3reo_duc_enable, ua3reo_tx_enable) allocate many buffers; when one of them fails the
// buffers this call already took are given back, so that a retry does not leak and no half-built stage is left behind.
struct AllocScope {
    ua3reo_ctx* c;
    size_t mark;
    bool ok = false;
    explicit AllocScope(ua3reo_ctx* c_) : c(c_), mark(c_->allocs.size()) {}
    ~AllocScope() {
        if (ok) return;
        cudaStreamSynchronize(c->stream);                 // the memsets of dev_alloc
        while (c->allocs.size() > mark) { cudaFree(c->allocs.back()); c->allocs.pop_back(); }
    }
};

extern "C" {

const char* ua3reo_version(void) { return "ua3reo_b200 0.1 sm_100a"; }
const char* ua3reo_last_error(void) { return g_err.c_str(); }

uint32_t ua3reo_phrase_from_frequency(uint32_t freq, int* iq_swap) {
    // functions.c:206-226, in double as the firmware does
    const uint32_t clk = 49152000u;
    bool inverted = false;
    uint32_t f = freq;
    if (f > clk / 2) {
        while (f > clk / 2) { f -= clk / 2; inverted = !inverted; }
        if (inverted) f = clk / 2 - f;
    }
    if (iq_swap) *iq_swap = inverted ? 1 : 0;
    return (uint32_t)std::round(((double)f / (double)clk) * 4194304.0);
}

static int ctx_free(ua3reo_ctx* c) {
    if (!c) return UA3_OK;
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    if (c->rx_stream) { cudaStreamSynchronize(c->rx_stream); cudaStreamDestroy(c->rx_stream); }
    if (c->rx_stream2) { cudaStreamSynchronize(c->rx_stream2); cudaStreamDestroy(c->rx_stream2); }
    if (c->ev_fft_done) cudaEventDestroy(c->ev_fft_done);
    if (c->ev_frames) cudaEventDestroy(c->ev_frames);
    for (int i = 0; i < 2; ++i) if (c->ev_rx_done[i]) cudaEventDestroy(c->ev_rx_done[i]);
    for (int i = 0; i < 2; ++i) if (c->ev_rxcopy[i]) cudaEventDestroy(c->ev_rxcopy[i]);
    for (void* p : c->allocs) cudaFree(p);
    for (cudaEvent_t e : c->prof_ev) cudaEventDestroy(e);
    if (c->copy_stream) { cudaStreamSynchronize(c->copy_stream); cudaStreamDestroy(c->copy_stream); }
    if (c->h2d_stream) { cudaStreamSynchronize(c->h2d_stream); cudaStreamDestroy(c->h2d_stream); }
    if (c->ev_h2d) cudaEventDestroy(c->ev_h2d);
    for (int i = 0; i < 2; ++i) if (c->ev_stage_free[i]) cudaEventDestroy(c->ev_stage_free[i]);
    if (c->ev_push) cudaEventDestroy(c->ev_push);
    for (int i = 0; i < 2; ++i) if (c->ev_copy[i]) cudaEventDestroy(c->ev_copy[i]);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
    return UA3_OK;
}

int ua3reo_create(int device, uint32_t n_channels, uint32_t max_block_samples, ua3reo_ctx** out) {
    if (!out || n_channels == 0) return fail(UA3_E_INVAL, "ua3reo_create: bad arguments");
    *out = nullptr;
    if (max_block_samples == 0) max_block_samples = 1u << 20;
    if (max_block_samples % UA3_ADC_PER_FRAME) return fail(UA3_E_INVAL, "max_block_samples must be a multiple of 1024");
    int n_dev = 0;
    if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev == 0)
        return fail(UA3_E_NODEV, "no CUDA device: libua3reo_b200 has no CPU fallback");
    if (device < 0 || device >= n_dev) return fail(UA3_E_INVAL, "device index out of range");
    cudaDeviceProp prop;
    UA3_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) return fail(UA3_E_NODEV, "device is not sm_100: kernels are built for sm_100a only");
    UA3_CUDA(cudaSetDevice(device));

    ua3reo_ctx* c = new (std::nothrow) ua3reo_ctx;
    if (!c) return fail(UA3_E_INVAL, "out of host memory");
    c->device = device;
    c->sm_count = prop.multiProcessorCount;
    c->n_ch = n_channels;
    c->n_ch_pad = (n_channels + 31u) & ~31u;
    c->max_block = max_block_samples;
    cudaError_t e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) { delete c; return fail(UA3_E_CUDA, "cudaStreamCreate", e); }
    e = cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->ev_push, cudaEventDisableTiming);
    for (int i = 0; i < 2 && e == cudaSuccess; ++i) e = cudaEventCreateWithFlags(&c->ev_copy[i], cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&c->h2d_stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) {
        int prio_lo = 0, prio_hi = 0;
        e = cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
        if (e == cudaSuccess) e = cudaStreamCreateWithPriority(&c->rx_stream, cudaStreamNonBlocking, prio_hi);
        // the FFT kernels yield to the audio kernels (the longer chain), both go before the DDC
        if (e == cudaSuccess) e = cudaStreamCreateWithPriority(&c->rx_stream2, cudaStreamNonBlocking, prio_hi < prio_lo - 1 ? prio_hi + 1 : prio_hi);
    }
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->ev_fft_done, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->ev_frames, cudaEventDisableTiming);
    for (int i = 0; i < 2 && e == cudaSuccess; ++i) e = cudaEventCreateWithFlags(&c->ev_rx_done[i], cudaEventDisableTiming);
    for (int i = 0; i < 2 && e == cudaSuccess; ++i) e = cudaEventCreateWithFlags(&c->ev_rxcopy[i], cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->ev_h2d, cudaEventDisableTiming);
    for (int i = 0; i < 2 && e == cudaSuccess; ++i) e = cudaEventCreateWithFlags(&c->ev_stage_free[i], cudaEventDisableTiming);
    if (e != cudaSuccess) { ctx_free(c); return fail(UA3_E_CUDA, "stream/event creation", e); }

    DdcBuffers& b = c->b;
    b.n_ch = c->n_ch; b.n_ch_pad = c->n_ch_pad;
    b.max_chunks = max_block_samples / kCicR;
    b.max_frames = max_block_samples / kFrameAdc;
    b.l_ch_stride = (kLHalo + b.max_chunks) * kLRec;
    b.yi_stride = (kYIHalo + b.max_frames + 7u) & ~7u;
    b.yq_stride = (kYQHalo + b.max_frames + 7u) & ~7u;
    uint32_t ring = 1024;
    while (ring < 2u * b.max_frames + 1024u) ring <<= 1;  // two pushes (async reads overlap the next push) + the <= 511 frames the STM32 stage has not consumed
    b.frame_ch_stride = ring;
    b.ring_mask = ring - 1u;
#define UA3_TRY(call) do { e = (call); if (e != cudaSuccess) { ctx_free(c); return fail(UA3_E_CUDA, #call, e); } } while (0)
    UA3_TRY(dev_alloc(c, &b.nco_tab, 2048));
    UA3_TRY(dev_alloc(c, &b.big_tab, (size_t)kNcoBigTabWords));
    UA3_TRY(dev_alloc(c, &b.adc9, (size_t)max_block_samples + 8));
    UA3_TRY(dev_alloc(c, &b.tile_counter, (size_t)4));
    UA3_TRY(ddc_prepare_kernels());
    {
        std::vector<uint32_t> bt((size_t)kNcoBigTabWords);
        build_nco_big_table(bt.data());
        UA3_TRY(cudaMemcpyAsync(b.big_tab, bt.data(), bt.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, c->stream));
        UA3_TRY(cudaStreamSynchronize(c->stream));
    }
#if !defined(UA3_HOST_EMU)
    {
        std::vector<uint32_t> th((size_t)kNcoBigTabWords);
        std::vector<uint8_t> tw((size_t)kTcWeightPlaneBytes);
        build_nco_half_table(th.data());
        build_tc_weight_planes(tw.data(), b.tc_fix);
        UA3_TRY(dev_alloc(c, &b.tab_h, (size_t)kNcoBigTabWords));
        UA3_TRY(dev_alloc(c, &b.tc_w, (size_t)kTcWeightPlaneBytes));
        UA3_TRY(dev_alloc(c, &b.adc_h, (size_t)max_block_samples + 8));
        UA3_TRY(dev_alloc(c, &b.wrap_flag, (size_t)b.max_chunks + 4));
        UA3_TRY(cudaMemcpyAsync(b.tab_h, th.data(), th.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, c->stream));
        UA3_TRY(cudaMemcpyAsync(b.tc_w, tw.data(), tw.size(), cudaMemcpyHostToDevice, c->stream));
        UA3_TRY(cudaStreamSynchronize(c->stream));
    }
#endif
    if (const char* v = std::getenv("UA3REO_TC_ADC_STAGE")) b.tc_adc_stage = std::atoi(v);
    if (const char* v = std::getenv("UA3REO_PDL")) b.pdl = std::atoi(v);
    if (const char* v = std::getenv("UA3REO_FRONT_VARIANT")) b.front_variant = std::atoi(v);   // 1 small table, 2 big table (CUDA cores), 3 tensor cores
    UA3_TRY(dev_alloc(c, &b.fcw, c->n_ch_pad));
    UA3_TRY(dev_alloc(c, &b.phase, c->n_ch_pad));
    UA3_TRY(dev_alloc(c, &b.L, (size_t)c->n_ch_pad * b.l_ch_stride));
    UA3_TRY(dev_alloc(c, &b.YI, (size_t)c->n_ch_pad * b.yi_stride));
    UA3_TRY(dev_alloc(c, &b.YQ, (size_t)c->n_ch_pad * b.yq_stride));
    UA3_TRY(dev_alloc(c, &b.frames, (size_t)c->n_ch * b.frame_ch_stride));
    UA3_TRY(dev_alloc(c, &c->adc_stage, (size_t)max_block_samples + UA3_ADC_PER_FRAME));
    UA3_TRY(dev_alloc(c, &c->adc_stage2[0], (size_t)max_block_samples));
    UA3_TRY(dev_alloc(c, &c->adc_stage2[1], (size_t)max_block_samples));
    uint32_t tab[2048];
    build_nco_table(tab);
    UA3_TRY(cudaMemcpyAsync(b.nco_tab, tab, sizeof tab, cudaMemcpyHostToDevice, c->stream));
    UA3_TRY(ddc_upload_constants());
    c->h_fcw.assign(c->n_ch_pad, 0u);
    for (uint32_t i = 0; i < c->n_ch; ++i) c->h_fcw[i] = 620407u;   // stm32_interface.v:56 power-on freq_out
    c->h_iq_swap.assign(c->n_ch, 0);
    UA3_TRY(cudaMemcpyAsync(b.fcw, c->h_fcw.data(), sizeof(uint32_t) * c->n_ch_pad, cudaMemcpyHostToDevice, c->stream));
    UA3_TRY(cudaStreamSynchronize(c->stream));
#undef UA3_TRY
    *out = c;
    return UA3_OK;
}

int ua3reo_destroy(ua3reo_ctx* ctx) { return ctx_free(ctx); }

static int rx_quiesce(ua3reo_ctx* c);

int ua3reo_reset(ua3reo_ctx* c) {
    if (!c) return fail(UA3_E_INVAL, "null context");
    UA3_CUDA(cudaSetDevice(c->device));
    { const int qrc = rx_quiesce(c); if (qrc != UA3_OK) return qrc; }
    const DdcBuffers& b = c->b;
    UA3_CUDA(cudaMemsetAsync(b.phase, 0, sizeof(uint32_t) * c->n_ch_pad, c->stream));
    UA3_CUDA(cudaMemsetAsync(b.L, 0, sizeof(uint64_t) * (size_t)c->n_ch_pad * b.l_ch_stride, c->stream));
    UA3_CUDA(cudaMemsetAsync(b.YI, 0, sizeof(int16_t) * (size_t)c->n_ch_pad * b.yi_stride, c->stream));
    UA3_CUDA(cudaMemsetAsync(b.YQ, 0, sizeof(int16_t) * (size_t)c->n_ch_pad * b.yq_stride, c->stream));
    c->carry = 0; c->last_frames = 0; c->pushed = false;
    UA3_CUDA(cudaStreamSynchronize(c->copy_stream));
    for (int i = 0; i < 2; ++i) { c->rx_done_valid[i] = false; c->rxcopy_pending[i] = false; c->copy_pending[i] = false; }
    c->w_pos = c->a_pos = c->f_pos = 0;
    c->last_audio_blocks = c->last_fft_frames = 0;
    if (c->rx_alloc) {
        int launches = 0;
        UA3_CUDA(rx_launch_init_state(c->rx, c->stream, &launches));
        c->launches += (uint64_t)launches;
    }
    if (c->duc_alloc) UA3_CUDA(cudaMemsetAsync(c->duc.state, 0, sizeof(DucState) * (size_t)c->n_ch, c->stream));
    if (c->tx_alloc) {
        int launches = 0;
        UA3_CUDA(tx_launch_init_state(c->tx, c->stream, &launches));
        c->launches += (uint64_t)launches;
    }
    c->last_tx = 0;
    return UA3_OK;
}

int ua3reo_ddc_set_clocking(ua3reo_ctx* c, int align_b, int d_i, int d_q) {
    if (!c) return fail(UA3_E_INVAL, "null context");
    if ((align_b != 0 && align_b != 1) || d_i < 0 || d_i > kMaxDI || d_q < 1 || d_q > kYQHalo)
        return fail(UA3_E_INVAL, "ua3reo_ddc_set_clocking: align_b in {0,1}, d_i in 0..3, d_q in 1..130");
    if (c->pushed) return fail(UA3_E_STATE, "ua3reo_ddc_set_clocking: only before the first push or after ua3reo_reset()");
    c->b.align_b = align_b; c->b.d_i = d_i; c->b.d_q = d_q;
    return UA3_OK;
}

int ua3reo_ddc_get_clocking(const ua3reo_ctx* c, int* align_b, int* d_i, int* d_q) {
    if (!c) return fail(UA3_E_INVAL, "null context");
    if (align_b) *align_b = c->b.align_b;
    if (d_i) *d_i = c->b.d_i;
    if (d_q) *d_q = c->b.d_q;
    return UA3_OK;
}

int ua3reo_reserve_sms(ua3reo_ctx* c, int n_sms) {
    if (!c) return fail(UA3_E_INVAL, "null context");
    if (n_sms < 0 || n_sms >= c->sm_count / 2) return fail(UA3_E_INVAL, "ua3reo_reserve_sms: 0 <= n_sms < half the device");
    c->comm_sms = n_sms;
    return UA3_OK;
}

uint32_t ua3reo_n_channels(const ua3reo_ctx* c) { return c ? c->n_ch : 0; }
uint32_t ua3reo_max_block_samples(const ua3reo_ctx* c) { return c ? c->max_block : 0; }
uint64_t ua3reo_launch_count(const ua3reo_ctx* c) { return c ? c->launches : 0; }

int ua3reo_set_fcw(ua3reo_ctx* c, uint32_t first, uint32_t n, const uint32_t* fcw22) {
    if (!c || !fcw22 || first > c->n_ch || n > c->n_ch - first) return fail(UA3_E_INVAL, "ua3reo_set_fcw: range");
    UA3_CUDA(cudaSetDevice(c->device));
    for (uint32_t i = 0; i < n; ++i) c->h_fcw[first + i] = fcw22[i] & 0x3FFFFFu;   // freq_out is 22 bits wide
    UA3_CUDA(cudaMemcpyAsync(c->b.fcw + first, c->h_fcw.data() + first, sizeof(uint32_t) * n, cudaMemcpyHostToDevice,
                             c->stream));
    UA3_CUDA(cudaStreamSynchronize(c->stream));   // h_fcw may be rewritten by the next call
    return UA3_OK;
}

int ua3reo_get_fcw(ua3reo_ctx* c, uint32_t first, uint32_t n, uint32_t* fcw22) {
    if (!c || !fcw22 || first > c->n_ch || n > c->n_ch - first) return fail(UA3_E_INVAL, "ua3reo_get_fcw: range");
    std::memcpy(fcw22, c->h_fcw.data() + first, sizeof(uint32_t) * n);
    return UA3_OK;
}

int ua3reo_set_frequency(ua3reo_ctx* c, uint32_t channel, uint32_t freq_hz) {
    if (!c || channel >= c->n_ch) return fail(UA3_E_INVAL, "ua3reo_set_frequency: channel");
    int swap = 0;
    const uint32_t w = ua3reo_phrase_from_frequency(freq_hz, &swap);
    c->h_iq_swap[channel] = (uint8_t)swap;
    return ua3reo_set_fcw(c, channel, 1, &w);
}

// Control-plane calls and synchronous result reads touch the STM32 stage's buffers through `stream`: they first wait
// for whatever that stage still has in flight on its own stream.  (What they enqueue on `stream` is ordered before
// the next push's STM32 kernels by ev_frames.)
static int rx_quiesce(ua3reo_ctx* c) {
    if (c->rx_stream) UA3_CUDA(cudaStreamSynchronize(c->rx_stream));
    return UA3_OK;
}

// Before a push writes ring slots: the STM32 stage of the push before the previous one may still be reading them.
static int rx_ring_guard(ua3reo_ctx* c) {
    const int slot = (int)(c->n_push & 1);
    if (c->rx_done_valid[slot]) {
        UA3_CUDA(cudaStreamWaitEvent(c->stream, c->ev_rx_done[slot], 0));
        c->rx_done_valid[slot] = false;
    }
    return UA3_OK;
}

// processRxAudio()/FFT_doFFT() over the whole audio blocks / FFT frames that the ring holds, on the STM32 stream.
static int rx_run_stage(ua3reo_ctx* c, cudaEvent_t* ev, int* launches) {
    const uint32_t nb = (uint32_t)((c->w_pos - c->a_pos) / UA3_AUDIO_BLOCK);
    const uint32_t nf = (uint32_t)((c->w_pos - c->f_pos) / UA3_FFT_SIZE);
    const int set = (int)(c->n_push & 1);
    c->rx.audio_out = c->rx_audio2[set]; c->rx.spectra = c->rx_spec2[set]; c->rx.waterfall = c->rx_wf2[set]; c->rx.cw_mag = c->rx_cw2[set];
    c->rx_last_slot = set;
    // processRxAudio and FFT_doFFT touch disjoint state and outputs: their kernels run on two streams side by side
    // (serialised on one stream only while profiling events are being recorded AROUND THEM - not when the profile covers a
    // DDC kernel only, as bench.py's timed region does: round 2 measured a 1024-channel step at 1.05 ms instead of 0.81 because
    // of exactly that)
    const bool prof_rx = ev && (((c->prof_mask >> kDdcKernels) & 7u) != 0u);
    cudaStream_t fft_st = prof_rx ? c->rx_stream : c->rx_stream2;
    UA3_CUDA(cudaEventRecord(c->ev_frames, c->stream));
    UA3_CUDA(cudaStreamWaitEvent(c->rx_stream, c->ev_frames, 0));
    if (fft_st != c->rx_stream) UA3_CUDA(cudaStreamWaitEvent(fft_st, c->ev_frames, 0));
    if (c->rxcopy_pending[set]) {                                // a pipelined read of push k-2 still owns this result set
        UA3_CUDA(cudaStreamWaitEvent(c->rx_stream, c->ev_rxcopy[set], 0));
        if (fft_st != c->rx_stream) UA3_CUDA(cudaStreamWaitEvent(fft_st, c->ev_rxcopy[set], 0));
        c->rxcopy_pending[set] = false;
    }
    UA3_CUDA(rx_launch_audio(c->rx, (uint32_t)(c->a_pos & c->b.ring_mask), nb, c->rx_stream, launches));
    if (ev && ((c->prof_mask >> (kDdcKernels + 1)) & 1u)) cudaEventRecord(ev[kDdcKernels + 1], c->rx_stream);
    UA3_CUDA(rx_launch_fft(c->rx, (uint32_t)(c->f_pos & c->b.ring_mask), nf, fft_st, launches));
    if (ev && ((c->prof_mask >> (kDdcKernels + 2)) & 1u)) cudaEventRecord(ev[kDdcKernels + 2], c->rx_stream);
    if (fft_st != c->rx_stream) {
        UA3_CUDA(cudaEventRecord(c->ev_fft_done, fft_st));
        UA3_CUDA(cudaStreamWaitEvent(c->rx_stream, c->ev_fft_done, 0));
    }
    const int slot = (int)(c->n_push & 1);
    UA3_CUDA(cudaEventRecord(c->ev_rx_done[slot], c->rx_stream));
    c->rx_done_valid[slot] = true;
    c->a_pos += (uint64_t)nb * UA3_AUDIO_BLOCK;
    c->f_pos += (uint64_t)nf * UA3_FFT_SIZE;
    c->last_audio_blocks = nb;
    c->last_fft_frames = nf;
    return UA3_OK;
}

static int push_common(ua3reo_ctx* c, const int16_t* src, size_t n, size_t* frames_out, cudaMemcpyKind kind) {
    if (!c || (!src && n)) return fail(UA3_E_INVAL, "ua3reo_ddc_push: null argument");
    if ((size_t)c->carry + n > (size_t)c->max_block + UA3_ADC_PER_FRAME - 1 ||
        (((size_t)c->carry + n) / UA3_ADC_PER_FRAME) * UA3_ADC_PER_FRAME > c->max_block)
        return fail(UA3_E_TOOBIG, "ua3reo_ddc_push: block exceeds max_block_samples");
    UA3_CUDA(cudaSetDevice(c->device));
    const size_t total = c->carry + n;
    const uint32_t n_proc = (uint32_t)((total / UA3_ADC_PER_FRAME) * UA3_ADC_PER_FRAME);
    const int16_t* proc_src = c->adc_stage;
    const bool in_place = (kind == cudaMemcpyDeviceToDevice) && c->carry == 0 && n_proc == n &&
                          ((uintptr_t)src % 16u) == 0;
    const bool staged = (kind == cudaMemcpyHostToDevice) && c->carry == 0 && n_proc == n && n > 0;
    int stage_slot = -1;
    if (in_place) {
        proc_src = src;
    } else if (staged) {
        // whole-block host push: copy on the H2D stream into the free staging buffer, kernels wait for the copy
        stage_slot = (int)(c->n_host_push++ & 1);
        if (c->stage_busy[stage_slot]) UA3_CUDA(cudaStreamWaitEvent(c->h2d_stream, c->ev_stage_free[stage_slot], 0));
        UA3_CUDA(cudaMemcpyAsync(c->adc_stage2[stage_slot], src, n * sizeof(int16_t), kind, c->h2d_stream));
        UA3_CUDA(cudaEventRecord(c->ev_h2d, c->h2d_stream));
        UA3_CUDA(cudaStreamWaitEvent(c->stream, c->ev_h2d, 0));
        proc_src = c->adc_stage2[stage_slot];
    } else if (n) {
        UA3_CUDA(cudaMemcpyAsync(c->adc_stage + c->carry, src, n * sizeof(int16_t), kind, c->stream));
    }
    // an asynchronous frame read issued after push k-2 still owns the ring slots this push overwrites
    if (c->copy_pending[c->n_push & 1]) {
        UA3_CUDA(cudaStreamWaitEvent(c->stream, c->ev_copy[c->n_push & 1], 0));
        c->copy_pending[c->n_push & 1] = false;
    }
    { const int rc = rx_ring_guard(c); if (rc != UA3_OK) return rc; }
    int launches = 0;
    cudaEvent_t* ev = nullptr;
    if (n_proc && c->prof_used < c->prof_cap) ev = c->prof_ev.data() + (size_t)(c->prof_used++) * kProfEvents;
    if (n_proc && c->adc_stats_on)
        UA3_CUDA(adc_stats_launch(proc_src, n_proc, c->adc_stats, c->sm_count, c->stream, &launches));
    if (n_proc) {
        // With the STM32 stage on, the persistent front kernel leaves a few SMs free for the previous push's STM32 kernels
        // (rx_filter / rx_post / rx_fft on the high-priority rx_stream), so that the two stages always run side by side:
        // a front CTA owns a whole SM (all shared memory and registers), nothing can share one with it.  The STM32 stage
        // needs about 4 % of the machine; it gets the SMs its resident warps can fill, at most kRxReserveSmall / kRxReserveLarge.
        int front_sms = c->sm_count - c->comm_sms;
        if (c->rx_on) {
            const char* env = std::getenv("UA3REO_RX_RESERVE_SMS");
            const int want = rx_audio_sms(c->n_ch);
            // From 2048 channels the stage's CTAs are packed onto whole SMs (rx.cu: rx_audio_kernel<2>) and nothing is set aside:
            // its kernels are launched before the front kernel and take idle SMs, and the FFT kernel that follows them gets the
            // SMs they vacate ahead of the waiting front CTAs (the rx streams have the higher priority).  Measured
            // (tools/gpu/sweep_rx_reserve2.sh) at 4096 channels 2.743 / 2.752 / 2.765 / 2.787 / 2.811 ms per step with 0 / 1 / 2 / 4 /
            // 6 SMs set aside, at 2048 channels 1.406 / 1.412 / 1.417 / 1.423 with 0 / 2 / 4 / 6, at 1024 channels (192-thread CTAs,
            // sweep_rx_reserve3.sh) 0.783 / 0.796 / 0.794 with 0 / 2 / 5.
            const int cap = c->n_ch < 2048u ? kRxReserveSmall : kRxReserveLarge;
            const int reserve = env ? std::atoi(env) : (want < cap ? want : cap);
            if (reserve > 0 && reserve < front_sms) front_sms -= reserve;
        }
        UA3_CUDA(ddc_launch_block(c->b, proc_src, n_proc, (uint32_t)(c->w_pos & c->b.ring_mask), front_sms, c->stream,
                                  &launches, ev, c->prof_mask));
    }
    c->w_pos += n_proc / UA3_ADC_PER_FRAME;
    c->last_audio_blocks = c->last_fft_frames = 0;
    if (c->rx_on) {
        const int rc = rx_run_stage(c, ev, &launches);
        if (rc != UA3_OK) return rc;
    } else {
        c->a_pos = c->f_pos = c->w_pos;
        if (ev && c->prof_mask == 0xFFFFFFFFu) { cudaEventRecord(ev[kDdcKernels + 1], c->stream); cudaEventRecord(ev[kDdcKernels + 2], c->stream); }
    }
    c->launches += (uint64_t)launches;
    UA3_CUDA(cudaEventRecord(c->ev_push, c->stream));
    if (stage_slot >= 0) {
        UA3_CUDA(cudaEventRecord(c->ev_stage_free[stage_slot], c->stream));
        c->stage_busy[stage_slot] = true;
    }
    c->n_push++;
    const uint32_t left = (uint32_t)(total - n_proc);
    if (!in_place && n_proc && left)
        UA3_CUDA(cudaMemcpyAsync(c->adc_stage, c->adc_stage + n_proc, left * sizeof(int16_t), cudaMemcpyDeviceToDevice,
                                 c->stream));
    c->carry = left;
    c->last_frames = n_proc / UA3_ADC_PER_FRAME;
    c->pushed = true;
    if (frames_out) *frames_out = c->last_frames;
    return UA3_OK;
}

int ua3reo_ddc_push(ua3reo_ctx* c, const int16_t* adc_host, size_t n, size_t* frames_out) {
    return push_common(c, adc_host, n, frames_out, cudaMemcpyHostToDevice);
}

int ua3reo_ddc_push_device(ua3reo_ctx* c, const int16_t* adc_dev, size_t n, size_t* frames_out) {
    return push_common(c, adc_dev, n, frames_out, cudaMemcpyDeviceToDevice);
}

static int read_frames_on(ua3reo_ctx* c, uint8_t* dst, size_t n_frames, cudaStream_t st, const char* who) {
    if (!c || (!dst && n_frames)) return fail(UA3_E_INVAL, who);
    if (!c->pushed) return fail(UA3_E_STATE, "ua3reo_ddc_read_frames: no push yet");
    if (n_frames != c->last_frames) return fail(UA3_E_INVAL, "ua3reo_ddc_read_frames: n_frames != frames of last push");
    UA3_CUDA(cudaSetDevice(c->device));
    if (n_frames) {
        const uint32_t ring = c->b.frame_ch_stride;
        const uint32_t first = (uint32_t)((c->w_pos - n_frames) & c->b.ring_mask);
        const size_t n1 = (first + n_frames <= ring) ? n_frames : (size_t)(ring - first);   // up to the wrap
        const size_t pitch = (size_t)ring * UA3_FRAME_BYTES;
        UA3_CUDA(cudaMemcpy2DAsync(dst, n_frames * UA3_FRAME_BYTES, c->b.frames + first, pitch, n1 * UA3_FRAME_BYTES,
                                   c->n_ch, cudaMemcpyDeviceToHost, st));
        if (n1 < n_frames)
            UA3_CUDA(cudaMemcpy2DAsync(dst + n1 * UA3_FRAME_BYTES, n_frames * UA3_FRAME_BYTES, c->b.frames, pitch,
                                       (n_frames - n1) * UA3_FRAME_BYTES, c->n_ch, cudaMemcpyDeviceToHost, st));
    }
    return UA3_OK;
}

int ua3reo_ddc_read_frames(ua3reo_ctx* c, uint8_t* dst, size_t n_frames) {
    const int rc = read_frames_on(c, dst, n_frames, c ? c->stream : nullptr, "ua3reo_ddc_read_frames: null argument");
    if (rc != UA3_OK) return rc;
    UA3_CUDA(cudaStreamSynchronize(c->stream));
    return UA3_OK;
}

int ua3reo_ddc_read_frames_async(ua3reo_ctx* c, uint8_t* dst, size_t n_frames) {
    if (!c) return fail(UA3_E_INVAL, "ua3reo_ddc_read_frames_async: null argument");
    if (c->b.frame_ch_stride < 2u * c->b.max_frames)
        return fail(UA3_E_STATE, "ua3reo_ddc_read_frames_async: frame ring cannot hold two pushes");
    UA3_CUDA(cudaSetDevice(c->device));
    UA3_CUDA(cudaStreamWaitEvent(c->copy_stream, c->ev_push, 0));
    const int rc = read_frames_on(c, dst, n_frames, c->copy_stream, "ua3reo_ddc_read_frames_async: null argument");
    if (rc != UA3_OK) return rc;
    const unsigned slot = (unsigned)((c->n_push + 1) & 1);      // n_push was already advanced by the push being read
    UA3_CUDA(cudaEventRecord(c->ev_copy[slot], c->copy_stream));
    c->copy_pending[slot] = true;
    return UA3_OK;
}

int ua3reo_ddc_frames_device(ua3reo_ctx* c, const uint8_t** base, size_t* first_frame, size_t* n_frames,
                             size_t* ring_frames, size_t* stride) {
    if (!c || !base) return fail(UA3_E_INVAL, "ua3reo_ddc_frames_device: null argument");
    *base = reinterpret_cast<const uint8_t*>(c->b.frames);
    if (first_frame) *first_frame = (size_t)((c->w_pos - c->last_frames) & c->b.ring_mask);
    if (n_frames) *n_frames = c->last_frames;
    if (ring_frames) *ring_frames = c->b.frame_ch_stride;
    if (stride) *stride = (size_t)c->b.frame_ch_stride * UA3_FRAME_BYTES;
    return UA3_OK;
}

// ------------------------------------------------------------------------------------------------
// STM32 stage
// ------------------------------------------------------------------------------------------------
void ua3reo_rx_defaults(ua3reo_rx_settings* s) {
    if (!s) return;
    std::memset(s, 0, sizeof *s);                 // settings.c:33-94
    s->mode = kModeLSB; s->agc = 1; s->agc_speed = 3; s->dnr = 0; s->notch = 0; s->mute = 0; s->volume = 20;
    s->rf_gain = 50; s->fm_sql_threshold = 1; s->fft_enabled = 1; s->fft_averaging = 4; s->fft_zoom = 1; s->iq_swap = 0;
    s->filter_width = 2700; s->ssb_hpf_pass = 300; s->notch_fc = 1000;
}

// slot -> channel permutation that groups channels taking the same branches of rx_filter_kernel / rx_post_kernel (stable sort)
static int rx_upload_order(ua3reo_ctx* c) {
    std::vector<uint32_t> order(c->n_ch);
    for (uint32_t i = 0; i < c->n_ch; ++i) order[i] = i;
    auto key = [&](uint32_t ch) {
        const RxParams& p = c->h_par[ch];
        return ((uint32_t)p.mode << 8) | ((uint32_t)p.dnr_on << 2) | ((uint32_t)p.notch_on << 1) | (uint32_t)p.cw_on;
    };
    std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return key(a) < key(b); });
    UA3_CUDA(cudaMemcpyAsync(c->rx.order, order.data(), sizeof(uint32_t) * c->n_ch, cudaMemcpyHostToDevice, c->stream));
    UA3_CUDA(cudaStreamSynchronize(c->stream));
    return UA3_OK;
}

static int rx_allocate(ua3reo_ctx* c) {
    if (c->rx_alloc) return UA3_OK;
    AllocScope scope(c);
    RxBuffers& r = c->rx;
    r.n_ch = c->n_ch;
    r.frames = c->b.frames; r.ring_mask = c->b.ring_mask; r.frame_ch_stride = c->b.frame_ch_stride;
    r.max_audio_blocks = (c->b.max_frames + UA3_AUDIO_BLOCK - 1) / UA3_AUDIO_BLOCK + 1;
    r.max_fft_frames = c->b.max_frames / UA3_FFT_SIZE + 1;
    r.audio_ch_stride = r.max_audio_blocks * 2 * UA3_AUDIO_BLOCK;
    r.spec_ch_stride = r.max_fft_frames * UA3_FFT_BINS;
    UA3_CUDA(dev_alloc(c, &r.params, (size_t)c->n_ch));
    UA3_CUDA(dev_alloc(c, &r.state, (size_t)c->n_ch));
    UA3_CUDA(dev_alloc(c, &r.order, (size_t)c->n_ch));
    for (int i = 0; i < 2; ++i) {
        UA3_CUDA(dev_alloc(c, &c->rx_audio2[i], (size_t)c->n_ch * r.audio_ch_stride));
        UA3_CUDA(dev_alloc(c, &c->rx_spec2[i], (size_t)c->n_ch * r.spec_ch_stride));
        UA3_CUDA(dev_alloc(c, &c->rx_wf2[i], (size_t)c->n_ch * r.spec_ch_stride));
        UA3_CUDA(dev_alloc(c, &c->rx_cw2[i], (size_t)c->n_ch * r.max_audio_blocks));
    }
    UA3_CUDA(dev_alloc(c, &r.scratch, rx_scratch_floats(c->n_ch, r.max_audio_blocks)));
    if (const char* v = std::getenv("UA3REO_RX_SPLIT")) r.split_audio = std::atoi(v) != 0;
    UA3_CUDA(dev_alloc(c, &r.fft_in, rx_fft_in_floats(c->n_ch, r.max_fft_frames)));
    r.audio_out = c->rx_audio2[0]; r.spectra = c->rx_spec2[0]; r.waterfall = c->rx_wf2[0]; r.cw_mag = c->rx_cw2[0];
    UA3_CUDA(dev_alloc(c, &r.wtf_hist, (size_t)c->n_ch * kWtfRows * kFftBins));
    UA3_CUDA(dev_alloc(c, &r.wtf_head, (size_t)c->n_ch));
    UA3_CUDA(dev_alloc(c, &r.wtf_pending_hz, (size_t)c->n_ch));
    UA3_CUDA(dev_alloc(c, &c->rx_flags, (size_t)c->n_ch));
    std::vector<float> win(kFftSize), tw(2 * kFftSize);
    rx_build_window(win.data());
    rx_build_twiddles(tw.data());
    uint16_t colors[32];
    rx_build_colors(colors);
    float zb[4][20], zf[4][4];
    rx_build_zoom(zb, zf);
    UA3_CUDA(rx_upload_constants(win.data(), tw.data(), colors, &zb[0][0], &zf[0][0]));
    int launches = 0;
    UA3_CUDA(rx_launch_init_state(r, c->stream, &launches));
    c->launches += (uint64_t)launches;
    // every channel starts with the firmware's default settings
    ua3reo_rx_settings d;
    ua3reo_rx_defaults(&d);
    c->h_set.assign(c->n_ch, d);
    c->h_par.assign(c->n_ch, RxParams());
    for (uint32_t i = 0; i < c->n_ch; ++i) {
        std::memset(&c->h_par[i], 0, sizeof(RxParams));
        bool cl, ch;
        if (!rx_derive(d, c->h_par[i], cl, ch)) return fail(UA3_E_INVAL, "internal: default settings rejected");
    }
    UA3_CUDA(cudaMemcpyAsync(r.params, c->h_par.data(), sizeof(RxParams) * c->n_ch, cudaMemcpyHostToDevice, c->stream));
    UA3_CUDA(cudaStreamSynchronize(c->stream));
    const int rc = rx_upload_order(c);
    if (rc != UA3_OK) return rc;
    c->rx_alloc = scope.ok = true;
    return UA3_OK;
}

int ua3reo_rx_enable(ua3reo_ctx* c, int enable) {
    if (!c) return fail(UA3_E_INVAL, "null context");
    UA3_CUDA(cudaSetDevice(c->device));
    { const int qrc = rx_quiesce(c); if (qrc != UA3_OK) return qrc; }
    if (enable) {
        const int rc = rx_allocate(c);
        if (rc != UA3_OK) return rc;
        if (!c->rx_on) c->a_pos = c->f_pos = c->w_pos;     // start consuming from "now"
    }
    c->rx_on = enable != 0;
    return UA3_OK;
}

// live == false: TRX_setMode() + ReinitAudioFilters() + InitNotchFilter() (filter tables reselected, their states cleared).
// live == true : only the fields processRxAudio()/FFT_doFFT() read from TRX on every call; the lattice tables, the
//                notch coefficients and the ZoomFFT decimator stay as the last full set left them, no state is cleared.
static int rx_apply_settings(ua3reo_ctx* c, uint32_t first, uint32_t n, const ua3reo_rx_settings* settings, bool live,
                             const char* who) {
    if (!c || !settings || first > c->n_ch || n > c->n_ch - first) return fail(UA3_E_INVAL, "ua3reo_rx_set: range");
    UA3_CUDA(cudaSetDevice(c->device));
    { const int qrc = rx_quiesce(c); if (qrc != UA3_OK) return qrc; }
    const int rc = rx_allocate(c);
    if (rc != UA3_OK) return rc;
    std::vector<RxParams> np(c->h_par.begin() + first, c->h_par.begin() + first + n);
    std::vector<uint8_t> flags(n, 0);
    for (uint32_t i = 0; i < n; ++i) {
        bool cl = false, chp = false;
        const RxParams old = np[i];
        ua3reo_rx_settings s = settings[i];
        if (live) { s.filter_width = c->h_set[first + i].filter_width; s.ssb_hpf_pass = c->h_set[first + i].ssb_hpf_pass; }
        if (!rx_derive(s, np[i], cl, chp)) {
            char msg[128];
            std::snprintf(msg, sizeof msg, "%s: channel %u: settings outside the firmware's tables", who, first + i);
            return fail(UA3_E_INVAL, msg);
        }
        if (live) {
            std::memcpy(np[i].lpf_k, old.lpf_k, sizeof old.lpf_k); std::memcpy(np[i].lpf_v, old.lpf_v, sizeof old.lpf_v);
            std::memcpy(np[i].hpf_k, old.hpf_k, sizeof old.hpf_k); std::memcpy(np[i].hpf_v, old.hpf_v, sizeof old.hpf_v);
            std::memcpy(np[i].notch, old.notch, sizeof old.notch);
            np[i].hpf_set = old.hpf_set; np[i].fft_zoom = old.fft_zoom;
            np[i].agc_step_up = old.agc_step_up; np[i].agc_step_down = old.agc_step_down;   // InitAGC() is not per call (agc.c:14-19)
            np[i].lpf_on = settings[i].filter_width > 0;          // `if (CurrentVFO()->Filter_Width > 0)` is read per block (audio_processor.c:448)
        } else {
            const bool zoom_changed = np[i].fft_zoom != old.fft_zoom;
            flags[i] = (uint8_t)((cl ? 1 : 0) | (chp ? 2 : 0) | (zoom_changed ? 4 : 0));
        }
    }
    for (uint32_t i = 0; i < n; ++i) {
        c->h_par[first + i] = np[i];
        if (live) {
            const uint16_t fw = c->h_set[first + i].filter_width, hp = c->h_set[first + i].ssb_hpf_pass, nf = c->h_set[first + i].notch_fc;
            const uint8_t zoom = c->h_set[first + i].fft_zoom, agcs = c->h_set[first + i].agc_speed;
            c->h_set[first + i] = settings[i];
            c->h_set[first + i].agc_speed = agcs;
            c->h_set[first + i].filter_width = fw; c->h_set[first + i].ssb_hpf_pass = hp; c->h_set[first + i].notch_fc = nf;
            c->h_set[first + i].fft_zoom = zoom;
        } else {
            c->h_set[first + i] = settings[i];
        }
    }
    UA3_CUDA(cudaMemcpyAsync(c->rx.params + first, c->h_par.data() + first, sizeof(RxParams) * n, cudaMemcpyHostToDevice,
                             c->stream));
    if (!live) {
        UA3_CUDA(cudaMemcpyAsync(c->rx_flags, flags.data(), n, cudaMemcpyHostToDevice, c->stream));
        int launches = 0;
        UA3_CUDA(rx_launch_clear(c->rx, c->rx_flags, first, n, c->stream, &launches));
        c->launches += (uint64_t)launches;
    }
    UA3_CUDA(cudaStreamSynchronize(c->stream));
    return rx_upload_order(c);
}

int ua3reo_rx_set(ua3reo_ctx* c, uint32_t first, uint32_t n, const ua3reo_rx_settings* settings) {
    return rx_apply_settings(c, first, n, settings, false, "ua3reo_rx_set");
}

int ua3reo_rx_set_live(ua3reo_ctx* c, uint32_t first, uint32_t n, const ua3reo_rx_settings* settings) {
    return rx_apply_settings(c, first, n, settings, true, "ua3reo_rx_set_live");
}

// InitNotchFilter() (audio_filters.c:341-346): new biquad coefficients for TRX.NotchFC, states untouched (the firmware's
// arm_biquad_cascade_df2T_init_f32 call there is commented out).
int ua3reo_rx_set_notch(ua3reo_ctx* c, uint32_t first, uint32_t n, const uint16_t* notch_fc) {
    if (!c || !notch_fc || first > c->n_ch || n > c->n_ch - first) return fail(UA3_E_INVAL, "ua3reo_rx_set_notch: range");
    UA3_CUDA(cudaSetDevice(c->device));
    { const int qrc = rx_quiesce(c); if (qrc != UA3_OK) return qrc; }
    const int rc = rx_allocate(c);
    if (rc != UA3_OK) return rc;
    for (uint32_t i = 0; i < n; ++i) {
        rx_notch_coeffs(notch_fc[i], c->h_par[first + i].notch);
        c->h_set[first + i].notch_fc = notch_fc[i];
    }
    UA3_CUDA(cudaMemcpyAsync(c->rx.params + first, c->h_par.data() + first, sizeof(RxParams) * n, cudaMemcpyHostToDevice,
                             c->stream));
    UA3_CUDA(cudaStreamSynchronize(c->stream));
    return UA3_OK;
}

// FFT_Init() (fft.c:185-210) alone: selects the ZoomFFT decimator for TRX.FFT_Zoom and - whenever the zoom is above 1,
// changed or not - clears the biquad and FIR-decimator states; the accumulation buffer FFTInput_ZOOMFFT and the averages
// are left as they are.
int ua3reo_rx_fft_init(ua3reo_ctx* c, uint32_t first, uint32_t n, const uint8_t* fft_zoom) {
    if (!c || !fft_zoom || first > c->n_ch || n > c->n_ch - first) return fail(UA3_E_INVAL, "ua3reo_rx_fft_init: range");
    UA3_CUDA(cudaSetDevice(c->device));
    const int rc = rx_allocate(c);
    if (rc != UA3_OK) return rc;
    { const int qrc = rx_quiesce(c); if (qrc != UA3_OK) return qrc; }
    std::vector<uint8_t> flags(n, 0);
    for (uint32_t i = 0; i < n; ++i) {
        const uint8_t z = fft_zoom[i] ? fft_zoom[i] : 1;
        if (z != 1 && z != 2 && z != 4 && z != 8 && z != 16) return fail(UA3_E_INVAL, "ua3reo_rx_fft_init: zoom must be 1, 2, 4, 8 or 16");
        flags[i] = z > 1 ? 4 : 0;
    }
    for (uint32_t i = 0; i < n; ++i) {
        const uint8_t z = fft_zoom[i] ? fft_zoom[i] : 1;
        c->h_par[first + i].fft_zoom = z;
        c->h_set[first + i].fft_zoom = z;
    }
    UA3_CUDA(cudaMemcpyAsync(c->rx.params + first, c->h_par.data() + first, sizeof(RxParams) * n, cudaMemcpyHostToDevice,
                             c->stream));
    UA3_CUDA(cudaMemcpyAsync(c->rx_flags, flags.data(), n, cudaMemcpyHostToDevice, c->stream));
    int launches = 0;
    UA3_CUDA(rx_launch_clear(c->rx, c->rx_flags, first, n, c->stream, &launches));
    c->launches += (uint64_t)launches;
    UA3_CUDA(cudaStreamSynchronize(c->stream));
    return UA3_OK;
}

// InitAGC() (agc.c:14-19) alone: the AGC step sizes for TRX.Agc_speed; the gain state is untouched.  The firmware calls it
// from the menu / encoder handlers right after changing the speed (encoder.c:140-143, lcd.c:885), not per audio block.
int ua3reo_rx_set_agc_speed(ua3reo_ctx* c, uint32_t first, uint32_t n, const uint8_t* agc_speed) {
    if (!c || !agc_speed || first > c->n_ch || n > c->n_ch - first) return fail(UA3_E_INVAL, "ua3reo_rx_set_agc_speed: range");
    UA3_CUDA(cudaSetDevice(c->device));
    const int rc = rx_allocate(c);
    if (rc != UA3_OK) return rc;
    { const int qrc = rx_quiesce(c); if (qrc != UA3_OK) return qrc; }
    for (uint32_t i = 0; i < n; ++i)
        if (agc_speed[i] == 0) return fail(UA3_E_INVAL, "ua3reo_rx_set_agc_speed: speed 0 (the firmware would divide by zero)");
    for (uint32_t i = 0; i < n; ++i) {
        c->h_par[first + i].agc_step_up = 500.0f / (float)agc_speed[i];
        c->h_par[first + i].agc_step_down = c->h_par[first + i].agc_step_up / 10.0f;
        c->h_set[first + i].agc_speed = agc_speed[i];
    }
    UA3_CUDA(cudaMemcpyAsync(c->rx.params + first, c->h_par.data() + first, sizeof(RxParams) * n, cudaMemcpyHostToDevice,
                             c->stream));
    UA3_CUDA(cudaStreamSynchronize(c->stream));
    return UA3_OK;
}

// The STM32 stage alone, over I/Q frames that come from elsewhere (a real FPGA, a recording, BASELINE config 1's
// synthetic 48 kSPS I/Q): the frames enter the same per-channel ring the DDC writes.
int ua3reo_rx_push_frames(ua3reo_ctx* c, const uint8_t* frames_host, size_t n) {
    if (!c || (!frames_host && n)) return fail(UA3_E_INVAL, "ua3reo_rx_push_frames: null argument");
    if (!c->rx_on) return fail(UA3_E_STATE, "ua3reo_rx_push_frames: call ua3reo_rx_enable first");
    if (n > c->b.max_frames) return fail(UA3_E_TOOBIG, "ua3reo_rx_push_frames: more frames than one block holds");
    UA3_CUDA(cudaSetDevice(c->device));
    if (c->copy_pending[c->n_push & 1]) {
        UA3_CUDA(cudaStreamWaitEvent(c->stream, c->ev_copy[c->n_push & 1], 0));
        c->copy_pending[c->n_push & 1] = false;
    }
    { const int rc = rx_ring_guard(c); if (rc != UA3_OK) return rc; }
    const uint32_t ring = c->b.frame_ch_stride;
    const uint32_t first = (uint32_t)(c->w_pos & c->b.ring_mask);
    const size_t n1 = (first + n <= ring) ? n : (size_t)(ring - first);
    const size_t pitch = (size_t)ring * UA3_FRAME_BYTES;
    uint8_t* base = reinterpret_cast<uint8_t*>(c->b.frames);
    if (n1)
        UA3_CUDA(cudaMemcpy2DAsync(base + (size_t)first * UA3_FRAME_BYTES, pitch, frames_host, n * UA3_FRAME_BYTES,
                                   n1 * UA3_FRAME_BYTES, c->n_ch, cudaMemcpyHostToDevice, c->stream));
    if (n1 < n)
        UA3_CUDA(cudaMemcpy2DAsync(base, pitch, frames_host + n1 * UA3_FRAME_BYTES, n * UA3_FRAME_BYTES,
                                   (n - n1) * UA3_FRAME_BYTES, c->n_ch, cudaMemcpyHostToDevice, c->stream));
    c->w_pos += n;
    int launches = 0;
    { const int rc = rx_run_stage(c, nullptr, &launches); if (rc != UA3_OK) return rc; }
    c->last_frames = n;
    c->pushed = true;
    c->launches += (uint64_t)launches;
    UA3_CUDA(cudaEventRecord(c->ev_push, c->stream));
    c->n_push++;
    return UA3_OK;
}

int ua3reo_rx_counts(ua3reo_ctx* c, size_t* audio_blocks, size_t* fft_frames) {
    if (!c) return fail(UA3_E_INVAL, "null context");
    if (audio_blocks) *audio_blocks = c->last_audio_blocks;
    if (fft_frames) *fft_frames = c->last_fft_frames;
    return UA3_OK;
}

int ua3reo_rx_read_audio(ua3reo_ctx* c, int32_t* dst, size_t n_blocks) {
    if (!c || (!dst && n_blocks)) return fail(UA3_E_INVAL, "ua3reo_rx_read_audio: null argument");
    if (!c->rx_alloc) return fail(UA3_E_STATE, "ua3reo_rx_read_audio: STM32 stage not enabled");
    if (n_blocks != c->last_audio_blocks) return fail(UA3_E_INVAL, "ua3reo_rx_read_audio: n_blocks != blocks of last push");
    UA3_CUDA(cudaSetDevice(c->device));
    { const int qrc = rx_quiesce(c); if (qrc != UA3_OK) return qrc; }
    const size_t row = n_blocks * 2 * UA3_AUDIO_BLOCK * sizeof(int32_t);
    if (n_blocks)
        UA3_CUDA(cudaMemcpy2DAsync(dst, row, c->rx.audio_out, (size_t)c->rx.audio_ch_stride * sizeof(int32_t), row, c->n_ch,
                                   cudaMemcpyDeviceToHost, c->stream));
    UA3_CUDA(cudaStreamSynchronize(c->stream));
    return UA3_OK;
}

int ua3reo_rx_read_spectra(ua3reo_ctx* c, float* dst, size_t n_frames) {
    if (!c || (!dst && n_frames)) return fail(UA3_E_INVAL, "ua3reo_rx_read_spectra: null argument");
    if (!c->rx_alloc) return fail(UA3_E_STATE, "ua3reo_rx_read_spectra: STM32 stage not enabled");
    if (n_frames != c->last_fft_frames) return fail(UA3_E_INVAL, "ua3reo_rx_read_spectra: n_frames != FFT frames of last push");
    UA3_CUDA(cudaSetDevice(c->device));
    { const int qrc = rx_quiesce(c); if (qrc != UA3_OK) return qrc; }
    const size_t row = n_frames * UA3_FFT_BINS * sizeof(float);
    if (n_frames)
        UA3_CUDA(cudaMemcpy2DAsync(dst, row, c->rx.spectra, (size_t)c->rx.spec_ch_stride * sizeof(float), row, c->n_ch,
                                   cudaMemcpyDeviceToHost, c->stream));
    UA3_CUDA(cudaStreamSynchronize(c->stream));
    return UA3_OK;
}

// Pipelined result reads: the copy is enqueued on the copy stream behind the STM32 stage of the last push and the
// call returns at once; results are double buffered, so the next push's DDC *and* STM32 kernels overlap the copy (the
// push after that waits for it before reusing the result set).  dst should be pinned host memory; ua3reo_sync() (or ua3reo_rx_sync) before it is read.
int ua3reo_rx_read_audio_async(ua3reo_ctx* c, int32_t* dst, size_t n_blocks) {
    if (!c || (!dst && n_blocks)) return fail(UA3_E_INVAL, "ua3reo_rx_read_audio_async: null argument");
    if (!c->rx_alloc) return fail(UA3_E_STATE, "ua3reo_rx_read_audio_async: STM32 stage not enabled");
    if (n_blocks != c->last_audio_blocks) return fail(UA3_E_INVAL, "ua3reo_rx_read_audio_async: n_blocks != blocks of last push");
    UA3_CUDA(cudaSetDevice(c->device));
    const size_t row = n_blocks * 2 * UA3_AUDIO_BLOCK * sizeof(int32_t);
    const int set = c->rx_last_slot;
    UA3_CUDA(cudaStreamWaitEvent(c->copy_stream, c->ev_rx_done[set], 0));
    if (n_blocks)
        UA3_CUDA(cudaMemcpy2DAsync(dst, row, c->rx.audio_out, (size_t)c->rx.audio_ch_stride * sizeof(int32_t), row, c->n_ch,
                                   cudaMemcpyDefault, c->copy_stream));
    UA3_CUDA(cudaEventRecord(c->ev_rxcopy[set], c->copy_stream));
    c->rxcopy_pending[set] = true;
    return UA3_OK;
}

int ua3reo_rx_read_spectra_async(ua3reo_ctx* c, float* dst, size_t n_frames) {
    if (!c || (!dst && n_frames)) return fail(UA3_E_INVAL, "ua3reo_rx_read_spectra_async: null argument");
    if (!c->rx_alloc) return fail(UA3_E_STATE, "ua3reo_rx_read_spectra_async: STM32 stage not enabled");
    if (n_frames != c->last_fft_frames) return fail(UA3_E_INVAL, "ua3reo_rx_read_spectra_async: n_frames != FFT frames of last push");
    UA3_CUDA(cudaSetDevice(c->device));
    const size_t row = n_frames * UA3_FFT_BINS * sizeof(float);
    const int set = c->rx_last_slot;
    UA3_CUDA(cudaStreamWaitEvent(c->copy_stream, c->ev_rx_done[set], 0));
    if (n_frames)
        UA3_CUDA(cudaMemcpy2DAsync(dst, row, c->rx.spectra, (size_t)c->rx.spec_ch_stride * sizeof(float), row, c->n_ch,
                                   cudaMemcpyDefault, c->copy_stream));
    UA3_CUDA(cudaEventRecord(c->ev_rxcopy[set], c->copy_stream));
    c->rxcopy_pending[set] = true;
    return UA3_OK;
}

int ua3reo_rx_read_audio_usb(ua3reo_ctx* c, int16_t* dst, size_t n_blocks) {
    if (!c || (!dst && n_blocks)) return fail(UA3_E_INVAL, "ua3reo_rx_read_audio_usb: null argument");
    if (!c->rx_alloc) return fail(UA3_E_STATE, "ua3reo_rx_read_audio_usb: STM32 stage not enabled");
    if (n_blocks != c->last_audio_blocks) return fail(UA3_E_INVAL, "ua3reo_rx_read_audio_usb: n_blocks != blocks of last push");
    UA3_CUDA(cudaSetDevice(c->device));
    { const int qrc = rx_quiesce(c); if (qrc != UA3_OK) return qrc; }
    if (!n_blocks) return UA3_OK;
    const size_t n_words = n_blocks * 2 * UA3_AUDIO_BLOCK;
    // audio_processor.c:420 undoes the volume with 1 / volume * 100.  Volume 0 makes that inf and the word 0 * inf = NaN, whose
    // conversion to int16 is undefined in C; here a muted-by-volume channel reads as silence.
    std::vector<float> undo(c->n_ch);
    for (uint32_t i = 0; i < c->n_ch; ++i) undo[i] = c->h_set[i].volume ? 1.0f / (float)c->h_set[i].volume * 100.0f : 0.0f;
    // scratch of the call, allocated once for the largest push (the firmware shim calls this for every 192-sample block:
    // a cudaMalloc / cudaFree pair per call would synchronise the device each time)
    if (!c->usb_undo) {
        AllocScope scope(c);
        float* u = nullptr;
        int16_t* o = nullptr;
        UA3_CUDA(dev_alloc(c, &u, (size_t)c->n_ch));
        UA3_CUDA(dev_alloc(c, &o, (size_t)c->n_ch * c->rx.max_audio_blocks * 2 * UA3_AUDIO_BLOCK));
        c->usb_undo = u; c->usb_out = o;
        scope.ok = true;
    }
    float* undo_dev = c->usb_undo;
    int16_t* out_dev = c->usb_out;
    int launches = 0;
    cudaError_t e = cudaMemcpyAsync(undo_dev, undo.data(), sizeof(float) * c->n_ch, cudaMemcpyHostToDevice, c->stream);
    if (e == cudaSuccess) e = rx_launch_usb_pack(c->rx, (uint32_t)n_blocks, undo_dev, out_dev, c->stream, &launches);
    if (e == cudaSuccess) e = cudaMemcpyAsync(dst, out_dev, sizeof(int16_t) * c->n_ch * n_words, cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    c->launches += (uint64_t)launches;
    if (e != cudaSuccess) return fail(UA3_E_CUDA, "ua3reo_rx_read_audio_usb", e);
    return UA3_OK;
}

int ua3reo_rx_read_waterfall(ua3reo_ctx* c, uint16_t* dst, size_t n_frames) {
    if (!c || (!dst && n_frames)) return fail(UA3_E_INVAL, "ua3reo_rx_read_waterfall: null argument");
    if (!c->rx_alloc) return fail(UA3_E_STATE, "ua3reo_rx_read_waterfall: STM32 stage not enabled");
    if (n_frames != c->last_fft_frames) return fail(UA3_E_INVAL, "ua3reo_rx_read_waterfall: n_frames != FFT frames of last push");
    UA3_CUDA(cudaSetDevice(c->device));
    { const int qrc = rx_quiesce(c); if (qrc != UA3_OK) return qrc; }
    const size_t row = n_frames * UA3_FFT_BINS * sizeof(uint16_t);
    if (n_frames)
        UA3_CUDA(cudaMemcpy2DAsync(dst, row, c->rx.waterfall, (size_t)c->rx.spec_ch_stride * sizeof(uint16_t), row, c->n_ch,
                                   cudaMemcpyDeviceToHost, c->stream));
    UA3_CUDA(cudaStreamSynchronize(c->stream));
    return UA3_OK;
}

// CurrentVFO()->Freq - currentFFTFreq of FFT_printFFT() (fft.c:347-351): recorded per channel and applied by the next
// FFT frame between its averaging and its row, where the firmware's display pass applies it.
int ua3reo_rx_move_waterfall(ua3reo_ctx* c, uint32_t first, uint32_t n, const int32_t* freq_diff_hz) {
    if (!c || !freq_diff_hz || first > c->n_ch || n > c->n_ch - first) return fail(UA3_E_INVAL, "ua3reo_rx_move_waterfall: range");
    if (!c->rx_alloc) return fail(UA3_E_STATE, "ua3reo_rx_move_waterfall: STM32 stage not enabled");
    UA3_CUDA(cudaSetDevice(c->device));
    { const int qrc = rx_quiesce(c); if (qrc != UA3_OK) return qrc; }
    // accumulate on the host copy so that two retunes between FFT frames add up like Freq - currentFFTFreq does
    std::vector<int32_t> pend(n);
    UA3_CUDA(cudaMemcpyAsync(pend.data(), c->rx.wtf_pending_hz + first, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, c->stream));
    UA3_CUDA(cudaStreamSynchronize(c->stream));
    for (uint32_t i = 0; i < n; ++i) pend[i] += freq_diff_hz[i];
    UA3_CUDA(cudaMemcpyAsync(c->rx.wtf_pending_hz + first, pend.data(), sizeof(int32_t) * n, cudaMemcpyHostToDevice, c->stream));
    UA3_CUDA(cudaStreamSynchronize(c->stream));
    return UA3_OK;
}

// wtf_buffer (fft.c:29) of every channel in the firmware's row order: row 0 is the newest.
int ua3reo_rx_read_waterfall_history(ua3reo_ctx* c, uint16_t* dst) {
    if (!c || !dst) return fail(UA3_E_INVAL, "ua3reo_rx_read_waterfall_history: null argument");
    if (!c->rx_alloc) return fail(UA3_E_STATE, "ua3reo_rx_read_waterfall_history: STM32 stage not enabled");
    UA3_CUDA(cudaSetDevice(c->device));
    { const int qrc = rx_quiesce(c); if (qrc != UA3_OK) return qrc; }
    const size_t per_ch = (size_t)kWtfRows * kFftBins;
    std::vector<uint16_t> ring((size_t)c->n_ch * per_ch);
    std::vector<uint32_t> head(c->n_ch);
    UA3_CUDA(cudaMemcpyAsync(ring.data(), c->rx.wtf_hist, ring.size() * sizeof(uint16_t), cudaMemcpyDeviceToHost, c->stream));
    UA3_CUDA(cudaMemcpyAsync(head.data(), c->rx.wtf_head, head.size() * sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
    UA3_CUDA(cudaStreamSynchronize(c->stream));
    for (uint32_t ch = 0; ch < c->n_ch; ++ch)
        for (int y = 0; y < kWtfRows; ++y)
            std::memcpy(dst + ch * per_ch + (size_t)y * kFftBins,
                        ring.data() + ch * per_ch + (size_t)((head[ch] + (uint32_t)y) % kWtfRows) * kFftBins, kFftBins * sizeof(uint16_t));
    return UA3_OK;
}

int ua3reo_rx_read_cw(ua3reo_ctx* c, float* dst, size_t n_blocks) {
    if (!c || (!dst && n_blocks)) return fail(UA3_E_INVAL, "ua3reo_rx_read_cw: null argument");
    if (!c->rx_alloc) return fail(UA3_E_STATE, "ua3reo_rx_read_cw: STM32 stage not enabled");
    if (n_blocks != c->last_audio_blocks) return fail(UA3_E_INVAL, "ua3reo_rx_read_cw: n_blocks != blocks of last push");
    UA3_CUDA(cudaSetDevice(c->device));
    { const int qrc = rx_quiesce(c); if (qrc != UA3_OK) return qrc; }
    const size_t row = n_blocks * sizeof(float);
    if (n_blocks)
        UA3_CUDA(cudaMemcpy2DAsync(dst, row, c->rx.cw_mag, (size_t)c->rx.max_audio_blocks * sizeof(float), row, c->n_ch,
                                   cudaMemcpyDeviceToHost, c->stream));
    UA3_CUDA(cudaStreamSynchronize(c->stream));
    return UA3_OK;
}

static int adc_stats_reset(ua3reo_ctx* c) {
    const int32_t init[3] = {2000, -2000, 0};          // stm32_interface.v:385-389
    UA3_CUDA(cudaMemcpyAsync(c->adc_stats, init, sizeof init, cudaMemcpyHostToDevice, c->stream));
    UA3_CUDA(cudaStreamSynchronize(c->stream));
    return UA3_OK;
}

int ua3reo_adc_stats(ua3reo_ctx* c, int16_t* adc_min, int16_t* adc_max, uint32_t* n_rail, int reset) {
    if (!c) return fail(UA3_E_INVAL, "null context");
    UA3_CUDA(cudaSetDevice(c->device));
    if (!c->adc_stats) {
        UA3_CUDA(dev_alloc(c, &c->adc_stats, 4));
        const int rc = adc_stats_reset(c);
        if (rc != UA3_OK) return rc;
        c->adc_stats_on = true;                        // tracking starts with the first call
    }
    int32_t h[3];
    UA3_CUDA(cudaMemcpyAsync(h, c->adc_stats, sizeof h, cudaMemcpyDeviceToHost, c->stream));
    UA3_CUDA(cudaStreamSynchronize(c->stream));
    if (adc_min) *adc_min = (int16_t)h[0];
    if (adc_max) *adc_max = (int16_t)h[1];
    if (n_rail) *n_rail = (uint32_t)h[2];
    return reset ? adc_stats_reset(c) : UA3_OK;
}

// The five bytes the FPGA puts on the bus for command 2 (stm32_interface.v:172-205) and what FPGA_fpgadata_getparam()
// (fpga.c:222-284) makes of them.  ADC_OTR: the AD9226 flag is a pin state; here it reports "a sample sat at either
// rail since the last read".  Front-panel keys and encoder bits are 0 (no front panel behind this library).
int ua3reo_get_params(ua3reo_ctx* c, uint8_t packet[5], int16_t* adc_min_amplitude, int16_t* adc_max_amplitude, int dac_otr) {
    if (!c) return fail(UA3_E_INVAL, "null context");
    int16_t mn = 0, mx = 0;
    uint32_t rail = 0;
    const int rc = ua3reo_adc_stats(c, &mn, &mx, &rail, /*reset=*/1);      // ADC_MINMAX_RESET=1 at k==203
    if (rc != UA3_OK) return rc;
    const uint32_t min12 = (uint32_t)mn & 0xFFFu, max12 = (uint32_t)mx & 0xFFFu;
    uint8_t b[5];
    b[0] = (uint8_t)((rail ? 1u : 0u) | (dac_otr ? 2u : 0u));
    b[1] = (uint8_t)(((min12 >> 8) << 4) | (max12 >> 8));
    b[2] = (uint8_t)(min12 & 0xFFu);
    b[3] = (uint8_t)(max12 & 0xFFu);
    // k == 204 assigns DATA_BUS_OUT[4:0] only (encoder = 0, no key): bits 7:5 still carry those of the byte before
    // (found by running fpga.c against the executed stm32_interface.v; tests/golden/bus_cases.npz)
    b[4] = (uint8_t)(b[3] & 0xE0u);
    if (packet) std::memcpy(packet, b, 5);
    // decode as the firmware does: the minimum is sign-extended through <<4, /16; the maximum is NOT (fpga.c:247-270)
    int16_t dmin = 0, dmax = 0;
    dmin = (int16_t)(dmin | ((b[1] & 0xF0) << 4));
    dmax = (int16_t)(dmax | ((b[1] & 0x0F) << 8));
    dmin = (int16_t)(dmin | b[2]);
    dmin = (int16_t)(dmin << 4);
    dmin = (int16_t)(dmin / 16);
    dmax = (int16_t)(dmax | b[3]);
    if (adc_min_amplitude) *adc_min_amplitude = dmin;
    if (adc_max_amplitude) *adc_max_amplitude = dmax;
    return UA3_OK;
}

// TRX_DoAutoGain() (trx_manager.c:268-356): one call of the input-stage state machine (stm32f4xx_it.c:422).
void ua3reo_autogain_init(ua3reo_autogain* st) { if (st) std::memset(st, 0, sizeof *st); }

void ua3reo_autogain_step(ua3reo_autogain* st, int16_t adc_max_amplitude) {
    if (!st) return;
    const double kLimit = 1100;                                   // AUTOGAIN_MAX_AMPLITUDE (settings.h:25)
    const double att = std::pow(10.0, 12 / 20.0);                 // db2rateV(ATT_DB)         (functions.c:262-265, settings.h:23)
    const double pre = std::pow(10.0, 20 / 20.0);                 // db2rateV(PREAMP_GAIN_DB) (settings.h:24)
    const uint8_t kWait = 7;                                      // AUTOGAIN_CORRECTOR_WAITSTEP (trx_manager.h:9)
    const double a = (double)adc_max_amplitude;
    st->lpf = 1; st->bpf = 1;
    auto settle = [&](bool can_raise) {
        if (can_raise) st->wait++; else st->wait = 0;
        if (st->wait >= kWait) { st->stage++; st->wait = 0; }
    };
    switch (st->stage) {
        case 0: st->preamp = 0; st->att = 1; st->stage++; st->wait = 0; break;               // -12 dB
        case 1: settle(a * att <= kLimit); break;
        case 2: st->preamp = 0; st->att = 0; st->stage++; st->wait = 0; break;               // 0 dB
        case 3: if (a > kLimit) st->stage -= 3; settle(a * pre / att <= kLimit); break;
        case 4: st->preamp = 1; st->att = 1; st->stage++; st->wait = 0; break;               // +8 dB
        case 5: if (a > kLimit) st->stage -= 3; settle(a * att <= kLimit); break;
        case 6: st->preamp = 1; st->att = 0; st->stage++; st->wait = 0; break;               // +20 dB
        case 7: if (a > kLimit) st->stage -= 3; break;
        default: st->stage = 0; break;
    }
}

#include "cw_host.cpp.inc"

int16_t ua3reo_smeter_dbm(float sample_max, float sample_min, uint8_t rf_gain) {
    // stm32f4xx_it.c:398-407, settings.h:11,19-21 (ADC_BITS 12, FPGA_BUS_BITS 16, ADC_VREF 1.0, ratio 4, calibration 0.2)
    float vpp = (sample_max / (float)rf_gain) - (sample_min / (float)rf_gain);
    for (int i = 0; i < (16 - 12); i++) vpp = vpp / 2;
    const float adc_vpp = vpp * 1.0f / ((float)std::pow(2.0, 12) - 1);
    const float vrms = adc_vpp * 0.3535f;
    float rf_in = (vrms / 4) * 0.2f;
    if (rf_in < 0.0000001f) rf_in = 0.0000001f;
    // log10f_fast (functions.c:247-262)
    const float X = (rf_in * rf_in) / (50.0f * 0.001f);
    int E;
    const float F = std::frexp(std::fabs(X), &E);
    float Y = 1.23149591368684f;
    Y *= F; Y += -4.11852516267426f;
    Y *= F; Y += 6.02197014179219f;
    Y *= F; Y += -3.13396450166353f;
    Y += E;
    return (int16_t)(10 * (Y * 0.3010299956639812f));
}

int ua3reo_rx_stage(ua3reo_ctx* c, uint32_t channel, int stage, float* buf_host, float* out_host, size_t n, int arg) {
    if (!c || !buf_host || channel >= c->n_ch) return fail(UA3_E_INVAL, "ua3reo_rx_stage: bad arguments");
    if (stage < UA3_STAGE_DC_FILTER || stage > UA3_STAGE_DNR) return fail(UA3_E_INVAL, "ua3reo_rx_stage: unknown stage");
    if (n == 0 || n > 4096) return fail(UA3_E_INVAL, "ua3reo_rx_stage: 1..4096 samples");
    if (stage == UA3_STAGE_DC_FILTER && (arg < 0 || arg > 5)) return fail(UA3_E_INVAL, "ua3reo_rx_stage: dc_filter stateNum is 0..5");
    if (stage == UA3_STAGE_DNR && (n != 64 || !out_host)) return fail(UA3_E_INVAL, "ua3reo_rx_stage: the DNR takes 64 samples and an output buffer");
    UA3_CUDA(cudaSetDevice(c->device));
    { const int rc = rx_allocate(c); if (rc != UA3_OK) return rc; }
    { const int qrc = rx_quiesce(c); if (qrc != UA3_OK) return qrc; }
    if (!c->stage_buf) UA3_CUDA(dev_alloc(c, &c->stage_buf, (size_t)2 * 4096));
    UA3_CUDA(cudaMemcpyAsync(c->stage_buf, buf_host, n * sizeof(float), cudaMemcpyHostToDevice, c->stream));
    if (stage == UA3_STAGE_DNR)                     // DNR off: the firmware returns without touching bufferOut
        UA3_CUDA(cudaMemcpyAsync(c->stage_buf + 4096, out_host, n * sizeof(float), cudaMemcpyHostToDevice, c->stream));
    int launches = 0;
    UA3_CUDA(rx_launch_stage(c->rx, stage, c->stage_buf, c->stage_buf + 4096, (uint32_t)n, channel, arg, c->stream, &launches));
    c->launches += (uint64_t)launches;
    if (stage == UA3_STAGE_DNR)
        UA3_CUDA(cudaMemcpyAsync(out_host, c->stage_buf + 4096, n * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    else
        UA3_CUDA(cudaMemcpyAsync(buf_host, c->stage_buf, n * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    UA3_CUDA(cudaStreamSynchronize(c->stream));
    return UA3_OK;
}

int ua3reo_rx_read_smeter(ua3reo_ctx* c, float* dst, int reset) {
    if (!c || !dst) return fail(UA3_E_INVAL, "ua3reo_rx_read_smeter: null argument");
    if (!c->rx_alloc) return fail(UA3_E_STATE, "ua3reo_rx_read_smeter: STM32 stage not enabled");
    UA3_CUDA(cudaSetDevice(c->device));
    { const int qrc = rx_quiesce(c); if (qrc != UA3_OK) return qrc; }
    const size_t off = offsetof(RxState, smeter_max);
    UA3_CUDA(cudaMemcpy2DAsync(dst, 2 * sizeof(float), reinterpret_cast<const uint8_t*>(c->rx.state) + off, sizeof(RxState),
                               2 * sizeof(float), c->n_ch, cudaMemcpyDeviceToHost, c->stream));
    if (reset)
        UA3_CUDA(cudaMemset2DAsync(reinterpret_cast<uint8_t*>(c->rx.state) + off, sizeof(RxState), 0, 2 * sizeof(float),
                                   c->n_ch, c->stream));
    UA3_CUDA(cudaStreamSynchronize(c->stream));
    return UA3_OK;
}

int ua3reo_sync(ua3reo_ctx* c) {
    if (!c) return fail(UA3_E_INVAL, "null context");
    UA3_CUDA(cudaSetDevice(c->device));
    UA3_CUDA(cudaStreamSynchronize(c->h2d_stream));
    UA3_CUDA(cudaStreamSynchronize(c->stream));
    UA3_CUDA(cudaStreamSynchronize(c->rx_stream));
    UA3_CUDA(cudaStreamSynchronize(c->copy_stream));
    return UA3_OK;
}

int ua3reo_copy_stream(ua3reo_ctx* c, void** stream) {
    if (!c || !stream) return fail(UA3_E_INVAL, "null argument");
    *stream = (void*)c->copy_stream;
    return UA3_OK;
}

int ua3reo_stream(ua3reo_ctx* c, void** stream) {
    if (!c || !stream) return fail(UA3_E_INVAL, "null argument");
    *stream = (void*)c->stream;
    return UA3_OK;
}

// ------------------------------------------------------------------------------------------------
// transmit DUC
// ------------------------------------------------------------------------------------------------
int ua3reo_duc_enable(ua3reo_ctx* c, uint32_t max_tx_samples) {
    if (!c || max_tx_samples == 0) return fail(UA3_E_INVAL, "ua3reo_duc_enable: bad arguments");
    if (c->duc_alloc) return (max_tx_samples <= c->duc.max_in) ? UA3_OK : fail(UA3_E_STATE, "ua3reo_duc_enable: already enabled with a smaller block");
    UA3_CUDA(cudaSetDevice(c->device));
    AllocScope scope(c);
    DucBuffers& d = c->duc;
    d.n_ch = c->n_ch; d.max_in = max_tx_samples; d.fcw = c->b.fcw;
    {
        int16_t* tab_dev = nullptr;
        UA3_CUDA(dev_alloc(c, &tab_dev, (size_t)kNcoBigTabWords));
        std::vector<int16_t> tab(kNcoBigTabWords);
        build_duc_nco_table(tab.data());
        UA3_CUDA(cudaMemcpyAsync(tab_dev, tab.data(), tab.size() * sizeof(int16_t), cudaMemcpyHostToDevice, c->stream));
        UA3_CUDA(cudaStreamSynchronize(c->stream));
        d.nco_tab = tab_dev;
    }
    UA3_CUDA(dev_alloc(c, &d.state, (size_t)c->n_ch));
    UA3_CUDA(dev_alloc(c, &d.iq_in, (size_t)c->n_ch * max_tx_samples * 2));
    UA3_CUDA(dev_alloc(c, &d.dac, (size_t)c->n_ch * max_tx_samples * 1024));
    UA3_CUDA(duc_upload_constants());
    UA3_CUDA(cudaStreamSynchronize(c->stream));
    c->duc_alloc = scope.ok = true;
    return UA3_OK;
}

int ua3reo_duc_push(ua3reo_ctx* c, const int16_t* iq_host, size_t n) {
    if (!c || (!iq_host && n)) return fail(UA3_E_INVAL, "ua3reo_duc_push: null argument");
    if (!c->duc_alloc) return fail(UA3_E_STATE, "ua3reo_duc_push: call ua3reo_duc_enable first");
    if (n > c->duc.max_in) return fail(UA3_E_TOOBIG, "ua3reo_duc_push: more samples than max_tx_samples");
    UA3_CUDA(cudaSetDevice(c->device));
    if (n)
        UA3_CUDA(cudaMemcpy2DAsync(c->duc.iq_in, (size_t)c->duc.max_in * 2 * sizeof(int16_t), iq_host, n * 2 * sizeof(int16_t),
                                   n * 2 * sizeof(int16_t), c->n_ch, cudaMemcpyHostToDevice, c->stream));
    int launches = 0;
    UA3_CUDA(duc_launch(c->duc, (uint32_t)n, c->stream, &launches));
    c->launches += (uint64_t)launches;
    c->last_tx = n;
    return UA3_OK;
}

int ua3reo_duc_push_wire(ua3reo_ctx* c, const uint8_t* wire_host, size_t n) {
    if (!c || (!wire_host && n)) return fail(UA3_E_INVAL, "ua3reo_duc_push_wire: null argument");
    if (!c->duc_alloc) return fail(UA3_E_STATE, "ua3reo_duc_push_wire: call ua3reo_duc_enable first");
    if (n > c->duc.max_in) return fail(UA3_E_TOOBIG, "ua3reo_duc_push_wire: more samples than max_tx_samples");
    // stm32_interface.v:206-227: k=300 Q_HOLD[15:8], 301 Q_HOLD[7:0], 302 I_HOLD[15:8], 303 I_HOLD[7:0] -> TX_I, TX_Q
    std::vector<int16_t> iq((size_t)c->n_ch * n * 2);
    for (size_t i = 0; i < (size_t)c->n_ch * n; ++i) {
        const uint8_t* w = wire_host + 4 * i;
        iq[2 * i] = (int16_t)(uint16_t)(((uint16_t)w[2] << 8) | w[3]);        // I
        iq[2 * i + 1] = (int16_t)(uint16_t)(((uint16_t)w[0] << 8) | w[1]);    // Q
    }
    const int rc = ua3reo_duc_push(c, iq.data(), n);
    if (rc != UA3_OK) return rc;
    UA3_CUDA(cudaStreamSynchronize(c->stream));      // iq is a temporary: the host->device copy must have finished
    return UA3_OK;
}

int ua3reo_duc_read_dac(ua3reo_ctx* c, uint16_t* dst, size_t n) {
    if (!c || (!dst && n)) return fail(UA3_E_INVAL, "ua3reo_duc_read_dac: null argument");
    if (!c->duc_alloc) return fail(UA3_E_STATE, "ua3reo_duc_read_dac: DUC not enabled");
    if (n != c->last_tx) return fail(UA3_E_INVAL, "ua3reo_duc_read_dac: n != samples of last push");
    UA3_CUDA(cudaSetDevice(c->device));
    const size_t row = n * 1024 * sizeof(uint16_t);
    if (n)
        UA3_CUDA(cudaMemcpy2DAsync(dst, row, c->duc.dac, (size_t)c->duc.max_in * 1024 * sizeof(uint16_t), row, c->n_ch,
                                   cudaMemcpyDeviceToHost, c->stream));
    UA3_CUDA(cudaStreamSynchronize(c->stream));
    return UA3_OK;
}

int ua3reo_duc_dac_device(ua3reo_ctx* c, const uint16_t** base, size_t* n_words, size_t* channel_stride_words) {
    if (!c || !base) return fail(UA3_E_INVAL, "ua3reo_duc_dac_device: null argument");
    if (!c->duc_alloc) return fail(UA3_E_STATE, "ua3reo_duc_dac_device: DUC not enabled");
    *base = c->duc.dac;
    if (n_words) *n_words = c->last_tx * 1024;
    if (channel_stride_words) *channel_stride_words = (size_t)c->duc.max_in * 1024;
    return UA3_OK;
}

int ua3reo_duc_read_otr(ua3reo_ctx* c, uint32_t* dst) {
    if (!c || !dst) return fail(UA3_E_INVAL, "ua3reo_duc_read_otr: null argument");
    if (!c->duc_alloc) return fail(UA3_E_STATE, "ua3reo_duc_read_otr: DUC not enabled");
    UA3_CUDA(cudaSetDevice(c->device));
    UA3_CUDA(cudaMemcpy2DAsync(dst, sizeof(uint32_t), reinterpret_cast<const uint8_t*>(c->duc.state) + offsetof(DucState, otr),
                               sizeof(DucState), sizeof(uint32_t), c->n_ch, cudaMemcpyDeviceToHost, c->stream));
    UA3_CUDA(cudaStreamSynchronize(c->stream));
    return UA3_OK;
}

// ------------------------------------------------------------------------------------------------
// transmit audio
// ------------------------------------------------------------------------------------------------
void ua3reo_tx_defaults(ua3reo_tx_settings* s) {
    if (!s) return;
    std::memset(s, 0, sizeof *s);
    s->mode = kModeUSB; s->rf_power = 20; s->volume = 20; s->filter_width = 2700; s->ssb_hpf_pass = 300;   // settings.c:33-94
}

int ua3reo_tx_enable(ua3reo_ctx* c, uint32_t max_blocks) {
    if (!c || max_blocks == 0) return fail(UA3_E_INVAL, "ua3reo_tx_enable: bad arguments");
    if (c->tx_alloc) return (max_blocks <= c->tx.max_blocks) ? UA3_OK : fail(UA3_E_STATE, "ua3reo_tx_enable: already enabled with fewer blocks");
    UA3_CUDA(cudaSetDevice(c->device));
    AllocScope scope(c);
    TxBuffers& t = c->tx;
    t.n_ch = c->n_ch; t.max_blocks = max_blocks;
    const size_t per_ch = (size_t)max_blocks * UA3_AUDIO_BLOCK * 2;
    UA3_CUDA(dev_alloc(c, &t.params, (size_t)c->n_ch));
    UA3_CUDA(dev_alloc(c, &t.state, (size_t)c->n_ch));
    UA3_CUDA(dev_alloc(c, &t.mic, (size_t)c->n_ch * per_ch));
    UA3_CUDA(dev_alloc(c, &t.iq_f, (size_t)c->n_ch * per_ch));
    UA3_CUDA(dev_alloc(c, &t.iq_w, (size_t)c->n_ch * per_ch));
    UA3_CUDA(dev_alloc(c, &t.loop_out, (size_t)c->n_ch * per_ch));
    UA3_CUDA(dev_alloc(c, &c->tx_flags, (size_t)c->n_ch));
    float T[513];
    cmsis_sin_table(T);
    UA3_CUDA(tx_upload_constants(T));
    int launches = 0;
    UA3_CUDA(tx_launch_init_state(t, c->stream, &launches));
    c->launches += (uint64_t)launches;
    ua3reo_tx_settings d;
    ua3reo_tx_defaults(&d);
    c->h_txpar.assign(c->n_ch, TxParams());
    for (uint32_t i = 0; i < c->n_ch; ++i) {
        std::memset(&c->h_txpar[i], 0, sizeof(TxParams));
        c->h_txpar[i].fm_index = 2.0f;                                   // static float32_t modulation_index = 2.0f
        bool a, b;
        if (!tx_derive(d, c->h_txpar[i], a, b)) return fail(UA3_E_INVAL, "internal: default TX settings rejected");
    }
    UA3_CUDA(cudaMemcpyAsync(t.params, c->h_txpar.data(), sizeof(TxParams) * c->n_ch, cudaMemcpyHostToDevice, c->stream));
    UA3_CUDA(cudaStreamSynchronize(c->stream));
    c->tx_alloc = scope.ok = true;
    return UA3_OK;
}

static int tx_apply_settings(ua3reo_ctx* c, uint32_t first, uint32_t n, const ua3reo_tx_settings* settings, bool live) {
    if (!c || !settings || first > c->n_ch || n > c->n_ch - first) return fail(UA3_E_INVAL, "ua3reo_tx_set: range");
    if (!c->tx_alloc) return fail(UA3_E_STATE, "ua3reo_tx_set: call ua3reo_tx_enable first");
    UA3_CUDA(cudaSetDevice(c->device));
    std::vector<TxParams> np(c->h_txpar.begin() + first, c->h_txpar.begin() + first + n);
    std::vector<uint8_t> flags(n, 0);
    for (uint32_t i = 0; i < n; ++i) {
        bool cl = false, chp = false;
        const TxParams old = np[i];
        if (!tx_derive(settings[i], np[i], cl, chp)) return fail(UA3_E_INVAL, "ua3reo_tx_set: settings outside the firmware's tables");
        if (live) {            // keep what the last ReinitAudioFilters() selected; nothing is cleared
            std::memcpy(np[i].lpf_k, old.lpf_k, sizeof old.lpf_k); std::memcpy(np[i].lpf_v, old.lpf_v, sizeof old.lpf_v);
            std::memcpy(np[i].hpf_k, old.hpf_k, sizeof old.hpf_k); std::memcpy(np[i].hpf_v, old.hpf_v, sizeof old.hpf_v);
            np[i].hpf_set = old.hpf_set;
        } else {
            flags[i] = (uint8_t)((cl ? 1 : 0) | (chp ? 2 : 0));
        }
    }
    for (uint32_t i = 0; i < n; ++i) c->h_txpar[first + i] = np[i];
    UA3_CUDA(cudaMemcpyAsync(c->tx.params + first, c->h_txpar.data() + first, sizeof(TxParams) * n, cudaMemcpyHostToDevice, c->stream));
    if (!live) {
        UA3_CUDA(cudaMemcpyAsync(c->tx_flags, flags.data(), n, cudaMemcpyHostToDevice, c->stream));
        int launches = 0;
        UA3_CUDA(tx_launch_clear(c->tx, c->tx_flags, first, n, c->stream, &launches));
        c->launches += (uint64_t)launches;
    }
    UA3_CUDA(cudaStreamSynchronize(c->stream));
    return UA3_OK;
}

int ua3reo_tx_set(ua3reo_ctx* c, uint32_t first, uint32_t n, const ua3reo_tx_settings* settings) {
    return tx_apply_settings(c, first, n, settings, false);
}

int ua3reo_tx_set_live(ua3reo_ctx* c, uint32_t first, uint32_t n, const ua3reo_tx_settings* settings) {
    return tx_apply_settings(c, first, n, settings, true);
}

int ua3reo_tx_process(ua3reo_ctx* c, const int16_t* mic_host, size_t n_blocks) {
    if (!c || (!mic_host && n_blocks)) return fail(UA3_E_INVAL, "ua3reo_tx_process: null argument");
    if (!c->tx_alloc) return fail(UA3_E_STATE, "ua3reo_tx_process: call ua3reo_tx_enable first");
    if (n_blocks > c->tx.max_blocks) return fail(UA3_E_TOOBIG, "ua3reo_tx_process: more blocks than max_blocks");
    UA3_CUDA(cudaSetDevice(c->device));
    const size_t row = n_blocks * UA3_AUDIO_BLOCK * 2 * sizeof(int16_t);
    if (n_blocks)
        UA3_CUDA(cudaMemcpy2DAsync(c->tx.mic, (size_t)c->tx.max_blocks * UA3_AUDIO_BLOCK * 2 * sizeof(int16_t), mic_host, row, row,
                                   c->n_ch, cudaMemcpyHostToDevice, c->stream));
    int launches = 0;
    UA3_CUDA(tx_launch_audio(c->tx, (uint32_t)n_blocks, c->stream, &launches));
    c->launches += (uint64_t)launches;
    c->last_tx_blocks = n_blocks;
    return UA3_OK;
}

int ua3reo_tx_read_iq(ua3reo_ctx* c, int16_t* iq_words, float* iq_float, size_t n_blocks) {
    if (!c) return fail(UA3_E_INVAL, "ua3reo_tx_read_iq: null argument");
    if (!c->tx_alloc) return fail(UA3_E_STATE, "ua3reo_tx_read_iq: transmit audio not enabled");
    if (n_blocks != c->last_tx_blocks) return fail(UA3_E_INVAL, "ua3reo_tx_read_iq: n_blocks != blocks of last call");
    UA3_CUDA(cudaSetDevice(c->device));
    const size_t n = n_blocks * UA3_AUDIO_BLOCK * 2, stride = (size_t)c->tx.max_blocks * UA3_AUDIO_BLOCK * 2;
    if (n && iq_words)
        UA3_CUDA(cudaMemcpy2DAsync(iq_words, n * sizeof(int16_t), c->tx.iq_w, stride * sizeof(int16_t), n * sizeof(int16_t), c->n_ch,
                                   cudaMemcpyDeviceToHost, c->stream));
    if (n && iq_float)
        UA3_CUDA(cudaMemcpy2DAsync(iq_float, n * sizeof(float), c->tx.iq_f, stride * sizeof(float), n * sizeof(float), c->n_ch,
                                   cudaMemcpyDeviceToHost, c->stream));
    UA3_CUDA(cudaStreamSynchronize(c->stream));
    return UA3_OK;
}

int ua3reo_tx_read_loopback(ua3reo_ctx* c, int32_t* dst, size_t n_blocks) {
    if (!c || (!dst && n_blocks)) return fail(UA3_E_INVAL, "ua3reo_tx_read_loopback: null argument");
    if (!c->tx_alloc) return fail(UA3_E_STATE, "ua3reo_tx_read_loopback: transmit audio not enabled");
    if (n_blocks != c->last_tx_blocks) return fail(UA3_E_INVAL, "ua3reo_tx_read_loopback: n_blocks != blocks of last call");
    UA3_CUDA(cudaSetDevice(c->device));
    const size_t n = n_blocks * UA3_AUDIO_BLOCK * 2, stride = (size_t)c->tx.max_blocks * UA3_AUDIO_BLOCK * 2;
    if (n)
        UA3_CUDA(cudaMemcpy2DAsync(dst, n * sizeof(int32_t), c->tx.loop_out, stride * sizeof(int32_t), n * sizeof(int32_t), c->n_ch,
                                   cudaMemcpyDeviceToHost, c->stream));
    UA3_CUDA(cudaStreamSynchronize(c->stream));
    return UA3_OK;
}

int ua3reo_tx_feed_duc(ua3reo_ctx* c) {
    if (!c) return fail(UA3_E_INVAL, "null context");
    if (!c->tx_alloc || !c->duc_alloc) return fail(UA3_E_STATE, "ua3reo_tx_feed_duc: enable the transmit audio stage and the DUC first");
    const size_t n = c->last_tx_blocks * UA3_AUDIO_BLOCK;
    if (n > c->duc.max_in) return fail(UA3_E_TOOBIG, "ua3reo_tx_feed_duc: DUC block smaller than the audio blocks");
    UA3_CUDA(cudaSetDevice(c->device));
    if (n)
        UA3_CUDA(cudaMemcpy2DAsync(c->duc.iq_in, (size_t)c->duc.max_in * 2 * sizeof(int16_t), c->tx.iq_w,
                                   (size_t)c->tx.max_blocks * UA3_AUDIO_BLOCK * 2 * sizeof(int16_t), n * 2 * sizeof(int16_t), c->n_ch,
                                   cudaMemcpyDeviceToDevice, c->stream));
    int launches = 0;
    UA3_CUDA(duc_launch(c->duc, (uint32_t)n, c->stream, &launches));
    c->launches += (uint64_t)launches;
    c->last_tx = n;
    return UA3_OK;
}

// ---------------------------------------------------------------------------------------------------------------------
// multi-device bank (one host process)
// ---------------------------------------------------------------------------------------------------------------------
struct ua3reo_bank {
    std::vector<ua3reo_ctx*> ctx;
    std::vector<uint32_t> first, count;
    uint32_t n_total = 0, max_block = 0;
    std::vector<int16_t*> adc[2];            // [slot][device slot] block buffers; slot alternates with the push number
    std::vector<cudaStream_t> fan;           // fan-out stream per device slot (on that device)
    std::vector<cudaEvent_t> ev_ready[2];    // block has arrived in adc[slot][i]
    std::vector<cudaEvent_t> ev_free[2];     // the push that read adc[slot][i] has finished
    std::vector<uint8_t> used[2];
    cudaEvent_t ev_in[2] = {nullptr, nullptr};   // host -> ingest device copy done
    uint64_t n_push = 0;
};

static void bank_slab(uint32_t n_total, int rank, int world, uint32_t* lo, uint32_t* n) {
    const uint32_t base = n_total / (uint32_t)world, rem = n_total % (uint32_t)world;
    *lo = (uint32_t)rank * base + std::min<uint32_t>((uint32_t)rank, rem);
    *n = base + ((uint32_t)rank < rem ? 1u : 0u);
}

int ua3reo_bank_destroy(ua3reo_bank* b) {
    if (!b) return UA3_OK;
    for (size_t i = 0; i < b->ctx.size(); ++i) {
        if (!b->ctx[i]) continue;
        cudaSetDevice(b->ctx[i]->device);
        if (i < b->fan.size() && b->fan[i]) { cudaStreamSynchronize(b->fan[i]); cudaStreamDestroy(b->fan[i]); }
        for (int s = 0; s < 2; ++s) {
            if (i < b->adc[s].size() && b->adc[s][i]) cudaFree(b->adc[s][i]);
            if (i < b->ev_ready[s].size() && b->ev_ready[s][i]) cudaEventDestroy(b->ev_ready[s][i]);
            if (i < b->ev_free[s].size() && b->ev_free[s][i]) cudaEventDestroy(b->ev_free[s][i]);
        }
    }
    if (!b->ctx.empty() && b->ctx[0]) { cudaSetDevice(b->ctx[0]->device); for (int s = 0; s < 2; ++s) if (b->ev_in[s]) cudaEventDestroy(b->ev_in[s]); }
    for (ua3reo_ctx* c : b->ctx) ctx_free(c);
    delete b;
    return UA3_OK;
}

int ua3reo_bank_create(int n_devices, const int* devices, uint32_t n_channels, uint32_t max_block_samples, ua3reo_bank** out) {
    if (!out || !devices || n_devices < 1 || n_channels < (uint32_t)n_devices) return fail(UA3_E_INVAL, "ua3reo_bank_create: bad arguments");
    *out = nullptr;
    if (max_block_samples == 0) max_block_samples = 1u << 20;
    ua3reo_bank* b = new (std::nothrow) ua3reo_bank;
    if (!b) return fail(UA3_E_INVAL, "out of host memory");
    b->n_total = n_channels; b->max_block = max_block_samples;
    b->ctx.assign((size_t)n_devices, nullptr);
    b->first.assign((size_t)n_devices, 0); b->count.assign((size_t)n_devices, 0);
    b->fan.assign((size_t)n_devices, nullptr);
    for (int s = 0; s < 2; ++s) {
        b->adc[s].assign((size_t)n_devices, nullptr);
        b->ev_ready[s].assign((size_t)n_devices, nullptr); b->ev_free[s].assign((size_t)n_devices, nullptr);
        b->used[s].assign((size_t)n_devices, 0);
    }
#define UA3_BTRY(call, what) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) { ua3reo_bank_destroy(b); return fail(UA3_E_CUDA, what, e__); } } while (0)
    for (int i = 0; i < n_devices; ++i) {
        bank_slab(n_channels, i, n_devices, &b->first[(size_t)i], &b->count[(size_t)i]);
        const int rc = ua3reo_create(devices[i], b->count[(size_t)i], max_block_samples, &b->ctx[(size_t)i]);
        if (rc != UA3_OK) { const std::string keep = g_err; ua3reo_bank_destroy(b); g_err = keep; return rc; }
        UA3_BTRY(cudaSetDevice(devices[i]), "cudaSetDevice");
        UA3_BTRY(cudaStreamCreateWithFlags(&b->fan[(size_t)i], cudaStreamNonBlocking), "cudaStreamCreate");
        for (int s = 0; s < 2; ++s) {
            UA3_BTRY(cudaMalloc((void**)&b->adc[s][(size_t)i], (size_t)max_block_samples * sizeof(int16_t)), "cudaMalloc");
            UA3_BTRY(cudaEventCreateWithFlags(&b->ev_ready[s][(size_t)i], cudaEventDisableTiming), "cudaEventCreate");
            UA3_BTRY(cudaEventCreateWithFlags(&b->ev_free[s][(size_t)i], cudaEventDisableTiming), "cudaEventCreate");
        }
        if (i > 0 && devices[i] != devices[0]) {            // direct NVLink path for the fan-out; "already enabled" is fine
            int can = 0;
            if (cudaDeviceCanAccessPeer(&can, devices[i], devices[0]) == cudaSuccess && can) {
                const cudaError_t e = cudaDeviceEnablePeerAccess(devices[0], 0);
                if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) UA3_BTRY(e, "cudaDeviceEnablePeerAccess");
                (void)cudaGetLastError();
            }
        }
    }
    UA3_BTRY(cudaSetDevice(devices[0]), "cudaSetDevice");
    for (int s = 0; s < 2; ++s) UA3_BTRY(cudaEventCreateWithFlags(&b->ev_in[s], cudaEventDisableTiming), "cudaEventCreate");
#undef UA3_BTRY
    *out = b;
    return UA3_OK;
}

int ua3reo_bank_n_devices(const ua3reo_bank* b) { return b ? (int)b->ctx.size() : 0; }

int ua3reo_bank_context(ua3reo_bank* b, int i, ua3reo_ctx** ctx, uint32_t* first, uint32_t* count) {
    if (!b || i < 0 || (size_t)i >= b->ctx.size()) return fail(UA3_E_INVAL, "ua3reo_bank_context: device slot");
    if (ctx) *ctx = b->ctx[(size_t)i];
    if (first) *first = b->first[(size_t)i];
    if (count) *count = b->count[(size_t)i];
    return UA3_OK;
}

// applies fn(ctx, slab-relative first, n, global offset) to the part of [first, first + n) each device owns
static int bank_route(ua3reo_bank* b, uint32_t first, uint32_t n, const std::function<int(ua3reo_ctx*, uint32_t, uint32_t, uint32_t)>& fn) {
    if (!b || first > b->n_total || n > b->n_total - first) return fail(UA3_E_INVAL, "bank: channel range");
    for (size_t i = 0; i < b->ctx.size(); ++i) {
        const uint32_t lo = std::max(first, b->first[i]), hi = std::min(first + n, b->first[i] + b->count[i]);
        if (lo >= hi) continue;
        const int rc = fn(b->ctx[i], lo - b->first[i], hi - lo, lo - first);
        if (rc != UA3_OK) return rc;
    }
    return UA3_OK;
}

int ua3reo_bank_set_fcw(ua3reo_bank* b, uint32_t first, uint32_t n, const uint32_t* fcw22) {
    if (!fcw22) return fail(UA3_E_INVAL, "ua3reo_bank_set_fcw: null argument");
    return bank_route(b, first, n, [&](ua3reo_ctx* c, uint32_t f, uint32_t m, uint32_t off) { return ua3reo_set_fcw(c, f, m, fcw22 + off); });
}

int ua3reo_bank_rx_enable(ua3reo_bank* b, int enable) {
    if (!b) return fail(UA3_E_INVAL, "null bank");
    for (ua3reo_ctx* c : b->ctx) { const int rc = ua3reo_rx_enable(c, enable); if (rc != UA3_OK) return rc; }
    return UA3_OK;
}

int ua3reo_bank_rx_set(ua3reo_bank* b, uint32_t first, uint32_t n, const ua3reo_rx_settings* settings) {
    if (!settings) return fail(UA3_E_INVAL, "ua3reo_bank_rx_set: null argument");
    return bank_route(b, first, n, [&](ua3reo_ctx* c, uint32_t f, uint32_t m, uint32_t off) { return ua3reo_rx_set(c, f, m, settings + off); });
}

int ua3reo_bank_push(ua3reo_bank* b, const int16_t* adc_host, size_t n, size_t* frames_out) {
    if (!b || (!adc_host && n)) return fail(UA3_E_INVAL, "ua3reo_bank_push: null argument");
    if (n > b->max_block) return fail(UA3_E_TOOBIG, "ua3reo_bank_push: block exceeds max_block_samples");
    const int s = (int)(b->n_push & 1);
    const size_t bytes = n * sizeof(int16_t);
    ua3reo_ctx* c0 = b->ctx[0];
    // ingest: host -> device 0, on device 0's fan-out stream, once the push that last read this slot is through
    UA3_CUDA(cudaSetDevice(c0->device));
    if (b->used[s][0]) UA3_CUDA(cudaStreamWaitEvent(b->fan[0], b->ev_free[s][0], 0));
    if (n) UA3_CUDA(cudaMemcpyAsync(b->adc[s][0], adc_host, bytes, cudaMemcpyHostToDevice, b->fan[0]));
    UA3_CUDA(cudaEventRecord(b->ev_in[s], b->fan[0]));
    UA3_CUDA(cudaEventRecord(b->ev_ready[s][0], b->fan[0]));
    // fan-out: copy engines, device 0 -> device i, each on the destination's own stream
    for (size_t i = 1; i < b->ctx.size(); ++i) {
        UA3_CUDA(cudaSetDevice(b->ctx[i]->device));
        UA3_CUDA(cudaStreamWaitEvent(b->fan[i], b->ev_in[s], 0));
        if (b->used[s][i]) UA3_CUDA(cudaStreamWaitEvent(b->fan[i], b->ev_free[s][i], 0));
        if (n) UA3_CUDA(cudaMemcpyPeerAsync(b->adc[s][i], b->ctx[i]->device, b->adc[s][0], c0->device, bytes, b->fan[i]));
        UA3_CUDA(cudaEventRecord(b->ev_ready[s][i], b->fan[i]));
    }
    // every device runs the chain for its slab over the whole block (zero copy when it is whole frames)
    size_t nf = 0;
    for (size_t i = 0; i < b->ctx.size(); ++i) {
        ua3reo_ctx* c = b->ctx[i];
        UA3_CUDA(cudaSetDevice(c->device));
        UA3_CUDA(cudaStreamWaitEvent(c->stream, b->ev_ready[s][i], 0));
        size_t f = 0;
        const int rc = ua3reo_ddc_push_device(c, b->adc[s][i], n, &f);
        if (rc != UA3_OK) return rc;
        UA3_CUDA(cudaEventRecord(b->ev_free[s][i], c->stream));
        b->used[s][i] = 1;
        if (i == 0) nf = f; else if (f != nf) return fail(UA3_E_STATE, "ua3reo_bank_push: devices disagree on the frame count");
    }
    // device 0's copy of the slot also feeds the peers: it is free when its own push AND their copies are done
    UA3_CUDA(cudaSetDevice(c0->device));
    for (size_t i = 1; i < b->ctx.size(); ++i) UA3_CUDA(cudaStreamWaitEvent(c0->stream, b->ev_ready[s][i], 0));
    UA3_CUDA(cudaEventRecord(b->ev_free[s][0], c0->stream));
    b->n_push++;
    if (frames_out) *frames_out = nf;
    return UA3_OK;
}

int ua3reo_bank_read_frames(ua3reo_bank* b, uint8_t* dst, size_t n_frames) {
    if (!b || (!dst && n_frames)) return fail(UA3_E_INVAL, "ua3reo_bank_read_frames: null argument");
    for (size_t i = 0; i < b->ctx.size(); ++i) {            // every device starts its copy, then all are awaited
        const int rc = ua3reo_ddc_read_frames_async(b->ctx[i], dst + (size_t)b->first[i] * n_frames * UA3_FRAME_BYTES, n_frames);
        if (rc != UA3_OK) return rc;
    }
    return ua3reo_bank_sync(b);
}

int ua3reo_bank_rx_counts(ua3reo_bank* b, size_t* audio_blocks, size_t* fft_frames) {
    if (!b) return fail(UA3_E_INVAL, "null bank");
    return ua3reo_rx_counts(b->ctx[0], audio_blocks, fft_frames);
}

int ua3reo_bank_rx_read_audio(ua3reo_bank* b, int32_t* dst, size_t n_blocks) {
    if (!b || (!dst && n_blocks)) return fail(UA3_E_INVAL, "ua3reo_bank_rx_read_audio: null argument");
    for (size_t i = 0; i < b->ctx.size(); ++i) {
        const int rc = ua3reo_rx_read_audio_async(b->ctx[i], dst + (size_t)b->first[i] * n_blocks * 2 * UA3_AUDIO_BLOCK, n_blocks);
        if (rc != UA3_OK) return rc;
    }
    return ua3reo_bank_sync(b);
}

int ua3reo_bank_rx_read_spectra(ua3reo_bank* b, float* dst, size_t n_frames) {
    if (!b || (!dst && n_frames)) return fail(UA3_E_INVAL, "ua3reo_bank_rx_read_spectra: null argument");
    for (size_t i = 0; i < b->ctx.size(); ++i) {
        const int rc = ua3reo_rx_read_spectra_async(b->ctx[i], dst + (size_t)b->first[i] * n_frames * UA3_FFT_BINS, n_frames);
        if (rc != UA3_OK) return rc;
    }
    return ua3reo_bank_sync(b);
}

int ua3reo_bank_sync(ua3reo_bank* b) {
    if (!b) return fail(UA3_E_INVAL, "null bank");
    for (size_t i = 0; i < b->ctx.size(); ++i) {
        UA3_CUDA(cudaSetDevice(b->ctx[i]->device));
        UA3_CUDA(cudaStreamSynchronize(b->fan[i]));
        const int rc = ua3reo_sync(b->ctx[i]);
        if (rc != UA3_OK) return rc;
    }
    return UA3_OK;
}

int ua3reo_profile_begin(ua3reo_ctx* c, uint32_t max_blocks) {
    if (!c) return fail(UA3_E_INVAL, "null context");
    UA3_CUDA(cudaSetDevice(c->device));
    while (c->prof_ev.size() < (size_t)max_blocks * kProfEvents) {
        cudaEvent_t e;
        UA3_CUDA(cudaEventCreate(&e));
        c->prof_ev.push_back(e);
    }
    c->prof_cap = max_blocks;
    c->prof_used = 0;
    c->prof_mask = 0xFFFFFFFFu;
    return UA3_OK;
}

// Same, but only the two events around kernel slot `kernel` are recorded, so that the measurement costs the measured run
// two event records per block instead of eight (bench.py times the front kernel this way inside its timed region).
int ua3reo_profile_begin_kernel(ua3reo_ctx* c, uint32_t max_blocks, uint32_t kernel) {
    if (kernel + 1 >= (uint32_t)kProfEvents) return fail(UA3_E_INVAL, "ua3reo_profile_begin_kernel: kernel slot out of range");
    const int rc = ua3reo_profile_begin(c, max_blocks);
    if (rc != UA3_OK) return rc;
    c->prof_mask = (1u << kernel) | (1u << (kernel + 1));
    return UA3_OK;
}

int ua3reo_profile_end(ua3reo_ctx* c, double* kernel_ms, uint32_t n_kernels, uint32_t* blocks) {
    if (!c || !kernel_ms) return fail(UA3_E_INVAL, "null argument");
    UA3_CUDA(cudaSetDevice(c->device));
    UA3_CUDA(cudaStreamSynchronize(c->stream));
    UA3_CUDA(cudaStreamSynchronize(c->rx_stream));     // the STM32 stage records its events on its own stream
    for (uint32_t k = 0; k < n_kernels; ++k) kernel_ms[k] = 0.0;
    for (uint32_t b = 0; b < c->prof_used; ++b)
        for (uint32_t k = 0; k < (uint32_t)kProfEvents - 1 && k < n_kernels; ++k) {
            float ms = 0.f;
            if (((c->prof_mask >> k) & 3u) != 3u) continue;          // this slot's two events were not recorded
            const cudaEvent_t* ev = c->prof_ev.data() + (size_t)b * kProfEvents;
            UA3_CUDA(cudaEventElapsedTime(&ms, ev[k], ev[k + 1]));
            kernel_ms[k] += (double)ms;
        }
    if (blocks) *blocks = c->prof_used;
    c->prof_cap = 0;
    c->prof_used = 0;
    c->prof_mask = 0xFFFFFFFFu;
    return UA3_OK;
}

int ua3reo_measure_int32_peak(int device, double* ops_per_s) {
    if (!ops_per_s) return fail(UA3_E_INVAL, "null argument");
    int n_dev = 0;
    if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev == 0) return fail(UA3_E_NODEV, "no CUDA device");
    if (device < 0 || device >= n_dev) return fail(UA3_E_INVAL, "device index out of range");
    UA3_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    UA3_CUDA(cudaGetDeviceProperties(&prop, device));
    UA3_CUDA(measure_int32_peak(prop.multiProcessorCount, nullptr, ops_per_s));
    return UA3_OK;
}

int ua3reo_measure_lds_peak(int device, double* wavefronts_per_s) {
    if (!wavefronts_per_s) return fail(UA3_E_INVAL, "null argument");
    int n_dev = 0;
    if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev == 0) return fail(UA3_E_NODEV, "no CUDA device");
    if (device < 0 || device >= n_dev) return fail(UA3_E_INVAL, "device index out of range");
    UA3_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    UA3_CUDA(cudaGetDeviceProperties(&prop, device));
    UA3_CUDA(measure_lds_peak(prop.multiProcessorCount, nullptr, wavefronts_per_s));
    return UA3_OK;
}

}  // extern "C"
