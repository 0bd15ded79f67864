// fanout.cu - ua3reo_fanout_*: the ADC block fanned out to one PROCESS per GPU without a collective kernel.
//
// Every rank owns an arena (cudaMalloc, exported with cudaIpcGetMemHandle): `n_buffers` block slots, one READY word per
// slot and - read on the ingest rank only - one CREDIT word per (rank, slot).  The ingest rank opens every peer's arena and
// writes block s into slot s % n_buffers of each of them with cudaMemcpyAsync (copy engines over NVLink, no SM), then stores
// s + 1 into that slot's READY word; a consumer makes ITS stream wait for READY == s + 1 (cuStreamWaitValue32: the wait is
// executed by the stream's front end, no kernel, no host round trip, no ordering between the processes' host threads) and
// after its kernels stores s + 1 into its CREDIT word in the ingest rank's arena, which the ingest rank's copy stream waits
// for before it refills the slot.  All words count blocks, so equality waits cannot be overtaken: a slot is refilled only
// after its credit came back.  (A collective KERNEL cannot overlap the receive path: a front CTA owns its whole SM, so the
// NCCL broadcast needed an SM set aside for it - ua3reo_reserve_sms; this needs none.)
#include "../../include/ua3reo_b200.h"
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstring>
#include <new>
#include <string>
#include <vector>

void ua3_set_last_error(const char* what);            // api.cu

namespace {

typedef CUresult (*WaitValue32Fn)(CUstream, CUdeviceptr, cuuint32_t, unsigned int);
typedef CUresult (*WriteValue32Fn)(CUstream, CUdeviceptr, cuuint32_t, unsigned int);

int ffail(int code, const std::string& what, cudaError_t e = cudaSuccess) {
    std::string msg = what;
    if (e != cudaSuccess) { msg += ": "; msg += cudaGetErrorString(e); }
    ua3_set_last_error(msg.c_str());                  // read back with ua3reo_last_error()
    return code;
}

constexpr size_t kAlign = 256;
size_t round_up(size_t x) { return (x + kAlign - 1) / kAlign * kAlign; }

}   // namespace

struct ua3reo_fanout {
    int device = 0, rank = 0, world = 1, src = 0, n_buf = 2;
    size_t block = 0;                  // samples per slot
    size_t slot_bytes = 0, ready_off = 0, credit_off = 0, arena_bytes = 0;
    uint8_t* arena = nullptr;          // this rank's arena
    std::vector<uint8_t*> peer;        // ingest rank: every rank's arena (own = arena); others: only [src] is set
    std::vector<bool> opened;
    uint32_t* stage_words = nullptr;   // local words for the staged form of a remote store (see signal())
    uint32_t n_stage = 0, stage_pos = 0;
    bool direct_remote_store = true;   // cuStreamWriteValue32 straight onto the peer mapping; else local word + 4-byte copy
    cudaStream_t send_stream = nullptr;    // ingest rank: credit waits, block copies, READY stores
    cudaStream_t credit_stream = nullptr;  // every rank: the CREDIT store, behind an event of the consumer stream
    cudaEvent_t ev_consumed = nullptr;
    uint64_t n_sent = 0, n_acquired = 0, n_released = 0;
    bool connected = false;
    WaitValue32Fn wait32 = nullptr;
    WriteValue32Fn write32 = nullptr;

    uint32_t* ready_word(uint8_t* base, int slot) const { return (uint32_t*)(base + ready_off) + slot * 32; }      // 128-byte lines
    uint32_t* credit_word(uint8_t* base, int r, int slot) const { return (uint32_t*)(base + credit_off) + (r * n_buf + slot) * 32; }
    int16_t* slot_ptr(uint8_t* base, int slot) const { return (int16_t*)(base + (size_t)slot * slot_bytes); }
};

#define UA3_FCUDA(call)                                                       \
    do {                                                                      \
        cudaError_t e__ = (call);                                             \
        if (e__ != cudaSuccess) return ffail(UA3_E_CUDA, #call, e__);         \
    } while (0)

// value -> *addr in stream order.  addr may be a peer mapping (CUDA IPC).
static int signal_word(ua3reo_fanout* f, cudaStream_t st, uint32_t* addr, uint32_t value, bool remote) {
    if (!remote || f->direct_remote_store) {
        CUresult r = f->write32((CUstream)st, (CUdeviceptr)(uintptr_t)addr, value, CU_STREAM_WRITE_VALUE_DEFAULT);
        if (r == CUDA_SUCCESS) return UA3_OK;
        if (!remote) return ffail(UA3_E_CUDA, "cuStreamWriteValue32 failed (code " + std::to_string((int)r) + ")");
        f->direct_remote_store = false;            // this driver does not store onto peer mappings: stage the word locally
    }
    uint32_t* w = f->stage_words + (f->stage_pos++ % f->n_stage);
    CUresult r = f->write32((CUstream)st, (CUdeviceptr)(uintptr_t)w, value, CU_STREAM_WRITE_VALUE_DEFAULT);
    if (r != CUDA_SUCCESS) return ffail(UA3_E_CUDA, "cuStreamWriteValue32 (staged) failed (code " + std::to_string((int)r) + ")");
    UA3_FCUDA(cudaMemcpyAsync(addr, w, sizeof(uint32_t), cudaMemcpyDefault, st));
    return UA3_OK;
}

static int wait_word(ua3reo_fanout* f, cudaStream_t st, uint32_t* addr, uint32_t value) {
    CUresult r = f->wait32((CUstream)st, (CUdeviceptr)(uintptr_t)addr, value, CU_STREAM_WAIT_VALUE_EQ);
    if (r != CUDA_SUCCESS) return ffail(UA3_E_CUDA, "cuStreamWaitValue32 failed (code " + std::to_string((int)r) + ")");
    return UA3_OK;
}

extern "C" {

int ua3reo_fanout_disconnect(ua3reo_fanout* f) {
    if (!f) return UA3_OK;
    cudaSetDevice(f->device);
    if (f->send_stream) cudaStreamSynchronize(f->send_stream);
    if (f->credit_stream) cudaStreamSynchronize(f->credit_stream);
    for (size_t r = 0; r < f->peer.size(); ++r)
        if (f->opened[r] && f->peer[r]) { cudaIpcCloseMemHandle(f->peer[r]); f->peer[r] = nullptr; f->opened[r] = false; }
    if (f->world > 1) f->connected = false;
    return UA3_OK;
}

int ua3reo_fanout_destroy(ua3reo_fanout* f) {
    if (!f) return UA3_OK;
    ua3reo_fanout_disconnect(f);
    if (f->ev_consumed) cudaEventDestroy(f->ev_consumed);
    if (f->send_stream) cudaStreamDestroy(f->send_stream);
    if (f->credit_stream) cudaStreamDestroy(f->credit_stream);
    if (f->stage_words) cudaFree(f->stage_words);
    if (f->arena) cudaFree(f->arena);
    delete f;
    return UA3_OK;
}

// what both directions share: the entry points of the stream memory operations, the arena, the local words, the streams
static int link_create(ua3reo_fanout* f, int device, int rank, int world, int src, size_t arena_bytes) {
    f->device = device; f->rank = rank; f->world = world; f->src = src;
    f->arena_bytes = arena_bytes;
    f->peer.assign((size_t)world, nullptr);
    f->opened.assign((size_t)world, false);
#define UA3_FTRY(call, what) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) return ffail(UA3_E_CUDA, what, e__); } while (0)
    UA3_FTRY(cudaSetDevice(device), "cudaSetDevice");
    cudaDriverEntryPointQueryResult q1, q2;
    void *p1 = nullptr, *p2 = nullptr;
    UA3_FTRY(cudaGetDriverEntryPoint("cuStreamWaitValue32", &p1, cudaEnableDefault, &q1), "cudaGetDriverEntryPoint(cuStreamWaitValue32)");
    UA3_FTRY(cudaGetDriverEntryPoint("cuStreamWriteValue32", &p2, cudaEnableDefault, &q2), "cudaGetDriverEntryPoint(cuStreamWriteValue32)");
    if (!p1 || !p2 || q1 != cudaDriverEntryPointSuccess || q2 != cudaDriverEntryPointSuccess)
        return ffail(UA3_E_STATE, "the driver has no stream memory operations");
    f->wait32 = (WaitValue32Fn)p1;
    f->write32 = (WriteValue32Fn)p2;
    UA3_FTRY(cudaMalloc((void**)&f->arena, f->arena_bytes), "cudaMalloc(arena)");
    UA3_FTRY(cudaMemset(f->arena, 0, f->arena_bytes), "cudaMemset(arena)");
    f->n_stage = 4096;
    UA3_FTRY(cudaMalloc((void**)&f->stage_words, f->n_stage * sizeof(uint32_t)), "cudaMalloc(words)");
    UA3_FTRY(cudaStreamCreateWithFlags(&f->send_stream, cudaStreamNonBlocking), "cudaStreamCreate");
    UA3_FTRY(cudaStreamCreateWithFlags(&f->credit_stream, cudaStreamNonBlocking), "cudaStreamCreate");
    UA3_FTRY(cudaEventCreateWithFlags(&f->ev_consumed, cudaEventDisableTiming), "cudaEventCreate");
    UA3_FTRY(cudaDeviceSynchronize(), "cudaDeviceSynchronize");
#undef UA3_FTRY
    f->peer[(size_t)rank] = f->arena;
    if (world == 1) f->connected = true;
    return UA3_OK;
}

int ua3reo_fanout_create(int device, int rank, int world, int src, size_t block_samples, int n_buffers, ua3reo_fanout** out) {
    if (!out || world < 1 || rank < 0 || rank >= world || src < 0 || src >= world || block_samples == 0 || n_buffers < 2 || n_buffers > 16)
        return ffail(UA3_E_INVAL, "ua3reo_fanout_create: bad arguments");
    *out = nullptr;
    ua3reo_fanout* f = new (std::nothrow) ua3reo_fanout;
    if (!f) return ffail(UA3_E_STATE, "ua3reo_fanout_create: out of host memory");
    f->n_buf = n_buffers; f->block = block_samples;
    f->slot_bytes = round_up(block_samples * sizeof(int16_t));
    f->ready_off = f->slot_bytes * (size_t)n_buffers;
    f->credit_off = f->ready_off + round_up((size_t)n_buffers * 128);
    const int rc = link_create(f, device, rank, world, src, f->credit_off + round_up((size_t)world * (size_t)n_buffers * 128));
    if (rc != UA3_OK) { ua3reo_fanout_destroy(f); return rc; }       // (the message of the failure is kept: destroy sets none)
    *out = f;
    return UA3_OK;
}

int ua3reo_fanout_handle(ua3reo_fanout* f, void* handle64) {
    if (!f || !handle64) return ffail(UA3_E_INVAL, "ua3reo_fanout_handle: null argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == UA3_FANOUT_HANDLE_BYTES, "handle size");
    UA3_FCUDA(cudaSetDevice(f->device));
    cudaIpcMemHandle_t h;
    UA3_FCUDA(cudaIpcGetMemHandle(&h, f->arena));
    memcpy(handle64, &h, sizeof h);
    return UA3_OK;
}

int ua3reo_fanout_connect(ua3reo_fanout* f, const void* handles) {
    if (!f || !handles) return ffail(UA3_E_INVAL, "ua3reo_fanout_connect: null argument");
    if (f->connected) return UA3_OK;
    UA3_FCUDA(cudaSetDevice(f->device));
    const uint8_t* h = (const uint8_t*)handles;
    for (int r = 0; r < f->world; ++r) {
        if (r == f->rank) continue;
        if (f->rank != f->src && r != f->src) continue;       // only the ingest rank talks to everybody
        cudaIpcMemHandle_t mh;
        memcpy(&mh, h + (size_t)r * UA3_FANOUT_HANDLE_BYTES, sizeof mh);
        void* p = nullptr;
        cudaError_t e = cudaIpcOpenMemHandle(&p, mh, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) return ffail(UA3_E_CUDA, "cudaIpcOpenMemHandle(rank " + std::to_string(r) + ")", e);
        f->peer[(size_t)r] = (uint8_t*)p;
        f->opened[(size_t)r] = true;
    }
    f->connected = true;
    return UA3_OK;
}

int ua3reo_fanout_send(ua3reo_fanout* f, const int16_t* block, size_t n) {
    if (!f || !block) return ffail(UA3_E_INVAL, "ua3reo_fanout_send: null argument");
    if (f->rank != f->src) return ffail(UA3_E_STATE, "ua3reo_fanout_send: only the ingest rank sends");
    if (!f->connected) return ffail(UA3_E_STATE, "ua3reo_fanout_send: not connected");
    if (n != f->block) return ffail(UA3_E_INVAL, "ua3reo_fanout_send: a send is one whole block");
    UA3_FCUDA(cudaSetDevice(f->device));
    const uint64_t s = f->n_sent;
    const int slot = (int)(s % (uint64_t)f->n_buf);
    cudaStream_t st = f->send_stream;
    // own slot first (host -> device or device -> device); the peers are fed from it, so a pinned host block crosses PCIe once
    int rc;
    if (s >= (uint64_t)f->n_buf && (rc = wait_word(f, st, f->credit_word(f->arena, f->rank, slot), (uint32_t)(s - f->n_buf + 1))) != UA3_OK) return rc;
    UA3_FCUDA(cudaMemcpyAsync(f->slot_ptr(f->arena, slot), block, n * sizeof(int16_t), cudaMemcpyDefault, st));
    for (int r = 0; r < f->world; ++r) {
        if (r == f->rank) continue;
        if (s >= (uint64_t)f->n_buf && (rc = wait_word(f, st, f->credit_word(f->arena, r, slot), (uint32_t)(s - f->n_buf + 1))) != UA3_OK) return rc;
        UA3_FCUDA(cudaMemcpyAsync(f->slot_ptr(f->peer[(size_t)r], slot), f->slot_ptr(f->arena, slot), n * sizeof(int16_t), cudaMemcpyDefault, st));
        if ((rc = signal_word(f, st, f->ready_word(f->peer[(size_t)r], slot), (uint32_t)(s + 1), true)) != UA3_OK) return rc;
    }
    if ((rc = signal_word(f, st, f->ready_word(f->arena, slot), (uint32_t)(s + 1), false)) != UA3_OK) return rc;
    f->n_sent = s + 1;
    return UA3_OK;
}

int ua3reo_fanout_acquire(ua3reo_fanout* f, void* consumer_stream, const int16_t** block_dev) {
    if (!f || !block_dev) return ffail(UA3_E_INVAL, "ua3reo_fanout_acquire: null argument");
    if (!f->connected) return ffail(UA3_E_STATE, "ua3reo_fanout_acquire: not connected");
    if (f->n_acquired != f->n_released) return ffail(UA3_E_STATE, "ua3reo_fanout_acquire: the previous block was not released");
    UA3_FCUDA(cudaSetDevice(f->device));
    const uint64_t s = f->n_acquired;
    const int slot = (int)(s % (uint64_t)f->n_buf);
    int rc = wait_word(f, (cudaStream_t)consumer_stream, f->ready_word(f->arena, slot), (uint32_t)(s + 1));
    if (rc != UA3_OK) return rc;
    *block_dev = f->slot_ptr(f->arena, slot);
    f->n_acquired = s + 1;
    return UA3_OK;
}

int ua3reo_fanout_release(ua3reo_fanout* f, void* consumer_stream) {
    if (!f) return ffail(UA3_E_INVAL, "ua3reo_fanout_release: null argument");
    if (f->n_released + 1 != f->n_acquired) return ffail(UA3_E_STATE, "ua3reo_fanout_release without an acquired block");
    UA3_FCUDA(cudaSetDevice(f->device));
    const uint64_t s = f->n_released;
    const int slot = (int)(s % (uint64_t)f->n_buf);
    // the credit travels on its own stream behind an event, so that the consumer stream's kernels stay back to back
    UA3_FCUDA(cudaEventRecord(f->ev_consumed, (cudaStream_t)consumer_stream));
    UA3_FCUDA(cudaStreamWaitEvent(f->credit_stream, f->ev_consumed, 0));
    uint8_t* home = f->peer[(size_t)f->src];
    int rc = signal_word(f, f->credit_stream, f->credit_word(home, f->rank, slot), (uint32_t)(s + 1), f->rank != f->src);
    if (rc != UA3_OK) return rc;
    f->n_released = s + 1;
    return UA3_OK;
}

int ua3reo_fanout_sync(ua3reo_fanout* f) {
    if (!f) return ffail(UA3_E_INVAL, "ua3reo_fanout_sync: null argument");
    UA3_FCUDA(cudaSetDevice(f->device));
    UA3_FCUDA(cudaStreamSynchronize(f->send_stream));
    UA3_FCUDA(cudaStreamSynchronize(f->credit_stream));
    return UA3_OK;
}

int ua3reo_fanout_info(const ua3reo_fanout* f, int* direct_remote_store, uint64_t* n_sent, uint64_t* n_acquired) {
    if (!f) return ffail(UA3_E_INVAL, "ua3reo_fanout_info: null argument");
    if (direct_remote_store) *direct_remote_store = f->direct_remote_store ? 1 : 0;
    if (n_sent) *n_sent = f->n_sent;
    if (n_acquired) *n_acquired = f->n_acquired;
    return UA3_OK;
}

// -----------------------------------------------------------------------------------------------------------------------
// the opposite direction: every rank's slab of results into one buffer on the root rank (ua3reo_gather_*).  Arena layout,
// flags first so that their offsets do not depend on the rank: ARRIVE[rank][slot] (used in the root's arena, stored by
// `rank` behind its copy), CREDIT[slot] (used in every rank's arena, stored by the root once it has consumed the slot), and -
// on the root only - the data: n_buffers x world x slab bytes.
// -----------------------------------------------------------------------------------------------------------------------
struct ua3reo_gather {
    ua3reo_fanout link;                // arena, mappings, entry points, streams (src = the root)
    size_t slab = 0, slab_pad = 0, arrive_off = 0, credit_off = 0, data_off = 0;
    uint64_t n_sent = 0, n_acquired = 0, n_released = 0;
    uint32_t* arrive(uint8_t* base, int r, int slot) const { return (uint32_t*)(base + arrive_off) + (r * link.n_buf + slot) * 32; }
    uint32_t* credit(uint8_t* base, int slot) const { return (uint32_t*)(base + credit_off) + slot * 32; }
    uint8_t* data(uint8_t* base, int slot, int r) const { return base + data_off + ((size_t)slot * link.world + r) * slab_pad; }
};

int ua3reo_gather_destroy(ua3reo_gather* g) {
    if (!g) return UA3_OK;
    ua3reo_fanout_disconnect(&g->link);
    ua3reo_fanout* f = &g->link;
    if (f->ev_consumed) cudaEventDestroy(f->ev_consumed);
    if (f->send_stream) cudaStreamDestroy(f->send_stream);
    if (f->credit_stream) cudaStreamDestroy(f->credit_stream);
    if (f->stage_words) cudaFree(f->stage_words);
    if (f->arena) cudaFree(f->arena);
    delete g;
    return UA3_OK;
}

int ua3reo_gather_disconnect(ua3reo_gather* g) { return g ? ua3reo_fanout_disconnect(&g->link) : UA3_OK; }

int ua3reo_gather_create(int device, int rank, int world, int root, size_t slab_bytes, int n_buffers, ua3reo_gather** out) {
    if (!out || world < 1 || rank < 0 || rank >= world || root < 0 || root >= world || slab_bytes == 0 || n_buffers < 2 || n_buffers > 16)
        return ffail(UA3_E_INVAL, "ua3reo_gather_create: bad arguments");
    *out = nullptr;
    ua3reo_gather* g = new (std::nothrow) ua3reo_gather;
    if (!g) return ffail(UA3_E_STATE, "ua3reo_gather_create: out of host memory");
    g->link.n_buf = n_buffers;
    g->slab = slab_bytes;
    g->slab_pad = round_up(slab_bytes);
    g->arrive_off = 0;
    g->credit_off = round_up((size_t)world * (size_t)n_buffers * 128);
    g->data_off = g->credit_off + round_up((size_t)n_buffers * 128);
    const size_t bytes = g->data_off + (rank == root ? (size_t)n_buffers * (size_t)world * g->slab_pad : 0);
    const int rc = link_create(&g->link, device, rank, world, root, bytes);
    if (rc != UA3_OK) { ua3reo_gather_destroy(g); return rc; }
    *out = g;
    return UA3_OK;
}

int ua3reo_gather_handle(ua3reo_gather* g, void* handle64) { return g ? ua3reo_fanout_handle(&g->link, handle64) : ffail(UA3_E_INVAL, "null gather"); }
int ua3reo_gather_connect(ua3reo_gather* g, const void* handles) { return g ? ua3reo_fanout_connect(&g->link, handles) : ffail(UA3_E_INVAL, "null gather"); }

int ua3reo_gather_send(ua3reo_gather* g, const void* slab_dev, void* producer_stream) {
    if (!g || !slab_dev) return ffail(UA3_E_INVAL, "ua3reo_gather_send: null argument");
    ua3reo_fanout* f = &g->link;
    if (!f->connected) return ffail(UA3_E_STATE, "ua3reo_gather_send: not connected");
    UA3_FCUDA(cudaSetDevice(f->device));
    cudaStream_t st = (cudaStream_t)producer_stream;       // everything in the producer's stream: the slab is read in stream order
    const uint64_t s = g->n_sent;
    const int slot = (int)(s % (uint64_t)f->n_buf);
    uint8_t* root = f->peer[(size_t)f->src];
    const bool remote = f->rank != f->src;
    int rc;
    if (s >= (uint64_t)f->n_buf && (rc = wait_word(f, st, g->credit(f->arena, slot), (uint32_t)(s - f->n_buf + 1))) != UA3_OK) return rc;
    UA3_FCUDA(cudaMemcpyAsync(g->data(root, slot, f->rank), slab_dev, g->slab, cudaMemcpyDefault, st));
    if ((rc = signal_word(f, st, g->arrive(root, f->rank, slot), (uint32_t)(s + 1), remote)) != UA3_OK) return rc;
    g->n_sent = s + 1;
    return UA3_OK;
}

int ua3reo_gather_acquire(ua3reo_gather* g, void* consumer_stream, const void** all_dev, size_t* rank_stride) {
    if (!g || !all_dev) return ffail(UA3_E_INVAL, "ua3reo_gather_acquire: null argument");
    ua3reo_fanout* f = &g->link;
    if (f->rank != f->src) return ffail(UA3_E_STATE, "ua3reo_gather_acquire: only the root rank receives");
    if (!f->connected) return ffail(UA3_E_STATE, "ua3reo_gather_acquire: not connected");
    if (g->n_acquired != g->n_released) return ffail(UA3_E_STATE, "ua3reo_gather_acquire: the previous slot was not released");
    UA3_FCUDA(cudaSetDevice(f->device));
    const uint64_t s = g->n_acquired;
    const int slot = (int)(s % (uint64_t)f->n_buf);
    for (int r = 0; r < f->world; ++r) {
        const int rc = wait_word(f, (cudaStream_t)consumer_stream, g->arrive(f->arena, r, slot), (uint32_t)(s + 1));
        if (rc != UA3_OK) return rc;
    }
    *all_dev = g->data(f->arena, slot, 0);
    if (rank_stride) *rank_stride = g->slab_pad;
    g->n_acquired = s + 1;
    return UA3_OK;
}

int ua3reo_gather_release(ua3reo_gather* g, void* consumer_stream) {
    if (!g) return ffail(UA3_E_INVAL, "ua3reo_gather_release: null argument");
    ua3reo_fanout* f = &g->link;
    if (g->n_released + 1 != g->n_acquired) return ffail(UA3_E_STATE, "ua3reo_gather_release without an acquired slot");
    UA3_FCUDA(cudaSetDevice(f->device));
    const uint64_t s = g->n_released;
    const int slot = (int)(s % (uint64_t)f->n_buf);
    UA3_FCUDA(cudaEventRecord(f->ev_consumed, (cudaStream_t)consumer_stream));
    UA3_FCUDA(cudaStreamWaitEvent(f->credit_stream, f->ev_consumed, 0));
    for (int r = 0; r < f->world; ++r) {
        const int rc = signal_word(f, f->credit_stream, g->credit(f->peer[(size_t)r], slot), (uint32_t)(s + 1), r != f->rank);
        if (rc != UA3_OK) return rc;
    }
    g->n_released = s + 1;
    return UA3_OK;
}

}   // extern "C"
