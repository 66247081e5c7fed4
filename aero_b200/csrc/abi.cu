// C ABI (include/aero_b200.h) over the sm_100a kernels: context, transform plans, segment / FRI
// handles.  Host-side work here is bookkeeping only (table construction, index lists for openings,
// byte packing); every field/hash operation over trace-sized data runs in a CUDA kernel.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <chrono>
#include <condition_variable>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <set>
#include <string>
#include <thread>
#include <vector>

#include "../../include/aero_b200.h"
#include "../host/copy_pool.hpp"
#include "kernels.cuh"
#include "ntt.cuh"

using namespace aero;

namespace aero {
std::atomic<unsigned long long> g_launch_count{0};
}

// -------------------------------------------------------------------------------------------------
// context
// -------------------------------------------------------------------------------------------------
struct PhaseStat {
    int calls = 0;
    double ms = 0;
};
struct PendingEvent {
    std::string name;
    cudaEvent_t a, b;
};
struct PowTableOwned {
    uint64_t *lo = nullptr, *hi = nullptr;
    int lo_bits = 0;
    PowTable view() const { return PowTable{lo, hi, lo_bits}; }
};

struct aero_upload {
    aero_ctx *ctx = nullptr;
    uint64_t *d = nullptr;          // n_cols x n_rows, contiguous columns
    cudaEvent_t done = nullptr;
    std::vector<const uint64_t *> cols;
    int col_begin = 0, col_end = 0;  // columns actually copied (a sharded context may take its own only)
    uint64_t row_begin = 0, row_end = 0;  // rows actually copied (likewise: its own row block)
    uint64_t n_rows = 0;
    bool queued = false;
};

constexpr int PUSH_PARTS = 4;
struct aero_ctx {
    int device = 0;
    cudaStream_t stream = 0;
    cudaStream_t copy_stream = nullptr;  // host->device uploads overlapped with compute
    cudaStream_t hash_stream = nullptr;  // row hashing of batch k overlapped with the LDE of batch k+1
    // coefficient pushes to the peers, beside the LDE of the columns already here: one copy engine moves
    // ~280 GB/s over NVLink, so every block goes out as PUSH_PARTS pieces on as many streams
    cudaStream_t push_stream[PUSH_PARTS] = {};
    int push_parts = PUSH_PARTS;         // "push_parts": pieces actually used (ranks sharing a device run short of hardware queues)
    cudaEvent_t ev_push = nullptr;
    unsigned long long push_epoch = 0;   // arrival flags: window + 1024 + 8 * (source rank * PUSH_PARTS + piece)
    cudaEvent_t ev_lde = nullptr, ev_hash = nullptr, ev_copy_order = nullptr;
    bool overlap_hash = false;           // aero_ctx_set_option("overlap_hash"); measured slower on B200, see DESIGN.md
    bool force_split_intt = false;       // test hook: take the B x size-n route of constraints_into_poly at any size
    int air_blocks_per_sm = 0;           // experiment hook: resident blocks per SM of the AIR evaluator, threads looping over steps (0 = one thread per step)
    std::map<std::string, uint64_t *> const_tables;
    std::vector<aero_upload *> deferred_uploads;  // queued behind the next segment commit's own copies
    // Exchange window (multi-GPU): one allocation per rank, the same size everywhere, mapped by every
    // peer (CUDA IPC across processes, plain peer access inside one).  [0, 4096): barrier flags and the time-out word; the rest is a bump heap
    // that all ranks of a sharded proof allocate from in the same order, so a buffer sits at the same
    // offset in every window and a kernel can store to its peer copies directly over NVLink.
    uint8_t *win = nullptr;
    size_t win_bytes = 0, win_off = 4096;
    int win_live = 0;
    int win_ranks = 0;                       // > 1 once the peers are attached
    uint8_t *win_base[AERO_MAX_RANKS] = {};  // every rank's window as mapped here (win_base[shard_rank] == win)
    bool win_ipc[AERO_MAX_RANKS] = {};       // mapped with cudaIpcOpenMemHandle (closed with the context)
    unsigned long long win_epoch = 0;
    // Barriers between the ranks are device-side flag barriers (no host involvement) once a proof shape is
    // warm.  While a context may still call cudaMalloc -- the first proof of a shape -- a rank spinning on
    // the device could block a peer's allocation, so those proofs synchronise on the host instead:
    // stream sync + the caller's rendezvous (aero_ctx_set_host_barrier).
    bool win_host_sync = true;
    aero_host_barrier_fn host_barrier = nullptr;
    void *host_barrier_user = nullptr;
    std::set<std::string> warm_shapes;
    std::string cur_shape;
    bool force_host_sync = false;            // test hook: never switch to device-side barriers
    bool own_stream = false;                 // `stream` was created by aero_ctx_create_stream
    int hash_blocks_per_sm = 2;          // grid cap of an overlapped row-hash launch ("hash_blocks_per_sm")
    int num_sms = 148;
    int form = AERO_FORM_MONTGOMERY;
    std::string err;
    bool profile = false;
    std::string profile_prefix;  // non-empty: only phases whose name starts with it are timed
    std::map<std::string, PhaseStat> stats;
    std::vector<PendingEvent> pending;
    std::vector<cudaEvent_t> event_pool;  // timing events of finished phases, reused (no driver call per phase)
    std::map<std::string, DftTables> plans;
    std::map<std::string, PowTableOwned> pow_tables;
    std::vector<void *> owned;  // device allocations freed with the context
    size_t lde_batch_bytes = (size_t)1 << 30;  // NTT scratch budget per column batch
    int upload_batch_cols = 8;                 // columns per host->device copy batch of aero_segment_commit
    int upload_edge_cols = -1;                 // size of its first / last batch (-1: half a batch, 0: uniform batches)
    size_t ntt_table_max_bytes = (size_t)1 << 30;  // largest full inter-pass twiddle table a plan may hold
    int hash_early_batches = 2;  // host-buffer commits: column batches hashed right after their extension ("hash_early_batches")
    int fri_fused = 0;       // fold a FRI layer and hash the next layer's leaves in one kernel: 0 never (default: measured slower), 1 small layers, 2 all ("fri_fused")
    int ntt_outer_log = -1;  // third factor of two-pass transforms: -1 = 2^(logn-20) above 2^20 points, 0 = never, k = force 2^k (tests)
    int shard_rank = 0, shard_world = 1;       // LDE coset shard of this context (multi-GPU)
    std::multimap<size_t, void *> free_blocks;  // exact-size cache of released device blocks
    std::map<void *, size_t> live_blocks;
    size_t cached_bytes = 0, cache_limit_bytes = (size_t)96 << 30;
    // staging pair of the opening phase (GatherBatch): pinned host + device, grown on demand
    uint8_t *h_stage = nullptr, *d_stage = nullptr;
    size_t stage_bytes = 0;
    // Pinned ring for the small host<->device transfers of a proof (roots, coefficient vectors, OOD point
    // tables, flags).  An async copy from or to PAGEABLE memory makes the calling thread wait inside the
    // driver until the stream gets there -- behind a rank barrier that wait can starve the very peers
    // the barrier is waiting for when several ranks share a process; pinned copies just queue.
    uint8_t *h_ring = nullptr;
    size_t ring_bytes = 0, ring_off = 0;
    // two pinned slots for bulk transfers whose host side is pageable (aero_segment_download_lde)
    uint8_t *h_bulk[2] = {nullptr, nullptr};
    size_t bulk_bytes = 0;
    cudaEvent_t ev_bulk[2] = {nullptr, nullptr};
    int bulk_slot = 0;
};

#define CTX_FAIL(ctx, code, ...)                         \
    do {                                                 \
        char _b[512];                                    \
        snprintf(_b, sizeof _b, __VA_ARGS__);            \
        (ctx)->err = _b;                                 \
        return (code);                                   \
    } while (0)
#define CUDA_TRY(ctx, expr)                                                                          \
    do {                                                                                             \
        cudaError_t _e = (expr);                                                                     \
        if (_e != cudaSuccess)                                                                       \
            CTX_FAIL(ctx, _e == cudaErrorMemoryAllocation ? AERO_ERR_NOMEM : AERO_ERR_CUDA, "%s: %s", #expr, \
                     cudaGetErrorString(_e));                                                        \
    } while (0)
#define TRY(expr)                        \
    do {                                 \
        aero_status _s = (expr);         \
        if (_s != AERO_OK) return _s;    \
    } while (0)

static bool is_pow2(uint64_t x) { return x && !(x & (x - 1)); }
static int ilog2(uint64_t x) {
    int l = 0;
    while ((1ULL << l) < x) l++;
    return l;
}

struct PhaseTimer {
    aero_ctx *ctx;
    PendingEvent ev;
    bool on;
    cudaStream_t s;
    PhaseTimer(aero_ctx *c, const char *name, cudaStream_t stream = nullptr, bool use_given = false)
        : ctx(c), on(c->profile), s(use_given ? stream : c->stream) {
        if (on && !c->profile_prefix.empty() && strncmp(name, c->profile_prefix.c_str(), c->profile_prefix.size()) != 0) on = false;
        if (!on) return;
        ev.name = name;
        ev.a = take(c);
        ev.b = take(c);
        cudaEventRecord(ev.a, s);
    }
    ~PhaseTimer() {
        if (!on) return;
        cudaEventRecord(ev.b, s);
        ctx->pending.push_back(ev);
    }
    static cudaEvent_t take(aero_ctx *c) {
        cudaEvent_t e = nullptr;
        if (!c->event_pool.empty()) {
            e = c->event_pool.back();
            c->event_pool.pop_back();
        } else {
            cudaEventCreate(&e);
        }
        return e;
    }
};
static void profile_flush(aero_ctx *ctx) {
    for (auto &p : ctx->pending) {
        cudaEventSynchronize(p.b);
        float ms = 0;
        cudaEventElapsedTime(&ms, p.a, p.b);
        auto &s = ctx->stats[p.name];
        s.calls++;
        s.ms += ms;
        ctx->event_pool.push_back(p.a);
        ctx->event_pool.push_back(p.b);
    }
    ctx->pending.clear();
}

// Device memory for segments / scratch comes from a per-context exact-size cache: every proof of a
// given shape requests the same sizes in the same order, so after the first proof no driver call is
// made at all.  (cudaMallocAsync's pool showed 50-1400 ms stalls when two processes re-shaped their
// pools at once.)  All work of a context is on one stream, so handing a block freed "now" to a later
// launch on that stream is ordered correctly without synchronisation.
static void cache_release_all(aero_ctx *ctx) {
    cudaStreamSynchronize(ctx->stream);
    for (auto &kv : ctx->free_blocks) cudaFree(kv.second);
    ctx->free_blocks.clear();
    ctx->cached_bytes = 0;
}
static aero_status dev_alloc(aero_ctx *ctx, void **p, size_t bytes) {
    if (bytes == 0) bytes = 8;
    bytes = (bytes + 511) & ~(size_t)511;
    auto it = ctx->free_blocks.find(bytes);
    if (it != ctx->free_blocks.end()) {
        *p = it->second;
        ctx->free_blocks.erase(it);
        ctx->cached_bytes -= bytes;
        ctx->live_blocks[*p] = bytes;
        return AERO_OK;
    }
    cudaError_t e = cudaMalloc(p, bytes);
    if (e != cudaSuccess) {  // give the cached blocks back and retry once
        cudaGetLastError();
        cache_release_all(ctx);
        e = cudaMalloc(p, bytes);
    }
    if (e != cudaSuccess) {
        cudaGetLastError();
        CTX_FAIL(ctx, AERO_ERR_NOMEM, "cudaMalloc(%zu bytes): %s", bytes, cudaGetErrorString(e));
    }
    ctx->live_blocks[*p] = bytes;
    return AERO_OK;
}
static bool win_owns(const aero_ctx *ctx, const void *p) {
    return ctx->win && (const uint8_t *)p >= ctx->win && (const uint8_t *)p < ctx->win + ctx->win_bytes;
}
static void dev_free(aero_ctx *ctx, void *p) {
    if (!p) return;
    if (win_owns(ctx, p)) {  // bump heap: space comes back when the last buffer of the proof goes
        if (--ctx->win_live == 0) ctx->win_off = 4096;
        return;
    }
    auto it = ctx->live_blocks.find(p);
    if (it == ctx->live_blocks.end()) return;
    const size_t bytes = it->second;
    ctx->live_blocks.erase(it);
    ctx->free_blocks.emplace(bytes, p);
    ctx->cached_bytes += bytes;
    if (ctx->cached_bytes > ctx->cache_limit_bytes) cache_release_all(ctx);
}
// temporary device blocks of one call: returned to the cache when the call ends, also on error paths
struct DevBlocks {
    aero_ctx *ctx;
    std::vector<void *> v;
    explicit DevBlocks(aero_ctx *c) : ctx(c) {}
    DevBlocks(const DevBlocks &) = delete;
    aero_status alloc(void **p, size_t bytes) {
        aero_status st = dev_alloc(ctx, p, bytes);
        if (st == AERO_OK) v.push_back(*p);
        return st;
    }
    aero_status alloc_shared(void **p, size_t bytes);
    ~DevBlocks() {
        for (void *p : v) dev_free(ctx, p);
    }
};
struct SegmentDeleter {
    void operator()(aero_segment *s) const { aero_segment_destroy(s); }
};
using SegmentGuard = std::unique_ptr<aero_segment, SegmentDeleter>;
struct FriDeleter {
    void operator()(aero_fri *f) const { aero_fri_destroy(f); }
};
using FriGuard = std::unique_ptr<aero_fri, FriDeleter>;

// `bytes` of pinned scratch, valid until RING_BYTES more have been taken (every proof synchronises its
// stream many times per lap)
constexpr size_t RING_BYTES = (size_t)8 << 20;
static aero_status ring_take(aero_ctx *ctx, size_t bytes, void **out) {
    bytes = (bytes + 63) & ~(size_t)63;
    if (!ctx->h_ring || bytes > ctx->ring_bytes) {
        CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
        if (ctx->h_ring) cudaFreeHost(ctx->h_ring);
        ctx->h_ring = nullptr;
        ctx->ring_bytes = std::max(RING_BYTES, 2 * bytes);
        CUDA_TRY(ctx, cudaMallocHost((void **)&ctx->h_ring, ctx->ring_bytes));
        ctx->ring_off = 0;
    }
    if (ctx->ring_off + bytes > ctx->ring_bytes) ctx->ring_off = 0;
    *out = ctx->h_ring + ctx->ring_off;
    ctx->ring_off += bytes;
    return AERO_OK;
}
// device -> caller memory through the ring: queue, synchronise, copy out
static aero_status download_small(aero_ctx *ctx, void *dst, const void *d_src, size_t bytes) {
    void *h = nullptr;
    TRY(ring_take(ctx, bytes, &h));
    CUDA_TRY(ctx, cudaMemcpyAsync(h, d_src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    memcpy(dst, h, bytes);
    return AERO_OK;
}
// caller memory -> device through the ring (the source may be reused as soon as this returns)
static aero_status upload_small(aero_ctx *ctx, void *d_dst, const void *src, size_t bytes) {
    void *h = nullptr;
    TRY(ring_take(ctx, bytes, &h));
    memcpy(h, src, bytes);
    CUDA_TRY(ctx, cudaMemcpyAsync(d_dst, h, bytes, cudaMemcpyHostToDevice, ctx->stream));
    return AERO_OK;
}

// Host copy split over a few threads: one thread moves ~10 GB/s, PCIe 5 x16 ~55 GB/s.  The workers are a
// process-wide pool created on first use (spawning threads per copy cost ~0.1 ms per 8 MB column, a third of
// the copy itself); the calling thread takes chunks too.  The pool is never destroyed -- its threads sleep
// on a condition variable and go away with the process -- so nothing runs in a static destructor.
static void parallel_memcpy(void *dst, const void *src, size_t bytes) {
    if (bytes < ((size_t)2 << 20)) {
        memcpy(dst, src, bytes);
        return;
    }
    static aero::host::CopyPool *pool = [] {
        const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
        return new aero::host::CopyPool(std::min(7u, hw > 1 ? hw - 1 : 0u));  // + the calling thread
    }();
    const size_t chunk = (size_t)1 << 20;  // 1 MiB pieces, handed out dynamically
    std::vector<aero::host::CopyPool::Chunk> chunks;
    chunks.reserve(bytes / chunk + 1);
    for (size_t a = 0; a < bytes; a += chunk)
        chunks.push_back({(uint8_t *)dst + a, (const uint8_t *)src + a, std::min(chunk, bytes - a)});
    pool->run(std::move(chunks));
}
// is this host pointer page-locked (cudaMallocHost / cudaHostRegister)?  Async copies from or to pageable
// memory are staged by the driver and block the caller.
static bool host_is_pinned(const void *p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return a.type == cudaMemoryTypeHost;
}

static aero_status bulk_reserve(aero_ctx *ctx, size_t bytes) {
    if (bytes <= ctx->bulk_bytes) return AERO_OK;
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    if (ctx->copy_stream) CUDA_TRY(ctx, cudaStreamSynchronize(ctx->copy_stream));
    for (int i = 0; i < 2; i++) {
        if (ctx->h_bulk[i]) cudaFreeHost(ctx->h_bulk[i]);
        ctx->h_bulk[i] = nullptr;
        if (!ctx->ev_bulk[i]) CUDA_TRY(ctx, cudaEventCreateWithFlags(&ctx->ev_bulk[i], cudaEventDisableTiming));
    }
    ctx->bulk_bytes = 0;
    for (int i = 0; i < 2; i++) CUDA_TRY(ctx, cudaMallocHost((void **)&ctx->h_bulk[i], bytes));
    ctx->bulk_bytes = bytes;
    return AERO_OK;
}
// Host -> device copy of `bytes` from possibly PAGEABLE memory on the copy stream: page-locked sources are
// handed to the copy engine as they are; pageable ones go through the two pinned bulk slots in pieces, the
// host copy of piece k + 1 (split over threads) running under the PCIe transfer of piece k.  (Left to the
// driver, a pageable source is staged by one thread at ~8-10 GB/s while the caller is blocked.)
static aero_status staged_h2d(aero_ctx *ctx, void *d_dst, const void *h_src, size_t bytes, bool pinned, int *slot_counter) {
    if (pinned) {
        CUDA_TRY(ctx, cudaMemcpyAsync(d_dst, h_src, bytes, cudaMemcpyHostToDevice, ctx->copy_stream));
        return AERO_OK;
    }
    const size_t piece = ctx->bulk_bytes;
    for (size_t off = 0; off < bytes; off += piece) {
        const size_t len = std::min(piece, bytes - off);
        const int slot = (*slot_counter)++ & 1;
        CUDA_TRY(ctx, cudaEventSynchronize(ctx->ev_bulk[slot]));  // the slot's previous transfer has left it
        parallel_memcpy(ctx->h_bulk[slot], (const uint8_t *)h_src + off, len);
        CUDA_TRY(ctx, cudaMemcpyAsync((uint8_t *)d_dst + off, ctx->h_bulk[slot], len, cudaMemcpyHostToDevice, ctx->copy_stream));
        CUDA_TRY(ctx, cudaEventRecord(ctx->ev_bulk[slot], ctx->copy_stream));
    }
    return AERO_OK;
}

// Pinned host / device staging pair for small result downloads (OOD frame, openings): results land
// in pinned memory so that several downloads can be queued before the single synchronisation.
static aero_status stage_reserve(aero_ctx *ctx, size_t total) {
    if (total <= ctx->stage_bytes) return AERO_OK;
    size_t cap = ctx->stage_bytes ? ctx->stage_bytes : ((size_t)1 << 20);
    while (cap < total) cap *= 2;
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    if (ctx->h_stage) cudaFreeHost(ctx->h_stage);
    if (ctx->d_stage) cudaFree(ctx->d_stage);
    ctx->h_stage = ctx->d_stage = nullptr;
    ctx->stage_bytes = 0;
    CUDA_TRY(ctx, cudaMallocHost((void **)&ctx->h_stage, cap));
    CUDA_TRY(ctx, cudaMalloc((void **)&ctx->d_stage, cap));
    ctx->stage_bytes = cap;
    return AERO_OK;
}
static bool ctx_sharded(const aero_ctx *ctx) { return ctx->shard_world > 1; }
static aero_status window_barrier(aero_ctx *ctx);
// Buffers that other ranks write into (coefficients, leaf digests, sub-roots, DEEP evaluations, opening
// results) live in the exchange window of a sharded context; otherwise this is dev_alloc.
static aero_status dev_alloc_shared(aero_ctx *ctx, void **p, size_t bytes) {
    if (ctx_sharded(ctx)) {
        if (ctx->win_ranks != ctx->shard_world)
            CTX_FAIL(ctx, AERO_ERR_STATE, "sharded context (%d ranks) has no exchange window attached", ctx->shard_world);
        // first allocation after the heap was reset: every rank must be done with the previous contents
        // before anyone stores into them again (the ranks allocate in the same order, so all arrive here)
        if (ctx->win_live == 0) TRY(window_barrier(ctx));
        bytes = (bytes + 511) & ~(size_t)511;
        if (ctx->win_off + bytes > ctx->win_bytes)
            CTX_FAIL(ctx, AERO_ERR_NOMEM, "exchange window too small: %zu bytes needed, %zu of %zu in use", bytes, ctx->win_off, ctx->win_bytes);
        *p = ctx->win + ctx->win_off;
        ctx->win_off += bytes;
        ctx->win_live++;
        return AERO_OK;
    }
    return dev_alloc(ctx, p, bytes);
}
// the copies of a buffer on every rank (only the own rank's for a buffer outside the window)
static RankPtrs rank_ptrs(const aero_ctx *ctx, const void *p) {
    RankPtrs q;
    for (int r = 0; r < AERO_MAX_RANKS; r++) q.p[r] = nullptr;
    q.p[ctx_sharded(ctx) ? ctx->shard_rank : 0] = const_cast<void *>(p);
    if (ctx_sharded(ctx) && win_owns(ctx, p)) {
        const size_t off = (const uint8_t *)p - ctx->win;
        for (int r = 0; r < ctx->win_ranks; r++) q.p[r] = ctx->win_base[r] + off;
    }
    return q;
}
static aero_status window_barrier(aero_ctx *ctx) {
    if (!ctx_sharded(ctx)) return AERO_OK;
    if (ctx->win_host_sync) {
        if (!ctx->host_barrier) CTX_FAIL(ctx, AERO_ERR_STATE, "sharded context needs a host barrier (aero_ctx_set_host_barrier) until a proof shape is warm");
        CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
        if (ctx->host_barrier(ctx->host_barrier_user) != AERO_OK) CTX_FAIL(ctx, AERO_ERR_STATE, "host barrier failed: a peer rank did not arrive");
        return AERO_OK;
    }
    peer_barrier(rank_ptrs(ctx, ctx->win), ctx->shard_world, ctx->shard_rank, ++ctx->win_epoch, (unsigned int *)(ctx->win + 2048),
                 ctx->stream);
    CUDA_TRY(ctx, cudaGetLastError());
    return AERO_OK;
}
aero_status DevBlocks::alloc_shared(void **p, size_t bytes) {
    aero_status st = dev_alloc_shared(ctx, p, bytes);
    if (st == AERO_OK) v.push_back(*p);
    return st;
}
// A peer that never reached a device-side barrier (it failed) is reported here instead of hanging the
// GPU.  The time-out count is cleared once reported, so the context stays usable after the ranks are
// back in lock-step.
// The check rides on a synchronisation the caller needs anyway: window_check_queue() queues the download
// of the time-out word, the caller queues its own downloads and synchronises the stream once, then
// window_check_result() looks at the word.
static aero_status window_check_queue(aero_ctx *ctx, const unsigned int **word) {
    *word = nullptr;
    if (!ctx_sharded(ctx) || ctx->win_host_sync) return AERO_OK;
    void *h = nullptr;
    TRY(ring_take(ctx, 4, &h));
    CUDA_TRY(ctx, cudaMemcpyAsync(h, ctx->win + 2048, 4, cudaMemcpyDeviceToHost, ctx->stream));
    *word = (const unsigned int *)h;
    return AERO_OK;
}
static aero_status window_check_result(aero_ctx *ctx, const unsigned int *word) {
    if (word && *word) {
        const unsigned int t = *word;
        cudaMemsetAsync(ctx->win + 2048, 0, 4, ctx->stream);
        CTX_FAIL(ctx, AERO_ERR_STATE, "exchange barrier timed out %u time(s): a peer rank did not arrive", t);
    }
    return AERO_OK;
}
// device -> caller memory and the barrier check with ONE stream synchronisation
static aero_status download_small_checked(aero_ctx *ctx, void *dst, const void *d_src, size_t bytes) {
    const unsigned int *word = nullptr;
    TRY(window_check_queue(ctx, &word));
    TRY(download_small(ctx, dst, d_src, bytes));
    return window_check_result(ctx, word);
}
template <typename T>
static aero_status upload_vec(aero_ctx *ctx, T **d, const std::vector<T> &h, bool own = true) {
    void *p = nullptr;
    CUDA_TRY(ctx, cudaMalloc(&p, std::max<size_t>(8, h.size() * sizeof(T))));
    CUDA_TRY(ctx, cudaMemcpy(p, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice));
    if (own) ctx->owned.push_back(p);
    *d = (T *)p;
    return AERO_OK;
}

static inline uint64_t to_canon(const aero_ctx *ctx, uint64_t x) {
    return ctx->form == AERO_FORM_MONTGOMERY ? gl::mont_to_canon(x) : gl::canon(x);
}
static inline uint64_t from_canon(const aero_ctx *ctx, uint64_t x) {
    return ctx->form == AERO_FORM_MONTGOMERY ? gl::canon_to_mont(x) : x;
}

// -------------------------------------------------------------------------------------------------
// transform plans
// -------------------------------------------------------------------------------------------------
// stage table for an M-point mixed-radix DIT evaluating on the coset sigma*<w_M> (layout: ntt.cuh).
// For the round that merges 2^R sub-transforms of size 2^s0 into size m = 2^(s0+R), sub-transform e
// (a bit-reversed residue rho = bitrev_R(e)) at offset low is multiplied by (sigma^(M/m) w_m^low)^rho.
static void fill_stage_table(uint64_t *tw, int logM, uint64_t sigma, uint64_t wM) {
    const uint64_t M = 1ULL << logM;
    tw[0] = 0;
    const NttRounds rounds(logM);
    int s0 = 0;
    for (int i = 0; i < rounds.count; i++) {
        const int R = rounds.log(i);
        const uint64_t m = 1ULL << (s0 + R);
        const uint64_t sg = gl::pow(sigma, M / m);  // sigma^(M/m)
        const uint64_t wm = gl::pow(wM, M / m);     // primitive m-th root
        uint64_t x = sg;                            // sigma_m * w_m^low
        for (uint64_t low = 0; low < (1ULL << s0); low++) {
            uint64_t pw = x;                        // x^rho
            for (uint32_t rho = 1; rho < (1u << R); rho++) {
                uint32_t e = 0;
                for (int b = 0; b < R; b++) e |= ((rho >> b) & 1u) << (R - 1 - b);
                tw[((uint64_t)e << s0) + low] = pw;
                pw = gl::mul(pw, x);
            }
            x = gl::mul(x, wm);
        }
        s0 += R;
    }
}
static void fill_pow_table(std::vector<uint64_t> &lo, std::vector<uint64_t> &hi, uint64_t base, int total_bits,
                           int lo_bits, uint64_t hi_scale) {
    lo.resize(1ULL << lo_bits);
    hi.resize(1ULL << std::max(0, total_bits - lo_bits));
    uint64_t x = 1;
    for (auto &v : lo) {
        v = x;
        x = gl::mul(x, base);
    }
    const uint64_t step = gl::pow(base, 1ULL << lo_bits);
    x = hi_scale;
    for (auto &v : hi) {
        v = x;
        x = gl::mul(x, step);
    }
}

// shifts: per-coset input shift s_r (evaluate p on s_r * <w_n>), or all 1.
// scale_c: constant folded into the result.  post_base != 0: out[i] *= post_base^i.
static aero_status get_plan(aero_ctx *ctx, const std::string &key_in, int logn, bool inverse,
                            const std::vector<uint64_t> &shifts, uint64_t scale_c, uint64_t post_base,
                            const DftTables **out) {
    const std::string key = key_in + "/o" + std::to_string(ctx->ntt_outer_log);  // the split is part of the plan
    auto it = ctx->plans.find(key);
    if (it != ctx->plans.end()) {
        *out = &it->second;
        return AERO_OK;
    }
    if (logn < 1 || logn > NTT_MAX_LOG) CTX_FAIL(ctx, AERO_ERR_UNSUPPORTED, "transform size 2^%d unsupported (1..%d)", logn, NTT_MAX_LOG);
    DftTables t;
    t.logn = logn;
    t.ncosets = (int)shifts.size();
    t.plain = true;
    t.inverse = inverse;
    for (uint64_t sh : shifts) t.plain = t.plain && sh == 1;
    const uint64_t n = 1ULL << logn;
    uint64_t w = gl::root_of_unity(logn);
    if (inverse) w = gl::inv(w);
    if (logn <= NTT_SINGLE_MAX_LOG) {
        t.log1 = 0;
        t.log2 = logn;
        std::vector<uint64_t> st((size_t)t.ncosets * n);
        for (int r = 0; r < t.ncosets; r++) fill_stage_table(st.data() + (size_t)r * n, logn, shifts[r], w);
        TRY(upload_vec(ctx, &t.stage2, st));
        t.single_scale = scale_c;
        if (post_base) {
            std::vector<uint64_t> pu(n);
            uint64_t x = scale_c;
            for (auto &v : pu) {
                v = x;
                x = gl::mul(x, post_base);
            }
            TRY(upload_vec(ctx, &t.post_u, pu));
        }
    } else {
        // Above 2^20 points a third factor n0 is split off so that both shared-memory passes stay 2^10-point
        // transforms (ntt.cuh); plans with an output scale (the coset interpolation, which also permutes its
        // output) keep two passes.
        if (!post_base) {
            if (ctx->ntt_outer_log < 0) t.log0 = logn > 20 ? logn - 20 : 0;
            else if (ctx->ntt_outer_log <= NTT_OUTER_MAX_LOG && logn - ctx->ntt_outer_log >= 12) t.log0 = ctx->ntt_outer_log;
        }
        const int rem = logn - t.log0;
        t.log2 = rem / 2;
        t.log1 = rem - t.log2;
        const uint64_t n1 = 1ULL << t.log1, n2 = 1ULL << t.log2, n0 = 1ULL << t.log0;
        const uint64_t nq = n2 * n0;  // tile columns of pass 1
        const uint64_t w1 = gl::pow(w, nq), w2 = gl::pow(w, n1 * n0);
        std::vector<uint64_t> st1((size_t)t.ncosets * n1), st2(n2), ib((size_t)t.ncosets * nq);
        for (int r = 0; r < t.ncosets; r++) {
            fill_stage_table(st1.data() + (size_t)r * n1, t.log1, gl::pow(shifts[r], nq), w1);
            uint64_t x = scale_c;
            for (uint64_t q = 0; q < nq; q++) {
                ib[(size_t)r * nq + q] = x;
                x = gl::mul(x, shifts[r]);
            }
        }
        fill_stage_table(st2.data(), t.log2, 1, w2);
        TRY(upload_vec(ctx, &t.stage1, st1));
        TRY(upload_vec(ctx, &t.stage2, st2));
        TRY(upload_vec(ctx, &t.inter_b, ib));
        if (t.log0) {
            // pass 2 of a three-pass plan: result (i2; j0) times w_(n2 n0)^(i2 j0), w_(n2 n0) = w^n1
            std::vector<uint64_t> pj((size_t)n0 * n2);
            const uint64_t wq = gl::pow(w, n1);
            uint64_t wj = 1;  // wq^j0
            for (uint64_t j0 = 0; j0 < n0; j0++) {
                uint64_t x = 1;
                for (uint64_t i2 = 0; i2 < n2; i2++) {
                    pj[j0 * n2 + i2] = x;
                    x = gl::mul(x, wj);
                }
                wj = gl::mul(wj, wq);
            }
            TRY(upload_vec(ctx, &t.post_j, pj));
        }
        t.lo_bits = (logn + 1) / 2;
        std::vector<uint64_t> lo, hi;
        fill_pow_table(lo, hi, w, logn, t.lo_bits, 1);
        TRY(upload_vec(ctx, &t.wlo, lo));
        TRY(upload_vec(ctx, &t.whi, hi));
        // full inter-pass table (one multiplication per element instead of two) while it stays small
        // against the transform's own data: [ncosets][n] entries
        if ((size_t)t.ncosets * n * 8 <= ctx->ntt_table_max_bytes) {
            void *p = nullptr;
            CUDA_TRY(ctx, cudaMalloc(&p, (size_t)t.ncosets * n * 8));
            ctx->owned.push_back(p);
            dft_fill_inter_table(t, (uint64_t *)p, ctx->stream);
            CUDA_TRY(ctx, cudaGetLastError());
            t.inter_full = (uint64_t *)p;
        }
        if (post_base) {
            std::vector<uint64_t> pu(n1), pv(n2);
            uint64_t x = 1;
            for (auto &v : pu) {
                v = x;
                x = gl::mul(x, post_base);
            }
            const uint64_t pb1 = gl::pow(post_base, n1);
            x = 1;
            for (auto &v : pv) {
                v = x;
                x = gl::mul(x, pb1);
            }
            TRY(upload_vec(ctx, &t.post_u, pu));
            TRY(upload_vec(ctx, &t.post_v, pv));
        }
    }
    auto ins = ctx->plans.emplace(key, t);
    *out = &ins.first->second;
    return AERO_OK;
}

static aero_status plan_intt(aero_ctx *ctx, int logn, bool input_mont, const DftTables **out) {
    char key[64];
    snprintf(key, sizeof key, "intt/%d/%d", logn, (int)input_mont);
    uint64_t c = gl::inv((1ULL << logn) % gl::P);
    if (input_mont) c = gl::mul(c, gl::MONT_R_INV);  // the DFT is linear: fold x*2^-64 into 1/n
    return get_plan(ctx, key, logn, true, {1}, c, 0, out);
}
static std::vector<uint64_t> coset_shifts(int logn, int log_blowup, uint64_t offset) {
    // coset r holds natural rows k = B*i + r : x = offset * g_N^(B*i + r) = (offset * g_N^r) * g_n^i
    const uint64_t gN = gl::root_of_unity(logn + log_blowup);
    std::vector<uint64_t> s(1ULL << log_blowup);
    uint64_t x = offset;
    for (auto &v : s) {
        v = x;
        x = gl::mul(x, gN);
    }
    return s;
}
static aero_status plan_lde(aero_ctx *ctx, int logn, int log_blowup, bool input_mont, const DftTables **out) {
    char key[64];
    snprintf(key, sizeof key, "lde/%d/%d/%d", logn, log_blowup, (int)input_mont);
    return get_plan(ctx, key, logn, false, coset_shifts(logn, log_blowup, gl::GENERATOR),
                    input_mont ? gl::MONT_R_INV : 1, 0, out);
}
// interpolate_poly_with_offset (fft/serial.rs:86-103): coeff[i] = (1/N) offset^-i IDFT(ev)[i]
static aero_status plan_coset_intt(aero_ctx *ctx, int logN, bool input_mont, const DftTables **out) {
    char key[64];
    snprintf(key, sizeof key, "cintt/%d/%d", logN, (int)input_mont);
    uint64_t c = gl::inv((1ULL << logN) % gl::P);
    if (input_mont) c = gl::mul(c, gl::MONT_R_INV);
    return get_plan(ctx, key, logN, true, {1}, c, gl::inv(gl::GENERATOR), out);
}
static aero_status get_pow_table(aero_ctx *ctx, const std::string &key, uint64_t base, int total_bits,
                                 uint64_t hi_scale, PowTable *out) {
    auto it = ctx->pow_tables.find(key);
    if (it == ctx->pow_tables.end()) {
        PowTableOwned o;
        o.lo_bits = (total_bits + 1) / 2;
        std::vector<uint64_t> lo, hi;
        fill_pow_table(lo, hi, base, total_bits, o.lo_bits, hi_scale);
        TRY(upload_vec(ctx, &o.lo, lo));
        TRY(upload_vec(ctx, &o.hi, hi));
        it = ctx->pow_tables.emplace(key, o).first;
    }
    *out = it->second.view();
    return AERO_OK;
}

// Transforms of 2^25 / 2^26 points (the top of BASELINE's NTT sweep): one outer radix-B step (B = 2, 4) over
// 2^24-point two-pass transforms, see large_combine in poly.cu.  kind 0: plain inverse transform with the
// 1/n scale (interpolate_poly), kind 1: coset LDE (evaluate_poly_with_offset), both column by column.
static aero_status dft_run_large(aero_ctx *ctx, int kind, int logn, int log_blowup, bool input_mont, const DftLaunch &l) {
    const int logm = NTT_MAX_LOG, logB = logn - logm, B = 1 << logB;
    if (logB < 1 || logB > 2) CTX_FAIL(ctx, AERO_ERR_UNSUPPORTED, "transform size 2^%d unsupported (max 2^%d)", logn, NTT_LARGE_MAX_LOG);
    const size_t m = (size_t)1 << logm, n = m << logB;
    const bool inverse = kind == 0;
    const int ncosets_all = inverse ? 1 : (1 << log_blowup);
    const int nc = inverse ? 1 : (l.coset_count ? l.coset_count : ncosets_all), cb = inverse ? 0 : l.coset_begin;
    char key[64];
    const DftTables *sub = nullptr;
    uint64_t *d_shifts = nullptr;
    uint64_t scale = 1;
    if (inverse) {
        TRY(plan_intt(ctx, logm, input_mont, &sub));  // carries 1/m (and 2^-64 for Montgomery input)
        scale = gl::inv((uint64_t)B);
    } else {
        const std::vector<uint64_t> sh = coset_shifts(logn, log_blowup, gl::GENERATOR);
        std::vector<uint64_t> shB(sh.size());
        for (size_t i = 0; i < sh.size(); i++) shB[i] = gl::pow(sh[i], (uint64_t)B);
        snprintf(key, sizeof key, "lde_sub/%d/%d/%d", logn, log_blowup, (int)input_mont);
        TRY(get_plan(ctx, key, logm, false, shB, input_mont ? gl::MONT_R_INV : 1, 0, &sub));
        snprintf(key, sizeof key, "shifts/%d/%d", logn, log_blowup);
        auto it = ctx->const_tables.find(key);
        if (it == ctx->const_tables.end()) {
            TRY(upload_vec(ctx, &d_shifts, sh));
            ctx->const_tables[key] = d_shifts;
        } else {
            d_shifts = it->second;
        }
    }
    PowTable wn;
    snprintf(key, sizeof key, "wn/%d/%d", logn, (int)inverse);
    const uint64_t root = gl::root_of_unity(logn);
    TRY(get_pow_table(ctx, key, inverse ? gl::inv(root) : root, logm, 1, &wn));
    const uint64_t w4 = inverse ? gl::inv(gl::root_of_unity(2)) : gl::root_of_unity(2);
    DevBlocks blk(ctx);
    uint64_t *xd = nullptr, *Y = nullptr, *tmp = nullptr;
    TRY(blk.alloc((void **)&xd, n * 8));
    TRY(blk.alloc((void **)&Y, (size_t)B * nc * m * 8));
    TRY(blk.alloc((void **)&tmp, (size_t)nc * m * 8));
    for (int c = 0; c < l.ncols; c++) {
        large_deinterleave(l.src + (size_t)c * l.src_col_stride, xd, logm, logB, ctx->stream);
        for (int r = 0; r < B; r++) {
            DftLaunch sl;
            sl.src = xd + (size_t)r * m;
            sl.dst = Y + (size_t)r * nc * m;
            sl.tmp = tmp;
            sl.src_col_stride = m;
            sl.dst_col_stride = (size_t)nc * m;
            sl.ncols = 1;
            sl.deinterleave_log = 0;
            sl.coset_begin = cb;
            sl.coset_count = inverse ? 0 : nc;
            dft_run(*sub, sl, ctx->stream);
        }
        large_combine(Y, l.dst + (size_t)c * l.dst_col_stride, wn, d_shifts ? d_shifts + cb : nullptr, scale, w4, logm, logB, nc,
                      ctx->stream);
    }
    CUDA_TRY(ctx, cudaGetLastError());
    return AERO_OK;
}

// -------------------------------------------------------------------------------------------------
// segment
// -------------------------------------------------------------------------------------------------
struct aero_segment {
    aero_ctx *ctx = nullptr;
    int ncols = 0, logn = 0, log_blowup = -1;  // log_blowup < 0: polys only, not yet extended
    uint64_t *polys = nullptr;                 // ncols x n, canonical coefficients (all columns, on every rank)
    uint64_t *lde = nullptr;                   // ncols x (coset_count * n), coset-major: (local coset q, i) -> natural B*i + coset_begin + q
    // Commitment.  With G ranks, rank r owns the leaf block [r*N/G, (r+1)*N/G) = LDE rows of
    // i in [r*nb, (r+1)*nb), nb = n/G, and the subtree above it (merkle/concurrent.rs:21-70 splits the
    // tree the same way); G = 1 is the whole tree.
    uint32_t *leaf_stage = nullptr;            // [B][nb] digests of the block, coset-major: leaf B*il + c at c*nb + il
    uint32_t *heap = nullptr;                  // B*nb digests: the block's subtree in heap layout, heap[1] = sub-root
    uint32_t *top = nullptr;                   // 2G digests: nodes 1 .. 2G-1 of the whole tree, top[1] = root
    // coset shard (multi-GPU): this rank stores cosets [coset_begin, coset_begin + coset_count) only,
    // compactly: lde[c][q][i] with q = r - coset_begin, column stride coset_count * n.
    int coset_begin = 0, coset_count = 0;
    int logG = 0;
    bool failed = false;  // an LDE batch could not be queued (ctx->err holds the reason)
    uint64_t n() const { return 1ULL << logn; }
    uint64_t N() const { return 1ULL << (logn + log_blowup); }
    uint64_t lde_stride() const { return (uint64_t)coset_count << logn; }
    uint32_t nb() const { return 1u << (logn - logG); }
    SegTreeView tree_view() const { return SegTreeView{leaf_stage, heap, top, logn, log_blowup, logG, ctx->shard_rank}; }
};

// a rank's share of the columns of a matrix (interpolation is sharded by column, extension by coset)
static void columns_of_rank(const aero_ctx *ctx, int rank, int ncols, int *cb, int *ce) {
    *cb = (int)((long long)rank * ncols / ctx->shard_world);
    *ce = (int)((long long)(rank + 1) * ncols / ctx->shard_world);
}
static void own_columns(const aero_ctx *ctx, int ncols, int *cb, int *ce) { columns_of_rank(ctx, ctx->shard_rank, ncols, cb, ce); }

// All digests of this rank's leaf block are in `leaf_stage` (after a barrier when other ranks wrote
// some of them): build the block's subtree, exchange the G sub-roots, finish the top levels.
static aero_status segment_finish_tree(aero_segment *seg, uint8_t root[32]) {
    aero_ctx *ctx = seg->ctx;
    const int G = ctx->shard_world;
    TRY(window_barrier(ctx));
    {
        PhaseTimer t(ctx, "merkle");
        merkle_build_block(seg->leaf_stage, seg->heap, seg->nb(), seg->log_blowup, ctx->stream);
        merkle_push_subroot(seg->heap, rank_ptrs(ctx, seg->top), G, ctx->shard_rank, ctx->stream);
    }
    TRY(window_barrier(ctx));
    merkle_top(seg->top, G, ctx->stream);
    CUDA_TRY(ctx, cudaGetLastError());
    // root == NULL: the caller collects the root later (aero_segments_roots), together with the barrier check
    if (root) return download_small_checked(ctx, root, seg->top + 8, 32);
    return AERO_OK;
}

// ---- building blocks of a segment commitment ----------------------------------------------------
static aero_status segment_alloc_lde(aero_segment *seg, int log_blowup, const DftTables **plan) {
    aero_ctx *ctx = seg->ctx;
    if (seg->lde) CTX_FAIL(ctx, AERO_ERR_STATE, "segment already committed");
    if (log_blowup < 1 || log_blowup > 6) CTX_FAIL(ctx, AERO_ERR_UNSUPPORTED, "blowup must be 2..64");
    if (seg->logn + log_blowup > 31) CTX_FAIL(ctx, AERO_ERR_UNSUPPORTED, "LDE domain too large");
    seg->log_blowup = log_blowup;
    const int B = 1 << log_blowup, G = ctx->shard_world;
    if (B % G) CTX_FAIL(ctx, AERO_ERR_UNSUPPORTED, "shard world size %d must divide the blowup factor %d", G, B);
    seg->logG = ilog2((uint64_t)G);
    if (seg->logn < seg->logG) CTX_FAIL(ctx, AERO_ERR_UNSUPPORTED, "trace of %llu rows is too short for %d ranks", (unsigned long long)seg->n(), G);
    seg->coset_count = B / G;
    seg->coset_begin = ctx->shard_rank * seg->coset_count;
    const size_t block = (size_t)seg->N() >> seg->logG;  // leaves of this rank's block
    TRY(dev_alloc(ctx, (void **)&seg->lde, (size_t)seg->ncols * seg->lde_stride() * 8));
    TRY(dev_alloc_shared(ctx, (void **)&seg->leaf_stage, block * 32));
    TRY(dev_alloc_shared(ctx, (void **)&seg->top, (size_t)2 * G * 32));
    TRY(dev_alloc(ctx, (void **)&seg->heap, block * 32));
    *plan = nullptr;  // above the two-pass NTT the extension goes through dft_run_large
    if (seg->logn > NTT_MAX_LOG) return AERO_OK;
    return plan_lde(ctx, seg->logn, log_blowup, false, plan);
}
static int segment_lde_batch_cols(aero_segment *seg) {
    const size_t per_col = (size_t)seg->lde_stride() * 8;
    int batch = (int)std::max<size_t>(1, seg->ctx->lde_batch_bytes / per_col);
    if (batch > 1) batch &= ~1;  // column ranges handed to the row hash start on a 64-byte block boundary
    return std::min(batch, seg->ncols);
}
// coset LDE of columns [c0, c0 + ncols) (tmp: >= ncols * lde_stride entries when two-pass)
static void segment_lde_batch(aero_segment *seg, const DftTables *plan, int c0, int ncols, uint64_t *tmp) {
    aero_ctx *ctx = seg->ctx;
    const uint64_t n = seg->n(), Nl = seg->lde_stride();
    char nm[32];
    snprintf(nm, sizeof nm, "lde_w%d", seg->ncols);
    PhaseTimer t(ctx, nm);
    DftLaunch l;
    l.src = seg->polys + (size_t)c0 * n;
    l.dst = seg->lde + (size_t)c0 * Nl;
    l.tmp = tmp;
    l.src_col_stride = n;
    l.dst_col_stride = Nl;
    l.ncols = ncols;
    l.deinterleave_log = 0;
    l.coset_begin = seg->coset_begin;
    l.coset_count = seg->coset_count;
    if (plan) dft_run(*plan, l, ctx->stream);
    else if (dft_run_large(ctx, 1, seg->logn, seg->log_blowup, false, l) != AERO_OK) seg->failed = true;
}
// Row hashing of columns [c0, c0 + nc) (chaining value kept in the leaf slot, see hash_rows_kernel).
// With ctx->overlap_hash the launch goes to a second stream, ordered after the LDE batch that
// produced those columns, so it shares the SMs with the next batch's NTT.  Off by default: on B200
// the mix is 1-8 % SLOWER than running the two back to back (40.3 ms serial vs 41.6-49.2 ms with
// 4..1 hash blocks per SM) -- IMAD.WIDE blocks the ALU issue port too (tools/int_peak.cu), so the
// NTT passes leave no usable ALU slack for the hash to fill.
static bool segment_hash_overlapped(const aero_segment *seg) {
    return seg->ctx->overlap_hash && seg->ncols > 2 && !ctx_sharded(seg->ctx);
}
static aero_status segment_hash_batch(aero_segment *seg, int c0, int nc) {
    aero_ctx *ctx = seg->ctx;
    const uint64_t Nl = seg->lde_stride();
    cudaStream_t hs = ctx->stream;
    if (segment_hash_overlapped(seg)) {
        if (!ctx->hash_stream) {
            CUDA_TRY(ctx, cudaStreamCreateWithFlags(&ctx->hash_stream, cudaStreamNonBlocking));
            CUDA_TRY(ctx, cudaEventCreateWithFlags(&ctx->ev_lde, cudaEventDisableTiming));
            CUDA_TRY(ctx, cudaEventCreateWithFlags(&ctx->ev_hash, cudaEventDisableTiming));
        }
        hs = ctx->hash_stream;
        CUDA_TRY(ctx, cudaEventRecord(ctx->ev_lde, ctx->stream));
        CUDA_TRY(ctx, cudaStreamWaitEvent(hs, ctx->ev_lde, 0));
    }
    char nm[32];
    snprintf(nm, sizeof nm, "hash_rows_w%d", seg->ncols);
    PhaseTimer t(ctx, nm, hs, true);
    // the last column range has no NTT beside it: give it the whole GPU
    const bool alone = hs == ctx->stream || c0 + nc == seg->ncols;
    hash_rows_lde(seg->lde, Nl, c0, nc, seg->ncols, seg->logn, (uint32_t)seg->coset_begin, (uint32_t)Nl, seg->logn - seg->logG,
                  rank_ptrs(ctx, seg->leaf_stage), alone ? 0 : ctx->hash_blocks_per_sm * ctx->num_sms, hs);
    return AERO_OK;
}
// all column ranges hashed -> join the hash stream, then the tree
static aero_status segment_tree_after_hash(aero_segment *seg, uint8_t root[32]) {
    aero_ctx *ctx = seg->ctx;
    if (segment_hash_overlapped(seg)) {
        CUDA_TRY(ctx, cudaEventRecord(ctx->ev_hash, ctx->hash_stream));
        CUDA_TRY(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_hash, 0));
    }
    CUDA_TRY(ctx, cudaMemsetAsync(seg->heap, 0, 32, ctx->stream));
    return segment_finish_tree(seg, root);
}

// coefficients already on the device (all columns, on every rank) -> LDE + commitment
// (CompositionPoly::evaluate + commit_to_rows)
static aero_status segment_extend_commit(aero_segment *seg, int log_blowup, uint8_t root[32]) {
    aero_ctx *ctx = seg->ctx;
    const DftTables *plan;
    TRY(segment_alloc_lde(seg, log_blowup, &plan));
    const int batch = segment_lde_batch_cols(seg);
    DevBlocks tmp(ctx);
    uint64_t *tmp_l = nullptr;
    if (plan && plan->log1 != 0) TRY(tmp.alloc((void **)&tmp_l, (size_t)batch * seg->lde_stride() * 8));
    if ((batch & 1) || !segment_hash_overlapped(seg)) {  // extend everything, then one hash launch
        for (int c0 = 0; c0 < seg->ncols; c0 += batch) segment_lde_batch(seg, plan, c0, std::min(batch, seg->ncols - c0), tmp_l);
        TRY(segment_hash_batch(seg, 0, seg->ncols));
    } else {
        for (int c0 = 0; c0 < seg->ncols; c0 += batch) {
            const int nc = std::min(batch, seg->ncols - c0);
            segment_lde_batch(seg, plan, c0, nc, tmp_l);
            TRY(segment_hash_batch(seg, c0, nc));
        }
    }
    return segment_tree_after_hash(seg, root);
}

// Column batches of an upload-overlapped commit: [edge, batch, batch, ..., batch, edge] when edge > 0
// (batch - 2*edge... the middle batches are full; the last one takes whatever is left, at most edge
// when the column count allows), else uniform batches.
static int upload_batch_size(int c0, int n_cols, int batch, int edge) {
    if (edge <= 0 || edge >= batch) return batch;
    if (c0 == 0) return edge;
    const int left = n_cols - c0;
    // keep the final batch short: stop the full batches `edge` columns before the end
    if (left > batch + edge) return batch;
    // (even, so that the next batch still starts on a 64-byte block boundary of the row hash)
    if (left > edge) return std::max(2, (left - edge) & ~1);
    return left;
}
// Where the column batches of an upload-overlapped commit come from (aero_segment_commit): prepare(b) makes
// sure the copy of batch b is queued on the copy stream and ev[b] recorded behind it.
struct BatchSource {
    std::vector<cudaEvent_t> ev;
    virtual aero_status prepare(int b) = 0;
    virtual ~BatchSource() {
        for (auto e : ev)
            if (e) cudaEventDestroy(e);
    }
};
// d_src: columns [src_col0, ...) of the matrix (column c at d_src + (c - src_col0) * src_stride), n values
// each in ABI form, on the device; this rank reads its own columns only (own_columns).  They are
// processed in batches (interpolate, extend); when `ready` is given, batch b first waits for ready[b] --
// the event that says its host->device copy has landed -- so uploads overlap compute.  In a sharded
// proof every batch of coefficients is also pushed into the peers' copies of `polys` as soon as it
// exists (NVLink stores under the next batch's upload and transform); after the barrier the columns
// the other ranks interpolated are extended too.
static aero_status segment_from_device(aero_ctx *ctx, const uint64_t *d_src, size_t src_stride, int src_col0, uint32_t n_cols,
                                       uint64_t n_rows, uint32_t blowup, int input_is_coeffs, aero_segment **out,
                                       uint8_t root[32], int batch_cols = 0, BatchSource *ready = nullptr, int edge_cols = 0) {
    if (!out) CTX_FAIL(ctx, AERO_ERR_INVALID, "null output handle");
    if (n_cols == 0 || n_cols > 255) CTX_FAIL(ctx, AERO_ERR_INVALID, "number of columns must be 1..255, got %u", n_cols);
    // Matrix::new (prover/src/matrix.rs:41-64): at least two rows, power of two
    if (n_rows < 2 || !is_pow2(n_rows)) CTX_FAIL(ctx, AERO_ERR_INVALID, "number of rows must be a power of two >= 2, got %llu", (unsigned long long)n_rows);
    if (!is_pow2(blowup) || blowup < 2) CTX_FAIL(ctx, AERO_ERR_INVALID, "blowup factor must be a power of two >= 2, got %u", blowup);
    const int logn = ilog2(n_rows);
    if (logn > NTT_LARGE_MAX_LOG) CTX_FAIL(ctx, AERO_ERR_UNSUPPORTED, "trace length 2^%d unsupported (max 2^%d)", logn, NTT_LARGE_MAX_LOG);
    const bool large = logn > NTT_MAX_LOG;
    SegmentGuard seg(new aero_segment());
    seg->ctx = ctx;
    seg->ncols = (int)n_cols;
    seg->logn = logn;
    const bool mont = ctx->form == AERO_FORM_MONTGOMERY;
    const bool sharded = ctx_sharded(ctx);
    const DftTables *iplan = nullptr, *lplan = nullptr;
    DevBlocks tmp(ctx);
    uint64_t *tmp_i = nullptr, *tmp_l = nullptr;
    TRY(dev_alloc_shared(ctx, (void **)&seg->polys, (size_t)n_cols * n_rows * 8));
    if (!input_is_coeffs && !large) TRY(plan_intt(ctx, logn, mont, &iplan));
    TRY(segment_alloc_lde(seg.get(), ilog2(blowup), &lplan));
    int cb = 0, ce = (int)n_cols;
    own_columns(ctx, (int)n_cols, &cb, &ce);
    const int lde_batch = segment_lde_batch_cols(seg.get());
    const int batch = batch_cols > 0 ? std::min<int>(batch_cols, std::max(1, ce - cb)) : lde_batch;
    // Inputs already on the device: the inverse transforms of many LDE batches share one launch
    // pair (a 16-column interpolation is only ~7 waves of blocks); uploads in flight keep the
    // per-batch order so that batch b waits for ready[b] only.
    int ibatch = batch;
    if (!ready && !input_is_coeffs) {
        const size_t fit = ctx->lde_batch_bytes / ((size_t)n_rows * 8);
        ibatch = (int)std::min<size_t>(n_cols, std::max<size_t>(batch, fit / batch * batch));
    }
    if (iplan && iplan->log1 != 0) TRY(tmp.alloc((void **)&tmp_i, (size_t)ibatch * n_rows * 8));
    if (lplan && lplan->log1 != 0) TRY(tmp.alloc((void **)&tmp_l, (size_t)std::max(batch, lde_batch) * seg->lde_stride() * 8));
    const bool per_batch_hash = !(batch & 1) && segment_hash_overlapped(seg.get());
    // Uploads in flight deliver a column batch slightly slower than it is extended, so the stream would wait
    // a little before every batch (profiles/r02_trace_gaps_host_kernels.txt: 1.2 ms per 72-column segment).
    // Hashing the first batches' columns early (chained row hash, same stream) builds a backlog of arrived
    // batches instead; the rest of the columns are hashed once at the end as before.
    const int early_batches = (ready && !per_batch_hash && !sharded && n_cols >= 32) ? ctx->hash_early_batches : 0;
    int hashed_cols = 0;
    char nm[32];
    snprintf(nm, sizeof nm, "interpolate_w%d", (int)n_cols);
    const RankPtrs polys_all = rank_ptrs(ctx, seg->polys);
    // edge_cols > 0 (uploads in flight): the first and the last batch are short -- the first so that
    // compute starts early, the last so that little work is left once the final copy has landed
    for (int c0 = cb, b = 0, nc = 0; c0 < ce; c0 += nc, b++) {
        nc = std::min(upload_batch_size(c0 - cb, ce - cb, batch, edge_cols), ce - c0);
        if (ready) {  // batch b's host->device copy: queue it if that has not happened yet, then order after it
            TRY(ready->prepare(b));
            CUDA_TRY(ctx, cudaStreamWaitEvent(ctx->stream, ready->ev[b], 0));
        }
        const uint64_t *src = d_src + (size_t)(c0 - src_col0) * src_stride;
        int done = nc;  // columns whose coefficients exist after this step, starting at c0
        if (input_is_coeffs) {
            uint64_t *dst = seg->polys + (size_t)c0 * n_rows;
            PhaseTimer t(ctx, "convert");
            for (int c = 0; c < nc; c++) {
                if (mont) convert_form(src + (size_t)c * src_stride, dst + (size_t)c * n_rows, n_rows, 0, ctx->stream);
                else CUDA_TRY(ctx, cudaMemcpyAsync(dst + (size_t)c * n_rows, src + (size_t)c * src_stride, n_rows * 8, cudaMemcpyDeviceToDevice, ctx->stream));
            }
        } else if (edge_cols > 0 || (c0 - cb) % ibatch == 0) {
            PhaseTimer t(ctx, nm);
            DftLaunch l;
            l.src = src;
            l.dst = seg->polys + (size_t)c0 * n_rows;
            l.tmp = tmp_i;
            l.src_col_stride = src_stride;
            l.dst_col_stride = n_rows;
            l.ncols = done = edge_cols > 0 ? nc : std::min(ibatch, ce - c0);
            l.deinterleave_log = 0;
            if (iplan) dft_run(*iplan, l, ctx->stream);
            else TRY(dft_run_large(ctx, 0, logn, 0, mont, l));
        } else {
            done = 0;  // part of an earlier, wider interpolation launch
        }
        if (sharded && done && ctx->win_host_sync) {
            PhaseTimer t(ctx, "push_polys");
            peer_push(polys_all, ctx->shard_world, ctx->shard_rank, (size_t)c0 * n_rows * 8, (size_t)done * n_rows * 8, ctx->stream);
        }
        if (sharded && !ctx->win_host_sync && c0 + nc == ce) CUDA_TRY(ctx, cudaEventRecord(ctx->ev_push, ctx->stream));  // all own coefficients exist
        segment_lde_batch(seg.get(), lplan, c0, nc, tmp_l);
        if (per_batch_hash) TRY(segment_hash_batch(seg.get(), c0, nc));
        else if (b < early_batches && c0 == hashed_cols && !((c0 + nc) & 1) && c0 + nc < ce) {
            TRY(segment_hash_batch(seg.get(), hashed_cols, c0 + nc - hashed_cols));
            hashed_cols = c0 + nc;
        }
    }
    if (sharded && ctx->win_host_sync) {  // the other ranks' coefficients have arrived: extend those columns too
        TRY(window_barrier(ctx));
        for (int part = 0; part < 2; part++) {
            const int lo = part ? ce : 0, hi = part ? (int)n_cols : cb;
            for (int c0 = lo; c0 < hi; c0 += lde_batch) segment_lde_batch(seg.get(), lplan, c0, std::min(lde_batch, hi - c0), tmp_l);
        }
    } else if (sharded) {
        // Warm shape: the exchange runs on its own stream beside the LDE.  Rank r sends its columns to
        // r+1, r+2, ... in turn, each transfer at the full NVLink rate of the sender, and raises an arrival
        // flag at the receiver after each; so rank r receives from r-1 first, then r-2, ... and extends
        // each rank's columns as soon as they are there, while the later ones are still in flight.
        const int G = ctx->shard_world, me = ctx->shard_rank;
        const unsigned long long epoch = ++ctx->push_epoch;
        if (ce == cb) CUDA_TRY(ctx, cudaEventRecord(ctx->ev_push, ctx->stream));
        const int parts = ctx->push_parts;
        const size_t own_bytes = (size_t)(ce - cb) * n_rows * 8, piece = ((own_bytes / parts) + 15) & ~(size_t)15;
        for (int part = 0; part < parts; part++) {
            cudaStream_t ps = ctx->push_stream[part];
            CUDA_TRY(ctx, cudaStreamWaitEvent(ps, ctx->ev_push, 0));
            PhaseTimer t(ctx, "push_polys", ps, true);
            const size_t p0 = std::min(own_bytes, piece * part), p1 = part + 1 == parts ? own_bytes : std::min(own_bytes, piece * (part + 1));
            for (int k = 1; k < G; k++) {
                const int dest = (me + k) % G;
                peer_send(polys_all, me, dest, (size_t)cb * n_rows * 8 + p0, p1 - p0,
                          (unsigned long long *)(ctx->win_base[dest] + 1024) + me * PUSH_PARTS + part, epoch, ps);
            }
        }
        for (int k = 1; k < G; k++) {
            const int src = (me - k + G) % G;
            int lo = 0, hi = 0;
            columns_of_rank(ctx, src, (int)n_cols, &lo, &hi);
            peer_wait((unsigned long long *)(ctx->win + 1024) + src * PUSH_PARTS, parts, epoch, (unsigned int *)(ctx->win + 2048), ctx->stream);
            for (int c0 = lo; c0 < hi; c0 += lde_batch) segment_lde_batch(seg.get(), lplan, c0, std::min(lde_batch, hi - c0), tmp_l);
        }
        CUDA_TRY(ctx, cudaGetLastError());
    }
    if (seg->failed) return AERO_ERR_UNSUPPORTED;
    if (!per_batch_hash) TRY(segment_hash_batch(seg.get(), hashed_cols, (int)n_cols - hashed_cols));
    TRY(segment_tree_after_hash(seg.get(), root));
    *out = seg.release();
    return AERO_OK;
}

// -------------------------------------------------------------------------------------------------
// Merkle batch-proof index list (crypto/src/merkle/mod.rs:188-250 restated over heap indices:
// leaf i is full[N + i], internal node j is full[j])
// -------------------------------------------------------------------------------------------------
static aero_status batch_proof_indices(aero_ctx *ctx, const uint64_t *positions, uint32_t n_pos, uint64_t N,
                                       std::vector<std::vector<uint32_t>> &out) {
    if (n_pos == 0) CTX_FAIL(ctx, AERO_ERR_INVALID, "at least one index is required");          // TooFewLeafIndexes
    if (n_pos > 255) CTX_FAIL(ctx, AERO_ERR_INVALID, "at most 255 indexes are allowed");        // TooManyLeafIndexes
    std::set<uint64_t> index_set;
    for (uint32_t i = 0; i < n_pos; i++) {
        if (positions[i] >= N) CTX_FAIL(ctx, AERO_ERR_INVALID, "leaf index %llu out of bounds (%llu leaves)", (unsigned long long)positions[i], (unsigned long long)N);
        if (!index_set.insert(positions[i]).second) CTX_FAIL(ctx, AERO_ERR_INVALID, "duplicate leaf index %llu", (unsigned long long)positions[i]);
    }
    std::set<uint64_t> norm;
    for (uint32_t i = 0; i < n_pos; i++) norm.insert(positions[i] & ~1ULL);
    out.clear();
    std::vector<uint64_t> next;
    for (uint64_t index : norm) {
        std::vector<uint32_t> v;
        for (uint64_t i = index; i < index + 2; i++)
            if (!index_set.count(i)) v.push_back((uint32_t)(N + i));
        out.push_back(v);
        next.push_back((index + N) >> 1);
    }
    const int depth = ilog2(N);
    for (int d = 1; d < depth; d++) {
        std::vector<uint64_t> cur;
        cur.swap(next);
        for (size_t i = 0; i < cur.size(); i++) {
            const uint64_t sib = cur[i] ^ 1;
            if (i + 1 < cur.size() && cur[i + 1] == sib) {
                i++;
            } else {
                out[i].push_back((uint32_t)sib);
            }
            next.push_back(sib >> 1);
        }
    }
    return AERO_OK;
}
// NOTE: in the reference loop `nodes[i].push(...)` uses the position i inside the *current* level's
// index list (merkle/mod.rs:226-240), including after `i += 1` skipped a sibling; the loop above
// mirrors that: when a sibling pair is merged, `i` has already advanced, exactly as in the Rust.

// BatchMerkleProof::serialize_nodes (merkle/proofs.rs:421-439) over the gathered digests `dig`
// (32 bytes per entry of idx, in order)
static aero_status serialize_batch_proof(aero_ctx *ctx, const std::vector<std::vector<uint32_t>> &idx, const uint8_t *dig,
                                         std::vector<uint8_t> &bytes) {
    bytes.clear();
    bytes.push_back((uint8_t)idx.size());
    size_t off = 0;
    for (auto &v : idx) {
        if (v.size() > 255) CTX_FAIL(ctx, AERO_ERR_INVALID, "too many nodes in a batch proof vector");
        bytes.push_back((uint8_t)v.size());
        bytes.insert(bytes.end(), dig + off * 32, dig + (off + v.size()) * 32);
        off += v.size();
    }
    return AERO_OK;
}

// -------------------------------------------------------------------------------------------------
// FRI handle
// -------------------------------------------------------------------------------------------------
struct FriLayerDev {
    uint64_t *evals = nullptr;  // M evaluations
    uint32_t M = 0;
    int log_cosets = 0;         // storage layout of evals
    uint32_t *full = nullptr;   // 2*rows digests
};
struct aero_fri {
    aero_ctx *ctx = nullptr;
    std::vector<FriLayerDev> layers;  // committed layers
    uint64_t *cur = nullptr;
    uint32_t curM = 0;
    int cur_log_cosets = 0;
    bool cur_committed = false;
};


// -------------------------------------------------------------------------------------------------
// Openings.  The query phase gathers a few hundred digests and rows out of trees and matrices that
// stay on the device.  A GatherBatch collects every index list first, uploads them in one copy, runs
// the gather kernels back to back and downloads all results with one copy and ONE stream
// synchronisation (a proof used to pay 17 host round trips of 20-100 us here; profiles/r01_trace_gaps_v5.txt).
// -------------------------------------------------------------------------------------------------
struct GatherBatch {
    enum Kind { DIGESTS, TREE_DIGESTS, SEG_ROWS, FRI_ROWS, COPY, COPY_NATURAL };
    struct Job {
        Kind kind;
        const void *src;
        const aero_segment *seg;
        uint32_t rows;
        int log_cosets;
        size_t idx_off;   // first index (u32 units)
        int count;
        size_t out_off, out_bytes;
    };
    aero_ctx *ctx;
    std::vector<uint32_t> idx;
    std::vector<Job> jobs;
    size_t out_bytes = 0;
    const uint8_t *host = nullptr;  // results, valid until the next batch of this context runs
    explicit GatherBatch(aero_ctx *c) : ctx(c) {}
    size_t push(Kind k, const void *src, const aero_segment *seg, uint32_t rows, int log_cosets, const uint32_t *ix,
                size_t n, size_t bytes) {
        Job j{k, src, seg, rows, log_cosets, idx.size(), (int)n, out_bytes, bytes};
        if (ix) idx.insert(idx.end(), ix, ix + n);
        out_bytes += (bytes + 15) & ~(size_t)15;
        jobs.push_back(j);
        return j.out_off;
    }
    // each returns the offset of its result inside `host`
    size_t digests(const uint32_t *full, const std::vector<std::vector<uint32_t>> &lists) {
        std::vector<uint32_t> flat;
        for (auto &v : lists) flat.insert(flat.end(), v.begin(), v.end());
        return push(DIGESTS, full, nullptr, 0, 0, flat.data(), flat.size(), flat.size() * 32);
    }
    size_t tree_digests(const aero_segment *seg, const std::vector<std::vector<uint32_t>> &lists) {
        std::vector<uint32_t> flat;
        for (auto &v : lists) flat.insert(flat.end(), v.begin(), v.end());
        return push(TREE_DIGESTS, nullptr, seg, 0, 0, flat.data(), flat.size(), flat.size() * 32);
    }
    size_t segment_rows(const aero_segment *seg, const std::vector<uint32_t> &pos);
    size_t fri_rows(const FriLayerDev &L, const std::vector<uint32_t> &pos) {
        return push(FRI_ROWS, L.evals, nullptr, L.M / 8, L.log_cosets, pos.data(), pos.size(), pos.size() * 64);
    }
    size_t copy(const void *d_src, size_t bytes) { return push(COPY, d_src, nullptr, 0, 0, nullptr, 0, bytes); }
    // `count` evaluations stored coset-major (2^log_cosets cosets), delivered in natural order
    size_t copy_natural(const uint64_t *d_src, uint32_t count, int log_cosets) {
        return push(COPY_NATURAL, d_src, nullptr, count, log_cosets, nullptr, 0, (size_t)count * 8);
    }
    aero_status run();
};

size_t GatherBatch::segment_rows(const aero_segment *seg, const std::vector<uint32_t> &pos) {
    return push(SEG_ROWS, seg->lde, seg, 0, 0, pos.data(), pos.size(), pos.size() * (size_t)seg->ncols * 8);
}
// Results of a sharded proof are assembled in the exchange window: each entry is written by the rank
// that owns it into every rank's result buffer (gather_rows / gather_tree_digests), one barrier later
// all ranks hold everything; replicated sources (FRI layers) are gathered locally.
aero_status GatherBatch::run() {
    const size_t idx_bytes = (idx.size() * 4 + 15) & ~(size_t)15;
    const size_t total = idx_bytes + out_bytes;
    TRY(stage_reserve(ctx, total));
    uint8_t *d = ctx->d_stage;
    DevBlocks win(ctx);
    uint8_t *d_res = d + idx_bytes;
    if (ctx_sharded(ctx) && out_bytes) TRY(win.alloc_shared((void **)&d_res, out_bytes));
    const RankPtrs res_all = rank_ptrs(ctx, d_res);
    if (!idx.empty()) {
        memcpy(ctx->h_stage, idx.data(), idx.size() * 4);
        CUDA_TRY(ctx, cudaMemcpyAsync(d, ctx->h_stage, idx.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
    }
    for (const Job &j : jobs) {
        if (j.out_bytes == 0) continue;
        const uint32_t *d_idx = (const uint32_t *)d + j.idx_off;
        uint8_t *d_out = d_res + j.out_off;
        RankPtrs out_all = res_all;
        for (int r = 0; r < AERO_MAX_RANKS; r++)
            if (out_all.p[r]) out_all.p[r] = (uint8_t *)out_all.p[r] + j.out_off;
        switch (j.kind) {
        case DIGESTS: gather_digests((const uint32_t *)j.src, d_idx, j.count, (uint32_t *)d_out, ctx->stream); break;
        case TREE_DIGESTS: gather_tree_digests(j.seg->tree_view(), d_idx, j.count, out_all, ctx->stream); break;
        case SEG_ROWS:
            gather_rows(j.seg->lde, j.seg->lde_stride(), j.seg->ncols, j.seg->logn, j.seg->log_blowup, j.seg->coset_begin,
                        j.seg->coset_count, ctx->shard_world, d_idx, j.count, out_all, ctx->stream);
            break;
        case FRI_ROWS:
            gather_fri_rows((const uint64_t *)j.src, j.rows, j.log_cosets, d_idx, j.count, (uint64_t *)d_out, ctx->stream);
            break;
        case COPY: CUDA_TRY(ctx, cudaMemcpyAsync(d_out, j.src, j.out_bytes, cudaMemcpyDeviceToDevice, ctx->stream)); break;
        case COPY_NATURAL:
            lde_to_natural((const uint64_t *)j.src, (uint64_t *)d_out, ilog2(j.rows) - j.log_cosets, j.log_cosets, 0, ctx->stream);
            break;
        }
    }
    CUDA_TRY(ctx, cudaGetLastError());
    TRY(window_barrier(ctx));
    const unsigned int *word = nullptr;
    TRY(window_check_queue(ctx, &word));
    if (out_bytes)
        CUDA_TRY(ctx, cudaMemcpyAsync(ctx->h_stage + idx_bytes, d_res, out_bytes, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    TRY(window_check_result(ctx, word));
    host = ctx->h_stage + idx_bytes;
    return AERO_OK;
}

// One segment's query: index lists now, bytes after the batch ran.
struct SegmentOpening {
    aero_segment *seg = nullptr;
    std::vector<std::vector<uint32_t>> idx;
    uint32_t n_pos = 0;
    size_t dig_off = 0, rows_off = 0;
    bool want_rows = false;
};
static aero_status segment_open_plan(aero_segment *seg, const uint64_t *positions, uint32_t n_pos, bool want_rows,
                                     GatherBatch &gb, SegmentOpening &o) {
    aero_ctx *ctx = seg->ctx;
    if (!seg->heap) CTX_FAIL(ctx, AERO_ERR_STATE, "segment has no commitment");
    o.seg = seg;
    o.n_pos = n_pos;
    o.want_rows = want_rows;
    TRY(batch_proof_indices(ctx, positions, n_pos, seg->N(), o.idx));
    o.dig_off = gb.tree_digests(seg, o.idx);
    if (want_rows) {
        std::vector<uint32_t> pos(n_pos);
        for (uint32_t i = 0; i < n_pos; i++) pos[i] = (uint32_t)positions[i];
        o.rows_off = gb.segment_rows(seg, pos);
    }
    return AERO_OK;
}
static aero_status segment_open_finish(const SegmentOpening &o, const GatherBatch &gb, uint64_t *rows_out,
                                       uint8_t *batch_nodes_out, size_t *len) {
    aero_ctx *ctx = o.seg->ctx;
    std::vector<uint8_t> bytes;
    TRY(serialize_batch_proof(ctx, o.idx, gb.host + o.dig_off, bytes));
    if (!batch_nodes_out || *len < bytes.size()) {
        *len = bytes.size();
        CTX_FAIL(ctx, AERO_ERR_BUFFER, "batch proof needs %zu bytes", bytes.size());
    }
    memcpy(batch_nodes_out, bytes.data(), bytes.size());
    *len = bytes.size();
    if (o.want_rows) memcpy(rows_out, gb.host + o.rows_off, (size_t)o.n_pos * o.seg->ncols * 8);
    return AERO_OK;
}

// fold_positions (fri/src/folding/mod.rs:159-176)
static std::vector<uint64_t> fold_positions(const std::vector<uint64_t> &pos, uint64_t source, uint64_t ff) {
    const uint64_t target = source / ff;
    std::vector<uint64_t> r;
    for (uint64_t p : pos) {
        p %= target;
        if (std::find(r.begin(), r.end(), p) == r.end()) r.push_back(p);
    }
    return r;
}
// FriProver::build_proof (fri/src/prover/mod.rs:231-275), same split
struct FriOpening {
    aero_fri *fri = nullptr;
    struct Layer {
        std::vector<std::vector<uint32_t>> idx;
        size_t n_pos = 0, vals_off = 0, dig_off = 0;
    };
    std::vector<Layer> layers;
    size_t rem_off = 0;
};
static aero_status fri_open_plan(aero_fri *fri, const uint64_t *positions, uint32_t n_pos, GatherBatch &gb, FriOpening &o) {
    aero_ctx *ctx = fri->ctx;
    if (fri->layers.empty()) CTX_FAIL(ctx, AERO_ERR_STATE, "FRI layers have not been built yet");  // prover/mod.rs:232-235
    if (n_pos == 0 || n_pos > 255) CTX_FAIL(ctx, AERO_ERR_INVALID, "number of positions must be 1..255");
    std::vector<uint64_t> pos(positions, positions + n_pos);
    for (uint64_t p : pos)
        if (p >= fri->layers[0].M) CTX_FAIL(ctx, AERO_ERR_INVALID, "query position out of range");
    const FriLayerDev &R = fri->layers.back();
    if ((size_t)R.M * 8 > 0xFFFF) CTX_FAIL(ctx, AERO_ERR_INVALID, "remainder too large for the wire format");
    o.fri = fri;
    const size_t nl = fri->layers.size() - 1;
    o.layers.resize(nl);
    uint64_t domain = fri->layers[0].M;
    for (size_t i = 0; i < nl; i++) {
        const FriLayerDev &L = fri->layers[i];
        pos = fold_positions(pos, domain, 8);
        // queried values: [E; 8] rows at the folded positions, canonical bytes
        std::vector<uint32_t> p32(pos.begin(), pos.end());
        o.layers[i].n_pos = p32.size();
        o.layers[i].vals_off = gb.fri_rows(L, p32);
        TRY(batch_proof_indices(ctx, pos.data(), (uint32_t)pos.size(), L.M / 8, o.layers[i].idx));
        o.layers[i].dig_off = gb.digests(L.full, o.layers[i].idx);
        domain /= 8;
    }
    // remainder = last committed layer in natural order (prover/mod.rs:258-268 un-transposes the
    // stored transposed copy; ours is stored natural already)
    // (with no FRI layer at all -- LDE domain <= max remainder size -- it is the coset-major DEEP layer)
    o.rem_off = R.log_cosets ? gb.copy_natural(R.evals, R.M, R.log_cosets) : gb.copy(R.evals, (size_t)R.M * 8);
    return AERO_OK;
}
static aero_status fri_open_finish(const FriOpening &o, const GatherBatch &gb, uint8_t *out_bytes, size_t *len) {
    aero_ctx *ctx = o.fri->ctx;
    std::vector<uint8_t> bytes;
    bytes.push_back((uint8_t)o.layers.size());
    for (const FriOpening::Layer &l : o.layers) {
        std::vector<uint8_t> paths;
        TRY(serialize_batch_proof(ctx, l.idx, gb.host + l.dig_off, paths));
        // FriProofLayer::write_into (fri/src/proof.rs:351-359)
        const uint32_t vlen = (uint32_t)(l.n_pos * 64), plen = (uint32_t)paths.size();
        bytes.insert(bytes.end(), (uint8_t *)&vlen, (uint8_t *)&vlen + 4);
        bytes.insert(bytes.end(), gb.host + l.vals_off, gb.host + l.vals_off + vlen);
        bytes.insert(bytes.end(), (uint8_t *)&plen, (uint8_t *)&plen + 4);
        bytes.insert(bytes.end(), paths.begin(), paths.end());
    }
    const uint16_t rl = (uint16_t)(o.fri->layers.back().M * 8);
    bytes.insert(bytes.end(), (uint8_t *)&rl, (uint8_t *)&rl + 2);
    bytes.insert(bytes.end(), gb.host + o.rem_off, gb.host + o.rem_off + rl);
    bytes.push_back(0);  // log2(num_partitions = 1), fri/src/proof.rs:50-52
    if (!out_bytes || *len < bytes.size()) {
        *len = bytes.size();
        CTX_FAIL(ctx, AERO_ERR_BUFFER, "FRI proof needs %zu bytes", bytes.size());
    }
    memcpy(out_bytes, bytes.data(), bytes.size());
    *len = bytes.size();
    return AERO_OK;
}

// -------------------------------------------------------------------------------------------------
// extern "C"
// -------------------------------------------------------------------------------------------------
// every entry point runs on the context's device, whatever the calling thread had current
static inline void enter(const aero_ctx *ctx) { cudaSetDevice(ctx->device); }

extern "C" {

const char *aero_version(void) { return "aero_b200 0.1 (sm_100a)"; }
uint64_t aero_launch_count(void) { return aero::g_launch_count.load(); }

aero_status aero_ctx_create(const int *device_ids, int n_devices, aero_ctx **out) {
    if (!out) return AERO_ERR_INVALID;
    *out = nullptr;
    if (n_devices != 1 && !(n_devices == 0 && device_ids == nullptr)) return AERO_ERR_UNSUPPORTED;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) return AERO_ERR_CUDA;  // no CPU fallback
    int dev = 0;
    if (device_ids) dev = device_ids[0];
    else if (cudaGetDevice(&dev) != cudaSuccess) return AERO_ERR_CUDA;
    if (dev < 0 || dev >= count) return AERO_ERR_INVALID;
    if (cudaSetDevice(dev) != cudaSuccess) return AERO_ERR_CUDA;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, dev) != cudaSuccess) return AERO_ERR_CUDA;
    if (prop.major < 10) return AERO_ERR_UNSUPPORTED;  // kernels are built for sm_100a only
    aero_ctx *ctx = new aero_ctx();
    ctx->device = dev;
    if (const char *e = getenv("AERO_FRI_FUSED")) ctx->fri_fused = std::max(0, std::min(2, atoi(e)));  // experiment hooks
    if (const char *e = getenv("AERO_HASH_EARLY")) ctx->hash_early_batches = std::max(0, std::min(16, atoi(e)));
    if (const char *e = getenv("AERO_NTT_OUTER")) {  // experiment hook: default of the "ntt_outer_log" option
        const int v = atoi(e);
        if (v >= -1 && v <= NTT_OUTER_MAX_LOG) ctx->ntt_outer_log = v;
    }
    ctx->num_sms = prop.multiProcessorCount;
    size_t free_b = 0, total_b = 0;
    // released blocks stay cached up to 90 % of the device (a 2^24-row proof holds ~130 GB; with the old
    // 50 % limit every such proof paid ~90 ms of cudaFree + cudaMalloc); a failing cudaMalloc drops the cache
    if (cudaMemGetInfo(&free_b, &total_b) == cudaSuccess) ctx->cache_limit_bytes = total_b / 10 * 9;
    *out = ctx;
    return AERO_OK;
}
void aero_ctx_destroy(aero_ctx *ctx) {
    if (!ctx) return;
    enter(ctx);
    cudaStreamSynchronize(ctx->stream);
    profile_flush(ctx);
    if (ctx->copy_stream) {
        cudaStreamSynchronize(ctx->copy_stream);
        cudaStreamDestroy(ctx->copy_stream);
        cudaEventDestroy(ctx->ev_copy_order);
    }
    if (ctx->hash_stream) {
        cudaStreamSynchronize(ctx->hash_stream);
        cudaStreamDestroy(ctx->hash_stream);
        cudaEventDestroy(ctx->ev_lde);
        cudaEventDestroy(ctx->ev_hash);
    }
    if (ctx->push_stream[0]) {
        for (cudaStream_t ps : ctx->push_stream) {
            cudaStreamSynchronize(ps);
            cudaStreamDestroy(ps);
        }
        cudaEventDestroy(ctx->ev_push);
    }
    for (int r = 0; r < AERO_MAX_RANKS; r++)
        if (ctx->win_ipc[r]) cudaIpcCloseMemHandle(ctx->win_base[r]);
    if (ctx->win) cudaFree(ctx->win);
    cache_release_all(ctx);
    for (auto &kv : ctx->live_blocks) cudaFree(kv.first);  // handles the caller forgot to destroy
    for (void *p : ctx->owned) cudaFree(p);
    if (ctx->h_stage) cudaFreeHost(ctx->h_stage);
    if (ctx->d_stage) cudaFree(ctx->d_stage);
    if (ctx->h_ring) cudaFreeHost(ctx->h_ring);
    for (int i = 0; i < 2; i++) {
        if (ctx->h_bulk[i]) cudaFreeHost(ctx->h_bulk[i]);
        if (ctx->ev_bulk[i]) cudaEventDestroy(ctx->ev_bulk[i]);
    }
    for (cudaEvent_t e : ctx->event_pool) cudaEventDestroy(e);
    if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}
const char *aero_last_error(aero_ctx *ctx) { return ctx ? ctx->err.c_str() : "null context"; }
aero_status aero_ctx_set_stream(aero_ctx *ctx, void *s) {
    if (!ctx) return AERO_ERR_INVALID;
    enter(ctx);
    // cached blocks are handed out on the assumption that all work is ordered on one stream
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
    ctx->own_stream = false;
    ctx->stream = (cudaStream_t)s;
    return AERO_OK;
}
aero_status aero_ctx_set_form(aero_ctx *ctx, int form) {
    if (!ctx) return AERO_ERR_INVALID;
    enter(ctx);
    if (form != AERO_FORM_MONTGOMERY && form != AERO_FORM_CANONICAL) CTX_FAIL(ctx, AERO_ERR_INVALID, "unknown element form %d", form);
    ctx->form = form;
    return AERO_OK;
}
int aero_ctx_get_form(aero_ctx *ctx) { return ctx ? ctx->form : -1; }
aero_status aero_ctx_set_option(aero_ctx *ctx, const char *key, long long value) {
    if (!ctx || !key) return AERO_ERR_INVALID;
    const std::string k(key);
    if (k == "overlap_hash") ctx->overlap_hash = value != 0;
    else if (k == "own_stream") {
        if (value && !ctx->own_stream) {
            cudaStream_t st = nullptr;
            CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
            CUDA_TRY(ctx, cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
            ctx->stream = st;
            ctx->own_stream = true;
        }
    }
    else if (k == "force_host_sync") ctx->force_host_sync = value != 0;
    else if (k == "push_parts" && value >= 1 && value <= PUSH_PARTS) ctx->push_parts = (int)value;
    else if (k == "force_split_intt") ctx->force_split_intt = value != 0;
    else if (k == "air_blocks_per_sm") ctx->air_blocks_per_sm = (int)value;
    else if (k == "hash_blocks_per_sm" && value > 0) ctx->hash_blocks_per_sm = (int)value;
    else if (k == "lde_batch_bytes" && value > 0) ctx->lde_batch_bytes = (size_t)value;
    else if (k == "ntt_table_max_bytes" && value >= 0) ctx->ntt_table_max_bytes = (size_t)value;
    else if (k == "fri_fused" && value >= 0 && value <= 2) ctx->fri_fused = (int)value;
    else if (k == "hash_early_batches" && value >= 0 && value <= 16) ctx->hash_early_batches = (int)value;
    else if (k == "ntt_outer_log" && value >= -1 && value <= NTT_OUTER_MAX_LOG) ctx->ntt_outer_log = (int)value;
    else if (k == "upload_batch_cols" && value >= 1 && value <= 255) ctx->upload_batch_cols = (int)value;
    else if (k == "upload_edge_cols" && value >= -1 && value <= 255) ctx->upload_edge_cols = (int)value;
    else if (k == "cache_limit_bytes" && value >= 0) {
        ctx->cache_limit_bytes = (size_t)value;
        if (ctx->cached_bytes > ctx->cache_limit_bytes) cache_release_all(ctx);
    }
    else CTX_FAIL(ctx, AERO_ERR_INVALID, "unknown option '%s'", key);
    return AERO_OK;
}
void aero_ctx_set_error(aero_ctx *ctx, const char *msg) {
    if (ctx) ctx->err = msg ? msg : "";
}
aero_status aero_ctx_profile_enable(aero_ctx *ctx, int enable) {
    if (!ctx) return AERO_ERR_INVALID;
    enter(ctx);
    ctx->profile = enable != 0;
    return AERO_OK;
}
aero_status aero_ctx_profile_filter(aero_ctx *ctx, const char *prefix) {
    if (!ctx) return AERO_ERR_INVALID;
    enter(ctx);
    ctx->profile_prefix = prefix ? prefix : "";
    return AERO_OK;
}
aero_status aero_ctx_profile_read(aero_ctx *ctx, char *json_out, size_t *len) {
    if (!ctx || !len) return AERO_ERR_INVALID;
    profile_flush(ctx);
    std::string s = "{";
    bool first = true;
    for (auto &kv : ctx->stats) {
        char b[160];
        snprintf(b, sizeof b, "%s\"%s\": [%d, %.6f]", first ? "" : ", ", kv.first.c_str(), kv.second.calls, kv.second.ms);
        s += b;
        first = false;
    }
    s += "}";
    if (!json_out || *len < s.size() + 1) {
        *len = s.size() + 1;
        CTX_FAIL(ctx, AERO_ERR_BUFFER, "profile buffer too small");
    }
    memcpy(json_out, s.c_str(), s.size() + 1);
    *len = s.size() + 1;
    ctx->stats.clear();
    return AERO_OK;
}

aero_status aero_device_alloc(aero_ctx *ctx, size_t bytes, void **d_ptr) {
    if (!ctx || !d_ptr) return AERO_ERR_INVALID;
    CUDA_TRY(ctx, cudaMalloc(d_ptr, bytes ? bytes : 8));
    return AERO_OK;
}
aero_status aero_device_free(aero_ctx *ctx, void *d_ptr) {
    if (!ctx) return AERO_ERR_INVALID;
    enter(ctx);
    CUDA_TRY(ctx, cudaFree(d_ptr));
    return AERO_OK;
}
aero_status aero_device_upload(aero_ctx *ctx, void *d_dst, const void *h_src, size_t bytes) {
    if (!ctx) return AERO_ERR_INVALID;
    enter(ctx);
    CUDA_TRY(ctx, cudaMemcpyAsync(d_dst, h_src, bytes, cudaMemcpyHostToDevice, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return AERO_OK;
}
aero_status aero_device_download(aero_ctx *ctx, void *h_dst, const void *d_src, size_t bytes) {
    if (!ctx) return AERO_ERR_INVALID;
    enter(ctx);
    CUDA_TRY(ctx, cudaMemcpyAsync(h_dst, d_src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return AERO_OK;
}
aero_status aero_device_sync(aero_ctx *ctx) {
    if (!ctx) return AERO_ERR_INVALID;
    enter(ctx);
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    CUDA_TRY(ctx, cudaGetLastError());
    return AERO_OK;
}

aero_status aero_measure_alu_peak(aero_ctx *ctx, double *lane_ops_per_s) {
    if (!ctx || !lane_ops_per_s) return AERO_ERR_INVALID;
    enter(ctx);
    DevBlocks blk(ctx);
    uint32_t *scratch = nullptr;
    TRY(blk.alloc((void **)&scratch, (size_t)ctx->num_sms * 4 * 256 * 4));
    *lane_ops_per_s = measure_alu_peak(ctx->num_sms, scratch, ctx->stream);
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    CUDA_TRY(ctx, cudaGetLastError());
    return AERO_OK;
}

aero_status aero_test_field_ops(aero_ctx *ctx, const uint64_t *a, const uint64_t *b, size_t n, uint64_t *out) {
    if (!ctx || !a || !b || !out || !n) return AERO_ERR_INVALID;
    DevBlocks blk(ctx);
    uint64_t *da = nullptr, *db = nullptr, *dout = nullptr;
    TRY(blk.alloc((void **)&da, n * 8));
    TRY(blk.alloc((void **)&db, n * 8));
    TRY(blk.alloc((void **)&dout, 12 * n * 8));
    CUDA_TRY(ctx, cudaMemcpyAsync(da, a, n * 8, cudaMemcpyHostToDevice, ctx->stream));
    CUDA_TRY(ctx, cudaMemcpyAsync(db, b, n * 8, cudaMemcpyHostToDevice, ctx->stream));
    field_ops(da, db, n, dout, ctx->stream);
    CUDA_TRY(ctx, cudaMemcpyAsync(out, dout, 12 * n * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return AERO_OK;
}

// ---- prefetched uploads ------------------------------------------------------------------------
static aero_status ensure_copy_stream(aero_ctx *ctx) {
    if (!ctx->copy_stream) {
        CUDA_TRY(ctx, cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
        CUDA_TRY(ctx, cudaEventCreateWithFlags(&ctx->ev_copy_order, cudaEventDisableTiming));
    }
    return AERO_OK;
}
// The destination of a copy on the copy stream is a cached block that kernels already queued on the
// compute stream may still read or write: order the copy stream after the compute stream's tail.
static aero_status copy_stream_after_compute(aero_ctx *ctx) {
    TRY(ensure_copy_stream(ctx));
    CUDA_TRY(ctx, cudaEventRecord(ctx->ev_copy_order, ctx->stream));
    CUDA_TRY(ctx, cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_copy_order, 0));
    return AERO_OK;
}
static aero_status upload_enqueue(aero_upload *u) {
    aero_ctx *ctx = u->ctx;
    if (u->queued) return AERO_OK;
    TRY(copy_stream_after_compute(ctx));
    bool pinned = true;
    for (int c = u->col_begin; c < u->col_end; c++) pinned = pinned && host_is_pinned(u->cols[c]);
    if (!pinned) TRY(bulk_reserve(ctx, (size_t)16 << 20));
    for (int c = u->col_begin; c < u->col_end; c++)
        TRY(staged_h2d(ctx, u->d + (size_t)c * u->n_rows + u->row_begin, u->cols[c] + u->row_begin,
                       (u->row_end - u->row_begin) * 8, pinned, &ctx->bulk_slot));
    CUDA_TRY(ctx, cudaEventRecord(u->done, ctx->copy_stream));
    u->queued = true;
    return AERO_OK;
}
static aero_status flush_deferred_uploads(aero_ctx *ctx) {
    std::vector<aero_upload *> list;
    list.swap(ctx->deferred_uploads);
    for (aero_upload *u : list) TRY(upload_enqueue(u));
    return AERO_OK;
}
aero_status aero_upload_start(aero_ctx *ctx, const uint64_t *const *cols, uint32_t n_cols, uint64_t n_rows, int defer,
                              int shard_mode, aero_upload **out) {
    if (!ctx) return AERO_ERR_INVALID;
    enter(ctx);
    if (!cols || !out || n_cols == 0 || n_rows == 0) CTX_FAIL(ctx, AERO_ERR_INVALID, "null or empty matrix");
    for (uint32_t c = 0; c < n_cols; c++)
        if (!cols[c]) CTX_FAIL(ctx, AERO_ERR_INVALID, "null column %u", c);
    aero_upload *u = new aero_upload();
    u->ctx = ctx;
    u->n_rows = n_rows;
    u->cols.assign(cols, cols + n_cols);
    u->col_begin = 0;
    u->col_end = (int)n_cols;
    u->row_begin = 0;
    u->row_end = n_rows;
    if (shard_mode == AERO_UPLOAD_OWN_COLUMNS) own_columns(ctx, (int)n_cols, &u->col_begin, &u->col_end);
    else if (shard_mode == AERO_UPLOAD_OWN_ROWS) {
        if (n_rows % (uint64_t)ctx->shard_world) CTX_FAIL(ctx, AERO_ERR_INVALID, "rows not divisible by the number of ranks");
        u->row_begin = n_rows / ctx->shard_world * ctx->shard_rank;
        u->row_end = u->row_begin + n_rows / ctx->shard_world;
    } else if (shard_mode != AERO_UPLOAD_ALL) CTX_FAIL(ctx, AERO_ERR_INVALID, "unknown upload shard mode %d", shard_mode);
    aero_status st = dev_alloc(ctx, (void **)&u->d, (size_t)n_cols * n_rows * 8);
    if (st == AERO_OK && cudaEventCreateWithFlags(&u->done, cudaEventDisableTiming) != cudaSuccess) {
        ctx->err = "cudaEventCreate failed";
        st = AERO_ERR_CUDA;
    }
    if (st == AERO_OK) {
        if (defer) ctx->deferred_uploads.push_back(u);
        else st = upload_enqueue(u);
    }
    if (st != AERO_OK) {
        dev_free(ctx, u->d);
        if (u->done) cudaEventDestroy(u->done);
        delete u;
        return st;
    }
    *out = u;
    return AERO_OK;
}
aero_status aero_upload_wait(aero_upload *u, const uint64_t **d_cols) {
    if (!u || !d_cols) return AERO_ERR_INVALID;
    aero_ctx *ctx = u->ctx;
    enter(ctx);
    if (!u->queued) {  // still deferred: no commit came in between
        auto &v = ctx->deferred_uploads;
        v.erase(std::remove(v.begin(), v.end(), u), v.end());
        TRY(upload_enqueue(u));
    }
    CUDA_TRY(ctx, cudaStreamWaitEvent(ctx->stream, u->done, 0));
    *d_cols = u->d;
    return AERO_OK;
}
void aero_upload_free(aero_upload *u) {
    if (!u) return;
    aero_ctx *ctx = u->ctx;
    enter(ctx);
    auto &v = ctx->deferred_uploads;
    v.erase(std::remove(v.begin(), v.end(), u), v.end());
    if (u->queued) cudaEventSynchronize(u->done);  // the host columns are free again, and so is the block
    cudaEventDestroy(u->done);
    dev_free(ctx, u->d);  // later users of the block are ordered on ctx->stream, which waited for `done`
    delete u;
}

// ---- segments -------------------------------------------------------------------------------
aero_status aero_segment_commit_device(aero_ctx *ctx, const uint64_t *d_cols, size_t col_stride, uint32_t n_cols,
                                       uint64_t n_rows, uint32_t blowup, int input_is_coeffs, aero_segment **out,
                                       uint8_t root[32]) {
    if (!ctx) return AERO_ERR_INVALID;
    enter(ctx);
    if (!d_cols) CTX_FAIL(ctx, AERO_ERR_INVALID, "null matrix");
    if (col_stride < n_rows) CTX_FAIL(ctx, AERO_ERR_INVALID, "column stride smaller than the number of rows");
    return segment_from_device(ctx, d_cols, col_stride, 0, n_cols, n_rows, blowup, input_is_coeffs, out, root);
}

aero_status aero_segment_commit(aero_ctx *ctx, const uint64_t *const *cols, uint32_t n_cols, uint64_t n_rows,
                                uint32_t blowup, int input_is_coeffs, aero_segment **out, uint8_t root[32]) {
    if (!ctx) return AERO_ERR_INVALID;
    enter(ctx);
    if (!cols) CTX_FAIL(ctx, AERO_ERR_INVALID, "null matrix");
    if (n_cols == 0 || n_cols > 255) CTX_FAIL(ctx, AERO_ERR_INVALID, "number of columns must be 1..255, got %u", n_cols);
    if (n_rows < 2 || !is_pow2(n_rows)) CTX_FAIL(ctx, AERO_ERR_INVALID, "number of rows must be a power of two >= 2, got %llu", (unsigned long long)n_rows);
    for (uint32_t c = 0; c < n_cols; c++)
        if (!cols[c]) CTX_FAIL(ctx, AERO_ERR_INVALID, "null column %u", c);
    // a sharded context uploads (and interpolates) its own columns only; the coefficients of the others
    // arrive over NVLink
    int cb = 0, ce = (int)n_cols;
    own_columns(ctx, (int)n_cols, &cb, &ce);
    const int own = ce - cb;
    DevBlocks blocks(ctx);
    uint64_t *stage = nullptr;
    TRY(blocks.alloc((void **)&stage, (size_t)std::max(own, 1) * n_rows * 8));
    // Uploads run on a second stream in column batches; batch b's transforms wait only for batch b,
    // so the PCIe copy of later columns hides behind the NTTs of earlier ones.
    int batch = (int)std::max<uint32_t>(1, std::min<uint32_t>((uint32_t)ctx->upload_batch_cols, ((uint32_t)own + 2) / 3));
    if (batch > 1 && (batch & 1)) batch++;  // even batches: each one's row hash can start as soon as it is extended
    // short first / last batches once there are enough columns for it to matter (see upload_batch_size)
    int edge = (own >= 4 * batch && batch >= 4) ? ((batch / 2) & ~1) : 0;
    if (edge && ctx->upload_edge_cols >= 0) edge = std::min(ctx->upload_edge_cols & ~1, batch);
    std::vector<int> sizes;
    for (int c0 = 0; c0 < own;) {
        const int nc = std::min(upload_batch_size(c0, own, batch, edge), own - c0);
        sizes.push_back(nc);
        c0 += nc;
    }
    const int nb = (int)sizes.size();
    struct HostColumns : BatchSource {
        aero_ctx *ctx;
        const uint64_t *const *cols;
        uint64_t *stage;
        uint64_t n_rows;
        std::vector<int> sizes, first;
        bool pinned = true;
        int queued = 0;  // batches whose copies are on the copy stream
        aero_status queue(int b) {
            for (int k = 0; k < sizes[b]; k++) {
                const int c = first[b] + k;
                TRY(staged_h2d(ctx, stage + (size_t)c * n_rows, cols[c], n_rows * 8, pinned, &ctx->bulk_slot));
            }
            CUDA_TRY(ctx, cudaEventRecord(ev[b], ctx->copy_stream));
            if (b + 1 == (int)sizes.size()) TRY(flush_deferred_uploads(ctx));  // prefetches ride behind this segment's copies, under its NTTs
            return AERO_OK;
        }
        aero_status prepare(int b) override {
            for (; queued <= b; queued++) TRY(queue(queued));
            return AERO_OK;
        }
    } src;
    src.ctx = ctx;
    src.cols = cols + cb;
    src.stage = stage;
    src.n_rows = n_rows;
    src.sizes = sizes;
    for (int b = 0, c = 0; b < nb; c += sizes[b], b++) src.first.push_back(c);
    for (int c = 0; c < own; c++) src.pinned = src.pinned && host_is_pinned(cols[cb + c]);
    src.ev.assign((size_t)nb, nullptr);
    for (auto &e : src.ev) CUDA_TRY(ctx, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    // the staging block may still be read by kernels already queued on the compute stream
    TRY(copy_stream_after_compute(ctx));
    if (src.pinned) {
        // page-locked columns: every copy is queued now, so the copy engine never waits for the host
        if (nb) TRY(src.prepare(nb - 1));
        else TRY(flush_deferred_uploads(ctx));
    } else {
        // pageable columns (a Rust Vec): each batch is staged through pinned memory right before its
        // transforms are queued, the host copy of batch b + 1 running under the transfer and the NTTs of batch b
        TRY(bulk_reserve(ctx, std::max<size_t>((size_t)16 << 20, (size_t)batch * n_rows * 8)));
        if (!nb) TRY(flush_deferred_uploads(ctx));
    }
    aero_status st = segment_from_device(ctx, stage, n_rows, cb, n_cols, n_rows, blowup, input_is_coeffs, out, root, batch,
                                         nb ? &src : nullptr, edge);
    if (nb && src.queued == nb) cudaEventSynchronize(src.ev[nb - 1]);  // the caller's host buffers are free again on return
    return st;
}

aero_status aero_ctx_set_shard(aero_ctx *ctx, int rank, int world) {
    if (!ctx) return AERO_ERR_INVALID;
    if (world < 1 || world > AERO_MAX_RANKS || rank < 0 || rank >= world || (world & (world - 1)))
        CTX_FAIL(ctx, AERO_ERR_INVALID, "bad shard %d of %d (world must be a power of two <= %d)", rank, world, AERO_MAX_RANKS);
    if (ctx->win_live) CTX_FAIL(ctx, AERO_ERR_STATE, "window buffers of a sharded proof are live");
    if (ctx->win_ranks && world > 1 && (world != ctx->win_ranks || ctx->win_base[rank] != ctx->win))
        CTX_FAIL(ctx, AERO_ERR_STATE, "the attached exchange window serves rank layout %d, not %d of %d", ctx->win_ranks, rank, world);
    ctx->shard_rank = rank;
    ctx->shard_world = world;
    return AERO_OK;
}
aero_status aero_ctx_window_create(aero_ctx *ctx, size_t bytes, uint8_t handle_out[64]) {
    if (!ctx) return AERO_ERR_INVALID;
    enter(ctx);
    if (ctx->win) CTX_FAIL(ctx, AERO_ERR_STATE, "exchange window already created");
    if (bytes < 8192) CTX_FAIL(ctx, AERO_ERR_INVALID, "exchange window must be at least 8 KiB");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    void *p = nullptr;
    CUDA_TRY(ctx, cudaMalloc(&p, bytes));
    cudaError_t e = cudaMemset(p, 0, 4096);
    cudaIpcMemHandle_t h;
    if (e == cudaSuccess && handle_out) e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) {
        cudaFree(p);
        CTX_FAIL(ctx, AERO_ERR_CUDA, "exchange window: %s", cudaGetErrorString(e));
    }
    if (handle_out) memcpy(handle_out, &h, 64);
    ctx->win = (uint8_t *)p;
    ctx->win_bytes = bytes;
    ctx->win_off = 4096;
    ctx->win_live = 0;
    return AERO_OK;
}
static aero_status ensure_push_stream(aero_ctx *ctx) {
    if (!ctx->push_stream[0]) {
        for (cudaStream_t &ps : ctx->push_stream) CUDA_TRY(ctx, cudaStreamCreateWithFlags(&ps, cudaStreamNonBlocking));
        CUDA_TRY(ctx, cudaEventCreateWithFlags(&ctx->ev_push, cudaEventDisableTiming));
        preload_exchange_kernels();
    }
    return AERO_OK;
}
static aero_status window_attach_check(aero_ctx *ctx, int n_ranks) {
    if (!ctx->win) CTX_FAIL(ctx, AERO_ERR_STATE, "create the exchange window first");
    if (ctx->win_ranks) CTX_FAIL(ctx, AERO_ERR_STATE, "exchange window already attached");
    if (n_ranks != ctx->shard_world || n_ranks < 2 || n_ranks > AERO_MAX_RANKS)
        CTX_FAIL(ctx, AERO_ERR_INVALID, "window ranks (%d) must equal the shard world size (%d), 2..%d", n_ranks, ctx->shard_world, AERO_MAX_RANKS);
    return AERO_OK;
}
aero_status aero_ctx_window_attach(aero_ctx *ctx, int n_ranks, const uint8_t *handles) {
    if (!ctx || !handles) return AERO_ERR_INVALID;
    enter(ctx);
    TRY(window_attach_check(ctx, n_ranks));
    for (int r = 0; r < n_ranks; r++) {
        if (r == ctx->shard_rank) {
            ctx->win_base[r] = ctx->win;
            continue;
        }
        cudaIpcMemHandle_t h;
        memcpy(&h, handles + (size_t)r * 64, 64);
        void *p = nullptr;
        cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) {
            cudaGetLastError();
            for (int i = 0; i < r; i++)
                if (ctx->win_ipc[i]) cudaIpcCloseMemHandle(ctx->win_base[i]), ctx->win_ipc[i] = false;
            CTX_FAIL(ctx, AERO_ERR_CUDA, "cudaIpcOpenMemHandle(rank %d): %s", r, cudaGetErrorString(e));
        }
        ctx->win_base[r] = (uint8_t *)p;
        ctx->win_ipc[r] = true;
    }
    ctx->win_ranks = n_ranks;
    return ensure_push_stream(ctx);
}
aero_status aero_ctx_window_attach_local(aero_ctx *ctx, int n_ranks, aero_ctx *const *ranks) {
    if (!ctx || !ranks) return AERO_ERR_INVALID;
    enter(ctx);
    TRY(window_attach_check(ctx, n_ranks));
    for (int r = 0; r < n_ranks; r++) {
        aero_ctx *o = ranks[r];
        if (!o || !o->win || o->win_bytes != ctx->win_bytes) CTX_FAIL(ctx, AERO_ERR_INVALID, "rank %d has no exchange window of the same size", r);
        if ((r == ctx->shard_rank) != (o == ctx)) CTX_FAIL(ctx, AERO_ERR_INVALID, "ranks[%d] does not match this context's shard rank %d", r, ctx->shard_rank);
        if (o->device != ctx->device) {
            int can = 0;
            CUDA_TRY(ctx, cudaDeviceCanAccessPeer(&can, ctx->device, o->device));
            if (!can) CTX_FAIL(ctx, AERO_ERR_UNSUPPORTED, "device %d cannot access device %d", ctx->device, o->device);
            cudaError_t e = cudaDeviceEnablePeerAccess(o->device, 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) CTX_FAIL(ctx, AERO_ERR_CUDA, "cudaDeviceEnablePeerAccess(%d): %s", o->device, cudaGetErrorString(e));
            cudaGetLastError();
        }
        ctx->win_base[r] = o->win;
    }
    ctx->win_ranks = n_ranks;
    return ensure_push_stream(ctx);
}
int aero_ctx_window_ranks(aero_ctx *ctx) { return (ctx && ctx->shard_world > 1) ? ctx->win_ranks : 0; }
aero_status aero_ctx_set_host_barrier(aero_ctx *ctx, aero_host_barrier_fn fn, void *user) {
    if (!ctx) return AERO_ERR_INVALID;
    ctx->host_barrier = fn;
    ctx->host_barrier_user = user;
    return AERO_OK;
}
aero_status aero_window_barrier(aero_ctx *ctx) {
    if (!ctx) return AERO_ERR_INVALID;
    enter(ctx);
    return window_barrier(ctx);
}
// A sharded proof of a shape this context has proved before makes no driver allocation (every block
// comes from the cache), so its barriers can be device-side; the first proof of a shape synchronises
// on the host (see aero_ctx::win_host_sync).
aero_status aero_ctx_shard_begin(aero_ctx *ctx, const char *shape_key) {
    if (!ctx || !shape_key) return AERO_ERR_INVALID;
    enter(ctx);
    if (!ctx_sharded(ctx)) return AERO_OK;
    if (ctx->win_ranks != ctx->shard_world) CTX_FAIL(ctx, AERO_ERR_STATE, "sharded context (%d ranks) has no exchange window attached", ctx->shard_world);
    const bool warm = ctx->warm_shapes.count(shape_key) != 0;
    ctx->win_host_sync = !warm || ctx->force_host_sync;
    ctx->cur_shape = shape_key;
    return AERO_OK;
}
aero_status aero_ctx_shard_end(aero_ctx *ctx, int ok) {
    if (!ctx) return AERO_ERR_INVALID;
    if (!ctx_sharded(ctx)) return AERO_OK;
    if (ok) ctx->warm_shapes.insert(ctx->cur_shape);
    else ctx->warm_shapes.erase(ctx->cur_shape);
    ctx->win_host_sync = true;
    return AERO_OK;
}

void aero_segment_destroy(aero_segment *seg) {
    if (!seg) return;
    enter(seg->ctx);
    dev_free(seg->ctx, seg->polys);
    dev_free(seg->ctx, seg->lde);
    dev_free(seg->ctx, seg->heap);
    dev_free(seg->ctx, seg->leaf_stage);
    dev_free(seg->ctx, seg->top);
    delete seg;
}
aero_status aero_segment_info(aero_segment *seg, uint32_t *n_cols, uint64_t *n_rows, uint32_t *blowup) {
    if (!seg) return AERO_ERR_INVALID;
    if (n_cols) *n_cols = (uint32_t)seg->ncols;
    if (n_rows) *n_rows = seg->n();
    if (blowup) *blowup = seg->log_blowup < 0 ? 0 : (1u << seg->log_blowup);
    return AERO_OK;
}
aero_status aero_segments_roots(aero_ctx *ctx, aero_segment *const *segs, uint32_t n_segs, uint8_t *roots_out) {
    if (!ctx) return AERO_ERR_INVALID;
    enter(ctx);
    if (!segs || !roots_out || !n_segs) CTX_FAIL(ctx, AERO_ERR_INVALID, "null argument");
    void *h = nullptr;
    TRY(ring_take(ctx, (size_t)n_segs * 32, &h));
    for (uint32_t i = 0; i < n_segs; i++) {
        if (!segs[i] || segs[i]->ctx != ctx || !segs[i]->top) CTX_FAIL(ctx, AERO_ERR_STATE, "segment %u has no commitment on this context", i);
        CUDA_TRY(ctx, cudaMemcpyAsync((uint8_t *)h + (size_t)i * 32, segs[i]->top + 8, 32, cudaMemcpyDeviceToHost, ctx->stream));
    }
    const unsigned int *word = nullptr;
    TRY(window_check_queue(ctx, &word));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    TRY(window_check_result(ctx, word));
    memcpy(roots_out, h, (size_t)n_segs * 32);
    return AERO_OK;
}
aero_status aero_segment_commit_polys(aero_segment *seg, uint32_t blowup, uint8_t root[32]) {
    if (!seg) return AERO_ERR_INVALID;
    enter(seg->ctx);
    if (!is_pow2(blowup) || blowup < 2) CTX_FAIL(seg->ctx, AERO_ERR_INVALID, "blowup factor must be a power of two >= 2, got %u", blowup);
    return segment_extend_commit(seg, ilog2(blowup), root);
}

// The extended trace for the host-side AIR evaluator: 8 * N bytes per column, 5.4 GB for a 2^20-row Miden
// trace -- the transfer that dominates a real proof once the rest of the path is on the GPU.
// Columns are converted to natural order two at a time into alternating device buffers; page-locked
// destinations receive them directly; pageable ones (a Rust Vec) through two pinned slots, the host copy of
// column c (split over threads) running under the PCIe transfer of column c + 1.
aero_status aero_segment_download_lde(aero_segment *seg, uint64_t *const *cols_out) {
    if (!seg || !cols_out) return AERO_ERR_INVALID;
    aero_ctx *ctx = seg->ctx;
    enter(ctx);
    if (!seg->lde) CTX_FAIL(ctx, AERO_ERR_STATE, "segment has no LDE");
    if (seg->coset_count != (1 << seg->log_blowup)) CTX_FAIL(ctx, AERO_ERR_UNSUPPORTED, "download_lde on a coset-sharded segment");
    const uint64_t N = seg->N();
    const size_t col_bytes = N * 8;
    for (int c = 0; c < seg->ncols; c++)
        if (!cols_out[c]) CTX_FAIL(ctx, AERO_ERR_INVALID, "null column %d", c);
    PhaseTimer t(ctx, "download_lde");
    DevBlocks blk(ctx);
    uint64_t *tmp[2] = {nullptr, nullptr};
    TRY(blk.alloc((void **)&tmp[0], col_bytes));
    TRY(blk.alloc((void **)&tmp[1], col_bytes));
    bool all_pinned = true;
    for (int c = 0; c < seg->ncols; c++) all_pinned = all_pinned && host_is_pinned(cols_out[c]);
    if (!all_pinned) TRY(bulk_reserve(ctx, col_bytes));
    TRY(ensure_copy_stream(ctx));
    // conversion kernels on the compute stream, transfers on the copy stream: ev_conv[s] = slot s converted,
    // ev_bulk[s] = slot s transferred (device buffer and pinned slot free again)
    cudaEvent_t ev_conv[2];
    for (auto &e : ev_conv) CUDA_TRY(ctx, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    if (!ctx->ev_bulk[0])
        for (int i = 0; i < 2; i++) CUDA_TRY(ctx, cudaEventCreateWithFlags(&ctx->ev_bulk[i], cudaEventDisableTiming));
    aero_status st = AERO_OK;
    const bool mont = ctx->form == AERO_FORM_MONTGOMERY;
    for (int c = 0; c <= seg->ncols && st == AERO_OK; c++) {
        const int s = c & 1;
        if (c < seg->ncols) {
            if (c >= 2) {  // slot s last served column c - 2: its transfer must be over, and (pageable) copied out
                if (cudaStreamWaitEvent(ctx->stream, ctx->ev_bulk[s], 0) != cudaSuccess) st = AERO_ERR_CUDA;
            }
            lde_to_natural(seg->lde + (size_t)c * N, tmp[s], seg->logn, seg->log_blowup, mont, ctx->stream);
            cudaEventRecord(ev_conv[s], ctx->stream);
            cudaStreamWaitEvent(ctx->copy_stream, ev_conv[s], 0);
        }
        if (c < seg->ncols) {  // (pinned slot s is free: column c - 2 was copied out of it in the previous iteration)
            void *dst = all_pinned ? (void *)cols_out[c] : (void *)ctx->h_bulk[s];
            if (cudaMemcpyAsync(dst, tmp[s], col_bytes, cudaMemcpyDeviceToHost, ctx->copy_stream) != cudaSuccess) st = AERO_ERR_CUDA;
            cudaEventRecord(ctx->ev_bulk[s], ctx->copy_stream);
        }
        if (!all_pinned && c >= 1) {  // column c - 1 has arrived in the other slot: copy it out under column c's transfer
            const int p = (c - 1) & 1;
            if (cudaEventSynchronize(ctx->ev_bulk[p]) != cudaSuccess) st = AERO_ERR_CUDA;
            parallel_memcpy(cols_out[c - 1], ctx->h_bulk[p], col_bytes);
        }
    }
    if (cudaStreamSynchronize(ctx->copy_stream) != cudaSuccess || cudaStreamSynchronize(ctx->stream) != cudaSuccess) st = AERO_ERR_CUDA;
    for (auto &e : ev_conv) cudaEventDestroy(e);
    if (st != AERO_OK) CTX_FAIL(ctx, st, "download_lde: %s", cudaGetErrorString(cudaGetLastError()));
    return AERO_OK;
}
aero_status aero_segment_download_polys(aero_segment *seg, uint64_t *const *cols_out) {
    if (!seg || !cols_out) return AERO_ERR_INVALID;
    aero_ctx *ctx = seg->ctx;
    enter(ctx);
    const uint64_t n = seg->n();
    DevBlocks blk(ctx);
    uint64_t *tmp = nullptr;
    const bool mont = ctx->form == AERO_FORM_MONTGOMERY;
    if (mont) TRY(blk.alloc((void **)&tmp, n * 8));
    for (int c = 0; c < seg->ncols; c++) {
        const uint64_t *src = seg->polys + (size_t)c * n;
        if (mont) {
            convert_form(src, tmp, n, 1, ctx->stream);
            src = tmp;
        }
        CUDA_TRY(ctx, cudaMemcpyAsync(cols_out[c], src, n * 8, cudaMemcpyDeviceToHost, ctx->stream));
    }
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return AERO_OK;
}
aero_status aero_segment_download_leaves(aero_segment *seg, uint8_t *leaves_out) {
    if (!seg || !leaves_out) return AERO_ERR_INVALID;
    aero_ctx *ctx = seg->ctx;
    enter(ctx);
    if (!seg->heap) CTX_FAIL(ctx, AERO_ERR_STATE, "segment has no commitment");
    if (seg->logG) CTX_FAIL(ctx, AERO_ERR_UNSUPPORTED, "download_leaves on a sharded segment");
    const uint64_t N = seg->N();
    DevBlocks tmp(ctx);
    uint32_t *nat = nullptr;
    TRY(tmp.alloc((void **)&nat, N * 32));
    leaves_to_natural(seg->leaf_stage, nat, seg->logn, seg->log_blowup, ctx->stream);
    CUDA_TRY(ctx, cudaMemcpyAsync(leaves_out, nat, N * 32, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return AERO_OK;
}

aero_status aero_segment_open(aero_segment *seg, const uint64_t *positions, uint32_t n_pos, uint64_t *rows_out,
                              uint8_t *batch_nodes_out, size_t *len) {
    if (!seg || !positions || !len) return AERO_ERR_INVALID;
    enter(seg->ctx);
    GatherBatch gb(seg->ctx);
    SegmentOpening o;
    TRY(segment_open_plan(seg, positions, n_pos, rows_out != nullptr, gb, o));
    TRY(gb.run());
    return segment_open_finish(o, gb, rows_out, batch_nodes_out, len);
}

aero_status aero_open_queries(aero_ctx *ctx, aero_fri *fri, aero_segment *const *segs, uint32_t n_segs,
                              const uint64_t *positions, uint32_t n_pos, uint8_t *fri_proof_bytes, size_t *fri_len,
                              uint64_t *const *rows_out, uint8_t *const *batch_nodes_out, size_t *batch_len) {
    if (!ctx) return AERO_ERR_INVALID;
    enter(ctx);
    if (!positions || (fri && !fri_len) || (n_segs && (!segs || !rows_out || !batch_nodes_out || !batch_len)))
        CTX_FAIL(ctx, AERO_ERR_INVALID, "null argument");
    static const bool timing = getenv("AERO_HOST_TIMING") != nullptr;  // diagnostic: host time of the planning step
    const auto t_begin = std::chrono::steady_clock::now();
    GatherBatch gb(ctx);
    FriOpening fo;
    std::vector<SegmentOpening> so(n_segs);
    if (fri) {
        if (fri->ctx != ctx) CTX_FAIL(ctx, AERO_ERR_INVALID, "FRI handle belongs to another context");
        TRY(fri_open_plan(fri, positions, n_pos, gb, fo));
    }
    for (uint32_t i = 0; i < n_segs; i++) {
        if (!segs[i] || segs[i]->ctx != ctx) CTX_FAIL(ctx, AERO_ERR_INVALID, "segment %u is null or belongs to another context", i);
        TRY(segment_open_plan(segs[i], positions, n_pos, true, gb, so[i]));
    }
    const auto t_plan = std::chrono::steady_clock::now();
    TRY(gb.run());
    if (timing)
        fprintf(stderr, "aero_open_queries: plan %.1f us, gather batch (upload, %zu launches, download, sync) %.1f us\n",
                std::chrono::duration<double, std::micro>(t_plan - t_begin).count(), gb.jobs.size(),
                std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t_plan).count());
    aero_status worst = AERO_OK;  // report every required size before failing with AERO_ERR_BUFFER
    if (fri) {
        aero_status st = fri_open_finish(fo, gb, fri_proof_bytes, fri_len);
        if (st != AERO_OK && st != AERO_ERR_BUFFER) return st;
        if (st != AERO_OK) worst = st;
    }
    for (uint32_t i = 0; i < n_segs; i++) {
        aero_status st = segment_open_finish(so[i], gb, rows_out[i], batch_nodes_out[i], &batch_len[i]);
        if (st != AERO_OK && st != AERO_ERR_BUFFER) return st;
        if (st != AERO_OK) worst = st;
    }
    return worst;
}

aero_status aero_commit_rows_device(aero_ctx *ctx, const uint64_t *d_m, size_t col_stride, uint32_t n_cols,
                                    uint64_t n_rows, uint8_t root[32]) {
    if (!ctx) return AERO_ERR_INVALID;
    enter(ctx);
    if (!d_m || !root) CTX_FAIL(ctx, AERO_ERR_INVALID, "null argument");
    if (n_cols == 0) CTX_FAIL(ctx, AERO_ERR_INVALID, "matrix must have at least one column");
    // MerkleTree::new (crypto/src/merkle/mod.rs:108-114): >= 2 leaves, power of two
    if (n_rows < 2) CTX_FAIL(ctx, AERO_ERR_INVALID, "a Merkle tree needs at least two leaves, got %llu", (unsigned long long)n_rows);
    if (!is_pow2(n_rows)) CTX_FAIL(ctx, AERO_ERR_INVALID, "number of leaves must be a power of two, got %llu", (unsigned long long)n_rows);
    if (n_rows > (1ULL << 31)) CTX_FAIL(ctx, AERO_ERR_UNSUPPORTED, "too many rows");
    uint32_t *full = nullptr;
    TRY(dev_alloc(ctx, (void **)&full, (size_t)2 * n_rows * 32));
    {
        PhaseTimer t(ctx, "hash_rows");
        hash_rows_natural(d_m, col_stride, (int)n_cols, (uint32_t)n_rows, full + (size_t)n_rows * 8, ctx->stream);
    }
    {
        PhaseTimer t(ctx, "merkle");
        merkle_build(full, n_rows, ctx->stream);
    }
    aero_status st = download_small(ctx, root, full + 8, 32);
    dev_free(ctx, full);
    if (st != AERO_OK) return st;
    CUDA_TRY(ctx, cudaGetLastError());
    return AERO_OK;
}

// ---- constraints ----------------------------------------------------------------------------
// ---- AIR constraint evaluation on the device ---------------------------------------------------
// Host-side radix-2 transform of a small vector in place (canonical elements): out[j] = sum_k a[k] w^(jk).
static void host_ntt(std::vector<uint64_t> &a, uint64_t w) {
    const size_t n = a.size();
    for (size_t i = 1, j = 0; i < n; i++) {  // bit reversal
        size_t bit = n >> 1;
        for (; j & bit; bit >>= 1) j ^= bit;
        j ^= bit;
        if (i < j) std::swap(a[i], a[j]);
    }
    for (size_t len = 2; len <= n; len <<= 1) {
        const uint64_t wl = gl::pow(w, n / len);
        for (size_t i = 0; i < n; i += len) {
            uint64_t t = 1;
            for (size_t k = 0; k < len / 2; k++) {
                const uint64_t u = a[i + k], v = gl::mul(a[i + k + len / 2], t);
                a[i + k] = gl::add(u, v);
                a[i + k + len / 2] = gl::sub(u, v);
                t = gl::mul(t, wl);
            }
        }
    }
}
// One column of PeriodicValueTable (prover/src/constraints/periodic_table.rs:41-56): the cycle values are
// interpolated over the subgroup of order c (Air::get_periodic_column_polys, air/src/air/mod.rs:336-341) and the
// polynomial is evaluated over offset^num_cycles * <w_(c * ce_blowup)> in natural order
// (fft::evaluate_poly_with_offset); appended to `table`.
static void periodic_column_table(std::vector<uint64_t> col /* c values */, uint64_t num_cycles, uint32_t ce_blowup,
                                  std::vector<uint64_t> &table) {
    const size_t c = col.size();
    const int logc = ilog2(c);
    host_ntt(col, gl::inv(gl::root_of_unity(logc)));
    const uint64_t cinv = gl::inv((uint64_t)c), offset = gl::pow(gl::GENERATOR, num_cycles);
    std::vector<uint64_t> ext(c * ce_blowup, 0);
    uint64_t s = cinv;  // coefficient j scaled by offset^j / c
    for (size_t j = 0; j < c; j++) {
        ext[j] = gl::mul(col[j], s);
        s = gl::mul(s, offset);
    }
    host_ntt(ext, gl::root_of_unity(logc + ilog2(ce_blowup)));
    table.insert(table.end(), ext.begin(), ext.end());
}
aero_status aero_periodic_column_table(const uint64_t *cycle_values, uint64_t cycle_len, uint64_t trace_len,
                                       uint32_t ce_blowup, uint64_t *out) {
    if (!cycle_values || !out) return AERO_ERR_INVALID;
    if (cycle_len < 2 || !is_pow2(cycle_len) || !is_pow2(trace_len) || cycle_len > trace_len) return AERO_ERR_INVALID;
    if (!is_pow2(ce_blowup) || ilog2(cycle_len) + ilog2(ce_blowup) > 32) return AERO_ERR_INVALID;
    std::vector<uint64_t> col(cycle_values, cycle_values + cycle_len), table;
    for (uint64_t v : col)
        if (v >= gl::P) return AERO_ERR_INVALID;
    periodic_column_table(col, trace_len / cycle_len, ce_blowup, table);
    std::copy(table.begin(), table.end(), out);
    return AERO_OK;
}

// tau0 / tau_count out: the coset-major range of evaluation-domain steps this rank evaluated (everything on one GPU)
static aero_status constraints_evaluate_impl(aero_ctx *ctx, aero_segment *const *trace_segs, uint32_t n_trace_segs,
                                             const aero_air_program *prog, const uint64_t *coeffs, uint32_t n_coeffs,
                                             uint32_t ce_blowup, uint32_t n_div, uint64_t *d_eval_cols, size_t col_stride,
                                             uint32_t *tau0_out, uint32_t *tau_count_out) {
    if (!ctx) return AERO_ERR_INVALID;
    enter(ctx);
    if (!trace_segs || !prog || !d_eval_cols || (n_coeffs && !coeffs)) CTX_FAIL(ctx, AERO_ERR_INVALID, "null argument");
    if (n_trace_segs == 0 || n_trace_segs > 4) CTX_FAIL(ctx, AERO_ERR_INVALID, "1..4 trace segments");
    if (n_div == 0 || n_div > 8) CTX_FAIL(ctx, AERO_ERR_INVALID, "number of divisors must be 1..8, got %u", n_div);
    if (prog->n_nodes == 0 || prog->n_nodes > (1u << 16) || !prog->nodes) CTX_FAIL(ctx, AERO_ERR_UNSUPPORTED, "transition program must have 1..65536 nodes");
    if (n_coeffs != 2 * (prog->n_transition + prog->n_boundary)) CTX_FAIL(ctx, AERO_ERR_INVALID, "expected a coefficient pair per constraint (%u), got %u elements", prog->n_transition + prog->n_boundary, n_coeffs);
    if ((prog->n_transition && (!prog->transition_out || !prog->transition_adj)) ||
        (prog->n_boundary && (!prog->boundary_col || !prog->boundary_value || !prog->boundary_adj || !prog->boundary_div)) ||
        (prog->n_consts && !prog->consts) || (prog->n_periodic && (!prog->periodic_len || !prog->periodic_values)))
        CTX_FAIL(ctx, AERO_ERR_INVALID, "null program array");
    AirSegs segs{};
    segs.nseg = (int)n_trace_segs;
    uint32_t width = 0;
    const aero_segment *s0 = trace_segs[0];
    for (uint32_t k = 0; k < n_trace_segs; k++) {
        const aero_segment *sg = trace_segs[k];
        if (!sg || !sg->lde) CTX_FAIL(ctx, AERO_ERR_STATE, "trace segment %u has not been committed", k);
        if (sg->logn != s0->logn || sg->log_blowup != s0->log_blowup || sg->coset_begin != s0->coset_begin || sg->coset_count != s0->coset_count)
            CTX_FAIL(ctx, AERO_ERR_INVALID, "trace segments of different shape");
        // a coset-sharded segment stores cosets [coset_begin, coset_begin + coset_count) compactly: rebase so that
        // the kernel's (LDE coset, i) index lands in it (only the rank's own cosets are ever read)
        segs.lde[k] = sg->lde - ((size_t)sg->coset_begin << sg->logn);
        segs.stride[k] = sg->lde_stride();
        segs.ncols[k] = sg->ncols;
        width += (uint32_t)sg->ncols;
    }
    const int logn = s0->logn, log_blowup = s0->log_blowup;
    if (!is_pow2(ce_blowup) || ce_blowup < 2 || ce_blowup > (1u << log_blowup)) CTX_FAIL(ctx, AERO_ERR_INVALID, "constraint evaluation blowup must be a power of two in 2..blowup");
    const int log_ce = ilog2(ce_blowup);
    const uint64_t CE = (uint64_t)ce_blowup << logn;
    if (col_stride < CE) CTX_FAIL(ctx, AERO_ERR_INVALID, "column stride smaller than the constraint evaluation domain");
    // validate the program: operands refer to earlier nodes, columns and constants exist
    for (uint32_t k = 0; k < prog->n_nodes; k++) {
        const aero_air_node &nd = prog->nodes[k];
        bool ok;
        switch (nd.op) {
        case AERO_AIR_CUR: case AERO_AIR_NEXT: ok = nd.a < width; break;
        case AERO_AIR_CONST: ok = nd.a < prog->n_consts; break;
        case AERO_AIR_PERIODIC: ok = nd.a < prog->n_periodic; break;
        case AERO_AIR_ADD: case AERO_AIR_SUB: case AERO_AIR_MUL: ok = nd.a < k && nd.b < k; break;
        default: ok = false;
        }
        if (!ok) CTX_FAIL(ctx, AERO_ERR_INVALID, "transition program: bad node %u (op %u, operands %u, %u)", k, nd.op, nd.a, nd.b);
    }
    auto is_arith = [](uint32_t op) { return op >= AERO_AIR_ADD && op <= AERO_AIR_MUL; };
    // PeriodicValueTable::new (periodic_table.rs:25-75), one column after another
    std::vector<uint64_t> periodic;
    std::vector<uint32_t> per_off(prog->n_periodic), per_mask(prog->n_periodic);
    {
        size_t src = 0, total = 0;
        for (uint32_t k = 0; k < prog->n_periodic; k++) {
            const uint64_t c = prog->periodic_len[k];
            // Air::get_periodic_column_polys' assertions (air/src/air/mod.rs:319-335)
            if (c < 2 || !is_pow2(c) || c > ((uint64_t)1 << logn)) CTX_FAIL(ctx, AERO_ERR_INVALID, "periodic column %u: cycle length %llu is not a power of two in 2..trace length", k, (unsigned long long)c);
            total += (size_t)c * ce_blowup;
        }
        if (total > ((size_t)1 << 28)) CTX_FAIL(ctx, AERO_ERR_UNSUPPORTED, "periodic value table too large (%zu values)", total);
        for (uint32_t k = 0; k < prog->n_periodic; k++) {
            const uint64_t c = prog->periodic_len[k];
            std::vector<uint64_t> col(c);
            for (uint64_t i = 0; i < c; i++) col[i] = to_canon(ctx, prog->periodic_values[src + i]);
            src += c;
            per_off[k] = (uint32_t)periodic.size();
            per_mask[k] = (uint32_t)(c * ce_blowup - 1);
            periodic_column_table(col, ((uint64_t)1 << logn) / c, ce_blowup, periodic);
        }
    }
    std::vector<uint64_t> adj;  // distinct degree adjustments
    auto adj_index = [&](uint64_t a) {
        for (size_t i = 0; i < adj.size(); i++)
            if (adj[i] == a) return (uint32_t)i;
        adj.push_back(a);
        return (uint32_t)(adj.size() - 1);
    };
    // one upload: [u64 section: consts, b_val, coeffs, adj][u32 section: nodes, t_out, t_adj, b_col, b_adj, b_div]
    const uint32_t nt = prog->n_transition, nb = prog->n_boundary;
    std::vector<uint32_t> t_adj(nt), b_adj(nb);
    for (uint32_t t = 0; t < nt; t++) {
        if (prog->transition_out[t] >= prog->n_nodes) CTX_FAIL(ctx, AERO_ERR_INVALID, "transition constraint %u: no such node", t);
        t_adj[t] = adj_index(prog->transition_adj[t]);
    }
    for (uint32_t j = 0; j < nb; j++) {
        if (prog->boundary_col[j] >= width) CTX_FAIL(ctx, AERO_ERR_INVALID, "boundary constraint %u: no such column", j);
        if (prog->boundary_div[j] == 0 || prog->boundary_div[j] >= n_div) CTX_FAIL(ctx, AERO_ERR_INVALID, "boundary constraint %u: divisor column out of range", j);
        b_adj[j] = adj_index(prog->boundary_adj[j]);
    }
    if (adj.size() > (size_t)AIR_MAX_ADJ) CTX_FAIL(ctx, AERO_ERR_UNSUPPORTED, "more than %d distinct degree adjustments", AIR_MAX_ADJ);
    std::vector<uint64_t> w64;
    const size_t o_consts = 0, o_bval = o_consts + prog->n_consts, o_coeffs = o_bval + nb, o_adj = o_coeffs + n_coeffs,
                 o_adjoff = o_adj + adj.size(), o_per = o_adjoff + adj.size();
    w64.resize(o_per + periodic.size());
    std::copy(periodic.begin(), periodic.end(), w64.begin() + o_per);
    for (uint32_t i = 0; i < prog->n_consts; i++) w64[o_consts + i] = to_canon(ctx, prog->consts[i]);
    for (uint32_t j = 0; j < nb; j++) w64[o_bval + j] = to_canon(ctx, prog->boundary_value[j]);
    for (uint32_t i = 0; i < n_coeffs; i++) w64[o_coeffs + i] = to_canon(ctx, coeffs[i]);
    for (size_t i = 0; i < adj.size(); i++) {
        if (adj[i] >> 32) CTX_FAIL(ctx, AERO_ERR_UNSUPPORTED, "degree adjustment %llu does not fit 32 bits", (unsigned long long)adj[i]);
        w64[o_adj + i] = adj[i];
        w64[o_adjoff + i] = gl::pow(gl::GENERATOR, adj[i]);  // domain offset^adjustment (domain.rs:109-117)
    }
    // Slots by liveness: a node's value occupies a slot from its evaluation to its last consumer; operands that
    // die at node k free their slots before k's result is placed (the kernel reads both operands, then writes).
    const uint32_t NN = prog->n_nodes;
    std::vector<uint32_t> last(NN, 0), slot(NN, 0);
    std::vector<bool> used(NN, false);
    for (uint32_t k = 0; k < NN; k++)
        if (is_arith(prog->nodes[k].op)) {
            last[prog->nodes[k].a] = k;
            last[prog->nodes[k].b] = k;
            used[prog->nodes[k].a] = used[prog->nodes[k].b] = true;
        }
    for (uint32_t t = 0; t < nt; t++) {
        last[prog->transition_out[t]] = 0xFFFFFFFFu;  // constraint values live to the end
        used[prog->transition_out[t]] = true;
    }
    std::vector<uint32_t> free_slots;
    uint32_t n_slots = 0;
    for (uint32_t k = 0; k < NN; k++) {
        const aero_air_node &nd = prog->nodes[k];
        if (is_arith(nd.op)) {
            if (last[nd.a] == k) free_slots.push_back(slot[nd.a]);
            if (nd.b != nd.a && last[nd.b] == k) free_slots.push_back(slot[nd.b]);
        }
        if (free_slots.empty()) slot[k] = n_slots++;
        else {
            slot[k] = free_slots.back();
            free_slots.pop_back();
        }
        if (!used[k]) free_slots.push_back(slot[k]);  // a value nobody reads
    }
    std::vector<uint32_t> w32;
    const size_t o_nodes = 0, o_tout = o_nodes + 4 * (size_t)NN, o_tadj = o_tout + nt, o_bcol = o_tadj + nt,
                 o_badj = o_bcol + nb, o_bdiv = o_badj + nb;
    w32.resize(o_bdiv + nb);
    for (uint32_t k = 0; k < NN; k++) {
        const aero_air_node &nd = prog->nodes[k];
        const bool arith = is_arith(nd.op), per = nd.op == AERO_AIR_PERIODIC;
        w32[o_nodes + 4 * k] = nd.op;
        w32[o_nodes + 4 * k + 1] = arith ? slot[nd.a] : per ? per_off[nd.a] : nd.a;
        w32[o_nodes + 4 * k + 2] = arith ? slot[nd.b] : per ? per_mask[nd.a] : 0;
        w32[o_nodes + 4 * k + 3] = slot[k];
    }
    for (uint32_t t = 0; t < nt; t++) {
        w32[o_tout + t] = slot[prog->transition_out[t]];
        w32[o_tadj + t] = t_adj[t];
    }
    for (uint32_t j = 0; j < nb; j++) {
        w32[o_bcol + j] = prog->boundary_col[j];
        w32[o_badj + j] = b_adj[j];
        w32[o_bdiv + j] = prog->boundary_div[j];
    }
    if (n_slots > 1024) CTX_FAIL(ctx, AERO_ERR_UNSUPPORTED, "transition program keeps %u values alive at once (max 1024)", n_slots);
    DevBlocks blk(ctx);
    uint64_t *d64 = nullptr;
    uint32_t *d32 = nullptr;
    TRY(blk.alloc((void **)&d64, std::max<size_t>(8, w64.size() * 8)));
    TRY(blk.alloc((void **)&d32, w32.size() * 4));
    if (!w64.empty()) TRY(upload_small(ctx, d64, w64.data(), w64.size() * 8));
    TRY(upload_small(ctx, d32, w32.data(), w32.size() * 4));
    AirProgramDev p{};
    p.nodes = d32 + o_nodes;
    p.t_out = d32 + o_tout;
    p.t_adj = d32 + o_tadj;
    p.b_col = d32 + o_bcol;
    p.b_adj = d32 + o_badj;
    p.b_div = d32 + o_bdiv;
    p.consts = d64 + o_consts;
    p.b_val = d64 + o_bval;
    p.coeffs = d64 + o_coeffs;
    p.adj = d64 + o_adj;
    p.adj_off = d64 + o_adjoff;
    p.periodic = d64 + o_per;
    p.n_nodes = (int)prog->n_nodes;
    p.n_slots = (int)std::max<uint32_t>(1, n_slots);
    p.nt = (int)nt;
    p.nb = (int)nb;
    p.n_adj = (int)adj.size();
    p.n_div = (int)n_div;
    PowTable g_ce;  // g_ce^e: the constraint evaluation domain without its offset (StarkDomain::ce_domain, domain.rs:53-60)
    {
        char key[32];
        snprintf(key, sizeof key, "g/%d", logn + log_ce);
        TRY(get_pow_table(ctx, key, gl::root_of_unity(logn + log_ce), logn + log_ce, 1, &g_ce));
    }
    // The evaluation-domain cosets whose LDE coset (rc << shift) this rank holds: a contiguous range, all of them
    // on one GPU.  Step s = i * ce_blowup + rc reads LDE coset rc << shift at i and i + 1 only, so a coset-sharded
    // rank evaluates its share of the domain without any exchange.
    const int shift = log_blowup - log_ce;
    const uint32_t rc_lo = (uint32_t)((s0->coset_begin + (1 << shift) - 1) >> shift);
    const uint32_t rc_hi = (uint32_t)((s0->coset_begin + s0->coset_count + (1 << shift) - 1) >> shift);
    const uint32_t tau0 = rc_lo << logn, tau_count = (rc_hi - rc_lo) << logn;
    {
        PhaseTimer t(ctx, "constraint_evaluate");
        air_evaluate(segs, p, logn, log_blowup, log_ce, g_ce, ctx->form == AERO_FORM_MONTGOMERY, d_eval_cols, col_stride, ctx->stream,
                     tau0, tau_count, ctx->num_sms, ctx->air_blocks_per_sm);
    }
    if (tau0_out) *tau0_out = tau0;
    if (tau_count_out) *tau_count_out = tau_count;
    CUDA_TRY(ctx, cudaGetLastError());
    return AERO_OK;
}

aero_status aero_constraints_evaluate_device(aero_ctx *ctx, aero_segment *const *trace_segs, uint32_t n_trace_segs,
                                             const aero_air_program *prog, const uint64_t *coeffs, uint32_t n_coeffs,
                                             uint32_t ce_blowup, uint32_t n_div, uint64_t *d_eval_cols, size_t col_stride) {
    return constraints_evaluate_impl(ctx, trace_segs, n_trace_segs, prog, coeffs, n_coeffs, ce_blowup, n_div, d_eval_cols,
                                     col_stride, nullptr, nullptr);
}

// own_cosets: a sharded context holds (and combines) the coset-major range [tau0, tau0 + tau_count) of the evaluation
// columns -- what constraints_evaluate_impl left there -- instead of its block of rows.
static aero_status constraints_into_poly_impl(aero_ctx *ctx, const uint64_t *d_eval_cols, size_t col_stride,
                                              const aero_divisor *divs, uint32_t n_div, uint64_t ce_domain_size,
                                              uint64_t trace_len, aero_segment **out, bool own_cosets, uint32_t tau0,
                                              uint32_t tau_count) {
    if (!ctx) return AERO_ERR_INVALID;
    enter(ctx);
    if (!d_eval_cols || !divs || !out) CTX_FAIL(ctx, AERO_ERR_INVALID, "null argument");
    if (col_stride < ce_domain_size) CTX_FAIL(ctx, AERO_ERR_INVALID, "column stride smaller than the domain");
    if (n_div == 0 || n_div > 8) CTX_FAIL(ctx, AERO_ERR_INVALID, "number of divisors must be 1..8, got %u", n_div);
    if (!is_pow2(ce_domain_size) || !is_pow2(trace_len)) CTX_FAIL(ctx, AERO_ERR_INVALID, "domain sizes must be powers of two");
    // CompositionPoly::new (composition_poly.rs:21-35): trace length smaller than the polynomial
    if (trace_len >= ce_domain_size) CTX_FAIL(ctx, AERO_ERR_INVALID, "trace length must be smaller than the constraint evaluation domain");
    const int logN = ilog2(ce_domain_size), logn = ilog2(trace_len);
    const uint64_t N = ce_domain_size;
    const int ncols = (int)(N / trace_len);
    if (ncols > 64) CTX_FAIL(ctx, AERO_ERR_UNSUPPORTED, "constraint blowup > 64");
    const bool mont = ctx->form == AERO_FORM_MONTGOMERY;
    PowTable gN;
    {
        char key[32];
        snprintf(key, sizeof key, "g/%d", logN);
        TRY(get_pow_table(ctx, key, gl::root_of_unity(logN), logN, 1, &gN));
    }
    const uint64_t *cols = d_eval_cols;
    uint64_t *combined = nullptr;
    // A sharded context combines its own block of rows [rank*N/G, (rank+1)*N/G) -- the only part of the
    // evaluation columns it reads, so the only part it has to upload -- and stores it into the peers.
    const int G = ctx->shard_world;
    if (N % (uint64_t)G) CTX_FAIL(ctx, AERO_ERR_UNSUPPORTED, "constraint evaluation domain smaller than the number of ranks");
    const uint64_t rows_per = N / G, row0 = rows_per * ctx->shard_rank;
    TRY(dev_alloc_shared(ctx, (void **)&combined, N * 8));
    std::vector<DivisorDev> dd(n_div);
    std::vector<uint64_t *> zbufs;
    aero_status st = AERO_OK;
    {
        PhaseTimer t(ctx, "constraint_combine");
        for (uint32_t d = 0; d < n_div && st == AERO_OK; d++) {
            const aero_divisor &v = divs[d];
            if (v.a == 0 || !is_pow2(v.a) || v.a > N) { ctx->err = "divisor degree must be a power of two <= domain size"; st = AERO_ERR_INVALID; break; }
            if (v.n_exemptions > 8) { ctx->err = "at most 8 exemption points"; st = AERO_ERR_INVALID; break; }
            DivisorDev &o = dd[d];
            o.a = v.a;
            o.b = to_canon(ctx, v.b);
            o.off_pow_a = gl::pow(gl::GENERATOR, v.a);
            o.nex = v.n_exemptions;
            for (uint32_t k = 0; k < 8; k++) o.ex[k] = k < v.n_exemptions ? to_canon(ctx, v.exemptions[k]) : 0;
            o.zn = (uint32_t)(N / v.a);
            uint64_t *z = nullptr;
            st = dev_alloc(ctx, (void **)&z, (size_t)o.zn * 8);
            if (st != AERO_OK) break;
            zbufs.push_back(z);
            divisor_inverses(o, z, logN, gN, ctx->stream);
            o.zinv = z;
        }
        if (st == AERO_OK && !own_cosets)
            constraint_combine(cols, col_stride, dd.data(), (int)n_div, logN, gl::GENERATOR, gN, (uint32_t)row0, (uint32_t)rows_per,
                               combined, ctx->stream);
        if (st == AERO_OK && own_cosets)  // `combined` is coset-major on this route
            constraint_combine(cols, col_stride, dd.data(), (int)n_div, logN, gl::GENERATOR, gN, tau0, tau_count, combined, ctx->stream,
                               logn, logN - logn);
    }
    if (st == AERO_OK && G > 1) {
        if (own_cosets) peer_push_words(rank_ptrs(ctx, combined), G, ctx->shard_rank, tau0, tau_count, ctx->stream);
        else peer_push_words(rank_ptrs(ctx, combined), G, ctx->shard_rank, row0, rows_per, ctx->stream);
        st = window_barrier(ctx);
    }
    aero_segment *seg = nullptr;
    if (st == AERO_OK) {
        seg = new aero_segment();
        seg->ctx = ctx;
        seg->ncols = ncols;
        seg->logn = logn;
        st = dev_alloc(ctx, (void **)&seg->polys, N * 8);
    }
    const int logB = logN - logn;
    if (st == AERO_OK && (logN > NTT_MAX_LOG || ctx->force_split_intt || own_cosets)) {
        // N = B*n beyond the two-pass NTT: B plain size-n interpolations (one per LDE coset) and a
        // B-point inverse DFT across them (coset_interp_combine in poly.cu)
        if (logB > 4 || logn > NTT_MAX_LOG) {
            ctx->err = "constraint evaluation domain too large (trace <= 2^24 rows, constraint blowup <= 16 at this size)";
            st = AERO_ERR_UNSUPPORTED;
        }
        const DftTables *plan = nullptr;
        PowTable ginv, oinv;
        uint64_t *M = nullptr, *cm = nullptr, *a = nullptr, *tmp = nullptr;
        char key[48];
        if (st == AERO_OK) st = plan_intt(ctx, logn, mont, &plan);
        snprintf(key, sizeof key, "ginv/%d/%d", logN, logn);
        if (st == AERO_OK) st = get_pow_table(ctx, key, gl::inv(gl::root_of_unity(logN)), logn, 1, &ginv);
        snprintf(key, sizeof key, "oinv/%d", logn);
        if (st == AERO_OK) st = get_pow_table(ctx, key, gl::inv(gl::GENERATOR), logn, 1, &oinv);
        snprintf(key, sizeof key, "cinterp/%d/%d", logn, logB);
        if (st == AERO_OK) {
            auto it = ctx->const_tables.find(key);
            if (it == ctx->const_tables.end()) {
                const int B = 1 << logB;
                std::vector<uint64_t> m((size_t)B * B);
                const uint64_t wBinv = gl::inv(gl::root_of_unity(logB));
                const uint64_t on_inv = gl::inv(gl::pow(gl::GENERATOR, trace_len));  // offset^-n
                uint64_t ck = gl::inv((uint64_t)B);
                for (int k = 0; k < B; k++) {
                    const uint64_t wk = gl::pow(wBinv, (uint64_t)k);
                    uint64_t x = ck;
                    for (int r = 0; r < B; r++) {
                        m[(size_t)k * B + r] = x;
                        x = gl::mul(x, wk);
                    }
                    ck = gl::mul(ck, on_inv);
                }
                st = upload_vec(ctx, &M, m);
                if (st == AERO_OK) ctx->const_tables[key] = M;
            } else {
                M = it->second;
            }
        }
        if (st == AERO_OK && !own_cosets) st = dev_alloc(ctx, (void **)&cm, N * 8);
        if (st == AERO_OK) st = dev_alloc(ctx, (void **)&a, N * 8);
        if (st == AERO_OK && plan->log1 != 0) st = dev_alloc(ctx, (void **)&tmp, N * 8);
        if (st == AERO_OK) {
            PhaseTimer t(ctx, "composition_poly");
            if (!own_cosets) natural_to_coset_major(combined, cm, logn, logB, ctx->stream);
            DftLaunch l;
            l.src = own_cosets ? combined : cm;
            l.dst = a;
            l.tmp = tmp;
            l.src_col_stride = trace_len;
            l.dst_col_stride = trace_len;
            l.ncols = ncols;
            l.deinterleave_log = 0;
            dft_run(*plan, l, ctx->stream);
            coset_interp_combine(a, ginv, oinv, M, logn, logB, seg->polys, ctx->stream);
        }
        dev_free(ctx, tmp);
        dev_free(ctx, a);
        dev_free(ctx, cm);
    } else if (st == AERO_OK) {
        const DftTables *plan;
        st = plan_coset_intt(ctx, logN, mont, &plan);
        if (st == AERO_OK) {
            PhaseTimer t(ctx, "composition_poly");
            uint64_t *tmp = nullptr;
            if (plan->log1 != 0) st = dev_alloc(ctx, (void **)&tmp, N * 8);
            if (st == AERO_OK) {
                DftLaunch l;
                l.src = combined;
                l.dst = seg->polys;
                l.tmp = tmp;
                l.src_col_stride = N;
                l.dst_col_stride = N;
                l.ncols = 1;
                l.deinterleave_log = ilog2((uint64_t)ncols);  // coefficient i -> column i % ncols, row i / ncols
                dft_run(*plan, l, ctx->stream);
                dev_free(ctx, tmp);
            }
        }
    }
    for (auto z : zbufs) dev_free(ctx, z);
    dev_free(ctx, combined);
    if (st == AERO_OK && cudaGetLastError() != cudaSuccess) { ctx->err = "kernel launch failed in constraints_into_poly"; st = AERO_ERR_CUDA; }
    if (st != AERO_OK) {
        aero_segment_destroy(seg);
        return st;
    }
    *out = seg;
    return AERO_OK;
}

aero_status aero_constraints_into_poly_device(aero_ctx *ctx, const uint64_t *d_eval_cols, size_t col_stride,
                                              const aero_divisor *divs, uint32_t n_div, uint64_t ce_domain_size,
                                              uint64_t trace_len, aero_segment **out) {
    return constraints_into_poly_impl(ctx, d_eval_cols, col_stride, divs, n_div, ce_domain_size, trace_len, out, false, 0, 0);
}

aero_status aero_constraints_into_poly(aero_ctx *ctx, const uint64_t *const *eval_cols, const aero_divisor *divs,
                                       uint32_t n_div, uint64_t ce_domain_size, uint64_t trace_len,
                                       aero_segment **out) {
    if (!ctx) return AERO_ERR_INVALID;
    enter(ctx);
    if (!eval_cols || !divs || !out) CTX_FAIL(ctx, AERO_ERR_INVALID, "null argument");
    if (n_div == 0 || n_div > 8) CTX_FAIL(ctx, AERO_ERR_INVALID, "number of divisors must be 1..8, got %u", n_div);
    if (!is_pow2(ce_domain_size)) CTX_FAIL(ctx, AERO_ERR_INVALID, "domain sizes must be powers of two");
    const uint64_t N = ce_domain_size;
    DevBlocks blk(ctx);
    uint64_t *cols = nullptr;
    TRY(blk.alloc((void **)&cols, (size_t)n_div * N * 8));
    {
        PhaseTimer t(ctx, "h2d");
        const uint64_t per = N / ctx->shard_world, r0 = per * ctx->shard_rank;  // own row block (all rows on one GPU)
        for (uint32_t d = 0; d < n_div; d++) {
            if (!eval_cols[d]) CTX_FAIL(ctx, AERO_ERR_INVALID, "null evaluation column %u", d);
            CUDA_TRY(ctx, cudaMemcpyAsync(cols + (size_t)d * N + r0, eval_cols[d] + r0, per * 8, cudaMemcpyHostToDevice, ctx->stream));
        }
    }
    return aero_constraints_into_poly_device(ctx, cols, N, divs, n_div, ce_domain_size, trace_len, out);
}

aero_status aero_constraints_evaluate_into_poly(aero_ctx *ctx, aero_segment *const *trace_segs, uint32_t n_trace_segs,
                                                const aero_air_program *prog, const uint64_t *coeffs, uint32_t n_coeffs,
                                                uint32_t ce_blowup, const aero_divisor *divs, uint32_t n_div, aero_segment **out) {
    if (!ctx) return AERO_ERR_INVALID;
    enter(ctx);
    if (!trace_segs || !n_trace_segs || !trace_segs[0] || !out) CTX_FAIL(ctx, AERO_ERR_INVALID, "null argument");
    if (n_div == 0 || n_div > 8 || ce_blowup < 2 || ce_blowup > 128) CTX_FAIL(ctx, AERO_ERR_INVALID, "bad divisor count or constraint evaluation blowup");
    const uint64_t n = trace_segs[0]->n(), CE = n * ce_blowup;
    DevBlocks blk(ctx);
    uint64_t *d_ce = nullptr;
    TRY(blk.alloc((void **)&d_ce, (size_t)n_div * CE * 8));
    uint32_t tau0 = 0, tau_count = 0;
    TRY(constraints_evaluate_impl(ctx, trace_segs, n_trace_segs, prog, coeffs, n_coeffs, ce_blowup, n_div, d_ce, CE, &tau0, &tau_count));
    // one GPU: the whole table is there, in natural order; several: each rank combines the cosets it evaluated
    return constraints_into_poly_impl(ctx, d_ce, CE, divs, n_div, CE, n, out, ctx_sharded(ctx), tau0, tau_count);
}

// ---- OOD + DEEP -----------------------------------------------------------------------------
// Evaluation of the columns of `seg` at `points`, results [col][point] at res + res_off.  A sharded context
// evaluates the columns it owns (own_columns; every rank holds all coefficients) and stores them into the
// result buffer of every rank.  Nothing is synchronised here.
static aero_status ood_enqueue(aero_ctx *ctx, aero_segment *seg, const std::vector<uint64_t> &points, uint64_t *res,
                               size_t res_off, DevBlocks &blk) {
    const int logn = seg->logn;
    const int np = (int)points.size();
    const uint64_t n = seg->n();
    int cb = 0, ce = seg->ncols;
    own_columns(ctx, seg->ncols, &cb, &ce);
    if (ce == cb) return AERO_OK;
    const int chunk_len = n < 4096 ? (int)n : 4096;
    const int nchunks = (int)(n / chunk_len);
    const int stride = 257 + nchunks + 16;  // layout: see ood_partial_kernel
    std::vector<uint64_t> tab((size_t)np * stride);
    for (int p = 0; p < np; p++) {
        uint64_t *t = tab.data() + (size_t)p * stride;
        uint64_t x = 1;
        for (int i = 0; i <= 256; i++) {
            t[i] = x;
            x = gl::mul(x, points[p]);
        }
        const uint64_t xc = gl::pow(points[p], (uint64_t)chunk_len);
        x = 1;
        for (int c = 0; c < nchunks; c++) {
            t[257 + c] = x;
            x = gl::mul(x, xc);
        }
        x = 1;
        for (int l = 0; l < 16; l++) {
            t[257 + nchunks + l] = x;
            x = gl::mul(x, t[256]);
        }
    }
    uint64_t *d_tab = nullptr, *d_scr = nullptr;
    TRY(blk.alloc((void **)&d_tab, tab.size() * 8));
    TRY(blk.alloc((void **)&d_scr, ood_scratch_elems(ce - cb, logn, np) * 8));
    TRY(upload_small(ctx, d_tab, tab.data(), tab.size() * 8));
    const size_t first = res_off + (size_t)cb * np, count = (size_t)(ce - cb) * np;
    ood_eval(seg->polys + (size_t)cb * n, n, ce - cb, logn, d_tab, np, res + first, d_scr, ctx->stream);
    peer_push_words(rank_ptrs(ctx, res), ctx->shard_world, ctx->shard_rank, first, count, ctx->stream);
    return AERO_OK;
}

aero_status aero_ood_eval(aero_ctx *ctx, aero_segment *const *trace_segs, uint32_t n_trace_segs, aero_segment *comp,
                          uint64_t z, uint64_t *out_trace, uint64_t *out_comp) {
    if (!ctx) return AERO_ERR_INVALID;
    enter(ctx);
    if (n_trace_segs && (!trace_segs || !out_trace)) CTX_FAIL(ctx, AERO_ERR_INVALID, "null argument");
    if (comp && !out_comp) CTX_FAIL(ctx, AERO_ERR_INVALID, "null argument");
    const uint64_t zc = to_canon(ctx, z);
    int W = 0;
    for (uint32_t s = 0; s < n_trace_segs; s++) {
        if (!trace_segs[s] || trace_segs[s]->logn != trace_segs[0]->logn) CTX_FAIL(ctx, AERO_ERR_INVALID, "trace segments must have equal length");
        W += trace_segs[s]->ncols;
    }
    const size_t total = (size_t)2 * W + (comp ? comp->ncols : 0);
    if (total == 0) return AERO_OK;
    // all segments are queued before the one synchronisation; results: [seg][col][point]
    DevBlocks blk(ctx);
    uint64_t *res = nullptr;
    TRY(blk.alloc_shared((void **)&res, total * 8));
    std::vector<size_t> seg_off(n_trace_segs + 1, 0);
    {
        PhaseTimer t(ctx, "ood_eval");
        size_t off = 0;
        if (n_trace_segs) {
            const uint64_t g = gl::root_of_unity(trace_segs[0]->logn);
            const std::vector<uint64_t> pts = {zc, gl::mul(zc, g)};
            for (uint32_t s = 0; s < n_trace_segs; s++) {
                seg_off[s] = off;
                TRY(ood_enqueue(ctx, trace_segs[s], pts, res, off, blk));
                off += (size_t)2 * trace_segs[s]->ncols;
            }
        }
        seg_off[n_trace_segs] = off;
        if (comp) {
            const std::vector<uint64_t> pts = {gl::pow(zc, (uint64_t)comp->ncols)};
            TRY(ood_enqueue(ctx, comp, pts, res, off, blk));
        }
        CUDA_TRY(ctx, cudaGetLastError());
    }
    TRY(window_barrier(ctx));
    std::vector<uint64_t> h(total);
    TRY(download_small_checked(ctx, h.data(), res, total * 8));
    int c0 = 0;
    for (uint32_t s = 0; s < n_trace_segs; s++) {
        const uint64_t *r = h.data() + seg_off[s];  // [col][point]
        const int nc = trace_segs[s]->ncols;
        for (int c = 0; c < nc; c++) {
            out_trace[c0 + c] = from_canon(ctx, r[(size_t)c * 2]);
            out_trace[W + c0 + c] = from_canon(ctx, r[(size_t)c * 2 + 1]);
        }
        c0 += nc;
    }
    if (comp) {
        const uint64_t *r = h.data() + seg_off[n_trace_segs];
        for (int c = 0; c < comp->ncols; c++) out_comp[c] = from_canon(ctx, r[c]);
    }
    return AERO_OK;
}

aero_status aero_deep_compose(aero_ctx *ctx, aero_segment *const *trace_segs, uint32_t n_trace_segs,
                              aero_segment *comp, uint64_t z, const uint64_t *ood_trace, const uint64_t *ood_comp,
                              const uint64_t *cc, aero_fri **out) {
    if (!ctx) return AERO_ERR_INVALID;
    enter(ctx);
    if (!trace_segs || !n_trace_segs || !comp || !ood_trace || !ood_comp || !cc || !out) CTX_FAIL(ctx, AERO_ERR_INVALID, "null argument");
    if (n_trace_segs > 4) CTX_FAIL(ctx, AERO_ERR_UNSUPPORTED, "at most 4 trace segments");
    const int logn = trace_segs[0]->logn;
    const int log_blowup = trace_segs[0]->log_blowup;
    if (log_blowup < 1) CTX_FAIL(ctx, AERO_ERR_STATE, "trace segments must be committed first");
    int W = 0;
    for (uint32_t s = 0; s < n_trace_segs; s++) {
        if (trace_segs[s]->logn != logn) CTX_FAIL(ctx, AERO_ERR_INVALID, "trace segments must have equal length");
        W += trace_segs[s]->ncols;
    }
    if (comp->logn != logn) CTX_FAIL(ctx, AERO_ERR_INVALID, "composition columns must have trace length");
    const int m = comp->ncols;
    const uint64_t n = 1ULL << logn;
    const uint64_t zc = to_canon(ctx, z);
    const uint64_t g = gl::root_of_unity(logn);
    const uint64_t zg = gl::mul(zc, g), zm = gl::pow(zc, (uint64_t)m);
    // coefficient vector for the kernel: (cc_i.0, cc_i.1) per trace column, then cc'_j
    std::vector<uint64_t> h_cc((size_t)2 * W + m);
    uint64_t k1 = 0, k2 = 0, kh = 0;
    for (int i = 0; i < W; i++) {
        const uint64_t c0 = to_canon(ctx, cc[3 * i]), c1 = to_canon(ctx, cc[3 * i + 1]);
        h_cc[2 * i] = c0;
        h_cc[2 * i + 1] = c1;
        k1 = gl::add(k1, gl::mul(c0, to_canon(ctx, ood_trace[i])));       // acc_trace_poly, composer/mod.rs:286
        k2 = gl::add(k2, gl::mul(c1, to_canon(ctx, ood_trace[W + i])));
    }
    for (int j = 0; j < m; j++) {
        const uint64_t c = to_canon(ctx, cc[3 * W + j]);
        h_cc[2 * W + j] = c;
        kh = gl::add(kh, gl::mul(c, to_canon(ctx, ood_comp[j])));
    }
    const uint64_t d0 = to_canon(ctx, cc[3 * W + m]), d1 = to_canon(ctx, cc[3 * W + m + 1]);
    const uint64_t h_consts[3] = {k1, k2, kh};

    DevBlocks tmp(ctx);
    uint64_t *d_cc = nullptr, *d_consts = nullptr, *t1 = nullptr, *t2 = nullptr, *hh = nullptr, *carry = nullptr, *coeffs = nullptr;
    TRY(tmp.alloc((void **)&d_cc, h_cc.size() * 8));
    TRY(tmp.alloc((void **)&d_consts, 3 * 8));
    // (a sharded context accumulates its own range of coefficient indices and stores it into the peers)
    TRY(tmp.alloc_shared((void **)&t1, n * 8));
    TRY(tmp.alloc_shared((void **)&t2, n * 8));
    TRY(tmp.alloc_shared((void **)&hh, n * 8));
    TRY(tmp.alloc((void **)&coeffs, n * 8));
    TRY(tmp.alloc((void **)&carry, (6 * (n / 256 + 1)) * 8));
    TRY(upload_small(ctx, d_cc, h_cc.data(), h_cc.size() * 8));
    TRY(upload_small(ctx, d_consts, h_consts, 24));
    {
        PhaseTimer t(ctx, "deep_compose");
        DeepSegs segs;
        segs.nseg = (int)n_trace_segs;
        for (uint32_t s = 0; s < n_trace_segs; s++) {
            segs.p[s] = trace_segs[s]->polys;
            segs.ncols[s] = trace_segs[s]->ncols;
        }
        const int G = ctx->shard_world;
        const uint32_t per = (uint32_t)(n / G), j0 = per * (uint32_t)ctx->shard_rank;
        deep_accumulate(segs, comp->polys, m, logn, j0, per, d_cc, d_consts, t1, t2, hh, ctx->stream);
        if (G > 1) {
            for (uint64_t *buf : {t1, t2, hh}) peer_push_words(rank_ptrs(ctx, buf), G, ctx->shard_rank, j0, per, ctx->stream);
            TRY(window_barrier(ctx));
        }
        const uint64_t bs[3] = {zc, zg, zm};
        syn_div3(t1, t2, hh, logn, bs, carry, ctx->stream);
        deep_finish(t1, t2, hh, logn, d0, d1, coeffs, ctx->stream);
    }
    FriGuard fri(new aero_fri());
    fri->ctx = ctx;
    const uint64_t N = n << log_blowup;
    const int B = 1 << log_blowup;
    if (B % ctx->shard_world) CTX_FAIL(ctx, AERO_ERR_UNSUPPORTED, "shard world size %d must divide the blowup factor %d", ctx->shard_world, B);
    TRY(dev_alloc_shared(ctx, (void **)&fri->cur, N * 8));
    fri->curM = (uint32_t)N;
    fri->cur_log_cosets = log_blowup;
    const DftTables *plan;
    TRY(plan_lde(ctx, logn, log_blowup, false, &plan));
    uint64_t *tmp_l = nullptr;
    // coset shard: own cosets only, written at their global slots and into the peers' copies
    const int coset_count = B / ctx->shard_world, coset_begin = ctx->shard_rank * coset_count;
    if (plan->log1 != 0) TRY(tmp.alloc((void **)&tmp_l, (size_t)coset_count * n * 8));
    {
        PhaseTimer t(ctx, "deep_lde");
        DftLaunch l;
        l.src = coeffs;
        l.tmp = tmp_l;
        l.src_col_stride = n;
        l.dst_col_stride = N;
        l.ncols = 1;
        l.deinterleave_log = 0;
        l.coset_begin = coset_begin;
        l.coset_count = coset_count;
        l.dst = fri->cur + (size_t)coset_begin * n;
        dft_run(*plan, l, ctx->stream);
    }
    if (ctx_sharded(ctx)) {
        {
            PhaseTimer t(ctx, "push_deep");
            peer_push(rank_ptrs(ctx, fri->cur), ctx->shard_world, ctx->shard_rank, (size_t)coset_begin * n * 8, (size_t)coset_count * n * 8, ctx->stream);
        }
        TRY(window_barrier(ctx));
    }
    if (cudaGetLastError() != cudaSuccess) CTX_FAIL(ctx, AERO_ERR_CUDA, "kernel launch failed in deep_compose");
    *out = fri.release();
    return AERO_OK;
}

// ---- FRI ------------------------------------------------------------------------------------
aero_status aero_fri_from_evaluations(aero_ctx *ctx, const uint64_t *evaluations, uint64_t count, aero_fri **out) {
    if (!ctx) return AERO_ERR_INVALID;
    enter(ctx);
    if (!evaluations || !out) CTX_FAIL(ctx, AERO_ERR_INVALID, "null argument");
    if (!is_pow2(count) || count < 16 || count > (1ULL << 31)) CTX_FAIL(ctx, AERO_ERR_INVALID, "number of evaluations must be a power of two >= 16");
    FriGuard fri(new aero_fri());
    fri->ctx = ctx;
    TRY(dev_alloc(ctx, (void **)&fri->cur, count * 8));
    CUDA_TRY(ctx, cudaMemcpyAsync(fri->cur, evaluations, count * 8, cudaMemcpyHostToDevice, ctx->stream));
    if (ctx->form == AERO_FORM_MONTGOMERY) convert_form(fri->cur, fri->cur, count, 0, ctx->stream);
    fri->curM = (uint32_t)count;
    fri->cur_log_cosets = 0;
    *out = fri.release();
    return AERO_OK;
}
aero_status aero_fri_download_evaluations(aero_fri *fri, uint64_t *out, uint64_t *count) {
    if (!fri || !count) return AERO_ERR_INVALID;
    aero_ctx *ctx = fri->ctx;
    enter(ctx);
    if (!out || *count < fri->curM) {
        *count = fri->curM;
        CTX_FAIL(ctx, AERO_ERR_BUFFER, "need room for %u evaluations", fri->curM);
    }
    *count = fri->curM;
    DevBlocks blk(ctx);
    uint64_t *tmp = nullptr;
    TRY(blk.alloc((void **)&tmp, (size_t)fri->curM * 8));
    const bool mont = ctx->form == AERO_FORM_MONTGOMERY;
    if (fri->cur_log_cosets) {
        lde_to_natural(fri->cur, tmp, ilog2(fri->curM) - fri->cur_log_cosets, fri->cur_log_cosets, mont, ctx->stream);
    } else if (mont) {
        convert_form(fri->cur, tmp, fri->curM, 1, ctx->stream);
    } else {
        CUDA_TRY(ctx, cudaMemcpyAsync(tmp, fri->cur, (size_t)fri->curM * 8, cudaMemcpyDeviceToDevice, ctx->stream));
    }
    CUDA_TRY(ctx, cudaMemcpyAsync(out, tmp, (size_t)fri->curM * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return AERO_OK;
}

// transpose_slice + hash_values + MerkleTree::new of the current evaluations, queued on the stream
// next_full != nullptr: the tree array was allocated by the previous layer's fused fold-and-hash, which has
// already written the leaf digests into it
static aero_status fri_commit_enqueue(aero_fri *fri, uint32_t *next_full = nullptr) {
    aero_ctx *ctx = fri->ctx;
    enter(ctx);
    if (!fri->cur) CTX_FAIL(ctx, AERO_ERR_STATE, "no evaluations to commit");
    if (fri->cur_committed) CTX_FAIL(ctx, AERO_ERR_STATE, "layer already committed; fold first");
    const uint32_t rows = fri->curM / 8;
    if (rows < 2) CTX_FAIL(ctx, AERO_ERR_INVALID, "FRI layer of %u evaluations is too small to commit", fri->curM);
    if (fri->cur_log_cosets && (rows >> fri->cur_log_cosets) == 0) CTX_FAIL(ctx, AERO_ERR_UNSUPPORTED, "layer too small for coset-major layout");
    FriLayerDev L;
    L.evals = fri->cur;
    L.M = fri->curM;
    L.log_cosets = fri->cur_log_cosets;
    if (next_full) L.full = next_full;
    else TRY(dev_alloc(ctx, (void **)&L.full, (size_t)2 * rows * 32));
    {
        PhaseTimer t(ctx, "fri_commit");
        CUDA_TRY(ctx, cudaMemsetAsync(L.full, 0, 64, ctx->stream));
        if (!next_full) fri_leaf_hash(L.evals, rows, L.log_cosets, L.full + (size_t)rows * 8, ctx->stream);
        merkle_build(L.full, rows, ctx->stream);
    }
    fri->layers.push_back(L);
    fri->cur_committed = true;
    return AERO_OK;
}
// apply_drp of the committed layer; the challenge comes by value or from device memory (alpha_dev)
// next_full_out != nullptr (challenge on the device): fused fold-and-hash where the shape allows -- the leaf
// digests of the folded layer go into a freshly allocated tree array returned in *next_full_out (nullptr if
// the plain fold ran).
static aero_status fri_fold_enqueue(aero_fri *fri, uint64_t alpha_canon, const uint64_t *alpha_dev, uint32_t **next_full_out = nullptr) {
    aero_ctx *ctx = fri->ctx;
    enter(ctx);
    if (next_full_out) *next_full_out = nullptr;
    if (!fri->cur_committed) CTX_FAIL(ctx, AERO_ERR_STATE, "commit the layer before folding it");
    const uint32_t M = fri->curM, rows = M / 8;
    const int logM = ilog2(M);
    PowTable xinv;
    {
        // x_j^-1 = (7 * g_M^j)^-1 = 7^-1 * (g_M^-1)^j,  j < rows
        char key[32];
        snprintf(key, sizeof key, "xinv/%d", logM);
        TRY(get_pow_table(ctx, key, gl::inv(gl::root_of_unity(logM)), logM - 3, gl::inv(gl::GENERATOR), &xinv));
    }
    const uint64_t w8i = gl::inv(gl::root_of_unity(3));
    const uint64_t w[4] = {1, w8i, gl::mul(w8i, w8i), gl::mul(gl::mul(w8i, w8i), w8i)};
    uint64_t *next = nullptr;
    TRY(dev_alloc(ctx, (void **)&next, (size_t)rows * 8));
    const uint32_t rows2 = rows / 8;  // leaves of the folded layer
    // Fused fold-and-hash (fri_fold_hash_kernel) runs rows / 8 threads of eight folds + four compressions each.
    // Measured on B200 (FRI of a 2^20-row proof, profiles/r02_bench_x5_*.json, r02_bench_x7_*.json): 0.490 ms
    // with separate kernels, 0.546 ms with every layer fused, 0.571 ms with only the layers of <= 2^14 folded
    // leaves fused -- one launch per layer is saved, but every thread's dependent chain gets eight folds longer
    // and the layers are latency-, not launch-bound.  Off by default ("fri_fused": 1 = small layers, 2 = all).
    const bool fuse = next_full_out && alpha_dev && ctx->fri_fused && (rows2 <= (1u << 14) || ctx->fri_fused >= 2) && rows % 8 == 0 && rows2 >= 2 &&
                      (fri->cur_log_cosets == 0 || (rows2 >> fri->cur_log_cosets) << fri->cur_log_cosets == rows2);
    {
        PhaseTimer t(ctx, "fri_fold");
        if (fuse) {
            uint32_t *nf = nullptr;
            TRY(dev_alloc(ctx, (void **)&nf, (size_t)2 * rows2 * 32));
            fri_fold_hash(fri->cur, rows, fri->cur_log_cosets, alpha_dev, xinv, w, gl::inv(8), next, nf + (size_t)rows2 * 8, ctx->stream);
            *next_full_out = nf;
        } else {
            fri_fold(fri->cur, rows, fri->cur_log_cosets, alpha_canon, alpha_dev, xinv, w, gl::inv(8), next, ctx->stream);
        }
    }
    fri->cur = next;  // the committed layer keeps ownership of the old buffer
    fri->curM = rows;
    fri->cur_log_cosets = 0;
    fri->cur_committed = false;
    CUDA_TRY(ctx, cudaGetLastError());
    return AERO_OK;
}

aero_status aero_fri_commit_layer(aero_fri *fri, uint8_t root[32]) {
    if (!fri || !root) return AERO_ERR_INVALID;
    aero_ctx *ctx = fri->ctx;
    enter(ctx);
    TRY(fri_commit_enqueue(fri));
    TRY(download_small(ctx, root, fri->layers.back().full + 8, 32));
    CUDA_TRY(ctx, cudaGetLastError());
    return AERO_OK;
}

aero_status aero_fri_fold(aero_fri *fri, uint64_t alpha) {
    if (!fri) return AERO_ERR_INVALID;
    return fri_fold_enqueue(fri, to_canon(fri->ctx, alpha), nullptr);
}

static aero_status fri_build_layers_impl(aero_fri *fri, const uint8_t coin_seed[32], uint32_t num_layers, int grinding_bits,
                                        uint8_t *roots_out, uint64_t *alphas_out, uint64_t *nonce_out) {
    aero_ctx *ctx = fri->ctx;
    if (!coin_seed || !roots_out || !alphas_out) CTX_FAIL(ctx, AERO_ERR_INVALID, "null argument");
    if (!fri->layers.empty() || fri->cur_committed) CTX_FAIL(ctx, AERO_ERR_STATE, "FRI layers have already been built");
    if (num_layers > 32) CTX_FAIL(ctx, AERO_ERR_INVALID, "too many FRI layers");
    if (grinding_bits > 40) CTX_FAIL(ctx, AERO_ERR_UNSUPPORTED, "grinding factor above 40 bits");
    const uint32_t nl = num_layers + 1;  // + the remainder commitment
    // staging layout (host mirror at the same offsets): seed[32] | roots[nl][32] | alphas[nl] | nonce
    const size_t off_roots = 32, off_alpha = 32 + (size_t)nl * 32, off_nonce = off_alpha + (size_t)nl * 8, total = off_nonce + 8;
    TRY(stage_reserve(ctx, total));
    memcpy(ctx->h_stage, coin_seed, 32);
    uint8_t *d = ctx->d_stage;
    CUDA_TRY(ctx, cudaMemcpyAsync(d, ctx->h_stage, 32, cudaMemcpyHostToDevice, ctx->stream));
    uint32_t *next_full = nullptr;  // tree array of the next layer when the fold already hashed its leaves
    for (uint32_t l = 0; l < nl; l++) {
        TRY(fri_commit_enqueue(fri, next_full));
        next_full = nullptr;
        uint64_t *d_alpha = (uint64_t *)(d + off_alpha) + l;
        fri_coin((uint32_t *)d, fri->layers.back().full + 8, d_alpha, (uint32_t *)(d + off_roots + (size_t)l * 32), ctx->stream);
        // the reference also draws a challenge for the remainder layer and discards the fold (prover/mod.rs:174-183)
        if (l < num_layers) TRY(fri_fold_enqueue(fri, 0, d_alpha, &next_full));
    }
    // Grinding rides on the same round trip: the coin's seed after the last layer is already on the device
    // (draws never move it).  One batch of 2^18 nonces finds a 16-bit nonce with probability 98 %; the
    // ascending batches below keep the result the MINIMUM nonce, like the reference's serial search.
    const uint32_t batch = 1u << 18;
    unsigned long long *d_best = (unsigned long long *)(d + off_nonce);
    if (grinding_bits >= 0) {
        PhaseTimer t(ctx, "grind");
        CUDA_TRY(ctx, cudaMemsetAsync(d_best, 0xFF, 8, ctx->stream));
        pow_search((const uint32_t *)d, 1, batch, (uint32_t)grinding_bits, d_best, ctx->stream);
    }
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->h_stage + off_roots, d + off_roots, total - off_roots, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    CUDA_TRY(ctx, cudaGetLastError());
    memcpy(roots_out, ctx->h_stage + off_roots, (size_t)nl * 32);
    const uint64_t *al = (const uint64_t *)(ctx->h_stage + off_alpha);
    for (uint32_t l = 0; l < nl; l++) {
        if (al[l] == ~0ULL) CTX_FAIL(ctx, AERO_ERR_STATE, "failed to draw a FRI challenge for layer %u", l);
        alphas_out[l] = from_canon(ctx, al[l]);
    }
    if (grinding_bits >= 0) {
        unsigned long long best = *(const unsigned long long *)(ctx->h_stage + off_nonce);
        for (uint64_t base = 1 + batch; best == ~0ULL; base += batch) {
            pow_search((const uint32_t *)d, base, batch, (uint32_t)grinding_bits, d_best, ctx->stream);
            TRY(download_small(ctx, &best, d_best, 8));
        }
        *nonce_out = best;
    }
    return AERO_OK;
}
aero_status aero_fri_build_layers(aero_fri *fri, const uint8_t coin_seed[32], uint32_t num_layers, uint8_t *roots_out,
                                  uint64_t *alphas_out) {
    if (!fri) return AERO_ERR_INVALID;
    enter(fri->ctx);
    return fri_build_layers_impl(fri, coin_seed, num_layers, -1, roots_out, alphas_out, nullptr);
}
aero_status aero_fri_build_layers_grind(aero_fri *fri, const uint8_t coin_seed[32], uint32_t num_layers, uint32_t grinding_bits,
                                        uint8_t *roots_out, uint64_t *alphas_out, uint64_t *nonce_out) {
    if (!fri) return AERO_ERR_INVALID;
    enter(fri->ctx);
    if (!nonce_out) CTX_FAIL(fri->ctx, AERO_ERR_INVALID, "null argument");
    return fri_build_layers_impl(fri, coin_seed, num_layers, (int)grinding_bits, roots_out, alphas_out, nonce_out);
}

aero_status aero_fri_open(aero_fri *fri, const uint64_t *positions, uint32_t n_pos, uint8_t *out_bytes, size_t *len) {
    if (!fri || !positions || !len) return AERO_ERR_INVALID;
    enter(fri->ctx);
    GatherBatch gb(fri->ctx);
    FriOpening o;
    TRY(fri_open_plan(fri, positions, n_pos, gb, o));
    TRY(gb.run());
    return fri_open_finish(o, gb, out_bytes, len);
}

void aero_fri_destroy(aero_fri *fri) {
    if (!fri) return;
    bool cur_owned_by_layer = false;
    for (auto &L : fri->layers) {
        if (L.evals == fri->cur) cur_owned_by_layer = true;
        dev_free(fri->ctx, L.evals);
        dev_free(fri->ctx, L.full);
    }
    if (!cur_owned_by_layer) dev_free(fri->ctx, fri->cur);
    delete fri;
}

// ---- auxiliary-segment construction (running-product columns) -------------------------------------
aero_status aero_running_product_columns_device(aero_ctx *ctx, const uint64_t *d_multiplicands, size_t m_stride,
                                                const uint64_t *init, uint32_t n_cols, uint64_t n_rows, uint64_t *d_out,
                                                size_t out_stride) {
    if (!ctx) return AERO_ERR_INVALID;
    enter(ctx);
    if (!d_multiplicands || !init || !d_out) CTX_FAIL(ctx, AERO_ERR_INVALID, "null argument");
    if (n_cols == 0 || n_cols > 255 || n_rows < 2) CTX_FAIL(ctx, AERO_ERR_INVALID, "need 1..255 columns of at least two rows");
    if (m_stride + 1 < n_rows || out_stride < n_rows) CTX_FAIL(ctx, AERO_ERR_INVALID, "column stride smaller than the number of rows");
    DevBlocks blk(ctx);
    uint64_t *scr = nullptr, *d_init = nullptr;
    TRY(blk.alloc((void **)&scr, running_product_scratch_elems((int)n_cols, n_rows) * 8));
    TRY(blk.alloc((void **)&d_init, (size_t)n_cols * 8));
    TRY(upload_small(ctx, d_init, init, (size_t)n_cols * 8));
    PhaseTimer t(ctx, "aux_running_product");
    running_product(d_multiplicands, m_stride, d_init, (int)n_cols, n_rows, ctx->form == AERO_FORM_MONTGOMERY, d_out, out_stride,
                    scr, ctx->stream);
    CUDA_TRY(ctx, cudaGetLastError());
    return AERO_OK;
}
aero_status aero_running_product_columns(aero_ctx *ctx, const uint64_t *const *multiplicands, const uint64_t *init,
                                         uint32_t n_cols, uint64_t n_rows, uint64_t *const *cols_out) {
    if (!ctx) return AERO_ERR_INVALID;
    enter(ctx);
    if (!multiplicands || !init || !cols_out) CTX_FAIL(ctx, AERO_ERR_INVALID, "null argument");
    if (n_cols == 0 || n_cols > 255 || n_rows < 2) CTX_FAIL(ctx, AERO_ERR_INVALID, "need 1..255 columns of at least two rows");
    DevBlocks blk(ctx);
    uint64_t *d_m = nullptr, *d_o = nullptr;
    TRY(blk.alloc((void **)&d_m, (size_t)n_cols * n_rows * 8));
    TRY(blk.alloc((void **)&d_o, (size_t)n_cols * n_rows * 8));
    for (uint32_t c = 0; c < n_cols; c++) {
        if (!multiplicands[c] || !cols_out[c]) CTX_FAIL(ctx, AERO_ERR_INVALID, "null column %u", c);
        CUDA_TRY(ctx, cudaMemcpyAsync(d_m + (size_t)c * n_rows, multiplicands[c], (n_rows - 1) * 8, cudaMemcpyHostToDevice, ctx->stream));
    }
    TRY(aero_running_product_columns_device(ctx, d_m, n_rows, init, n_cols, n_rows, d_o, n_rows));
    for (uint32_t c = 0; c < n_cols; c++)
        CUDA_TRY(ctx, cudaMemcpyAsync(cols_out[c], d_o + (size_t)c * n_rows, n_rows * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return AERO_OK;
}
aero_status aero_batch_inverse(aero_ctx *ctx, const uint64_t *values, uint64_t count, uint64_t *out) {
    if (!ctx) return AERO_ERR_INVALID;
    enter(ctx);
    if (!values || !out) CTX_FAIL(ctx, AERO_ERR_INVALID, "null argument");
    if (!count) return AERO_OK;
    DevBlocks blk(ctx);
    uint64_t *d_v = nullptr, *d_o = nullptr;
    TRY(blk.alloc((void **)&d_v, count * 8));
    TRY(blk.alloc((void **)&d_o, count * 8));
    CUDA_TRY(ctx, cudaMemcpyAsync(d_v, values, count * 8, cudaMemcpyHostToDevice, ctx->stream));
    batch_inverse(d_v, count, ctx->form == AERO_FORM_MONTGOMERY, d_o, ctx->stream);
    CUDA_TRY(ctx, cudaMemcpyAsync(out, d_o, count * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    CUDA_TRY(ctx, cudaGetLastError());
    return AERO_OK;
}

// ---- grinding -------------------------------------------------------------------------------
aero_status aero_pow_min_nonce(aero_ctx *ctx, const uint8_t seed[32], uint32_t grinding_bits, uint64_t *nonce) {
    if (!ctx) return AERO_ERR_INVALID;
    enter(ctx);
    if (!seed || !nonce) CTX_FAIL(ctx, AERO_ERR_INVALID, "null argument");
    if (grinding_bits > 40) CTX_FAIL(ctx, AERO_ERR_UNSUPPORTED, "grinding factor above 40 bits");
    PhaseTimer t(ctx, "grind");
    DevBlocks blk(ctx);
    uint32_t *d_seed = nullptr;
    unsigned long long *d_best = nullptr;
    TRY(blk.alloc((void **)&d_seed, 32));
    TRY(blk.alloc((void **)&d_best, 8));
    TRY(upload_small(ctx, d_seed, seed, 32));
    CUDA_TRY(ctx, cudaMemsetAsync(d_best, 0xFF, 8, ctx->stream));
    const uint32_t batch = 1u << 18;
    uint64_t base = 1;
    unsigned long long best = ~0ULL;
    for (;;) {
        pow_search(d_seed, base, batch, grinding_bits, d_best, ctx->stream);
        TRY(download_small(ctx, &best, d_best, 8));
        if (best != ~0ULL) break;
        base += batch;
    }
    *nonce = best;
    return AERO_OK;
}

}  // extern "C"
