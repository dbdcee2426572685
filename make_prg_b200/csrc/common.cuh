// Shared definitions for libmprg (sm_100a).  See include/mprg.h for the ABI and DESIGN.md for the
// data layout.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <mutex>
#include <sched.h>
#include <string>
#include <vector>

#include "../../include/mprg.h"

namespace mprg {

constexpr int SYM_GAP = MPRG_SYM_GAP;
constexpr int SYM_N = MPRG_SYM_N;
constexpr int SYM_PAD = MPRG_SYM_PAD;
// the four unambiguous bases are the odd codes below 8 (see mprg.h)
__host__ __device__ __forceinline__ bool sym_is_base(int code) { return (code & 9) == 1; }
// R Y K M S W: what is neither gap, base, N nor padding
__host__ __device__ __forceinline__ bool sym_is_ambiguous(int code) {
    return code != SYM_GAP && code != SYM_N && code != SYM_PAD && !sym_is_base(code);
}
constexpr int COLS_PER_CHUNK = 32;  // one 16-byte vector load = 32 columns of one row
constexpr int CHUNK_BYTES = 16;

// Position of a column inside a packed row.  A chunk (32 columns, 16 bytes) is stored word-interleaved:
// column c of the chunk lives in 32-bit word (c & 3), nibble (c >> 2).  With this layout the four
// per-word "which nibbles are gaps" indicators of a chunk interleave, by three shifts and ORs, into
// one 32-bit mask whose bit c is column c (scan.cu).
__host__ __device__ __forceinline__ int packed_byte_of(int col) {
    const int c = col & 31;
    return ((col >> 5) << 4) + ((c & 3) << 2) + (c >> 3);
}
__host__ __device__ __forceinline__ int packed_shift_of(int col) { return ((col >> 2) & 1) << 2; }
__host__ __device__ __forceinline__ int packed_sym(const uint8_t *row, int col) {
    return (row[packed_byte_of(col)] >> packed_shift_of(col)) & 15;
}

// Device-side task descriptor (one sub-alignment).
struct DTask {
    long long base;   // byte offset of the locus in the packed arena
    int stride;       // bytes per packed row (multiple of 16)
    int rows_off;     // offset into the row-index arena, -1 => rows are 0..n_rows-1
    int n_rows;
    int c0, c1;       // column window [c0, c1) in locus coordinates
    int col_off;      // offset (columns, multiple of 32) of this task's per-column outputs; the
                      // outputs cover the chunk-aligned window [c0 & ~31, roundup(c1, 32))
    int iv_off;       // offset into the interval arena (capacity max(1, c1-c0))
    int flags;        // bit0: locus may hold N (use the wildcard path)
};

struct DInterval {
    int start, stop, type;
};

// A unit of scan work: a tile (row range x chunk range, at most 32 chunks) of one task, handled by
// one warp.  Self-contained (48 bytes, three vector loads) so that a warp needs a single dependent
// global load before it can issue its first row loads.
struct ScanUnit {
    long long base;  // byte offset of the locus in the packed arena
    int stride;      // bytes per packed row
    int rows_off;    // offset of the tile's first row index in the row arena, -1 => rows are consecutive
    int row_begin;   // first row (when rows_off < 0)
    int row_count;
    int c0, c1;      // column window of the task
    int col_off;     // per-column outputs of the task (see DTask)
    int ch_begin;    // first chunk of the tile (locus coordinates)
    int ch_count;    // chunks of the tile, 1..32
    int flags;       // bit0: the locus holds even symbol codes (M S W N): exact gap test
};
static_assert(sizeof(ScanUnit) == 48, "ScanUnit is loaded as three 16-byte vectors");

// Growable device buffer (never shrinks); all launches of a context share one stream, so reuse
// across calls is ordered.
struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        // buffers grow level by level within a build and batch by batch across builds: double while that is cheap
        // (every reallocation is a cudaFree that synchronises the device), a quarter of headroom above 1 GiB
        size_t want = bytes < ((size_t)1 << 30) ? 2 * bytes + 256 : bytes + bytes / 4 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess && want > bytes + 256) {
            cudaGetLastError();
            want = bytes + 256;
            e = cudaMalloc(&p, want);
        }
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
    template <typename T>
    T *as() const {
        return reinterpret_cast<T *>(p);
    }
};

struct PinnedBuf {
    void *p = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFreeHost(p);
        p = nullptr;
        cap = 0;
        size_t want = bytes + bytes / 4 + 256;
        cudaError_t e = cudaMallocHost(&p, want);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() {
        if (p) cudaFreeHost(p);
        p = nullptr;
        cap = 0;
    }
    template <typename T>
    T *as() const {
        return reinterpret_cast<T *>(p);
    }
};

}  // namespace mprg

struct mprg_batch {
    int n_loci = 0;
    std::vector<int> n_rows, n_cols, stride;
    std::vector<long long> base;  // byte offsets into d_packed
    std::vector<int> flags;       // alphabet flags per locus (host copy)
    long long packed_bytes = 0;
    uint8_t *d_packed = nullptr;
    size_t packed_capacity = 0;
    std::atomic<bool> any_n{false};  // set by whichever upload finds an N; a stale read only picks the slower scan variant
};

struct mprg_ctx {
    int device = 0;
    int sm_count = 0;
    int cc_major = 0, cc_minor = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t stream_side = nullptr;  // result copies that overlap the PRG assembly (engine_dev.cu)
    cudaEvent_t ev_side = nullptr;
    std::string err;
    long long launches = 0;
    int n_workers = 1;                // host threads / streams mprg_build may use
    int wait_mode = 0;                // mprg_set_wait_mode: 0 cudaStreamSynchronize, 1 sleep on ev_wait, 2 poll + yield
    cudaEvent_t ev_wait = nullptr;    // cudaEventBlockingSync | cudaEventDisableTiming, made on first use
    std::vector<mprg_ctx *> workers;  // lazily created worker contexts (same device)
    long long h2d_bytes = 0, d2h_bytes = 0;
    // which variants the engine chose (mprg_path_counts): see MPRG_PATH_* in mprg.h
    long long path_counts[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    cudaEvent_t ev_t0 = nullptr, ev_t1 = nullptr;  // mprg_timer
    // per-launch log of the scan kernel (algorithmic bytes, device ms), newest last
    std::vector<double> scan_log_bytes, scan_log_ms;
    // scan-kernel accounting for the roofline object
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    double scan_ms = 0, scan_bytes = 0;
    long long scan_launches = 0;
    // scratch (device)
    mprg::DevBuf d_tasks, d_units, d_rows, d_colwords, d_colB, d_cls, d_reach, d_iv, d_ivcnt, d_misc,
        d_stage;
    // scratch for clustering
    mprg::DevBuf d_c[16];
    // buffers of the device-resident level loop (engine_dev.cu): tree, row pool, alleles, task / problem tables
    mprg::DevBuf d_dev[32];
    mprg::DevBuf d_ref;  // scratch of the whole-grid one-reference-like check (refcheck_grid.cu)
    mprg::PinnedBuf h_cnt;  // the counter block the host reads twice per level
    mprg::PinnedBuf h_setup;  // staging of a range's start state (counters, locus table, root nodes, pending list)
    bool pending_scan = false;  // a scan launch whose events have not been read yet
    double pending_scan_bytes = 0;
    // device time of the KMeans launches of the level loop (mprg_kmeans_stats): event pairs of the current
    // clustering pass, read after the next synchronisation
    cudaEvent_t ev_km[2][12] = {};
    int km_pending = 0;
    long long km_pending_problems = 0;
    double km_ms = 0;
    long long km_launches = 0, km_problems = 0;
    // scratch (pinned host)
    mprg::PinnedBuf h_a, h_b, h_c, h_d;
    // packed arenas of freed batches, kept for the next batch of about the same size: cudaMalloc and
    // cudaFree of 100 MB cost more than a millisecond each and cudaFree synchronises the device
    std::mutex arena_mutex;
    std::vector<std::pair<uint8_t *, size_t>> idle_arenas;
};

namespace mprg {
// The host waits for everything queued on `st` (a stream of this context): see mprg_set_wait_mode.
inline cudaError_t wait_stream(mprg_ctx *c, cudaStream_t st) {
    if (c->wait_mode == 0) return cudaStreamSynchronize(st);
    cudaError_t e = cudaSuccess;
    if (c->wait_mode == 2) {
        // poll, and hand the core to any other runnable thread between polls: as quick as a spin on an idle host,
        // and the waiting lanes of several builds in flight do not starve the lanes that are launching
        while ((e = cudaStreamQuery(st)) == cudaErrorNotReady) sched_yield();
        return e;
    }
    if (!c->ev_wait) e = cudaEventCreateWithFlags(&c->ev_wait, cudaEventBlockingSync | cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventRecord(c->ev_wait, st);
    if (e == cudaSuccess) e = cudaEventSynchronize(c->ev_wait);
    return e;
}
inline cudaError_t copy_h2d(mprg_ctx *c, void *dst, const void *src, size_t n, cudaStream_t st) {
    c->h2d_bytes += (long long)n;
    return cudaMemcpyAsync(dst, src, n, cudaMemcpyHostToDevice, st);
}
inline cudaError_t copy_d2h(mprg_ctx *c, void *dst, const void *src, size_t n, cudaStream_t st) {
    c->d2h_bytes += (long long)n;
    return cudaMemcpyAsync(dst, src, n, cudaMemcpyDeviceToHost, st);
}
}  // namespace mprg

#define MPRG_CUDA(ctx, call)                                                                   \
    do {                                                                                       \
        cudaError_t e__ = (call);                                                              \
        if (e__ != cudaSuccess) {                                                              \
            (ctx)->err = std::string(#call) + ": " + cudaGetErrorString(e__);                  \
            return MPRG_E_CUDA;                                                                \
        }                                                                                      \
    } while (0)

#define MPRG_FAIL(ctx, code, msg) \
    do {                          \
        (ctx)->err = (msg);       \
        return (code);            \
    } while (0)
