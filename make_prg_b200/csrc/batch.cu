// Context management and the loader -> HBM step: ASCII MSAs are copied to the device and packed to
// 4 bits per symbol there (two columns per byte, rows padded to 16 bytes).
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <thread>

#include "common.cuh"
#include "kernels.cuh"

namespace mprg {

__constant__ uint8_t c_sym_lut[256];

static void build_lut(uint8_t *lut) {
    for (int i = 0; i < 256; ++i) lut[i] = SYM_PAD;
    const char *alphabet = MPRG_ALPHABET;
    for (int i = 0; i < 16; ++i) {
        const char ch = alphabet[i];
        if (ch == '?') continue;
        lut[(uint8_t)ch] = (uint8_t)i;
        if (ch >= 'A' && ch <= 'Z') lut[(uint8_t)(ch - 'A' + 'a')] = (uint8_t)i;
    }
}

// One warp per (locus,row): lanes produce consecutive packed 32-bit words (8 columns each).
// row_prefix[l] = number of rows of loci < l (row_prefix[n_loci] = total rows).
__global__ void __launch_bounds__(256)
pack_rows_kernel(const uint8_t *__restrict__ ascii, const long long *__restrict__ ascii_off,
                 const long long *__restrict__ row_prefix, const int *__restrict__ n_cols,
                 const long long *__restrict__ base, const int *__restrict__ stride, int n_loci,
                 long long total_rows, uint8_t *__restrict__ packed, int *__restrict__ flags) {
    const int lane = threadIdx.x & 31;
    const long long warp = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (warp >= total_rows) return;
    // binary search the locus of this row
    int lo = 0, hi = n_loci;  // row_prefix[lo] <= warp < row_prefix[hi]
    while (hi - lo > 1) {
        int mid = (lo + hi) >> 1;
        if (row_prefix[mid] <= warp) lo = mid; else hi = mid;
    }
    const int l = lo;
    const int r = (int)(warp - row_prefix[l]);
    const int C = n_cols[l];
    const uint8_t *src = ascii + ascii_off[l] + (long long)r * C;
    uint32_t *dst = reinterpret_cast<uint32_t *>(packed + base[l] + (long long)r * stride[l]);
    const int n_words = stride[l] >> 2;
    int f = 0;
    for (int w = lane; w < n_words; w += 32) {
        uint32_t word = 0;
        // word-interleaved chunk layout: word (w & 3) of chunk (w >> 2) holds columns 4k + (w & 3)
        const int cbase = (w >> 2) * 32 + (w & 3);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            uint32_t code = SYM_PAD;
            if (cbase + 4 * j < C) {
                const uint8_t ch = src[cbase + 4 * j];
                code = c_sym_lut[ch];
                if (code == SYM_PAD) f |= 1;
                else if (code == SYM_N) f |= 2;
                else if (sym_is_ambiguous(code)) f |= 4;
                if (code != SYM_GAP && !(code & 1u)) f |= 8;  // even code: the scan needs its exact gap test
            }
            word |= code << (4 * j);
        }
        dst[w] = word;
    }
    f = __reduce_or_sync(0xffffffffu, f);
    if (lane == 0 && f) atomicOr(&flags[l], f);
}

// The same packing with the ASCII rows read straight from PINNED HOST memory (zero copy, opt-in): a big
// cudaMemcpyAsync sits in the copy engine's queue in front of the small task-table and result copies of
// the ranges that are already being built (a saturating stream of 26 MB copies beside mprg_build slows
// it from 7.5 to 110 ms), and the staged copy is read once more from HBM.  Each warp stages the aligned
// 16-byte vectors that cover 512 columns of its row in shared memory with coalesced 128-bit loads over
// PCIe (an aligned vector that holds one valid byte lies in a mapped page), then packs from there; here
// ascii_off are offsets into the caller's buffer.
__global__ void __launch_bounds__(256)
pack_rows_hostmem_kernel(const uint8_t *__restrict__ ascii, const long long *__restrict__ ascii_off,
                         const long long *__restrict__ row_prefix, const int *__restrict__ n_cols,
                         const long long *__restrict__ base, const int *__restrict__ stride, int n_loci,
                         long long total_rows, uint8_t *__restrict__ packed, int *__restrict__ flags) {
    __shared__ uint4 s_stage[8][34];
    const int lane = threadIdx.x & 31;
    const long long warp = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (warp >= total_rows) return;
    int lo = 0, hi = n_loci;
    while (hi - lo > 1) {
        int mid = (lo + hi) >> 1;
        if (row_prefix[mid] <= warp) lo = mid; else hi = mid;
    }
    const int l = lo;
    const int r = (int)(warp - row_prefix[l]);
    const int C = n_cols[l];
    const uint8_t *src = ascii + ascii_off[l] + (long long)r * C;
    uint32_t *dst = reinterpret_cast<uint32_t *>(packed + base[l] + (long long)r * stride[l]);
    const int n_words = stride[l] >> 2;
    uint4 *stage = s_stage[threadIdx.x >> 5];
    int f = 0;
    for (int c0 = 0; c0 < n_words * 8; c0 += 512) {
        const int len = min(512, C - c0);  // valid columns of this span (<= 0: padding words only)
        const uint8_t *first = src + c0;
        const int lead = (int)(reinterpret_cast<unsigned long long>(first) & 15ull);
        if (len > 0) {
            const uint4 *aligned = reinterpret_cast<const uint4 *>(first - lead);
            const int n_vec = (lead + len + 15) >> 4;  // <= 33
            for (int k = lane; k < n_vec; k += 32) stage[k] = aligned[k];
        }
        __syncwarp();
        const uint8_t *sb = reinterpret_cast<const uint8_t *>(stage) + lead;
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            const int ws = lane + 32 * half;  // word of the span
            const int w = (c0 >> 3) + ws;
            if (w < n_words) {
                const int cbase = (ws >> 2) * 32 + (ws & 3);
                uint32_t word = 0;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    uint32_t code = SYM_PAD;
                    if (cbase + 4 * j < len) {
                        code = c_sym_lut[sb[cbase + 4 * j]];
                        if (code == SYM_PAD) f |= 1;
                        else if (code == SYM_N) f |= 2;
                        else if (sym_is_ambiguous(code)) f |= 4;
                        if (code != SYM_GAP && !(code & 1u)) f |= 8;
                    }
                    word |= code << (4 * j);
                }
                dst[w] = word;
            }
        }
        __syncwarp();
    }
    f = __reduce_or_sync(0xffffffffu, f);
    if (lane == 0 && f) atomicOr(&flags[l], f);
}

}  // namespace mprg

using namespace mprg;

extern "C" int mprg_create(int device_ordinal, mprg_ctx **out) {
    if (!out) return MPRG_E_BAD_ARG;
    *out = nullptr;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0 || device_ordinal < 0 || device_ordinal >= n)
        return MPRG_E_NO_DEVICE;
    mprg_ctx *ctx = new mprg_ctx();
    ctx->device = device_ordinal;
    if (cudaSetDevice(device_ordinal) != cudaSuccess) {
        delete ctx;
        return MPRG_E_NO_DEVICE;
    }
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, device_ordinal);
    ctx->sm_count = prop.multiProcessorCount;
    ctx->cc_major = prop.major;
    ctx->cc_minor = prop.minor;
    if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreate(&ctx->ev0) != cudaSuccess || cudaEventCreate(&ctx->ev1) != cudaSuccess ||
        cudaEventCreate(&ctx->ev_t0) != cudaSuccess || cudaEventCreate(&ctx->ev_t1) != cudaSuccess) {
        delete ctx;
        return MPRG_E_CUDA;
    }
    cudaDeviceSetLimit(cudaLimitStackSize, 8192);  // recursive pairwise summation in kmeans.cu
    uint8_t lut[256];
    build_lut(lut);
    if (cudaMemcpyToSymbol(c_sym_lut, lut, 256) != cudaSuccess) {
        delete ctx;
        return MPRG_E_CUDA;
    }
    if (const char *env = getenv("MPRG_WORKERS")) {
        const int n = atoi(env);
        if (n >= 1 && n <= 64) ctx->n_workers = n;
    } else {
        // at most 8 (sweeps in DESIGN.md section 4): half of the host cores for one process, the whole
        // share of this process when torchrun says several ranks drive their GPUs from this host
        // (4 ranks on 32 cores: 8 workers each 447 k loci/s, 4 each 404 k)
        const unsigned hc = std::thread::hardware_concurrency();
        unsigned share = hc / 2;
        if (const char *lw = getenv("LOCAL_WORLD_SIZE")) {
            const int n = atoi(lw);
            if (n > 1) share = hc / (unsigned)n;
        }
        ctx->n_workers = (int)std::max(hc >= 4 ? 2u : 1u, std::min(8u, share));
    }
    *out = ctx;
    return MPRG_OK;
}

extern "C" void mprg_destroy(mprg_ctx *ctx) {
    if (!ctx) return;
    for (mprg_ctx *w : ctx->workers) mprg_destroy(w);
    ctx->workers.clear();
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    DevBuf *bufs[] = {&ctx->d_tasks, &ctx->d_units, &ctx->d_rows, &ctx->d_colwords, &ctx->d_colB,
                      &ctx->d_cls,   &ctx->d_reach, &ctx->d_iv,   &ctx->d_ivcnt,    &ctx->d_misc,
                      &ctx->d_stage};
    for (DevBuf *b : bufs) b->release();
    for (DevBuf &b : ctx->d_c) b.release();
    for (DevBuf &b : ctx->d_dev) b.release();
    ctx->h_cnt.release();
    ctx->h_setup.release();
    ctx->d_ref.release();
    for (auto &a : ctx->idle_arenas) cudaFree(a.first);
    ctx->idle_arenas.clear();
    ctx->h_a.release();
    ctx->h_b.release();
    ctx->h_c.release();
    ctx->h_d.release();
    for (int a = 0; a < 2; ++a)
        for (int r = 0; r < 12; ++r)
            if (ctx->ev_km[a][r]) cudaEventDestroy(ctx->ev_km[a][r]);
    if (ctx->ev0) cudaEventDestroy(ctx->ev0);
    if (ctx->ev1) cudaEventDestroy(ctx->ev1);
    if (ctx->ev_t0) cudaEventDestroy(ctx->ev_t0);
    if (ctx->ev_t1) cudaEventDestroy(ctx->ev_t1);
    if (ctx->stream_side) cudaStreamDestroy(ctx->stream_side);
    if (ctx->ev_side) cudaEventDestroy(ctx->ev_side);
    if (ctx->ev_wait) cudaEventDestroy(ctx->ev_wait);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

extern "C" const char *mprg_last_error(const mprg_ctx *ctx) { return ctx ? ctx->err.c_str() : ""; }

extern "C" int mprg_device_info(const mprg_ctx *ctx, int *sm_count, int *cc_major, int *cc_minor) {
    if (!ctx) return MPRG_E_BAD_ARG;
    if (sm_count) *sm_count = ctx->sm_count;
    if (cc_major) *cc_major = ctx->cc_major;
    if (cc_minor) *cc_minor = ctx->cc_minor;
    return MPRG_OK;
}

extern "C" int64_t mprg_launch_count(const mprg_ctx *ctx) { return ctx ? ctx->launches : 0; }

extern "C" int mprg_scan_stats(mprg_ctx *ctx, double *ms, double *bytes, int64_t *launches,
                               int reset) {
    if (!ctx) return MPRG_E_BAD_ARG;
    if (ms) *ms = ctx->scan_ms;
    if (bytes) *bytes = ctx->scan_bytes;
    if (launches) *launches = ctx->scan_launches;
    if (reset) {
        ctx->scan_ms = 0;
        ctx->scan_bytes = 0;
        ctx->scan_launches = 0;
    }
    return MPRG_OK;
}

namespace mprg {

// Host metadata of a batch and its (still empty) packed arena in HBM.
int batch_prepare(mprg_ctx *ctx, const int32_t *n_rows, const int32_t *n_cols, int32_t n_loci, mprg_batch **out) {
    *out = nullptr;
    cudaSetDevice(ctx->device);
    mprg_batch *b = new mprg_batch();
    b->n_loci = n_loci;
    b->n_rows.assign(n_rows, n_rows + n_loci);
    b->n_cols.assign(n_cols, n_cols + n_loci);
    b->stride.resize(n_loci);
    b->base.resize(n_loci);
    b->flags.assign(n_loci, 0);
    long long packed = 0;
    for (int l = 0; l < n_loci; ++l) {
        if (n_rows[l] < 0 || n_cols[l] < 0) {
            delete b;
            MPRG_FAIL(ctx, MPRG_E_BAD_ARG, "negative MSA dimension");
        }
        b->stride[l] = ((n_cols[l] + COLS_PER_CHUNK - 1) / COLS_PER_CHUNK) * CHUNK_BYTES;
        b->base[l] = packed;
        packed += (long long)b->stride[l] * n_rows[l];
    }
    b->packed_bytes = packed;
    const size_t need = (size_t)packed + 16;
    {
        std::lock_guard<std::mutex> lock(ctx->arena_mutex);
        int best = -1;
        for (int i = 0; i < (int)ctx->idle_arenas.size(); ++i)
            if (ctx->idle_arenas[i].second >= need && (best < 0 || ctx->idle_arenas[i].second < ctx->idle_arenas[best].second))
                best = i;
        if (best >= 0 && ctx->idle_arenas[best].second <= 2 * need + ((size_t)64 << 20)) {
            b->d_packed = ctx->idle_arenas[best].first;
            b->packed_capacity = ctx->idle_arenas[best].second;
            ctx->idle_arenas.erase(ctx->idle_arenas.begin() + best);
            *out = b;
            return MPRG_OK;
        }
    }
    b->packed_capacity = need + need / 16;
    const cudaError_t e = cudaMalloc(&b->d_packed, b->packed_capacity);
    if (e != cudaSuccess) {
        ctx->err = std::string("cudaMalloc packed: ") + cudaGetErrorString(e);
        delete b;
        return MPRG_E_CUDA;
    }
    *out = b;
    return MPRG_OK;
}

// One big host-to-device copy at a time per GPU, across the ranges of one build AND across builds that are in
// flight side by side (device.BuildPipeline: the upload of build k+1 runs while build k is in its level loop):
// copies that share the link finish together, copies in turn let the first one's kernels start early.
static std::mutex &link_mutex(int device) {
    static std::mutex m[64];
    return m[(unsigned)device & 63u];
}

// Loci [l0, l1) of the batch: ASCII host -> device stage of `ctx` (its stream), packed on the device
// into the batch arena, alphabet flags back.  Ranges are independent, so the worker contexts of
// mprg_build_ascii upload theirs concurrently and the copies overlap the other workers' kernels.
int batch_upload_range(mprg_ctx *ctx, mprg_batch *b, const uint8_t *h_ascii, const int64_t *h_offsets, int l0,
                       int l1) {
    const int n = l1 - l0;
    if (n <= 0) return MPRG_OK;
    cudaSetDevice(ctx->device);
    long long ascii_total = 0, total_rows = 0;
    std::vector<long long> row_prefix(n + 1, 0), aoff(n);
    for (int i = 0; i < n; ++i) {
        const int l = l0 + i;
        aoff[i] = ascii_total;
        ascii_total += (long long)b->n_rows[l] * b->n_cols[l];
        row_prefix[i] = total_rows;
        total_rows += b->n_rows[l];
    }
    row_prefix[n] = total_rows;
    if (total_rows <= 0 || ascii_total <= 0) return MPRG_OK;
    // stage: ascii | ascii_off | row_prefix | base | n_cols | stride | flags
    const size_t o_ascii = 0;
    const size_t o_aoff = (ascii_total + 15) & ~15LL;
    const size_t o_rp = o_aoff + sizeof(long long) * n;
    const size_t o_base = o_rp + sizeof(long long) * (n + 1);
    const size_t o_nc = o_base + sizeof(long long) * n;
    const size_t o_st = o_nc + sizeof(int) * n;
    const size_t o_fl = o_st + sizeof(int) * n;
    const size_t total = o_fl + sizeof(int) * n;
    MPRG_CUDA(ctx, ctx->d_stage.reserve(total));
    uint8_t *d = ctx->d_stage.as<uint8_t>();
    cudaStream_t s = ctx->stream;
    // The rows are staged in HBM with cudaMemcpyAsync and packed from there.  With MPRG_ZEROCOPY=1 pinned
    // (device-visible) host memory is read by the pack kernel itself instead; measured slower here (11.7
    // against 10.8 ms per 1,000 loci end to end: SM reads over PCIe do not reach the copy engine's rate),
    // kept as a checked alternative for hosts where the copy engines are the contended resource.
    const uint8_t *d_host_view = nullptr;
    if (getenv("MPRG_ZEROCOPY")) {
        cudaPointerAttributes attr;
        if (cudaPointerGetAttributes(&attr, h_ascii) == cudaSuccess && attr.type == cudaMemoryTypeHost &&
            attr.devicePointer != nullptr)
            d_host_view = static_cast<const uint8_t *>(attr.devicePointer);
        else
            cudaGetLastError();
    }
    // One range at a time on the PCIe link: concurrent copies of several workers would share the
    // bandwidth and finish together; in turn, the first range is being built while the next is copied.
    std::unique_lock<std::mutex> link(link_mutex(ctx->device));
    if (d_host_view) {
        for (int i = 0; i < n; ++i) aoff[i] = h_offsets[l0 + i];  // offsets into the caller's buffer
    } else {
        // loci that are contiguous in the caller's buffer go in one copy
        bool contiguous = true;
        for (int i = 0; i < n; ++i) contiguous &= (h_offsets[l0 + i] == h_offsets[l0] + aoff[i]);
        if (contiguous) {
            MPRG_CUDA(ctx, mprg::copy_h2d(ctx, d + o_ascii, h_ascii + h_offsets[l0], (size_t)ascii_total, s));
        } else {
            for (int i = 0; i < n; ++i) {
                const size_t nbytes = (size_t)b->n_rows[l0 + i] * b->n_cols[l0 + i];
                if (!nbytes) continue;
                MPRG_CUDA(ctx, mprg::copy_h2d(ctx, d + o_ascii + aoff[i], h_ascii + h_offsets[l0 + i], nbytes, s));
            }
        }
        MPRG_CUDA(ctx, cudaStreamSynchronize(s));
        link.unlock();
    }
    MPRG_CUDA(ctx, mprg::copy_h2d(ctx, d + o_aoff, aoff.data(), sizeof(long long) * n, s));
    MPRG_CUDA(ctx, mprg::copy_h2d(ctx, d + o_rp, row_prefix.data(), sizeof(long long) * (n + 1), s));
    MPRG_CUDA(ctx, mprg::copy_h2d(ctx, d + o_base, b->base.data() + l0, sizeof(long long) * n, s));
    MPRG_CUDA(ctx, mprg::copy_h2d(ctx, d + o_nc, b->n_cols.data() + l0, sizeof(int) * n, s));
    MPRG_CUDA(ctx, mprg::copy_h2d(ctx, d + o_st, b->stride.data() + l0, sizeof(int) * n, s));
    MPRG_CUDA(ctx, cudaMemsetAsync(d + o_fl, 0, sizeof(int) * n, s));
    const int warps_per_block = 8;
    const long long blocks = (total_rows + warps_per_block - 1) / warps_per_block;
    if (d_host_view) {
        pack_rows_hostmem_kernel<<<(unsigned)blocks, warps_per_block * 32, 0, s>>>(
            d_host_view, (const long long *)(d + o_aoff), (const long long *)(d + o_rp), (const int *)(d + o_nc),
            (const long long *)(d + o_base), (const int *)(d + o_st), n, total_rows, b->d_packed, (int *)(d + o_fl));
        ctx->h2d_bytes += ascii_total;  // read over PCIe by the kernel
    } else {
        pack_rows_kernel<<<(unsigned)blocks, warps_per_block * 32, 0, s>>>(
            d + o_ascii, (const long long *)(d + o_aoff), (const long long *)(d + o_rp), (const int *)(d + o_nc),
            (const long long *)(d + o_base), (const int *)(d + o_st), n, total_rows, b->d_packed, (int *)(d + o_fl));
    }
    ctx->launches++;
    MPRG_CUDA(ctx, cudaGetLastError());
    MPRG_CUDA(ctx, mprg::copy_d2h(ctx, b->flags.data() + l0, d + o_fl, sizeof(int) * n, s));
    MPRG_CUDA(ctx, cudaStreamSynchronize(s));
    if (link.owns_lock()) link.unlock();
    for (int l = l0; l < l1; ++l)
        if (b->flags[l] & 2) b->any_n = true;
    return MPRG_OK;
}

// Loci [l0, l1) from host rows in the packed layout: one copy per run of loci that are contiguous in the
// caller's buffer, straight into the batch arena (the loader lays them out exactly like the arena).
int batch_upload_range_packed(mprg_ctx *ctx, mprg_batch *b, const uint8_t *h_packed, const int64_t *h_offsets,
                              const int32_t *h_flags, int l0, int l1) {
    if (l1 <= l0) return MPRG_OK;
    cudaSetDevice(ctx->device);
    cudaStream_t s = ctx->stream;
    std::unique_lock<std::mutex> link(link_mutex(ctx->device));  // one range at a time on the PCIe link
    int run0 = l0;
    for (int l = l0; l < l1; ++l) {
        const long long bytes_l = (long long)b->stride[l] * b->n_rows[l];
        const bool last = l + 1 == l1;
        const bool joins = !last && h_offsets[l + 1] == h_offsets[l] + bytes_l && b->base[l + 1] == b->base[l] + bytes_l;
        if (joins) continue;
        const long long nbytes = b->base[l] + bytes_l - b->base[run0];
        if (nbytes > 0)
            MPRG_CUDA(ctx, mprg::copy_h2d(ctx, b->d_packed + b->base[run0], h_packed + h_offsets[run0], (size_t)nbytes, s));
        run0 = l + 1;
    }
    MPRG_CUDA(ctx, mprg::wait_stream(ctx, s));
    link.unlock();
    for (int l = l0; l < l1; ++l) {
        b->flags[l] = h_flags ? h_flags[l] : 0;
        if (b->flags[l] & 2) b->any_n = true;
    }
    return MPRG_OK;
}

}  // namespace mprg

extern "C" int mprg_batch_upload(mprg_ctx *ctx, const uint8_t *h_ascii, const int64_t *h_offsets,
                                 const int32_t *n_rows, const int32_t *n_cols, int32_t n_loci,
                                 mprg_batch **out) {
    if (!ctx || !out || n_loci < 0 || (n_loci > 0 && (!h_ascii || !h_offsets || !n_rows || !n_cols)))
        return MPRG_E_BAD_ARG;
    mprg_batch *b = nullptr;
    int rc = mprg::batch_prepare(ctx, n_rows, n_cols, n_loci, &b);
    if (rc != MPRG_OK) return rc;
    rc = mprg::batch_upload_range(ctx, b, h_ascii, h_offsets, 0, n_loci);
    if (rc != MPRG_OK) {
        cudaFree(b->d_packed);
        delete b;
        *out = nullptr;
        return rc;
    }
    *out = b;
    return MPRG_OK;
}

extern "C" void mprg_batch_free(mprg_ctx *ctx, mprg_batch *batch) {
    if (!batch) return;
    if (ctx) cudaSetDevice(ctx->device);
    if (batch->d_packed && ctx && batch->packed_capacity <= ((size_t)8 << 30)) {
        // every launch that read the arena has been waited for (mprg_build returns host results)
        std::lock_guard<std::mutex> lock(ctx->arena_mutex);
        if (ctx->idle_arenas.size() < 2) {
            ctx->idle_arenas.emplace_back(batch->d_packed, batch->packed_capacity);
            batch->d_packed = nullptr;
        }
    }
    if (batch->d_packed) cudaFree(batch->d_packed);
    delete batch;
}

extern "C" int mprg_batch_flags(mprg_ctx *ctx, const mprg_batch *batch, int32_t *h_flags) {
    if (!ctx || !batch || !h_flags) return MPRG_E_BAD_ARG;
    for (int l = 0; l < batch->n_loci; ++l) h_flags[l] = batch->flags[l];
    return MPRG_OK;
}

extern "C" int mprg_batch_download_packed(mprg_ctx *ctx, const mprg_batch *batch, int32_t locus,
                                          uint8_t *h_out, int64_t capacity, int32_t *stride) {
    if (!ctx || !batch || locus < 0 || locus >= batch->n_loci) return MPRG_E_BAD_ARG;
    cudaSetDevice(ctx->device);
    if (stride) *stride = batch->stride[locus];
    const long long nbytes = (long long)batch->stride[locus] * batch->n_rows[locus];
    if (!h_out) return MPRG_OK;
    if (capacity < nbytes) MPRG_FAIL(ctx, MPRG_E_BAD_ARG, "download buffer too small");
    if (nbytes) {
        MPRG_CUDA(ctx, mprg::copy_d2h(ctx, h_out, batch->d_packed + batch->base[locus], (size_t)nbytes, ctx->stream));
        MPRG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    return MPRG_OK;
}
