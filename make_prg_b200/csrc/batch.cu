// Context management and the loader -> HBM step: ASCII MSAs are copied to the device and packed to
// 4 bits per symbol there (two columns per byte, rows padded to 16 bytes).
#include <algorithm>
#include <cstdlib>
#include <cstring>
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
        const unsigned hc = std::thread::hardware_concurrency();
        ctx->n_workers = (int)std::max(1u, std::min(4u, hc ? hc : 1u));
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
    ctx->h_a.release();
    ctx->h_b.release();
    ctx->h_c.release();
    ctx->h_d.release();
    if (ctx->ev0) cudaEventDestroy(ctx->ev0);
    if (ctx->ev1) cudaEventDestroy(ctx->ev1);
    if (ctx->ev_t0) cudaEventDestroy(ctx->ev_t0);
    if (ctx->ev_t1) cudaEventDestroy(ctx->ev_t1);
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

extern "C" int mprg_batch_upload(mprg_ctx *ctx, const uint8_t *h_ascii, const int64_t *h_offsets,
                                 const int32_t *n_rows, const int32_t *n_cols, int32_t n_loci,
                                 mprg_batch **out) {
    if (!ctx || !out || n_loci < 0 || (n_loci > 0 && (!h_ascii || !h_offsets || !n_rows || !n_cols)))
        return MPRG_E_BAD_ARG;
    *out = nullptr;
    cudaSetDevice(ctx->device);
    mprg_batch *b = new mprg_batch();
    b->n_loci = n_loci;
    b->n_rows.assign(n_rows, n_rows + n_loci);
    b->n_cols.assign(n_cols, n_cols + n_loci);
    b->stride.resize(n_loci);
    b->base.resize(n_loci);
    b->flags.assign(n_loci, 0);
    long long packed = 0, ascii_total = 0, total_rows = 0;
    std::vector<long long> row_prefix(n_loci + 1, 0), aoff(n_loci);
    for (int l = 0; l < n_loci; ++l) {
        if (n_rows[l] < 0 || n_cols[l] < 0) {
            delete b;
            MPRG_FAIL(ctx, MPRG_E_BAD_ARG, "negative MSA dimension");
        }
        b->stride[l] = ((n_cols[l] + COLS_PER_CHUNK - 1) / COLS_PER_CHUNK) * CHUNK_BYTES;
        b->base[l] = packed;
        packed += (long long)b->stride[l] * n_rows[l];
        aoff[l] = ascii_total;
        ascii_total += (long long)n_rows[l] * n_cols[l];
        row_prefix[l] = total_rows;
        total_rows += n_rows[l];
    }
    row_prefix[n_loci] = total_rows;
    b->packed_bytes = packed;
    auto fail = [&](cudaError_t e, const char *what) {
        ctx->err = std::string(what) + ": " + cudaGetErrorString(e);
        if (b->d_packed) cudaFree(b->d_packed);
        delete b;
        return MPRG_E_CUDA;
    };
    cudaError_t e;
    if ((e = cudaMalloc(&b->d_packed, (size_t)packed + 16)) != cudaSuccess) return fail(e, "cudaMalloc packed");
    if (total_rows > 0 && ascii_total > 0) {
        // stage: ascii | ascii_off | row_prefix | base | n_cols | stride | flags
        size_t o_ascii = 0;
        size_t o_aoff = (ascii_total + 15) & ~15LL;
        size_t o_rp = o_aoff + sizeof(long long) * n_loci;
        size_t o_base = o_rp + sizeof(long long) * (n_loci + 1);
        size_t o_nc = o_base + sizeof(long long) * n_loci;
        size_t o_st = o_nc + sizeof(int) * n_loci;
        size_t o_fl = o_st + sizeof(int) * n_loci;
        size_t total = o_fl + sizeof(int) * n_loci;
        if ((e = ctx->d_stage.reserve(total)) != cudaSuccess) return fail(e, "reserve stage");
        uint8_t *d = ctx->d_stage.as<uint8_t>();
        cudaStream_t s = ctx->stream;
        // the ASCII copy goes straight from the caller's buffer locus by locus when the loci are
        // contiguous in the caller's buffer this is one copy
        bool contiguous = true;
        for (int l = 0; l < n_loci; ++l) contiguous &= (h_offsets[l] == h_offsets[0] + aoff[l]);
        if (contiguous) {
            if ((e = mprg::copy_h2d(ctx, d + o_ascii, h_ascii + h_offsets[0], (size_t)ascii_total, s)) != cudaSuccess)
                return fail(e, "H2D ascii");
        } else {
            for (int l = 0; l < n_loci; ++l) {
                const size_t nbytes = (size_t)n_rows[l] * n_cols[l];
                if (!nbytes) continue;
                if ((e = mprg::copy_h2d(ctx, d + o_ascii + aoff[l], h_ascii + h_offsets[l], nbytes, s)) != cudaSuccess)
                    return fail(e, "H2D ascii");
            }
        }
        mprg::copy_h2d(ctx, d + o_aoff, aoff.data(), sizeof(long long) * n_loci, s);
        mprg::copy_h2d(ctx, d + o_rp, row_prefix.data(), sizeof(long long) * (n_loci + 1), s);
        mprg::copy_h2d(ctx, d + o_base, b->base.data(), sizeof(long long) * n_loci, s);
        mprg::copy_h2d(ctx, d + o_nc, b->n_cols.data(), sizeof(int) * n_loci, s);
        mprg::copy_h2d(ctx, d + o_st, b->stride.data(), sizeof(int) * n_loci, s);
        cudaMemsetAsync(d + o_fl, 0, sizeof(int) * n_loci, s);
        const int warps_per_block = 8;
        const long long blocks = (total_rows + warps_per_block - 1) / warps_per_block;
        pack_rows_kernel<<<(unsigned)blocks, warps_per_block * 32, 0, s>>>(
            d + o_ascii, (const long long *)(d + o_aoff), (const long long *)(d + o_rp),
            (const int *)(d + o_nc), (const long long *)(d + o_base), (const int *)(d + o_st), n_loci,
            total_rows, b->d_packed, (int *)(d + o_fl));
        ctx->launches++;
        if ((e = cudaGetLastError()) != cudaSuccess) return fail(e, "pack_rows_kernel");
        if ((e = mprg::copy_d2h(ctx, b->flags.data(), d + o_fl, sizeof(int) * n_loci, s)) != cudaSuccess)
            return fail(e, "D2H flags");
        if ((e = cudaStreamSynchronize(s)) != cudaSuccess) return fail(e, "sync upload");
    }
    for (int f : b->flags) b->any_n |= (f & 2) != 0;
    *out = b;
    return MPRG_OK;
}

extern "C" void mprg_batch_free(mprg_ctx *ctx, mprg_batch *batch) {
    if (!batch) return;
    if (ctx) cudaSetDevice(ctx->device);
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
