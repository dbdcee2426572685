// C-ABI entry points for the individual kernels (parity tests and the Python mirrors of the
// reference's functions call these); the whole-path entry point mprg_build lives in engine.cu.
#include <algorithm>
#include <cstring>

#include "common.cuh"
#include "kernels.cuh"

using namespace mprg;

extern "C" int mprg_scan_tasks(mprg_ctx *ctx, const mprg_batch *batch, const mprg_task *h_tasks,
                               int32_t n_tasks, const int32_t *h_rows, int64_t n_row_entries,
                               const int64_t *h_col_offsets, uint8_t *h_consensus,
                               int32_t *h_gap_reach) {
    if (!ctx || !batch || n_tasks < 0 || (n_tasks > 0 && (!h_tasks || !h_col_offsets)))
        return MPRG_E_BAD_ARG;
    Level lv;
    int rc = level_run(ctx, batch, h_tasks, n_tasks, h_rows, n_row_entries, 1, false, lv);
    if (rc != MPRG_OK) return rc;
    if (n_tasks == 0) return MPRG_OK;
    std::vector<uint8_t> cls((size_t)lv.total_cols);
    std::vector<int> reach((size_t)lv.total_cols);
    MPRG_CUDA(ctx, mprg::copy_d2h(ctx, cls.data(), ctx->d_cls.p, cls.size(), ctx->stream));
    MPRG_CUDA(ctx, mprg::copy_d2h(ctx, reach.data(), ctx->d_reach.p, sizeof(int) * reach.size(), ctx->stream));
    MPRG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    account_scan(ctx, lv);
    for (int i = 0; i < n_tasks; ++i) {
        const DTask &t = lv.tasks[i];
        const int n = t.c1 - t.c0, shift = t.c0 & 31;
        if (h_consensus && n) memcpy(h_consensus + h_col_offsets[i], cls.data() + t.col_off + shift, n);
        if (h_gap_reach && n)
            memcpy(h_gap_reach + h_col_offsets[i], reach.data() + t.col_off + shift, sizeof(int) * n);
    }
    return MPRG_OK;
}

extern "C" int mprg_partition_tasks(mprg_ctx *ctx, const mprg_batch *batch, const mprg_task *h_tasks,
                                    int32_t n_tasks, const int32_t *h_rows, int64_t n_row_entries,
                                    int32_t min_match_length, const int64_t *h_iv_offsets,
                                    mprg_interval *h_intervals, int32_t *h_iv_counts) {
    if (!ctx || !batch || n_tasks < 0 ||
        (n_tasks > 0 && (!h_tasks || !h_iv_offsets || !h_intervals || !h_iv_counts)))
        return MPRG_E_BAD_ARG;
    Level lv;
    int rc = level_run(ctx, batch, h_tasks, n_tasks, h_rows, n_row_entries, min_match_length, true, lv);
    if (rc != MPRG_OK) return rc;
    if (n_tasks == 0) return MPRG_OK;
    std::vector<DInterval> iv((size_t)lv.total_iv);
    std::vector<int> cnt(n_tasks + 1);
    MPRG_CUDA(ctx, mprg::copy_d2h(ctx, iv.data(), ctx->d_iv.p, sizeof(DInterval) * iv.size(), ctx->stream));
    MPRG_CUDA(ctx, mprg::copy_d2h(ctx, cnt.data(), ctx->d_ivcnt.p, sizeof(int) * cnt.size(), ctx->stream));
    MPRG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    account_scan(ctx, lv);
    if (cnt[n_tasks]) MPRG_FAIL(ctx, MPRG_E_PARTITION, "Failed interval partitioning");
    for (int i = 0; i < n_tasks; ++i) {
        h_iv_counts[i] = cnt[i];
        for (int k = 0; k < cnt[i]; ++k) {
            const DInterval &d = iv[lv.tasks[i].iv_off + k];
            h_intervals[h_iv_offsets[i] + k] = mprg_interval{d.start, d.stop, d.type};
        }
    }
    return MPRG_OK;
}

extern "C" int mprg_partition_consensus(mprg_ctx *ctx, const uint8_t *h_consensus,
                                        const int32_t *h_gap_reach, int32_t n_cols,
                                        int32_t min_match_length, mprg_interval *h_intervals,
                                        int32_t capacity, int32_t *h_count) {
    if (!ctx || n_cols < 0 || (n_cols > 0 && !h_consensus) || !h_intervals || !h_count)
        return MPRG_E_BAD_ARG;
    cudaSetDevice(ctx->device);
    cudaStream_t s = ctx->stream;
    const int n = n_cols;
    const size_t ivcap = std::max(n, 1);
    MPRG_CUDA(ctx, ctx->d_cls.reserve(std::max(n, 1)));
    MPRG_CUDA(ctx, ctx->d_reach.reserve(sizeof(int) * std::max(n, 1)));
    MPRG_CUDA(ctx, ctx->d_misc.reserve(sizeof(uint32_t) * (n / 32 + 2)));
    MPRG_CUDA(ctx, ctx->d_iv.reserve(sizeof(DInterval) * ivcap));
    MPRG_CUDA(ctx, ctx->d_ivcnt.reserve(sizeof(int) * 2));
    if (n) MPRG_CUDA(ctx, mprg::copy_h2d(ctx, ctx->d_cls.p, h_consensus, n, s));
    if (n && h_gap_reach)
        MPRG_CUDA(ctx, mprg::copy_h2d(ctx, ctx->d_reach.p, h_gap_reach, sizeof(int) * n, s));
    MPRG_CUDA(ctx, cudaMemsetAsync(ctx->d_ivcnt.p, 0, sizeof(int) * 2, s));
    int *cnt = ctx->d_ivcnt.as<int>();
    MPRG_CUDA(ctx, launch_partition_consensus(s, ctx->d_cls.as<uint8_t>(),
                                              h_gap_reach ? ctx->d_reach.as<int>() : nullptr, n,
                                              min_match_length, ctx->d_misc.as<uint32_t>(),
                                              ctx->d_iv.as<DInterval>(), cnt, cnt + 1));
    ctx->launches++;
    int hc[2] = {0, 0};
    std::vector<DInterval> iv(ivcap);
    MPRG_CUDA(ctx, mprg::copy_d2h(ctx, hc, cnt, sizeof(int) * 2, s));
    MPRG_CUDA(ctx, mprg::copy_d2h(ctx, iv.data(), ctx->d_iv.p, sizeof(DInterval) * ivcap, s));
    MPRG_CUDA(ctx, cudaStreamSynchronize(s));
    if (hc[1]) MPRG_FAIL(ctx, MPRG_E_PARTITION, "Failed interval partitioning");
    if (hc[0] > capacity) MPRG_FAIL(ctx, MPRG_E_BAD_ARG, "interval capacity too small");
    *h_count = hc[0];
    for (int k = 0; k < hc[0]; ++k) h_intervals[k] = mprg_interval{iv[k].start, iv[k].stop, iv[k].type};
    return MPRG_OK;
}

extern "C" int mprg_copy_stats(mprg_ctx *ctx, int64_t *h2d_bytes, int64_t *d2h_bytes, int reset) {
    if (!ctx) return MPRG_E_BAD_ARG;
    if (h2d_bytes) *h2d_bytes = ctx->h2d_bytes;
    if (d2h_bytes) *d2h_bytes = ctx->d2h_bytes;
    if (reset) ctx->h2d_bytes = ctx->d2h_bytes = 0;
    return MPRG_OK;
}

extern "C" int mprg_kmeans_stats(mprg_ctx *ctx, double *ms, int64_t *launches, int64_t *problems, int reset) {
    if (!ctx) return MPRG_E_BAD_ARG;
    double t = ctx->km_ms;
    long long l = ctx->km_launches, p = ctx->km_problems;
    for (mprg_ctx *w : ctx->workers) {
        t += w->km_ms;
        l += w->km_launches;
        p += w->km_problems;
    }
    if (ms) *ms = t;
    if (launches) *launches = l;
    if (problems) *problems = p;
    if (reset) {
        ctx->km_ms = 0;
        ctx->km_launches = ctx->km_problems = 0;
        for (mprg_ctx *w : ctx->workers) {
            w->km_ms = 0;
            w->km_launches = w->km_problems = 0;
        }
    }
    return MPRG_OK;
}

extern "C" int mprg_path_counts(mprg_ctx *ctx, int64_t *out, int reset) {
    if (!ctx || !out) return MPRG_E_BAD_ARG;
    for (int k = 0; k < MPRG_PATH_COUNT; ++k) {
        out[k] = ctx->path_counts[k];
        for (mprg_ctx *w : ctx->workers) out[k] += w->path_counts[k];
    }
    if (reset) {
        for (int k = 0; k < MPRG_PATH_COUNT; ++k) {
            ctx->path_counts[k] = 0;
            for (mprg_ctx *w : ctx->workers) w->path_counts[k] = 0;
        }
    }
    return MPRG_OK;
}

extern "C" int mprg_timer(mprg_ctx *ctx, int op, double *ms) {
    if (!ctx) return MPRG_E_BAD_ARG;
    cudaSetDevice(ctx->device);
    if (op == 0) {
        MPRG_CUDA(ctx, cudaEventRecord(ctx->ev_t0, ctx->stream));
    } else {
        MPRG_CUDA(ctx, cudaEventRecord(ctx->ev_t1, ctx->stream));
        MPRG_CUDA(ctx, cudaEventSynchronize(ctx->ev_t1));
        float f = 0;
        MPRG_CUDA(ctx, cudaEventElapsedTime(&f, ctx->ev_t0, ctx->ev_t1));
        if (ms) *ms = f;
    }
    return MPRG_OK;
}

namespace mprg {
// bare streaming read with the scan kernel's tile shape (measurement aid, mprg_read_yardstick)
__global__ void __launch_bounds__(128)
read_yardstick_kernel(const uint4 *__restrict__ p, long long n_tiles, long long n_vec, unsigned *sink) {
    const int lane = threadIdx.x & 31;
    const long long tile = (long long)blockIdx.x * 4 + (threadIdx.x >> 5);
    if (tile >= n_tiles) return;
    const long long first = tile * 16 * 32 + lane;  // 16 rows of 32 vectors
    uint32_t a = 0;
    uint4 v[8];
#pragma unroll
    for (int half = 0; half < 2; ++half) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const long long i = first + (half * 8 + u) * 32;
            v[u] = make_uint4(0, 0, 0, 0);
            if (i < n_vec)
                asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
                             : "=r"(v[u].x), "=r"(v[u].y), "=r"(v[u].z), "=r"(v[u].w) : "l"(p + i));
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) a |= v[u].x | v[u].y | v[u].z | v[u].w;
    }
    if (a == 0x5A5A5A5Au) atomicOr(sink, a);  // never true for packed symbols in practice; keeps the loads
}
}  // namespace mprg

extern "C" int mprg_read_yardstick(mprg_ctx *ctx, const mprg_batch *batch, double *bytes, double *ms) {
    if (!ctx || !batch || !batch->d_packed) return MPRG_E_BAD_ARG;
    cudaSetDevice(ctx->device);
    const long long n_vec = batch->packed_bytes / 16;
    const long long n_tiles = (n_vec + 511) / 512;
    if (n_tiles <= 0) return MPRG_E_BAD_ARG;
    MPRG_CUDA(ctx, ctx->d_misc.reserve(64));
    MPRG_CUDA(ctx, cudaEventRecord(ctx->ev_t0, ctx->stream));
    mprg::read_yardstick_kernel<<<(unsigned)((n_tiles + 3) / 4), 128, 0, ctx->stream>>>(
        reinterpret_cast<const uint4 *>(batch->d_packed), n_tiles, n_vec, ctx->d_misc.as<unsigned>());
    MPRG_CUDA(ctx, cudaGetLastError());
    MPRG_CUDA(ctx, cudaEventRecord(ctx->ev_t1, ctx->stream));
    MPRG_CUDA(ctx, cudaEventSynchronize(ctx->ev_t1));
    float f = 0;
    MPRG_CUDA(ctx, cudaEventElapsedTime(&f, ctx->ev_t0, ctx->ev_t1));
    ctx->launches++;
    if (ms) *ms = f;
    if (bytes) *bytes = (double)n_vec * 16.0;
    return MPRG_OK;
}

extern "C" int mprg_scan_log(mprg_ctx *ctx, double *bytes, double *ms, int32_t capacity, int32_t *n,
                             int reset) {
    if (!ctx || !n) return MPRG_E_BAD_ARG;
    const int have = (int)ctx->scan_log_bytes.size();
    const int take = have < capacity ? have : capacity;
    for (int i = 0; i < take; ++i) {
        if (bytes) bytes[i] = ctx->scan_log_bytes[have - take + i];
        if (ms) ms[i] = ctx->scan_log_ms[have - take + i];
    }
    *n = take;
    if (reset) {
        ctx->scan_log_bytes.clear();
        ctx->scan_log_ms.clear();
    }
    return MPRG_OK;
}
