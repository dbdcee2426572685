// Host side of one recursion level: task table, scan units, launches.
#include <algorithm>
#include <cstdlib>

#include "common.cuh"
#include "kernels.cuh"

namespace mprg {

constexpr int TILE_CHUNKS = 32;      // widest tile: one 512-byte row segment per warp load
constexpr int TILE_ITER_QUANTUM = 4; // warp iterations one trip of the scan kernel consumes
constexpr int MIN_TILE_ITERS = 12, MAX_TILE_ITERS = 32;  // sweeps in DESIGN.md section 7
constexpr int SCAN_RESIDENT_WARPS_PER_SM = 32;  // 64 registers per thread (__launch_bounds__(128, 8))

static inline int pow2_ceil(int x) {
    int p = 1;
    while (p < x) p <<= 1;
    return p;
}

// call after the stream has been synchronised: device time of the level's scan launch
void account_scan(mprg_ctx *ctx, const Level &lv) {
    float ms = 0;
    if (!lv.units.empty() && cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1) == cudaSuccess) {
        ctx->scan_ms += ms;
        ctx->scan_bytes += lv.algo_bytes;
        ctx->scan_launches += 1;
        if (ctx->scan_log_bytes.size() < (1u << 20)) {
            ctx->scan_log_bytes.push_back(lv.algo_bytes);
            ctx->scan_log_ms.push_back(ms);
        }
    }
}

int level_run(mprg_ctx *ctx, const mprg_batch *batch, const mprg_task *h_tasks, int n_tasks,
              const int32_t *h_rows, long long n_row_entries, int mml, bool do_partition, Level &lv) {
    cudaSetDevice(ctx->device);
    cudaStream_t s = ctx->stream;
    lv = Level();
    lv.n_tasks = n_tasks;
    lv.tasks.resize(n_tasks);
    lv.has_n = batch->any_n;
    long long col_off = 0, iv_off = 0;
    // A warp streams the rows of its tile with a bounded number of loads in flight, so a tile's
    // duration grows with its height whatever the load of the machine: cut the level into enough tiles
    // to fill every SM for a few waves (but never below a few trips' worth of rows).  Heights are
    // counted in warp iterations: a tile of W (power of two) chunks takes 32 / W rows per iteration.
    long long level_iters = 0;
    for (int i = 0; i < n_tasks; ++i) {
        const mprg_task &ht = h_tasks[i];
        if (ht.n_rows <= 0 || ht.c1 < ht.c0) continue;
        const int nch = std::max(((ht.c1 + 31) >> 5) - (ht.c0 >> 5), 1);
        const int rem = nch % TILE_CHUNKS;
        level_iters += (long long)(nch / TILE_CHUNKS) * ht.n_rows;
        if (rem) level_iters += (ht.n_rows + (32 / pow2_ceil(rem)) - 1) / (32 / pow2_ceil(rem));
    }
    // 3.5 waves of resident warps: 16 iterations for the 1,000 x (200 x 1,000) root level (sweep: 12: 3.00, 16: 3.18,
    // 20: 3.06, 24: 2.94 TB/s), 32 for launches 8x that size (12: 4.2, 20: 5.1, 28-32: 5.3 TB/s)
    const long long target_tiles = (long long)std::max(ctx->sm_count, 1) * SCAN_RESIDENT_WARPS_PER_SM * 7 / 2;
    int tile_iters = (int)((level_iters + target_tiles - 1) / target_tiles);
    tile_iters = ((tile_iters + TILE_ITER_QUANTUM - 1) / TILE_ITER_QUANTUM) * TILE_ITER_QUANTUM;
    tile_iters = std::min(std::max(tile_iters, MIN_TILE_ITERS), MAX_TILE_ITERS);
    if (const char *env = getenv("MPRG_TILE_ITERS")) {  // tuning knob for profiling runs
        const int v = atoi(env);
        if (v >= TILE_ITER_QUANTUM && v <= 1024) tile_iters = v;
    }
    for (int i = 0; i < n_tasks; ++i) {
        const mprg_task &ht = h_tasks[i];
        if (ht.locus < 0 || ht.locus >= batch->n_loci) MPRG_FAIL(ctx, MPRG_E_BAD_ARG, "task locus out of range");
        const int R = batch->n_rows[ht.locus], C = batch->n_cols[ht.locus];
        if (ht.c0 < 0 || ht.c1 < ht.c0 || ht.c1 > C || ht.n_rows < 0 ||
            (ht.rows_off < 0 && ht.n_rows > R) ||
            (ht.rows_off >= 0 && (long long)ht.rows_off + ht.n_rows > n_row_entries))
            MPRG_FAIL(ctx, MPRG_E_BAD_ARG, "task window / rows out of range");
        DTask &t = lv.tasks[i];
        t.base = batch->base[ht.locus];
        t.stride = batch->stride[ht.locus];
        t.rows_off = ht.rows_off;
        t.n_rows = ht.n_rows;
        t.c0 = ht.c0;
        t.c1 = ht.c1;
        const int a0 = ht.c0 & ~31, a1 = (ht.c1 + 31) & ~31;
        const int aligned = std::max(a1 - a0, 32);
        if (col_off + aligned > 0x7fffffffLL) MPRG_FAIL(ctx, MPRG_E_BAD_ARG, "level too wide");
        t.col_off = (int)col_off;
        t.iv_off = (int)iv_off;
        t.flags = batch->any_n ? 1 : 0;
        col_off += aligned;
        iv_off += std::max(ht.c1 - ht.c0, 1);
        const double r = ht.n_rows, c = ht.c1 - ht.c0;
        lv.algo_bytes += r * c / 2 + (ht.rows_off >= 0 ? 4.0 * r : 0.0) + 5.0 * c;
        // tiles: full 32-chunk column strips, then the remainder strip; each strip cut into row ranges
        // of tile_iters warp iterations (a short remainder joins the last tile)
        const int ch0 = ht.c0 >> 5, ch1 = std::max((ht.c1 + 31) >> 5, ch0 + 1);
        for (int cb = ch0; cb < ch1; cb += TILE_CHUNKS) {
            const int bn = std::min(TILE_CHUNKS, ch1 - cb);
            const int tile_rows = tile_iters * (32 / pow2_ceil(bn));
            for (int rb = 0; rb < ht.n_rows;) {
                int cnt = std::min(tile_rows, ht.n_rows - rb);
                const int left = ht.n_rows - rb - cnt;
                if (left > 0 && left < tile_rows / 2) cnt += left;
                lv.units.push_back(ScanUnit{t.base, t.stride, t.rows_off >= 0 ? t.rows_off + rb : -1, rb, cnt, t.c0,
                                        t.c1, t.col_off, cb, bn, (batch->flags[ht.locus] & 8) ? 1 : 0});
                rb += cnt;
            }
        }
    }
    lv.total_cols = col_off;
    lv.total_iv = iv_off;
    if (n_tasks == 0) return MPRG_OK;
    // validate row indices are inside their locus (cheap, protects the kernels)
    for (int i = 0; i < n_tasks; ++i) {
        const mprg_task &ht = h_tasks[i];
        if (ht.rows_off < 0) continue;
        const int R = batch->n_rows[ht.locus];
        for (int k = 0; k < ht.n_rows; ++k) {
            const int r = h_rows[ht.rows_off + k];
            if (r < 0 || r >= R) MPRG_FAIL(ctx, MPRG_E_BAD_ARG, "row index out of range");
        }
    }

    const size_t n_units = lv.units.size();
    MPRG_CUDA(ctx, ctx->d_tasks.reserve(sizeof(DTask) * n_tasks));
    MPRG_CUDA(ctx, ctx->d_units.reserve(sizeof(ScanUnit) * std::max<size_t>(n_units, 1)));
    MPRG_CUDA(ctx, ctx->d_rows.reserve(sizeof(int) * std::max<long long>(n_row_entries, 1)));
    // colwords: colOR | colNOR (total_cols/8 words each) ; colB: total_cols unsigned
    const size_t words = (size_t)lv.total_cols / 8;
    MPRG_CUDA(ctx, ctx->d_colwords.reserve(sizeof(uint32_t) * 2 * words));
    MPRG_CUDA(ctx, ctx->d_colB.reserve(sizeof(unsigned) * lv.total_cols));
    MPRG_CUDA(ctx, ctx->d_cls.reserve((size_t)lv.total_cols));
    MPRG_CUDA(ctx, ctx->d_reach.reserve(sizeof(int) * lv.total_cols));
    MPRG_CUDA(ctx, ctx->d_misc.reserve(sizeof(uint32_t) * (lv.total_cols / 32) + 64));
    MPRG_CUDA(ctx, ctx->d_iv.reserve(sizeof(DInterval) * lv.total_iv));
    MPRG_CUDA(ctx, ctx->d_ivcnt.reserve(sizeof(int) * (n_tasks + 1)));

    MPRG_CUDA(ctx, mprg::copy_h2d(ctx, ctx->d_tasks.p, lv.tasks.data(), sizeof(DTask) * n_tasks, s));
    if (n_units)
        MPRG_CUDA(ctx, mprg::copy_h2d(ctx, ctx->d_units.p, lv.units.data(), sizeof(ScanUnit) * n_units, s));
    if (n_row_entries > 0)
        MPRG_CUDA(ctx, mprg::copy_h2d(ctx, ctx->d_rows.p, h_rows, sizeof(int) * n_row_entries, s));
    MPRG_CUDA(ctx, cudaMemsetAsync(ctx->d_colwords.p, 0, sizeof(uint32_t) * 2 * words, s));
    MPRG_CUDA(ctx, cudaMemsetAsync(ctx->d_colB.p, 0, sizeof(unsigned) * lv.total_cols, s));
    // last int of d_ivcnt is the partition error flag
    MPRG_CUDA(ctx, cudaMemsetAsync(ctx->d_ivcnt.p, 0, sizeof(int) * (n_tasks + 1), s));

    uint32_t *colOR = ctx->d_colwords.as<uint32_t>();
    uint32_t *colNOR = colOR + words;
    MPRG_CUDA(ctx, cudaEventRecord(ctx->ev0, s));
    MPRG_CUDA(ctx, launch_scan(s, lv.has_n, batch->d_packed, ctx->d_units.as<ScanUnit>(), (int)n_units,
                               ctx->d_rows.as<int>(), colOR, colNOR, ctx->d_colB.as<unsigned>()));
    MPRG_CUDA(ctx, cudaEventRecord(ctx->ev1, s));
    ctx->launches += n_units ? 1 : 0;
    MPRG_CUDA(ctx, launch_classify(s, ctx->d_tasks.as<DTask>(), n_tasks, colOR, colNOR,
                                   ctx->d_colB.as<unsigned>(), ctx->d_cls.as<uint8_t>(),
                                   ctx->d_reach.as<int>(), ctx->d_misc.as<uint32_t>()));
    ctx->launches++;
    if (do_partition) {
        int *cnt = ctx->d_ivcnt.as<int>();
        MPRG_CUDA(ctx, launch_partition(s, ctx->d_tasks.as<DTask>(), n_tasks, ctx->d_misc.as<uint32_t>(),
                                        ctx->d_reach.as<int>(), mml, ctx->d_iv.as<DInterval>(), cnt,
                                        cnt + n_tasks));
        MPRG_CUDA(ctx, launch_demote(s, batch->d_packed, ctx->d_tasks.as<DTask>(), n_tasks,
                                     ctx->d_rows.as<int>(), ctx->d_iv.as<DInterval>(), cnt));
        ctx->launches += 2;
    }
    return MPRG_OK;
}

}  // namespace mprg
