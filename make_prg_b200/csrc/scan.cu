// Kernel (a): column-wise scan of 4-bit packed sub-alignments.
//
// Replaces, for every task (row subset x column window) of a recursion level at once:
//   get_consensus_from_MSA        make_prg/utils/seq_utils.py:219-239   (column classes)
//   has_empty_sequence            make_prg/utils/seq_utils.py:37-42     (as gap-run reach)
//   all-gap column detection      make_prg/utils/seq_utils.py:201-207
//
// Per row and 16-byte chunk (32 columns) one 128-bit load.  Per column two accumulators over the rows
// of the task, kept nibble-parallel inside 32-bit words (8 columns per word):
//     OR  of the 4-bit symbol codes      -> colOR
//     OR  of the complemented codes      -> colNOR   (== ~AND, so both are OR-reductions, identity 0)
// A column is uniform iff OR == AND, i.e. (colOR ^ colNOR) == 0xF in its nibble; the symbol is then the
// OR nibble.  N (code 11) is a wildcard in the HAS_N variant (it contributes to neither accumulator).
//
// Gap runs: for each row, each maximal run of '-' inside the window contributes
//     B[run start] = max(B[run start], run end + 1)
// (coordinates relative to the chunk-aligned window start).  The partition kernel turns B into
// gap_reach by a prefix maximum: has_empty_sequence([s, e]) == (gap_reach[s] >= e).  Runs crossing a
// 1024-column block are stitched with a per-row carry while the CTA walks the blocks right to left.
//
// Work decomposition: one CTA per ScanUnit (a row range of one task, all of its columns).  Inside a
// warp the 32 lanes cover (32 / nchp) rows x nchp chunks, nchp = chunks in the block rounded to a
// power of two, so narrow windows still use all lanes.  Partial results are merged through shared
// memory, then with atomicOr / atomicMax into the (zero-initialised) per-task column arrays, which
// also merges units that split the rows of a very tall task.
#include "common.cuh"
#include "kernels.cuh"

namespace mprg {

constexpr int SCAN_THREADS = 128;
constexpr int SCAN_WARPS = SCAN_THREADS / 32;
constexpr int SCAN_BLOCK_CHUNKS = 32;                                  // chunks per column block
constexpr int SCAN_BLOCK_COLS = SCAN_BLOCK_CHUNKS * COLS_PER_CHUNK;    // 1024
constexpr int SCAN_UNROLL = 4;

__device__ __forceinline__ uint4 ld_stream(const uint4 *p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}

// non-zero iff some nibble of w equals the gap code
__device__ __forceinline__ uint32_t has_gap_nibble(uint32_t w) {
    const uint32_t t = w ^ 0x44444444u;
    return (t - 0x11111111u) & ~t & 0x88888888u;
}

// 8-bit mask (bit j = column j of the word) of nibbles equal to `pattern` (pattern repeated x8)
__device__ __forceinline__ uint32_t nibble_eq_mask8(uint32_t w, uint32_t pattern) {
    uint32_t t = w ^ pattern;
    t |= t >> 1;
    t |= t >> 2;
    uint32_t y = ~t & 0x11111111u;  // 1 at the low bit of each matching nibble
    y = (y | (y >> 3)) & 0x03030303u;
    y = (y | (y >> 6)) & 0x000F000Fu;
    y = (y | (y >> 12)) & 0xFFu;
    return y;
}

// nibble-wide (0xF) mask of the nibbles equal to `pattern`
__device__ __forceinline__ uint32_t nibble_eq_maskF(uint32_t w, uint32_t pattern) {
    uint32_t t = w ^ pattern;
    t |= t >> 1;
    t |= t >> 2;
    return (~t & 0x11111111u) * 15u;
}

__device__ __forceinline__ uint32_t gap_mask32(const uint4 &v) {
    return nibble_eq_mask8(v.x, 0x44444444u) | (nibble_eq_mask8(v.y, 0x44444444u) << 8) |
           (nibble_eq_mask8(v.z, 0x44444444u) << 16) | (nibble_eq_mask8(v.w, 0x44444444u) << 24);
}

template <bool HAS_N>
__global__ void __launch_bounds__(SCAN_THREADS)
scan_kernel(const uint8_t *__restrict__ packed, const DTask *__restrict__ tasks,
            const ScanUnit *__restrict__ units, const int *__restrict__ rows_arena,
            uint32_t *__restrict__ colOR, uint32_t *__restrict__ colNOR,
            unsigned *__restrict__ colB) {
    __shared__ unsigned B_s[SCAN_BLOCK_COLS];
    __shared__ uint32_t acc_s[SCAN_WARPS][SCAN_BLOCK_CHUNKS][8];
    extern __shared__ int carry_s[];  // one int per row of the unit (only used when nblocks > 1)

    const ScanUnit unit = units[blockIdx.x];
    const DTask t = tasks[unit.task];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int ch0 = t.c0 >> 5;
    const int ch1 = (t.c1 + 31) >> 5;
    const int nch = ch1 - ch0;
    if (nch <= 0 || unit.row_count <= 0) return;
    const int a0 = ch0 << 5;  // chunk-aligned window start (columns)
    const int nblocks = (nch + SCAN_BLOCK_CHUNKS - 1) / SCAN_BLOCK_CHUNKS;
    const int *rows = t.rows_off >= 0 ? rows_arena + t.rows_off + unit.row_begin : nullptr;
    const uint8_t *msa = packed + t.base;

    for (int blk = nblocks - 1; blk >= 0; --blk) {
        const int bch0 = ch0 + blk * SCAN_BLOCK_CHUNKS;
        const int bn = min(SCAN_BLOCK_CHUNKS, ch1 - bch0);
        int lg = 0;
        while ((1 << lg) < bn) ++lg;
        const int nchp = 1 << lg;       // lanes per row
        const int rpw = 32 >> lg;       // rows per warp iteration
        const int lane_chunk = lane & (nchp - 1);
        const int lane_slot = lane >> lg;
        const bool chunk_valid = lane_chunk < bn;
        const int chunk = bch0 + lane_chunk;
        const int colbase = chunk << 5;
        // window mask: bit i set iff c0 <= colbase + i < c1
        uint32_t wmask = 0;
        if (chunk_valid) {
            const int lo = max(t.c0 - colbase, 0);
            const int hi = min(t.c1 - colbase, 32);
            if (hi > lo) wmask = (hi - lo == 32) ? 0xffffffffu : (((1u << (hi - lo)) - 1u) << lo);
        }
        const bool right_block_exists = blk < nblocks - 1;

        for (int i = threadIdx.x; i < SCAN_BLOCK_COLS; i += SCAN_THREADS) B_s[i] = 0;
        __syncthreads();

        uint32_t a_or[4] = {0, 0, 0, 0}, a_nor[4] = {0, 0, 0, 0};
        const int n_iters = (unit.row_count + rpw - 1) / rpw;
        for (int it0 = warp; it0 < n_iters; it0 += SCAN_WARPS * SCAN_UNROLL) {
            uint4 v[SCAN_UNROLL];
            bool valid[SCAN_UNROLL];
            int rl[SCAN_UNROLL];
#pragma unroll
            for (int u = 0; u < SCAN_UNROLL; ++u) {
                const int it = it0 + u * SCAN_WARPS;
                rl[u] = it * rpw + lane_slot;
                valid[u] = chunk_valid && it < n_iters && rl[u] < unit.row_count;
                v[u] = make_uint4(0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu);
                if (valid[u]) {
                    const int row = rows ? rows[rl[u]] : unit.row_begin + rl[u];
                    v[u] = ld_stream(reinterpret_cast<const uint4 *>(
                        msa + (long long)row * t.stride + (long long)chunk * CHUNK_BYTES));
                }
            }
#pragma unroll
            for (int u = 0; u < SCAN_UNROLL; ++u) {
                const int it = it0 + u * SCAN_WARPS;
                if (it >= n_iters) break;  // warp-uniform
                const uint32_t w[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
                if (valid[u]) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        if (HAS_N) {
                            const uint32_t keep = ~nibble_eq_maskF(w[j], 0xBBBBBBBBu);
                            a_or[j] |= w[j] & keep;
                            a_nor[j] |= ~w[j] & keep;
                        } else {
                            a_or[j] |= w[j];
                            a_nor[j] |= ~w[j];
                        }
                    }
                }
                uint32_t hg = 0;
                if (valid[u])
                    hg = has_gap_nibble(w[0]) | has_gap_nibble(w[1]) | has_gap_nibble(w[2]) |
                         has_gap_nibble(w[3]);
                if (!__any_sync(0xffffffffu, hg != 0)) {
                    if (nblocks > 1 && lane_chunk == 0 && rl[u] < unit.row_count) carry_s[rl[u]] = 0;
                    continue;
                }
                // ---- slow path: this warp iteration holds at least one gap ----
                const uint32_t g = valid[u] ? (gap_mask32(v[u]) & wmask) : 0u;
                const bool full = g == 0xffffffffu;
                int val = full ? 32 : (__ffs(~g) - 1);  // gap columns at the start of the chunk
                bool f = full;
                // segmented suffix scan over the nchp lanes of this row: E(j) = lead(j) + (full(j) ? E(j+1) : 0)
                for (int d = 1; d < nchp; d <<= 1) {
                    const int oval = __shfl_down_sync(0xffffffffu, val, d);
                    const int of = __shfl_down_sync(0xffffffffu, (int)f, d);
                    if (lane_chunk + d < nchp) {
                        if (f) val += oval;
                        f = f && (of != 0);
                    }
                }
                int carry_in = 0;
                if (right_block_exists && rl[u] < unit.row_count) carry_in = carry_s[rl[u]];
                const int e_total = val + (f ? carry_in : 0);
                int ext = __shfl_down_sync(0xffffffffu, e_total, 1);
                if (lane_chunk == nchp - 1) ext = carry_in;
                if (nblocks > 1 && lane_chunk == 0 && rl[u] < unit.row_count) carry_s[rl[u]] = e_total;
                uint32_t prev = __shfl_up_sync(0xffffffffu, g, 1);
                prev = (lane_chunk == 0) ? 0u : (prev >> 31);
                uint32_t starts = g & ~((g << 1) | prev);
                while (starts) {
                    const int i = __ffs(starts) - 1;
                    starts &= starts - 1;
                    const uint32_t x = ~(g >> i);
                    const int ones = x ? (__ffs(x) - 1) : 32;
                    int end_col = colbase + i + ones - 1;
                    if (i + ones == 32) end_col += ext;
                    atomicMax(&B_s[(lane_chunk << 5) + i], (unsigned)(end_col - a0 + 1));
                }
            }
        }
        // merge row slots inside the warp
        for (int d = nchp; d < 32; d <<= 1) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                a_or[j] |= __shfl_xor_sync(0xffffffffu, a_or[j], d);
                a_nor[j] |= __shfl_xor_sync(0xffffffffu, a_nor[j], d);
            }
        }
        if (lane_slot == 0 && chunk_valid) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                acc_s[warp][lane_chunk][j] = a_or[j];
                acc_s[warp][lane_chunk][4 + j] = a_nor[j];
            }
        }
        __syncthreads();
        // merge warps, publish
        const long long word_base = ((long long)t.col_off >> 3) + (long long)(bch0 - ch0) * 4;
        for (int i = threadIdx.x; i < bn * 8; i += SCAN_THREADS) {
            const int c = i >> 3, j = i & 7;
            uint32_t x = 0;
#pragma unroll
            for (int w = 0; w < SCAN_WARPS; ++w) x |= acc_s[w][c][j];
            if (x) {
                if (j < 4) atomicOr(&colOR[word_base + c * 4 + j], x);
                else atomicOr(&colNOR[word_base + c * 4 + (j - 4)], x);
            }
        }
        const long long col_base = (long long)t.col_off + (long long)(bch0 - ch0) * 32;
        for (int i = threadIdx.x; i < bn * 32; i += SCAN_THREADS) {
            const unsigned b = B_s[i];
            if (b) atomicMax(&colB[col_base + i], b);
        }
        __syncthreads();
    }
}

cudaError_t launch_scan(cudaStream_t stream, bool has_n, const uint8_t *packed, const DTask *d_tasks,
                        const ScanUnit *d_units, int n_units, int max_unit_rows,
                        const int *d_rows, uint32_t *colOR, uint32_t *colNOR, unsigned *colB) {
    if (n_units <= 0) return cudaSuccess;
    const size_t smem = sizeof(int) * (size_t)max_unit_rows;
    if (has_n)
        scan_kernel<true><<<n_units, SCAN_THREADS, smem, stream>>>(packed, d_tasks, d_units, d_rows,
                                                                    colOR, colNOR, colB);
    else
        scan_kernel<false><<<n_units, SCAN_THREADS, smem, stream>>>(packed, d_tasks, d_units, d_rows,
                                                                     colOR, colNOR, colB);
    return cudaGetLastError();
}

}  // namespace mprg
