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
constexpr int SCAN_UNROLL = 4;

__device__ __forceinline__ uint4 ld_stream(const uint4 *p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}

// bit 4k+3 set iff nibble k of w is zero (the gap code), exact: no carry crosses a nibble
__device__ __forceinline__ uint32_t zero_nibbles(uint32_t w) {
    static_assert(SYM_GAP == 0, "the zero-nibble test assumes gap == 0");
    return ~(((w & 0x77777777u) + 0x77777777u) | w) & 0x88888888u;
}

// 32-bit gap mask of a chunk, bit c = column c: with the word-interleaved chunk layout (column c in
// word c & 3, nibble c >> 2) the four per-word indicators interleave with three shifts
__device__ __forceinline__ uint32_t gap_mask32(const uint4 &v) {
    return (zero_nibbles(v.x) >> 3) | (zero_nibbles(v.y) >> 2) | (zero_nibbles(v.z) >> 1) | zero_nibbles(v.w);
}

// nibble-wide (0xF) mask of the nibbles equal to `pattern`
__device__ __forceinline__ uint32_t nibble_eq_maskF(uint32_t w, uint32_t pattern) {
    uint32_t t = w ^ pattern;
    t |= t >> 1;
    t |= t >> 2;
    return (~t & 0x11111111u) * 15u;
}

// Gap runs that stay inside the lane's chunk (neither its first nor its last column is a gap) need no
// cooperation: the common case for short indels.  Returns the mask when the lane needs the
// cooperative path (a run touches a chunk border), 0 otherwise.
__device__ __forceinline__ uint32_t scan_gap_local(uint32_t g, int colbase, int a0, unsigned *B) {
    if (g & 0x80000001u) return g;
    uint32_t starts = g & ~(g << 1);
    while (starts) {
        const int i = __ffs(starts) - 1;
        starts &= starts - 1;
        const int ones = __ffs(~(g >> i)) - 1;
        atomicMax(&B[colbase - a0 + i], (unsigned)(colbase + i + ones - a0));
    }
    return 0u;
}

// Slow path of one warp iteration (one row per `nchp` lanes) that holds at least one gap: record
// B[run start] = run end + 1 for every gap run, stitching runs across lanes (ballots + one shuffle)
// and across 1024-column blocks (carry_s).
__device__ __forceinline__ void scan_gap_rows(uint32_t g, int lane,
                                              int lane_chunk, int nchp, int colbase, int a0, int rl,
                                              int row_count, bool right_block_exists, bool multi_block,
                                              uint32_t topmask, int *carry_s, unsigned *B) {
    const uint32_t firstb = __ballot_sync(0xffffffffu, g & 1u);
    const uint32_t lastb = __ballot_sync(0xffffffffu, g >> 31);
    const uint32_t fullb = __ballot_sync(0xffffffffu, g == 0xffffffffu);
    const int seg_base = lane - lane_chunk;
    const uint32_t seg_mask = (nchp == 32 ? 0xffffffffu : ((1u << nchp) - 1u)) << seg_base;
    int carry_in = 0;
    if (right_block_exists && rl < row_count) carry_in = carry_s[rl];
    // does any run of this warp iteration continue past its chunk?  (warp-uniform: `topmask` marks the
    // last lane of every row segment, whose right neighbour is the block to the right, not lane + 1)
    const uint32_t cont = lastb & (((firstb >> 1) & ~topmask) | (right_block_exists ? topmask : 0u));
    int ext = 0;  // gap columns that follow the end of my chunk in the same row
    const int lead = (g == 0xffffffffu) ? 32 : (__ffs(~g) - 1);
    if (cont != 0u || multi_block) {
        // consecutive full chunks after mine (inside the row's lanes), then the partial lead of the next
        const uint32_t after = (lane == 31) ? 0u : ((fullb & seg_mask) >> (lane + 1));
        int nfull = (after == 0xffffffffu) ? 32 : (__ffs(~after) - 1);
        const int remaining = nchp - 1 - lane_chunk;
        nfull = min(nfull, remaining);
        const int k = lane + 1 + nfull;  // first lane after the full ones
        const bool k_in_seg = (nfull < remaining);
        const int lead_k = __shfl_sync(0xffffffffu, lead, k_in_seg ? k : lane);
        ext = 32 * nfull + (k_in_seg ? lead_k : carry_in);
        if (multi_block && lane_chunk == 0 && rl < row_count)
            carry_s[rl] = (g == 0xffffffffu) ? 32 + ext : lead;
    }
    const uint32_t prev = (lane_chunk == 0) ? 0u : ((lastb >> (lane - 1)) & 1u);
    uint32_t starts = g & ~((g << 1) | prev);
    while (starts) {
        const int i = __ffs(starts) - 1;
        starts &= starts - 1;
        const uint32_t x = ~(g >> i);
        const int ones = x ? (__ffs(x) - 1) : 32;
        int end_col = colbase + i + ones - 1;
        if (i + ones == 32) end_col += ext;
        atomicMax(&B[colbase - a0 + i], (unsigned)(end_col - a0 + 1));
    }
}

template <bool HAS_N>
__global__ void __launch_bounds__(SCAN_THREADS)
scan_kernel(const uint8_t *__restrict__ packed, const DTask *__restrict__ tasks,
            const ScanUnit *__restrict__ units, const int *__restrict__ rows_arena,
            uint32_t *__restrict__ colOR, uint32_t *__restrict__ colNOR,
            unsigned *__restrict__ colB) {
    __shared__ uint32_t acc_s[SCAN_WARPS][SCAN_BLOCK_CHUNKS][8];
    extern __shared__ int carry_s[];  // one int per row of the unit (only used when nblocks > 1)

    const ScanUnit t = units[blockIdx.x];
    const ScanUnit &unit = t;
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int ch0 = t.c0 >> 5;
    const int ch1 = (t.c1 + 31) >> 5;
    const int nch = ch1 - ch0;
    if (nch <= 0 || unit.row_count <= 0) return;
    const int a0 = ch0 << 5;  // chunk-aligned window start (columns)
    const int nblocks = (nch + SCAN_BLOCK_CHUNKS - 1) / SCAN_BLOCK_CHUNKS;
    const bool multi_block = nblocks > 1;
    const int *rows = t.rows_off >= 0 ? rows_arena + t.rows_off : nullptr;
    const uint8_t *msa = packed + t.base;
    const int row_count = unit.row_count;

    for (int blk = nblocks - 1; blk >= 0; --blk) {
        const int bch0 = ch0 + blk * SCAN_BLOCK_CHUNKS;
        const int bn = min(SCAN_BLOCK_CHUNKS, ch1 - bch0);
        const int lg = bn > 1 ? 32 - __clz(bn - 1) : 0;
        const int nchp = 1 << lg;       // lanes per row
        const int rpw = 32 >> lg;       // rows per warp iteration
        const int lane_chunk = lane & (nchp - 1);
        const int lane_slot = lane >> lg;
        const bool chunk_valid = lane_chunk < bn;
        const int chunk = bch0 + min(lane_chunk, bn - 1);  // lanes beyond the block re-read its last chunk
        const int colbase = (bch0 + lane_chunk) << 5;
        // window mask: bit i set iff c0 <= colbase + i < c1 (0 for lanes beyond the block)
        uint32_t wmask = 0;
        if (chunk_valid) {
            const int lo = max(t.c0 - colbase, 0);
            const int hi = min(t.c1 - colbase, 32);
            if (hi > lo) wmask = (hi - lo == 32) ? 0xffffffffu : (((1u << (hi - lo)) - 1u) << lo);
        }
        const bool right_block_exists = blk < nblocks - 1;
        const uint32_t topmask = __ballot_sync(0xffffffffu, lane_chunk == nchp - 1);
        const uint8_t *col_ptr = msa + (long long)chunk * CHUNK_BYTES;

        uint32_t a_or[4] = {0, 0, 0, 0};
        uint32_t a_and[4] = {0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu};
        const int n_iters = (row_count + rpw - 1) / rpw;
        // Every reduction here is idempotent (OR, AND, max), so missing rows at the end of the unit are
        // replaced by its last row: no tail code, and all lanes of every trip do useful, uniform work.
        const int last_row = row_count - 1;
        auto load_trip = [&](int first_it, uint4 *dst) {
#pragma unroll
            for (int u = 0; u < SCAN_UNROLL; ++u) {
                const int rl = min((first_it + u * SCAN_WARPS) * rpw + lane_slot, last_row);
                const int row = rows ? rows[rl] : unit.row_begin + rl;
                dst[u] = ld_stream(reinterpret_cast<const uint4 *>(col_ptr + (long long)row * t.stride));
            }
        };
        int it0 = warp;
        uint4 vnext[SCAN_UNROLL];
        if (it0 < n_iters) load_trip(it0, vnext);
        // gap-run ends go straight to the task's (zero-initialised) B array with atomicMax: about one
        // run per row, far cheaper than clearing and flushing a shared tile per CTA
        unsigned *B = colB + (long long)t.col_off;
        // software pipelined by one trip: the loads of trip i+1 are in flight while trip i is processed
        for (; it0 < n_iters; it0 += SCAN_WARPS * SCAN_UNROLL) {
            uint4 v[SCAN_UNROLL];
#pragma unroll
            for (int u = 0; u < SCAN_UNROLL; ++u) v[u] = vnext[u];
            const int nit = it0 + SCAN_WARPS * SCAN_UNROLL;
            if (nit < n_iters) load_trip(nit, vnext);
            uint32_t gm[SCAN_UNROLL];  // gap mask of my chunk in each of the rows (0 outside the window)
#pragma unroll
            for (int u = 0; u < SCAN_UNROLL; u += 2) {
                const uint32_t wa[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
                const uint32_t wb[4] = {v[u + 1].x, v[u + 1].y, v[u + 1].z, v[u + 1].w};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    if (HAS_N) {
                        const uint32_t ka = ~nibble_eq_maskF(wa[j], 0xBBBBBBBBu);
                        const uint32_t kb = ~nibble_eq_maskF(wb[j], 0xBBBBBBBBu);
                        a_or[j] |= (wa[j] & ka) | (wb[j] & kb);
                        a_and[j] &= (wa[j] | ~ka) & (wb[j] | ~kb);
                    } else {
                        a_or[j] |= wa[j] | wb[j];   // one LOP3 for two rows
                        a_and[j] &= wa[j] & wb[j];  // one LOP3 for two rows
                    }
                }
                gm[u] = gap_mask32(v[u]) & wmask;
                gm[u + 1] = gap_mask32(v[u + 1]) & wmask;
            }
            // lanes that hold a gap resolve their interior runs on their own; only runs touching a chunk
            // border (or multi-block rows, which must hand a carry to the next block) go cooperative
            uint32_t gb[SCAN_UNROLL];
            uint32_t any_gb = 0;
#pragma unroll
            for (int u = 0; u < SCAN_UNROLL; ++u) {
                gb[u] = 0u;
                if (gm[u]) gb[u] = scan_gap_local(gm[u], colbase, a0, B);
                any_gb |= gb[u];
            }
            if (multi_block || __any_sync(0xffffffffu, any_gb != 0u)) {
#pragma unroll
                for (int u = 0; u < SCAN_UNROLL; ++u) {
                    const int rl = min((it0 + u * SCAN_WARPS) * rpw + lane_slot, last_row);
                    if (multi_block || __any_sync(0xffffffffu, gb[u] != 0u))
                        scan_gap_rows(gb[u], lane, lane_chunk, nchp, colbase, a0, rl, row_count,
                                      right_block_exists, multi_block, topmask, carry_s, B);
                }
            }
        }
        // merge row slots inside the warp
        for (int d = nchp; d < 32; d <<= 1) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                a_or[j] |= __shfl_xor_sync(0xffffffffu, a_or[j], d);
                a_and[j] &= __shfl_xor_sync(0xffffffffu, a_and[j], d);
            }
        }
        if (lane_slot == 0 && chunk_valid) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                acc_s[warp][lane_chunk][j] = a_or[j];
                acc_s[warp][lane_chunk][4 + j] = ~a_and[j];  // published as OR of complements
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
