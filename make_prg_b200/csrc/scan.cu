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
// gap_reach by a prefix maximum: has_empty_sequence([s, e]) == (gap_reach[s] >= e).
//
// Work decomposition: one WARP per ScanUnit = a tile (row range x up to 32 chunks) of one task; warps
// never cooperate, there is no shared memory and no barrier.  Inside a warp the 32 lanes cover
// (32 / W) rows x W chunks, W = chunks of the tile rounded up to a power of two, so narrow windows
// (deep recursion levels) still use every lane and a wide tile reads 512 contiguous bytes per row.
// A gap run that reaches the end of its chunk is continued by the warp reading on in that row (32
// chunks per step), a run that starts in the first column of a chunk looks one column back: tiles need
// no carry between them, so tall tasks are cut by rows and wide ones by columns freely.  Partial
// results reach the (zero-initialised) per-task column arrays with atomicOr / atomicMax.
//
// Rows are fetched with 128-bit ld.global.nc (L1 no-allocate), four rows per trip, software pipelined
// by one trip.  A TMA variant (per-warp shared-memory ring filled by cp.async.bulk on mbarriers) and a
// persistent-grid variant with a global tile counter were built and measured slower on B200 for this
// access pattern (profiles/r1_scan_variants.txt); they are not kept.
#include <cstdlib>

#include "common.cuh"
#include "kernels.cuh"

namespace mprg {

constexpr int SCAN_THREADS = 128;
constexpr int SCAN_WARPS = SCAN_THREADS / 32;
constexpr int SCAN_UNROLL = 4;  // warp iterations per trip

__device__ __forceinline__ uint4 ld_stream(const uint4 *p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}

// ---- TMA (bulk async copy) variant: a tile whose rows are consecutive AND complete (the tile spans the whole
// packed row) is one contiguous block of global memory, so ONE lane fetches it with ONE cp.async.bulk into the
// warp's shared-memory slab and the warp waits on an mbarrier; the row loop then reads 128-bit vectors from
// shared memory.  Many warps per SM keep many such copies in flight.  (Round 1 measured a ring of per-lane 2 KB
// copies per warp, which serialised in the issue path.)
constexpr int SCAN_TMA_SLAB = 8192;  // bytes per warp: 16 rows of 512 bytes (1,000-column loci)

__device__ __forceinline__ uint32_t smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_addr(dst)),
                 "l"(src), "r"(bytes), "r"(smem_addr(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_addr(bar)),
        "r"(parity)
        : "memory");
}

// bit 4k+3 set iff nibble k of w is zero (the gap code), exact: no carry crosses a nibble
__device__ __forceinline__ uint32_t zero_nibbles(uint32_t w) {
    static_assert(SYM_GAP == 0, "the zero-nibble test assumes gap == 0");
    static_assert((SYM_PAD & 1) == 1, "padding must pass the odd-code (no gap) filter");
    return ~(((w & 0x77777777u) + 0x77777777u) | w) & 0x88888888u;
}

// 32-bit gap mask of a chunk, bit c = column c: with the word-interleaved chunk layout (column c in
// word c & 3, nibble c >> 2) the four per-word indicators interleave with three shifts
__device__ __forceinline__ uint32_t gap_mask32(const uint4 &v) {
    return (zero_nibbles(v.x) >> 3) | (zero_nibbles(v.y) >> 2) | (zero_nibbles(v.z) >> 1) | zero_nibbles(v.w);
}

// The same mask for a chunk without even symbol codes (M S W N): a nibble is the gap iff its bit 0 is
// clear.  Bit 0 of every nibble of the four words is gathered with three bit-selects (LOP3) on
// shifted words: 7 instructions instead of 17.
__device__ __forceinline__ uint32_t gap_mask32_odd(const uint4 &v) {
    const uint32_t z01 = (v.x & 0x11111111u) | ((v.y << 1) & ~0x11111111u);  // bit 0: word 0, bit 1: word 1
    const uint32_t z23 = (v.z & 0x11111111u) | ((v.w << 1) & ~0x11111111u);  // bit 0: word 2, bit 1: word 3
    return ~((z01 & 0x33333333u) | ((z23 << 2) & ~0x33333333u));
}

// nibble-wide (0xF) mask of the nibbles equal to `pattern`
__device__ __forceinline__ uint32_t nibble_eq_maskF(uint32_t w, uint32_t pattern) {
    uint32_t t = w ^ pattern;
    t |= t >> 1;
    t |= t >> 2;
    return (~t & 0x11111111u) * 15u;
}

// The warp reads on in the row `rp` from chunk k0: number of consecutive gap columns that follow, up
// to the end of the window (32 chunks per step; only runs longer than a chunk, or runs that leave the
// tile, come here).  Warp-uniform call, every lane returns the count.
__device__ __forceinline__ int scan_walk(const uint8_t *rp, int k0, int lane, int c1) {
    int e = 0;
    while (true) {
        const int k = k0 + lane;
        uint32_t m = 0u;
        if ((k << 5) < c1) {
            m = gap_mask32(ld_stream(reinterpret_cast<const uint4 *>(rp + ((long long)k << 4))));
            const int hi = c1 - (k << 5);
            if (hi < 32) m &= (1u << hi) - 1u;
        }
        const uint32_t fullb = __ballot_sync(0xffffffffu, m == 0xffffffffu);
        if (fullb == 0xffffffffu) {
            e += 1024;
            k0 += 32;
            continue;
        }
        const int nf = __ffs(~fullb) - 1;
        const int lead = __ffs(~m) - 1;
        return e + 32 * nf + __shfl_sync(0xffffffffu, lead, nf);
    }
}

// Gap path of one warp iteration (warp-uniform call).  g = gap mask of the lane's chunk clipped to the
// task window, todo = the lane has runs to record.  Records B[run start] = run end + 1 (columns
// relative to a0) for every maximal gap run of the window that starts in the lane's chunk.  The masks
// of the two neighbouring chunks come from the neighbouring lanes; only a run that covers the whole
// next chunk, or leaves the tile, makes the warp read on in that row (scan_walk).
struct ScanLane {
    int chunk, lane, lane_chunk, bn, c0, c1, a0;
};
// Inlined four times (once per row of a trip).  A single out-of-line copy (smaller loop, fewer
// instruction-cache misses) was measured at 2.1 instead of 3.2 TB/s: the call pins the accumulators
// to the ABI's registers (profiles/r1_scan_variants.txt).
#ifdef MPRG_SCAN_NOINLINE_GAPS
#define MPRG_GAP_INLINE __noinline__
#else
#define MPRG_GAP_INLINE __forceinline__
#endif
template <typename RowPtr>
__device__ MPRG_GAP_INLINE void scan_gap_rows(uint32_t g, bool todo, const ScanLane &L, RowPtr row_ptr,
                                              unsigned *B) {
    const int colbase = L.chunk << 5;
    uint32_t gn = __shfl_down_sync(0xffffffffu, g, 1);
    uint32_t gp = __shfl_up_sync(0xffffffffu, g, 1);
    // tile edges: the neighbour is not a lane of this row
    if (L.lane_chunk >= L.bn - 1) gn = (colbase + 32 < L.c1) ? 0xffffffffu : 0u;  // unknown => "read on"
    uint32_t prev = gp >> 31;
    if (L.lane_chunk == 0) {
        prev = 0u;
        // first chunk of a tile that is not the first of the window: look one column back
        // (column 31 of the previous chunk = high nibble of its last byte)
        if (todo && (g & 1u) && colbase > L.c0)
            prev = ((row_ptr()[(L.chunk << 4) - 1] >> 4) == SYM_GAP) ? 1u : 0u;
    }
    const uint32_t starts = g & ~((g << 1) | prev);
    const bool reaches_end = (g >> 31) != 0u;
    // gap columns that follow the end of my chunk in this row
    int ext = reaches_end ? ((gn == 0xffffffffu) ? 32 : (__ffs(~gn) - 1)) : 0;
    uint32_t need = __ballot_sync(0xffffffffu, todo && reaches_end && gn == 0xffffffffu && starts != 0u);
    if (need) {
        const uint8_t *mine = row_ptr();
        while (need) {
            const int src = __ffs(need) - 1;
            need &= need - 1;
            const uint8_t *rp = reinterpret_cast<const uint8_t *>(
                __shfl_sync(0xffffffffu, reinterpret_cast<unsigned long long>(mine), src));
            const int k0 = __shfl_sync(0xffffffffu, L.chunk, src) + 1;
            const int e = scan_walk(rp, k0, L.lane, L.c1);
            if (L.lane == src) ext = e;
        }
    }
    if (todo) {
        uint32_t st = starts;
        while (st) {
            const int i = __ffs(st) - 1;
            st &= st - 1;
            const uint32_t x = ~(g >> i);
            const int ones = x ? (__ffs(x) - 1) : 32;
            int end = colbase + i + ones;  // exclusive
            if (i + ones == 32) end += ext;
            atomicMax(&B[colbase + i - L.a0], (unsigned)(end - L.a0));
        }
    }
}

template <bool HAS_N, bool TMA>
#ifndef MPRG_SCAN_MIN_BLOCKS
#define MPRG_SCAN_MIN_BLOCKS 8  // 64 registers (11 words spilled): 32 resident warps per SM instead of 28, +6..11 % measured
#endif
__global__ void __launch_bounds__(SCAN_THREADS, MPRG_SCAN_MIN_BLOCKS)
scan_kernel(const uint8_t *__restrict__ packed, const ScanUnit *__restrict__ units, int n_units,
            const int *__restrict__ rows_arena, uint32_t *__restrict__ colOR,
            uint32_t *__restrict__ colNOR, unsigned *__restrict__ colB) {
    const int lane = threadIdx.x & 31;
    const int ui = blockIdx.x * SCAN_WARPS + (threadIdx.x >> 5);
    __shared__ __align__(128) uint8_t s_slab[TMA ? SCAN_WARPS : 1][TMA ? SCAN_TMA_SLAB : 16];
    __shared__ uint64_t s_bar[SCAN_WARPS];
    if (TMA) {
        if (threadIdx.x < SCAN_WARPS) mbar_init(&s_bar[threadIdx.x], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        __syncthreads();
    }
    if (ui >= n_units) return;
    const ScanUnit t = units[ui];
    const int bn = t.ch_count;
    const int row_count = t.row_count;
    const int a0 = (t.c0 >> 5) << 5;  // chunk-aligned window start (columns)
    const int *rows = t.rows_off >= 0 ? rows_arena + t.rows_off : nullptr;
    const uint8_t *msa = packed + t.base;

    const int lg = bn > 1 ? 32 - __clz(bn - 1) : 0;
    const int rpw = 32 >> lg;  // rows per warp iteration
    const int lane_chunk = lane & ((1 << lg) - 1);
    const int lane_slot = lane >> lg;
    const bool chunk_valid = lane_chunk < bn;
    const int chunk = t.ch_begin + min(lane_chunk, bn - 1);  // lanes beyond the tile re-read its last chunk
    const int colbase = chunk << 5;
    // window mask: bit i set iff c0 <= colbase + i < c1 (0 for lanes beyond the tile)
    uint32_t wmask = 0;
    if (chunk_valid) {
        const int lo = max(t.c0 - colbase, 0);
        const int hi = min(t.c1 - colbase, 32);
        if (hi > lo) wmask = (hi - lo == 32) ? 0xffffffffu : (((1u << (hi - lo)) - 1u) << lo);
    }
    const uint8_t *col_ptr = msa + ((long long)chunk << 4);
    unsigned *B = colB + (long long)t.col_off;

    uint32_t a_or[4] = {0, 0, 0, 0};
    uint32_t a_and[4] = {0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu};
    const int n_iters = (row_count + rpw - 1) / rpw;
    // Every reduction here is idempotent (OR, AND, max), so missing rows at the end of the tile are
    // replaced by its last row: no tail code, and all lanes of every trip do useful, uniform work.
    const int last_row = row_count - 1;
    auto row_of = [&](int it) {
        const int rl = min(it * rpw + lane_slot, last_row);
        return rows ? rows[rl] : t.row_begin + rl;
    };
    const ScanLane L = {chunk, lane, lane_chunk, bn, t.c0, t.c1, a0};
    uint32_t g_done = 0u;  // last gap mask (interior runs only) this lane has recorded
    const bool exact = HAS_N || (t.flags & 1);

    // consecutive rows: one pointer per lane, advanced by whole iterations (no index arithmetic per row)
    const uint8_t *lane_ptr = col_ptr + (long long)(t.row_begin + lane_slot) * t.stride;
    const long long it_bytes = (long long)rpw * t.stride;
    auto load_trip = [&](int first_it, uint4 *dst) {
        if (rows == nullptr && (first_it + SCAN_UNROLL) * rpw <= row_count) {
            const uint8_t *p = lane_ptr + first_it * it_bytes;
#pragma unroll
            for (int u = 0; u < SCAN_UNROLL; ++u) dst[u] = ld_stream(reinterpret_cast<const uint4 *>(p + u * it_bytes));
        } else {
#pragma unroll
            for (int u = 0; u < SCAN_UNROLL; ++u)
                dst[u] = ld_stream(reinterpret_cast<const uint4 *>(col_ptr + (long long)row_of(first_it + u) * t.stride));
        }
    };
    // TMA: the whole tile in one bulk copy when it is one contiguous block that fits the slab
    const uint8_t *slab = s_slab[TMA ? (threadIdx.x >> 5) : 0];
    bool staged = false;
    if (TMA) {
        const long long tile_bytes = (long long)row_count * t.stride;
        staged = rows == nullptr && t.ch_begin == 0 && bn * CHUNK_BYTES == t.stride && tile_bytes <= SCAN_TMA_SLAB;
        if (staged) {
            uint64_t *bar = &s_bar[threadIdx.x >> 5];
            if (lane == 0) {
                mbar_expect_tx(bar, (uint32_t)tile_bytes);
                bulk_g2s(const_cast<uint8_t *>(slab), msa + (long long)t.row_begin * t.stride, (uint32_t)tile_bytes, bar);
            }
            mbar_wait(bar, 0);
        }
    }
    const int slab_col = (chunk - t.ch_begin) << 4;
    auto load_staged = [&](int first_it, uint4 *dst) {
#pragma unroll
        for (int u = 0; u < SCAN_UNROLL; ++u) {
            const int rl = min((first_it + u) * rpw + lane_slot, last_row);
            dst[u] = *reinterpret_cast<const uint4 *>(slab + rl * t.stride + slab_col);
        }
    };
    uint4 vnext[SCAN_UNROLL];
    if (TMA && staged) load_staged(0, vnext);
    else load_trip(0, vnext);
    // software pipelined by one trip: the loads of trip i+1 are in flight while trip i is processed
    for (int it0 = 0; it0 < n_iters; it0 += SCAN_UNROLL) {
        uint4 v[SCAN_UNROLL];
#pragma unroll
        for (int u = 0; u < SCAN_UNROLL; ++u) v[u] = vnext[u];
        if (it0 + SCAN_UNROLL < n_iters) {
            if (TMA && staged) load_staged(it0 + SCAN_UNROLL, vnext);
            else load_trip(it0 + SCAN_UNROLL, vnext);
        }
        uint32_t gm[SCAN_UNROLL];  // gap mask of my chunk in each of the rows (0 outside the window)
#pragma unroll
        for (int u = 0; u < SCAN_UNROLL; u += 2) {
            const uint32_t wa[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
            const uint32_t wb[4] = {v[u + 1].x, v[u + 1].y, v[u + 1].z, v[u + 1].w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (HAS_N) {
                    const uint32_t ka = ~nibble_eq_maskF(wa[j], SYM_N * 0x11111111u);
                    const uint32_t kb = ~nibble_eq_maskF(wb[j], SYM_N * 0x11111111u);
                    a_or[j] |= (wa[j] & ka) | (wb[j] & kb);
                    a_and[j] &= (wa[j] | ~ka) & (wb[j] | ~kb);
                } else {
                    a_or[j] |= wa[j] | wb[j];   // one LOP3 for two rows
                    a_and[j] &= wa[j] & wb[j];  // one LOP3 for two rows
                }
            }
            // Bases and padding are odd codes and the gap is 0: without even codes in the locus the gap
            // mask is bit 0 of every nibble, complemented
            if (exact) {
                gm[u] = gap_mask32(v[u]) & wmask;
                gm[u + 1] = gap_mask32(v[u + 1]) & wmask;
            } else {
                gm[u] = gap_mask32_odd(v[u]) & wmask;
                gm[u + 1] = gap_mask32_odd(v[u + 1]) & wmask;
            }
        }
#pragma unroll
        for (int u = 0; u < SCAN_UNROLL; ++u) {
            // Rows of one clade share their deletions: a mask equal to the last one this lane recorded
            // (runs strictly inside the chunk, so nothing depends on the neighbours) adds nothing.
            const bool todo = gm[u] != 0u && gm[u] != g_done;
            if (__any_sync(0xffffffffu, todo)) {
                scan_gap_rows(gm[u], todo, L, [&]() { return msa + (long long)row_of(it0 + u) * t.stride; }, B);
                if (todo && (gm[u] & 0x80000001u) == 0u) g_done = gm[u];
            }
        }
    }
    // merge the row slots of the warp, publish
    for (int d = 1 << lg; d < 32; d <<= 1) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            a_or[j] |= __shfl_xor_sync(0xffffffffu, a_or[j], d);
            a_and[j] &= __shfl_xor_sync(0xffffffffu, a_and[j], d);
        }
    }
    if (lane_slot == 0 && chunk_valid) {
        // col_off and the chunk base are multiples of 32 columns, so the lane's 4 words are 16-byte aligned:
        // two 64-bit reductions per array instead of four 32-bit ones (the publishing atomics measured ~6 % of
        // the 105 MB launch)
        const long long word = (((long long)t.col_off + (colbase - a0)) >> 3);
        unsigned long long *o64 = reinterpret_cast<unsigned long long *>(colOR + word);
        unsigned long long *n64 = reinterpret_cast<unsigned long long *>(colNOR + word);
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const unsigned long long vo = (unsigned long long)a_or[2 * j] | ((unsigned long long)a_or[2 * j + 1] << 32);
            const unsigned long long vn =
                (unsigned long long)(~a_and[2 * j]) | ((unsigned long long)(~a_and[2 * j + 1]) << 32);
            if (vo) atomicOr(&o64[j], vo);
            if (vn) atomicOr(&n64[j], vn);  // published as OR of complements
        }
    }
}

cudaError_t launch_scan(cudaStream_t stream, bool has_n, const uint8_t *packed, const ScanUnit *d_units,
                        int n_units, const int *d_rows, uint32_t *colOR, uint32_t *colNOR, unsigned *colB) {
    if (n_units <= 0) return cudaSuccess;
    const int grid = (n_units + SCAN_WARPS - 1) / SCAN_WARPS;
    // MPRG_SCAN_TMA=1: tiles that are one contiguous block of rows come in by one bulk async copy per warp
    // (profiles/r2_scan_tma.txt has the comparison)
    static const bool tma = getenv("MPRG_SCAN_TMA") != nullptr;
    if (has_n)
        scan_kernel<true, false><<<grid, SCAN_THREADS, 0, stream>>>(packed, d_units, n_units, d_rows, colOR, colNOR, colB);
    else if (tma)
        scan_kernel<false, true><<<grid, SCAN_THREADS, 0, stream>>>(packed, d_units, n_units, d_rows, colOR, colNOR, colB);
    else
        scan_kernel<false, false><<<grid, SCAN_THREADS, 0, stream>>>(packed, d_units, n_units, d_rows, colOR, colNOR, colB);
    return cudaGetLastError();
}

}  // namespace mprg
