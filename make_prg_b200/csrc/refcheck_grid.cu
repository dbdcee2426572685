// One-reference-like check of a DEEP clustering problem with the whole grid, straight from the 4-bit packed
// MSA (cluster_sequences.py:59-104: per cluster the majority symbol of every column over the member rows with
// multiplicity, ties to the first-seen symbol; then the Hamming distance of every member row to that string
// against the threshold).  Replaces refcheck_big_majority_kernel / refcheck_big_hamming_kernel (one thread per
// column walking all rows of the unpacked copy: 6.7 ms + 2.4 ms on BASELINE config #4) for loci whose alphabet
// is ACGT- (no flag but "valid": what the pack kernel reports); other loci keep the byte-wise kernels.
//
// The integer contraction here is  counts[cluster][column][symbol] = sum_rows onehot(cluster) x onehot(symbol).
// As a tensor-core GEMM it would need the one-hot expansion of every 4-bit symbol to 5..16 int8 operands in
// shared memory (8..32 x the bytes of the packed rows) for an M = 10 "clusters" dimension padded to 64/128;
// bit-sliced SIMT counting does the same sums in ~10 integer instructions per 8 symbols on the packed words
// themselves, so the kernel is bound by reading the rows once from HBM (profiles/).
//
//   cluster_rows_kernel   member rows of every cluster, in member order (one warp per cluster)
//   majority_count_kernel one warp per (cluster, 240 rows, 1,024 columns), a 128-bit load per lane and row: SWAR
//                         bit-plane counters in registers -- codes A C G T = 1 3 5 7 share bit 0, so four sums
//                         (b0, b0 b1, b0 b2, b0 b1 b2) give the five symbol counts -- nibble fields spilled to byte
//                         fields every 15 rows, published as one 64-bit atomic (4 x 16 bits) per column and tile
//   majority_pick_kernel  arg-max per (cluster, column); exact ties resolved by reading the rows in order
//   hamming_packed_kernel one warp per member row: popcount of the non-zero nibbles of row XOR majority
#include "common.cuh"
#include "kernels.cuh"

namespace mprg {

#ifndef MPRG_RG_WARP_ROWS
#define MPRG_RG_WARP_ROWS 60
#endif
constexpr int RG_WARP_ROWS = MPRG_RG_WARP_ROWS;  // rows a warp counts in registers (groups of 15; byte fields hold 255)
constexpr int RG_ROWS = 4 * RG_WARP_ROWS;     // rows per CTA of the counting kernel: its four warps share a column group
constexpr unsigned NIB = 0x11111111u;

struct RefGrid {
    long long base;   // locus in the packed arena
    int stride;       // bytes per packed row
    int rows_off;     // task rows in the row arena, -1 => rows are 0..R-1
    int R;            // rows of the task
    int c0, c1;       // column window
    int ch0, n_words; // first chunk, 32-bit words per row that cover the window
    int K_max;
};

__device__ __forceinline__ const uint32_t *row_words(const uint8_t *packed, const RefGrid &g, const int *rows_arena,
                                                     int task_row) {
    const int lr = g.rows_off >= 0 ? rows_arena[g.rows_off + task_row] : task_row;
    return reinterpret_cast<const uint32_t *>(packed + g.base + (long long)lr * g.stride + (long long)g.ch0 * 16);
}

// crows[c * R + i] = i-th member row (task-local) of cluster c in member order; ccount[c] = how many
__global__ void __launch_bounds__(320)
cluster_rows_kernel(const ClusterState *__restrict__ states, int q, RefGrid g, const int *__restrict__ mem_off_all,
                    const int *__restrict__ mem_rows_all, const int *__restrict__ assign_all, int *__restrict__ crows,
                    int *__restrict__ ccount) {
    const ClusterState &st = states[q];
    const int lane = threadIdx.x & 31, c = threadIdx.x >> 5;
    if (st.status != 0 || c >= st.K) {
        if (lane == 0 && c < g.K_max) ccount[c] = 0;
        return;
    }
    const int *mem_off = mem_off_all + st.mem_off;
    const int *mem_rows = mem_rows_all + st.mem_rows_off;
    const int *assign = assign_all + st.assign_off;
    int *out = crows + (long long)c * g.R;
    int total = 0;
    // four chunks of 32 sequences per trip: their loads are issued together (the loop is bound by the latency
    // of these reads: 313 dependent trips for 10,000 sequences took 470 us), the ordered placement follows
    for (int j0 = 0; j0 < st.n; j0 += 128) {
        int m0[4], cnt[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int j = j0 + 32 * u + lane;
            const bool mine = j < st.n && assign[j] == c;
            m0[u] = mine ? mem_off[j] : 0;
            cnt[u] = mine ? mem_off[j + 1] - m0[u] : 0;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            int incl = cnt[u];
            for (int d = 1; d < 32; d <<= 1) {
                const int o = __shfl_up_sync(0xffffffffu, incl, d);
                if (lane >= d) incl += o;
            }
            const int pos = total + incl - cnt[u];
            for (int m = 0; m < cnt[u]; ++m) out[pos + m] = mem_rows[m0[u] + m];
            total += __shfl_sync(0xffffffffu, incl, 31);
        }
    }
    if (lane == 0) ccount[c] = total;
}

// counts[c * w + col]: four 16-bit fields in one 64-bit word -- rows with a base | bases with bit 1 (C or T) << 16 |
// bases with bit 2 (G or T) << 32 | bases with both (T) << 48 -- so a tile publishes ONE atomic per column (the
// first version did four 32-bit ones per column and tile of 128 rows: 7 M atomics per pass were what bounded it).
// One CTA per (cluster, RG_ROWS rows, 32 chunks = 1,024 columns): each of its four warps counts a quarter of the
// rows in registers (a lane owns one chunk: 32 columns, a 128-bit load per row), the warps add up through shared
// memory and the CTA publishes one atomic per column.
#ifndef MPRG_RG_MIN_BLOCKS
#define MPRG_RG_MIN_BLOCKS 6
#endif
__global__ void __launch_bounds__(128, MPRG_RG_MIN_BLOCKS)
majority_count_kernel(const ClusterState *__restrict__ states, int q, RefGrid g, const uint8_t *__restrict__ packed,
                      const int *__restrict__ rows_arena, const int *__restrict__ crows, const int *__restrict__ ccount,
                      unsigned long long *__restrict__ counts) {
    const ClusterState &st = states[q];
    if (st.status != 0) return;
    __shared__ unsigned long long s_cnt[4][1024];  // [warp][column of the group]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int n_chunks = g.n_words >> 2;
    const int ch = blockIdx.x * 32 + lane;  // chunk of the window
    // row block -> (cluster, first row)
    int c = 0, blk = blockIdx.y, nrows = 0;
    for (; c < st.K; ++c) {
        nrows = ccount[c];
        const int nb = (nrows + RG_ROWS - 1) / RG_ROWS;
        if (blk < nb) break;
        blk -= nb;
    }
    if (c >= st.K) return;
    const bool in_window = ch < n_chunks;
    const int i0 = min(blk * RG_ROWS + warp * RG_WARP_ROWS, nrows), i1 = min(i0 + RG_WARP_ROWS, nrows);
    const int *rows = crows + (long long)c * g.R;
    unsigned a[4][4];      // [word][plane] nibble fields
    unsigned lo[4][4], hi[4][4];  // byte fields of the even / odd nibbles
#pragma unroll
    for (int wd = 0; wd < 4; ++wd)
#pragma unroll
        for (int p = 0; p < 4; ++p) a[wd][p] = lo[wd][p] = hi[wd][p] = 0;
    auto add = [&](const uint4 &v) {
        const unsigned x[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int wd = 0; wd < 4; ++wd) {
            const unsigned t1 = x[wd] >> 1, t2 = x[wd] >> 2;
            const unsigned p1 = x[wd] & t1 & NIB;
            a[wd][0] += x[wd] & NIB;
            a[wd][1] += p1;
            a[wd][2] += x[wd] & t2 & NIB;
            a[wd][3] += p1 & t2;
        }
    };
    auto spill = [&]() {
#pragma unroll
        for (int wd = 0; wd < 4; ++wd)
#pragma unroll
            for (int p = 0; p < 4; ++p) {
                lo[wd][p] += a[wd][p] & 0x0f0f0f0fu;
                hi[wd][p] += (a[wd][p] >> 4) & 0x0f0f0f0fu;
                a[wd][p] = 0;
            }
    };
    const uint4 zero = make_uint4(0, 0, 0, 0);
    for (int i = i0; i < i1; i += 15) {  // nibble fields hold 15 rows
        const int n15 = min(15, i1 - i);
        // the row numbers of the group: one coalesced read, handed out by shuffles
        int rid = 0;
        if (lane < n15) {
            const int tr = rows[i + lane];
            rid = g.rows_off >= 0 ? rows_arena[g.rows_off + tr] : tr;
        }
        for (int u0 = 0; u0 < n15; u0 += 5) {  // five independent 128-bit loads in flight per lane
            uint4 v[5];
#pragma unroll
            for (int u = 0; u < 5; ++u) {
                const int r = __shfl_sync(0xffffffffu, rid, min(u0 + u, 14));
                const uint4 *rw = reinterpret_cast<const uint4 *>(packed + g.base + (long long)r * g.stride +
                                                                  (long long)g.ch0 * 16);
                v[u] = (in_window && u0 + u < n15) ? __ldg(rw + ch) : zero;
            }
#pragma unroll
            for (int u = 0; u < 5; ++u) add(v[u]);
        }
        spill();
    }
    // the lane's 32 columns into the warp's slab (column 4 j + wd of chunk `lane`)
#pragma unroll
    for (int wd = 0; wd < 4; ++wd)
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int sh = (j >> 1) * 8;
            unsigned long long v = 0;
#pragma unroll
            for (int p = 0; p < 4; ++p)
                v |= (unsigned long long)((((j & 1) ? hi[wd][p] : lo[wd][p]) >> sh) & 255u) << (16 * p);
            s_cnt[warp][lane * 32 + 4 * j + wd] = v;
        }
    __syncthreads();
    const int w = g.c1 - g.c0;
    const int col0 = (g.ch0 + blockIdx.x * 32) * 32;  // first column of the group
    for (int k = threadIdx.x; k < 1024; k += 128) {
        const int col = col0 + k;
        if (col < g.c0 || col >= g.c1) continue;
        const unsigned long long v = s_cnt[0][k] + s_cnt[1][k] + s_cnt[2][k] + s_cnt[3][k];
        if (v) atomicAdd(&counts[(long long)c * w + (col - g.c0)], v);
    }
}

// one thread per (cluster, window word): the eight majority symbols of the word, as bytes (maj) and packed (majw)
__global__ void __launch_bounds__(128)
majority_pick_kernel(const ClusterState *__restrict__ states, int q, RefGrid g, const uint8_t *__restrict__ packed,
                     const int *__restrict__ rows_arena, const int *__restrict__ crows, const int *__restrict__ ccount,
                     const unsigned long long *__restrict__ counts, uint8_t *__restrict__ maj_all,
                     uint32_t *__restrict__ majw) {
    const ClusterState &st = states[q];
    if (st.status != 0) return;
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    const int c = blockIdx.y;
    if (c >= st.K || k >= g.n_words) return;
    const int nrows = ccount[c];
    const int w = g.c1 - g.c0;
    uint8_t *maj = maj_all + st.maj_off + (long long)c * w;
    const int col_base = (g.ch0 + (k >> 2)) * 32 + (k & 3);
    uint32_t word = 0;
    for (int j = 0; j < 8; ++j) {
        const int col = col_base + 4 * j;
        if (col < g.c0 || col >= g.c1 || nrows == 0) continue;
        const unsigned long long cn = counts[(long long)c * w + (col - g.c0)];
        const int b = (int)(cn & 0xffff), p1 = (int)((cn >> 16) & 0xffff), p2 = (int)((cn >> 32) & 0xffff),
                  p3 = (int)(cn >> 48);
        // symbol codes: gap 0, A 1, C 3, G 5, T 7
        int cnt[5] = {nrows - b, b - p1 - p2 + p3, p1 - p3, p2 - p3, p3};
        const int code[5] = {0, 1, 3, 5, 7};
        int best = 0, n_best = 1;
        for (int s = 1; s < 5; ++s) {
            if (cnt[s] > cnt[best]) {
                best = s;
                n_best = 1;
            } else if (cnt[s] == cnt[best]) {
                ++n_best;
            }
        }
        int sym = code[best];
        if (n_best > 1) {
            // Counter.most_common(1) on a tie: the symbol seen first among the member rows, in member order
            const int top = cnt[best];
            const int *rows = crows + (long long)c * g.R;
            for (int i = 0; i < nrows; ++i) {
                const uint32_t x = row_words(packed, g, rows_arena, rows[i])[k];
                const int s_code = (x >> (4 * j)) & 15;
                const int s_idx = s_code == 0 ? 0 : (s_code + 1) >> 1;  // 1 3 5 7 -> 1 2 3 4
                if (cnt[s_idx] == top) {
                    sym = s_code;
                    break;
                }
            }
        }
        maj[col - g.c0] = (uint8_t)sym;
        word |= (uint32_t)sym << (4 * j);
    }
    majw[(long long)c * g.n_words + k] = word;
}

// one warp per member row (rows of cluster 0, then cluster 1, ...): a row further than the threshold from the
// majority string of its cluster marks the cluster (and the problem) as not one-reference-like
__global__ void __launch_bounds__(128)
hamming_packed_kernel(const ClusterState *__restrict__ states, int q, RefGrid g, const uint8_t *__restrict__ packed,
                      const int *__restrict__ rows_arena, const int *__restrict__ crows, const int *__restrict__ ccount,
                      const uint32_t *__restrict__ majw, int *__restrict__ flags) {
    const ClusterState &st = states[q];
    if (st.status != 0) return;
    const int lane = threadIdx.x & 31;
    int i = blockIdx.x * 4 + (threadIdx.x >> 5);
    int c = 0;
    for (; c < st.K; ++c) {
        const int n = ccount[c];
        if (i < n) break;
        i -= n;
    }
    if (c >= st.K) return;
    const int w = g.c1 - g.c0;
    const int thr = w < 5 ? 1 : (int)(0.2 * (double)w);
    const uint4 *rw = reinterpret_cast<const uint4 *>(row_words(packed, g, rows_arena, crows[(long long)c * g.R + i]));
    const uint4 *mw = reinterpret_cast<const uint4 *>(majw + (long long)c * g.n_words);
    // nibbles of the first / last chunk that lie outside the window do not count
    const int lead = g.c0 & 31, tail = g.c1 & 31;
    const int n_chunks = g.n_words >> 2;
    auto differing = [&](uint32_t x, int k) {  // k = word of the window
        uint32_t z = (x | (x >> 1) | (x >> 2) | (x >> 3)) & NIB;
        if (k < 4 && lead) {
            // column of nibble j in this word: 4 j + (k & 3); keep those >= lead
            const int first_j = (lead - (k & 3) + 3) >> 2;  // smallest j with 4 j + (k & 3) >= lead
            z &= first_j >= 8 ? 0u : (0xffffffffu << (4 * max(first_j, 0)));
        }
        if (k >= g.n_words - 4 && tail) {
            const int wlane = k & 3;
            const int n_j = tail > wlane ? (tail - wlane + 3) >> 2 : 0;  // nibbles j with 4 j + wlane < tail
            z &= n_j >= 8 ? 0xffffffffu : ((1u << (4 * n_j)) - 1u);
        }
        return __popc(z);
    };
    int d = 0;
    for (int k0 = lane; k0 < n_chunks; k0 += 64) {  // two independent 128-bit loads in flight per lane
        const int k1 = k0 + 32;
        const uint4 x0 = __ldg(rw + k0);
        const uint4 x1 = k1 < n_chunks ? __ldg(rw + k1) : make_uint4(0, 0, 0, 0);
        const uint4 m0 = mw[k0];
        const uint4 m1 = k1 < n_chunks ? mw[k1] : make_uint4(0, 0, 0, 0);
        if (k0 == 0 || k0 == n_chunks - 1) {
            d += differing(x0.x ^ m0.x, 4 * k0) + differing(x0.y ^ m0.y, 4 * k0 + 1) + differing(x0.z ^ m0.z, 4 * k0 + 2) +
                 differing(x0.w ^ m0.w, 4 * k0 + 3);
        } else {
            const uint32_t e0 = x0.x ^ m0.x, e1 = x0.y ^ m0.y, e2 = x0.z ^ m0.z, e3 = x0.w ^ m0.w;
            d += __popc((e0 | (e0 >> 1) | (e0 >> 2) | (e0 >> 3)) & NIB) + __popc((e1 | (e1 >> 1) | (e1 >> 2) | (e1 >> 3)) & NIB) +
                 __popc((e2 | (e2 >> 1) | (e2 >> 2) | (e2 >> 3)) & NIB) + __popc((e3 | (e3 >> 1) | (e3 >> 2) | (e3 >> 3)) & NIB);
        }
        if (k1 < n_chunks) {
            if (k1 == n_chunks - 1) {
                d += differing(x1.x ^ m1.x, 4 * k1) + differing(x1.y ^ m1.y, 4 * k1 + 1) + differing(x1.z ^ m1.z, 4 * k1 + 2) +
                     differing(x1.w ^ m1.w, 4 * k1 + 3);
            } else {
                const uint32_t e0 = x1.x ^ m1.x, e1 = x1.y ^ m1.y, e2 = x1.z ^ m1.z, e3 = x1.w ^ m1.w;
                d += __popc((e0 | (e0 >> 1) | (e0 >> 2) | (e0 >> 3)) & NIB) + __popc((e1 | (e1 >> 1) | (e1 >> 2) | (e1 >> 3)) & NIB) +
                     __popc((e2 | (e2 >> 1) | (e2 >> 2) | (e2 >> 3)) & NIB) + __popc((e3 | (e3 >> 1) | (e3 >> 2) | (e3 >> 3)) & NIB);
            }
        }
    }
    d = __reduce_add_sync(0xffffffffu, d);
    if (d > thr && lane == 0) {
        flags[0] = 1;
        flags[1 + c] = 1;
    }
}

// scratch ints a problem of R rows and a window of w columns / n_words words needs (K_max clusters)
long long refgrid_scratch_ints(int R, int w, int n_words, int K_max) {
    return (long long)K_max * R + 16 + 2LL * K_max * w + 4 + (long long)K_max * n_words + 64;
}

cudaError_t launch_refcheck_grid(cudaStream_t s, ClusterState *states, int q, const DTask &t, int K_max,
                                 const uint8_t *packed, const int *rows_arena, const int *mem_off, const int *mem_rows,
                                 int *assign, uint8_t *maj, int *scratch, int n_member_rows, int *flags) {
    RefGrid g;
    g.base = t.base;
    g.stride = t.stride;
    g.rows_off = t.rows_off;
    g.R = t.n_rows;
    g.c0 = t.c0;
    g.c1 = t.c1;
    g.ch0 = t.c0 >> 5;
    g.n_words = (((t.c1 + 31) >> 5) - g.ch0) * 4;
    g.K_max = K_max;
    const int w = t.c1 - t.c0;
    int *crows = scratch;
    int *ccount = crows + (long long)K_max * g.R;
    // (16-byte aligned: the packed majority words that follow are read as 128-bit vectors)
    unsigned long long *counts = reinterpret_cast<unsigned long long *>(
        (reinterpret_cast<uintptr_t>(ccount + 16) + 15) & ~(uintptr_t)15);
    uint32_t *majw = reinterpret_cast<uint32_t *>(counts + (long long)K_max * w);
    cudaError_t e = cudaMemsetAsync(counts, 0, sizeof(unsigned long long) * (size_t)K_max * w, s);
    if (e != cudaSuccess) return e;
    cluster_rows_kernel<<<1, 32 * K_max, 0, s>>>(states, q, g, mem_off, mem_rows, assign, crows, ccount);
    const int col_groups = ((g.n_words >> 2) + 31) / 32;  // 32 chunks = 1,024 columns per CTA
    const int row_blocks = (n_member_rows + RG_ROWS - 1) / RG_ROWS + K_max;
    majority_count_kernel<<<dim3(col_groups, row_blocks), 128, 0, s>>>(states, q, g, packed, rows_arena, crows, ccount,
                                                                      counts);
    majority_pick_kernel<<<dim3((g.n_words + 127) / 128, K_max), 128, 0, s>>>(states, q, g, packed, rows_arena, crows,
                                                                             ccount, counts, maj, majw);
    hamming_packed_kernel<<<(n_member_rows + 3) / 4, 128, 0, s>>>(states, q, g, packed, rows_arena, crows, ccount, majw,
                                                                 flags);
    return cudaGetLastError();
}

}  // namespace mprg
