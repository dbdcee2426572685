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
//   majority_count_kernel one warp per (cluster, 128 rows, 256 columns): SWAR bit-plane counters in registers
//                         -- codes A C G T = 1 3 5 7 share bit 0, so four sums (b0, b0 b1, b0 b2, b0 b1 b2) give
//                         the five symbol counts -- nibble fields spilled to byte fields every 15 rows
//   majority_pick_kernel  arg-max per (cluster, column); exact ties resolved by reading the rows in order
//   hamming_packed_kernel one warp per member row: popcount of the non-zero nibbles of row XOR majority
#include "common.cuh"
#include "kernels.cuh"

namespace mprg {

constexpr int RG_ROWS = 128;   // rows per tile of the counting kernel (byte fields hold up to 255)
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
    for (int j0 = 0; j0 < st.n; j0 += 32) {
        const int j = j0 + lane;
        const bool mine = j < st.n && assign[j] == c;
        const int m0 = mine ? mem_off[j] : 0;
        const int cnt = mine ? mem_off[j + 1] - m0 : 0;
        int incl = cnt;
        for (int d = 1; d < 32; d <<= 1) {
            const int o = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += o;
        }
        const int pos = total + incl - cnt;
        for (int m = 0; m < cnt; ++m) out[pos + m] = mem_rows[m0 + m];
        total += __shfl_sync(0xffffffffu, incl, 31);
    }
    if (lane == 0) ccount[c] = total;
}

// counts[(c * w + col) * 4 + p]: p = 0 rows with a base, 1 bases with bit 1 (C or T), 2 bases with bit 2 (G or T),
// 3 bases with both (T)
__global__ void __launch_bounds__(128)
majority_count_kernel(const ClusterState *__restrict__ states, int q, RefGrid g, const uint8_t *__restrict__ packed,
                      const int *__restrict__ rows_arena, const int *__restrict__ crows, const int *__restrict__ ccount,
                      int *__restrict__ counts) {
    const ClusterState &st = states[q];
    if (st.status != 0) return;
    const int lane = threadIdx.x & 31;
    const int k = (blockIdx.x * 4 + (threadIdx.x >> 5)) * 32 + lane;  // word of the window
    // row block -> (cluster, first row)
    int c = 0, blk = blockIdx.y, nrows = 0;
    for (; c < st.K; ++c) {
        nrows = ccount[c];
        const int nb = (nrows + RG_ROWS - 1) / RG_ROWS;
        if (blk < nb) break;
        blk -= nb;
    }
    if (c >= st.K) return;
    const bool in_window = k < g.n_words;
    const int i0 = blk * RG_ROWS, i1 = min(i0 + RG_ROWS, nrows);
    const int *rows = crows + (long long)c * g.R;
    unsigned a0 = 0, a1 = 0, a2 = 0, a3 = 0;                                  // nibble fields
    unsigned l0 = 0, h0 = 0, l1 = 0, h1 = 0, l2 = 0, h2 = 0, l3 = 0, h3 = 0;  // byte fields (even / odd nibbles)
    int pending = 0;
    auto spill = [&]() {
        l0 += a0 & 0x0f0f0f0fu; h0 += (a0 >> 4) & 0x0f0f0f0fu;
        l1 += a1 & 0x0f0f0f0fu; h1 += (a1 >> 4) & 0x0f0f0f0fu;
        l2 += a2 & 0x0f0f0f0fu; h2 += (a2 >> 4) & 0x0f0f0f0fu;
        l3 += a3 & 0x0f0f0f0fu; h3 += (a3 >> 4) & 0x0f0f0f0fu;
        a0 = a1 = a2 = a3 = 0;
        pending = 0;
    };
    auto add = [&](unsigned x) {
        const unsigned t1 = x >> 1, t2 = x >> 2;
        const unsigned p1 = x & t1 & NIB;
        a0 += x & NIB;
        a1 += p1;
        a2 += x & t2 & NIB;
        a3 += p1 & t2;
    };
    int i = i0;
    for (; i + 5 <= i1; i += 5) {  // five independent row loads in flight per lane, three trips per spill
        unsigned x[5];
#pragma unroll
        for (int u = 0; u < 5; ++u) {
            const uint32_t *rw = row_words(packed, g, rows_arena, rows[i + u]);
            x[u] = in_window ? __ldg(rw + k) : 0u;
        }
#pragma unroll
        for (int u = 0; u < 5; ++u) add(x[u]);
        pending += 5;
        if (pending == 15) spill();
    }
    for (; i < i1; ++i) {
        const uint32_t *rw = row_words(packed, g, rows_arena, rows[i]);
        add(in_window ? __ldg(rw + k) : 0u);
        if (++pending == 15) spill();
    }
    spill();
    if (!in_window) return;
    const int w = g.c1 - g.c0;
    const int col_base = (g.ch0 + (k >> 2)) * 32 + (k & 3);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int col = col_base + 4 * j;
        if (col < g.c0 || col >= g.c1) continue;
        const int sh = (j >> 1) * 8;
        const unsigned v0 = (((j & 1) ? h0 : l0) >> sh) & 255u, v1 = (((j & 1) ? h1 : l1) >> sh) & 255u;
        const unsigned v2 = (((j & 1) ? h2 : l2) >> sh) & 255u, v3 = (((j & 1) ? h3 : l3) >> sh) & 255u;
        int *dst = counts + ((long long)c * w + (col - g.c0)) * 4;
        if (v0) atomicAdd(dst + 0, (int)v0);
        if (v1) atomicAdd(dst + 1, (int)v1);
        if (v2) atomicAdd(dst + 2, (int)v2);
        if (v3) atomicAdd(dst + 3, (int)v3);
    }
}

// one thread per (cluster, window word): the eight majority symbols of the word, as bytes (maj) and packed (majw)
__global__ void __launch_bounds__(128)
majority_pick_kernel(const ClusterState *__restrict__ states, int q, RefGrid g, const uint8_t *__restrict__ packed,
                     const int *__restrict__ rows_arena, const int *__restrict__ crows, const int *__restrict__ ccount,
                     const int *__restrict__ counts, uint8_t *__restrict__ maj_all, uint32_t *__restrict__ majw) {
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
        const int *cn = counts + ((long long)c * w + (col - g.c0)) * 4;
        const int b = cn[0], p1 = cn[1], p2 = cn[2], p3 = cn[3];
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
    const uint32_t *rw = row_words(packed, g, rows_arena, crows[(long long)c * g.R + i]);
    const uint32_t *mw = majw + (long long)c * g.n_words;
    // nibbles of the first / last chunk that lie outside the window do not count
    const int lead = g.c0 & 31, tail = g.c1 & 31;
    int d = 0;
    for (int k = lane; k < g.n_words; k += 32) {
        const uint32_t x = __ldg(rw + k) ^ mw[k];
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
        d += __popc(z);
    }
    d = __reduce_add_sync(0xffffffffu, d);
    if (d > thr && lane == 0) {
        flags[0] = 1;
        flags[1 + c] = 1;
    }
}

// scratch ints a problem of R rows and a window of w columns / n_words words needs (K_max clusters)
long long refgrid_scratch_ints(int R, int w, int n_words, int K_max) {
    return (long long)K_max * R + 16 + 4LL * K_max * w + (long long)K_max * n_words + 64;
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
    int *counts = ccount + 16;
    uint32_t *majw = reinterpret_cast<uint32_t *>(counts + 4LL * K_max * w);
    cudaError_t e = cudaMemsetAsync(counts, 0, sizeof(int) * 4 * (size_t)K_max * w, s);
    if (e != cudaSuccess) return e;
    cluster_rows_kernel<<<1, 32 * K_max, 0, s>>>(states, q, g, mem_off, mem_rows, assign, crows, ccount);
    const int col_groups = (g.n_words + 127) / 128;  // four warps of 32 words per CTA
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
