// Kernels (b) and the integer part of (c): everything kmeans_cluster_seqs does around KMeans.
//
//   unpack_kernel      gapped rows of a task -> one byte per symbol (scratch G, row-major)
//   dedupe_kernel      distinct ungapped / gapped rows in first-seen order
//                      (cluster_sequences.py:220-233, seq_utils.py:58-70); hashes are only a filter,
//                      every merge is verified symbol by symbol
//   kmer_kernel        ungapped k-mer extraction, first-occurrence column numbering and the dense
//                      count matrix (count_distinct_kmers / count_kmer_occurrences,
//                      cluster_sequences.py:26-56)
//   refcheck_kernel    cluster_further / sequences_are_one_reference_like (majority string with
//                      first-seen tie-break, Hamming distance, threshold; cluster_sequences.py:59-111)
//                      plus the loop control of kmeans_cluster_seqs (:256-261)
#include <algorithm>

#include "common.cuh"
#include "kernels.cuh"

namespace mprg {

__device__ __forceinline__ int sym_of(const uint8_t *row, int col) {
    return packed_sym(row, col);
}

__device__ __forceinline__ uint64_t mix64(uint64_t x) {
    x ^= x >> 33;
    x *= 0xff51afd7ed558ccdULL;
    x ^= x >> 33;
    x *= 0xc4ceb9fe1a85ec53ULL;
    x ^= x >> 33;
    return x;
}

// ---- unpack -------------------------------------------------------------------------------------
// one CTA (or, for deep tasks, gridDim.y CTAs) per task, warps over rows, lanes over columns
__global__ void __launch_bounds__(256)
unpack_kernel(const uint8_t *__restrict__ packed, const DTask *__restrict__ tasks,
              const int *__restrict__ rows_arena, const long long *__restrict__ g_off,
              uint8_t *__restrict__ G) {
    const DTask t = tasks[blockIdx.x];
    const int w = t.c1 - t.c0;
    const int *rows = t.rows_off >= 0 ? rows_arena + t.rows_off : nullptr;
    uint8_t *out = G + g_off[blockIdx.x];
    if (w >= 64) {
        // wide windows: warps over rows, lanes over columns
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
        for (int r = warp + nw * blockIdx.y; r < t.n_rows; r += nw * gridDim.y) {
            const uint8_t *row = packed + t.base + (long long)(rows ? rows[r] : r) * t.stride;
            for (int i = lane; i < w; i += 32) out[(long long)r * w + i] = (uint8_t)sym_of(row, t.c0 + i);
        }
    } else {
        // narrow windows (the non-match intervals of a pangenome level are a few columns wide): the
        // threads stride over the (row, column) pairs, so every lane has a symbol to fetch
        const long long total = (long long)t.n_rows * w;
        for (long long e = threadIdx.x + (long long)blockDim.x * blockIdx.y; e < total;
             e += (long long)blockDim.x * gridDim.y) {
            const int r = (int)(e / w), i = (int)(e - (long long)r * w);
            const uint8_t *row = packed + t.base + (long long)(rows ? rows[r] : r) * t.stride;
            out[e] = (uint8_t)sym_of(row, t.c0 + i);
        }
    }
}

// ---- helpers shared by the kernels below ----
// Per clustering problem (one CTA): the n distinct long sequences are rows seq_rows[0..n) (task-local
// row positions) of the task's unpacked block G.

__device__ __forceinline__ uint64_t kmer_hash(const uint8_t *p, int k) {
    uint64_t h = 0x243f6a8885a308d3ULL;
    for (int i = 0; i < k; ++i) h = (h ^ (uint64_t)(p[i] + 1)) * 0x9e3779b97f4a7c15ULL + (h >> 32);
    h = mix64(h);
    return h == ~0ULL ? 0 : h;  // ~0 marks an empty slot
}

// block-wide exclusive prefix sum over `n` ints in place (values small); returns the total
__device__ int block_exclusive_scan(int *a, int n, int *s_warp /* >= 33 ints */) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    __shared__ int carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < n; base += blockDim.x) {
        const int i = base + threadIdx.x;
        const int v = i < n ? a[i] : 0;
        int x = v;
        for (int d = 1; d < 32; d <<= 1) {
            const int o = __shfl_up_sync(0xffffffffu, x, d);
            if (lane >= d) x += o;
        }
        if (lane == 31) s_warp[warp] = x;
        __syncthreads();
        if (warp == 0) {
            int y = lane < nw ? s_warp[lane] : 0;
            for (int d = 1; d < 32; d <<= 1) {
                const int o = __shfl_up_sync(0xffffffffu, y, d);
                if (lane >= d) y += o;
            }
            s_warp[lane] = y;  // inclusive over warps
        }
        __syncthreads();
        const int warp_excl = warp ? s_warp[warp - 1] : 0;
        if (i < n) a[i] = carry + warp_excl + x - v;
        __syncthreads();
        if (threadIdx.x == 0) carry += s_warp[nw - 1];
        __syncthreads();
    }
    return carry;
}

// ---- dedupe ---------------------------------------------------------------------------------------
struct RowSig {
    uint64_t hu, hg;
    int len;
    int pad;
};

__device__ bool same_ungapped(const uint8_t *a, const uint8_t *b, int w) {
    int i = 0, j = 0;
    while (true) {
        while (i < w && a[i] == SYM_GAP) ++i;
        while (j < w && b[j] == SYM_GAP) ++j;
        if (i >= w || j >= w) return (i >= w) == (j >= w);
        if (a[i] != b[j]) return false;
        ++i;
        ++j;
    }
}

// Small tasks (most of a pangenome level: tens of thousands of tasks of a few columns) take one WARP each.
// Rows are taken 32 at a time in row order; a row is compared with the distinct rows found so far (a short
// list in shared memory whose position IS the first-seen group number) and, if new, with the other new rows
// of its 32 (match.any), so the work per row is O(#distinct) and there is no block barrier.  Same outputs as
// dedupe_kernel (group, ulen, leaders, leader_len, the counts).
constexpr int DW_ROWS = 512, DW_WARPS = 4;  // rows per task at most / tasks per CTA
constexpr int DW_LIST = 64;                 // distinct rows a warp keeps; a task with more goes to dedupe_kernel
constexpr int DW_OVERFLOW = -1;             // n_ungapped[task] of such a task after dedupe_warp_kernel
// one word into the running key: the full finaliser per word (a multiply and a fold alone let differences in
// the top nibbles of consecutive words cancel -- found by the verification below); still a bijection of x
__device__ __forceinline__ uint64_t word_step(uint64_t h, uint64_t x) { return mix64(h ^ x) + 0x9e3779b97f4a7c15ULL; }
__device__ __forceinline__ bool dedupe_small(int R, int w) { return R <= DW_ROWS && (long long)R * w <= 32768; }

__global__ void __launch_bounds__(DW_WARPS * 32)
dedupe_warp_kernel(const DTask *__restrict__ tasks, int n_tasks, const long long *__restrict__ g_off,
                   const uint8_t *__restrict__ G, const long long *__restrict__ row_off, int *__restrict__ group,
                   int *__restrict__ ulen, int *__restrict__ leaders, int *__restrict__ leader_len,
                   int *__restrict__ n_ungapped, int *__restrict__ n_gapped, int *__restrict__ err) {
    __shared__ uint64_t s_ku[DW_WARPS][DW_LIST], s_kg[DW_WARPS][DW_LIST];  // keys of the distinct rows so far
    __shared__ int s_ru[DW_WARPS][DW_LIST], s_rg[DW_WARPS][DW_LIST];        // ... and their rows
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int ti = blockIdx.x * DW_WARPS + warp;
    if (ti >= n_tasks) return;
    const DTask t = tasks[ti];
    const int w = t.c1 - t.c0, R = t.n_rows;
    if (!dedupe_small(R, w)) return;
    const uint8_t *g = G + g_off[ti];
    const long long ro = row_off[ti];
    uint64_t *ku = s_ku[warp], *kg = s_kg[warp];
    int *ru = s_ru[warp], *rg = s_rg[warp];
    int nu = 0, ng = 0;
    for (int r0 = 0; r0 < R; r0 += 32) {
        const int r = r0 + lane;
        const bool valid = r < R;
        const uint8_t *row = g + (long long)r * w;
        // keys over 16-symbol words of 4-bit codes (one 64-bit mixing step per word, not per symbol).  Every
        // step is a bijection of the word, so for w <= 16 the keys ARE the rows and need no verification
        // (non-gap codes are non-zero: the ungapped word also fixes the length).
        uint64_t hu = 0x9e3779b97f4a7c15ULL, hg = 0x2545f4914f6cdd1dULL, wu = 0, wg = 0;
        int len = 0;
        if (valid) {
            for (int i = 0; i < w; ++i) {
                const uint64_t c = row[i];
                wg |= c << (4 * (i & 15));
                if ((i & 15) == 15) {
                    hg = word_step(hg, wg);
                    wg = 0;
                }
                if (c != SYM_GAP) {
                    wu |= c << (4 * (len & 15));
                    if ((++len & 15) == 0) {
                        hu = word_step(hu, wu);
                        wu = 0;
                    }
                }
            }
            ulen[ro + r] = len;
        }
        const bool exact = w <= 16;
        const uint64_t mu = word_step(hu, wu) ^ (exact ? 0ULL : (uint64_t)len * 0xd6e8feb86659fd93ULL);
        const uint64_t mg = word_step(hg, wg);
        // ---- ungapped: group number = position of the row's distinct sequence in first-seen order ----
        int gi = -1;
        if (valid)
            for (int j = 0; j < nu; ++j)
                if (ku[j] == mu) {
                    gi = j;
                    break;
                }
        {
            const bool fresh = valid && gi < 0;
            const unsigned fresh_mask = __ballot_sync(0xffffffffu, fresh);
            const unsigned peers = __match_any_sync(0xffffffffu, mu) & fresh_mask;
            const int first = __ffs(peers) - 1;  // lowest fresh lane with my key
            const bool lead = fresh && first == lane;
            const unsigned lead_mask = __ballot_sync(0xffffffffu, lead);
            if (nu + __popc(lead_mask) > DW_LIST) {  // more distinct rows than the list holds: dedupe_kernel redoes it
                if (lane == 0) n_ungapped[ti] = DW_OVERFLOW;
                return;
            }
            if (fresh) gi = nu + __popc(lead_mask & ((1u << (first & 31)) - 1u));
            if (lead) {
                ku[gi] = mu;
                ru[gi] = r;
                leaders[ro + gi] = r;
                leader_len[ro + gi] = len;
            }
            nu += __popc(lead_mask);
            __syncwarp();
        }
        // ---- gapped: only the number of distinct rows is kept ----
        int gj = -1;
        if (valid)
            for (int j = 0; j < ng; ++j)
                if (kg[j] == mg) {
                    gj = j;
                    break;
                }
        {
            const bool fresh = valid && gj < 0;
            const unsigned fresh_mask = __ballot_sync(0xffffffffu, fresh);
            const unsigned peers = __match_any_sync(0xffffffffu, mg) & fresh_mask;
            const int first = __ffs(peers) - 1;
            const bool lead = fresh && first == lane;
            const unsigned lead_mask = __ballot_sync(0xffffffffu, lead);
            if (ng + __popc(lead_mask) > DW_LIST) {
                if (lane == 0) n_ungapped[ti] = DW_OVERFLOW;
                return;
            }
            if (fresh) gj = ng + __popc(lead_mask & ((1u << (first & 31)) - 1u));
            if (lead) {
                kg[gj] = mg;
                rg[gj] = r;
            }
            ng += __popc(lead_mask);
            __syncwarp();
        }
        if (valid) {
            group[ro + r] = gi;
            // verify the merges exactly (key equality is only a filter)
            const int a = ru[gi], b = rg[gj];
#ifdef MPRG_DEDUPE_DEBUG
            if (!exact && a != r && !same_ungapped(g + (long long)a * w, row, w))
                printf("U ti %d r %d a %d w %d R %d len %d gi %d nu %d key %llx akey %llx\n", ti, r, a, w, R, len, gi, nu,
                       (unsigned long long)mu, (unsigned long long)ku[gi]);
#endif
            if (!exact && a != r && !same_ungapped(g + (long long)a * w, row, w)) atomicExch(err, 2);
            if (!exact && b != r) {
                const uint8_t *x = g + (long long)b * w;
                bool eq = true;
                for (int i = 0; i < w && eq; ++i) eq = x[i] == row[i];
#ifdef MPRG_DEDUPE_DEBUG
                if (!eq)
                {
                    printf("G ti %d r %d b %d w %d R %d gj %d ng %d key %llx bkey %llx\n", ti, r, b, w, R, gj, ng,
                           (unsigned long long)mg, (unsigned long long)kg[gj]);
                    for (int i = 0; i < w; ++i) printf("%d:%d/%d ", i, (int)x[i], (int)row[i]);
                    printf("\n");
                }
#endif
                if (!eq) atomicExch(err, 2);
            }
        }
    }
    if (lane == 0) {
        n_ungapped[ti] = nu;
        n_gapped[ti] = ng;
    }
}

// one CTA per task.  Outputs per row (at row_off[t]): group = index of the row's distinct ungapped
// sequence in first-seen order, ulen = ungapped length; per task the distinct counts.
__device__ void dedupe_task(int ti, const DTask &t, const long long *__restrict__ g_off,
                            const uint8_t *__restrict__ G, const long long *__restrict__ row_off,
                            RowSig *__restrict__ sig, int *__restrict__ leader_u, int *__restrict__ leader_g,
                            int *__restrict__ group, int *__restrict__ ulen, int *__restrict__ leaders,
                            int *__restrict__ leader_len, int *__restrict__ n_ungapped, int *__restrict__ n_gapped,
                            int *__restrict__ err) {
    const int w = t.c1 - t.c0, R = t.n_rows;
    const uint8_t *g = G + g_off[ti];
    const long long ro = row_off[ti];
    RowSig *s = sig + ro;
    int *lu = leader_u + ro, *lg = leader_g + ro;
    for (int r = threadIdx.x; r < R; r += blockDim.x) {
        const uint8_t *row = g + (long long)r * w;
        uint64_t hu = 0x9e3779b97f4a7c15ULL, hg = 0x2545f4914f6cdd1dULL;
        int len = 0;
        for (int i = 0; i < w; ++i) {
            const uint64_t c = row[i];
            hg = (hg ^ (c + 1)) * 0x100000001b3ULL;
            hg ^= hg >> 29;
            if (c != SYM_GAP) {
                hu = (hu ^ (c + 1)) * 0x9fb21c651e98df25ULL;
                hu ^= hu >> 31;
                ++len;
            }
        }
        s[r] = RowSig{mix64(hu), mix64(hg), len, 0};
        ulen[ro + r] = len;
    }
    __syncthreads();
    for (int r = threadIdx.x; r < R; r += blockDim.x) {
        const RowSig me = s[r];
        int a = r, b = r;
        for (int q = 0; q < r; ++q) {
            const RowSig o = s[q];
            if (a == r && o.hu == me.hu && o.len == me.len) a = q;
            if (b == r && o.hg == me.hg) b = q;
            if (a != r && b != r) break;
        }
        // verify the merges exactly (hash equality is only a filter)
        if (a != r && !same_ungapped(g + (long long)a * w, g + (long long)r * w, w)) atomicExch(err, 2);
        if (b != r) {
            const uint8_t *x = g + (long long)b * w, *y = g + (long long)r * w;
            bool eq = true;
            for (int i = 0; i < w && eq; ++i) eq = x[i] == y[i];
            if (!eq) atomicExch(err, 2);
        }
        lu[r] = a;
        lg[r] = b;
    }
    __syncthreads();
    // first-seen numbering of the distinct rows: exclusive prefix sum of the leader flags
    __shared__ int s_warp[33];
    __shared__ int s_ng;
    if (threadIdx.x == 0) s_ng = 0;
    int ng = 0;
    for (int r = threadIdx.x; r < R; r += blockDim.x) {
        group[ro + r] = lu[r] == r ? 1 : 0;
        ng += lg[r] == r;
    }
    __syncthreads();
    if (ng) atomicAdd(&s_ng, ng);
    const int nu = block_exclusive_scan(group + ro, R, s_warp);
    for (int r = threadIdx.x; r < R; r += blockDim.x)
        if (lu[r] == r) {
            leaders[ro + group[ro + r]] = r;
            leader_len[ro + group[ro + r]] = ulen[ro + r];
        }
    __syncthreads();
    for (int r = threadIdx.x; r < R; r += blockDim.x) {
        const int a = lu[r];
        if (a != r) group[ro + r] = group[ro + a];  // a is a leader (a < r): its entry is final
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        n_ungapped[ti] = nu;
        n_gapped[ti] = s_ng;
    }
}

// The CTAs walk the task list (most entries are small tasks that dedupe_warp_kernel has done, or empty slots of
// the cluster-task table: a CTA per entry was tens of thousands of CTAs that left at once).
__global__ void __launch_bounds__(256)
dedupe_kernel(const DTask *__restrict__ tasks, int n_tasks, const long long *__restrict__ g_off,
              const uint8_t *__restrict__ G, const long long *__restrict__ row_off,
              RowSig *__restrict__ sig, int *__restrict__ leader_u, int *__restrict__ leader_g,
              int *__restrict__ group, int *__restrict__ ulen, int *__restrict__ leaders,
              int *__restrict__ leader_len, int *__restrict__ n_ungapped, int *__restrict__ n_gapped,
              int *__restrict__ err) {
    for (int ti = blockIdx.x; ti < n_tasks; ti += gridDim.x) {
        const DTask t = tasks[ti];
        if (dedupe_small(t.n_rows, t.c1 - t.c0) && n_ungapped[ti] != DW_OVERFLOW) continue;  // dedupe_warp_kernel did it
        dedupe_task(ti, t, g_off, G, row_off, sig, leader_u, leader_g, group, ulen, leaders, leader_len, n_ungapped,
                    n_gapped, err);
        __syncthreads();
    }
}

// ---- dedupe of deep tasks -------------------------------------------------------------------------
// The same outputs as dedupe_kernel from four launches that spread every task over the grid (one CTA
// per task is the wrong shape for a 10,000 x 20,000 task): signatures with one warp per row (the hashes
// are sums of mixed (position, symbol) words, so the lanes of a warp add up their columns in any order;
// the ungapped sequence of the row is compacted next to it for the exact check), first-seen search
// with one thread per row, exact verification of every merge with one warp per row, and the
// first-seen numbering of the distinct rows by a block-wide prefix sum.
constexpr uint64_t SALT_U = 0x9e3779b97f4a7c15ULL, SALT_G = 0x2545f4914f6cdd1dULL;

__global__ void __launch_bounds__(256)
dedupe_big_sig_kernel(const DTask *__restrict__ tasks, const long long *__restrict__ g_off,
                      const uint8_t *__restrict__ G, const long long *__restrict__ row_off,
                      RowSig *__restrict__ sig, int *__restrict__ ulen, uint8_t *__restrict__ U) {
    const int ti = blockIdx.y;
    const DTask t = tasks[ti];
    const int w = t.c1 - t.c0, R = t.n_rows;
    const int lane = threadIdx.x & 31;
    for (int r = blockIdx.x * 8 + (threadIdx.x >> 5); r < R; r += gridDim.x * 8) {  // warp-uniform
        const uint8_t *row = G + g_off[ti] + (long long)r * w;
        uint8_t *urow = U + g_off[ti] + (long long)r * w;
        uint64_t hu = 0, hg = 0;
        int base = 0;
        for (int i0 = 0; i0 < w; i0 += 32) {
            const int i = i0 + lane;
            const bool in = i < w;
            const uint32_t c = in ? row[i] : (uint32_t)SYM_GAP;
            if (in) hg += mix64((((uint64_t)i << 8) | c) ^ SALT_G);
            const bool solid = in && c != SYM_GAP;
            const uint32_t m = __ballot_sync(0xffffffffu, solid);
            if (solid) {
                const int u = base + __popc(m & ((1u << lane) - 1u));
                hu += mix64((((uint64_t)u << 8) | c) ^ SALT_U);
                urow[u] = (uint8_t)c;
            }
            base += __popc(m);
        }
        for (int d = 16; d; d >>= 1) {
            hu += __shfl_xor_sync(0xffffffffu, hu, d);
            hg += __shfl_xor_sync(0xffffffffu, hg, d);
        }
        if (lane == 0) {
            sig[row_off[ti] + r] = RowSig{mix64(hu ^ SALT_U), mix64(hg ^ SALT_G), base, 0};
            ulen[row_off[ti] + r] = base;
        }
    }
}

__global__ void __launch_bounds__(256)
dedupe_big_match_kernel(const DTask *__restrict__ tasks, const long long *__restrict__ row_off,
                        const RowSig *__restrict__ sig, int *__restrict__ leader_u, int *__restrict__ leader_g) {
    const int ti = blockIdx.y;
    const int R = tasks[ti].n_rows;
    const long long ro = row_off[ti];
    const RowSig *s = sig + ro;
    for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < R; r += gridDim.x * blockDim.x) {
        const RowSig me = s[r];
        int a = r, b = r;
        for (int q = 0; q < r; ++q) {
            const RowSig o = s[q];
            if (a == r && o.hu == me.hu && o.len == me.len) a = q;
            if (b == r && o.hg == me.hg) b = q;
            if (a != r && b != r) break;
        }
        leader_u[ro + r] = a;
        leader_g[ro + r] = b;
    }
}

__global__ void __launch_bounds__(256)
dedupe_big_verify_kernel(const DTask *__restrict__ tasks, const long long *__restrict__ g_off,
                         const uint8_t *__restrict__ G, const uint8_t *__restrict__ U,
                         const long long *__restrict__ row_off, const int *__restrict__ leader_u,
                         const int *__restrict__ leader_g, const int *__restrict__ ulen, int *__restrict__ err) {
    const int ti = blockIdx.y;
    const DTask t = tasks[ti];
    const int w = t.c1 - t.c0, R = t.n_rows;
    const int lane = threadIdx.x & 31;
    const long long ro = row_off[ti];
    for (int r = blockIdx.x * 8 + (threadIdx.x >> 5); r < R; r += gridDim.x * 8) {
        const int a = leader_u[ro + r], b = leader_g[ro + r];
        bool bad = false;
        if (a != r) {
            const int len = ulen[ro + r];
            const uint8_t *x = U + g_off[ti] + (long long)a * w, *y = U + g_off[ti] + (long long)r * w;
            bad |= ulen[ro + a] != len;
            for (int i = lane; i < len && !bad; i += 32) bad |= x[i] != y[i];
        }
        if (b != r) {
            const uint8_t *x = G + g_off[ti] + (long long)b * w, *y = G + g_off[ti] + (long long)r * w;
            for (int i = lane; i < w && !bad; i += 32) bad |= x[i] != y[i];
        }
        if (bad) atomicExch(err, 2);  // hash equality is only a filter: a false merge fails the call
    }
}

__global__ void __launch_bounds__(256)
dedupe_big_number_kernel(const DTask *__restrict__ tasks, const long long *__restrict__ row_off,
                         const int *__restrict__ leader_u, const int *__restrict__ leader_g,
                         const int *__restrict__ ulen, int *__restrict__ group, int *__restrict__ leaders,
                         int *__restrict__ leader_len, int *__restrict__ n_ungapped, int *__restrict__ n_gapped) {
    __shared__ int s_warp[33];
    __shared__ int s_ng;
    const int ti = blockIdx.x;
    const int R = tasks[ti].n_rows;
    const long long ro = row_off[ti];
    if (threadIdx.x == 0) s_ng = 0;
    int ng = 0;
    for (int r = threadIdx.x; r < R; r += blockDim.x) {
        group[ro + r] = leader_u[ro + r] == r ? 1 : 0;
        ng += leader_g[ro + r] == r;
    }
    __syncthreads();
    atomicAdd(&s_ng, ng);
    const int nu = block_exclusive_scan(group + ro, R, s_warp);  // leaders: their first-seen index
    for (int r = threadIdx.x; r < R; r += blockDim.x)
        if (leader_u[ro + r] == r) {
            leaders[ro + group[ro + r]] = r;
            leader_len[ro + group[ro + r]] = ulen[ro + r];
        }
    __syncthreads();
    for (int r = threadIdx.x; r < R; r += blockDim.x) {
        const int a = leader_u[ro + r];
        if (a != r) group[ro + r] = group[ro + a];  // leaders come first (a < r) and are never rewritten
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        n_ungapped[ti] = nu;
        n_gapped[ti] = s_ng;
    }
}

__global__ void __launch_bounds__(256)
kmer_kernel(const KmerProb *__restrict__ probs, const int *__restrict__ seq_rows,
            const uint8_t *__restrict__ G, int k, uint8_t *__restrict__ useq_all,
            int *__restrict__ ints_all, uint64_t *__restrict__ keys_all, int *__restrict__ ming_all,
            int *__restrict__ out_F, int *__restrict__ err) {
    __shared__ int s_warp[33];
    const KmerProb p = probs[blockIdx.x];
    if (p.big & 1) return;  // launch_kmer_big
    const uint8_t *g = G + p.g_off;
    const int *srow = seq_rows + p.seq_off;
    uint8_t *useq = useq_all + p.useq_off;
    int *ulen = ints_all + p.pos_off;
    int *pos = ulen + p.n;
    int *mref = pos + p.n + 1;
    int *kid = mref + p.Pmax;
    uint64_t *keys = keys_all + p.tab_off;
    int *ming = ming_all + p.tab_off;
    const int w = p.w, n = p.n;

    for (int j = threadIdx.x; j < n; j += blockDim.x) {
        const uint8_t *row = g + (long long)srow[j] * w;
        uint8_t *u = useq + (long long)j * w;
        int len = 0;
        for (int i = 0; i < w; ++i)
            if (row[i] != SYM_GAP) u[len++] = row[i];
        ulen[j] = len;
    }
    for (int i = threadIdx.x; i < p.T; i += blockDim.x) {
        keys[i] = ~0ULL;
        ming[i] = 0x7fffffff;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int acc = 0;
        for (int j = 0; j < n; ++j) {
            pos[j] = acc;
            acc += ulen[j] - k + 1;  // every sequence here has ulen >= k
        }
        pos[n] = acc;
    }
    __syncthreads();
    const int P = pos[n];
    const int mask = p.T - 1;
    // insert: slot key claimed by CAS, first-occurrence position by atomicMin
    for (int gidx = threadIdx.x; gidx < P; gidx += blockDim.x) {
        int lo = 0, hi = n;  // pos[lo] <= gidx < pos[hi]
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (pos[mid] <= gidx) lo = mid; else hi = mid;
        }
        const uint8_t *km = useq + (long long)lo * w + (gidx - pos[lo]);
        const uint64_t key = kmer_hash(km, k);
        int slot = (int)(key & mask);
        while (true) {
            const uint64_t old = atomicCAS((unsigned long long *)&keys[slot], ~0ULL, key);
            if (old == ~0ULL || old == key) {
                atomicMin(&ming[slot], gidx);
                break;
            }
            slot = (slot + 1) & mask;
        }
        mref[gidx] = slot;
    }
    __syncthreads();
    for (int gidx = threadIdx.x; gidx < P; gidx += blockDim.x) {
        const int first = ming[mref[gidx]];
        kid[gidx] = first == gidx ? 1 : 0;
        mref[gidx] = first;
        if (first != gidx) {  // verify: same hash must mean same k-mer
            int lo = 0, hi = n;
            while (hi - lo > 1) {
                const int mid = (lo + hi) >> 1;
                if (pos[mid] <= gidx) lo = mid; else hi = mid;
            }
            const uint8_t *a = useq + (long long)lo * w + (gidx - pos[lo]);
            int lo2 = 0, hi2 = n;
            while (hi2 - lo2 > 1) {
                const int mid = (lo2 + hi2) >> 1;
                if (pos[mid] <= first) lo2 = mid; else hi2 = mid;
            }
            const uint8_t *b = useq + (long long)lo2 * w + (first - pos[lo2]);
            bool eq = true;
            for (int i = 0; i < k && eq; ++i) eq = a[i] == b[i];
            if (!eq) atomicExch(err, 3);
        }
    }
    __syncthreads();
    const int F = block_exclusive_scan(kid, P, s_warp);  // kid[g] = column id when g is a first occurrence
    if (threadIdx.x == 0) out_F[blockIdx.x] = F;
}

// Second phase, once the host has sized the count matrices exactly (n x F doubles per problem, zeroed):
// X[sequence of position g][column of the k-mer at g] += 1.  grid (problems, stripes of positions).
__global__ void __launch_bounds__(256)
kmer_fill_kernel(const KmerProb *__restrict__ probs, const int *__restrict__ ints_all,
                 const int *__restrict__ F_all, double *__restrict__ X_all) {
    const KmerProb p = probs[blockIdx.x];
    if (p.big & 2) return;  // launch_kmer_fill_big
    const int n = p.n, F = F_all[blockIdx.x];
    const int *pos = ints_all + p.pos_off + p.n;
    const int *mref = pos + p.n + 1;
    const int *kid = mref + p.Pmax;
    double *X = X_all + p.x_off;
    const int P = pos[n];
    for (int gidx = blockIdx.y * blockDim.x + threadIdx.x; gidx < P; gidx += gridDim.y * blockDim.x) {
        int lo = 0, hi = n;
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (pos[mid] <= gidx) lo = mid; else hi = mid;
        }
        atomicAdd(&X[(long long)lo * F + kid[mref[gidx]]], 1.0);
    }
}

// ---- big problems: the same numbering and counting with the whole grid ---------------------------
// One problem at a time (deep loci: 10^4 sequences x 2*10^4 columns = 2*10^8 k-mer positions).  Same
// results as kmer_kernel: column ids in first-occurrence order, exact verification of every hash match.
constexpr int KB_THREADS = 256;
constexpr int KB_CHUNK = 4096;  // positions per CTA in the numbering passes

__device__ __forceinline__ int kb_seq_of(const int *pos, int n, int gidx) {
    int lo = 0, hi = n;  // pos[lo] <= gidx < pos[hi]
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (pos[mid] <= gidx) lo = mid; else hi = mid;
    }
    return lo;
}

// one warp per sequence: ordered compaction of the non-gap symbols, 32 columns at a time
__global__ void __launch_bounds__(KB_THREADS)
kmer_big_ungap_kernel(const KmerProb *__restrict__ probs, int q, const int *__restrict__ seq_rows,
                      const uint8_t *__restrict__ G, uint8_t *__restrict__ useq_all, int *__restrict__ ints_all) {
    const KmerProb p = probs[q];
    const int lane = threadIdx.x & 31;
    const int j = blockIdx.x * (KB_THREADS / 32) + (threadIdx.x >> 5);
    if (j >= p.n) return;
    const uint8_t *row = G + p.g_off + (long long)seq_rows[p.seq_off + j] * p.w;
    uint8_t *u = useq_all + p.useq_off + (long long)j * p.w;
    int len = 0;
    for (int c0 = 0; c0 < p.w; c0 += 32) {
        const int c = c0 + lane;
        const int sym = c < p.w ? row[c] : SYM_GAP;
        const unsigned keep = __ballot_sync(0xffffffffu, sym != SYM_GAP);
        if (sym != SYM_GAP) u[len + __popc(keep & ((1u << lane) - 1u))] = (uint8_t)sym;
        len += __popc(keep);
    }
    if (lane == 0) (ints_all + p.pos_off)[j] = len;
}

// one CTA: pos = exclusive prefix of (ulen - k + 1), table reset is done by the host (memset)
__global__ void __launch_bounds__(KB_THREADS)
kmer_big_prefix_kernel(const KmerProb *__restrict__ probs, int q, int k, int *__restrict__ ints_all) {
    __shared__ int s_warp[33];
    const KmerProb p = probs[q];
    int *ulen = ints_all + p.pos_off;
    int *pos = ulen + p.n;
    for (int j = threadIdx.x; j < p.n; j += blockDim.x) pos[j] = ulen[j] - k + 1;
    __syncthreads();
    const int total = block_exclusive_scan(pos, p.n, s_warp);
    if (threadIdx.x == 0) pos[p.n] = total;
}

// every position: claim / find the slot of its k-mer, keep the smallest position per slot.  The table
// is read before any atomic: once a k-mer is in and its first occurrence is known, later occurrences
// cost two cached loads and no atomic (10^8 positions hit ~10^4 hot slots on a deep locus).
__global__ void __launch_bounds__(KB_THREADS)
kmer_big_insert_kernel(const KmerProb *__restrict__ probs, int q, int k, const uint8_t *__restrict__ useq_all,
                       int *__restrict__ ints_all, uint64_t *__restrict__ keys_all, int *__restrict__ ming_all) {
    const KmerProb p = probs[q];
    const int *pos = ints_all + p.pos_off + p.n;
    int *mref = ints_all + p.pos_off + 2 * p.n + 1;
    const uint8_t *useq = useq_all + p.useq_off;
    uint64_t *keys = keys_all + p.tab_off;
    int *ming = ming_all + p.tab_off;
    const int P = pos[p.n], mask = p.T - 1;
    for (int gidx = blockIdx.x * blockDim.x + threadIdx.x; gidx < P; gidx += gridDim.x * blockDim.x) {
        const int j = kb_seq_of(pos, p.n, gidx);
        const uint64_t key = kmer_hash(useq + (long long)j * p.w + (gidx - pos[j]), k);
        int slot = (int)(key & mask);
        while (true) {
            uint64_t cur = __ldcg((const unsigned long long *)&keys[slot]);
            if (cur == ~0ULL) cur = atomicCAS((unsigned long long *)&keys[slot], ~0ULL, key);
            if (cur == ~0ULL || cur == key) break;
            slot = (slot + 1) & mask;
        }
        if (__ldcg(&ming[slot]) > gidx) atomicMin(&ming[slot], gidx);
        mref[gidx] = slot;
    }
}

// mref[g] <- first position of g's k-mer (verified symbol by symbol); per chunk the number of first
// occurrences
__global__ void __launch_bounds__(KB_THREADS)
kmer_big_first_kernel(const KmerProb *__restrict__ probs, int q, int k, const uint8_t *__restrict__ useq_all,
                      int *__restrict__ ints_all, const int *__restrict__ ming_all, int *__restrict__ block_counts,
                      int *__restrict__ err) {
    __shared__ int s_count;
    const KmerProb p = probs[q];
    const int *pos = ints_all + p.pos_off + p.n;
    int *mref = ints_all + p.pos_off + 2 * p.n + 1;
    const uint8_t *useq = useq_all + p.useq_off;
    const int *ming = ming_all + p.tab_off;
    const int P = pos[p.n];
    if (threadIdx.x == 0) s_count = 0;
    __syncthreads();
    int mine = 0;
    const int g0 = blockIdx.x * KB_CHUNK;
    for (int gidx = g0 + threadIdx.x; gidx < min(g0 + KB_CHUNK, P); gidx += blockDim.x) {
        const int first = ming[mref[gidx]];
        mref[gidx] = first;
        if (first == gidx) {
            ++mine;
        } else {
            const int ja = kb_seq_of(pos, p.n, gidx), jb = kb_seq_of(pos, p.n, first);
            const uint8_t *a = useq + (long long)ja * p.w + (gidx - pos[ja]);
            const uint8_t *b = useq + (long long)jb * p.w + (first - pos[jb]);
            bool eq = true;
            for (int i = 0; i < k && eq; ++i) eq = a[i] == b[i];
            if (!eq) atomicExch(err, 3);
        }
    }
    if (mine) atomicAdd(&s_count, mine);
    __syncthreads();
    if (threadIdx.x == 0) block_counts[blockIdx.x] = s_count;
}

// one CTA: exclusive scan of the chunk counts, total = F
__global__ void __launch_bounds__(KB_THREADS)
kmer_big_scan_kernel(int *__restrict__ block_counts, int n_chunks, int *__restrict__ out_F) {
    __shared__ int s_warp[33];
    const int F = block_exclusive_scan(block_counts, n_chunks, s_warp);
    if (threadIdx.x == 0) *out_F = F;
}

// kid[g] = column id for every first occurrence g (first-occurrence order = position order)
__global__ void __launch_bounds__(KB_THREADS)
kmer_big_number_kernel(const KmerProb *__restrict__ probs, int q, int *__restrict__ ints_all,
                       const int *__restrict__ block_counts) {
    __shared__ int s_warp[KB_THREADS / 32];
    __shared__ int s_base;
    const KmerProb p = probs[q];
    const int *pos = ints_all + p.pos_off + p.n;
    const int *mref = ints_all + p.pos_off + 2 * p.n + 1;
    int *kid = ints_all + p.pos_off + 2 * p.n + 1 + p.Pmax;
    const int P = pos[p.n];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_base = block_counts[blockIdx.x];
    __syncthreads();
    const int g0 = blockIdx.x * KB_CHUNK;
    for (int base = g0; base < min(g0 + KB_CHUNK, P); base += KB_THREADS) {
        const int gidx = base + threadIdx.x;
        const bool first = gidx < P && gidx < g0 + KB_CHUNK && mref[gidx] == gidx;
        const unsigned b = __ballot_sync(0xffffffffu, first);
        if (lane == 0) s_warp[warp] = __popc(b);
        __syncthreads();
        int before = 0, total = 0;
        for (int w2 = 0; w2 < KB_THREADS / 32; ++w2) {
            if (w2 < warp) before += s_warp[w2];
            total += s_warp[w2];
        }
        if (first) kid[gidx] = s_base + before + __popc(b & ((1u << lane) - 1u));
        __syncthreads();
        if (threadIdx.x == 0) s_base += total;
        __syncthreads();
    }
}

// counts of a big problem: one CTA per sequence, histogram in shared memory (F int counters), then the
// row goes out as doubles in one coalesced sweep
__global__ void __launch_bounds__(KB_THREADS)
kmer_big_fill_kernel(const KmerProb *__restrict__ probs, int q, int F, const int *__restrict__ ints_all,
                     double *__restrict__ X_all) {
    extern __shared__ int s_hist[];
    const KmerProb p = probs[q];
    const int *pos = ints_all + p.pos_off + p.n;
    const int *mref = ints_all + p.pos_off + 2 * p.n + 1;
    const int *kid = mref + p.Pmax;
    for (int j = blockIdx.x; j < p.n; j += gridDim.x) {
        for (int f = threadIdx.x; f < F; f += blockDim.x) s_hist[f] = 0;
        __syncthreads();
        for (int gidx = pos[j] + threadIdx.x; gidx < pos[j + 1]; gidx += blockDim.x)
            atomicAdd(&s_hist[kid[mref[gidx]]], 1);
        __syncthreads();
        double *x = X_all + p.x_off + (long long)j * F;
        for (int f = threadIdx.x; f < F; f += blockDim.x) x[f] = (double)s_hist[f];
        __syncthreads();
    }
}

// ---- one-reference-like check + loop control ------------------------------------------------------
__global__ void __launch_bounds__(128)
refcheck_kernel(ClusterState *__restrict__ states, const uint8_t *__restrict__ G,
                const int *__restrict__ mem_off_all, const int *__restrict__ mem_rows_all,
                int *__restrict__ assign_all, uint8_t *__restrict__ maj_all, int max_clusters,
                int *__restrict__ flags_out) {
    constexpr int REF_ROWS = 1024;
    __shared__ int cnt_s[16][128];
    __shared__ int first_s[16][128];
    __shared__ int s_rows[REF_ROWS];
    __shared__ int s_bad, s_bad_c;
    ClusterState &st = states[blockIdx.x];
    if (st.status != 0 || (st.big_ref && !flags_out)) return;  // big problems: launch_refcheck_big
    const uint8_t *g = G + st.g_off;
    const int w = st.w, n = st.n, K = st.K;
    const int *mem_off = mem_off_all + st.mem_off;   // n + 1 entries, into mem_rows
    const int *mem_rows = mem_rows_all + st.mem_rows_off;
    const int *assign = assign_all + st.assign_off;
    uint8_t *maj = maj_all + st.maj_off;
    const int thr = w < 5 ? 1 : (int)(0.2 * (double)w);
    if (threadIdx.x == 0) s_bad = 0;
    __syncthreads();
    for (int c = 0; c < K; ++c) {
        if (threadIdx.x == 0) s_bad_c = 0;
        // the member rows of cluster c in member order (sequence j = threadIdx.x copies its rows behind
        // those of the earlier sequences of the cluster); clusters of more than REF_ROWS rows take the
        // plain loops below
        int my_start = 0, total_rows = 0;
        for (int j = 0; j < n; ++j) {
            if (assign[j] != c) continue;
            const int cnt = mem_off[j + 1] - mem_off[j];
            if (j < (int)threadIdx.x) my_start += cnt;
            total_rows += cnt;
        }
        const bool listed = total_rows <= REF_ROWS;
        if (listed) {
            for (int j = threadIdx.x; j < n; j += blockDim.x) {
                if (assign[j] != c) continue;
                int start = 0;
                if (j == (int)threadIdx.x) {
                    start = my_start;
                } else {
                    for (int q = 0; q < j; ++q)
                        if (assign[q] == c) start += mem_off[q + 1] - mem_off[q];
                }
                for (int m = mem_off[j]; m < mem_off[j + 1]; ++m) s_rows[start + (m - mem_off[j])] = mem_rows[m];
            }
        }
        __syncthreads();
        // majority symbol per column over the members of cluster c in member order
        if (listed && w <= (int)blockDim.x / 2) {
            // narrow windows: the threads are (row slice, column) pairs; slice s counts rows s, s + S, ...,
            // then the first slice of each column adds the slices up (first-seen = smallest row position)
            const int S = blockDim.x / w;
            const int col = threadIdx.x % w, slice = threadIdx.x / w;
            if (slice < S) {
#pragma unroll
                for (int sy = 0; sy < 16; ++sy) cnt_s[sy][threadIdx.x] = 0;
                for (int k = slice; k < total_rows; k += S) {
                    const int sy = g[(long long)s_rows[k] * w + col];
                    if (cnt_s[sy][threadIdx.x]++ == 0) first_s[sy][threadIdx.x] = k;
                }
            }
            __syncthreads();
            if (slice == 0) {
                int best = -1, best_cnt = 0, best_first = 0;
                for (int sy = 0; sy < 16; ++sy) {
                    int cn = 0, fi = 0x7fffffff;
                    for (int q = 0; q < S; ++q) {
                        const int cq = cnt_s[sy][q * w + col];
                        if (cq) {
                            cn += cq;
                            fi = min(fi, first_s[sy][q * w + col]);
                        }
                    }
                    if (cn == 0) continue;
                    if (cn > best_cnt || (cn == best_cnt && fi < best_first)) {
                        best = sy;
                        best_cnt = cn;
                        best_first = fi;
                    }
                }
                maj[col] = (uint8_t)best;
            }
        } else {
            for (int col0 = 0; col0 < w; col0 += blockDim.x) {
                const int col = col0 + threadIdx.x;
                if (col < w) {
#pragma unroll
                    for (int sy = 0; sy < 16; ++sy) cnt_s[sy][threadIdx.x] = 0;
                    int order = 0;
                    for (int j = 0; j < n; ++j) {
                        if (assign[j] != c) continue;
                        for (int m = mem_off[j]; m < mem_off[j + 1]; ++m) {
                            const int sy = g[(long long)mem_rows[m] * w + col];
                            if (cnt_s[sy][threadIdx.x]++ == 0) first_s[sy][threadIdx.x] = order;
                            ++order;
                        }
                    }
                    int best = -1, best_cnt = 0, best_first = 0;
                    for (int sy = 0; sy < 16; ++sy) {
                        const int cn = cnt_s[sy][threadIdx.x];
                        if (cn == 0) continue;
                        const int fi = first_s[sy][threadIdx.x];
                        if (cn > best_cnt || (cn == best_cnt && fi < best_first)) {
                            best = sy;
                            best_cnt = cn;
                            best_first = fi;
                        }
                    }
                    maj[col] = (uint8_t)best;
                }
            }
        }
        __syncthreads();
        // Hamming distance of every member to the majority string
        if (listed) {
            for (int k = threadIdx.x; k < total_rows; k += blockDim.x) {
                const uint8_t *row = g + (long long)s_rows[k] * w;
                int d = 0;
                for (int i = 0; i < w; ++i) d += row[i] != maj[i];
                if (d > thr) {
                    s_bad = 1;
                    s_bad_c = 1;
                }
            }
        } else {
            for (int j = 0; j < n; ++j) {
                if (assign[j] != c) continue;
                const int m0 = mem_off[j], m1 = mem_off[j + 1];
                for (int m = m0 + threadIdx.x; m < m1; m += blockDim.x) {
                    const uint8_t *row = g + (long long)mem_rows[m] * w;
                    int d = 0;
                    for (int i = 0; i < w; ++i) d += row[i] != maj[i];
                    if (d > thr) {
                        s_bad = 1;
                        s_bad_c = 1;
                    }
                }
            }
        }
        __syncthreads();
        if (flags_out) {
            if (threadIdx.x == 0) flags_out[st.assign_off + c] = s_bad_c ? 0 : 1;
        } else if (s_bad) {
            break;
        }
        __syncthreads();
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        // kmeans_cluster_seqs loop control (cluster_sequences.py:256-261)
        if (!s_bad) {
            st.status = 1;
        } else {
            st.K = K + 1;
            if (st.K > max_clusters || st.K == n) st.status = 1;
            else st.run_kmeans = 1;
        }
    }
}

// ---- the same check for a deep locus, with the whole grid -----------------------------------------
// grid (column strips of 128, clusters): majority symbol per column of cluster blockIdx.y, members in
// member order, ties to the first-seen symbol -- the loop of refcheck_kernel, one strip per CTA
__global__ void __launch_bounds__(128)
refcheck_big_majority_kernel(const ClusterState *__restrict__ states, int q, const uint8_t *__restrict__ G,
                             const int *__restrict__ mem_off_all, const int *__restrict__ mem_rows_all,
                             const int *__restrict__ assign_all, uint8_t *__restrict__ maj_all) {
    __shared__ int cnt_s[16][128];
    __shared__ int first_s[16][128];
    const ClusterState &st = states[q];
    const int c = blockIdx.y;
    if (st.status != 0 || c >= st.K) return;
    const uint8_t *g = G + st.g_off;
    const int w = st.w, n = st.n;
    const int *mem_off = mem_off_all + st.mem_off;
    const int *mem_rows = mem_rows_all + st.mem_rows_off;
    const int *assign = assign_all + st.assign_off;
    uint8_t *maj = maj_all + st.maj_off + (long long)c * w;
    const int col = blockIdx.x * 128 + threadIdx.x;
    if (col >= w) return;
#pragma unroll
    for (int s = 0; s < 16; ++s) cnt_s[s][threadIdx.x] = 0;
    int order = 0;
    for (int j = 0; j < n; ++j) {
        if (assign[j] != c) continue;
        for (int m = mem_off[j]; m < mem_off[j + 1]; ++m) {
            const int s = g[(long long)mem_rows[m] * w + col];
            if (cnt_s[s][threadIdx.x]++ == 0) first_s[s][threadIdx.x] = order;
            ++order;
        }
    }
    int best = -1, best_cnt = 0, best_first = 0;
    for (int s = 0; s < 16; ++s) {
        const int cn = cnt_s[s][threadIdx.x];
        if (cn == 0) continue;
        const int fi = first_s[s][threadIdx.x];
        if (cn > best_cnt || (cn == best_cnt && fi < best_first)) {
            best = s;
            best_cnt = cn;
            best_first = fi;
        }
    }
    maj[col] = (uint8_t)best;
}

// one warp per distinct sequence j: Hamming distance of each of its member rows to the majority string
// of its cluster; a row further than the threshold marks the cluster (and the problem) as not
// one-reference-like
__global__ void __launch_bounds__(128)
refcheck_big_hamming_kernel(const ClusterState *__restrict__ states, int q, const uint8_t *__restrict__ G,
                            const int *__restrict__ mem_off_all, const int *__restrict__ mem_rows_all,
                            const int *__restrict__ assign_all, const uint8_t *__restrict__ maj_all,
                            int *__restrict__ flags) {
    const ClusterState &st = states[q];
    if (st.status != 0) return;
    const int lane = threadIdx.x & 31;
    const int j = blockIdx.x * 4 + (threadIdx.x >> 5);
    if (j >= st.n) return;
    const uint8_t *g = G + st.g_off;
    const int w = st.w;
    const int *mem_off = mem_off_all + st.mem_off;
    const int *mem_rows = mem_rows_all + st.mem_rows_off;
    const int c = assign_all[st.assign_off + j];
    const uint8_t *maj = maj_all + st.maj_off + (long long)c * w;
    const int thr = w < 5 ? 1 : (int)(0.2 * (double)w);
    for (int m = mem_off[j]; m < mem_off[j + 1]; ++m) {
        const uint8_t *row = g + (long long)mem_rows[m] * w;
        int d = 0;
        for (int i = lane; i < w; i += 32) d += row[i] != maj[i];
        d = __reduce_add_sync(0xffffffffu, d);
        if (d > thr) {
            if (lane == 0) {
                flags[0] = 1;
                flags[1 + c] = 1;
            }
            break;
        }
    }
}

// kmeans_cluster_seqs loop control (cluster_sequences.py:256-261), then the flags are reset
__global__ void refcheck_big_control_kernel(ClusterState *__restrict__ states, int q, int max_clusters,
                                            int *__restrict__ flags) {
    ClusterState &st = states[q];
    if (st.status != 0) return;
    if (!flags[0]) {
        st.status = 1;
    } else {
        st.K = st.K + 1;
        if (st.K > max_clusters || st.K == st.n) st.status = 1;
        else st.run_kmeans = 1;
    }
    for (int i = 0; i < REFCHECK_FLAG_INTS; ++i) flags[i] = 0;
}

// ---- small device-side bookkeeping so that only O(#distinct) data crosses PCIe ---------------------
// exclusive scan of counts[0..n) into offs[0..n] (single CTA)
__global__ void __launch_bounds__(1024) scan_counts_kernel(const int *__restrict__ counts, int n,
                                                           int *__restrict__ offs) {
    __shared__ int s_warp[33];
    __shared__ int carry;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < n; base += blockDim.x) {
        const int i = base + threadIdx.x;
        const int v = i < n ? counts[i] : 0;
        int x = v;
        for (int d = 1; d < 32; d <<= 1) {
            const int o = __shfl_up_sync(0xffffffffu, x, d);
            if (lane >= d) x += o;
        }
        if (lane == 31) s_warp[warp] = x;
        __syncthreads();
        if (warp == 0) {
            int y = lane < nw ? s_warp[lane] : 0;
            for (int d = 1; d < 32; d <<= 1) {
                const int o = __shfl_up_sync(0xffffffffu, y, d);
                if (lane >= d) y += o;
            }
            s_warp[lane] = y;
        }
        __syncthreads();
        if (i < n) offs[i] = carry + (warp ? s_warp[warp - 1] : 0) + x - v;
        __syncthreads();
        if (threadIdx.x == 0) carry += s_warp[nw - 1];
        __syncthreads();
    }
    if (threadIdx.x == 0) offs[n] = carry;
}

// dst[dst_off[t] + i] = src[src_off[t] + i] for i < counts[t], two int arrays at once; one warp per t
__global__ void __launch_bounds__(256)
gather2_kernel(const long long *__restrict__ src_off, const int *__restrict__ counts,
               const int *__restrict__ dst_off, int n, const int *__restrict__ src_a,
               const int *__restrict__ src_b, int *__restrict__ dst_a, int *__restrict__ dst_b) {
    const int lane = threadIdx.x & 31;
    const int t = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (t >= n) return;
    const long long so = src_off[t];
    const int d = dst_off[t], c = counts[t];
    for (int i = lane; i < c; i += 32) {
        dst_a[d + i] = src_a[so + i];
        if (src_b) dst_b[d + i] = src_b[so + i];
    }
}

// member lists of a clustering problem: rows of every distinct long sequence, in row order.  One warp per
// problem: the long-sequence index of every group by a ballot prefix, the member counts by atomics on the
// problem's own offset array, the stable placement chunk by chunk (rows of one chunk that share a sequence
// are ranked with __match_any_sync, so every list keeps the row order).
__global__ void __launch_bounds__(32)
members_kernel(const MemberProb *__restrict__ probs, const int *__restrict__ group,
               const int *__restrict__ leader_len, int *__restrict__ long_of_group,
               int *__restrict__ mem_off_all, int *__restrict__ mem_rows_all) {
    const int lane = threadIdx.x;
    const MemberProb p = probs[blockIdx.x];
    const int *grp = group + p.row_off;
    const int *ll = leader_len + p.row_off;
    int *lg = long_of_group + p.row_off;
    int *mem_off = mem_off_all + p.mem_off;
    int *mem_rows = mem_rows_all + p.mem_rows_off;
    int n = 0;
    for (int g0 = 0; g0 < p.n_groups; g0 += 32) {
        const int g = g0 + lane;
        const bool is_long = g < p.n_groups && ll[g] >= p.k;
        const unsigned m = __ballot_sync(0xffffffffu, is_long);
        if (g < p.n_groups) lg[g] = is_long ? n + __popc(m & ((1u << lane) - 1u)) : -1;
        n += __popc(m);
    }
    for (int j = lane; j <= n; j += 32) mem_off[j] = 0;
    __syncwarp();
    for (int r = lane; r < p.R; r += 32) {
        const int j = lg[grp[r]];
        if (j >= 0) atomicAdd(&mem_off[j + 1], 1);
    }
    __syncwarp();
    // inclusive prefix over mem_off[1..n] (n <= R; chunks of 32 with a running carry)
    int carry = 0;
    for (int j0 = 1; j0 <= n; j0 += 32) {
        const int j = j0 + lane;
        int v = j <= n ? mem_off[j] : 0;
        for (int d = 1; d < 32; d <<= 1) {
            const int o = __shfl_up_sync(0xffffffffu, v, d);
            if (lane >= d) v += o;
        }
        v += carry;
        if (j <= n) mem_off[j] = v;
        carry = __shfl_sync(0xffffffffu, v, 31);
    }
    __syncwarp();
    // stable placement in row order: mem_off[j] is advanced as the cursor of list j (it starts at the
    // exclusive offset because mem_off[j] holds the inclusive sum of the lists before j), then shifted back
    for (int r0 = 0; r0 < p.R; r0 += 32) {
        const int r = r0 + lane;
        const int j = r < p.R ? lg[grp[r]] : -1;
        const unsigned peers = __match_any_sync(0xffffffffu, j);
        if (j >= 0) mem_rows[mem_off[j] + __popc(peers & ((1u << lane) - 1u))] = r;
        __syncwarp();
        if (j >= 0 && (peers & ((1u << lane) - 1u)) == 0) mem_off[j] += __popc(peers);
        __syncwarp();
    }
    // the cursors now stand at the END of their lists: mem_off[j] = offset of list j + 1; shift back
    for (int j0 = ((n + 32) / 32) * 32 - 32; j0 >= 0; j0 -= 32) {  // highest chunk first
        const int j = j0 + lane;
        const int v = (j >= 1 && j <= n) ? mem_off[j - 1] : 0;
        __syncwarp();
        if (j >= 1 && j <= n) mem_off[j] = v;
        __syncwarp();
    }
    if (lane == 0) mem_off[0] = 0;
}

cudaError_t launch_scan_counts(cudaStream_t s, const int *counts, int n, int *offs) {
    scan_counts_kernel<<<1, 1024, 0, s>>>(counts, n, offs);
    return cudaGetLastError();
}

// the number of distinct k-mers of every problem goes from the numbering kernel's output straight into the
// loop state (levels of small problems do not wait for it on the host: their matrices are laid out for
// the upper bound "every k-mer position is distinct")
__global__ void set_features_kernel(ClusterState *__restrict__ states, const int *__restrict__ F, int n) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q < n) states[q].F = F[q];
}

cudaError_t launch_set_features(cudaStream_t s, ClusterState *states, const int *F, int n) {
    if (n <= 0) return cudaSuccess;
    set_features_kernel<<<(n + 127) / 128, 128, 0, s>>>(states, F, n);
    return cudaGetLastError();
}

cudaError_t launch_gather2(cudaStream_t s, const long long *src_off, const int *counts, const int *dst_off,
                           int n, const int *src_a, const int *src_b, int *dst_a, int *dst_b) {
    if (n <= 0) return cudaSuccess;
    gather2_kernel<<<(n + 7) / 8, 256, 0, s>>>(src_off, counts, dst_off, n, src_a, src_b, dst_a, dst_b);
    return cudaGetLastError();
}

cudaError_t launch_members(cudaStream_t s, const MemberProb *probs, int n, const int *group,
                           const int *leader_len, int *long_of_group, int *mem_off, int *mem_rows) {
    if (n <= 0) return cudaSuccess;
    members_kernel<<<n, 32, 0, s>>>(probs, group, leader_len, long_of_group, mem_off, mem_rows);
    return cudaGetLastError();
}

cudaError_t launch_unpack(cudaStream_t s, const uint8_t *packed, const DTask *d_tasks, int n_tasks,
                          const int *d_rows, const long long *g_off, uint8_t *G, int row_split) {
    if (n_tasks <= 0) return cudaSuccess;
    // (a warp per task on the wide levels was measured: 122 us against 92 us per step for one CTA per task)
    unpack_kernel<<<dim3(n_tasks, std::max(1, std::min(row_split, 4096))), 256, 0, s>>>(packed, d_tasks, d_rows,
                                                                                     g_off, G);
    return cudaGetLastError();
}

cudaError_t launch_dedupe_big(cudaStream_t s, const DTask *d_tasks, int n_tasks, int max_rows, const long long *g_off,
                              const uint8_t *G, uint8_t *U, const long long *row_off, void *sig, int *leader_u,
                              int *leader_g, int *group, int *ulen, int *leaders, int *leader_len, int *n_ungapped,
                              int *n_gapped, int *err) {
    if (n_tasks <= 0 || n_tasks > 65535) return n_tasks <= 0 ? cudaSuccess : cudaErrorInvalidValue;
    const int by_warp = std::max(1, std::min((max_rows + 7) / 8, 8192));
    const int by_thread = std::max(1, std::min((max_rows + 255) / 256, 8192));
    dedupe_big_sig_kernel<<<dim3(by_warp, n_tasks), 256, 0, s>>>(d_tasks, g_off, G, row_off, (RowSig *)sig, ulen, U);
    dedupe_big_match_kernel<<<dim3(by_thread, n_tasks), 256, 0, s>>>(d_tasks, row_off, (const RowSig *)sig, leader_u,
                                                                     leader_g);
    dedupe_big_verify_kernel<<<dim3(by_warp, n_tasks), 256, 0, s>>>(d_tasks, g_off, G, U, row_off, leader_u, leader_g,
                                                                    ulen, err);
    dedupe_big_number_kernel<<<n_tasks, 256, 0, s>>>(d_tasks, row_off, leader_u, leader_g, ulen, group, leaders,
                                                     leader_len, n_ungapped, n_gapped);
    return cudaGetLastError();
}

cudaError_t launch_dedupe(cudaStream_t s, const DTask *d_tasks, int n_tasks, int max_rows, const long long *g_off,
                          const uint8_t *G, const long long *row_off, void *sig, int *leader_u,
                          int *leader_g, int *group, int *ulen, int *leaders, int *leader_len,
                          int *n_ungapped, int *n_gapped, int *err) {
    if (n_tasks <= 0) return cudaSuccess;
    // a warp for every small task; a CTA for the others and for the small ones with more than DW_LIST distinct rows
    (void)max_rows;
    dedupe_warp_kernel<<<(n_tasks + DW_WARPS - 1) / DW_WARPS, DW_WARPS * 32, 0, s>>>(
        d_tasks, n_tasks, g_off, G, row_off, group, ulen, leaders, leader_len, n_ungapped, n_gapped, err);
    dedupe_kernel<<<std::min(n_tasks, 148 * 8), 256, 0, s>>>(d_tasks, n_tasks, g_off, G, row_off, (RowSig *)sig, leader_u,
                                                             leader_g, group, ulen, leaders, leader_len, n_ungapped,
                                                             n_gapped, err);
    return cudaGetLastError();
}

size_t rowsig_bytes() { return sizeof(RowSig); }

cudaError_t launch_kmer(cudaStream_t s, const void *d_probs, int n_probs, const int *seq_rows,
                        const uint8_t *G, int k, uint8_t *useq, int *ints, uint64_t *keys, int *ming,
                        int *out_F, int *err) {
    if (n_probs <= 0) return cudaSuccess;
    kmer_kernel<<<n_probs, 256, 0, s>>>((const KmerProb *)d_probs, seq_rows, G, k, useq, ints, keys, ming,
                                        out_F, err);
    return cudaGetLastError();
}

// h_prob: host copy of probs[q] (sizes for the grids)
cudaError_t launch_kmer_big(cudaStream_t s, const void *d_probs, int q, const void *h_prob, const int *seq_rows,
                            const uint8_t *G, int k, uint8_t *useq, int *ints, uint64_t *keys, int *ming,
                            int *block_counts, int *out_F, int *err) {
    const KmerProb &hp = *(const KmerProb *)h_prob;
    const KmerProb *probs = (const KmerProb *)d_probs;
    const int n_chunks = (hp.Pmax + KB_CHUNK - 1) / KB_CHUNK;
    cudaError_t e;
    if ((e = cudaMemsetAsync(keys + hp.tab_off, 0xff, sizeof(uint64_t) * (size_t)hp.T, s)) != cudaSuccess) return e;
    if ((e = cudaMemsetAsync(ming + hp.tab_off, 0x7f, sizeof(int) * (size_t)hp.T, s)) != cudaSuccess) return e;
    kmer_big_ungap_kernel<<<(hp.n + KB_THREADS / 32 - 1) / (KB_THREADS / 32), KB_THREADS, 0, s>>>(probs, q, seq_rows, G,
                                                                                                  useq, ints);
    kmer_big_prefix_kernel<<<1, KB_THREADS, 0, s>>>(probs, q, k, ints);
    kmer_big_insert_kernel<<<std::min(n_chunks, 148 * 16), KB_THREADS, 0, s>>>(probs, q, k, useq, ints, keys, ming);
    kmer_big_first_kernel<<<n_chunks, KB_THREADS, 0, s>>>(probs, q, k, useq, ints, ming, block_counts, err);
    kmer_big_scan_kernel<<<1, KB_THREADS, 0, s>>>(block_counts, n_chunks, out_F);
    kmer_big_number_kernel<<<n_chunks, KB_THREADS, 0, s>>>(probs, q, ints, block_counts);
    return cudaGetLastError();
}

cudaError_t launch_kmer_fill_big(cudaStream_t s, const void *d_probs, int q, const void *h_prob, int F,
                                 const int *ints, double *X) {
    const KmerProb &hp = *(const KmerProb *)h_prob;
    const size_t smem = sizeof(int) * (size_t)F;
    if (smem > 200 * 1024) return cudaErrorInvalidValue;  // caller falls back to kmer_fill_kernel
    static bool attr = false;
    if (!attr) {
        cudaFuncSetAttribute(kmer_big_fill_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        attr = true;
    }
    kmer_big_fill_kernel<<<std::min(hp.n, 148 * 8), KB_THREADS, smem, s>>>((const KmerProb *)d_probs, q, F, ints, X);
    return cudaGetLastError();
}

cudaError_t launch_kmer_fill(cudaStream_t s, const void *d_probs, int n_probs, long long max_positions,
                             const int *ints, const int *F, double *X) {
    if (n_probs <= 0) return cudaSuccess;
    const long long stripes = (max_positions + 256 * 16 - 1) / (256 * 16);
    const dim3 grid(n_probs, (unsigned)std::min<long long>(std::max<long long>(stripes, 1), 256));
    kmer_fill_kernel<<<grid, 256, 0, s>>>((const KmerProb *)d_probs, ints, F, X);
    return cudaGetLastError();
}

cudaError_t launch_refcheck(cudaStream_t s, ClusterState *states, int n_probs, const uint8_t *G,
                            const int *mem_off, const int *mem_rows, int *assign, uint8_t *maj,
                            int max_clusters, int *flags_out) {
    if (n_probs <= 0) return cudaSuccess;
    refcheck_kernel<<<n_probs, 128, 0, s>>>(states, G, mem_off, mem_rows, assign, maj, max_clusters,
                                            flags_out);
    return cudaGetLastError();
}

cudaError_t launch_refcheck_big(cudaStream_t s, ClusterState *states, int q, int w, int rows, const uint8_t *G,
                                const int *mem_off, const int *mem_rows, int *assign, uint8_t *maj, int max_clusters,
                                int *flag_blocks) {
    int *flags = flag_blocks;
    refcheck_big_majority_kernel<<<dim3((w + 127) / 128, max_clusters), 128, 0, s>>>(states, q, G, mem_off, mem_rows,
                                                                                     assign, maj);
    refcheck_big_hamming_kernel<<<(rows + 3) / 4, 128, 0, s>>>(states, q, G, mem_off, mem_rows, assign, maj, flags);
    refcheck_big_control_kernel<<<1, 1, 0, s>>>(states, q, max_clusters, flags);
    return cudaGetLastError();
}

cudaError_t launch_refcheck_big_control(cudaStream_t s, ClusterState *states, int q, int max_clusters, int *flags) {
    refcheck_big_control_kernel<<<1, 1, 0, s>>>(states, q, max_clusters, flags);
    return cudaGetLastError();
}

}  // namespace mprg
