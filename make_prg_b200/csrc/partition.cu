// Kernel (a'): per-task column classes, gap reach and the interval partition.
//
// Replaces IntervalPartitioner (make_prg/from_msa/interval_partition.py:81-252) as called from
// NodeFactory._get_vertical_partition (make_prg/recursion_tree.py:500-513):
//   classify_kernel   colOR/colNOR/B  ->  consensus byte per column ('*' = non-match), gap_reach
//                     (prefix maximum of B), and a bit per column (1 = non-match)
//   partition_kernel  the left-to-right run state machine of __init__/_add_interval (:99-110,
//                     :143-185) on the run bitmask, one thread per task, output = intervals sorted by
//                     start; followed by the bijection check (:219-252)
//   demote_kernel     enforce_multisequence_nonmatch_intervals (:187-217): a non-match interval whose
//                     rows all spell the same ungapped, unambiguous sequence becomes a match interval
#include "common.cuh"
#include "kernels.cuh"

namespace mprg {

__device__ __forceinline__ int sym_at(const uint8_t *row, int col) {
    return packed_sym(row, col);
}

// one warp per task
__global__ void __launch_bounds__(128)
classify_kernel(const DTask *__restrict__ tasks, int n_tasks, const uint32_t *__restrict__ colOR,
                const uint32_t *__restrict__ colNOR, const unsigned *__restrict__ colB,
                uint8_t *__restrict__ cls, int *__restrict__ reach, uint32_t *__restrict__ starbits) {
    const int lane = threadIdx.x & 31;
    const int ti = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (ti >= n_tasks) return;
    const DTask t = tasks[ti];
    const int n = t.c1 - t.c0;
    const int shift = t.c0 & 31;  // c0 - a0
    unsigned run_max = 0;         // prefix maximum of B (a0-relative, +1 encoded)
    for (int i0 = 0; i0 < n; i0 += 32) {
        const int i = i0 + lane;
        const bool in = i < n;
        const int wi = i + shift;  // a0-relative column
        uint32_t orw = 0, norw = 0;
        unsigned b = 0;
        if (in) {
            // column accumulators use the packed layout: chunk -> word (c & 3) -> nibble (c >> 2)
            const int cc = wi & 31;
            const long long widx = (((long long)t.col_off + wi) >> 5) * 4 + (cc & 3);
            orw = (colOR[widx] >> ((cc >> 2) * 4)) & 15u;
            norw = (colNOR[widx] >> ((cc >> 2) * 4)) & 15u;
            b = colB[(long long)t.col_off + wi];
        }
        const bool uniform = ((orw ^ norw) == 15u);
        const bool is_match = in && uniform && sym_is_base((int)orw);
        // inclusive prefix max over the warp, seeded with the running maximum
        unsigned m = b;
        for (int d = 1; d < 32; d <<= 1) {
            const unsigned o = __shfl_up_sync(0xffffffffu, m, d);
            if (lane >= d) m = max(m, o);
        }
        m = max(m, run_max);
        run_max = __shfl_sync(0xffffffffu, m, 31);
        const uint32_t star = __ballot_sync(0xffffffffu, in && !is_match);
        if (in) {
            cls[(long long)t.col_off + wi] = is_match ? (uint8_t)(MPRG_ALPHABET[orw]) : (uint8_t)'*';
            // end (c0-relative) of the furthest gap run covering column i, else i - 1
            reach[(long long)t.col_off + wi] = (m >= (unsigned)(wi + 1)) ? (int)m - 1 - shift : i - 1;
        }
        if (lane == 0) starbits[((long long)t.col_off >> 5) + (i0 >> 5)] = star;
    }
}

// ---- state machine -----------------------------------------------------------------------------
struct IvStack {
    DInterval *iv;
    int n;
};

__device__ __forceinline__ int find_last(const IvStack &s, int type) {
    for (int i = s.n - 1; i >= 0; --i)
        if (s.iv[i].type == type) return i;
    return -1;
}

// _add_interval (interval_partition.py:143-185).  Returns true when `cur` stays the current run.
__device__ bool add_interval(IvStack &s, DInterval &cur, int mml, const int *reach, bool end) {
    if (cur.type == MPRG_IV_MATCH) {
        const int len = cur.stop - cur.start + 1;
        if (len < mml) {
            const int k = find_last(s, MPRG_IV_NONMATCH);
            DInterval last;
            if (k >= 0) {
                // _pop(NonMatch): by construction the last non-match is the top of the stack
                last = s.iv[k];
                for (int i = k; i + 1 < s.n; ++i) s.iv[i] = s.iv[i + 1];
                s.n--;
                last.stop += len + 1;
            } else {
                last.type = MPRG_IV_NONMATCH;
                last.start = cur.start;
                last.stop = cur.stop + 1;
            }
            if (end) {
                last.stop -= 1;
                s.iv[s.n++] = last;
            }
            cur = last;
            return true;
        }
    } else {
        const int km = find_last(s, MPRG_IV_MATCH);
        if (km >= 0 && reach != nullptr && reach[cur.start] >= cur.stop) {
            const int len_match = s.iv[km].stop - s.iv[km].start + 1;
            if (len_match - 1 < mml) {
                for (int i = km; i + 1 < s.n; ++i) s.iv[i] = s.iv[i + 1];
                s.n--;
                cur.start -= len_match;
                const int kn = find_last(s, MPRG_IV_NONMATCH);
                if (kn >= 0) {
                    s.iv[kn].stop += cur.stop - cur.start + 1;
                    return false;
                }
            } else {
                s.iv[km].stop -= 1;
                cur.start -= 1;
            }
        }
    }
    s.iv[s.n++] = cur;
    return false;
}

__device__ void run_partition(const uint32_t *starbits, const int *reach, int n, int mml,
                              DInterval *out, int *out_count, int *err) {
    IvStack s{out, 0};
    if (n < mml) {
        if (n > 0) {
            bool any_star = false;
            for (int w = 0; w * 32 < n; ++w) any_star |= starbits[w] != 0;
            out[0] = DInterval{0, n - 1, any_star ? MPRG_IV_NONMATCH : MPRG_IV_MATCH};
            s.n = 1;
        }
    } else {
        // walk the runs of equal class using the bitmask
        DInterval cur{0, 0, (int)(starbits[0] & 1u)};
        int pos = 1;  // next column to look at
        while (true) {
            // find the next column >= pos whose class differs from cur.type (or n)
            int nxt = n;
            for (int w = pos >> 5; w * 32 < n; ++w) {
                uint32_t bits = starbits[w];
                if (cur.type == MPRG_IV_NONMATCH) bits = ~bits;  // look for class != cur.type
                if (w == (pos >> 5)) bits &= ~0u << (pos & 31);
                if (bits) {
                    const int c = w * 32 + __ffs(bits) - 1;
                    if (c < n) nxt = c;
                    break;
                }
            }
            cur.stop = nxt - 1;  // simple extension up to the class change
            if (nxt >= n) break;
            const int letter_type = cur.type ^ 1;
            if (!add_interval(s, cur, mml, reach, false)) cur = DInterval{nxt, nxt, letter_type};
            // when add_interval keeps `cur` (short match absorbed) it already covers column nxt
            pos = nxt + 1;
            if (cur.stop < nxt) cur.stop = nxt;
        }
        add_interval(s, cur, mml, reach, true);
    }
    // sort by start (stack order is already increasing; insertion sort keeps it exact and cheap)
    for (int i = 1; i < s.n; ++i) {
        DInterval x = out[i];
        int j = i - 1;
        while (j >= 0 && out[j].start > x.start) {
            out[j + 1] = out[j];
            --j;
        }
        out[j + 1] = x;
    }
    // enforce_alignment_interval_bijection (interval_partition.py:219-252)
    int expect = 0;
    bool ok = true;
    for (int i = 0; i < s.n; ++i) {
        ok &= out[i].start == expect && out[i].stop >= out[i].start;
        expect = out[i].stop + 1;
    }
    ok &= expect == n;
    if (!ok) *err = 1;
    *out_count = s.n;
}

// One warp per task.  The state machine is inherently sequential (lane 0 runs it), but every push / pop of its
// interval stack used to be a round trip to global memory (162 us at the root level of the bench batch): the
// stack now lives in shared memory whenever the task cannot produce more intervals than fit -- the number of
// class runs of its consensus bounds them and the warp counts it first -- and the warp copies the result out.
constexpr int PART_WARPS = 4;
constexpr int PART_SMEM_IV = 448;  // intervals per warp in shared memory (21 KB per CTA)

__global__ void __launch_bounds__(32 * PART_WARPS)
partition_kernel(const DTask *__restrict__ tasks, int n_tasks, const uint32_t *__restrict__ starbits,
                 const int *__restrict__ reach, int mml, DInterval *__restrict__ intervals,
                 int *__restrict__ iv_count, int *__restrict__ err) {
    __shared__ DInterval s_iv[PART_WARPS][PART_SMEM_IV];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int ti = blockIdx.x * PART_WARPS + warp;
    if (ti >= n_tasks) return;
    const DTask t = tasks[ti];
    const int n = t.c1 - t.c0;
    const int shift = t.c0 & 31;
    // starbits / reach of this task are stored at a0-relative positions; the state machine wants
    // c0-relative ones.  classify_kernel wrote starbits already c0-relative (bit i0+lane), and reach
    // at col_off + shift + i.
    const uint32_t *sb = starbits + ((long long)t.col_off >> 5);
    // class runs of the consensus = 1 + transitions between neighbouring columns
    int transitions = 0;
    const int n_words = (n + 31) >> 5;
    for (int w = lane; w < n_words; w += 32) {
        const uint32_t x = sb[w];
        const uint32_t nxt = (w + 1 < n_words) ? sb[w + 1] : 0u;
        uint32_t d = x ^ ((x >> 1) | (nxt << 31));  // bit i: column 32 w + i differs from column 32 w + i + 1
        const int valid = min(32, n - 1 - (w << 5));  // pairs (i, i + 1) with i + 1 < n
        if (valid < 32) d &= valid <= 0 ? 0u : ((1u << valid) - 1u);
        transitions += __popc(d);
    }
    transitions = __reduce_add_sync(0xffffffffu, transitions);
    const bool in_smem = transitions + 2 <= PART_SMEM_IV;
    DInterval *out = intervals + t.iv_off;
    if (lane == 0)
        run_partition(sb, reach + (long long)t.col_off + shift, n, mml, in_smem ? s_iv[warp] : out, iv_count + ti, err);
    __syncwarp();
    if (in_smem) {
        const int count = iv_count[ti];
        const int *src = reinterpret_cast<const int *>(s_iv[warp]);
        int *dst = reinterpret_cast<int *>(out);
        for (int i = lane; i < 3 * count; i += 32) dst[i] = src[i];
    }
}

// consensus supplied by the caller (mprg_partition_consensus): single task, c0 = 0
__global__ void partition_consensus_kernel(const uint8_t *__restrict__ cons, const int *__restrict__ reach,
                                           int n, int mml, uint32_t *__restrict__ starbits,
                                           DInterval *__restrict__ intervals, int *__restrict__ iv_count,
                                           int *__restrict__ err) {
    const int lane = threadIdx.x;
    for (int i0 = 0; i0 < n; i0 += 32) {
        const int i = i0 + lane;
        const uint32_t star = __ballot_sync(0xffffffffu, i < n && cons[i] == '*');
        if (lane == 0) starbits[i0 >> 5] = star;
    }
    __syncwarp();
    if (lane == 0) run_partition(starbits, reach, n, mml, intervals, iv_count, err);
}

// ---- demotion ------------------------------------------------------------------------------------
// One CTA per task; warp w takes non-match intervals w, w+4, ...; lanes take rows and compare them
// with the first row of the task in ungapped order.  All equal and no RYKMSW (N-free by contract)
// => fewer than two expanded sequences => the interval becomes a match interval.
// wpt warps per task (4: the CTA's warps share the intervals of one task -- a root level, few tasks with ~60
// intervals each; 1: a warp per task -- the wide levels, tens of thousands of tasks with one or two intervals)
__global__ void __launch_bounds__(128)
demote_kernel(const uint8_t *__restrict__ packed, const DTask *__restrict__ tasks, int n_tasks, int wpt,
              const int *__restrict__ rows_arena, DInterval *__restrict__ intervals,
              const int *__restrict__ iv_count) {
    const int lane = threadIdx.x & 31;
    const int cta_warp = threadIdx.x >> 5;
    const int ti = wpt == 1 ? blockIdx.x * (blockDim.x >> 5) + cta_warp : blockIdx.x;
    if (ti >= n_tasks) return;
    const DTask t = tasks[ti];
    if (t.n_rows <= 0) return;  // "for testing convenience" (interval_partition.py:200-201)
    const int warp = wpt == 1 ? 0 : cta_warp;
    const int nw = wpt == 1 ? 1 : (blockDim.x >> 5);
    const int n_iv = iv_count[ti];
    const int *rows = t.rows_off >= 0 ? rows_arena + t.rows_off : nullptr;
    const uint8_t *msa = packed + t.base;
    const uint8_t *row0 = msa + (long long)(rows ? rows[0] : 0) * t.stride;
    for (int k = warp; k < n_iv; k += nw) {
        DInterval iv = intervals[t.iv_off + k];
        if (iv.type != MPRG_IV_NONMATCH) continue;
        const int s = t.c0 + iv.start, e = t.c0 + iv.stop;
        // ambiguity / N in the first row's ungapped content => at least two expansions or the row
        // is dropped; with any other row equal to it the same holds, so only row 0 needs the test
        bool bad = false;
        for (int c = s + lane; c <= e; c += 32) {
            const int sym = sym_at(row0, c);
            bad |= sym != SYM_GAP && !sym_is_base(sym);  // RYKMSW / N
        }
        bad = __any_sync(0xffffffffu, bad);
        bool differ = false;
        if (!bad) {
            for (int r0 = 1; r0 < t.n_rows && !differ; r0 += 32) {
                const int r = r0 + lane;
                bool d = false;
                if (r < t.n_rows) {
                    const uint8_t *row = msa + (long long)(rows ? rows[r] : r) * t.stride;
                    int i = s, j = s;  // i walks row0, j walks row r
                    while (true) {
                        while (i <= e && sym_at(row0, i) == SYM_GAP) ++i;
                        while (j <= e && sym_at(row, j) == SYM_GAP) ++j;
                        if (i > e || j > e) {
                            d = (i > e) != (j > e);
                            break;
                        }
                        if (sym_at(row0, i) != sym_at(row, j)) {
                            d = true;
                            break;
                        }
                        ++i;
                        ++j;
                    }
                }
                differ = __any_sync(0xffffffffu, d);
            }
        }
        if (!bad && !differ && lane == 0) intervals[t.iv_off + k].type = MPRG_IV_MATCH;
    }
}

cudaError_t launch_classify(cudaStream_t stream, const DTask *d_tasks, int n_tasks,
                            const uint32_t *colOR, const uint32_t *colNOR, const unsigned *colB,
                            uint8_t *cls, int *reach, uint32_t *starbits) {
    if (n_tasks <= 0) return cudaSuccess;
    classify_kernel<<<(n_tasks + 3) / 4, 128, 0, stream>>>(d_tasks, n_tasks, colOR, colNOR, colB, cls,
                                                          reach, starbits);
    return cudaGetLastError();
}

cudaError_t launch_partition(cudaStream_t stream, const DTask *d_tasks, int n_tasks,
                             const uint32_t *starbits, const int *reach, int mml, DInterval *intervals,
                             int *iv_count, int *err) {
    if (n_tasks <= 0) return cudaSuccess;
    partition_kernel<<<(n_tasks + PART_WARPS - 1) / PART_WARPS, 32 * PART_WARPS, 0, stream>>>(
        d_tasks, n_tasks, starbits, reach, mml, intervals, iv_count, err);
    return cudaGetLastError();
}

cudaError_t launch_partition_consensus(cudaStream_t stream, const uint8_t *cons, const int *reach, int n,
                                       int mml, uint32_t *starbits, DInterval *intervals, int *iv_count,
                                       int *err) {
    partition_consensus_kernel<<<1, 32, 0, stream>>>(cons, reach, n, mml, starbits, intervals, iv_count,
                                                     err);
    return cudaGetLastError();
}

cudaError_t launch_demote(cudaStream_t stream, const uint8_t *packed, const DTask *d_tasks, int n_tasks,
                          const int *d_rows, DInterval *intervals, const int *iv_count) {
    if (n_tasks <= 0) return cudaSuccess;
    const int wpt = n_tasks >= 4096 ? 1 : 4;
    const int grid = wpt == 1 ? (n_tasks + 3) / 4 : n_tasks;
    demote_kernel<<<grid, 128, 0, stream>>>(packed, d_tasks, n_tasks, wpt, d_rows, intervals, iv_count);
    return cudaGetLastError();
}

}  // namespace mprg
