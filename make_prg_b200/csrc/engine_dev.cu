// The device-resident level loop: NodeFactory.build's recursion (make_prg/recursion_tree.py:401-471) with
// the tree, the row subsets, the task tables, the clustering problem tables and the allele list all living
// in HBM.  At every recursion depth the kernels below turn the partition / clustering results into the next
// level's tasks on the device; the host only reads a block of counters twice per level (to size buffers
// and grids) and launches.  Nothing else crosses PCIe until the node table and the allele strings come
// back at the end.
//
//   level_begin -> prepare_tasks -> count_units   [counters -> host: sync A]   -> fill_units
//   scan -> classify -> partition -> demote -> expand_partition            (children, leaves, cluster tasks)
//   unpack -> dedupe -> make_problems                                      [counters -> host: sync C]
//   members -> k-mer numbering -> counts -> [one-ref check, KMeans] x K -> expand_clusters
//
// Arenas are bump-allocated with warp-aggregated atomics: where an object lands is arbitrary, the ORDER
// inside an object (children of a node, alleles of a leaf, rows of a cluster) is the reference's.
// Clustering levels that hold a deep problem (whole-grid kernels, exact-F count matrices) fetch the problem
// table and run the host-driven sequence of engine.cu on it (run_problems_host).
#include <algorithm>
#include <cstring>
#include <thread>

#include <nvtx3/nvToolsExt.h>  // header-only: ranges show up in Nsight Systems / Compute, no cost without a tool

#include "engine.cuh"

namespace mprg {

// NVTX range for the phases of the level loop (tracing)
struct NvtxRange {
    explicit NvtxRange(const char *name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
};

struct DLocus {
    long long base;
    int stride, n_rows, n_cols, flags;
    int as_root;  // 0: built below an existing node (mprg_build_sub)
    int status;
};

struct DevCounters {
    // whole build
    int n_nodes, n_alleles;
    long long pool_size, allele_bytes;
    int err, n_next;
    // partition pass of the current level
    int n_tasks, n_units;
    long long total_cols, total_iv, level_iters, sum_rw, sum_rows, max_rw, algo_bytes;
    int max_rows, n_ctasks;
    long long g_total, row_total, scratch_total;
    // clustering problems of the current level
    int np, max_n, n_big, n_rounds_pad;
    long long useq_total, ints_total, tab_total, x_total, memoff_total, memrows_total, assign_total, maj_total,
        kmd_total, kmi_total, seqrows_total, max_elements, max_P;
};

constexpr int ERR_PARTITION = 1, ERR_HASH = 2, ERR_LOOP = 8, ERR_OVERFLOW = 16;

// ---- warp-aggregated bump allocation: every lane of the warp calls, inactive lanes ask for 0 ----------
__device__ __forceinline__ long long warp_alloc(long long *counter, long long amount) {
    const int lane = threadIdx.x & 31;
    long long incl = amount;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const long long o = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += o;
    }
    const long long total = __shfl_sync(0xffffffffu, incl, 31);
    long long base = 0;
    if (lane == 31 && total) base = (long long)atomicAdd((unsigned long long *)counter, (unsigned long long)total);
    base = __shfl_sync(0xffffffffu, base, 31);
    return base + incl - amount;
}
__device__ __forceinline__ int warp_alloc(int *counter, int amount) {
    const int lane = threadIdx.x & 31;
    int incl = amount;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int o = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += o;
    }
    const int total = __shfl_sync(0xffffffffu, incl, 31);
    int base = 0;
    if (lane == 31 && total) base = atomicAdd(counter, total);
    base = __shfl_sync(0xffffffffu, base, 31);
    return base + incl - amount;
}
__device__ __forceinline__ long long warp_sum(long long v) {
#pragma unroll
    for (int d = 16; d; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
    return v;
}
__device__ __forceinline__ long long warp_max(long long v) {
#pragma unroll
    for (int d = 16; d; d >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, d));
    return v;
}
// Block-wide bump allocation: EVERY thread of the block calls (inactive ones ask for 0); one atomic per block and
// counter instead of one per warp -- the arenas of a level are bumped by thousands of warps, and atomics on one
// address serialise (prepare_tasks: 57 us for 30,000 tasks with warp-level atomics).  s_buf: 34 long longs.
__device__ __forceinline__ long long block_alloc(long long *counter, long long amount, long long *s_buf) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    long long incl = amount;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const long long o = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += o;
    }
    if (lane == 31) s_buf[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        long long v = lane < nw ? s_buf[lane] : 0;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const long long o = __shfl_up_sync(0xffffffffu, v, d);
            if (lane >= d) v += o;
        }
        s_buf[lane] = v;  // inclusive over the warps
        const long long total = __shfl_sync(0xffffffffu, v, 31);
        if (lane == 0) s_buf[32] = total ? (long long)atomicAdd((unsigned long long *)counter, (unsigned long long)total) : 0;
    }
    __syncthreads();
    const long long base = s_buf[32] + (warp ? s_buf[warp - 1] : 0) + incl - amount;
    __syncthreads();
    return base;
}
__device__ __forceinline__ int block_alloc(int *counter, int amount, long long *s_buf) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    int incl = amount;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int o = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += o;
    }
    if (lane == 31) s_buf[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        long long v = lane < nw ? s_buf[lane] : 0;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const long long o = __shfl_up_sync(0xffffffffu, v, d);
            if (lane >= d) v += o;
        }
        s_buf[lane] = v;
        const long long total = __shfl_sync(0xffffffffu, v, 31);
        if (lane == 0) s_buf[32] = total ? (long long)atomicAdd(counter, (int)total) : 0;
    }
    __syncthreads();
    const int base = (int)(s_buf[32] + (warp ? s_buf[warp - 1] : 0)) + incl - amount;
    __syncthreads();
    return base;
}

// N allocations of a block at once: the scans as above, then ONE round of atomics -- lane k of warp 0 bumps
// counter k, all N are in flight together -- instead of N global round trips one after the other on every
// CTA's critical path.  amt[k]: this thread's request in, its offset out.  s_buf: N * 33 words.
struct AllocCounter {
    void *p;
    bool wide;  // long long counter (else int)
};
template <int N>
__device__ __forceinline__ void block_alloc_n(const AllocCounter (&ctr)[N], long long (&amt)[N], long long *s_buf) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    long long incl[N];
#pragma unroll
    for (int k = 0; k < N; ++k) {
        long long v = amt[k];
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const long long o = __shfl_up_sync(0xffffffffu, v, d);
            if (lane >= d) v += o;
        }
        incl[k] = v;
        if (lane == 31) s_buf[k * 33 + warp] = v;
    }
    __syncthreads();
    if (warp == 0) {
        long long total[N], base[N];
#pragma unroll
        for (int k = 0; k < N; ++k) {
            long long v = lane < nw ? s_buf[k * 33 + lane] : 0;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const long long o = __shfl_up_sync(0xffffffffu, v, d);
                if (lane >= d) v += o;
            }
            s_buf[k * 33 + lane] = v;  // inclusive over the warps
            total[k] = __shfl_sync(0xffffffffu, v, 31);
        }
#pragma unroll
        for (int k = 0; k < N; ++k) {  // issued back to back: nothing below waits for one of them before the next
            base[k] = 0;
            if (lane == k && total[k]) {
                if (ctr[k].wide)
                    base[k] = (long long)atomicAdd((unsigned long long *)ctr[k].p, (unsigned long long)total[k]);
                else
                    base[k] = (long long)atomicAdd((int *)ctr[k].p, (int)total[k]);
            }
        }
#pragma unroll
        for (int k = 0; k < N; ++k)
            if (lane == k) s_buf[k * 33 + 32] = base[k];
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < N; ++k) amt[k] = s_buf[k * 33 + 32] + (warp ? s_buf[k * 33 + warp - 1] : 0) + incl[k] - amt[k];
    __syncthreads();
}

__device__ __forceinline__ int pow2_ceil_dev(int x) {
    int p = 1;
    while (p < x) p <<= 1;
    return p;
}

// ---- zero-filling several buffers with one launch (a level needs five, a clustering pass three; as
// cudaMemsetAsync calls they were 40 of the ~190 stream operations of a build) ---------------------------
struct ZeroSegs {
    void *p[5];
    unsigned long long bytes[5];  // multiples of 4, pointers 4-byte aligned
    int n;
};
__global__ void __launch_bounds__(256) zero_segments_kernel(ZeroSegs z) {
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, nthr = (size_t)gridDim.x * blockDim.x;
    for (int k = 0; k < z.n; ++k) {
        if (((uintptr_t)z.p[k] & 15) == 0) {
            uint4 *q = reinterpret_cast<uint4 *>(z.p[k]);
            const size_t n16 = z.bytes[k] >> 4;
            for (size_t i = tid; i < n16; i += nthr) q[i] = make_uint4(0u, 0u, 0u, 0u);
            uint32_t *t = reinterpret_cast<uint32_t *>(q + n16);
            const size_t rest = (z.bytes[k] & 15) >> 2;
            if (tid < rest) t[tid] = 0u;
        } else {
            uint32_t *q = reinterpret_cast<uint32_t *>(z.p[k]);
            const size_t n4 = z.bytes[k] >> 2;
            for (size_t i = tid; i < n4; i += nthr) q[i] = 0u;
        }
    }
}
static cudaError_t launch_zero(cudaStream_t s, std::initializer_list<std::pair<void *, size_t>> segs) {
    ZeroSegs z;
    z.n = 0;
    size_t most = 0;
    for (const auto &sg : segs) {
        if (sg.second == 0) continue;
        z.p[z.n] = sg.first;
        z.bytes[z.n] = sg.second;
        most = std::max(most, sg.second);
        ++z.n;
    }
    if (z.n == 0) return cudaSuccess;
    const size_t blocks = std::min<size_t>(std::max<size_t>((most / 16 + 255) / 256, 1), 148 * 8);
    zero_segments_kernel<<<(unsigned)blocks, 256, 0, s>>>(z);
    return cudaGetLastError();
}

// ---- level set-up ---------------------------------------------------------------------------------------
__global__ void level_begin_kernel(DevCounters *C) {
    C->n_tasks = C->n_next;
    C->n_next = 0;
    C->n_units = 0;
    C->total_cols = C->total_iv = C->level_iters = C->sum_rw = C->sum_rows = C->max_rw = C->algo_bytes = 0;
    C->max_rows = C->n_ctasks = 0;
    C->g_total = C->row_total = C->scratch_total = 0;
    C->np = C->max_n = C->n_big = 0;
    C->useq_total = C->ints_total = C->tab_total = C->x_total = C->memoff_total = C->memrows_total = 0;
    C->assign_total = C->maj_total = C->kmd_total = C->kmi_total = C->seqrows_total = C->max_elements = C->max_P = 0;
}

// one thread per pending node: its task descriptor, the offsets of its per-column outputs and of its
// interval arena, and the sums the host sizes the level by (level.cu's host loop, on the device)
__global__ void __launch_bounds__(256)
prepare_tasks_kernel(DevCounters *C, const int *__restrict__ pending, const DNode *__restrict__ nodes,
                     const DLocus *__restrict__ loci, int l0, int any_n, DTask *__restrict__ tasks) {
    // the host knows only a bound of the task count: a machine-sized grid walks the tasks
    __shared__ long long s_buf[2 * 33];
    __shared__ unsigned long long s_sum[4];
    __shared__ long long s_max[2];
    const int n_tasks = C->n_tasks;
    for (int base = blockIdx.x * blockDim.x; base < n_tasks; base += gridDim.x * blockDim.x) {
    const int i = base + threadIdx.x;
    const bool active = i < n_tasks;
    long long aligned = 0, ivn = 0, iters = 0, rw = 0, rows = 0, algo = 0;
    DTask t;
    if (active) {
        const DNode nd = nodes[pending[i]];
        const DLocus lc = loci[nd.locus - l0];
        t.base = lc.base;
        t.stride = lc.stride;
        t.rows_off = nd.row_off < 0 ? -1 : (int)nd.row_off;
        t.n_rows = nd.n_rows;
        t.c0 = nd.c0;
        t.c1 = nd.c1;
        t.flags = any_n ? 1 : 0;
        const int a0 = nd.c0 & ~31, a1 = (nd.c1 + 31) & ~31;
        aligned = max(a1 - a0, 32);
        ivn = max(nd.c1 - nd.c0, 1);
        const int nch = max(((nd.c1 + 31) >> 5) - (nd.c0 >> 5), 1);
        const int rem = nch % 32;
        iters = (long long)(nch / 32) * nd.n_rows;
        if (rem) {
            const int per = 32 / pow2_ceil_dev(rem);
            iters += (nd.n_rows + per - 1) / per;
        }
        rows = nd.n_rows;
        rw = (long long)nd.n_rows * (nd.c1 - nd.c0);
        algo = rw / 2 + (nd.row_off >= 0 ? 4LL * nd.n_rows : 0) + 5LL * (nd.c1 - nd.c0);
    }
    if (threadIdx.x < 4) s_sum[threadIdx.x] = 0;
    if (threadIdx.x < 2) s_max[threadIdx.x] = 0;
    const AllocCounter ctr[2] = {{&C->total_cols, true}, {&C->total_iv, true}};
    long long off[2] = {aligned, ivn};
    block_alloc_n<2>(ctr, off, s_buf);  // (its barriers order the zeroing above)
    const long long col_off = off[0], iv_off = off[1];
    const long long s_it = warp_sum(iters), s_rw = warp_sum(rw), s_rows = warp_sum(rows), s_algo = warp_sum(algo);
    const long long m_rw = warp_max(rw), m_rows = warp_max(rows);
    if ((threadIdx.x & 31) == 0 && s_rows) {
        atomicAdd(&s_sum[0], (unsigned long long)s_it);
        atomicAdd(&s_sum[1], (unsigned long long)s_rw);
        atomicAdd(&s_sum[2], (unsigned long long)s_rows);
        atomicAdd(&s_sum[3], (unsigned long long)s_algo);
        atomicMax(&s_max[0], m_rw);
        atomicMax(&s_max[1], m_rows);
    }
    __syncthreads();
    if (threadIdx.x == 0 && s_sum[2]) {
        atomicAdd((unsigned long long *)&C->level_iters, s_sum[0]);
        atomicAdd((unsigned long long *)&C->sum_rw, s_sum[1]);
        atomicAdd((unsigned long long *)&C->sum_rows, s_sum[2]);
        atomicAdd((unsigned long long *)&C->algo_bytes, s_sum[3]);
        atomicMax((long long *)&C->max_rw, s_max[0]);
        atomicMax(&C->max_rows, (int)s_max[1]);
    }
    if (active) {
        t.col_off = (int)col_off;
        t.iv_off = (int)iv_off;
        if (col_off + aligned > 0x7fffffffLL || iv_off + ivn > 0x7fffffffLL) atomicOr(&C->err, ERR_OVERFLOW);
        tasks[i] = t;
    }
    __syncthreads();  // s_sum / s_max are reused by the next trip
    }
}

__device__ __forceinline__ int tile_iters_of(long long level_iters, int sm_count, int forced) {
    // level.cu: 3.5 waves of resident warps, in quanta of 4 warp iterations, 12..32
    if (forced > 0) return forced;
    const long long target = (long long)max(sm_count, 1) * 32 * 7 / 2;
    int ti = (int)((level_iters + target - 1) / target);
    ti = ((ti + 3) / 4) * 4;
    return min(max(ti, 12), 32);
}

// scan tiles of every task: full 32-chunk column strips, then the remainder strip; each strip cut into
// row ranges of tile_iters warp iterations (a short remainder joins the last tile) -- level.cu's loop
__device__ __forceinline__ int tiles_of_strip(int n_rows, int tile_rows) {
    if (n_rows <= 0) return 0;
    const int full = n_rows / tile_rows, rem = n_rows % tile_rows;
    if (full == 0) return 1;
    return full + ((rem > 0 && rem >= tile_rows / 2) ? 1 : 0);
}

// pass 1: how many tiles every task needs (the host sizes the tile table by the total)
__global__ void __launch_bounds__(256)
count_units_kernel(DevCounters *C, const DTask *__restrict__ tasks, int sm_count, int forced_iters,
                   int *__restrict__ unit_off) {
    __shared__ long long s_buf[34];
    const int n_tasks = C->n_tasks;
    const int tile_iters = tile_iters_of(C->level_iters, sm_count, forced_iters);
    for (int base = blockIdx.x * blockDim.x; base < n_tasks; base += gridDim.x * blockDim.x) {  // (as prepare_tasks)
        const int i = base + threadIdx.x;
        const bool active = i < n_tasks;
        int count = 0;
        if (active) {
            const DTask t = tasks[i];
            if (t.n_rows > 0) {
                const int ch0 = t.c0 >> 5, ch1 = max((t.c1 + 31) >> 5, ch0 + 1);
                for (int cb = ch0; cb < ch1; cb += 32) {
                    const int bn = min(32, ch1 - cb);
                    count += tiles_of_strip(t.n_rows, tile_iters * (32 / pow2_ceil_dev(bn)));
                }
            }
        }
        const int off = block_alloc(&C->n_units, count, s_buf);
        if (active) unit_off[i] = off;
    }
}

// pass 2: the tiles
__global__ void __launch_bounds__(256)
fill_units_kernel(const DevCounters *C, const int *__restrict__ pending, const DNode *__restrict__ nodes,
                  const DLocus *__restrict__ loci, int l0, const DTask *__restrict__ tasks, int sm_count,
                  int forced_iters, const int *__restrict__ unit_off, ScanUnit *__restrict__ units) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= C->n_tasks) return;
    const int tile_iters = tile_iters_of(C->level_iters, sm_count, forced_iters);
    const DTask t = tasks[i];
    if (t.n_rows <= 0) return;
    const int uflags = (loci[nodes[pending[i]].locus - l0].flags & 8) ? 1 : 0;
    const int ch0 = t.c0 >> 5, ch1 = max((t.c1 + 31) >> 5, ch0 + 1);
    int u = unit_off[i];
    for (int cb = ch0; cb < ch1; cb += 32) {
        const int bn = min(32, ch1 - cb);
        const int tile_rows = tile_iters * (32 / pow2_ceil_dev(bn));
        for (int rb = 0; rb < t.n_rows;) {
            int cnt = min(tile_rows, t.n_rows - rb);
            const int left = t.n_rows - rb - cnt;
            if (left > 0 && left < tile_rows / 2) cnt += left;
            units[u++] = ScanUnit{t.base, t.stride, t.rows_off >= 0 ? t.rows_off + rb : -1, rb, cnt, t.c0, t.c1, t.col_off,
                                  cb, bn, uflags};
            rb += cnt;
        }
    }
}

// ---- partition results -> tree -------------------------------------------------------------------------
struct ClusterTaskArrays {
    DTask *tasks;        // cluster tasks (zero-filled beyond n_ctasks: empty tasks for the kernels)
    long long *g_off;    // unpacked rows
    long long *row_off;  // per-row arrays
    int *R;
    int *node;           // node of the task
    int *want;           // nesting_level + 1 < max_nesting
    long long *sc_off;   // scratch ints of expand_clusters (2 R + 32 per task)
};

__device__ __forceinline__ int first_row_of(const DNode &nd, const int *pool) {
    return nd.row_off < 0 ? 0 : pool[nd.row_off];
}

// NodeFactory.build's decision per task (recursion_tree.py:436-471): single match interval -> leaf;
// several intervals (or a locus root) -> MultiIntervalNode with one child per interval, pure match
// children become leaves at once; a single non-match interval -> clustering task.
// LPT lanes per task: 1 for the wide levels (one thread per task, a few children each), 32 for the levels with
// few tasks and many children (a locus root has ~60 intervals: one thread per root was 57 us of serial stores on
// 8 SMs); the children of a task are then counted and written by a warp, 32 at a time.
template <int LPT>
__global__ void __launch_bounds__(128)
expand_partition_kernel(DevCounters *C, const int *__restrict__ pending, DNode *__restrict__ nodes,
                        DLocus *__restrict__ loci, int l0, const DTask *__restrict__ tasks,
                        const DInterval *__restrict__ iv, const int *__restrict__ iv_cnt, const int *__restrict__ pool,
                        int max_nesting, int *__restrict__ next, ExtractItem *__restrict__ items,
                        ClusterTaskArrays ct, int node_capacity, int item_capacity) {
    const int gt = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = gt / LPT, sub = gt % LPT;
    const unsigned lt = LPT == 32 ? ((1u << sub) - 1u) : 0u;
    const bool active = i < C->n_tasks;
    if (gt == 0 && iv_cnt[C->n_tasks]) atomicOr(&C->err, ERR_PARTITION);  // bijection check of partition_kernel
    int ni = -1, c = 0, n_child = 0, n_match = 0, n_non = 0, is_ct = 0;
    long long match_bytes = 0, rw = 0, R = 0;
    DNode nd;
    DTask t;
    const DInterval *ivs = nullptr;
    int mode = 0;  // 1 leaf, 2 interval node, 3 cluster task, 4 bad root
    if (active) {
        ni = pending[i];
        nd = nodes[ni];
        t = tasks[i];
        ivs = iv + t.iv_off;
        c = iv_cnt[i];
        const bool is_root = nd.parent < 0 && loci[nd.locus - l0].as_root;
        if (c == 1 && ivs[0].type != MPRG_IV_NONMATCH) {
            mode = 1;
            n_match = 1;
            match_bytes = nd.c1 - nd.c0;
        } else if (c > 1 || is_root) {
            if (c == 0) {
                mode = 4;
            } else {
                mode = 2;
                n_child = c;
                for (int k = sub; k < c; k += LPT) {
                    if (ivs[k].type == MPRG_IV_MATCH) {
                        ++n_match;
                        match_bytes += ivs[k].stop + 1 - ivs[k].start;
                    } else {
                        ++n_non;
                    }
                }
            }
        } else {
            mode = 3;
            is_ct = 1;
            R = nd.n_rows;
            rw = (long long)nd.n_rows * (nd.c1 - nd.c0);
        }
    }
    if (LPT == 32) {  // the task's totals on its first lane, which is the one that asks for the space
        n_match = (int)warp_sum(mode == 2 ? n_match : (sub == 0 ? n_match : 0));
        n_non = (int)warp_sum(n_non);
        match_bytes = warp_sum(mode == 2 ? match_bytes : (sub == 0 ? match_bytes : 0));
    }
    const bool asks = sub == 0;
    __shared__ long long s_buf[8 * 33];
    const AllocCounter ctr[8] = {{&C->n_nodes, false},     {&C->n_alleles, false}, {&C->allele_bytes, true},
                                 {&C->n_next, false},      {&C->n_ctasks, false},  {&C->g_total, true},
                                 {&C->row_total, true},    {&C->scratch_total, true}};
    long long off[8] = {n_child, n_match, match_bytes, n_non, is_ct, rw, R, is_ct ? 2LL * R + 32 : 0LL};
    if (!asks) {
#pragma unroll
        for (int k = 0; k < 8; ++k) off[k] = 0;
    }
    block_alloc_n<8>(ctr, off, s_buf);
    if (LPT == 32) {
#pragma unroll
        for (int k = 0; k < 8; ++k) off[k] = __shfl_sync(0xffffffffu, off[k], 0);
    }
    const int child0 = (int)off[0], item0 = (int)off[1];
    const long long byte0 = off[2];
    const int next0 = (int)off[3], q = (int)off[4];
    const long long g0 = off[5], r0 = off[6], s0 = off[7];
    if (!active) return;
    if (child0 + n_child > node_capacity || item0 + n_match > item_capacity) {
        if (asks) atomicOr(&C->err, ERR_OVERFLOW);
        return;
    }
    const DLocus lc = loci[nd.locus - l0];
    if (mode == 1) {
        if (!asks) return;
        nd.kind = MPRG_NODE_LEAF;
        nd.allele_first = item0;
        nd.allele_count = 1;
        items[item0] = ExtractItem{lc.base, lc.stride, first_row_of(nd, pool), nd.c0, nd.c1, byte0};
        nodes[ni] = nd;
    } else if (mode == 2) {
        if (asks) {
            nd.kind = MPRG_NODE_INTERVAL;
            nd.first_child = child0;
            nd.n_children = c;
            nodes[ni] = nd;
        }
        int a0 = item0, nx0 = next0;  // of the current group of LPT children
        long long by0 = byte0;
        const int frow = first_row_of(nd, pool);
        for (int k0 = 0; k0 < c; k0 += LPT) {
            const int k = k0 + sub;
            const bool valid = k < c;
            const DInterval me = valid ? ivs[k] : DInterval{0, 0, 0};
            const bool is_match = valid && me.type == MPRG_IV_MATCH;
            const long long width = is_match ? me.stop + 1 - me.start : 0;
            int a = a0, nx = nx0;
            long long by = by0;
            if (LPT == 32) {
                const unsigned m_mask = __ballot_sync(0xffffffffu, is_match);
                const unsigned n_mask = __ballot_sync(0xffffffffu, valid && !is_match);
                long long incl = width;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const long long o = __shfl_up_sync(0xffffffffu, incl, d);
                    if (sub >= d) incl += o;
                }
                a += __popc(m_mask & lt);
                nx += __popc(n_mask & lt);
                by += incl - width;
                a0 += __popc(m_mask);
                nx0 += __popc(n_mask);
                by0 += __shfl_sync(0xffffffffu, incl, 31);
            } else {
                a0 += is_match ? 1 : 0;
                nx0 += (valid && !is_match) ? 1 : 0;
                by0 += width;
            }
            if (!valid) continue;
            DNode ch;
            ch.locus = nd.locus;
            ch.parent = ni;
            ch.level = nd.level;
            ch.kind = -1;
            ch.c0 = nd.c0 + me.start;
            ch.c1 = nd.c0 + me.stop + 1;
            ch.row_off = nd.row_off;
            ch.n_rows = nd.n_rows;
            ch.first_child = -1;
            ch.n_children = 0;
            ch.allele_first = -1;
            ch.allele_count = 0;
            ch.pad = 0;
            if (is_match) {
                // a pure match interval re-partitions to itself: leaf without another scan
                ch.kind = MPRG_NODE_LEAF;
                ch.allele_first = a;
                ch.allele_count = 1;
                items[a] = ExtractItem{lc.base, lc.stride, frow, ch.c0, ch.c1, by};
            } else {
                next[nx] = child0 + k;
            }
            nodes[child0 + k] = ch;
        }
    } else if (mode == 3) {
        if (!asks) return;
        ct.tasks[q] = t;
        ct.g_off[q] = g0;
        ct.row_off[q] = r0;
        ct.R[q] = nd.n_rows;
        ct.node[q] = ni;
        ct.want[q] = (nd.level + 1 < max_nesting) ? 1 : 0;
        ct.sc_off[q] = s0;
    } else if (asks) {
        loci[nd.locus - l0].status = 2;  // zero-column root: the reference trips an assertion here
    }
}

// ---- clustering problems of a level (cluster_sequences.py:226-246, on the device) ------------------------
struct ProblemArrays {
    KmerProb *kp;
    MemberProb *mp;
    ClusterState *st;
    int *seq_rows;       // leader rows of the long sequences of every problem
    int *prob_of_ctask;  // per cluster task: its problem or -1
    int *clustered;      // per cluster task: kmeans_cluster_seqs is evaluated for it
    int *task_of_prob;
    long long *P_of_prob;
};

__global__ void __launch_bounds__(128)
make_problems_kernel(DevCounters *C, ClusterTaskArrays ct, const int *__restrict__ n_ungapped,
                     const int *__restrict__ n_gapped, const int *__restrict__ leaders,
                     const int *__restrict__ leader_len, int kmer_size, ProblemArrays pa, int prob_capacity,
                     const int *__restrict__ hash_err) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    const bool active = q < C->n_ctasks;
    if (q == 0 && *hash_err) atomicOr(&C->err, ERR_HASH);  // a hash match that failed its exact verification
    int n = 0, is_prob = 0, nu = 0, R = 0, w = 0;
    long long P = 0, ro = 0;
    if (active) {
        nu = n_ungapped[q];
        const int ng = n_gapped[q];
        R = ct.R[q];
        w = ct.tasks[q].c1 - ct.tasks[q].c0;
        ro = ct.row_off[q];
        // NodeFactory._alignment_has_issues (recursion_tree.py:475-494) discards the clustering anyway
        const bool clustered = ct.want[q] && R > 0 && !(nu <= 2 || nu < ng);
        pa.clustered[q] = clustered ? 1 : 0;
        if (clustered) {
            for (int g = 0; g < nu; ++g) {
                const int len = leader_len[ro + g];
                if (len >= kmer_size) {
                    ++n;
                    P += len - kmer_size + 1;
                }
            }
            is_prob = n > 2;  // too few sequences: single cluster (no_clustering stays true)
        }
        if (!is_prob) pa.prob_of_ctask[q] = -1;
    }
    const long long nn = is_prob ? n : 0, PP = is_prob ? P : 0;
    int T = 64;
    while (T < 2 * PP && T < (1 << 30)) T <<= 1;
    const bool big_ref = is_prob && (long long)R * w >= REFCHECK_BIG_SYMBOLS;
    __shared__ long long s_buf[12 * 33];
    const AllocCounter ctr[12] = {{&C->np, false},           {&C->seqrows_total, true}, {&C->useq_total, true},
                                  {&C->ints_total, true},    {&C->tab_total, true},     {&C->x_total, true},
                                  {&C->memoff_total, true},  {&C->memrows_total, true}, {&C->assign_total, true},
                                  {&C->maj_total, true},     {&C->kmd_total, true},     {&C->kmi_total, true}};
    long long off[12] = {is_prob,
                         nn,
                         nn * w,
                         is_prob ? 2 * nn + 1 + 2 * PP : 0LL,
                         is_prob ? (long long)T : 0LL,
                         nn * PP,
                         is_prob ? nn + 1 : 0LL,
                         is_prob ? (long long)R : 0LL,
                         nn,
                         is_prob ? (big_ref ? 10LL * w : (long long)w) : 0LL,
                         is_prob ? km_dscratch_doubles(nn, PP) : 0LL,
                         is_prob ? km_iscratch_ints(nn) : 0LL};
    block_alloc_n<12>(ctr, off, s_buf);
    const int pq = (int)off[0];
    const long long seq_off = off[1], useq_off = off[2], pos_off = off[3], tab_off = off[4], x_off = off[5];
    const long long mem_off = off[6], mem_rows_off = off[7], assign_off = off[8], maj_off = off[9];
    const long long kmd_off = off[10], kmi_off = off[11];
    const long long m_n = warp_max(nn), m_el = warp_max(nn * PP), m_P = warp_max(PP);
    const bool big = is_prob && (PP >= KMER_BIG_POSITIONS || nn * PP >= KMEANS_BIG_ELEMENTS || big_ref ||
                                 PP > 0x3fffffffLL);
    const unsigned any_big = __ballot_sync(0xffffffffu, big);
    if ((threadIdx.x & 31) == 0) {
        if (m_n) {
            atomicMax(&C->max_n, (int)m_n);
            atomicMax((long long *)&C->max_elements, m_el);
            atomicMax((long long *)&C->max_P, m_P);
        }
        if (any_big) atomicAdd(&C->n_big, __popc(any_big));
    }
    if (!active || !is_prob) return;
    if (pq >= prob_capacity) {
        atomicOr(&C->err, ERR_OVERFLOW);
        return;
    }
    pa.prob_of_ctask[q] = pq;
    pa.task_of_prob[pq] = q;
    pa.P_of_prob[pq] = P;
    KmerProb k;
    k.g_off = ct.g_off[q];
    k.w = w;
    k.n = n;
    k.seq_off = (int)seq_off;
    k.useq_off = useq_off;
    k.pos_off = pos_off;
    k.tab_off = tab_off;
    k.T = T;
    k.Pmax = (int)min(P, 0x7fffffffLL);
    k.x_off = x_off;
    k.big = 0;
    k.pad = 0;
    pa.kp[pq] = k;
    MemberProb m;
    m.row_off = ro;
    m.R = R;
    m.n_groups = nu;
    m.k = kmer_size;
    m.mem_off = (int)mem_off;
    m.mem_rows_off = (int)mem_rows_off;
    pa.mp[pq] = m;
    ClusterState c;
    c.status = 0;
    c.run_kmeans = 0;
    c.K = 1;
    c.n = n;
    c.F = (int)min(P, 0x7fffffffLL);  // the bound "every k-mer position is distinct"; set_features puts the real F
    c.w = w;
    c.g_off = k.g_off;
    c.mem_off = m.mem_off;
    c.mem_rows_off = m.mem_rows_off;
    c.assign_off = (int)assign_off;
    c.maj_off = maj_off;
    c.x_off = x_off;
    c.kmd_off = kmd_off;
    c.kmi_off = kmi_off;
    c.big = 0;
    c.big_ref = 0;
    pa.st[pq] = c;
    int j = 0;
    for (int g = 0; g < nu; ++g)
        if (leader_len[ro + g] >= kmer_size) pa.seq_rows[seq_off + j++] = leaders[ro + g];
}

// ---- clustering results -> tree (recursion_tree.py:448-471; cluster_sequences.py:276-296) -----------------
// One warp per cluster task.  Either a MultiClusterNode whose children keep the input row order
// (recursion_tree.py:558-572; cluster 0 holds the first row, merge_clusters), or a forced leaf with one
// allele per distinct ungapped row in first-seen order.
__global__ void __launch_bounds__(128)
expand_clusters_kernel(DevCounters *C, ClusterTaskArrays ct, ProblemArrays pa, const ClusterState *__restrict__ states,
                       const int *__restrict__ assign_all, const int *__restrict__ n_ungapped,
                       const int *__restrict__ n_gapped, const int *__restrict__ group,
                       const int *__restrict__ leaders, const int *__restrict__ long_of_group,
                       int *__restrict__ scratch, DNode *__restrict__ nodes, const DLocus *__restrict__ loci, int l0,
                       int *__restrict__ pool, int *__restrict__ next, ExtractItem *__restrict__ items,
                       int node_capacity, int item_capacity, long long pool_capacity, const int *__restrict__ hash_err) {
    const int lane = threadIdx.x & 31;
    const int q = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (q >= C->n_ctasks) return;  // warp-uniform
    if (q == 0 && lane == 0 && *hash_err) atomicOr(&C->err, ERR_HASH);
    const int ni = ct.node[q];
    DNode nd = nodes[ni];
    const DLocus lc = loci[nd.locus - l0];
    const long long ro = ct.row_off[q];
    const int R = ct.R[q], w = nd.c1 - nd.c0;
    const int nu = n_ungapped[q], ng = n_gapped[q];
    const int pq = pa.prob_of_ctask[q];
    const bool has_issues = nu <= 2 || nu < ng;
    bool no_clustering = true;
    int n_labels = 0;
    const ClusterState *st = nullptr;
    if (pq >= 0) {
        st = states + pq;
        if (st->status != 1 && lane == 0) atomicOr(&C->err, ERR_LOOP);
        const int K = st->K;
        if (!(K == 1 || K == st->n)) {  // ClusterResult.no_clustering (cluster_sequences.py:276)
            no_clustering = false;
            n_labels = min(K, 10);  // K == 11 keeps the 10-cluster assignment
        }
    }
    const bool further = ct.want[q] && !has_issues && pa.clustered[q] && !no_clustering;
    const int *rows = nd.row_off < 0 ? nullptr : pool + nd.row_off;
    if (!further) {
        // forced leaf: one allele per distinct ungapped row, first-seen order
        int a0 = 0;
        long long b0 = 0;
        if (lane == 0) {
            a0 = atomicAdd(&C->n_alleles, nu);
            b0 = (long long)atomicAdd((unsigned long long *)&C->allele_bytes, (unsigned long long)nu * (unsigned long long)w);
        }
        a0 = __shfl_sync(0xffffffffu, a0, 0);
        b0 = __shfl_sync(0xffffffffu, b0, 0);
        if (a0 + nu > item_capacity) {
            if (lane == 0) atomicOr(&C->err, ERR_OVERFLOW);
            return;
        }
        for (int g = lane; g < nu; g += 32) {
            const int pos = leaders[ro + g];
            items[a0 + g] = ExtractItem{lc.base, lc.stride, rows ? rows[pos] : pos, nd.c0, nd.c1, b0 + (long long)g * w};
        }
        if (lane == 0) {
            nd.kind = MPRG_NODE_LEAF;
            nd.allele_first = a0;
            nd.allele_count = nu;
            nodes[ni] = nd;
        }
        return;
    }
    // cluster index of every distinct ungapped sequence: KMeans label for the long ones, one cluster per
    // small one behind them; the cluster of row 0 moves to the front (merge_clusters), the others keep
    // their order
    // (the region is allocated with the task: offsets derived from row_off and q would not be disjoint, the two
    // are bump-allocated from different counters and need not come in the same order)
    int *idx_of_group = scratch + ct.sc_off[q];  // nu ints
    int *cl_end = idx_of_group + R + 16;              // up to 10 + nu ints
    const int *lg = long_of_group + ro;
    const int *assign = assign_all + st->assign_off;
    int n_small = 0;
    for (int g0 = 0; g0 < nu; g0 += 32) {
        const int g = g0 + lane;
        const bool in = g < nu;
        const int l = in ? lg[g] : 0;
        const bool small = in && l < 0;
        const unsigned m = __ballot_sync(0xffffffffu, small);
        if (in) idx_of_group[g] = small ? n_labels + n_small + __popc(m & ((1u << lane) - 1u)) : assign[l];
        n_small += __popc(m);
    }
    const int n_cl = n_labels + n_small;
    __syncwarp();
    const int first = idx_of_group[group[ro]];
    for (int c = lane; c < n_cl; c += 32) cl_end[c] = 0;
    __syncwarp();
    // pass 1: rows per cluster
    for (int r0 = 0; r0 < R; r0 += 32) {
        const int r = r0 + lane;
        int c = -1;
        if (r < R) {
            c = idx_of_group[group[ro + r]];
            c = c == first ? 0 : (c < first ? c + 1 : c);
        }
        const unsigned peers = __match_any_sync(0xffffffffu, c);
        if (c >= 0 && (peers & ((1u << lane) - 1u)) == 0) cl_end[c] += __popc(peers);
        __syncwarp();
    }
    // exclusive starts (kept in cl_end as cursors), number of non-empty clusters
    int n_ne = 0;
    if (lane == 0) {
        int acc = 0;
        for (int c = 0; c < n_cl; ++c) {
            const int cnt = cl_end[c];
            n_ne += cnt > 0;
            cl_end[c] = acc;
            acc += cnt;
        }
    }
    n_ne = __shfl_sync(0xffffffffu, n_ne, 0);
    long long pool0 = 0;
    int child0 = 0, next0 = 0;
    if (lane == 0) {
        pool0 = (long long)atomicAdd((unsigned long long *)&C->pool_size, (unsigned long long)R);
        child0 = atomicAdd(&C->n_nodes, n_ne);
        next0 = atomicAdd(&C->n_next, n_ne);
    }
    pool0 = __shfl_sync(0xffffffffu, pool0, 0);
    child0 = __shfl_sync(0xffffffffu, child0, 0);
    next0 = __shfl_sync(0xffffffffu, next0, 0);
    if (child0 + n_ne > node_capacity || pool0 + R > pool_capacity || pool0 + R > 0x7fffffffLL) {
        if (lane == 0) atomicOr(&C->err, ERR_OVERFLOW);
        return;
    }
    __syncwarp();
    // pass 2: stable placement (every child keeps the input row order)
    for (int r0 = 0; r0 < R; r0 += 32) {
        const int r = r0 + lane;
        int c = -1;
        if (r < R) {
            c = idx_of_group[group[ro + r]];
            c = c == first ? 0 : (c < first ? c + 1 : c);
        }
        const unsigned peers = __match_any_sync(0xffffffffu, c);
        if (c >= 0) {
            const int rank = __popc(peers & ((1u << lane) - 1u));
            pool[pool0 + cl_end[c] + rank] = rows ? rows[r] : r;
        }
        __syncwarp();
        if (c >= 0 && (peers & ((1u << lane) - 1u)) == 0) cl_end[c] += __popc(peers);
        __syncwarp();
    }
    if (lane == 0) {
        nd.kind = MPRG_NODE_CLUSTER;
        nd.level += 1;
        nd.first_child = child0;
        nd.n_children = n_ne;
        nodes[ni] = nd;
        int k = 0, begin = 0;
        for (int c = 0; c < n_cl; ++c) {
            const int end = cl_end[c];  // cursors now stand at the end of their cluster
            if (end == begin) continue;
            DNode ch;
            ch.locus = nd.locus;
            ch.parent = ni;
            ch.level = nd.level;
            ch.kind = -1;
            ch.c0 = nd.c0;
            ch.c1 = nd.c1;
            ch.row_off = pool0 + begin;
            ch.n_rows = end - begin;
            ch.first_child = -1;
            ch.n_children = 0;
            ch.allele_first = -1;
            ch.allele_count = 0;
            ch.pad = 0;
            nodes[child0 + k] = ch;
            next[next0 + k] = child0 + k;
            ++k;
            begin = end;
        }
    }
}

// ---- PRG strings on the device (recursion_tree.py:194-300, prg_builder.py:100-110) -------------------------
// Loci without RYKMSW: the distinct ungapped rows of a leaf are its alleles, so the PRG of a locus is the
// pre-order concatenation of allele strings and site markers.  allele_len: ungapped length of every allele;
// leaf_len: text length of every leaf; prg_walk: one thread per locus walks its tree once and records where
// every node's text starts and ends and its site number; a prefix sum gives the position of every locus in the
// blob; prg_emit (one thread per node) writes the markers and tells every allele where it goes; the extraction
// kernel then cuts the alleles straight into the blob.
struct PrgInfo {
    long long off, len;
    int n_sites, n_nodes, status, pad;
};

__global__ void __launch_bounds__(128)
allele_len_kernel(const uint8_t *__restrict__ packed, const ExtractItem *__restrict__ items, int n_items,
                  int *__restrict__ len_out) {
    const int lane = threadIdx.x & 31;
    const int it = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (it >= n_items) return;
    const ExtractItem e = items[it];
    const uint8_t *row = packed + e.base + (long long)e.row * e.stride;
    int len = 0;
    for (int c0 = e.c0; c0 < e.c1; c0 += 32) {
        const int c = c0 + lane;
        const bool solid = c < e.c1 && packed_sym(row, c) != SYM_GAP;
        len += __popc(__ballot_sync(0xffffffffu, solid));
    }
    if (lane == 0) len_out[it] = len;
}

__device__ __forceinline__ int marker_len(int m) {
    int d = 1;
    while (m >= 10) {
        m /= 10;
        ++d;
    }
    return d + 2;
}
__device__ __forceinline__ void marker_put(char *dst, int m) {
    const int n = marker_len(m);
    dst[0] = ' ';
    dst[n - 1] = ' ';
    for (int k = n - 2; k >= 1; --k) {
        dst[k] = (char)('0' + m % 10);
        m /= 10;
    }
}

constexpr int PRG_STACK = 64;  // nesting depth of the tree: at most 2 * max_nesting + 2 in practice

// what the per-locus walk needs of a node, in one 32-byte sector (a DNode is 56 bytes and the leaf length a
// second dependent load): the walk is a chain of dependent loads, one per visit
struct __align__(32) WalkNode {
    int kind, n_children, first_child, allele_count;
    long long leaf_len;  // text length of a leaf without its markers (sum of its alleles' ungapped lengths)
    long long pad;
};

// one thread per node, so that the per-locus walks below touch WalkNodes only, never alleles
__global__ void __launch_bounds__(256)
leaf_len_kernel(const DNode *__restrict__ nodes, int n_nodes, const int *__restrict__ allele_len,
                WalkNode *__restrict__ walk) {
    const int ni = blockIdx.x * blockDim.x + threadIdx.x;
    if (ni >= n_nodes) return;
    const DNode nd = nodes[ni];
    long long sum = 0;
    if (nd.kind == MPRG_NODE_LEAF)
        for (int a = 0; a < nd.allele_count; ++a) sum += allele_len[nd.allele_first + a];
    walk[ni] = WalkNode{nd.kind, nd.n_children, nd.first_child, nd.allele_count, sum, 0};
}

// one thread per locus walks its tree once (pre-order, the order site numbers are handed out in) and records
// for every node where its text starts and ends relative to the locus (node_at / node_end) and its site number
// (node_site; 0 = none); lengths, sites and node counts per locus go to info.  Everything that writes is then
// per node (prg_emit_kernel).
__global__ void __launch_bounds__(128)
prg_walk_kernel(const WalkNode *__restrict__ nodes, const DLocus *__restrict__ loci, int nl,
                PrgInfo *__restrict__ info, long long *__restrict__ node_at,
                long long *__restrict__ node_end, int *__restrict__ node_site, int *__restrict__ err) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nl) return;
    if (loci[i].status != MPRG_LOCUS_OK) {
        info[i] = PrgInfo{0, 0, 0, 0, loci[i].status, 0};
        return;
    }
    int st_node[PRG_STACK], st_next[PRG_STACK], st_site[PRG_STACK];
    int depth = 0;
    st_node[0] = i;  // the root of locus i is node i
    st_next[0] = 0;
    st_site[0] = 0;
    long long at = 0;
    int site = 5, n_nodes = 0;
    while (depth >= 0) {
        const int ni = st_node[depth];
        const WalkNode nd = nodes[ni];
        if (st_next[depth] == 0) {
            ++n_nodes;
            node_at[ni] = at;
            if (nd.kind == MPRG_NODE_LEAF) {
                if (nd.allele_count == 1) {
                    node_site[ni] = 0;
                    at += nd.leaf_len;
                } else {
                    const int sn = site;
                    site += 2;
                    node_site[ni] = sn;
                    // " sn " a1 " sn+1 " a2 ... " sn+1 " ak " sn "
                    at += 2LL * marker_len(sn) + (long long)(nd.allele_count - 1) * marker_len(sn + 1) + nd.leaf_len;
                }
                node_end[ni] = at;
                --depth;
                continue;
            }
            // the children are contiguous: ask for all of them now, so that only the first of them is a wait
            for (int k = 1; k < nd.n_children && k < 64; ++k)
                asm volatile("prefetch.global.L1 [%0];" ::"l"(nodes + nd.first_child + k));
            if (nd.kind == MPRG_NODE_CLUSTER) {
                st_site[depth] = site;
                site += 2;
                at += marker_len(st_site[depth]);
            }
            node_site[ni] = st_site[depth];
        } else if (nd.kind == MPRG_NODE_CLUSTER) {
            const int m = st_next[depth] < nd.n_children ? st_site[depth] + 1 : st_site[depth];  // after a child
            at += marker_len(m);
        }
        if (st_next[depth] < nd.n_children) {
            const int ch = nd.first_child + st_next[depth];
            st_next[depth]++;
            if (depth + 1 >= PRG_STACK) {
                atomicOr(err, ERR_OVERFLOW);
                return;
            }
            ++depth;
            st_node[depth] = ch;
            st_next[depth] = 0;
            st_site[depth] = 0;
        } else {
            node_end[ni] = at;
            --depth;
        }
    }
    info[i] = PrgInfo{0, at, (site - 5) / 2, n_nodes, MPRG_LOCUS_OK, 0};
}

// one thread per node: the opening marker of a cluster node, the marker that follows a child of a cluster node
// (" s+1 " between children, " s " after the last), and for a leaf where each of its alleles goes
// (items[a].out_off) with the markers between them.  node_at < 0: the walk never reached the node.
__global__ void __launch_bounds__(256)
prg_emit_kernel(const DNode *__restrict__ nodes, int n_nodes, const DLocus *__restrict__ loci, int l0,
                const PrgInfo *__restrict__ info, const int *__restrict__ allele_len,
                const long long *__restrict__ node_at, const long long *__restrict__ node_end,
                const int *__restrict__ node_site, ExtractItem *__restrict__ items, char *__restrict__ blob) {
    const int ni = blockIdx.x * blockDim.x + threadIdx.x;
    if (ni >= n_nodes) return;
    const DNode nd = nodes[ni];
    if (loci[nd.locus - l0].status != MPRG_LOCUS_OK || node_at[ni] < 0) return;
    char *base = blob + info[nd.locus - l0].off;
    if (nd.parent >= 0) {
        const DNode pa = nodes[nd.parent];
        if (pa.kind == MPRG_NODE_CLUSTER) {
            const int ps = node_site[nd.parent];
            marker_put(base + node_end[ni], ni - pa.first_child + 1 < pa.n_children ? ps + 1 : ps);
        }
    }
    if (nd.kind == MPRG_NODE_CLUSTER) marker_put(base + node_at[ni], node_site[ni]);
    if (nd.kind != MPRG_NODE_LEAF || nd.allele_count <= 0) return;
    long long at = (base - blob) + node_at[ni];
    if (nd.allele_count == 1) {
        items[nd.allele_first].out_off = at;
        return;
    }
    const int sn = node_site[ni];
    marker_put(blob + at, sn);
    at += marker_len(sn);
    for (int a = 0; a < nd.allele_count; ++a) {
        items[nd.allele_first + a].out_off = at;
        at += allele_len[nd.allele_first + a];
        const int m = a + 1 < nd.allele_count ? sn + 1 : sn;
        marker_put(blob + at, m);
        at += marker_len(m);
    }
}

// exclusive prefix sum of the PRG lengths (single CTA); total -> *total_out
__global__ void __launch_bounds__(1024) prg_offsets_kernel(PrgInfo *__restrict__ info, int nl, long long *total_out) {
    __shared__ long long s_warp[33];
    __shared__ long long carry;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < nl; base += blockDim.x) {
        const int i = base + threadIdx.x;
        const long long v = i < nl ? info[i].len : 0;
        long long x = v;
        for (int d = 1; d < 32; d <<= 1) {
            const long long o = __shfl_up_sync(0xffffffffu, x, d);
            if (lane >= d) x += o;
        }
        if (lane == 31) s_warp[warp] = x;
        __syncthreads();
        if (warp == 0) {
            long long y = lane < nw ? s_warp[lane] : 0;
            for (int d = 1; d < 32; d <<= 1) {
                const long long o = __shfl_up_sync(0xffffffffu, y, d);
                if (lane >= d) y += o;
            }
            s_warp[lane] = y;
        }
        __syncthreads();
        if (i < nl) info[i].off = carry + (warp ? s_warp[warp - 1] : 0) + x - v;
        __syncthreads();
        if (threadIdx.x == 0) carry += s_warp[nw - 1];
        __syncthreads();
    }
    if (threadIdx.x == 0) *total_out = carry;
}

// ---- growable device buffers that keep their contents -----------------------------------------------------
static cudaError_t reserve_keep(DevBuf &b, size_t bytes, size_t used, cudaStream_t s) {
    if (bytes <= b.cap) return cudaSuccess;
    void *p = nullptr;
    const size_t want = bytes + bytes / 2 + 4096;
    cudaError_t e = cudaMalloc(&p, want);
    if (e != cudaSuccess) return e;
    if (b.p && used) {
        e = cudaMemcpyAsync(p, b.p, std::min(used, b.cap), cudaMemcpyDeviceToDevice, s);
        if (e != cudaSuccess) return e;
        e = cudaStreamSynchronize(s);
        if (e != cudaSuccess) return e;
    }
    if (b.p) cudaFree(b.p);
    b.p = p;
    b.cap = want;
    return cudaSuccess;
}

// buffers of the device-resident loop (ctx->d_dev)
enum {
    V_COUNTERS = 0, V_LOCI, V_NODES, V_POOL, V_ITEMS, V_PEND_A, V_PEND_B, V_CT_TASKS, V_CT_MISC, V_G, V_SIG, V_ROWINTS,
    V_PROBS, V_PROB_MISC, V_SCRATCH, V_USEQ, V_INTS, V_KEYS, V_MING, V_X, V_STATE_MISC, V_B14, V_KM, V_OUT, V_OUTLEN, V_UNIT_OFF, V_PRGINFO, V_LEAF,
    V_COUNT
};

int build_range_dev(mprg_ctx *ctx, mprg_batch *batch, int l_begin, int l_end, int32_t max_nesting,
                    int32_t min_match_length, mprg_result *res, bool allow_trace, const int32_t *root_levels) {
    cudaSetDevice(ctx->device);
    cudaStream_t s = ctx->stream;
    const int nl = l_end - l_begin;
    if (nl <= 0) return MPRG_OK;
    PhaseTrace trace;
    const bool trace_all = getenv("MPRG_TRACE_ALL") != nullptr;
    if (allow_trace || trace_all) g_trace = trace.on ? &trace : nullptr;
    DevBuf *V = ctx->d_dev;
    static const int forced_iters = []() {
        const char *e = getenv("MPRG_TILE_ITERS");
        const int v = e ? atoi(e) : 0;
        return (v >= 4 && v <= 1024) ? v : 0;
    }();

    // ---- loci, root nodes, first pending list ----
    std::vector<DLocus> h_loci(nl);
    std::vector<DNode> h_roots(nl);
    std::vector<int> h_pending;
    h_pending.reserve(nl);
    for (int i = 0; i < nl; ++i) {
        const int l = l_begin + i;
        LocusResult &L = res->loci[l];
        DLocus &d = h_loci[i];
        d.base = batch->base[l];
        d.stride = batch->stride[l];
        d.n_rows = batch->n_rows[l];
        d.n_cols = batch->n_cols[l];
        d.flags = batch->flags[l];
        d.as_root = !(root_levels && root_levels[l] >= 0);
        d.status = MPRG_LOCUS_OK;
        if (batch->flags[l] & 1) d.status = MPRG_LOCUS_CURATION_ERROR;
        else if ((batch->flags[l] & 2) || batch->n_rows[l] <= 0) d.status = 2;  // N present or empty alignment
        L.as_root = d.as_root != 0;
        DNode &r = h_roots[i];
        memset(&r, 0, sizeof(r));
        r.locus = l;
        r.parent = -1;
        r.level = (root_levels && root_levels[l] >= 0) ? root_levels[l] : 0;
        r.kind = -1;
        r.c0 = 0;
        r.c1 = batch->n_cols[l];
        r.row_off = -1;
        r.n_rows = batch->n_rows[l];
        r.first_child = -1;
        r.allele_first = -1;
        if (d.status == MPRG_LOCUS_OK) {
            h_pending.push_back(i);
        }
    }
    MPRG_CUDA(ctx, V[V_COUNTERS].reserve(sizeof(DevCounters)));
    MPRG_CUDA(ctx, V[V_LOCI].reserve(sizeof(DLocus) * nl));
    long long node_cap = std::max<long long>(V[V_NODES].cap / sizeof(DNode), 0);
    long long item_cap = std::max<long long>(V[V_ITEMS].cap / sizeof(ExtractItem), 0);
    long long pend_cap = std::max<long long>(std::min(V[V_PEND_A].cap, V[V_PEND_B].cap) / sizeof(int), 0);
    long long pool_cap = std::max<long long>(V[V_POOL].cap / sizeof(int), 0);
    // grow (keeping the contents) so that the kernels of one pass can never run out: the bounds are what
    // the pass can create at most
    long long n_nodes_known = nl, n_items_known = 0, pool_known = 0;
    auto ensure = [&](long long nodes_more, long long items_more, long long pend_need, long long pool_more) -> int {
        if (n_nodes_known + nodes_more > node_cap) {
            MPRG_CUDA(ctx, reserve_keep(V[V_NODES], sizeof(DNode) * (size_t)(n_nodes_known + nodes_more),
                                        sizeof(DNode) * (size_t)n_nodes_known, s));
            node_cap = V[V_NODES].cap / sizeof(DNode);
        }
        if (n_items_known + items_more > item_cap) {
            MPRG_CUDA(ctx, reserve_keep(V[V_ITEMS], sizeof(ExtractItem) * (size_t)(n_items_known + items_more),
                                        sizeof(ExtractItem) * (size_t)n_items_known, s));
            item_cap = V[V_ITEMS].cap / sizeof(ExtractItem);
        }
        if (pend_need > pend_cap) {
            // the list being read is kept, the one being written only has to be large enough
            MPRG_CUDA(ctx, reserve_keep(V[V_PEND_A], sizeof(int) * (size_t)pend_need, V[V_PEND_A].cap, s));
            MPRG_CUDA(ctx, reserve_keep(V[V_PEND_B], sizeof(int) * (size_t)pend_need, V[V_PEND_B].cap, s));
            pend_cap = std::min(V[V_PEND_A].cap, V[V_PEND_B].cap) / sizeof(int);
        }
        if (pool_known + pool_more > pool_cap) {
            MPRG_CUDA(ctx, reserve_keep(V[V_POOL], sizeof(int) * (size_t)(pool_known + pool_more),
                                        sizeof(int) * (size_t)pool_known, s));
            pool_cap = V[V_POOL].cap / sizeof(int);
        }
        return MPRG_OK;
    };
    int rc = ensure(0, 0, std::max<long long>(nl, 1), 1);
    if (rc != MPRG_OK) return rc;
    DevCounters h_cnt;
    memset(&h_cnt, 0, sizeof(h_cnt));
    h_cnt.n_nodes = nl;
    h_cnt.n_next = (int)h_pending.size();
    DevCounters *d_cnt = V[V_COUNTERS].as<DevCounters>();
    DLocus *d_loci = V[V_LOCI].as<DLocus>();
    {
        // start state of the range through ONE pinned staging block: copies from pageable vectors are staged
        // synchronously by the driver, one at a time (4 of them cost ~0.1 ms of a 3 ms build)
        const size_t o_loci = (sizeof(DevCounters) + 63) & ~(size_t)63;
        const size_t o_roots = (o_loci + sizeof(DLocus) * (size_t)nl + 63) & ~(size_t)63;
        const size_t o_pend = (o_roots + sizeof(DNode) * (size_t)nl + 63) & ~(size_t)63;
        MPRG_CUDA(ctx, ctx->h_setup.reserve(o_pend + sizeof(int) * (size_t)nl + 64));
        unsigned char *st = ctx->h_setup.as<unsigned char>();
        memcpy(st, &h_cnt, sizeof(h_cnt));
        memcpy(st + o_loci, h_loci.data(), sizeof(DLocus) * (size_t)nl);
        memcpy(st + o_roots, h_roots.data(), sizeof(DNode) * (size_t)nl);
        if (!h_pending.empty()) memcpy(st + o_pend, h_pending.data(), sizeof(int) * h_pending.size());
        MPRG_CUDA(ctx, mprg::copy_h2d(ctx, d_cnt, st, sizeof(h_cnt), s));
        MPRG_CUDA(ctx, mprg::copy_h2d(ctx, d_loci, st + o_loci, sizeof(DLocus) * nl, s));
        MPRG_CUDA(ctx, mprg::copy_h2d(ctx, V[V_NODES].p, st + o_roots, sizeof(DNode) * nl, s));
        if (!h_pending.empty())
            MPRG_CUDA(ctx, mprg::copy_h2d(ctx, V[V_PEND_B].p, st + o_pend, sizeof(int) * h_pending.size(), s));
    }
    int cur = V_PEND_B, nxt = V_PEND_A;  // level_begin turns "next" into "pending": the lists swap first
    long long pending_bound = (long long)h_pending.size();
    PinnedBuf &hc = ctx->h_cnt;
    MPRG_CUDA(ctx, hc.reserve(sizeof(DevCounters)));
    DevCounters *cnt = hc.as<DevCounters>();
    *cnt = h_cnt;  // what the result phase reads when no level runs (no buildable locus)
    const int any_n = batch->any_n ? 1 : 0;
    // device time of the level's scan launch, read after the next synchronisation (roofline object of bench.py)
    // Both are called while the GPU is busy with freshly launched work, not in the idle gap after a
    // synchronisation (a dozen cudaEventElapsedTime calls are ~20 us of host time).
    auto account_km = [&]() {
        for (int r = 0; r < ctx->km_pending; ++r) {
            float ms = 0;
            if (cudaEventElapsedTime(&ms, ctx->ev_km[0][r], ctx->ev_km[1][r]) == cudaSuccess) ctx->km_ms += ms;
        }
        ctx->km_launches += ctx->km_pending;
        ctx->km_problems += ctx->km_pending_problems;
        ctx->km_pending = 0;
        ctx->km_pending_problems = 0;
    };
    auto account_scan = [&]() {  // before ev0 / ev1 are recorded again
        if (!ctx->pending_scan) return;
        ctx->pending_scan = false;
        float ms = 0;
        if (cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1) != cudaSuccess) return;
        ctx->scan_ms += ms;
        ctx->scan_bytes += ctx->pending_scan_bytes;
        ctx->scan_launches += 1;
        if (ctx->scan_log_bytes.size() < (1u << 20)) {
            ctx->scan_log_bytes.push_back(ctx->pending_scan_bytes);
            ctx->scan_log_ms.push_back(ms);
        }
    };
    TRACE("dev: setup");

    NvtxRange nvtx_build("mprg: device-resident level loop");
    for (int level = 0;; ++level) {
        if (pending_bound == 0) break;
        char nvtx_name[48];
        snprintf(nvtx_name, sizeof(nvtx_name), "mprg: recursion level %d", level);
        NvtxRange nvtx_level(nvtx_name);
        // ---- tasks and scan tiles of the level ----
        const int pb = (int)std::min<long long>((pending_bound + 255) / 256, (long long)std::max(ctx->sm_count, 1) * 8);
        MPRG_CUDA(ctx, ctx->d_tasks.reserve(sizeof(DTask) * (size_t)pending_bound));
        level_begin_kernel<<<1, 1, 0, s>>>(d_cnt);
        prepare_tasks_kernel<<<pb, 256, 0, s>>>(d_cnt, V[cur].as<int>(), V[V_NODES].as<DNode>(), d_loci, l_begin, any_n,
                                                ctx->d_tasks.as<DTask>());
        MPRG_CUDA(ctx, V[V_UNIT_OFF].reserve(sizeof(int) * (size_t)pending_bound));
        count_units_kernel<<<pb, 256, 0, s>>>(d_cnt, ctx->d_tasks.as<DTask>(), ctx->sm_count, forced_iters,
                                              V[V_UNIT_OFF].as<int>());
        ctx->launches += 3;
        MPRG_CUDA(ctx, cudaGetLastError());
        MPRG_CUDA(ctx, mprg::copy_d2h(ctx, cnt, d_cnt, sizeof(DevCounters), s));
        MPRG_CUDA(ctx, mprg::wait_stream(ctx, s));  // ---- sync A ----
        TRACE("dev: prepare+sync A");
        if (cnt->err) break;
        const int nt = cnt->n_tasks;
        if (nt == 0) break;
        n_nodes_known = cnt->n_nodes;
        n_items_known = cnt->n_alleles;
        pool_known = cnt->pool_size;
        const long long total_cols = cnt->total_cols, total_iv = cnt->total_iv;
        const int n_units = cnt->n_units;
        // what the two passes of this level can create at most
        rc = ensure(total_iv + cnt->sum_rows, total_iv + cnt->sum_rows, total_iv + cnt->sum_rows + 16, cnt->sum_rows);
        if (rc != MPRG_OK) return rc;
        MPRG_CUDA(ctx, ctx->d_units.reserve(sizeof(ScanUnit) * (size_t)std::max(n_units, 1)));
        fill_units_kernel<<<(nt + 255) / 256, 256, 0, s>>>(d_cnt, V[cur].as<int>(), V[V_NODES].as<DNode>(), d_loci, l_begin,
                                                          ctx->d_tasks.as<DTask>(), ctx->sm_count, forced_iters,
                                                          V[V_UNIT_OFF].as<int>(), ctx->d_units.as<ScanUnit>());
        ctx->launches++;
        account_scan();  // (the previous level's launch, when that level had no clustering pass)
        const size_t words = (size_t)total_cols / 8;
        MPRG_CUDA(ctx, ctx->d_colwords.reserve(sizeof(uint32_t) * 2 * words));
        MPRG_CUDA(ctx, ctx->d_colB.reserve(sizeof(unsigned) * total_cols));
        MPRG_CUDA(ctx, ctx->d_cls.reserve((size_t)total_cols));
        MPRG_CUDA(ctx, ctx->d_reach.reserve(sizeof(int) * total_cols));
        MPRG_CUDA(ctx, ctx->d_misc.reserve(sizeof(uint32_t) * (total_cols / 32) + 64));
        MPRG_CUDA(ctx, ctx->d_iv.reserve(sizeof(DInterval) * total_iv));
        MPRG_CUDA(ctx, ctx->d_ivcnt.reserve(sizeof(int) * (nt + 1)));
        // cluster-task arrays (bounded by the level's tasks), zero-filled: entries past n_ctasks are empty tasks
        const size_t ct_misc_bytes = (sizeof(long long) * 3 + sizeof(int) * 8) * (size_t)nt + 64;
        MPRG_CUDA(ctx, V[V_CT_TASKS].reserve(sizeof(DTask) * (size_t)nt));
        MPRG_CUDA(ctx, V[V_CT_MISC].reserve(ct_misc_bytes));
        ClusterTaskArrays ct;
        ct.tasks = V[V_CT_TASKS].as<DTask>();
        ct.g_off = V[V_CT_MISC].as<long long>();
        ct.row_off = ct.g_off + nt;
        ct.sc_off = ct.row_off + nt;
        ct.R = reinterpret_cast<int *>(ct.sc_off + nt);
        ct.node = ct.R + nt;
        ct.want = ct.node + nt;
        int *d_nu = ct.want + nt, *d_ng = d_nu + nt;
        int *d_prob_of_ctask = d_ng + nt, *d_clustered = d_prob_of_ctask + nt;
        int *d_err = d_clustered + nt;  // [0] hash collision codes of the clustering kernels

        MPRG_CUDA(ctx, launch_zero(s, {{V[V_CT_TASKS].p, sizeof(DTask) * (size_t)nt},
                                       {V[V_CT_MISC].p, ct_misc_bytes},
                                       {ctx->d_colwords.p, sizeof(uint32_t) * 2 * (size_t)words},
                                       {ctx->d_colB.p, sizeof(unsigned) * (size_t)total_cols},
                                       {ctx->d_ivcnt.p, sizeof(int) * (size_t)(nt + 1)}}));
        ctx->launches += 1;
        uint32_t *colOR = ctx->d_colwords.as<uint32_t>();
        uint32_t *colNOR = colOR + words;
        const int *d_pool = V[V_POOL].as<int>();
        MPRG_CUDA(ctx, cudaEventRecord(ctx->ev0, s));
        MPRG_CUDA(ctx, launch_scan(s, batch->any_n, batch->d_packed, ctx->d_units.as<ScanUnit>(), n_units, d_pool, colOR,
                                   colNOR, ctx->d_colB.as<unsigned>()));
        MPRG_CUDA(ctx, cudaEventRecord(ctx->ev1, s));
        MPRG_CUDA(ctx, launch_classify(s, ctx->d_tasks.as<DTask>(), nt, colOR, colNOR, ctx->d_colB.as<unsigned>(),
                                       ctx->d_cls.as<uint8_t>(), ctx->d_reach.as<int>(), ctx->d_misc.as<uint32_t>()));
        int *ivcnt = ctx->d_ivcnt.as<int>();
        MPRG_CUDA(ctx, launch_partition(s, ctx->d_tasks.as<DTask>(), nt, ctx->d_misc.as<uint32_t>(), ctx->d_reach.as<int>(),
                                        min_match_length, ctx->d_iv.as<DInterval>(), ivcnt, ivcnt + nt));
        MPRG_CUDA(ctx, launch_demote(s, batch->d_packed, ctx->d_tasks.as<DTask>(), nt, d_pool, ctx->d_iv.as<DInterval>(),
                                     ivcnt));
        if (nt <= 16384)  // few tasks (a root level: many children each): a warp per task
            expand_partition_kernel<32><<<(nt + 3) / 4, 128, 0, s>>>(
                d_cnt, V[cur].as<int>(), V[V_NODES].as<DNode>(), d_loci, l_begin, ctx->d_tasks.as<DTask>(),
                ctx->d_iv.as<DInterval>(), ivcnt, d_pool, max_nesting, V[nxt].as<int>(), V[V_ITEMS].as<ExtractItem>(), ct,
                (int)std::min<long long>(node_cap, 0x7fffffff), (int)std::min<long long>(item_cap, 0x7fffffff));
        else
            expand_partition_kernel<1><<<(nt + 127) / 128, 128, 0, s>>>(
                d_cnt, V[cur].as<int>(), V[V_NODES].as<DNode>(), d_loci, l_begin, ctx->d_tasks.as<DTask>(),
                ctx->d_iv.as<DInterval>(), ivcnt, d_pool, max_nesting, V[nxt].as<int>(), V[V_ITEMS].as<ExtractItem>(), ct,
                (int)std::min<long long>(node_cap, 0x7fffffff), (int)std::min<long long>(item_cap, 0x7fffffff));
        ctx->launches += (n_units ? 1 : 0) + 4;
        MPRG_CUDA(ctx, cudaGetLastError());
        {
            // scan accounting for the roofline object (bench.py): event times are read after the next sync
            ctx->pending_scan_bytes = (double)cnt->algo_bytes;
            ctx->pending_scan = n_units > 0;
        }
        account_km();  // the previous level's KMeans launches
        TRACE("dev: partition pass launch");

        // ---- clustering pass: a locus root never clusters, so level 0 of from_msa skips it ----
        const bool may_cluster = !(level == 0 && !root_levels);
        if (may_cluster) {
            NvtxRange nvtx_cluster("mprg: clustering pass");
            // sized by the bound "every task of the level clusters"
            const long long g_bound = std::max<long long>(cnt->sum_rw, 1), r_bound = std::max<long long>(cnt->sum_rows, 1);
            MPRG_CUDA(ctx, V[V_G].reserve((size_t)g_bound));
            MPRG_CUDA(ctx, V[V_SIG].reserve(rowsig_bytes() * (size_t)r_bound));
            MPRG_CUDA(ctx, V[V_ROWINTS].reserve(sizeof(int) * (size_t)(6 * r_bound + 16)));
            int *d_leader_u = V[V_ROWINTS].as<int>();
            int *d_leader_g = d_leader_u + r_bound, *d_group = d_leader_g + r_bound, *d_ulen = d_group + r_bound;
            int *d_leaders = d_ulen + r_bound, *d_leadlen = d_leaders + r_bound;
            const bool deep = (cnt->max_rw >= DEDUPE_BIG_SYMBOLS || getenv("MPRG_FORCE_BIG_DEDUPE")) && nt <= 65535;
            MPRG_CUDA(ctx, launch_unpack(s, batch->d_packed, ct.tasks, nt, d_pool, ct.g_off, V[V_G].as<uint8_t>(),
                                         deep ? (cnt->max_rows + 7) / 8 : 1));
            if (deep) {
                MPRG_CUDA(ctx, V[V_USEQ].reserve((size_t)g_bound));  // compacted rows (scratch of this pass)
                MPRG_CUDA(ctx, launch_dedupe_big(s, ct.tasks, nt, cnt->max_rows, ct.g_off, V[V_G].as<uint8_t>(),
                                                 V[V_USEQ].as<uint8_t>(), ct.row_off, V[V_SIG].p, d_leader_u, d_leader_g,
                                                 d_group, d_ulen, d_leaders, d_leadlen, d_nu, d_ng, d_err));
                ctx->launches += 3;
                ctx->path_counts[MPRG_PATH_DEDUPE_GRID]++;
            } else {
                MPRG_CUDA(ctx, launch_dedupe(s, ct.tasks, nt, cnt->max_rows, ct.g_off, V[V_G].as<uint8_t>(), ct.row_off, V[V_SIG].p,
                                             d_leader_u, d_leader_g, d_group, d_ulen, d_leaders, d_leadlen, d_nu, d_ng,
                                             d_err));
                ctx->launches += 1;  // two kernels: a warp per small task, a CTA per other task
            }
            // problem tables: at most one problem per cluster task
            const size_t prob_bytes = (sizeof(KmerProb) + sizeof(MemberProb) + sizeof(ClusterState)) * (size_t)nt + 64;
            MPRG_CUDA(ctx, V[V_PROBS].reserve(prob_bytes));
            MPRG_CUDA(ctx, V[V_PROB_MISC].reserve((sizeof(long long) + sizeof(int) * 2) * (size_t)nt + sizeof(int) * (size_t)r_bound + 64));
            ProblemArrays pa;
            pa.kp = V[V_PROBS].as<KmerProb>();
            pa.mp = reinterpret_cast<MemberProb *>(pa.kp + nt);
            pa.st = reinterpret_cast<ClusterState *>(pa.mp + nt);
            pa.P_of_prob = V[V_PROB_MISC].as<long long>();
            pa.task_of_prob = reinterpret_cast<int *>(pa.P_of_prob + nt);
            pa.seq_rows = pa.task_of_prob + nt;  // at most one entry per row of the level
            pa.prob_of_ctask = d_prob_of_ctask;
            pa.clustered = d_clustered;
            make_problems_kernel<<<(nt + 127) / 128, 128, 0, s>>>(d_cnt, ct, d_nu, d_ng, d_leaders, d_leadlen,
                                                                  min_match_length, pa, nt, d_err);
            ctx->launches += 3;
            MPRG_CUDA(ctx, cudaGetLastError());
            MPRG_CUDA(ctx, mprg::copy_d2h(ctx, cnt, d_cnt, sizeof(DevCounters), s));
            MPRG_CUDA(ctx, mprg::wait_stream(ctx, s));  // ---- sync C ----
            TRACE("dev: dedupe+problems+sync C");
            if (cnt->err) break;
            const int np = cnt->np, n_ct = cnt->n_ctasks;
            const ClusterState *d_states = pa.st;
            const int *d_assign = nullptr;
            if (np > 0) {
                const bool fallback = cnt->n_big > 0 || cnt->x_total > (16LL << 20) || getenv("MPRG_EXACT_F") != nullptr;
                if (!fallback) {
                    // ---- every problem of the level is small: the whole clustering loop from device tables ----
                    MPRG_CUDA(ctx, V[V_USEQ].reserve((size_t)std::max<long long>(cnt->useq_total, 1)));
                    MPRG_CUDA(ctx, V[V_INTS].reserve(sizeof(int) * (size_t)std::max<long long>(cnt->ints_total, 1)));
                    MPRG_CUDA(ctx, V[V_KEYS].reserve(sizeof(uint64_t) * (size_t)cnt->tab_total));
                    MPRG_CUDA(ctx, V[V_MING].reserve(sizeof(int) * (size_t)cnt->tab_total));
                    MPRG_CUDA(ctx, V[V_X].reserve(sizeof(double) * (size_t)std::max<long long>(cnt->x_total, 1)));
                    MPRG_CUDA(ctx, V[V_STATE_MISC].reserve(sizeof(int) * 2 * (size_t)np + 64));
                    const size_t o_memrows = sizeof(int) * (size_t)cnt->memoff_total;
                    const size_t o_assign = o_memrows + sizeof(int) * (size_t)cnt->memrows_total;
                    const size_t o_newlab = o_assign + sizeof(int) * (size_t)cnt->assign_total;
                    const size_t o_maj = o_newlab + sizeof(int) * (size_t)cnt->assign_total;
                    MPRG_CUDA(ctx, V[V_B14].reserve(o_maj + (size_t)cnt->maj_total + 16));
                    MPRG_CUDA(ctx, V[V_KM].reserve(sizeof(double) * (size_t)cnt->kmd_total + sizeof(int) * (size_t)cnt->kmi_total + 64));
                    uint8_t *b14 = V[V_B14].as<uint8_t>();
                    int *d_memoff = reinterpret_cast<int *>(b14);
                    int *d_memrows = reinterpret_cast<int *>(b14 + o_memrows);
                    int *d_asg = reinterpret_cast<int *>(b14 + o_assign);
                    int *d_newlab = reinterpret_cast<int *>(b14 + o_newlab);
                    uint8_t *d_maj = b14 + o_maj;
                    int *d_F = V[V_STATE_MISC].as<int>();
                    int *d_tickets = d_F + np;
                    double *d_kmd = V[V_KM].as<double>();
                    int *d_kmi = reinterpret_cast<int *>(d_kmd + cnt->kmd_total);
                    ClusterState *states = pa.st;
                    MPRG_CUDA(ctx, launch_zero(s, {{d_tickets, sizeof(int) * (size_t)np},
                                                   {V[V_X].p, sizeof(double) * (size_t)cnt->x_total},
                                                   {d_asg, sizeof(int) * (size_t)cnt->assign_total}}));
                    ctx->launches += 1;
                    MPRG_CUDA(ctx, launch_members(s, pa.mp, np, d_group, d_leadlen, d_leader_u, d_memoff, d_memrows));
                    MPRG_CUDA(ctx, launch_kmer(s, pa.kp, np, pa.seq_rows, V[V_G].as<uint8_t>(), min_match_length,
                                               V[V_USEQ].as<uint8_t>(), V[V_INTS].as<int>(), V[V_KEYS].as<uint64_t>(),
                                               V[V_MING].as<int>(), d_F, d_err));
                    MPRG_CUDA(ctx, launch_kmer_fill(s, pa.kp, np, cnt->max_P, V[V_INTS].as<int>(), d_F, V[V_X].as<double>()));
                    MPRG_CUDA(ctx, launch_set_features(s, states, d_F, np));
                    MPRG_CUDA(ctx, launch_refcheck(s, states, np, V[V_G].as<uint8_t>(), d_memoff, d_memrows, d_asg, d_maj, 10));
                    MPRG_CUDA(ctx, launch_kmeans_prepare(s, states, np, V[V_X].as<double>(), d_kmd, d_kmi));
                    ctx->launches += 6;
                    ctx->path_counts[MPRG_PATH_REFCHECK_CTA]++;
                    // a problem with n distinct sequences stops at K == n: rounds 2 .. max n - 1 at most
                    const int last_round = std::min(10, cnt->max_n - 1);
                    for (int round = 2; round <= last_round; ++round) {
                        const int slot = ctx->km_pending < 12 ? ctx->km_pending : -1;
                        if (slot >= 0) {
                            for (int a = 0; a < 2; ++a)
                                if (!ctx->ev_km[a][slot]) MPRG_CUDA(ctx, cudaEventCreate(&ctx->ev_km[a][slot]));
                            MPRG_CUDA(ctx, cudaEventRecord(ctx->ev_km[0][slot], s));
                        }
                        MPRG_CUDA(ctx, launch_kmeans(s, states, np, V[V_X].as<double>(), d_kmd, d_kmi, d_asg, d_newlab,
                                                     d_tickets, cnt->max_elements));
                        if (slot >= 0) {
                            MPRG_CUDA(ctx, cudaEventRecord(ctx->ev_km[1][slot], s));
                            ctx->km_pending++;
                            ctx->km_pending_problems += np;
                        }
                        MPRG_CUDA(ctx, launch_refcheck(s, states, np, V[V_G].as<uint8_t>(), d_memoff, d_memrows, d_asg,
                                                       d_maj, 10));
                        ctx->launches += 2;
                        ctx->path_counts[MPRG_PATH_KMEANS_CTA]++;
                        ctx->path_counts[MPRG_PATH_REFCHECK_CTA]++;
                    }
                    d_states = states;
                    d_assign = d_asg;
                    TRACE("dev: clustering loop launch");
                } else {
                    // ---- a deep problem (or exact-F matrices): fetch the problem table, host-driven sequence ----
                    std::vector<KmerProb> kp(np);
                    std::vector<MemberProb> mp(np);
                    std::vector<long long> Pq(np);
                    std::vector<int> task_of(np), seq_rows((size_t)cnt->seqrows_total);
                    std::vector<DTask> h_ct((size_t)n_ct);
                    MPRG_CUDA(ctx, mprg::copy_d2h(ctx, h_ct.data(), ct.tasks, sizeof(DTask) * (size_t)n_ct, s));
                    MPRG_CUDA(ctx, mprg::copy_d2h(ctx, kp.data(), pa.kp, sizeof(KmerProb) * np, s));
                    MPRG_CUDA(ctx, mprg::copy_d2h(ctx, mp.data(), pa.mp, sizeof(MemberProb) * np, s));
                    MPRG_CUDA(ctx, mprg::copy_d2h(ctx, Pq.data(), pa.P_of_prob, sizeof(long long) * np, s));
                    MPRG_CUDA(ctx, mprg::copy_d2h(ctx, task_of.data(), pa.task_of_prob, sizeof(int) * np, s));
                    if (!seq_rows.empty())
                        MPRG_CUDA(ctx, mprg::copy_d2h(ctx, seq_rows.data(), pa.seq_rows, sizeof(int) * seq_rows.size(), s));
                    MPRG_CUDA(ctx, mprg::wait_stream(ctx, s));
                    // problems in the order of the device table; each names its slice of seq_rows
                    std::vector<HostProblem> hp(np);
                    std::vector<int> seq_sorted;
                    seq_sorted.reserve(seq_rows.size());
                    for (int q = 0; q < np; ++q) {
                        HostProblem &h = hp[q];
                        h.task = task_of[q];
                        h.n = kp[q].n;
                        h.P = Pq[q];
                        h.w = kp[q].w;
                        h.R = mp[q].R;
                        h.n_groups = mp[q].n_groups;
                        h.g_off = kp[q].g_off;
                        h.row_off = mp[q].row_off;
                        {
                            const DTask &t = h_ct[(size_t)h.task];
                            h.base = t.base;
                            h.stride = t.stride;
                            h.rows_off = t.rows_off;
                            h.c0 = t.c0;
                            // the locus of the task: the last one whose arena offset is not above the task's
                            int lo = 0, hi = nl;
                            while (hi - lo > 1) {
                                const int mid = (lo + hi) >> 1;
                                if (h_loci[mid].base <= t.base) lo = mid; else hi = mid;
                            }
                            h.alpha_flags = h_loci[lo].flags;
                        }
                        seq_sorted.insert(seq_sorted.end(), seq_rows.begin() + kp[q].seq_off,
                                          seq_rows.begin() + kp[q].seq_off + kp[q].n);
                    }
                    ProblemRun run;
                    rc = run_problems_host(ctx, s, hp, seq_sorted, min_match_length, V[V_G].as<uint8_t>(), d_group,
                                           d_leadlen, d_leader_u, d_err, false, run, batch->d_packed, d_pool);
                    if (rc != MPRG_OK) return rc;
                    d_states = run.d_states;
                    d_assign = run.d_assign;
                    TRACE("dev: clustering loop (host-driven)");
                }
            }
            if (n_ct > 0) {
                MPRG_CUDA(ctx, V[V_SCRATCH].reserve(sizeof(int) * (size_t)(cnt->scratch_total + 64)));
                expand_clusters_kernel<<<(n_ct + 3) / 4, 128, 0, s>>>(
                    d_cnt, ct, pa, d_states, d_assign, d_nu, d_ng, d_group, d_leaders, d_leader_u, V[V_SCRATCH].as<int>(),
                    V[V_NODES].as<DNode>(), d_loci, l_begin, V[V_POOL].as<int>(), V[nxt].as<int>(),
                    V[V_ITEMS].as<ExtractItem>(), (int)std::min<long long>(node_cap, 0x7fffffff),
                    (int)std::min<long long>(item_cap, 0x7fffffff), pool_cap, d_err);
                ctx->launches++;
                MPRG_CUDA(ctx, cudaGetLastError());
            }
        }
        // bound of the next level's tasks: non-match children of this level's intervals + cluster children
        pending_bound = total_iv + (may_cluster ? cnt->sum_rows : 0);
        std::swap(cur, nxt);
    }

    account_scan();  // the last level's launches (the loop ends right after a synchronisation)
    account_km();
    // ---- results: node table, row pool, allele strings ----
    if (cnt->err) {
        if (allow_trace || trace_all) g_trace = nullptr;
        if (cnt->err & ERR_PARTITION) MPRG_FAIL(ctx, MPRG_E_PARTITION, "Failed interval partitioning");
        if (cnt->err & ERR_OVERFLOW) MPRG_FAIL(ctx, MPRG_E_INTERNAL, "device arena overflow in the level loop");
        if (cnt->err & ERR_HASH) MPRG_FAIL(ctx, MPRG_E_INTERNAL, "hash collision detected while de-duplicating rows or numbering k-mers");
        MPRG_FAIL(ctx, MPRG_E_INTERNAL, "clustering loop did not terminate");
    }
    const int n_nodes = cnt->n_nodes, na = cnt->n_alleles;
    const long long pool_size = cnt->pool_size, out_total = cnt->allele_bytes;
    bool expansion = getenv("MPRG_HOST_ASSEMBLY") != nullptr;  // RYKMSW: the cartesian product is made on the host
    for (int i = 0; i < nl && !expansion; ++i) expansion = h_loci[i].status == MPRG_LOCUS_OK && (h_loci[i].flags & 4);
    if (!expansion) {
        NvtxRange nvtx_prg("mprg: PRG strings on the device");
        // ---- PRG strings assembled on the device: only the strings and the raw tree cross PCIe ----
        MPRG_CUDA(ctx, V[V_OUTLEN].reserve(sizeof(int) * (size_t)std::max(na, 1)));
        MPRG_CUDA(ctx, V[V_PRGINFO].reserve(sizeof(PrgInfo) * (size_t)nl + 64));
        PrgInfo *d_info = V[V_PRGINFO].as<PrgInfo>();
        long long *d_total = reinterpret_cast<long long *>(d_info + nl);
        int *d_len = V[V_OUTLEN].as<int>();
        ExtractItem *d_items = V[V_ITEMS].as<ExtractItem>();
        int *d_errflag = &d_cnt->err;
        const size_t nn1 = (size_t)std::max(n_nodes, 1);
        MPRG_CUDA(ctx, V[V_LEAF].reserve((sizeof(WalkNode) + sizeof(long long) * 2 + sizeof(int)) * nn1 + 64));
        WalkNode *d_walk = V[V_LEAF].as<WalkNode>();
        long long *d_node_at = reinterpret_cast<long long *>(d_walk + nn1);
        long long *d_node_end = d_node_at + nn1;
        int *d_node_site = reinterpret_cast<int *>(d_node_end + nn1);
        MPRG_CUDA(ctx, cudaMemsetAsync(d_node_at, 0xFF, sizeof(long long) * nn1, s));  // < 0: not reached by the walk
        // the raw tree (node table + row pool, final once the level loop has ended) travels on a second stream
        // while the strings are laid out: its 4-5 MB would otherwise sit between the walk and the emission
        if (!ctx->stream_side) MPRG_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->stream_side, cudaStreamNonBlocking));
        if (!ctx->ev_side) MPRG_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_side, cudaEventDisableTiming));
        RawTree raw;
        raw.l_begin = l_begin;
        raw.l_end = l_end;
        raw.n_nodes = n_nodes;
        raw.pool_size = pool_size;
        raw.nodes = pinned_acquire(sizeof(DNode) * (size_t)std::max(n_nodes, 1));
        raw.pool = pinned_acquire(sizeof(int) * (size_t)std::max<long long>(pool_size, 1));
        if (!raw.nodes.p || !raw.pool.p) MPRG_FAIL(ctx, MPRG_E_CUDA, "pinned allocation for the result failed");
        auto drop_raw = [&]() {
            mprg::wait_stream(ctx, ctx->stream_side);
            pinned_release(raw.nodes);
            pinned_release(raw.pool);
        };
        {
            cudaError_t e = cudaEventRecord(ctx->ev_side, s);
            if (e == cudaSuccess) e = cudaStreamWaitEvent(ctx->stream_side, ctx->ev_side, 0);
            if (e == cudaSuccess)
                e = mprg::copy_d2h(ctx, raw.nodes.p, V[V_NODES].p, sizeof(DNode) * (size_t)n_nodes, ctx->stream_side);
            if (e == cudaSuccess && pool_size > 0)
                e = mprg::copy_d2h(ctx, raw.pool.p, V[V_POOL].p, sizeof(int) * (size_t)pool_size, ctx->stream_side);
            if (e != cudaSuccess) {
                drop_raw();
                MPRG_CUDA(ctx, e);
            }
        }
        if (na > 0) allele_len_kernel<<<(na + 3) / 4, 128, 0, s>>>(batch->d_packed, d_items, na, d_len);
        leaf_len_kernel<<<(n_nodes + 255) / 256, 256, 0, s>>>(V[V_NODES].as<DNode>(), n_nodes, d_len, d_walk);
        prg_walk_kernel<<<(nl + 127) / 128, 128, 0, s>>>(d_walk, d_loci, nl, d_info, d_node_at, d_node_end, d_node_site,
                                                        d_errflag);
        prg_offsets_kernel<<<1, 1024, 0, s>>>(d_info, nl, d_total);
        ctx->launches += 4;
        MPRG_CUDA(ctx, cudaGetLastError());
        MPRG_CUDA(ctx, ctx->h_d.reserve(sizeof(PrgInfo) * (size_t)nl + 64));
        PrgInfo *h_info = ctx->h_d.as<PrgInfo>();
        long long *h_total = reinterpret_cast<long long *>(h_info + nl);
        MPRG_CUDA(ctx, mprg::copy_d2h(ctx, h_total, d_total, sizeof(long long), s));
        MPRG_CUDA(ctx, mprg::wait_stream(ctx, s));
        const long long blob_bytes = *h_total;
        TRACE("dev: measure + raw tree D2H");
        PinnedBlock blob = pinned_acquire((size_t)std::max<long long>(blob_bytes, 1));
        if (!blob.p) {
            drop_raw();
            MPRG_FAIL(ctx, MPRG_E_CUDA, "pinned allocation for the result failed");
        }
        MPRG_CUDA(ctx, V[V_OUT].reserve((size_t)std::max<long long>(blob_bytes, 1)));
        prg_emit_kernel<<<(n_nodes + 255) / 256, 256, 0, s>>>(V[V_NODES].as<DNode>(), n_nodes, d_loci, l_begin, d_info,
                                                            d_len, d_node_at, d_node_end, d_node_site, d_items,
                                                            V[V_OUT].as<char>());
        if (na > 0)
            MPRG_CUDA(ctx, launch_extract(s, batch->d_packed, d_items, na, V[V_OUT].as<uint8_t>(), d_len));
        ctx->launches += 2;
        MPRG_CUDA(ctx, cudaGetLastError());
        if (blob_bytes > 0) MPRG_CUDA(ctx, mprg::copy_d2h(ctx, blob.p, V[V_OUT].p, (size_t)blob_bytes, s));
        MPRG_CUDA(ctx, mprg::copy_d2h(ctx, h_info, d_info, sizeof(PrgInfo) * (size_t)nl, s));
        MPRG_CUDA(ctx, mprg::copy_d2h(ctx, cnt, d_cnt, sizeof(DevCounters), s));
        MPRG_CUDA(ctx, mprg::wait_stream(ctx, s));
        MPRG_CUDA(ctx, mprg::wait_stream(ctx, ctx->stream_side));
        TRACE("dev: strings D2H");
        if (cnt->err & ERR_OVERFLOW) {
            pinned_release(blob);
            drop_raw();
            MPRG_FAIL(ctx, MPRG_E_INTERNAL, "tree deeper than the PRG walk supports");
        }
        int raw_index;
        {
            std::lock_guard<std::mutex> lock(res->raw_mutex);
            raw_index = (int)res->raw.size();
            res->raw.push_back(raw);
            res->blobs.push_back(blob);
        }
        const char *base = static_cast<const char *>(blob.p);
        for (int i = 0; i < nl; ++i) {
            LocusResult &L = res->loci[l_begin + i];
            L.status = h_info[i].status;
            if (L.status != MPRG_LOCUS_OK) continue;
            L.prg_data = base + h_info[i].off;
            L.prg_size = h_info[i].len;
            L.n_sites = h_info[i].n_sites;
            L.n_nodes = h_info[i].n_nodes;
            L.raw_index = raw_index;
            L.raw_root = i;
            L.tables_ready = false;
        }
        TRACE("dev: result records");
        if (allow_trace || trace_all) {
            trace.report("mprg_build (device-resident loop)");
            g_trace = nullptr;
        }
        return MPRG_OK;
    }
    MPRG_CUDA(ctx, ctx->h_a.reserve(sizeof(DNode) * (size_t)std::max(n_nodes, 1)));
    MPRG_CUDA(ctx, ctx->h_c.reserve((size_t)std::max<long long>(out_total, 1) + 16));
    MPRG_CUDA(ctx, ctx->h_d.reserve((sizeof(int) + sizeof(ExtractItem)) * (size_t)std::max(na, 1) + 64));
    DNode *h_nodes = ctx->h_a.as<DNode>();
    const size_t pool_bytes = (sizeof(int) * (size_t)std::max<long long>(pool_size, 1) + 15) & ~(size_t)15;
    MPRG_CUDA(ctx, ctx->h_b.reserve(pool_bytes + sizeof(DLocus) * nl));
    int *h_pool = ctx->h_b.as<int>();
    DLocus *h_loci_back = reinterpret_cast<DLocus *>(reinterpret_cast<uint8_t *>(h_pool) + ((sizeof(int) * (size_t)std::max<long long>(pool_size, 1) + 15) & ~(size_t)15));
    uint8_t *h_out = ctx->h_c.as<uint8_t>();
    ExtractItem *h_items = ctx->h_d.as<ExtractItem>();
    int *h_len = reinterpret_cast<int *>(h_items + std::max(na, 1));
    if (na > 0) {
        MPRG_CUDA(ctx, V[V_OUT].reserve((size_t)std::max<long long>(out_total, 1)));
        MPRG_CUDA(ctx, V[V_OUTLEN].reserve(sizeof(int) * (size_t)na));
        MPRG_CUDA(ctx, launch_extract(s, batch->d_packed, V[V_ITEMS].as<ExtractItem>(), na, V[V_OUT].as<uint8_t>(),
                                      V[V_OUTLEN].as<int>()));
        ctx->launches++;
        MPRG_CUDA(ctx, mprg::copy_d2h(ctx, h_out, V[V_OUT].p, (size_t)out_total, s));
        MPRG_CUDA(ctx, mprg::copy_d2h(ctx, h_len, V[V_OUTLEN].p, sizeof(int) * (size_t)na, s));
        MPRG_CUDA(ctx, mprg::copy_d2h(ctx, h_items, V[V_ITEMS].p, sizeof(ExtractItem) * (size_t)na, s));
    }
    MPRG_CUDA(ctx, mprg::copy_d2h(ctx, h_nodes, V[V_NODES].p, sizeof(DNode) * (size_t)n_nodes, s));
    if (pool_size > 0) MPRG_CUDA(ctx, mprg::copy_d2h(ctx, h_pool, V[V_POOL].p, sizeof(int) * (size_t)pool_size, s));
    MPRG_CUDA(ctx, mprg::copy_d2h(ctx, h_loci_back, d_loci, sizeof(DLocus) * nl, s));
    MPRG_CUDA(ctx, mprg::wait_stream(ctx, s));
    TRACE("dev: results D2H");

    // ---- per-locus node tables (children contiguous, local indices), then the PRG strings ----
    std::vector<long long> out_off((size_t)std::max(na, 1));
    for (int a = 0; a < na; ++a) out_off[a] = h_items[a].out_off;
    std::vector<long long> prg_bound((size_t)nl, 0);
    const int n_threads = std::max(1, std::min(ctx->n_workers, 8));
    auto convert = [&](int lo, int hi) {
        std::vector<int> order;  // global node indices of the locus in local order
        for (int i = lo; i < hi; ++i) {
            const int l = l_begin + i;
            LocusResult &L = res->loci[l];
            L.status = h_loci_back[i].status;
            if (L.status != MPRG_LOCUS_OK) continue;
            order.clear();
            order.push_back(i);  // the root of locus i is node i
            L.nodes.clear();
            L.row_pool.clear();
            long long bound = 0;
            for (size_t k = 0; k < order.size(); ++k) {
                const DNode &g = h_nodes[order[k]];
                HNode h;
                h.kind = g.kind;
                h.parent = -1;  // set by the parent below
                h.level = g.level;
                h.c0 = g.c0;
                h.c1 = g.c1;
                h.row_off = -1;
                h.n_rows = g.n_rows;
                h.first_child = g.n_children ? (int)order.size() : -1;
                h.n_children = g.n_children;
                h.allele_first = g.allele_first;
                h.allele_count = g.allele_count;
                if (g.kind == MPRG_NODE_LEAF) bound += (long long)g.allele_count * (g.c1 - g.c0 + 12);
                L.nodes.push_back(h);
                for (int c = 0; c < g.n_children; ++c) order.push_back(g.first_child + c);
            }
            // parents and row subsets: interval children share their parent's rows, cluster children own theirs
            for (size_t k = 0; k < order.size(); ++k) {
                HNode &h = L.nodes[k];
                const DNode &g = h_nodes[order[k]];
                for (int c = 0; c < h.n_children; ++c) {
                    HNode &ch = L.nodes[h.first_child + c];
                    const DNode &gc = h_nodes[g.first_child + c];
                    ch.parent = (int)k;
                    if (gc.row_off < 0) {
                        ch.row_off = -1;
                    } else if (gc.row_off == g.row_off) {
                        ch.row_off = h.row_off;
                    } else {
                        ch.row_off = (long long)L.row_pool.size();
                        L.row_pool.insert(L.row_pool.end(), h_pool + gc.row_off, h_pool + gc.row_off + gc.n_rows);
                    }
                }
            }
            prg_bound[i] = bound;
        }
    };
    if (n_threads <= 1 || nl < 64) {
        convert(0, nl);
    } else {
        std::vector<std::thread> th;
        for (int t = 0; t < n_threads; ++t) {
            const int lo = (int)((long long)nl * t / n_threads), hi = (int)((long long)nl * (t + 1) / n_threads);
            if (t + 1 == n_threads) convert(lo, hi);
            else th.emplace_back(convert, lo, hi);
        }
        for (auto &t : th) t.join();
    }
    TRACE("dev: node tables");
    assemble_prgs(batch, res, l_begin, l_end, out_off.data(), h_len, h_out, prg_bound.data(), n_threads);
    TRACE("dev: prg strings");
    if (allow_trace || trace_all) {
        trace.report("mprg_build (device-resident loop)");
        g_trace = nullptr;
    }
    return MPRG_OK;
}

}  // namespace mprg
