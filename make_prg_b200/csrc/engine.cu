// Host side of the hot path: the level-synchronous replacement of NodeFactory.build's recursion
// (make_prg/recursion_tree.py:401-471) and of kmeans_cluster_seqs' control flow
// (make_prg/from_msa/cluster_sequences.py:211-296), PRG emission (recursion_tree.py:194-300,
// prg_builder.py:100-110) and the C-ABI entry points built on them.
//
// At every recursion depth all pending sub-alignments of all loci form ONE device batch:
//   level_run()      scan -> classify -> partition -> demote            (scan.cu, partition.cu)
//   cluster_level()  unpack -> dedupe -> k-mer counts -> [refcheck, KMeans] x K  (cluster.cu, kmeans.cu)
// The host only keeps the tree bookkeeping (nodes, row subsets, child tasks), the cluster-merge
// ordering rules and the string assembly.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <unordered_set>

#include "engine.cuh"

using namespace mprg;

namespace mprg {

thread_local PhaseTrace *g_trace = nullptr;

// ---- MT19937 as numpy's RandomState(seed) / random_sample ---------------------------------------
static void randomstate_doubles(uint32_t seed, double *out, int count) {
    const int N = 624, M = 397;
    uint32_t mt[N];
    mt[0] = seed;
    for (int i = 1; i < N; ++i) mt[i] = 1812433253u * (mt[i - 1] ^ (mt[i - 1] >> 30)) + (uint32_t)i;
    int idx = N;
    auto gen = [&]() -> uint32_t {
        if (idx >= N) {
            for (int k = 0; k < N; ++k) {
                const uint32_t y = (mt[k] & 0x80000000u) | (mt[(k + 1) % N] & 0x7fffffffu);
                mt[k] = mt[(k + M) % N] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
            }
            idx = 0;
        }
        uint32_t y = mt[idx++];
        y ^= y >> 11;
        y ^= (y << 7) & 0x9d2c5680u;
        y ^= (y << 15) & 0xefc60000u;
        y ^= y >> 18;
        return y;
    };
    for (int i = 0; i < count; ++i) {
        const uint32_t a = gen() >> 5, b = gen() >> 6;
        out[i] = (a * 67108864.0 + b) / 9007199254740992.0;
    }
}

static bool g_rand_uploaded = false;
static std::mutex g_rand_mutex;  // first builds of two contexts may arrive together (device.BuildPipeline)
int ensure_rand(mprg_ctx *ctx) {
    std::lock_guard<std::mutex> lock(g_rand_mutex);
    if (g_rand_uploaded) return MPRG_OK;
    double r[KM_RAND_COUNT];
    randomstate_doubles(2u, r, KM_RAND_COUNT);
    MPRG_CUDA(ctx, kmeans_upload_rand(r));
    g_rand_uploaded = true;
    return MPRG_OK;
}

// ---- allele extraction ---------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
extract_kernel(const uint8_t *__restrict__ packed, const ExtractItem *__restrict__ items, int n_items,
               uint8_t *__restrict__ out, int *__restrict__ out_len) {
    // one warp per item: ordered compaction of the non-gap symbols, 32 columns at a time
    const int lane = threadIdx.x & 31;
    const int it = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (it >= n_items) return;
    const ExtractItem e = items[it];
    const uint8_t *row = packed + e.base + (long long)e.row * e.stride;
    const char *alphabet = MPRG_ALPHABET;
    int len = 0;
    for (int c0 = e.c0; c0 < e.c1; c0 += 32) {
        const int c = c0 + lane;
        int sym = SYM_GAP;
        if (c < e.c1) {
            sym = packed_sym(row, c);
        }
        const unsigned keep = __ballot_sync(0xffffffffu, sym != SYM_GAP);
        if (sym != SYM_GAP) out[e.out_off + len + __popc(keep & ((1u << lane) - 1u))] = (uint8_t)alphabet[sym];
        len += __popc(keep);
    }
    if (lane == 0) out_len[it] = len;
}

cudaError_t launch_extract(cudaStream_t s, const uint8_t *packed, const ExtractItem *items, int n_items, uint8_t *out,
                           int *out_len) {
    if (n_items <= 0) return cudaSuccess;
    extract_kernel<<<(n_items + 3) / 4, 128, 0, s>>>(packed, items, n_items, out, out_len);
    return cudaGetLastError();
}

// ---- clustering of a level ------------------------------------------------------------------------
struct ClusterOut {
    int n_ungapped = 0, n_gapped = 0;
    // task-local row position of each distinct ungapped sequence (first-seen order) and its ungapped
    // length: n_ungapped entries each, views into the context's pinned result buffer (valid until the next
    // clustering level on that context) -- a level has tens of thousands of tasks, no vector per task
    const int *leaders = nullptr;
    const int *leader_len = nullptr;
    bool no_clustering = true;    // ClusteringResult.no_clustering
    bool clustered = false;       // kmeans_cluster_seqs was evaluated for this task
    int n_labels = 0;             // KMeans clusters (labels 0..n_labels-1) when !no_clustering
    std::vector<int> assign;      // label of each distinct LONG sequence, in first-seen order
    std::vector<int> group;       // per task row: distinct ungapped sequence index (only when requested)
    std::vector<int> ulen;        // per task row: ungapped length (only when requested)
};

// ClusteringResult.clustered_ids as task-local row positions, from the per-row groups:
// extract_clusters (cluster_sequences.py:114-133) + one cluster per distinct small sequence +
// merge_clusters (:194-208).  Needs o.group.
static void clusters_from_rows(const ClusterOut &o, int kmer_size, std::vector<std::vector<int>> &out) {
    out.clear();
    const int R = (int)o.group.size();
    std::vector<int> long_index(o.n_ungapped, -1), small_index(o.n_ungapped, -1);
    int n_long = 0, n_small = 0;
    for (int g = 0; g < o.n_ungapped; ++g) {
        if (o.leader_len[g] >= kmer_size) long_index[g] = n_long++;
        else small_index[g] = n_small++;
    }
    if (o.no_clustering) {
        // all long ids in distinct-sequence order, then all small ones; first record moved to the front
        std::vector<std::vector<int>> lg(n_long), sg(n_small);
        for (int r = 0; r < R; ++r) {
            const int g = o.group[r];
            if (long_index[g] >= 0) lg[long_index[g]].push_back(r);
            else sg[small_index[g]].push_back(r);
        }
        std::vector<int> all;
        for (auto &v : lg) all.insert(all.end(), v.begin(), v.end());
        for (auto &v : sg) all.insert(all.end(), v.begin(), v.end());
        auto it = std::find(all.begin(), all.end(), 0);
        if (it != all.end()) {
            all.erase(it);
            all.insert(all.begin(), 0);
        }
        out.push_back(std::move(all));
        return;
    }
    std::vector<std::vector<std::vector<int>>> per_label(o.n_labels);  // label -> long seq -> rows
    std::vector<std::vector<int>> cl(o.n_labels + n_small);
    // rows inside a KMeans cluster are ordered by (distinct sequence, row)
    std::vector<std::vector<int>> lg(n_long), sg(n_small);
    for (int r = 0; r < R; ++r) {
        const int g = o.group[r];
        if (long_index[g] >= 0) lg[long_index[g]].push_back(r);
        else sg[small_index[g]].push_back(r);
    }
    for (int j = 0; j < n_long; ++j) cl[o.assign[j]].insert(cl[o.assign[j]].end(), lg[j].begin(), lg[j].end());
    for (int j = 0; j < n_small; ++j) cl[o.n_labels + j] = sg[j];
    size_t fi = 0;
    for (size_t a = 0; a < cl.size(); ++a)
        if (std::find(cl[a].begin(), cl[a].end(), 0) != cl[a].end()) fi = a;
    std::vector<int> first = cl[fi];
    first.erase(std::find(first.begin(), first.end(), 0));
    first.insert(first.begin(), 0);
    out.push_back(std::move(first));
    for (size_t a = 0; a < cl.size(); ++a)
        if (a != fi) out.push_back(std::move(cl[a]));
}

// The same partition as clusters_from_rows, as one cluster index per row (cluster 0 holds row 0, the
// others keep the order of ClusteringResult.clustered_ids); no per-cluster row lists are built.  The
// order of the rows inside a cluster is not represented: sub-alignments keep the input row order
// (recursion_tree.py:558-572).  Returns the number of clusters.
static int cluster_of_rows(const ClusterOut &o, int kmer_size, std::vector<int> &index_of_group,
                           std::vector<int> &cluster_of_row) {
    const int R = (int)o.group.size();
    cluster_of_row.resize(R);
    if (o.no_clustering) {
        std::fill(cluster_of_row.begin(), cluster_of_row.end(), 0);
        return R > 0 ? 1 : 0;
    }
    index_of_group.resize(o.n_ungapped);
    int n_long = 0, n_small = 0;
    for (int g = 0; g < o.n_ungapped; ++g) {
        if (o.leader_len[g] >= kmer_size) index_of_group[g] = o.assign[n_long++];
        else index_of_group[g] = o.n_labels + n_small++;
    }
    const int n_cl = o.n_labels + n_small;
    // the cluster of row 0 moves to the front (merge_clusters), the others keep their order
    const int first = R > 0 ? index_of_group[o.group[0]] : 0;
    for (int g = 0; g < o.n_ungapped; ++g) {
        const int c = index_of_group[g];
        index_of_group[g] = c == first ? 0 : (c < first ? c + 1 : c);
    }
    for (int r = 0; r < R; ++r) cluster_of_row[r] = index_of_group[o.group[r]];
    return n_cl;
}

// The clustering loop of kmeans_cluster_seqs (cluster_sequences.py:249-274) for the problems of one level,
// driven from the host: member lists, k-mer numbering (whole-grid kernels for deep problems), count matrices
// sized exactly (or by the bound F <= positions when every problem is small), then <= 9 rounds of
// [KMeans, one-reference-like check] with the loop state on the device.  The unpacked rows (d_G), the
// per-row groups and the leader lengths are the outputs of the de-duplication of the same level.
// fetch: copy the final loop states and assignments back (the device-resident loop reads them in place).
int run_problems_host(mprg_ctx *ctx, cudaStream_t s, const std::vector<HostProblem> &hp, const std::vector<int> &seq_rows,
                      int kmer_size, const uint8_t *d_G, const int *d_group, const int *d_leadlen, int *d_leader_u,
                      int *d_err, bool fetch, ProblemRun &run, const uint8_t *d_packed, const int *d_rows_arena) {
    const int np = (int)hp.size();
    DevBuf *B = ctx->d_c;
    std::vector<ClusterState> &st = run.st;
    std::vector<int> &h_assign = run.h_assign;
    st.assign(np, ClusterState());
    long long seq_total = 0;
    // ---- member lists + k-mer count matrices ----
    std::vector<KmerProb> kp(np);
    std::vector<MemberProb> mp(np);
    long long useq_total = 0, ints_total = 0, tab_total = 0, x_total = 0;
    long long maj_total = 0, assign_total = 0, memoff_total = 0, memrows_total = 0;
    std::vector<int> big_ref_q;  // deep loci: one-reference-like check with the whole grid
    for (int q = 0; q < np; ++q) {
        const HostProblem &p = hp[q];
        const int w = p.w;
        if (p.P > 0x3fffffffLL) MPRG_FAIL(ctx, MPRG_E_BAD_ARG, "clustering problem too large");
        KmerProb &k = kp[q];
        k.g_off = p.g_off;
        k.w = w;
        k.n = p.n;
        k.seq_off = (int)seq_total;
        seq_total += p.n;
        k.useq_off = useq_total;
        useq_total += (long long)p.n * w;
        k.pos_off = ints_total;
        ints_total += 2LL * p.n + 1 + 2 * p.P;
        int T = 64;
        while (T < 2 * p.P) T <<= 1;
        k.T = T;
        k.tab_off = tab_total;
        tab_total += T;
        k.Pmax = (int)p.P;
        k.big = p.P >= KMER_BIG_POSITIONS ? 1 : 0;
        k.pad = 0;
        k.x_off = 0;  // set below, once the number of distinct k-mers is known
        MemberProb &m = mp[q];
        m.row_off = p.row_off;
        m.R = p.R;
        m.n_groups = p.n_groups;
        m.k = kmer_size;
        m.mem_off = (int)memoff_total;
        m.mem_rows_off = (int)memrows_total;
        ClusterState &c = st[q];
        memset(&c, 0, sizeof(c));
        c.K = 1;
        c.n = p.n;
        c.w = w;
        c.g_off = p.g_off;
        c.mem_off = m.mem_off;
        c.mem_rows_off = m.mem_rows_off;
        memoff_total += p.n + 1;
        memrows_total += p.R;
        c.assign_off = (int)assign_total;
        assign_total += p.n;
        c.maj_off = maj_total;
        if ((long long)p.R * w >= REFCHECK_BIG_SYMBOLS) {
            c.big_ref = 1 + (int)big_ref_q.size();
            big_ref_q.push_back(q);
            maj_total += 10LL * w;  // one majority string per cluster
        } else {
            maj_total += w;
        }
    }
    // 7 kprobs|mprobs, 8 useq, 9 ints, 10 keys, 11 ming, 12 X, 13 states|F, 14 seq_rows|mem|assign|newlab|maj, 15 kmeans scratch
    MPRG_CUDA(ctx, B[7].reserve(sizeof(KmerProb) * np + sizeof(MemberProb) * np));
    MPRG_CUDA(ctx, B[8].reserve((size_t)useq_total));
    MPRG_CUDA(ctx, B[9].reserve(sizeof(int) * ints_total));
    MPRG_CUDA(ctx, B[10].reserve(sizeof(uint64_t) * tab_total));
    MPRG_CUDA(ctx, B[11].reserve(sizeof(int) * tab_total));
    MPRG_CUDA(ctx, B[13].reserve(sizeof(ClusterState) * np + sizeof(int) * 2 * np + KM_GROUP_WORDS * sizeof(unsigned) +
                                  sizeof(int) * REFCHECK_FLAG_INTS * big_ref_q.size() + 64));
    const size_t o_seqrows = 0;
    const size_t o_memoff = o_seqrows + sizeof(int) * seq_rows.size();
    const size_t o_memrows = o_memoff + sizeof(int) * memoff_total;
    const size_t o_assign = o_memrows + sizeof(int) * memrows_total;
    const size_t o_newlab = o_assign + sizeof(int) * assign_total;
    const size_t o_maj = o_newlab + sizeof(int) * assign_total;
    MPRG_CUDA(ctx, B[14].reserve(o_maj + (size_t)maj_total + 16));
    uint8_t *b14 = B[14].as<uint8_t>();
    int *d_seqrows = reinterpret_cast<int *>(b14 + o_seqrows);
    int *d_memoff = reinterpret_cast<int *>(b14 + o_memoff);
    int *d_memrows = reinterpret_cast<int *>(b14 + o_memrows);
    int *d_assign = reinterpret_cast<int *>(b14 + o_assign);
    int *d_newlab = reinterpret_cast<int *>(b14 + o_newlab);
    uint8_t *d_maj = b14 + o_maj;
    KmerProb *d_kp = B[7].as<KmerProb>();
    MemberProb *d_mp = reinterpret_cast<MemberProb *>(d_kp + np);
    MPRG_CUDA(ctx, mprg::copy_h2d(ctx, d_kp, kp.data(), sizeof(KmerProb) * np, s));
    MPRG_CUDA(ctx, mprg::copy_h2d(ctx, d_mp, mp.data(), sizeof(MemberProb) * np, s));
    MPRG_CUDA(ctx, mprg::copy_h2d(ctx, d_seqrows, seq_rows.data(), sizeof(int) * seq_rows.size(), s));
    ClusterState *d_states = B[13].as<ClusterState>();
    int *d_F = reinterpret_cast<int *>(d_states + np);
    int *d_tickets = d_F + np;  // per-problem "initialisations finished" counters of kmeans_kernel
    // barriers and broadcast slots of launch_kmeans_group (8-byte aligned)
    unsigned *d_bars = reinterpret_cast<unsigned *>(d_tickets + np);  // d_F + 2 * np: 8-byte aligned
    int *d_refflags = reinterpret_cast<int *>(d_bars + KM_GROUP_WORDS);
    MPRG_CUDA(ctx, cudaMemsetAsync(d_tickets, 0,
                                   sizeof(int) * np + KM_GROUP_WORDS * sizeof(unsigned) +
                                       sizeof(int) * REFCHECK_FLAG_INTS * big_ref_q.size(),
                                   s));
    int refcheck_round = 1;  // the check of round r sees up to r clusters
    auto refcheck_all = [&]() -> cudaError_t {
        cudaError_t e = launch_refcheck(s, d_states, np, d_G, d_memoff, d_memrows, d_assign, d_maj, 10);
        ctx->launches++;
        ctx->path_counts[MPRG_PATH_REFCHECK_CTA]++;
        for (size_t b = 0; b < big_ref_q.size() && e == cudaSuccess; ++b) {
            const int q = big_ref_q[b];
            const HostProblem &p = hp[q];
            // loci over ACGT- only (no N, no RYKMSW, no even code): bit-sliced counting on the packed rows;
            // MPRG_REFCHECK_BYTES=1 keeps the byte-wise kernels (the checked alternative)
            static const bool bytes_only = getenv("MPRG_REFCHECK_BYTES") != nullptr;
            if (d_packed && !(p.alpha_flags & (2 | 4 | 8)) && !bytes_only && p.R <= 65535) {  // 16-bit count fields
                DTask t;
                t.base = p.base;
                t.stride = p.stride;
                t.rows_off = p.rows_off;
                t.n_rows = p.R;
                t.c0 = p.c0;
                t.c1 = p.c0 + p.w;
                t.col_off = t.iv_off = t.flags = 0;
                const int n_words = (((t.c1 + 31) >> 5) - (t.c0 >> 5)) * 4;
                cudaError_t e2 = ctx->d_ref.reserve(sizeof(int) * (size_t)refgrid_scratch_ints(p.R, p.w, n_words, 10));
                if (e2 != cudaSuccess) return e2;
                int *flags = d_refflags + REFCHECK_FLAG_INTS * b;
                e = launch_refcheck_grid(s, d_states, q, t, 10, d_packed, d_rows_arena, d_memoff, d_memrows, d_assign, d_maj,
                                         ctx->d_ref.as<int>(), p.R, flags);
                if (e == cudaSuccess) e = launch_refcheck_big_control(s, d_states, q, 10, flags);
                ctx->launches += 6;
            } else {
                e = launch_refcheck_big(s, d_states, q, st[q].w, st[q].n, d_G, d_memoff, d_memrows,
                                        d_assign, d_maj, 10, d_refflags + REFCHECK_FLAG_INTS * b);
                ctx->launches += 3;
            }
            ctx->path_counts[MPRG_PATH_REFCHECK_GRID]++;
            if (refcheck_round > 1) ctx->path_counts[MPRG_PATH_REFCHECK_GRID_MULTI]++;
        }
        return e;
    };
    // d_leader_u is free after dedupe: scratch for the group -> long-sequence map
    MPRG_CUDA(ctx, launch_members(s, d_mp, np, d_group, d_leadlen, d_leader_u, d_memoff, d_memrows));
    MPRG_CUDA(ctx, launch_kmer(s, d_kp, np, d_seqrows, d_G, kmer_size, B[8].as<uint8_t>(),
                               B[9].as<int>(), B[10].as<uint64_t>(), B[11].as<int>(), d_F, d_err));
    ctx->launches += 2;
    // deep loci: problems with very many k-mer positions go through the whole-grid kernels, one by one
    for (int q = 0; q < np; ++q) {
        if (!kp[q].big) continue;
        const size_t n_chunks = ((size_t)kp[q].Pmax + 4095) / 4096;
        MPRG_CUDA(ctx, B[6].reserve(sizeof(int) * (n_chunks + 1)));
        MPRG_CUDA(ctx, launch_kmer_big(s, d_kp, q, &kp[q], d_seqrows, d_G, kmer_size,
                                       B[8].as<uint8_t>(), B[9].as<int>(), B[10].as<uint64_t>(), B[11].as<int>(),
                                       B[6].as<int>(), d_F + q, d_err));
        ctx->launches += 6;
        ctx->path_counts[MPRG_PATH_KMER_GRID]++;
    }
    // The number of distinct k-mers F of a problem sizes its count matrix and KMeans scratch.  A level
    // of small problems (a pangenome level: n <= 8, <= 241 positions) does not wait for it: it lays its
    // matrices out for the upper bound F <= P (every k-mer position distinct) and the device passes the
    // real F from the numbering kernel to the loop state; levels with a big problem fetch F first.
    bool bounded = !getenv("MPRG_EXACT_F");
    {
        long long bound_elems = 0;
        for (int q = 0; q < np && bounded; ++q) {
            if (kp[q].big || (long long)hp[q].n * hp[q].P >= KMEANS_BIG_ELEMENTS) bounded = false;
            bound_elems += (long long)hp[q].n * hp[q].P;
        }
        if (bound_elems > (16LL << 20)) bounded = false;  // 128 MB of doubles for the level
    }
    std::vector<int> h_F(np + 1, 0);
    if (bounded) {
        for (int q = 0; q < np; ++q) h_F[q] = (int)hp[q].P;
    } else {
        MPRG_CUDA(ctx, mprg::copy_d2h(ctx, h_F.data(), d_F, sizeof(int) * np, s));
        MPRG_CUDA(ctx, mprg::copy_d2h(ctx, &h_F[np], d_err, sizeof(int), s));
        MPRG_CUDA(ctx, cudaStreamSynchronize(s));
        if (h_F[np]) MPRG_FAIL(ctx, MPRG_E_INTERNAL, "hash collision detected while numbering k-mers");
    }
    TRACE("cl: kmer setup+run+sync");

    // ---- count matrices, sized exactly now that every F is known ----
    long long max_P = 0;
    for (int q = 0; q < np; ++q) {
        kp[q].x_off = st[q].x_off = x_total;
        x_total += (long long)hp[q].n * h_F[q];
        if (kp[q].big && (size_t)h_F[q] * sizeof(int) <= 200 * 1024) kp[q].big |= 2;
        if (!(kp[q].big & 2)) max_P = std::max(max_P, hp[q].P);
    }
    if (g_trace) {
        long long big_n = 0, big_F = 0, big_P = 0;
        for (int q = 0; q < np; ++q) {
            big_n = std::max<long long>(big_n, hp[q].n);
            big_F = std::max<long long>(big_F, h_F[q]);
            big_P = std::max<long long>(big_P, hp[q].P);
        }
        fprintf(stderr, "[mprg trace] clustering level: %d problems, max n %lld, max F %lld, max positions %lld, X %.1f MB\n",
                np, big_n, big_F, big_P, 8e-6 * (double)x_total);
    }
    // ---- KMeans loop (cluster_sequences.py:256-274) ----
    long long kmd_total = 0, kmi_total = 0;
    for (int q = 0; q < np; ++q) {
        st[q].F = h_F[q];
        st[q].big = (long long)st[q].n * st[q].F >= KMEANS_BIG_ELEMENTS ? 1 : 0;
        st[q].kmd_off = kmd_total;
        st[q].kmi_off = kmi_total;
        kmd_total += kmeans_dscratch_doubles(st[q].n, st[q].F);
        kmi_total += kmeans_iscratch_ints(st[q].n);
    }
    const double want = 8.0 * (double)x_total + 8.0 * (double)kmd_total + 4.0 * (double)kmi_total;
    if (want > 1e9) {  // cudaMemGetInfo is a slow, serialising driver call: only deep loci ask
        size_t free_b = 0, total_b = 0;
        cudaMemGetInfo(&free_b, &total_b);
        const double have = (double)free_b + (double)B[12].cap + (double)B[15].cap;
        if (want > 0.9 * have)
            MPRG_FAIL(ctx, MPRG_E_BAD_ARG, "k-mer count matrices and KMeans scratch of this level do not fit in device memory");
    }
    MPRG_CUDA(ctx, B[12].reserve(sizeof(double) * std::max<long long>(x_total, 1)));
    MPRG_CUDA(ctx, cudaMemsetAsync(B[12].p, 0, sizeof(double) * x_total, s));
    MPRG_CUDA(ctx, mprg::copy_h2d(ctx, d_kp, kp.data(), sizeof(KmerProb) * np, s));
    MPRG_CUDA(ctx, launch_kmer_fill(s, d_kp, np, max_P, B[9].as<int>(), d_F, B[12].as<double>()));
    ctx->launches++;
    for (int q = 0; q < np; ++q)
        if (kp[q].big & 2) {
            MPRG_CUDA(ctx, launch_kmer_fill_big(s, d_kp, q, &kp[q], h_F[q], B[9].as<int>(), B[12].as<double>()));
            ctx->launches++;
        }
    MPRG_CUDA(ctx, B[15].reserve(sizeof(double) * kmd_total + sizeof(int) * kmi_total + 64));
    double *d_kmd = B[15].as<double>();
    int *d_kmi = reinterpret_cast<int *>(d_kmd + kmd_total);
    MPRG_CUDA(ctx, mprg::copy_h2d(ctx, d_states, st.data(), sizeof(ClusterState) * np, s));
    if (bounded) {
        MPRG_CUDA(ctx, launch_set_features(s, d_states, d_F, np));
        ctx->launches++;
    }
    MPRG_CUDA(ctx, cudaMemsetAsync(d_assign, 0, sizeof(int) * assign_total, s));
    const int MAX_CLUSTERS = 10;
    MPRG_CUDA(ctx, refcheck_all());
    // every problem that ever runs KMeans runs it in the K == 2 round: centre its data once, now
    MPRG_CUDA(ctx, launch_kmeans_prepare(s, d_states, np, B[12].as<double>(), d_kmd, d_kmi));
    ctx->launches++;
    std::vector<int> big_q;  // deep loci: these problems get the whole GPU, one after the other
    for (int q = 0; q < np; ++q)
        if (st[q].big) big_q.push_back(q);
    // A problem with n distinct sequences stops at K == n (cluster_sequences.py:257-259), so its last
    // KMeans round is K = n - 1: the level needs rounds 2 .. max n - 1, not always 2 .. 10 (a deep level
    // of a pangenome batch has max n = 3 or 4: one or two rounds instead of nine)
    int max_n = 0;
    long long max_elements = 0;
    for (int q = 0; q < np; ++q) {
        max_n = std::max(max_n, hp[q].n);
        if (!st[q].big) max_elements = std::max(max_elements, (long long)st[q].n * st[q].F);
    }
    const int last_round = std::min(MAX_CLUSTERS, max_n - 1);
    for (int round = 2; round <= last_round; ++round) {
        for (int q : big_q) {
            MPRG_CUDA(ctx, launch_kmeans_group(s, d_states, q, B[12].as<double>(), d_kmd, d_kmi, d_assign,
                                               d_newlab, d_bars, round == 2, ctx->sm_count));
            ctx->launches += round == 2 ? 2 : 1;
            ctx->path_counts[MPRG_PATH_KMEANS_GROUP]++;
        }
        refcheck_round = round;
        ctx->path_counts[MPRG_PATH_KMEANS_CTA]++;
        MPRG_CUDA(ctx, launch_kmeans(s, d_states, np, B[12].as<double>(), d_kmd, d_kmi, d_assign, d_newlab,
                                     d_tickets, max_elements));
        MPRG_CUDA(ctx, refcheck_all());
        ctx->launches++;
    }
    run.d_states = d_states;
    run.d_assign = d_assign;
    if (fetch) {
        h_assign.resize((size_t)assign_total);
        MPRG_CUDA(ctx, mprg::copy_d2h(ctx, st.data(), d_states, sizeof(ClusterState) * np, s));
        MPRG_CUDA(ctx, mprg::copy_d2h(ctx, h_assign.data(), d_assign, sizeof(int) * assign_total, s));
        if (bounded) MPRG_CUDA(ctx, mprg::copy_d2h(ctx, &h_F[np], d_err, sizeof(int), s));
        MPRG_CUDA(ctx, cudaStreamSynchronize(s));
        if (bounded && h_F[np]) MPRG_FAIL(ctx, MPRG_E_INTERNAL, "hash collision detected while numbering k-mers");
    }
    return MPRG_OK;
}


// Runs kmeans_cluster_seqs for every task of a level.
//   want_clusters[t] == 0   only the de-duplication outputs are needed
//   want_rows[t] != 0       also return the per-row group / ungapped length (O(rows) over PCIe)
//   rows_if_clustered       return the per-row groups of every task that ends up clustered
static int cluster_level(mprg_ctx *ctx, const mprg_batch *batch, const mprg_task *h_tasks, int n_tasks,
                         const int32_t *h_rows, long long n_row_entries, int kmer_size,
                         const uint8_t *want_clusters, const uint8_t *want_rows, bool rows_if_clustered,
                         std::vector<ClusterOut> &out, bool skip_if_issues = false) {
    out.assign(n_tasks, ClusterOut());
    if (n_tasks == 0) return MPRG_OK;
    cudaSetDevice(ctx->device);
    cudaStream_t s = ctx->stream;
    int rc = ensure_rand(ctx);
    if (rc != MPRG_OK) return rc;

    std::vector<DTask> tasks(n_tasks);
    std::vector<long long> g_off(n_tasks), row_off(n_tasks);
    std::vector<int> h_R(n_tasks);
    long long g_total = 0, row_total = 0;
    for (int i = 0; i < n_tasks; ++i) {
        const mprg_task &ht = h_tasks[i];
        if (ht.locus < 0 || ht.locus >= batch->n_loci || ht.c0 < 0 || ht.c1 < ht.c0 ||
            ht.c1 > batch->n_cols[ht.locus] || ht.n_rows < 0 ||
            (ht.rows_off < 0 && ht.n_rows > batch->n_rows[ht.locus]) ||
            (ht.rows_off >= 0 && (long long)ht.rows_off + ht.n_rows > n_row_entries))
            MPRG_FAIL(ctx, MPRG_E_BAD_ARG, "cluster task out of range");
        DTask &t = tasks[i];
        t.base = batch->base[ht.locus];
        t.stride = batch->stride[ht.locus];
        t.rows_off = ht.rows_off;
        t.n_rows = ht.n_rows;
        t.c0 = ht.c0;
        t.c1 = ht.c1;
        t.col_off = t.iv_off = t.flags = 0;
        g_off[i] = g_total;
        row_off[i] = row_total;
        h_R[i] = ht.n_rows;
        g_total += (long long)ht.n_rows * (ht.c1 - ht.c0);
        row_total += ht.n_rows;
    }
    if (row_total > 0x7fffffffLL) MPRG_FAIL(ctx, MPRG_E_BAD_ARG, "too many rows in one clustering level");
    DevBuf *B = ctx->d_c;
    // 0 tasks, 1 g_off, 2 row_off, 3 G, 4 sig,
    // 5 ints: leader_u | leader_g | group | ulen | leaders | leader_len | nu | ng | lead_off(n+1) | err | R(n)
    MPRG_CUDA(ctx, B[0].reserve(sizeof(DTask) * n_tasks));
    MPRG_CUDA(ctx, B[1].reserve(sizeof(long long) * n_tasks));
    MPRG_CUDA(ctx, B[2].reserve(sizeof(long long) * n_tasks));
    MPRG_CUDA(ctx, B[3].reserve((size_t)std::max<long long>(g_total, 1)));
    MPRG_CUDA(ctx, B[4].reserve(rowsig_bytes() * std::max<long long>(row_total, 1)));
    const long long n_int = 6 * row_total + 4LL * n_tasks + 2;
    MPRG_CUDA(ctx, B[5].reserve(sizeof(int) * n_int));
    MPRG_CUDA(ctx, ctx->d_rows.reserve(sizeof(int) * std::max<long long>(n_row_entries, 1)));
    MPRG_CUDA(ctx, mprg::copy_h2d(ctx, B[0].p, tasks.data(), sizeof(DTask) * n_tasks, s));
    MPRG_CUDA(ctx, mprg::copy_h2d(ctx, B[1].p, g_off.data(), sizeof(long long) * n_tasks, s));
    MPRG_CUDA(ctx, mprg::copy_h2d(ctx, B[2].p, row_off.data(), sizeof(long long) * n_tasks, s));
    if (n_row_entries > 0)
        MPRG_CUDA(ctx, mprg::copy_h2d(ctx, ctx->d_rows.p, h_rows, sizeof(int) * n_row_entries, s));
    int *d_leader_u = B[5].as<int>();
    int *d_leader_g = d_leader_u + row_total;
    int *d_group = d_leader_g + row_total;
    int *d_ulen = d_group + row_total;
    int *d_leaders = d_ulen + row_total;
    int *d_leadlen = d_leaders + row_total;
    int *d_nu = d_leadlen + row_total;
    int *d_ng = d_nu + n_tasks;
    int *d_leadoff = d_ng + n_tasks;  // n_tasks + 1
    int *d_err = d_leadoff + n_tasks + 1;
    int *d_R = d_err + 1;
    MPRG_CUDA(ctx, cudaMemsetAsync(d_err, 0, sizeof(int), s));
    MPRG_CUDA(ctx, mprg::copy_h2d(ctx, d_R, h_R.data(), sizeof(int) * n_tasks, s));
    // a level that holds a deep task (config #4: 10,000 x 20,000) spreads every task over the grid
    long long biggest = 0;
    int max_rows = 0;
    for (int i = 0; i < n_tasks; ++i) {
        biggest = std::max(biggest, (long long)h_R[i] * (tasks[i].c1 - tasks[i].c0));
        max_rows = std::max(max_rows, h_R[i]);
    }
    const bool deep = (biggest >= DEDUPE_BIG_SYMBOLS || getenv("MPRG_FORCE_BIG_DEDUPE")) && n_tasks <= 65535;
    MPRG_CUDA(ctx, launch_unpack(s, batch->d_packed, B[0].as<DTask>(), n_tasks, ctx->d_rows.as<int>(),
                                 B[1].as<long long>(), B[3].as<uint8_t>(), deep ? (max_rows + 7) / 8 : 1));
    if (deep) {
        MPRG_CUDA(ctx, B[15].reserve((size_t)std::max<long long>(g_total, 1)));  // compacted rows; free again below
        MPRG_CUDA(ctx, launch_dedupe_big(s, B[0].as<DTask>(), n_tasks, max_rows, B[1].as<long long>(),
                                         B[3].as<uint8_t>(), B[15].as<uint8_t>(), B[2].as<long long>(), B[4].p,
                                         d_leader_u, d_leader_g, d_group, d_ulen, d_leaders, d_leadlen, d_nu, d_ng,
                                         d_err));
        ctx->launches += 3;
        ctx->path_counts[MPRG_PATH_DEDUPE_GRID]++;
    } else {
        MPRG_CUDA(ctx, launch_dedupe(s, B[0].as<DTask>(), n_tasks, max_rows, B[1].as<long long>(), B[3].as<uint8_t>(),
                                     B[2].as<long long>(), B[4].p, d_leader_u, d_leader_g, d_group, d_ulen,
                                     d_leaders, d_leadlen, d_nu, d_ng, d_err));
        ctx->launches += 1;  // two kernels: a warp per small task, a CTA per other task
    }
    // Results of the de-duplication in ONE trip when the level is small (nu | ng | err and the per-row
    // leader arrays as they are, O(rows)); big levels first fetch the counts, compact the leaders on the
    // device and fetch O(#distinct) -- a second synchronisation that a pangenome level does not need.
    const bool one_trip = 2LL * row_total * (long long)sizeof(int) <= (256LL << 10);  // root levels: 30 tasks of 200 rows per locus
    if (!one_trip) MPRG_CUDA(ctx, launch_scan_counts(s, d_nu, n_tasks, d_leadoff));
    ctx->launches += one_trip ? 2 : 3;
    TRACE("cl: setup+launch dedupe");
    // nu | ng | lead_off | err are contiguous: one small copy
    std::vector<int> h_small((size_t)(3LL * n_tasks + 2));
    std::vector<int> lead_base(n_tasks);  // where task i's leaders start in h_dense
    long long lead_total = 0;
    int *h_dense = nullptr;
    if (one_trip) {
        MPRG_CUDA(ctx, ctx->h_a.reserve(sizeof(int) * 2 * std::max<long long>(row_total, 1)));
        h_dense = ctx->h_a.as<int>();
        MPRG_CUDA(ctx, mprg::copy_d2h(ctx, h_dense, d_leaders, sizeof(int) * 2 * row_total, s));
        lead_total = row_total;  // leader lengths follow the leaders at this distance
        for (int i = 0; i < n_tasks; ++i) lead_base[i] = (int)row_off[i];
    }
    MPRG_CUDA(ctx, mprg::copy_d2h(ctx, h_small.data(), d_nu, sizeof(int) * h_small.size(), s));
    MPRG_CUDA(ctx, cudaStreamSynchronize(s));
    const int *h_nu = h_small.data();
    const int *h_ng = h_nu + n_tasks;
    const int *h_leadoff = h_ng + n_tasks;
    if (h_leadoff[n_tasks + 1]) MPRG_FAIL(ctx, MPRG_E_INTERNAL, "hash collision detected while de-duplicating rows");
    if (!one_trip) {
        lead_total = h_leadoff[n_tasks];
        for (int i = 0; i < n_tasks; ++i) lead_base[i] = h_leadoff[i];
        // dense leaders | leader lengths
        MPRG_CUDA(ctx, B[6].reserve(sizeof(int) * 2 * std::max<long long>(lead_total, 1)));
        int *d_dense = B[6].as<int>();
        MPRG_CUDA(ctx, launch_gather2(s, B[2].as<long long>(), d_nu, d_leadoff, n_tasks, d_leaders, d_leadlen,
                                      d_dense, d_dense + lead_total));
        ctx->launches++;
        MPRG_CUDA(ctx, ctx->h_a.reserve(sizeof(int) * 2 * std::max<long long>(lead_total, 1)));
        h_dense = ctx->h_a.as<int>();
        MPRG_CUDA(ctx, mprg::copy_d2h(ctx, h_dense, d_dense, sizeof(int) * 2 * lead_total, s));
        MPRG_CUDA(ctx, cudaStreamSynchronize(s));
    }
    TRACE("cl: dedupe sync+D2H");

    // host: distinct sequences per task, trivial outcomes, KMeans problems -- O(#distinct)
    struct Prob {
        int task, n;
        long long P;
    };
    std::vector<Prob> probs;
    std::vector<int> seq_rows;  // leader rows of the long sequences of every problem
    for (int i = 0; i < n_tasks; ++i) {
        ClusterOut &o = out[i];
        o.n_ungapped = h_nu[i];
        o.n_gapped = h_ng[i];
        o.leaders = h_dense + lead_base[i];
        o.leader_len = h_dense + lead_total + lead_base[i];
        if (want_clusters && !want_clusters[i]) continue;
        if (h_R[i] == 0) continue;
        // NodeFactory._alignment_has_issues (recursion_tree.py:475-494) discards the clustering anyway
        if (skip_if_issues && (o.n_ungapped <= 2 || o.n_ungapped < o.n_gapped)) continue;
        o.clustered = true;
        int n = 0;
        long long P = 0;
        for (int g = 0; g < o.n_ungapped; ++g)
            if (o.leader_len[g] >= kmer_size) {
                ++n;
                P += o.leader_len[g] - kmer_size + 1;
            }
        if (n <= 2) continue;  // too few sequences: single cluster (no_clustering stays true)
        probs.push_back(Prob{i, n, P});
    }
    const int np = (int)probs.size();
    TRACE("cl: host grouping");

    std::vector<ClusterState> st;
    std::vector<int> h_assign;
    if (np > 0) {
        std::vector<HostProblem> hp(np);
        for (int q = 0; q < np; ++q) {
            const Prob &p = probs[q];
            const mprg_task &ht = h_tasks[p.task];
            HostProblem &h = hp[q];
            h.task = p.task;
            h.n = p.n;
            h.P = p.P;
            h.w = ht.c1 - ht.c0;
            h.R = ht.n_rows;
            h.n_groups = out[p.task].n_ungapped;
            h.g_off = g_off[p.task];
            h.row_off = row_off[p.task];
            h.base = batch->base[ht.locus];
            h.stride = batch->stride[ht.locus];
            h.rows_off = ht.rows_off;
            h.c0 = ht.c0;
            h.alpha_flags = batch->flags[ht.locus];
            const ClusterOut &o = out[p.task];
            for (int g = 0; g < o.n_ungapped; ++g)
                if (o.leader_len[g] >= kmer_size) seq_rows.push_back(o.leaders[g]);
        }
        ProblemRun run;
        rc = run_problems_host(ctx, s, hp, seq_rows, kmer_size, B[3].as<uint8_t>(), d_group, d_leadlen, d_leader_u, d_err,
                               true, run, batch->d_packed, ctx->d_rows.as<int>());
        if (rc != MPRG_OK) return rc;
        st.swap(run.st);
        h_assign.swap(run.h_assign);
        const int MAX_CLUSTERS = 10;
        TRACE("cl: kmeans loop+sync");
        for (int q = 0; q < np; ++q) {
            const Prob &p = probs[q];
            ClusterOut &o = out[p.task];
            const ClusterState &c = st[q];
            if (c.status != 1) MPRG_FAIL(ctx, MPRG_E_INTERNAL, "clustering loop did not terminate");
            const int K = c.K;
            if (K == 1 || K == p.n) continue;  // no_clustering (cluster_sequences.py:276)
            o.no_clustering = false;
            o.n_labels = std::min(K, MAX_CLUSTERS);  // K == 11 keeps the 10-cluster assignment
            o.assign.assign(h_assign.begin() + c.assign_off, h_assign.begin() + c.assign_off + p.n);
            for (int a : o.assign)
                if (a < 0 || a >= o.n_labels) MPRG_FAIL(ctx, MPRG_E_INTERNAL, "label out of range");
        }
    }
    // a task that is clustered but has small sequences only next to <= 2 long ones keeps no_clustering;
    // clusters made only of small-sequence groups are not produced by the reference in that branch

    // ---- per-row groups for the tasks that need them ----
    std::vector<int> need;
    for (int i = 0; i < n_tasks; ++i) {
        const bool w = (want_rows && want_rows[i]) || (rows_if_clustered && out[i].clustered && !out[i].no_clustering);
        if (w && h_R[i] > 0) need.push_back(i);
    }
    if (!need.empty()) {
        const int nn = (int)need.size();
        std::vector<long long> src(nn);
        std::vector<int> cnt(nn), dst(nn);
        long long total = 0;
        for (int q = 0; q < nn; ++q) {
            src[q] = row_off[need[q]];
            cnt[q] = h_R[need[q]];
            dst[q] = (int)total;
            total += cnt[q];
        }
        MPRG_CUDA(ctx, B[6].reserve(sizeof(int) * 2 * total + sizeof(long long) * nn + sizeof(int) * 2 * nn + 64));
        uint8_t *b6 = B[6].as<uint8_t>();
        long long *d_src = reinterpret_cast<long long *>(b6);
        int *d_cnt = reinterpret_cast<int *>(d_src + nn);
        int *d_dst = d_cnt + nn;
        int *d_out = d_dst + nn + ((nn & 1) ? 1 : 0);
        MPRG_CUDA(ctx, mprg::copy_h2d(ctx, d_src, src.data(), sizeof(long long) * nn, s));
        MPRG_CUDA(ctx, mprg::copy_h2d(ctx, d_cnt, cnt.data(), sizeof(int) * nn, s));
        MPRG_CUDA(ctx, mprg::copy_h2d(ctx, d_dst, dst.data(), sizeof(int) * nn, s));
        MPRG_CUDA(ctx, launch_gather2(s, d_src, d_cnt, d_dst, nn, d_group, d_ulen, d_out, d_out + total));
        ctx->launches++;
        MPRG_CUDA(ctx, ctx->h_b.reserve(sizeof(int) * 2 * total));
        int *h_out = ctx->h_b.as<int>();
        MPRG_CUDA(ctx, mprg::copy_d2h(ctx, h_out, d_out, sizeof(int) * 2 * total, s));
        MPRG_CUDA(ctx, cudaStreamSynchronize(s));
        for (int q = 0; q < nn; ++q) {
            ClusterOut &o = out[need[q]];
            o.group.assign(h_out + dst[q], h_out + dst[q] + cnt[q]);
            o.ulen.assign(h_out + total + dst[q], h_out + total + dst[q] + cnt[q]);
        }
    }
    TRACE("cl: rows D2H");
    return MPRG_OK;
}

}  // namespace mprg

// =================================================================================================
// C ABI: clustering entry points
// =================================================================================================
extern "C" int mprg_dedupe_rows(mprg_ctx *ctx, const mprg_batch *batch, const mprg_task *h_tasks,
                                int32_t n_tasks, const int32_t *h_rows, int64_t n_row_entries,
                                const int64_t *h_row_offsets, int32_t *h_group, int32_t *h_ungapped_len,
                                int32_t *h_n_ungapped, int32_t *h_n_gapped) {
    if (!ctx || !batch || n_tasks < 0 || (n_tasks > 0 && (!h_tasks || !h_row_offsets))) return MPRG_E_BAD_ARG;
    std::vector<ClusterOut> out;
    std::vector<uint8_t> want(std::max(n_tasks, 1), 0), rows(std::max(n_tasks, 1), 1);
    int rc = cluster_level(ctx, batch, h_tasks, n_tasks, h_rows, n_row_entries, 1, want.data(), rows.data(),
                           false, out);
    if (rc != MPRG_OK) return rc;
    for (int i = 0; i < n_tasks; ++i) {
        const int R = h_tasks[i].n_rows;
        if (h_group && R) memcpy(h_group + h_row_offsets[i], out[i].group.data(), sizeof(int) * R);
        if (h_ungapped_len && R) memcpy(h_ungapped_len + h_row_offsets[i], out[i].ulen.data(), sizeof(int) * R);
        if (h_n_ungapped) h_n_ungapped[i] = out[i].n_ungapped;
        if (h_n_gapped) h_n_gapped[i] = out[i].n_gapped;
    }
    return MPRG_OK;
}

extern "C" int mprg_cluster_tasks(mprg_ctx *ctx, const mprg_batch *batch, const mprg_task *h_tasks,
                                  int32_t n_tasks, const int32_t *h_rows, int64_t n_row_entries,
                                  int32_t kmer_size, const int64_t *h_row_offsets, int32_t *h_cluster,
                                  int32_t *h_n_clusters) {
    if (!ctx || !batch || n_tasks < 0 || kmer_size < 1 ||
        (n_tasks > 0 && (!h_tasks || !h_row_offsets || !h_cluster || !h_n_clusters)))
        return MPRG_E_BAD_ARG;
    std::vector<ClusterOut> out;
    std::vector<uint8_t> rows(std::max(n_tasks, 1), 1);
    int rc = cluster_level(ctx, batch, h_tasks, n_tasks, h_rows, n_row_entries, kmer_size, nullptr, rows.data(),
                           false, out);
    if (rc != MPRG_OK) return rc;
    std::vector<std::vector<int>> clusters;
    for (int i = 0; i < n_tasks; ++i) {
        if (h_tasks[i].n_rows == 0) {
            h_n_clusters[i] = 0;
            continue;
        }
        clusters_from_rows(out[i], kmer_size, clusters);
        h_n_clusters[i] = (int)clusters.size();
        for (size_t c = 0; c < clusters.size(); ++c)
            for (int r : clusters[c]) h_cluster[h_row_offsets[i] + r] = (int)c;
    }
    return MPRG_OK;
}

// mode 0: the engine's choice (CTA groups from KMEANS_BIG_ELEMENTS on), 1: one CTA per initialisation,
// 2: CTA groups
extern "C" int mprg_kmeans_mode(mprg_ctx *ctx, const double *h_X, int32_t n, int32_t F, int32_t K,
                                int32_t *h_labels, double *h_inertia, int32_t mode) {
    if (!ctx || !h_X || !h_labels || n < 1 || F < 1 || K < 1 || K > 10 || K > n || mode < 0 || mode > 2)
        return MPRG_E_BAD_ARG;
    cudaSetDevice(ctx->device);
    cudaStream_t s = ctx->stream;
    int rc = ensure_rand(ctx);
    if (rc != MPRG_OK) return rc;
    DevBuf *B = ctx->d_c;
    const long long nd = kmeans_dscratch_doubles(n, F), ni = kmeans_iscratch_ints(n);
    const bool group = mode == 2 || (mode == 0 && (long long)n * F >= KMEANS_BIG_ELEMENTS);
    MPRG_CUDA(ctx, B[12].reserve(sizeof(double) * (size_t)n * F));
    MPRG_CUDA(ctx, B[15].reserve(sizeof(double) * (nd + 1) + sizeof(int) * (ni + 2 * n + 4 + KM_GROUP_WORDS)));
    double *d_d = B[15].as<double>();
    double *d_inertia = d_d + nd;
    int *d_i = reinterpret_cast<int *>(d_inertia + 1);
    int *d_labels = d_i + ni;
    int *d_assign = d_labels + n;
    int *d_ticket = d_assign + n;  // 1 ticket (+ pad to 8 bytes) + the group communication area
    int *d_comm = d_ticket + 1 + ((ni + 2 * n + 1) & 1);
    MPRG_CUDA(ctx, cudaMemsetAsync(d_ticket, 0, sizeof(int) * (2 + KM_GROUP_WORDS), s));
    MPRG_CUDA(ctx, mprg::copy_h2d(ctx, B[12].p, h_X, sizeof(double) * (size_t)n * F, s));
    if (group) {
        ClusterState c;
        memset(&c, 0, sizeof(c));
        c.run_kmeans = 1;
        c.K = K;
        c.n = n;
        c.F = F;
        c.big = 1;
        MPRG_CUDA(ctx, B[13].reserve(sizeof(ClusterState)));
        MPRG_CUDA(ctx, mprg::copy_h2d(ctx, B[13].p, &c, sizeof(c), s));
        MPRG_CUDA(ctx, launch_kmeans_group(s, B[13].as<ClusterState>(), 0, B[12].as<double>(), d_d, d_i, d_assign,
                                           d_labels, reinterpret_cast<unsigned *>(d_comm), true,
                                           ctx->sm_count, d_inertia));
        ctx->launches += 2;
    } else {
        MPRG_CUDA(ctx, launch_kmeans_single(s, B[12].as<double>(), n, F, K, d_d, d_i, d_labels, d_inertia, d_ticket));
        ctx->launches++;
    }
    MPRG_CUDA(ctx, mprg::copy_d2h(ctx, h_labels, d_labels, sizeof(int) * n, s));
    double inertia = 0;
    MPRG_CUDA(ctx, mprg::copy_d2h(ctx, &inertia, d_inertia, sizeof(double), s));
    MPRG_CUDA(ctx, cudaStreamSynchronize(s));
    if (h_inertia) *h_inertia = inertia;
    return MPRG_OK;
}

extern "C" int mprg_kmeans(mprg_ctx *ctx, const double *h_X, int32_t n, int32_t F, int32_t K,
                           int32_t *h_labels, double *h_inertia) {
    return mprg_kmeans_mode(ctx, h_X, n, F, K, h_labels, h_inertia, 0);
}

extern "C" int mprg_kmer_counts(mprg_ctx *ctx, const mprg_batch *batch, const mprg_task *h_task,
                                const int32_t *h_rows, int32_t kmer_size, int32_t *n_seqs,
                                int32_t *n_kmers, double *h_counts, int64_t capacity) {
    if (!ctx || !batch || !h_task || kmer_size < 1 || !n_seqs || !n_kmers) return MPRG_E_BAD_ARG;
    std::vector<ClusterOut> out;
    uint8_t want = 0;
    mprg_task t = *h_task;
    const long long n_row_entries = t.rows_off >= 0 ? (long long)t.rows_off + t.n_rows : 0;
    int rc = cluster_level(ctx, batch, &t, 1, h_rows, n_row_entries, kmer_size, &want, nullptr, false, out);
    if (rc != MPRG_OK) return rc;
    cudaStream_t s = ctx->stream;
    DevBuf *B = ctx->d_c;
    const ClusterOut &o = out[0];
    std::vector<int> leaders;
    long long P = 0;
    for (int g = 0; g < o.n_ungapped; ++g)
        if (o.leader_len[g] >= kmer_size) {
            leaders.push_back(o.leaders[g]);
            P += o.leader_len[g] - kmer_size + 1;
        }
    const int n = (int)leaders.size();
    *n_seqs = n;
    *n_kmers = 0;
    if (n == 0) return MPRG_OK;
    const int w = t.c1 - t.c0;
    KmerProb k;
    memset(&k, 0, sizeof(k));
    k.g_off = 0;
    k.w = w;
    k.n = n;
    k.seq_off = 0;
    k.useq_off = 0;
    k.pos_off = 0;
    k.tab_off = 0;
    int T = 64;
    while (T < 2 * P) T <<= 1;
    k.T = T;
    k.Pmax = (int)P;
    k.x_off = 0;
    k.big = P >= KMER_BIG_POSITIONS ? 1 : 0;
    MPRG_CUDA(ctx, B[7].reserve(sizeof(KmerProb)));
    MPRG_CUDA(ctx, B[14].reserve(sizeof(int) * n));
    MPRG_CUDA(ctx, B[8].reserve((size_t)n * w + 1));
    MPRG_CUDA(ctx, B[9].reserve(sizeof(int) * (2LL * n + 1 + 2 * P)));
    MPRG_CUDA(ctx, B[10].reserve(sizeof(uint64_t) * T));
    MPRG_CUDA(ctx, B[11].reserve(sizeof(int) * T));
    MPRG_CUDA(ctx, B[13].reserve(sizeof(int) * 2));
    int *d_F = B[13].as<int>();
    MPRG_CUDA(ctx, cudaMemsetAsync(d_F, 0, sizeof(int) * 2, s));
    MPRG_CUDA(ctx, mprg::copy_h2d(ctx, B[7].p, &k, sizeof(k), s));
    MPRG_CUDA(ctx, mprg::copy_h2d(ctx, B[14].p, leaders.data(), sizeof(int) * n, s));
    if (k.big) {
        MPRG_CUDA(ctx, B[6].reserve(sizeof(int) * ((size_t)(P + 4095) / 4096 + 1)));
        MPRG_CUDA(ctx, launch_kmer_big(s, B[7].p, 0, &k, B[14].as<int>(), B[3].as<uint8_t>(), kmer_size,
                                       B[8].as<uint8_t>(), B[9].as<int>(), B[10].as<uint64_t>(), B[11].as<int>(),
                                       B[6].as<int>(), d_F, d_F + 1));
        ctx->launches += 6;
    } else {
        MPRG_CUDA(ctx, launch_kmer(s, B[7].p, 1, B[14].as<int>(), B[3].as<uint8_t>(), kmer_size, B[8].as<uint8_t>(),
                                   B[9].as<int>(), B[10].as<uint64_t>(), B[11].as<int>(), d_F, d_F + 1));
        ctx->launches++;
    }
    int hF[2] = {0, 0};
    MPRG_CUDA(ctx, mprg::copy_d2h(ctx, hF, d_F, sizeof(int) * 2, s));
    MPRG_CUDA(ctx, cudaStreamSynchronize(s));
    if (hF[1]) MPRG_FAIL(ctx, MPRG_E_INTERNAL, "hash collision detected while numbering k-mers");
    *n_kmers = hF[0];
    MPRG_CUDA(ctx, B[12].reserve(sizeof(double) * std::max<size_t>((size_t)n * hF[0], 1)));
    MPRG_CUDA(ctx, cudaMemsetAsync(B[12].p, 0, sizeof(double) * (size_t)n * hF[0], s));
    if (k.big && (size_t)hF[0] * sizeof(int) <= 200 * 1024) {
        k.big |= 2;
        MPRG_CUDA(ctx, mprg::copy_h2d(ctx, B[7].p, &k, sizeof(k), s));
        MPRG_CUDA(ctx, launch_kmer_fill_big(s, B[7].p, 0, &k, hF[0], B[9].as<int>(), B[12].as<double>()));
    } else {
        MPRG_CUDA(ctx, launch_kmer_fill(s, B[7].p, 1, P, B[9].as<int>(), d_F, B[12].as<double>()));
    }
    ctx->launches++;
    if (h_counts) {
        if (capacity < (int64_t)n * hF[0]) MPRG_FAIL(ctx, MPRG_E_BAD_ARG, "count matrix capacity too small");
        MPRG_CUDA(ctx, mprg::copy_d2h(ctx, h_counts, B[12].p, sizeof(double) * (size_t)n * hF[0], s));
        MPRG_CUDA(ctx, cudaStreamSynchronize(s));
    }
    return MPRG_OK;
}

extern "C" int mprg_one_ref_like(mprg_ctx *ctx, const mprg_batch *batch, const mprg_task *h_task,
                                 const int32_t *h_rows, const int32_t *h_cluster_of_row,
                                 int32_t n_clusters, int32_t *h_flags) {
    if (!ctx || !batch || !h_task || !h_cluster_of_row || !h_flags || n_clusters < 1) return MPRG_E_BAD_ARG;
    std::vector<ClusterOut> out;
    uint8_t want = 0;
    mprg_task t = *h_task;
    const long long n_row_entries = t.rows_off >= 0 ? (long long)t.rows_off + t.n_rows : 0;
    int rc = cluster_level(ctx, batch, &t, 1, h_rows, n_row_entries, 1, &want, nullptr, false, out);
    if (rc != MPRG_OK) return rc;
    cudaStream_t s = ctx->stream;
    DevBuf *B = ctx->d_c;
    const int R = t.n_rows, w = t.c1 - t.c0;
    for (int r = 0; r < R; ++r)
        if (h_cluster_of_row[r] < 0 || h_cluster_of_row[r] >= n_clusters) MPRG_FAIL(ctx, MPRG_E_BAD_ARG, "cluster index out of range");
    ClusterState c;
    memset(&c, 0, sizeof(c));
    c.K = n_clusters;
    c.n = R;
    c.w = w;
    std::vector<int> mem_off(R + 1), mem_rows(R);
    for (int r = 0; r < R; ++r) {
        mem_off[r] = r;
        mem_rows[r] = r;
    }
    mem_off[R] = R;
    const size_t o_memrows = sizeof(int) * (R + 1), o_assign = o_memrows + sizeof(int) * R;
    const size_t o_flags = o_assign + sizeof(int) * R, o_maj = o_flags + sizeof(int) * n_clusters;
    MPRG_CUDA(ctx, B[13].reserve(sizeof(ClusterState)));
    MPRG_CUDA(ctx, B[14].reserve(o_maj + w + 16));
    uint8_t *b = B[14].as<uint8_t>();
    MPRG_CUDA(ctx, mprg::copy_h2d(ctx, B[13].p, &c, sizeof(c), s));
    MPRG_CUDA(ctx, mprg::copy_h2d(ctx, b, mem_off.data(), sizeof(int) * (R + 1), s));
    MPRG_CUDA(ctx, mprg::copy_h2d(ctx, b + o_memrows, mem_rows.data(), sizeof(int) * R, s));
    MPRG_CUDA(ctx, mprg::copy_h2d(ctx, b + o_assign, h_cluster_of_row, sizeof(int) * R, s));
    MPRG_CUDA(ctx, launch_refcheck(s, B[13].as<ClusterState>(), 1, B[3].as<uint8_t>(), (int *)b, (int *)(b + o_memrows),
                                   (int *)(b + o_assign), b + o_maj, 10, (int *)(b + o_flags)));
    ctx->launches++;
    MPRG_CUDA(ctx, mprg::copy_d2h(ctx, h_flags, b + o_flags, sizeof(int) * n_clusters, s));
    MPRG_CUDA(ctx, cudaStreamSynchronize(s));
    return MPRG_OK;
}

// =================================================================================================
// The whole path: PrgBuilder.__init__ + build_prg for every locus of a batch
// =================================================================================================
namespace mprg {

static const char *iupac_alternatives(char c) {
    switch (c) {
        case 'R': return "GA";
        case 'Y': return "TC";
        case 'K': return "GT";
        case 'M': return "AC";
        case 'S': return "GC";
        case 'W': return "AT";
        default: return nullptr;
    }
}

// SequenceExpander.get_expanded_sequences (seq_utils.py:116-153) on already distinct sequences
static bool expand_sequences(const std::vector<std::string> &seqs, std::vector<std::string> &out) {
    out.clear();
    std::unordered_set<std::string> seen;
    for (const std::string &sq : seqs) {
        if (sq.find('N') != std::string::npos) continue;
        bool plain = true;
        for (char c : sq) plain &= (c == 'A' || c == 'C' || c == 'G' || c == 'T');
        if (plain) {
            if (seen.insert(sq).second) out.push_back(sq);
            continue;
        }
        // itertools.product order: the leftmost position varies slowest
        std::vector<int> amb;
        for (size_t i = 0; i < sq.size(); ++i)
            if (iupac_alternatives(sq[i])) amb.push_back((int)i);
        const size_t total = (size_t)1 << amb.size();
        for (size_t m = 0; m < total; ++m) {
            std::string e = sq;
            for (size_t a = 0; a < amb.size(); ++a) {
                const int bit = (int)((m >> (amb.size() - 1 - a)) & 1u);
                e[amb[a]] = iupac_alternatives(sq[amb[a]])[bit];
            }
            if (seen.insert(e).second) out.push_back(e);
        }
    }
    return !out.empty();
}

}  // namespace mprg


namespace mprg {
static void assemble_prgs_range(const mprg_batch *batch, mprg_result *res, int l_begin, int lo, int hi,
                                const long long *out_off, const int *h_len, const uint8_t *h_out,
                                const long long *prg_bound) {
    std::vector<std::string> raw, expanded;
    struct Frame {
        int node;
        int next_child;
        int site;
    };
    std::vector<Frame> stack;
    for (int l = lo; l < hi; ++l) {
        LocusResult &L = res->loci[l];
        if (L.status != MPRG_LOCUS_OK) continue;
        // upper bound of the string: gapped width of every allele + one marker per allele and node, so
        // that a 200 MB PRG (deep locus) is not grown by doubling
        L.prg.reserve((size_t)prg_bound[l - l_begin] + 12 * L.nodes.size() + 64);
        int site = 5;
        L.preorder.reserve(L.nodes.size());
        stack.clear();
        stack.push_back(Frame{0, 0, 0});
        auto emit_marker = [&](int m) {
            char buf[16];
            int k = 15;
            buf[k] = ' ';
            do {
                buf[--k] = (char)('0' + m % 10);
                m /= 10;
            } while (m);
            buf[--k] = ' ';
            L.prg.append(buf + k, (size_t)(16 - k));
        };
        while (!stack.empty() && L.status == MPRG_LOCUS_OK) {
            Frame &f = stack.back();
            HNode &nd = L.nodes[f.node];
            if (f.next_child == 0) {
                L.preorder.push_back(f.node);
                if (nd.kind == MPRG_NODE_LEAF && !(batch->flags[l] & 4)) {
                    // no RYKMSW anywhere in the locus: the distinct ungapped rows are the alleles
                    if (nd.allele_count == 1) {
                        const long long it_off = out_off[nd.allele_first];
                        L.prg.append(reinterpret_cast<const char *>(h_out + it_off), (size_t)h_len[nd.allele_first]);
                    } else {
                        const int sn = site;
                        site += 2;
                        emit_marker(sn);
                        for (int a = 0; a < nd.allele_count; ++a) {
                            const long long it_off = out_off[nd.allele_first + a];
                            L.prg.append(reinterpret_cast<const char *>(h_out + it_off),
                                         (size_t)h_len[nd.allele_first + a]);
                            emit_marker(a + 1 < nd.allele_count ? sn + 1 : sn);
                        }
                    }
                    stack.pop_back();
                    continue;
                }
                if (nd.kind == MPRG_NODE_LEAF) {
                    raw.clear();
                    for (int a = 0; a < nd.allele_count; ++a) {
                        const long long it_off = out_off[nd.allele_first + a];
                        raw.emplace_back(reinterpret_cast<const char *>(h_out + it_off),
                                         (size_t)h_len[nd.allele_first + a]);
                    }
                    if (!expand_sequences(raw, expanded)) {
                        L.status = MPRG_LOCUS_CURATION_ERROR;
                        break;
                    }
                    if (expanded.size() == 1) {
                        L.prg += expanded[0];
                    } else {
                        const int sn = site;
                        site += 2;
                        emit_marker(sn);
                        for (size_t a = 0; a < expanded.size(); ++a) {
                            L.prg += expanded[a];
                            emit_marker(a + 1 < expanded.size() ? sn + 1 : sn);
                        }
                    }
                    stack.pop_back();
                    continue;
                }
                if (nd.kind == MPRG_NODE_CLUSTER) {
                    f.site = site;
                    site += 2;
                    emit_marker(f.site);
                }
            } else if (nd.kind == MPRG_NODE_CLUSTER) {
                // separator after child (next_child - 1)
                emit_marker(f.next_child < nd.n_children ? f.site + 1 : f.site);
            }
            if (f.next_child < nd.n_children) {
                const int ch = nd.first_child + f.next_child;
                f.next_child++;
                stack.push_back(Frame{ch, 0, 0});
            } else {
                stack.pop_back();
            }
        }
        L.n_sites = (site - 5) / 2;
        if (L.status != MPRG_LOCUS_OK) {
            L.prg.clear();
            L.preorder.clear();
        }
    }
}

void assemble_prgs(const mprg_batch *batch, mprg_result *res, int l_begin, int l_end, const long long *out_off,
                   const int *h_len, const uint8_t *h_out, const long long *prg_bound, int n_threads) {
    const int n = l_end - l_begin;
    n_threads = std::max(1, std::min(n_threads, n / 16));
    if (n_threads <= 1) {
        assemble_prgs_range(batch, res, l_begin, l_begin, l_end, out_off, h_len, h_out, prg_bound);
        return;
    }
    // contiguous ranges of about equal output size
    std::vector<long long> prefix((size_t)n + 1, 0);
    for (int i = 0; i < n; ++i) prefix[i + 1] = prefix[i] + prg_bound[i] + 64;
    std::vector<std::thread> threads;
    int lo = l_begin;
    for (int t = 0; t < n_threads; ++t) {
        int hi = l_end;
        if (t + 1 < n_threads) {
            const long long target = prefix[n] * (t + 1) / n_threads;
            hi = l_begin + (int)(std::lower_bound(prefix.begin(), prefix.end(), target) - prefix.begin());
            hi = std::max(lo, std::min(hi, l_end));
        }
        if (t + 1 == n_threads) {
            assemble_prgs_range(batch, res, l_begin, lo, hi, out_off, h_len, h_out, prg_bound);
        } else {
            threads.emplace_back(assemble_prgs_range, batch, res, l_begin, lo, hi, out_off, h_len, h_out, prg_bound);
        }
        lo = hi;
    }
    for (auto &t : threads) t.join();
}
}  // namespace mprg

// Builds loci [l_begin, l_end) of the batch on one context (one stream, one host thread).
// root_levels (may be null): per locus -1 = the alignment is a locus root (from_msa), >= 0 = it is built
// below an existing node of that nesting level, NodeFactory.build(alignment, builder, parent_node)
// (recursion_tree.py:431-432: no forced MultiIntervalNode, nesting starts at the parent's level).
static int build_range(mprg_ctx *ctx, mprg_batch *batch, int l_begin, int l_end, int32_t max_nesting,
                       int32_t min_match_length, mprg_result *res, bool allow_trace, const int32_t *root_levels) {
    cudaSetDevice(ctx->device);
    cudaStream_t s = ctx->stream;
    struct Pending {
        int locus, node;
    };
    struct Allele {
        int locus, row, c0, c1;
    };
    std::vector<Pending> pending, next;
    std::vector<Allele> alleles;
    auto fail = [&](int code) { return code; };
    for (int l = l_begin; l < l_end; ++l) {
        LocusResult &L = res->loci[l];
        if (batch->flags[l] & 1) {
            L.status = MPRG_LOCUS_CURATION_ERROR;
            continue;
        }
        if ((batch->flags[l] & 2) || batch->n_rows[l] <= 0) {
            L.status = 2;  // N present (loader must replace it first) or empty alignment
            continue;
        }
        HNode root;
        root.parent = -1;
        root.level = (root_levels && root_levels[l] >= 0) ? root_levels[l] : 0;
        L.as_root = !(root_levels && root_levels[l] >= 0);
        root.c0 = 0;
        root.c1 = batch->n_cols[l];
        root.row_off = -1;
        root.n_rows = batch->n_rows[l];
        L.nodes.push_back(root);
        pending.push_back(Pending{l, 0});
    }
    auto first_row = [&](const LocusResult &L, const HNode &nd) {
        return nd.row_off < 0 ? 0 : L.row_pool[nd.row_off];
    };
    auto make_match_leaf = [&](int l, HNode &nd) {
        LocusResult &L = res->loci[l];
        nd.kind = MPRG_NODE_LEAF;
        nd.allele_first = (int)alleles.size();
        nd.allele_count = 1;
        alleles.push_back(Allele{l, first_row(L, nd), nd.c0, nd.c1});
    };

    std::vector<mprg_task> tasks;
    std::vector<int32_t> arena;
    PhaseTrace trace;
    const bool trace_all = getenv("MPRG_TRACE_ALL") != nullptr;
    if (allow_trace || trace_all) g_trace = trace.on ? &trace : nullptr;
    while (!pending.empty()) {
        const int nt = (int)pending.size();
        tasks.resize(nt);
        arena.clear();
        for (int i = 0; i < nt; ++i) {
            const LocusResult &L = res->loci[pending[i].locus];
            const HNode &nd = L.nodes[pending[i].node];
            mprg_task &t = tasks[i];
            t.locus = pending[i].locus;
            t.n_rows = nd.n_rows;
            t.c0 = nd.c0;
            t.c1 = nd.c1;
            if (nd.row_off < 0) {
                t.rows_off = -1;
            } else {
                if (arena.size() + (size_t)nd.n_rows > 0x7fffffffULL) {
                    ctx->err = "row arena overflow";
                    return fail(MPRG_E_BAD_ARG);
                }
                t.rows_off = (int)arena.size();
                arena.insert(arena.end(), L.row_pool.begin() + nd.row_off,
                             L.row_pool.begin() + nd.row_off + nd.n_rows);
            }
        }
        TRACE("build: task table");
        Level lv;
        int rc = level_run(ctx, batch, tasks.data(), nt, arena.data(), (long long)arena.size(),
                           min_match_length, true, lv);
        if (rc != MPRG_OK) return fail(rc);
        TRACE("build: level_run (host+launch)");
        cudaError_t e;
        if ((e = ctx->h_d.reserve(sizeof(DInterval) * (size_t)lv.total_iv + sizeof(int) * (nt + 1))) != cudaSuccess) {
            ctx->err = std::string("pinned alloc: ") + cudaGetErrorString(e);
            return fail(MPRG_E_CUDA);
        }
        DInterval *iv = ctx->h_d.as<DInterval>();
        int *cnt = reinterpret_cast<int *>(iv + lv.total_iv);
        if ((e = mprg::copy_d2h(ctx, iv, ctx->d_iv.p, sizeof(DInterval) * (size_t)lv.total_iv, s)) != cudaSuccess ||
            (e = mprg::copy_d2h(ctx, cnt, ctx->d_ivcnt.p, sizeof(int) * (nt + 1), s)) != cudaSuccess ||
            (e = cudaStreamSynchronize(s)) != cudaSuccess) {
            ctx->err = std::string("level D2H: ") + cudaGetErrorString(e);
            return fail(MPRG_E_CUDA);
        }
        TRACE("build: level sync+D2H");
        account_scan(ctx, lv);
        if (cnt[nt]) {
            ctx->err = "Failed interval partitioning";
            return fail(MPRG_E_PARTITION);
        }
        next.clear();
        std::vector<int> cluster_idx;  // indices into pending/tasks
        for (int i = 0; i < nt; ++i) {
            const int l = pending[i].locus;
            LocusResult &L = res->loci[l];
            const int ni = pending[i].node;
            const DInterval *ivs = iv + lv.tasks[i].iv_off;
            const int c = cnt[i];
            const bool is_root = L.nodes[ni].parent < 0 && L.as_root;
            if (c == 1 && ivs[0].type != MPRG_IV_NONMATCH) {
                make_match_leaf(l, L.nodes[ni]);
            } else if (c > 1 || is_root) {
                if (c == 0) {
                    L.status = 2;  // zero-column root: the reference trips an assertion here
                    continue;
                }
                L.nodes[ni].kind = MPRG_NODE_INTERVAL;
                for (int k = 0; k < c; ++k) {
                    HNode ch;
                    ch.parent = ni;
                    ch.level = L.nodes[ni].level;
                    ch.c0 = L.nodes[ni].c0 + ivs[k].start;
                    ch.c1 = L.nodes[ni].c0 + ivs[k].stop + 1;
                    ch.row_off = L.nodes[ni].row_off;
                    ch.n_rows = L.nodes[ni].n_rows;
                    const int ci = (int)L.nodes.size();
                    L.nodes.push_back(ch);
                    if (L.nodes[ni].n_children++ == 0) L.nodes[ni].first_child = ci;
                    // a pure match interval re-partitions to itself: leaf without another scan
                    if (ivs[k].type == MPRG_IV_MATCH) make_match_leaf(l, L.nodes[ci]);
                    else next.push_back(Pending{l, ci});
                }
            } else {
                cluster_idx.push_back(i);
            }
        }
        TRACE("build: host nodes");
        if (!cluster_idx.empty()) {
            const int nc = (int)cluster_idx.size();
            std::vector<mprg_task> ctasks(nc);
            std::vector<uint8_t> want(nc);
            for (int q = 0; q < nc; ++q) {
                const int i = cluster_idx[q];
                ctasks[q] = tasks[i];
                const HNode &nd = res->loci[pending[i].locus].nodes[pending[i].node];
                want[q] = (nd.level + 1 < max_nesting) ? 1 : 0;
            }
            std::vector<ClusterOut> cout_;
            rc = cluster_level(ctx, batch, ctasks.data(), nc, arena.data(), (long long)arena.size(),
                               min_match_length, want.data(), nullptr, true, cout_, true);
            if (rc != MPRG_OK) return fail(rc);
            TRACE("cl: finalize");
            std::vector<int> index_of_group, cl_of_row, cl_count, cl_start;
            for (int q = 0; q < nc; ++q) {
                const int i = cluster_idx[q];
                const int l = pending[i].locus, ni = pending[i].node;
                LocusResult &L = res->loci[l];
                const ClusterOut &o = cout_[q];
                const bool has_issues = o.n_ungapped <= 2 || o.n_ungapped < o.n_gapped;
                const bool further = want[q] && !has_issues && o.clustered && !o.no_clustering;
                auto row_id = [&](int pos) {
                    const HNode &nd = L.nodes[ni];
                    return nd.row_off < 0 ? pos : L.row_pool[nd.row_off + pos];
                };
                if (further) {
                    // counting sort of the rows by cluster: every child keeps the input row order
                    const int n_cl = cluster_of_rows(o, min_match_length, index_of_group, cl_of_row);
                    const int R = (int)cl_of_row.size();
                    cl_count.assign(n_cl, 0);
                    for (int r = 0; r < R; ++r) cl_count[cl_of_row[r]]++;
                    cl_start.resize(n_cl);
                    const long long pool0 = (long long)L.row_pool.size();
                    long long at = pool0;
                    for (int c = 0; c < n_cl; ++c) {
                        cl_start[c] = (int)(at - pool0);
                        at += cl_count[c];
                    }
                    L.row_pool.resize((size_t)at);
                    {
                        const HNode &nd = L.nodes[ni];
                        int *dst = L.row_pool.data() + pool0;
                        if (nd.row_off < 0) {
                            for (int r = 0; r < R; ++r) dst[cl_start[cl_of_row[r]]++] = r;
                        } else {
                            const int *src = L.row_pool.data() + nd.row_off;
                            for (int r = 0; r < R; ++r) dst[cl_start[cl_of_row[r]]++] = src[r];
                        }
                    }
                    L.nodes[ni].kind = MPRG_NODE_CLUSTER;
                    L.nodes[ni].level += 1;
                    long long off = pool0;
                    for (int c = 0; c < n_cl; ++c) {
                        if (cl_count[c] == 0) continue;
                        HNode ch;
                        ch.parent = ni;
                        ch.level = L.nodes[ni].level;
                        ch.c0 = L.nodes[ni].c0;
                        ch.c1 = L.nodes[ni].c1;
                        ch.row_off = off;
                        ch.n_rows = cl_count[c];
                        off += cl_count[c];
                        const int ci = (int)L.nodes.size();
                        L.nodes.push_back(ch);
                        if (L.nodes[ni].n_children++ == 0) L.nodes[ni].first_child = ci;
                        next.push_back(Pending{l, ci});
                    }
                } else {
                    // forced leaf: one allele per distinct ungapped row, first-seen order
                    HNode &nd = L.nodes[ni];
                    nd.kind = MPRG_NODE_LEAF;
                    nd.allele_first = (int)alleles.size();
                    for (int g = 0; g < o.n_ungapped; ++g)
                        alleles.push_back(Allele{l, row_id(o.leaders[g]), nd.c0, nd.c1});
                    nd.allele_count = o.n_ungapped;
                }
            }
        }
        TRACE("build: host cluster nodes");
        pending.swap(next);
    }

    // ---- leaf alleles: ungapped strings cut on the device ----
    const int na = (int)alleles.size();
    std::vector<ExtractItem> items(na);
    long long out_total = 0;
    std::vector<long long> prg_bound((size_t)std::max(l_end - l_begin, 1), 0);
    for (int a = 0; a < na; ++a) {
        const Allele &al = alleles[a];
        items[a] = ExtractItem{batch->base[al.locus], batch->stride[al.locus], al.row, al.c0, al.c1, out_total};
        out_total += al.c1 - al.c0;
        prg_bound[al.locus - l_begin] += (al.c1 - al.c0) + 12;
    }
    {
        cudaError_t e1 = ctx->h_c.reserve((size_t)std::max<long long>(out_total, 1) + sizeof(int) * (size_t)std::max(na, 1) + 16);
        if (e1 != cudaSuccess) {
            ctx->err = std::string("pinned alloc: ") + cudaGetErrorString(e1);
            return fail(MPRG_E_CUDA);
        }
    }
    int *h_len = ctx->h_c.as<int>();
    uint8_t *h_out = reinterpret_cast<uint8_t *>(h_len + std::max(na, 1));
    if (na > 0) {
        DevBuf *B = ctx->d_c;
        cudaError_t e;
        if ((e = B[0].reserve(sizeof(ExtractItem) * na)) != cudaSuccess ||
            (e = B[3].reserve((size_t)std::max<long long>(out_total, 1))) != cudaSuccess ||
            (e = B[5].reserve(sizeof(int) * na)) != cudaSuccess ||
            (e = mprg::copy_h2d(ctx, B[0].p, items.data(), sizeof(ExtractItem) * na, s)) != cudaSuccess) {
            ctx->err = std::string("extract setup: ") + cudaGetErrorString(e);
            return fail(MPRG_E_CUDA);
        }
        extract_kernel<<<(na + 3) / 4, 128, 0, s>>>(batch->d_packed, B[0].as<ExtractItem>(), na, B[3].as<uint8_t>(),
                                                    B[5].as<int>());
        ctx->launches++;
        if ((e = cudaGetLastError()) != cudaSuccess ||
            (e = mprg::copy_d2h(ctx, h_out, B[3].p, (size_t)out_total, s)) != cudaSuccess ||
            (e = mprg::copy_d2h(ctx, h_len, B[5].p, sizeof(int) * na, s)) != cudaSuccess ||
            (e = cudaStreamSynchronize(s)) != cudaSuccess) {
            ctx->err = std::string("extract: ") + cudaGetErrorString(e);
            return fail(MPRG_E_CUDA);
        }
    }

    TRACE("build: allele extraction");
    {
        std::vector<long long> off(std::max(na, 1));
        for (int a = 0; a < na; ++a) off[a] = items[a].out_off;
        assemble_prgs(batch, res, l_begin, l_end, off.data(), h_len, h_out, prg_bound.data(), 1);
    }
    TRACE("build: prg strings");
    if (allow_trace || trace_all) {
        trace.report("mprg_build");
        g_trace = nullptr;
    }
    return MPRG_OK;
}

extern "C" int mprg_set_wait_mode(mprg_ctx *ctx, int32_t mode) {
    if (!ctx || mode < 0 || mode > 2) return MPRG_E_BAD_ARG;
    ctx->wait_mode = mode;
    for (mprg_ctx *w : ctx->workers) w->wait_mode = mode;
    return MPRG_OK;
}

extern "C" int mprg_set_workers(mprg_ctx *ctx, int32_t n_workers) {
    if (!ctx || n_workers < 1 || n_workers > 64) return MPRG_E_BAD_ARG;
    ctx->n_workers = n_workers;
    return MPRG_OK;
}

// Loci are independent, so a batch is cut into contiguous, cost-balanced ranges that are built
// concurrently: one host thread + one stream + one scratch set per range, all reading the same
// packed batch.  This overlaps the host bookkeeping of one range with the kernels of the others.
// With h_ascii != nullptr every range is first uploaded and packed by its own worker, so the
// host-to-device copies of some ranges overlap the kernels of the others (mprg_build_ascii).
static int build_ranges(mprg_ctx *ctx, mprg_batch *batch, int32_t max_nesting, int32_t min_match_length,
                        const uint8_t *h_ascii, const int64_t *h_offsets, mprg_result **out_res,
                        const int32_t *root_levels = nullptr, const int32_t *h_packed_flags = nullptr,
                        bool host_is_packed = false) {
    *out_res = nullptr;
    cudaSetDevice(ctx->device);
    const int n_loci = batch->n_loci;
    mprg_result *res = new mprg_result();
    res->loci.resize(n_loci);
    int rc = ensure_rand(ctx);
    if (rc != MPRG_OK) {
        delete res;
        return rc;
    }
    const auto t_call = std::chrono::steady_clock::now();
    const bool trace_all = getenv("MPRG_TRACE") && getenv("MPRG_TRACE_ALL");
    auto since_call = [&]() {
        return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_call).count();
    };
    auto run = [&](mprg_ctx *c, int l0, int l1, bool trace) {
        if (h_ascii) {
            const double t0 = since_call();
            const int r = host_is_packed ? batch_upload_range_packed(c, batch, h_ascii, h_offsets, h_packed_flags, l0, l1)
                                         : batch_upload_range(c, batch, h_ascii, h_offsets, l0, l1);
            if (r != MPRG_OK) return r;
            if (trace_all) fprintf(stderr, "[mprg trace] range %d..%d: upload %.2f -> %.2f ms after the call\n", l0, l1, t0, since_call());
        }
        // the device-resident level loop is the product path; MPRG_HOST_LOOP=1 keeps the host-driven one
        // (the checked alternative: both must produce identical results)
        static const bool host_loop = getenv("MPRG_HOST_LOOP") != nullptr;
        const int rc_b = host_loop
                             ? build_range(c, batch, l0, l1, max_nesting, min_match_length, res, trace, root_levels)
                             : build_range_dev(c, batch, l0, l1, max_nesting, min_match_length, res, trace, root_levels);
        if (trace_all) fprintf(stderr, "[mprg trace] range %d..%d: built %.2f ms after the call\n", l0, l1, since_call());
        return rc_b;
    };
    int W = std::max(1, std::min(ctx->n_workers, n_loci / 8));
    {
        // The device-resident loop needs no host threads to hide bookkeeping: one range per build when the
        // batch is resident, two from host buffers (the upload of one overlaps the kernels of the other).
        // MPRG_DEV_RANGES overrides.
        static const bool host_loop = getenv("MPRG_HOST_LOOP") != nullptr;
        static const int dev_ranges = []() {
            const char *e = getenv("MPRG_DEV_RANGES");
            return e ? std::max(1, atoi(e)) : 0;
        }();
        if (!host_loop) W = std::min(W, dev_ranges ? dev_ranges : (h_ascii ? 2 : 1));
    }
    if (W <= 1) {
        rc = run(ctx, 0, n_loci, true);
        if (rc != MPRG_OK) {
            delete res;
            return rc;
        }
        *out_res = res;
        return MPRG_OK;
    }
    while ((int)ctx->workers.size() < W - 1) {
        mprg_ctx *w = nullptr;
        rc = mprg_create(ctx->device, &w);
        if (rc != MPRG_OK) {
            delete res;
            MPRG_FAIL(ctx, rc, "could not create a worker context");
        }
        w->wait_mode = ctx->wait_mode;
        ctx->workers.push_back(w);
    }
    // contiguous ranges of roughly equal rows x cols, dealt round-robin; from host ASCII there are
    // two ranges per worker, so that copies (one range at a time) and kernels stay overlapped
    static const int ranges_per_worker = []() {
        const char *e = getenv("MPRG_RANGES_PER_WORKER");
        return e ? std::max(1, atoi(e)) : 0;
    }();
    static const bool dynamic = getenv("MPRG_DYNAMIC_RANGES") != nullptr;
    const int rpw = ranges_per_worker ? ranges_per_worker : 1;  // sweeps: profiles/r1_ranges_sweep.txt
    const int R = std::max(W, std::min(rpw * W, n_loci / 8));
    std::vector<double> prefix(n_loci + 1, 0.0);
    for (int l = 0; l < n_loci; ++l) prefix[l + 1] = prefix[l] + (double)batch->n_rows[l] * batch->n_cols[l] + 1.0;
    std::vector<int> cut(R + 1, n_loci);
    cut[0] = 0;
    for (int r = 1; r < R; ++r) {
        const double target = prefix[n_loci] * r / R;
        cut[r] = (int)(std::lower_bound(prefix.begin(), prefix.end(), target) - prefix.begin());
        cut[r] = std::max(cut[r], cut[r - 1]);
    }
    std::vector<int> rcs(W, MPRG_OK);
    // range r always goes to worker r % W: the same worker sees the same loci on every call, so its
    // scratch buffers stop growing after the first call (a reallocation synchronises the device)
    std::atomic<int> next_range{0};
    auto work = [&](mprg_ctx *c, int w) {
        // static: range r goes to worker r % W; dynamic: the next range goes to whichever worker is free
        for (int r = dynamic ? next_range.fetch_add(1) : w; r < R; r = dynamic ? next_range.fetch_add(1) : r + W) {
            const int rc_r = run(c, cut[r], cut[r + 1], false);
            if (rc_r != MPRG_OK) {
                rcs[w] = rc_r;
                break;
            }
        }
    };
    std::vector<std::thread> threads;
    for (int w = 1; w < W; ++w) {
        mprg_ctx *wc = ctx->workers[w - 1];
        threads.emplace_back([&, w, wc]() { work(wc, w); });
    }
    work(ctx, 0);
    for (auto &t : threads) t.join();
    for (int w = 1; w < W; ++w) {
        mprg_ctx *wc = ctx->workers[w - 1];
        ctx->launches += wc->launches;
        ctx->h2d_bytes += wc->h2d_bytes;
        ctx->d2h_bytes += wc->d2h_bytes;
        ctx->scan_ms += wc->scan_ms;
        ctx->scan_bytes += wc->scan_bytes;
        ctx->scan_launches += wc->scan_launches;
        ctx->scan_log_bytes.insert(ctx->scan_log_bytes.end(), wc->scan_log_bytes.begin(), wc->scan_log_bytes.end());
        ctx->scan_log_ms.insert(ctx->scan_log_ms.end(), wc->scan_log_ms.begin(), wc->scan_log_ms.end());
        wc->launches = wc->h2d_bytes = wc->d2h_bytes = 0;
        wc->scan_ms = wc->scan_bytes = 0;
        wc->scan_launches = 0;
        wc->scan_log_bytes.clear();
        wc->scan_log_ms.clear();
        if (rcs[w] != MPRG_OK && rcs[0] == MPRG_OK) {
            rcs[0] = rcs[w];
            ctx->err = wc->err;
        }
    }
    if (rcs[0] != MPRG_OK) {
        delete res;
        return rcs[0];
    }
    *out_res = res;
    return MPRG_OK;
}

extern "C" int mprg_build(mprg_ctx *ctx, mprg_batch *batch, int32_t max_nesting,
                          int32_t min_match_length, mprg_result **out_res) {
    if (!ctx || !batch || !out_res || min_match_length < 1) return MPRG_E_BAD_ARG;
    return build_ranges(ctx, batch, max_nesting, min_match_length, nullptr, nullptr, out_res);
}

extern "C" int mprg_build_sub(mprg_ctx *ctx, mprg_batch *batch, int32_t max_nesting, int32_t min_match_length,
                              const int32_t *h_parent_level, mprg_result **out_res) {
    if (!ctx || !batch || !out_res || min_match_length < 1) return MPRG_E_BAD_ARG;
    return build_ranges(ctx, batch, max_nesting, min_match_length, nullptr, nullptr, out_res, h_parent_level);
}

extern "C" int mprg_build_ascii(mprg_ctx *ctx, const uint8_t *h_ascii, const int64_t *h_offsets,
                                const int32_t *n_rows, const int32_t *n_cols, int32_t n_loci,
                                int32_t max_nesting, int32_t min_match_length, mprg_batch **out_batch,
                                mprg_result **out_res) {
    if (!ctx || !out_batch || !out_res || min_match_length < 1 || n_loci < 0 ||
        (n_loci > 0 && (!h_ascii || !h_offsets || !n_rows || !n_cols)))
        return MPRG_E_BAD_ARG;
    *out_batch = nullptr;
    *out_res = nullptr;
    mprg_batch *b = nullptr;
    int rc = batch_prepare(ctx, n_rows, n_cols, n_loci, &b);
    if (rc != MPRG_OK) return rc;
    rc = build_ranges(ctx, b, max_nesting, min_match_length, h_ascii, h_offsets, out_res);
    if (rc != MPRG_OK) {
        mprg_batch_free(ctx, b);
        return rc;
    }
    *out_batch = b;
    return MPRG_OK;
}

extern "C" int mprg_build_packed(mprg_ctx *ctx, const uint8_t *h_packed, const int64_t *h_offsets,
                                 const int32_t *n_rows, const int32_t *n_cols, const int32_t *h_flags, int32_t n_loci,
                                 int32_t max_nesting, int32_t min_match_length, mprg_batch **out_batch,
                                 mprg_result **out_res) {
    if (!ctx || !out_batch || !out_res || min_match_length < 1 || n_loci < 0 ||
        (n_loci > 0 && (!h_packed || !h_offsets || !n_rows || !n_cols)))
        return MPRG_E_BAD_ARG;
    *out_batch = nullptr;
    *out_res = nullptr;
    mprg_batch *b = nullptr;
    int rc = batch_prepare(ctx, n_rows, n_cols, n_loci, &b);
    if (rc != MPRG_OK) return rc;
    rc = build_ranges(ctx, b, max_nesting, min_match_length, h_packed, h_offsets, out_res, nullptr, h_flags, true);
    if (rc != MPRG_OK) {
        mprg_batch_free(ctx, b);
        return rc;
    }
    *out_batch = b;
    return MPRG_OK;
}

extern "C" int mprg_result_from_prgs(const char *const *prgs, const int64_t *lengths, int32_t n, mprg_result **out) {
    if (!out || n < 0 || (n > 0 && (!prgs || !lengths))) return MPRG_E_BAD_ARG;
    mprg_result *res = new mprg_result();
    res->loci.resize(n);
    for (int i = 0; i < n; ++i) res->loci[i].prg.assign(prgs[i] ? prgs[i] : "", (size_t)std::max<int64_t>(lengths[i], 0));
    *out = res;
    return MPRG_OK;
}

// ---- pinned blocks recycled across results -----------------------------------------------------------
namespace mprg {
static std::mutex g_pin_mutex;
static std::vector<PinnedBlock> g_pin_free;

PinnedBlock pinned_acquire(size_t bytes) {
    bytes = std::max<size_t>(bytes, 64);
    {
        std::lock_guard<std::mutex> lock(g_pin_mutex);
        int best = -1;
        for (int i = 0; i < (int)g_pin_free.size(); ++i)
            if (g_pin_free[i].cap >= bytes && (best < 0 || g_pin_free[i].cap < g_pin_free[best].cap)) best = i;
        if (best >= 0 && g_pin_free[best].cap <= 4 * bytes + ((size_t)1 << 20)) {
            PinnedBlock b = g_pin_free[best];
            g_pin_free.erase(g_pin_free.begin() + best);
            return b;
        }
    }
    PinnedBlock b;
    const size_t want = bytes + bytes / 4 + 4096;
    if (cudaMallocHost(&b.p, want) == cudaSuccess) b.cap = want;
    else b.p = nullptr;
    return b;
}

void pinned_release(PinnedBlock b) {
    if (!b.p) return;
    {
        std::lock_guard<std::mutex> lock(g_pin_mutex);
        size_t held = 0;
        for (const PinnedBlock &f : g_pin_free) held += f.cap;
        if (g_pin_free.size() < 24 && held + b.cap <= ((size_t)4 << 30)) {
            g_pin_free.push_back(b);
            return;
        }
    }
    cudaFreeHost(b.p);
}

void ensure_tables(mprg_result *res, int l) {
    LocusResult &L = res->loci[l];
    if (L.tables_ready) return;
    std::lock_guard<std::mutex> lock(res->lazy_mutex);
    if (L.tables_ready) return;
    const RawTree &raw = res->raw[L.raw_index];
    const DNode *nodes = static_cast<const DNode *>(raw.nodes.p);
    const int *pool = static_cast<const int *>(raw.pool.p);
    std::vector<int> order;  // raw node indices in local order: children of a node contiguous
    order.push_back(L.raw_root);
    L.nodes.clear();
    L.row_pool.clear();
    for (size_t k = 0; k < order.size(); ++k) {
        const DNode &g = nodes[order[k]];
        HNode h;
        h.kind = g.kind;
        h.parent = -1;
        h.level = g.level;
        h.c0 = g.c0;
        h.c1 = g.c1;
        h.row_off = -1;
        h.n_rows = g.n_rows;
        h.first_child = g.n_children ? (int)order.size() : -1;
        h.n_children = g.n_children;
        h.allele_first = g.allele_first;
        h.allele_count = g.allele_count;
        L.nodes.push_back(h);
        for (int c = 0; c < g.n_children; ++c) order.push_back(g.first_child + c);
    }
    // parents and row subsets: interval children share their parent's rows, cluster children own theirs
    for (size_t k = 0; k < order.size(); ++k) {
        HNode &h = L.nodes[k];
        const DNode &g = nodes[order[k]];
        for (int c = 0; c < h.n_children; ++c) {
            HNode &ch = L.nodes[h.first_child + c];
            const DNode &gc = nodes[g.first_child + c];
            ch.parent = (int)k;
            if (gc.row_off < 0) {
                ch.row_off = -1;
            } else if (gc.row_off == g.row_off) {
                ch.row_off = h.row_off;
            } else {
                ch.row_off = (long long)L.row_pool.size();
                L.row_pool.insert(L.row_pool.end(), pool + gc.row_off, pool + gc.row_off + gc.n_rows);
            }
        }
    }
    // pre-order == node_id order
    L.preorder.clear();
    L.preorder.reserve(L.nodes.size());
    std::vector<std::pair<int, int>> stack;
    stack.emplace_back(0, 0);
    L.preorder.push_back(0);
    while (!stack.empty()) {
        auto &f = stack.back();
        const HNode &nd = L.nodes[f.first];
        if (f.second < nd.n_children) {
            const int ch = nd.first_child + f.second++;
            L.preorder.push_back(ch);
            stack.emplace_back(ch, 0);
        } else {
            stack.pop_back();
        }
    }
    L.tables_ready = true;
}
}  // namespace mprg

mprg_result::~mprg_result() {
    for (mprg::RawTree &r : raw) {
        mprg::pinned_release(r.nodes);
        mprg::pinned_release(r.pool);
    }
    for (mprg::PinnedBlock &b : blobs) mprg::pinned_release(b);
}

extern "C" void mprg_result_free(mprg_result *res) { delete res; }
extern "C" int32_t mprg_result_n_loci(const mprg_result *res) { return res ? (int32_t)res->loci.size() : 0; }
extern "C" int32_t mprg_result_status(const mprg_result *res, int32_t l) {
    return (res && l >= 0 && l < (int)res->loci.size()) ? res->loci[l].status : -1;
}
extern "C" int mprg_result_statuses(const mprg_result *res, int32_t *h_status, int64_t *h_prg_length) {
    if (!res) return MPRG_E_BAD_ARG;
    for (size_t l = 0; l < res->loci.size(); ++l) {
        if (h_status) h_status[l] = res->loci[l].status;
        if (h_prg_length)
            h_prg_length[l] = res->loci[l].prg_data ? (int64_t)res->loci[l].prg_size : (int64_t)res->loci[l].prg.size();
    }
    return MPRG_OK;
}
extern "C" const char *mprg_result_prg(const mprg_result *res, int32_t l, int64_t *length) {
    if (!res || l < 0 || l >= (int)res->loci.size()) return nullptr;
    const LocusResult &L = res->loci[l];
    if (L.prg_data) {
        if (length) *length = (int64_t)L.prg_size;
        return L.prg_data;
    }
    if (length) *length = (int64_t)L.prg.size();
    return L.prg.data();
}
extern "C" int32_t mprg_result_n_nodes(const mprg_result *res, int32_t l) {
    if (!res || l < 0 || l >= (int)res->loci.size()) return 0;
    const LocusResult &L = res->loci[l];
    return L.n_nodes >= 0 ? L.n_nodes : (int32_t)L.preorder.size();
}
extern "C" int32_t mprg_result_n_sites(const mprg_result *res, int32_t l) {
    return (res && l >= 0 && l < (int)res->loci.size()) ? res->loci[l].n_sites : 0;
}
extern "C" int mprg_result_nodes(const mprg_result *res, int32_t l, int32_t *kind, int32_t *parent,
                                 int32_t *nesting_level, int32_t *c0, int32_t *c1, int32_t *n_rows,
                                 int64_t *row_off, int32_t *n_children) {
    if (!res || l < 0 || l >= (int)res->loci.size()) return MPRG_E_BAD_ARG;
    ensure_tables(const_cast<mprg_result *>(res), l);
    const LocusResult &L = res->loci[l];
    std::vector<int> new_id(L.nodes.size(), -1);
    for (size_t i = 0; i < L.preorder.size(); ++i) new_id[L.preorder[i]] = (int)i;
    for (size_t i = 0; i < L.preorder.size(); ++i) {
        const HNode &nd = L.nodes[L.preorder[i]];
        if (kind) kind[i] = nd.kind;
        if (parent) parent[i] = nd.parent < 0 ? -1 : new_id[nd.parent];
        if (nesting_level) nesting_level[i] = nd.level;
        if (c0) c0[i] = nd.c0;
        if (c1) c1[i] = nd.c1;
        if (n_rows) n_rows[i] = nd.n_rows;
        if (row_off) row_off[i] = nd.row_off;
        if (n_children) n_children[i] = nd.n_children;
    }
    return MPRG_OK;
}
extern "C" int64_t mprg_result_row_pool_size(const mprg_result *res, int32_t l) {
    if (!res || l < 0 || l >= (int)res->loci.size()) return 0;
    ensure_tables(const_cast<mprg_result *>(res), l);
    return (int64_t)res->loci[l].row_pool.size();
}
extern "C" int mprg_result_row_pool(const mprg_result *res, int32_t l, int32_t *h_rows) {
    if (!res || l < 0 || l >= (int)res->loci.size() || !h_rows) return MPRG_E_BAD_ARG;
    ensure_tables(const_cast<mprg_result *>(res), l);
    const LocusResult &L = res->loci[l];
    if (!L.row_pool.empty()) memcpy(h_rows, L.row_pool.data(), sizeof(int) * L.row_pool.size());
    return MPRG_OK;
}
