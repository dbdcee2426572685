// Shared between the two level loops (engine.cu: host-driven, kept for deep clustering levels and as the
// checked alternative; engine_dev.cu: device-resident): result containers, allele extraction, PRG assembly.
#pragma once
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <mutex>
#include <string>
#include <vector>

#include "common.cuh"
#include "kernels.cuh"

namespace mprg {

// ---- optional wall-clock phase trace (MPRG_TRACE=1) ------------------------------------------------
struct PhaseTrace {
    bool on;
    std::vector<std::pair<std::string, double>> acc;
    std::chrono::steady_clock::time_point t;
    PhaseTrace() : on(getenv("MPRG_TRACE") != nullptr), t(std::chrono::steady_clock::now()) {}
    void mark(const char *name) {
        if (!on) return;
        auto now = std::chrono::steady_clock::now();
        const double ms = std::chrono::duration<double, std::milli>(now - t).count();
        t = now;
        for (auto &p : acc)
            if (p.first == name) {
                p.second += ms;
                return;
            }
        acc.emplace_back(name, ms);
    }
    void report(const char *title) {
        if (!on) return;
        double tot = 0;
        for (auto &p : acc) tot += p.second;
        fprintf(stderr, "[mprg trace] %s total %.2f ms\n", title, tot);
        for (auto &p : acc) fprintf(stderr, "    %-28s %8.2f ms\n", p.first.c_str(), p.second);
    }
};
extern thread_local PhaseTrace *g_trace;
#define TRACE(name) do { if (g_trace) g_trace->mark(name); } while (0)

struct HNode {
    int kind = -1;
    int parent = -1;
    int level = 0;
    int c0 = 0, c1 = 0;
    long long row_off = -1;  // into the locus row pool, -1 = all rows
    int n_rows = 0;
    // the children of a node are created one after the other: nodes [first_child, first_child + n_children)
    int first_child = -1, n_children = 0;
    int allele_first = -1, allele_count = 0;  // extract items of a leaf
};

struct LocusResult {
    int status = MPRG_LOCUS_OK;
    bool as_root = true;            // false: built below an existing node (mprg_build_sub)
    std::vector<HNode> nodes;       // creation (level) order; node 0 is the root
    std::vector<int> row_pool;
    std::vector<int> preorder;      // node indices in pre-order == node_id order
    std::string prg;
    int n_sites = 0;
    // device-assembled results: the PRG lies in a blob of the result, the node table is made from the raw
    // device tree (mprg_result::raw[raw_index]) the first time it is asked for
    const char *prg_data = nullptr;
    long long prg_size = 0;
    int n_nodes = -1;
    int raw_index = -1, raw_root = -1;
    bool tables_ready = true;
};

// node and locus records of the device-resident loop (engine_dev.cu)
struct DNode {
    int locus, parent, level, kind;
    int c0, c1;
    long long row_off;  // into the device row pool, -1 = all rows of the locus
    int n_rows, first_child, n_children, allele_first, allele_count, pad;
};

// pinned host memory recycled across results (cudaMallocHost / cudaFreeHost cost milliseconds)
struct PinnedBlock {
    void *p = nullptr;
    size_t cap = 0;
};
PinnedBlock pinned_acquire(size_t bytes);
void pinned_release(PinnedBlock b);

// what one range of the device-resident loop brought back: node table and row pool as the device made them
struct RawTree {
    int l_begin = 0, l_end = 0;
    PinnedBlock nodes, pool;
    int n_nodes = 0;
    long long pool_size = 0;
};

// one allele of a leaf: ungapped symbols of (row, [c0, c1)) of a locus, written at out_off
struct ExtractItem {
    long long base;
    int stride, row, c0, c1;
    long long out_off;
};
cudaError_t launch_extract(cudaStream_t s, const uint8_t *packed, const ExtractItem *items, int n_items, uint8_t *out,
                           int *out_len);

}  // namespace mprg

struct mprg_result {
    std::vector<mprg::LocusResult> loci;
    std::vector<mprg::RawTree> raw;          // one per range of the device-resident loop
    std::vector<mprg::PinnedBlock> blobs;    // PRG strings assembled on the device
    std::mutex lazy_mutex, raw_mutex;
    ~mprg_result();
};

namespace mprg {
// Pre-order numbering and PRG strings (recursion_tree.py:194-300, prg_builder.py:100-110) of loci
// [l_begin, l_end) from their node tables; allele a of the call lies at h_out + out_off[a], h_len[a] bytes.
// n_threads > 1 spreads the loci over host threads.
void assemble_prgs(const mprg_batch *batch, mprg_result *res, int l_begin, int l_end, const long long *out_off,
                   const int *h_len, const uint8_t *h_out, const long long *prg_bound, int n_threads);
// node table, row pool and pre-order of one locus from the raw device tree (first use)
void ensure_tables(mprg_result *res, int l);
int ensure_rand(mprg_ctx *ctx);
// one clustering problem as the host-driven clustering loop takes it: the n distinct long sequences of
// cluster task `task` (w columns, R rows, n_groups distinct ungapped sequences), P k-mer positions
struct HostProblem {
    int task, n;
    long long P;
    int w, R, n_groups;
    long long g_off, row_off;  // unpacked rows / per-row arrays of the task
    // the task in the packed arena (whole-grid one-reference-like check reads the 4-bit rows directly)
    long long base = 0;
    int stride = 0, rows_off = -1, c0 = 0;
    int alpha_flags = 0;  // alphabet flags of its locus (mprg_batch_flags)
};
struct ProblemRun {
    ClusterState *d_states = nullptr;  // final loop states, in problem order
    int *d_assign = nullptr;           // cluster of every distinct long sequence, at states[q].assign_off
    std::vector<ClusterState> st;
    std::vector<int> h_assign;
};
int run_problems_host(mprg_ctx *ctx, cudaStream_t s, const std::vector<HostProblem> &hp, const std::vector<int> &seq_rows,
                      int kmer_size, const uint8_t *d_G, const int *d_group, const int *d_leadlen, int *d_leader_u,
                      int *d_err, bool fetch, ProblemRun &run, const uint8_t *d_packed, const int *d_rows_arena);
// engine_dev.cu: the device-resident level loop over loci [l_begin, l_end)
int build_range_dev(mprg_ctx *ctx, mprg_batch *batch, int l_begin, int l_end, int32_t max_nesting,
                    int32_t min_match_length, mprg_result *res, bool allow_trace, const int32_t *root_levels);
}  // namespace mprg
