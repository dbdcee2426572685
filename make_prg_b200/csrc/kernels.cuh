// Launchers of the device kernels (definitions in scan.cu, partition.cu, cluster.cu, kmeans.cu).
#pragma once
#include "common.cuh"
#include "km_layout.cuh"

namespace mprg {

cudaError_t launch_scan(cudaStream_t stream, bool has_n, const uint8_t *packed, const ScanUnit *d_units,
                        int n_units, const int *d_rows, uint32_t *colOR, uint32_t *colNOR, unsigned *colB);
cudaError_t launch_classify(cudaStream_t stream, const DTask *d_tasks, int n_tasks,
                            const uint32_t *colOR, const uint32_t *colNOR, const unsigned *colB,
                            uint8_t *cls, int *reach, uint32_t *starbits);
cudaError_t launch_partition(cudaStream_t stream, const DTask *d_tasks, int n_tasks,
                             const uint32_t *starbits, const int *reach, int mml, DInterval *intervals,
                             int *iv_count, int *err);
cudaError_t launch_partition_consensus(cudaStream_t stream, const uint8_t *cons, const int *reach, int n,
                                       int mml, uint32_t *starbits, DInterval *intervals, int *iv_count,
                                       int *err);
cudaError_t launch_demote(cudaStream_t stream, const uint8_t *packed, const DTask *d_tasks, int n_tasks,
                          const int *d_rows, DInterval *intervals, const int *iv_count);

// batch.cu: batch metadata + arena, and the upload of a range of its loci on a context's stream
int batch_prepare(mprg_ctx *ctx, const int32_t *n_rows, const int32_t *n_cols, int32_t n_loci, mprg_batch **out);
int batch_upload_range(mprg_ctx *ctx, mprg_batch *b, const uint8_t *h_ascii, const int64_t *h_offsets, int l0,
                       int l1);
// the same for host rows already in the packed layout: straight into the arena, flags from the caller
int batch_upload_range_packed(mprg_ctx *ctx, mprg_batch *b, const uint8_t *h_packed, const int64_t *h_offsets,
                              const int32_t *h_flags, int l0, int l1);

// Host-side description of one level of tasks, resident on the device after level_upload().
struct Level {
    int n_tasks = 0;
    std::vector<DTask> tasks;
    std::vector<ScanUnit> units;
    long long total_cols = 0;  // sum of chunk-aligned window widths (multiple of 32)
    long long total_iv = 0;    // interval arena entries
    double algo_bytes = 0;     // sum r*c/2 + 4*r (row subsets) + 5*c
    bool has_n = false;
};

// Builds the level (device task table, scan units), uploads it together with the row-index arena,
// and runs scan -> classify [-> partition -> demote].  Results stay in the context's device buffers:
//   d_cls (uint8 per aligned column), d_reach (int per aligned column), d_iv / d_ivcnt.
void account_scan(mprg_ctx *ctx, const Level &lv);
int level_run(mprg_ctx *ctx, const mprg_batch *batch, const mprg_task *h_tasks, int n_tasks,
              const int32_t *h_rows, long long n_row_entries, int mml, bool do_partition, Level &lv);

// ---- clustering (cluster.cu, kmeans.cu) -----------------------------------------------------------
constexpr int KM_RAND_COUNT = 400;  // doubles of RandomState(2) a 10-init, K<=10 fit can consume

// k-mer counting problem: the n distinct long sequences of one clustering task
struct KmerProb {
    long long g_off;     // unpacked rows of the task (bytes into G)
    int w;               // gapped width
    int n;               // distinct long sequences
    int seq_off;         // offset into seq_rows (task-local row positions of the n sequences)
    long long useq_off;  // scratch bytes: ungapped symbols, n * w
    long long pos_off;   // scratch ints: ulen[n] | pos[n+1] | mref[Pmax] | kid[Pmax]
    long long tab_off;   // hash table slots (keys u64 / ming int), T entries
    int T;               // table size (power of two >= 2 * Pmax)
    int Pmax;            // upper bound on the number of k-mer positions
    long long x_off;     // count matrix (doubles), n * F, set by the host once F is known
    int big;             // bit0: numbered by the whole-grid kernels (launch_kmer_big), bit1: counted by
                         // launch_kmer_fill_big (shared-memory histograms)
    int pad;
};
constexpr long long KMER_BIG_POSITIONS = 1LL << 19;  // problems with at least this many k-mer positions are "big"

// per clustering problem: loop state of kmeans_cluster_seqs (cluster_sequences.py:249-274)
struct ClusterState {
    int status;      // 0 = looping, 1 = finished
    int run_kmeans;  // set by refcheck_kernel when KMeans(K) has to run
    int K;           // num_clusters
    int n;           // distinct long sequences
    int F;           // distinct k-mers
    int w;           // gapped width of the task
    long long g_off;         // unpacked rows of the task
    int mem_off;             // n + 1 offsets into mem_rows (members of each distinct sequence)
    int mem_rows_off;        // base of this problem's member list
    int assign_off;          // n ints: cluster of each distinct sequence
    long long maj_off;       // w bytes scratch: majority string
    long long x_off;         // count matrix
    long long kmd_off;       // KMeans double scratch
    long long kmi_off;       // KMeans int scratch
    int big;                 // KMeans of this problem runs on CTA groups (launch_kmeans_group)
    int big_ref;             // 1 + index of its flag block: checked by launch_refcheck_big (whole grid)
};
constexpr long long REFCHECK_BIG_SYMBOLS = 1LL << 22;  // rows * width from which the one-reference-like check is "big"
constexpr int REFCHECK_FLAG_INTS = 16;                 // flag block of a big problem: [0] any bad, [1 + c] cluster c bad
constexpr long long KMEANS_BIG_ELEMENTS = 1LL << 21;  // n * F from which a problem is "big"

constexpr long long DEDUPE_BIG_SYMBOLS = 1LL << 22;  // rows * width from which a level de-duplicates with the whole grid
// row_split: CTAs per task (deep tasks); 1 for the many small tasks of a pangenome level
cudaError_t launch_unpack(cudaStream_t s, const uint8_t *packed, const DTask *d_tasks, int n_tasks,
                          const int *d_rows, const long long *g_off, uint8_t *G, int row_split);
// dedupe_kernel's outputs for levels that hold a deep task; U: scratch of the size of G
cudaError_t launch_dedupe_big(cudaStream_t s, const DTask *d_tasks, int n_tasks, int max_rows, const long long *g_off,
                              const uint8_t *G, uint8_t *U, const long long *row_off, void *sig, int *leader_u,
                              int *leader_g, int *group, int *ulen, int *leaders, int *leader_len, int *n_ungapped,
                              int *n_gapped, int *err);
cudaError_t launch_dedupe(cudaStream_t s, const DTask *d_tasks, int n_tasks, int max_rows, const long long *g_off,
                          const uint8_t *G, const long long *row_off, void *sig, int *leader_u,
                          int *leader_g, int *group, int *ulen, int *leaders, int *leader_len,
                          int *n_ungapped, int *n_gapped, int *err);
// member lists (rows of each distinct long sequence) of one clustering problem
struct MemberProb {
    long long row_off;  // per-row arrays of the task
    int R;              // rows of the task
    int n_groups;       // distinct ungapped sequences of the task
    int k;              // kmer size: sequences shorter than k are "small"
    int mem_off;        // n + 1 ints
    int mem_rows_off;   // up to R ints
};
cudaError_t launch_scan_counts(cudaStream_t s, const int *counts, int n, int *offs);
cudaError_t launch_gather2(cudaStream_t s, const long long *src_off, const int *counts, const int *dst_off,
                           int n, const int *src_a, const int *src_b, int *dst_a, int *dst_b);
cudaError_t launch_members(cudaStream_t s, const MemberProb *probs, int n, const int *group,
                           const int *leader_len, int *long_of_group, int *mem_off, int *mem_rows);
size_t rowsig_bytes();
cudaError_t launch_kmer(cudaStream_t s, const void *d_probs, int n_probs, const int *seq_rows,
                        const uint8_t *G, int k, uint8_t *useq, int *ints, uint64_t *keys, int *ming,
                        int *out_F, int *err);
cudaError_t launch_kmer_big(cudaStream_t s, const void *d_probs, int q, const void *h_prob, const int *seq_rows,
                            const uint8_t *G, int k, uint8_t *useq, int *ints, uint64_t *keys, int *ming,
                            int *block_counts, int *out_F, int *err);
cudaError_t launch_kmer_fill_big(cudaStream_t s, const void *d_probs, int q, const void *h_prob, int F,
                                 const int *ints, double *X);
cudaError_t launch_set_features(cudaStream_t s, ClusterState *states, const int *F, int n);
cudaError_t launch_kmer_fill(cudaStream_t s, const void *d_probs, int n_probs, long long max_positions,
                             const int *ints, const int *F, double *X);
cudaError_t launch_refcheck(cudaStream_t s, ClusterState *states, int n_probs, const uint8_t *G,
                            const int *mem_off, const int *mem_rows, int *assign, uint8_t *maj,
                            int max_clusters, int *flags_out = nullptr);
// deep loci: the same check for one problem with the whole grid (majority per column strip and cluster, one warp
// per member row for the Hamming distances); maj holds max_clusters * w bytes for such a problem
cudaError_t launch_refcheck_big(cudaStream_t s, ClusterState *states, int q, int w, int rows, const uint8_t *G,
                                const int *mem_off, const int *mem_rows, int *assign, uint8_t *maj, int max_clusters,
                                int *flag_blocks);
// refcheck_grid.cu: the same check straight from the 4-bit packed rows (loci whose alphabet is ACGT-):
// bit-sliced counting, bound by one read of the member rows per pass; scratch: refgrid_scratch_ints ints
long long refgrid_scratch_ints(int R, int w, int n_words, int K_max);
cudaError_t launch_refcheck_grid(cudaStream_t s, ClusterState *states, int q, const DTask &t, int K_max,
                                 const uint8_t *packed, const int *rows_arena, const int *mem_off, const int *mem_rows,
                                 int *assign, uint8_t *maj, int *scratch, int n_member_rows, int *flags);
cudaError_t launch_refcheck_big_control(cudaStream_t s, ClusterState *states, int q, int max_clusters, int *flags);
long long kmeans_dscratch_doubles(long long n, long long F);
long long kmeans_iscratch_ints(long long n);
cudaError_t kmeans_upload_rand(const double *h_rand);
cudaError_t launch_kmeans_prepare(cudaStream_t s, const ClusterState *states, int n_probs, const double *X,
                                  double *dscratch, int *iscratch);
// deep loci: one problem, every initialisation on a group of co-resident CTAs (kmeans.cu compiled a second
// time with MPRG_KM_GROUP and L2-only loads); bars: KM_GROUP_WORDS zero-initialised unsigneds, 8-byte aligned
constexpr int KM_GROUP_WORDS = 32 + 16 * 10;
cudaError_t launch_kmeans_group(cudaStream_t s, ClusterState *states, int q, const double *X, double *dscratch,
                                int *iscratch, int *assign, int *newlab, unsigned *bars, bool prepare, int sm_count,
                                double *inertia_out = nullptr);
cudaError_t kmeans_group_upload_rand(const double *h_rand);
cudaError_t launch_kmeans(cudaStream_t s, ClusterState *states, int n_probs, const double *X,
                          double *dscratch, int *iscratch, int *assign, int *newlab, int *tickets,
                          long long max_elements);
cudaError_t launch_kmeans_single(cudaStream_t s, const double *X0, int n, int F, int K, double *dscratch,
                                 int *iscratch, int *labels, double *inertia, int *ticket);

}  // namespace mprg
