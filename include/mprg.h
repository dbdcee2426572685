/*
 * mprg.h -- C ABI of libmprg.so, the sm_100a implementation of make_prg's `from_msa` compute core.
 *
 * The reference (iqbal-lab-org/make_prg v0.5.0) is pure Python and has no FFI layer; the seams this
 * library sits behind are the Python calls listed next to each entry point (paths relative to the
 * reference root).  INTEGRATION.md shows the ctypes binding a maintainer would add.
 *
 * Conventions
 *   - every function returns 0 on success, a negative MPRG_E_* code otherwise; no C++ exception
 *     crosses the ABI; mprg_last_error(ctx) gives the text of the last failure on that context.
 *   - plain pointers and sizes only.  Every pointer an entry point takes (h_*) is HOST memory owned by the
 *     caller (pinned or pageable; the library never frees caller memory).  Device memory is owned by the
 *     library: batches (mprg_batch) hold the packed MSAs in HBM, results (mprg_result) hold what came back.
 *   - a context is used from one host thread at a time; all its work is enqueued on the context's
 *     stream; entry points that return host results synchronise that stream.  Several contexts of one
 *     GPU may build side by side from their own host threads (that is how throughput is reached:
 *     INTEGRATION.md section 6, make_prg_b200.device.BuildPipeline); batches and results may be
 *     freed from any thread, a batch may be built by any context of its GPU.
 *   - there is NO CPU fallback: without a CUDA device mprg_create fails with MPRG_E_NO_DEVICE.
 *
 * Symbol codes of the 4-bit packed MSA (rows padded to 16 bytes = 32 columns with MPRG_SYM_PAD; inside
 * a 32-column chunk, column c is nibble c / 4 of the 32-bit little-endian word c % 4):
 *   '-' = 0;  A C G T = 1 3 5 7;  R Y K = 9 11 13;  M S W = 2 4 6;  N = 8;  pad/disallowed = 15
 *   (MPRG_ALPHABET[code] is the character).  The gap is the zero nibble and every common symbol --
 *   the four bases and the padding -- is odd, so "these rows hold no gap in these columns" is one AND
 *   over the words and a test of bit 0 of every nibble (scan kernel); the even codes (M S W N, rare)
 *   only send the scan through its exact path.
 */
#ifndef MPRG_H
#define MPRG_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MPRG_OK 0
#define MPRG_E_NO_DEVICE (-1)
#define MPRG_E_CUDA (-2)
#define MPRG_E_BAD_ARG (-3)
#define MPRG_E_PARTITION (-4) /* PartitioningError (interval_partition.py:12) */
#define MPRG_E_INTERNAL (-5)

#define MPRG_SYM_GAP 0
#define MPRG_SYM_N 8
#define MPRG_ALPHABET "-AMCSGWTNR?Y?K??"
#define MPRG_SYM_PAD 15

#define MPRG_IV_MATCH 0
#define MPRG_IV_NONMATCH 1

/* node kinds of the recursion tree (recursion_tree.py:176-391) */
#define MPRG_NODE_LEAF 0
#define MPRG_NODE_INTERVAL 1
#define MPRG_NODE_CLUSTER 2

/* per-locus status of mprg_build */
#define MPRG_LOCUS_OK 0
#define MPRG_LOCUS_CURATION_ERROR 1 /* SequenceCurationError => locus skipped (from_msa.py:147-151) */

typedef struct mprg_ctx mprg_ctx;
typedef struct mprg_batch mprg_batch;   /* a set of loci resident in HBM (4-bit packed) */
typedef struct mprg_result mprg_result; /* trees + PRG strings of one mprg_build call */

/* A sub-alignment: rows (subset, input order) x columns [c0, c1) of one locus of a batch.
 * Mirrors what NodeFactory.build receives (recursion_tree.py:401-406): every sub-alignment the
 * reference ever builds is a (row subset, contiguous column range) of the root MSA. */
typedef struct {
    int32_t locus;    /* index into the batch */
    int32_t rows_off; /* offset into the caller's row-index array, or -1 for "all rows" */
    int32_t n_rows;
    int32_t c0, c1;
} mprg_task;

typedef struct {
    int32_t start, stop; /* closed interval, columns relative to the task's c0 */
    int32_t type;        /* MPRG_IV_MATCH / MPRG_IV_NONMATCH */
} mprg_interval;

/* ---- context ------------------------------------------------------------------------------- */
int mprg_create(int device_ordinal, mprg_ctx **out);
void mprg_destroy(mprg_ctx *ctx);
const char *mprg_last_error(const mprg_ctx *ctx);
int mprg_device_info(const mprg_ctx *ctx, int *sm_count, int *cc_major, int *cc_minor);
/* number of kernels this context has launched so far (bench.py's gpu_launches) */
int64_t mprg_launch_count(const mprg_ctx *ctx);
/* Which kernel variants the engine picked since the last reset (tests pin the deep-locus paths with it):
 * out[MPRG_PATH_x] = launches of that variant, worker contexts included. */
#define MPRG_PATH_KMEANS_CTA 0        /* kmeans_kernel: one CTA (1-4 warps) per initialisation */
#define MPRG_PATH_KMEANS_GROUP 1      /* kmeans_group_kernel: CTA groups per initialisation (deep loci) */
#define MPRG_PATH_REFCHECK_CTA 2      /* refcheck_kernel: one CTA per problem */
#define MPRG_PATH_REFCHECK_GRID 3     /* whole-grid one-reference-like check */
#define MPRG_PATH_REFCHECK_GRID_MULTI 4 /* ... of which with more than one cluster */
#define MPRG_PATH_KMER_GRID 5         /* whole-grid k-mer numbering */
#define MPRG_PATH_DEDUPE_GRID 6       /* whole-grid de-duplication */
#define MPRG_PATH_COUNT 8
int mprg_path_counts(mprg_ctx *ctx, int64_t *out, int reset);
/* Device time (ms, CUDA events around every launch) of the KMeans launches of the level loop since the last reset,
 * their number, and the problems they covered (a launch covers every clustering problem of its level; a problem
 * that has left the loop costs an early exit): bench.py's "kmeans" object.  Worker contexts included. */
int mprg_kmeans_stats(mprg_ctx *ctx, double *ms, int64_t *launches, int64_t *problems, int reset);
/* device time (ms, CUDA events on the context's stream) and algorithmic bytes of the column-scan
 * kernel accumulated since the last reset; used by bench.py for the roofline object */
int mprg_scan_stats(mprg_ctx *ctx, double *ms, double *bytes, int64_t *launches, int reset);

/* Upper bound of the ranges (one host thread + stream + scratch set each) a build is cut into; default
 * min(8, host cores / ranks on this host) or the MPRG_WORKERS environment variable.  The device-resident level
 * loop uses ONE range for a resident batch and TWO from host buffers (the upload of one overlaps the kernels of
 * the other; MPRG_DEV_RANGES overrides); the full count only applies to the host-driven loop kept as the checked
 * alternative (MPRG_HOST_LOOP=1) and to the host threads that assemble PRG strings of loci holding RYKMSW. */
int mprg_set_workers(mprg_ctx *ctx, int32_t n_workers);
/* How the host thread of a build waits for the device at the synchronisation points of the level loop and of the
 * uploads: 0 (default) = cudaStreamSynchronize (the driver's choice, in practice a spin: lowest latency for one
 * build at a time), 1 = the thread sleeps on an event created with cudaEventBlockingSync, 2 = it polls
 * cudaStreamQuery and yields the core between polls.  Contexts that run side by side (device.BuildPipeline:
 * several builds in flight per GPU, one host thread each) choose by measurement (profiles/r2_lanes_sweep.txt). */
int mprg_set_wait_mode(mprg_ctx *ctx, int32_t mode);
/* host<->device bytes copied by this context since the last reset (bench.py's e2e object) */
int mprg_copy_stats(mprg_ctx *ctx, int64_t *h2d_bytes, int64_t *d2h_bytes, int reset);
/* CUDA-event stopwatch on the context's stream: op 0 records the start, op 1 records the stop,
 * waits for it and returns the elapsed device time in ms */
int mprg_timer(mprg_ctx *ctx, int op, double *ms);
/* per-launch log of the column-scan kernel (algorithmic bytes, device ms), newest last */
int mprg_scan_log(mprg_ctx *ctx, double *bytes, double *ms, int32_t capacity, int32_t *n, int reset);
/* Measurement aid for the roofline of the scan kernel: stream the packed batch once with the scan's
 * own access pattern (one warp per 16-row x 512-byte tile, 128-bit ld.global.nc) and nothing else;
 * *ms = device time of that one launch (CUDA events on the context's stream), *bytes = bytes read.
 * What a bare read of the same bytes reaches at this launch size bounds what the scan can reach. */
int mprg_read_yardstick(mprg_ctx *ctx, const mprg_batch *batch, double *bytes, double *ms);

/* ---- loader -> HBM (replaces the in-memory Biopython MSA of io_utils.py:17-49) --------------- */
/* h_ascii: concatenated row-major ASCII matrices (upper or lower case, N already replaced by the
 * host loader); locus i occupies n_rows[i]*n_cols[i] bytes starting at h_offsets[i].
 * Copies to the device, packs to 4 bits there, records per-locus alphabet flags. */
int mprg_batch_upload(mprg_ctx *ctx, const uint8_t *h_ascii, const int64_t *h_offsets,
                      const int32_t *n_rows, const int32_t *n_cols, int32_t n_loci,
                      mprg_batch **out);
void mprg_batch_free(mprg_ctx *ctx, mprg_batch *batch);
/* flags[i] bit0: locus holds a character outside ACGTRYKMSWN- (=> SequenceCurationError),
 * bit1: holds N, bit2: holds RYKMSW */
int mprg_batch_flags(mprg_ctx *ctx, const mprg_batch *batch, int32_t *h_flags);
/* debug/parity: copy the packed rows of one locus back (n_rows * stride bytes) */
int mprg_batch_download_packed(mprg_ctx *ctx, const mprg_batch *batch, int32_t locus,
                               uint8_t *h_out, int64_t capacity, int32_t *stride);

/* ---- kernel (a): column scan  (get_consensus_from_MSA seq_utils.py:219-239,
 *      has_empty_sequence seq_utils.py:37-42 in its gap-reach form) ---------- */
/* For each task: h_consensus gets c1-c0 bytes ('*' for non-match, else the base), h_gap_reach gets
 * c1-c0 int32 (relative to c0): has_empty_sequence([s,e]) == (gap_reach[s] >= e).
 * Outputs of task t start at h_col_offsets[t] (the caller provides the prefix sums). */
int mprg_scan_tasks(mprg_ctx *ctx, const mprg_batch *batch, const mprg_task *h_tasks,
                    int32_t n_tasks, const int32_t *h_rows, int64_t n_row_entries,
                    const int64_t *h_col_offsets, uint8_t *h_consensus, int32_t *h_gap_reach);

/* ---- kernel (a'): interval partition (IntervalPartitioner interval_partition.py:81-252,
 *      NodeFactory._get_vertical_partition recursion_tree.py:500-513) --------------------------- */
/* Full vertical partition of each task: scan + run state machine + single-sequence demotion
 * (enforce_multisequence_nonmatch_intervals) + bijection check.  h_intervals receives, per task,
 * h_iv_counts[t] intervals sorted by start at h_iv_offsets[t] (capacity max(1, c1-c0) each). */
int mprg_partition_tasks(mprg_ctx *ctx, const mprg_batch *batch, const mprg_task *h_tasks,
                         int32_t n_tasks, const int32_t *h_rows, int64_t n_row_entries,
                         int32_t min_match_length, const int64_t *h_iv_offsets,
                         mprg_interval *h_intervals, int32_t *h_iv_counts);
/* The state machine alone on a caller-supplied consensus string (the reference's unit tests drive
 * IntervalPartitioner with hand-written consensus strings and an empty alignment,
 * tests/from_msa/test_interval_partition.py:80-136).  gap_reach may be NULL (no empty rows). */
int mprg_partition_consensus(mprg_ctx *ctx, const uint8_t *h_consensus, const int32_t *h_gap_reach,
                             int32_t n_cols, int32_t min_match_length, mprg_interval *h_intervals,
                             int32_t capacity, int32_t *h_count);

/* ---- kernels (b), (c): clustering of one sub-alignment per task
 *      (kmeans_cluster_seqs cluster_sequences.py:211-296) --------------------------------------- */
/* Row de-duplication (cluster_sequences.py:220-233, seq_utils.py:58-70): per row of each task the
 * index (in first-seen order) of its distinct ungapped sequence, the ungapped length, and per task
 * the number of distinct ungapped / gapped rows.  Row outputs start at h_row_offsets[t]. */
int mprg_dedupe_rows(mprg_ctx *ctx, const mprg_batch *batch, const mprg_task *h_tasks,
                     int32_t n_tasks, const int32_t *h_rows, int64_t n_row_entries,
                     const int64_t *h_row_offsets, int32_t *h_group, int32_t *h_ungapped_len,
                     int32_t *h_n_ungapped, int32_t *h_n_gapped);
/* k-mer count matrix of one task (count_distinct_kmers / count_kmer_occurrences
 * cluster_sequences.py:26-56): distinct ungapped sequences of length >= k in first-seen order are
 * the rows, k-mers in first-occurrence order the columns.  h_counts (row-major doubles, capacity
 * given in elements) may be NULL to query the shape only. */
int mprg_kmer_counts(mprg_ctx *ctx, const mprg_batch *batch, const mprg_task *h_task,
                     const int32_t *h_rows, int32_t kmer_size, int32_t *n_seqs, int32_t *n_kmers,
                     double *h_counts, int64_t capacity);
/* KMeans(n_clusters=K, random_state=2, algorithm="elkan") of scikit-learn 1.3.0 (n_init=10),
 * .fit(X).predict(X) as called at cluster_sequences.py:262-266.  X row-major [n, F] doubles. */
int mprg_kmeans(mprg_ctx *ctx, const double *h_X, int32_t n, int32_t F, int32_t K,
                int32_t *h_labels, double *h_inertia);
/* The same with the execution shape chosen by the caller (tests): mode 0 = the engine's choice, 1 = one
 * CTA per initialisation, 2 = every initialisation on a group of co-resident CTAs (deep loci).  All
 * modes return bit-identical labels and inertia. */
int mprg_kmeans_mode(mprg_ctx *ctx, const double *h_X, int32_t n, int32_t F, int32_t K,
                     int32_t *h_labels, double *h_inertia, int32_t mode);
/* sequences_are_one_reference_like over clusters of gapped rows (cluster_sequences.py:59-111):
 * cluster_of_row[i] in [0, n_clusters) for each row of the task (in member order = the order the
 * caller lists them in h_rows); out_flags[c] = 1 when cluster c is one-reference-like. */
int mprg_one_ref_like(mprg_ctx *ctx, const mprg_batch *batch, const mprg_task *h_task,
                      const int32_t *h_rows, const int32_t *h_cluster_of_row, int32_t n_clusters,
                      int32_t *h_flags);
/* Whole kmeans_cluster_seqs for a list of tasks.  Per row (at h_row_offsets[t]) the cluster index in
 * the order of ClusteringResult.clustered_ids; h_n_clusters[t] == 1 means "no clustering". */
int mprg_cluster_tasks(mprg_ctx *ctx, const mprg_batch *batch, const mprg_task *h_tasks,
                       int32_t n_tasks, const int32_t *h_rows, int64_t n_row_entries,
                       int32_t kmer_size, const int64_t *h_row_offsets, int32_t *h_cluster,
                       int32_t *h_n_clusters);

/* ---- the whole path: PrgBuilder.__init__ + build_prg for every locus of a batch
 *      (prg_builder.py:24-42,100-110; NodeFactory.build recursion_tree.py:401-471) -------------- */
int mprg_build(mprg_ctx *ctx, mprg_batch *batch, int32_t max_nesting, int32_t min_match_length,
               mprg_result **out);
/* The same for alignments that are built BELOW an existing node, NodeFactory.build(alignment, builder,
 * parent_node) as LeafNode._update_leaf calls it (recursion_tree.py:373-376, 431-432): h_parent_level[l]
 * is parent_node.nesting_level for locus l, or -1 to build it as a root.  A non-root alignment is never
 * forced into a MultiIntervalNode and its nesting starts at the parent's level; node ids in the result
 * count from 0 (the caller adds PrgBuilder.next_node_id).  h_parent_level == NULL is mprg_build. */
int mprg_build_sub(mprg_ctx *ctx, mprg_batch *batch, int32_t max_nesting, int32_t min_match_length,
                   const int32_t *h_parent_level, mprg_result **out);
/* The same from HOST ASCII in one call: loader output in, batch + result out.  The loci are cut into
 * the ranges mprg_build uses and every range is copied, packed and built by its own worker thread and
 * stream, so the host-to-device copy of one range overlaps the kernels of the others (the end-to-end
 * path of `make_prg from_msa`, from_msa.py:114-123 per locus).  Arguments as mprg_batch_upload; the
 * caller frees *out_batch with mprg_batch_free and *out_res with mprg_result_free. */
int mprg_build_ascii(mprg_ctx *ctx, const uint8_t *h_ascii, const int64_t *h_offsets,
                     const int32_t *n_rows, const int32_t *n_cols, int32_t n_loci, int32_t max_nesting,
                     int32_t min_match_length, mprg_batch **out_batch, mprg_result **out_res);
/* The same from HOST rows that are already in the 4-bit device layout (mprg_fasta_packed / mprg_pack_rows):
 * locus i occupies n_rows[i] * 16 * ceil(n_cols[i] / 32) bytes at h_packed + h_offsets[i]; h_flags[i] are
 * its alphabet flags.  Half the bytes of mprg_build_ascii cross PCIe and no pack kernel runs. */
int mprg_build_packed(mprg_ctx *ctx, const uint8_t *h_packed, const int64_t *h_offsets, const int32_t *n_rows,
                      const int32_t *n_cols, const int32_t *h_flags, int32_t n_loci, int32_t max_nesting,
                      int32_t min_match_length, mprg_batch **out_batch, mprg_result **out_res);
/* A result that only holds the given PRG strings (no trees): lets the writers below serve PRGs that
 * did not come out of mprg_build, e.g. PrgBuilder.build_prg() of a host-side tree (prg_builder.py:100-105) */
int mprg_result_from_prgs(const char *const *prgs, const int64_t *lengths, int32_t n, mprg_result **out);
void mprg_result_free(mprg_result *res);
int32_t mprg_result_n_loci(const mprg_result *res);
int32_t mprg_result_status(const mprg_result *res, int32_t locus);
/* all loci at once: status and PRG length per locus (either array may be NULL) */
int mprg_result_statuses(const mprg_result *res, int32_t *h_status, int64_t *h_prg_length);
/* PRG string of a locus (not NUL-terminated); valid until mprg_result_free */
const char *mprg_result_prg(const mprg_result *res, int32_t locus, int64_t *length);
int32_t mprg_result_n_nodes(const mprg_result *res, int32_t locus);
int32_t mprg_result_n_sites(const mprg_result *res, int32_t locus);
/* Node table of a locus in pre-order (== node_id order).  Arrays of n_nodes entries; rows of node i
 * are h_rows[row_off[i] .. row_off[i]+n_rows[i]) of the locus' own row-index pool. */
int mprg_result_nodes(const mprg_result *res, int32_t locus, int32_t *kind, int32_t *parent,
                      int32_t *nesting_level, int32_t *c0, int32_t *c1, int32_t *n_rows,
                      int64_t *row_off, int32_t *n_children);
int64_t mprg_result_row_pool_size(const mprg_result *res, int32_t locus);
int mprg_result_row_pool(const mprg_result *res, int32_t locus, int32_t *h_rows);

/* ---- host-side I/O either side of the path (plain host threads) ---------
 * Loader: load_alignment_file (io_utils.py:17-49) for FASTA / FASTA.gz files.  Every file becomes one
 * locus of an mprg_msa_set: upper-cased row-major ASCII rows in ONE host buffer (pinned when pin != 0
 * and a device is present) laid out exactly as mprg_build_ascii / mprg_batch_upload take it.  Record
 * ids are the first whitespace token of each title (Biopython).  N replacement (io_utils.py:35-47,
 * seq_utils.py:246-290) is done by the loader, bit for bit as the reference's sha256-seeded
 * random.Random does it (mprg_replace_n); loci that held N carry MPRG_LOAD_FLAG_HAS_N. */
#define MPRG_LOAD_OK 0
#define MPRG_LOAD_NO_RECORDS 1 /* ValueError("No records found in handle") => EmptyMSAError */
#define MPRG_LOAD_RAGGED 2     /* ValueError("Sequences must all be the same length") */
#define MPRG_LOAD_IO_ERROR 3   /* cannot open / read / inflate: the caller re-raises through Python's open */
#define MPRG_LOAD_NOT_ASCII 4  /* bytes >= 0x80: left to the caller's text decoder */
#define MPRG_LOAD_FLAG_HAS_N 1
/* two records share an id (first token of the title): the reference selects cluster sub-alignments by id
 * (recursion_tree.py:558-572), so its sub-alignments overlap on such a file; this engine cuts by row */
#define MPRG_LOAD_FLAG_DUPLICATE_IDS 2
typedef struct mprg_msa_set mprg_msa_set;
/* N replacement alone on one upper-cased row-major matrix, in place */
int mprg_replace_n(uint8_t *h_ascii, int32_t n_rows, int32_t n_cols);
/* mode: MPRG_LOADMODE_PIN = the output buffer is pinned host memory (pooled); MPRG_LOADMODE_PACKED = the
 * matrices come out in the 4-bit device layout (mprg_fasta_packed, for mprg_build_packed: half the bytes
 * cross PCIe and no pack kernel runs) and the text is dropped unless MPRG_LOADMODE_KEEP_ASCII is set */
#define MPRG_LOADMODE_PIN 1
#define MPRG_LOADMODE_PACKED 2
#define MPRG_LOADMODE_KEEP_ASCII 4
int mprg_fasta_load(const char *const *paths, int32_t n_files, int32_t n_threads, int32_t mode,
                    mprg_msa_set **out);
void mprg_fasta_free(mprg_msa_set *set);
/* Borrowed views, valid until mprg_fasta_free; any out pointer may be NULL.  Locus i occupies
 * n_rows[i] * n_cols[i] bytes at h_ascii + h_offsets[i] (0 bytes unless status[i] == MPRG_LOAD_OK). */
/* packed mode: locus i occupies n_rows[i] * stride(n_cols[i]) bytes at h_packed + h_packed_offsets[i]
 * (stride = 16 * ceil(cols / 32)); alphabet_flags[i] = what mprg_batch_flags reports after a device pack */
int mprg_fasta_packed(const mprg_msa_set *set, uint8_t **h_packed, int64_t *packed_bytes,
                      const int64_t **h_packed_offsets, const int32_t **alphabet_flags);
/* ASCII rows (row-major, n_rows * n_cols bytes, either case) -> the packed layout on the host;
 * h_packed holds n_rows * stride(n_cols) bytes; *flags = alphabet flags as the device pack reports them */
int mprg_pack_rows(const uint8_t *h_ascii, int32_t n_rows, int32_t n_cols, uint8_t *h_packed, int64_t capacity,
                   int32_t *flags);
int mprg_fasta_info(const mprg_msa_set *set, int32_t *n_loci, uint8_t **h_ascii, int64_t *ascii_bytes,
                    const int64_t **h_offsets, const int32_t **n_rows, const int32_t **n_cols,
                    const int32_t **status, const int32_t **flags);
/* the title lines of a locus (text after '>'), joined by '\n', not NUL-terminated */
const char *mprg_fasta_titles(const mprg_msa_set *set, int32_t locus, int64_t *length);

/* Writers.  mprg_encode_prg: PrgEncoder.encode (prg_encoder.py:44-91), the uint32 values that
 * PrgEncoder.write stores little-endian.  mprg_prg_to_gfa: GFA_Output.write_gfa's text (gfa.py:39-109,
 * header included).  Both return *n = elements / bytes needed and fill `out` only when capacity >= *n;
 * > 0 return codes are the reference's exceptions. */
#define MPRG_ENC_INVALID_UNIT 1         /* EncodeError / "Invalid prg sequence" */
#define MPRG_ENC_ODD_MARKER_REPEATED 2  /* ValueError: odd site marker found > 2 times */
#define MPRG_ENC_OVERFLOW 3             /* marker does not fit 4 bytes (OverflowError in to_bytes) */
int mprg_encode_prg(const char *prg, int64_t length, uint32_t *out, int64_t capacity, int64_t *n);
int mprg_prg_to_gfa(const char *prg, int64_t length, char *out, int64_t capacity, int64_t *n);
/* Final files of a run as InputOutputFiles.create_final_files lays them out
 * (input_output_files.py:70-135): <prefix>.prg.fa (records sorted by "<name>.prg.fa"), and for ONE
 * locus <prefix>.prg.bin / <prefix>.prg.gfa, for several <prefix>.prg.bin.zip / <prefix>.prg.gfa.zip
 * (stored archives of <name>.bin / <name>.gfa in the order added).  mprg_writer_add encodes loci
 * h_loci[0..n) of a result on n_threads host threads and appends them; results can be freed afterwards. */
#define MPRG_WRITE_PRG 1
#define MPRG_WRITE_BIN 2
#define MPRG_WRITE_GFA 4
/* this writer produces one PART of a run (one GPU shard): archives even for a single locus; the parts are
 * turned into the final files by mprg_merge_outputs */
#define MPRG_WRITE_PART 8
/* mprg_merge_outputs only: also merge the parts' <prefix>.update_DS.zip archives */
#define MPRG_WRITE_DS 16
typedef struct mprg_writer mprg_writer;
int mprg_writer_open(const char *output_prefix, int32_t what, mprg_writer **out);
int mprg_writer_add(mprg_writer *w, const mprg_result *res, const int32_t *h_loci, const char *const *names,
                    int32_t n, int32_t n_threads);
/* <prefix>.update_DS.zip (prg_builder.py:145-147, input_output_files.py:95-104): one member per locus, named by
 * the locus, holding what its PrgBuilder is made of as tables (pre-order node table, row subsets, record titles,
 * the root alignment as 4-bit packed rows, the PRG string; layout "MPRGDS01" in csrc/hostio.cpp) instead of a
 * pickle of Python objects -- the Python host builds the objects on load.  msas: the loader's set the result was
 * built from; h_loci index both. */
int mprg_writer_add_ds(mprg_writer *w, const mprg_result *res, const mprg_msa_set *msas, const int32_t *h_loci,
                       const char *const *names, int32_t n, int32_t max_nesting, int32_t min_match_length,
                       int32_t n_threads);
/* finishes the files and frees the writer on success; on failure read mprg_writer_error, then abort */
int mprg_writer_close(mprg_writer *w, int64_t *n_loci, int64_t *bytes_written);
void mprg_writer_abort(mprg_writer *w);
const char *mprg_writer_error(const mprg_writer *w);
/* Final files of a run built as several parts (replaces the concatenation of per-process files in
 * make_prg/utils/input_output_files.py:70-135): .prg.fa records merged in sorted order, archive members
 * appended part by part, plain .prg.bin / .prg.gfa when the whole run holds one locus.  Missing parts hold no
 * locus; the parts are removed.  Files appear under their final names only when complete. */
int mprg_merge_outputs(const char *const *part_prefixes, int32_t n_parts, const char *output_prefix, int32_t what,
                       int64_t *n_loci, char *err_buf, int64_t err_capacity);

#ifdef __cplusplus
}
#endif
#endif /* MPRG_H */
