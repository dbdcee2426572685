// Launchers of the device kernels (definitions in scan.cu, partition.cu, cluster.cu, kmeans.cu).
#pragma once
#include "common.cuh"

namespace mprg {

cudaError_t launch_scan(cudaStream_t stream, bool has_n, const uint8_t *packed, const DTask *d_tasks,
                        const ScanUnit *d_units, int n_units, int max_unit_rows, const int *d_rows,
                        uint32_t *colOR, uint32_t *colNOR, unsigned *colB);
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

// Host-side description of one level of tasks, resident on the device after level_upload().
struct Level {
    int n_tasks = 0;
    std::vector<DTask> tasks;
    std::vector<ScanUnit> units;
    int max_unit_rows = 1;
    long long total_cols = 0;  // sum of chunk-aligned window widths (multiple of 32)
    long long total_iv = 0;    // interval arena entries
    double algo_bytes = 0;     // sum r*c/2 + 4*r (row subsets) + 5*c
    bool has_n = false;
};

// Builds the level (device task table, scan units), uploads it together with the row-index arena,
// and runs scan -> classify [-> partition -> demote].  Results stay in the context's device buffers:
//   d_cls (uint8 per aligned column), d_reach (int per aligned column), d_iv / d_ivcnt.
int level_run(mprg_ctx *ctx, const mprg_batch *batch, const mprg_task *h_tasks, int n_tasks,
              const int32_t *h_rows, long long n_row_entries, int mml, bool do_partition, Level &lv);

}  // namespace mprg
