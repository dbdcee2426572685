// Entry points declared in include/mprg.h that are not implemented yet.
#include "common.cuh"
#define STUB(ctx) do { if (ctx) (ctx)->err = "not implemented"; return MPRG_E_INTERNAL; } while (0)
extern "C" int mprg_dedupe_rows(mprg_ctx *ctx, const mprg_batch *, const mprg_task *, int32_t, const int32_t *, int64_t, const int64_t *, int32_t *, int32_t *, int32_t *, int32_t *) { STUB(ctx); }
extern "C" int mprg_kmer_counts(mprg_ctx *ctx, const mprg_batch *, const mprg_task *, const int32_t *, int32_t, int32_t *, int32_t *, double *, int64_t) { STUB(ctx); }
extern "C" int mprg_kmeans(mprg_ctx *ctx, const double *, int32_t, int32_t, int32_t, int32_t *, double *) { STUB(ctx); }
extern "C" int mprg_one_ref_like(mprg_ctx *ctx, const mprg_batch *, const mprg_task *, const int32_t *, const int32_t *, int32_t, int32_t *) { STUB(ctx); }
extern "C" int mprg_cluster_tasks(mprg_ctx *ctx, const mprg_batch *, const mprg_task *, int32_t, const int32_t *, int64_t, int32_t, const int64_t *, int32_t *, int32_t *) { STUB(ctx); }
extern "C" int mprg_build(mprg_ctx *ctx, mprg_batch *, int32_t, int32_t, mprg_result **) { STUB(ctx); }
extern "C" void mprg_result_free(mprg_result *) {}
extern "C" int32_t mprg_result_n_loci(const mprg_result *) { return 0; }
extern "C" int32_t mprg_result_status(const mprg_result *, int32_t) { return 0; }
extern "C" const char *mprg_result_prg(const mprg_result *, int32_t, int64_t *) { return nullptr; }
extern "C" int32_t mprg_result_n_nodes(const mprg_result *, int32_t) { return 0; }
extern "C" int32_t mprg_result_n_sites(const mprg_result *, int32_t) { return 0; }
extern "C" int mprg_result_nodes(const mprg_result *, int32_t, int32_t *, int32_t *, int32_t *, int32_t *, int32_t *, int32_t *, int64_t *, int32_t *) { return MPRG_E_INTERNAL; }
extern "C" int64_t mprg_result_row_pool_size(const mprg_result *, int32_t) { return 0; }
extern "C" int mprg_result_row_pool(const mprg_result *, int32_t, int32_t *) { return MPRG_E_INTERNAL; }
