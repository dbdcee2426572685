// Sizes of the KMeans scratch of one clustering problem; no CUDA dependency (also compiled by the host
// emulation of tests/hostemu), shared by kmeans.cu, the host engine and the device-side problem builder.
#pragma once

namespace mprg {

constexpr int KM_MAXK = 10;   // most clusters a fit can ask for (cluster_sequences.py:257: k > 10 ends the loop)
constexpr int KM_NINIT = 10;  // initialisations per fit (scikit-learn 1.3.0 default)
// KMeans scratch of one problem: shared block | KM_NINIT init blocks | best_c [KM_MAXK * F] | cc [KM_MAXK];
// ints: KM_NINIT init blocks (layout: kmeans.cu, km_bind_init)
__device__ __host__ inline long long km_shared_doubles(long long n, long long F) { return n * F + 2 * F + n + 8; }
__device__ __host__ inline long long km_init_doubles(long long n, long long F) {
    return 2LL * KM_MAXK * F + n * KM_MAXK + 3 * n + 4 * n + n + KM_MAXK * KM_MAXK + 4 * KM_MAXK +
           8;  // + [inertia, final-centres selector] in the last 8
}
__device__ __host__ inline long long km_init_ints(long long n) { return 2 * n + 8; }
__device__ __host__ inline long long km_dscratch_doubles(long long n, long long F) {
    return km_shared_doubles(n, F) + KM_NINIT * km_init_doubles(n, F) + KM_MAXK * F + KM_MAXK + 8;
}
__device__ __host__ inline long long km_iscratch_ints(long long n) { return KM_NINIT * km_init_ints(n) + 8; }

}  // namespace mprg
