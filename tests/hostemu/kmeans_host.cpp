// Test-only: host build of the KMeans device code (see cuda_shim.h).
#include "cuda_shim.h"
#include "../../make_prg_b200/csrc/kmeans.cu"
#include <vector>
extern "C" int emu_kmeans(const double *rand400, const double *X, int n, int F, int K, int *labels,
                          double *inertia) {
    using namespace mprg;
    memcpy(c_rand, rand400, sizeof(double) * KM_RAND_COUNT);
    std::vector<double> d(kmeans_dscratch_doubles(n, F));
    std::vector<int> ii(kmeans_iscratch_ints(n));
    kmeans_single_problem_kernel(X, n, F, K, d.data(), ii.data(), labels, inertia);
    return 0;
}
