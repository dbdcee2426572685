// Test-only shim: lets the *device* functions of make_prg_b200/csrc/kmeans.cu be compiled as plain
// C++ (one "thread", blockDim.x == 1) so their arithmetic can be debugged in a container without a
// GPU.  Never part of libmprg.so; the product has no CPU path.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#define MPRG_HOST_EMU 1
#define __device__
#define __host__
#define __global__
#define __constant__ static
#define __shared__ static
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(...)
struct Dim3Shim { int x; };
static Dim3Shim threadIdx = {0}, blockDim = {1}, blockIdx = {0};
static inline void __syncthreads() {}
static inline double __dadd_rn(double a, double b) { volatile double r = a + b; return r; }
static inline double __dsub_rn(double a, double b) { volatile double r = a - b; return r; }
static inline double __dmul_rn(double a, double b) { volatile double r = a * b; return r; }
static inline int __popc(unsigned x) { return __builtin_popcount(x); }
namespace mprg {
constexpr int KM_RAND_COUNT = 400;
struct ClusterState {
    int status, run_kmeans, K, n, F, w;
    long long g_off; int mem_off, mem_rows_off, assign_off; long long maj_off, x_off, kmd_off, kmi_off; int big, pad;
};
}
