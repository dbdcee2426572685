// Micro-benchmark (GPU box): what a pure streaming read of the bench-sized packed batch can reach on
// this B200, as a yardstick for scan_kernel.  nvcc -arch=sm_100a -O3 -o membench membench.cu
#include <cstdio>
#include <cstdint>
#include <vector>
#include <algorithm>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA %s line %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

__device__ __forceinline__ uint4 ld_stream(const uint4 *p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
// K2: grid-stride, UNR independent 16-byte loads per thread per iteration
template <int UNR>
__global__ void k_stride(const uint4 *__restrict__ p, long long n, uint32_t *out) {
    uint32_t a = 0;
    const long long stride = (long long)gridDim.x * blockDim.x;
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + (UNR - 1) * stride < n; i += UNR * stride) {
        uint4 v[UNR];
#pragma unroll
        for (int u = 0; u < UNR; ++u) v[u] = ld_stream(p + i + u * stride);
#pragma unroll
        for (int u = 0; u < UNR; ++u) a |= v[u].x | v[u].y | v[u].z | v[u].w;
    }
    for (; i < n; i += stride) { uint4 v = ld_stream(p + i); a |= v.x | v.y | v.z | v.w; }
    if (a == 0x12345678u) out[0] = a;
}
// K1: one warp per tile of `rows` rows x 512 B (consecutive), 4 rows per trip, one trip of lookahead
__global__ void __launch_bounds__(128) k_tile(const uint4 *__restrict__ p, int n_tiles, int rows, uint32_t *out) {
    const int lane = threadIdx.x & 31;
    const int tile = blockIdx.x * 4 + (threadIdx.x >> 5);
    if (tile >= n_tiles) return;
    const uint4 *base = p + (long long)tile * rows * 32 + lane;
    uint32_t a = 0;
    uint4 nx[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) nx[u] = ld_stream(base + u * 32);
    for (int r = 0; r < rows; r += 4) {
        uint4 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) v[u] = nx[u];
        if (r + 4 < rows) {
#pragma unroll
            for (int u = 0; u < 4; ++u) nx[u] = ld_stream(base + (r + 4 + u) * 32);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) a |= v[u].x | v[u].y | v[u].z | v[u].w;
    }
    if (a == 0x12345678u) out[0] = a;
}
// K3: one warp per tile, whole tile fetched with one TMA bulk copy per 2 KB trip into a 4-stage ring
__device__ __forceinline__ uint32_t smem_u32(const void *q) { return (uint32_t)__cvta_generic_to_shared(q); }
__global__ void __launch_bounds__(128) k_tma(const uint8_t *__restrict__ p, int n_tiles, int rows, uint32_t *out) {
    extern __shared__ __align__(128) uint8_t sm[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int tile = blockIdx.x * 4 + warp;
    if (tile >= n_tiles) return;
    const uint32_t ring = smem_u32(sm) + warp * 8192, bars = smem_u32(sm) + 32768 + warp * 32;
    if (lane == 0) {
        for (int s = 0; s < 4; ++s) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bars + s * 8));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    const uint8_t *src = p + (long long)tile * rows * 512;
    const int n_trips = (rows + 3) / 4;
    auto issue = [&](int t) {
        if (lane == 0) {
            const uint32_t bytes = (uint32_t)min(4, rows - t * 4) * 512u, st = t & 3;
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bars + st * 8), "r"(bytes) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(ring + st * 2048), "l"(src + (long long)t * 2048), "r"(bytes), "r"(bars + st * 8) : "memory");
        }
    };
    for (int t = 0; t < min(4, n_trips); ++t) issue(t);
    uint32_t a = 0;
    for (int t = 0; t < n_trips; ++t) {
        const uint32_t st = t & 3, par = (t >> 2) & 1;
        asm volatile("{\n.reg .pred p;\nW:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D;\nbra W;\nD:\n}" ::"r"(bars + st * 8), "r"(par) : "memory");
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            uint4 v;
            asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(ring + st * 2048 + u * 512 + lane * 16));
            a |= v.x | v.y | v.z | v.w;
        }
        __syncwarp();
        if (t + 4 < n_trips) issue(t + 4);
    }
    if (a == 0x12345678u) out[0] = a;
}

int main() {
    const long long bytes = 1000LL * 200 * 512;  // the bench batch
    uint8_t *d; uint32_t *out; uint8_t *flush;
    CK(cudaMalloc(&d, bytes)); CK(cudaMalloc(&out, 64)); CK(cudaMalloc(&flush, 256 << 20));
    CK(cudaMemset(d, 0x11, bytes));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    CK(cudaFuncSetAttribute(k_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768 + 128));
    auto timeit = [&](const char *name, auto launch) {
        std::vector<float> ms;
        for (int rep = 0; rep < 12; ++rep) {
            cudaMemsetAsync(flush, rep, 256 << 20);
            cudaDeviceSynchronize();
            cudaEventRecord(e0); launch(); cudaEventRecord(e1); cudaEventSynchronize(e1);
            float t; cudaEventElapsedTime(&t, e0, e1); if (rep >= 2) ms.push_back(t);
        }
        std::sort(ms.begin(), ms.end());
        const float med = ms[ms.size() / 2];
        printf("%-34s median %.2f us  min %.2f us  -> %.0f GB/s (median)\n", name, med * 1e3, ms[0] * 1e3, bytes / (med * 1e-3) / 1e9);
        return 0;
    };
    const long long n16 = bytes / 16;
    for (int mult : {4, 8, 16}) {
        char nm[64]; snprintf(nm, 64, "grid-stride x8, %d CTAs/SM of 256", mult);
        timeit(nm, [&] { k_stride<8><<<148 * mult, 256>>>((const uint4 *)d, n16, out); });
    }
    timeit("grid-stride x4, 8 CTAs/SM of 256", [&] { k_stride<4><<<148 * 8, 256>>>((const uint4 *)d, n16, out); });
    for (int rows : {16, 28, 40, 100, 200}) {
        const int n_tiles = (int)(bytes / 512 / rows);
        char nm[64]; snprintf(nm, 64, "warp tiles LDG, %d rows", rows);
        timeit(nm, [&] { k_tile<<<(n_tiles + 3) / 4, 128>>>((const uint4 *)d, n_tiles, rows, out); });
        snprintf(nm, 64, "warp tiles TMA ring, %d rows", rows);
        timeit(nm, [&] { k_tma<<<(n_tiles + 3) / 4, 128, 32768 + 128>>>(d, n_tiles, rows, out); });
    }
    CK(cudaDeviceSynchronize());
    return 0;
}
