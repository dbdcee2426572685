// Kernel (c): KMeans exactly as make_prg calls it
//     KMeans(n_clusters=K, random_state=2, algorithm="elkan").fit(X); .predict(X)
// (make_prg/from_msa/cluster_sequences.py:262-266) with the semantics of the pinned scikit-learn 1.3.0
// (n_init=10, k-means++ seeding from one RandomState(2), Elkan iterations, best-of-10 by strict
// inertia, labels from predict).  float64 throughout, compiled with -fmad=false and written with
// explicit __dadd_rn/__dmul_rn so that every sum that scikit-learn evaluates sequentially
// (_euclidean_dense_dense in blocks of four features, centre accumulation in sample order, inertia in
// sample order, numpy's pairwise summation for the tolerance and the centre-shift total) is evaluated
// in the same order here.  The three places where scikit-learn goes through BLAS/einsum (k-means++
// candidate distances, centre-centre distances used for pruning, predict) use a plain left-to-right
// dot product; results can differ from scikit-learn there only by rounding of quantities that are
// mathematically tied ("KMeans parity").
//
// One CTA per problem; samples (or features, for the centre update) are spread over the threads.
#ifndef MPRG_HOST_EMU  // tests/hostemu compiles the device functions below as plain C++
#include <cstdio>
#include <cstdlib>

#include "common.cuh"
#include "kernels.cuh"
#endif
#include "km_layout.cuh"

namespace mprg {
// the device functions of the two compilations of this file (single-CTA, CTA-group) must not collide
#ifdef MPRG_KM_GROUP
#define KM_NS km_group
#else
#define KM_NS km_single
#endif
namespace KM_NS {

constexpr int KM_THREADS = 128;
constexpr int KM_MAXITER = 300;

#ifdef MPRG_KM_GROUP
#define c_rand c_rand_group  // the second compilation of this file keeps its own copy
#endif
__constant__ double c_rand[KM_RAND_COUNT];  // RandomState(2).random_sample stream

__device__ __forceinline__ double sqdist(const double *a, const double *b, int F) {
    // _euclidean_dense_dense (sklearn/cluster/_k_means_common.pyx): 4-way unrolled, then remainder
    double res = 0.0;
    const int m = F >> 2;
    int i = 0;
    for (int q = 0; q < m; ++q, i += 4) {
        const double d0 = __dsub_rn(a[i], b[i]), d1 = __dsub_rn(a[i + 1], b[i + 1]);
        const double d2 = __dsub_rn(a[i + 2], b[i + 2]), d3 = __dsub_rn(a[i + 3], b[i + 3]);
        const double s = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(d0, d0), __dmul_rn(d1, d1)),
                                             __dmul_rn(d2, d2)), __dmul_rn(d3, d3));
        res = __dadd_rn(res, s);
    }
    for (; i < F; ++i) {
        const double d = __dsub_rn(a[i], b[i]);
        res = __dadd_rn(res, __dmul_rn(d, d));
    }
    return res;
}

// ---- dot products exactly as numpy's matmul (OpenBLAS 0.3.30 dgemm, "TN" operand layout) rounds them
// One output element C[i, j] = sum_k a[k] * b[k] of a product with inner dimension F, `mi`/`M` the
// index/extent along the operand that OpenBLAS sees as M and `ni`/`N` along N:
//   F < 32                      left-to-right chain of fused multiply-adds
//   F >= 32 and M * N <= 1200   small-matrix kernel: eight lanes (k mod 8), each a chain of fused
//                               multiply-adds, masked tail, then a horizontal reduction whose shape
//                               depends on the 4x4 tile the element falls in
//   otherwise                   blocked driver: chains of fused multiply-adds over K blocks of at most
//                               384 (OpenBLAS level-3 split rule), block results added in order
// Checked bit-for-bit against numpy matmul over thousands of random shapes (DESIGN.md "KMeans parity");
// odd inner dimensions above 384 in the blocked regime are the one known gap.
__device__ __forceinline__ double dot_fma_range(const double *a, const double *b, int lo, int hi) {
    double acc = 0.0;
    for (int i = lo; i < hi; ++i) acc = fma(a[i], b[i], acc);
    return acc;
}

__device__ double gemm_dot(const double *a, const double *b, int F, int mi, int M, int ni, int N) {
    if (F < 32) return dot_fma_range(a, b, 0, F);
    if ((long long)M * N <= 1200) {
        double l[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        const int full = F & ~7;
        for (int k = 0; k < full; k += 8) {
#pragma unroll
            for (int q = 0; q < 8; ++q) l[q] = fma(a[k + q], b[k + q], l[q]);
        }
        for (int k = full; k < F; ++k) l[k - full] = fma(a[k], b[k], l[k - full]);
        if (mi < (M & ~3) || ni < (N & ~3)) {
            // adjacent pairs
            const double p0 = __dadd_rn(l[0], l[1]), p1 = __dadd_rn(l[2], l[3]);
            const double p2 = __dadd_rn(l[4], l[5]), p3 = __dadd_rn(l[6], l[7]);
            return __dadd_rn(__dadd_rn(p0, p1), __dadd_rn(p2, p3));
        }
        // halves (_mm512_reduce_add_pd)
        const double h0 = __dadd_rn(l[0], l[4]), h1 = __dadd_rn(l[1], l[5]);
        const double h2 = __dadd_rn(l[2], l[6]), h3 = __dadd_rn(l[3], l[7]);
        return __dadd_rn(__dadd_rn(h0, h2), __dadd_rn(h1, h3));
    }
    double tot = 0.0;
    int ls = 0;
    while (ls < F) {
        int m = F - ls;
        if (m >= 768) m = 384;
        else if (m > 384) m = ((m / 2 + 3) / 4) * 4;
        tot = __dadd_rn(tot, dot_fma_range(a, b, ls, ls + m));
        ls += m;
    }
    return tot;
}

// c0 + alpha * dot as dgemm(alpha, beta = 1) leaves it in C: identical to gemm_dot except in the
// blocked regime, where every K block is scaled and added to C separately
__device__ double gemm_axpy(const double *a, const double *b, int F, int mi, int M, int ni, int N,
                            double alpha, double c0) {
    if (F < 32 || (long long)M * N <= 1200 || F <= 384)
        return __dadd_rn(c0, __dmul_rn(alpha, gemm_dot(a, b, F, mi, M, ni, N)));
    double v = c0;
    int ls = 0;
    while (ls < F) {
        int m = F - ls;
        if (m >= 768) m = 384;
        else if (m > 384) m = ((m / 2 + 3) / 4) * 4;
        v = __dadd_rn(v, __dmul_rn(alpha, dot_fma_range(a, b, ls, ls + m)));
        ls += m;
    }
    return v;
}

// ---- OpenBLAS dgemv_t (numpy matmul with a single row / a single column operand) -------------
// Output j of nout, dot length len (x86_64 dgemv_t_4.c + Haswell/SkylakeX micro-kernel): outputs are
// produced four at a time by a 4-lane fused-multiply-add kernel; the 1-3 trailing outputs go through
// the 2-output (two lanes, separate multiply and add) and 1-output (four lanes, separate multiply and
// add) kernels; the len % 4 trailing elements are added as one expression.  Bit-identical to numpy on
// every shape tried (DESIGN.md "KMeans parity").
__device__ __forceinline__ double gemv_tail_expr(const double *a, const double *b, int m3) {
    if (m3 == 1) return __dmul_rn(a[0], b[0]);
    const double p = __dmul_rn(a[1], b[1]);
    const double s = fma(a[0], b[0], p);
    return m3 == 2 ? s : fma(a[2], b[2], s);
}

__device__ double gemv_t_dot(const double *a, const double *b, int len, int j, int nout) {
    const int m3 = len & 3, main_n = len - m3;
    if (main_n == 0) return m3 ? gemv_tail_expr(a, b, m3) : 0.0;
    const int n4 = nout & ~3, rem = nout & 3;
    int kind = 0;  // 0: four lanes fused, 1: two lanes unfused, 2: four lanes unfused
    if (j >= n4) kind = (rem >= 2 && j - n4 < 2) ? 1 : 2;
    double s;
    if (kind == 1) {
        double l0 = 0.0, l1 = 0.0;
        for (int i = 0; i < main_n; i += 2) {
            l0 = __dadd_rn(__dmul_rn(a[i], b[i]), l0);
            l1 = __dadd_rn(__dmul_rn(a[i + 1], b[i + 1]), l1);
        }
        s = __dadd_rn(l0, l1);
    } else {
        double l[4] = {0.0, 0.0, 0.0, 0.0};
        for (int i = 0; i < main_n; i += 4) {
#pragma unroll
            for (int q = 0; q < 4; ++q)
                l[q] = kind == 0 ? fma(a[i + q], b[i + q], l[q]) : __dadd_rn(__dmul_rn(a[i + q], b[i + q]), l[q]);
        }
        s = __dadd_rn(__dadd_rn(l[0], l[2]), __dadd_rn(l[1], l[3]));
    }
    if (m3 == 1) return fma(a[main_n], b[main_n], s);
    if (m3) return __dadd_rn(s, gemv_tail_expr(a + main_n, b + main_n, m3));
    return s;
}

// the same kernel with b == 1 everywhere: row sums `D @ sample_weight.reshape(-1, 1)`
__device__ double gemv_t_sum(const double *a, int len, int j, int nout) {
    const int m3 = len & 3, main_n = len - m3;
    auto tail = [&](const double *p) {
        if (m3 == 1) return p[0];
        const double s2 = __dadd_rn(p[0], p[1]);
        return m3 == 2 ? s2 : __dadd_rn(p[2], s2);
    };
    if (main_n == 0) return m3 ? tail(a) : 0.0;
    const int n4 = nout & ~3, rem = nout & 3;
    const bool two_lanes = j >= n4 && rem >= 2 && j - n4 < 2;
    double s;
    if (two_lanes) {
        double l0 = 0.0, l1 = 0.0;
        for (int i = 0; i < main_n; i += 2) {
            l0 = __dadd_rn(a[i], l0);
            l1 = __dadd_rn(a[i + 1], l1);
        }
        s = __dadd_rn(l0, l1);
    } else {
        double l[4] = {0.0, 0.0, 0.0, 0.0};
        for (int i = 0; i < main_n; i += 4) {
#pragma unroll
            for (int q = 0; q < 4; ++q) l[q] = __dadd_rn(a[i + q], l[q]);
        }
        s = __dadd_rn(__dadd_rn(l[0], l[2]), __dadd_rn(l[1], l[3]));
    }
    if (m3 == 1) return __dadd_rn(a[main_n], s);
    if (m3) return __dadd_rn(s, tail(a + main_n));
    return s;
}

// OpenBLAS ddot (SkylakeX micro-kernel) with y == 1: `closest_dist_sq @ sample_weight`
__device__ double ddot_ones(const double *x, int n) {
    const int n1 = n & -16;
    double dot = 0.0;
    if (n1) {
        double z[4][8], a[4][4];
        for (int k = 0; k < 4; ++k)
            for (int l = 0; l < 8; ++l) z[k][l] = 0.0;
        const int n32 = n1 & ~31;
        int i = 0;
        for (; i < n32; i += 32)
            for (int k = 0; k < 4; ++k)
                for (int l = 0; l < 8; ++l) z[k][l] = __dadd_rn(x[i + 8 * k + l], z[k][l]);
        for (int k = 0; k < 4; ++k)
            for (int l = 0; l < 4; ++l) a[k][l] = __dadd_rn(z[k][l], z[k][l + 4]);
        for (; i < n1; i += 16)
            for (int k = 0; k < 4; ++k)
                for (int l = 0; l < 4; ++l) a[k][l] = __dadd_rn(x[i + 4 * k + l], a[k][l]);
        double t[4];
        for (int l = 0; l < 4; ++l)
            t[l] = __dadd_rn(__dadd_rn(__dadd_rn(a[0][l], a[1][l]), a[2][l]), a[3][l]);
        dot = __dadd_rn(__dadd_rn(t[0], t[2]), __dadd_rn(t[1], t[3]));
    }
    for (int i = n1; i < n; ++i) dot = __dadd_rn(x[i], dot);
    return dot;
}

// row_norms(X, squared=True) == np.einsum("ij,ij->i", X, X): numpy's two-lane SIMD inner loop
// (einsum_sumprod.c.src, contig_contig_outstride0_two, 128-bit baseline, separate multiply and add):
// blocks of 8 elements processed back to front per lane, tail two at a time, lanes added at the end.
// Bit-identical to numpy for every length.
__device__ __forceinline__ double einsum_self(const double *a, int n) {
    double acc0 = 0.0, acc1 = 0.0;
    int i = 0;
    for (; n - i >= 8; i += 8) {
#pragma unroll
        for (int q = 3; q >= 0; --q) {
            acc0 = __dadd_rn(__dmul_rn(a[i + 2 * q], a[i + 2 * q]), acc0);
            acc1 = __dadd_rn(__dmul_rn(a[i + 2 * q + 1], a[i + 2 * q + 1]), acc1);
        }
    }
    for (; i < n; i += 2) {
        acc0 = __dadd_rn(__dmul_rn(a[i], a[i]), acc0);
        if (i + 1 < n) acc1 = __dadd_rn(__dmul_rn(a[i + 1], a[i + 1]), acc1);
    }
    return __dadd_rn(acc0, acc1);
}

// numpy's pairwise summation of a contiguous float64 vector (numpy/core/src/umath/loops_utils.h)
__device__ double np_pairwise_sum(const double *a, int n) {
    if (n < 8) {
        double res = 0.0;
        for (int i = 0; i < n; ++i) res = __dadd_rn(res, a[i]);
        return res;
    }
    if (n <= 128) {
        double r[8];
        for (int k = 0; k < 8; ++k) r[k] = a[k];
        int i = 8;
        for (; i < n - (n % 8); i += 8)
            for (int k = 0; k < 8; ++k) r[k] = __dadd_rn(r[k], a[i + k]);
        double res = __dadd_rn(__dadd_rn(__dadd_rn(r[0], r[1]), __dadd_rn(r[2], r[3])),
                               __dadd_rn(__dadd_rn(r[4], r[5]), __dadd_rn(r[6], r[7])));
        for (; i < n; ++i) res = __dadd_rn(res, a[i]);
        return res;
    }
    int n2 = n / 2;
    n2 -= n2 % 8;
    return __dadd_rn(np_pairwise_sum(a, n2), np_pairwise_sum(a + n2, n - n2));
}

// numpy pairwise summation of (x[f] - c[f])^2 over f in [lo, lo + cnt), evaluated on the fly
__device__ double np_pairwise_sqdiff(const double *x, const double *c, int lo, int cnt) {
    if (cnt < 8) {
        double r0 = 0.0;
        for (int f = 0; f < cnt; ++f) {
            const double d = __dsub_rn(x[lo + f], c[lo + f]);
            r0 = __dadd_rn(r0, __dmul_rn(d, d));
        }
        return r0;
    }
    if (cnt <= 128) {
        double r[8];
        for (int q = 0; q < 8; ++q) {
            const double d = __dsub_rn(x[lo + q], c[lo + q]);
            r[q] = __dmul_rn(d, d);
        }
        int i = 8;
        for (; i < cnt - (cnt % 8); i += 8)
            for (int q = 0; q < 8; ++q) {
                const double d = __dsub_rn(x[lo + i + q], c[lo + i + q]);
                r[q] = __dadd_rn(r[q], __dmul_rn(d, d));
            }
        double res = __dadd_rn(__dadd_rn(__dadd_rn(r[0], r[1]), __dadd_rn(r[2], r[3])),
                               __dadd_rn(__dadd_rn(r[4], r[5]), __dadd_rn(r[6], r[7])));
        for (; i < cnt; ++i) {
            const double d = __dsub_rn(x[lo + i], c[lo + i]);
            res = __dadd_rn(res, __dmul_rn(d, d));
        }
        return res;
    }
    int n2 = cnt / 2;
    n2 -= n2 % 8;
    return __dadd_rn(np_pairwise_sqdiff(x, c, lo, n2), np_pairwise_sqdiff(x, c, lo + n2, cnt - n2));
}

struct KM {
    int n, F, K;
    const double *X0;
    double *Xc, *mean, *tmpF, *xx, *ca, *cb, *best_c, *lb, *ub, *closest, *cum, *D, *half, *nxt, *shift,
        *wts, *cc, *dist, *res, *tol;
    int *labels, *labels_old, *best_labels;
    // the threads that work on this initialisation: one CTA (tid = threadIdx.x), or a group of G
    // co-resident CTAs that meet at a barrier in global memory (deep loci, kmeans_group_kernel)
    int tid, nthr, G;
    unsigned *bar;
};

__device__ __forceinline__ void km_sync(const KM &k) {
#if defined(MPRG_KM_GROUP)
    __syncthreads();
    if (k.G > 1) {
        if (threadIdx.x == 0) {
            __threadfence();
            volatile unsigned *gen = k.bar + 1;
            const unsigned g = *gen;
            if (atomicAdd(k.bar, 1u) == (unsigned)k.G - 1u) {
                *(volatile unsigned *)k.bar = 0u;
                __threadfence();
                atomicAdd(k.bar + 1, 1u);
            } else {
                while (*gen == g) __nanosleep(64);
            }
            __threadfence();
        }
        __syncthreads();
    }
#else
    __syncthreads();
#endif
}

// squared distance candidate -> sample through  -2 x.y + |x|^2 + |y|^2  clamped at 0
__device__ __forceinline__ double eucl_sq(const KM &k, int cand, int i, int t, int trials) {
    // numpy: X[candidate_ids] @ X.T  ->  M = n samples (index i), N = trials (index t); a single
    // candidate row (the first centre) goes through dgemv instead of dgemm
    const double *a = k.Xc + (long long)cand * k.F, *b = k.Xc + (long long)i * k.F;
    const double d = trials == 1 ? gemv_t_dot(a, b, k.F, i, k.n) : gemm_dot(a, b, k.F, i, k.n, t, trials);
    double v = __dadd_rn(__dadd_rn(__dmul_rn(-2.0, d), k.xx[cand]), k.xx[i]);
    return v > 0.0 ? v : 0.0;
}

__device__ void center_half_distances(const KM &k, const double *C) {
    const int K = k.K, F = k.F;
    for (int j = k.tid; j < K; j += k.nthr) k.cc[j] = einsum_self(C + (long long)j * F, F);
    km_sync(k);
    for (int p = k.tid; p < K * K; p += k.nthr) {
        const int a = p / K, b = p % K;
        double v = 0.0;
        if (a != b) {
            const double d = gemm_dot(C + (long long)a * F, C + (long long)b * F, F, b, K, a, K);
            v = __dadd_rn(__dadd_rn(__dmul_rn(-2.0, d), k.cc[a]), k.cc[b]);
            v = v > 0.0 ? v : 0.0;
            v = sqrt(v) / 2.0;
        }
        k.half[p] = v;
    }
    km_sync(k);
    // distance_next_center = np.partition(half, kth=1, axis=0)[1]: second smallest per column
    for (int j = k.tid; j < K; j += k.nthr) {
        double m1 = 1e300, m2 = 1e300;
        for (int a = 0; a < K; ++a) {
            const double v = k.half[a * K + j];
            if (v < m1) { m2 = m1; m1 = v; }
            else if (v < m2) m2 = v;
        }
        k.nxt[j] = m2;
    }
    km_sync(k);
}

// E step of _update_chunk_dense (sklearn/cluster/_k_means_elkan.pyx) for every sample
__device__ void elkan_e_step(const KM &k, const double *C) {
    const int K = k.K, F = k.F;
    for (int i = k.tid; i < k.n; i += k.nthr) {
        double u = k.ub[i];
        bool tight = false;
        int lab = k.labels[i];
        const double *x = k.Xc + (long long)i * F;
        double *lbi = k.lb + (long long)i * K;
        if (!(k.nxt[lab] >= u)) {
            for (int j = 0; j < K; ++j) {
                if (j != lab && u > lbi[j] && u > k.half[lab * K + j]) {
                    if (!tight) {
                        u = sqrt(sqdist(x, C + (long long)lab * F, F));
                        lbi[lab] = u;
                        tight = true;
                    }
                    if (u > lbi[j] || u > k.half[lab * K + j]) {
                        const double d = sqrt(sqdist(x, C + (long long)j * F, F));
                        lbi[j] = d;
                        if (d < u) {
                            lab = j;
                            u = d;
                        }
                    }
                }
            }
            k.labels[i] = lab;
            k.ub[i] = u;
        }
    }
    km_sync(k);
}

// one k-means run from k-means++ seeds; returns inertia in *out_inertia (thread 0 valid), final
// centres in *out_C
__device__ void kmeans_single(KM &k, int &rand_pos, double tol, double *s_scalar, int *s_int,
                              const double **out_C) {
    const int n = k.n, F = k.F, K = k.K;
    double *C = k.ca, *Cn = k.cb;
    const int trials = 2 + (int)log((double)K);
    // ---- k-means++ (_kmeans_plusplus, sklearn/cluster/_kmeans.py) ----
    if (k.tid == 0) {
        // random_state.choice(n, p=1/n): cdf = cumsum(p); cdf /= cdf[-1]; searchsorted(u, 'right')
        const double p = 1.0 / (double)n;
        double acc = 0.0;
        for (int i = 0; i < n; ++i) {
            acc = __dadd_rn(acc, p);
            k.cum[i] = acc;
        }
        const double total = k.cum[n - 1], u = c_rand[rand_pos];
        int idx = 0;
        while (idx < n && k.cum[idx] / total <= u) ++idx;
        s_int[0] = idx < n ? idx : n - 1;
    }
    rand_pos += 1;
    km_sync(k);
    const int c0 = s_int[0];
#ifdef MPRG_HOST_EMU_DEBUG
    printf("  kpp first %d (u=%.17g)\n", c0, c_rand[rand_pos - 1]);
#endif
    for (int f = k.tid; f < F; f += k.nthr) C[f] = k.Xc[(long long)c0 * F + f];
    for (int i = k.tid; i < n; i += k.nthr) k.closest[i] = eucl_sq(k, c0, i, 0, 1);
    km_sync(k);
    if (k.tid == 0) {
        s_scalar[0] = ddot_ones(k.closest, n);
    }
    km_sync(k);
    for (int c = 1; c < K; ++c) {
        if (k.tid == 0) {
            const double pot = s_scalar[0];
            double acc = 0.0;
            for (int i = 0; i < n; ++i) {
                acc = __dadd_rn(acc, k.closest[i]);
                k.cum[i] = acc;
            }
            for (int t = 0; t < trials; ++t) {
                const double val = __dmul_rn(c_rand[rand_pos + t], pot);
                int lo = 0, hi = n;  // first index with cum[idx] >= val  (searchsorted 'left')
                while (lo < hi) {
                    const int mid = (lo + hi) >> 1;
                    if (k.cum[mid] < val) lo = mid + 1; else hi = mid;
                }
                s_int[1 + t] = lo < n - 1 ? lo : n - 1;
            }
        }
        rand_pos += trials;
        km_sync(k);
        for (int p = k.tid; p < trials * n; p += k.nthr) {
            const int t = p / n, i = p % n;
            const double d = eucl_sq(k, s_int[1 + t], i, t, trials);
            const double cl = k.closest[i];
            k.D[p] = cl < d ? cl : d;
        }
        km_sync(k);
        if (k.tid == 0) {
            int best = 0;
            double best_pot = 0.0;
            for (int t = 0; t < trials; ++t) {
                const double s = gemv_t_sum(k.D + (long long)t * n, n, t, trials);
                if (t == 0 || s < best_pot) {
                    best = t;
                    best_pot = s;
                }
            }
            s_scalar[0] = best_pot;
            s_int[0] = best;
#ifdef MPRG_HOST_EMU_DEBUG
            printf("  kpp c=%d cands", c);
            for (int t = 0; t < trials; ++t) printf(" %d", s_int[1 + t]);
            printf(" best %d pot %.17g\n", best, best_pot);
#endif
        }
        km_sync(k);
        const int best = s_int[0], cand = s_int[1 + best];
        for (int i = k.tid; i < n; i += k.nthr) k.closest[i] = k.D[best * n + i];
        for (int f = k.tid; f < F; f += k.nthr) C[(long long)c * F + f] = k.Xc[(long long)cand * F + f];
        km_sync(k);
    }

    // ---- Elkan (_kmeans_single_elkan) ----
    for (int i = k.tid; i < n; i += k.nthr) {
        k.labels[i] = -1;
        k.labels_old[i] = -1;
        k.ub[i] = 0.0;
    }
    for (long long p = k.tid; p < (long long)n * K; p += k.nthr) k.lb[p] = 0.0;
    for (long long p = k.tid; p < (long long)K * F; p += k.nthr) Cn[p] = 0.0;
    km_sync(k);
    center_half_distances(k, C);
    // init_bounds_dense
    for (int i = k.tid; i < n; i += k.nthr) {
        const double *x = k.Xc + (long long)i * F;
        int best = 0;
        double md = sqrt(sqdist(x, C, F));
        k.lb[(long long)i * K] = md;
        for (int j = 1; j < K; ++j) {
            if (md > k.half[best * K + j]) {
                const double d = sqrt(sqdist(x, C + (long long)j * F, F));
                k.lb[(long long)i * K + j] = d;
                if (d < md) {
                    md = d;
                    best = j;
                }
            }
        }
        k.labels[i] = best;
        k.ub[i] = md;
    }
    km_sync(k);
    bool strict = false;
    for (int it = 0; it < KM_MAXITER; ++it) {
        elkan_e_step(k, C);
        // M step: member sums in sample order (thread per feature), weights
        for (int f = k.tid; f < F; f += k.nthr) {
            double acc[KM_MAXK];
#pragma unroll
            for (int j = 0; j < KM_MAXK; ++j) acc[j] = 0.0;
            for (int i = 0; i < n; ++i) {
                const int lab = k.labels[i];
                const double x = k.Xc[(long long)i * F + f];
#pragma unroll
                for (int j = 0; j < KM_MAXK; ++j)
                    if (j == lab) acc[j] = __dadd_rn(acc[j], x);
            }
#pragma unroll
            for (int j = 0; j < KM_MAXK; ++j)
                if (j < K) Cn[(long long)j * F + f] = acc[j];
        }
        if (k.tid == 0) {
            for (int j = 0; j < K; ++j) k.wts[j] = 0.0;
            for (int i = 0; i < n; ++i) k.wts[k.labels[i]] += 1.0;
            int emask = 0;
            for (int j = 0; j < K; ++j)
                if (k.wts[j] == 0.0) emask |= 1 << j;
            s_int[0] = emask;
        }
        km_sync(k);
        const int emask = s_int[0];
        if (emask != 0) {
            // _relocate_empty_clusters_dense: farthest samples re-seed the empty clusters
            for (int i = k.tid; i < n; i += k.nthr) {
                const double *x = k.Xc + (long long)i * F;
                const double *c = C + (long long)k.labels[i] * F;
                // ((X - C[labels])**2).sum(axis=1): numpy pairwise summation over the features
                const double res = np_pairwise_sqdiff(x, c, 0, F);
                k.dist[i] = res;
            }
            km_sync(k);
            if (k.tid == 0) {
                double mx = 0.0;
                for (int i = 0; i < n; ++i) mx = k.dist[i] > mx ? k.dist[i] : mx;
                s_int[1] = mx != 0.0;
            }
            km_sync(k);
            if (s_int[1]) {
                // one empty cluster at a time (ascending id), farthest remaining sample first
                for (int e = 0; e < K; ++e) {
                    if (!((emask >> e) & 1)) continue;  // the empty set is fixed before relocating
                    if (k.tid == 0) {
                        int far = 0;
                        double best = -1.0;
                        for (int i = 0; i < n; ++i)
                            if (k.dist[i] > best) {
                                best = k.dist[i];
                                far = i;
                            }
                        k.dist[far] = -2.0;  // taken
                        s_int[2] = far;
                    }
                    km_sync(k);
                    const int far = s_int[2], old = k.labels[far];
                    for (int f = k.tid; f < F; f += k.nthr) {
                        const double x = k.Xc[(long long)far * F + f];
                        Cn[(long long)old * F + f] = __dsub_rn(Cn[(long long)old * F + f], x);
                        Cn[(long long)e * F + f] = x;
                    }
                    km_sync(k);
                    if (k.tid == 0) {
                        k.wts[e] = 1.0;
                        k.wts[old] -= 1.0;
                    }
                    km_sync(k);
                }
            }
        }
        // _average_centers (thread per feature, clusters in ascending order as the reference loop)
        if (k.tid == 0) {
            int amax = 0;
            for (int j = 1; j < K; ++j)
                if (k.wts[j] > k.wts[amax]) amax = j;
            s_int[5] = amax;  // its own slot: slot 0 (emask) may still be unread by slower threads
        }
        km_sync(k);
        {
            const int amax = s_int[5];
            for (int f = k.tid; f < F; f += k.nthr) {
                for (int j = 0; j < K; ++j) {
                    if (k.wts[j] > 0.0) {
                        const double alpha = 1.0 / k.wts[j];
                        Cn[(long long)j * F + f] = __dmul_rn(Cn[(long long)j * F + f], alpha);
                    } else {
                        Cn[(long long)j * F + f] = Cn[(long long)amax * F + f];
                    }
                }
            }
        }
        km_sync(k);
        // _center_shift, bounds update
        for (int j = k.tid; j < K; j += k.nthr)
            k.shift[j] = sqrt(sqdist(Cn + (long long)j * F, C + (long long)j * F, F));
        km_sync(k);
        for (int i = k.tid; i < n; i += k.nthr) {
            k.ub[i] = __dadd_rn(k.ub[i], k.shift[k.labels[i]]);
            for (int j = 0; j < K; ++j) {
                double v = __dsub_rn(k.lb[(long long)i * K + j], k.shift[j]);
                k.lb[(long long)i * K + j] = v < 0.0 ? 0.0 : v;
            }
        }
        km_sync(k);
        center_half_distances(k, Cn);
        {
            double *t = C;
            C = Cn;
            Cn = t;
        }
        // convergence
        if (k.tid == 0) {
            bool same = true;
            for (int i = 0; i < n && same; ++i) same = k.labels[i] == k.labels_old[i];
            int stop = 0;
            if (same) stop = 1;
            else {
                double sq[KM_MAXK];
                for (int j = 0; j < K; ++j) sq[j] = __dmul_rn(k.shift[j], k.shift[j]);
                if (np_pairwise_sum(sq, K) <= tol) stop = 2;
            }
            s_int[6] = stop;
        }
        km_sync(k);
        const int stop = s_int[6];
        if (stop == 1) {
            strict = true;
            break;
        }
        if (stop == 2) break;
        for (int i = k.tid; i < n; i += k.nthr) k.labels_old[i] = k.labels[i];
        km_sync(k);
    }
    if (!strict) elkan_e_step(k, C);
    // _inertia_dense: sequential over samples
    for (int i = k.tid; i < n; i += k.nthr)
        k.dist[i] = sqdist(k.Xc + (long long)i * F, C + (long long)k.labels[i] * F, F);
    km_sync(k);
    if (k.tid == 0) {
        double inertia = 0.0;
        for (int i = 0; i < n; ++i) inertia = __dadd_rn(inertia, k.dist[i]);
        s_scalar[1] = inertia;
    }
    km_sync(k);
    *out_C = C;
}

// ---- one fit = preparation + KM_NINIT independent initialisations + a best-of selection ----------
// Scratch of one problem: the shared block (centred copy of X, column means, squared row norms,
// tolerance: written once by kmeans_prepare, read-only afterwards, for every K the clustering loop
// tries), then KM_NINIT per-initialisation blocks (so the initialisations can run in different CTAs),
// then the selection block.
// (km_shared_doubles / km_init_doubles / km_init_ints: kernels.cuh, shared with the device-side problem builder)

// d0 = scratch of the problem, init = which initialisation block to bind
__device__ void km_bind_init(KM &k, int n, int F, int K, const double *X0, double *d0, int *i0, int init) {
    k.n = n; k.F = F; k.K = K; k.X0 = X0;
    double *d = d0;
    k.Xc = d; d += (long long)n * F;
    k.mean = d; d += F;
    k.tmpF = d; d += F;
    k.xx = d; d += n;
    k.tol = d;
    d = d0 + km_shared_doubles(n, F) + init * km_init_doubles(n, F);
    k.ca = d; d += (long long)KM_MAXK * F;
    k.cb = d; d += (long long)KM_MAXK * F;
    k.lb = d; d += (long long)n * KM_MAXK;
    k.ub = d; d += n;
    k.closest = d; d += n;
    k.cum = d; d += n;
    k.D = d; d += 4LL * n;
    k.dist = d; d += n;
    k.half = d; d += KM_MAXK * KM_MAXK;
    k.nxt = d; d += KM_MAXK;
    k.shift = d; d += KM_MAXK;
    k.wts = d; d += KM_MAXK;
    k.cc = d; d += KM_MAXK;
    k.res = d;  // res[0] = inertia, res[1] = 0/1: final centres in ca/cb
    int *ii = i0 + init * km_init_ints(n);
    k.labels = ii; ii += n;
    k.labels_old = ii; ii += n;
    k.best_c = nullptr;
    k.best_labels = nullptr;
    k.tid = threadIdx.x;
    k.nthr = blockDim.x;
    k.G = 1;
    k.bar = nullptr;
}

// once per problem: tolerance on the un-centred data, column means, centring, squared row norms
// (KMeans._tolerance, X -= X.mean(axis=0), row_norms(X, squared=True) in sklearn/cluster/_kmeans.py)
__device__ void kmeans_prepare(KM &k) {
    const int n = k.n, F = k.F;
    for (int f = k.tid; f < F; f += k.nthr) {
        double s = 0.0;
        for (int i = 0; i < n; ++i) s = __dadd_rn(s, k.X0[(long long)i * F + f]);
        const double m = s / (double)n;
        double v = 0.0;
        for (int i = 0; i < n; ++i) {
            const double d = __dsub_rn(k.X0[(long long)i * F + f], m);
            v = __dadd_rn(v, __dmul_rn(d, d));
            k.Xc[(long long)i * F + f] = d;
        }
        k.mean[f] = m;
        k.tmpF[f] = v / (double)n;
    }
    km_sync(k);
    if (k.tid == 0) k.tol[0] = __dmul_rn(np_pairwise_sum(k.tmpF, F) / (double)F, 1e-4);
    for (int i = k.tid; i < n; i += k.nthr) k.xx[i] = einsum_self(k.Xc + (long long)i * F, F);
    km_sync(k);
}

// one initialisation: k-means++ from the init-th slice of the RandomState(2) stream, Elkan; leaves
// labels, inertia and the final centres in the init's block
__device__ void kmeans_run_init(KM &k, int init, double *s_scalar, int *s_int) {
    const int K = k.K;
    const double tol = k.tol[0];
    // every initialisation consumes 1 + (K-1) * (2 + int(ln K)) doubles of the shared stream
    const int trials = 2 + (int)log((double)K);
    int rand_pos = init * (1 + (K - 1) * trials);
    const double *C = nullptr;
    kmeans_single(k, rand_pos, tol, s_scalar, s_int, &C);
    if (k.tid == 0) {
        k.res[0] = s_scalar[1];
        k.res[1] = (C == k.ca) ? 0.0 : 1.0;
    }
    km_sync(k);
}

// best-of-n_init selection (strict inertia improvement and not the same clustering), then predict on
// the un-centred data: argmin_j |c_j|^2 - 2 x.c_j  (lloyd _update_chunk_dense, chunks of 256 samples)
__device__ void kmeans_select_predict(int n, int F, int K, const double *X0, double *d0, int *i0,
                                      int *s_int, int *out_labels, double *out_inertia) {
    double *best_c = d0 + km_shared_doubles(n, F) + KM_NINIT * km_init_doubles(n, F);
    double *cc = best_c + (long long)KM_MAXK * F;
    if (threadIdx.x == 0) {
        int best = 0;
        for (int init = 1; init < KM_NINIT; ++init) {
            KM a, b;
            km_bind_init(a, n, F, K, X0, d0, i0, init);
            km_bind_init(b, n, F, K, X0, d0, i0, best);
            if (a.res[0] < b.res[0]) {
                // _is_same_clustering(labels, best_labels, K)
                int mapping[KM_MAXK];
                for (int j = 0; j < KM_MAXK; ++j) mapping[j] = -1;
                bool same = true;
                for (int i = 0; i < n && same; ++i) {
                    const int la = a.labels[i], lb = b.labels[i];
                    if (mapping[la] == -1) mapping[la] = lb;
                    else if (mapping[la] != lb) same = false;
                }
                if (!same) best = init;
            }
        }
        s_int[0] = best;
    }
    __syncthreads();
    KM w;
    km_bind_init(w, n, F, K, X0, d0, i0, s_int[0]);
    const double *C = (w.res[1] == 0.0) ? w.ca : w.cb;
    for (long long p = threadIdx.x; p < (long long)K * F; p += blockDim.x)
        best_c[p] = __dadd_rn(C[p], w.mean[p % F]);
    __syncthreads();
    for (int j = threadIdx.x; j < K; j += blockDim.x) cc[j] = einsum_self(best_c + (long long)j * F, F);
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const double *x = X0 + (long long)i * F;
        int lab = 0;
        double best = 0.0;
        const int chunk = n > 256 ? 256 : n;
        const int ci = i % chunk, cn = (i / chunk) * chunk + chunk <= n ? chunk : n % chunk;
        for (int j = 0; j < K; ++j) {
            const double v = gemm_axpy(x, best_c + (long long)j * F, F, j, K, ci, cn, -2.0, cc[j]);
            if (j == 0 || v < best) {
                best = v;
                lab = j;
            }
        }
        out_labels[i] = lab;
    }
    if (threadIdx.x == 0 && out_inertia) *out_inertia = w.res[0];
    __syncthreads();
}

}  // namespace KM_NS
using namespace KM_NS;

#ifndef MPRG_KM_GROUP
// scratch per problem: shared block | KM_NINIT init blocks | best_c [KM_MAXK * F] | cc [KM_MAXK] ;
// ints: KM_NINIT init blocks
long long kmeans_dscratch_doubles(long long n, long long F) { return km_dscratch_doubles(n, F); }
long long kmeans_iscratch_ints(long long n) { return km_iscratch_ints(n); }
#endif

#if !defined(MPRG_HOST_EMU) && !defined(MPRG_KM_GROUP)
// one CTA per problem that is about to run KMeans for the first time (K == 2 round)
__global__ void __launch_bounds__(KM_THREADS)
kmeans_prepare_kernel(const ClusterState *__restrict__ states, const double *__restrict__ X_all,
                      double *__restrict__ dscratch, int *__restrict__ iscratch) {
    const ClusterState &st = states[blockIdx.x];
    if (st.status != 0 || !st.run_kmeans || st.big) return;
    KM k;
    km_bind_init(k, st.n, st.F, st.K, X_all + st.x_off, dscratch + st.kmd_off, iscratch + st.kmi_off, 0);
    kmeans_prepare(k);
}

// grid (problems, KM_NINIT): every CTA runs one initialisation; the last CTA of a problem to finish
// (ticket counter) does the selection, the prediction and the loop control of kmeans_cluster_seqs
// Scratch of ONE initialisation in shared memory (doubles then ints), laid out for the problem's own K: the
// tiny problems of a pangenome level (n <= 8, F <= 70, K = 2..3) spend their time in sequential float64 sections
// on one lane, and every store-then-load of the global scratch was an L2 round trip.
__device__ __forceinline__ long long km_smem_bytes(int n, int F, int K) {
    const long long d = 2LL * K * F + (long long)n * K + 8LL * n + (long long)K * K + 4LL * K + 8;
    return 8 * d + 4 * 2LL * n;
}
__device__ __forceinline__ void km_bind_smem(KM &k, double *sm) {
    const int n = k.n, F = k.F, K = k.K;
    double *d = sm;
    k.ca = d; d += (long long)K * F;
    k.cb = d; d += (long long)K * F;
    k.lb = d; d += (long long)n * K;
    k.ub = d; d += n;
    k.closest = d; d += n;
    k.cum = d; d += n;
    k.D = d; d += 4LL * n;
    k.dist = d; d += n;
    k.half = d; d += K * K;
    k.nxt = d; d += K;
    k.shift = d; d += K;
    k.wts = d; d += K;
    k.cc = d; d += K;
    k.res = d; d += 8;
    int *ii = reinterpret_cast<int *>(d);
    k.labels = ii; ii += n;
    k.labels_old = ii;
}

extern __shared__ double km_dyn_smem[];

__device__ __forceinline__ void
kmeans_kernel_body(ClusterState *__restrict__ states, const double *__restrict__ X_all,
                   double *__restrict__ dscratch, int *__restrict__ iscratch, int *__restrict__ assign_all,
                   int *__restrict__ newlab_all, int *__restrict__ tickets, int smem_bytes) {
    __shared__ double s_scalar[4];
    __shared__ int s_int[8];
    ClusterState &st = states[blockIdx.x];
    if (st.status != 0 || !st.run_kmeans || st.big) return;
    const int n = st.n, F = st.F, K = st.K, init = blockIdx.y;
    const double *X0 = X_all + st.x_off;
    double *d0 = dscratch + st.kmd_off;
    int *i0 = iscratch + st.kmi_off;
    KM k;
    km_bind_init(k, n, F, K, X0, d0, i0, init);
    const bool in_smem = km_smem_bytes(n, F, K) <= (long long)smem_bytes;
    KM g = k;  // where the selection step (any CTA of the problem) reads this initialisation's results
    if (in_smem) km_bind_smem(k, km_dyn_smem);
    kmeans_run_init(k, init, s_scalar, s_int);
    if (in_smem) {
        // results out: inertia + centre selector, labels, final centres
        const double *Cs = (k.res[1] == 0.0) ? k.ca : k.cb;
        double *Cg = (k.res[1] == 0.0) ? g.ca : g.cb;
        for (int p = threadIdx.x; p < K * F; p += blockDim.x) Cg[p] = Cs[p];
        for (int i = threadIdx.x; i < n; i += blockDim.x) g.labels[i] = k.labels[i];
        if (threadIdx.x == 0) {
            g.res[0] = k.res[0];
            g.res[1] = k.res[1];
        }
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_int[7] = atomicAdd(&tickets[blockIdx.x], 1);
    __syncthreads();
    if (s_int[7] != KM_NINIT - 1) return;
    __threadfence();
    int *newlab = newlab_all + st.assign_off;
    int *assign = assign_all + st.assign_off;
    kmeans_select_predict(n, F, K, X0, d0, i0, s_int, newlab, nullptr);
    if (threadIdx.x == 0) {
        // cluster_sequences.py:267-274: fewer distinct labels than K => keep the previous assignment
        unsigned seen = 0;
        for (int i = 0; i < n; ++i) seen |= 1u << newlab[i];
        const int distinct = __popc(seen);
        if (distinct < K) {
            st.K -= 1;
            st.status = 1;
        } else {
            for (int i = 0; i < n; ++i) assign[i] = newlab[i];
        }
        tickets[blockIdx.x] = 0;
        __threadfence();
        st.run_kmeans = 0;
    }
}

__global__ void __launch_bounds__(KM_THREADS)
kmeans_kernel(ClusterState *__restrict__ states, const double *__restrict__ X_all,
              double *__restrict__ dscratch, int *__restrict__ iscratch, int *__restrict__ assign_all,
              int *__restrict__ newlab_all, int *__restrict__ tickets, int smem_bytes) {
    kmeans_kernel_body(states, X_all, dscratch, iscratch, assign_all, newlab_all, tickets, smem_bytes);
}

// The same for the one-warp CTAs of a pangenome level (thousands of problems with n <= 8, F <= 70): the level
// is bound by latency (sequential float64 sums on one lane while the others wait), so what counts is how many
// initialisations are resident.  At 128 registers an SM holds 16 one-warp CTAs; bounded to 64 registers
// (spills stay in L1) it holds the 32 the CTA slots allow.
__global__ void __launch_bounds__(32, 32)
kmeans_kernel_w32(ClusterState *__restrict__ states, const double *__restrict__ X_all,
                  double *__restrict__ dscratch, int *__restrict__ iscratch, int *__restrict__ assign_all,
                  int *__restrict__ newlab_all, int *__restrict__ tickets, int smem_bytes) {
    kmeans_kernel_body(states, X_all, dscratch, iscratch, assign_all, newlab_all, tickets, smem_bytes);
}

// stand-alone problem (mprg_kmeans): prepare <<<1>>> then grid (1, KM_NINIT)
__global__ void __launch_bounds__(KM_THREADS)
kmeans_single_prepare_kernel(const double *X0, int n, int F, int K, double *dscratch, int *iscratch) {
    KM k;
    km_bind_init(k, n, F, K, X0, dscratch, iscratch, 0);
    kmeans_prepare(k);
}
__global__ void __launch_bounds__(KM_THREADS)
kmeans_single_problem_kernel(const double *X0, int n, int F, int K, double *dscratch, int *iscratch,
                             int *labels, double *inertia, int *ticket) {
    __shared__ double s_scalar[4];
    __shared__ int s_int[8];
    const int init = blockIdx.y;
    KM k;
    km_bind_init(k, n, F, K, X0, dscratch, iscratch, init);
    kmeans_run_init(k, init, s_scalar, s_int);
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_int[7] = atomicAdd(ticket, 1);
    __syncthreads();
    if (s_int[7] != KM_NINIT - 1) return;
    __threadfence();
    kmeans_select_predict(n, F, K, X0, dscratch, iscratch, s_int, labels, inertia);
    if (threadIdx.x == 0) *ticket = 0;
}
#elif defined(MPRG_KM_GROUP)
// ---- deep loci: one problem on the whole GPU ----------------------------------------------------------
// Every initialisation runs on a group of G co-resident CTAs (cooperative launch): the loops over
// samples and over features are spread over the G * KM_THREADS threads of the group, the sequential
// parts stay with its first thread, and __syncthreads becomes a barrier in global memory (km_sync).
// Every dot product, sum and comparison is the one the single-CTA kernel does, in the same order, so
// labels and inertia are bit-identical.  This object is compiled with -dlcm=cg: data written by one
// CTA of the group is read by the others through L2, never from a stale L1 line.
// bars (8-byte aligned, 32 + 16 * KM_NINIT zero-initialised words): [2 * init] barrier of initialisation
// init | [2 * KM_NINIT] barrier of the preparation | [2 * KM_NINIT + 2] ticket | [32 + 16 * init]
// broadcast slots of initialisation init (4 doubles, 8 ints)
__global__ void __launch_bounds__(KM_THREADS)
kmeans_group_prepare_kernel(const ClusterState *__restrict__ states, int q, const double *__restrict__ X_all,
                            double *__restrict__ dscratch, int *__restrict__ iscratch, unsigned *bars) {
    const ClusterState &st = states[q];
    if (st.status != 0 || !st.run_kmeans) return;
    KM k;
    km_bind_init(k, st.n, st.F, st.K, X_all + st.x_off, dscratch + st.kmd_off, iscratch + st.kmi_off, 0);
    k.tid = blockIdx.x * blockDim.x + threadIdx.x;
    k.nthr = gridDim.x * blockDim.x;
    k.G = gridDim.x;
    k.bar = bars + 2 * KM_NINIT;
    kmeans_prepare(k);
}

// (measured: bounding the kernel to 80 registers for 6 CTAs per SM and groups of 88 CTAs is no faster -- config #4
// 5.23 s against 5.11 s -- the group barriers and the sequential sections, not the resident threads, bound it)
__global__ void __launch_bounds__(KM_THREADS)
kmeans_group_kernel(ClusterState *__restrict__ states, int q, const double *__restrict__ X_all,
                    double *__restrict__ dscratch, int *__restrict__ iscratch, int *__restrict__ assign_all,
                    int *__restrict__ newlab_all, unsigned *bars, int G, double *inertia_out) {
    __shared__ int s_last;
    __shared__ int s_sel[8];
    ClusterState &st = states[q];
    if (st.status != 0 || !st.run_kmeans) return;
    const int n = st.n, F = st.F, K = st.K;
    const int init = blockIdx.x / G, rank = blockIdx.x % G;
    const double *X0 = X_all + st.x_off;
    double *d0 = dscratch + st.kmd_off;
    int *i0 = iscratch + st.kmi_off;
    KM k;
    km_bind_init(k, n, F, K, X0, d0, i0, init);
    k.tid = rank * blockDim.x + threadIdx.x;
    k.nthr = G * blockDim.x;
    k.G = G;
    k.bar = bars + 2 * init;
    // the broadcast slots of kmeans_single (thread 0 of the group writes, everybody reads after the
    // barrier) live in the communication area next to the barriers: 16 words per initialisation
    unsigned *slots = bars + 32 + 16 * init;
    kmeans_run_init(k, init, reinterpret_cast<double *>(slots), reinterpret_cast<int *>(slots + 8));
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = atomicAdd(&bars[2 * KM_NINIT + 2], 1u) == (unsigned)(KM_NINIT * G - 1);
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    int *newlab = newlab_all + st.assign_off;
    int *assign = assign_all + st.assign_off;
    kmeans_select_predict(n, F, K, X0, d0, i0, s_sel, newlab, inertia_out);
    if (threadIdx.x == 0) {
        // cluster_sequences.py:267-274: fewer distinct labels than K => keep the previous assignment
        unsigned seen = 0;
        for (int i = 0; i < n; ++i) seen |= 1u << newlab[i];
        const int distinct = __popc(seen);
        if (distinct < K) {
            st.K -= 1;
            st.status = 1;
        } else {
            for (int i = 0; i < n; ++i) assign[i] = newlab[i];
        }
        bars[2 * KM_NINIT + 2] = 0u;
        __threadfence();
        st.run_kmeans = 0;
    }
}
#else
// host emulation (tests/hostemu): the same device functions, initialisations one after the other
void kmeans_single_problem_kernel(const double *X0, int n, int F, int K, double *dscratch, int *iscratch,
                                  int *labels, double *inertia) {
    static double s_scalar[4];
    static int s_int[8];
    KM k;
    km_bind_init(k, n, F, K, X0, dscratch, iscratch, 0);
    kmeans_prepare(k);
    for (int init = 0; init < KM_NINIT; ++init) {
        km_bind_init(k, n, F, K, X0, dscratch, iscratch, init);
        kmeans_run_init(k, init, s_scalar, s_int);
    }
    kmeans_select_predict(n, F, K, X0, dscratch, iscratch, s_int, labels, inertia);
}
#endif

#if defined(MPRG_KM_GROUP)
cudaError_t kmeans_group_upload_rand(const double *h_rand) {
    return cudaMemcpyToSymbol(c_rand, h_rand, sizeof(double) * KM_RAND_COUNT);
}

cudaError_t launch_kmeans_group(cudaStream_t s, ClusterState *states, int q, const double *X, double *dscratch,
                                int *iscratch, int *assign, int *newlab, unsigned *bars, bool prepare, int sm_count,
                                double *inertia_out) {
    static int occ = 0, occ_prep = 0;
    if (!occ) {
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kmeans_group_kernel, KM_THREADS, 0);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_prep, kmeans_group_prepare_kernel, KM_THREADS, 0);
        occ = occ > 0 ? occ : 1;
        occ_prep = occ_prep > 0 ? occ_prep : 1;
    }
    cudaError_t e;
    if (prepare) {
        int gp = std::max(1, std::min(sm_count * occ_prep, sm_count));
        if (const char *env = getenv("MPRG_KM_G")) gp = std::max(1, std::min(atoi(env), gp));  // debugging knob
        const dim3 grid((unsigned)gp);
        void *args[] = {(void *)&states, (void *)&q, (void *)&X, (void *)&dscratch, (void *)&iscratch, (void *)&bars};
        e = cudaLaunchCooperativeKernel((const void *)kmeans_group_prepare_kernel, grid, dim3(KM_THREADS), args, 0, s);
        if (e != cudaSuccess) return e;
        if (getenv("MPRG_DEBUG_SYNC")) {
            e = cudaStreamSynchronize(s);
            fprintf(stderr, "[mprg debug] group prepare grid %u: %s\n", grid.x, cudaGetErrorString(e));
            if (e != cudaSuccess) return e;
        }
    }
    // as many CTAs per initialisation as stay co-resident (B200 at 136 registers: 148 x 3 / 10 = 44)
    int G = std::max(1, std::min((sm_count * occ) / KM_NINIT, 64));
    if (const char *env = getenv("MPRG_KM_G")) G = std::max(1, std::min(atoi(env), G));  // debugging knob
    const dim3 grid((unsigned)(KM_NINIT * G));
    void *args[] = {(void *)&states, (void *)&q, (void *)&X, (void *)&dscratch, (void *)&iscratch,
                    (void *)&assign, (void *)&newlab, (void *)&bars, (void *)&G, (void *)&inertia_out};
    e = cudaLaunchCooperativeKernel((const void *)kmeans_group_kernel, grid, dim3(KM_THREADS), args, 0, s);
    if (e == cudaSuccess && getenv("MPRG_DEBUG_SYNC")) {
        e = cudaStreamSynchronize(s);
        fprintf(stderr, "[mprg debug] group kernel grid %u (G %d): %s\n", grid.x, G, cudaGetErrorString(e));
    }
    return e;
}
#elif !defined(MPRG_HOST_EMU)
cudaError_t kmeans_upload_rand(const double *h_rand) {
    const cudaError_t e = kmeans_group_upload_rand(h_rand);
    if (e != cudaSuccess) return e;
    return cudaMemcpyToSymbol(c_rand, h_rand, sizeof(double) * KM_RAND_COUNT);
}

// max_elements: the largest n * F of the level.  The loops of an initialisation stride over whatever
// threads it has (the results do not depend on their number: test_kmeans_cta_groups_equal_single_cta),
// so the many tiny problems of a pangenome level (n <= 8, F <= 70) get one warp per CTA: four times as
// many initialisations resident per SM at 128 registers per thread, cheaper barriers.
cudaError_t launch_kmeans(cudaStream_t s, ClusterState *states, int n_probs, const double *X,
                          double *dscratch, int *iscratch, int *assign, int *newlab, int *tickets,
                          long long max_elements) {
    if (n_probs <= 0) return cudaSuccess;
    // (max_elements is counted with the bound F <= positions when the level does not wait for F: 8 x 241)
    int threads = max_elements <= 2048 ? 32 : (max_elements <= 16384 ? 64 : KM_THREADS);
    if (const char *e = getenv("MPRG_KM_THREADS")) threads = std::max(32, std::min(atoi(e) & ~31, KM_THREADS));
    static const bool w32 = getenv("MPRG_KM_NO_W32") == nullptr;
    // shared-memory scratch per initialisation (problems that do not fit keep the global scratch): 4 KB keep 32
    // one-warp CTAs per SM, 16 KB are free for the few big CTAs; MPRG_KM_SMEM overrides (0 = global scratch only)
    static const int smem_env = getenv("MPRG_KM_SMEM") ? atoi(getenv("MPRG_KM_SMEM")) : -1;
    if (threads == 32 && w32) {
        const int smem = smem_env >= 0 ? smem_env : 4096;
        kmeans_kernel_w32<<<dim3(n_probs, KM_NINIT), 32, smem, s>>>(states, X, dscratch, iscratch, assign, newlab, tickets,
                                                                   smem);
    } else {
        const int smem = smem_env >= 0 ? smem_env : 16384;
        kmeans_kernel<<<dim3(n_probs, KM_NINIT), threads, smem, s>>>(states, X, dscratch, iscratch, assign, newlab,
                                                                     tickets, smem);
    }
    return cudaGetLastError();
}

cudaError_t launch_kmeans_prepare(cudaStream_t s, const ClusterState *states, int n_probs, const double *X,
                                  double *dscratch, int *iscratch) {
    if (n_probs <= 0) return cudaSuccess;
    kmeans_prepare_kernel<<<n_probs, KM_THREADS, 0, s>>>(states, X, dscratch, iscratch);
    return cudaGetLastError();
}

cudaError_t launch_kmeans_single(cudaStream_t s, const double *X0, int n, int F, int K, double *dscratch,
                                 int *iscratch, int *labels, double *inertia, int *ticket) {
    kmeans_single_prepare_kernel<<<1, KM_THREADS, 0, s>>>(X0, n, F, K, dscratch, iscratch);
    kmeans_single_problem_kernel<<<dim3(1, KM_NINIT), KM_THREADS, 0, s>>>(X0, n, F, K, dscratch, iscratch, labels,
                                                                          inertia, ticket);
    return cudaGetLastError();
}

#endif  // MPRG_HOST_EMU

}  // namespace mprg
