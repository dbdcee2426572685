"""CPU restatement of scikit-learn 1.3.0 `KMeans(n_clusters=K, random_state=2, algorithm="elkan")`
`.fit(X)` followed by `.predict(X)` exactly as make_prg calls it
(reference call site: make_prg/from_msa/cluster_sequences.py:262-266).

TEST INFRASTRUCTURE ONLY (oracle).  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference leg may import this file; the product never does.

The arithmetic lives in a third-party dependency that is absent from /root/reference:
scikit-learn pinned at 1.3.0 (poetry.lock:1115-1116), files sklearn/cluster/_kmeans.py,
_k_means_elkan.pyx, _k_means_common.pyx, _k_means_lloyd.pyx.  In 1.3.0 the default is n_init=10
(it became 1 for k-means++ in >= 1.4); the reference's golden outputs need n_init=10 and
sequential (single OpenMP thread) reductions.

Pinned: tests/test_oracle_kmeans.py compares this restatement against the installed scikit-learn
(forced to n_init=10, one OpenMP thread) for labels and bit-identical inertia on tie-prone integer
matrices and on every count matrix the reference feeds to KMeans on its own fixtures
(tests/golden/kmeans_cases.npz).

Operation order that matters (published algorithm, restated):
  * tol = mean(var(X, axis=0)) * 1e-4 on the un-centred X  (_kmeans.py `_tolerance`, called from
    `_check_params_vs_input` before the mean is subtracted)
  * X <- X - mean(X, axis=0); xx = row_norms(X)^2
  * ONE RandomState(2) shared by the 10 initialisations
  * k-means++ (`_kmeans_plusplus`): first centre rs.choice(n, p=1/n); per further centre
    rs.uniform(size=2+int(ln K))*pot, searchsorted(cumsum(closest)), clip, candidate distances
    through  -2 X Y^T + |x|^2 + |y|^2  clamped at 0, keep the candidate of least potential
  * Elkan (`init_bounds_dense`, `elkan_iter_chunked_dense`/`_update_chunk_dense`) with
    `_euclidean_dense_dense`: blocks of four features, `res += (d0^2+d1^2+d2^2+d3^2)`, remainder one
    by one; distances compared as square roots; new centre = member sum times (1.0/weight)
  * empty-cluster relocation (`_relocate_empty_clusters_dense`), `_average_centers`, `_center_shift`
  * stop on label equality (strict) or sum(shift^2) <= tol (then one label-only pass)
  * inertia = sequential sum over samples of squared distance to own centre (`_inertia_dense`)
  * a run replaces the best iff inertia < best and not `_is_same_clustering`
  * predict: first minimum over j of |c_j|^2 - 2 x.c_j on the un-centred X, centres shifted back
    (`lloyd_iter_chunked_dense` / `_update_chunk_dense`, chunks of 256 samples)
"""
import math

import numpy as np

N_INIT = 10
MAX_ITER = 300
TOL = 1e-4
SEED = 2
CHUNK = 256


def _sqdist(a, b):
    """_euclidean_dense_dense(..., squared=True): 4-blocked accumulation, plain mul/add."""
    n = a.shape[0]
    res = 0.0
    m = n // 4
    i = 0
    for _ in range(m):
        d0 = a[i] - b[i]
        d1 = a[i + 1] - b[i + 1]
        d2 = a[i + 2] - b[i + 2]
        d3 = a[i + 3] - b[i + 3]
        res += (d0 * d0 + d1 * d1 + d2 * d2 + d3 * d3)
        i += 4
    for j in range(i, n):
        d = a[j] - b[j]
        res += d * d
    return float(res)


def _sqdist_rows(X, c):
    """Vectorised over samples: identical per-sample operation order to _sqdist(X[i], c)."""
    n, F = X.shape
    D = X - c[None, :]
    D *= D
    m = F // 4
    res = np.zeros(n)
    if m:
        Q = D[:, :4 * m].reshape(n, m, 4)
        blocks = ((Q[:, :, 0] + Q[:, :, 1]) + Q[:, :, 2]) + Q[:, :, 3]
        for b in range(m):
            res += blocks[:, b]
    for j in range(4 * m, F):
        res += D[:, j]
    return res


def row_norms_sq(X):
    return np.einsum("ij,ij->i", X, X)


def _eucl_sq(X, Y, YY):
    """sklearn.metrics.pairwise._euclidean_distances(X, Y, Y_norm_squared=YY, squared=True)."""
    XX = row_norms_sq(X)[:, None]
    D = -2 * (X @ Y.T)
    D += XX
    D += YY.reshape(1, -1)
    np.maximum(D, 0, out=D)
    return D


def _center_half_distances(C):
    """euclidean_distances(centers) / 2 (X is Y: diagonal forced to 0)."""
    XX = row_norms_sq(C)[:, None]
    D = -2 * (C @ C.T)
    D += XX
    D += XX.T
    np.maximum(D, 0, out=D)
    np.fill_diagonal(D, 0)
    return np.sqrt(D) / 2


def _kmeans_plusplus(X, K, xx, rs):
    n, F = X.shape
    centers = np.empty((K, F))
    sw = np.ones(n)
    trials = 2 + int(np.log(K))
    cid = rs.choice(n, p=sw / sw.sum())
    centers[0] = X[cid]
    closest = _eucl_sq(centers[0, None], X, xx)
    pot = closest @ sw
    for c in range(1, K):
        rand_vals = rs.uniform(size=trials) * pot
        cand = np.searchsorted(np.cumsum(sw * closest), rand_vals)
        np.clip(cand, None, closest.size - 1, out=cand)
        D = _eucl_sq(X[cand], X, xx)
        np.minimum(closest, D, out=D)
        pots = D @ sw.reshape(-1, 1)
        best = int(np.argmin(pots))
        pot = pots[best]
        closest = D[best]
        centers[c] = X[cand[best]]
    return centers


def _elkan(X, centers, max_iter, tol):
    n, F = X.shape
    K = centers.shape[0]
    centers = centers.copy()
    centers_new = np.zeros_like(centers)
    labels = np.full(n, -1, np.int32)
    labels_old = labels.copy()
    half = _center_half_distances(centers)
    nxt = np.partition(half, kth=1, axis=0)[1]
    ub = np.zeros(n)
    lb = np.zeros((n, K))
    shift = np.zeros(K)

    # init_bounds_dense
    d0 = np.sqrt(_sqdist_rows(X, centers[0]))
    for i in range(n):
        best = 0
        md = d0[i]
        lb[i, 0] = md
        for j in range(1, K):
            if md > half[best, j]:
                d = math.sqrt(_sqdist(X[i], centers[j]))
                lb[i, j] = d
                if d < md:
                    md = d
                    best = j
        labels[i] = best
        ub[i] = md

    def e_step(update):
        w = np.zeros(K)
        if update:
            centers_new[:] = 0
        for i in range(n):
            u = ub[i]
            tight = False
            lab = int(labels[i])
            if not nxt[lab] >= u:
                for j in range(K):
                    if j != lab and u > lb[i, j] and u > half[lab, j]:
                        if not tight:
                            u = math.sqrt(_sqdist(X[i], centers[lab]))
                            lb[i, lab] = u
                            tight = True
                        if u > lb[i, j] or u > half[lab, j]:
                            d = math.sqrt(_sqdist(X[i], centers[j]))
                            lb[i, j] = d
                            if d < u:
                                lab = j
                                u = d
                labels[i] = lab
                ub[i] = u
            if update:
                w[lab] += 1.0
                centers_new[lab] += X[i]
        return w

    strict = False
    n_iter = 0
    for n_iter in range(max_iter):
        w = e_step(True)
        # _relocate_empty_clusters_dense
        empty = np.where(w == 0)[0]
        if len(empty):
            dist = ((X - centers[labels]) ** 2).sum(axis=1)
            if np.max(dist) != 0:
                far = np.argpartition(dist, -len(empty))[:-len(empty) - 1:-1]
                for idx, new_c in enumerate(empty):
                    f = int(far[idx])
                    old_c = int(labels[f])
                    centers_new[old_c] -= X[f]
                    centers_new[new_c] = X[f]
                    w[new_c] = 1.0
                    w[old_c] -= 1.0
        # _average_centers
        amax = int(np.argmax(w))
        for j in range(K):
            if w[j] > 0:
                centers_new[j] *= (1.0 / w[j])
            else:
                centers_new[j] = centers_new[amax]
        # _center_shift
        for j in range(K):
            shift[j] = math.sqrt(_sqdist(centers_new[j], centers[j]))
        ub += shift[labels]
        lb -= shift[None, :]
        np.maximum(lb, 0, out=lb)
        half = _center_half_distances(centers_new)
        nxt = np.partition(half, kth=1, axis=0)[1]
        centers, centers_new = centers_new, centers
        if np.array_equal(labels, labels_old):
            strict = True
            break
        if (shift ** 2).sum() <= tol:
            break
        labels_old[:] = labels
    if not strict:
        e_step(False)
    inertia = 0.0
    for i in range(n):
        inertia += _sqdist(X[i], centers[labels[i]])
    return labels.copy(), inertia, centers.copy(), n_iter + 1


def _is_same_clustering(l1, l2, K):
    mapping = np.full(K, -1)
    for a, b in zip(l1, l2):
        if mapping[a] == -1:
            mapping[a] = b
        elif mapping[a] != b:
            return False
    return True


def kmeans_fit(X0, K, seed=SEED, n_init=N_INIT):
    """Returns (fit_labels, inertia, centers_uncentred)."""
    X = np.array(X0, dtype=np.float64, order="C")
    tol = np.mean(np.var(X, axis=0)) * TOL
    rs = np.random.RandomState(seed)
    mean = X.mean(axis=0)
    X -= mean
    xx = row_norms_sq(X)
    best = None
    for _ in range(n_init):
        c0 = _kmeans_plusplus(X, K, xx, rs)
        labels, inertia, centers, _ = _elkan(X, c0, MAX_ITER, tol)
        if best is None or (inertia < best[1] and not _is_same_clustering(labels, best[0], K)):
            best = (labels, inertia, centers)
    return best[0], best[1], best[2] + mean


def kmeans_predict(X0, centers):
    X = np.ascontiguousarray(X0, dtype=np.float64)
    cc = row_norms_sq(centers)
    out = np.empty(X.shape[0], np.int32)
    n = X.shape[0]
    chunk = CHUNK if n > CHUNK else n
    for s in range(0, n, chunk):
        Xc = X[s:s + chunk]
        P = np.repeat(cc[None, :], Xc.shape[0], axis=0)
        P += -2.0 * (Xc @ centers.T)
        out[s:s + chunk] = np.argmin(P, axis=1)
    return out


def kmeans_fit_predict(X0, K):
    """Labels as the reference sees them: KMeans(...).fit(X) then .predict(X)."""
    fit_labels, inertia, centers = kmeans_fit(X0, K)
    return kmeans_predict(X0, centers), inertia, centers, fit_labels


def sklearn_fit_predict(X0, K):
    """The same call through the INSTALLED scikit-learn (>= 1.4) forced to the pinned 1.3.0 behaviour
    (n_init=10; run with OMP_NUM_THREADS=1 for sequential reductions) -- the library the unmodified reference
    itself runs on in this image (oracle/run_reference.py).  For clustering problems too large for the
    pure-Python restatement above (deep loci: thousands of sequences x 4^k k-mers); the restatement is pinned
    against this very function by tests/test_oracle_kmeans.py.  Same return shape as kmeans_fit_predict."""
    from sklearn.cluster import KMeans

    X = np.ascontiguousarray(X0, dtype=np.float64)
    km = KMeans(n_clusters=K, random_state=SEED, algorithm="elkan", n_init=N_INIT).fit(X)
    return km.predict(X), float(km.inertia_), km.cluster_centers_, km.labels_


def hybrid_fit_predict(X0, K, limit=1 << 16):
    """Restatement for small problems, installed scikit-learn above `limit` matrix elements."""
    if X0.shape[0] * X0.shape[1] <= limit:
        return kmeans_fit_predict(X0, K)
    return sklearn_fit_predict(X0, K)
