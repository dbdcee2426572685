"""-m gpu parity: kernels (b) and (c) through the C ABI against the oracle and the golden vectors
(tests/golden/units.json, kmeans_cases.npz generated from the reference + scikit-learn)."""
import numpy as np
import pytest

import kmeans13
import make_prg_oracle as mo
from helpers import kmeans_cases, rows_to_matrix, unit_cases
from make_prg_b200 import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    from make_prg_b200 import device

    return device.Context(0)


def _oracle_dedupe(S):
    ung = mo.ungapped_rows(S)
    first = {}
    group = []
    for u in ung:
        group.append(first.setdefault(u, len(first)))
    return group, [len(u) for u in ung], len(first), len({bytes(r) for r in S})


def test_dedupe_rows(ctx):
    rng = np.random.default_rng(3)
    mats = [synth.synth_msa(60, 200, 5, var_frac=0.1, n_dels=10),
            synth.synth_msa(300, 90, 6, var_frac=0.2, n_dels=5)]
    batch = ctx.upload(mats)
    tasks = []
    for l, M in enumerate(mats):
        for _ in range(20):
            c0 = int(rng.integers(0, M.shape[1] - 1))
            c1 = int(rng.integers(c0 + 1, min(M.shape[1], c0 + 60) + 1))
            rows = np.sort(rng.choice(M.shape[0], int(rng.integers(1, M.shape[0] + 1)), replace=False))
            tasks.append((l, rows, c0, c1))
        tasks.append((l, None, 0, M.shape[1]))
    for (l, rows, c0, c1), (group, ulen, nu, ng) in zip(tasks, ctx.dedupe_rows(batch, tasks)):
        S = mats[l][:, c0:c1] if rows is None else mats[l][rows, c0:c1]
        g, ul, n_u, n_g = _oracle_dedupe(S)
        assert group.tolist() == g and ulen.tolist() == ul and (nu, ng) == (n_u, n_g)


def test_dedupe_rows_around_the_warp_list_capacity(ctx):
    """Small tasks de-duplicate on one warp that keeps at most 64 distinct rows (dedupe_warp_kernel); a task with
    more is redone by the one-CTA kernel.  63 / 64 / 65 / 130 distinct rows, with and without gap-only differences,
    first occurrences spread over the 32-row passes."""
    rng = np.random.default_rng(11)
    mats, tasks = [], []
    for distinct in (1, 31, 63, 64, 65, 130):
        pats = set()
        while len(pats) < distinct:
            pats.add(bytes(rng.choice(list(b"ACGT-"), size=24).astype(np.uint8)))
        pats = [np.frombuffer(p, np.uint8) for p in sorted(pats)]
        rows = [pats[int(i)] for i in rng.permutation(np.concatenate([np.arange(distinct),
                                                                    rng.integers(0, distinct, 200 - min(distinct, 200))]))]
        mats.append(np.stack(rows))
        tasks.append((len(mats) - 1, None, 0, 24))
        tasks.append((len(mats) - 1, np.sort(rng.choice(len(rows), len(rows) // 2, replace=False)), 3, 21))
    batch = ctx.upload(mats)
    for (l, rows, c0, c1), (group, ulen, nu, ng) in zip(tasks, ctx.dedupe_rows(batch, tasks)):
        S = mats[l][:, c0:c1] if rows is None else mats[l][rows, c0:c1]
        g, ul, n_u, n_g = _oracle_dedupe(S)
        assert group.tolist() == g and ulen.tolist() == ul and (nu, ng) == (n_u, n_g), (l, c0, c1)


def test_dedupe_rows_whole_grid_path(ctx, monkeypatch):
    """Deep tasks de-duplicate with the whole grid (dedupe_big_* kernels): same outputs as the one-CTA
    kernel, on the forced path for small windows and on a task that takes it by size."""
    rng = np.random.default_rng(8)
    mats = [synth.synth_msa(70, 260, 21, var_frac=0.1, n_dels=12), rows_to_matrix(["A--C", "AC--", "-A-C", "----", "----"]),
            synth.synth_msa(2500, 1800, 22, n_haps=300, var_frac=0.05, n_dels=6, private_snp=0.0005)]
    mats[2][5] = mats[2][3]  # exact duplicates far apart
    mats[2][2400] = mats[2][3]
    batch = ctx.upload(mats)
    tasks = [(0, None, 0, 260), (1, None, 0, 4), (1, np.array([3, 4]), 1, 3)]
    for _ in range(25):
        c0 = int(rng.integers(0, 259))
        c1 = int(rng.integers(c0 + 1, min(260, c0 + 70) + 1))
        tasks.append((0, np.sort(rng.choice(70, int(rng.integers(1, 71)), replace=False)), c0, c1))
    plain = ctx.dedupe_rows(batch, tasks)
    monkeypatch.setenv("MPRG_FORCE_BIG_DEDUPE", "1")
    forced = ctx.dedupe_rows(batch, tasks)
    monkeypatch.delenv("MPRG_FORCE_BIG_DEDUPE")
    for (l, rows, c0, c1), a, b in zip(tasks, plain, forced):
        S = mats[l][:, c0:c1] if rows is None else mats[l][rows, c0:c1]
        g, ul, n_u, n_g = _oracle_dedupe(S)
        for group, ulen, nu, ng in (a, b):
            assert group.tolist() == g and ulen.tolist() == ul and (nu, ng) == (n_u, n_g)
    # 2,500 x 1,800 = 4.5 M symbols: above DEDUPE_BIG_SYMBOLS, next to a small task of the same level
    (group, ulen, nu, ng), small = ctx.dedupe_rows(batch, [(2, None, 0, 1800), (0, None, 10, 40)])
    g, ul, n_u, n_g = _oracle_dedupe(mats[2])
    assert group.tolist() == g and ulen.tolist() == ul and (nu, ng) == (n_u, n_g)
    g, ul, n_u, n_g = _oracle_dedupe(mats[0][:, 10:40])
    assert small[0].tolist() == g and small[1].tolist() == ul and small[2:] == (n_u, n_g)


def test_kmer_count_matrix(ctx):
    rng = np.random.default_rng(4)
    mats = [synth.synth_msa(50, 120, 15, var_frac=0.15, n_dels=6),
            rows_to_matrix(["AAAAT", "AATA-", "AAAAT", "TTTTT"])]
    mats[0][3, 10:20] = np.frombuffer(b"RYKMSWRYKM", np.uint8)  # k-mers are raw substrings
    batch = ctx.upload(mats)
    for l, c0, c1, k in [(0, 0, 120, 7), (0, 5, 60, 3), (0, 30, 50, 1), (0, 0, 120, 15), (1, 0, 5, 3)]:
        S = mats[l][:, c0:c1]
        seqs = list(dict.fromkeys(u for u in mo.ungapped_rows(S) if len(u) >= k))
        want = mo.count_kmer_occurrences(seqs, mo.count_distinct_kmers(seqs, k))
        got = ctx.kmer_counts(batch, (l, None, c0, c1), k)
        assert got.shape == want.shape and np.array_equal(got, want)


def test_kmer_count_matrix_big_problem(ctx):
    """Problems with >= 2^19 k-mer positions take the whole-grid path (hash numbering over all CTAs,
    shared-memory histograms): same first-occurrence column order and counts as the oracle, both for few
    distinct k-mers (shared-memory fill) and for very many (global-atomic fill)."""
    few = synth.synth_msa(320, 2000, 77, n_haps=300, var_frac=0.05, n_dels=5, private_snp=0.002)  # F ~ 10^4
    many = synth.synth_msa(320, 2000, 78, n_haps=300, var_frac=0.05, n_dels=5, private_snp=0.05)
    batch = ctx.upload([few, many])
    for l, M, k in [(0, few, 7), (1, many, 9)]:
        seqs = list(dict.fromkeys(u for u in mo.ungapped_rows(M) if len(u) >= k))
        assert sum(len(u) - k + 1 for u in seqs) >= 1 << 19
        want = mo.count_kmer_occurrences(seqs, mo.count_distinct_kmers(seqs, k))
        got = ctx.kmer_counts(batch, (l, None, 0, M.shape[1]), k)
        assert got.shape == want.shape and np.array_equal(got, want)
        assert (want.shape[1] * 4 > 200 * 1024) == (l == 1)  # only the second case is past the smem histogram


def test_kmeans_cta_groups_equal_single_cta(ctx):
    """Deep loci run every initialisation on a group of co-resident CTAs (global-memory barrier, loads
    through L2).  Same operations in the same order: labels and inertia must be bit-identical to the
    one-CTA kernel, which is pinned to scikit-learn by the golden cases -- on those cases, and on count
    matrices of the size that takes the group path in the engine."""
    n = 0
    for X, K, labels, inertia in kmeans_cases():
        if n % 9 == 0:
            got, got_inertia = ctx.kmeans(X, K, mode=2)
            assert np.array_equal(got, labels) and got_inertia == inertia, (X.shape, K)
        n += 1
    rng = np.random.default_rng(11)
    for rows, F, K in [(700, 3100, 2), (1500, 1500, 5), (400, 6000, 10)]:
        centres = rng.integers(0, 4, (K + 1, F))
        X = (centres[rng.integers(0, K + 1, rows)] + (rng.random((rows, F)) < 0.02)).astype(np.float64)
        a, ia = ctx.kmeans(X, K, mode=1)
        b, ib = ctx.kmeans(X, K, mode=2)
        assert np.array_equal(a, b) and ia == ib, (rows, F, K)


def test_kmeans_golden_cases(ctx):
    """Labels identical and inertia bit-identical to what scikit-learn returned inside the reference on
    every count matrix it was handed (north_star asks for identical labels, inertia within 1e-6)."""
    n = 0
    for X, K, labels, inertia in kmeans_cases():
        got, got_inertia = ctx.kmeans(X, K)
        assert np.array_equal(got, labels), (X.shape, K)
        assert got_inertia == inertia, (X.shape, K)
        n += 1
    assert n > 500


def test_kmeans_random_tie_prone_against_oracle(ctx):
    rng = np.random.default_rng(11)
    for t in range(60):
        n = int(rng.integers(3, 14))
        F = int(rng.integers(1, 24))
        X = rng.integers(0, 3, size=(n, F)).astype(float)
        if t % 3 == 0:
            X = (rng.random((n, F)) < 0.25).astype(float)
        for K in range(2, min(10, n - 1) + 1):
            want, inertia, _, _ = kmeans13.kmeans_fit_predict(X, K)
            got, got_inertia = ctx.kmeans(X, K)
            assert np.array_equal(got, want), (t, n, F, K)
            assert got_inertia == inertia


def test_kmeans_larger_problems_against_sklearn(ctx):
    sk = pytest.importorskip("sklearn.cluster")
    from threadpoolctl import threadpool_limits

    rng = np.random.default_rng(21)
    with threadpool_limits(limits=1, user_api="openmp"):
        for n, F, K in [(40, 200, 4), (120, 64, 10), (150, 500, 7), (60, 33, 3), (200, 40, 9), (35, 700, 5)]:
            base = rng.integers(0, 3, (max(2, n // 5), F))
            X = base[rng.integers(0, len(base), n)].astype(float)
            X[rng.random((n, F)) < 0.03] += 1
            m = sk.KMeans(n_clusters=K, random_state=2, algorithm="elkan", n_init=10).fit(X)
            got, got_inertia = ctx.kmeans(X, K)
            assert np.array_equal(got, m.predict(X)), (n, F, K)
            assert got_inertia == m.inertia_


def test_one_ref_like(ctx):
    rng = np.random.default_rng(8)
    M = synth.synth_msa(40, 60, 21, var_frac=0.3, n_dels=4)
    batch = ctx.upload([M])
    for w0, w1 in [(0, 60), (3, 7), (10, 40), (20, 24)]:
        for _ in range(6):
            K = int(rng.integers(1, 5))
            cl = rng.integers(0, K, 40)
            cl[:K] = np.arange(K)
            got = ctx.one_ref_like(batch, (0, None, w0, w1), cl, K)
            for c in range(K):
                seqs = [M[r, w0:w1].tobytes().decode() for r in range(40) if cl[r] == c]
                assert bool(got[c]) == mo.sequences_are_one_reference_like(seqs)


def test_cluster_tasks_unit_vectors(ctx):
    cases = [r for r in unit_cases() if r["clustered_ids"] is not None and "N" not in "".join(r["rows"])]
    mats = [rows_to_matrix(r["rows"]) for r in cases]
    batch = ctx.upload(mats)
    for L in sorted({r["L"] for r in cases}):
        idx = [i for i, r in enumerate(cases) if r["L"] == L]
        res = ctx.cluster_tasks(batch, [(i, None, 0, mats[i].shape[1]) for i in idx], L)
        for i, clusters in zip(idx, res):
            want = [sorted(int(s[1:]) for s in cl) for cl in cases[i]["clustered_ids"]]
            assert clusters == want, (cases[i]["rows"], L)
