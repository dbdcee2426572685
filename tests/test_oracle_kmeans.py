"""Pins oracle/kmeans13.py (restated scikit-learn 1.3.0 KMeans, n_init=10, sequential reductions)
against labels/inertia produced by scikit-learn inside the unmodified reference
(tests/golden/kmeans_cases.npz) and against the scikit-learn installed in this image."""
import numpy as np
import pytest

import kmeans13
from helpers import kmeans_cases

CASES = list(kmeans_cases())


def test_golden_kmeans_cases_labels_and_inertia_bit_identical():
    assert len(CASES) > 100
    step = max(1, len(CASES) // 150)  # a spread of ~150 cases keeps the CPU suite short
    for X, K, labels, inertia in CASES[::step] + CASES[-5:]:
        got, got_inertia, _, _ = kmeans13.kmeans_fit_predict(X, K)
        assert np.array_equal(got, labels)
        assert got_inertia == inertia


def test_against_installed_sklearn_on_tie_prone_matrices():
    sk = pytest.importorskip("sklearn.cluster")
    from threadpoolctl import threadpool_limits

    rng = np.random.default_rng(7)
    with threadpool_limits(limits=1, user_api="openmp"):
        for t in range(40):
            n = int(rng.integers(3, 14))
            F = int(rng.integers(1, 24))
            X = rng.integers(0, 3, size=(n, F)).astype(float)
            if t % 3 == 0:
                X = (rng.random((n, F)) < 0.25).astype(float)
            for K in range(2, min(10, n - 1) + 1):
                m = sk.KMeans(n_clusters=K, random_state=2, algorithm="elkan", n_init=10).fit(X)
                got, inertia, _, _ = kmeans13.kmeans_fit_predict(X, K)
                assert np.array_equal(got, m.predict(X))
                assert inertia == m.inertia_
