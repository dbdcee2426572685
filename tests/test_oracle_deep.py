"""CPU: the oracle on deep-clade loci (config #4 class) against runs of the unmodified reference
(tests/golden/deep.json, oracle/gen_golden_deep.py): k-mer count matrices bit-equal to what the reference
handed to scikit-learn, PRG byte-identical (by sha256) when the recorded scikit-learn labels are replayed, and
the installed scikit-learn (forced to n_init=10) reproducing the recorded labels and inertia."""
import hashlib
import os

import numpy as np
import pytest

os.environ.setdefault("OMP_NUM_THREADS", "1")

import kmeans13
from helpers import deep_cases, deep_kmeans_problems, deep_msa


@pytest.mark.parametrize("name", ["deep_300", "deep_6clades_L5", "odd_F_a", "odd_F_b", "deep_11clades"])
def test_oracle_reproduces_reference_on_deep_loci(name):
    case = deep_cases()[name]
    M = deep_msa(case)
    problems, prg = deep_kmeans_problems(case, M)
    assert len(problems) == len(case["kmeans"]) >= 4
    assert hashlib.sha256(prg.encode()).hexdigest() == case["prg_sha256"] and len(prg) == case["prg_len"]


def test_installed_sklearn_reproduces_recorded_deep_kmeans():
    case = deep_cases()["odd_F_a"]
    problems, _ = deep_kmeans_problems(case, deep_msa(case))
    for X, K, g in problems:
        assert X.shape[1] % 2 == 1 and X.shape[1] > 384  # the odd-F blocked-dgemm regime
        labels, inertia, _, _ = kmeans13.sklearn_fit_predict(X, K)
        assert labels.tolist() == g["labels"] and float(inertia).hex() == g["inertia"]
