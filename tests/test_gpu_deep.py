"""-m gpu: deep-clade loci (BASELINE configs[3] class) through the C ABI against runs of the UNMODIFIED
reference (tests/golden/deep.json): the clustering loop of kmeans_cluster_seqs
(make_prg/from_msa/cluster_sequences.py:256-274) really runs KMeans on hundreds to thousands of distinct long
sequences here, so these tests pin what round 1 left unpinned:
  (i)   the CTA-group KMeans chosen BY THE ENGINE (n * F >= 2^21),
  (ii)  the whole-grid one-reference-like check with more than one cluster,
  (iii) count matrices with an odd number of k-mers above 384 (OpenBLAS' blocked-dgemm regime).
PRG strings byte-identical (sha256), recursion trees identical, KMeans labels identical and inertia within
1e-6 relative (north_star's tolerance; bit-equality is reported)."""
import hashlib

import numpy as np
import pytest

from helpers import deep_cases, deep_kmeans_problems, deep_msa, tree_dump

pytestmark = pytest.mark.gpu

# case -> kernel variants the engine must have chosen (Context.path_counts)
EXPECT = {
    "deep_300": ("kmeans_group",),
    "deep_1500": ("kmeans_group", "refcheck_grid_multi", "kmer_grid", "dedupe_grid"),
    "deep_6clades_L5": ("kmeans_cta",),
    "deep_11clades": ("kmeans_group",),
    "odd_F_a": ("kmeans_cta",),
    "odd_F_b": ("kmeans_cta",),
}


@pytest.fixture(scope="module")
def ctx():
    from make_prg_b200 import device

    return device.Context(0)


@pytest.mark.parametrize("name", sorted(EXPECT))
def test_deep_locus_equals_reference(ctx, name):
    case = deep_cases()[name]
    M = deep_msa(case)
    ctx.path_counts(reset=True)
    batch = ctx.upload([M])
    res = ctx.build(batch, case["N"], case["L"])
    paths = ctx.path_counts(reset=True)
    assert res.status(0) == 0
    prg = res.prg(0)
    assert len(prg) == case["prg_len"]
    assert hashlib.sha256(prg.encode()).hexdigest() == case["prg_sha256"]
    assert res.n_nodes(0) == case["n_nodes"] and res.n_sites(0) == case["n_sites"]
    assert tree_dump(res, 0, M) == [list(t) for t in case["tree"]]
    for variant in EXPECT[name]:
        assert paths[variant] > 0, (variant, paths)
    res.free()
    batch.free()


@pytest.mark.parametrize("name", ["deep_300", "odd_F_a", "odd_F_b", "deep_6clades_L5", "deep_11clades"])
def test_deep_kmeans_problems_equal_sklearn(ctx, name):
    """Every big (X, K) the reference handed to scikit-learn on this locus, through mprg_kmeans with the
    engine's own choice of kernel: identical labels, inertia within 1e-6 relative."""
    case = deep_cases()[name]
    problems, _ = deep_kmeans_problems(case, deep_msa(case))
    assert problems
    bit_equal = 0
    for X, K, g in problems:
        labels, inertia = ctx.kmeans(X, K)
        want = float.fromhex(g["inertia"])
        assert labels.tolist() == g["labels"], (name, X.shape, K)
        assert abs(inertia - want) <= 1e-6 * abs(want), (name, X.shape, K, inertia, want)
        bit_equal += int(float(inertia).hex() == g["inertia"])
    print(f"{name}: {len(problems)} problems, inertia bit-equal in {bit_equal}")


def test_deep_1500_first_kmeans_rounds(ctx):
    """The 1500 x 16384 problems of deep_1500 (K = 2, 3 and 8) on the CTA-group kernel."""
    case = deep_cases()["deep_1500"]
    problems, _ = deep_kmeans_problems(case, deep_msa(case))
    for X, K, g in problems:
        if K not in (2, 3, 8):
            continue
        labels, inertia = ctx.kmeans(X, K)
        want = float.fromhex(g["inertia"])
        assert labels.tolist() == g["labels"], K
        assert abs(inertia - want) <= 1e-6 * abs(want)
