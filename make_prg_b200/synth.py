"""Seeded synthetic MSA generator for the BASELINE.json configs.

Haplotype-pool model: a random ACGT root; `n_haps` haplotypes related by a random recursive
bipartition (the clades); a fraction of columns is variable, each mutating one size-weighted random
clade to a different base; a few clade-level deletions of 1-8 columns; rows draw haplotypes with
Zipf(1) probabilities.  Pure numpy, deterministic in `seed`.
"""
import numpy as np

BASES = np.frombuffer(b"ACGT", dtype=np.uint8)
GAP = ord("-")

# seeds per config
CONFIG2_SEED0 = 1000
CONFIG3_SEED0 = 2_000_000
CONFIG5_SEED0 = 3_000_000


def _clades(n_haps, rng):
    order = rng.permutation(n_haps)
    clades = []
    stack = [(0, n_haps)]
    while stack:
        lo, hi = stack.pop()
        clades.append(order[lo:hi])
        if hi - lo >= 2:
            cut = int(rng.integers(lo + 1, hi))
            stack.append((cut, hi))
            stack.append((lo, cut))
    return clades


def synth_msa(rows, cols, seed, n_haps=None, var_frac=0.04, n_dels=3, private_snp=0.0):
    """Returns a uint8[rows, cols] ASCII matrix over ACGT and '-'."""
    rng = np.random.default_rng(seed)
    if n_haps is None:
        n_haps = max(2, rows // 5)
    root = BASES[rng.integers(0, 4, cols)]
    clades = _clades(n_haps, rng)
    sizes = np.array([len(c) for c in clades], dtype=np.float64)
    weights = sizes / sizes.sum()
    H = np.tile(root, (n_haps, 1))
    n_var = int(round(var_frac * cols))
    var_cols = rng.choice(cols, size=n_var, replace=False)
    clade_pick = rng.choice(len(clades), size=n_var, p=weights)
    shift = rng.integers(1, 4, size=n_var)
    for c, ci, sh in zip(var_cols, clade_pick, shift):
        old = int(np.searchsorted(BASES, root[c]))
        H[clades[ci], c] = BASES[(old + sh) % 4]
    for _ in range(n_dels):
        ci = int(rng.choice(len(clades), p=weights))
        length = int(rng.integers(1, 9))
        start = int(rng.integers(0, max(1, cols - length)))
        H[clades[ci], start:start + length] = GAP
    p = 1.0 / (np.arange(n_haps) + 1.0)
    p /= p.sum()
    M = H[rng.choice(n_haps, size=rows, p=p)].copy()
    if private_snp > 0:
        mask = (rng.random(M.shape) < private_snp) & (M != GAP)
        sh = rng.integers(1, 4, size=int(mask.sum()))
        old = np.searchsorted(BASES, M[mask])
        M[mask] = BASES[(old + sh) % 4]
    return M


def synth_deep_msa(rows, cols, seed, n_clades=8, clade_div=0.30, n_haps=None, var_frac=0.04, n_dels=3,
                   private_snp=0.01, flank=0):
    """A locus whose rows fall into `n_clades` deep clades: every clade founder differs from the root at
    `clade_div` of the columns (pairwise divergence between clades ~ 2 * clade_div > the 20 % one-reference-
    like threshold, cluster_sequences.py:59-104), so the KMeans loop of kmeans_cluster_seqs (:256-274) runs
    until the clades are separated.  Inside a clade the rows follow the haplotype-pool model of synth_msa
    (`var_frac` clade-consistent variable columns, deletions, `private_snp` per-row SNPs).  Clade sizes are
    Zipf-like.  `flank` conserved columns at both ends give the root its match intervals."""
    rng = np.random.default_rng(seed)
    root = BASES[rng.integers(0, 4, cols)]
    w = 1.0 / (np.arange(n_clades) + 2.0)
    sizes = np.maximum(2, np.floor(rows * w / w.sum()).astype(np.int64))
    sizes[0] += rows - int(sizes.sum())
    if n_haps is None:
        n_haps = max(2 * n_clades, rows // 5)
    parts = []
    inner = slice(flank, cols - flank) if flank else slice(0, cols)
    n_inner = inner.stop - inner.start
    for c in range(n_clades):
        founder = root.copy()
        hit = inner.start + rng.choice(n_inner, size=int(round(clade_div * n_inner)), replace=False)
        old = np.searchsorted(BASES, founder[hit])
        founder[hit] = BASES[(old + rng.integers(1, 4, size=hit.size)) % 4]
        haps = max(2, int(n_haps * sizes[c] / rows))
        clades = _clades(haps, rng)
        csz = np.array([len(x) for x in clades], dtype=np.float64)
        cw = csz / csz.sum()
        H = np.tile(founder, (haps, 1))
        n_var = int(round(var_frac * n_inner))
        var_cols = inner.start + rng.choice(n_inner, size=n_var, replace=False)
        pick = rng.choice(len(clades), size=n_var, p=cw)
        shift = rng.integers(1, 4, size=n_var)
        for col, ci, sh in zip(var_cols, pick, shift):
            o = int(np.searchsorted(BASES, founder[col]))
            H[clades[ci], col] = BASES[(o + sh) % 4]
        for _ in range(n_dels):
            ci = int(rng.choice(len(clades), p=cw))
            length = int(rng.integers(1, 9))
            start = int(rng.integers(inner.start, max(inner.start + 1, inner.stop - length)))
            H[clades[ci], start:start + length] = GAP
        p = 1.0 / (np.arange(haps) + 1.0)
        p /= p.sum()
        parts.append(H[rng.choice(haps, size=int(sizes[c]), p=p)])
    M = np.concatenate(parts)[rng.permutation(rows)].copy()
    if private_snp > 0:
        mask = (rng.random(M.shape) < private_snp) & (M != GAP)
        if flank:
            mask[:, :flank] = False
            mask[:, cols - flank:] = False
        sh = rng.integers(1, 4, size=int(mask.sum()))
        old = np.searchsorted(BASES, M[mask])
        M[mask] = BASES[(old + sh) % 4]
    return M


def config_msa(config, index, rows=None, cols=None):
    """The i-th locus of BASELINE config 2, 3, 4 or 5."""
    if config == 2:
        return synth_msa(rows or 200, cols or 1000, CONFIG2_SEED0 + index)
    if config == 3:
        rng = np.random.default_rng(CONFIG3_SEED0 + index)
        c = cols or int(rng.integers(600, 1401))
        return synth_msa(rows or 500, c, CONFIG3_SEED0 + index)
    if config == 4:
        # BASELINE configs[3] "kmer-count + KMeans GEMM dominated": deep clades, so that the clustering loop
        # really runs KMeans on thousands of distinct long sequences (round 1's generator never did: every
        # row was within 20 % of the majority string); "4flat" below keeps that older shape
        return synth_deep_msa(rows or 10_000, cols or 20_000, 4_000_000 + index, n_clades=8, clade_div=0.30,
                              n_haps=2000, var_frac=0.04, private_snp=0.01)
    if config == "4flat":
        return synth_msa(rows or 10_000, cols or 20_000, 4_000_000 + index, n_haps=2000,
                         var_frac=0.04, private_snp=0.01)
    if config == 5:
        return synth_msa(rows or 200, cols or 1000, CONFIG5_SEED0 + index, var_frac=0.25)
    raise ValueError(f"unknown config {config}")


def to_fasta(M, prefix="s"):
    lines = []
    for i, row in enumerate(M):
        lines.append(f">{prefix}{i}\n{row.tobytes().decode()}\n")
    return "".join(lines)
