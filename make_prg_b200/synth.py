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


def config_msa(config, index, rows=None, cols=None):
    """The i-th locus of BASELINE config 2, 3, 4 or 5."""
    if config == 2:
        return synth_msa(rows or 200, cols or 1000, CONFIG2_SEED0 + index)
    if config == 3:
        rng = np.random.default_rng(CONFIG3_SEED0 + index)
        c = cols or int(rng.integers(600, 1401))
        return synth_msa(rows or 500, c, CONFIG3_SEED0 + index)
    if config == 4:
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
