"""Golden vectors for DEEP loci (BASELINE configs[3] class: the clustering loop really runs KMeans on
hundreds to thousands of distinct long sequences), produced by the UNMODIFIED reference through
oracle/run_reference.py (real scikit-learn forced to n_init=10, one OpenMP thread).

TEST INFRASTRUCTURE ONLY.  Run once in the build container (`python oracle/gen_golden_deep.py`); the output
tests/golden/deep.json is committed because /root/reference does not exist on the GPU box.  Per case:
the generator call (make_prg_b200.synth, regenerated from the seed by the tests, pinned by msa_sha256), the
reference's PRG (sha256 + length), its pre-order tree dump, and every KMeans problem the reference handed to
scikit-learn in call order: (n, F, K, predict labels, inertia as float.hex()).
"""
import hashlib
import json
import os
import sys
import tempfile
import time
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
REPO = HERE.parent
sys.path.insert(0, str(REPO))
sys.path.insert(0, str(HERE))

import run_reference as rr  # noqa: E402
from make_prg_b200 import synth  # noqa: E402

GOLD = REPO / "tests" / "golden"

# name -> (generator kwargs, N, L).  What each case is there to exercise on the device:
#   deep_300     n*F >= 2^21: the CTA-group KMeans inside the engine; one-CTA one-reference-like check
#   deep_1500    rows*width >= 2^22: whole-grid one-reference-like check with K > 1, whole-grid k-mer numbering
#   odd_F_*      count matrices with an ODD number of distinct k-mers above 384 (blocked-dgemm regime)
CASES = {
    "deep_300": (dict(rows=300, cols=800, seed=4_000_000, n_clades=8, n_haps=60), 10, 7),
    "deep_1500": (dict(rows=1500, cols=4000, seed=4_000_000, n_clades=8, n_haps=300), 10, 7),
    "deep_6clades_L5": (dict(rows=400, cols=1200, seed=4_000_017, n_clades=6, n_haps=80, clade_div=0.35), 10, 5),
    "deep_11clades": (dict(rows=600, cols=1000, seed=4_000_023, n_clades=11, n_haps=120), 10, 7),
    "odd_F_a": (dict(rows=120, cols=500, seed=4_000_101, n_clades=5, n_haps=30), 10, 7),
    "odd_F_b": (dict(rows=120, cols=500, seed=4_000_103, n_clades=5, n_haps=30), 10, 7),
}


def make_msa(kw):
    return synth.synth_deep_msa(**kw)


def tap_kmeans(calls):
    import make_prg.from_msa.cluster_sequences as cs

    base = cs.KMeans

    class Tap(base):
        def fit(self, X, *a, **k):
            res = super().fit(X, *a, **k)
            calls.append({"n": int(X.shape[0]), "F": int(X.shape[1]), "K": int(self.n_clusters),
                          "labels": self.predict(X).astype(int).tolist(), "inertia": float(self.inertia_).hex(),
                          "x_sha256": hashlib.sha256(np.ascontiguousarray(X, np.float64).tobytes()).hexdigest()})
            return res

    cs.KMeans = Tap


def main():
    only = set(sys.argv[1:])
    rr.load_reference()
    calls = []
    tap_kmeans(calls)
    out_path = GOLD / "deep.json"
    out = json.loads(out_path.read_text()) if out_path.exists() else {}
    for name, (kw, N, L) in CASES.items():
        if only and name not in only:
            continue
        M = make_msa(kw)
        with tempfile.NamedTemporaryFile("w", suffix=".fa", delete=False) as fh:
            fh.write(synth.to_fasta(M))
            path = fh.name
        del calls[:]
        t0 = time.time()
        builder, prg = rr.ref_build(path, N, L, locus_name=name)
        os.unlink(path)
        big = [c for c in calls if c["n"] * c["F"] >= 4096]
        out[name] = {"gen": kw, "N": N, "L": L, "shape": list(M.shape),
                     "msa_sha256": hashlib.sha256(M.tobytes()).hexdigest(),
                     "prg_sha256": hashlib.sha256(prg.encode()).hexdigest(), "prg_len": len(prg),
                     "n_nodes": builder.next_node_id, "n_sites": (builder.site_num - 5) // 2,
                     "tree": rr.dump_tree(builder), "n_kmeans_calls": len(calls), "kmeans": big,
                     "reference_seconds": round(time.time() - t0, 1)}
        print(name, M.shape, "nodes", builder.next_node_id, "kmeans calls", len(calls), "big",
              [(c["n"], c["F"], c["K"]) for c in big], f"{time.time() - t0:.0f} s", flush=True)
        with open(out_path, "w") as fh:
            json.dump(out, fh)


if __name__ == "__main__":
    main()
