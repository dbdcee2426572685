"""Full-size BASELINE config #4 (one 10,000 x 20,000 deep-clade locus, -N 10 -L 7) through the ORACLE in the
build container: sha256 of its PRG -> tests/golden/config4_deep_oracle.json (the GPU box compares the
sha256 of the engine's PRG with it, scripts/run_configs.py).

TEST INFRASTRUCTURE ONLY.  KMeans problems of this size (10,000 x 16,384) go to the installed scikit-learn
forced to the pinned 1.3.0 behaviour (kmeans13.hybrid_fit_predict), one OpenMP thread; everything else is the
numpy restatement.  The reduced-size deep loci of tests/golden/deep.json come from the unmodified reference
itself and pin this oracle (tests/test_oracle_deep.py)."""
import hashlib
import json
import os
import sys
import time
from pathlib import Path

os.environ["OMP_NUM_THREADS"] = "1"
HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent))
sys.path.insert(0, str(HERE))

import kmeans13  # noqa: E402
import make_prg_oracle as mo  # noqa: E402
from make_prg_b200 import synth  # noqa: E402

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000
cols = int(sys.argv[2]) if len(sys.argv) > 2 else 20_000
calls = []


def km(X, K):
    t0 = time.time()
    r = kmeans13.hybrid_fit_predict(X, K)
    calls.append({"n": int(X.shape[0]), "F": int(X.shape[1]), "K": int(K), "seconds": round(time.time() - t0, 1),
                  "inertia": float(r[1]).hex()})
    print("kmeans", calls[-1], flush=True)
    return r


M = synth.config_msa(4, 0, rows, cols)
t0 = time.time()
prg, b = mo.build_prg_from_matrix([f"s{i}" for i in range(M.shape[0])], M, 10, 7, kmeans=km)
dt = time.time() - t0
rec = {"config": 4, "generator": "synth.config_msa(4, 0): synth_deep_msa, 8 clades at 30 % divergence", "rows": rows,
       "cols": cols, "N": 10, "L": 7, "msa_sha256": hashlib.sha256(M.tobytes()).hexdigest(), "prg_len": len(prg),
       "prg_sha256": hashlib.sha256(prg.encode()).hexdigest(), "oracle_seconds_1core": dt,
       "kmeans_calls": [c for c in calls if c["n"] * c["F"] >= 4096], "n_kmeans_calls": len(calls)}
name = "config4_deep_oracle.json" if (rows, cols) == (10_000, 20_000) else f"config4_deep_oracle_{rows}x{cols}.json"
(HERE.parent / "tests" / "golden" / name).write_text(json.dumps(rec))
print(json.dumps({k: v for k, v in rec.items() if k != "kmeans_calls"}))
