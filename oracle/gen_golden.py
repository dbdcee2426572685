"""Generate the golden vectors under tests/golden/ by running the UNMODIFIED reference
(/root/reference, through oracle/run_reference.py) in the build container.

TEST INFRASTRUCTURE ONLY.  Run once here (`python oracle/gen_golden.py`); the outputs are committed
because /root/reference does not exist on the GPU box.

Outputs
  tests/golden/ref/...            copies of the reference's own test DATA (inputs and truth outputs of
                                  tests/integration_tests/data; the two large amira inputs gzipped)
  tests/golden/synthetic.json     reference PRG + pre-order tree dump on seeded synthetic MSAs
                                  (make_prg_b200.synth, regenerated from the seed by the tests)
  tests/golden/units.json         per-function known answers from the reference on random small MSAs:
                                  consensus, intervals, has_empty_sequence, kmeans_cluster_seqs, expansion
  tests/golden/kmeans_cases.npz   every distinct (count matrix, K) the reference handed to KMeans on
                                  its fixtures and the synthetic sets, with sklearn's labels and inertia
"""
import gzip
import hashlib
import json
import os
import shutil
import sys
import tempfile
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
REPO = HERE.parent
sys.path.insert(0, str(REPO))
sys.path.insert(0, str(HERE))

import run_reference as rr  # noqa: E402
from make_prg_b200 import synth  # noqa: E402

GOLD = REPO / "tests" / "golden"


def copy_reference_fixtures():
    out = GOLD / "ref"
    if out.exists():
        shutil.rmtree(out)
    (out / "truth").mkdir(parents=True)
    D = rr.DATA
    for fa in sorted(D.glob("*.fa")) + sorted(D.glob("*.fa.gz")):
        shutil.copy(fa, out / fa.name)
    for sub in ("sample_example", "several", "several_compressed", "several_empty"):
        shutil.copytree(D / sub, out / sub)
    (out / "amira_MSAs").mkdir()
    for fa in sorted((D / "amira_MSAs").glob("*.fasta")):
        with open(fa, "rb") as src, gzip.GzipFile(out / "amira_MSAs" / (fa.name + ".gz"), "wb",
                                                  mtime=0) as dst:
            dst.write(src.read())
    for case_dir in sorted((D / "truth_output").iterdir()):
        dst = out / "truth" / case_dir.name
        dst.mkdir()
        for f in case_dir.iterdir():
            if f.name.endswith(".update_DS.zip"):
                continue  # pickles of Biopython objects: not checkable here
            shutil.copy(f, dst / f.name)
    os.system(f"chmod -R u+w {out}")


SYNTH_CASES = (
    [dict(config=2, index=i, N=5, L=7) for i in range(6)]
    + [dict(config=3, index=i, N=5, L=7) for i in range(2)]
    + [dict(config=5, index=0, N=5, L=L) for L in (3, 5, 7, 9, 11, 13, 15)]
    + [dict(config=5, index=1, N=5, L=L) for L in (3, 7)]
    + [dict(config=2, index=7, N=2, L=7), dict(config=2, index=8, N=1, L=7),
       dict(config=2, index=9, N=10, L=5)]
    # "4flat": round 1's config-#4 generator (no deep clades: the clustering loop ends before KMeans);
    # the deep-clade config #4 has its own vectors (oracle/gen_golden_deep.py -> tests/golden/deep.json)
    + [dict(config="4flat", index=0, N=10, L=7, rows=300, cols=1500)]
)


def _kmeans_tap(store):
    """Record every (X, K) the reference passes to KMeans, with labels/inertia."""
    import make_prg.from_msa.cluster_sequences as cs

    base = cs.KMeans

    class Tap(base):
        def fit(self, X, *a, **k):
            res = super().fit(X, *a, **k)
            key = hashlib.sha256(X.tobytes() + bytes([self.n_clusters])).hexdigest()
            if key not in store:
                store[key] = (X.copy(), self.n_clusters,
                              self.predict(X).astype(np.int32), float(self.inertia_))
            return res

    cs.KMeans = Tap
    return base


def random_small_msa(rng):
    """Tie-prone small MSAs with gaps and occasional IUPAC codes."""
    rows = int(rng.integers(1, 13))
    cols = int(rng.integers(1, 61))
    n_haps = int(rng.integers(1, 5))
    alphabet = np.frombuffer(b"ACGT", np.uint8)
    root = alphabet[rng.integers(0, 4, cols)]
    haps = np.tile(root, (n_haps, 1))
    for h in range(n_haps):
        m = rng.random(cols) < rng.choice([0.0, 0.05, 0.15, 0.4])
        haps[h, m] = alphabet[rng.integers(0, 4, int(m.sum()))]
        for _ in range(int(rng.integers(0, 3))):
            s = int(rng.integers(0, cols))
            haps[h, s:s + int(rng.integers(1, 9))] = ord("-")
    M = haps[rng.integers(0, n_haps, rows)].copy()
    noise = rng.random(M.shape) < rng.choice([0.0, 0.02, 0.1])
    M[noise] = alphabet[rng.integers(0, 4, int(noise.sum()))]
    if rng.random() < 0.3:
        amb = rng.random(M.shape) < 0.03
        M[amb] = np.frombuffer(b"RYKMSW", np.uint8)[rng.integers(0, 6, int(amb.sum()))]
    if rng.random() < 0.15:
        M[:, int(rng.integers(0, cols))] = ord("-")
    return M


def unit_vectors(n_cases=400, seed=12345):
    rr.load_reference()
    from Bio.Align import MultipleSeqAlignment
    from Bio.Seq import Seq
    from Bio.SeqRecord import SeqRecord
    from make_prg.from_msa.cluster_sequences import kmeans_cluster_seqs
    from make_prg.from_msa.interval_partition import IntervalPartitioner
    from make_prg.utils.seq_utils import (SequenceCurationError, SequenceExpander,
                                          get_consensus_from_MSA, has_empty_sequence)

    rng = np.random.default_rng(seed)
    out = []
    for case in range(n_cases):
        M = random_small_msa(rng)
        rows = [r.tobytes().decode() for r in M]
        msa = MultipleSeqAlignment([SeqRecord(Seq(s), id=f"s{i}", name=f"s{i}", description=f"s{i}")
                                    for i, s in enumerate(rows)])
        L = int(rng.choice([1, 2, 3, 5, 7, 9]))
        rec = {"rows": rows, "L": L}
        cons = get_consensus_from_MSA(msa)
        rec["consensus"] = cons
        queries = []
        for _ in range(6):
            a = int(rng.integers(0, M.shape[1]))
            b = int(rng.integers(a, M.shape[1]))
            queries.append([a, b, bool(has_empty_sequence(msa, (a, b)))])
        rec["has_empty"] = queries
        try:
            m, n, a = IntervalPartitioner(cons, L, msa).get_intervals()
            rec["match"] = [[i.start, i.stop] for i in m]
            rec["nonmatch"] = [[i.start, i.stop] for i in n]
            rec["all"] = [[i.start, i.stop, 0 if i in m else 1] for i in a]
        except SequenceCurationError:
            rec["match"] = rec["nonmatch"] = rec["all"] = None
        try:
            rec["expanded"] = SequenceExpander.get_expanded_sequences_from_MSA(msa)
        except SequenceCurationError:
            rec["expanded"] = None
        try:
            res = kmeans_cluster_seqs(msa, L)
            rec["clustered_ids"] = res.clustered_ids
        except SequenceCurationError:
            rec["clustered_ids"] = None
        out.append(rec)
    return out


def main():
    GOLD.mkdir(parents=True, exist_ok=True)
    assert rr.check_reference_against_its_own_truth(verbose=False), "reference harness broken"
    copy_reference_fixtures()

    rr.load_reference()
    store = {}
    _kmeans_tap(store)

    # KMeans problems on the reference's own fixtures
    for f in ("alsB.fasta", "group_18516.fasta"):
        rr.ref_build(rr.DATA / "amira_MSAs" / f, 5, 7)
    for case, L in rr.SMALL_CASES.items():
        rr.ref_build(rr.DATA / f"{case}.fa", 5, L)

    synth_out = []
    for spec in SYNTH_CASES:
        M = synth.config_msa(spec["config"], spec["index"], spec.get("rows"), spec.get("cols"))
        with tempfile.NamedTemporaryFile("w", suffix=".fa", delete=False) as fh:
            fh.write(synth.to_fasta(M))
            path = fh.name
        builder, prg = rr.ref_build(path, spec["N"], spec["L"], locus_name="synth")
        os.unlink(path)
        rec = dict(spec)
        rec.update(shape=list(M.shape), msa_sha256=hashlib.sha256(M.tobytes()).hexdigest(),
                   prg=prg, prg_sha256=hashlib.sha256(prg.encode()).hexdigest(),
                   gfa_sha256=hashlib.sha256(rr.ref_gfa(prg).encode()).hexdigest(),
                   bin_sha256=hashlib.sha256(rr.ref_bin(prg)).hexdigest(),
                   n_nodes=builder.next_node_id, n_sites=(builder.site_num - 5) // 2,
                   tree=rr.dump_tree(builder))
        synth_out.append(rec)
        print("synthetic", spec, "nodes", rec["n_nodes"], "sites", rec["n_sites"], flush=True)
    with open(GOLD / "synthetic.json", "w") as fh:
        json.dump(synth_out, fh)

    units = unit_vectors()
    with open(GOLD / "units.json", "w") as fh:
        json.dump(units, fh)
    print("unit cases", len(units))

    items = sorted(store.values(), key=lambda t: (t[0].shape, t[1], t[0].tobytes()))
    arrays = {}
    for i, (X, K, labels, inertia) in enumerate(items):
        arrays[f"X{i}"] = X.astype(np.int16) if (X == X.astype(np.int16)).all() else X
        arrays[f"K{i}"] = np.int32(K)
        arrays[f"labels{i}"] = labels
        arrays[f"inertia{i}"] = np.float64(inertia)
    arrays["count"] = np.int32(len(items))
    np.savez_compressed(GOLD / "kmeans_cases.npz", **arrays)
    print("kmeans cases", len(items), "largest", max(t[0].shape for t in items))


if __name__ == "__main__":
    main()
