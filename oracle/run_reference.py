"""Harness that runs the UNMODIFIED reference source (/root/reference) in this container.

TEST INFRASTRUCTURE ONLY -- never imported by make_prg_b200/.  It cannot travel to the GPU
box (/root/reference does not exist there); it is used here to
  * validate the oracle restatement (oracle/make_prg_oracle.py) against the real reference, and
  * generate the golden vectors committed under tests/golden/ (oracle/gen_golden.py).

Recipe:
  1. oracle/refshim supplies the Biopython / intervaltree surface the from_msa path touches.
  2. importlib.metadata.version("make_prg") is patched to return "0.5.0".
  3. make_prg.from_msa.cluster_sequences.KMeans is replaced by a subclass that adds n_init=10
     (the default of the pinned scikit-learn 1.3.0, poetry.lock:1115-1116).
  4. OMP_NUM_THREADS=1 so sklearn's inertia reduction is sequential (the reference's own golden
     files correspond to that setting).
"""
import os
import sys
from pathlib import Path

os.environ.setdefault("OMP_NUM_THREADS", "1")

SHIM_ROOT = Path(__file__).resolve().parent / "refshim"
# /root/reference in the build container; on the GPU box the copy staged by oracle/stage_reference.py
_STAGED = Path(__file__).resolve().parent / "_ref"
REFERENCE_ROOT = Path(os.environ.get(
    "MAKE_PRG_REFERENCE",
    "/root/reference" if Path("/root/reference/make_prg").exists() else str(_STAGED)))
DATA = REFERENCE_ROOT / "tests" / "integration_tests" / "data"

_loaded = False


def reference_available() -> bool:
    return (REFERENCE_ROOT / "make_prg" / "recursion_tree.py").exists()


def load_reference():
    """Import the reference package under the stand-ins; returns the make_prg module."""
    global _loaded
    import importlib.metadata as md

    if not _loaded:
        if not reference_available():
            raise RuntimeError(f"reference source not found at {REFERENCE_ROOT}")
        for p in (str(REFERENCE_ROOT), str(SHIM_ROOT)):
            if p not in sys.path:
                sys.path.insert(0, p)
        real_version = md.version
        md.version = lambda name: "0.5.0" if name == "make_prg" else real_version(name)
        import warnings

        warnings.filterwarnings("ignore")
        from loguru import logger

        logger.remove()
        import make_prg.from_msa.cluster_sequences as cs
        from sklearn.cluster import KMeans as _KMeans

        class KMeansSk13(_KMeans):
            """sklearn >= 1.4 defaults to n_init=1 for k-means++; the pinned 1.3.0 used 10."""

            def __init__(self, n_clusters, random_state, algorithm):
                super().__init__(n_clusters=n_clusters, random_state=random_state,
                                 algorithm=algorithm, n_init=10)

        cs.KMeans = KMeansSk13
        _loaded = True
    import make_prg

    return make_prg


def ref_build(msa_path, max_nesting=5, min_match_length=7, locus_name=None):
    """PrgBuilder(...) + build_prg() through the reference; returns (builder, prg_string)."""
    load_reference()
    from make_prg.prg_builder import PrgBuilder

    msa_path = Path(msa_path)
    builder = PrgBuilder(locus_name or msa_path.name, msa_path, "fasta", max_nesting,
                         min_match_length)
    return builder, builder.build_prg()


def ref_gfa(prg: str) -> str:
    load_reference()
    from make_prg.utils.gfa import GFA_Output

    g = GFA_Output("H\tVN:Z:1.0\tbn:Z:--linear --singlearr\n")
    g.build_gfa_string(prg_string=prg)
    return g.gfa_string


def ref_bin(prg: str) -> bytes:
    load_reference()
    from make_prg.utils.prg_encoder import PrgEncoder, to_bytes

    return b"".join(map(to_bytes, PrgEncoder().encode(prg)))


def ref_cli(argv):
    """Run `make_prg <argv...>` through the reference's own entry point."""
    load_reference()
    from make_prg.__main__ import main

    old = sys.argv
    sys.argv = ["make_prg"] + [str(a) for a in argv]
    try:
        main()
    finally:
        sys.argv = old


def dump_tree(builder):
    """Pre-order dump [(class, node_id, nesting_level, n_rows, n_cols_after_gap_removal, n_children)]."""
    out = []

    def walk(node):
        out.append((type(node).__name__, node.node_id, node.nesting_level, len(node.alignment),
                    node.alignment.get_alignment_length(), len(node.children)))
        for child in node.children:
            walk(child)

    walk(builder.root)
    return out


SMALL_CASES = {
    "match": 7, "match.nonmatch": 7, "match.nonmatch.match": 7, "match.nonmatch.shortmatch": 7,
    "match.staggereddash": 7, "nonmatch": 7, "nonmatch.match": 7, "nonmatch.shortmatch": 7,
    "shortmatch.nonmatch": 7, "shortmatch.nonmatch.match": 7, "contains_n": 7,
    "contains_n_and_RYKMSW": 7, "contains_n_no_variants": 7, "contains_RYKMSW": 7,
    "a_column_full_of_Ns": 7, "nested_snps_seq_backgrounds": 3,
    "nested_snps_seq_backgrounds_more_seqs": 3, "nested_snps_deletion": 1,
}


def check_reference_against_its_own_truth(verbose=True):
    """Reproduce every from_msa golden output of the reference byte for byte."""
    import glob
    import re

    ok = True
    for case, L in SMALL_CASES.items():
        _, prg = ref_build(DATA / f"{case}.fa", 5, L)
        truth = (DATA / "truth_output" / case / f"{case}.prg.fa").read_text().split("\n")[1]
        gfa_truth = (DATA / "truth_output" / case / f"{case}.prg.gfa").read_text()
        bin_truth = (DATA / "truth_output" / case / f"{case}.prg.bin").read_bytes()
        good = prg == truth and ref_gfa(prg) == gfa_truth and ref_bin(prg) == bin_truth
        ok &= good
        if verbose:
            print(f"{case:45s} {'ok' if good else 'MISMATCH'}")
    for setname, pattern in (("sample_example", "*.fa"), ("amira_MSAs", "*.fasta")):
        lines = (DATA / "truth_output" / setname / f"{setname}.prg.fa").read_text().split("\n")
        truth = {lines[i][1:]: lines[i + 1] for i in range(0, len(lines) - 1, 2)}
        for f in sorted(glob.glob(str(DATA / setname / pattern))):
            name = re.sub(r"\.(fa|fasta)(\.gz)?$", "", Path(f).name)
            _, prg = ref_build(f, 5, 7)
            good = prg == truth.get(name)
            ok &= good
            if verbose:
                print(f"{setname}/{name:35s} {'ok' if good else 'MISMATCH'}")
    return ok


if __name__ == "__main__":
    sys.exit(0 if check_reference_against_its_own_truth() else 1)
