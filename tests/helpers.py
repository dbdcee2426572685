"""Shared helpers for the test-suite (golden fixture access, seeded synthetic inputs)."""
import gzip
import json
import re
from pathlib import Path

import numpy as np

REPO = Path(__file__).resolve().parent.parent
GOLDEN = REPO / "tests" / "golden"
REF = GOLDEN / "ref"

# case -> min_match_length used by the reference's integration tests
# (tests/integration_tests/test_from_msa.py:35-220 in the reference)
SMALL_CASES = {
    "match": 7, "match.nonmatch": 7, "match.nonmatch.match": 7, "match.nonmatch.shortmatch": 7,
    "match.staggereddash": 7, "nonmatch": 7, "nonmatch.match": 7, "nonmatch.shortmatch": 7,
    "shortmatch.nonmatch": 7, "shortmatch.nonmatch.match": 7, "contains_n": 7,
    "contains_n_and_RYKMSW": 7, "contains_n_no_variants": 7, "contains_RYKMSW": 7,
    "a_column_full_of_Ns": 7, "nested_snps_seq_backgrounds": 3,
    "nested_snps_seq_backgrounds_more_seqs": 3, "nested_snps_deletion": 1,
}


def truth_prg(case):
    return (REF / "truth" / case / f"{case}.prg.fa").read_text().split("\n")[1]


def truth_multi(setname):
    lines = (REF / "truth" / setname / f"{setname}.prg.fa").read_text().split("\n")
    return {lines[i][1:]: lines[i + 1] for i in range(0, len(lines) - 1, 2)}


def locus_name(path):
    return re.sub(r"\.(fa|fasta)(\.gz)?$", "", Path(path).name)


def synthetic_cases():
    with open(GOLDEN / "synthetic.json") as fh:
        return json.load(fh)


def sub_build_cases():
    """NodeFactory.build(alignment, builder, parent_node) vectors from the unmodified reference
    (oracle/gen_golden_sub.py)."""
    with open(GOLDEN / "sub_builds.json") as fh:
        return json.load(fh)


def unit_cases():
    with open(GOLDEN / "units.json") as fh:
        return json.load(fh)


def kmeans_cases():
    z = np.load(GOLDEN / "kmeans_cases.npz")
    for i in range(int(z["count"])):
        yield (z[f"X{i}"].astype(np.float64), int(z[f"K{i}"]), z[f"labels{i}"],
               float(z[f"inertia{i}"]))


def rows_to_matrix(rows):
    if not rows or len(rows[0]) == 0:
        return np.zeros((len(rows), 0), np.uint8)
    return np.frombuffer("".join(rows).encode(), np.uint8).reshape(len(rows), -1).copy()


def prg_to_regex(prg):
    """A PRG string as a regular expression over ACGT: site ` s a1 s+1 a2 ... s ` -> (?:a1|a2|...).
    Every path through the PRG is a match; used for the size-independent property
    'each input sequence is spelled by the PRG'."""
    import re

    tokens = re.findall(r" \d+ |[ACGT]+", prg)
    out = []
    stack = []  # open odd markers
    for tok in tokens:
        if tok[0] == " ":
            m = int(tok)
            if m % 2 == 1:
                if stack and stack[-1] == m:
                    stack.pop()
                    out.append(")")
                else:
                    stack.append(m)
                    out.append("(?:")
            else:
                out.append("|")
        else:
            out.append(tok)
    assert not stack, "unbalanced PRG"
    return "".join(out)


def prg_spells_all_rows(prg, M):
    import re

    pat = re.compile(prg_to_regex(prg))
    seen = set()
    for row in M:
        s = bytes(row[row != ord("-")]).decode()
        if s in seen:
            continue
        seen.add(s)
        if not pat.fullmatch(s):
            return False
    return True


def deep_cases():
    """Deep-clade loci (config #4 class) run through the unmodified reference (oracle/gen_golden_deep.py)."""
    with open(GOLDEN / "deep.json") as fh:
        return json.load(fh)


def deep_msa(case):
    """The MSA of a deep case, regenerated from its seed and pinned by the recorded sha256."""
    import hashlib

    from make_prg_b200 import synth

    M = synth.synth_deep_msa(**case["gen"])
    assert hashlib.sha256(M.tobytes()).hexdigest() == case["msa_sha256"], "synthetic generator drifted"
    return M


def deep_kmeans_problems(case, M):
    """[(X, K, golden record)] for the big KMeans problems of a deep case: the count matrices come from the
    oracle's own k-mer counting (checked against the sha256 of what the reference handed to scikit-learn),
    the answers from the recorded scikit-learn results, so nothing slow runs."""
    import hashlib

    import make_prg_oracle as mo

    golden = iter(case["kmeans"])
    recorded = {}
    out = []

    def replay(X, K):
        # replays the reference's answers in call order; small problems go to the restatement
        if X.shape[0] * X.shape[1] < 4096:
            return mo.kmeans13.kmeans_fit_predict(X, K)
        g = next(golden)
        assert (g["n"], g["F"], g["K"]) == (X.shape[0], X.shape[1], K)
        assert hashlib.sha256(np.ascontiguousarray(X, np.float64).tobytes()).hexdigest() == g["x_sha256"]
        out.append((np.ascontiguousarray(X, np.float64).copy(), K, g))
        return np.array(g["labels"]), float.fromhex(g["inertia"]), None, None

    ids = [f"s{i}" for i in range(M.shape[0])]
    prg, builder = mo.build_prg_from_matrix(ids, M, case["N"], case["L"], kmeans=replay)
    recorded["prg"] = prg
    return out, prg


def tree_dump(res, locus, M):
    """Pre-order dump of a built locus in the shape of oracle/run_reference.dump_tree:
    [class, node_id, nesting_level, rows, columns after all-gap removal, children]."""
    t = res.nodes(locus)
    names = {0: "LeafNode", 1: "MultiIntervalNode", 2: "MultiClusterNode"}
    out = []
    for i in range(len(t["kind"])):
        rows = (np.arange(M.shape[0]) if t["row_off"][i] < 0
                else t["row_pool"][t["row_off"][i]:t["row_off"][i] + t["n_rows"][i]])
        S = M[rows, t["c0"][i]:t["c1"][i]]
        keep = int((~(S == ord("-")).all(axis=0)).sum())
        out.append([names[int(t["kind"][i])], i, int(t["nesting_level"][i]), len(rows), keep,
                    int(t["n_children"][i])])
    return out
